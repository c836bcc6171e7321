"""OnPolicyRunner (drop-in for rsl_rl/runners/on_policy_runner.py:60-320).

Same constructor `(env, train_cfg, log_dir, device)`, `learn`, `log`, `save`, `load`, `get_inference_policy` and the same
TensorBoard scalars / console table.  The rollout loop itself has no host synchronisation: per-step episode statistics are
accumulated on the device and fetched once per iteration (the reference calls `.cpu()` inside the step loop, :149-150)."""
import os
import statistics
import time
from collections import deque
from pathlib import Path

import torch
import yaml

from .. import _ops
from ..algorithms import PPO
from ..modules import ActorCritic
from ...envs.env_arrays import class_to_dict

try:
    from torch.utils.tensorboard import SummaryWriter
except Exception:  # tensorboard is optional on the bench box
    SummaryWriter = None


class OnPolicyRunner:
    def __init__(self, env, train_cfg, log_dir=None, device='cpu'):
        self.cfg = train_cfg["runner"]
        self.alg_cfg = train_cfg["algorithm"]
        self.policy_cfg = train_cfg["policy"]
        self.device = device
        self.env = env
        num_critic_obs = self.env.num_privileged_obs if self.env.num_privileged_obs is not None else self.env.num_obs
        actor_critic_class = {"ActorCritic": ActorCritic}[self.cfg["policy_class_name"]]
        actor_critic = actor_critic_class(self.env.num_obs, num_critic_obs, self.env.num_actions, **self.policy_cfg)
        alg_class = {"PPO": PPO}[self.cfg["algorithm_class_name"]]
        seed = train_cfg.get("seed", 1)
        self.alg = alg_class(actor_critic, device=self.device, seed=seed, env_offset=getattr(env, "_A", None).env_offset if hasattr(env, "_A") else 0,
                             **self.alg_cfg)
        self.num_steps_per_env = self.cfg["num_steps_per_env"]
        self.save_interval = self.cfg["save_interval"]
        self.alg.init_storage(self.env.num_envs, self.num_steps_per_env, [self.env.num_obs], [self.env.num_privileged_obs], [self.env.num_actions])
        self.log_dir = log_dir
        self.writer = None
        self.tot_timesteps = 0
        self.tot_time = 0
        self.current_learning_iteration = 0
        _, _ = self.env.reset()
        if self.log_dir is not None and self.env.cfg.env.test is False:
            Path(self.log_dir).mkdir(parents=True, exist_ok=True)
            all_cfg = {"train_cfg": train_cfg, "env_cfg": class_to_dict(self.env.cfg)}
            yaml.safe_dump(all_cfg, open(os.path.join(self.log_dir, 'config.yaml'), 'w'))
        self.robogauge_client = None  # external evaluation service: out of scope
        N = self.env.num_envs
        self._cur_reward_sum = torch.zeros(N, device=self.device)
        self._cur_episode_length = torch.zeros(N, device=self.device)
        # per-iteration device-side episode log: [T, N] finished-episode returns / lengths (NaN where no episode ended)
        self._done_rew = torch.full((self.num_steps_per_env, N), float("nan"), device=self.device)
        self._done_len = torch.full((self.num_steps_per_env, N), float("nan"), device=self.device)
        self._rollout_graphs = _ops.GraphSet()

    # ---- rollout (on_policy_runner.py:136-153; on_policy_runner_cts.py:147-170) ----------------------------------------------
    # The loop is shared with OnPolicyRunnerCTS, which overrides the two hooks: how the policy is fed and what happens between the env step and the
    # transition bookkeeping (the observation history of the CTS family).
    def _policy_act(self, obs, priv):
        return self.alg.act(obs, priv if priv is not None else obs)

    def _after_env_step(self, obs, dones):
        pass

    def _rollout_steps(self, log, dev):
        """The num_steps_per_env act -> step -> process_env_step loop.  dev: step parameters / sampling counters are device-resident
        (begin_rollout), so the launch sequence depends on nothing the host computes per step and can be captured in one CUDA graph."""
        env, alg = self.env, self.alg
        obs, priv = env.get_observations(), env.get_privileged_observations()
        nan = float("nan")
        ep_infos = []
        alg.storage.step = 0
        for i in range(self.num_steps_per_env):
            actions = self._policy_act(obs, priv)
            obs, priv, rewards, dones, infos = env.step_dev(actions, i) if dev else env.step(actions)
            self._after_env_step(obs, dones)
            alg.process_env_step(rewards, dones, infos)
            if log:
                if not dev and 'episode' in infos:
                    ep_infos.append((infos['episode'], infos.get('episode_valid')))
                self._cur_reward_sum += rewards
                self._cur_episode_length += 1
                self._done_rew[i] = torch.where(dones, self._cur_reward_sum, nan)
                self._done_len[i] = torch.where(dones, self._cur_episode_length, nan)
                self._cur_reward_sum *= ~dones
                self._cur_episode_length *= ~dones
        return ep_infos

    def collect(self, log=False):
        """One rollout.  Replayed as a single CUDA graph (GO2_GRAPH=0 or an env without begin_rollout: eager per-step launches).
        Returns the list of per-step infos['episode'] dicts when logging."""
        env, alg, T = self.env, self.alg, self.num_steps_per_env
        with torch.inference_mode():
            if self._rollout_graphs.enabled and hasattr(env, "begin_rollout") and env.begin_rollout(T):
                alg.begin_rollout(T)
                try:
                    self._rollout_graphs.run(("rollout", bool(log)), lambda: self._rollout_steps(log, True))
                finally:
                    alg.end_rollout(T)
                return env.end_rollout(fetch=bool(log))      # un-logged: no device -> host read at the end of the rollout
            ep_infos = self._rollout_steps(log, False)
            # eager steps serve extras["episode"] without a host sync; rows from before the env's first reset are dropped here, once per rollout
            flags = [v for _, v in ep_infos if v is not None]
            keep = torch.stack(flags).cpu().tolist() if flags else []
            return [e for (e, v), k in zip(ep_infos, keep or [1.0] * len(ep_infos)) if k > 0]

    def collect_host(self, h_actions, h_obs, h_priv, h_rew, h_reset):
        """One rollout with the env reached through its HOST-buffer entry point (`go2_env_step_host_begin / _end`, the calls an external simulator loop
        or a logging host would make): every step copies the sampled actions into the pinned host tensor `h_actions`, the C call uploads them, steps
        and downloads observations / privileged observations / rewards / resets into the other four host tensors.  The policy inference + sampling in
        front of the call and the transition bookkeeping behind it are replayed as per-step CUDA graphs; the download of step t (5 MB at 4096 envs)
        runs on the library's copy stream beside that bookkeeping and the next inference, and is complete — the host buffers hold step t — before
        the actions of step t + 1 go up."""
        env, alg, T = self.env, self.alg, self.num_steps_per_env
        if not hasattr(self, "_host_graphs"):
            self._host_graphs = _ops.GraphSet()
        with torch.inference_mode():
            alg.begin_rollout(T)                      # device-resident sampling counters, as in collect()
            for t in range(T):
                alg.storage.step = t
                self._host_graphs.run(("act", t), lambda: self._host_act(t))
                # the policy output goes to the pinned host tensor on the same stream the C call uploads it from: stream order, no host wait in between
                h_actions.copy_(self._host_actions(t), non_blocking=True)
                env.step_host_end()                               # step t - 1 is on the host (no-op at t = 0)
                env.step_host_begin(h_actions, h_obs, h_priv, h_rew, h_reset)
                alg.storage.step = t
                self._host_graphs.run(("proc", t), self._host_proc)
            env.step_host_end()
            alg.end_rollout(T)

    def _host_act(self, t):
        env, alg = self.env, self.alg
        alg.act(env.obs_buf, env.privileged_obs_buf)
        if getattr(alg, "_join_pending", False):      # the critic's side stream rejoins before the actions leave the device
            alg._side.join(); alg._join_pending = False

    def _host_actions(self, t):
        return self.alg.storage.actions[t]

    def _host_proc(self):
        env = self.env
        self.alg.process_env_step(env.rew_buf, env.reset_buf, {"time_outs": env.time_out_buf})

    def run_iteration_host(self, h_actions, h_obs, h_priv, h_rew, h_reset):
        """collect_host() + returns + update: one iteration end to end through host buffers (bench.py `e2e`)."""
        self.collect_host(h_actions, h_obs, h_priv, h_rew, h_reset)
        env = self.env
        with torch.inference_mode():
            self.alg.compute_returns(env.privileged_obs_buf)
        return self.alg.update()                      # ends with the D2H read of the losses

    def run_iteration(self, sync=None, fetch_losses=True):
        """One un-logged iteration (rollout + returns + update) — the timing loop of bench.py / tools."""
        self.collect(False)
        priv = self.env.get_privileged_observations()
        cobs = priv if priv is not None else self.env.get_observations()
        with torch.inference_mode():
            if sync is not None:
                sync()
            self.alg.compute_returns(cobs)
        return self.alg.update(fetch=fetch_losses)

    def learn(self, num_learning_iterations, init_at_random_ep_len=False):
        if self.log_dir is not None and self.writer is None and SummaryWriter is not None:
            self.writer = SummaryWriter(log_dir=self.log_dir, flush_secs=10)
        if init_at_random_ep_len:
            self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf, high=int(self.env.max_episode_length))
        obs = self.env.get_observations()
        privileged_obs = self.env.get_privileged_observations()
        critic_obs = privileged_obs if privileged_obs is not None else obs
        self.alg.actor_critic.train()
        ep_infos = []
        rewbuffer, lenbuffer = deque(maxlen=100), deque(maxlen=100)
        tot_iter = self.current_learning_iteration + num_learning_iterations
        nan = float("nan")
        it = self.current_learning_iteration
        for it in range(self.current_learning_iteration, tot_iter):
            start = time.time()
            ep_infos = self.collect(log=self.log_dir is not None)
            privileged_obs = self.env.get_privileged_observations()
            critic_obs = privileged_obs if privileged_obs is not None else self.env.get_observations()
            with torch.inference_mode():
                if self.log_dir is not None:  # one fetch per iteration, time-major order like the reference's per-step extend
                    dr, dl = self._done_rew.flatten(), self._done_len.flatten()
                    keep = ~torch.isnan(dr)
                    rewbuffer.extend(dr[keep].cpu().numpy().tolist())
                    lenbuffer.extend(dl[keep].cpu().numpy().tolist())
                torch.cuda.synchronize()
                stop = time.time()
                collection_time = stop - start
                start = stop
                self.alg.compute_returns(critic_obs)
            mean_value_loss, mean_surrogate_loss = self.alg.update()
            stop = time.time()
            learn_time = stop - start
            if self.log_dir is not None:
                self.log(locals())
                if it % self.save_interval == 0:
                    self.save(os.path.join(self.log_dir, 'model_{}.pt'.format(it)), it, False)
            ep_infos.clear()
        self.current_learning_iteration += num_learning_iterations
        if self.log_dir is not None:
            self.save(os.path.join(self.log_dir, 'model_{}.pt'.format(self.current_learning_iteration)), it, True)

    def log(self, locs, width=80, pad=35):
        self.tot_timesteps += self.num_steps_per_env * self.env.num_envs
        self.tot_time += locs['collection_time'] + locs['learn_time']
        iteration_time = locs['collection_time'] + locs['learn_time']
        ep_string = ''
        if locs['ep_infos']:
            for key in locs['ep_infos'][0]:
                vals = []
                for ep_info in locs['ep_infos']:
                    v = ep_info[key]
                    v = v if isinstance(v, torch.Tensor) else torch.tensor([float(v)], device=self.device)
                    vals.append(v.reshape(-1).to(self.device))
                value = torch.mean(torch.cat(vals))
                if self.writer:
                    self.writer.add_scalar('Episode/' + key, value, locs['it'])
                ep_string += f"""{f'Mean episode {key}:':>{pad}} {value:.4f}\n"""
        mean_std = self.alg.actor_critic.std.mean()
        fps = int(self.num_steps_per_env * self.env.num_envs / (locs['collection_time'] + locs['learn_time']))
        if self.writer:
            w = self.writer
            w.add_scalar('Loss/value_function', locs['mean_value_loss'], locs['it'])
            w.add_scalar('Loss/surrogate', locs['mean_surrogate_loss'], locs['it'])
            w.add_scalar('Loss/learning_rate', self.alg.learning_rate, locs['it'])
            w.add_scalar('Policy/mean_noise_std', mean_std.item(), locs['it'])
            w.add_scalar('Perf/total_fps', fps, locs['it'])
            w.add_scalar('Perf/collection time', locs['collection_time'], locs['it'])
            w.add_scalar('Perf/learning_time', locs['learn_time'], locs['it'])
            if len(locs['rewbuffer']) > 0:
                w.add_scalar('Train/mean_reward', statistics.mean(locs['rewbuffer']), locs['it'])
                w.add_scalar('Train/mean_episode_length', statistics.mean(locs['lenbuffer']), locs['it'])
                w.add_scalar('Train/mean_reward/time', statistics.mean(locs['rewbuffer']), self.tot_time)
                w.add_scalar('Train/mean_episode_length/time', statistics.mean(locs['lenbuffer']), self.tot_time)
        head = f" \033[1m Learning iteration {locs['it']}/{self.current_learning_iteration + locs['num_learning_iterations']} \033[0m "
        log_string = (f"""{'#' * width}\n{head.center(width, ' ')}\n\n"""
                      f"""{'Computation:':>{pad}} {fps:.0f} steps/s (collection: {locs['collection_time']:.3f}s, learning {locs['learn_time']:.3f}s)\n"""
                      f"""{'Value function loss:':>{pad}} {locs['mean_value_loss']:.4f}\n"""
                      f"""{'Surrogate loss:':>{pad}} {locs['mean_surrogate_loss']:.4f}\n"""
                      f"""{'Mean action noise std:':>{pad}} {mean_std.item():.2f}\n""")
        if len(locs['rewbuffer']) > 0:
            log_string += (f"""{'Mean reward:':>{pad}} {statistics.mean(locs['rewbuffer']):.2f}\n"""
                           f"""{'Mean episode length:':>{pad}} {statistics.mean(locs['lenbuffer']):.2f}\n""")
        log_string += ep_string
        log_string += (f"""{'-' * width}\n{'Total timesteps:':>{pad}} {self.tot_timesteps}\n"""
                       f"""{'Iteration time:':>{pad}} {iteration_time:.2f}s\n{'Total time:':>{pad}} {self.tot_time:.2f}s\n"""
                       f"""{'ETA:':>{pad}} {self.tot_time / (locs['it'] + 1) * (locs['num_learning_iterations'] - locs['it']):.1f}s\n""")
        print(log_string)

    def save(self, path, it, last_model, infos=None):
        torch.save({'model_state_dict': self.alg.actor_critic.state_dict(), 'optimizer_state_dict': self.alg.optimizer_state_dict(),
                    'iter': self.current_learning_iteration, 'infos': infos}, path)

    def load(self, path, load_optimizer=True):
        loaded_dict = torch.load(path, map_location=self.device, weights_only=False)
        self.alg.actor_critic.load_state_dict(loaded_dict['model_state_dict'])
        if load_optimizer and loaded_dict.get('optimizer_state_dict') is not None:
            self.alg.load_optimizer_state_dict(loaded_dict['optimizer_state_dict'])
        self.current_learning_iteration = loaded_dict['iter']
        # the action-sampling stream is keyed (seed, env, step): continue it where the saved run stood instead of replaying iteration 0's draws
        self.alg._act_step = self.current_learning_iteration * self.num_steps_per_env
        return loaded_dict['infos']

    def get_inference_policy(self, device=None):
        self.alg.actor_critic.eval()
        return self.alg.actor_critic.act_inference

"""OnPolicyRunnerCTS (drop-in for rsl_rl/runners/on_policy_runner_cts.py:63-360): the PPO runner plus the rolling 5-frame
observation history (zeroed on done, shift-append: :155-156), 3-argument act, teacher / student reward buffers, two optimisers
in the checkpoint."""
import os
import statistics
import time
from collections import deque
from pathlib import Path

import torch
import yaml

from .. import _ops
from .on_policy_runner import OnPolicyRunner
from ..algorithms import CTS, MoECTS, MoENGCTS, ACMoECTS, DualMoECTS, MCPCTS
from ..modules import (ActorCriticCTS, ActorCriticMoECTS, ActorCriticMoENGCTS, ActorCriticACMoECTS, ActorCriticDualMoECTS,
                       ActorCriticMCPCTS)
from ...utils.cfg_dict import class_to_dict
from .on_policy_runner import SummaryWriter


class OnPolicyRunnerCTS:
    def __init__(self, env, train_cfg, log_dir=None, device='cpu'):
        self.cfg, self.alg_cfg, self.policy_cfg = train_cfg["runner"], train_cfg["algorithm"], train_cfg["policy"]
        self.device, self.env = device, env
        history_length = train_cfg["history_length"]
        self.history_length = history_length
        num_critic_obs = self.env.num_privileged_obs if self.env.num_privileged_obs is not None else self.env.num_obs
        model_class = {"ActorCriticCTS": ActorCriticCTS, "ActorCriticMoECTS": ActorCriticMoECTS, "ActorCriticMoENGCTS": ActorCriticMoENGCTS,
                       "ActorCriticACMoECTS": ActorCriticACMoECTS, "ActorCriticDualMoECTS": ActorCriticDualMoECTS,
                       "ActorCriticMCPCTS": ActorCriticMCPCTS}[self.cfg["policy_class_name"]]
        model = model_class(self.env.num_obs, num_critic_obs, self.env.num_actions, self.env.num_envs, history_length, **self.policy_cfg)
        alg_class = {"CTS": CTS, "MoECTS": MoECTS, "MoENGCTS": MoENGCTS, "ACMoECTS": ACMoECTS, "DualMoECTS": DualMoECTS,
                     "MCPCTS": MCPCTS}[self.cfg["algorithm_class_name"]]
        # the value of the MoE-actor variants needs the actor's gate, hence the observations (on_policy_runner_cts.py:182-185)
        self._returns_need_obs = self.cfg["algorithm_class_name"] in ("ACMoECTS", "DualMoECTS")
        off = env._A.env_offset if hasattr(env, "_A") else 0
        self.alg = alg_class(model, self.env.num_envs, history_length, device=self.device, seed=train_cfg.get("seed", 0), env_offset=off, **self.alg_cfg)
        self.num_steps_per_env, self.save_interval = self.cfg["num_steps_per_env"], self.cfg["save_interval"]
        self.alg.init_storage(self.env.num_envs, self.num_steps_per_env, [self.env.num_obs], [self.env.num_privileged_obs], [self.env.num_actions])
        self.history = torch.zeros(self.env.num_envs, history_length, self.env.num_obs, device=self.device)
        self.log_dir, self.writer = log_dir, None
        self.tot_timesteps, self.tot_time, self.current_learning_iteration = 0, 0, 0
        _, _ = self.env.reset()
        if self.log_dir is not None and self.env.cfg.env.test is False:
            Path(self.log_dir).mkdir(parents=True, exist_ok=True)
            yaml.safe_dump({"train_cfg": train_cfg, "env_cfg": class_to_dict(self.env.cfg)}, open(os.path.join(self.log_dir, 'config.yaml'), 'w'))
        N = self.env.num_envs
        self._cur_reward_sum, self._cur_episode_length = torch.zeros(N, device=device), torch.zeros(N, device=device)
        self._done_rew = torch.full((self.num_steps_per_env, N), float("nan"), device=device)
        self._done_len = torch.full((self.num_steps_per_env, N), float("nan"), device=device)
        self._rollout_graphs = _ops.GraphSet()
        self._is_teacher = torch.zeros(N, dtype=torch.bool, device=device)
        self._is_teacher[self.alg.teacher_env_idxs] = True

    def _roll_history(self, obs, dones):
        d8 = None if dones is None else (dones.view(torch.uint8) if dones.dtype == torch.bool else dones.to(torch.uint8))
        _ops.call("go2_history_update", _ops.ptr(self.history), _ops.ptr(obs), _ops.ptr(d8), self.env.num_envs, self.history_length, self.env.num_obs)

    # ---- rollout (on_policy_runner_cts.py:147-170): OnPolicyRunner's loop with the history in the policy input and the history roll between the env
    # step and the transition bookkeeping; replayed as one CUDA graph by the shared collect()
    def _policy_act(self, obs, priv):
        return self.alg.act(obs, priv, self.history.flatten(1))

    def _after_env_step(self, obs, dones):
        self._roll_history(obs, dones)

    _rollout_steps = OnPolicyRunner._rollout_steps
    collect = OnPolicyRunner.collect

    # host-buffer rollout: same loop as OnPolicyRunner.collect_host (history in the policy input, history roll before the bookkeeping)
    collect_host = OnPolicyRunner.collect_host

    def _host_act(self, t):
        env, alg = self.env, self.alg
        alg.act(env.obs_buf, env.privileged_obs_buf, self.history.flatten(1))
        if getattr(alg, "_join_pending", False):
            alg._side.join(); alg._join_pending = False

    def _host_actions(self, t):
        return self.alg._actions_env

    def _host_proc(self):
        env = self.env
        self._roll_history(env.obs_buf, env.reset_buf)
        self.alg.process_env_step(env.rew_buf, env.reset_buf, {"time_outs": env.time_out_buf})

    def run_iteration_host(self, h_actions, h_obs, h_priv, h_rew, h_reset):
        env = self.env
        if not getattr(self, "_hist_primed", False):
            self._roll_history(env.get_observations(), None)
            self._hist_primed = True
        self.collect_host(h_actions, h_obs, h_priv, h_rew, h_reset)
        with torch.inference_mode():
            self._compute_returns(env.obs_buf, env.privileged_obs_buf)
        return self.alg.update()

    def run_iteration(self, sync=None, fetch_losses=True):
        """One un-logged iteration (rollout + returns + both update passes) — the timing loop of bench.py / tools."""
        env, alg = self.env, self.alg
        if not getattr(self, "_hist_primed", False):
            self._roll_history(env.get_observations(), None)
            self._hist_primed = True
        self.collect(False)
        with torch.inference_mode():
            if sync is not None:
                sync()
            self._compute_returns(env.get_observations(), env.get_privileged_observations())
        return alg.update(fetch=fetch_losses)

    def _compute_returns(self, obs, privileged_obs):
        if self._returns_need_obs:
            self.alg.compute_returns(obs, privileged_obs, self.history.flatten(1))
        else:
            self.alg.compute_returns(privileged_obs, self.history.flatten(1))

    def learn(self, num_learning_iterations, init_at_random_ep_len=False):
        if self.log_dir is not None and self.writer is None and SummaryWriter is not None:
            self.writer = SummaryWriter(log_dir=self.log_dir, flush_secs=10)
        if init_at_random_ep_len:
            self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf, high=int(self.env.max_episode_length))
        obs, privileged_obs = self.env.get_observations(), self.env.get_privileged_observations()
        assert privileged_obs is not None
        self._roll_history(obs, None)
        self.alg.model.train()
        ep_infos = []
        bufs = {k: deque(maxlen=100) for k in ("teacher_rew", "teacher_len", "student_rew", "student_len")}
        nan = float("nan")
        tot_iter = self.current_learning_iteration + num_learning_iterations
        it = self.current_learning_iteration
        for it in range(self.current_learning_iteration, tot_iter):
            start = time.time()
            ep_infos = self.collect(log=self.log_dir is not None)
            privileged_obs = self.env.get_privileged_observations()
            with torch.inference_mode():
                if self.log_dir is not None:
                    for who, mask in (("teacher", self._is_teacher), ("student", ~self._is_teacher)):
                        dr, dl = self._done_rew[:, mask].flatten(), self._done_len[:, mask].flatten()
                        keep = ~torch.isnan(dr)
                        bufs[who + "_rew"].extend(dr[keep].cpu().numpy().tolist())
                        bufs[who + "_len"].extend(dl[keep].cpu().numpy().tolist())
                torch.cuda.synchronize()
                stop = time.time()
                collection_time = stop - start
                start = stop
                self._compute_returns(self.env.get_observations(), privileged_obs)
            losses = self.alg.update()
            stop = time.time()
            learn_time = stop - start
            self.current_learning_iteration += 1
            if self.log_dir is not None:
                self.log(locals())
                if it % self.save_interval == 0:
                    self.save(os.path.join(self.log_dir, 'model_{}.pt'.format(it)), it, False)
            ep_infos.clear()
        if self.log_dir is not None:
            self.save(os.path.join(self.log_dir, 'model_{}.pt'.format(self.current_learning_iteration)), it, True)

    def log(self, locs, width=80, pad=35):
        self.tot_timesteps += self.num_steps_per_env * self.env.num_envs
        self.tot_time += locs['collection_time'] + locs['learn_time']
        names = ["value_function", "surrogate", "entropy", "latent", "load_balance", "actor_load_balance"]      # on_policy_runner_cts.py:227-237
        fps = int(self.num_steps_per_env * self.env.num_envs / (locs['collection_time'] + locs['learn_time']))
        std = getattr(self.alg.model, "std", None)       # the MCP policy has no std parameter (on_policy_runner_cts.py:226-227): log the last batch's mean sigma
        mean_std = std.mean() if std is not None else self.alg._sigma.mean()
        ep_string = ''
        if locs['ep_infos']:
            for key in locs['ep_infos'][0]:
                vals = [(v if isinstance(v, torch.Tensor) else torch.tensor([float(v)], device=self.device)).reshape(-1).to(self.device)
                        for v in (e[key] for e in locs['ep_infos'])]
                value = torch.mean(torch.cat(vals))
                if self.writer:
                    self.writer.add_scalar('Episode/' + key, value, locs['it'])
                ep_string += f"""{f'Mean episode {key}:':>{pad}} {value:.4f}\n"""
        if self.writer:
            for n, v in zip(names, locs['losses']):
                self.writer.add_scalar('Loss/' + n, v, locs['it'])
            self.writer.add_scalar('Loss/learning_rate', self.alg.learning_rate, locs['it'])
            self.writer.add_scalar('Policy/mean_noise_std', mean_std.item(), locs['it'])
            self.writer.add_scalar('Perf/total_fps', fps, locs['it'])
            self.writer.add_scalar('Perf/collection time', locs['collection_time'], locs['it'])
            self.writer.add_scalar('Perf/learning_time', locs['learn_time'], locs['it'])
            for who in ("teacher", "student"):
                if len(locs['bufs'][who + "_rew"]) > 0:
                    self.writer.add_scalar(f'Train/mean_{who}_reward', statistics.mean(locs['bufs'][who + "_rew"]), locs['it'])
                    self.writer.add_scalar(f'Train/mean_{who}_episode_length', statistics.mean(locs['bufs'][who + "_len"]), locs['it'])
        head = f" \033[1m Learning iteration {locs['it']}/{locs['tot_iter']} \033[0m "
        s = f"""{'#' * width}\n{head.center(width, ' ')}\n\n{'Computation:':>{pad}} {fps:.0f} steps/s (collection: {locs['collection_time']:.3f}s, learning {locs['learn_time']:.3f}s)\n"""
        for n, v in zip(names, locs['losses']):
            s += f"""{n + ' loss:':>{pad}} {v:.4f}\n"""
        s += f"""{'Mean action noise std:':>{pad}} {mean_std.item():.2f}\n"""
        for who in ("teacher", "student"):
            if len(locs['bufs'][who + "_rew"]) > 0:
                s += f"""{f'Mean {who} reward:':>{pad}} {statistics.mean(locs['bufs'][who + '_rew']):.2f}\n"""
        print(s + ep_string)

    def save(self, path, it, last_model, infos=None):
        torch.save({'model_state_dict': self.alg.model.state_dict(), 'optimizer1_state_dict': self.alg.optimizer1_state_dict(),
                    'optimizer2_state_dict': self.alg.optimizer2_state_dict(), 'iter': self.current_learning_iteration, 'infos': infos}, path)

    def load(self, path, load_optimizer=True):
        d = torch.load(path, map_location=self.device, weights_only=False)
        self.alg.model.load_state_dict(d['model_state_dict'])
        if load_optimizer:
            self.alg.load_optimizer_state_dicts(d.get('optimizer1_state_dict'), d.get('optimizer2_state_dict'))
        self.current_learning_iteration = d['iter']
        # the action-sampling stream is keyed (seed, env, step): continue it where the saved run stood instead of replaying iteration 0's draws
        self.alg._act_step = self.current_learning_iteration * self.num_steps_per_env
        return d['infos']

    def get_inference_policy(self, device=None):
        self.alg.model.eval()
        return self.alg.model.act_inference

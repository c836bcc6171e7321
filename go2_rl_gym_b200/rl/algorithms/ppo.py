"""PPO (drop-in for rsl_rl/algorithms/ppo.py:37-187) on hand-written kernels.

Same constructor and method names.  Differences a caller can observe, all deliberate and documented in DESIGN.md:
  * act() samples with the library's Philox normal (keyed by seed / env / step) instead of torch's global generator;
  * update() runs without any host synchronisation: the KL-adaptive learning rate (ppo.py:139-151) lives in a device scalar
    read by the fused clip+Adam kernel; `learning_rate` on the host is refreshed once per update() for logging;
  * with torch.distributed initialised, each optimiser step all-reduces the flat gradient (one NCCL call, scalar tail included)."""
import os

import torch
import torch.distributed as dist

from .. import _ops, dist_utils
from ..modules import ActorCritic
from ..storage import RolloutStorage


def adam_group_template():
    """The param_group keys (with their defaults) of the installed torch's Adam: the key set differs between torch versions (2.11 adds
    decoupled_weight_decay; the reference recommends 2.3.1), and the reference's Optimizer.load_state_dict takes the saved group verbatim."""
    g = dict(torch.optim.Adam([torch.zeros(1)], lr=1e-3).param_groups[0])
    g.pop("params")
    return g


_CRITIC_JOIN_EARLY = os.environ.get("GO2_CRITIC_JOIN", "late") == "early"

class PPO:
    actor_critic: ActorCritic

    def __init__(self, actor_critic, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95,
                 value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True,
                 schedule="fixed", desired_kl=0.01, device='cpu', seed=1, env_offset=0):
        self.device = device
        self.desired_kl = desired_kl
        self.schedule = schedule
        self.learning_rate = learning_rate
        self.actor_critic = actor_critic
        self.storage = None
        self.transition = RolloutStorage.Transition()
        self.clip_param = clip_param
        self.num_learning_epochs = num_learning_epochs
        self.num_mini_batches = num_mini_batches
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.gamma = gamma
        self.lam = lam
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.seed = int(seed)
        self.env_offset = int(env_offset)
        self._act_step = 0
        self._dev_steps = None
        self._opt_step = 0
        self.world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        # GO2_DIST_GRAPH=1 (opt-in): capture the all-reduces inside the update graph.  Measured on 2 B200: 16.9 vs 17.6 ms / iteration (go2),
        # 35.3 vs 37.4 ms (go2_moe_cts), identical parameters on all ranks - but the process then hangs in destroy_process_group() / interpreter
        # exit while graphs holding NCCL kernels are alive (profiles/r01m_dist_graph_check.txt), so the segmented path stays the default.
        self._dist_graph = os.environ.get("GO2_DIST_GRAPH", "0") == "1"
        self.optimizer = None  # Adam state lives in flat vectors; see optimizer_state_dict()

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        dev = self.device
        self.storage = RolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape, dev)
        T = num_transitions_per_env
        self.mini_batch_size = (num_envs * T) // self.num_mini_batches
        ac = self.actor_critic
        ac.flatten_(dev, max(num_envs, self.mini_batch_size), train_rows=self.mini_batch_size)
        n = ac.flat_params.numel()
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        self._lr = torch.zeros(4, device=dev)      # {lr, optimiser step, Adam bias corrections} — device resident
        self._lr[0] = float(self.learning_rate)
        # env-sharded over GPUs: the flat gradient lives in a symmetric buffer and the per-step exchange is the library's own NVLink kernel
        # (dist_utils.P2PReducer), so the 20 optimiser steps and their exchanges replay as ONE graph; otherwise (gloo, no peer access) NCCL / gloo
        # all-reduces between graph segments
        self._red = dist_utils.reducer_for(ac.flat_grads) if self.world_size > 1 else None
        self._graphs = _ops.GraphSet()
        self._side = _ops.SideStream(dev)          # the critic's chain runs next to the actor's (see _ops.SideStream)
        self._join_pending = False
        self._scal = self._red.tail if self._red is not None else torch.zeros(20, device=dev)
        self._log = torch.zeros(5, device=dev)
        self._scratch = torch.zeros(1025, device=dev)
        self._dmu = torch.empty(self.mini_batch_size, action_shape[0], device=dev)
        self._dval = torch.empty(self.mini_batch_size + 4, 1, device=dev)[:self.mini_batch_size]
        # [1+1, mb] view of the value gradient for the tensor-core wgrad: row 0 = dV^T (same storage), then the engine's ones row is not needed here
        self._mu_b = torch.empty(self.mini_batch_size, action_shape[0], device=dev)
        self._val_b = torch.empty(self.mini_batch_size, 1, device=dev)
        self._last_values = torch.empty(num_envs, 1, device=dev)

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    # ---- rollout ---------------------------------------------------------------------------------------------
    def begin_rollout(self, T):
        """The next T act() calls read their sampling step counter from a device array (same values act() would pass by value), so the
        whole rollout can be replayed as one CUDA graph.  end_rollout() returns to host-side counters."""
        self._dev_steps = _ops.upload_steps(getattr(self, "_dev_steps_buf", None), self._act_step, T, self.device)
        self._dev_steps_buf = self._dev_steps
        self._act_step += T
        self.actor_critic.actor_engine.mark_dirty(); self.actor_critic.critic_engine.mark_dirty()    # the graph always refreshes the padded weights

    def end_rollout(self, T):
        self._dev_steps = None
        self.storage.step = T

    def act(self, obs, critic_obs):
        st, t, ac = self.storage, self.storage.step, self.actor_critic
        if t >= st.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        N, A = st.num_envs, st.actions.shape[-1]
        st.observations[t].copy_(obs)
        cobs = (st.privileged_observations if st.privileged_observations is not None else st.observations)[t]
        cobs.copy_(critic_obs)
        # the value is first needed by process_env_step: the critic reads the STORED row (the env overwrites its own buffer during step()) on the
        # side stream, next to the actor, the sampling and — inside a rollout opened by begin_rollout() — the env step itself
        sd = self._side
        sd.fork()
        with sd:
            ac.evaluate(cobs, out=st.values[t])
        mu = ac._actor_forward(st.observations[t])
        if self._dev_steps is not None:      # rollout opened by begin_rollout(): the Philox step counter comes from device memory
            _ops.call("go2_sample_actions_dev", _ops.ptr(mu), _ops.ptr(ac.std.data), _ops.ptr(st.actions[t]), _ops.ptr(st.actions_log_prob[t]),
                      _ops.ptr(st.mu[t]), _ops.ptr(st.sigma[t]), N, A, self.seed, self._dev_steps.data_ptr() + 4 * t, self.env_offset)
        else:
            self._act_step += 1
            _ops.call("go2_sample_actions", _ops.ptr(mu), _ops.ptr(ac.std.data), _ops.ptr(st.actions[t]), _ops.ptr(st.actions_log_prob[t]),
                      _ops.ptr(st.mu[t]), _ops.ptr(st.sigma[t]), N, A, self.seed, self._act_step, self.env_offset)
        if self._dev_steps is not None:
            if _CRITIC_JOIN_EARLY:           # A/B: the critic rejoins before the env step (it then only runs beside the actor and the sampling)
                sd.join()
            else:
                self._join_pending = True        # joined in process_env_step
        else:
            sd.join()
        self.transition.actions = st.actions[t]
        self.transition.values = st.values[t]
        return st.actions[t]

    def process_env_step(self, rewards, dones, infos):
        st, t = self.storage, self.storage.step
        if self._join_pending:
            self._side.join(); self._join_pending = False
        tout = infos.get('time_outs') if isinstance(infos, dict) else None
        d8 = dones.view(torch.uint8) if dones.dtype == torch.bool else dones.to(torch.uint8)
        t8 = None if tout is None else (tout.view(torch.uint8) if tout.dtype == torch.bool else tout.to(torch.uint8))
        _ops.call("go2_process_env_step", _ops.ptr(rewards), _ops.ptr(d8), _ops.ptr(t8), _ops.ptr(st.values[t]), _ops.ptr(st.rewards[t]),
                  _ops.ptr(st.dones[t]), st.num_envs, self.gamma, 0)
        st.step += 1
        self.transition.clear()
        self.actor_critic.reset(dones)

    def compute_returns(self, last_critic_obs):
        self.actor_critic.evaluate(last_critic_obs, out=self._last_values)
        self.storage.compute_returns(self._last_values, self.gamma, self.lam, reduce_stats=self._reduce_adv_stats if self.world_size > 1 else None)

    def _reduce_adv_stats(self, stats):
        return dist_utils.allreduce_adv_stats(stats, self.storage.num_envs * self.storage.num_transitions_per_env)

    # ---- update ----------------------------------------------------------------------------------------------
    def update(self, indices=None, fetch=True):
        """fetch=False: no host read of the logged means at the end (an un-logged iteration; the host runs ahead of the device, the learning rate
        lives on the device anyway) -> returns None and leaves `learning_rate` at its last fetched value."""
        st, ac = self.storage, self.actor_critic
        mb, A = self.mini_batch_size, st.actions.shape[-1]
        sh_w = (st.privileged_observations if st.privileged_observations is not None else st.observations).shape[-1]
        if indices is None:
            indices = torch.randperm(self.num_mini_batches * mb, device=self.device)
        tc = _ops.use_tc()
        pads = {"obs": _ops.pad_in(st.observations.shape[-1]), "critic_obs": _ops.pad_in(sh_w)} if tc else {}
        sh = st.shuffled(indices, pads)
        total = indices.numel()
        self._log.zero_()
        self._sh, self._total, self._tc = sh, total, tc
        if self.world_size == 1:
            self._graphs.run("update", self._update_body)
        elif (self._red is not None or self._dist_graph) and "update_dist" not in self._graphs._failed:
            # the 20 optimiser steps INCLUDING their NCCL all-reduces as one CUDA graph (NCCL collectives are capturable); a failed capture
            # falls back to the segmented path below for the rest of the run
            self._graphs.run("update_dist", self._update_body_dist)
        else:
            for epoch in range(self.num_learning_epochs):
                for i in range(self.num_mini_batches):
                    self._graphs.run(("grad", i), lambda: self._grad_part(i))
                    self._allreduce_grads()           # one collective per optimiser step: flat gradient + the scalar tail (KL / loss sums)
                    self._graphs.run("step", self._step_part)
                    ac.actor_engine.mark_dirty(); ac.critic_engine.mark_dirty()     # replays skip the Python side of _step_part
        ac.actor_engine.mark_dirty(); ac.critic_engine.mark_dirty()
        self._opt_step += self.num_learning_epochs * self.num_mini_batches
        num_updates = self.num_learning_epochs * self.num_mini_batches
        if not fetch:
            st.clear()
            return None
        log = self._log.tolist()          # the single host sync of update()
        self.learning_rate = log[3]
        st.clear()
        return log[0] / num_updates, log[1] / num_updates

    def _update_body(self):
        for epoch in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                self._grad_part(i)
                self._step_part()

    def _update_body_dist(self):
        for epoch in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                self._grad_part(i)
                self._allreduce_grads()           # one collective per optimiser step: flat gradient + the scalar tail (KL / loss sums)
                self._step_part()

    def _grad_part(self, i):
        """Forward, loss and backward of mini-batch i: fills the flat gradient and the scalar tail."""
        st, ac = self.storage, self.actor_critic
        sh, total, tc = self._sh, self._total, self._tc
        mb, A = self.mini_batch_size, st.actions.shape[-1]
        inv_count = 1.0 / (mb * self.world_size)
        s = slice(i * mb, (i + 1) * mb)
        obs_b, cobs_b = sh["obs"][s], sh["critic_obs"][s]
        # padded gathers carry a 1 in their first padding column (the bias-gradient column of the row-major wgrad)
        sd = self._side
        sd.fork()
        with sd:
            ac.critic_engine.forward(cobs_b, cobs_b.shape[1], mb, self._val_b, 1, train=True, x_ones=tc)
        ac.actor_engine.forward(obs_b, obs_b.shape[1], mb, self._mu_b, A, train=True, x_ones=tc)
        sd.join()
        _ops.call("go2_ppo_loss", _ops.ptr(self._mu_b), _ops.ptr(ac.std.data), _ops.ptr(self._val_b), _ops.ptr(sh["actions"][s]),
                  _ops.ptr(sh["old_logp"][s]), _ops.ptr(sh["adv"][s]), _ops.ptr(sh["values"][s]), _ops.ptr(sh["returns"][s]),
                  _ops.ptr(sh["old_mu"][s]), _ops.ptr(sh["old_sigma"][s]), _ops.ptr(self._dmu), 0,
                  _ops.ptr(self._dval), _ops.ptr(self._scal), mb, A, self.clip_param, self.value_loss_coef, self.entropy_coef,
                  int(self.use_clipped_value_loss), inv_count, mb, inv_count, inv_count)
        sd.fork()
        with sd:
            ac.critic_engine.backward(self._dval, 1)
        ac.actor_engine.backward(self._dmu, A)
        ac._gviews["std"].copy_(self._scal[4:4 + A])
        sd.join()

    def _step_part(self):
        """KL-adaptive learning rate (ppo.py:139-151) + clip + Adam on the (all-reduced) flat gradient."""
        ac = self.actor_critic
        mb = self.mini_batch_size
        adaptive = self.desired_kl is not None and self.schedule == 'adaptive'
        red = self._red         # the exchanged sums (same layout) when the gradient went through the NVLink all-reduce kernel
        scal, grads = (red.out_tail, red.out_grads) if red is not None else (self._scal, ac.flat_grads)
        _ops.call("go2_kl_adaptive_lr", _ops.ptr(scal), float(mb * self.world_size), float(self.desired_kl) if adaptive else -1.0,
                  _ops.ptr(self._lr), _ops.ptr(self._log), float(mb * self.world_size), 1.0)
        _ops.call("go2_adam_clip_step", _ops.ptr(ac.flat_params), _ops.ptr(grads), _ops.ptr(self.exp_avg), _ops.ptr(self.exp_avg_sq),
                  ac.flat_params.numel(), self.max_grad_norm, _ops.ptr(self._lr), 1.0, _ops.ptr(self._scratch))
        ac.actor_engine.mark_dirty(); ac.critic_engine.mark_dirty()

    def _allreduce_grads(self):
        if self._red is not None:
            self._red.allreduce(0, dist_utils.TAIL + self._red.n)       # [scalar tail | flat gradient] in one launch
        else:
            self._comm = dist_utils.allreduce_grads_and_tail(self.actor_critic.flat_grads, self._scal[:4], getattr(self, "_comm", None))

    # ---- checkpoint interop (torch.optim.Adam layout, ppo.py:67) ---------------------------------------------
    def optimizer_state_dict(self):
        state, offs = {}, self.actor_critic._offsets
        for i, (name, p) in enumerate(self.actor_critic.named_parameters()):
            k, off = p.numel(), offs[name]
            state[i] = {"step": torch.tensor(float(self._opt_step)), "exp_avg": self.exp_avg[off:off + k].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + k].view(p.shape).clone()}
        group = dict(adam_group_template(), lr=self.learning_rate, params=list(range(len(state))))
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd):
        offs = self.actor_critic._offsets
        for i, (name, p) in enumerate(self.actor_critic.named_parameters()):
            k, off = p.numel(), offs[name]
            s = sd["state"].get(i)
            if s is not None:
                self.exp_avg[off:off + k].copy_(s["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + k].copy_(s["exp_avg_sq"].reshape(-1))
                self._opt_step = int(float(s["step"]))
                self._lr[1] = float(self._opt_step)
        if sd.get("param_groups"):
            self.learning_rate = float(sd["param_groups"][0]["lr"])
            self._lr[0] = self.learning_rate

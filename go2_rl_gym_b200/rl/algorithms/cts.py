"""CTS and MoECTS — Concurrent Teacher-Student PPO (drop-ins for rsl_rl/algorithms/cts.py:39-285 and moe_cts.py:40-234).

Same constructor arguments and methods.  75 % of the envs (i % 4 != 0) act through the TEACHER encoder (privileged obs), the rest
through the STUDENT encoder (5-frame observation history, evaluated without gradient in pass 1); transitions are stored teacher
first.  update(): pass 1 = PPO on [latent | obs] with optimizer 1 (teacher encoder + critic + actor + std, KL-adaptive LR);
pass 2 = latent reconstruction (+ MoE load balance) on the student rows with optimizer 2.  Everything runs on the library's
kernels; the two Adam states are two contiguous segments of one flat vector."""
import os

import torch
import torch.distributed as dist

from .. import _ops, dist_utils
from .._ops import call, ptr
from ..storage.rollout_storage_cts import RolloutStorageCTS
from .ppo import adam_group_template


_CRITIC_JOIN_EARLY = os.environ.get("GO2_CRITIC_JOIN", "late") == "early"

class CTS:
    def __init__(self, model, num_envs, history_length, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95,
                 value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, student_encoder_learning_rate=1e-3, max_grad_norm=1.0,
                 use_clipped_value_loss=True, schedule="fixed", desired_kl=0.01, teacher_env_ratio=0.75, device='cpu', load_balance_coef=0.0,
                 seed=0, env_offset=0):
        self.device = device
        self.desired_kl, self.schedule, self.learning_rate = desired_kl, schedule, learning_rate
        self.student_encoder_learning_rate = student_encoder_learning_rate
        self.history_length = history_length
        self.model = model
        self.storage = None
        self.transition = RolloutStorageCTS.Transition()
        self.clip_param, self.num_learning_epochs, self.num_mini_batches = clip_param, num_learning_epochs, num_mini_batches
        self.value_loss_coef, self.entropy_coef, self.load_balance_coef = value_loss_coef, entropy_coef, load_balance_coef
        self.gamma, self.lam, self.max_grad_norm, self.use_clipped_value_loss = gamma, lam, max_grad_norm, use_clipped_value_loss
        self.teacher_num_envs = max(int(num_envs * teacher_env_ratio), 1)
        self.student_num_envs = num_envs - self.teacher_num_envs
        student_env_ratio = 1 - teacher_env_ratio
        k = int(1 / student_env_ratio)
        # global env ids decide the role (cts.py:93-97), so a shard of a multi-GPU run keeps the single-process assignment
        ids = torch.arange(num_envs)
        gids = ids + env_offset
        self.teacher_env_idxs = ids[gids % k != 0].to(device)
        self.student_env_idxs = ids[gids % k == 0].to(device)
        assert len(self.teacher_env_idxs) == self.teacher_num_envs, f"{len(self.teacher_env_idxs)=} != {self.teacher_num_envs=}"
        assert len(self.student_env_idxs) == self.student_num_envs, f"{len(self.student_env_idxs)=} != {self.student_num_envs=}"
        self.perm = torch.cat([self.teacher_env_idxs, self.student_env_idxs]).contiguous()           # storage row -> env
        self.inv_perm = torch.empty_like(self.perm)
        self.inv_perm[self.perm] = torch.arange(num_envs, device=device)                             # env -> storage row
        self.seed, self.env_offset = int(seed), int(env_offset)
        self._act_step = 0
        self._dev_steps = None
        self.world_size = dist_utils.world_size()
        # GO2_DIST_GRAPH=1 (opt-in): capture the all-reduces inside the update graph.  Measured on 2 B200: 16.9 vs 17.6 ms / iteration (go2),
        # 35.3 vs 37.4 ms (go2_moe_cts), identical parameters on all ranks - but the process then hangs in destroy_process_group() / interpreter
        # exit while graphs holding NCCL kernels are alive (profiles/r01m_dist_graph_check.txt), so the segmented path stays the default.
        self._dist_graph = os.environ.get("GO2_DIST_GRAPH", "0") == "1"
        self.optimizer1 = self.optimizer2 = None

    # ---- set-up ------------------------------------------------------------------------------------------------------
    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        dev, T = self.device, num_transitions_per_env
        self.storage = RolloutStorageCTS(num_envs, self.teacher_num_envs, self.history_length, T, actor_obs_shape, critic_obs_shape, action_shape, dev)
        nb = self.num_mini_batches
        self.tm, self.sm = self.teacher_num_envs * T // nb, self.student_num_envs * T // nb
        self.mb = self.tm + self.sm
        m = self.model
        m.flatten_(dev, max(num_envs, self.mb), self.mb, self.tm, self.sm)
        A, D, N = action_shape[0], m.latent_dim, num_envs
        z = lambda *s: torch.zeros(*s, device=dev)
        self.exp_avg, self.exp_avg_sq = z(m.flat_params.numel()), z(m.flat_params.numel())
        self._lr1, self._lr2 = z(4), z(4)
        self._lr1[0], self._lr2[0] = float(self.learning_rate), float(self.student_encoder_learning_rate)
        # env-sharded over GPUs: symmetric gradient buffer + the library's NVLink all-reduce kernel (see PPO.init_storage)
        self._red = dist_utils.reducer_for(m.flat_grads) if self.world_size > 1 else None
        self._scal, self._log, self._log2 = (self._red.tail if self._red is not None else z(20)), z(5), z(2)
        self._scratch, self._acc = z(1025), z(1)
        rows = max(N, self.mb)
        self._lat = z(rows, D)
        self._lat_t = z(max(self.sm, 1), D)          # teacher latents of the student rows (pass 2 target)
        # Neither the student encoder (pass 1) nor the teacher encoder (pass 2) changes during the pass that only READS it, and every epoch walks the
        # same mini-batches (rollout_storage_cts.py:168-216): the no-grad latents are computed once per mini-batch and update(), not once per epoch
        self._lat_s_cache = z(nb, max(self.sm, 1), D)
        self._lat_t_cache = z(nb, max(self.sm, 1), D)
        self._xa = z(rows, _ops.pad_in(D + actor_obs_shape[0]))
        self._xc = z(rows, _ops.pad_in(D + critic_obs_shape[0]))
        self._mu, self._val = z(rows, A), z(rows, 1)
        self._dmu, self._dval = z(self.mb, A), z(self.mb + 4, 1)[:self.mb]
        self._dlat = z(self.mb, D)
        self._dls = z(max(self.sm, 1), D)
        self._last_values = z(N, 1)
        self._hist_p = z(N, self.history_length * actor_obs_shape[0])
        self._rew_p, self._actions_env = z(N), z(N, A)
        self._graphs = _ops.GraphSet()
        self._side = _ops.SideStream(dev)          # independent chains (teacher / student, actor / critic) run side by side: _ops.SideStream
        self._join_pending = False

    def test_mode(self):
        self.model.eval()

    def train_mode(self):
        self.model.train()

    # ---- shared forward pieces -----------------------------------------------------------------------------------------
    def _latents(self, priv, hist, n_t, n_s, train_teacher=False, cache=None):
        """self._lat[0:n_t] = teacher latent of the first n_t rows, self._lat[n_t:n_t+n_s] = student latent of the rest (no grad).
        cache = i: the student latents of mini-batch i were computed by _pre1() (same weights, same rows, every epoch)."""
        m = self.model
        sd = self._side
        sd.fork()
        with sd:         # the teacher's rows next to the student's
            if n_t:
                m.teacher_latent(priv[:n_t], n_t, self._lat[:n_t], train=train_teacher, x_ones=train_teacher and _ops.use_tc())
        if n_s and cache is not None:
            self._lat[n_t:n_t + n_s].copy_(self._lat_s_cache[cache, :n_s])
        elif n_s:
            m.student.forward(hist[n_t:n_t + n_s], n_s, self._lat[n_t:n_t + n_s])
        sd.join()

    def _heads(self, obs, priv, M, train=False, values_out=None):
        m, D = self.model, self.model.latent_dim
        tc = _ops.use_tc()
        # go2_concat2 writes a 1 into the first padding column of its output (bias-gradient column of the row-major wgrad)
        call("go2_concat2", ptr(self._lat), D, D, ptr(obs), m.num_obs, obs.stride(0), ptr(self._xa), self._xa.shape[1], 0, M)
        call("go2_concat2", ptr(self._lat), D, D, ptr(priv), m.num_critic_obs, priv.stride(0), ptr(self._xc), self._xc.shape[1], 0, M)
        ones = self._xa.shape[1] > D + m.num_obs
        sd = self._side
        sd.fork()
        with sd:         # the critic next to the actor
            m.critic_engine.forward(self._xc, self._xc.shape[1], M, self._val[:M], 1, train=train, x_ones=self._xc.shape[1] > D + m.num_critic_obs)
            if values_out is not None:       # rollout: the value is first needed by process_env_step — the critic also runs next to the env step
                values_out.copy_(self._val[:M])
        m.actor_engine.forward(self._xa, self._xa.shape[1], M, self._mu[:M], m.num_actions, train=train, x_ones=ones)
        if values_out is not None:
            if _CRITIC_JOIN_EARLY:           # A/B: the critic rejoins before the env step (it then only runs beside the actor and the sampling)
                sd.join()
            else:
                self._join_pending = True
        else:
            sd.join()

    # ---- rollout -------------------------------------------------------------------------------------------------------
    def begin_rollout(self, T):
        """See PPO.begin_rollout: device-resident sampling counters for a graph-replayed rollout."""
        self._dev_steps = _ops.upload_steps(getattr(self, "_dev_steps_buf", None), self._act_step, T, self.device)
        self._dev_steps_buf = self._dev_steps
        self._act_step += T
        self.model.mark_dirty()

    def end_rollout(self, T):
        self._dev_steps = None
        self.storage.step = T

    def act(self, obs, privileged_obs, history):
        st, t, m = self.storage, self.storage.step, self.model
        if t >= st.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        N, A = st.num_envs, st.actions.shape[-1]
        gather = lambda src, w, dst: call("go2_gather_rows", ptr(src.contiguous()), w, ptr(self.perm), ptr(dst), w, 0, N)
        gather(obs, obs.shape[1], st.observations[t])
        gather(privileged_obs, privileged_obs.shape[1], st.privileged_observations[t])
        gather(history, history.shape[1], st.history[t])
        nt, ns = self.teacher_num_envs, self.student_num_envs
        self._latents(st.privileged_observations[t], st.history[t], nt, ns)
        if type(self)._heads is CTS._heads and self._dev_steps is not None:      # graph-replayed rollout: critic + value copy on the side stream, joined
            self._heads(st.observations[t], st.privileged_observations[t], N, values_out=st.values[t])      # in process_env_step
        else:
            self._heads(st.observations[t], st.privileged_observations[t], N)
            st.values[t].copy_(self._val[:N])
        self._sample(t, N, A)
        call("go2_gather_rows", ptr(st.actions[t]), A, ptr(self.inv_perm), ptr(self._actions_env), A, 0, N)   # back to env order
        return self._actions_env

    def _sample(self, t, N, A):
        """actions ~ N(mu, std), their log-prob and the (mu, sigma) rows of transition t (cts.py:114-131)."""
        st, m = self.storage, self.model
        if self._dev_steps is not None:      # rollout opened by begin_rollout(): the Philox step counter comes from device memory
            call("go2_sample_actions_dev", ptr(self._mu), ptr(m.std.data), ptr(st.actions[t]), ptr(st.actions_log_prob[t]), ptr(st.mu[t]), ptr(st.sigma[t]),
                 N, A, self.seed, self._dev_steps.data_ptr() + 4 * t, self.env_offset)
        else:
            self._act_step += 1
            call("go2_sample_actions", ptr(self._mu), ptr(m.std.data), ptr(st.actions[t]), ptr(st.actions_log_prob[t]), ptr(st.mu[t]), ptr(st.sigma[t]),
                 N, A, self.seed, self._act_step, self.env_offset)

    def process_env_step(self, rewards, dones, infos):
        st, t = self.storage, self.storage.step
        if self._join_pending:
            self._side.join(); self._join_pending = False
        tout = infos.get('time_outs') if isinstance(infos, dict) else None
        d8 = dones.view(torch.uint8) if dones.dtype == torch.bool else dones.to(torch.uint8)
        t8 = None if tout is None else (tout.view(torch.uint8) if tout.dtype == torch.bool else tout.to(torch.uint8))
        call("go2_process_env_step", ptr(rewards), ptr(d8), ptr(t8), ptr(st.values[t]), ptr(st.rewards[t]), ptr(st.dones[t]), st.num_envs, self.gamma,
             ptr(self.perm))
        st.step += 1
        self.transition.clear()
        self.model.reset(dones)

    def compute_returns(self, last_privileged_obs, last_history):
        N = self.storage.num_envs
        call("go2_gather_rows", ptr(last_privileged_obs.contiguous()), last_privileged_obs.shape[1], ptr(self.perm), ptr(self._xc_p(N)), last_privileged_obs.shape[1], 0, N)
        call("go2_gather_rows", ptr(last_history.contiguous()), last_history.shape[1], ptr(self.perm), ptr(self._hist_p), last_history.shape[1], 0, N)
        priv_p = self._xc_p(N)
        self._latents(priv_p, self._hist_p, self.teacher_num_envs, self.student_num_envs)
        D, m = self.model.latent_dim, self.model
        call("go2_concat2", ptr(self._lat), D, D, ptr(priv_p), m.num_critic_obs, priv_p.stride(0), ptr(self._xc), self._xc.shape[1], 0, N)
        m.critic_engine.forward(self._xc, self._xc.shape[1], N, self._last_values, 1)
        self.storage.compute_returns(self._last_values, self.gamma, self.lam,
                                     reduce_stats=(lambda s: dist_utils.allreduce_adv_stats(s, N * self.storage.num_transitions_per_env)) if self.world_size > 1 else None)

    def _xc_p(self, N):
        if not hasattr(self, "_priv_p"):
            self._priv_p = torch.zeros(N, self.model.num_critic_obs, device=self.device)
        return self._priv_p

    # ---- update --------------------------------------------------------------------------------------------------------
    def update(self, teacher_perm=None, student_perm=None, fetch=True):
        """fetch=False: no host read of the logged means at the end (see PPO.update)."""
        st, m = self.storage, self.model
        idx, tm, sm = st.batch_indices(self.num_mini_batches, teacher_perm, student_perm)
        tc = _ops.use_tc()
        pads = {"critic_obs": _ops.pad_in(m.num_critic_obs), "history": _ops.pad_in(st.history.shape[-1])} if tc else {}
        self._sh = st.shuffled(idx, pads)
        self._total = idx.numel()
        self._log.zero_(); self._log2.zero_()
        if self.world_size == 1:
            self._graphs.run("update", self._update_body)
        elif (self._red is not None or self._dist_graph) and "update_dist" not in self._graphs._failed:
            self._graphs.run("update_dist", self._update_body_dist)       # both passes with their gradient exchanges as one CUDA graph
        else:
            ws, n1, G = self.world_size, m.n1, m.flat_grads
            self._graphs.run("pre1", self._pre1)
            for epoch in range(self.num_learning_epochs):
                for i in range(self.num_mini_batches):
                    self._graphs.run(("grad1", i), lambda: self._grad1(i))
                    self._comm1 = dist_utils.allreduce_grads_and_tail(G[:n1], self._scal, getattr(self, "_comm1", None))
                    self._graphs.run("step1", self._step1)
                    m.mark_dirty()                    # replays skip the Python side of the step
            if self.sm > 0:
                self._graphs.run("pre2", self._pre2)
                for epoch in range(self.num_learning_epochs):
                    for i in range(self.num_mini_batches):
                        self._graphs.run(("grad2", i), lambda: self._grad2(i))
                        dist.all_reduce(G[n1:])
                        self._graphs.run("step2", self._step2)
                        m.mark_dirty()
        m.mark_dirty()
        n = self.num_learning_epochs * self.num_mini_batches
        if not fetch:
            st.clear()
            return None
        if self.world_size > 1:      # pass 2's logged means are per-rank sums (pass 1's travel in the gradient's scalar tail): one small collective per iteration
            self._reduce_logs()
        log, log2 = self._log.tolist(), self._log2.tolist()      # the single host sync of update()
        self.learning_rate = log[3]
        st.clear()
        return self._losses(log, log2, n)

    _moe = False

    def _reduce_logs(self):
        dist.all_reduce(self._log2)
        self._log2 /= self.world_size

    def _losses(self, log, log2, n):
        """update()'s return value: mean value / surrogate / entropy / latent loss (cts.py:280-285) + the student's load-balance loss (moe_cts.py:229-234)."""
        out = (log[0] / n, log[1] / n, log[4] / n, log2[0] / n)
        return out + (log2[1] / n,) if self._moe else out

    def _pre1(self):
        """student latents (no grad) of every mini-batch's student rows, once per update()"""
        sh, tm, sm, mb = self._sh, self.tm, self.sm, self.mb
        for i in range(self.num_mini_batches if sm > 0 else 0):
            self.model.student.forward(sh["history"][i * mb + tm:(i + 1) * mb], sm, self._lat_s_cache[i, :sm])

    def _pre2(self):
        """teacher latents (the target of pass 2, no grad) of every mini-batch's student rows, once per update(), after pass 1 has moved the teacher"""
        sh, tm, sm, mb = self._sh, self.tm, self.sm, self.mb
        for i in range(self.num_mini_batches):
            self.model.teacher_latent(sh["critic_obs"][i * mb + tm:(i + 1) * mb], sm, self._lat_t_cache[i, :sm])

    def _update_body(self):
        # pass 1: PPO on teacher + student rows, optimizer 1 (moe_cts.py:114-195)
        self._pre1()
        for epoch in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                self._grad1(i)
                self._step1()
        # pass 2: student encoder towards the (updated) teacher latent, optimizer 2 (moe_cts.py:197-224)
        if self.sm == 0:
            return
        self._pre2()
        for epoch in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                self._grad2(i)
                self._step2()

    def _update_body_dist(self):
        m, n1, G = self.model, self.model.n1, self.model.flat_grads
        self._pre1()
        for epoch in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                self._grad1(i)
                if self._red is not None:
                    self._red.allreduce(0, dist_utils.TAIL + n1)          # [scalar tail | pass-1 gradient] in one launch
                else:
                    self._comm1 = dist_utils.allreduce_grads_and_tail(G[:n1], self._scal, getattr(self, "_comm1", None))
                self._step1()
        if self.sm == 0:
            return
        self._pre2()
        for epoch in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                self._grad2(i)
                if self._red is not None:
                    self._red.allreduce(dist_utils.TAIL + n1, m.n2)
                else:
                    dist.all_reduce(G[n1:])
                self._step2()

    def _grad1(self, i):
        st, m, sh, total = self.storage, self.model, self._sh, self._total
        tm, sm, mb, A, D = self.tm, self.sm, self.mb, st.actions.shape[-1], m.latent_dim
        tc, ws = _ops.use_tc(), self.world_size
        s = slice(i * mb, (i + 1) * mb)
        obs_b, priv_b, hist_b = sh["obs"][s], sh["critic_obs"][s], sh["history"][s]
        self._latents(priv_b, hist_b, tm, sm, train_teacher=True, cache=i)
        self._heads(obs_b, priv_b, mb, train=True)
        call("go2_ppo_loss", ptr(self._mu), ptr(m.std.data), ptr(self._val), ptr(sh["actions"][s]), ptr(sh["old_logp"][s]), ptr(sh["adv"][s]),
             ptr(sh["values"][s]), ptr(sh["returns"][s]), ptr(sh["old_mu"][s]), ptr(sh["old_sigma"][s]), ptr(self._dmu),
             0, ptr(self._dval), ptr(self._scal), mb, A, self.clip_param, self.value_loss_coef, self.entropy_coef,
             int(self.use_clipped_value_loss), 1.0 / (mb * ws), tm, 1.0 / (tm * ws), 1.0 / (max(sm, 1) * ws))
        sd = self._side
        sd.fork()
        with sd:
            m.critic_engine.backward(self._dval, 1)
        m.actor_engine.backward(self._dmu, A)
        # d loss / d latent = first D columns of the actor's input gradient, teacher rows only (student latents carry no grad)
        m.teacher_backward(m.actor_engine.dx, m.actor_engine.kpad0, self._lat, D, tm)
        m._gviews["std"].copy_(self._scal[4:4 + A])
        sd.join()

    def _step1(self):
        m, ws = self.model, self.world_size
        adaptive = self.desired_kl is not None and self.schedule == 'adaptive'
        red = self._red      # the exchanged sums (same layout) when the gradient went through the NVLink all-reduce kernel
        scal, grads = (red.out_tail, red.out_grads) if red is not None else (self._scal, m.flat_grads)
        call("go2_kl_adaptive_lr", ptr(scal), float(self.mb * ws), float(self.desired_kl) if adaptive else -1.0, ptr(self._lr1), ptr(self._log),
             float(self.tm * ws), float(max(self.sm, 1) * ws))
        call("go2_adam_clip_step", ptr(m.flat_params), ptr(grads), ptr(self.exp_avg), ptr(self.exp_avg_sq), m.n1, self.max_grad_norm, ptr(self._lr1),
             1.0, ptr(self._scratch))
        for e in m.pass1_engines():
            e.mark_dirty()

    def _grad2(self, i):
        m, sh, total = self.model, self._sh, self._total
        tm, sm, mb, D = self.tm, self.sm, self.mb, m.latent_dim
        tc = _ops.use_tc()
        s = slice(i * mb + tm, (i + 1) * mb)
        priv_b, hist_b = sh["critic_obs"][s], sh["history"][s]
        m.student.forward(hist_b, sm, self._lat[:sm], train=True, x_ones=tc)
        call("go2_latent_loss", ptr(self._lat), ptr(self._lat_t_cache[i]), ptr(self._dls), ptr(self._acc), sm, D)      # target: _pre2()
        m.student.backward(self._dls, self.load_balance_coef)
        call("go2_cts_log", ptr(self._acc), ptr(m.student.usage) if self._moe else 0, ptr(self._log2), sm * D, m.student.E if self._moe else 1)

    def _step2(self):
        m, n1 = self.model, self.model.n1
        # with world_size > 1 the gradient arrives SUM-reduced: fold the 1/world_size of the mean loss into the step
        grads = self._red.out_grads if self._red is not None else m.flat_grads
        call("go2_adam_clip_step", ptr(m.flat_params) + 4 * n1, ptr(grads) + 4 * n1, ptr(self.exp_avg) + 4 * n1, ptr(self.exp_avg_sq) + 4 * n1, m.n2,
             self.max_grad_norm, ptr(self._lr2), 1.0 / self.world_size, ptr(self._scratch))
        for e in m.student.engines():
            e.mark_dirty()
        m.student.mark_dirty()

    # ---- checkpoint interop (two torch.optim.Adam state dicts, on_policy_runner_cts.py:287-294) --------------------------
    def _opt_state(self, names, lr, lr_state):
        named = dict(self.model.named_parameters())
        state = {}
        for i, k in enumerate(names):
            p = named[k]
            n, off = p.numel(), self.model._offsets[k]
            state[i] = {"step": lr_state[1].detach().cpu().clone(), "exp_avg": self.exp_avg[off:off + n].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view(p.shape).clone()}
        # one param_group per top-level module, in the order the reference hands them to torch.optim.Adam (cts.py:73-80: teacher_encoder, critic,
        # actor, std for optimizer 1; the student encoder alone for optimizer 2), each with the installed torch's full Adam key set (ppo.adam_group_template), so that the reference's
        # Optimizer.load_state_dict (which checks the group count and sizes) accepts a checkpoint written here
        groups, prev = [], None
        for i, k in enumerate(names):
            top = k.split(".")[0]
            if top != prev:
                groups.append(dict(adam_group_template(), lr=lr, params=[]))
                prev = top
            groups[-1]["params"].append(i)
        return {"state": state, "param_groups": groups}

    def optimizer1_state_dict(self):
        return self._opt_state(self.model.seg1_names, self.learning_rate, self._lr1)

    def optimizer2_state_dict(self):
        return self._opt_state(self.model.seg2_names, self.student_encoder_learning_rate, self._lr2)

    def load_optimizer_state_dicts(self, sd1, sd2):
        named = dict(self.model.named_parameters())
        for sd, names, lrs in ((sd1, self.model.seg1_names, self._lr1), (sd2, self.model.seg2_names, self._lr2)):
            if sd is None:
                continue
            for i, k in enumerate(names):
                n, off = named[k].numel(), self.model._offsets[k]
                s = sd["state"].get(i)
                if s is not None:
                    self.exp_avg[off:off + n].copy_(s["exp_avg"].reshape(-1))
                    self.exp_avg_sq[off:off + n].copy_(s["exp_avg_sq"].reshape(-1))
                    lrs[1] = float(s["step"])
            if sd.get("param_groups"):
                lrs[0] = float(sd["param_groups"][0]["lr"])
        self.learning_rate = float(self._lr1[0])


class MoECTS(CTS):
    _moe = True

    def __init__(self, model, num_envs, history_length, load_balance_coef=0.01, **kwargs):
        super().__init__(model, num_envs, history_length, load_balance_coef=load_balance_coef, **kwargs)


class MoENGCTS(MoECTS):
    """MoENGCTS (rsl_rl/algorithms/moe_ng_cts.py:40-234): identical to MoECTS; the no-goal column selection lives in the model's student encoder."""


class ACMoECTS(CTS):
    """ACMoECTS (rsl_rl/algorithms/ac_moe_cts.py:40-277): CTS with a mixture-of-experts actor and value experts weighted by the actor's
    gate (ActorCriticACMoECTS); pass 1 adds the load-balance loss of the actor's gate over the whole mini-batch (:225-235); pass 2 is plain
    CTS (MLP student).  compute_returns takes the last observations too (:136-142): the value needs the actor's gate.

    The reference evaluates the gate twice per row (act and evaluate); both see the same input and weights, so their gradients are summed
    into ONE backward of the gate here.  Gradients the reference lets flow into the student encoder during pass 1 are discarded there
    (optimizer 2 zeroes them before pass 2, :261-262) and never computed here."""
    _moe = False        # pass 2: no student load-balance term
    _n_losses = 5

    def __init__(self, model, num_envs, history_length, load_balance_coef=0.01, **kwargs):
        super().__init__(model, num_envs, history_length, load_balance_coef=load_balance_coef, **kwargs)

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        super().init_storage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape)
        dev = self.device
        self._log3, self._zero1 = torch.zeros(2, device=dev), torch.zeros(1, device=dev)
        self._obs_p = torch.zeros(num_envs, actor_obs_shape[0], device=dev)

    def _heads(self, obs, priv, M, train=False):
        m, D = self.model, self.model.latent_dim
        call("go2_concat2", ptr(self._lat), D, D, ptr(obs), m.num_obs, obs.stride(0), ptr(self._xa), self._xa.shape[1], 0, M)
        call("go2_concat2", ptr(self._lat), D, D, ptr(priv), m.num_critic_obs, priv.stride(0), ptr(self._xc), self._xc.shape[1], 0, M)
        m.heads_forward(self._xa, self._xc, M, self._mu, self._val, train=train)

    def compute_returns(self, last_obs, last_privileged_obs, last_history):
        N, m = self.storage.num_envs, self.model
        g = lambda src, dst: call("go2_gather_rows", ptr(src.contiguous()), src.shape[1], ptr(self.perm), ptr(dst), src.shape[1], 0, N)
        g(last_obs, self._obs_p); g(last_privileged_obs, self._xc_p(N)); g(last_history, self._hist_p)
        self._latents(self._priv_p, self._hist_p, self.teacher_num_envs, self.student_num_envs)
        self._heads(self._obs_p, self._priv_p, N)
        self._last_values.copy_(self._val[:N])
        self.storage.compute_returns(self._last_values, self.gamma, self.lam,
                                     reduce_stats=(lambda s: dist_utils.allreduce_adv_stats(s, N * self.storage.num_transitions_per_env)) if self.world_size > 1 else None)

    def update(self, teacher_perm=None, student_perm=None, fetch=True):
        self._log3.zero_()
        return super().update(teacher_perm, student_perm, fetch=fetch)

    def _grad1(self, i):
        st, m, sh = self.storage, self.model, self._sh
        tm, sm, mb, A, D = self.tm, self.sm, self.mb, st.actions.shape[-1], m.latent_dim
        ws = self.world_size
        s = slice(i * mb, (i + 1) * mb)
        obs_b, priv_b, hist_b = sh["obs"][s], sh["critic_obs"][s], sh["history"][s]
        self._latents(priv_b, hist_b, tm, sm, train_teacher=True, cache=i)
        self._heads(obs_b, priv_b, mb, train=True)
        call("go2_ppo_loss", ptr(self._mu), ptr(m.std.data), ptr(self._val), ptr(sh["actions"][s]), ptr(sh["old_logp"][s]), ptr(sh["adv"][s]),
             ptr(sh["values"][s]), ptr(sh["returns"][s]), ptr(sh["old_mu"][s]), ptr(sh["old_sigma"][s]), ptr(self._dmu),
             0, ptr(self._dval), ptr(self._scal), mb, A, self.clip_param, self.value_loss_coef, self.entropy_coef,
             int(self.use_clipped_value_loss), 1.0 / (mb * ws), tm, 1.0 / (tm * ws), 1.0 / (max(sm, 1) * ws))
        # the gradient is SUM-reduced over the ranks: each rank's load-balance term (on its own rows' mean usage) enters with 1 / world_size
        dx = m.heads_backward(self._dmu, self._dval, mb, self.load_balance_coef / ws)
        m.teacher_backward(dx, dx.shape[1], self._lat, D, tm)           # teacher rows only: student latents carry no gradient in pass 1
        m._gviews["std"].copy_(self._scal[4:4 + A])
        call("go2_cts_log", ptr(self._zero1), ptr(m.actor_head.usage), ptr(self._log3), 1, m.actor_head.E)

    def _reduce_logs(self):
        super()._reduce_logs()
        dist.all_reduce(self._log3)
        self._log3 /= self.world_size

    def _losses(self, log, log2, n):
        actor_lb = self._log3.tolist()[1] / n
        return (log[0] / n, log[1] / n, log[4] / n, log2[0] / n, actor_lb)


class DualMoECTS(ACMoECTS):
    """DualMoECTS (rsl_rl/algorithms/dual_moe_cts.py:40-287): ACMoECTS with the MoE student encoder of MoECTS; update() returns the student's and
    the actor's load-balance losses (:287)."""
    _moe = True

    def _losses(self, log, log2, n):
        out = super()._losses(log, log2, n)
        return out[:4] + (log2[1] / n, out[4])


class MCPCTS(CTS):
    """MCPCTS (rsl_rl/algorithms/mcp_cts.py:40-220): CTS with the multiplicative-compositional actor of ActorCriticMCPCTS.  The action sigma is a
    network output, so sampling, the PPO loss and its backward run on the state-dependent-sigma kernels (csrc/mcp_kernels.cu) and optimizer 1
    holds teacher encoder + critic + actor_mcp (no std parameter, :72-78)."""

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        super().init_storage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape)
        dev, m, A = self.device, self.model, action_shape[0]
        rows = max(num_envs, self.mb)
        self._sigma, self._dsigma = torch.zeros(rows, A, device=dev), torch.zeros(self.mb, A, device=dev)
        self._ng = torch.zeros(rows, m.num_obs_no_goal, device=dev)
        self._xng = torch.zeros(rows, _ops.pad_in(m.ng_dim), device=dev)

    def _heads(self, obs, priv, M, train=False):
        m, D = self.model, self.model.latent_dim
        ng = m.no_goal(obs, M, self._ng)
        call("go2_concat2", ptr(self._lat), D, D, ptr(obs), m.num_obs, obs.stride(0), ptr(self._xa), self._xa.shape[1], 0, M)
        call("go2_concat2", ptr(self._lat), D, D, ptr(ng), m.num_obs_no_goal, ng.stride(0), ptr(self._xng), self._xng.shape[1], 0, M)
        call("go2_concat2", ptr(self._lat), D, D, ptr(priv), m.num_critic_obs, priv.stride(0), ptr(self._xc), self._xc.shape[1], 0, M)
        m.heads_forward(self._xa, self._xng, self._xc, M, self._mu, self._sigma, self._val, train=train)

    def _sample(self, t, N, A):
        st = self.storage
        if self._dev_steps is not None:
            step, d_step = 0, self._dev_steps.data_ptr() + 4 * t
        else:
            self._act_step += 1
            step, d_step = self._act_step, 0
        call("go2_sample_actions_sigma", ptr(self._mu), ptr(self._sigma), ptr(st.actions[t]), ptr(st.actions_log_prob[t]), ptr(st.mu[t]), ptr(st.sigma[t]),
             N, A, self.seed, step, d_step, self.env_offset)

    def _grad1(self, i):
        st, m, sh = self.storage, self.model, self._sh
        tm, sm, mb, A, D = self.tm, self.sm, self.mb, st.actions.shape[-1], m.latent_dim
        ws = self.world_size
        s = slice(i * mb, (i + 1) * mb)
        obs_b, priv_b, hist_b = sh["obs"][s], sh["critic_obs"][s], sh["history"][s]
        self._latents(priv_b, hist_b, tm, sm, train_teacher=True, cache=i)
        self._heads(obs_b, priv_b, mb, train=True)
        call("go2_ppo_loss_sigma", ptr(self._mu), ptr(self._sigma), ptr(self._val), ptr(sh["actions"][s]), ptr(sh["old_logp"][s]), ptr(sh["adv"][s]),
             ptr(sh["values"][s]), ptr(sh["returns"][s]), ptr(sh["old_mu"][s]), ptr(sh["old_sigma"][s]), ptr(self._dmu),
             ptr(self._dsigma), ptr(self._dval), ptr(self._scal), mb, A, self.clip_param, self.value_loss_coef, self.entropy_coef,
             int(self.use_clipped_value_loss), 1.0 / (mb * ws), tm, 1.0 / (tm * ws), 1.0 / (max(sm, 1) * ws))
        dlat = m.heads_backward(self._dmu, self._dsigma, self._dval, mb)
        m.teacher_backward(dlat, dlat.shape[1], self._lat, D, tm)       # teacher rows only: student latents carry no gradient in pass 1

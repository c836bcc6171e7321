"""ActorCriticCTS / ActorCriticMoECTS (drop-ins for rsl_rl/modules/actor_critic_cts.py:18-160 and
actor_critic_moe_cts.py:20-141 + modules/utils.py:24-151).

Same constructor arguments and `state_dict` keys (teacher_encoder.*, student_encoder.* / student_moe_encoder.moe.*,
actor.*, critic.*, std), parameters re-homed in ONE flat vector split into two contiguous optimiser segments:
  segment 1 (optimizer1, cts.py:72-79): teacher_encoder, critic, actor, std        segment 2 (optimizer2): the student encoder
and every forward / backward evaluated by the library's kernels (GEMMs on tcgen05, the small pieces in csrc/cts_kernels.cu)."""
import os

import torch
import torch.nn as nn

from .. import _ops, dist_utils
from .._ops import call, ptr


class MLP(nn.Module):  # parameter container with the reference's key layout (modules/utils.py:51-67)
    def __init__(self, dims, last_activation=False):
        super().__init__()
        layers = []
        for i in range(len(dims) - 1):
            layers.append(nn.Linear(dims[i], dims[i + 1]))
            if i < len(dims) - 2 or last_activation:
                layers.append(nn.ELU())
        self.network = nn.Sequential(*layers)
        self.dims, self.last_activation = list(dims), last_activation

    def forward(self, x):
        return self.network(x)


class L2Norm(nn.Module):
    def forward(self, x):
        return torch.nn.functional.normalize(x, p=2.0, dim=-1)


class Experts(nn.Module):
    def __init__(self, expert_num, input_dim, backbone_hidden_dims, expert_hidden_dim, output_dim):
        super().__init__()
        self.expert_num, self.output_dim = expert_num, output_dim
        self.backbone = MLP([input_dim, *backbone_hidden_dims, expert_num * expert_hidden_dim], last_activation=True)
        self.experts = nn.Conv1d(expert_num * expert_hidden_dim, expert_num * output_dim, kernel_size=1, groups=expert_num)


class MoE(nn.Module):
    def __init__(self, expert_num, input_dim, hidden_dims, output_dim):
        super().__init__()
        self.experts = Experts(expert_num, input_dim, hidden_dims[:-1], hidden_dims[-1], output_dim)
        self.gating_network = nn.Sequential(MLP([input_dim, *hidden_dims[:-1], expert_num]), nn.Softmax(dim=-1))


class StudentMoEEncoder(nn.Module):
    def __init__(self, expert_num, input_dim, hidden_dims, output_dim):
        super().__init__()
        self.norm_layer = L2Norm()
        self.moe = MoE(expert_num, input_dim, hidden_dims, output_dim)


def _seq_mlp(dims):  # ActorCriticCTS keeps plain nn.Sequential keys (actor_critic_cts.py:51-104)
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append(nn.ELU())
    return layers


def _linear_names(prefix, n_layers):
    return [f"{prefix}.{2 * i}" for i in range(n_layers)]


class _CTSBase(nn.Module):
    is_recurrent = False

    # ---- flat storage ---------------------------------------------------------------------------------------------
    def _segments(self):
        raise NotImplementedError

    def flatten_(self, device, max_rows, train_rows_1, train_rows_t, train_rows_s):
        """train_rows_1: mini-batch rows of pass 1; _t / _s: its teacher / student parts."""
        seg1, seg2 = self._segments()
        named = dict(self.named_parameters())
        n1 = sum((named[k].numel() + 3) // 4 * 4 for k in seg1)     # every parameter starts 16-byte aligned (TMA operand rule);
        n2 = sum((named[k].numel() + 3) // 4 * 4 for k in seg2)     # the padding stays zero (zero gradient -> Adam leaves it at zero)
        flat = torch.zeros(n1 + n2, device=device)
        grad = dist_utils.new_flat_grad(n1 + n2, device)    # a symmetric (peer-mapped) buffer when the envs are sharded over GPUs
        self._views, self._gviews, self._offsets, off = {}, {}, {}, 0
        for k in seg1 + seg2:
            p = named[k]
            n = p.numel()
            self._offsets[k] = off
            flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = flat[off:off + n].view(p.shape)
            self._views[k] = p.data
            self._gviews[k] = grad[off:off + n].view(p.shape)
            off += (n + 3) // 4 * 4
        self._flat, self._grad, self.n1, self.n2 = flat, grad, n1, n2
        self.seg1_names, self.seg2_names = seg1, seg2
        self.device = torch.device(device)
        self.history = self.history.to(device)
        self._build_engines(device, max_rows, train_rows_1, train_rows_t, train_rows_s)
        return self

    def _engine(self, names, dims, max_rows, train_rows, **kw):
        W = [self._views[n + ".weight"] for n in names]
        b = [self._views[n + ".bias"] for n in names]
        gW = [self._gviews[n + ".weight"] for n in names]
        gb = [self._gviews[n + ".bias"] for n in names]
        return _ops.MlpEngine(dims, W, b, gW, gb, max_rows, self.device, train_rows=train_rows, **kw)

    @property
    def flat_params(self):
        return self._flat

    @property
    def flat_grads(self):
        return self._grad

    def pass1_engines(self):
        """Engines whose weights optimizer 1 changes (cts.py:72-79)."""
        return [self.teacher_engine, self.actor_engine, self.critic_engine]

    def engines(self):
        return self.pass1_engines() + self.student.engines()

    def mark_dirty(self):
        for e in self.engines():
            e.mark_dirty()
        self.student.mark_dirty()

    def load_state_dict(self, state_dict, strict=True):
        if getattr(self, "_flat", None) is None:
            return super().load_state_dict(state_dict, strict)
        own = dict(self.named_parameters())
        missing = [k for k in own if k not in state_dict]
        unexpected = [k for k in state_dict if k not in own]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing}, unexpected {unexpected}")
        with torch.no_grad():
            for k, p in own.items():
                if k in state_dict:
                    p.data.copy_(state_dict[k].to(p.device))
        self.mark_dirty()
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    # ---- reference API --------------------------------------------------------------------------------------------
    def reset(self, dones=None):
        if dones is not None:
            # masked_fill_ with a broadcast mask: no nonzero() / host sync (the reference's boolean-index assignment, actor_critic_cts.py:100-101)
            self.history.masked_fill_((dones > 0).view(-1, *([1] * (self.history.dim() - 1))), 0.0)

    def forward(self):
        raise NotImplementedError

    def teacher_latent(self, priv, M, out, train=False, x_ones=False):
        """out[M, latent] = L2Norm(teacher_encoder(priv))  (actor_critic_moe_cts.py:116-117)"""
        e = self.teacher_engine
        e.forward(priv, priv.shape[1] if priv.dim() == 2 else 0, M, self._t_pre[:M], self.latent_dim, train=train, x_ones=x_ones)
        call("go2_l2norm_forward", ptr(self._t_pre), self.latent_dim, ptr(out), out.stride(0), ptr(self._t_norm), M, self.latent_dim)

    def teacher_backward(self, dlatent, lddl, latent, ldl, M):
        call("go2_l2norm_backward", ptr(dlatent), lddl, ptr(latent), ldl, ptr(self._t_norm), ptr(self._t_dpre), self.latent_dim, 0, M, self.latent_dim)
        self.teacher_engine.backward(self._t_dpre, self.latent_dim)

    # ---- the reference's module-level API (actor_critic_cts.py:106-160 and its siblings): the algorithm classes above do not need it (their
    # rollout / update paths call the engines directly with fused sampling), external callers written against the reference do
    def _api_buffers(self, M):
        if getattr(self, "_api_rows", 0) < M:
            dev, D = self.device, self.latent_dim
            self._api_lat = torch.zeros(M, D, device=dev)
            self._api_xa = torch.zeros(M, _ops.pad_in(D + self.num_obs), device=dev)
            self._api_xc = torch.zeros(M, _ops.pad_in(D + self.num_critic_obs), device=dev)
            self._api_mu, self._api_sigma, self._api_val = (torch.zeros(M, self.num_actions, device=dev), torch.zeros(M, self.num_actions, device=dev),
                                                            torch.zeros(M, 1, device=dev))
            self._api_rows = M

    def _api_latent(self, privileged_obs, history, is_teacher, M):
        assert M <= self.teacher_engine.max_rows, "batch larger than the engines were sized for (flatten_ max_rows)"
        self._api_buffers(M)
        if is_teacher:
            self.teacher_latent(privileged_obs.contiguous(), M, self._api_lat[:M])
        else:
            self.student.forward(history.contiguous(), M, self._api_lat[:M])
        return self._api_lat

    def _api_heads(self, obs, privileged_obs, M, actor=True, critic=True):
        """mu (and sigma) / value of the [latent | obs] / [latent | privileged obs] rows; the plain-head families."""
        D = self.latent_dim
        if actor:
            obs = obs.contiguous()
            call("go2_concat2", ptr(self._api_lat), D, D, ptr(obs), self.num_obs, obs.shape[1], ptr(self._api_xa), self._api_xa.shape[1], 0, M)
            self.actor_engine.forward(self._api_xa, self._api_xa.shape[1], M, self._api_mu[:M], self.num_actions, x_ones=self._api_xa.shape[1] > D + self.num_obs)
            self._api_sigma[:M] = self.std.data
        if critic:
            priv = privileged_obs.contiguous()
            call("go2_concat2", ptr(self._api_lat), D, D, ptr(priv), self.num_critic_obs, priv.shape[1], ptr(self._api_xc), self._api_xc.shape[1], 0, M)
            self.critic_engine.forward(self._api_xc, self._api_xc.shape[1], M, self._api_val[:M], 1, x_ones=self._api_xc.shape[1] > D + self.num_critic_obs)

    def update_distribution(self, obs, privileged_obs, history, is_teacher):
        M = obs.shape[0]
        self._api_latent(privileged_obs, history, is_teacher, M)
        self._api_heads(obs, privileged_obs, M, critic=False)
        self.distribution = torch.distributions.Normal(self._api_mu[:M].clone(), self._api_sigma[:M].clone())

    def act(self, obs, privileged_obs, history, is_teacher, **kwargs):
        """actor_critic_cts.py:129-139: sample from the policy of the teacher (privileged obs) or student (history) branch."""
        self.update_distribution(obs, privileged_obs, history, is_teacher)
        return self.distribution.sample()

    def evaluate(self, privileged_obs, history, is_teacher, **kwargs):
        """actor_critic_cts.py:152-160: value of [latent (detached) | privileged obs]."""
        M = privileged_obs.shape[0]
        self._api_latent(privileged_obs, history, is_teacher, M)
        self._api_heads(None, privileged_obs, M, actor=False)
        return self._api_val[:M].clone()

    def get_actions_log_prob(self, actions):
        return self.distribution.log_prob(actions).sum(dim=-1)

    @property
    def action_mean(self):
        return self.distribution.mean

    @property
    def action_std(self):
        return self.distribution.stddev

    @property
    def entropy(self):
        return self.distribution.entropy().sum(dim=-1)

    def act_inference(self, obs):
        """Student policy (actor_critic_moe_cts.py:127-132): roll the history, encode it, act on [latent | obs]."""
        N = obs.shape[0]
        call("go2_history_update", ptr(self.history), ptr(obs.contiguous()), 0, N, self.history_length, self.num_obs)
        lat = self._inf_lat[:N]
        self.student.forward(self.history.view(N, -1), N, lat)
        xa = self._inf_xa[:N]
        call("go2_concat2", ptr(lat), self.latent_dim, self.latent_dim, ptr(obs), self.num_obs, obs.shape[1], ptr(xa), xa.shape[1], 0, N)
        out = self._inf_mu[:N]
        self.actor_engine.forward(xa, xa.shape[1], N, out, self.num_actions)
        return out.clone()


def combine_backward(mod, dpre, gates, eo, lb_coef, deo, dlogits, M, E, D):
    """go2_moe_combine_backward with the mean gate usage taken over ALL ranks' rows when the envs are sharded (mod.usage_sum: dist_utils.SmallSum)."""
    us = getattr(mod, "usage_sum", None)
    if us is None or lb_coef == 0.0:
        call("go2_moe_combine_backward", ptr(dpre), ptr(gates), ptr(eo), ptr(mod.usage), float(lb_coef), ptr(deo), 0, ptr(dlogits), 0, M, E, D)
        return
    call("go2_gate_usage", ptr(gates), ptr(us.inp), M, E, 1.0 / dist_utils.world_size())
    mod.usage.copy_(us.reduce())
    call("go2_moe_combine_backward_given_usage", ptr(dpre), ptr(gates), ptr(eo), ptr(mod.usage), float(lb_coef), ptr(deo), 0, ptr(dlogits), 0, M, E, D)


class _StudentMLP:
    """Plain student encoder: MLP + L2Norm (actor_critic_cts.py:73-89)."""

    def __init__(self, model, names, dims, max_rows, train_rows):
        self.m, self.D = model, dims[-1]
        self.engine = model._engine(names, dims, max_rows, train_rows)
        dev = model.device
        self.pre = torch.empty(max_rows, self.D, device=dev)
        self.norm = torch.empty(max_rows, device=dev)
        self.dpre = torch.empty(max(train_rows, 1), self.D, device=dev)

    def engines(self):
        return [self.engine]

    def mark_dirty(self):
        pass

    def forward(self, hist, M, out, train=False, x_ones=False):
        self.engine.forward(hist, hist.shape[1], M, self.pre[:M], self.D, train=train, x_ones=x_ones)
        call("go2_l2norm_forward", ptr(self.pre), self.D, ptr(out), out.stride(0), ptr(self.norm), M, self.D)
        self._out, self._M = out, M

    def backward(self, dlatent, lb_coef=0.0):
        M = self._M
        call("go2_l2norm_backward", ptr(dlatent), self.D, ptr(self._out), self._out.stride(0), ptr(self.norm), ptr(self.dpre), self.D, 0, M, self.D)
        self.engine.backward(self.dpre, self.D)
        return None


class _StudentMoE:
    """StudentMoEEncoder (modules/utils.py:69-151): shared backbone -> 8 block-diagonal experts, softmax gate, weighted sum, L2Norm."""

    def __init__(self, model, prefix, in_dim, hidden_dims, E, D, max_rows, train_rows, names=None, expert_cols=None):
        """names: parameter names {backbone: [...], gate: [...], experts: name} when they differ from the MoE-CTS layout (the no-goal variant);
        expert_cols: column indices of the history the experts' backbone sees (gate: the full history), actor_critic_moe_ng_cts.py:185-188."""
        self.m, self.E, self.D, self.H = model, E, D, hidden_dims[-1]
        dev = model.device
        self.expert_cols = None if expert_cols is None else expert_cols.to(dev)
        b_in = in_dim if expert_cols is None else int(expert_cols.numel())
        bdims = [b_in, *hidden_dims[:-1], E * self.H]
        gdims = [in_dim, *hidden_dims[:-1], E]
        nb = len(bdims) - 1
        if names is None:
            names = {"backbone": _linear_names(prefix + ".moe.experts.backbone.network", nb),
                     "gate": _linear_names(prefix + ".moe.gating_network.0.network", len(gdims) - 1), "experts": prefix + ".moe.experts.experts"}
        self.backbone = model._engine(names["backbone"], bdims, max_rows, train_rows, last_act=True)
        self.gate = model._engine(names["gate"], gdims, max_rows, train_rows)
        self._hsel = torch.zeros(max_rows, b_in, device=dev) if expert_cols is not None else None
        self.We = model._views[names["experts"] + ".weight"].view(E * D, self.H)     # Conv1d [E*D, H, 1] -> E blocks of [D, H]
        self.be = model._views[names["experts"] + ".bias"]
        self.gWe = model._gviews[names["experts"] + ".weight"].view(E * D, self.H)
        self.gbe = model._gviews[names["experts"] + ".bias"]
        tr = max(train_rows, 1)
        z = lambda *s: torch.empty(*s, device=dev)
        self.eo, self.logits, self.gates = z(max_rows, E * D), z(max_rows, E), z(max_rows, E)
        self.pre, self.norm = z(max_rows, D), z(max_rows)
        self.dpre, self.deo = z(tr, D), z(tr, E * D)
        self.dlogits = z(tr, E)
        self.dfeat = z(tr, E * self.H)
        self.usage = torch.zeros(E, device=dev)
        self.usage_sum = dist_utils.make_small_sum(E, dev) if train_rows else None      # env-sharded: the load-balance term sees the global mean usage
        self.Wet = torch.zeros(E * self.H, D, device=dev)   # per expert W_e^T [H, D], stacked
        self.pool = _ops.StreamPool(dev, 4)                 # the E experts are independent small GEMMs: four lanes, one split-K workspace each
        self.side = _ops.SideStream(dev)                    # the gate runs next to backbone + experts
        self.work = torch.empty(64 * 128 * (self.H + 4), device=dev)
        self.works = [self.work] + [torch.empty_like(self.work) for _ in range(self.pool.n - 1)]
        self._dirty = True
        self.train_rows = train_rows

    def engines(self):
        return [self.backbone, self.gate]

    def mark_dirty(self):
        self._dirty = True

    def forward(self, hist, M, out, train=False, x_ones=False):
        E, D, H = self.E, self.D, self.H
        self.side.fork()
        with self.side:
            self.gate.forward(hist, hist.shape[1], M, self.logits[:M], E, train=train, x_ones=x_ones)
        if self.expert_cols is None:
            self.backbone.forward(hist, hist.shape[1], M, train=train, x_ones=x_ones)
        else:   # the experts see the history without its command columns (gather into a dense buffer; the engine pads it itself)
            torch.index_select(hist[:M], 1, self.expert_cols, out=self._hsel[:M])
            self.backbone.forward(self._hsel, self._hsel.shape[1], M, train=train, x_ones=False)
        feat, ldf = self.backbone.out, self.backbone.ld_out
        self.pool.fork()
        for e in range(E):  # block-diagonal expert layer = Conv1d(groups=E, kernel 1)
            fn = "go2_linear_forward_tc" if _ops.use_tc() else "go2_linear_forward_simt"
            with self.pool.lane(e):
                call(fn, ptr(feat) + 4 * e * H, ldf, ptr(self.We) + 4 * e * D * H, H, ptr(self.be) + 4 * e * D, ptr(self.eo) + 4 * e * D, E * D, 0, 0,
                     M, D, H, 0)
        self.pool.join()
        self.side.join()
        call("go2_moe_combine_forward", ptr(self.logits), ptr(self.eo), ptr(self.gates), ptr(self.pre), M, E, D)
        call("go2_l2norm_forward", ptr(self.pre), D, ptr(out), out.stride(0), ptr(self.norm), M, D)
        self._out, self._M = out, M

    def backward(self, dlatent, lb_coef=0.0):
        E, D, H, M, tr = self.E, self.D, self.H, self._M, self.train_rows
        tc = _ops.use_tc()
        call("go2_l2norm_backward", ptr(dlatent), D, ptr(self._out), self._out.stride(0), ptr(self.norm), ptr(self.dpre), D, 0, M, D)
        combine_backward(self, self.dpre, self.gates, self.eo, lb_coef, self.deo, self.dlogits, M, E, D)
        if self._dirty and tc:      # W_e^T of the 8 experts in one launch
            import ctypes as C
            vp, ia = (C.c_void_p * E), (C.c_int * E)
            call("go2_refresh_weights", E, vp(*[ptr(self.We) + 4 * e * D * H for e in range(E)]), ia(*([H] * E)),
                 vp(*[ptr(self.Wet) + 4 * e * H * D for e in range(E)]), ia(*([D] * E)), ia(*([D] * E)), ia(*([H] * E)), ia(*([1] * E)))
            self._dirty = False
        call("go2_colsum", ptr(self.deo), E * D, ptr(self.gbe), M, E * D, ptr(self.work))
        feat, ldf = self.backbone.out, self.backbone.ld_out
        self.side.fork()
        with self.side:
            self.gate.backward(self.dlogits, E)
        self.pool.fork()
        for e in range(E):
            work = self.works[e % self.pool.n]
            with self.pool.lane(e):
                if tc:
                    call("go2_linear_wgrad_tc_rm", ptr(self.deo) + 4 * e * D, E * D, ptr(feat) + 4 * e * H, ldf, ptr(self.gWe) + 4 * e * D * H, H, 0,
                         M, D, H, ptr(work), work.numel())
                    call("go2_linear_dgrad_tc", ptr(self.deo) + 4 * e * D, E * D, ptr(self.Wet) + 4 * e * H * D, D, ptr(feat) + 4 * e * H, ldf, 0, 0,
                         ptr(self.dfeat) + 4 * e * H, E * H, 0, 0, M, D, H)
                else:
                    call("go2_linear_wgrad_simt", ptr(self.deo) + 4 * e * D, E * D, ptr(feat) + 4 * e * H, ldf, ptr(self.gWe) + 4 * e * D * H, H, 0,
                         M, D, H, ptr(work), work.numel())
                    call("go2_linear_dgrad_simt", ptr(self.deo) + 4 * e * D, E * D, ptr(self.We) + 4 * e * D * H, H, ptr(feat) + 4 * e * H, ldf,
                         ptr(self.dfeat) + 4 * e * H, E * H, 0, 0, M, D, H)
        self.pool.join()
        self.backbone.backward(self.dfeat, E * H)
        self.side.join()


class ActorCriticMoECTS(_CTSBase):
    def __init__(self, num_obs, num_critic_obs, num_actions, num_envs, history_length, actor_hidden_dims=[512, 256, 128],
                 critic_hidden_dims=[512, 256, 128], teacher_encoder_hidden_dims=[512, 256], student_encoder_hidden_dims=[512, 256, 256],
                 expert_num=8, activation='elu', init_noise_std=1.0, latent_dim=32, norm_type='l2norm', **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str([key for key in kwargs.keys()]))
        if activation != 'elu' or norm_type != 'l2norm':
            raise NotImplementedError("fused epilogues implement ELU / L2Norm (the go2_moe_cts configuration)")
        super().__init__()
        self.num_obs, self.num_critic_obs, self.num_actions = num_obs, num_critic_obs, num_actions
        self.history_length, self.latent_dim, self.expert_num = history_length, latent_dim, expert_num
        self.register_buffer("history", torch.zeros((num_envs, history_length, num_obs)), persistent=False)
        self.t_dims = [num_critic_obs, *teacher_encoder_hidden_dims, latent_dim]
        self.s_hidden = list(student_encoder_hidden_dims)
        self.a_dims = [latent_dim + num_obs, *actor_hidden_dims, num_actions]
        self.c_dims = [latent_dim + num_critic_obs, *critic_hidden_dims, 1]
        self.teacher_encoder = nn.Sequential(MLP(self.t_dims), L2Norm())
        self.student_moe_encoder = StudentMoEEncoder(expert_num, num_obs * history_length, self.s_hidden, latent_dim)
        self.actor = MLP(self.a_dims)
        self.critic = MLP(self.c_dims)
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        self.student_prefix = "student_moe_encoder"

    def _segments(self):
        names = [k for k, _ in self.named_parameters()]
        seg1 = [k for k in names if k.startswith("teacher_encoder.")] + [k for k in names if k.startswith("critic.")] + \
               [k for k in names if k.startswith("actor.")] + ["std"]
        seg2 = [k for k in names if k.startswith("student_moe_encoder.")]
        return seg1, seg2

    def _build_engines(self, dev, max_rows, tr1, trt, trs):
        self.teacher_engine = self._engine(_linear_names("teacher_encoder.0.network", len(self.t_dims) - 1), self.t_dims, max_rows, max(trt, trs))
        self.actor_engine = self._engine(_linear_names("actor.network", len(self.a_dims) - 1), self.a_dims, max_rows, tr1, need_dx=True)
        self.critic_engine = self._engine(_linear_names("critic.network", len(self.c_dims) - 1), self.c_dims, max_rows, tr1)
        self.student = _StudentMoE(self, "student_moe_encoder", self.num_obs * self.history_length, self.s_hidden, self.expert_num, self.latent_dim,
                                   max_rows, trs)
        self._common_buffers(dev, max_rows, max(trt, trs))

    def _common_buffers(self, dev, max_rows, trt):
        D = self.latent_dim
        self._t_pre = torch.empty(max_rows, D, device=dev)
        self._t_norm = torch.empty(max_rows, device=dev)
        self._t_dpre = torch.empty(max(trt, 1), D, device=dev)
        self._inf_lat = torch.empty(max_rows, D, device=dev)
        self._inf_xa = torch.zeros(max_rows, _ops.pad_in(self.a_dims[0]), device=dev)
        self._inf_mu = torch.empty(max_rows, self.num_actions, device=dev)


class ActorCriticCTS(_CTSBase):
    def __init__(self, num_actor_obs, num_critic_obs, num_actions, num_envs, history_length, actor_hidden_dims=[512, 256, 128],
                 critic_hidden_dims=[512, 256, 128], teacher_encoder_hidden_dims=[512, 256], student_encoder_hidden_dims=[512, 256],
                 activation='elu', init_noise_std=1.0, latent_dim=32, norm_type='l2norm', **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str([key for key in kwargs.keys()]))
        if activation != 'elu' or norm_type != 'l2norm':
            raise NotImplementedError("fused epilogues implement ELU / L2Norm (the go2_cts configuration)")
        super().__init__()
        self.num_obs, self.num_critic_obs, self.num_actions = num_actor_obs, num_critic_obs, num_actions
        self.history_length, self.latent_dim = history_length, latent_dim
        self.register_buffer("history", torch.zeros((num_envs, history_length, num_actor_obs)), persistent=False)
        self.t_dims = [num_critic_obs, *teacher_encoder_hidden_dims, latent_dim]
        self.s_dims = [num_actor_obs * history_length, *student_encoder_hidden_dims, latent_dim]
        self.a_dims = [latent_dim + num_actor_obs, *actor_hidden_dims, num_actions]
        self.c_dims = [latent_dim + num_critic_obs, *critic_hidden_dims, 1]
        self.teacher_encoder = nn.Sequential(*_seq_mlp(self.t_dims), L2Norm())
        self.student_encoder = nn.Sequential(*_seq_mlp(self.s_dims), L2Norm())
        self.actor = nn.Sequential(*_seq_mlp(self.a_dims))
        self.critic = nn.Sequential(*_seq_mlp(self.c_dims))
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))

    def _segments(self):
        names = [k for k, _ in self.named_parameters()]
        seg1 = [k for k in names if k.startswith("teacher_encoder.")] + [k for k in names if k.startswith("critic.")] + \
               [k for k in names if k.startswith("actor.")] + ["std"]
        seg2 = [k for k in names if k.startswith("student_encoder.")]
        return seg1, seg2

    def _build_engines(self, dev, max_rows, tr1, trt, trs):
        self.teacher_engine = self._engine(_linear_names("teacher_encoder", len(self.t_dims) - 1), self.t_dims, max_rows, max(trt, trs))
        self.actor_engine = self._engine(_linear_names("actor", len(self.a_dims) - 1), self.a_dims, max_rows, tr1, need_dx=True)
        self.critic_engine = self._engine(_linear_names("critic", len(self.c_dims) - 1), self.c_dims, max_rows, tr1)
        self.student = _StudentMLP(self, _linear_names("student_encoder", len(self.s_dims) - 1), self.s_dims, max_rows, trs)
        ActorCriticMoECTS._common_buffers(self, dev, max_rows, max(trt, trs))


class _NGStudentParams(nn.Module):
    """Parameter container with the reference's key layout (actor_critic_moe_ng_cts.py:187-230): experts_backbone.{0,2,..}, experts_hidden.0,
    experts_out (grouped 1x1 conv), gating_network.{0,2,..}."""

    def __init__(self, expert_dim, gating_dim, hidden_dims, expert_num, expert_hidden_dim, latent_dim):
        super().__init__()
        self.norm_layer = L2Norm()
        layers, last = [], expert_dim
        for h in hidden_dims:
            layers += [nn.Linear(last, h), nn.ELU()]
            last = h
        self.experts_backbone = nn.Sequential(*layers)
        self.experts_hidden = nn.Sequential(nn.Linear(last, expert_num * expert_hidden_dim), nn.ELU())
        self.experts_out = nn.Conv1d(expert_num * expert_hidden_dim, expert_num * latent_dim, kernel_size=1, groups=expert_num)
        layers, last = [], gating_dim
        for h in hidden_dims:
            layers += [nn.Linear(last, h), nn.ELU()]
            last = h
        layers += [nn.Linear(last, expert_num), nn.Softmax(dim=-1)]
        self.gating_network = nn.Sequential(*layers)


class ActorCriticMoENGCTS(ActorCriticMoECTS):
    """ActorCriticMoENGCTS (rsl_rl/modules/actor_critic_moe_ng_cts.py:18-188): MoE-CTS whose experts see the observation history WITHOUT the
    command ("goal") columns while the gate sees all of it.  Same kernels as ActorCriticMoECTS; the expert input is a column gather."""

    def __init__(self, num_obs, num_critic_obs, num_actions, num_envs, history_length, obs_no_goal_mask, actor_hidden_dims=[512, 256, 128],
                 critic_hidden_dims=[512, 256, 128], teacher_encoder_hidden_dims=[512, 256], student_encoder_hidden_dims=[512, 256],
                 student_expert_num=8, activation='elu', init_noise_std=1.0, latent_dim=32, norm_type='l2norm', expert_hidden_dim=256, **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str([key for key in kwargs.keys()]))
        if activation != 'elu' or norm_type != 'l2norm':
            raise NotImplementedError("fused epilogues implement ELU / L2Norm (the go2_moe_ng_cts configuration)")
        _CTSBase.__init__(self)
        self.num_obs, self.num_critic_obs, self.num_actions = num_obs, num_critic_obs, num_actions
        self.history_length, self.latent_dim, self.expert_num = history_length, latent_dim, student_expert_num
        self.register_buffer("obs_no_goal_mask", torch.tensor(obs_no_goal_mask, dtype=torch.bool), persistent=False)
        self.register_buffer("history", torch.zeros((num_envs, history_length, num_obs)), persistent=False)
        n_ng = int(self.obs_no_goal_mask.sum())
        self.t_dims = [num_critic_obs, *teacher_encoder_hidden_dims, latent_dim]
        self.s_hidden = [*student_encoder_hidden_dims, expert_hidden_dim]
        self.a_dims = [latent_dim + num_obs, *actor_hidden_dims, num_actions]
        self.c_dims = [latent_dim + num_critic_obs, *critic_hidden_dims, 1]
        self.teacher_encoder = nn.Sequential(*_seq_mlp(self.t_dims), L2Norm())
        self.student_moe_encoder = _NGStudentParams(n_ng * history_length, num_obs * history_length, student_encoder_hidden_dims, student_expert_num,
                                                    expert_hidden_dim, latent_dim)
        self.actor = nn.Sequential(*_seq_mlp(self.a_dims))
        self.critic = nn.Sequential(*_seq_mlp(self.c_dims))
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        self.student_prefix = "student_moe_encoder"
        # columns of the flattened [history_length x num_obs] history the experts see
        cols = torch.nonzero(self.obs_no_goal_mask).flatten()
        self._expert_cols = torch.cat([cols + t * num_obs for t in range(history_length)])

    def _build_engines(self, dev, max_rows, tr1, trt, trs):
        self.teacher_engine = self._engine(_linear_names("teacher_encoder", len(self.t_dims) - 1), self.t_dims, max_rows, max(trt, trs))
        self.actor_engine = self._engine(_linear_names("actor", len(self.a_dims) - 1), self.a_dims, max_rows, tr1, need_dx=True)
        self.critic_engine = self._engine(_linear_names("critic", len(self.c_dims) - 1), self.c_dims, max_rows, tr1)
        p, nh = "student_moe_encoder", len(self.s_hidden) - 1
        names = {"backbone": _linear_names(p + ".experts_backbone", nh) + [p + ".experts_hidden.0"],
                 "gate": _linear_names(p + ".gating_network", nh + 1), "experts": p + ".experts_out"}
        self.student = _StudentMoE(self, p, self.num_obs * self.history_length, self.s_hidden, self.expert_num, self.latent_dim, max_rows, trs,
                                   names=names, expert_cols=self._expert_cols)
        self._common_buffers(dev, max_rows, max(trt, trs))


# ---- MoE actor + gated critic experts (ac_moe_cts / dual_moe_cts) -----------------------------------------------------------------
class _ExpertLayer:
    """Conv1d(E*H -> E*D, kernel 1, groups=E) (modules/utils.py:83-88) = E block-diagonal Linear(H -> D) over the backbone's feature
    blocks.  Narrow experts (D <= 16, H <= 128: the 12-action and 1-value experts of the default configuration) run on the streaming
    fp32 kernels, wider ones on the GEMM path the student's experts use."""

    def __init__(self, model, pname, E, H, D, max_rows, train_rows):
        dev = model.device
        self.E, self.H, self.D = E, H, D
        self.W = model._views[pname + ".weight"].view(E * D, H)
        self.b = model._views[pname + ".bias"]
        self.gW = model._gviews[pname + ".weight"].view(E * D, H)
        self.gb = model._gviews[pname + ".bias"]
        self.small = D <= 16 and H <= 128
        # wider experts (the student encoder's 8 x (256 -> 32)): ONE grouped fp32 launch per direction for all experts (go2_grouped_linear_*);
        # GO2_EXPERTS=tc keeps the earlier one-tensor-core-GEMM-per-expert path for A / B
        self.grouped = not self.small and os.environ.get("GO2_EXPERTS", "grouped") == "grouped"
        self.tc = _ops.use_tc() and not self.small and not self.grouped and D % 4 == 0
        self.out = torch.empty(max_rows, E * D, device=dev)
        self.dfeat = torch.empty(max(train_rows, 1), E * H, device=dev)
        self.Wt = torch.zeros(E * H, D, device=dev) if self.tc else None
        self.work = torch.empty(max(296 * (D * H + D), 64 * 128 * (H + 4), 64 * E * D), device=dev)
        self.gwork = torch.empty(((max(train_rows, 1) + 255) // 256) * E * D * H, device=dev) if self.grouped else None   # go2_grouped_linear_wgrad_workspace
        self._dirty = True

    def mark_dirty(self):
        self._dirty = True

    def forward(self, feat, ldf, M):
        E, D, H = self.E, self.D, self.H
        if self.grouped:
            call("go2_grouped_linear_forward", ptr(feat), ldf, ptr(self.W), ptr(self.b), ptr(self.out), E * D, M, E, D, H)
            return
        for e in range(E):
            x, w, b, y = ptr(feat) + 4 * e * H, ptr(self.W) + 4 * e * D * H, ptr(self.b) + 4 * e * D, ptr(self.out) + 4 * e * D
            if self.small:
                call("go2_linear_forward_smalln", x, ldf, w, H, b, y, E * D, M, D, H)
            else:
                call("go2_linear_forward_tc" if self.tc else "go2_linear_forward_simt", x, ldf, w, H, b, y, E * D, 0, 0, M, D, H, 0)

    def backward(self, dout, feat, ldf, M):
        """dout [M, E*D] -> gW, gb and self.dfeat [M, E*H] = gradient w.r.t. the backbone's last LINEAR output (ELU' applied)."""
        E, D, H = self.E, self.D, self.H
        if self.tc and self._dirty:
            import ctypes as C
            vp, ia = (C.c_void_p * E), (C.c_int * E)
            call("go2_refresh_weights", E, vp(*[ptr(self.W) + 4 * e * D * H for e in range(E)]), ia(*([H] * E)),
                 vp(*[ptr(self.Wt) + 4 * e * H * D for e in range(E)]), ia(*([D] * E)), ia(*([D] * E)), ia(*([H] * E)), ia(*([1] * E)))
            self._dirty = False
        if not self.small:
            call("go2_colsum", ptr(dout), E * D, ptr(self.gb), M, E * D, ptr(self.work))
        if self.grouped:
            call("go2_grouped_linear_wgrad", ptr(dout), E * D, ptr(feat), ldf, ptr(self.gW), M, E, D, H, ptr(self.gwork), self.gwork.numel())
            call("go2_grouped_linear_dgrad", ptr(dout), E * D, ptr(self.W), ptr(feat), ldf, ptr(self.dfeat), E * H, M, E, D, H)
            return
        for e in range(E):
            d, x, gw = ptr(dout) + 4 * e * D, ptr(feat) + 4 * e * H, ptr(self.gW) + 4 * e * D * H
            df = ptr(self.dfeat) + 4 * e * H
            if self.small:
                call("go2_linear_wgrad_smalln", d, E * D, x, ldf, gw, H, ptr(self.gb) + 4 * e * D, M, D, H, ptr(self.work), self.work.numel())
                call("go2_linear_dgrad_smalln", d, E * D, ptr(self.W) + 4 * e * D * H, H, x, ldf, df, E * H, M, D, H)
            elif self.tc:
                call("go2_linear_wgrad_tc_rm", d, E * D, x, ldf, gw, H, 0, M, D, H, ptr(self.work), self.work.numel())
                call("go2_linear_dgrad_tc", d, E * D, ptr(self.Wt) + 4 * e * H * D, D, x, ldf, 0, 0, df, E * H, 0, 0, M, D, H)
            else:
                call("go2_linear_wgrad_simt", d, E * D, x, ldf, gw, H, 0, M, D, H, ptr(self.work), self.work.numel())
                call("go2_linear_dgrad_simt", d, E * D, ptr(self.W) + 4 * e * D * H, H, x, ldf, df, E * H, 0, 0, M, D, H)


class _MoEHead:
    """MoE (modules/utils.py:96-126) as a policy head: shared backbone -> E experts, softmax gate on the same input, weighted sum.
    Both the backbone and the gate return the gradient w.r.t. their input (the [latent | obs] row)."""

    def __init__(self, model, prefix, in_dim, hidden_dims, E, D, max_rows, train_rows):
        dev = model.device
        self.E, self.D, self.H = E, D, hidden_dims[-1]
        bdims = [in_dim, *hidden_dims[:-1], E * self.H]
        gdims = [in_dim, *hidden_dims[:-1], E]
        self.backbone = model._engine(_linear_names(prefix + ".experts.backbone.network", len(bdims) - 1), bdims, max_rows, train_rows,
                                      last_act=True, need_dx=True)
        self.gate = model._engine(_linear_names(prefix + ".gating_network.0.network", len(gdims) - 1), gdims, max_rows, train_rows, need_dx=True)
        self.layer = _ExpertLayer(model, prefix + ".experts.experts", E, self.H, D, max_rows, train_rows)
        tr = max(train_rows, 1)
        self.logits, self.gates = torch.empty(max_rows, E, device=dev), torch.empty(max_rows, E, device=dev)
        self.dlogits, self.deo = torch.empty(tr, E, device=dev), torch.empty(tr, E * D, device=dev)
        self.usage = torch.zeros(E, device=dev)
        self.usage_sum = dist_utils.make_small_sum(E, dev) if train_rows else None      # env-sharded: the load-balance term sees the global mean usage

    def engines(self):
        """Everything that keeps operand copies derived from the weights (refreshed after an optimiser step)."""
        return [self.backbone, self.gate, self.layer]

    def mark_dirty(self):
        self.layer.mark_dirty()

    def forward(self, x, ldx, M, out, train=False, x_ones=False):
        """out[M, D] (dense) = sum_e softmax(gate(x))_e * expert_e(backbone(x))"""
        self.backbone.forward(x, ldx, M, train=train, x_ones=x_ones)
        self.layer.forward(self.backbone.out, self.backbone.ld_out, M)
        self.gate.forward(x, ldx, M, self.logits[:M], self.E, train=train, x_ones=x_ones)
        call("go2_moe_combine_forward", ptr(self.logits), ptr(self.layer.out), ptr(self.gates), ptr(out), M, self.E, self.D)
        self._M = M

    def backward(self, dout, lb_coef=0.0, extra_dlogits=None):
        """dout [M, D] dense.  lb_coef: load-balance term on the mean gate usage of these M rows (ac_moe_cts.py:226-228).  extra_dlogits: the
        gate's gradient from another consumer of the same gates (the critic's weighted value), already through the softmax."""
        M, E, D = self._M, self.E, self.D
        combine_backward(self, dout, self.gates, self.layer.out, lb_coef, self.deo, self.dlogits, M, E, D)
        if extra_dlogits is not None:
            self.dlogits[:M].add_(extra_dlogits[:M])
        self.layer.backward(self.deo, self.backbone.out, self.backbone.ld_out, M)
        self.backbone.backward(self.layer.dfeat, E * self.H)
        self.gate.backward(self.dlogits, E)


class _GatedExpertsHead:
    """Experts (modules/utils.py:69-94) with one output each, weighted by gates computed elsewhere (the actor's gating network,
    actor_critic_ac_moe_cts.py:139-145): value = sum_e gates_e * v_e."""

    def __init__(self, model, prefix, in_dim, backbone_hidden_dims, H, E, max_rows, train_rows):
        dev = model.device
        self.E, self.H = E, H
        bdims = [in_dim, *backbone_hidden_dims, E * H]
        self.backbone = model._engine(_linear_names(prefix + ".backbone.network", len(bdims) - 1), bdims, max_rows, train_rows, last_act=True)
        self.layer = _ExpertLayer(model, prefix + ".experts", E, H, 1, max_rows, train_rows)
        tr = max(train_rows, 1)
        self.gates = torch.empty(max_rows, E, device=dev)
        self.deo, self.dlogits = torch.empty(tr, E, device=dev), torch.empty(tr, E, device=dev)
        self.usage = torch.zeros(E, device=dev)

    def engines(self):
        return [self.backbone, self.layer]

    def mark_dirty(self):
        self.layer.mark_dirty()

    def forward(self, x, ldx, M, logits, out, train=False, x_ones=False):
        self.backbone.forward(x, ldx, M, train=train, x_ones=x_ones)
        self.layer.forward(self.backbone.out, self.backbone.ld_out, M)
        call("go2_moe_combine_forward", ptr(logits), ptr(self.layer.out), ptr(self.gates), ptr(out), M, self.E, 1)
        self._M = M

    def backward(self, dvalue):
        """dvalue [M, 1] dense -> parameter gradients; self.dlogits [M, E] = the gate logits' gradient from the value."""
        M, E = self._M, self.E
        call("go2_moe_combine_backward", ptr(dvalue), ptr(self.gates), ptr(self.layer.out), ptr(self.usage), 0.0, ptr(self.deo), 0,
             ptr(self.dlogits), 0, M, E, 1)
        self.layer.backward(self.deo, self.backbone.out, self.backbone.ld_out, M)
        self.backbone.backward(self.layer.dfeat, E * self.H)


class ActorCriticACMoECTS(_CTSBase):
    """ActorCriticACMoECTS (rsl_rl/modules/actor_critic_ac_moe_cts.py:21-146): CTS whose actor is a mixture of experts on [latent | obs] and
    whose critic is a set of value experts on [latent | privileged obs] weighted by the ACTOR's gate.  Same constructor arguments and
    state_dict keys (teacher_encoder.0.network.*, student_encoder.0.network.*, actor_moe.*, critic_experts.*, std)."""
    moe_heads = True

    def __init__(self, num_obs, num_critic_obs, num_actions, num_envs, history_length, actor_hidden_dims=[512, 256, 128],
                 critic_hidden_dims=[512, 256, 128], teacher_encoder_hidden_dims=[512, 256], student_encoder_hidden_dims=[512, 256],
                 expert_num=8, activation='elu', init_noise_std=1.0, latent_dim=32, norm_type='l2norm', **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str([key for key in kwargs.keys()]))
        if activation != 'elu' or norm_type != 'l2norm':
            raise NotImplementedError("fused epilogues implement ELU / L2Norm (the go2_ac_moe_cts / go2_dual_moe_cts configuration)")
        super().__init__()
        self.num_obs, self.num_critic_obs, self.num_actions = num_obs, num_critic_obs, num_actions
        self.history_length, self.latent_dim, self.expert_num = history_length, latent_dim, expert_num
        self.register_buffer("history", torch.zeros((num_envs, history_length, num_obs)), persistent=False)
        self.t_dims = [num_critic_obs, *teacher_encoder_hidden_dims, latent_dim]
        self.a_hidden, self.c_hidden = list(actor_hidden_dims), list(critic_hidden_dims)
        self.a_dims = [latent_dim + num_obs, *actor_hidden_dims, num_actions]       # input / output widths (the hidden part is the MoE's)
        self.c_dims = [latent_dim + num_critic_obs, *critic_hidden_dims, 1]
        self.teacher_encoder = nn.Sequential(MLP(self.t_dims), L2Norm())
        self._make_student(num_obs * history_length, list(student_encoder_hidden_dims), latent_dim, expert_num)
        self.actor_moe = MoE(expert_num, self.a_dims[0], self.a_hidden, num_actions)
        self.critic_experts = Experts(expert_num, self.c_dims[0], self.c_hidden[:-1], self.c_hidden[-1], 1)
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))

    student_prefix = "student_encoder"

    def _make_student(self, in_dim, hidden, latent_dim, expert_num):
        self.s_dims = [in_dim, *hidden, latent_dim]
        self.student_encoder = nn.Sequential(MLP(self.s_dims), L2Norm())

    def _segments(self):
        names = [k for k, _ in self.named_parameters()]
        seg1 = [k for k in names if k.startswith("teacher_encoder.")] + [k for k in names if k.startswith("critic_experts.")] + \
               [k for k in names if k.startswith("actor_moe.")] + ["std"]
        seg2 = [k for k in names if k.startswith(self.student_prefix + ".")]
        return seg1, seg2

    def _build_student(self, max_rows, trs):
        return _StudentMLP(self, _linear_names("student_encoder.0.network", len(self.s_dims) - 1), self.s_dims, max_rows, trs)

    def _build_engines(self, dev, max_rows, tr1, trt, trs):
        E = self.expert_num
        self.teacher_engine = self._engine(_linear_names("teacher_encoder.0.network", len(self.t_dims) - 1), self.t_dims, max_rows, max(trt, trs))
        self.actor_head = _MoEHead(self, "actor_moe", self.a_dims[0], self.a_hidden, E, self.num_actions, max_rows, tr1)
        self.critic_head = _GatedExpertsHead(self, "critic_experts", self.c_dims[0], self.c_hidden[:-1], self.c_hidden[-1], E, max_rows, tr1)
        self.student = self._build_student(max_rows, trs)
        ActorCriticMoECTS._common_buffers(self, dev, max_rows, max(trt, trs))
        self._dx = torch.zeros(max(tr1, 1), self.actor_head.backbone.kpad0, device=dev)

    def pass1_engines(self):
        return [self.teacher_engine] + self.actor_head.engines() + self.critic_head.engines()

    def engines(self):
        return self.pass1_engines() + self.student.engines()

    def mark_dirty(self):
        super().mark_dirty()
        self.actor_head.mark_dirty()
        self.critic_head.mark_dirty()

    def heads_forward(self, xa, xc, M, mu, val, train=False):
        """mu[M, A], val[M, 1] from the padded [latent | obs] / [latent | privileged obs] rows (actor_critic_ac_moe_cts.py:103-145)."""
        a, c = self.actor_head, self.critic_head
        a.forward(xa, xa.shape[1], M, mu, train=train, x_ones=xa.shape[1] > self.a_dims[0])
        c.forward(xc, xc.shape[1], M, a.logits, val, train=train, x_ones=xc.shape[1] > self.c_dims[0])

    def heads_backward(self, dmu, dval, M, lb_coef):
        """-> [M, kpad0] whose first latent_dim columns are d loss / d latent through the actor's backbone AND gate (the critic's input
        latent is detached, actor_critic_ac_moe_cts.py:142; its gate gradient arrives through the shared gates)."""
        a, c = self.actor_head, self.critic_head
        c.backward(dval)
        a.backward(dmu, lb_coef, extra_dlogits=c.dlogits)
        torch.add(a.backbone.dx[:M], a.gate.dx[:M], out=self._dx[:M])
        return self._dx

    def _api_heads(self, obs, privileged_obs, M, actor=True, critic=True):
        D = self.latent_dim
        obs, priv = obs.contiguous(), privileged_obs.contiguous()
        call("go2_concat2", ptr(self._api_lat), D, D, ptr(obs), self.num_obs, obs.shape[1], ptr(self._api_xa), self._api_xa.shape[1], 0, M)
        call("go2_concat2", ptr(self._api_lat), D, D, ptr(priv), self.num_critic_obs, priv.shape[1], ptr(self._api_xc), self._api_xc.shape[1], 0, M)
        self.heads_forward(self._api_xa, self._api_xc, M, self._api_mu, self._api_val)        # the value needs the actor's gate
        self._api_sigma[:M] = self.std.data

    def evaluate(self, obs, privileged_obs, history, is_teacher, **kwargs):
        """actor_critic_ac_moe_cts.py:134-146: (value, gate weights); the value experts are weighted by the ACTOR's gate, hence obs."""
        M = obs.shape[0]
        self._api_latent(privileged_obs, history, is_teacher, M)
        self._api_heads(obs, privileged_obs, M)
        return self._api_val[:M].clone(), self.actor_head.gates[:M].clone()

    def act_inference(self, obs):
        """Student policy (actor_critic_ac_moe_cts.py:127-132)."""
        N = obs.shape[0]
        call("go2_history_update", ptr(self.history), ptr(obs.contiguous()), 0, N, self.history_length, self.num_obs)
        lat = self._inf_lat[:N]
        self.student.forward(self.history.view(N, -1), N, lat)
        xa = self._inf_xa[:N]
        call("go2_concat2", ptr(lat), self.latent_dim, self.latent_dim, ptr(obs), self.num_obs, obs.shape[1], ptr(xa), xa.shape[1], 0, N)
        out = self._inf_mu[:N]
        self.actor_head.forward(xa, xa.shape[1], N, out, x_ones=xa.shape[1] > self.a_dims[0])
        return out.clone()


class ActorCriticDualMoECTS(ActorCriticACMoECTS):
    """ActorCriticDualMoECTS (rsl_rl/modules/actor_critic_dual_moe_cts.py:21-149): ActorCriticACMoECTS with the MoE student encoder of
    ActorCriticMoECTS (student_moe_encoder.moe.*)."""
    student_prefix = "student_moe_encoder"

    def __init__(self, num_obs, num_critic_obs, num_actions, num_envs, history_length, student_encoder_hidden_dims=[512, 256, 256], **kwargs):
        super().__init__(num_obs, num_critic_obs, num_actions, num_envs, history_length, student_encoder_hidden_dims=student_encoder_hidden_dims, **kwargs)

    def _make_student(self, in_dim, hidden, latent_dim, expert_num):
        self.s_hidden = hidden
        self.student_moe_encoder = StudentMoEEncoder(expert_num, in_dim, hidden, latent_dim)

    def _build_student(self, max_rows, trs):
        return _StudentMoE(self, "student_moe_encoder", self.num_obs * self.history_length, self.s_hidden, self.expert_num, self.latent_dim, max_rows, trs)


# ---- multiplicative compositional policy (mcp_cts) ---------------------------------------------------------------------------------
class _ActorMCPParams(nn.Module):
    """Parameter container with the reference's key layout (actor_critic_mcp_cts.py:183-218): gating_network.{0,2,..}, experts_backbone.{0,2,..},
    experts_hidden.0, experts_out (grouped 1x1 conv with 2 * action_dim outputs per expert: mean and log-std)."""

    def __init__(self, input_dim, input_dim_no_goal, action_dim, hidden_dims, expert_num, expert_hidden_dim):
        super().__init__()
        layers, last = [], input_dim
        for h in hidden_dims:
            layers += [nn.Linear(last, h), nn.ELU()]
            last = h
        layers += [nn.Linear(last, expert_num), nn.Sigmoid()]
        self.gating_network = nn.Sequential(*layers)
        layers, last = [], input_dim_no_goal
        for h in hidden_dims:
            layers += [nn.Linear(last, h), nn.ELU()]
            last = h
        self.experts_backbone = nn.Sequential(*layers)
        self.experts_hidden = nn.Sequential(nn.Linear(last, expert_num * expert_hidden_dim), nn.ELU())
        self.experts_out = nn.Conv1d(expert_num * expert_hidden_dim, expert_num * action_dim * 2, kernel_size=1, groups=expert_num)


class ActorCriticMCPCTS(_CTSBase):
    """ActorCriticMCPCTS (rsl_rl/modules/actor_critic_mcp_cts.py:19-180): CTS whose actor composes E expert Gaussians multiplicatively
    (MCP, arXiv 1905.09808): a sigmoid gate on [latent | obs], experts on [latent | obs without the command columns], action distribution
    N(mu, sigma) with a STATE-DEPENDENT sigma (no `std` parameter).  Same constructor arguments and state_dict keys."""
    mcp_head = True
    EXPERT_HIDDEN = 256          # ActorMCP's expert_hidden_dim default; ActorCriticMCPCTS never overrides it (:94-101)

    def __init__(self, num_obs, num_critic_obs, num_actions, num_envs, history_length, obs_no_goal_mask, actor_hidden_dims=[512, 256],
                 critic_hidden_dims=[512, 256, 128], teacher_encoder_hidden_dims=[512, 256], student_encoder_hidden_dims=[512, 256],
                 student_expert_num=8, activation='elu', latent_dim=32, norm_type='l2norm', **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str([key for key in kwargs.keys()]))
        if activation != 'elu' or norm_type != 'l2norm':
            raise NotImplementedError("fused epilogues implement ELU / L2Norm (the go2_mcp_cts configuration)")
        super().__init__()
        self.num_obs, self.num_critic_obs, self.num_actions = num_obs, num_critic_obs, num_actions
        self.history_length, self.latent_dim, self.expert_num = history_length, latent_dim, student_expert_num
        self.register_buffer("obs_no_goal_mask", torch.tensor(obs_no_goal_mask, dtype=torch.bool), persistent=False)
        self.num_obs_no_goal = int(self.obs_no_goal_mask.sum())
        self.register_buffer("history", torch.zeros((num_envs, history_length, num_obs)), persistent=False)
        self.t_dims = [num_critic_obs, *teacher_encoder_hidden_dims, latent_dim]
        self.s_dims = [num_obs * history_length, *student_encoder_hidden_dims, latent_dim]
        self.a_hidden = list(actor_hidden_dims)
        self.a_dims = [latent_dim + num_obs, *actor_hidden_dims, num_actions]
        self.ng_dim = latent_dim + self.num_obs_no_goal
        self.c_dims = [latent_dim + num_critic_obs, *critic_hidden_dims, 1]
        self.teacher_encoder = nn.Sequential(*_seq_mlp(self.t_dims), L2Norm())
        self.student_encoder = nn.Sequential(*_seq_mlp(self.s_dims), L2Norm())
        self.actor_mcp = _ActorMCPParams(self.a_dims[0], self.ng_dim, num_actions, self.a_hidden, student_expert_num, self.EXPERT_HIDDEN)
        self.critic = nn.Sequential(*_seq_mlp(self.c_dims))
        self._ng_cols = torch.nonzero(self.obs_no_goal_mask).flatten()

    def _segments(self):
        names = [k for k, _ in self.named_parameters()]
        seg1 = [k for k in names if k.startswith("teacher_encoder.")] + [k for k in names if k.startswith("critic.")] + \
               [k for k in names if k.startswith("actor_mcp.")]
        seg2 = [k for k in names if k.startswith("student_encoder.")]
        return seg1, seg2

    def _build_engines(self, dev, max_rows, tr1, trt, trs):
        E, A, H, nh = self.expert_num, self.num_actions, self.EXPERT_HIDDEN, len(self.a_hidden)
        self.teacher_engine = self._engine(_linear_names("teacher_encoder", len(self.t_dims) - 1), self.t_dims, max_rows, max(trt, trs))
        self.critic_engine = self._engine(_linear_names("critic", len(self.c_dims) - 1), self.c_dims, max_rows, tr1)
        self.gate_engine = self._engine(_linear_names("actor_mcp.gating_network", nh + 1), [self.a_dims[0], *self.a_hidden, E], max_rows, tr1, need_dx=True)
        self.backbone_engine = self._engine(_linear_names("actor_mcp.experts_backbone", nh) + ["actor_mcp.experts_hidden.0"],
                                            [self.ng_dim, *self.a_hidden, E * H], max_rows, tr1, last_act=True, need_dx=True)
        self.layer = _ExpertLayer(self, "actor_mcp.experts_out", E, H, 2 * A, max_rows, tr1)
        self.student = _StudentMLP(self, _linear_names("student_encoder", len(self.s_dims) - 1), self.s_dims, max_rows, trs)
        ActorCriticMoECTS._common_buffers(self, dev, max_rows, max(trt, trs))
        tr = max(tr1, 1)
        z = lambda *s: torch.zeros(*s, device=dev)
        self.logits, self.gates = z(max_rows, E), z(max_rows, E)
        self.dlogits, self.deo, self._dlat = z(tr, E), z(tr, E * 2 * A), z(tr, self.latent_dim)
        self._ng_cols = self._ng_cols.to(dev)
        self._inf_ng = z(max_rows, self.num_obs_no_goal)
        self._inf_xng = z(max_rows, _ops.pad_in(self.ng_dim))
        self._inf_sigma = z(max_rows, A)

    def pass1_engines(self):
        return [self.teacher_engine, self.critic_engine, self.gate_engine, self.backbone_engine, self.layer]

    def mark_dirty(self):
        super().mark_dirty()
        self.layer.mark_dirty()

    def heads_forward(self, xa, xng, xc, M, mu, sigma, val, train=False):
        """mu, sigma [M, A] and val [M, 1] from the padded [latent | obs], [latent | obs without commands], [latent | privileged obs] rows."""
        E, A = self.expert_num, self.num_actions
        self.gate_engine.forward(xa, xa.shape[1], M, self.logits[:M], E, train=train, x_ones=xa.shape[1] > self.a_dims[0])
        self.backbone_engine.forward(xng, xng.shape[1], M, train=train, x_ones=xng.shape[1] > self.ng_dim)
        self.layer.forward(self.backbone_engine.out, self.backbone_engine.ld_out, M)
        call("go2_mcp_compose_forward", ptr(self.layer.out), ptr(self.logits), ptr(self.gates), ptr(mu), ptr(sigma), M, E, A)
        if val is not None:
            self.critic_engine.forward(xc, xc.shape[1], M, val[:M], 1, train=train, x_ones=xc.shape[1] > self.c_dims[0])

    def heads_backward(self, dmu, dsigma, dval, M):
        """-> [M, latent_dim]: d loss / d latent through the gate AND the experts' backbone (the critic's input latent is detached, :178)."""
        E, A, D = self.expert_num, self.num_actions, self.latent_dim
        call("go2_mcp_compose_backward", ptr(dmu), ptr(dsigma), ptr(self.layer.out), ptr(self.gates), ptr(self.deo), ptr(self.dlogits), M, E, A)
        self.layer.backward(self.deo, self.backbone_engine.out, self.backbone_engine.ld_out, M)
        self.backbone_engine.backward(self.layer.dfeat, E * self.EXPERT_HIDDEN)
        self.gate_engine.backward(self.dlogits, E)
        self.critic_engine.backward(dval, 1)
        torch.add(self.gate_engine.dx[:M, :D], self.backbone_engine.dx[:M, :D], out=self._dlat[:M])
        return self._dlat

    def _api_heads(self, obs, privileged_obs, M, actor=True, critic=True):
        D = self.latent_dim
        if actor:
            obs = obs.contiguous()
            if getattr(self, "_api_ng_rows", 0) < M:
                self._api_ng = torch.zeros(M, self.num_obs_no_goal, device=self.device)
                self._api_xng = torch.zeros(M, _ops.pad_in(self.ng_dim), device=self.device)
                self._api_ng_rows = M
            ng = self.no_goal(obs, M, self._api_ng)
            call("go2_concat2", ptr(self._api_lat), D, D, ptr(obs), self.num_obs, obs.shape[1], ptr(self._api_xa), self._api_xa.shape[1], 0, M)
            call("go2_concat2", ptr(self._api_lat), D, D, ptr(ng), self.num_obs_no_goal, ng.shape[1], ptr(self._api_xng), self._api_xng.shape[1], 0, M)
            self.heads_forward(self._api_xa, self._api_xng, None, M, self._api_mu, self._api_sigma, None)
        if critic:
            priv = privileged_obs.contiguous()
            call("go2_concat2", ptr(self._api_lat), D, D, ptr(priv), self.num_critic_obs, priv.shape[1], ptr(self._api_xc), self._api_xc.shape[1], 0, M)
            self.critic_engine.forward(self._api_xc, self._api_xc.shape[1], M, self._api_val[:M], 1, x_ones=self._api_xc.shape[1] > D + self.num_critic_obs)

    def no_goal(self, obs, M, out):
        """out[M, n_ng] = obs[:, obs_no_goal_mask]  (actor_critic_mcp_cts.py:156)"""
        torch.index_select(obs[:M], 1, self._ng_cols, out=out[:M])
        return out

    def act_inference(self, obs):
        """Student policy, mean action (actor_critic_mcp_cts.py:166-173)."""
        N, D = obs.shape[0], self.latent_dim
        call("go2_history_update", ptr(self.history), ptr(obs.contiguous()), 0, N, self.history_length, self.num_obs)
        lat = self._inf_lat[:N]
        self.student.forward(self.history.view(N, -1), N, lat)
        xa, xng, ng = self._inf_xa[:N], self._inf_xng[:N], self.no_goal(obs, N, self._inf_ng)
        call("go2_concat2", ptr(lat), D, D, ptr(obs), self.num_obs, obs.shape[1], ptr(xa), xa.shape[1], 0, N)
        call("go2_concat2", ptr(lat), D, D, ptr(ng), self.num_obs_no_goal, ng.shape[1], ptr(xng), xng.shape[1], 0, N)
        out = self._inf_mu[:N]
        self.heads_forward(xa, xng, None, N, out, self._inf_sigma, None)
        return out.clone()

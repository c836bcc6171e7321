"""Policy export (drop-in for legged_gym/utils/exporter.py:13-58): TorchScript `policy.pt` / state-dict `policy.pkl` of a trained
actor, with the call signatures the deployment loops expect (deploy/deploy_mujoco/deploy_go2.py:235-240):

    ActorCritic        forward(obs[1,45]) -> action[1,12]                                  (exporter.py:127-128)
    ActorCriticCTS     forward(obs)       -> (action, (None, latent[1,32]))                (exporter.py:130-135)
    ActorCriticMoECTS  forward(obs)       -> (action, (gate weights[1,8], latent[1,32]))   (exporter.py:145-150)
    ActorCriticMoENGCTS forward(obs)      -> (action, (gate weights[1,8], latent[1,32]))   (exporter.py:137-143; experts on the no-goal history)

The CTS variants keep the rolling observation history ([1, H, 45], shift-append) inside the module and expose `reset()`.
The exported module is plain PyTorch built from the policy's state_dict (inference on the robot / in MuJoCo has no B200);
it is rebuilt from weights rather than deep-copied because this package's modules evaluate through the CUDA library.
The three remaining ablation variants (MCP / AC-MoE / Dual-MoE; all seven registered tasks train on the CUDA kernels) are exported from any
module or checkpoint with the reference's key layout; ONNX export and the recurrent policy are out of scope."""
import os
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def _mlp_from(sd, prefix, last_activation=False):
    """nn.Sequential of Linear / ELU from the `prefix.{0,2,4,..}.{weight,bias}` entries of a state dict."""
    idx = sorted({int(k[len(prefix) + 1:].split(".")[0]) for k in sd if k.startswith(prefix + ".") and k.endswith(".weight")})
    if not idx:
        raise ValueError(f"no '{prefix}.*' layers in the policy state dict")
    layers = []
    for n, i in enumerate(idx):
        w, b = sd[f"{prefix}.{i}.weight"], sd[f"{prefix}.{i}.bias"]
        lin = nn.Linear(w.shape[1], w.shape[0])
        lin.weight.data.copy_(w.detach().cpu())
        lin.bias.data.copy_(b.detach().cpu())
        layers.append(lin)
        if n < len(idx) - 1 or last_activation:
            layers.append(nn.ELU())
    return nn.Sequential(*layers)


class _ActorPolicy(nn.Module):
    def __init__(self, sd, normalizer=None):
        super().__init__()
        self.actor = _mlp_from(sd, "actor")
        self.normalizer = normalizer if normalizer is not None else nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.actor(self.normalizer(x))

    @torch.jit.export
    def reset(self):
        pass


class _CTSPolicy(nn.Module):
    def __init__(self, sd, history_length, num_obs, normalizer=None):
        super().__init__()
        self.student_encoder = _mlp_from(sd, "student_encoder")
        self.actor = _mlp_from(sd, "actor.network" if any(k.startswith("actor.network.") for k in sd) else "actor")
        self.normalizer = normalizer if normalizer is not None else nn.Identity()
        self.history = torch.zeros(1, history_length, num_obs)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, Tuple[Optional[torch.Tensor], torch.Tensor]]:
        x = self.normalizer(x)
        self.history = torch.cat([self.history[:, 1:], x.unsqueeze(1)], dim=1)
        latent = F.normalize(self.student_encoder(self.history.flatten(1)), p=2.0, dim=-1)
        none: Optional[torch.Tensor] = None
        return self.actor(torch.cat([latent, x], dim=1)), (none, latent)

    @torch.jit.export
    def reset(self):
        self.history = torch.zeros_like(self.history)


class _MoECTSPolicy(nn.Module):
    def __init__(self, sd, history_length, num_obs, expert_num, normalizer=None):
        super().__init__()
        p = "student_moe_encoder.moe."
        self.backbone = _mlp_from(sd, p + "experts.backbone.network", last_activation=True)
        w = sd[p + "experts.experts.weight"].detach().cpu()          # Conv1d(E*H -> E*D, k=1, groups=E): [E*D, H, 1]
        self.expert_num = int(expert_num)
        self.out_dim = w.shape[0] // self.expert_num
        self.register_buffer("expert_w", w.reshape(self.expert_num, self.out_dim, w.shape[1]).clone())      # [E, D, H]
        self.register_buffer("expert_b", sd[p + "experts.experts.bias"].detach().cpu().reshape(self.expert_num, self.out_dim).clone())
        self.gate = _mlp_from(sd, p + "gating_network.0.network")
        self.actor = _mlp_from(sd, "actor.network")
        self.normalizer = normalizer if normalizer is not None else nn.Identity()
        self.history = torch.zeros(1, history_length, num_obs)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
        x = self.normalizer(x)
        self.history = torch.cat([self.history[:, 1:], x.unsqueeze(1)], dim=1)
        h = self.history.flatten(1)
        weights = torch.softmax(self.gate(h), dim=-1)                                   # [B, E]
        feat = self.backbone(h).reshape(-1, self.expert_num, self.expert_w.shape[2])      # [B, E, H]
        outs = torch.einsum("beh,edh->bed", feat, self.expert_w) + self.expert_b          # grouped 1x1 convolution
        latent = F.normalize(torch.sum(weights.unsqueeze(-1) * outs, dim=1), p=2.0, dim=-1)
        return self.actor(torch.cat([latent, x], dim=1)), (weights, latent)

    @torch.jit.export
    def reset(self):
        self.history = torch.zeros_like(self.history)


class _MoENGCTSPolicy(nn.Module):
    """MoE student whose experts see the history without its command columns (exporter.py:137-143, actor_critic_moe_ng_cts.py:185-230)."""

    def __init__(self, sd, history_length, num_obs, obs_no_goal_mask, normalizer=None):
        super().__init__()
        p = "student_moe_encoder."
        self.backbone = _mlp_from(sd, p + "experts_backbone", last_activation=True)
        self.hidden = _mlp_from(sd, p + "experts_hidden", last_activation=True)
        self.gate = _mlp_from(sd, p + "gating_network")
        self.expert_num = int(self.gate[len(self.gate) - 1].out_features)
        w = sd[p + "experts_out.weight"].detach().cpu()              # Conv1d(E*H -> E*D, k=1, groups=E): [E*D, H, 1]
        self.out_dim = w.shape[0] // self.expert_num
        self.register_buffer("expert_w", w.reshape(self.expert_num, self.out_dim, w.shape[1]).clone())
        self.register_buffer("expert_b", sd[p + "experts_out.bias"].detach().cpu().reshape(self.expert_num, self.out_dim).clone())
        self.actor = _mlp_from(sd, "actor")
        self.normalizer = normalizer if normalizer is not None else nn.Identity()
        self.history_length = int(history_length)
        self.register_buffer("obs_no_goal_mask", torch.as_tensor(obs_no_goal_mask, dtype=torch.bool).cpu().clone())
        self.history = torch.zeros(1, history_length, num_obs)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
        x = self.normalizer(x)
        self.history = torch.cat([self.history[:, 1:], x.unsqueeze(1)], dim=1)
        no_goal = self.history.reshape(1, self.history_length, -1)[:, :, self.obs_no_goal_mask].reshape(1, -1)
        weights = torch.softmax(self.gate(self.history.flatten(1)), dim=-1)
        feat = self.hidden(self.backbone(no_goal)).reshape(-1, self.expert_num, self.expert_w.shape[2])
        outs = torch.einsum("beh,edh->bed", feat, self.expert_w) + self.expert_b
        latent = F.normalize(torch.sum(weights.unsqueeze(-1) * outs, dim=1), p=2.0, dim=-1)
        return self.actor(torch.cat([latent, x], dim=1)), (weights, latent)

    @torch.jit.export
    def reset(self):
        self.history = torch.zeros_like(self.history)


class _MoEBlock(nn.Module):
    """modules/utils.py:96-126 (`MoE`): shared backbone -> E block-diagonal experts (grouped 1x1 conv), softmax gate, weighted sum."""

    def __init__(self, sd, prefix):
        super().__init__()
        self.backbone = _mlp_from(sd, prefix + ".experts.backbone.network", last_activation=True)
        self.gate = _mlp_from(sd, prefix + ".gating_network.0.network")
        self.expert_num = int(self.gate[len(self.gate) - 1].out_features)
        w = sd[prefix + ".experts.experts.weight"].detach().cpu()
        self.out_dim = w.shape[0] // self.expert_num
        self.register_buffer("expert_w", w.reshape(self.expert_num, self.out_dim, w.shape[1]).clone())
        self.register_buffer("expert_b", sd[prefix + ".experts.experts.bias"].detach().cpu().reshape(self.expert_num, self.out_dim).clone())

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        weights = torch.softmax(self.gate(x), dim=-1)
        feat = self.backbone(x).reshape(-1, self.expert_num, self.expert_w.shape[2])
        outs = torch.einsum("beh,edh->bed", feat, self.expert_w) + self.expert_b
        return torch.sum(weights.unsqueeze(-1) * outs, dim=1), weights


class _ACMoECTSPolicy(nn.Module):
    """MLP student encoder + MoE actor (exporter.py:162-168, actor_critic_ac_moe_cts.py:127-133)."""

    def __init__(self, sd, history_length, num_obs, normalizer=None):
        super().__init__()
        self.student_encoder = _mlp_from(sd, "student_encoder.0.network")
        self.actor = _MoEBlock(sd, "actor_moe")
        self.normalizer = normalizer if normalizer is not None else nn.Identity()
        self.history = torch.zeros(1, history_length, num_obs)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
        x = self.normalizer(x)
        self.history = torch.cat([self.history[:, 1:], x.unsqueeze(1)], dim=1)
        latent = F.normalize(self.student_encoder(self.history.flatten(1)), p=2.0, dim=-1)
        mean, weights = self.actor(torch.cat([latent, x], dim=1))
        return mean, (weights, latent)

    @torch.jit.export
    def reset(self):
        self.history = torch.zeros_like(self.history)


class _DualMoECTSPolicy(nn.Module):
    """MoE student encoder + MoE actor (exporter.py:170-176, actor_critic_dual_moe_cts.py)."""

    def __init__(self, sd, history_length, num_obs, normalizer=None):
        super().__init__()
        self.student = _MoEBlock(sd, "student_moe_encoder.moe")
        self.actor = _MoEBlock(sd, "actor_moe")
        self.normalizer = normalizer if normalizer is not None else nn.Identity()
        self.history = torch.zeros(1, history_length, num_obs)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        x = self.normalizer(x)
        self.history = torch.cat([self.history[:, 1:], x.unsqueeze(1)], dim=1)
        lat, student_weights = self.student(self.history.flatten(1))
        latent = F.normalize(lat, p=2.0, dim=-1)
        mean, actor_weights = self.actor(torch.cat([latent, x], dim=1))
        return mean, (student_weights, actor_weights, latent)

    @torch.jit.export
    def reset(self):
        self.history = torch.zeros_like(self.history)


class _MCPCTSPolicy(nn.Module):
    """MLP student encoder + multiplicative-composition actor (exporter.py:152-160, actor_critic_mcp_cts.py:172-247): E Gaussian primitives
    on [latent | obs without commands], sigmoid gate on [latent | obs]; the composite mean is the precision-weighted mean of the primitives."""

    def __init__(self, sd, history_length, num_obs, obs_no_goal_mask, normalizer=None):
        super().__init__()
        self.student_encoder = _mlp_from(sd, "student_encoder")
        p = "actor_mcp."
        self.gate = _mlp_from(sd, p + "gating_network")
        self.backbone = _mlp_from(sd, p + "experts_backbone", last_activation=True)
        self.hidden = _mlp_from(sd, p + "experts_hidden", last_activation=True)
        self.expert_num = int(self.gate[len(self.gate) - 1].out_features)
        w = sd[p + "experts_out.weight"].detach().cpu()              # [E * 2A, H, 1]
        self.out_dim = w.shape[0] // self.expert_num                # 2 * actions: mean | log std
        self.register_buffer("expert_w", w.reshape(self.expert_num, self.out_dim, w.shape[1]).clone())
        self.register_buffer("expert_b", sd[p + "experts_out.bias"].detach().cpu().reshape(self.expert_num, self.out_dim).clone())
        self.normalizer = normalizer if normalizer is not None else nn.Identity()
        self.register_buffer("obs_no_goal_mask", torch.as_tensor(obs_no_goal_mask, dtype=torch.bool).cpu().clone())
        self.history = torch.zeros(1, history_length, num_obs)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
        x = self.normalizer(x)
        self.history = torch.cat([self.history[:, 1:], x.unsqueeze(1)], dim=1)
        latent = F.normalize(self.student_encoder(self.history.flatten(1)), p=2.0, dim=-1)
        full = torch.cat([latent, x], dim=1)
        no_goal = torch.cat([latent, x[:, self.obs_no_goal_mask]], dim=1)
        weights = torch.sigmoid(self.gate(full)).unsqueeze(-1)                                   # [B, E, 1]
        feat = self.hidden(self.backbone(no_goal)).reshape(-1, self.expert_num, self.expert_w.shape[2])
        out = torch.einsum("beh,edh->bed", feat, self.expert_w) + self.expert_b                  # [B, E, 2A]
        mu, log_std = torch.chunk(out, 2, dim=-1)
        var = torch.exp(2 * torch.clamp(log_std, -5.0, 2.0)) + 1e-9
        var_total = 1.0 / (torch.sum(weights / var, dim=1) + 1e-9)
        mean = var_total * torch.sum(weights * mu / var, dim=1)
        return mean, (weights.squeeze(-1), latent)

    @torch.jit.export
    def reset(self):
        self.history = torch.zeros_like(self.history)


def build_export_module(policy, normalizer=None):
    """The plain-PyTorch inference module of `policy` (ActorCritic / ActorCriticCTS / ActorCriticMoECTS, or any module whose state
    dict has the reference's key layout, e.g. one loaded from a reference checkpoint)."""
    sd = {k: v for k, v in policy.state_dict().items()}
    if getattr(policy, "is_recurrent", False):
        raise NotImplementedError("recurrent policies are not part of the go2 tasks (SURVEY section 2)")
    hist = getattr(policy, "history", None)
    if any(k.startswith("actor_mcp.") for k in sd):
        return _MCPCTSPolicy(sd, hist.shape[1], hist.shape[2], policy.obs_no_goal_mask, normalizer)
    if any(k.startswith("actor_moe.") for k in sd):
        if any(k.startswith("student_moe_encoder.") for k in sd):
            return _DualMoECTSPolicy(sd, hist.shape[1], hist.shape[2], normalizer)
        return _ACMoECTSPolicy(sd, hist.shape[1], hist.shape[2], normalizer)
    if "student_moe_encoder.experts_out.weight" in sd:
        return _MoENGCTSPolicy(sd, hist.shape[1], hist.shape[2], policy.obs_no_goal_mask, normalizer)
    if any(k.startswith("student_moe_encoder.") for k in sd):
        E = sd["student_moe_encoder.moe.gating_network.0.network.%d.weight" % max(
            int(k.split(".")[-2]) for k in sd if k.startswith("student_moe_encoder.moe.gating_network.0.network.") and k.endswith(".weight"))].shape[0]
        return _MoECTSPolicy(sd, hist.shape[1], hist.shape[2], E, normalizer)
    if any(k.startswith("student_encoder.") for k in sd):
        return _CTSPolicy(sd, hist.shape[1], hist.shape[2], normalizer)
    if any(k.startswith("actor.") for k in sd):
        return _ActorPolicy(sd, normalizer)
    raise ValueError("Policy does not have an actor/student module.")


def export_policy_as_jit(policy, path, normalizer=None, filename="policy.pt"):
    """TorchScript file of the inference policy (same arguments as the reference's export_policy_as_jit)."""
    os.makedirs(path, exist_ok=True)
    module = build_export_module(policy, normalizer).to("cpu").eval()
    scripted = torch.jit.script(module)
    scripted.save(os.path.join(path, filename))
    return os.path.join(path, filename)


def export_policy_as_pkl(policy, path, filename="policy.pkl"):
    """state_dict pickle (exporter.py:44-58)."""
    os.makedirs(path, exist_ok=True)
    torch.save({k: v.detach().cpu().clone() for k, v in policy.state_dict().items()}, os.path.join(path, filename))
    return os.path.join(path, filename)


def export_policy_as_onnx(policy, path, normalizer=None, filename="policy.onnx", verbose=False):
    raise NotImplementedError("ONNX export is out of scope (the deployment loops in deploy/ load policy.pt)")

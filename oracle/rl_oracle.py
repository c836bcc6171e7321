"""CPU ORACLE for the trainer path — TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs).

A plain fp32 PyTorch restatement (autograd + torch.optim.Adam) of
  rsl_rl/modules/actor_critic.py:38-136      ActorCritic forward, Normal log-prob / entropy
  rsl_rl/storage/rollout_storage.py:123-137  GAE + advantage normalisation
  rsl_rl/algorithms/ppo.py:120-187           PPO.update (KL-adaptive LR, clipped surrogate / value loss, clip_grad_norm_, Adam)
PINNED: tests/golden/rl_ppo.npz was produced by the reference's own rsl_rl imported from /root/reference
(tests/golden/make_golden_rl.py); tests/test_rl_oracle_golden.py checks this file against it."""
import math

import torch
import torch.nn.functional as F


def mlp_forward(sd, prefix, x):
    idx = sorted({int(k.split(".")[1]) for k in sd if k.startswith(prefix + ".") and k.endswith(".weight")})
    for n, i in enumerate(idx):
        x = F.linear(x, sd[f"{prefix}.{i}.weight"], sd[f"{prefix}.{i}.bias"])
        if n < len(idx) - 1:
            x = F.elu(x)
    return x


def log_prob(mu, std, a):
    return (-((a - mu) ** 2) / (2 * std * std) - torch.log(std) - math.log(math.sqrt(2 * math.pi))).sum(-1)


def gae(rewards, values, dones, last_values, gamma, lam):
    """rewards/values/dones [T,N,1]; returns (returns, normalised advantages)."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    adv = 0
    for t in reversed(range(T)):
        nv = last_values if t == T - 1 else values[t + 1]
        nt = 1.0 - dones[t].float()
        delta = rewards[t] + nt * gamma * nv - values[t]
        adv = delta + nt * gamma * lam * adv
        returns[t] = adv + values[t]
    a = returns - values
    return returns, (a - a.mean()) / (a.std() + 1e-8)


def ppo_update(sd, data, indices, cfg):
    """sd: state_dict (std, actor.*, critic.*) of float32 CPU tensors, updated IN PLACE semantics returned as new dict.
    data: dict of flattened [T*N, ·] tensors obs, critic_obs, actions, values, returns, old_logp, adv, old_mu, old_sigma.
    Returns (new_sd, mean_value_loss, mean_surrogate_loss, lr, per-step records)."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    order = ["std"] + [k for k in sd if k != "std"]
    opt = torch.optim.Adam([params[k] for k in order], lr=cfg["learning_rate"])
    lr = cfg["learning_rate"]
    mb = indices.numel() // cfg["num_mini_batches"]
    mvl = msl = 0.0
    rec = []
    for epoch in range(cfg["num_learning_epochs"]):
        for i in range(cfg["num_mini_batches"]):
            b = indices[i * mb:(i + 1) * mb]
            mu = mlp_forward(params, "actor", data["obs"][b])
            std = params["std"]
            sigma = mu * 0.0 + std
            lp = log_prob(mu, sigma, data["actions"][b])
            value = mlp_forward(params, "critic", data["critic_obs"][b])
            entropy = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma)).sum(-1)
            old_mu, old_sigma = data["old_mu"][b], data["old_sigma"][b]
            if cfg.get("desired_kl") is not None and cfg.get("schedule") == "adaptive":
                with torch.no_grad():
                    kl = torch.sum(torch.log(sigma / old_sigma + 1.e-5) + (old_sigma ** 2 + (old_mu - mu) ** 2) / (2.0 * sigma ** 2) - 0.5, axis=-1)
                    kl_mean = kl.mean()
                    if kl_mean > cfg["desired_kl"] * 2.0:
                        lr = max(1e-5, lr / 1.5)
                    elif kl_mean < cfg["desired_kl"] / 2.0 and kl_mean > 0.0:
                        lr = min(1e-2, lr * 1.5)
                    for g in opt.param_groups:
                        g["lr"] = lr
            adv = data["adv"][b].squeeze(-1)
            ratio = torch.exp(lp - data["old_logp"][b].squeeze(-1))
            surr = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1.0 - cfg["clip_param"], 1.0 + cfg["clip_param"])).mean()
            tv, ret = data["values"][b], data["returns"][b]
            if cfg["use_clipped_value_loss"]:
                vc = tv + (value - tv).clamp(-cfg["clip_param"], cfg["clip_param"])
                vloss = torch.max((value - ret).pow(2), (vc - ret).pow(2)).mean()
            else:
                vloss = (ret - value).pow(2).mean()
            loss = surr + cfg["value_loss_coef"] * vloss - cfg["entropy_coef"] * entropy.mean()
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_([params[k] for k in order], cfg["max_grad_norm"])
            opt.step()
            mvl += vloss.item(); msl += surr.item()
            rec.append((float(vloss.detach()), float(surr.detach()), float(lr)))
    n = cfg["num_learning_epochs"] * cfg["num_mini_batches"]
    return {k: v.detach() for k, v in params.items()}, mvl / n, msl / n, lr, rec

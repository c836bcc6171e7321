#!/usr/bin/env python3
"""Generate tests/golden/rl_ppo.npz from the REFERENCE's rsl_rl (imported unmodified from /root/reference/rsl_rl):
ActorCritic forward, RolloutStorage.compute_returns, PPO.update on seeded synthetic rollouts with an injected permutation.
Runs only in the build container.  Usage: python tests/golden/make_golden_rl.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path[:0] = ["/root/reference/rsl_rl", HERE]

from rsl_rl.algorithms.ppo import PPO  # noqa: E402
from rsl_rl.modules.actor_critic import ActorCritic  # noqa: E402
import rsl_rl.storage.rollout_storage as RS  # noqa: E402

from rl_cfg import CFG  # noqa: E402


def main():
    torch.manual_seed(0)
    N, T, H = 32, 24, [64, 32, 16]
    ac = ActorCritic(45, 263, 12, actor_hidden_dims=H, critic_hidden_dims=H, activation="elu", init_noise_std=1.0)
    sd0 = {k: v.clone() for k, v in ac.state_dict().items()}
    alg = PPO(ac, device="cpu", **CFG)
    alg.init_storage(N, T, [45], [263], [12])
    g = torch.Generator().manual_seed(1)
    obs = torch.randn(T + 1, N, 45, generator=g)
    priv = torch.randn(T + 1, N, 263, generator=g)
    rew = 0.1 * torch.randn(T, N, generator=g)
    dones = (torch.rand(T, N, generator=g) < 0.03)
    touts = dones & (torch.rand(T, N, generator=g) < 0.5)
    with torch.inference_mode():
        for t in range(T):
            alg.act(obs[t], priv[t])
            alg.process_env_step(rew[t], dones[t], {"time_outs": touts[t]})
        alg.compute_returns(priv[T])
    st = alg.storage
    save = {f"sd0_{k}": v.numpy() for k, v in sd0.items()}
    for k in ("observations", "privileged_observations", "actions", "rewards", "dones", "values", "returns", "advantages", "actions_log_prob", "mu", "sigma"):
        save["st_" + k] = getattr(st, k).clone().numpy()
    save["in_obs"], save["in_priv"], save["in_rew"], save["in_dones"], save["in_touts"] = obs.numpy(), priv.numpy(), rew.numpy(), dones.numpy(), touts.numpy()
    perm = torch.randperm(T * N, generator=g)
    save["perm"] = perm.numpy()
    orig = RS.torch.randperm
    RS.torch.randperm = lambda n, **kw: perm.clone()
    try:
        mvl, msl = alg.update()
    finally:
        RS.torch.randperm = orig
    for k, v in ac.state_dict().items():
        save[f"sd1_{k}"] = v.detach().clone().numpy()
    save["mean_value_loss"], save["mean_surrogate_loss"], save["lr"] = mvl, msl, alg.learning_rate
    osd = alg.optimizer.state_dict()
    save["adam_exp_avg_1"] = osd["state"][1]["exp_avg"].numpy()   # actor.0.weight
    path = os.path.join(HERE, "rl_ppo.npz")
    np.savez_compressed(path, **save)
    print("wrote", path, f"{os.path.getsize(path)/1024:.0f} KiB  vloss {mvl:.5f} surr {msl:.5f} lr {alg.learning_rate}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Generate tests/golden/rl_moe_cts.npz from the REFERENCE's rsl_rl (ActorCriticMoECTS + MoECTS + RolloutStorageCTS, imported
unmodified from /root/reference/rsl_rl): a seeded synthetic rollout through act / process_env_step / compute_returns and one
update() with injected teacher / student permutations.  `--variant moe_ng_cts` does the same for ActorCriticMoENGCTS + MoENGCTS (-> rl_moe_ng_cts.npz), `--variant mcp_cts` for ActorCriticMCPCTS + MCPCTS, `--variant ac_moe_cts` / `dual_moe_cts` for ActorCriticACMoECTS + ACMoECTS / ActorCriticDualMoECTS + DualMoECTS, `--variant cts` for ActorCriticCTS + CTS (-> rl_cts.npz; the
reference allocates that module's history on 'cuda' at construction, actor_critic_cts.py:48, so torch.zeros is wrapped to drop the
device while the module is built).  Build container only.  Usage (from /tmp): python /root/repo/tests/golden/make_golden_cts.py [--variant cts]"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path[:0] = ["/root/reference/rsl_rl", HERE]

from rsl_rl.algorithms.moe_cts import MoECTS  # noqa: E402
from rsl_rl.modules.actor_critic_moe_cts import ActorCriticMoECTS  # noqa: E402
import rsl_rl.storage.rollout_storage_cts as RS  # noqa: E402
from cts_cfg import ALG, ALG_CTS, POLICY, POLICY_AC, POLICY_CTS, POLICY_DUAL, POLICY_MCP, POLICY_NG  # noqa: E402


def main():
    variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else "moe_cts"
    torch.manual_seed(0)
    N, T, H = 32, (24 if variant == "moe_cts" else 12), 5
    if variant == "moe_cts":
        model = ActorCriticMoECTS(45, 263, 12, N, H, **POLICY)
    elif variant == "moe_ng_cts":     # experts on the history without its command columns (actor_critic_moe_ng_cts.py)
        from rsl_rl.modules.actor_critic_moe_ng_cts import ActorCriticMoENGCTS
        model = ActorCriticMoENGCTS(45, 263, 12, N, H, **POLICY_NG)
    elif variant == "ac_moe_cts":     # MoE actor, value experts weighted by the actor's gate (actor_critic_ac_moe_cts.py)
        from rsl_rl.modules.actor_critic_ac_moe_cts import ActorCriticACMoECTS
        model = ActorCriticACMoECTS(45, 263, 12, N, H, **POLICY_AC)
    elif variant == "mcp_cts":        # multiplicative compositional actor with a state-dependent sigma (actor_critic_mcp_cts.py)
        from rsl_rl.modules.actor_critic_mcp_cts import ActorCriticMCPCTS
        model = ActorCriticMCPCTS(45, 263, 12, N, H, **POLICY_MCP)
    elif variant == "dual_moe_cts":   # the same with the MoE student encoder (actor_critic_dual_moe_cts.py)
        from rsl_rl.modules.actor_critic_dual_moe_cts import ActorCriticDualMoECTS
        model = ActorCriticDualMoECTS(45, 263, 12, N, H, **POLICY_DUAL)
    else:
        from rsl_rl.algorithms.cts import CTS
        import rsl_rl.modules.actor_critic_cts as ACC
        zeros = torch.zeros
        ACC.torch.zeros = lambda *a, **kw: zeros(*a, **{k: v for k, v in kw.items() if k != "device"})
        try:
            model = ACC.ActorCriticCTS(45, 263, 12, N, H, **POLICY_CTS)
        finally:
            ACC.torch.zeros = zeros
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    if variant == "moe_ng_cts":
        from rsl_rl.algorithms.moe_ng_cts import MoENGCTS
        alg = MoENGCTS(model, N, H, device="cpu", **ALG)
    elif variant == "ac_moe_cts":
        from rsl_rl.algorithms.ac_moe_cts import ACMoECTS
        alg = ACMoECTS(model, N, H, device="cpu", **ALG)
    elif variant == "mcp_cts":
        from rsl_rl.algorithms.mcp_cts import MCPCTS
        alg = MCPCTS(model, N, H, device="cpu", **ALG_CTS)
    elif variant == "dual_moe_cts":
        from rsl_rl.algorithms.dual_moe_cts import DualMoECTS
        alg = DualMoECTS(model, N, H, device="cpu", **ALG)
    else:
        alg = MoECTS(model, N, H, device="cpu", **ALG) if variant == "moe_cts" else CTS(model, N, H, device="cpu", **ALG_CTS)
    alg.init_storage(N, T, [45], [263], [12])
    g = torch.Generator().manual_seed(1)
    obs = torch.randn(T + 1, N, 45, generator=g)
    priv = torch.randn(T + 1, N, 263, generator=g)
    hist = torch.randn(T + 1, N, H * 45, generator=g)
    rew = 0.1 * torch.randn(T, N, generator=g)
    dones = (torch.rand(T, N, generator=g) < 0.03)
    touts = dones & (torch.rand(T, N, generator=g) < 0.5)
    with torch.inference_mode():
        for t in range(T):
            alg.act(obs[t], priv[t], hist[t])
            alg.process_env_step(rew[t], dones[t], {"time_outs": touts[t]})
        if variant in ("ac_moe_cts", "dual_moe_cts"):      # the value needs the actor's gate, hence the observations (ac_moe_cts.py:136)
            alg.compute_returns(obs[T], priv[T], hist[T])
        else:
            alg.compute_returns(priv[T], hist[T])
    st = alg.storage
    save = {f"sd0_{k}": v.numpy() for k, v in sd0.items()}
    for k in ("observations", "privileged_observations", "history", "actions", "rewards", "dones", "values", "returns", "advantages",
              "actions_log_prob", "mu", "sigma"):
        save["st_" + k] = getattr(st, k).clone().numpy()
    save.update(in_obs=obs.numpy(), in_priv=priv.numpy(), in_hist=hist.numpy(), in_rew=rew.numpy(), in_dones=dones.numpy(), in_touts=touts.numpy())
    nt, ns = alg.teacher_num_envs * T, alg.student_num_envs * T
    tperm, sperm = torch.randperm(nt, generator=g), torch.randperm(ns, generator=g)
    save["tperm"], save["sperm"] = tperm.numpy(), sperm.numpy()
    queue = [tperm.clone(), sperm.clone()]
    orig = RS.torch.randperm
    RS.torch.randperm = lambda n, **kw: queue.pop(0)
    try:
        losses = alg.update()
    finally:
        RS.torch.randperm = orig
    for k, v in model.state_dict().items():
        save[f"sd1_{k}"] = v.detach().clone().numpy()
    save["losses"] = np.array(losses, dtype=np.float64)
    save["lr"] = alg.learning_rate
    path = os.path.join(HERE, f"rl_{variant}.npz")
    np.savez_compressed(path, **save)
    print("wrote", path, f"{os.path.getsize(path)/1024:.0f} KiB losses", losses, "lr", alg.learning_rate)


if __name__ == "__main__":
    main()

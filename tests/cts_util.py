"""Shared set-up of the CTS-family parity tests (CPU: tests/test_emu_rl_cpu.py; GPU: tests/test_gpu_x_moe_heads.py)."""
import torch

STORAGE_KEYS = ("observations", "privileged_observations", "history", "actions", "rewards", "dones", "values", "returns", "advantages",
                "actions_log_prob", "mu", "sigma")


def make_cts(variant, Z, device):
    """(model, algorithm, T, N) of a CTS-family variant with the fixture's initial weights (shared with tests/test_gpu_cts.py)."""
    from golden import cts_cfg as cc
    from go2_rl_gym_b200.rl import algorithms as A, modules as Mo
    T, N = Z["st_rewards"].shape[:2]
    model_cls, alg_cls, policy, alg_kw = {
        "moe_cts": (Mo.ActorCriticMoECTS, A.MoECTS, cc.POLICY, cc.ALG),
        "moe_ng_cts": (Mo.ActorCriticMoENGCTS, A.MoENGCTS, cc.POLICY_NG, cc.ALG),
        "cts": (Mo.ActorCriticCTS, A.CTS, cc.POLICY_CTS, cc.ALG_CTS),
        "ac_moe_cts": (getattr(Mo, "ActorCriticACMoECTS", None), getattr(A, "ACMoECTS", None), getattr(cc, "POLICY_AC", None), cc.ALG),
        "dual_moe_cts": (getattr(Mo, "ActorCriticDualMoECTS", None), getattr(A, "DualMoECTS", None), getattr(cc, "POLICY_DUAL", None), cc.ALG),
        "mcp_cts": (getattr(Mo, "ActorCriticMCPCTS", None), getattr(A, "MCPCTS", None), getattr(cc, "POLICY_MCP", None), cc.ALG_CTS),
    }[variant]
    model = model_cls(45, 263, 12, N, 5, **policy)
    model.load_state_dict({k[4:]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith("sd0_")})
    alg = alg_cls(model, N, 5, device=device, **alg_kw)
    alg.init_storage(N, T, [45], [263], [12])
    return model, alg, T, N

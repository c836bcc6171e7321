"""Checkpoint interop: the modules expose exactly the reference's state_dict keys and shapes (fixtures made by the reference)."""
import os

import numpy as np

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _keys(z):
    return {k[4:]: z[k].shape for k in z.files if k.startswith("sd0_")}


def test_actor_critic_keys():
    from go2_rl_gym_b200.rl.modules import ActorCritic
    ref = _keys(np.load(os.path.join(G, "rl_ppo.npz")))
    m = ActorCritic(45, 263, 12, actor_hidden_dims=[64, 32, 16], critic_hidden_dims=[64, 32, 16])
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(s) for k, s in ref.items()}


def test_actor_critic_moe_cts_keys():
    from go2_rl_gym_b200.rl.modules import ActorCriticMoECTS
    from golden.cts_cfg import POLICY
    ref = _keys(np.load(os.path.join(G, "rl_moe_cts.npz")))
    m = ActorCriticMoECTS(45, 263, 12, 32, 5, **POLICY)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(s) for k, s in ref.items()}
    full = ActorCriticMoECTS(45, 263, 12, 32, 5)
    assert sum(p.numel() for p in full.parameters()) == 1884609


def test_actor_critic_moe_ng_cts_keys():
    from go2_rl_gym_b200.rl.modules import ActorCriticMoENGCTS
    from golden.cts_cfg import POLICY_NG
    ref = _keys(np.load(os.path.join(G, "rl_moe_ng_cts.npz")))
    m = ActorCriticMoENGCTS(45, 263, 12, 32, 5, **POLICY_NG)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(s) for k, s in ref.items()}
    assert [k for k, _ in m.named_parameters()] == [k[4:] for k in np.load(os.path.join(G, "rl_moe_ng_cts.npz")).files if k.startswith("sd0_")]
    full = ActorCriticMoENGCTS(45, 263, 12, 32, 5, POLICY_NG["obs_no_goal_mask"])
    assert sum(p.numel() for p in full.parameters()) == 1876929          # the reference's ActorCriticMoENGCTS at GO2CfgMoENGCTS widths


def test_actor_critic_ac_moe_and_dual_moe_cts_keys():
    """Keys, shapes AND parameter order (the optimiser state dicts index parameters by position) of the MoE-actor variants."""
    from go2_rl_gym_b200.rl.modules import ActorCriticACMoECTS, ActorCriticDualMoECTS
    from golden.cts_cfg import POLICY_AC, POLICY_DUAL
    for cls, pol, name in ((ActorCriticACMoECTS, POLICY_AC, "ac_moe_cts"), (ActorCriticDualMoECTS, POLICY_DUAL, "dual_moe_cts")):
        z = np.load(os.path.join(G, f"rl_{name}.npz"))
        ref = _keys(z)
        m = cls(45, 263, 12, 32, 5, **pol)
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(s) for k, s in ref.items()}
        assert [k for k, _ in m.named_parameters()] == [k[4:] for k in z.files if k.startswith("sd0_")]


def test_actor_critic_mcp_cts_keys():
    from go2_rl_gym_b200.rl.modules import ActorCriticMCPCTS
    from golden.cts_cfg import POLICY_MCP
    z = np.load(os.path.join(G, "rl_mcp_cts.npz"))
    m = ActorCriticMCPCTS(45, 263, 12, 32, 5, **POLICY_MCP)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(s) for k, s in _keys(z).items()}
    assert [k for k, _ in m.named_parameters()] == [k[4:] for k in z.files if k.startswith("sd0_")]
    assert "std" not in m.state_dict()          # sigma is a network output (actor_critic_mcp_cts.py:236-247)


def test_shim_packages_export_the_reference_names():
    """`rsl_rl.*` / `legged_gym.utils` at the repo root re-export every name the reference's packages export (rsl_rl/rsl_rl/{algorithms,modules,
    runners,storage}/__init__.py, legged_gym/utils/__init__.py) except the recurrent policy, which no go2 task uses."""
    import importlib
    import sys
    for k in [k for k in sys.modules if k.split(".")[0] in ("rsl_rl", "legged_gym")]:
        del sys.modules[k]
    want = {"rsl_rl.algorithms": ["PPO", "CTS", "MoENGCTS", "MCPCTS", "ACMoECTS", "DualMoECTS", "MoECTS"],
            "rsl_rl.modules": ["ActorCritic", "ActorCriticCTS", "ActorCriticMoENGCTS", "ActorCriticMCPCTS", "ActorCriticACMoECTS", "ActorCriticDualMoECTS",
                               "ActorCriticMoECTS"],
            "rsl_rl.runners": ["OnPolicyRunner", "OnPolicyRunnerCTS"], "rsl_rl.storage": ["RolloutStorage", "RolloutStorageCTS"], "rsl_rl.env": ["VecEnv"],
            "legged_gym.utils": ["class_to_dict", "get_load_path", "get_args", "set_seed", "update_class_from_dict", "task_registry", "Logger", "Terrain",
                                 "quat_apply_yaw", "wrap_to_pi", "torch_rand_sqrt_float"],
            "legged_gym.utils.exporter": ["export_policy_as_jit", "export_policy_as_pkl", "export_policy_as_onnx"]}
    for mod, names in want.items():
        m = importlib.import_module(mod)
        assert m.__file__.startswith(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), (mod, m.__file__)      # the shim, not the reference
        for n in names:
            assert hasattr(m, n), (mod, n)

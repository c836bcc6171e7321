"""GPU parity tests of the fused step kernel, called through the C ABI (libgo2b200.so).

Bars: integer / flag state (reset_buf, time_out_buf, episode_length_buf, terrain_levels, last_is_limit_vel) bit-exact;
float state within the fp32 tolerances of golden_util.TOL for ONE step from an identical state (the kernel evaluates
the same algorithm with a different, warp-cooperative operation order and FMA contraction).  Contact dynamics amplify
rounding differences step over step, so multi-step checks re-synchronise the state after every compared step."""
import numpy as np
import pytest
import torch

from golden_util import load_case, compare_step, TOL
from go2_rl_gym_b200.envs.env_arrays import EnvArrays
from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg

pytestmark = pytest.mark.gpu

FLOAT_KEYS = ["obs_buf", "privileged_obs_buf", "rew_buf", "root_states", "dof_pos", "dof_vel", "torques", "commands",
              "commands_resampling_step", "commands_xy_accumulation", "env_origins", "max_move_distance", "motor_strengths",
              "motor_zero_offsets", "p_gains_multiplier", "d_gains_multiplier", "episode_sums", "base_lin_vel", "base_ang_vel",
              "projected_gravity", "measured_heights", "last_actions", "last_last_actions", "last_dof_vel", "contact_forces",
              "feet_pos", "feet_vel"]
INT_KEYS = ["reset_buf", "time_out_buf", "episode_length_buf", "terrain_levels", "last_is_limit_vel"]


def _cmp(Tg, Tc, loose=1.0):
    bad, worst = [], {}
    for k in INT_KEYS:
        if not torch.equal(Tg[k].cpu().long(), Tc[k].long()):
            bad.append((k, "exact", int((Tg[k].cpu().long() != Tc[k].long()).sum())))
    for k in FLOAT_KEYS:
        rtol, atol = TOL.get(k, TOL["default"])
        g, c = Tg[k].cpu().numpy(), Tc[k].numpy()
        err = np.abs(g - c)
        worst[k] = float(err.max())
        if not np.allclose(g, c, rtol=rtol * loose, atol=atol * loose):
            bad.append((k, float(err.max()), np.unravel_index(err.argmax(), err.shape)))
    return bad, worst


@pytest.mark.parametrize("name", ["rough", "plane"])
def test_cuda_step_matches_reference_golden(name):
    """First recorded step of the fixture: CUDA kernel vs the REFERENCE's Python run over the oracle physics."""
    from cuda_util import CudaEnv
    z, A = load_case(name, device="cuda")
    env = CudaEnv(A)
    env.common_step_counter = int(z["meta_start_counter"])
    env.step(torch.from_numpy(z["actions"][0]))
    bad = compare_step(z, 0, A.tensors)
    assert not bad, bad


@pytest.mark.parametrize("plane", [False, True])
def test_cuda_matches_oracle_rollout(plane):
    """64-step rollout, state re-synchronised to the oracle after each compared step."""
    from cuda_util import CudaEnv, copy_state
    from oracle.oracle import OracleEnv
    N = 256
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "plane" if plane else "heightfield"; cfg.seed = 11
    Ac, Ag = EnvArrays(cfg, "cpu", seed=11), EnvArrays(GO2Cfg.__new__(GO2Cfg) if False else cfg, "cuda", seed=11)
    orc, env = OracleEnv(Ac), CudaEnv(Ag)
    orc.common_step_counter = env.common_step_counter = 24 * 900
    orc.reset_all(); env.reset_all(); torch.cuda.synchronize()
    bad, _ = _cmp(Ag.tensors, Ac.tensors)
    bad = [b for b in bad if b[0] in ("root_states", "dof_pos", "commands", "motor_strengths", "episode_length_buf")]
    assert not bad, f"reset_all: {bad}"
    g = torch.Generator().manual_seed(5)
    Ac.tensors["episode_length_buf"].copy_(torch.randint(0, 1250, (N,), generator=g).int())
    copy_state(Ac.tensors, Ag.tensors)
    n_reset, worst_all = 0, {}
    for step in range(64):
        a = 0.6 * torch.randn(N, 12, generator=g)
        orc.step(a); env.step(a)
        bad, worst = _cmp(Ag.tensors, Ac.tensors)
        for k, v in worst.items():
            worst_all[k] = max(worst_all.get(k, 0.0), v)
        assert not bad, f"step {step}: {bad}"
        n_reset += int(Ac.tensors["reset_buf"].sum())
        copy_state(Ac.tensors, Ag.tensors)
    print("worst abs errors:", {k: f"{v:.2e}" for k, v in worst_all.items()})
    assert n_reset > 0


def test_cuda_substeps_match_oracle():
    from cuda_util import CudaEnv, copy_state
    from oracle.oracle import OracleEnv
    N = 128
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"; cfg.seed = 2
    Ac, Ag = EnvArrays(cfg, "cpu", seed=2), EnvArrays(cfg, "cuda", seed=2)
    orc, env = OracleEnv(Ac), CudaEnv(Ag)
    orc.reset_all()
    g = torch.Generator().manual_seed(1)
    for _ in range(20):
        orc.step(0.5 * torch.randn(N, 12, generator=g))
    copy_state(Ac.tensors, Ag.tensors)
    tau = 8.0 * torch.randn(N, 12, generator=g)
    orc.substeps(tau, 1); env.substeps(tau, 1)
    for k in ("root_states", "dof_pos", "dof_vel", "contact_forces"):
        rtol, atol = TOL.get(k, TOL["default"])
        assert np.allclose(Ag.tensors[k].cpu().numpy(), Ac.tensors[k].numpy(), rtol=rtol, atol=atol), k


def test_go2robot_vecenv_contract():
    from go2_rl_gym_b200.envs.go2.go2_env import Go2Robot
    cfg = GO2Cfg(); cfg.env.num_envs = 128; cfg.terrain.mesh_type = "heightfield"
    env = Go2Robot(cfg, None, None, "cuda:0", True)
    obs, priv = env.reset()
    assert obs.shape == (128, 45) and priv.shape == (128, 263) and obs.is_cuda
    env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length))
    total_resets = 0
    for _ in range(30):
        obs, priv, rew, dones, extras = env.step(torch.randn(128, 12, device="cuda"))
        assert rew.shape == (128,) and dones.dtype == torch.bool and "time_outs" in extras
        assert torch.isfinite(obs).all() and torch.isfinite(priv).all() and torch.isfinite(rew).all()
        total_resets += int(dones.sum())
    assert env.common_step_counter == 31
    assert set(["rew_tracking_lin_vel", "terrain_level_all"]).issubset(extras["episode"].keys())

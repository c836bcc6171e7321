"""CPU coverage of the kernel's own source: env_step_core.cuh compiled with g++ and executed lane by lane
(tests/emu) must reproduce the reference-made golden fixtures and track the oracle over a rollout."""
import numpy as np
import pytest
import torch

from golden_util import load_case, compare_step, TOL
from go2_rl_gym_b200.envs.env_arrays import EnvArrays
from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
from emu.emu import EmuEnv
from oracle.oracle import OracleEnv
from cuda_util import copy_state


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("name", ["rough", "plane"])
def test_emulated_kernel_matches_reference_golden(name, packed):
    z, A = load_case(name)
    env = EmuEnv(A, packed=packed)
    env.common_step_counter = int(z["meta_start_counter"])
    env.step(torch.from_numpy(z["actions"][0]))
    bad = compare_step(z, 0, A.tensors)
    assert not bad, bad


@pytest.mark.parametrize("packed,N", [(False, 64), (True, 64), (True, 61)])
def test_emulated_kernel_tracks_oracle(packed, N):
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"; cfg.seed = 11
    Ac, Ae = EnvArrays(cfg, "cpu", seed=11), EnvArrays(cfg, "cpu", seed=11)
    orc, env = OracleEnv(Ac), EmuEnv(Ae, packed=packed)
    orc.common_step_counter = env.common_step_counter = 24 * 900
    orc.reset_all(); env.reset_all()
    for k in ("root_states", "dof_pos", "commands", "motor_strengths", "p_gains_multiplier", "commands_resampling_step"):
        assert torch.equal(Ac.tensors[k], Ae.tensors[k]), k      # reset draws are bit-exact
    g = torch.Generator().manual_seed(5)
    Ac.tensors["episode_length_buf"].copy_(torch.randint(0, 1250, (N,), generator=g).int())
    copy_state(Ac.tensors, Ae.tensors)
    n_reset = 0
    for step in range(40):
        a = 0.6 * torch.randn(N, 12, generator=g)
        orc.step(a); env.step(a)
        for k in ("reset_buf", "time_out_buf", "episode_length_buf", "terrain_levels", "last_is_limit_vel"):
            assert torch.equal(Ac.tensors[k], Ae.tensors[k]), (step, k)
        for k in ("obs_buf", "privileged_obs_buf", "rew_buf", "root_states", "dof_pos", "dof_vel", "contact_forces", "episode_sums"):
            rtol, atol = TOL.get(k, TOL["default"])
            assert np.allclose(Ae.tensors[k].numpy(), Ac.tensors[k].numpy(), rtol=rtol, atol=atol), (step, k)
        n_reset += int(Ac.tensors["reset_buf"].sum())
        copy_state(Ac.tensors, Ae.tensors)
    assert n_reset > 0


def test_packed_map_is_bit_identical_to_warp_per_env():
    """The two thread maps run the same arithmetic in the same order: every buffer must be bit-identical over a rollout with resets."""
    N = 29   # partial last group
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"; cfg.seed = 4
    A0, A1 = EnvArrays(cfg, "cpu", seed=4), EnvArrays(cfg, "cpu", seed=4)
    e0, e1 = EmuEnv(A0, packed=False), EmuEnv(A1, packed=True)
    e0.common_step_counter = e1.common_step_counter = 24 * 100
    e0.reset_all(); e1.reset_all()
    g = torch.Generator().manual_seed(9)
    ep = torch.randint(1200, 1250, (N,), generator=g).int()      # time-outs within the rollout
    A0.tensors["episode_length_buf"].copy_(ep); A1.tensors["episode_length_buf"].copy_(ep)
    n_reset = 0
    for step in range(60):
        a = 1.5 * torch.randn(N, 12, generator=g)
        e0.step(a); e1.step(a)
        n_reset += int(A0.tensors["reset_buf"].sum())
        for k, v in A0.tensors.items():
            if k == "ep_accum":
                continue        # scratch of the cross-env logging sums (summation order differs)
            assert torch.equal(v, A1.tensors[k]) or (torch.isnan(v) == torch.isnan(A1.tensors[k])).all() and torch.equal(torch.nan_to_num(v), torch.nan_to_num(A1.tensors[k])), (step, k)
    assert n_reset > 5


@pytest.mark.parametrize("packed", [False, True])
def test_emulated_kernel_crosses_command_range_curriculum(packed):
    """Fixture made by the reference across learning iteration 20 000 (the command-range curriculum widens lin_vel_x / ang_vel_yaw,
    legged_robot.py:433-446): the kernel source must draw the same commands, timers and flags at every step of the window."""
    z, A = load_case("cmdcur")
    env = EmuEnv(A, packed=packed)
    env.common_step_counter = int(z["meta_start_counter"])
    acts = torch.from_numpy(z["actions"])
    assert env.common_step_counter < 24 * 20000 <= env.common_step_counter + int(z["meta_K"])
    for i in range(int(z["meta_K"])):
        env.step(acts[i])
        bad = compare_step(z, i, A.tensors, keys=("commands", "commands_resampling_step", "last_is_limit_vel", "reset_buf", "time_out_buf",
                                                  "episode_length_buf", "terrain_levels", "motor_strengths", "p_gains_multiplier"))
        assert not bad, (i, bad)

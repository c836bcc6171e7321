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


@pytest.mark.parametrize("name", ["rough", "plane"])
def test_emulated_kernel_matches_reference_golden(name):
    z, A = load_case(name)
    env = EmuEnv(A)
    env.common_step_counter = int(z["meta_start_counter"])
    env.step(torch.from_numpy(z["actions"][0]))
    bad = compare_step(z, 0, A.tensors)
    assert not bad, bad


def test_emulated_kernel_tracks_oracle():
    N = 64
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"; cfg.seed = 11
    Ac, Ae = EnvArrays(cfg, "cpu", seed=11), EnvArrays(cfg, "cpu", seed=11)
    orc, env = OracleEnv(Ac), EmuEnv(Ae)
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

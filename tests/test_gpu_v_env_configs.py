"""GPU parity of the library's env step against the CPU oracle for configurations other than GO2 training: the evaluation set-up of
legged_gym/scripts/play.py and two mixes of the remaining config switches (the CPU twin, on the kernel-source emulation, is
tests/test_emu_cpu.py::test_emulated_kernel_tracks_oracle_off_the_training_defaults; the oracle side of the same switches is pinned against the
reference's own Python by tests/tools/fuzz_reference_parity.py --switches)."""
import numpy as np
import pytest
import torch

from go2_rl_gym_b200.envs.env_arrays import EnvArrays
from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
from golden_util import BARE, FLIP, ODD, PLAY, TOL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,overrides", [("play", PLAY), ("odd", ODD), ("bare", BARE), ("flip", FLIP)])
def test_cuda_tracks_oracle_off_the_training_defaults(name, overrides):
    from cuda_util import CudaEnv, copy_state
    from oracle.oracle import OracleEnv
    N = 256
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"; cfg.seed = 17
    for path, val in overrides.items():
        node, parts = cfg, path.split(".")
        for p in parts[:-1]:
            node = getattr(node, p)
        setattr(node, parts[-1], val)
    Ac, Ag = EnvArrays(cfg, "cpu", seed=17), EnvArrays(cfg, "cuda", seed=17)
    orc, env = OracleEnv(Ac), CudaEnv(Ag)
    orc.common_step_counter = env.common_step_counter = 24 * 3000
    orc.reset_all(); env.reset_all(); torch.cuda.synchronize()
    g = torch.Generator().manual_seed(9)
    Ac.tensors["episode_length_buf"].copy_(torch.randint(0, int(Ac.max_episode_length), (N,), generator=g).int())
    copy_state(Ac.tensors, Ag.tensors)
    n_reset = 0
    for step in range(30):
        a = 0.6 * torch.randn(N, 12, generator=g)
        orc.step(a); env.step(a)
        for k in ("reset_buf", "time_out_buf", "episode_length_buf", "terrain_levels", "last_is_limit_vel"):
            assert torch.equal(Ac.tensors[k], Ag.tensors[k].cpu()), (name, step, k)
        for k in ("obs_buf", "privileged_obs_buf", "rew_buf", "root_states", "dof_pos", "dof_vel", "commands", "episode_sums", "xrew_sums", "turn_over_timer"):
            rtol, atol = TOL.get(k, TOL["default"])
            assert np.allclose(Ag.tensors[k].cpu().numpy(), Ac.tensors[k].numpy(), rtol=rtol, atol=atol), (name, step, k)
        n_reset += int(Ac.tensors["reset_buf"].sum())
        copy_state(Ac.tensors, Ag.tensors)
    assert n_reset > 0


@pytest.mark.parametrize("mode", ["Q4", "Q2", "H14"])
def test_q_thread_maps_agree_with_the_default(mode):
    """"Q4" / "Q2" (4 / 2 envs packed per 128- / 64-thread CTA, 4 / 8 CTAs per SM; go2_env_set_step_mode) against the default "P2" map: the
    comparison of tests/test_gpu_properties.py::test_thread_maps_agree, incl. a partially filled last CTA and envs that time out."""
    from test_gpu_properties import test_thread_maps_agree
    test_thread_maps_agree(mode)

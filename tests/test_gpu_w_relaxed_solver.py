"""GPU parity of the SECOND build of the library (libgo2b200_relaxed.so: step kernel compiled with -DGO2_RELAXED_SOLVER=1 — the convergent
contact / joint-limit solver and the state guard of DESIGN.md section 3) against the CPU oracle, through the C ABI.

The default library keeps the code every measurement of round 1 was made with; this variant was validated on the CPU only (kernel-source
emulation vs the oracle, tests/test_emu_cpu.py) because the round's GPU budget was spent.  This file is its first run on hardware and is
collected after the verified GPU files."""
import os

import numpy as np
import pytest
import torch

from go2_rl_gym_b200.envs.env_arrays import EnvArrays
from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
from golden_util import TOL as _TOL

TOL = {k: (3 * r, 3 * a) for k, (r, a) in _TOL.items()}      # first hardware run: 3 x the one-step bars of the verified kernel (golden_util.TOL)

pytestmark = pytest.mark.gpu
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "go2_rl_gym_b200", "libgo2b200_relaxed.so")


def _cfg(N, relaxed=True, guard=0, seed=11):
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"; cfg.seed = seed
    if relaxed:
        cfg.sim.b200.limit_relax, cfg.sim.b200.contact_relax, cfg.sim.b200.limit_erp = 0.5, 0.7, 0.8
    cfg.sim.b200.state_guard = guard
    return cfg


def test_default_library_rejects_the_new_settings():
    from cuda_util import CudaEnv
    with pytest.raises(RuntimeError, match="relaxed"):
        CudaEnv(EnvArrays(_cfg(32), "cuda", seed=11))


@pytest.mark.parametrize("mode", ["P2", "8p"])
def test_relaxed_kernel_tracks_oracle(mode):
    """25 policy-scale steps on rough terrain with resets, state re-synchronised to the oracle after each compared
    step: flags / counters / levels exact, floats within golden_util.TOL — the same bar as the default kernel in tests/test_gpu_env.py."""
    from cuda_util import CudaEnv, copy_state
    from oracle.oracle import OracleEnv
    N = 256
    Ac, Ag = EnvArrays(_cfg(N), "cpu", seed=11), EnvArrays(_cfg(N), "cuda", seed=11)
    orc, env = OracleEnv(Ac), CudaEnv(Ag, mode=mode, lib_path=LIB)
    orc.common_step_counter = env.common_step_counter = 24 * 900
    orc.reset_all(); env.reset_all(); torch.cuda.synchronize()
    g = torch.Generator().manual_seed(5)
    Ac.tensors["episode_length_buf"].copy_(torch.randint(0, 1250, (N,), generator=g).int())
    copy_state(Ac.tensors, Ag.tensors)
    n_reset = 0
    for step in range(25):
        a = 0.6 * torch.randn(N, 12, generator=g)          # the action scale of tests/test_gpu_env.py's rollout (the bar was measured there)
        orc.step(a); env.step(a)
        for k in ("reset_buf", "time_out_buf", "episode_length_buf", "terrain_levels"):
            assert torch.equal(Ac.tensors[k], Ag.tensors[k].cpu()), (step, k)
        for k in ("obs_buf", "rew_buf", "root_states", "dof_pos", "dof_vel", "contact_forces"):
            rtol, atol = TOL.get(k, TOL["default"])
            assert np.allclose(Ag.tensors[k].cpu().numpy(), Ac.tensors[k].numpy(), rtol=rtol, atol=atol), (step, k)
        n_reset += int(Ac.tensors["reset_buf"].sum())
        copy_state(Ac.tensors, Ag.tensors)
    assert n_reset > 0


def test_relaxed_library_with_the_first_solver_equals_the_default_library():
    """limit_relax = 0 / contact_relax = 1 / state_guard = 0 on the relaxed build = the default library's results (same algorithm; the two
    builds may contract floating-point operations differently, hence the one-step tolerances)."""
    from cuda_util import CudaEnv, copy_state
    N = 512
    A0, A1 = EnvArrays(_cfg(N, relaxed=False), "cuda", seed=11), EnvArrays(_cfg(N, relaxed=False), "cuda", seed=11)
    e0, e1 = CudaEnv(A0), CudaEnv(A1, lib_path=LIB)
    e0.common_step_counter = e1.common_step_counter = 24 * 900
    e0.reset_all(); e1.reset_all(); torch.cuda.synchronize()
    for k in ("root_states", "dof_pos", "commands", "motor_strengths"):
        assert torch.equal(A0.tensors[k], A1.tensors[k]), k
    g = torch.Generator().manual_seed(2)
    for step in range(8):
        a = 0.6 * torch.randn(N, 12, generator=g)
        e0.step(a); e1.step(a)
        for k in ("reset_buf", "time_out_buf", "episode_length_buf"):
            assert torch.equal(A0.tensors[k], A1.tensors[k]), (step, k)
        for k in ("obs_buf", "rew_buf", "root_states", "dof_pos", "dof_vel"):
            rtol, atol = TOL.get(k, TOL["default"])
            assert torch.allclose(A0.tensors[k], A1.tensors[k], rtol=rtol, atol=atol), (step, k)
        copy_state(A0.tensors, A1.tensors)


def test_state_guard_contains_a_diverged_env_on_the_gpu():
    """sim.b200.state_guard = 1: poisoned envs restart and reset in the same step, every output stays finite, neighbours are untouched bit for bit
    (the CPU twin of this test runs the oracle and the kernel-source emulation: tests/test_emu_cpu.py)."""
    from cuda_util import CudaEnv
    N = 64

    def run(poison):
        A = EnvArrays(_cfg(N, guard=1, seed=4), "cuda", seed=4)
        env = CudaEnv(A, lib_path=LIB)
        env.common_step_counter = 24 * 100
        env.reset_all()
        g = torch.Generator().manual_seed(1)
        T, outs = A.tensors, []
        for step in range(4):
            if poison and step == 1:
                T["root_states"][3, 2] = float("nan")
                T["dof_vel"][5, 7] = float("inf")
                T["root_states"][9, 7:10] = torch.tensor([5000.0, 0.0, 0.0], device="cuda")
            env.step(0.3 * torch.randn(N, 12, generator=g))
            outs.append({k: T[k].clone() for k in ("obs_buf", "privileged_obs_buf", "rew_buf", "reset_buf", "time_out_buf", "root_states", "dof_pos",
                                                   "dof_vel", "torques", "contact_forces", "episode_sums")})
        return outs

    clean, dirty = run(False), run(True)
    o = dirty[1]
    assert bool(o["reset_buf"][3]) and bool(o["reset_buf"][5]) and not bool(o["time_out_buf"][3]) and not bool(o["time_out_buf"][5])
    for step in range(1, 4):
        for k, v in dirty[step].items():
            assert torch.isfinite(v.float()).all(), (step, k)
    others = [e for e in range(N) if e not in (3, 5, 9)]
    for step in range(4):
        for k in ("obs_buf", "rew_buf", "root_states", "dof_pos", "reset_buf"):
            assert torch.equal(dirty[step][k][others], clean[step][k][others]), (step, k)
    assert float(dirty[1]["root_states"][9, 7:10].norm()) <= 1000.0 * (1 + 1e-5) or bool(dirty[1]["reset_buf"][9])


@pytest.mark.parametrize("name", ["ctrl_v_pos", "ctrl_t", "heading", "rough_relaxed", "plane_relaxed"])
def test_env_switches_match_reference_golden(name):
    """control_type 'V' / 'T' and only_positive_rewards (legged_robot.py:605-618,266-267) against the fixture made by the REFERENCE's
    own Python: first recorded step, same tolerances as the kernel-source emulation (tests/test_emu_cpu.py)."""
    from cuda_util import CudaEnv
    from golden_util import compare_step, load_case
    z, A = load_case(name, device="cuda")
    env = CudaEnv(A, lib_path=LIB)
    env.common_step_counter = int(z["meta_start_counter"])
    env.step(torch.from_numpy(z["actions"][0]))
    tol = dict(TOL, dof_vel=(1e-3, 0.1), last_dof_vel=(1e-3, 0.1), torques=(1e-3, 0.1), privileged_obs_buf=(1e-3, 1e-2), obs_buf=(1e-3, 1e-2))
    bad = compare_step(z, 0, A.tensors, tol=tol)
    assert not bad, bad

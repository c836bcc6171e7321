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


# ---- the physics of the kernel's own source against mechanics (same checks as tests/test_oracle_physics.py makes on the oracle) ----------
def _emu_env(num_envs=2, packed=False, **dr_off):
    cfg = GO2Cfg(); cfg.terrain.mesh_type = "plane"; cfg.env.num_envs = num_envs; cfg.domain_rand.push_robots = False
    for k in ("randomize_motor_strength", "randomize_pd_gains", "randomize_motor_zero_offset", "randomize_friction"):
        setattr(cfg.domain_rand, k, False)
    A = EnvArrays(cfg, "cpu", seed=3)
    env = EmuEnv(A, packed=packed)
    env.reset_all()
    return A, env


@pytest.mark.parametrize("packed", [False, True])
def test_emulated_kernel_free_flight_momentum(packed):
    """No contacts, no torques: linear momentum changes by -M g t, horizontal and angular (about the COM) momentum are conserved to the
    first-order accuracy of the 5 ms semi-implicit step (fp32 kernel source; the fp64 oracle test shows the error is O(dt))."""
    from physics_checks import mechanics
    A, env = _emu_env(packed=packed)
    T = A.tensors
    T["root_states"][:, 2] = 10.0
    T["root_states"][:, 7:13] = torch.tensor([0.3, -0.2, 0.5, 1.0, -2.0, 1.5])
    T["dof_vel"][:] = torch.linspace(-3, 3, 12)
    mech = lambda: mechanics(A.model_json, T["body_inertia"][0].numpy(), T["root_states"][0].numpy(), T["dof_pos"][0].numpy().astype(np.float64),
                             T["dof_vel"][0].numpy().astype(np.float64))
    M, c0, KE0, PE0, P0, L0 = mech()
    env.substeps(torch.zeros(2, 12), 20)
    M, c, KE, PE, P, L = mech()
    assert abs(P[2] - (P0[2] - M * 9.81 * 0.1)) < 2e-2
    assert np.linalg.norm(P[:2] - P0[:2]) < 2e-2
    assert np.linalg.norm((L - np.cross(c, P)) - (L0 - np.cross(c0, P0))) < 2e-2


def test_emulated_kernel_stance_and_drop():
    """Standing on the plane under the PD controller: the trunk settles, the feet stick and carry m g; a robot dropped from 0.6 m lands, stays
    finite and comes to rest on its feet or terminates on base contact (never tunnels below the ground)."""
    A, env = _emu_env(num_envs=2)
    T = A.tensors
    T["root_states"][:] = 0; T["root_states"][:, 6] = 1
    T["root_states"][0, 2] = 0.34; T["root_states"][1, 2] = 0.60
    T["dof_pos"][:] = torch.tensor(A.default_dof_pos_np); T["dof_vel"][:] = 0
    min_z = 1.0
    for _ in range(300):
        env.step(torch.zeros(2, 12))
        for k in ("root_states", "dof_pos", "dof_vel", "contact_forces", "obs_buf"):
            assert torch.isfinite(T[k]).all(), k
        min_z = min(min_z, float(T["root_states"][:, 2].min()))
    assert min_z > 0.05                                                 # nothing sinks through the plane
    z = float(T["root_states"][0, 2])
    assert 0.22 < z < 0.32, z
    mass = float(T["body_inertia"][0, :, 0].sum())
    fz = float(T["contact_forces"][0, :, 2].sum())
    assert abs(fz - mass * 9.81) < 0.05 * mass * 9.81
    assert float(T["feet_vel"][0].abs().max()) < 0.02                   # feet stick
    assert float(T["projected_gravity"][0, 2]) < -0.99

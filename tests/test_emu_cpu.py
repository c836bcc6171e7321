"""CPU coverage of the kernel's own source: env_step_core.cuh compiled with g++ and executed lane by lane
(tests/emu) must reproduce the reference-made golden fixtures and track the oracle over a rollout."""
import numpy as np
import pytest
import torch

from golden_util import load_case, compare_step, TOL, PLAY, ODD, BARE, FLIP
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


@pytest.mark.parametrize("packed", [False, True, 14])
@pytest.mark.parametrize("name", ["ctrl_v_pos", "ctrl_t", "heading", "cmdcur", "xrew", "xrew_pos", "turn_over"])
def test_emulated_kernel_matches_reference_golden_with_env_switches(name, packed):
    """control_type 'V' / 'T', only_positive_rewards, heading commands, the 14 reward functions no registered task switches on and the command-range curriculum boundary:
    every recorded step of the reference-made fixture, state re-synchronised to the fixture between steps."""
    z, A = load_case(name)
    env = EmuEnv(A, packed=packed)
    env.common_step_counter = int(z["meta_start_counter"])
    # velocity control differentiates the joint velocity over one 5 ms substep (kd (qd - last_qd) / sim_dt): the kernel-vs-oracle operation-order
    # noise on qd is amplified by it for the robots lying on their side, hence 0.1 rad/s instead of 0.03 on the velocities of this fixture
    tol = dict(TOL, dof_vel=(1e-3, 0.1), last_dof_vel=(1e-3, 0.1), torques=(1e-3, 0.1), privileged_obs_buf=(1e-3, 1e-2), obs_buf=(1e-3, 1e-2))
    for i in range(int(z["meta_K"])):
        env.step(torch.from_numpy(z["actions"][i]))
        bad = compare_step(z, i, A.tensors, tol=tol)
        assert not bad, (i, bad)
        for k in z.files:          # continue from the reference's own state
            if k.startswith(f"out{i}_") and k[len(f"out{i}_"):] in A.tensors and not k.endswith(("obs_buf", "rew_buf")):
                t = A.tensors[k[len(f"out{i}_"):]]
                t.copy_(torch.from_numpy(z[k]).to(t.dtype).reshape(t.shape))


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


@pytest.mark.parametrize("packed", [True, 4, 2, 14])
def test_packed_map_is_bit_identical_to_warp_per_env(packed):
    """The thread maps (warp per env; 8 envs packed per CTA = "P2"; 4 envs packed per CTA = "Q4"; 14 envs per group on half-warps with dedicated
    leg warps = "H14") run the same arithmetic in the same order:
    every buffer must be bit-identical over a rollout with resets."""
    N = 29   # partial last group
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"; cfg.seed = 4
    A0, A1 = EnvArrays(cfg, "cpu", seed=4), EnvArrays(cfg, "cpu", seed=4)
    e0, e1 = EmuEnv(A0, packed=False), EmuEnv(A1, packed=packed)
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
            # incl. "ep_accum" / "ep_stats": the finished episodes' reward sums are accumulated with integer atomics in 2^-20 fixed point, so the
            # logged means do not depend on the order in which the groups run
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


# ---- the convergent contact / joint-limit solver (default: sim.b200.limit_relax = 0.5) vs round 1's first solver ------
def _legacy(cfg):
    """round 1's first solver: limit rows stepping with D_j, plain mass splitting, soft limit ERP, no state guard"""
    cfg.sim.b200.limit_relax, cfg.sim.b200.contact_relax, cfg.sim.b200.limit_erp, cfg.sim.b200.state_guard = 0.0, 1.0, 0.2, 0
    return cfg


def _relaxed_cfg(N, iters=4):
    cfg = GO2Cfg(); cfg.terrain.mesh_type = "plane"; cfg.env.num_envs = N; cfg.domain_rand.push_robots = False
    for k in ("randomize_motor_strength", "randomize_pd_gains", "randomize_motor_zero_offset", "randomize_friction", "randomize_action_delay"):
        setattr(cfg.domain_rand, k, False)
    cfg.sim.b200.solver_iterations = iters
    assert (cfg.sim.b200.limit_relax, cfg.sim.b200.contact_relax, cfg.sim.b200.limit_erp, cfg.sim.b200.state_guard) == (0.5, 0.7, 0.8, 1)   # the defaults
    return cfg


def _limit_probe(make_env, cfg, steps=150):
    """8 robots standing on the plane, each driven hard against different joint stops by its PD targets; -> worst overshoot per robot."""
    A = EnvArrays(cfg, "cpu", seed=3); env = make_env(A); env.reset_all(); T = A.tensors
    hi = torch.tensor([A.model.q_upper[j] for j in range(12)]); lo = torch.tensor([A.model.q_lower[j] for j in range(12)])
    T["root_states"][:] = 0; T["root_states"][:, 6] = 1; T["root_states"][:, 2] = 0.34
    T["dof_pos"][:] = torch.tensor(A.default_dof_pos_np); T["dof_vel"][:] = 0
    acts = torch.zeros(8, 12)
    acts[1, 2::3] = -12.0; acts[2, 2::3] = 6.0; acts[3, 0::3] = 8.0; acts[4, 1::3] = 12.0; acts[5, 1::3] = -12.0
    acts[6] = 10 * torch.randn(12, generator=torch.Generator().manual_seed(1)); acts[7] = 20 * torch.randn(12, generator=torch.Generator().manual_seed(2))
    worst = torch.zeros(8)
    for _ in range(steps):
        env.step(acts)
        q = T["dof_pos"]
        if not torch.isfinite(q).all() or not torch.isfinite(T["root_states"]).all():
            return None
        worst = torch.maximum(worst, torch.maximum(torch.relu(q - hi), torch.relu(lo - q)).max(dim=1).values)
    return worst


def test_relaxed_solver_converges_where_the_first_one_diverges():
    """Oracle (the physics spec): with the first solver the Jacobi sweeps DIVERGE when more of them are run (joint-limit rows stepping with D_j
    are over-relaxed); with the exact joint-space diagonal and a 0.5 step (contacts 0.7) nothing becomes non-finite and the overshoot shrinks as
    sweeps are added.  At the shipped 4 sweeps the joint stops get ~20x stiffer."""
    legacy4 = _limit_probe(OracleEnv, _legacy(_relaxed_cfg(8, 4)))
    assert legacy4 is not None and float(legacy4.max()) > 1.0                      # the documented weakness of the first solver
    assert _limit_probe(OracleEnv, _legacy(_relaxed_cfg(8, 16)), steps=40) is None           # ... and its divergence with more sweeps
    worst = []
    for iters in (4, 8, 16):
        cfg = _relaxed_cfg(8, iters)
        cfg.sim.b200.limit_relax, cfg.sim.b200.contact_relax, cfg.sim.b200.limit_erp = 0.5, 0.7, 0.8
        w = _limit_probe(OracleEnv, cfg)
        assert w is not None, iters
        worst.append(float(w.max()))
    assert max(worst) < 0.2 and worst[2] < 0.5 * worst[0], worst            # measured: 0.10, 0.11, 0.03 rad (first solver at 4 sweeps: 2.2 rad)


@pytest.mark.parametrize("packed", [False, True])
def test_emulated_kernel_tracks_oracle_with_relaxed_solver(packed):
    """The kernel source with sim.b200.limit_relax / contact_relax set: same single-step agreement with the oracle as the default solver, on
    rough terrain with resets, and the same stiffer joint stops in the probe."""
    N = 32
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"; cfg.seed = 11
    cfg.sim.b200.limit_relax, cfg.sim.b200.contact_relax, cfg.sim.b200.limit_erp = 0.5, 0.7, 0.8
    Ac, Ae = EnvArrays(cfg, "cpu", seed=11), EnvArrays(cfg, "cpu", seed=11)
    orc, env = OracleEnv(Ac), EmuEnv(Ae, packed=packed)
    orc.common_step_counter = env.common_step_counter = 24 * 900
    orc.reset_all(); env.reset_all()
    g = torch.Generator().manual_seed(5)
    Ac.tensors["episode_length_buf"].copy_(torch.randint(0, 1250, (N,), generator=g).int())
    copy_state(Ac.tensors, Ae.tensors)
    for step in range(25):
        a = 2.0 * torch.randn(N, 12, generator=g)          # large actions: joints reach their stops
        orc.step(a); env.step(a)
        for k in ("reset_buf", "time_out_buf", "episode_length_buf", "terrain_levels"):
            assert torch.equal(Ac.tensors[k], Ae.tensors[k]), (step, k)
        for k in ("obs_buf", "rew_buf", "root_states", "dof_pos", "dof_vel", "contact_forces"):
            rtol, atol = TOL.get(k, TOL["default"])
            assert np.allclose(Ae.tensors[k].numpy(), Ac.tensors[k].numpy(), rtol=rtol, atol=atol), (step, k)
        copy_state(Ac.tensors, Ae.tensors)
    cfgp = _relaxed_cfg(8, 4)
    cfgp.sim.b200.limit_relax, cfgp.sim.b200.contact_relax, cfgp.sim.b200.limit_erp = 0.5, 0.7, 0.8
    w = _limit_probe(lambda A: EmuEnv(A, packed=packed), cfgp)
    assert w is not None and float(w.max()) < 0.2, w


@pytest.mark.parametrize("kind", ["oracle", "emu"])
def test_relaxed_solver_holds_joint_stops_in_free_flight(kind):
    """Free flight, joints driven against their stops by a sustained torque (5 % / 20 % of the actuator limit): the parent link recoils, which is
    where the first solver's rows over-relax until the state becomes non-finite (after ~100 substeps, asserted here as the documented defect);
    the relaxed solver parks the joints at the stops (overshoot < 0.2 rad, finite)."""
    def run(scale, relaxed):
        cfg = _relaxed_cfg(2, 4)
        if not relaxed:
            _legacy(cfg)
        A = EnvArrays(cfg, "cpu", seed=3)
        env = OracleEnv(A) if kind == "oracle" else EmuEnv(A)
        env.reset_all(); T = A.tensors
        hi = torch.tensor([A.model.q_upper[j] for j in range(12)]); lo = torch.tensor([A.model.q_lower[j] for j in range(12)])
        eff = torch.tensor([A.model.effort[j] for j in range(12)])
        T["root_states"][:, 2] = 10.0
        T["dof_pos"][:] = torch.tensor(A.default_dof_pos_np); T["dof_vel"][:] = 0
        tau, worst = scale * torch.stack([eff, -eff]), 0.0
        for _ in range(60):
            env.substeps(tau, 5)
            if not torch.isfinite(T["dof_pos"]).all():
                return None
            worst = max(worst, float(torch.relu(T["dof_pos"][0] - hi).max()), float(torch.relu(lo - T["dof_pos"][1]).max()))
        return worst
    for scale in (0.05, 0.2):
        w = run(scale, True)
        assert w is not None and w < 0.2, (scale, w)
    legacy = run(0.05, False)
    assert legacy is None or legacy > 1.0, legacy          # the first solver: runaway joints / non-finite state


@pytest.mark.parametrize("kind", ["oracle", "emu", "emu_packed"])
def test_state_guard_contains_a_diverged_env(kind):
    """sim.b200.state_guard = 1 (include/go2_b200.h): envs whose state turns non-finite restart from the initial pose and reset in the same step;
    every output stays finite, their neighbours are untouched bit for bit, and the base twist is clamped at asset.max_*_velocity."""
    N = 16
    make = {"oracle": OracleEnv, "emu": lambda A: EmuEnv(A, packed=False), "emu_packed": lambda A: EmuEnv(A, packed=True)}[kind]

    def run(poison):
        cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"; cfg.seed = 4
        cfg.sim.b200.limit_relax, cfg.sim.b200.contact_relax, cfg.sim.b200.limit_erp, cfg.sim.b200.state_guard = 0.5, 0.7, 0.8, 1
        A = EnvArrays(cfg, "cpu", seed=4)
        env = make(A)
        env.common_step_counter = 24 * 100
        env.reset_all()
        g = torch.Generator().manual_seed(1)
        T = A.tensors
        outs = []
        for step in range(4):
            if poison and step == 1:
                T["root_states"][3, 2] = float("nan")
                T["dof_vel"][5, 7] = float("inf")
                T["root_states"][9, 7:10] = torch.tensor([5000.0, 0.0, 0.0])      # finite but absurd: clamped, not reset
            env.step(0.3 * torch.randn(N, 12, generator=g))
            outs.append({k: T[k].clone() for k in ("obs_buf", "privileged_obs_buf", "rew_buf", "reset_buf", "time_out_buf", "root_states", "dof_pos",
                                                   "dof_vel", "torques", "contact_forces", "episode_sums", "last_dof_vel")})
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
    # the restarted envs continue like any freshly reset env
    assert float(dirty[3]["root_states"][3, 2]) > -5.0 and float(dirty[3]["dof_pos"][5].abs().max()) < 10.0


@pytest.mark.parametrize("packed", [False, True])
def test_heading_mode_with_stop_flags_kernel_source_equals_oracle(packed):
    """heading_command with the GO2 defaults stop_heading_at_limit = True / limit_ang_vel_at_zero_command_prob > 0.  The reference itself cannot
    run this combination (it clips the masked yaw command with unmasked bounds, legged_robot.py:415-419, and raises once an env holds its
    heading), so the evident intent is pinned between the oracle and the kernel source: identical stop flags and commands, stopped envs keep
    their yaw command, the others follow 0.5 * wrap_to_pi(target - heading) clipped to their yaw range."""
    N = 64
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"; cfg.seed = 21
    cfg.commands.heading_command = True
    cfg.commands.resampling_time = 0.2                       # resample every 10 steps: many draws in a short run
    Ac, Ae = EnvArrays(cfg, "cpu", seed=21), EnvArrays(cfg, "cpu", seed=21)
    orc, env = OracleEnv(Ac), EmuEnv(Ae, packed=packed)
    orc.common_step_counter = env.common_step_counter = 24 * 2000      # zero-command probability at its final value
    orc.reset_all(); env.reset_all()
    g = torch.Generator().manual_seed(3)
    seen_stop = 0
    for step in range(40):
        a = 0.4 * torch.randn(N, 12, generator=g)
        prev_cmd, prev_stop = Ac.tensors["commands"].clone(), Ac.tensors["stop_heading"].clone()
        orc.step(a); env.step(a)
        for k in ("stop_heading", "reset_buf", "last_is_limit_vel"):
            assert torch.equal(Ac.tensors[k], Ae.tensors[k]), (step, k)
        assert torch.allclose(Ac.tensors["commands"], Ae.tensors["commands"], atol=1e-5), step
        stop = Ac.tensors["stop_heading"].bool()
        seen_stop += int(stop.sum())
        held = stop & prev_stop.bool() & ~Ac.tensors["reset_buf"].bool() & (Ac.tensors["commands"][:, 3] == prev_cmd[:, 3])
        assert torch.equal(Ac.tensors["commands"][held, 2], prev_cmd[held, 2])               # a held heading keeps its yaw command
        q = Ac.tensors["root_states"][:, 3:7]
        fwd_x = 1 - 2 * (q[:, 1] ** 2 + q[:, 2] ** 2); fwd_y = 2 * (q[:, 0] * q[:, 1] + q[:, 3] * q[:, 2])
        err = Ac.tensors["commands"][:, 3] - torch.atan2(fwd_y, fwd_x)
        want = torch.clip(0.5 * (torch.remainder(err + np.pi, 2 * np.pi) - np.pi), Ac.tensors["env_command_ranges"][:, 4], Ac.tensors["env_command_ranges"][:, 5])
        free = ~stop & ~Ac.tensors["reset_buf"].bool()
        assert torch.allclose(Ac.tensors["commands"][free, 2], want[free], atol=2e-4), step
        copy_state(Ac.tensors, Ae.tensors)
        Ae.tensors["stop_heading"].copy_(Ac.tensors["stop_heading"])
    assert seen_stop > 0


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("name,overrides", [("play", PLAY), ("odd", ODD), ("bare", BARE), ("flip", FLIP)])
def test_emulated_kernel_tracks_oracle_off_the_training_defaults(name, overrides, packed):
    """The kernel's branches for configurations other than GO2 training: legged_gym/scripts/play.py's evaluation set-up (7 x 7 terrain without
    curriculum, noise / pushes / most randomisation off) and two mixes of the remaining switches (single-interval command sampling, no dynamic
    sigma, no reward / zero-command curricula, short episodes, other clips and scales).  The oracle side of the same switches is pinned against
    the reference by tests/tools/fuzz_reference_parity.py --switches."""
    N = 40
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"; cfg.seed = 17
    for path, val in overrides.items():
        node, parts = cfg, path.split(".")
        for p in parts[:-1]:
            node = getattr(node, p)
        assert hasattr(node, parts[-1]) or parts[1] in ("scales", "turn_over_scales"), path
        setattr(node, parts[-1], val)
    Ac, Ae = EnvArrays(cfg, "cpu", seed=17), EnvArrays(cfg, "cpu", seed=17)
    orc, env = OracleEnv(Ac), EmuEnv(Ae, packed=packed)
    orc.common_step_counter = env.common_step_counter = 24 * 3000
    orc.reset_all(); env.reset_all()
    g = torch.Generator().manual_seed(9)
    Ac.tensors["episode_length_buf"].copy_(torch.randint(0, int(Ac.max_episode_length), (N,), generator=g).int())
    copy_state(Ac.tensors, Ae.tensors)
    n_reset = 0
    for step in range(30):
        a = 0.7 * torch.randn(N, 12, generator=g)
        orc.step(a); env.step(a)
        for k in ("reset_buf", "time_out_buf", "episode_length_buf", "terrain_levels", "last_is_limit_vel"):
            assert torch.equal(Ac.tensors[k], Ae.tensors[k]), (name, step, k)
        for k in ("obs_buf", "privileged_obs_buf", "rew_buf", "root_states", "dof_pos", "dof_vel", "commands", "episode_sums", "xrew_sums", "xrew_state", "turn_over_timer"):
            rtol, atol = TOL.get(k, TOL["default"])
            assert np.allclose(Ae.tensors[k].numpy(), Ac.tensors[k].numpy(), rtol=rtol, atol=atol), (name, step, k)
        n_reset += int(Ac.tensors["reset_buf"].sum())
        copy_state(Ac.tensors, Ae.tensors)
    assert n_reset > 0
    if name == "flip":
        assert float(Ac.tensors["turn_over_timer"].max()) > 0 and float(Ac.tensors["xrew_sums"].abs().max()) > 0


@pytest.mark.parametrize("kind", ["oracle", "emu"])
def test_episode_statistics_are_reserved_on_steps_without_a_reset(kind):
    """extras["episode"] semantics (legged_robot.py:229-242, on_policy_runner.py:145-146): no 'episode' entry before the first reset (valid
    flag 0), a fresh row on a step with resets, and on later steps WITHOUT a reset the previous step's row served again (copied forward in the ring),
    not whatever an older rollout left in that slot."""
    from go2_rl_gym_b200 import _abi
    cfg = GO2Cfg(); cfg.terrain.mesh_type = "plane"; cfg.env.num_envs = 4; cfg.domain_rand.push_robots = False
    A = EnvArrays(cfg, "cpu", seed=3)
    env = OracleEnv(A) if kind == "oracle" else EmuEnv(A)
    env.reset_all(); T = A.tensors
    T["ep_stats"].fill_(7.0)                                    # stale garbage in every slot
    T["ep_stats"][:, _abi.NUM_REW + 11] = 0
    rows = []
    for k in range(4):
        if k == 1:
            T["episode_length_buf"][0] = 1250                   # env 0 times out in this step
        sp = env.step(torch.zeros(4, 12))
        rows.append(T["ep_stats"][sp.ep_slot].clone())
        assert int(T["reset_buf"].sum()) == (1 if k == 1 else 0)
    v = _abi.NUM_REW + 11
    assert float(rows[0][v]) == 0.0                             # nothing to serve yet
    assert float(rows[1][v]) == 1.0 and float(rows[1][_abi.NUM_REW + 10]) == 1.0 and not torch.equal(rows[1][:14], torch.full((14,), 7.0))
    assert torch.equal(rows[2], rows[1]) and torch.equal(rows[3], rows[1])

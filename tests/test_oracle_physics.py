"""Sanity of the physics spec as restated by the oracle (PhysX parity is unpinned, so the physics is checked against
mechanics): conservation laws in free flight to first order in dt, and a stable stance with feet at rest carrying m*g."""
import numpy as np
import torch

from go2_rl_gym_b200.envs.env_arrays import EnvArrays
from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
from oracle.oracle import OracleEnv
from physics_checks import mechanics


def _free_flight(sim_dt):
    cfg = GO2Cfg(); cfg.terrain.mesh_type = "plane"; cfg.env.num_envs = 2; cfg.sim.dt = sim_dt
    A = EnvArrays(cfg, "cpu", seed=3); env = OracleEnv(A, double=True); env.reset_all(); T = A.tensors
    T["root_states"][:, 2] = 10.0
    T["root_states"][:, 7:13] = torch.tensor([0.3, -0.2, 0.5, 1.0, -2.0, 1.5])
    T["dof_vel"][:] = torch.linspace(-3, 3, 12)
    # start inside the joint limits: _reset_dofs draws q0 * U[0.5, 1.5] unclamped (the calf can start 0.088 rad past its stop), and the stop's
    # position correction limit_erp * penetration / dt is a velocity that GROWS as dt shrinks: not the first-order integration error measured here
    T["dof_pos"][:] = torch.tensor(A.default_dof_pos_np)
    mech = lambda: mechanics(A.model_json, T["body_inertia"][0].numpy(), T["root_states"][0].numpy(), T["dof_pos"][0].numpy().astype(np.float64),
                             T["dof_vel"][0].numpy().astype(np.float64))
    M, c0, KE0, PE0, P0, L0 = mech()
    env.substeps(torch.zeros(2, 12), int(round(0.1 / sim_dt)))
    M, c, KE, PE, P, L = mech()
    t = 0.1
    return (abs(P[2] - (P0[2] - M * 9.81 * t)), np.linalg.norm(P[:2] - P0[:2]),
            np.linalg.norm((L - np.cross(c, P)) - (L0 - np.cross(c0, P0))))


def test_free_flight_momentum_first_order_in_dt():
    e1, e2 = _free_flight(0.005), _free_flight(0.00125)
    for a, b in zip(e1, e2):
        assert a < 2e-2                      # momentum / angular-momentum drift over 0.1 s at 5 ms
        assert b < 0.45 * a + 1e-6           # shrinks ~4x with dt/4


def test_stance_is_stable_and_carries_weight():
    cfg = GO2Cfg(); cfg.terrain.mesh_type = "plane"; cfg.env.num_envs = 2; cfg.domain_rand.push_robots = False
    for k in ("randomize_motor_strength", "randomize_pd_gains", "randomize_motor_zero_offset", "randomize_friction"):
        setattr(cfg.domain_rand, k, False)
    A = EnvArrays(cfg, "cpu", seed=3); env = OracleEnv(A); env.reset_all(); T = A.tensors
    T["root_states"][:] = 0; T["root_states"][:, 2] = 0.34; T["root_states"][:, 6] = 1
    T["dof_pos"][:] = torch.tensor(A.default_dof_pos_np); T["dof_vel"][:] = 0
    for _ in range(300):
        env.step(torch.zeros(2, 12))
    z = float(T["root_states"][0, 2])
    assert 0.22 < z < 0.32, z
    assert not bool(T["reset_buf"].any())
    mass = float(T["body_inertia"][0, :, 0].sum())
    fz = float(T["contact_forces"][0, :, 2].sum())
    assert abs(fz - mass * 9.81) < 0.05 * mass * 9.81
    assert float(T["feet_vel"][0].abs().max()) < 0.02          # feet stick
    assert float(T["projected_gravity"][0, 2]) < -0.99

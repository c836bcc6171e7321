"""Init-time parity (VERDICT r1 'common-mode init'): every config-derived constant and per-env table the kernel receives from the
product's own `EnvArrays` is compared with what the UNMODIFIED reference's init code derives for the same config — `LeggedRobot._parse_cfg`,
`_get_env_origins`, `_process_dof_props`, `_init_buffers`, `_init_height_points`, `_get_noise_scale_vec`, `_prepare_reward_function`,
`_update_env_command_ranges` (legged_robot.py:765-940,1054-1091,1172-1186; go2_env.py:9-21), run on a `Go2Robot` object built without Isaac Gym
(tests/golden/make_golden_env.py: build_reference_env).  The parity tests of the step feed both sides from `EnvArrays`; this test is what stands
between an `EnvArrays` bug and both sides agreeing on it.  Runs in a subprocess (the reference tree and this repo's shims share package names)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import json, sys, os
sys.path.insert(0, os.path.join(%(root)r, "tests", "golden"))
import numpy as np, torch
import make_golden_env as G            # puts tests/ref_stub and /root/reference in front of the repo's shims
from go2_rl_gym_b200 import _abi
N, plane, seed = %(N)d, %(plane)r, %(seed)d
cfg = G.MyGO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "plane" if plane else "heightfield"; cfg.seed = seed
A = G.EnvArrays(cfg, "cpu", seed=seed)
O = G.OracleEnv(A)
import legged_gym.envs
from legged_gym.envs.go2.go2_config import GO2Cfg as RefGO2Cfg
ref_cfg = RefGO2Cfg(); ref_cfg.env.num_envs = N; ref_cfg.terrain.mesh_type = cfg.terrain.mesh_type
G.install_rng(G.Draws(seed), N)
r = G.build_reference_env(A, O, ref_cfg)
C, T, M = A.config, A.tensors, A.model
arr = lambda x: np.asarray(list(x), dtype=np.float64)
out = {}
def cmp(name, ref, mine, tol=0.0):
    ref, mine = np.asarray(ref, dtype=np.float64).reshape(-1), np.asarray(mine, dtype=np.float64).reshape(-1)
    out[name] = [int(ref.size), int(mine.size), float(np.abs(ref - mine).max()) if ref.size == mine.size and ref.size else -1.0, tol]
cmp("kp", r.p_gains, arr(C.kp)); cmp("kd", r.d_gains, arr(C.kd)); cmp("default_dof_pos", r.default_dof_pos, arr(C.default_dof_pos))
cmp("soft_dof_limit_lo", r.dof_pos_limits[:, 0], arr(C.soft_dof_limit_lo), 1e-6); cmp("soft_dof_limit_hi", r.dof_pos_limits[:, 1], arr(C.soft_dof_limit_hi), 1e-6)
cmp("torque_limits", r.torque_limits, arr(M.effort)); cmp("dof_vel_limits", r.dof_vel_limits, arr(M.vel_limit))
cmp("noise_scale_vec", r.noise_scale_vec, arr(C.noise_scale_vec), 1e-7)
cmp("height_points_x", r.height_points[0, :, 0], arr([C.height_points[i][0] for i in range(_abi.NUM_HEIGHT)]), 1e-6)
cmp("height_points_y", r.height_points[0, :, 1], arr([C.height_points[i][1] for i in range(_abi.NUM_HEIGHT)]), 1e-6)
cmp("base_height_mask", r.base_height_scan_mask, arr(C.base_height_mask)); cmp("num_base_height_points", [r.num_base_height_scan_points], [C.num_base_height_points])
cmp("reward_scales", [r.reward_scales.get(n, 0.0) for n in _abi.REWARD_NAMES], arr(C.reward_scales), 1e-7)
cmp("scalars", [r.dt, r.max_episode_length, r.max_episode_length_s, r.cfg.control.action_scale, r.cfg.normalization.clip_actions, r.cfg.normalization.clip_observations,
                r.obs_scales.lin_vel, r.obs_scales.ang_vel, r.obs_scales.dof_pos, r.obs_scales.dof_vel, r.obs_scales.height_measurements,
                r.cfg.domain_rand.push_interval, r.cfg.rewards.tracking_sigma, r.cfg.rewards.base_height_target, r.cfg.commands.resampling_time],
    [C.dt, C.max_episode_length, C.max_episode_length_s, C.action_scale, C.clip_actions, C.clip_obs, C.obs_scale_lin_vel, C.obs_scale_ang_vel,
     C.obs_scale_dof_pos, C.obs_scale_dof_vel, C.obs_scale_height, C.push_interval, C.tracking_sigma, C.base_height_target, C.resampling_time], 1e-6)
cmp("base_init_state", r.base_init_state, arr(C.base_init_state))
cmp("env_command_ranges", torch.cat([r.env_command_ranges[k] for k in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw")], 1), T["env_command_ranges"], 1e-6)
if not plane:
    cmp("env_origins", r.env_origins, T["env_origins"]); cmp("terrain_levels", r.terrain_levels, T["terrain_levels"]); cmp("terrain_types", r.terrain_types, T["terrain_types"])
    cmp("terrain_ids", r.terrain_ids, T["terrain_ids"]); cmp("terrain_origins", r.terrain_origins, T["terrain_origins"])
    cmp("terrain_max_sigmas", r.terrain_max_sigmas, arr(C.ds_max_sigma), 1e-7)
    cmp("terrain_dims", [r.terrain.tot_rows, r.terrain.tot_cols, r.cfg.terrain.horizontal_scale, r.cfg.terrain.vertical_scale, r.cfg.terrain.border_size, r.max_terrain_level],
        [C.hf_rows, C.hf_cols, C.hscale, C.vscale, C.border, C.num_levels], 1e-7)
print("INIT_PARITY " + json.dumps(out))
'''


def _run(N, plane, seed):
    code = SCRIPT % {"root": ROOT, "N": N, "plane": plane, "seed": seed}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""), cwd="/tmp")
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("INIT_PARITY ")][-1]
    return json.loads(line[len("INIT_PARITY "):])


@pytest.mark.skipif(not os.path.exists("/root/reference/legged_gym"), reason="reference tree not present")
@pytest.mark.parametrize("N,plane,seed", [(4096, False, 1), (120, False, 7), (64, True, 3)])
def test_envarrays_init_equals_the_reference_init(N, plane, seed):
    out = _run(N, plane, seed)
    assert len(out) >= (22 if not plane else 16)
    for name, (n_ref, n_mine, err, tol) in out.items():
        assert n_ref == n_mine and n_ref > 0, (name, n_ref, n_mine)
        assert 0.0 <= err <= tol, (name, err, tol)

"""Shared helpers: rebuild the env of a golden fixture and compare per-step outputs."""
import os

import numpy as np
import torch

from go2_rl_gym_b200.envs.env_arrays import EnvArrays
from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerances: integer / flag state is bit-exact; float state within fp32 rounding of a different summation order
EXACT = ["reset_buf", "time_out_buf", "episode_length_buf", "terrain_levels", "last_is_limit_vel", "stop_heading"]
# TOL_TIGHT: oracle post-physics vs the reference's Python over the SAME physics code (only summation order differs).
TOL_TIGHT = {"default": (2e-5, 2e-5), "rew_buf": (1e-5, 1e-5), "episode_sums": (1e-4, 1e-5), "privileged_obs_buf": (1e-4, 1e-4),
             "contact_forces": (1e-3, 1e-3), "torques": (1e-4, 1e-4), "dof_vel": (1e-4, 1e-4), "root_states": (1e-4, 1e-4),
             "obs_buf": (1e-4, 1e-4), "feet_vel": (1e-4, 1e-4), "last_dof_vel": (1e-4, 1e-4)}
# TOL: the CUDA kernel (or its host emulation) vs the oracle / golden after ONE step from an identical state, (rtol, atol).
# The kernel runs the same fp32 algorithm in block-structured form with a different operation order (and FMA contraction on
# the GPU); one 5 ms contact solve amplifies that to ~1e-2 rad/s on joint velocities of light links (measured: the oracle's
# own float-vs-double difference is 100x larger than these bounds).
TOL = {"default": (1e-3, 1e-3), "dof_pos": (1e-4, 2e-4), "root_states": (1e-4, 1e-3), "dof_vel": (1e-3, 3e-2),
       "last_dof_vel": (1e-3, 3e-2), "torques": (1e-3, 2e-2), "contact_forces": (2e-3, 0.3), "obs_buf": (1e-3, 3e-3),
       "privileged_obs_buf": (1e-3, 3e-3), "rew_buf": (1e-3, 2e-4), "episode_sums": (1e-3, 2e-4), "feet_vel": (1e-3, 1e-2),
       "feet_pos": (1e-4, 2e-4), "measured_heights": (0, 1e-6), "env_origins": (0, 0), "commands": (1e-6, 1e-6),
       "motor_strengths": (0, 0), "motor_zero_offsets": (0, 0), "p_gains_multiplier": (0, 0), "d_gains_multiplier": (0, 0)}


def load_case(name, device="cpu", **kw):
    z = np.load(os.path.join(GOLDEN, f"env_{name}.npz"))
    N, seed, plane = int(z["meta_N"]), int(z["meta_seed"]), bool(z["meta_plane"])
    cfg = GO2Cfg()
    cfg.env.num_envs = N
    cfg.terrain.mesh_type = "plane" if plane else "heightfield"
    cfg.seed = seed
    if "meta_control_type" in z.files:       # fixtures with env switches outside the GO2 defaults (make_golden_env.py --switches)
        cfg.control.control_type = "PVT"[int(z["meta_control_type"])]
        cfg.rewards.only_positive_rewards = bool(z["meta_only_positive"])
        cfg.commands.heading_command = bool(z["meta_heading"]) if "meta_heading" in z.files else False
    if "meta_overrides" in z.files:          # config switches applied to both sides by make_golden_env.make_case(overrides=...)
        import ast
        for path, val in ast.literal_eval(str(z["meta_overrides"])).items():
            node, parts = cfg, path.split(".")
            for p in parts[:-1]:
                node = getattr(node, p)
            setattr(node, parts[-1], val)
    if cfg.commands.heading_command:      # the settings under which the reference's heading mode runs at all (make_golden_env.py; applied last there too)
        cfg.commands.stop_heading_at_limit, cfg.commands.limit_ang_vel_at_zero_command_prob = False, 0.0
    A = EnvArrays(cfg, device, seed=seed, **kw)
    for k in z.files:
        if k.startswith("s0_"):
            A.tensors[k[3:]].copy_(torch.from_numpy(z[k]).to(A.tensors[k[3:]].dtype))
    return z, A


def compare_step(z, i, tensors, keys=None, skip=(), tol=None):
    tol = TOL if tol is None else tol
    bad = []
    names = [k[len(f"out{i}_"):] for k in z.files if k.startswith(f"out{i}_")]
    for name in names:
        if name.startswith("ep_") or name in skip or (keys is not None and name not in keys) or name not in tensors:
            continue
        ref = z[f"out{i}_{name}"]
        got = tensors[name].detach().cpu().numpy().reshape(ref.shape)
        if name in EXACT:
            if not np.array_equal(ref.astype(np.int64), got.astype(np.int64)):
                bad.append((name, "exact", np.argwhere(ref.astype(np.int64) != got.astype(np.int64))[:4].tolist()))
        else:
            rtol, atol = tol.get(name, tol["default"])
            if not np.allclose(got, ref, rtol=rtol, atol=atol, equal_nan=True):      # NaN == NaN: x_command_hip_regular is 0 / 0 at a zero command, in the reference too
                err = np.abs(got - ref)
                bad.append((name, float(err.max()), np.unravel_index(err.argmax(), err.shape)))
    return bad


# configurations other than GO2 training (tests/test_emu_cpu.py, tests/test_gpu_v_env_configs.py): play.py's evaluation set-up and two switch mixes
PLAY = {"terrain.num_rows": 7, "terrain.num_cols": 7, "terrain.curriculum": False, "noise.add_noise": False, "domain_rand.randomize_friction": False,
        "domain_rand.push_robots": False, "domain_rand.randomize_base_mass": False, "domain_rand.randomize_link_mass": False,
        "domain_rand.randomize_base_com": False, "domain_rand.randomize_pd_gains": False, "domain_rand.randomize_motor_zero_offset": False}
ODD = {"domain_rand.randomize_action_delay": False, "domain_rand.randomize_motor_strength": False, "commands.limit_vel_prob": 0.5,
       "commands.limit_vel_invert_when_continuous": False, "commands.limit_ang_vel_at_zero_command_prob": 0.6, "commands.resampling_time": 0.2,
       "terrain.move_down_by_accumulated_xy_command": False, "rewards.dynamic_sigma": None, "rewards.curriculum_rewards": [],
       "commands.dynamic_resample_commands": False, "env.episode_length_s": 2, "normalization.clip_observations": 5.0, "control.action_scale": 0.5}
BARE = {"commands.zero_command_curriculum": None, "commands.limit_vel_prob": 0.0, "domain_rand.push_interval_s": 0.3, "rewards.soft_dof_pos_limit": 0.5,
        "rewards.base_height_target": 0.3, "rewards.tracking_sigma": 0.5, "normalization.clip_actions": 1.0}
# the f-3 switches together (tests/test_emu_cpu.py, tests/test_gpu_v_env_configs.py): turn_over with short episodes (flipped resets every few steps, zero-command timers,
# turn_over reward scales) and every reward function the registered tasks leave off (x_command_hip_regular excepted: 0 / 0 at zero commands)
FLIP = {"init_state.turn_over": True, "init_state.turn_over_proportions": [0.3, 0.3, 0.4], "env.episode_length_s": 0.4, "rewards.turn_over_scales.torques": -2e-4,
        "rewards.turn_over_scales.feet_air_time": 0.5, "rewards.scales.orientation": -0.2, "rewards.scales.base_height": -1.0, "rewards.scales.dof_vel": -1e-4,
        "rewards.scales.termination": -2.0, "rewards.scales.dof_vel_limits": -0.5, "rewards.scales.torque_limits": -0.01, "rewards.scales.feet_air_time": 1.0,
        "rewards.scales.stumble": -0.5, "rewards.scales.stand_still": -0.1, "rewards.scales.feet_contact_forces": -0.01, "rewards.scales.similar_to_default": -0.02,
        "rewards.scales.legs_distance": -1.5, "rewards.soft_dof_vel_limit": 0.1, "rewards.soft_torque_limit": 0.3, "rewards.max_contact_force": 20.0}

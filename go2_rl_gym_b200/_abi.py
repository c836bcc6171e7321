"""ctypes mirror of include/go2_b200.h and the loader of the CUDA library.

The product path has NO CPU fallback: `load_library()` raises if `libgo2b200.so` is missing
(build it with `python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes as C
import os

NUM_DOF, NUM_DYN, NUM_REPORT, NUM_COL = 12, 13, 19, 32
NUM_OBS, NUM_PRIV, NUM_HEIGHT, NUM_REW, NUM_CMD = 45, 263, 187, 14, 4
INERTIA_STRIDE = 10
EP_STATS = NUM_REW + 12
EP_ACC_FIXED_OFF = (EP_STATS + 2 + 1) // 2 * 2       # include/go2_b200.h: GO2_EP_ACC_FIXED_OFF / GO2_EP_ACCUM_FLOATS
EP_ACCUM_FLOATS = EP_ACC_FIXED_OFF + 2 * NUM_REW

REWARD_NAMES = ["tracking_lin_vel", "tracking_ang_vel", "lin_vel_z", "ang_vel_xy", "dof_acc", "dof_power", "torques",
                "correct_base_height", "action_rate", "action_smoothness", "collision", "dof_pos_limits",
                "feet_regulation", "hip_to_default"]

# the reward functions that are inactive in every registered go2 task (include/go2_b200.h: enum Go2XReward; legged_robot.py:1236-1441, go2_env.py:62-68)
XREWARD_NAMES = ["orientation", "base_height", "dof_vel", "termination", "dof_vel_limits", "torque_limits", "feet_air_time", "stumble",
                 "stand_still", "feet_contact_forces", "similar_to_default", "upright", "legs_distance", "x_command_hip_regular"]
NUM_XREW = len(XREWARD_NAMES)
EP_SLOTS = 64
XREW_LOG_BYTES = NUM_XREW * 8 + EP_SLOTS * NUM_XREW * 4

f32, i32, u32 = C.c_float, C.c_int32, C.c_uint32


class Go2Model(C.Structure):
    _fields_ = [("joint_origin", f32 * 3 * NUM_DOF), ("joint_axis", i32 * NUM_DOF),
                ("q_lower", f32 * NUM_DOF), ("q_upper", f32 * NUM_DOF), ("effort", f32 * NUM_DOF),
                ("vel_limit", f32 * NUM_DOF), ("col_pos", f32 * 3 * NUM_COL), ("col_radius", f32 * NUM_COL),
                ("col_dyn", i32 * NUM_COL), ("col_report", i32 * NUM_COL), ("foot_offset", f32 * 3 * 4)]


class Go2EnvConfig(C.Structure):
    _fields_ = [
        ("num_envs", i32), ("env_offset", i32), ("seed_lo", u32), ("seed_hi", u32),
        ("sim_dt", f32), ("decimation", i32), ("gravity_z", f32),
        ("kp", f32 * NUM_DOF), ("kd", f32 * NUM_DOF), ("default_dof_pos", f32 * NUM_DOF),
        ("action_scale", f32), ("clip_actions", f32), ("clip_obs", f32),
        ("randomize_action_delay", i32), ("randomize_motor_strength", i32), ("randomize_motor_zero_offset", i32),
        ("randomize_pd_gains", i32), ("push_robots", i32), ("add_noise", i32),
        ("motor_strength_range", f32 * 2), ("motor_zero_offset_range", f32 * 2), ("kp_mult_range", f32 * 2),
        ("kd_mult_range", f32 * 2),
        ("push_interval", i32), ("max_push_vel_xy", f32), ("max_push_ang_vel", f32),
        ("solver_iters", i32), ("erp", f32), ("limit_erp", f32), ("contact_offset", f32), ("max_depen_vel", f32),
        ("bounce_threshold", f32), ("penetration_slop", f32), ("terrain_friction", f32), ("terrain_restitution", f32),
        ("mesh_type", i32), ("hf_rows", i32), ("hf_cols", i32), ("hscale", f32), ("vscale", f32), ("border", f32),
        ("num_levels", i32), ("num_types", i32), ("terrain_length", f32),
        ("terrain_curriculum", i32), ("move_down_by_accumulated_xy_command", i32), ("custom_origins", i32),
        ("resampling_time", f32), ("dynamic_resample_commands", i32), ("limit_vel_prob", f32),
        ("limit_vel_invert_when_continuous", i32), ("limit_ang_vel_at_zero_command_prob", f32),
        ("max_episode_length", i32), ("max_episode_length_s", f32), ("dt", f32),
        ("reward_scales", f32 * NUM_REW), ("tracking_sigma", f32), ("base_height_target", f32),
        ("soft_dof_limit_lo", f32 * NUM_DOF), ("soft_dof_limit_hi", f32 * NUM_DOF),
        ("dynamic_sigma", i32), ("ds_min_lin", f32), ("ds_max_lin", f32), ("ds_min_ang", f32), ("ds_max_ang", f32),
        ("ds_max_sigma", f32 * 9),
        ("obs_scale_lin_vel", f32), ("obs_scale_ang_vel", f32), ("obs_scale_dof_pos", f32), ("obs_scale_dof_vel", f32),
        ("obs_scale_height", f32), ("noise_scale_vec", f32 * NUM_OBS),
        ("height_points", f32 * 2 * NUM_HEIGHT), ("base_height_mask", f32 * NUM_HEIGHT),
        ("num_base_height_points", f32), ("base_init_state", f32 * 13),
        ("limit_relax", f32), ("contact_relax", f32),
        ("state_guard", i32), ("max_base_lin_vel", f32), ("max_base_ang_vel", f32),
        ("control_type", i32), ("only_positive_rewards", i32),
        ("heading_command", i32), ("stop_heading_at_limit", i32), ("ext_stop_heading_lo", u32), ("ext_stop_heading_hi", u32),
        ("ext_heading_ranges_lo", u32), ("ext_heading_ranges_hi", u32),
        ("num_xrew", i32), ("xrew_scales", f32 * NUM_XREW),
        ("soft_dof_vel_limit", f32), ("soft_torque_limit", f32), ("max_contact_force", f32), ("min_legs_distance", f32),
        ("ext_xrew_sums_lo", u32), ("ext_xrew_sums_hi", u32), ("ext_xrew_state_lo", u32), ("ext_xrew_state_hi", u32),
        ("ext_xrew_log_lo", u32), ("ext_xrew_log_hi", u32),
        ("turn_over", i32), ("turn_over_proportions", f32 * 3), ("turn_over_back_height", f32 * 2), ("turn_over_side_height", f32 * 2),
        ("turn_over_zero_time_back", f32), ("turn_over_zero_time_side", f32), ("turn_over_roll_threshold", f32),
        ("to_scales", f32 * NUM_REW), ("to_xscales", f32 * NUM_XREW),
        ("ext_turn_over_timer_lo", u32), ("ext_turn_over_timer_hi", u32),
    ]


class Go2StepParams(C.Structure):
    _fields_ = [("common_step_counter", u32), ("reward_curriculum", f32 * NUM_REW), ("zero_command_proba", f32),
                ("max_lin_vel", f32), ("ep_slot", i32), ("xrew_curriculum", f32 * NUM_XREW)]


_PTR_FIELDS = [
    "root_states", "dof_pos", "dof_vel", "torques", "contact_forces", "feet_pos", "feet_vel",
    "actions", "last_actions", "last_last_actions", "last_dof_vel", "obs_buf", "privileged_obs_buf", "rew_buf",
    "reset_buf", "time_out_buf", "episode_length_buf",
    "base_lin_vel", "base_ang_vel", "projected_gravity", "measured_heights",
    "commands", "commands_resampling_step", "commands_xy_accumulation", "last_is_limit_vel", "env_command_ranges",
    "terrain_levels", "terrain_types", "terrain_ids", "env_origins", "max_move_distance", "terrain_origins",
    "height_samples",
    "motor_strengths", "motor_zero_offsets", "p_gains_multiplier", "d_gains_multiplier", "friction_coeffs",
    "restitutions", "body_inertia", "episode_sums", "ep_stats", "ep_accum",
]


class Go2EnvBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _PTR_FIELDS]


PTR_FIELDS = tuple(_PTR_FIELDS)

_LIB = None
_LIB_PATH = os.environ.get("GO2_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgo2b200.so")     # GO2_B200_LIB: tuning builds of the same library


def library_path():
    return _LIB_PATH


_VARIANTS = {}


def load_library(path=None):
    """Load libgo2b200.so (hand-written sm_100a kernels + C ABI). Raises if it has not been built.
    path: another build of the same library (e.g. a tuning build), loaded side by side with its own handle."""
    global _LIB
    if path is None and _LIB is not None:
        return _LIB
    if path is not None and path in _VARIANTS:
        return _VARIANTS[path]
    lib_path = path or _LIB_PATH
    if not os.path.exists(lib_path):
        raise RuntimeError(
            f"{lib_path} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with: python -c 'import __graft_entry__ as g; g.build()'")
    lib = C.CDLL(lib_path)
    vp = C.c_void_p
    lib.go2_env_create.argtypes = [C.POINTER(Go2EnvConfig), C.POINTER(Go2Model), C.POINTER(Go2EnvBuffers), C.POINTER(vp)]
    lib.go2_env_create.restype = C.c_int
    lib.go2_env_destroy.argtypes = [vp]
    lib.go2_env_destroy.restype = None
    lib.go2_env_step.argtypes = [vp, vp, C.POINTER(Go2StepParams), vp]
    lib.go2_env_step.restype = C.c_int
    lib.go2_env_set_step_mode.argtypes = [vp, C.c_char_p]
    lib.go2_env_set_step_mode.restype = C.c_int
    lib.go2_env_step_dev.argtypes = [vp, vp, vp, vp]
    lib.go2_env_step_dev.restype = C.c_int
    lib.go2_env_step_host.argtypes = [vp, vp, C.POINTER(Go2StepParams), vp, vp, vp, vp, vp]
    lib.go2_env_step_host.restype = C.c_int
    lib.go2_env_step_host_begin.argtypes = [vp, vp, C.POINTER(Go2StepParams), vp, vp, vp, vp, vp]
    lib.go2_env_step_host_begin.restype = C.c_int
    lib.go2_env_step_host_end.argtypes = [vp]
    lib.go2_env_step_host_end.restype = C.c_int
    lib.go2_env_reset_all.argtypes = [vp, C.POINTER(Go2StepParams), vp]
    lib.go2_env_reset_all.restype = C.c_int
    lib.go2_env_substeps.argtypes = [vp, vp, C.c_int, vp]
    lib.go2_env_substeps.restype = C.c_int
    lib.go2_last_error.restype = C.c_char_p
    lib.go2_kernel_launch_count.restype = C.c_longlong
    if path is None:
        _LIB = lib
    else:
        _VARIANTS[path] = lib
    return lib


def check(rc, lib=None):
    if rc != 0:
        lib = lib or load_library()
        raise RuntimeError(f"libgo2b200 error {rc}: {lib.go2_last_error().decode()}")

"""LeggedRobotCfg and the train configs (public API).

Same attribute names, nesting and default values as the reference
(legged_gym/envs/base/legged_robot_config.py:4-409) — these classes ARE the drop-in boundary, user code
subclasses and mutates them.  Values the B200 kernels consume are packed into `Go2EnvConfig`
(include/go2_b200.h) by go2_rl_gym_b200/envs/env_arrays.py.  `sim.physx` is kept for API compatibility;
of its fields only contact_offset, bounce_threshold_velocity and max_depenetration_velocity have a meaning
for the B200 physics; `sim.b200` holds the solver's own knobs.
"""
import math
from .base_config import BaseConfig

_PI2 = 1.57


def _cmd(x, y, yaw):
    return {'lin_vel_x': [-x, x], 'lin_vel_y': [-y, y], 'ang_vel_yaw': [-yaw, yaw], 'heading': [-_PI2, _PI2]}


class LeggedRobotCfg(BaseConfig):
    class env:
        num_envs = 4096
        num_observations = 48
        num_privileged_obs = None
        num_actions = 12
        env_spacing = 3.
        send_timeouts = True
        episode_length_s = 20
        test = False

    class terrain:
        mesh_type = 'trimesh'
        horizontal_scale = 0.1
        vertical_scale = 0.005
        border_size = 25
        curriculum = True
        static_friction = 1.0
        dynamic_friction = 1.0
        restitution = 0.
        measure_heights = True
        measured_points_x = [round(-0.8 + 0.1 * i, 1) for i in range(17)]
        measured_points_y = [round(-0.5 + 0.1 * i, 1) for i in range(11)]
        selected = False
        terrain_kwargs = None
        max_init_terrain_level = 5
        terrain_length = 8.
        terrain_width = 8.
        num_rows = 10
        num_cols = 20
        terrain_spacing = 0.5
        terrain_proportions = [0.1, 0.1, 0.1, 0.2, 0.2, 0.1, 0.1, 0.1, 0.0]
        slope_treshold = 0.75
        move_down_by_accumulated_xy_command = False

    class commands:
        curriculum = False
        max_curriculum = 1.
        num_commands = 4
        resampling_time = 10.
        heading_command = False
        zero_command_curriculum = None
        limit_ang_vel_at_zero_command_prob = 0.0
        limit_vel_prob = 0.0
        limit_vel_invert_when_continuous = True
        limit_vel = {"lin_vel_x": [-1, 1], "lin_vel_y": [-1, 1], "ang_vel_yaw": [-1, 0, 1]}
        stop_heading_at_limit = True
        dynamic_resample_commands = False
        command_range_curriculum = []
        turn_over_zero_time = {"backflip": 5.0, "sideflip": 3.0}
        # wave, slope, rough slope, stairs up, stairs down, obstacles, stepping stones, gap, flat
        terrain_max_command_ranges = [_cmd(1.5, 1.5, 1.5)] * 3 + [_cmd(1.0, 1.0, 1.5)] * 5 + [_cmd(2.0, 1.5, 1.5)]

        class ranges:
            lin_vel_x = [-1.0, 1.0]
            lin_vel_y = [-0.5, 0.5]
            ang_vel_yaw = [-1, 1]
            heading = [-3.14, 3.14]

    class init_state:
        pos = [0.0, 0.0, 1.]
        rot = [0.0, 0.0, 0.0, 1.0]
        lin_vel = [0.0, 0.0, 0.0]
        ang_vel = [0.0, 0.0, 0.0]
        default_joint_angles = {"joint_a": 0., "joint_b": 0.}
        turn_over = False
        turn_over_proportions = [0.0, 0.2, 0.8]
        turn_over_init_heights = {'backflip': [0.10, 0.15], 'sideflip': [0.16, 0.21]}

    class control:
        control_type = 'P'
        stiffness = {'joint_a': 10.0, 'joint_b': 15.}
        damping = {'joint_a': 1.0, 'joint_b': 1.5}
        action_scale = 0.5
        decimation = 4

    class asset:
        file = ""
        name = "legged_robot"
        foot_name = "None"
        penalize_contacts_on = []
        terminate_after_contacts_on = []
        disable_gravity = False
        collapse_fixed_joints = True
        fix_base_link = False
        default_dof_drive_mode = 3
        self_collisions = 0
        replace_cylinder_with_capsule = True
        flip_visual_attachments = True
        density = 0.001
        angular_damping = 0.
        linear_damping = 0.
        max_angular_velocity = 1000.
        max_linear_velocity = 1000.
        armature = 0.
        thickness = 0.01

    class domain_rand:
        robot_properties_update = None
        randomize_friction = True
        friction_range = [0.2, 1.25]
        randomize_base_mass = True
        added_mass_range = [-1., 1.]
        randomize_link_mass = True
        multiplied_link_mass_range = [0.9, 1.1]
        randomize_base_com = True
        added_base_com_range = [-0.03, 0.03]
        randomize_restitution = False
        restitution_range = [0.0, 0.2]
        randomize_pd_gains = True
        stiffness_multiplier_range = [0.9, 1.1]
        damping_multiplier_range = [0.9, 1.1]
        randomize_motor_zero_offset = True
        motor_zero_offset_range = [-0.035, 0.035]
        randomize_motor_strength = False
        motor_strength_range = [0.8, 1.2]
        push_robots = True
        push_interval_s = 4
        max_push_vel_xy = 0.4
        max_push_ang_vel = 0.6
        randomize_action_delay = False

    class rewards:
        class scales:
            termination = -0.0
            tracking_lin_vel = 1.0
            tracking_ang_vel = 0.5
            lin_vel_z = -2.0
            ang_vel_xy = -0.05
            orientation = -0.
            torques = -0.00001
            dof_vel = -0.
            dof_acc = -2.5e-7
            base_height = -0.
            feet_air_time = 1.0
            collision = -1.
            feet_stumble = -0.0
            action_rate = -0.01
            stand_still = -0.

        class turn_over_scales:
            upright = 1.0

        only_positive_rewards = True
        tracking_sigma = 0.25
        soft_dof_pos_limit = 1.
        soft_dof_vel_limit = 1.
        soft_torque_limit = 1.
        base_height_target = 1.
        max_contact_force = 100.
        curriculum_rewards = None
        dynamic_sigma = None
        turn_over_roll_threshold = math.pi / 4
        min_legs_distance = 0.1

    class normalization:
        class obs_scales:
            lin_vel = 2.0
            ang_vel = 0.25
            dof_pos = 1.0
            dof_vel = 0.05
            height_measurements = 2.5
        clip_observations = 100.
        clip_actions = 100.

    class noise:
        add_noise = True
        noise_level = 1.0

        class noise_scales:
            dof_pos = 0.01
            dof_vel = 1.5
            lin_vel = 0.1
            ang_vel = 0.2
            gravity = 0.05
            height_measurements = 0.1

    class viewer:
        ref_env = 0
        pos = [10, 0, 6]
        lookat = [11., 5, 3.]

    class sim:
        dt = 0.005
        substeps = 1
        gravity = [0., 0., -9.81]
        up_axis = 1

        class physx:
            num_threads = 10
            solver_type = 1
            num_position_iterations = 4
            num_velocity_iterations = 0
            contact_offset = 0.01
            rest_offset = 0.0
            bounce_threshold_velocity = 0.5
            max_depenetration_velocity = 1.0
            max_gpu_contact_pairs = 2**23
            default_buffer_size_multiplier = 5
            contact_collection = 2

        class b200:
            """Knobs of the B200 contact / joint-limit impulse solver (no reference counterpart)."""
            solver_iterations = 4
            erp = 0.2
            penetration_slop = 0.004
            # relaxation of the Jacobi sweeps (include/go2_b200.h): joint-limit rows step with limit_relax / (M^-1)_jj,
            # contact blocks with contact_relax / (active contacts of the group); the sweeps converge with these values (0 / 1 / limit_erp 0.2
            # selects round 1's first solver, whose limit rows are over-relaxed: soft stops, divergent beyond 4 sweeps)
            limit_relax = 0.5
            contact_relax = 0.7
            limit_erp = 0.8
            # clamp the base twist at asset.max_linear_velocity / max_angular_velocity and reset any env whose state went non-finite
            state_guard = 1


class _TrainCfgBase(BaseConfig):
    class runner:
        num_steps_per_env = 24
        max_iterations = 1500
        save_interval = 50
        experiment_name = 'test'
        run_name = ''
        resume = False
        load_run = -1
        checkpoint = -1
        resume_path = None

    class robogauge:
        enabled = False
        port = 9973


class LeggedRobotCfgPPO(_TrainCfgBase):
    seed = 1
    runner_class_name = 'OnPolicyRunner'

    class policy:
        init_noise_std = 1.0
        actor_hidden_dims = [512, 256, 128]
        critic_hidden_dims = [512, 256, 128]
        activation = 'elu'

    class algorithm:
        value_loss_coef = 1.0
        use_clipped_value_loss = True
        clip_param = 0.2
        entropy_coef = 0.01
        num_learning_epochs = 5
        num_mini_batches = 4
        learning_rate = 1.e-3
        schedule = 'adaptive'
        gamma = 0.99
        lam = 0.95
        desired_kl = 0.01
        max_grad_norm = 1.

    class runner(_TrainCfgBase.runner):
        policy_class_name = 'ActorCritic'
        algorithm_class_name = 'PPO'


class LeggedRobotCfgCTS(_TrainCfgBase):
    seed = 0
    runner_class_name = "OnPolicyRunnerCTS"
    history_length = 5

    class policy(LeggedRobotCfgPPO.policy):
        teacher_encoder_hidden_dims = [512, 256]
        student_encoder_hidden_dims = [512, 256]
        latent_dim = 32
        norm_type = 'l2norm'

    class algorithm(LeggedRobotCfgPPO.algorithm):
        student_encoder_learning_rate = 1e-3
        teacher_env_ratio = 0.75

    class runner(_TrainCfgBase.runner):
        policy_class_name = 'ActorCriticCTS'
        algorithm_class_name = 'CTS'


def _variant(policy_name, alg_name, policy_extra, alg_extra=None):
    """CTS ablation variants differ only in a few policy/algorithm keys and the class names."""
    ns_policy = type('policy', (LeggedRobotCfgCTS.policy,), dict(policy_extra))
    ns_alg = type('algorithm', (LeggedRobotCfgCTS.algorithm,), dict(alg_extra or {}))
    ns_runner = type('runner', (LeggedRobotCfgCTS.runner,),
                     {'policy_class_name': policy_name, 'algorithm_class_name': alg_name})
    return {'policy': ns_policy, 'algorithm': ns_alg, 'runner': ns_runner}


LeggedRobotCfgMoENGCTS = type('LeggedRobotCfgMoENGCTS', (LeggedRobotCfgCTS,), _variant(
    'ActorCriticMoENGCTS', 'MoENGCTS', {'obs_no_goal_mask': None, 'student_expert_num': 8}, {'load_balance_coef': 0.01}))
LeggedRobotCfgMCPCTS = type('LeggedRobotCfgMCPCTS', (LeggedRobotCfgCTS,), _variant(
    'ActorCriticMCPCTS', 'MCPCTS', {'obs_no_goal_mask': None, 'student_expert_num': 8}))
LeggedRobotCfgACMoECTS = type('LeggedRobotCfgACMoECTS', (LeggedRobotCfgCTS,), _variant(
    'ActorCriticACMoECTS', 'ACMoECTS', {'expert_num': 8}))
LeggedRobotCfgDualMoECTS = type('LeggedRobotCfgDualMoECTS', (LeggedRobotCfgCTS,), _variant(
    'ActorCriticDualMoECTS', 'DualMoECTS', {'expert_num': 8, 'student_encoder_hidden_dims': [512, 256, 256]}))
LeggedRobotCfgMoECTS = type('LeggedRobotCfgMoECTS', (LeggedRobotCfgCTS,), _variant(
    'ActorCriticMoECTS', 'MoECTS', {'expert_num': 8, 'student_encoder_hidden_dims': [512, 256, 256]},
    {'load_balance_coef': 0.01}))

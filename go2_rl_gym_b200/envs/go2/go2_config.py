"""Go2 environment + train configs (public API; values from legged_gym/envs/go2/go2_config.py:4-284)."""
import math
from ..base.legged_robot_config import (LeggedRobotCfg, LeggedRobotCfgPPO, LeggedRobotCfgCTS, LeggedRobotCfgMoECTS,
                                        LeggedRobotCfgMoENGCTS, LeggedRobotCfgMCPCTS, LeggedRobotCfgACMoECTS,
                                        LeggedRobotCfgDualMoECTS, _cmd)

_LEGS = ('FL', 'FR', 'RL', 'RR')


class GO2Cfg(LeggedRobotCfg):
    class init_state(LeggedRobotCfg.init_state):
        pos = [0.0, 0.0, 0.42]
        default_joint_angles = {**{f'{l}_hip_joint': (0.1 if l[1] == 'L' else -0.1) for l in _LEGS},
                                **{f'{l}_thigh_joint': (0.8 if l[0] == 'F' else 1.0) for l in _LEGS},
                                **{f'{l}_calf_joint': -1.5 for l in _LEGS}}
        turn_over = False
        turn_over_proportions = [0.0, 0.2, 0.8]
        turn_over_init_heights = {'backflip': [0.10, 0.15], 'sideflip': [0.16, 0.21]}

    class env(LeggedRobotCfg.env):
        num_envs = 8192
        num_observations = 45
        num_privileged_obs = 45 + 3 + 4 + 12 + 12 + 187
        episode_length_s = 25

    class domain_rand(LeggedRobotCfg.domain_rand):
        friction_range = [0.0, 2.0]
        randomize_restitution = True
        restitution_range = [0.0, 0.5]
        randomize_motor_strength = True
        randomize_action_delay = True

    class control(LeggedRobotCfg.control):
        control_type = 'P'
        stiffness = {'joint': 20.0}
        damping = {'joint': 0.5}
        action_scale = 0.25
        decimation = 4

    class terrain(LeggedRobotCfg.terrain):
        max_init_terrain_level = 5
        # wave, slope, rough_slope, stairs up, stairs down, obstacles, stepping_stones, gap, flat
        terrain_proportions = [0.05, 0.20, 0.05, 0.25, 0.10, 0.20, 0.0, 0.0, 0.15]
        move_down_by_accumulated_xy_command = True

    class commands(LeggedRobotCfg.commands):
        resampling_time = 5.
        zero_command_curriculum = {'start_iter': 0, 'end_iter': 1500, 'start_value': 0.0, 'end_value': 0.1}
        limit_ang_vel_at_zero_command_prob = 0.2
        limit_vel_prob = 0.2
        dynamic_resample_commands = True
        command_range_curriculum = [
            {'iter': 20000, 'lin_vel_x': [-1.0, 1.0], 'lin_vel_y': [-1.0, 1.0], 'ang_vel_yaw': [-1.5, 1.5], 'heading': [-1.57, 1.57]},
            {'iter': 50000, 'lin_vel_x': [-2.0, 2.0], 'lin_vel_y': [-1.0, 1.0], 'ang_vel_yaw': [-2.0, 2.0], 'heading': [-1.57, 1.57]},
        ]
        terrain_max_command_ranges = [_cmd(1.5, 1.0, 1.5)] * 3 + [_cmd(1.0, 1.0, 1.5)] * 5 + [_cmd(2.0, 1.0, 2.0)]

        class ranges:
            lin_vel_x = [-0.5, 0.5]
            lin_vel_y = [-0.5, 0.5]
            ang_vel_yaw = [-1.0, 1.0]
            heading = [-1.57, 1.57]

    class asset(LeggedRobotCfg.asset):
        file = '{LEGGED_GYM_ROOT_DIR}/resources/robots/go2/urdf/go2.urdf'
        name = "go2"
        foot_name = "foot"
        penalize_contacts_on = ["thigh", "calf"]
        terminate_after_contacts_on = ["base"]
        self_collisions = 1

    class rewards(LeggedRobotCfg.rewards):
        soft_dof_pos_limit = 0.9
        base_height_target = 0.38
        only_positive_rewards = False
        max_contact_force = 147.
        curriculum_rewards = [
            {'reward_name': 'lin_vel_z', 'start_iter': 0, 'end_iter': 1500, 'start_value': 1.0, 'end_value': 0.0},
            {'reward_name': 'correct_base_height', 'start_iter': 0, 'end_iter': 5000, 'start_value': 1.0, 'end_value': 10.0},
        ]
        tracking_sigma = 0.25
        dynamic_sigma = {"min_lin_vel": 0.5, "max_lin_vel": 1.5, "min_ang_vel": 1.0, "max_ang_vel": 2.0,
                         "max_sigma": [5/12, 1/4, 1/4, 1/2, 1/2, 3/4, 1, 1, 1/4]}
        min_legs_distance = 0.1

        class scales:
            tracking_lin_vel = 1.0
            tracking_ang_vel = 0.5
            lin_vel_z = -2.0
            ang_vel_xy = -0.05
            dof_acc = -2.5e-7
            dof_power = -2e-5
            torques = -1e-4
            correct_base_height = -1.0
            action_rate = -0.01
            action_smoothness = -0.01
            collision = -1.0
            dof_pos_limits = -2.0
            feet_regulation = -0.05
            hip_to_default = -0.05

        turn_over_roll_threshold = math.pi / 4

        class turn_over_scales:
            upright = 1.0

    class noise(LeggedRobotCfg.noise):
        add_noise = True


def _train(base, experiment, policy_extra=None, alg_extra=None, runner_extra=None):
    ns = {'runner': type('runner', (base.runner,), {'run_name': '', 'experiment_name': experiment,
                                                    'max_iterations': 150000, 'save_interval': 500,
                                                    **(runner_extra or {})})}
    if policy_extra:
        ns['policy'] = type('policy', (base.policy,), dict(policy_extra))
    if alg_extra:
        ns['algorithm'] = type('algorithm', (base.algorithm,), dict(alg_extra))
    return ns


_NO_GOAL = [True] * 6 + [False] * 3 + [True] * 36

GO2CfgPPO = type('GO2CfgPPO', (LeggedRobotCfgPPO,), _train(LeggedRobotCfgPPO, 'go2_ppo', alg_extra={'entropy_coef': 0.01}))
GO2CfgCTS = type('GO2CfgCTS', (LeggedRobotCfgCTS,), _train(LeggedRobotCfgCTS, 'go2_cts', {'latent_dim': 32, 'norm_type': 'l2norm'},
                                                             runner_extra={'num_steps_per_env': 24}))
GO2CfgMoECTS = type('GO2CfgMoECTS', (LeggedRobotCfgMoECTS,), _train(LeggedRobotCfgMoECTS, 'go2_moe_cts', {'expert_num': 8}))
GO2CfgMoENGCTS = type('GO2CfgMoENGCTS', (LeggedRobotCfgMoENGCTS,), _train(
    LeggedRobotCfgMoENGCTS, 'go2_moe_no_goal_cts', {'obs_no_goal_mask': _NO_GOAL, 'student_expert_num': 8}, {'load_balance_coef': 0.01}))
GO2CfgMCPCTS = type('GO2CfgMCPCTS', (LeggedRobotCfgMCPCTS,), _train(
    LeggedRobotCfgMCPCTS, 'go2_mcp_cts', {'obs_no_goal_mask': _NO_GOAL, 'student_expert_num': 8}))
GO2CfgACMoECTS = type('GO2CfgACMoECTS', (LeggedRobotCfgACMoECTS,), _train(LeggedRobotCfgACMoECTS, 'go2_ac_moe_cts', {'expert_num': 8}))
GO2CfgDualMoECTS = type('GO2CfgDualMoECTS', (LeggedRobotCfgDualMoECTS,), _train(LeggedRobotCfgDualMoECTS, 'go2_dual_moe_cts', {'expert_num': 8}))

"""Hyper-parameters of the MoE-CTS golden fixture (LeggedRobotCfgMoECTS / CTS, legged_robot_config.py:313-339,399-409), small widths."""
POLICY = dict(actor_hidden_dims=[64, 32, 16], critic_hidden_dims=[64, 32, 16], teacher_encoder_hidden_dims=[64, 32],
              student_encoder_hidden_dims=[64, 32, 32], expert_num=8, activation="elu", init_noise_std=1.0, latent_dim=32, norm_type="l2norm")
ALG = dict(value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.01, num_learning_epochs=5, num_mini_batches=4,
           learning_rate=1e-3, student_encoder_learning_rate=1e-3, schedule="adaptive", gamma=0.99, lam=0.95, desired_kl=0.01, max_grad_norm=1.0,
           teacher_env_ratio=0.75, load_balance_coef=0.01)
# plain CTS (LeggedRobotCfgCTS, legged_robot_config.py:289-339): MLP student encoder, no load-balance term
POLICY_CTS = dict(actor_hidden_dims=[64, 32, 16], critic_hidden_dims=[64, 32, 16], teacher_encoder_hidden_dims=[64, 32],
                  student_encoder_hidden_dims=[64, 32], activation="elu", init_noise_std=1.0, latent_dim=32, norm_type="l2norm")
ALG_CTS = {k: v for k, v in ALG.items() if k != "load_balance_coef"}
# MoE-CTS with no-goal experts (LeggedRobotCfgMoENGCTS, legged_robot_config.py:361-371; GO2CfgMoENGCTS go2_config.py:231-243)
NO_GOAL_MASK = [True] * 6 + [False] * 3 + [True] * 36
POLICY_NG = dict(obs_no_goal_mask=NO_GOAL_MASK, actor_hidden_dims=[64, 32, 16], critic_hidden_dims=[64, 32, 16], teacher_encoder_hidden_dims=[64, 32],
                 student_encoder_hidden_dims=[64, 32], student_expert_num=8, activation="elu", init_noise_std=1.0, latent_dim=32, norm_type="l2norm")
# MoE actor + gated value experts (LeggedRobotCfgACMoECTS / DualMoECTS, legged_robot_config.py:382-397): MLP student / MoE student
POLICY_AC = dict(actor_hidden_dims=[64, 32, 16], critic_hidden_dims=[64, 32, 16], teacher_encoder_hidden_dims=[64, 32],
                 student_encoder_hidden_dims=[64, 32], expert_num=8, activation="elu", init_noise_std=1.0, latent_dim=32, norm_type="l2norm")
POLICY_DUAL = dict(POLICY_AC, student_encoder_hidden_dims=[64, 32, 32])
# multiplicative compositional actor (LeggedRobotCfgMCPCTS, legged_robot_config.py:373-380; GO2CfgMCPCTS go2_config.py:245-254)
POLICY_MCP = dict(obs_no_goal_mask=NO_GOAL_MASK, actor_hidden_dims=[64, 32], critic_hidden_dims=[64, 32, 16], teacher_encoder_hidden_dims=[64, 32],
                  student_encoder_hidden_dims=[64, 32], student_expert_num=8, activation="elu", latent_dim=32, norm_type="l2norm")

"""PPO hyper-parameters of the RL golden fixture (LeggedRobotCfgPPO.algorithm, legged_robot_config.py:274-287)."""
CFG = dict(value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.01, num_learning_epochs=5, num_mini_batches=4,
           learning_rate=1e-3, schedule="adaptive", gamma=0.99, lam=0.95, desired_kl=0.01, max_grad_norm=1.0)

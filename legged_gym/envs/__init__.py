from go2_rl_gym_b200.envs import *  # noqa: F401,F403  (registers go2, go2_cts, go2_moe_cts)
from go2_rl_gym_b200.envs import task_registry, Go2Robot, LeggedRobot, GO2Cfg, GO2CfgPPO  # noqa: F401

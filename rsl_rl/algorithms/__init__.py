from go2_rl_gym_b200.rl.algorithms import *  # noqa: F401,F403
from go2_rl_gym_b200.rl.algorithms import PPO  # noqa: F401

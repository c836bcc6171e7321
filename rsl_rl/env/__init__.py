from go2_rl_gym_b200.rl.env import *  # noqa: F401,F403
from go2_rl_gym_b200.rl.env import VecEnv  # noqa: F401

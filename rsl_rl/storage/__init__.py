from go2_rl_gym_b200.rl.storage import *  # noqa: F401,F403
from go2_rl_gym_b200.rl.storage import RolloutStorage  # noqa: F401

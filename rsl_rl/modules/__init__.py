from go2_rl_gym_b200.rl.modules import *  # noqa: F401,F403
from go2_rl_gym_b200.rl.modules import ActorCritic  # noqa: F401

from go2_rl_gym_b200.rl.runners import *  # noqa: F401,F403
from go2_rl_gym_b200.rl.runners import OnPolicyRunner  # noqa: F401
from go2_rl_gym_b200.rl.runners import OnPolicyRunnerCTS  # noqa: F401

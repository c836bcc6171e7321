"""`python legged_gym/scripts/train.py --task=go2 [--num_envs N --headless --max_iterations K]`
(same flow as the reference's script, legged_gym/scripts/train.py:11-19)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from legged_gym.envs import *  # noqa: E402,F401,F403
from legged_gym.utils import get_args, task_registry  # noqa: E402


def train(args):
    env, env_cfg = task_registry.make_env(name=args.task, args=args)
    ppo_runner, train_cfg = task_registry.make_alg_runner(env=env, name=args.task, args=args)
    env.common_step_counter = ppo_runner.current_learning_iteration * env.num_steps_per_env
    env.update_reward_curriculum(force_update=True)
    ppo_runner.learn(num_learning_iterations=train_cfg.runner.max_iterations, init_at_random_ep_len=True)


if __name__ == '__main__':
    train(get_args())

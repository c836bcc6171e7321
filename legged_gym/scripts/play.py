"""`python legged_gym/scripts/play.py --task=go2 [--load_run R --checkpoint K --num_steps S]`: headless evaluation of a trained policy
(same flow as the reference's script, legged_gym/scripts/play.py:15-66: 7 x 7 non-curriculum terrain, noise / pushes / most
randomisation off, deterministic `act_inference`, policy export).  There is no viewer; the loop prints tracking statistics instead
(mean terrain level, velocity-tracking error, mean reward), which serve as this package's own regression metric."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch  # noqa: E402

from legged_gym.envs import *  # noqa: E402,F401,F403
from legged_gym.utils import get_args, task_registry  # noqa: E402
from legged_gym.utils.exporter import export_policy_as_jit, export_policy_as_pkl  # noqa: E402

EXPORT_POLICY = True
FIX_COMMAND = True


def play(args, num_steps=None, runner=None, export_dir=None):
    env_cfg, train_cfg = task_registry.get_cfgs(name=args.task)
    # override some parameters for testing (play.py:18-32)
    env_cfg.env.num_envs = min(env_cfg.env.num_envs, 100) if getattr(args, "num_envs", None) is None else args.num_envs
    env_cfg.terrain.num_rows = 7
    env_cfg.terrain.num_cols = 7
    env_cfg.terrain.curriculum = False
    env_cfg.noise.add_noise = False
    env_cfg.domain_rand.randomize_friction = False
    env_cfg.domain_rand.push_robots = False
    env_cfg.domain_rand.randomize_base_mass = False
    env_cfg.domain_rand.randomize_link_mass = False
    env_cfg.domain_rand.randomize_base_com = False
    env_cfg.domain_rand.randomize_pd_gains = False
    env_cfg.domain_rand.randomize_motor_zero_offset = False
    env_cfg.env.test = True
    args.num_envs = env_cfg.env.num_envs
    env, _ = task_registry.make_env(name=args.task, args=args, env_cfg=env_cfg)
    obs = env.get_observations()
    if runner is None:      # load the policy from the latest (or the selected) run, like the reference
        train_cfg.runner.resume = True
        runner, train_cfg = task_registry.make_alg_runner(env=env, name=args.task, args=args, train_cfg=train_cfg)
    else:                   # evaluate the policy of a live runner (tests): same weights, this env
        src = runner.alg.actor_critic if hasattr(runner.alg, "actor_critic") else runner.alg.model
        runner, train_cfg = task_registry.make_alg_runner(env=env, name=args.task, args=args, train_cfg=train_cfg, log_root=None)
        (runner.alg.actor_critic if hasattr(runner.alg, "actor_critic") else runner.alg.model).load_state_dict(src.state_dict())
    model = runner.alg.actor_critic if hasattr(runner.alg, "actor_critic") else runner.alg.model
    policy = runner.get_inference_policy(device=env.device)
    if EXPORT_POLICY:
        from go2_rl_gym_b200.utils.task_registry import LEGGED_GYM_ROOT_DIR
        path = export_dir or os.path.join(LEGGED_GYM_ROOT_DIR, 'logs', train_cfg.runner.experiment_name, 'exported', 'policies')
        export_policy_as_jit(model, path)
        export_policy_as_pkl(model, path)
        print('Exported policy as jit script / pkl to: ', path)
    n = 10 * int(env.max_episode_length) if num_steps is None else int(num_steps)
    rew_sum = torch.zeros((), device=env.device)
    err_sum = torch.zeros((), device=env.device)
    with torch.inference_mode():
        for i in range(n):
            actions = policy(obs.detach())
            if FIX_COMMAND:
                env.commands[:, 0] = 1.0
                env.commands[:, 1] = 0.0
                env.commands[:, 2] = 0.0
            obs, _, rews, dones, infos = env.step(actions.detach())
            rew_sum += rews.mean()
            err_sum += (env.commands[:, :2] - env.base_lin_vel[:, :2]).norm(dim=1).mean()
    stats = {"steps": n, "mean_reward": float(rew_sum) / max(n, 1), "mean_lin_vel_tracking_error": float(err_sum) / max(n, 1),
             "mean_terrain_level": float(env.terrain_levels.float().mean())}
    print(stats)
    return stats


if __name__ == '__main__':
    import argparse
    extra = argparse.ArgumentParser(add_help=False)
    extra.add_argument("--num_steps", type=int, default=None)
    ns, rest = extra.parse_known_args()
    play(get_args(rest), num_steps=ns.num_steps)

#!/usr/bin/env python3
"""Learning curve of the B200 path: `--task=go2` (or any registered task) trained through task_registry.make_env / make_alg_runner exactly as
legged_gym/scripts/train.py does, with the per-iteration statistics the reference logs (rsl_rl/runners/on_policy_runner.py:203-207:
Train/mean_reward, Train/mean_episode_length over the last 100 finished episodes) plus the mean terrain level, printed as one row per iteration in
the format of tests/tools/train_cpu_curve.py (the CPU oracle env under the UNMODIFIED reference PPO: profiles/r01n_cpu_oracle_learning_curve*.txt),
so the two curves can be laid side by side.

  python tools/train_gpu_curve.py --task go2 --num_envs 4096 --iterations 300 > profiles/r02_gpu_learning_curve_go2.txt
"""
import argparse
import contextlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def train(task="go2", num_envs=4096, iterations=300, seed=None, report=print):
    """-> list of per-iteration dicts"""
    import torch
    from go2_rl_gym_b200.envs import task_registry
    from go2_rl_gym_b200.utils import get_args
    argv = ["--task", task, "--num_envs", str(num_envs), "--headless"] + (["--seed", str(seed)] if seed is not None else [])
    args = get_args(argv)
    env_cfg, _ = task_registry.get_cfgs(task)
    env_cfg.terrain.mesh_type = "heightfield"
    with contextlib.redirect_stdout(sys.stderr):
        env, env_cfg = task_registry.make_env(task, args, env_cfg)
        runner, train_cfg = task_registry.make_alg_runner(env, task, args, log_root=None)
    alg = runner.alg
    is_cts = hasattr(runner, "history")
    # train.py:14-16
    env.common_step_counter = runner.current_learning_iteration * env.num_steps_per_env
    env.update_reward_curriculum(force_update=True)
    env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length))     # learn(init_at_random_ep_len=True)
    if is_cts:
        runner._roll_history(env.get_observations(), None)
        runner._hist_primed = True
    model = alg.model if is_cts else alg.actor_critic
    model.train()
    report(f"# {task} rough-terrain heightfield, {num_envs} envs, B200 path (fused step kernel + CUDA trainer) through task_registry; columns:")
    report("# iter  mean_reward/step  mean_episode_return(last 100)  mean_episode_length(last 100)  mean_terrain_level  action_std  lr  ms/iter")
    done_rew, done_len, rows = [], [], []
    for it in range(iterations):
        torch.cuda.synchronize()
        t0 = time.time()
        runner.collect(log=True)
        with torch.inference_mode():
            if is_cts:
                runner._compute_returns(env.get_observations(), env.get_privileged_observations())
            else:
                alg.compute_returns(env.get_privileged_observations())
        alg.update()
        torch.cuda.synchronize()
        ms = 1e3 * (time.time() - t0)
        dr, dl = runner._done_rew.flatten(), runner._done_len.flatten()        # time-major, like the reference's per-step rewbuffer.extend
        keep = ~torch.isnan(dr)
        done_rew = (done_rew + dr[keep].cpu().tolist())[-100:]
        done_len = (done_len + dl[keep].cpu().tolist())[-100:]
        std = getattr(model, "std", None)
        row = {"iter": it, "rew_step": float(env.rew_buf.mean()) if False else float(alg.storage.rewards.mean()),
               "ret": sum(done_rew) / max(len(done_rew), 1), "len": sum(done_len) / max(len(done_len), 1),
               "level": float(env.terrain_levels.float().mean()), "std": float(std.mean()) if std is not None else float("nan"),
               "lr": alg.learning_rate, "ms": ms}
        rows.append(row)
        report(f"{it:4d}  {row['rew_step']:9.5f}  {row['ret']:9.3f}  {row['len']:8.1f}  {row['level']:6.3f}  {row['std']:6.3f}  {row['lr']:.2e}  {ms:7.2f}")
    return rows


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="go2")
    ap.add_argument("--num_envs", type=int, default=4096)
    ap.add_argument("--iterations", type=int, default=300)
    ap.add_argument("--seed", type=int, default=None)
    a = ap.parse_args()
    train(a.task, a.num_envs, a.iterations, a.seed, report=lambda s: print(s, flush=True))

#!/usr/bin/env python3
"""Learning curve of --task=go2 on the host: the UNMODIFIED reference trainer (rsl_rl `PPO` + `ActorCritic` + `RolloutStorage`, imported from
/root/reference/rsl_rl, device='cpu') driving the CPU oracle of this package's env (oracle/go2_oracle.cpp: the physics spec of DESIGN.md section 3
and the post-physics path pinned against the reference's own Python).  Build container only.

Purpose: evidence that the environment this package simulates is LEARNABLE with the reference's own algorithm and hyper-parameters (reward, episode
length and terrain level rise), independent of the CUDA path — which is compared with this oracle step by step elsewhere.  It is not the
"mean episode return within 1e-3 of the reference" check of BASELINE.json: that needs PhysX (closed source, absent).

Usage: python tests/tools/train_cpu_curve.py [--num_envs 1024] [--iterations 150] > profiles/<round>_cpu_oracle_learning_curve.txt
"""
import argparse
import contextlib
import io
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = ["/root/reference/rsl_rl", ROOT]

from go2_rl_gym_b200.envs.env_arrays import EnvArrays  # noqa: E402
from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg  # noqa: E402
from go2_rl_gym_b200 import _abi  # noqa: E402
from oracle.oracle import OracleEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num_envs", type=int, default=1024)
    ap.add_argument("--iterations", type=int, default=150)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--relaxed", action="store_true", help="oracle physics with the relaxed contact solver + state guard (second library build's settings)")
    args = ap.parse_args()
    from rsl_rl.algorithms import PPO
    from rsl_rl.modules import ActorCritic
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(args.seed)
    N, T = args.num_envs, 24
    cfg = GO2Cfg()
    cfg.env.num_envs, cfg.terrain.mesh_type, cfg.seed = N, "heightfield", args.seed
    if args.relaxed:
        cfg.sim.b200.limit_relax, cfg.sim.b200.contact_relax, cfg.sim.b200.limit_erp, cfg.sim.b200.state_guard = 0.5, 0.7, 0.8, 1
    A = EnvArrays(cfg, "cpu", seed=args.seed)
    env = OracleEnv(A)
    env.reset_all()
    env.step(torch.zeros(N, 12))
    Tn = A.tensors
    with contextlib.redirect_stdout(io.StringIO()):
        ac = ActorCritic(45, 263, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128], activation="elu", init_noise_std=1.0)
        alg = PPO(ac, device="cpu", value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.01, num_learning_epochs=5,
                  num_mini_batches=4, learning_rate=1e-3, schedule="adaptive", gamma=0.99, lam=0.95, desired_kl=0.01, max_grad_norm=1.0)
    alg.init_storage(N, T, [45], [263], [12])
    Tn["episode_length_buf"].copy_(torch.randint(0, int(cfg.env.episode_length_s / (cfg.control.decimation * cfg.sim.dt)), (N,)).int())   # init_at_random_ep_len
    cur_rew, cur_len = torch.zeros(N), torch.zeros(N)
    done_rew, done_len = [], []
    i_track = _abi.REWARD_NAMES.index("tracking_lin_vel")
    print(f"# {'RELAXED contact solver + state guard; ' if args.relaxed else ''}go2 rough-terrain heightfield, {N} envs, reference rsl_rl PPO (cpu, {torch.get_num_threads()} threads) on the oracle env; columns:")
    print("# iter  mean_reward/step  mean_episode_return(last 100)  mean_episode_length(last 100)  mean_terrain_level  tracking_lin_vel/step  action_std  lr  s/iter")
    for it in range(args.iterations):
        t0 = time.time()
        rew_sum = track_sum = 0.0
        with torch.inference_mode():
            for _ in range(T):
                obs, priv = Tn["obs_buf"].clone(), Tn["privileged_obs_buf"].clone()
                act = alg.act(obs, priv)
                sums0 = Tn["episode_sums"][:, i_track].clone()
                env.step(act)
                rew, dones, touts = Tn["rew_buf"].clone(), Tn["reset_buf"].bool().clone(), Tn["time_out_buf"].bool().clone()
                alg.process_env_step(rew, dones, {"time_outs": touts})
                rew_sum += float(rew.mean())
                d = Tn["episode_sums"][:, i_track] - sums0
                track_sum += float(d[~dones].mean()) if (~dones).any() else 0.0
                cur_rew += rew; cur_len += 1
                if dones.any():
                    done_rew += cur_rew[dones].tolist(); done_len += cur_len[dones].tolist()
                    cur_rew[dones] = 0; cur_len[dones] = 0
            alg.compute_returns(Tn["privileged_obs_buf"].clone())
        alg.update()
        done_rew, done_len = done_rew[-100:], done_len[-100:]
        mr = sum(done_rew) / max(len(done_rew), 1)
        ml = sum(done_len) / max(len(done_len), 1)
        print(f"{it:4d}  {rew_sum / T:9.5f}  {mr:9.3f}  {ml:8.1f}  {float(Tn['terrain_levels'].float().mean()):6.3f}  {track_sum / T:8.5f}  "
              f"{float(ac.std.mean()):6.3f}  {alg.learning_rate:.2e}  {time.time() - t0:5.2f}", flush=True)


if __name__ == "__main__":
    main()

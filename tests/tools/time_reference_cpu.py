#!/usr/bin/env python3
"""Time the UNMODIFIED reference's own Python on the host cores (SURVEY.md 8(d) "CPU baseline beside it", items 1-2).  Build container only:
it imports /root/reference, which does not exist on the GPU box (bench.py's `cpu_baseline` / `--impl reference` legs time the oracle port there).

  (1) env   the reference's `Go2Robot.step()` (legged_gym/envs/base/legged_robot.py:60-142 and everything it calls) through the isaacgym
            stand-in of tests/ref_stub, with `gym.simulate` served by the oracle's physics (PhysX is closed source and absent): wall time
            per step, split into the simulate() calls (oracle C++, NOT the reference) and the rest (the reference's torch code on CPU)
  (2) RL    the reference's rsl_rl `PPO` + `ActorCritic` + `RolloutStorage` on device='cpu' (all host threads): 24 x act /
            process_env_step, compute_returns, update at the bench's shapes

Usage: python tests/tools/time_reference_cpu.py [--num_envs 4096] [--steps 10] > profiles/<round>_reference_cpu_in_container.txt
"""
import argparse
import contextlib
import io
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_env as H  # noqa: E402  (sets sys.path: ref_stub, /root/reference, /root/reference/rsl_rl, repo root)


def time_env(N, steps):
    cfg = H.MyGO2Cfg()
    cfg.env.num_envs, cfg.terrain.mesh_type, cfg.seed = N, "heightfield", 1
    A = H.EnvArrays(cfg, "cpu", seed=1)
    O = H.OracleEnv(A)
    O.reset_all()
    g = torch.Generator().manual_seed(0)
    for _ in range(5):
        O.step(0.5 * torch.randn(N, 12, generator=g))
    A.tensors["episode_length_buf"].copy_(torch.randint(0, 1200, (N,), generator=g).int())
    import legged_gym.envs  # noqa: F401
    from legged_gym.envs.go2.go2_config import GO2Cfg as RefGO2Cfg
    ref_cfg = RefGO2Cfg()
    ref_cfg.env.num_envs, ref_cfg.terrain.mesh_type = N, "heightfield"
    draws = H.Draws(1)
    H.install_rng(draws, N)
    with contextlib.redirect_stdout(io.StringIO()):
        r = H.build_reference_env(A, O, ref_cfg)
    H.load_state_into_reference(r, A)
    r.common_step_counter = 5
    r.update_reward_curriculum(force_update=True)
    r.zero_command_proba = A.step_params(r.common_step_counter).zero_command_proba
    sim_s = [0.0]
    orig = r.gym.simulate

    def simulate(sim):
        t0 = time.perf_counter()
        orig(sim)
        sim_s[0] += time.perf_counter() - t0
    r.gym.simulate = simulate
    tot = 0.0
    for k in range(steps + 2):
        a = 0.5 * torch.randn(N, 12, generator=g)
        draws.step = r.common_step_counter + 1
        if k == 2:
            sim_s[0], tot = 0.0, 0.0
        t0 = time.perf_counter()
        r.step(a)
        tot += time.perf_counter() - t0
    ms, sim_ms = 1e3 * tot / steps, 1e3 * sim_s[0] / steps
    print(f"env   N={N:5d}: reference Go2Robot.step() {ms:8.2f} ms/step = {N / ms * 1e3:10.0f} env-steps/s   "
          f"[4 x simulate() on the oracle physics {sim_ms:7.2f} ms | the reference's own torch code {ms - sim_ms:7.2f} ms = "
          f"{N / (ms - sim_ms) * 1e3:10.0f} env-steps/s]")


def time_rl(N, T, iters):
    from rsl_rl.algorithms import PPO
    from rsl_rl.modules import ActorCritic
    torch.manual_seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        ac = ActorCritic(45, 263, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128], activation="elu", init_noise_std=1.0)
        alg = PPO(ac, device="cpu", value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.01, num_learning_epochs=5,
                  num_mini_batches=4, learning_rate=1e-3, schedule="adaptive", gamma=0.99, lam=0.95, desired_kl=0.01, max_grad_norm=1.0)
    alg.init_storage(N, T, [45], [263], [12])
    g = torch.Generator().manual_seed(2)
    t_act = t_ret = t_upd = 0.0
    for it in range(iters + 1):
        if it == 1:
            t_act = t_ret = t_upd = 0.0
        obs, priv = torch.randn(N, 45, generator=g), torch.randn(N, 263, generator=g)
        t0 = time.perf_counter()
        with torch.inference_mode():
            for _ in range(T):
                alg.act(obs, priv)
                alg.process_env_step(0.01 * obs[:, 0], obs[:, 1] > 2.5, {"time_outs": obs[:, 2] > 3.0})
            t1 = time.perf_counter()
            alg.compute_returns(priv)
        t2 = time.perf_counter()
        alg.update()
        t3 = time.perf_counter()
        t_act += t1 - t0; t_ret += t2 - t1; t_upd += t3 - t2
    a, r, u = (1e3 * x / iters for x in (t_act, t_ret, t_upd))
    print(f"RL    N={N:5d}: reference rsl_rl PPO on cpu: 24 x act/process_env_step {a:8.1f} ms, compute_returns {r:6.1f} ms, update {u:9.1f} ms "
          f"-> {N * T / (a + r + u) * 1e3:9.0f} env-steps/s for the trainer alone")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num_envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--rl_iters", type=int, default=2)
    args = ap.parse_args()
    n = os.cpu_count()
    torch.set_num_threads(n)
    print(f"# unmodified reference Python on this container's {n} host threads (torch {torch.__version__}, {torch.get_num_threads()} intra-op threads)")
    for N in (64, args.num_envs):
        time_env(N, args.steps)
    time_rl(args.num_envs, 24, args.rl_iters)


if __name__ == "__main__":
    main()

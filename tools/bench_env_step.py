#!/usr/bin/env python3
"""Micro-benchmark of the fused env step kernel alone (CUDA events, device-resident actions)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
from go2_rl_gym_b200.envs.go2.go2_env import Go2Robot

ap = argparse.ArgumentParser()
ap.add_argument("--num_envs", type=int, nargs="+", default=[4096, 8192])
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--mesh", default="heightfield")
ap.add_argument("--modes", nargs="+", default=["P2"], help="thread maps to time: P2 (default kernel), P3, 8p, 4")
args = ap.parse_args()
for N, mode in [(n, m) for n in args.num_envs for m in args.modes]:
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = args.mesh
    if os.environ.get("GO2_RELAXED_CFG") == "1":      # the second library build's settings (GO2_B200_LIB=.../libgo2b200_relaxed.so)
        cfg.sim.b200.limit_relax, cfg.sim.b200.contact_relax, cfg.sim.b200.limit_erp, cfg.sim.b200.state_guard = 0.5, 0.7, 0.8, 1
    t0 = time.time()
    env = Go2Robot(cfg, None, None, "cuda:0", True)
    env.set_step_mode(mode)
    env.reset()
    torch.cuda.synchronize()
    t_init = time.time() - t0
    acts = [0.5 * torch.randn(N, 12, device="cuda") for _ in range(8)]
    for i in range(10):
        env.step(acts[i % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        env.step(acts[i % 8])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(f"N={N} mesh={args.mesh} mode={mode}: {ms*1e3:.1f} us/step  {N/ms*1e3/1e6:.2f} M env-steps/s  (init {t_init:.1f}s, resets/step ~{float(env.reset_buf.float().mean())*N:.1f})", flush=True)
    del env

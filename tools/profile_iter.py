#!/usr/bin/env python3
"""Kernel-time table of one PPO iteration via torch.profiler (CUPTI): no serialisation, real overlap, per-kernel totals."""
import argparse, os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from go2_rl_gym_b200.envs import task_registry
from go2_rl_gym_b200.utils import get_args

ap = argparse.ArgumentParser(); ap.add_argument("--num_envs", type=int, default=4096); ap.add_argument("--task", default="go2")
a = ap.parse_args()
args = get_args(["--task", a.task, "--num_envs", str(a.num_envs), "--headless"])
env_cfg, _ = task_registry.get_cfgs(a.task); env_cfg.terrain.mesh_type = "heightfield"
env, _ = task_registry.make_env(a.task, args, env_cfg)
runner, _ = task_registry.make_alg_runner(env, a.task, args, log_root=None)
alg = runner.alg
iteration = runner.run_iteration

for _ in range(3):
    iteration()
torch.cuda.synchronize()
t0 = time.time()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    iteration()
    torch.cuda.synchronize()
wall = time.time() - t0
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name.split("(")[0][:80]
        agg[k][0] += 1; agg[k][1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
tot = sum(v[1] for v in agg.values())
print(f"wall (with profiler) {wall*1e3:.1f} ms, sum of kernel time {tot/1e3:.1f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{v[1]/1e3:9.2f} ms {v[0]:6d}x {v[1]/v[0]:8.1f} us  {100*v[1]/tot:5.1f}%  {k}")

#!/usr/bin/env python3
"""Quick timing of one PPO iteration (24 fused env steps + update) on one GPU, split into collection and learning."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from go2_rl_gym_b200.envs import task_registry
from go2_rl_gym_b200.utils import get_args

ap = argparse.ArgumentParser(); ap.add_argument("--num_envs", type=int, default=4096); ap.add_argument("--iters", type=int, default=5); ap.add_argument("--task", default="go2")
a = ap.parse_args()
args = get_args(["--task", a.task, "--num_envs", str(a.num_envs), "--headless"])
env_cfg, _ = task_registry.get_cfgs(a.task); env_cfg.terrain.mesh_type = "heightfield"
env, _ = task_registry.make_env(a.task, args, env_cfg)
runner, _ = task_registry.make_alg_runner(env, a.task, args, log_root=None)
alg = runner.alg
T1 = [0.0]
def mark():
    torch.cuda.synchronize(); T1[0] = time.time()
for it in range(a.iters + 2):
    torch.cuda.synchronize(); t0 = time.time()
    losses = runner.run_iteration(sync=mark)
    torch.cuda.synchronize(); t2 = time.time(); t1 = T1[0]
    print(f"it {it}: collect {1e3*(t1-t0):.1f} ms  learn {1e3*(t2-t1):.1f} ms  -> {a.num_envs*24/(t2-t0)/1e6:.2f} M env-steps/s  losses {[round(float(x), 4) for x in losses]} lr {alg.learning_rate:.2e}", flush=True)

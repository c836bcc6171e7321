#!/usr/bin/env python3
"""torchrun --nproc-per-node 2 tools/check_dist_graph.py [--task go2]: trains a few iterations on 2+ GPUs with the all-reduces captured inside
the update graph (default) or issued eagerly between graph segments (GO2_DIST_GRAPH=0) and prints a parameter checksum + iteration time; the
two modes must agree to float-atomic noise and every rank must hold identical parameters."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser(); ap.add_argument("--task", default="go2"); ap.add_argument("--num_envs", type=int, default=4096); ap.add_argument("--iters", type=int, default=6)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from go2_rl_gym_b200.envs import task_registry
from go2_rl_gym_b200.utils import get_args
from go2_rl_gym_b200.envs.go2.go2_env import Go2Robot
from go2_rl_gym_b200.utils.cfg_dict import class_to_dict
from go2_rl_gym_b200.rl import runners
import contextlib
env_cfg, train_cfg = task_registry.get_cfgs(a.task)
env_cfg.env.num_envs = a.num_envs; env_cfg.terrain.mesh_type = "heightfield"; env_cfg.seed = train_cfg.seed
with contextlib.redirect_stdout(sys.stderr):
    env = Go2Robot(env_cfg, None, None, f"cuda:{local}", True, env_offset=rank * a.num_envs, num_envs_global=world * a.num_envs)
    torch.manual_seed(train_cfg.seed)
    runner = getattr(runners, train_cfg.runner_class_name)(env, class_to_dict(train_cfg), log_dir=None, device=f"cuda:{local}")
if a.task != "go2":
    runner._roll_history(env.get_observations(), None); runner._hist_primed = True
t = []
for it in range(a.iters):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
    losses = runner.run_iteration()
    torch.cuda.synchronize(); t.append(time.time() - t0)
model = runner.alg.actor_critic if hasattr(runner.alg, "actor_critic") else runner.alg.model
p = model.flat_params
chk = torch.stack([p.double().sum(), p.double().abs().sum(), p.double().pow(2).sum()])
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(torch.equal(allc[0], c) for c in allc)
if rank == 0:
    g = runner.alg._graphs
    print(f"task {a.task} world {world} p2p {getattr(runner.alg, '_red', None) is not None} dist_graph {os.environ.get('GO2_DIST_GRAPH', '0')} graphs {sorted(map(str, g._g))} failed {sorted(map(str, g._failed))} "
          f"ms/iter {1e3 * sum(t[2:]) / len(t[2:]):.2f} checksum {[f'{float(x):.9e}' for x in chk]} ranks_identical {same} losses {[round(float(x), 5) for x in losses]} lr {runner.alg.learning_rate:.3e}", flush=True)
dist.destroy_process_group()

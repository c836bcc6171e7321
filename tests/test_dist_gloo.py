"""Host logic of the multi-GPU path on CPU: (1) env shards keep GLOBAL env ids, so a shard reproduces its slice of the
single-process run bit for bit; (2) world_size-2 gloo run of the trainer's collectives: averaged gradient + scalar tail equal
the full-batch values, all-reduced advantage statistics equal the global normalisation."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from go2_rl_gym_b200.envs.env_arrays import EnvArrays
from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
from emu.emu import EmuEnv


def test_env_shards_reproduce_the_global_run():
    NG = 40                                    # divisible by num_cols = 20 -> terrain types by global index
    cfg = GO2Cfg(); cfg.env.num_envs = NG; cfg.terrain.mesh_type = "heightfield"; cfg.seed = 5
    full = EnvArrays(cfg, "cpu", seed=5)
    e_full = EmuEnv(full); e_full.common_step_counter = 24 * 800; e_full.reset_all()
    shards = []
    for r in range(2):
        c = GO2Cfg(); c.env.num_envs = NG // 2; c.terrain.mesh_type = "heightfield"; c.seed = 5
        A = EnvArrays(c, "cpu", num_envs=NG // 2, env_offset=r * NG // 2, num_envs_global=NG, seed=5)
        e = EmuEnv(A); e.common_step_counter = 24 * 800; e.reset_all()
        shards.append((A, e))
    g = torch.Generator().manual_seed(0)
    for _ in range(6):
        a = 0.5 * torch.randn(NG, 12, generator=g)
        e_full.step(a)
        for r, (A, e) in enumerate(shards):
            e.step(a[r * NG // 2:(r + 1) * NG // 2])
    for r, (A, _) in enumerate(shards):
        sl = slice(r * NG // 2, (r + 1) * NG // 2)
        for k in ("root_states", "dof_pos", "obs_buf", "privileged_obs_buf", "rew_buf", "commands", "terrain_levels", "terrain_types",
                  "friction_coeffs", "body_inertia", "reset_buf", "episode_length_buf"):
            assert torch.equal(A.tensors[k], full.tensors[k][sl]), (r, k)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from go2_rl_gym_b200.rl import dist_utils
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ELU(), torch.nn.Linear(16, 3))
    X, Y = torch.randn(64, 8), torch.randn(64, 3)
    mb = 64 // world
    xs, ys = X[rank * mb:(rank + 1) * mb], Y[rank * mb:(rank + 1) * mb]
    # per-rank gradient of sum-loss scaled by 1 / (global rows), as PPO.update does (inv_count = 1 / (mb * world))
    loss = ((net(xs) - ys) ** 2).sum() / 64
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    tail = torch.tensor([float(((net(xs) - ys) ** 2).sum()), float(mb), 0.0, 0.0])
    dist_utils.allreduce_grads_and_tail(flat, tail)
    adv = torch.arange(64, dtype=torch.float64)[rank * mb:(rank + 1) * mb] ** 1.5
    stats = torch.stack([adv.sum(), (adv * adv).sum()])
    count = dist_utils.allreduce_adv_stats(stats, mb)
    if rank == 0:
        q.put((flat.numpy(), tail.numpy(), stats.numpy(), count))
    dist.destroy_process_group()


def test_gloo_world2_collectives():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat, tail, stats, count = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ELU(), torch.nn.Linear(16, 3))
    X, Y = torch.randn(64, 8), torch.randn(64, 3)
    loss = ((net(X) - Y) ** 2).mean() * 3          # = sum / 64
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).numpy()
    assert np.allclose(flat, ref, atol=1e-6)
    assert np.isclose(tail[0], float(((net(X) - Y) ** 2).sum()), rtol=1e-6) and tail[1] == 64
    adv = np.arange(64, dtype=np.float64) ** 1.5
    assert count == 64 and np.isclose(stats[0], adv.sum()) and np.isclose(stats[1], (adv * adv).sum())
    mean = stats[0] / count
    std = np.sqrt((stats[1] - count * mean * mean) / (count - 1))
    assert np.isclose(std, adv.std(ddof=1))

#!/usr/bin/env python3
"""bench.py — the headline benchmark: go2 env-steps/sec of a full PPO iteration at num_envs = 4096 per GPU.

One "step" = one PPO iteration of `--task=go2` on the rough-terrain heightfield (BASELINE.json configs[1]):
24 x (policy forward + sampling -> fused env step kernel -> transition write) + GAE + 5 epochs x 4 mini-batches of
forward / loss / backward / clip+Adam.  Reported exactly like the reference's own speed figure
(rsl_rl/runners/on_policy_runner.py:194): env-steps/s = num_steps_per_env * num_envs / (collection + learning time).

  python bench.py [--gpus N] [--steps K] [--warmup W]        this repo's arm (torchrun launches N ranks, NCCL)
  python bench.py --impl reference ...                       the reference's algorithm on the host cores (oracle port)

Timing: CUDA events on the launching stream (torch's current stream = the legacy default stream the library launches on),
barrier + synchronize on both sides, max over ranks.  Every iteration touches ~300 MB of rollout / shuffled-batch
buffers (> 126 MB L2), so iterations do not find their inputs in L2.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# two OpenMP runtimes meet in the CPU legs (the oracle's libgomp and torch's): spinning idle threads would starve each other
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
os.environ.setdefault("GOMP_SPINCOUNT", "0")

NUM_ENVS_PER_GPU = 4096
STEPS_PER_ENV = 24
ALGO_BYTES_PER_ENV_STEP = 2350          # algorithmic HBM bytes of the fused step kernel per env-step
WORKLOAD = "go2 rough-terrain heightfield, num_envs=4096 per GPU, PPO iteration (24 steps + 5 epochs x 4 mini-batches)"


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt, self.proc = index, [], threading.Event(), None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self._stop_evt.is_set():
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(self.rows)}


def cpu_worker(num_envs, warmup, steps, budget_s, out):
    """Runs in a child process: the CPU pipeline for `steps` iterations (or until the time budget), prints one JSON line."""
    import torch
    from oracle.cpu_pipeline import CpuPipeline
    cores = min(os.cpu_count() or 1, 64)             # beyond ~64 threads the 4096-env steps are dominated by fork/join
    os.environ["OMP_NUM_THREADS"] = str(cores)
    torch.set_num_threads(cores)
    pipe = CpuPipeline(num_envs, threads=cores)
    for _ in range(warmup):
        pipe.iteration()
    t0, n_steps, iters, tc, tl = time.time(), 0, 0, 0.0, 0.0
    while iters < steps and (iters == 0 or time.time() - t0 < budget_s):
        n, c, l = pipe.iteration()
        n_steps += n; iters += 1; tc += c; tl += l
    dt = time.time() - t0
    print(json.dumps({"value": n_steps / dt, "iters": iters, "seconds": dt, "cores": cores, "collect_s": tc, "learn_s": tl}), file=out, flush=True)


def run_cpu_worker(num_envs, warmup, steps, budget_s, hard_timeout_s):
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu_worker", str(num_envs), str(warmup), str(steps), str(budget_s)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=hard_timeout_s, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:  # the baseline must never take the headline number down with it
        return {"value": None, "error": f"{type(e).__name__}: {str(e)[:200]}"}


def reference_arm(args, rank, world, out):
    """The reference's algorithm on the host cores: oracle env (C++, OpenMP) + fp32 PyTorch PPO (all threads)."""
    if rank != 0:
        return
    n_sample = 1024                                  # bounded sample of the 4096-env workload per step
    r = run_cpu_worker(n_sample, min(args.warmup, 1), args.steps, 120.0, 240.0)
    v, cores = r.get("value"), r.get("cores", os.cpu_count())
    dt = r.get("seconds", 0.0) or 0.0
    if v is None:
        print(json.dumps({"impl": "reference", "unavailable": r.get("error", "cpu pipeline failed")}), file=out, flush=True)
        return
    line = {"impl": "reference", "metric": "go2 env-steps/sec (PPO iteration)", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(r.get("iters", 1), 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": f"{n_sample} envs x 24 steps + full PPO update per step (oracle physics: PhysX is closed source)"},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=out, flush=True)


def _claim_stdout():
    """stdout must carry exactly ONE JSON line: libraries write to file descriptor 1 on their own (NCCL prints its version banner there),
    so fd 1 is pointed at stderr for the whole run and the line goes to a private duplicate of the original stdout."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def main():
    out = _claim_stdout()
    try:
        _main(out)
    finally:
        out.flush()


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--num_envs", type=int, default=NUM_ENVS_PER_GPU)
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--cpu_worker", nargs=4, default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.cpu_worker:
        return cpu_worker(int(args.cpu_worker[0]), int(args.cpu_worker[1]), int(args.cpu_worker[2]), float(args.cpu_worker[3]), out)
    if args.impl == "reference":
        return reference_arm(args, rank, world, out)

    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from go2_rl_gym_b200 import _abi
    from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg, GO2CfgPPO
    from go2_rl_gym_b200.envs.go2.go2_env import Go2Robot
    from go2_rl_gym_b200.rl.runners import OnPolicyRunner
    from go2_rl_gym_b200.utils.cfg_dict import class_to_dict

    N = args.num_envs
    env_cfg, train_cfg = GO2Cfg(), GO2CfgPPO()
    env_cfg.env.num_envs = N
    env_cfg.terrain.mesh_type = "heightfield"
    env_cfg.seed = train_cfg.seed
    dev = f"cuda:{local}"
    env = Go2Robot(env_cfg, None, None, dev, True, env_offset=rank * N, num_envs_global=world * N)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):       # the module prints its layer tables like the reference does; stdout carries the JSON line only
        runner = OnPolicyRunner(env, class_to_dict(train_cfg), log_dir=None, device=dev)
    alg = runner.alg
    lib = _abi.load_library()
    env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length))
    obs, cobs = env.get_observations(), env.get_privileged_observations()
    step_ev = []

    def iteration(timed_kernel=False):
        nonlocal obs, cobs
        with torch.inference_mode():
            for _ in range(STEPS_PER_ENV):
                act = alg.act(obs, cobs)
                if timed_kernel:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                obs, cobs, rew, dones, infos = env.step(act)
                if timed_kernel:
                    e1.record()
                    step_ev.append((e0, e1))
                alg.process_env_step(rew, dones, infos)
            alg.compute_returns(cobs)
        return alg.update()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # timed region: the runner's own iteration (rollout replayed as ONE CUDA graph over device-resident step parameters, update() as another)
    for _ in range(max(args.warmup, 3)):
        runner.run_iteration()
    sync()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    from go2_rl_gym_b200.rl._ops import GraphSet
    l0 = lib.go2_kernel_launch_count() + GraphSet.replayed_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        runner.run_iteration()
    ev1.record()
    sync()
    launches = lib.go2_kernel_launch_count() + GraphSet.replayed_launches - l0      # direct launches + launches replayed from CUDA graphs
    clocks = sampler.stop()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms) / args.steps
    value = world * N * STEPS_PER_ENV / (ms * 1e-3)
    # split of the iteration like the reference's log line (collection / learning, on_policy_runner.py:138-165), device-timed
    sp_ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    col_ms = lrn_ms = 0.0
    for _ in range(3):
        sp_ev[0].record()
        runner.collect(False)
        sp_ev[1].record()
        with torch.inference_mode():
            alg.compute_returns(env.get_privileged_observations())
        alg.update()
        sp_ev[2].record()
        torch.cuda.synchronize()
        col_ms += sp_ev[0].elapsed_time(sp_ev[1]) / 3
        lrn_ms += sp_ev[1].elapsed_time(sp_ev[2]) / 3
    # the fused step kernel's own duration: the same iterations launched step by step with CUDA events around every env.step()
    obs, cobs = env.get_observations(), env.get_privileged_observations()
    for _ in range(2):
        iteration(timed_kernel=True)
    sync()
    kern_ms = sum(a.elapsed_time(b) for a, b in step_ev) / len(step_ev)          # fused step kernel (+ its 2 us finalize kernel)
    peaks, peak_kind = _peaks()
    achieved = ALGO_BYTES_PER_ENV_STEP * N / (kern_ms * 1e-3) / 1e9

    # ---- e2e: the same iteration with the env reached through its HOST-buffer C-ABI call (go2_env_step_host)
    h_act = torch.empty(N, 12).pin_memory()
    h_obs, h_priv, h_rew = torch.empty(N, 45).pin_memory(), torch.empty(N, 263).pin_memory(), torch.empty(N).pin_memory()
    h_reset = torch.empty(N, dtype=torch.uint8).pin_memory()

    def iteration_host():
        with torch.inference_mode():
            for _ in range(STEPS_PER_ENV):
                act = alg.act(env.obs_buf, env.privileged_obs_buf)
                h_act.copy_(act)                                              # D2H of the policy output (synchronises)
                env.step_host(h_act.numpy(), h_obs.numpy(), h_priv.numpy(), h_rew.numpy(), h_reset.numpy())
                alg.process_env_step(env.rew_buf, env.reset_buf, {"time_outs": env.time_out_buf})
            alg.compute_returns(env.privileged_obs_buf)
        return alg.update()                                                   # ends with the D2H read of the losses

    iteration_host()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(2, args.steps // 2)
    e0.record()
    for _ in range(n_e2e):
        iteration_host()
    e1.record()
    sync()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * N * STEPS_PER_ENV / (float(ms2) / n_e2e * 1e-3)
    h2d = STEPS_PER_ENV * N * 12 * 4
    d2h = STEPS_PER_ENV * (N * 12 * 4 + N * (45 + 263 + 1) * 4 + N) + 16

    # ---- second roofline: the largest GEMM of the update (critic layer 0 forward + transposed copy, one mini-batch), timed alone
    gemm = None
    if rank == 0:
        from go2_rl_gym_b200.rl import _ops
        Mg, Ng, Kg = 6 * N, 512, 264
        Xg, Wg, bg = torch.randn(Mg, Kg, device=dev), torch.randn(Ng, Kg, device=dev) / 16, torch.randn(Ng, device=dev)
        Yg, Ytg = torch.empty(Mg, Ng, device=dev), torch.ones(Ng + 1, Mg, device=dev)
        run = lambda: _ops.call("go2_linear_forward_tc", Xg.data_ptr(), Kg, Wg.data_ptr(), Kg, bg.data_ptr(), Yg.data_ptr(), Ng, Ytg.data_ptr(), Mg, Mg, Ng, Kg, 1)
        for _ in range(5):
            run()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        g0.record()
        for _ in range(20):
            run()
        g1.record()
        torch.cuda.synchronize()
        us = g0.elapsed_time(g1) / 20 * 1e3
        gbytes = 4.0 * (Mg * Kg + Ng * Kg + 2 * Mg * Ng)          # operands read once, both output copies written once
        gemm = {"kernel": "go2::gemm_tf32_persist_kernel<128> (Y = ELU(X W^T + b), + transposed copy)", "shape": [Mg, Ng, Kg], "bound": "hbm",
                "achieved": gbytes / us / 1e3, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbytes / us / 1e3 / peaks["hbm_gbs"],
                "kernel_us": us, "tflops": 2.0 * Mg * Ng * Kg / us / 1e6,
                "note": "fp32 activations make every MLP GEMM of the update HBM/L2-bound (33 flop/B); 20 back-to-back launches, includes host launch gaps"}
    traffic = None
    try:   # ncu --set full capture of the step kernel of this build at this size (profiles/): dram__bytes_read.sum + dram__bytes_write.sum per launch
        d = json.load(open(os.path.join(ROOT, "profiles", "step_kernel_dram.json")))
        if d.get("num_envs") == N:
            traffic = d["dram_bytes_read"] + d["dram_bytes_write"]
    except Exception:  # noqa: BLE001
        pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = run_cpu_worker(N, 1, 3, 20.0, 150.0)
        cpu = {"value": r.get("value"), "unit": "env-steps/s", "cores": r.get("cores"), "kind": "port",
               "sample": f"{r.get('iters')} PPO iteration(s) of the same {N}-env workload in {r.get('seconds', 0):.1f} s "
                         "(oracle physics + fp32 PyTorch PPO on the host threads)", "error": r.get("error")}
    if rank == 0:
        line = {"metric": "go2 env-steps/sec (PPO iteration)", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (env, losses, optimiser), tf32 multiply / f32 accumulate (MLP GEMMs)", "data": "synthetic",
                "config": {"workload": WORKLOAD, "num_envs_per_gpu": N, "task": "go2", "l2": "per-iteration working set ~300 MB > 126 MB L2",
                           "parallelism": f"env-sharded dp{world}, NCCL all-reduce per optimiser step"},
                "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches),
                "roofline": {"kernel": "go2::step_kernel_packed<2>" if os.environ.get("GO2_STEP_MODE", "P2").startswith("P") else "go2::step_kernel", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_kind": peak_kind,
                             "kernel_us": kern_ms * 1e3, "env_steps_per_s_kernel_only": N / (kern_ms * 1e-3),
                             "note": "bound by the latency of one serial articulated-body chain per 8-env CTA (2 CTAs / SM by registers), not by bytes: "
                                     "DESIGN.md 5.1, profiles/r01k_step_kernel_phase_cycles.txt; traffic = ncu dram bytes per launch (profiles/step_kernel_dram.json)"},
                "roofline_gemm": gemm, "clocks": clocks, "cpu_baseline": cpu,
                "split_ms": {"collection": col_ms, "learning": lrn_ms}}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        try:    # CUDA graphs that captured NCCL kernels (GO2_DIST_GRAPH=1) must be gone before their communicator is: drop every graph set first
            import gc
            for o in (alg, runner):
                for name in ("_graphs", "_rollout_graphs"):
                    if hasattr(o, name):
                        setattr(o, name, None)
            gc.collect()
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001 - teardown must not turn a finished measurement into a failed run
            pass
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

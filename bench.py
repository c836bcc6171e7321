#!/usr/bin/env python3
"""bench.py — the headline benchmark: env-steps/sec of a full PPO iteration of a registered go2 task, num_envs per GPU.

One "step" = one PPO iteration on the rough-terrain heightfield (default: `--task=go2`, 4096 envs per GPU = BASELINE.json configs[1]):
24 x (policy forward + sampling -> fused env step kernel -> transition write) + GAE + 5 epochs x 4 mini-batches of
forward / loss / backward / clip+Adam (CTS-family tasks: both update passes).  Reported exactly like the reference's own speed figure
(rsl_rl/runners/on_policy_runner.py:194): env-steps/s = num_steps_per_env * num_envs / (collection + learning time).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--task go2|go2_cts|go2_moe_cts|...] [--num_envs PER_GPU]
                                                             this repo's arm (torchrun launches N ranks, NCCL)
  python bench.py --impl reference ...                       the reference's algorithm on the host cores (oracle port), SAME config

The default line also carries `other_configs`: BASELINE.json configs[2..4] measured in the same run at this N (go2_cts at 8192 envs per GPU;
go2_moe_cts at 4096 and 8192 envs per GPU = 16 384 envs on 4 GPUs / 65 536 envs on 8 GPUs when the driver runs N = 4 / 8).

Timing: CUDA events on the launching stream (torch's current stream = the legacy default stream the library launches on),
barrier + synchronize on both sides, max over ranks.  Every iteration touches ~300 MB of rollout / shuffled-batch
buffers (> 126 MB L2), so iterations do not find their inputs in L2.
"""
import argparse
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
OTHER_CONFIGS = [("go2_cts", 8192), ("go2_moe_cts", 4096), ("go2_moe_cts", 8192)]      # BASELINE.json configs[2], [3] (per-GPU share), [4]
STEP_KERNELS = {"H14": "go2::step_kernel_half", "P2": "go2::step_kernel_packed<2>", "P3": "go2::step_kernel_packed<3>", "Q4": "go2::step_kernel_quad<4,4>",
                "Q2": "go2::step_kernel_quad<2,8>", "8p": "go2::step_kernel_wide<8,2,2>", "4": "go2::step_kernel"}


def workload_config(task, num_envs, world):
    """The `config` object — identical in both arms (the driver compares them)."""
    upd = "5 epochs x 4 mini-batches" + ("" if task == "go2" else ", both CTS update passes")
    return {"workload": f"{task} rough-terrain heightfield, num_envs={num_envs} per GPU, PPO iteration (24 steps + {upd})",
            "task": task, "num_envs_per_gpu": num_envs, "n_gpus": world,
            "l2": "per-iteration working set ~300 MB > 126 MB L2",
            "parallelism": f"env-sharded dp{world}, one gradient all-reduce per optimiser step (own NVLink peer-memory kernel inside the update graph)"}


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1590.0}, "fallback"


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


# ---------------------------------------------------------------------------------------------------- CPU legs (oracle port)
def cpu_worker(num_envs, warmup, steps, budget_s, out):
    """Runs in a child process: the CPU pipeline for `warmup` + `steps` iterations (timed: the `steps`; stops early only when the
    time budget is exhausted, and says how many iterations it timed), prints one JSON line."""
    import torch
    from oracle.cpu_pipeline import CpuPipeline
    cores = min(os.cpu_count() or 1, 64)             # beyond ~64 threads the 4096-env steps are dominated by fork/join
    os.environ["OMP_NUM_THREADS"] = str(cores)
    torch.set_num_threads(cores)
    pipe = CpuPipeline(num_envs, threads=cores)
    t_start, done_w = time.time(), 0
    for _ in range(warmup):
        pipe.iteration()
        done_w += 1
        if time.time() - t_start > 0.3 * budget_s:
            break
    t0, n_steps, iters, tc, tl = time.time(), 0, 0, 0.0, 0.0
    while iters < steps and (iters == 0 or time.time() - t_start < budget_s):
        n, c, l = pipe.iteration()
        n_steps += n; iters += 1; tc += c; tl += l
    dt = time.time() - t0
    print(json.dumps({"value": n_steps / dt, "iters": iters, "warmup_iters": done_w, "seconds": dt, "cores": cores, "collect_s": tc, "learn_s": tl}),
          file=out, flush=True)


def run_cpu_worker(num_envs, warmup, steps, budget_s, hard_timeout_s):
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu_worker", str(num_envs), str(warmup), str(steps), str(budget_s)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=hard_timeout_s, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:  # the baseline must never take the headline number down with it
        return {"value": None, "error": f"{type(e).__name__}: {str(e)[:200]}"}


def reference_arm(args, rank, world, out):
    """The reference's algorithm on the host cores — oracle env (C++, OpenMP over envs; PhysX is closed source) + the reference's PPO
    restated in fp32 PyTorch on all host threads — on THIS arm's config: the same task, the same num_envs, full PPO iterations, the
    requested warm-up and step counts (a 4096-env iteration takes ~2 s on 16 host threads; a 300 s budget bounds the run)."""
    if rank != 0:
        return
    if args.task != "go2":
        print(json.dumps({"impl": "reference", "unavailable": f"the CPU port restates PPO / ActorCritic (task go2) only, not {args.task}"}), file=out, flush=True)
        return
    r = run_cpu_worker(args.num_envs, args.warmup, args.steps, 300.0, 600.0)
    v, cores = r.get("value"), r.get("cores", os.cpu_count())
    if v is None:
        print(json.dumps({"impl": "reference", "unavailable": r.get("error", "cpu pipeline failed")}), file=out, flush=True)
        return
    iters = int(r.get("iters", 0))
    line = {"impl": "reference", "metric": "go2 env-steps/sec (PPO iteration)", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": iters,
            "warmup": int(r.get("warmup_iters", 0)), "ms_per_step": 1e3 * r["seconds"] / max(iters, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.task, args.num_envs, args.gpus),
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": f"{iters} full PPO iteration(s) of the {args.num_envs}-env workload of ONE rank ({args.num_envs} envs x 24 steps + 5 epochs x 4 "
                                       f"mini-batches) after {r.get('warmup_iters')} warm-up iteration(s), {r['seconds']:.1f} s: oracle C++ physics + post-physics "
                                       "(PhysX is closed source) and the reference's PPO restated in fp32 PyTorch, all host threads"},
            "split_ms": {"collection": 1e3 * r["collect_s"] / max(iters, 1), "learning": 1e3 * r["learn_s"] / max(iters, 1)},
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


# ---------------------------------------------------------------------------------------------------- this repo's arm
def make_runner(task, N, rank, world, dev):
    """task_registry's own construction (make_env / make_alg_runner) with the shard arguments of the multi-GPU layout."""
    import contextlib
    from go2_rl_gym_b200.envs import task_registry
    from go2_rl_gym_b200.envs.go2.go2_env import Go2Robot
    from go2_rl_gym_b200.rl import runners
    from go2_rl_gym_b200.utils.cfg_dict import class_to_dict
    env_cfg, train_cfg = task_registry.get_cfgs(task)
    env_cfg.env.num_envs = N
    env_cfg.terrain.mesh_type = "heightfield"
    env_cfg.seed = train_cfg.seed
    env = Go2Robot(env_cfg, None, None, dev, True, env_offset=rank * N, num_envs_global=world * N)
    with contextlib.redirect_stdout(sys.stderr):       # the modules print their layer tables like the reference does; stdout carries the JSON line only
        runner = getattr(runners, train_cfg.runner_class_name)(env, class_to_dict(train_cfg), log_dir=None, device=dev)
    import torch
    env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length))       # on_policy_runner.py:118
    return env, runner


def time_iterations(runner, steps, warmup, sync, world, dev):
    """-> ms per iteration, device-timed, max over ranks"""
    import torch
    import torch.distributed as dist
    for _ in range(warmup):
        runner.run_iteration()
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        runner.run_iteration()
    ev1.record()
    sync()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) / steps


def split_iteration(runner, reps=3):
    """collection / learning split like the reference's log line (on_policy_runner.py:138-165), device-timed"""
    import torch
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    col = lrn = 0.0
    mark = lambda: ev[1].record()
    for _ in range(reps):
        ev[0].record()
        runner.run_iteration(sync=mark)
        ev[2].record()
        torch.cuda.synchronize()
        col += ev[0].elapsed_time(ev[1]) / reps
        lrn += ev[1].elapsed_time(ev[2]) / reps
    return {"collection": col, "learning": lrn}


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--task", default="go2", help="registered task (go2, go2_cts, go2_moe_cts, go2_moe_ng_cts, go2_mcp_cts, go2_ac_moe_cts, go2_dual_moe_cts)")
    ap.add_argument("--num_envs", type=int, default=NUM_ENVS_PER_GPU, help="envs PER GPU (weak scaling)")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_other_configs", action="store_true")
    ap.add_argument("--cpu_worker", nargs=4, default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.cpu_worker:
        return cpu_worker(int(args.cpu_worker[0]), int(args.cpu_worker[1]), int(args.cpu_worker[2]), float(args.cpu_worker[3]), out)
    if args.impl == "reference":
        return reference_arm(args, rank, world, out)

    import gc
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
    from go2_rl_gym_b200.rl import _ops
    from go2_rl_gym_b200.rl._ops import GraphSet

    N, task, warmup = args.num_envs, args.task, max(args.warmup, 3)
    dev = f"cuda:{local}"
    lib = _abi.load_library()
    passes = _ops.lib().go2_gemm_get_passes()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def drop(*objs):
        """CUDA graphs that captured NCCL kernels must be gone before their communicator is; also frees a finished config's buffers"""
        for o in objs:
            for name in ("_graphs", "_rollout_graphs"):
                for holder in (o, getattr(o, "alg", None)):
                    if holder is not None and hasattr(holder, name):
                        setattr(holder, name, None)
        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    env, runner = make_runner(task, N, rank, world, dev)
    alg = runner.alg
    is_cts = hasattr(runner, "history")

    # ---- timed region: the runner's own iteration (rollout replayed as ONE CUDA graph over device-resident step parameters, update() as another)
    for _ in range(warmup):
        runner.run_iteration()
    sync()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = lib.go2_kernel_launch_count() + GraphSet.replayed_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        runner.run_iteration()                                       # every iteration ends with the host read of its 5 logged means (losses, lr)
    ev1.record()
    sync()
    launches = lib.go2_kernel_launch_count() + GraphSet.replayed_launches - l0      # direct launches + launches replayed from CUDA graphs
    clocks = sampler.stop()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms) / args.steps
    value = world * N * STEPS_PER_ENV / (ms * 1e-3)
    split = split_iteration(runner)

    # ---- the fused step kernel's own duration: a rollout launched step by step with CUDA events around every env.step()
    step_ev = []

    def act(obs, priv):
        return alg.act(obs, priv, runner.history.flatten(1)) if is_cts else alg.act(obs, priv)

    def after_step(obs, dones):
        if is_cts:
            runner._roll_history(obs, dones)

    def rollout_eager_timed():
        """the rollout launched step by step (no graph) through the device-parameter entry point the graph replays (go2_env_step_dev): the events
        bracket exactly the step kernel + its finalize kernel, no host -> device parameter copy in between"""
        obs, priv = env.get_observations(), env.get_privileged_observations()
        alg.storage.step = 0
        with torch.inference_mode():
            dev_params = env.begin_rollout(STEPS_PER_ENV)
            if dev_params:
                alg.begin_rollout(STEPS_PER_ENV)
            for i in range(STEPS_PER_ENV):
                a = act(obs, priv)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                obs, priv, rew, dones, infos = env.step_dev(a, i) if dev_params else env.step(a)
                e1.record()
                step_ev.append((e0, e1))
                after_step(obs, dones)
                alg.process_env_step(rew, dones, infos)
            if dev_params:
                alg.end_rollout(STEPS_PER_ENV)
                env.end_rollout(fetch=False)
    for _ in range(2):
        rollout_eager_timed()
    sync()
    kern_ms = sum(a.elapsed_time(b) for a, b in step_ev) / len(step_ev)          # fused step kernel (+ its 2 us finalize kernel)
    peaks, peak_kind = _peaks()
    achieved = ALGO_BYTES_PER_ENV_STEP * N / (kern_ms * 1e-3) / 1e9

    # ---- e2e: the same iteration with the env reached through its HOST-buffer C-ABI calls (go2_env_step_host_begin / _end): every step uploads the
    # actions from pinned host memory and reads observations / privileged observations / rewards / resets back
    h_act = torch.empty(N, 12).pin_memory()
    h_obs, h_priv, h_rew = torch.empty(N, 45).pin_memory(), torch.empty(N, 263).pin_memory(), torch.empty(N).pin_memory()
    h_reset = torch.empty(N, dtype=torch.uint8).pin_memory()

    def iteration_host():
        """the runner's own host-buffer iteration (rl/runners: run_iteration_host -> collect_host -> env.step_host_begin / _end = go2_env_step_host_begin / _end)"""
        return runner.run_iteration_host(h_act, h_obs, h_priv, h_rew, h_reset)

    for _ in range(2):                                                        # eager pass, then the pass that captures the per-step graphs
        iteration_host()
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
    step_mode = os.environ.get("GO2_STEP_MODE", getattr(env, "step_mode", "H14"))
    drop(runner)
    del env, runner, alg
    gc.collect(); torch.cuda.empty_cache()

    # ---- second roofline: the largest GEMM of the update (critic layer 0 forward, one mini-batch of 6 N rows), timed alone at both precisions
    gemm = None
    if rank == 0:
        Mg, Ng, Kg = 6 * N, 512, 264
        Xg, Wg, bg = torch.randn(Mg, Kg, device=dev), torch.randn(Ng, Kg, device=dev) / 16, torch.randn(Ng, device=dev)
        Yg = torch.empty(Mg, Ng, device=dev)
        flush = torch.empty(64 * 1024 * 1024, device=dev)                    # 256 MB > 126 MB L2, rewritten between launches
        run = lambda: _ops.call("go2_linear_forward_tc", Xg.data_ptr(), Kg, Wg.data_ptr(), Kg, bg.data_ptr(), Yg.data_ptr(), Ng, 0, 0, Mg, Ng, Kg, 1)
        us_by_mode = {}
        L_ = _ops.lib()
        for mode, pz, pair in (("tf32", 1, 0), ("3xtf32_pair", 3, 1), ("3xtf32_1cta", 3, 0)):
            L_.go2_gemm_set_passes(pz); L_.go2_gemm_set_pair(pair)
            for _ in range(3):
                run()
            tot = 0.0
            for _ in range(10):
                flush.zero_()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(); run(); g1.record()
                torch.cuda.synchronize()
                tot += g0.elapsed_time(g1)
            us_by_mode[mode] = tot / 10 * 1e3
        L_.go2_gemm_set_passes(passes); L_.go2_gemm_set_pair(1)
        us = us_by_mode["3xtf32_pair" if passes == 3 else "tf32"]
        gbytes = 4.0 * (Mg * Kg + Ng * Kg + Mg * Ng)                 # ALGORITHMIC bytes: operands read once, ONE output written once
        tfl = 2.0 * Mg * Ng * Kg / us / 1e6
        tpeak = peaks.get("bf16_tflops", peaks.get("bf16_tflops_sustained"))      # a kernel timed alone: the burst figure
        kname = "go2::gemm_tf32_pair_kernel<256,true> (tcgen05.mma.cta_group::2, 3xTF32)" if passes == 3 else "go2::gemm_tf32_persist_kernel<128,false>"
        gemm = {"kernel": f"{kname} (Y = ELU(X W^T + b), critic layer 0 forward)",
                "shape": [Mg, Ng, Kg], "passes": passes, "bound": "tensor", "achieved": tfl * passes, "peak": tpeak / 2 if tpeak else None, "unit": "TFLOP/s",
                "frac": (tfl * passes) / (tpeak / 2) if tpeak else None, "kernel_us": us, "kernel_us_by_mode": us_by_mode,
                "useful_tflops": tfl, "hbm_gbs": gbytes / us / 1e3, "hbm_frac": gbytes / us / 1e3 / peaks["hbm_gbs"],
                "note": "tf32 runs at half the bf16 tensor rate: peak = MEASURED_PEAKS bf16_tflops (burst, kernel timed alone) / 2; achieved = tensor flops ISSUED (3 MMAs per operand pair), "
                        "useful_tflops = fp32-class product flops.  Measured bound (profiles/r02*_gemm_*): shared-memory bandwidth of the operand stream "
                        "(128 B/clk/SM) — a 128x128x8 tf32 MMA reads 8 KB per 64 cycles; CTA pairs halve it.  hbm_* = algorithmic bytes (operands once + "
                        "one output, 76.8 MB at 4096 envs); L2 flushed between launches"}
        del Xg, Wg, bg, Yg, flush
    traffic, issue_active = None, None
    try:   # ncu --set full capture of the step kernel of this build at this size (profiles/): dram bytes per launch, issue-slot utilisation
        d = json.load(open(os.path.join(ROOT, "profiles", "step_kernel_dram.json")))
        if d.get("num_envs") == N:
            traffic = d["dram_bytes_read"] + d["dram_bytes_write"]
            issue_active = d.get("issue_active_pct")
    except Exception:  # noqa: BLE001
        pass

    # ---- BASELINE.json configs[2..4] at this N (short runs: warm-up 3, 5 timed iterations each)
    others = None
    if task == "go2" and N == NUM_ENVS_PER_GPU and not args.no_other_configs:
        others = []
        for otask, on in OTHER_CONFIGS:
            try:
                oenv, orunner = make_runner(otask, on, rank, world, dev)
                oms = time_iterations(orunner, 5, 3, sync, world, dev)
                osplit = split_iteration(orunner)
                others.append({"config": workload_config(otask, on, world), "value": world * on * STEPS_PER_ENV / (oms * 1e-3), "unit": "env-steps/s",
                               "ms_per_step": oms, "steps": 5, "warmup": 3, "total_envs": world * on, "split_ms": osplit})
                drop(orunner)
                del oenv, orunner
                gc.collect(); torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001 - a side measurement must not take the headline down
                others.append({"config": workload_config(otask, on, world), "error": f"{type(e).__name__}: {str(e)[:300]}"})

    cpu = None
    if rank == 0 and world == 1 and task == "go2" and not args.no_cpu_baseline:
        r = run_cpu_worker(N, 1, 3, 30.0, 150.0)
        cpu = {"value": r.get("value"), "unit": "env-steps/s", "cores": r.get("cores"), "kind": "port",
               "sample": f"{r.get('iters')} PPO iteration(s) of the same {N}-env workload in {r.get('seconds', 0):.1f} s after 1 warm-up iteration "
                         "(oracle physics + fp32 PyTorch PPO on the host threads)", "error": r.get("error")}
    if rank == 0:
        gemm_dtype = {3: "3xTF32 split on tcgen05 (fp32-class products, f32 accumulate)", 1: "tf32 multiply / f32 accumulate"}[passes] \
            if _ops.use_tc() else "f32 CUDA-core GEMMs"
        line = {"metric": "go2 env-steps/sec (PPO iteration)", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": f"f32 (env, losses, optimiser); MLP GEMMs: {gemm_dtype}", "data": "synthetic",
                "config": workload_config(task, N, world),
                "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches),
                "roofline": {"kernel": STEP_KERNELS.get(step_mode, step_mode), "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "issue_active": issue_active, "peak_kind": peak_kind,
                             "kernel_us": kern_ms * 1e3, "env_steps_per_s_kernel_only": N / (kern_ms * 1e-3),
                             "note": "latency / issue bound, not byte bound (DESIGN.md 5.1): issue_active = ncu sm__inst_issued / cycles (%), "
                                     "traffic = ncu dram bytes per launch (both from profiles/step_kernel_dram.json, same build and size)"},
                "roofline_gemm": gemm, "clocks": clocks, "cpu_baseline": cpu, "split_ms": split, "other_configs": others}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        try:
            gc.collect()
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001 - teardown must not turn a finished measurement into a failed run
            pass
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

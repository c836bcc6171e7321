#!/usr/bin/env python3
"""The trainer's own GEMM calls (go2 PPO update, M = 24576 rows per mini-batch; rollout inference, M = 4096), per multiply mode:
   tf32      one tf32 pass (round 1's kernel)
   3x        3xTF32, lo-only split (hardware truncation of the raw word is the hi part), CTA pairs (cta_group::2) — the default
   3x-1cta   the same arithmetic on the one-CTA persistent kernel (go2_gemm_set_pair(0))
   3x-rw     3xTF32, stage rewritten with rn_tf32 (go2_gemm_set_split(1); only with --rw)
Columns: time per call (CUDA events, L2 flushed by the working set of the shape list) and relative error (norm-wise) against fp64."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from go2_rl_gym_b200.rl import _ops

import ctypes
L = _ops.lib()
L.go2_gemm_set_debug.argtypes = [ctypes.c_void_p]
MODES = (("tf32", 1, 0, 0), ("3x", 3, 0, 1), ("3x-1cta", 3, 0, 0)) + ((("3x-rw", 3, 1, 1),) if "--rw" in sys.argv else ())


def set_mode(passes, rewrite, pair=1):
    assert L.go2_gemm_set_passes(passes) == 0 and L.go2_gemm_set_split(rewrite) == 0 and L.go2_gemm_set_pair(pair) == 0


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


NAMES = ["prod_total", "prod_w_empty", "mma_total", "mma_w_tempty", "mma_w_full", "mma_w_lofull", "stages", "split_w_full", "split_w_loempty", "split_busy",
         "epi0_total", "epi0_w_tfull", "epi1_total", "epi1_w_tfull", "epi0_w_aux", "epi1_w_aux"]
DBG = None


def profile(fn, label):
    """one launch with the per-role cycle counters on (go2_gemm_set_debug): mean over the CTAs that ran, cycles per stage"""
    global DBG
    if "--dbg" not in sys.argv:
        return
    if DBG is None:
        DBG = torch.zeros(148 * 24, dtype=torch.int64, device="cuda")
    fn(); torch.cuda.synchronize()
    DBG.zero_()
    assert L.go2_gemm_set_debug(DBG.data_ptr()) == 0
    fn(); torch.cuda.synchronize()
    assert L.go2_gemm_set_debug(None) == 0
    d = DBG.view(148, 24).double().cpu()
    d = d[d[:, 6] > 0]
    g = d[:, 16:20]
    t0 = g[:, 0].min()
    span = f"; ns from first CTA entry: prologue done {float((g[:, 1] - t0).mean()):.0f} (mean), dependency wait done {float((g[:, 2] - t0).mean()):.0f}, exit mean {float((g[:, 3] - t0).mean()):.0f} / max {float((g[:, 3] - t0).max()):.0f}, entry spread {float((g[:, 0] - t0).max()):.0f}"
    m = d.mean(0); st = float(m[6])
    print(f"    [{label}] {len(d)} CTAs, {st:.1f} stages/CTA; cycles per stage: " + ", ".join(f"{n} {float(m[i]) / st:.0f}" for i, n in enumerate(NAMES) if i != 6 and m[i] > 0) + span, flush=True)


def rel(a, b):
    return float((a.double() - b).norm() / b.norm())


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = []
    for M in (24576, 4096):
        for (N, K) in ((512, 48), (512, 264), (256, 512), (128, 256)):
            X = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
            b = torch.randn(N, device="cuda", generator=g); Y = torch.empty(M, N + 4, device="cuda")
            ref = torch.nn.functional.elu(X.double() @ W.double().t() + b.double())
            out = [f"fwd   M={M:6d} N={N:4d} K={K:4d}"]
            for name, ps, rw, pr in MODES:
                set_mode(ps, rw, pr)
                fn = lambda: _ops.call("go2_linear_forward_tc", X.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Y.data_ptr(), N + 4, 0, 0, M, N, K, 1)
                us = timeit(fn)
                out.append(f"{name} {us:7.1f} us err {rel(Y[:, :N], ref):.2e}")
            rows.append("   ".join(out)); print(rows[-1], flush=True)
            for name, ps, rw, pr in MODES[:3]:
                set_mode(ps, rw, pr); profile(fn, name)
    M = 24576
    for (N, K) in ((256, 512), (128, 256), (512, 256), (256, 128)):      # dZ [M, N] -> dX [M, K]
        dZ = torch.randn(M, N, device="cuda", generator=g); Wt = torch.randn(K, N, device="cuda", generator=g) / math.sqrt(N)
        act = torch.nn.functional.elu(torch.randn(M, K + 4, device="cuda", generator=g)); dX = torch.empty(M, K, device="cuda")
        a = act[:, :K].double()
        ref = (dZ.double() @ Wt.double().t()) * torch.where(a > 0, torch.ones_like(a), a + 1)
        out = [f"dgrad M={M:6d} N={N:4d} K={K:4d}"]
        for name, ps, rw, pr in MODES:
            set_mode(ps, rw, pr)
            fn = lambda: _ops.call("go2_linear_dgrad_tc", dZ.data_ptr(), N, Wt.data_ptr(), N, act.data_ptr(), K + 4, 0, 0, dX.data_ptr(), K, 0, 0, M, N, K)
            us = timeit(fn)
            out.append(f"{name} {us:7.1f} us err {rel(dX, ref):.2e}")
        rows.append("   ".join(out)); print(rows[-1], flush=True)
        for name, ps, rw, pr in MODES[:3]:
            set_mode(ps, rw, pr); profile(fn, name)
    for (N, K) in ((512, 48), (512, 264), (256, 512), (128, 256)):       # dW [N, K] = dZ^T X
        dZ = torch.randn(M, N, device="cuda", generator=g); X = torch.ones(M, K + 4, device="cuda"); X[:, :K] = torch.nn.functional.elu(torch.randn(M, K, device="cuda", generator=g))
        dW, db = torch.empty(N, K, device="cuda"), torch.empty(N, device="cuda")
        work = torch.empty(64 * ((N + 127) // 128 * 128) * ((K + 4) // 4 * 4), device="cuda")
        ref = dZ.double().t() @ X[:, :K].double()
        r32 = rel(dZ.t() @ X[:, :K], ref)
        out = [f"wgrad M={M:6d} N={N:4d} K={K:4d}"]
        for name, ps, rw, pr in MODES:
            set_mode(ps, rw, pr)
            fn = lambda: _ops.call("go2_linear_wgrad_tc_rm", dZ.data_ptr(), N, X.data_ptr(), K + 4, dW.data_ptr(), K, db.data_ptr(), M, N, K, work.data_ptr(), work.numel())
            us = timeit(fn)
            out.append(f"{name} {us:7.1f} us err {rel(dW, ref):.2e}")
        out.append(f"(torch fp32 matmul err {r32:.2e})")
        rows.append("   ".join(out)); print(rows[-1], flush=True)
        for name, ps, rw, pr in MODES[:3]:
            set_mode(ps, rw, pr); profile(fn, name)
    set_mode(3, 0, 1)


if __name__ == "__main__":
    main()

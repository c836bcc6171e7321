#!/usr/bin/env python3
"""Per-shape timing of the tensor-core GEMM entry points (CUDA events, 30 repetitions after 5 warm-ups)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from go2_rl_gym_b200.rl import _ops

def timeit(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

print(f"{'op':8s} {'M':>6s} {'N':>5s} {'K':>5s} {'us':>8s} {'TFLOP/s':>8s} {'GB/s':>8s}")
for M in (24576, 4096):
    for (N, K) in ((512, 48), (512, 264), (256, 512), (128, 256), (12, 128)):
        X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / math.sqrt(K); b = torch.randn(N, device="cuda")
        Y = torch.empty(M, N, device="cuda"); Yt = torch.empty(N, M, device="cuda")
        for name, yt in (("fwd", 0), ("fwd+T", Yt.data_ptr())):
            us = timeit(lambda: _ops.call("go2_linear_forward_tc", X.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Y.data_ptr(), N, yt, M, M, N, K, 1))
            byts = 4 * (M * K + N * K + M * N * (2 if yt else 1))
            print(f"{name:8s} {M:6d} {N:5d} {K:5d} {us:8.1f} {2*M*N*K/us/1e6:8.1f} {byts/us/1e3:8.0f}")
M = 24576
for (N, K) in ((512, 264), (256, 512), (128, 256), (12, 128)):
    dZt = torch.randn(N, M, device="cuda"); Xt = torch.randn(K, M, device="cuda"); dW = torch.empty(N, K, device="cuda")
    work = torch.empty(64 * N * ((K + 3) // 4 * 4), device="cuda")
    us = timeit(lambda: _ops.call("go2_linear_wgrad_tc", dZt.data_ptr(), M, Xt.data_ptr(), M, dW.data_ptr(), K, 0, M, N, K, work.data_ptr(), work.numel()))
    print(f"{'wgrad':8s} {M:6d} {N:5d} {K:5d} {us:8.1f} {2*M*N*K/us/1e6:8.1f} {4*(M*N+M*K)/us/1e3:8.0f}")
for (N, K) in ((256, 512), (128, 256), (12, 128)):
    dZ = torch.randn(M, N, device="cuda"); Wt = torch.randn(K, N, device="cuda"); At = torch.randn(K, M, device="cuda")
    dX = torch.empty(M, K, device="cuda"); dXt = torch.empty(K, M, device="cuda")
    us = timeit(lambda: _ops.call("go2_linear_dgrad_tc", dZ.data_ptr(), N, Wt.data_ptr(), N, 0, 0, At.data_ptr(), M, dX.data_ptr(), K, dXt.data_ptr(), M, M, N, K))
    print(f"{'dgrad':8s} {M:6d} {N:5d} {K:5d} {us:8.1f} {2*M*N*K/us/1e6:8.1f} {4*(M*N+3*M*K)/us/1e3:8.0f}")

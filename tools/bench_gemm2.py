#!/usr/bin/env python3
"""GPU-time table of every GEMM of one PPO mini-batch step (M = 24576) and of one rollout step (M = 4096), replayed from a CUDA
graph so host launch cost is excluded.  Columns: us per call, TFLOP/s, algorithmic GB/s (operands read once + outputs written once)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from go2_rl_gym_b200.rl import _ops

def graph_time(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

def P(t): return t.data_ptr()
print(f"{'op':8s} {'M':>6s} {'N':>5s} {'K':>5s} {'us':>8s} {'TFLOP/s':>8s} {'GB/s':>8s}")
tot = {}
for M in (24576, 4096):
    tot[M] = 0.0
    for (N, K) in ((512, 48), (512, 264), (256, 512), (128, 256), (12, 128)):
        X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / math.sqrt(K); b = torch.randn(N, device="cuda")
        Y = torch.empty(M, N, device="cuda"); Yt = torch.ones(N + 1, M, device="cuda")
        train = M == 24576
        us = graph_time(lambda: _ops.call("go2_linear_forward_tc", P(X), K, P(W), K, P(b), P(Y), N, P(Yt) if train else 0, M, M, N, K, 1))
        byts = 4 * (M * K + N * K + M * N * (2 if train else 1))
        mult = 1 if (N, K) in ((512, 48), (512, 264), (12, 128)) else 2
        tot[M] += us * mult
        print(f"{'fwd+T' if train else 'fwd':8s} {M:6d} {N:5d} {K:5d} {us:8.1f} {2*M*N*K/us/1e6:8.1f} {byts/us/1e3:8.0f}")
M = 24576
for (N, K) in ((512, 48), (512, 264), (256, 512), (128, 256), (12, 128)):
    dZt = torch.randn(N, M, device="cuda"); Xt = torch.ones(K + 1, M, device="cuda"); dW = torch.empty(N, K, device="cuda"); db = torch.empty(N, device="cuda")
    work = torch.empty(64 * 128 * ((N + 127) // 128) * ((K + 4) // 4 * 4), device="cuda")
    us = graph_time(lambda: _ops.call("go2_linear_wgrad_tc", P(dZt), M, P(Xt), M, P(dW), K, P(db), M, N, K, P(work), work.numel()))
    tot[M] += us * (1 if (N, K) in ((512, 48), (512, 264), (12, 128)) else 2)
    print(f"{'wgrad':8s} {M:6d} {N:5d} {K:5d} {us:8.1f} {2*M*N*K/us/1e6:8.1f} {4*(M*N+M*K)/us/1e3:8.0f}")
for (N, K) in ((256, 512), (128, 256), (12, 128)):
    dZ = torch.randn(M, N, device="cuda"); Wt = torch.randn(K, (N + 3) // 4 * 4, device="cuda"); A = torch.randn(M, K, device="cuda"); At = torch.randn(K + 1, M, device="cuda")
    dX = torch.empty(M, K, device="cuda"); dXt = torch.empty(K, M, device="cuda")
    fn = "go2_linear_dgrad_tc" if N % 4 == 0 else None
    if fn is None: continue
    us = graph_time(lambda: _ops.call(fn, P(dZ), N, P(Wt), Wt.shape[1], P(A), K, P(At), M, P(dX), K, P(dXt), M, M, N, K))
    tot[M] += us * 2
    print(f"{'dgrad':8s} {M:6d} {N:5d} {K:5d} {us:8.1f} {2*M*N*K/us/1e6:8.1f} {4*(M*N+3*M*K)/us/1e3:8.0f}")
print(f"sum over one PPO mini-batch step (actor+critic, without the two 12/1-wide heads' fwd/dgrad): {tot[24576]:.0f} us; one rollout step fwd: {tot[4096]:.0f} us")

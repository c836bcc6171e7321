import math, os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from go2_rl_gym_b200 import _abi
lib = _abi.load_library()
vp, i = C.c_void_p, C.c_int
lib.go2_linear_forward_tc_dbg.argtypes = [vp, i, vp, i, vp, vp, i, vp, i, i, i, i, i, vp, vp]
for (M, N, K) in ((4096, 512, 48), (24576, 512, 264)):
    X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
    Y = torch.empty(M, N, device="cuda"); Yt = torch.empty(N, M, device="cuda")
    nct = ((M + 127) // 128) * ((N + 127) // 128)
    dbg = torch.zeros(nct, 8, dtype=torch.int64, device="cuda")
    for _ in range(3):
        lib.go2_linear_forward_tc_dbg(X.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Y.data_ptr(), N, Yt.data_ptr(), M, M, N, K, 1, dbg.data_ptr(), None)
    torch.cuda.synchronize()
    d = dbg.cpu().double()
    t0 = d[:, 0].min()
    print(f"M={M} N={N} K={K}: ctas={nct}")
    print("  setup      (cycles) mean %.0f" % (d[:, 1] - d[:, 0]).mean().item())
    print("  mainloop   (cycles) mean %.0f" % (d[:, 2] - d[:, 1]).mean().item())
    print("  epilogue   (cycles) mean %.0f" % (d[:, 3] - d[:, 2]).mean().item())
    print("  teardown   (cycles) mean %.0f" % (d[:, 4] - d[:, 3]).mean().item())
    print("  first start %.0f, last end %.0f (span cycles), start spread p50 %.0f p100 %.0f" % (0, (d[:, 4].max() - t0).item(), (d[:, 0] - t0).median().item(), (d[:, 0] - t0).max().item()))

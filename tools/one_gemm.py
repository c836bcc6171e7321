import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from go2_rl_gym_b200.rl import _ops
M, N, K = 24576, 512, 264
X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / math.sqrt(K); b = torch.randn(N, device="cuda")
Y = torch.empty(M, N, device="cuda"); Yt = torch.empty(N, M, device="cuda")
for _ in range(4):
    _ops.call("go2_linear_forward_tc", X.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Y.data_ptr(), N, Yt.data_ptr(), M, M, N, K, 1)
torch.cuda.synchronize()

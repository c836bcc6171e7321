TAG=${1:-r02p}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_cts.py tests/test_gpu_x_moe_heads.py -m gpu -q -rf --tb=short 2>&1 | grep -v "^Actor MLP\|^Critic MLP\|Linear(\|ELU(\|^)\|Sequential\|^$" | cut -c1-400 | tail -12
for t in go2_cts go2_moe_cts; do timeout 300 python tools/bench_iter.py --task $t --num_envs 8192 --iters 3 2>&1 | grep "^it" | tail -1; done

TAG=${1:-r02i}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_rl.py tests/test_gpu_cts.py tests/test_gpu_x_moe_heads.py -m gpu -q -rf --tb=short --deselect tests/test_gpu_rl.py::test_go2_learns_on_the_gpu 2>&1 | grep -v "^Actor MLP\|^Critic MLP\|Linear(\|ELU(\|^)\|Sequential\|^$" | cut -c1-400 | tail -30 > $O/gpu_tests_$TAG.log
tail -8 $O/gpu_tests_$TAG.log
for t in go2 go2_cts go2_moe_cts; do for ts in 0 1; do echo "$t TWO_STREAMS=$ts"; GO2_TWO_STREAMS=$ts timeout 300 python tools/bench_iter.py --task $t --num_envs 4096 --iters 3 2>&1 | grep "^it\|Warn\|warn" | tail -1; done; done

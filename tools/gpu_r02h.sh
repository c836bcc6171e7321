TAG=${1:-r02h}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_rl.py -m gpu -q -rf --tb=short -x --deselect tests/test_gpu_rl.py::test_go2_learns_on_the_gpu 2>&1 | grep -v "^Actor MLP\|^Critic MLP\|Linear(\|ELU(\|^)\|Sequential\|^$" | cut -c1-400 | tail -30 > $O/gpu_tests_$TAG.log
tail -5 $O/gpu_tests_$TAG.log
for ts in 0 1; do echo "TWO_STREAMS=$ts"; GO2_TWO_STREAMS=$ts timeout 300 python tools/bench_iter.py --task go2 --num_envs 4096 --iters 3 2>&1 | grep "^it\|Warn\|warn" | tail -2; done

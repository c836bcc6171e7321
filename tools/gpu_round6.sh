TAG=${1:-r01l}
O=gpurun_out
mkdir -p $O
set -x
timeout 1200 python -m pytest tests/test_gpu_cts.py tests/test_gpu_rl.py -m gpu -q 2>&1 | tail -40 > $O/gpu_tests_$TAG.log
timeout 300 python tools/bench_iter.py --task go2_moe_ng_cts --num_envs 8192 --iters 5 2>&1 | grep "^it" | tail -2 > $O/iter_ng_$TAG.log
tail -40 $O/gpu_tests_$TAG.log; cat $O/iter_ng_$TAG.log

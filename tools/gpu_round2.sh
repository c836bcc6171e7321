# GPU call: phase timing of the packed step kernel, thread-map A/B, full GPU test-suite (no -x), bench line
TAG=${1:-r01d}
O=gpurun_out
mkdir -p $O
set -x
GO2_B200_LIB=$PWD/go2_rl_gym_b200/libgo2b200_timing.so timeout 200 python tools/phase_timing.py --num_envs 4096 > $O/phase_timing_$TAG.log 2>&1
timeout 240 python tools/bench_env_step.py --num_envs 4096 8192 --modes P2 P3 8p --steps 100 > $O/env_step_ab_$TAG.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/gpu_tests_$TAG.log
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_line_$TAG.json 2> $O/bench_err_$TAG.log
cat $O/phase_timing_$TAG.log; cat $O/env_step_ab_$TAG.log; tail -30 $O/gpu_tests_$TAG.log; cut -c1-300 $O/bench_line_$TAG.json

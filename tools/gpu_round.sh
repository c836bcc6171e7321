# One GPU call of a build round: A/B of the step kernel's thread maps, the GPU test-suite, the bench line (+ reference arm), ncu captures.
# Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh <tag>      outputs: gpurun_out/*_<tag>.*
TAG=${1:-r01c}
O=gpurun_out
mkdir -p $O
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.log 2>&1
timeout 240 python tools/bench_env_step.py --num_envs 4096 8192 --modes P2 P3 8p --steps 100 > $O/env_step_ab_$TAG.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/gpu_tests_$TAG.log
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_line_$TAG.json 2> $O/bench_err_$TAG.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_ref_line_$TAG.json 2>> $O/bench_err_$TAG.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:step_kernel_packed -s 5 -c 2 -f -o $O/prof_step_$TAG python tools/bench_env_step.py --num_envs 4096 --steps 3 > $O/ncu_step_stdout_$TAG.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 -c 1500 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline > $O/ncu_launch_stdout_$TAG.log 2>&1
cat $O/env_step_ab_$TAG.log; tail -4 $O/gpu_tests_$TAG.log; cut -c1-700 $O/bench_line_$TAG.json; tail -3 $O/bench_err_$TAG.log

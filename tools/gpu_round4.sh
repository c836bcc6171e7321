# GPU call: GPU tests of the trainer + bench line + launch list of one replayed iteration
TAG=${1:-r01i}
O=gpurun_out
mkdir -p $O
set -x
timeout 1200 python -m pytest tests/test_gpu_rl.py -m gpu -q 2>&1 | tail -30 > $O/gpu_tests_$TAG.log
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no_cpu_baseline > $O/bench_line_$TAG.json 2> $O/bench_err_$TAG.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 -c 1500 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline > $O/ncu_launch_stdout_$TAG.log 2>&1
tail -30 $O/gpu_tests_$TAG.log; python -c "
import json;d=json.load(open('$O/bench_line_$TAG.json'));print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','split_ms') if k in d}); print(d['roofline'])"
tail -3 $O/bench_err_$TAG.log

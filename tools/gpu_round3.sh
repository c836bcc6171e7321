# GPU call: GPU test-suite + bench line (+ optional env-step A/B)
TAG=${1:-r01g}
O=gpurun_out
mkdir -p $O
set -x
timeout 240 python tools/bench_env_step.py --num_envs 4096 8192 --modes P2 --steps 100 > $O/env_step_ab_$TAG.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/gpu_tests_$TAG.log
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_line_$TAG.json 2> $O/bench_err_$TAG.log
cat $O/env_step_ab_$TAG.log; tail -30 $O/gpu_tests_$TAG.log; python -c "
import json;d=json.load(open('$O/bench_line_$TAG.json'));print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','split_ms') if k in d}); print(d['roofline'])"
tail -3 $O/bench_err_$TAG.log

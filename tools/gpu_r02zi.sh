timeout 300 python -m pytest tests/test_gpu_env.py tests/test_gpu_rl.py -m gpu -q -x -k "step_host or graph_rollout or host" --tb=short 2>&1 | tail -4
timeout 600 python bench.py --no_other_configs --no_cpu_baseline > gpurun_out/r02zi_bench_line.json 2> gpurun_out/r02zi_bench_err.log; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zi_bench_line.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['split_ms'])
PY
tail -3 gpurun_out/r02zi_bench_err.log

TAG=${1:-r02zl}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -5 | tee $O/${TAG}_gpu_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $O/${TAG}_gpu_tests.txt
timeout 700 python bench.py > $O/${TAG}_bench_line.json 2> $O/${TAG}_bench_err.log; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zl_bench_line.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['split_ms'], d['roofline']['kernel_us'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])
PY

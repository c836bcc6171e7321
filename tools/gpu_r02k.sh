TAG=${1:-r02k}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rf --tb=short 2>&1 | grep -v "^Actor MLP\|^Critic MLP\|Linear(\|ELU(\|^)\|Sequential\|^$" | cut -c1-500 | tail -60 > $O/gpu_tests_$TAG.log
tail -8 $O/gpu_tests_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_line_$TAG.json 2> $O/bench_err_$TAG.log
cut -c1-400 $O/bench_line_$TAG.json; tail -3 $O/bench_err_$TAG.log
python - <<PY
import json
d=json.load(open("$O/bench_line_$TAG.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "split", d["split_ms"], "launches", d["gpu_launches"])
print("gemm", d["roofline_gemm"]["kernel_us_by_mode"], d["roofline_gemm"]["frac"])
for o in d["other_configs"]: print(o["config"]["task"], o["config"]["num_envs_per_gpu"], o.get("value"), o.get("ms_per_step"), o.get("split_ms"), o.get("error"))
PY

TAG=${1:-r02d}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rf --tb=short 2>&1 | grep -v "^Actor MLP\|^Critic MLP\|Linear(\|ELU(\|^)\|Sequential\|^$" | cut -c1-500 | tail -80 > $O/gpu_tests_$TAG.log
tail -30 $O/gpu_tests_$TAG.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_line_$TAG.json 2> $O/bench_err_$TAG.log
cut -c1-1500 $O/bench_line_$TAG.json; tail -3 $O/bench_err_$TAG.log

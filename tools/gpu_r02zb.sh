TAG=${1:-r02zb}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12 | tee $O/${TAG}_gpu_tests.txt
timeout 300 python tools/bench_env_step.py --num_envs 4096 8192 16384 --modes P2 H14 --steps 200 2>&1 | grep "^N=" | tee $O/${TAG}_env_step_modes.txt
timeout 700 python bench.py > $O/${TAG}_bench_line.json 2> $O/${TAG}_bench_err.log
tail -c 600 $O/${TAG}_bench_line.json

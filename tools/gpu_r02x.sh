TAG=${1:-r02x}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_v_env_configs.py -m gpu -q -x -k "H14" --tb=short 2>&1 | tail -15
timeout 300 python tools/bench_env_step.py --num_envs 4096 8192 16384 --modes P2 H14 --steps 200 2>&1 | grep "^N=" | tee $O/${TAG}_env_step_modes.txt

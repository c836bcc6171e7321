# quick A/B of the step kernel: map test + timing
timeout 300 python -m pytest tests/test_gpu_v_env_configs.py -m gpu -q -x -k "H14" --tb=short 2>&1 | tail -3
timeout 300 python tools/bench_env_step.py --num_envs 4096 8192 --modes ${MODES:-H14} --steps 200 2>&1 | grep "^N="

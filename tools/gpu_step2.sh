timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -4
timeout 300 python tools/bench_env_step.py --num_envs 4096 8192 --modes H14 --steps 200 2>&1 | grep "^N="
GO2_B200_LIB=go2_rl_gym_b200/libgo2b200_timing.so timeout 200 python tools/phase_timing.py --mode H14 --raw > gpurun_out/${TAG:-r02z}_phase_cycles_H14.txt 2>&1
tail -8 gpurun_out/${TAG:-r02z}_phase_cycles_H14.txt

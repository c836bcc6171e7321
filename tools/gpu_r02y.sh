TAG=${1:-r02y}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -8 | tee $O/${TAG}_gpu_tests.txt
GO2_B200_LIB=go2_rl_gym_b200/libgo2b200_timing.so timeout 200 python tools/phase_timing.py --mode H14 > $O/${TAG}_phase_cycles_H14.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 1 -f -o $O/prof_step_$TAG python tools/bench_env_step.py --num_envs 4096 --steps 3 --modes H14 > $O/ncu_step_stdout_$TAG.log 2>&1
timeout 600 python bench.py --no_other_configs > $O/${TAG}_bench_line.json 2> $O/${TAG}_bench_err.log
tail -c 1500 $O/${TAG}_bench_line.json

TAG=${1:-r02o}
O=gpurun_out
mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_pair_kernel -s 6 -c 4 -f -o $O/prof_gemm_pair_$TAG python tools/bench_gemm_trainer.py > $O/ncu_gemm_pair_stdout_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 1 -f -o $O/prof_step_$TAG python tools/bench_env_step.py --num_envs 4096 --steps 3 > $O/ncu_step_stdout_$TAG.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 -c 1200 --csv --log-file $O/launches_go2_$TAG.csv python bench.py --steps 2 --warmup 3 --no_other_configs --no_cpu_baseline > $O/ncu_bench_stdout_$TAG.log 2>&1
ls -la $O/*$TAG*

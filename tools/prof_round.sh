set -x
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/gpu_tests_r01b.log
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_line_r01b.json 2> gpurun_out/bench_err_r01b.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref_line_r01b.json 2>> gpurun_out/bench_err_r01b.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline > gpurun_out/ncu_launch_stdout.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 2 -f -o gpurun_out/prof_step_r01b python tools/bench_env_step.py --num_envs 4096 --steps 3 > gpurun_out/ncu_step_stdout.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_persist -s 6 -c 2 -f -o gpurun_out/prof_gemm_r01b python tools/bench_gemm2.py > gpurun_out/ncu_gemm_stdout.log 2>&1
tail -2 gpurun_out/gpu_tests_r01b.log; cat gpurun_out/bench_line_r01b.json | cut -c1-600

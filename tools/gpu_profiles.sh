# GPU call that produces the evidence copied into profiles/: tests, smoke, bench lines (both arms), per-task iteration timings, the ncu
# launch list of a replayed iteration, one ncu --set full capture of the step kernel, the per-phase cycle counts of the packed step kernel.
TAG=${1:-r01k}
O=gpurun_out
mkdir -p $O
set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/gpu_tests_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1
timeout 500 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_line_$TAG.json 2> $O/bench_err_$TAG.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_ref_line_$TAG.json 2>> $O/bench_err_$TAG.log
for cfg in "go2 4096" "go2 8192" "go2_cts 8192" "go2_moe_cts 8192"; do set -- $cfg; timeout 300 python tools/bench_iter.py --task $1 --num_envs $2 --iters 6 2>&1 | grep "^it" | tail -3 | sed "s/^/$1 $2: /" >> $O/iter_tasks_$TAG.log; done
timeout 240 python tools/bench_env_step.py --num_envs 4096 8192 16384 --modes P2 8p --steps 100 > $O/env_step_ab_$TAG.log 2>&1
GO2_B200_LIB=$PWD/go2_rl_gym_b200/libgo2b200_timing.so timeout 200 python tools/phase_timing.py --num_envs 4096 > $O/phase_timing_$TAG.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:step_kernel_packed -s 5 -c 2 -f -o $O/prof_step_$TAG python tools/bench_env_step.py --num_envs 4096 --steps 3 > $O/ncu_step_stdout_$TAG.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1400 -c 1400 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline > $O/ncu_launch_stdout_$TAG.log 2>&1
tail -3 $O/gpu_tests_$TAG.log; cat $O/smoke_$TAG.log | tail -2; cat $O/iter_tasks_$TAG.log; cat $O/env_step_ab_$TAG.log; cat $O/phase_timing_$TAG.log; cut -c1-400 $O/bench_line_$TAG.json; cut -c1-300 $O/bench_ref_line_$TAG.json

# First GPU call of round 2: everything that was built after round 1's GPU budget was spent gets its hardware run and its
# numbers in ONE call.  Usage (repo root, on the GPU box):  bash tools/gpu_round2_first.sh [tag]      outputs: gpurun_out/*_<tag>.*
# Budget: ~12 min of box time.
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.log 2>&1
# 1. the whole GPU suite WITHOUT -x: the staged files (test_gpu_v_*, test_gpu_w_*, test_gpu_x_*) report every failure, not just the first
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -60 > $O/gpu_tests_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1
# 2. headline bench (default library) + reference arm
timeout 500 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_line_$TAG.json 2> $O/bench_err_$TAG.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_ref_line_$TAG.json 2>> $O/bench_err_$TAG.log
# 3. per-task iteration timings, all seven registered tasks
for cfg in "go2 4096" "go2 8192" "go2_cts 8192" "go2_moe_cts 8192" "go2_moe_ng_cts 8192" "go2_ac_moe_cts 8192" "go2_dual_moe_cts 8192" "go2_mcp_cts 8192"; do
  set -- $cfg; timeout 300 python tools/bench_iter.py --task $1 --num_envs $2 --iters 6 2>&1 | grep "^it\|Error\|error" | tail -3 | sed "s/^/$1 $2: /" >> $O/iter_tasks_$TAG.log
done
# 4. the second library build (relaxed solver + state guard): step-kernel A/B against the default build, then the bench line with it
timeout 240 python tools/bench_env_step.py --num_envs 4096 8192 --modes P2 Q4 Q2 8p --steps 100 > $O/env_step_default_$TAG.log 2>&1
GO2_B200_LIB=$PWD/go2_rl_gym_b200/libgo2b200_relaxed.so GO2_RELAXED_CFG=1 timeout 240 python tools/bench_env_step.py --num_envs 4096 8192 --modes P2 --steps 100 > $O/env_step_relaxed_$TAG.log 2>&1
# 5. ncu: launch list of one iteration of the MCP task (new kernels), full capture of the step kernel of the relaxed build
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 -c 2500 --csv --log-file $O/launches_mcp_$TAG.csv python tools/bench_iter.py --task go2_mcp_cts --num_envs 4096 --iters 1 > $O/ncu_mcp_stdout_$TAG.log 2>&1
GO2_B200_LIB=$PWD/go2_rl_gym_b200/libgo2b200_relaxed.so GO2_RELAXED_CFG=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:step_kernel_packed -s 5 -c 2 -f -o $O/prof_step_relaxed_$TAG python tools/bench_env_step.py --num_envs 4096 --steps 3 > $O/ncu_step_relaxed_stdout_$TAG.log 2>&1
tail -30 $O/gpu_tests_$TAG.log; tail -2 $O/smoke_$TAG.log; cat $O/iter_tasks_$TAG.log; cat $O/env_step_default_$TAG.log $O/env_step_relaxed_$TAG.log; cut -c1-400 $O/bench_line_$TAG.json

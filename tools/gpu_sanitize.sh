# compute-sanitizer over the default (H14) step kernel: memcheck on the reference-made fixture replays, racecheck (shared-memory hazards across the
# bar.arrive / bar.sync narrow-phase overlap and the aliased height-scan / impulse rows) and synccheck on a small rollout
O=gpurun_out; mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_env.py -m gpu -q -x -k "replays_reference_golden" 2>&1 | grep -v "^$" | tail -6 | tee $O/r02zf_sanitizer.txt
for tool in racecheck synccheck; do
  echo "== $tool" | tee -a $O/r02zf_sanitizer.txt
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/bench_env_step.py --num_envs 61 --steps 3 --modes H14 2>&1 | grep "^N=\|ERROR SUMMARY\|Race\|Barrier\|hazard" | head -12 | tee -a $O/r02zf_sanitizer.txt
done

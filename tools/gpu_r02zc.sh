for i in 1 2; do
echo "== rolled cold phases (default lib)"; timeout 200 python tools/bench_env_step.py --num_envs 4096 --modes H14 --steps 300 2>&1 | grep "^N="
echo "== unrolled"; GO2_B200_LIB=go2_rl_gym_b200/libgo2b200_unrolled.so timeout 200 python tools/bench_env_step.py --num_envs 4096 --modes H14 --steps 300 2>&1 | grep "^N="
done
timeout 200 python -m pytest tests/test_gpu_v_env_configs.py tests/test_gpu_env.py -m gpu -q -x 2>&1 | tail -2
echo "== critic joins late (default)"; timeout 200 python tools/bench_iter.py --iters 6 2>&1 | grep "^it" | tail -3
echo "== critic joins early"; GO2_CRITIC_JOIN=early timeout 200 python tools/bench_iter.py --iters 6 2>&1 | grep "^it" | tail -3
echo "== go2_moe_cts 8192 late / early"; timeout 300 python tools/bench_iter.py --task go2_moe_cts --num_envs 8192 --iters 4 2>&1 | grep "^it" | tail -2
GO2_CRITIC_JOIN=early timeout 300 python tools/bench_iter.py --task go2_moe_cts --num_envs 8192 --iters 4 2>&1 | grep "^it" | tail -2

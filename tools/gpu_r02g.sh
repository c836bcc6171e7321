TAG=${1:-r02g}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_rl.py -m gpu -q -rf --tb=short -x -k "tensor_core or wgrad or ppo_update or graph_rollout" 2>&1 | grep -v "^Actor MLP\|^Critic MLP\|Linear(\|ELU(\|^)\|Sequential\|^$" | cut -c1-400 | tail -30 > $O/gpu_tests_$TAG.log
tail -5 $O/gpu_tests_$TAG.log
for pdl in 0 1; do echo "PDL=$pdl"; GO2_GEMM_PDL=$pdl timeout 300 python tools/bench_iter.py --task go2 --num_envs 4096 --iters 3 2>&1 | grep "^it\|Warn\|warn" | tail -3; done
GO2_GEMM_PDL=1 timeout 300 python tools/bench_iter.py --task go2_moe_cts --num_envs 4096 --iters 3 2>&1 | grep "^it\|Warn\|warn" | tail -2

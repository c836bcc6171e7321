TAG=${1:-r02e}
O=gpurun_out
mkdir -p $O
timeout 300 python tools/bench_gemm_trainer.py --dbg 2>&1 | grep -v "\[tf32\]" > $O/gemm_trainer_$TAG.log
echo "rc=$?" >> $O/gemm_trainer_$TAG.log
cat $O/gemm_trainer_$TAG.log | cut -c1-330
timeout 600 python -m pytest tests/test_gpu_rl.py -m gpu -q -rf --tb=short -x -k "tensor_core or wgrad or ppo_update" 2>&1 | grep -v "^Actor MLP\|^Critic MLP\|Linear(\|ELU(\|^)\|Sequential\|^$" | cut -c1-400 | tail -30 > $O/gpu_tests_$TAG.log
tail -3 $O/gpu_tests_$TAG.log
timeout 300 python tools/bench_iter.py --task go2 --num_envs 4096 --iters 3 2>&1 | grep "^it" | tail -2

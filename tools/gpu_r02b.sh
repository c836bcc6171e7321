# Round-2 GPU call b: previously failing tests with short tracebacks, the trainer-GEMM table per multiply mode, iteration timing.
TAG=${1:-r02b}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_env.py tests/test_gpu_properties.py tests/test_gpu_rl.py -m gpu -q -rf --tb=short -x --deselect tests/test_gpu_rl.py::test_go2_learns_on_the_gpu 2>&1 | grep -v "^Actor MLP\|^Critic MLP\|Linear(\|ELU(\|^)\|Sequential\|^$" | cut -c1-400 | tail -120 > $O/gpu_tests_$TAG.log
tail -40 $O/gpu_tests_$TAG.log
timeout 600 python tools/bench_gemm_trainer.py > $O/gemm_trainer_$TAG.log 2>&1
cat $O/gemm_trainer_$TAG.log
timeout 300 python tools/bench_iter.py --task go2 --num_envs 4096 --iters 4 2>&1 | grep "^it" > $O/iter_$TAG.log
cat $O/iter_$TAG.log

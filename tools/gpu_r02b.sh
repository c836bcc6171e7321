# Round-2 GPU call: all GPU tests with short tracebacks, the trainer-GEMM table per multiply mode (+ per-role wait cycles), iteration timing.
TAG=${1:-r02b}
PH=${2:-tgi}
O=gpurun_out
mkdir -p $O
case $PH in *t*)
timeout 1500 python -m pytest tests -m gpu -q -rf --tb=short 2>&1 | grep -v "^Actor MLP\|^Critic MLP\|Linear(\|ELU(\|^)\|Sequential\|^$" | cut -c1-500 | tail -150 > $O/gpu_tests_$TAG.log
tail -60 $O/gpu_tests_$TAG.log ;;
esac
case $PH in *g*)
timeout 600 python tools/bench_gemm_trainer.py --dbg > $O/gemm_trainer_$TAG.log 2>&1
cat $O/gemm_trainer_$TAG.log ;;
esac
case $PH in *i*)
for t in go2 go2_moe_cts; do timeout 300 python tools/bench_iter.py --task $t --num_envs 4096 --iters 3 2>&1 | grep "^it" | tail -2 > $O/iter_${t}_$TAG.log; cat $O/iter_${t}_$TAG.log; done ;;
esac

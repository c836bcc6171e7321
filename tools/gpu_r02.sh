# Round-2 GPU call: tests + smoke + bench (both arms) + thread-map A/B + learning curve + ncu captures.   bash tools/gpu_r02.sh <tag> [phases]
# phases: any of t (tests) b (bench) m (step-kernel thread maps) c (learning curve) n (ncu)      outputs: gpurun_out/*_<tag>.*
TAG=${1:-r02a}
PH=${2:-tbmcn}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.log 2>&1
nproc >> $O/smi_$TAG.log
case $PH in *t*)
  timeout 1700 python -m pytest tests -m gpu -q -rf -s 2>&1 | grep -v "^Actor MLP\|^Critic MLP\|Linear(\|ELU(\|^)\|Sequential\|^$" | tail -150 > $O/gpu_tests_$TAG.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1
  tail -5 $O/gpu_tests_$TAG.log; tail -2 $O/smoke_$TAG.log ;;
esac
case $PH in *b*)
  timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_line_$TAG.json 2> $O/bench_err_$TAG.log
  timeout 400 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_ref_line_$TAG.json 2>> $O/bench_err_$TAG.log
  cut -c1-600 $O/bench_line_$TAG.json; tail -3 $O/bench_err_$TAG.log ;;
esac
case $PH in *m*)
  timeout 300 python tools/bench_env_step.py --num_envs 4096 8192 --modes P2 Q4 Q2 8p --steps 100 > $O/env_step_$TAG.log 2>&1
  cat $O/env_step_$TAG.log ;;
esac
case $PH in *c*)
  timeout 300 python tools/train_gpu_curve.py --task go2 --num_envs 4096 --iterations 300 > $O/gpu_learning_curve_go2_$TAG.txt 2> $O/curve_err_$TAG.log
  tail -3 $O/gpu_learning_curve_go2_$TAG.txt; tail -3 $O/curve_err_$TAG.log ;;
esac
case $PH in *n*)
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 1500 --csv --log-file $O/launches_$TAG.csv python tools/bench_iter.py --task go2 --num_envs 4096 --iters 2 > $O/ncu_launch_stdout_$TAG.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 1 -f -o $O/prof_step_$TAG python tools/bench_env_step.py --num_envs 4096 --steps 3 > $O/ncu_step_stdout_$TAG.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_persist -s 40 -c 3 -f -o $O/prof_gemm_$TAG python tools/bench_iter.py --task go2 --num_envs 4096 --iters 1 > $O/ncu_gemm_stdout_$TAG.log 2>&1
  ls -la $O/*.ncu-rep ;;
esac

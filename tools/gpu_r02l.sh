TAG=${1:-r02l}
O=gpurun_out
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6000 -c 3000 --csv --log-file $O/launches_moe_$TAG.csv python tools/bench_iter.py --task go2_moe_cts --num_envs 8192 --iters 1 > $O/ncu_moe_stdout_$TAG.log 2>&1
tail -2 $O/ncu_moe_stdout_$TAG.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2500 -c 1500 --csv --log-file $O/launches_go2_$TAG.csv python tools/bench_iter.py --task go2 --num_envs 4096 --iters 1 > $O/ncu_go2_stdout_$TAG.log 2>&1
tail -2 $O/ncu_go2_stdout_$TAG.log

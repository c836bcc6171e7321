O=gpurun_out; mkdir -p $O; set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/check_dist_graph.py --task go2 2> $O/dist_err1.log | tee $O/dist_graph_check.log
GO2_DIST_GRAPH=0 timeout 300 $TR --master-port 29522 tools/check_dist_graph.py --task go2 2> $O/dist_err2.log | tee -a $O/dist_graph_check.log
timeout 300 $TR --master-port 29523 tools/check_dist_graph.py --task go2_moe_cts --num_envs 2048 2> $O/dist_err3.log | tee -a $O/dist_graph_check.log
GO2_DIST_GRAPH=0 timeout 300 $TR --master-port 29524 tools/check_dist_graph.py --task go2_moe_cts --num_envs 2048 2> $O/dist_err4.log | tee -a $O/dist_graph_check.log
timeout 400 $TR --master-port 29525 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu_r01m.json 2> $O/bench_2gpu_err_r01m.log
cut -c1-260 $O/bench_2gpu_r01m.json; tail -3 $O/dist_err1.log; tail -3 $O/dist_err3.log

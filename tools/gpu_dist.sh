# 2-GPU check of the data-parallel trainer: parameters identical across ranks, captured (GO2_DIST_GRAPH=1) vs segmented all-reduce path.
# NOTE: with GO2_DIST_GRAPH=1 the processes print their result and then hang at exit (graphs holding NCCL kernels are alive when the
# process group is destroyed) -> short timeouts; budget ~2 x 4 min of box time.   Usage: gpurun --gpus 2 -- bash tools/gpu_dist.sh
O=gpurun_out; mkdir -p $O; set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
GO2_DIST_GRAPH=1 timeout 75 $TR --master-port 29521 tools/check_dist_graph.py --task go2 2> $O/dist_err1.log | tee $O/dist_graph_check.log
GO2_DIST_GRAPH=0 timeout 75 $TR --master-port 29522 tools/check_dist_graph.py --task go2 2> $O/dist_err2.log | tee -a $O/dist_graph_check.log
GO2_DIST_GRAPH=1 timeout 90 $TR --master-port 29523 tools/check_dist_graph.py --task go2_moe_cts --num_envs 2048 2> $O/dist_err3.log | tee -a $O/dist_graph_check.log
GO2_DIST_GRAPH=0 timeout 90 $TR --master-port 29524 tools/check_dist_graph.py --task go2_moe_cts --num_envs 2048 2> $O/dist_err4.log | tee -a $O/dist_graph_check.log
timeout 150 $TR --master-port 29525 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu_err.log
cut -c1-260 $O/bench_2gpu.json

# 2-GPU check of the data-parallel trainer: parameters identical across ranks; exchange = the library's NVLink all-reduce kernel inside the update graph
# (default) vs NCCL all-reduces between graph segments (GO2_DIST_P2P=0).   Usage: gpurun --gpus 2 -- bash tools/gpu_dist.sh [tag]
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
L=$O/dist_check_$TAG.log; : > $L
GO2_DIST_P2P=1 timeout 120 $TR --master-port 29521 tools/check_dist_graph.py --task go2 2> $O/dist_err1.log | tee -a $L; echo "rc=$?" >> $L
GO2_DIST_P2P=0 timeout 120 $TR --master-port 29522 tools/check_dist_graph.py --task go2 2> $O/dist_err2.log | tee -a $L; echo "rc=$?" >> $L
GO2_DIST_P2P=1 timeout 150 $TR --master-port 29523 tools/check_dist_graph.py --task go2_moe_cts --num_envs 2048 2> $O/dist_err3.log | tee -a $L; echo "rc=$?" >> $L
GO2_DIST_P2P=0 timeout 150 $TR --master-port 29524 tools/check_dist_graph.py --task go2_moe_cts --num_envs 2048 2> $O/dist_err4.log | tee -a $L; echo "rc=$?" >> $L
grep -i "warn\|error\|Traceback" -A3 $O/dist_err1.log | head -30
timeout 200 $TR --master-port 29525 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu_$TAG.json 2> $O/bench_2gpu_err.log; echo "bench rc=$?"
cut -c1-300 $O/bench_2gpu_$TAG.json; tail -3 $O/bench_2gpu_err.log

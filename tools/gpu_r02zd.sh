# weak scaling of the final build: 8, 4, 2, 1 GPUs back to back on one 8-GPU box (4096 envs per GPU, go2 PPO)
O=gpurun_out; mkdir -p $O
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no_other_configs --no_cpu_baseline > $O/r02zd_bench_line_${n}gpu.json 2> $O/r02zd_bench_${n}gpu_err.log; echo "n=$n rc=$?"
  cut -c1-260 $O/r02zd_bench_line_${n}gpu.json
done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no_other_configs --no_cpu_baseline > $O/r02zd_bench_line_1gpu.json 2> $O/r02zd_bench_1gpu_err.log; cut -c1-260 $O/r02zd_bench_line_1gpu.json

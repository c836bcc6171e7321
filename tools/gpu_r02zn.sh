# the north-star configuration with the final build: go2_moe_cts, 8192 envs per GPU, 8 GPUs = 65 536 envs (BASELINE configs[4])
O=gpurun_out; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --task go2_moe_cts --num_envs 8192 --steps 6 --warmup 3 --no_other_configs --no_cpu_baseline > $O/r02zn_bench_line_8gpu_go2_moe_cts_65536.json 2> $O/r02zn_err.log; echo "rc=$?"
cut -c1-330 $O/r02zn_bench_line_8gpu_go2_moe_cts_65536.json; tail -2 $O/r02zn_err.log | cut -c1-200

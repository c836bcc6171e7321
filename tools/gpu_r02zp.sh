O=gpurun_out; mkdir -p $O
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 -c 1200 --csv --log-file $O/launches_go2_r02zp.csv python bench.py --steps 2 --warmup 3 --no_other_configs --no_cpu_baseline > $O/ncu_bench_stdout_r02zp.log 2>&1; echo "ncu rc=$?"
wc -l $O/launches_go2_r02zp.csv

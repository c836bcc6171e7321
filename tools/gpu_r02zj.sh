timeout 600 python -m pytest tests/test_gpu_cts.py tests/test_gpu_x_moe_heads.py -m gpu -q -x --tb=short 2>&1 | grep -v "^Actor MLP\|^Critic MLP\|Linear(\|ELU(\|^)\|Sequential\|^$" | cut -c1-300 | tail -8
for t in "go2_moe_cts 8192" "go2_moe_cts 4096" "go2_moe_ng_cts 4096"; do set -- $t
  echo "== $1 $2 grouped (default)"; timeout 300 python tools/bench_iter.py --task $1 --num_envs $2 --iters 4 2>&1 | grep "^it" | tail -2
  echo "== $1 $2 GO2_EXPERTS=tc"; GO2_EXPERTS=tc timeout 300 python tools/bench_iter.py --task $1 --num_envs $2 --iters 4 2>&1 | grep "^it" | tail -2
done

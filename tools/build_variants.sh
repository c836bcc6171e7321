#!/bin/bash
# Tuning aid: builds libgo2b200_mb{N}.so with the step kernel's __launch_bounds__ min-blocks set to N (register cap / occupancy trade-off).
set -e
cd "$(dirname "$0")/.."
C=go2_rl_gym_b200/csrc
for mb in "$@"; do
  sed "s/__launch_bounds__(32 \* WARPS_PER_CTA, 4)/__launch_bounds__(32 * WARPS_PER_CTA, $mb)/" $C/env_step.cu > $C/_v_env_step.cu
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -c $C/_v_env_step.cu -o build/_v_env_step_$mb.o
  nvcc -shared -o go2_rl_gym_b200/libgo2b200_mb$mb.so build/_v_env_step_$mb.o build/common.cu.o build/rl_kernels.cu.o build/gemm_tc.cu.o build/cts_kernels.cu.o -lcudart -lcuda
  rm -f $C/_v_env_step.cu
  echo built libgo2b200_mb$mb.so
done

# relaxed-solver variant of the whole library: GO2_B200_LIB=go2_rl_gym_b200/libgo2b200_relaxed.so + sim.b200.limit_relax = 0.5
if [ "$RELAXED" = "1" ]; then
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DGO2_RELAXED_SOLVER=1 -c $C/env_step.cu -o build/_relaxed_env_step.o
  nvcc -shared -o go2_rl_gym_b200/libgo2b200_relaxed.so build/_relaxed_env_step.o build/common.cu.o build/rl_kernels.cu.o build/gemm_tc.cu.o build/cts_kernels.cu.o -lcudart -lcuda
  echo built libgo2b200_relaxed.so
fi

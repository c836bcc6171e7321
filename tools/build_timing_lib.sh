#!/bin/bash
# Tuning aid: libgo2b200_timing.so = the library with per-phase clock64() instrumentation of one CTA of the packed step kernel
# (GO2_PHASE_TIMING=<cta index>); read back by tools/phase_timing.py.  Needs build/*.o of a normal build.
set -e
cd "$(dirname "$0")/.."
C=go2_rl_gym_b200/csrc
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DGO2_PHASE_TIMING=${1:-100} -c $C/env_step.cu -o build/_timing_env_step.o
nvcc -shared -o go2_rl_gym_b200/libgo2b200_timing.so build/_timing_env_step.o build/common.cu.o build/rl_kernels.cu.o build/gemm_tc.cu.o build/cts_kernels.cu.o build/mcp_kernels.cu.o build/dist_kernels.cu.o -lcudart -lcuda
echo built go2_rl_gym_b200/libgo2b200_timing.so

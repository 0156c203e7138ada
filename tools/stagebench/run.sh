#!/bin/bash
# Sweep of tools/stagebench (run under gpurun; ~1 s per point).  Reads: which switch removes the fixed per-stage cost?
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I redtail_b200/csrc/kernels -I include tools/stagebench/swztest.cu -o gpurun_out/swztest && timeout 60 gpurun_out/swztest
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I redtail_b200/csrc/kernels -I include tools/stagebench/stagebench.cu -o gpurun_out/stagebench || exit 1
B=gpurun_out/stagebench
for N in 64 128 192; do
  for R in 1 3; do
    for S in 2 3 4; do
      for M in 0 1 2 8 16 9 25 64 66 72 192 200 208; do
        timeout 20 $B $N $R $S 2000 $M 2>&1 || true
      done
    done
  done
done

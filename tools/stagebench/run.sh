#!/bin/bash
# Sweep of tools/stagebench (run under gpurun; ~1 s per point).  Reads: which switch removes the fixed per-stage cost?
set -e
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I redtail_b200/csrc/kernels -I include tools/stagebench/stagebench.cu -o gpurun_out/stagebench
B=gpurun_out/stagebench
for N in 64 128 192; do
  for R in 1 3; do
    for S in 2 4; do
      for M in 0 1 3 7 8 9 16 32 25; do
        $B $N $R $S 2000 $M || true
      done
    done
  done
done

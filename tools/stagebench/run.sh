#!/bin/bash
# Sweep of tools/stagebench (run under gpurun; < 1 s per point).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I redtail_b200/csrc/kernels -I include tools/stagebench/stagebench.cu -o gpurun_out/stagebench || exit 1
B=gpurun_out/stagebench
run() { timeout 10 $B "$@" 2>&1 || echo "FAILED: $*"; }
echo "== MMA rate: operands resident, one chain, no per-stage commit (mode 13) / with commit (9)"
for N in 32 64 96 128 192 256; do for R in 1 3; do run $N $R 2 2000 13; run $N $R 2 2000 9; done; done
echo "== chunked accumulators + drains, operands resident (mode 1), without TMEM reads (17)"
for N in 64 128 192; do for R in 1 3; do run $N $R 2 2000 1; run $N $R 2 2000 17; done; done
echo "== TMA only, no MMA: hot tile (258), distinct A tiles from a 256 MB tensor (322), decoupled; slots 2 and 4"
for N in 64 128; do for R in 1 3; do for S in 2 4; do run $N $R $S 2000 266; run $N $R $S 2000 330; run $N $R $S 2000 458; done; done; done
echo "== full pipeline: hand-off (0), distinct A (64), distinct A + resident B (192); single chain variants (8, 72, 200)"
for N in 64 128 192; do for R in 1 3; do for S in 2 3 4; do for M in 0 64 192 8 72 200 10 74; do run $N $R $S 2000 $M; done; done; done; done

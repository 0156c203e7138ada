// Micro-benchmark for the open question of round 1 (profiles/r01_ncu_full_kernels.md): the ~1 k clk
// FIXED cost per pipeline stage of conv3d_umma_kernel.  A stripped-down copy of its pipeline -- one TMA producer thread, one
// MMA-issuing warp, eight epilogue warps, S smem slots, 512/N TMEM buffers, persistent one CTA per SM -- with every piece
// switchable, so the cost can be attributed in one GPU run:
//
//   stagebench N ROWS S STAGES MODE      N: accumulator columns (64..256), ROWS: A/B tile pairs per stage (4 K=16 MMAs each,
//                                        kc = 64), S: smem slots, STAGES: stages per CTA, MODE: bit mask
//     1  no TMA: operands are loaded once, the MMA warp never waits for data
//     2  TMA decoupled: the producer streams every stage's loads without waiting for free slots and the MMA warp does not
//        wait for data (timing only: same traffic, no hand-off in either direction)
//     4  no tcgen05.commit on the slot's empty barrier (implies the producer must free-run: use with 2 or 1)
//     8  one accumulation chain for the whole run (no per-stage accumulator switch, no accumulator barriers)
//    16  epilogue warps do not read TMEM (they only hand the buffer back)
//    32  no tcgen05.fence::after_thread_sync in the MMA warp
//    64  every A tile load reads a DIFFERENT 16 KB tile of a 256 MB tensor (L2/HBM traffic like the real kernel; B stays hot)
//   128  B tiles are loaded once per CTA (weights resident): only A streams
//   256  the MMA warp issues no MMAs (TMA / hand-off cost alone)
//
// Build (no library dependencies):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 \
//        -I redtail_b200/csrc/kernels -I include tools/stagebench/stagebench.cu -o gpurun_out/stagebench
// Prints clk per stage (SM clock) and the MMA-only lower bound.  tools/stagebench/run.sh sweeps the interesting points.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>

#include "tma.cuh"

using namespace rt;

constexpr int kThreads = 320;       // producer warp, MMA warp, 8 epilogue warps
constexpr int kKC = 64;             // K elements per tile row (128 B, SWIZZLE_128B)

struct Params { int n, rows, slots, stages, mode, nbuf, atiles; };

__global__ void __launch_bounds__(kThreads, 1)
stage_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ Params p,
             long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int a_bytes = 128 * kKC * 2, b_bytes = p.n * kKC * 2;
    const int stage_bytes = p.rows * (a_bytes + b_bytes);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.slots) * stage_bytes);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* tmem_full = empty_bar + 8;
    uint64_t* tmem_empty = tmem_full + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long t0 = clock64();

    if (warp == 0 && lane == 0) {
        // ---- producer ----
        const int loads = (p.mode & 1) ? p.slots : p.stages;       // mode 1: fill every slot once
        int slot = 0; uint32_t phase = 0;
        for (int k = 0; k < loads; ++k) {
            if (!(p.mode & (1 | 2))) mbar_wait(&empty_bar[slot], phase ^ 1);
            uint8_t* st = smem + static_cast<size_t>(slot) * stage_bytes;
            mbar_arrive_expect_tx(&full_bar[slot], p.rows * (a_bytes + (((p.mode & 128) && k >= p.slots) ? 0 : b_bytes)));
            for (int r = 0; r < p.rows; ++r) {
                const int arow = (p.mode & 64) ? static_cast<int>(((static_cast<long long>(blockIdx.x) * p.stages + k) * p.rows + r) % p.atiles) * 128 : 0;
                tma_load_2d(st + r * a_bytes, &map_a, &full_bar[slot], 0, arow);
                if (!(p.mode & 128) || k < p.slots) tma_load_2d(st + p.rows * a_bytes + r * b_bytes, &map_b, &full_bar[slot], 0, 0);
            }
            if (++slot == p.slots) { slot = 0; phase ^= 1; }
        }
        if (p.mode & 2) {
            // free-running producer: nobody consumed the full barriers, so wait here until the last load of every slot has
            // landed (a CTA must not exit with bulk copies in flight)
            for (int s = 0; s < p.slots && s < loads; ++s) {
                const int n_s = (loads - s + p.slots - 1) / p.slots;
                mbar_wait(&full_bar[s], static_cast<uint32_t>((n_s - 1) & 1));
            }
            __nanosleep(4000);
        }
    } else if (warp == 1) {
        // ---- MMA issuer ----
        const uint32_t idesc = umma_idesc_f16(128, p.n);
        const uint64_t desc_hi = (static_cast<uint64_t>(((8u * 128u) >> 4) | (1u << 14) | (2u << 29))) << 32;
        const uint32_t ring = smem_u32(smem);
        int slot = 0; uint32_t phase = 0; int buf = 0; uint32_t bphase = 0;
        if (p.mode & 1) {                                           // data is loaded once: wait for every slot here
            for (int s = 0; s < p.slots; ++s) mbar_wait(&full_bar[s], 0);
        }
        for (int k = 0; k < p.stages; ++k) {
            if (!(p.mode & 3)) mbar_wait(&full_bar[slot], phase);
            if (!(p.mode & 32)) tc_fence_after();
            if (!(p.mode & 8)) {
                mbar_wait(&tmem_empty[buf], bphase ^ 1);
                if (!(p.mode & 32)) tc_fence_after();
            }
            const uint32_t st = ring + static_cast<uint32_t>(slot) * stage_bytes;
            const uint32_t d = tmem_base + static_cast<uint32_t>((p.mode & 8) ? 0 : buf * p.n);
            if (elect_one_sync()) {
                for (int r = 0; r < p.rows; ++r) {
                    uint32_t xa = ((st + r * a_bytes) >> 4) | (1u << 16);
                    uint32_t xb = ((st + p.rows * a_bytes + r * b_bytes) >> 4) | (1u << 16);
                    for (int kk = 0; kk < kKC / 16 && !(p.mode & 256); ++kk) {
                        const uint32_t acc = (p.mode & 8) ? (k > 0 || r > 0 || kk > 0) : (r > 0 || kk > 0);
                        umma_f16(d, desc_hi | xa, desc_hi | xb, idesc, acc ? 1u : 0u);
                        xa += 2; xb += 2;
                    }
                }
                if (!(p.mode & 4)) umma_commit(&empty_bar[slot]);
                if (!(p.mode & 8)) umma_commit(&tmem_full[buf]);
                else if (k == p.stages - 1) umma_commit(&tmem_full[0]);
            }
            __syncwarp();
            if (++slot == p.slots) { slot = 0; phase ^= 1; }
            if (!(p.mode & 8)) { if (++buf == p.nbuf) { buf = 0; bphase ^= 1; } }
        }
    } else if (warp >= 2) {
        // ---- epilogue (whole warps only: tcgen05.ld is .sync.aligned) ----
        const int q = warp & 3, half = (warp - 2) >> 2;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        float sink = 0.f;
        if (!(p.mode & 8)) {
            int buf = 0; uint32_t bphase = 0;
            for (int k = 0; k < p.stages; ++k) {
                mbar_wait(&tmem_full[buf], bphase);
                tc_fence_after();
                if (!(p.mode & 16)) {
                    for (int c0 = half * (p.n / 2); c0 < (half + 1) * (p.n / 2); c0 += 16) {
                        uint32_t v[16];
                        tmem_ld16(lane_base + static_cast<uint32_t>(buf * p.n + c0), v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) sink += __uint_as_float(v[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                if (++buf == p.nbuf) { buf = 0; bphase ^= 1; }
            }
        } else {
            mbar_wait(&tmem_full[0], 0);
        }
        if (sink == 123.456f) cycles[1] = 1;        // keep the loads alive
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = clock64() - t0;
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

int main(int argc, char** argv) {
    Params p{};
    p.n = argc > 1 ? atoi(argv[1]) : 128;
    p.rows = argc > 2 ? atoi(argv[2]) : 3;
    p.slots = argc > 3 ? atoi(argv[3]) : 2;
    p.stages = argc > 4 ? atoi(argv[4]) : 2000;
    p.mode = argc > 5 ? atoi(argv[5]) : 0;
    if (p.n < 16 || p.n > 256 || p.n % 16 || p.rows < 1 || p.rows > 4 || p.slots < 1 || p.slots > 8) { fprintf(stderr, "bad arguments\n"); return 1; }
    p.nbuf = 512 / p.n > 8 ? 8 : 512 / p.n;
    const size_t stage_bytes = static_cast<size_t>(p.rows) * (128 + p.n) * kKC * 2;
    const size_t smem = p.slots * stage_bytes + 1024 + 512;
    if (smem > 227 * 1024) { fprintf(stderr, "stage ring does not fit (%zu bytes)\n", smem); return 1; }
    __half *a, *b; long long* cyc;
    p.atiles = (p.mode & 64) ? 16384 : 1;
    cudaMalloc(&a, static_cast<size_t>(p.atiles) * 128 * kKC * 2); cudaMalloc(&b, 256 * kKC * 2); cudaMalloc(&cyc, 16);
    cudaMemset(a, 0, static_cast<size_t>(p.atiles) * 128 * kKC * 2); cudaMemset(b, 0, 256 * kKC * 2); cudaMemset(cyc, 0, 16);
    CUtensorMap ma, mb;
    const uint64_t da[2] = {kKC, static_cast<uint64_t>(p.atiles) * 128}, db[2] = {kKC, 256}, st[1] = {kKC * 2};
    const uint32_t ba[2] = {kKC, 128}, bb[2] = {kKC, static_cast<uint32_t>(p.n)};
    if (make_tensor_map(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a, da, st, ba, nullptr, CU_TENSOR_MAP_SWIZZLE_128B) ||
        make_tensor_map(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, b, db, st, bb, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) { fprintf(stderr, "tensor map failed\n"); return 2; }
    cudaFuncSetAttribute(stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int it = 0; it < 3; ++it) stage_kernel<<<sms, kThreads, smem>>>(ma, mb, p, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
    long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_stage = static_cast<double>(c) / p.stages;
    const double mma_floor = p.rows * 4 * (p.n / 2.0);              // K=16 MMA of N columns: N/2 clk at the tensor peak
    const double smem_floor = p.rows * 4 * ((128 + p.n) * 32 / 128.0) + p.rows * (128 + p.n) * 128 / 128.0;   // operand reads + TMA fill at 128 B/clk
    printf("N=%d rows=%d slots=%d stages=%d mode=%d: %.0f clk/stage  (tensor floor %.0f, smem read+fill floor %.0f)\n",
           p.n, p.rows, p.slots, p.stages, p.mode, per_stage, mma_floor, smem_floor);
    return 0;
}

// Hardware question behind the "one A box for all nine (dh, dw) taps" plan: does tcgen05.mma read a SWIZZLE_128B K-major
// operand correctly when the descriptor's start address is NOT 1024-byte aligned (a tile that starts `off` rows into a
// swizzle atom) and when the 8-row group stride (SBO) is not a multiple of 1024 bytes?  The TMA wrote the box with the
// swizzle as a function of the absolute shared-memory address; if the tensor core applies the XOR on absolute address
// bits too, any row offset works and a (th+2) x (tw+2) box serves all nine taps of a 3x3 filter plane.
//
//   swztest        runs a list of (row offset, SBO bytes, base-offset field) triples; prints max |D - reference| for each.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>

#include "tma.cuh"

using namespace rt;

constexpr int kRows = 512;   // rows of the A array in shared memory (64 KB)
constexpr int kN = 64;

__global__ void __launch_bounds__(128, 1)
swz_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int off, int sbo, int bo, float* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sa = smem;
    uint8_t* sb = smem + kRows * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sb + kN * 128);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<64>(slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, kRows * 128 + kN * 128);
        tma_load_2d(sa, &map_a, bar, 0, 0);
        tma_load_2d(sa + 256 * 128, &map_a, bar, 0, 256);
        tma_load_2d(sb, &map_b, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16(128, kN);
        const uint64_t hi_a = (static_cast<uint64_t>((static_cast<uint32_t>(sbo) >> 4) | (1u << 14) | (static_cast<uint32_t>(bo & 7) << 17) | (2u << 29))) << 32;
        const uint64_t hi_b = (static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29))) << 32;
        uint32_t xa = ((smem_u32(sa) + off * 128) >> 4) | (1u << 16);
        uint32_t xb = (smem_u32(sb) >> 4) | (1u << 16);
        for (int kk = 0; kk < 4; ++kk) {
            umma_f16(tmem, hi_a | xa, hi_b | xb, idesc, kk > 0 ? 1u : 0u);
            xa += 2; xb += 2;
        }
        umma_commit(done);
    }
    __syncwarp();
    mbar_wait(done, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < kN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(warp * 32 + (threadIdx.x & 31)) * kN + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem);
}

int main() {
    std::vector<__half> a(kRows * 64), b(kN * 64);
    std::vector<float> af(kRows * 64), bf(kN * 64);
    srand(7);
    for (size_t i = 0; i < a.size(); ++i) { af[i] = static_cast<float>(rand() % 9 - 4); a[i] = __float2half(af[i]); }
    for (size_t i = 0; i < b.size(); ++i) { bf[i] = static_cast<float>(rand() % 9 - 4); b[i] = __float2half(bf[i]); }
    __half *da, *db; float* dout;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 128 * kN * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap ma, mb;
    const uint64_t dima[2] = {64, kRows}, dimb[2] = {64, kN}, st[1] = {128};
    const uint32_t boxa[2] = {64, 256}, boxb[2] = {64, kN};
    if (make_tensor_map(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, da, dima, st, boxa, nullptr, CU_TENSOR_MAP_SWIZZLE_128B) ||
        make_tensor_map(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, dimb, st, boxb, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) { fprintf(stderr, "tensor map failed\n"); return 2; }
    const size_t smem = kRows * 128 + kN * 128 + 1024 + 256;
    cudaFuncSetAttribute(swz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    struct Cfg { int off, sbo, bo; };
    const Cfg cfgs[] = {{0, 1024, 0}, {8, 1024, 0}, {1, 1024, 0}, {2, 1024, 0}, {5, 1024, 0}, {1, 1024, 1}, {5, 1024, 5}, {0, 2048, 0},
                        {0, 1280, 0}, {1, 1280, 0}, {11, 1280, 0}, {22, 1280, 0}, {1, 1280, 1}, {0, 1152, 0}, {3, 2304, 0}, {0, 128, 0}, {3, 256, 0}};
    std::vector<float> out(128 * kN);
    for (const Cfg& c : cfgs) {
        cudaMemset(dout, 0, out.size() * 4);
        swz_kernel<<<1, 128, smem>>>(ma, mb, c.off, c.sbo, c.bo, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("off=%d sbo=%d bo=%d: kernel failed: %s\n", c.off, c.sbo, c.bo, cudaGetErrorString(e)); return 3; }
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        int bad = 0;
        for (int m = 0; m < 128; ++m) {
            const int row = c.off + (m / 8) * (c.sbo / 128) + (m % 8);
            for (int n = 0; n < kN; ++n) {
                double r = 0;
                for (int k = 0; k < 64; ++k) r += static_cast<double>(af[row * 64 + k]) * bf[n * 64 + k];
                const double d = fabs(r - out[m * kN + n]);
                if (d > maxerr) maxerr = d;
                if (d > 1e-3) ++bad;
            }
        }
        printf("off=%2d sbo=%4d base_offset=%d: max err %.1f, %d / %d wrong  -> %s\n", c.off, c.sbo, c.bo, maxerr, bad, 128 * kN, bad ? "FAIL" : "ok");
    }
    return 0;
}

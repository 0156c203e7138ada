// RT_LAYOUT_SPLIT16 producers/consumers outside the convolution kernel:
//   * cost volume written directly in the layout conv3D_1 consumes (same values as CostVolumePlugin kDefault,
//     lib/kernels.cu:50-97: out[d, c] = L[c]; out[d, C+c, y, x] = x >= d ? R[c, y, x-d] : 0), so the 1 GB tensor is written
//     once as fp16 hi/lo channels-last instead of fp32 planes + a re-layout pass;
//   * dense fp32 <-> split16 converters (engine boundaries, tests).
// All HBM-bound: one pass, 16-byte stores, source staged through shared memory for the channel transpose.
#include "common.cuh"

namespace rt {
namespace {

// One CTA: 64 consecutive x of one (n, y) row, all 2C channels, all D disparities.
// The fp32 -> (hi, lo) split and the channel transpose are done ONCE per source element into shared memory, already in
// the output's 16-byte channel-vector order; the D output runs are then plain LDS.128 -> STG.128 copies (the right half
// at a per-disparity pixel shift), so the kernel is bound by HBM writes, not by conversion instructions.
//   smem (uint4 = 8 fp16 channels):  L_hi/L_lo [64][C/8],  R_hi/R_lo [64 + D - 1][C/8]
__global__ void __launch_bounds__(256)
cost_volume_split16_kernel(const float* __restrict__ left, const float* __restrict__ right, __half* __restrict__ out,
                           int c, int h, int w, int disp) {
    extern __shared__ uint4 smv[];
    const int g4 = c >> 3;                                   // channel vectors per side
    const int rw = 64 + disp - 1;
    uint4* l_hi = smv;
    uint4* l_lo = l_hi + 64 * g4;
    uint4* r_hi = l_lo + 64 * g4;
    uint4* r_lo = r_hi + rw * g4;
    const int x0 = blockIdx.x * 64, y = blockIdx.y, n = blockIdx.z;
    const long long hw = static_cast<long long>(h) * w;
    const float* lp = left + static_cast<long long>(n) * c * hw + static_cast<long long>(y) * w;
    const float* rp = right + static_cast<long long>(n) * c * hw + static_cast<long long>(y) * w;
    // stage + split: item = (pixel j, channel vector g); consecutive threads take consecutive pixels (coalesced reads)
    for (int i = threadIdx.x; i < (64 + rw) * g4; i += 256) {
        const bool is_r = i >= 64 * g4;
        const int k = is_r ? i - 64 * g4 : i;
        const int span = is_r ? rw : 64;
        const int g = k / span, j = k % span;
        const int x = is_r ? x0 - (disp - 1) + j : x0 + j;
        const float* src = (is_r ? rp : lp) + static_cast<long long>(g) * 8 * hw + x;
        __align__(16) __half hv[8];
        __align__(16) __half lv[8];
        const bool inb = x >= 0 && x < w;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float v = inb ? __ldg(src + q * hw) : 0.f;
            v = fminf(fmaxf(v, -65504.f), 65504.f);
            const __half hh = __float2half_rn(v);
            hv[q] = hh;
            lv[q] = __float2half_rn((v - __half2float(hh)) * 2048.f);
        }
        (is_r ? r_hi : l_hi)[j * g4 + g] = *reinterpret_cast<const uint4*>(hv);
        (is_r ? r_lo : l_lo)[j * g4 + g] = *reinterpret_cast<const uint4*>(lv);
    }
    __syncthreads();
    const int c2 = 2 * c, groups = 2 * g4;
    const long long plane = static_cast<long long>(disp) * hw * c2;          // elements of one fp16 plane
    __half* hi = out + static_cast<long long>(n) * 2 * plane;
    __half* lo = hi + plane;
    const int items = 64 * groups;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (int d = 0; d < disp; ++d) {
        const long long row = ((static_cast<long long>(d) * h + y) * w + x0) * c2;
        for (int i = threadIdx.x; i < items; i += 256) {
            const int x = i / groups, g = i % groups;
            if (x0 + x >= w) continue;
            uint4 vh, vl;
            if (g < g4) {
                vh = l_hi[x * g4 + g];
                vl = l_lo[x * g4 + g];
            } else if (x0 + x >= d) {
                const int j = x + (disp - 1) - d;
                vh = r_hi[j * g4 + g - g4];
                vl = r_lo[j * g4 + g - g4];
            } else {
                vh = zero;
                vl = zero;
            }
            const long long o = row + static_cast<long long>(x) * c2 + g * 8;
            *reinterpret_cast<uint4*>(hi + o) = vh;
            *reinterpret_cast<uint4*>(lo + o) = vl;
        }
    }
}

// dense fp32 [n][d][c][h][w] -> split16; one CTA per 64 x of one (n, d, y) row.
__global__ void __launch_bounds__(256)
dense_to_split16_kernel(const float* __restrict__ x, __half* __restrict__ out, int d_ext, int c, int h, int w) {
    extern __shared__ float sm[];   // [c][65]
    const int x0 = blockIdx.x * 64, y = blockIdx.y;
    const int d = blockIdx.z % d_ext, n = blockIdx.z / d_ext;
    const long long hw = static_cast<long long>(h) * w;
    const float* src = x + ((static_cast<long long>(n) * d_ext + d) * c) * hw + static_cast<long long>(y) * w;
    for (int i = threadIdx.x; i < c * 64; i += 256) {
        const int ch = i >> 6, xx = i & 63;
        sm[ch * 65 + xx] = (x0 + xx < w) ? __ldg(src + ch * hw + x0 + xx) : 0.f;
    }
    __syncthreads();
    const int groups = c >> 3;
    const long long plane = static_cast<long long>(d_ext) * hw * c;
    __half* hi = out + static_cast<long long>(n) * 2 * plane;
    __half* lo = hi + plane;
    const long long row = ((static_cast<long long>(d) * h + y) * w + x0) * c;
    for (int i = threadIdx.x; i < 64 * groups; i += 256) {
        const int xx = i / groups, g = i % groups;
        if (x0 + xx >= w) continue;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = sm[(g * 8 + k) * 65 + xx];
        const long long o = row + static_cast<long long>(xx) * c + g * 8;
        split_store8(v, hi + o, lo + o);
    }
}

__global__ void __launch_bounds__(256)
split16_to_dense_kernel(const __half* __restrict__ x, float* __restrict__ out, int d_ext, int c, int h, int w) {
    extern __shared__ float sm[];   // [c][65]
    const int x0 = blockIdx.x * 64, y = blockIdx.y;
    const int d = blockIdx.z % d_ext, n = blockIdx.z / d_ext;
    const long long hw = static_cast<long long>(h) * w;
    const long long plane = static_cast<long long>(d_ext) * hw * c;
    const __half* hi = x + static_cast<long long>(n) * 2 * plane;
    const __half* lo = hi + plane;
    const long long row = ((static_cast<long long>(d) * h + y) * w + x0) * c;
    for (int i = threadIdx.x; i < 64 * c; i += 256) {
        const int xx = i / c, ch = i % c;
        if (x0 + xx < w) {
            const long long o = row + static_cast<long long>(xx) * c + ch;
            sm[ch * 65 + xx] = fmaf(__half2float(lo[o]), 1.f / 2048.f, __half2float(hi[o]));
        }
    }
    __syncthreads();
    float* dst = out + ((static_cast<long long>(n) * d_ext + d) * c) * hw + static_cast<long long>(y) * w;
    for (int i = threadIdx.x; i < c * 64; i += 256) {
        const int ch = i >> 6, xx = i & 63;
        if (x0 + xx < w) dst[ch * hw + x0 + xx] = sm[ch * 65 + xx];
    }
}

}  // namespace
}  // namespace rt

using namespace rt;

extern "C" {

int rt_cost_volume_split16(const void* left, const void* right, void* out, int n, int c, int h, int w, int max_disp, void* stream) {
    if (!left || !right || !out || n < 0 || c <= 0 || h <= 0 || w <= 0 || max_disp <= 0) return RT_ERR_ARG;
    if ((2 * c) % 8 != 0) return RT_ERR_UNSUPPORTED;
    if (n == 0) return RT_OK;
    if (h > 65535 || n > 65535) return RT_ERR_UNSUPPORTED;
    if (c % 8 != 0) return RT_ERR_UNSUPPORTED;
    const size_t smem = static_cast<size_t>(c / 8) * (64 + 64 + max_disp - 1) * 2 * sizeof(uint4);
    if (smem > 200 * 1024) return RT_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) RT_CUDA(cudaFuncSetAttribute(cost_volume_split16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dim3 grid((w + 63) / 64, h, n);
    cost_volume_split16_kernel<<<grid, 256, smem, as_stream(stream)>>>(static_cast<const float*>(left), static_cast<const float*>(right),
                                                                      static_cast<__half*>(out), c, h, w, max_disp);
    note_launch("cost_volume_split16");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

int rt_dense_to_split16(const void* x, void* y, int n, int d, int c, int h, int w, void* stream) {
    if (!x || !y || n < 0 || d <= 0 || c <= 0 || h <= 0 || w <= 0) return RT_ERR_ARG;
    if (c % 8 != 0 || c > 512 || static_cast<long long>(n) * d > 65535 || h > 65535) return RT_ERR_UNSUPPORTED;
    if (n == 0) return RT_OK;
    const size_t smem = static_cast<size_t>(c) * 65 * sizeof(float);
    if (smem > 48 * 1024) RT_CUDA(cudaFuncSetAttribute(dense_to_split16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dim3 grid((w + 63) / 64, h, n * d);
    dense_to_split16_kernel<<<grid, 256, smem, as_stream(stream)>>>(static_cast<const float*>(x), static_cast<__half*>(y), d, c, h, w);
    note_launch("dense_to_split16");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

int rt_split16_to_dense(const void* x, void* y, int n, int d, int c, int h, int w, void* stream) {
    if (!x || !y || n < 0 || d <= 0 || c <= 0 || h <= 0 || w <= 0) return RT_ERR_ARG;
    if (c % 8 != 0 || c > 512 || static_cast<long long>(n) * d > 65535 || h > 65535) return RT_ERR_UNSUPPORTED;
    if (n == 0) return RT_OK;
    const size_t smem = static_cast<size_t>(c) * 65 * sizeof(float);
    if (smem > 48 * 1024) RT_CUDA(cudaFuncSetAttribute(split16_to_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dim3 grid((w + 63) / 64, h, n * d);
    split16_to_dense_kernel<<<grid, 256, smem, as_stream(stream)>>>(static_cast<const __half*>(x), static_cast<float*>(y), d, c, h, w);
    note_launch("split16_to_dense");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

}  // extern "C"

// Soft-arg{max,min} over the disparity axis: y[p] = sum_d d * softmax_d(+-x[d, p]).
// Replaces the reference's five cuDNN passes (copy, scale -1, softmax ACCURATE, mul by index, reduce-sum;
// lib/softargmax_plugin.cpp:167-205, workspace 2x input) with ONE pass over the volume.
//
// HBM-bound: algorithmic bytes = D*HW*s read + HW*s write (127.6 MB for NVSmall [96,321*1025] fp32).
// Mapping: a warp owns 32 consecutive pixels (one coalesced 128 B line per disparity plane); the D axis is split
// across the WARPS of a block, each warp running an online (max, sum, weighted-sum) recurrence over its planes with
// kUnroll independent loads in flight; the per-warp partials are merged through shared memory with the usual
// log-sum-exp rescaling.  Splitting D (instead of one thread per pixel looping over all D) multiplies the number
// of bytes in flight per SM, which is what a latency-bound streaming reduction needs.
#include "common.cuh"

namespace rt {
namespace {

constexpr int kWarps = 8;        // D-slices per block
constexpr int kUnroll = 4;

struct Partial { float m, s, ws; };

__device__ __forceinline__ void online_update(Partial& p, float v, float idx) {
    if (v > p.m) {
        const float sc = expf(p.m - v);     // p.m = -inf on first use -> sc = 0
        p.s = p.s * sc + 1.f;
        p.ws = p.ws * sc + idx;
        p.m = v;
    } else {
        // v == -inf while the running max is still -inf would give expf(NaN); such an entry has weight 0
        // (the cuDNN SOFTMAX_ACCURATE pass this replaces returns finite weights for -inf inputs).
        const float e = (v == -INFINITY) ? 0.f : expf(v - p.m);
        p.s += e;
        p.ws += e * idx;
    }
}

template <typename T>
__global__ void __launch_bounds__(kWarps * 32)
softargmax_kernel(const T* __restrict__ x, T* __restrict__ y, int d, int64_t hw, float sign) {
    __shared__ Partial part[kWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t pix = static_cast<int64_t>(blockIdx.x) * 32 + lane;
    const int64_t nofs = static_cast<int64_t>(blockIdx.y) * d * hw;
    Partial p{-INFINITY, 0.f, 0.f};
    if (pix < hw) {
        const T* src = x + nofs + pix;
        int dd = warp;
        for (; dd + (kUnroll - 1) * kWarps < d; dd += kUnroll * kWarps) {
            float v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) v[u] = sign * to_f32(src[static_cast<int64_t>(dd + u * kWarps) * hw]);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) online_update(p, v[u], static_cast<float>(dd + u * kWarps));
        }
        for (; dd < d; dd += kWarps) online_update(p, sign * to_f32(src[static_cast<int64_t>(dd) * hw]), static_cast<float>(dd));
    }
    part[warp][lane] = p;
    __syncthreads();
    if (warp == 0 && pix < hw) {
        float m = part[0][lane].m;
#pragma unroll
        for (int w = 1; w < kWarps; ++w) m = fmaxf(m, part[w][lane].m);
        float s = 0.f, ws = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const Partial q = part[w][lane];
            const float sc = (q.m == -INFINITY) ? 0.f : expf(q.m - m);
            s += q.s * sc;
            ws += q.ws * sc;
        }
        y[static_cast<int64_t>(blockIdx.y) * hw + pix] = from_f32<T>(ws / s);
    }
}

}  // namespace
}  // namespace rt

using namespace rt;

extern "C" int rt_softargmax(int dtype, int is_min, const void* x, void* y, int n, int d, int64_t hw, void* stream) {
    if (!x || !y || n < 0 || d <= 0 || hw < 0) return RT_ERR_ARG;
    if (n == 0 || hw == 0) return RT_OK;
    if (n > 65535) return RT_ERR_UNSUPPORTED;
    dim3 grid(static_cast<unsigned>(ceil_div(hw, 32)), n);
    const float sign = is_min ? -1.f : 1.f;
    if (dtype == RT_F32)
        softargmax_kernel<float><<<grid, kWarps * 32, 0, as_stream(stream)>>>(static_cast<const float*>(x), static_cast<float*>(y), d, hw, sign);
    else if (dtype == RT_F16)
        softargmax_kernel<__half><<<grid, kWarps * 32, 0, as_stream(stream)>>>(static_cast<const __half*>(x), static_cast<__half*>(y), d, hw, sign);
    else
        return RT_ERR_UNSUPPORTED;
    note_launch("softargmax_online");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

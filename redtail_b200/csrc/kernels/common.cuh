// Shared helpers for the sm_100a kernels of libredtail_b200.so.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "redtail_b200.h"

namespace rt {

extern std::atomic<uint64_t> g_launches;
extern const char* g_last_kernel;

inline void note_launch(const char* name, uint64_t n = 1) {
    g_launches.fetch_add(n, std::memory_order_relaxed);
    g_last_kernel = name;
}

inline int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Async error convention of the reference (lib/kernels.cu:19-24): report launch errors, never sync in release.
#define RT_CHECK_LAUNCH()                                  \
    do {                                                   \
        cudaError_t e_ = cudaGetLastError();               \
        if (e_ != cudaSuccess) return static_cast<int>(e_);\
    } while (0)

#define RT_CUDA(call)                                      \
    do {                                                   \
        cudaError_t e_ = (call);                           \
        if (e_ != cudaSuccess) return static_cast<int>(e_);\
    } while (0)

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// e^v - 1 for v <= 0, within 0.9 ulp (checked against float64 over 2.2 M arguments): Cody-Waite reduction v = n ln2 + g,
// degree-7 Taylor polynomial of e^g - 1 on |g| <= 0.35, result = p 2^n + (2^n - 1) with one rounding.  A third of
// expm1f's instruction count -- the fused ELU epilogues evaluate it for every output element of the 3-D stack.
__device__ __forceinline__ float expm1_nonpos(float v) {
    v = fmaxf(v, -80.f);
    const float n = rintf(v * 1.4426950408889634f);
    float g = fmaf(n, -0.693359375f, v);
    g = fmaf(n, 2.12194440e-4f, g);
    float p = 1.f / 5040.f;
    p = fmaf(p, g, 1.f / 720.f);
    p = fmaf(p, g, 1.f / 120.f);
    p = fmaf(p, g, 1.f / 24.f);
    p = fmaf(p, g, 1.f / 6.f);
    p = fmaf(p, g, 0.5f);
    p = fmaf(p * g, g, g);
    const float s = __int_as_float((static_cast<int>(n) + 127) << 23);      // 2^n, n in [-116, 0]
    return fmaf(p, s, s - 1.f);
}

// ELU (alpha = 1), the one definition every kernel of the library uses (lib/elu_plugin.cpp:93,132).
__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1_nonpos(v); }

// Cheaper ELU for the one kernel that is bound by its instruction count (cvconv_combine_kernel writes 127 M elements with
// ELU + fp16 split each): degree-6 Taylor polynomial on (-1/8, 0] (truncation < 1e-9 relative), ex2.approx - 1 below
// (absolute error <= 1.2e-7 on results of magnitude >= 0.117, i.e. ~2 ulp instead of elu1's 0.9).
__device__ __forceinline__ float elu1_approx(float v) {
    float p = 1.f / 720.f;
    p = fmaf(p, v, 1.f / 120.f);
    p = fmaf(p, v, 1.f / 24.f);
    p = fmaf(p, v, 1.f / 6.f);
    p = fmaf(p, v, 0.5f);
    p = fmaf(p, v * v, v);
    float e2;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(v * 1.4426950408889634f));
    const float neg = v > -0.125f ? p : e2 - 1.f;
    return v > 0.f ? v : neg;
}

// Two fp32 -> packed fp16x2 (x0 in the low half), round to nearest, saturating at +-65504.
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float x0, float x1) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
    return r;
}
// RT_LAYOUT_SPLIT16 split of 8 values: hi = fp16(x) (saturating), lo = fp16((x - hi) * 2048); 4 packed words each.
__device__ __forceinline__ void split8_packed(const float (&v)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        h[q] = cvt_f16x2_sat(v[2 * q], v[2 * q + 1]);
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[q]));
        l[q] = cvt_f16x2_sat((v[2 * q] - hf.x) * 2048.f, (v[2 * q + 1] - hf.y) * 2048.f);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// RT_LAYOUT_SPLIT16 element split: x = hi + lo / 2048 (hi = fp16(x), lo = fp16((x - hi) * 2048)), 8 channels = one 16-byte
// vector per plane.
__device__ __forceinline__ void split_store8(const float (&v)[8], __half* hi, __half* lo) {
    uint4 h, l;
    split8_packed(v, h, l);
    *reinterpret_cast<uint4*>(hi) = h;
    *reinterpret_cast<uint4*>(lo) = l;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace rt

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

// ---- packed fp32 pairs (sm_100a FADD2 / FMUL2 / FFMA2: two IEEE round-to-nearest operations per issue slot).  The fused
// epilogues of the tensor-core kernels are bound by their instruction count, not by latency or memory.
#ifndef RT_NO_PACKED_F32
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 p, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
#else   // A/B build (make PACKED=0): the same interface on scalar instructions, bit-identical results
struct f32x2 { float lo, hi; };
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { return f32x2{lo, hi}; }
__device__ __forceinline__ void upk2(f32x2 p, float& lo, float& hi) { lo = p.lo; hi = p.hi; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { return f32x2{__fadd_rn(a.lo, b.lo), __fadd_rn(a.hi, b.hi)}; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { return f32x2{__fsub_rn(a.lo, b.lo), __fsub_rn(a.hi, b.hi)}; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { return f32x2{__fmul_rn(a.lo, b.lo), __fmul_rn(a.hi, b.hi)}; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { return f32x2{__fmaf_rn(a.lo, b.lo, c.lo), __fmaf_rn(a.hi, b.hi, c.hi)}; }
#endif
__device__ __forceinline__ f32x2 bc2(float x) { return pk2(x, x); }
// a[k] += b[k] for an even number of consecutive floats, two per instruction.
template <int N>
__device__ __forceinline__ void add_pairs(float* a, const float* b) {
#pragma unroll
    for (int k = 0; k < N; k += 2) {
        const f32x2 r = add2(pk2(a[k], a[k + 1]), pk2(b[k], b[k + 1]));
        upk2(r, a[k], a[k + 1]);
    }
}

// e^v - 1 for v <= 0, within 0.9 ulp (checked against float64 over 2.2 M arguments): Cody-Waite reduction v = n ln2 + g,
// degree-7 Taylor polynomial of e^g - 1 on |g| <= 0.35, result = p 2^n + (2^n - 1) with one rounding.  A third of
// expm1f's instruction count -- the fused ELU epilogues evaluate it for every output element of the 3-D stack.
// n = rint(v log2 e) by the 1.5 * 2^23 trick and 2^n from the low mantissa bits of the same sum: no FRND / F2I, which run
// on the quarter-rate XU pipe.
__device__ __forceinline__ float expm1_nonpos(float v) {
    v = fmaxf(v, -80.f);
    const float t = fmaf(v, 1.4426950408889634f, 12582912.f);
    const float n = t - 12582912.f;
    float g = fmaf(n, -0.693359375f, v);
    g = fmaf(n, 2.12194440e-4f, g);
    float p = 1.f / 5040.f;
    p = fmaf(p, g, 1.f / 720.f);
    p = fmaf(p, g, 1.f / 120.f);
    p = fmaf(p, g, 1.f / 24.f);
    p = fmaf(p, g, 1.f / 6.f);
    p = fmaf(p, g, 0.5f);
    p = fmaf(p * g, g, g);
    const float s = __int_as_float((__float_as_int(t) << 23) + 0x3f800000);  // 2^n, n in [-116, 0]
    return fmaf(p, s, s - 1.f);
}

// ELU (alpha = 1), the one definition every kernel of the library uses (lib/elu_plugin.cpp:93,132).
__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1_nonpos(v); }

// The same function on two values at once (identical operations in identical order, so identical results).
__device__ __forceinline__ void elu1_x2(float& v0, float& v1) {
    const f32x2 v = pk2(fmaxf(v0, -80.f), fmaxf(v1, -80.f));
    const f32x2 t = fma2(v, bc2(1.4426950408889634f), bc2(12582912.f));
    const f32x2 n = sub2(t, bc2(12582912.f));
    f32x2 g = fma2(n, bc2(-0.693359375f), v);
    g = fma2(n, bc2(2.12194440e-4f), g);
    f32x2 p = bc2(1.f / 5040.f);
    p = fma2(p, g, bc2(1.f / 720.f));
    p = fma2(p, g, bc2(1.f / 120.f));
    p = fma2(p, g, bc2(1.f / 24.f));
    p = fma2(p, g, bc2(1.f / 6.f));
    p = fma2(p, g, bc2(0.5f));
    p = fma2(mul2(p, g), g, g);
    float t0, t1;
    upk2(t, t0, t1);
    const f32x2 s = pk2(__int_as_float((__float_as_int(t0) << 23) + 0x3f800000), __int_as_float((__float_as_int(t1) << 23) + 0x3f800000));
    float r0, r1;
    upk2(fma2(p, s, sub2(s, bc2(1.f))), r0, r1);
    v0 = v0 > 0.f ? v0 : r0;
    v1 = v1 > 0.f ? v1 : r1;
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
        float l0, l1;                            // (v - hi) * 2048: both steps are exact, two lanes per instruction
        upk2(mul2(sub2(pk2(v[2 * q], v[2 * q + 1]), pk2(hf.x, hf.y)), bc2(2048.f)), l0, l1);
        l[q] = cvt_f16x2_sat(l0, l1);
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

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace rt

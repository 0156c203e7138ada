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

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }

// RT_LAYOUT_SPLIT16 element split: x = hi + lo / 2048 (hi = fp16(x), lo = fp16((x - hi) * 2048)), 8 channels = one 16-byte
// vector per plane.
__device__ __forceinline__ void split_store8(const float (&v)[8], __half* hi, __half* lo) {
    __align__(16) __half hv[8];
    __align__(16) __half lv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x = fminf(fmaxf(v[k], -65504.f), 65504.f);
        const __half h = __float2half_rn(x);
        hv[k] = h;
        lv[k] = __float2half_rn((x - __half2float(h)) * 2048.f);
    }
    *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(hv);
    *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(lv);
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace rt

// Cost volume of the stereo nets (CostVolumePlugin, lib/cost_volume_plugin.cpp:122-139).
//
//   kDefault     out[n, d, ch,   y, x] = left [n, ch, y, x]                      (lib/kernels.cu:50-70,  K1)
//                out[n, d, C+ch, y, x] = x >= d ? right[n, ch, y, x-d] : 0       (lib/kernels.cu:72-97,  K2)
//   kCorrelation out[n, d, y, x]       = x >= d ? sum_ch L[ch,y,x]*R[ch,y,x-d] : 0  (lib/kernels.cu:168-200, K4)
//
// kDefault is a pure data mover: 2*C*H*W elements read, D*2C*H*W written (1 036 MB per NVSmall pair), so the only
// roofline is HBM write bandwidth.  The reference issues D strided 4-byte stores per thread from two kernels.
// Here: ONE kernel; a CTA owns a 4096-element run of one flattened (channel) plane, stages it (plus a 256-element
// left halo for the disparity shift) in shared memory with 17 TMA box loads on one mbarrier -- the source is read
// from HBM once, not D times -- and then emits the D output runs with 128-bit stores.  Because every extent of
// the dense plugin tensors is odd (513 x 161), the 16-byte phase of the destination run differs for every (d, ch)
// plane: each run is written as scalar head + aligned 16-byte body + scalar tail, the body re-aligned from shared
// memory with two conflict-free LDS.128 and a funnel shift (the phase is uniform across the CTA).
#include <cstdlib>
#include "common.cuh"
#include "tma.cuh"

namespace rt {
namespace {

constexpr int kTile = 4096;       // elements of one plane per CTA
constexpr int kHalo = 256;        // left halo (max_disp - 1 <= 256), one TMA box
constexpr int kBox = 256;         // TMA box (1-D tensor map, boxDim <= 256)
constexpr int kThreads = 256;

// Simple validation kernel (one thread per source element, loops over d) -- used when TMA constraints fail
// (max_disp > 257 or > 2^32 elements) and by the tests to cross-check the TMA kernel on the device.
template <typename T>
__global__ void cost_volume_simple_kernel(const T* __restrict__ left, const T* __restrict__ right, T* __restrict__ out,
                                          int c, int64_t hw, int w, int disp) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int ch2 = blockIdx.y, n = blockIdx.z;
    if (j >= hw) return;
    const bool is_right = ch2 >= c;
    const int ch = is_right ? ch2 - c : ch2;
    const T* src = (is_right ? right : left) + (static_cast<int64_t>(n) * c + ch) * hw;
    const int x = static_cast<int>(j % w);
    T* dst = out + (static_cast<int64_t>(n) * disp * 2 * c + ch2) * hw + j;
    const int64_t dstride = static_cast<int64_t>(2) * c * hw;
    const T zero = from_f32<T>(0.f);
    for (int d = 0; d < disp; ++d) dst[d * dstride] = is_right ? (x >= d ? src[j - d] : zero) : src[j];
}

template <int Q>
__device__ __forceinline__ uint4 realign(const uint4& lo, const uint4& hi, int sh) {
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint4 o;
    o.x = __funnelshift_r(w[Q + 0], w[Q + 1], sh);
    o.y = __funnelshift_r(w[Q + 1], w[Q + 2], sh);
    o.z = __funnelshift_r(w[Q + 2], w[Q + 3], sh);
    o.w = __funnelshift_r(w[Q + 3], w[Q + 4], sh);
    return o;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
cost_volume_tma_kernel(const __grid_constant__ CUtensorMap map_left, const __grid_constant__ CUtensorMap map_right,
                       T* __restrict__ out, int c, int hw, int w, int disp) {
    constexpr int V = 16 / sizeof(T);                       // elements per 128-bit word
    constexpr int kIters = kTile / (kThreads * V);          // vector stores per thread per disparity
    // TMA needs a 16-byte aligned global start address, so the staged window starts at the plane position rounded
    // DOWN to a 16-byte boundary (a0 elements early) and is one box longer; + V: the realign window may read one
    // 16-byte word past the last consumed element.
    __shared__ __align__(1024) T stage[kHalo + kTile + kBox + V];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int ch2 = blockIdx.y, n = blockIdx.z;
    const bool is_right = ch2 >= c;
    const int ch = is_right ? ch2 - c : ch2;
    const int j0 = blockIdx.x * kTile;                      // first plane element of this CTA
    const int len = min(kTile, hw - j0);

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (tid < V) stage[kHalo + kTile + kBox + tid] = from_f32<T>(0.f);
    __syncthreads();
    // Flat element coordinate of the first wanted element (plane position j0 - kHalo) and its 16-byte aligned floor.
    const long long want = (static_cast<long long>(n) * c + ch) * hw + j0 - kHalo;
    const long long start = want & ~static_cast<long long>(V - 1);
    const int a0 = static_cast<int>(want - start);          // 0 .. V-1
    if (tid == 0) {
        const CUtensorMap* m = is_right ? &map_right : &map_left;
        // Negative / past-the-end boxes are zero-filled by TMA.  Elements that belong to a neighbouring plane are
        // never consumed (the x >= d mask removes exactly those).
        mbar_arrive_expect_tx(&bar, (kHalo + kTile + kBox) * sizeof(T));
#pragma unroll 1
        for (int b = 0; b < (kHalo + kTile + kBox) / kBox; ++b)
            tma_load_1d(stage + b * kBox, m, &bar, static_cast<int>(start + b * kBox));
    }
    const int org = kHalo + a0;                              // stage index of plane element j0

    // Row position (x) of this thread's vector slots, before the per-plane head shift.
    int xg[kIters];
#pragma unroll
    for (int i = 0; i < kIters; ++i) xg[i] = (j0 + (tid + i * kThreads) * V) % w;
    const int xs = (j0 + tid) % w;                           // for head/tail scalars (tid < V)

    mbar_wait(&bar, 0);

    const T zero = from_f32<T>(0.f);
    const int64_t plane0 = (static_cast<int64_t>(n) * disp * 2 * c + ch2) * hw + j0;
    const int64_t dstride = static_cast<int64_t>(2) * c * hw;
#pragma unroll 1
    for (int d = 0; d < disp; ++d) {
        T* dst = out + plane0 + d * dstride;
        const int shift = is_right ? d : 0;
        int head = static_cast<int>((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15) / static_cast<int>(sizeof(T));
        if (head > len) head = len;
        const int nvec = (len - head) / V;
        const int tail = len - head - nvec * V;
        // scalar head and tail
        if (tid < head) {
            const int x = xs;                                // (j0 + tid) % w
            dst[tid] = (!is_right || x >= d) ? stage[org + tid - shift] : zero;
        }
        if (tid < tail) {
            const int jj = head + nvec * V + tid;
            const int x = (j0 + jj) % w;
            dst[jj] = (!is_right || x >= d) ? stage[org + jj - shift] : zero;
        }
        // aligned body
        const int sbase = org + head - shift;              // stage index of body element 0 (>= 0)
        const int mb = (sbase * static_cast<int>(sizeof(T))) & 15;
        const uint4* sp = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(stage) +
                                                         ((sbase * static_cast<int>(sizeof(T))) & ~15));
        const int q = mb >> 2, sh = (mb & 3) * 8;
        uint4* d4 = reinterpret_cast<uint4*>(dst + head);
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
            const int g = tid + i * kThreads;
            if (g < nvec) {
                const uint4 lo = sp[g], hi = sp[g + 1];
                uint4 v;
                switch (q) {
                    case 0: v = realign<0>(lo, hi, sh); break;
                    case 1: v = realign<1>(lo, hi, sh); break;
                    case 2: v = realign<2>(lo, hi, sh); break;
                    default: v = realign<3>(lo, hi, sh); break;
                }
                if (is_right) {
                    int x0 = xg[i] + head;
                    while (x0 >= w) x0 -= w;
                    if (x0 < d || x0 + V > w) {              // some element of this word is masked or wraps a row
                        T* e = reinterpret_cast<T*>(&v);
#pragma unroll
                        for (int k = 0; k < V; ++k) {
                            int xe = x0 + k;
                            while (xe >= w) xe -= w;
                            if (xe < d) e[k] = zero;
                        }
                    }
                }
                d4[g] = v;
            }
        }
    }
}

template <typename T>
int launch_cost_volume(const void* left, const void* right, void* out, int n, int c, int h, int w, int disp,
                       cudaStream_t s) {
    const int64_t hw = static_cast<int64_t>(h) * w;
    const int64_t total = static_cast<int64_t>(n) * c * hw;
    static const bool force_simple = getenv("RT_COSTVOL_SIMPLE") != nullptr;
    const bool tma_ok = !force_simple && disp - 1 <= kHalo && total < (1ll << 31) && hw < (1ll << 31) - kTile &&
                        2 * c <= 65535 && n <= 65535 &&
                        (reinterpret_cast<uintptr_t>(left) & 15) == 0 && (reinterpret_cast<uintptr_t>(right) & 15) == 0;
    if (tma_ok) {
        CUtensorMap ml, mr;
        const uint64_t dims[1] = {static_cast<uint64_t>(total)};
        const uint32_t box[1] = {kBox};
        const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        int rc = make_tensor_map(&ml, dt, 1, left, dims, nullptr, box, nullptr, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc == 0) rc = make_tensor_map(&mr, dt, 1, right, dims, nullptr, box, nullptr, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc == 0) {
            dim3 grid(static_cast<unsigned>(ceil_div(hw, kTile)), 2 * c, n);
            cost_volume_tma_kernel<T><<<grid, kThreads, 0, s>>>(ml, mr, static_cast<T*>(out), c, static_cast<int>(hw), w, disp);
            note_launch("cost_volume_tma");
            RT_CHECK_LAUNCH();
            return RT_OK;
        }
        return rc > 0 ? rc : RT_ERR_UNSUPPORTED;   // fail loudly: a tensor-map failure is a bug, not a reason to fall back
    }
    if (2 * c > 65535 || n > 65535) return RT_ERR_UNSUPPORTED;
    dim3 grid(static_cast<unsigned>(ceil_div(hw, 256)), 2 * c, n);
    cost_volume_simple_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(left), static_cast<const T*>(right),
                                                      static_cast<T*>(out), c, hw, w, disp);
    note_launch("cost_volume_simple");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

// Correlation cost volume.  Tiny (0.25 GFLOP, 37 MB): one thread per (y, x) keeps the C left values in registers
// is not possible for C = 32 x D = 48 without re-reads, so a CTA stages a row segment of both feature maps in
// shared memory (left: 128 px, right: 128 + D px, all C channels) and each thread produces one (d, x) output.
template <typename T, int CMAX>
__global__ void __launch_bounds__(128)
corr_cost_volume_kernel(const T* __restrict__ left, const T* __restrict__ right, T* __restrict__ out,
                        int c, int h, int w, int disp) {
    extern __shared__ float sm[];
    const int xt = blockIdx.x * 128, y = blockIdx.y, n = blockIdx.z;
    const int rw = 128 + disp - 1;                       // right segment: x in [xt - disp + 1, xt + 128)
    float* sl = sm;                                      // [c][128]
    float* sr = sm + c * 128;                            // [c][rw]
    const int64_t hw = static_cast<int64_t>(h) * w;
    const T* lp = left + static_cast<int64_t>(n) * c * hw + static_cast<int64_t>(y) * w;
    const T* rp = right + static_cast<int64_t>(n) * c * hw + static_cast<int64_t>(y) * w;
    for (int i = threadIdx.x; i < c * 128; i += 128) {
        const int ch = i / 128, x = xt + (i % 128);
        sl[i] = x < w ? to_f32(lp[ch * hw + x]) : 0.f;
    }
    for (int i = threadIdx.x; i < c * rw; i += 128) {
        const int ch = i / rw, x = xt - (disp - 1) + (i % rw);
        sr[i] = (x >= 0 && x < w) ? to_f32(rp[ch * hw + x]) : 0.f;
    }
    __syncthreads();
    const int x = xt + threadIdx.x;
    if (x >= w) return;
    T* op = out + (static_cast<int64_t>(n) * disp * h + y) * w + x;
    for (int d = 0; d < disp; ++d) {
        float acc = 0.f;
        if (x >= d) {
            const int ro = threadIdx.x + (disp - 1) - d;
            for (int ch = 0; ch < c; ++ch) acc = fmaf(sl[ch * 128 + threadIdx.x], sr[ch * rw + ro], acc);
        }
        op[static_cast<int64_t>(d) * hw] = from_f32<T>(acc);
    }
}

}  // namespace
}  // namespace rt

using namespace rt;

extern "C" {

int rt_cost_volume(int dtype, const void* left, const void* right, void* out, int n, int c, int h, int w,
                   int max_disp, void* stream) {
    if (!left || !right || !out || n < 0 || c <= 0 || h <= 0 || w <= 0 || max_disp <= 0) return RT_ERR_ARG;
    if (n == 0) return RT_OK;
    if (dtype == RT_F32) return launch_cost_volume<float>(left, right, out, n, c, h, w, max_disp, as_stream(stream));
    if (dtype == RT_F16) return launch_cost_volume<__half>(left, right, out, n, c, h, w, max_disp, as_stream(stream));
    return RT_ERR_UNSUPPORTED;
}

int rt_corr_cost_volume(int dtype, const void* left, const void* right, void* out, int n, int c, int h, int w,
                        int max_disp, void* stream) {
    if (!left || !right || !out || n < 0 || c <= 0 || h <= 0 || w <= 0 || max_disp <= 0) return RT_ERR_ARG;
    if (n == 0) return RT_OK;
    if (h > 65535 || n > 65535) return RT_ERR_UNSUPPORTED;
    const size_t smem = static_cast<size_t>(c) * (128 + 128 + max_disp - 1) * sizeof(float);
    if (smem > 200 * 1024) return RT_ERR_UNSUPPORTED;
    dim3 grid(static_cast<unsigned>(ceil_div(w, 128)), h, n);
    cudaStream_t s = as_stream(stream);
    if (dtype == RT_F32) {
        if (smem > 48 * 1024) RT_CUDA(cudaFuncSetAttribute(corr_cost_volume_kernel<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        corr_cost_volume_kernel<float, 0><<<grid, 128, smem, s>>>(static_cast<const float*>(left), static_cast<const float*>(right), static_cast<float*>(out), c, h, w, max_disp);
    } else if (dtype == RT_F16) {
        if (smem > 48 * 1024) RT_CUDA(cudaFuncSetAttribute(corr_cost_volume_kernel<__half, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        corr_cost_volume_kernel<__half, 0><<<grid, 128, smem, s>>>(static_cast<const __half*>(left), static_cast<const __half*>(right), static_cast<__half*>(out), c, h, w, max_disp);
    } else {
        return RT_ERR_UNSUPPORTED;
    }
    note_launch("corr_cost_volume");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

}  // extern "C"

// Last two layers of the stereo nets as ONE kernel: Conv3DTransposePlugin (32 -> 1 channel, 3x3x3, stride 2) + SlicePlugin +
// SoftargmaxPlugin (lib/conv3d_transpose_plugin.cpp:205-243, lib/softargmax_plugin.cpp:167-205; builders:
// sample_app/nvsmall_1025x321_net.cpp:398-420).  The [96, 321, 1025] fp32 volume between them (126 MB written, then read
// again) never exists: the soft-argmin runs online over the depth planes as the transposed convolution produces them.
//
// Round-1 numbers for the pair: 0.39 ms + 0.05 ms for 633 MB of algorithmic traffic (0.24 of the HBM roofline); the generic
// kernel pays ~900 clk of hand-off per pipeline stage and this layer has almost no math per stage.  Here:
//
//   job    = one 8 x 16 patch of the INPUT lattice (16 x 32 output pixels), all depth planes; persistent CTAs
//   stage  = one input plane i: four TMA boxes [9 rows x 16 positions x 32 ch] (W shift 0 / 1, fp16 hi / lo planes); the two
//            H shifts read the same box at row offsets 0 / 1 KB.  Every input plane crosses HBM -> SM once.
//   MMAs   = 4 shifts x 2 K-steps x (A_hi x [W_hi ; W_lo], N = 32  +  A_lo x W_hi, N = 16) into one 32-column TMEM buffer
//            (8 buffers): columns c = (ph * 2 + pw) * 4 + t_d hold the contribution of input plane i to the output planes
//            2i + t_d (filter plane t_d = 0, 1, 2) at output parity (ph, pw) -- depth-stationary like conv3d_ds.cu, so the
//            sub-pixel formulation needs 16 MMAs per plane instead of 32.  The 8 KB of weights stay resident in shared memory.
//   issue    = two warps on alternate planes, each with its own ring slots and TMEM buffers (see conv3d_ds.cu)
//   epilogue = 16 warps, one (ph, pw) output pixel per thread: plane 2i = T0(i) + T2(i-1) (carried in a register),
//            plane 2i+1 = T1(i); + bias; online (max, sum, weighted sum) soft-argmin update; after the last plane one
//            fp32 disparity per pixel is written.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "conv3d_internal.h"
#include "tma.cuh"

namespace rt {
namespace {

constexpr int kEpiWarps = 16;
constexpr int kMmaWarps = 2;                         // two issuing warps on alternate planes, fixed ring slots (see conv3d_ds.cu)
constexpr int kEpiBase = 1 + kMmaWarps;
constexpr int kThreads = 32 * (kEpiBase + kEpiWarps);
constexpr int kStages = 4;
constexpr int kTh = 8, kTw = 16;
constexpr int kCin = 32;
constexpr int kBox = (kTh + 1) * kTw * kCin * 2;     // 9 216 B: one activation box
constexpr int kStageBytes = 4 * kBox;                // hi w+0 | hi w+1 | lo w+0 | lo w+1
constexpr int kBufCols = 32;                         // D0 (16) | D1 (16)
constexpr int kNumBuf = 8;

struct DsaParams {
    int njobs, tiles_w, tiles_per_sample;
    int depth, in_h, in_w;         // input lattice
    int planes_out, out_h, out_w;  // output planes after the slice, output extent
    int nb;                        // rows of a weight tile: 32 ([W_hi ; W_lo]) or 16
    float sign;                    // -1: soft-argmin, +1: soft-argmax
    float bias;
};

struct DsaPlanImpl {
    DsaParams p{};
    __half* w_dev = nullptr;
    CUtensorMap map_w{};
    bool wlo = true;
    size_t in_elems = 0;
    int smem_bytes = 0;
};

struct Online { float m, s, ws; };

__device__ __forceinline__ void online_update(Online& p, float v, float idx) {     // same recurrence as softargmax.cu
    if (v > p.m) {
        const float sc = expf(p.m - v);
        p.s = p.s * sc + 1.f;
        p.ws = p.ws * sc + idx;
        p.m = v;
    } else {
        const float e = (v == -INFINITY) ? 0.f : expf(v - p.m);
        p.s += e;
        p.ws += e * idx;
    }
}

template <bool WLO>
__global__ void __launch_bounds__(kThreads, 1)
deconv_softargmax_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                         const __grid_constant__ CUtensorMap map_w, const __grid_constant__ DsaParams p, float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* wsm = smem;                                              // 4 resident weight tiles (1 KB slots x 2)
    uint8_t* ring = smem + 4 * 2048;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + kStages * kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + kNumBuf;
    uint64_t* w_bar = tmem_empty + kNumBuf;
    uint32_t* tmem_addr_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        prefetch_tensormap(&map_a_hi);
        prefetch_tensormap(&map_a_lo);
        prefetch_tensormap(&map_w);
        for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < kNumBuf; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarps); }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_addr_slot);      // one CTA per SM (> half of the shared memory)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (*tmem_addr_slot != 0u) __trap();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(w_bar, 4 * p.nb * kCin * 2);
            for (int t = 0; t < 4; ++t) tma_load_2d(wsm + t * 2048, &map_w, w_bar, 0, t * p.nb);
            int stage = 0;
            uint32_t phase = 0;
            for (int job = blockIdx.x; job < p.njobs; job += gridDim.x) {
                const int n = job / p.tiles_per_sample, r = job - n * p.tiles_per_sample;
                const int w0 = (r % p.tiles_w) * kTw, h0 = (r / p.tiles_w) * kTh;
                for (int pl = 0; pl < p.depth; ++pl) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = ring + stage * kStageBytes;
                    mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
                    tma_load_5d(st, &map_a_hi, &full_bar[stage], 0, w0, h0, pl, n);
                    tma_load_5d(st + kBox, &map_a_hi, &full_bar[stage], 0, w0 + 1, h0, pl, n);
                    tma_load_5d(st + 2 * kBox, &map_a_lo, &full_bar[stage], 0, w0, h0, pl, n);
                    tma_load_5d(st + 3 * kBox, &map_a_lo, &full_bar[stage], 0, w0 + 1, h0, pl, n);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < kEpiBase) {
        // ===================== MMA issuers: warp 1 + mw issues the planes with (plane counter & 1) == mw =====================
        const uint32_t mw = static_cast<uint32_t>(warp - 1);
        constexpr uint32_t pitch = kCin * 2;                                       // 64-byte operand rows, SWIZZLE_64B
        constexpr uint64_t desc_hi = (static_cast<uint64_t>(((8u * pitch) >> 4) | (1u << 14) | (4u << 29))) << 32;
        constexpr uint32_t idesc1 = umma_idesc_f16(128, WLO ? 32 : 16);
        constexpr uint32_t idesc2 = umma_idesc_f16(128, 16);
        constexpr uint32_t row16 = (kTw * pitch) >> 4;                             // one lattice row of A, in 16-byte units
        const uint32_t ring_lo = (smem_u32(ring) >> 4) | (1u << 16);
        const uint32_t w_lo = (smem_u32(wsm) >> 4) | (1u << 16);
        mbar_wait(w_bar, 0);
        uint32_t c = 0;                                                            // plane (= stage = chunk) counter over the whole kernel
        for (int job = blockIdx.x; job < p.njobs; job += gridDim.x) {
            for (int pl = 0; pl < p.depth; ++pl, ++c) {
                if ((c & 1u) != mw) continue;
                const uint32_t stage = c & (kStages - 1), buf = c & (kNumBuf - 1);    // even ring / buffer counts: fixed ownership
                mbar_wait(&tmem_empty[buf], ((c / kNumBuf) & 1u) ^ 1u);
                mbar_wait(&full_bar[stage], (c / kStages) & 1u);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t d0 = buf * kBufCols;
                    const uint32_t st = ring_lo + stage * (kStageBytes >> 4);
#pragma unroll
                    for (int ow = 0; ow < 2; ++ow) {
#pragma unroll
                        for (int oh = 0; oh < 2; ++oh) {
                            const uint32_t a_hi = st + ow * (kBox >> 4) + oh * row16;
                            const uint32_t a_lo = a_hi + 2 * (kBox >> 4);
                            const uint32_t b = w_lo + static_cast<uint32_t>(oh * 2 + ow) * (2048u >> 4);
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) {
                                const uint32_t accum = (ow == 0 && oh == 0 && kk == 0) ? 0u : 1u;
                                umma_f16(d0, desc_hi | (a_hi + 2 * kk), desc_hi | (b + 2 * kk), idesc1, accum);
                                umma_f16(d0 + 16, desc_hi | (a_lo + 2 * kk), desc_hi | (b + 2 * kk), idesc2, WLO ? 1u : accum);
                            }
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    umma_commit(&tmem_full[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue: one output pixel (ph, pw) of one lattice position per thread =====================
        const int q = warp & 3;                          // TMEM lane quarter
        const int e = (warp - kEpiBase) >> 2;            // output parity index ph * 2 + pw
        const int m = q * 32 + lane;
        const int hl = m / kTw, wl = m % kTw;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        int buf = 0;
        uint32_t bphase = 0;
        for (int job = blockIdx.x; job < p.njobs; job += gridDim.x) {
            const int n = job / p.tiles_per_sample, r = job - n * p.tiles_per_sample;
            const int oy = 2 * ((r / p.tiles_w) * kTh + hl) + (e >> 1), ox = 2 * ((r % p.tiles_w) * kTw + wl) + (e & 1);
            Online st{-INFINITY, 0.f, 0.f};
            float carry = 0.f;                           // T2 of the previous input plane
            for (int pl = 0; pl < p.depth; ++pl) {
                mbar_wait(&tmem_full[buf], bphase);
                tc_fence_after();
                uint32_t x0[4], x1[4];                  // this pixel's columns e * 4 + t_d of D0 and of D1
                tmem_ld4(lane_base + static_cast<uint32_t>(buf * kBufCols + e * 4), x0);
                tmem_ld4(lane_base + static_cast<uint32_t>(buf * kBufCols + 16 + e * 4), x1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                if (++buf == kNumBuf) { buf = 0; bphase ^= 1; }
                float t[3];
#pragma unroll
                for (int td = 0; td < 3; ++td) t[td] = fmaf(__uint_as_float(x1[td]), 1.f / 2048.f, __uint_as_float(x0[td]));
                if (2 * pl < p.planes_out) online_update(st, p.sign * (t[0] + carry + p.bias), static_cast<float>(2 * pl));
                if (2 * pl + 1 < p.planes_out) online_update(st, p.sign * (t[1] + p.bias), static_cast<float>(2 * pl + 1));
                carry = t[2];
            }
            if (2 * p.depth < p.planes_out) online_update(st, p.sign * (carry + p.bias), static_cast<float>(2 * p.depth));
            if (oy < p.out_h && ox < p.out_w)
                out[(static_cast<long long>(n) * p.out_h + oy) * p.out_w + ox] = st.ws / st.s;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<256>(0u);
}

uint16_t f2h(float f) {
    __half h = __float2half_rn(f);
    uint16_t b;
    memcpy(&b, &h, 2);
    return b;
}
float h2f(uint16_t b) {
    __half h;
    memcpy(&h, &b, 2);
    return __half2float(h);
}

}  // namespace

bool dsa_shape_supported(const rt_conv3d_desc& d) {
    if (!d.transposed || d.fuse_softargmax == 0 || d.precision != RT_PREC_FP32) return false;
    if (d.v != 3 || d.r != 3 || d.s != 3 || d.k != kCin || d.c != 1) return false;
    if (d.stride[0] != 2 || d.stride[1] != 2 || d.stride[2] != 2 || d.pad[0] != 0 || d.pad[1] != 1 || d.pad[2] != 1) return false;
    if (d.in_layout != RT_LAYOUT_SPLIT16 || d.out_layout != RT_LAYOUT_DENSE || d.fuse_elu) return false;
    if (d.in_dims[0] != kCin) return false;
    const int dd = d.in_dims[1], h = d.in_dims[2], w = d.in_dims[3];
    if (d.out_dims[0] != 2 * dd + 1 || d.out_dims[1] != 1 || d.out_dims[2] != 2 * h - 1 || d.out_dims[3] != 2 * w - 1) return false;
    if (d.slice_d < 0 || d.slice_d >= d.out_dims[0]) return false;
    return true;
}

int dsa_plan_init(rt_conv3d_plan* plan, const std::vector<float>& w, const std::vector<float>& bias) {
    const rt_conv3d_desc& d = plan->desc;
    if (!dsa_shape_supported(d) || !get_encode_tiled()) return RT_ERR_UNSUPPORTED;
    DsaPlanImpl* t = new DsaPlanImpl();
    DsaParams& p = t->p;
    p.depth = d.in_dims[1]; p.in_h = d.in_dims[2]; p.in_w = d.in_dims[3];
    p.planes_out = d.out_dims[0] - d.slice_d; p.out_h = d.out_dims[2]; p.out_w = d.out_dims[3];
    p.tiles_w = (p.in_w + kTw - 1) / kTw;
    p.tiles_per_sample = p.tiles_w * ((p.in_h + kTh - 1) / kTh);
    p.sign = d.fuse_softargmax == 1 ? -1.f : 1.f;
    p.bias = bias.empty() ? 0.f : bias[0];
    t->in_elems = static_cast<size_t>(kCin) * p.depth * p.in_h * p.in_w;
    bool wlo = false;
    if (!getenv("REDTAIL_TC_NO_WLO_SKIP")) {
        for (float v : w) {
            const float c = v > 65504.f ? 65504.f : (v < -65504.f ? -65504.f : v);
            if (h2f(f2h(c)) != c) { wlo = true; break; }
        }
    } else wlo = true;
    t->wlo = wlo;
    p.nb = wlo ? 32 : 16;
    // Weight tile of shift (oh, ow): row c = (ph * 2 + pw) * 4 + t_d holds W[k, t_d, t_h, t_w] with the filter taps that output
    // parity (ph, pw) reads at that shift: parity 0 <- (shift 0, tap 1); parity 1 <- (shift 0, tap 2), (shift 1, tap 0)
    // (out = 2 * in - 1 + tap).  Rows with t_d = 3 are zero; rows 16 + c carry W_lo.
    auto tap_of = [](int parity, int shift) { return parity == 0 ? (shift == 0 ? 1 : -1) : (shift == 0 ? 2 : 0); };
    std::vector<uint16_t> pk(static_cast<size_t>(4) * p.nb * kCin, 0);
    for (int oh = 0; oh < 2; ++oh)
        for (int ow = 0; ow < 2; ++ow)
            for (int td = 0; td < 3; ++td)
                for (int ph = 0; ph < 2; ++ph)
                    for (int pw = 0; pw < 2; ++pw) {
                        const int th2 = tap_of(ph, oh), tw2 = tap_of(pw, ow);
                        if (th2 < 0 || tw2 < 0) continue;
                        const int c = (ph * 2 + pw) * 4 + td;
                        for (int k = 0; k < kCin; ++k) {
                            float val = w[((static_cast<size_t>(k) * 3 + td) * 3 + th2) * 3 + tw2];      // KVCRS, C = 1
                            val = val > 65504.f ? 65504.f : (val < -65504.f ? -65504.f : val);
                            const size_t tile = static_cast<size_t>(oh * 2 + ow) * p.nb;
                            const uint16_t hb = f2h(val);
                            pk[(tile + c) * kCin + k] = hb;
                            if (wlo) pk[(tile + 16 + c) * kCin + k] = f2h((val - h2f(hb)) * 2048.f);
                        }
                    }
    if (cudaMalloc(&t->w_dev, pk.size() * 2) != cudaSuccess ||
        cudaMemcpy(t->w_dev, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(t->w_dev);
        delete t;
        return static_cast<int>(cudaErrorMemoryAllocation);
    }
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(kCin), static_cast<uint64_t>(4) * p.nb};
        const uint64_t strides[1] = {static_cast<uint64_t>(kCin) * 2};
        const uint32_t box[2] = {static_cast<uint32_t>(kCin), static_cast<uint32_t>(p.nb)};
        const int rc = make_tensor_map(&t->map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, t->w_dev, dims, strides, box, nullptr,
                                       CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc != 0) { cudaFree(t->w_dev); delete t; return RT_ERR_UNSUPPORTED; }
    }
    t->smem_bytes = 4 * 2048 + kStages * kStageBytes + 1024 /*align slack*/ + 512 /*barriers*/;
    plan->dsa = t;
    return RT_OK;
}

void dsa_plan_destroy(rt_conv3d_plan* plan) {
    DsaPlanImpl* t = static_cast<DsaPlanImpl*>(plan->dsa);
    if (!t) return;
    cudaFree(t->w_dev);
    delete t;
    plan->dsa = nullptr;
}

int dsa_enqueue(const rt_conv3d_plan* plan, int n, const void* x, void* y, cudaStream_t s) {
    const DsaPlanImpl* t = static_cast<const DsaPlanImpl*>(plan->dsa);
    DsaParams p = t->p;
    p.njobs = p.tiles_per_sample * n;
    if (p.njobs == 0) return RT_OK;
    const __half* hi = static_cast<const __half*>(x);
    const __half* lo = hi + t->in_elems;
    CUtensorMap ma_hi, ma_lo;
    {
        const uint64_t dims[5] = {static_cast<uint64_t>(kCin), static_cast<uint64_t>(p.in_w), static_cast<uint64_t>(p.in_h),
                                  static_cast<uint64_t>(p.depth), static_cast<uint64_t>(n)};
        const uint64_t st[4] = {static_cast<uint64_t>(kCin) * 2, static_cast<uint64_t>(kCin) * 2 * p.in_w,
                                static_cast<uint64_t>(kCin) * 2 * p.in_w * p.in_h, static_cast<uint64_t>(t->in_elems) * 4};
        const uint32_t box[5] = {static_cast<uint32_t>(kCin), static_cast<uint32_t>(kTw), static_cast<uint32_t>(kTh + 1), 1u, 1u};
        int rc = make_tensor_map(&ma_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, hi, dims, st, box, nullptr, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc == 0) rc = make_tensor_map(&ma_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, lo, dims, st, box, nullptr, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc != 0) return rc > 0 ? rc : RT_ERR_UNSUPPORTED;
    }
    int grid = num_sms();
    if (grid > p.njobs) grid = p.njobs;
    static bool attr_set[2][64] = {};
    int dev = 0;
    RT_CUDA(cudaGetDevice(&dev));
    const int wi = t->wlo ? 1 : 0;
    if (dev < 0 || dev >= 64 || !attr_set[wi][dev]) {
        if (t->wlo) RT_CUDA(cudaFuncSetAttribute(deconv_softargmax_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        else RT_CUDA(cudaFuncSetAttribute(deconv_softargmax_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (dev >= 0 && dev < 64) attr_set[wi][dev] = true;
    }
    if (t->wlo) deconv_softargmax_kernel<true><<<grid, kThreads, t->smem_bytes, s>>>(ma_hi, ma_lo, t->map_w, p, static_cast<float*>(y));
    else deconv_softargmax_kernel<false><<<grid, kThreads, t->smem_bytes, s>>>(ma_hi, ma_lo, t->map_w, p, static_cast<float*>(y));
    note_launch(t->wlo ? "deconv_softargmax_fp16x2split" : "deconv_softargmax_fp16x2split_w16");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

}  // namespace rt

// 3-D convolution / transposed convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM,
// operands staged by TMA) -- the one dense contraction of the stereo nets (98.9 % of NVSmall's FLOPs).
//
// Replaces cudnnConvolutionForward + cudnnAddTensor (lib/conv3d_plugin.cpp:204-210) and
// cudnnConvolutionBackwardData + addDBiasTo3DConv (lib/conv3d_transpose_plugin.cpp:223-237), plus -- fused in the
// epilogue -- the Transform, Slice, skip-add and ELU passes that follow them in the generated builders.
//
// Formulation: implicit GEMM, one output tile = 128 output positions (a th x tw patch of one output depth plane)
// x Cout channels.  GEMM-M = positions, GEMM-N = Cout, GEMM-K = taps x Cin, walked one (tap, 64-channel block) at a
// time.  Activations are first re-laid out channels-last as fp16 ([N][D][H][W][C], see pack kernel) so that the A
// operand of a tap is a plain TMA box [tw x th x KC] at a shifted coordinate -- zero padding (D, H and W) is TMA
// out-of-bounds fill, never materialised; stride-2 convolutions use TMA traversal strides.  A transposed convolution
// of stride s is computed as a sub-pixel convolution: its s^3 output parities ride along GEMM-N (merged while N <= 64,
// the rest are separate classes), A tiles are loaded once per SHIFT instead of once per tap, and the epilogue does the
// depth-to-space scatter.
//
// L2 -> SM traffic (what bounds the 32-channel layers) is cut three ways: (1) row groups -- the 3 filter rows of a
// stride-1 conv read one TMA box that is 2 patch rows taller, at swizzle-period-aligned row offsets (stride-2 convs: 8-wide
// patches and a box with every input row, see hs2); (2) two H-stacked M tiles per
// job share those halo rows and the weight tiles; (3) activations between two convolutions stay in RT_LAYOUT_SPLIT16
// (channels-last fp16 hi/lo planes written by the epilogue), so no re-layout / Transform / Padding pass ever runs.
//
// Numerics (RT_PREC_FP32): the fp32 tolerance of the plugin path (1e-3 px disparity after 11 chained layers) cannot be
// held by a single fp16/bf16/tf32 product.  Every operand x is split as x = hi + 2^-11 * lo with hi = fp16(x),
// lo = fp16((x - hi) * 2^11) -- 22 significant bits -- and the kernel accumulates the three leading products in fp32
// TMEM:  D0 += A_hi * W_hi ;  D1 += A_hi * W_lo + A_lo * W_hi ;  result = D0 + 2^-11 * D1  (dropped term ~2^-22).
// The first two products share one MMA of N = 2*Cout ([W_hi ; W_lo] stacked along N), so A_hi is read once.
// RT_PREC_FP16 issues only the hi*hi product (the reference's fp16 configs, 1e-2 tolerance).
//
// The tensor core accumulates in fp32 with TRUNCATION (measured: a 108-step chain drifts NVSmall's disparity by 6e-3 px),
// so TMEM chains are kept to one pipeline stage (6-12 K-steps, a "chunk"; shorter, closed inside a stage, on request) and
// the epilogue warps add every chunk into fp32 registers with round-to-nearest while the next chunk runs in another TMEM
// buffer.  What a chunk boundary costs and what it does not (TMEM reads are free) is measured in profiles/r01_ncu_full_kernels.md.
//
// CTA = 11 or 19 warps: warp 0 TMA producer, warps 1-2 MMA issuers on alternate accumulation chunks (elect.sync; warp 1 also owns
// the TMEM allocation, all 512 columns = 2..8 accumulator buffers), 8 or 16 epilogue warps (tcgen05.ld -> registers -> bias / skip /
// S-ReLU / ELU -> dense fp32 or split16 stores; each warp owns one TMEM lane quarter and one of 2 / 4 column groups).  Persistent
// grid (one CTA per SM), static round-robin tile schedule, 2..8-stage smem operand ring.  Forward convolutions with more than 128
// output channels run as <= 128-channel parts (rt_conv3d_create); 32 -> 32 layers and the final 32 -> 1 transposed conv have their own
// depth-stationary kernels (conv3d_ds.cu, deconv_softargmax.cu).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "conv3d_internal.h"
#include "tma.cuh"

namespace rt {
namespace {

constexpr int kMaxTaps = 27;
constexpr int kMaxClasses = 8;
// producer warp + two MMA-issuing warps + epilogue warps.  The two issuing warps take alternate accumulation chunks (each
// chunk has its own TMEM buffer and its own smem stages, so no ordering between them is needed): the ~800 clk one warp spends
// per stage on its own barrier round trips (two mbarrier.try_wait, elect/reconverge, two tcgen05.commit -- measured as the
// 0.11 ms "barrier skeleton" of conv3D_4, profiles/r02_conv_kernel_experiments.md) overlap with the other warp's MMAs.
constexpr int kMmaWarps = 2;
constexpr int kThreadsOf(int epi_warps, int /*mt*/) { return 32 * (1 + kMmaWarps + epi_warps); }
constexpr int kTileM = 128;

// One pipeline stage = one A box + nr weight tiles.  nr > 1 ("row group"): the filter taps dh, dh+1, .. dh+nr-1 of a
// stride-1 convolution read the SAME A box, nr-1 patch rows taller, at row offsets 0, tw, 2*tw (1 KB multiples, so the
// swizzle phase is unchanged) -- the activations cross L2 -> SM once per (dd, dw) instead of once per tap.
struct TapEntry { int8_t dd, dh, dw; uint8_t nr; uint8_t widx[3]; uint8_t pad_; };

// A "class" is one independent implicit GEMM over the input lattice.  Forward conv: one class, all taps.
// Transposed conv of stride s (sub-pixel formulation): output parities that are MERGED ride along GEMM-N (columns =
// [parity_d][parity_h][channel][parity_w]) and share the A tiles; the remaining parities are separate classes.
struct ClassInfo {
    int ntaps;               // number of A shifts of this class
    int ed, eh, ew;          // output parity of the un-merged dims (0 for merged dims / forward conv)
    int dc, hc, wc;          // class-local extent of the GEMM-M lattice
    int tiles_h, tiles_w;
    int job_begin;           // first job index of this class (per sample)
    unsigned long long nr_pack;   // taps[t].nr, 2 bits per tap (kept in a register by the MMA issuer)
    TapEntry taps[kMaxTaps];
};

// Per accumulator column: element offset of its output relative to the thread's (parity-0) output position, its
// channel and its parity index pd*4 + ph*2 + pw (8 = column maps to nothing).  Built on the host, tile independent.
struct ColInfo { long long off; int ch; int pidx; };

struct TcParams {
    int nclasses;
    int jobs_per_sample, njobs;
    int in_s[3];             // A coordinate = idx * in_s + tap offset   (conv stride, 1 for transposed)
    int out_s[3];            // output position = idx * out_s + e        (1 for conv, stride for transposed)
    int th, tw;
    int kc, ncb;             // channels per K block, K blocks per tap
    int cout, nb;            // real output channels, GEMM-N of the B tile (2*cout_pad in fp32 mode)
    int cout_pad;            // accumulator columns (power of two, >= 16)
    int ncols;               // columns that map to an output: cpc << (ld + lh + lw)
    int lc, ld, lh, lw;      // log2 of: channels per parity class, merged parities along D, H, W
    int split;               // 1: hi/lo operands (RT_PREC_FP32), 0: hi only
    int wlo;                 // split mode: 1 = weight tiles are [W_hi ; W_lo] (nb = 2 * cout_pad), 0 = every weight is exactly
                             // representable in fp16 (the reference's trt_weights_fp16.bin) so the A_hi x W_lo product is
                             // identically zero and is not issued: two products (A_hi x W, A_lo x W) instead of three
    int stages;
    int chunk_kb;            // K blocks (stages) accumulated in TMEM before the epilogue adds them up in fp32 registers
    int chunk_rows;          // < gr: a chunk is closed after this many filter rows inside a stage (chunk_kb == 1)
    int gr;                  // taps per row group (1..3)
    int mt;                  // M tiles (128 positions each, stacked along H) per job: they share the A halo and the weights
    int a_bytes, b_bytes, b_tx, stage_bytes;   // b_bytes: 1 KB-rounded slot, b_tx: bytes the weight TMA actually delivers
    int a_tx;                // bytes one A box delivers (a_bytes is its 1 KB-rounded slot)
    int hs2;                 // 1: stride-2 conv with a row group: the A box holds EVERY input row ((th-1)*2 + 3 rows of tw = 8
                             // positions), filter row i starts at box row i and consecutive 8-row groups are 2 box rows apart
    int out_d, out_h, out_w; // output extent actually written
    long long out_sn, out_sc, out_sd;   // element strides of sample, channel, depth in the dense fp32 output
    int fuse_elu;
    int fuse_act;            // 1: per-channel S-ReLU  y = max(v*s1 + b1, 0)*s2 + b2  after bias / skip (parameters in shared memory)
    int out_split;           // 1: y (and skip) are RT_LAYOUT_SPLIT16, 0: dense fp32
    int dbg;                 // timing experiments only (REDTAIL_TC_DEBUG bit mask; results are garbage): 1 = epilogue skips the
                             // tcgen05.ld drains, 2 = no MMAs are issued, 4 = the producer moves no data, 8 = no output phase (bias/ELU/stores);
                             // A/B switches with correct results: 16 / 64 = L2 / L1 prefetch of the skip tensor at tile start, 32 = skip vectors loaded per batch
    int out_c;               // channels of the output tensor (split16 addressing)
    long long out_lo;        // split16: offset of the lo plane, in halves
    int interleave;          // > 0 (experiment, REDTAIL_TC_INTERLEAVE=1): job index = tile * nclasses + class over the common tile
                             // grid (grid_d x grid_th x grid_tw), so that the un-merged parity classes of a transposed conv, which
                             // read the SAME input tile, run at the same time on different SMs and share it in L2
    int grid_d, grid_th, grid_tw;
    ClassInfo cls[kMaxClasses];
};

// ---------------------------------------------------------------------------------------------------------------
// Pack: dense fp32 [n][..] with arbitrary (d, c) strides -> channels-last fp16 hi (and lo) [n][D][H][W][C].
// One CTA handles 64 consecutive w of one (n, d, h) row for all C channels through a padded smem transpose.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_split_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, int d_ext, int c_src, int c_ext,
                  int h_ext, int w_ext, long long s_n, long long s_d, long long s_c, long long o_sn) {
    extern __shared__ float tile[];                    // [c_ext][65]; channels c_src .. c_ext-1 are zero padding
    const int w0 = blockIdx.x * 64;
    const int h = blockIdx.y;
    const int d = blockIdx.z % d_ext, n = blockIdx.z / d_ext;
    const float* src = x + n * s_n + d * s_d + static_cast<long long>(h) * w_ext;
    for (int i = threadIdx.x; i < c_ext * 64; i += 256) {
        const int c = i >> 6, w = i & 63;
        tile[c * 65 + w] = (w0 + w < w_ext && c < c_src) ? __ldg(src + c * s_c + w0 + w) : 0.f;
    }
    __syncthreads();
    const long long pix0 = (static_cast<long long>(d) * h_ext + h) * w_ext + w0;      // within the sample
    const int groups = c_ext >> 3;                     // 8 channels (16 bytes of fp16) per thread-item
    for (int i = threadIdx.x; i < 64 * groups; i += 256) {
        const int w = i / groups, g = i % groups;
        if (w0 + w >= w_ext) continue;
        __align__(16) __half hv[8];
        __align__(16) __half lv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = tile[(g * 8 + k) * 65 + w];
            v = fminf(fmaxf(v, -65504.f), 65504.f);    // fp16 range (documented limit of the split scheme)
            const __half hh = __float2half_rn(v);
            hv[k] = hh;
            lv[k] = __float2half_rn((v - __half2float(hh)) * 2048.f);
        }
        const long long o = n * o_sn + (pix0 + w) * c_ext + g * 8;
        *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(hv);
        if (lo) *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(lv);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Main kernel.
// ---------------------------------------------------------------------------------------------------------------
struct JobCoord { int cls, n, d, h0, w0; bool empty; };

__device__ __forceinline__ JobCoord decode_job(const TcParams& p, int job) {
    JobCoord j;
    j.n = job / p.jobs_per_sample;
    int r = job - j.n * p.jobs_per_sample;
    j.empty = false;
    if (p.interleave) {
        j.cls = r % p.nclasses;
        r /= p.nclasses;
        j.w0 = (r % p.grid_tw) * p.tw;
        r /= p.grid_tw;
        j.h0 = (r % p.grid_th) * (p.th * p.mt);
        j.d = r / p.grid_th;
        const ClassInfo& ci = p.cls[j.cls];
        j.empty = j.d >= ci.dc || j.h0 >= ci.hc || j.w0 >= ci.wc;     // this class has no tile here (its lattice is one shorter)
        return j;
    }
    int c = 0;
    while (c + 1 < p.nclasses && r >= p.cls[c + 1].job_begin) ++c;
    j.cls = c;
    r -= p.cls[c].job_begin;
    const int tw_n = p.cls[c].tiles_w, th_n = p.cls[c].tiles_h;
    j.w0 = (r % tw_n) * p.tw;
    r /= tw_n;
    j.h0 = (r % th_n) * (p.th * p.mt);
    j.d = r / th_n;
    return j;
}

// All tcgen05.mma of one pipeline stage, fully unrolled (the issuing warp executes ~5 uniform-datapath instructions per
// MMA and nothing else between them).  Measured on the round-1 kernel with ncu's warp-state sampling: the issuing warp
// spent 75 % of its time in its own scalar control code -- 125 instructions per stage plus 63 per filter row at ~6 clk each
// for a lone warp -- while the tensor pipe was 48 % active; the "fixed ~1.1 k clk per stage" of round 1 was this code, not
// a hardware hand-off cost.  nr (1..3 filter rows in this stage) is the only run-time quantity left.
template <int KC16, int MT, bool SPLIT, int COUT_PAD, int ACC_COLS>
__device__ __forceinline__ void umma_issue_stage(int nr, uint32_t lo_a, uint32_t lo_l, uint32_t lo_b, uint32_t grp_a16,
                                                 uint32_t grp_b16, uint32_t tile_a16, uint64_t desc_hi, uint64_t desc_b_hi,
                                                 uint32_t idesc_full, uint32_t idesc_half, uint32_t d0, uint32_t acc_first, uint32_t acc_first_lo) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < nr) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const uint32_t xa = lo_a + i * grp_a16 + mt * tile_a16, xl = lo_l + i * grp_a16 + mt * tile_a16;
                const uint32_t xb = lo_b + i * grp_b16;
                const uint32_t dt = d0 + mt * ACC_COLS;
#pragma unroll
                for (int kk = 0; kk < KC16; ++kk) {          // +32 bytes = one K=16 slice inside the swizzle atom
                    umma_f16(dt, desc_hi | (xa + 2 * kk), desc_b_hi | (xb + 2 * kk), idesc_full, (i == 0 && kk == 0) ? acc_first : 1u);
                    if (SPLIT) umma_f16(dt + COUT_PAD, desc_hi | (xl + 2 * kk), desc_b_hi | (xb + 2 * kk), idesc_half, (i == 0 && kk == 0) ? acc_first_lo : 1u);
                }
            }
        }
    }
}

// Epilogue register budget: each of the 8 epilogue warps owns one TMEM lane quarter (warp_id % 4) and one half of the
// output channels: EW = 8 epilogue warps -> 2 column groups, EW = 16 -> 4 column groups (the output phase -- bias, skip, ELU,
// fp16 split, stores -- and the per-chunk register adds are the throughput limit of the transposed-conv layers, which
// have 8x less MMA work per output; twice the warps halve that work per warp).  CPH = cout_pad / (EW / 4) columns of D0
// (and of D1 in split mode) per thread.
template <int CPH, bool SPLIT, int MT, int EW>
// One CTA per SM.  The register file is four 16 K banks, one per SM sub-partition: with 10 warps three of them share a bank
// (<= 170 registers per thread, not 65536 / 320 = 204 -- a kernel compiled for 200 fails to launch with "too many resources"),
// with 18 warps five do (<= 102); __launch_bounds__ lets ptxas apply exactly that rule.
__global__ void __launch_bounds__(kThreadsOf(EW, MT), 1)
conv3d_umma_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_w, const __grid_constant__ TcParams p,
                   const float* __restrict__ bias, const ColInfo* __restrict__ cols, const float* __restrict__ skip,
                   float* __restrict__ out, const float* __restrict__ act) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // Carve: [stages x stage_bytes] operand ring (1024-aligned) | barriers | tmem address | bias
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* ring = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + static_cast<size_t>(p.stages) * p.stage_bytes);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* tmem_full = empty_bar + 8;
    uint64_t* tmem_empty = tmem_full + 8;
    uint32_t* tmem_addr_slot = reinterpret_cast<uint32_t*>(tmem_empty + 8);
    float* s_bias = reinterpret_cast<float*>(tmem_addr_slot + 4);
    ColInfo* s_col = reinterpret_cast<ColInfo*>(s_bias + 128);
    float* s_act = reinterpret_cast<float*>(s_col + 128);     // [4][128]: s1, b1, s2, b2 (fuse_act)

    constexpr int kCoutPad = (EW / 4) * CPH;
    constexpr int kThreads = kThreadsOf(EW, MT);
    constexpr int kEpiBase = 1 + kMmaWarps;                  // first epilogue warp (warps 1, 2 issue the MMAs)
    constexpr int kAccCols = SPLIT ? 2 * kCoutPad : kCoutPad;   // TMEM columns of one accumulator buffer (= p.nb)
    constexpr int kBufCols = MT * kAccCols;                     // one buffer = the accumulators of the MT tiles of a job
    constexpr int kNumBuf = (512 / kBufCols) > 8 ? 8 : (512 / kBufCols);   // buffers the MMA warp may run ahead by
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&map_a_hi);
        if (SPLIT) prefetch_tensormap(&map_a_lo);
        prefetch_tensormap(&map_w);
        // A slot is free again when BOTH issuing warps are done with it: the one that issued its MMAs (tcgen05.commit) and the one
        // that only observed its full barrier (plain arrive) -- see the observer branch of the issuing loop.  With sub-stage chunks
        // only warp 1 issues (and observes), so one arrival frees the slot.
        const uint32_t slot_arrivals = p.chunk_rows >= p.gr ? kMmaWarps : 1;
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], slot_arrivals); }
        for (int i = 0; i < kNumBuf; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EW); }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < kCoutPad; i += kThreads) {
        s_bias[i] = i < p.cout ? bias[i] : 0.f;
        s_col[i] = cols[i];
    }
    if (p.fuse_act)
        for (int i = threadIdx.x; i < 4 * 128; i += kThreads) s_act[i] = ((i & 127) < p.cout) ? act[(i >> 7) * p.cout + (i & 127)] : ((i >> 7) & 1 ? 0.f : 1.f);
    if (warp == 1) tmem_alloc<512>(tmem_addr_slot);      // one CTA per SM: take the whole TMEM (2 accumulator buffers)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // The CTA owns all 512 TMEM columns, so the allocation starts at lane 0 / column 0: TMEM addresses below are
    // plain compile-time offsets (keeps them in uniform registers; no per-MMA broadcast).  Trap if that ever fails.
    if (*tmem_addr_slot != 0u) __trap();
    constexpr uint32_t tmem_base = 0u;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int job = blockIdx.x; job < p.njobs; job += gridDim.x) {
                const JobCoord jc = decode_job(p, job);
                if (jc.empty) continue;
                const ClassInfo& ci = p.cls[jc.cls];
                for (int t = 0; t < ci.ntaps; ++t) {
                    const TapEntry te = ci.taps[t];
                    const int cw = jc.w0 * p.in_s[2] + te.dw;
                    const int chh = jc.h0 * p.in_s[1] + te.dh;
                    const int cd = jc.d * p.in_s[0] + te.dd;
                    for (int cb = 0; cb < p.ncb; ++cb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* st = ring + static_cast<size_t>(stage) * p.stage_bytes;
                        if (p.dbg & 4) {
                            mbar_arrive(&full_bar[stage]);
                            if (++stage == p.stages) { stage = 0; phase ^= 1; }
                            continue;
                        }
                        mbar_arrive_expect_tx(&full_bar[stage], p.a_tx * (SPLIT ? 2 : 1) + te.nr * p.b_tx);
                        tma_load_5d(st, &map_a_hi, &full_bar[stage], cb * p.kc, cw, chh, cd, jc.n);
                        if (SPLIT) tma_load_5d(st + p.a_bytes, &map_a_lo, &full_bar[stage], cb * p.kc, cw, chh, cd, jc.n);
                        for (int i = 0; i < te.nr; ++i)
                            tma_load_2d(st + p.a_bytes * (SPLIT ? 2 : 1) + i * p.b_bytes, &map_w, &full_bar[stage], 0,
                                        (static_cast<int>(te.widx[i]) * p.ncb + cb) * p.nb);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp < kEpiBase) {
        // ===================== MMA issuers: warp 1 + mw issues the chunks with (chunk index & 1) == mw =====================
        const int mw = warp - 1;
        // The tensor core accumulates in fp32 with truncation, so a long accumulation chain drifts (measured: 6e-3 px
        // of disparity over NVSmall).  Chains are therefore kept short: one chunk = p.chunk_kb K blocks (one filter
        // tap in split mode) accumulates in TMEM, then the epilogue warps add it into fp32 registers (round-to-nearest)
        // while the next chunk runs in the other TMEM buffer.
        // All 32 lanes run the (warp-uniform) control flow so that addresses and descriptors live in uniform registers;
        // lane 0 issues.  Per MMA the only arithmetic is one 32-bit add on the descriptor's start-address field.
        {
            const uint32_t pitch = p.kc * 2;                                   // bytes per operand row = swizzle span
            const uint32_t swz = p.kc == 64 ? 2u : (p.kc == 32 ? 4u : 6u);     // SWIZZLE_128B / 64B / 32B
            // descriptor high word: SBO (8 rows) | version 1 (bit 46) | swizzle (bits 61-63); low word: start>>4 | LBO=1
            const uint64_t desc_b_hi = (static_cast<uint64_t>(((8u * pitch) >> 4) | (1u << 14) | (swz << 29))) << 32;
            // A in hs2 mode: the next 8-row group (= next patch row, tw = 8) is two box rows further
            const uint64_t desc_hi = (static_cast<uint64_t>((((p.hs2 ? 16u : 8u) * pitch) >> 4) | (1u << 14) | (swz << 29))) << 32;
            const uint32_t idesc_half = umma_idesc_f16(kTileM, kCoutPad);
            const uint32_t idesc_full = (SPLIT && !p.wlo) ? idesc_half : umma_idesc_f16(kTileM, kAccCols);
            const uint32_t ring_addr = smem_u32(ring);
            const uint32_t stage_bytes = p.stage_bytes, a_bytes = p.a_bytes;
            const int kc16 = p.kc >> 4, stages = p.stages, chunk_kb = p.chunk_kb, chunk_rows = p.chunk_rows, ncb = p.ncb, gr = p.gr;
            const uint32_t grp_a16 = (static_cast<uint32_t>(p.tw) * pitch) >> 4;   // one patch row of A, in 16-byte units
            const uint32_t grp_b16 = static_cast<uint32_t>(p.b_bytes) >> 4;
            const uint32_t tile_a16 = (static_cast<uint32_t>(kTileM) * pitch) >> 4;    // next M tile of the job (th patch rows down)
            int stage = 0;
            uint32_t phase = 0;
            int buf = 0;
            uint32_t bphase = 0;
            if (chunk_rows >= gr) {
                // ---- fast path: a chunk is `chunk_kb` whole stages (every layer of the nets; sub-stage chunks below) ----
                uint32_t st_lo = (ring_addr >> 4) | (1u << 16);                  // descriptor low word of the current slot
                uint32_t chunk_idx = 0;
                const uint32_t st_lo0 = st_lo, stage16 = stage_bytes >> 4, a16 = a_bytes >> 4;
                for (int job = blockIdx.x; job < p.njobs; job += gridDim.x) {
                    int cls = 0;
                    if (p.interleave) {
                        const JobCoord jc = decode_job(p, job);
                        if (jc.empty) continue;
                        cls = jc.cls;
                    } else {
                        const int r = job % p.jobs_per_sample;
                        while (cls + 1 < p.nclasses && r >= p.cls[cls + 1].job_begin) ++cls;
                    }
                    const int nkb = p.cls[cls].ntaps * ncb;
                    unsigned long long nrp = p.cls[cls].nr_pack;
                    int cb_left = 0, nr = 1, kb_in_chunk = 0;
                    for (int kb = 0; kb < nkb; ++kb) {
                        if (cb_left == 0) {
                            nr = gr == 1 ? 1 : static_cast<int>(nrp & 3ull);
                            nrp >>= 2;
                            cb_left = ncb;
                        }
                        --cb_left;
                        const bool first = kb_in_chunk == 0;
                        const bool close = (++kb_in_chunk == chunk_kb) || (kb == nkb - 1);
                        if ((chunk_idx & 1u) != static_cast<uint32_t>(mw)) {    // the other issuing warp's chunk: only keep the counters in step
                            // ... and observe the stage's full barrier: an mbarrier has ONE phase bit, so a warp that waited for
                            // phase p+1 of a slot before phase p had completed would see "complete" at once (parity aliasing).
                            // Watching every phase of every slot in order keeps both warps within one phase of each barrier.
                            mbar_wait(&full_bar[stage], phase);
                            // ... and tell the producer so.  Observing alone is not enough: a slot whose consecutive fills all belong
                            // to the OTHER warp (even ring sizes with one-stage chunks, or two-stage chunks on a 3-slot ring) could be
                            // consumed and refilled twice while this warp is held up elsewhere (its own MMAs queue behind the other
                            // warp's); it would then wait for the fill after next, run two fills late for the rest of the kernel
                            // and hang on the last one, which never comes (seen as an intermittent hang of the ResNet-18 decoder
                            // layers at batch 4..8).  The slot's empty barrier therefore needs this warp's arrival too.
                            if (lane == 0) mbar_arrive(&empty_bar[stage]);
                            if (close) {
                                kb_in_chunk = 0; ++chunk_idx;
                                if (++buf == kNumBuf) { buf = 0; bphase ^= 1; }
                            }
                            if (++stage == stages) { stage = 0; phase ^= 1; st_lo = st_lo0; } else st_lo += stage16;
                            continue;
                        }
                        if (first) mbar_wait(&tmem_empty[buf], bphase ^ 1);      // epilogue drained this buffer
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (elect_one_sync()) {
                            const uint32_t d0 = tmem_base + static_cast<uint32_t>(buf * kBufCols);
                            const uint32_t lo_l = st_lo + a16, lo_b = st_lo + (SPLIT ? 2u : 1u) * a16;
                            const uint32_t acc_first = first ? 0u : 1u;
                            // D1 is initialised by the first MMA (N covers D0|D1) when the weight tile carries W_lo rows; without them
                            // (fp16-exact weights) the A_lo x W_hi product is the first to touch D1 and must overwrite it
                            const uint32_t acc_first_lo = p.wlo ? 1u : acc_first;
                            if (p.dbg & 2) {
                            } else if (kc16 == 4)
                                umma_issue_stage<4, MT, SPLIT, kCoutPad, kAccCols>(nr, st_lo, lo_l, lo_b, grp_a16, grp_b16, tile_a16, desc_hi, desc_b_hi, idesc_full, idesc_half, d0, acc_first, acc_first_lo);
                            else if (kc16 == 2)
                                umma_issue_stage<2, MT, SPLIT, kCoutPad, kAccCols>(nr, st_lo, lo_l, lo_b, grp_a16, grp_b16, tile_a16, desc_hi, desc_b_hi, idesc_full, idesc_half, d0, acc_first, acc_first_lo);
                            else
                                umma_issue_stage<1, MT, SPLIT, kCoutPad, kAccCols>(nr, st_lo, lo_l, lo_b, grp_a16, grp_b16, tile_a16, desc_hi, desc_b_hi, idesc_full, idesc_half, d0, acc_first, acc_first_lo);
                            umma_commit(&empty_bar[stage]);                      // slot free once these MMAs retire
                            if (close) umma_commit(&tmem_full[buf]);             // chunk complete
                        }
                        __syncwarp();
                        if (close) {
                            kb_in_chunk = 0; ++chunk_idx;
                            if (++buf == kNumBuf) { buf = 0; bphase ^= 1; }
                        }
                        if (++stage == stages) { stage = 0; phase ^= 1; st_lo = st_lo0; } else st_lo += stage16;
                    }
                }
            } else if (mw == 0)                          // sub-stage chunks (short chains on request): one issuing warp
            for (int job = blockIdx.x; job < p.njobs; job += gridDim.x) {
                const JobCoord jc = decode_job(p, job);
                if (jc.empty) continue;
                const int nkb = p.cls[jc.cls].ntaps * ncb;
                // A chunk (one TMEM accumulation chain) is closed after `chunk_rows` filter rows inside a stage when chains
                // are shorter than a stage, else after `chunk_kb` whole stages; the epilogue walks the same sequence.
                bool open = false;
                int rows_in_chunk = 0, kb_in_chunk = 0;
                uint32_t d0 = 0;
                // (no divisions or dependent constant loads between the MMAs of consecutive stages: the tensor pipe idles
                // whenever this warp is not blocked issuing, so the per-stage scalar path is kept minimal)
                unsigned long long nrp = p.cls[jc.cls].nr_pack;
                int cb_left = 0, nr = 1;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (cb_left == 0) {                                        // next tap (row group): ncb stages share its nr
                        nr = gr == 1 ? 1 : static_cast<int>(nrp & 3ull);
                        nrp >>= 2;
                        cb_left = ncb;
                    }
                    --cb_left;
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t st_addr = ring_addr + static_cast<uint32_t>(stage) * stage_bytes;
                    const uint32_t lo_a = (st_addr >> 4) | (1u << 16);
                    const uint32_t lo_l = ((st_addr + a_bytes) >> 4) | (1u << 16);
                    const uint32_t lo_b = ((st_addr + (SPLIT ? 2u : 1u) * a_bytes) >> 4) | (1u << 16);
                    ++kb_in_chunk;
                    // The rows of a stage are issued in segments that end where a chunk ends; with chains of a stage or
                    // longer (chunk_rows == gr) a segment is the whole stage.
                    for (int i0 = 0; i0 < nr;) {
                        if (!open) {
                            mbar_wait(&tmem_empty[buf], bphase ^ 1);           // epilogue drained this buffer
                            tc_fence_after();
                            d0 = tmem_base + static_cast<uint32_t>(buf * kBufCols);
                        }
                        const int i1 = chunk_rows < gr ? min(nr, i0 + chunk_rows - rows_in_chunk) : nr;
                        const bool last_seg = i1 == nr;
                        const bool close = chunk_rows < gr ? true : (kb_in_chunk == chunk_kb || kb == nkb - 1);
                        if (elect_one_sync()) {
                            for (int i = i0; i < i1; ++i) {                    // taps of the row group share the A box
#pragma unroll
                                for (int mt = 0; mt < MT; ++mt) {              // M tiles of the job share the weight tile
                                    uint32_t xa = lo_a + i * grp_a16 + mt * tile_a16, xl = lo_l + i * grp_a16 + mt * tile_a16;
                                    uint32_t xb = lo_b + i * grp_b16;
                                    const uint32_t dt = d0 + mt * kAccCols;
                                    for (int kk = 0; kk < kc16; ++kk) {
                                        umma_f16(dt, desc_hi | xa, desc_b_hi | xb, idesc_full, (open || i > i0 || kk > 0) ? 1u : 0u);
                                        if (SPLIT) umma_f16(dt + kCoutPad, desc_hi | xl, desc_b_hi | xb, idesc_half, (p.wlo || open || i > i0 || kk > 0) ? 1u : 0u);
                                        xa += 2; xl += 2; xb += 2;             // +32 bytes = one K=16 slice inside the swizzle atom
                                    }
                                }
                            }
                            if (last_seg) umma_commit(&empty_bar[stage]);      // slot free once these MMAs retire
                            if (close) umma_commit(&tmem_full[buf]);           // chunk complete
                        }
                        __syncwarp();
                        rows_in_chunk += i1 - i0;
                        i0 = i1;
                        open = !close;
                        if (close) {
                            rows_in_chunk = 0; kb_in_chunk = 0;
                            if (++buf == kNumBuf) { buf = 0; bphase ^= 1; }
                        }
                    }
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access (warp_id % 4)
        const int half = (warp - kEpiBase) >> 2;                // which group of the output channels (2 groups for EW = 8, 4 for EW = 16)
        const int m = q * 32 + lane;                     // tile row = output position within the patch
        const int hl = m / p.tw, wl = m % p.tw;
        const int col0 = half * CPH;
        int buf = 0;
        uint32_t bphase = 0;
        const int chunk_kb = p.chunk_kb, chunk_rows = p.chunk_rows;
        for (int job = blockIdx.x; job < p.njobs; job += gridDim.x) {
            const JobCoord jc = decode_job(p, job);
            if (jc.empty) continue;
            const ClassInfo& ci = p.cls[jc.cls];
            const int nkb = ci.ntaps * p.ncb;
            // Experiment (REDTAIL_TC_DEBUG=16): ask L2 for this tile's skip lines at the start of the tile.  Measured on NVSmall:
            // deconv3D_2 0.514 -> 0.568 ms, deconv3D_1 0.182 -> 0.196 ms WITH the prefetch -- the extra requests cost more than
            // the (already overlapped) latency they hide; off by default.
            if (skip != nullptr && p.out_split && (p.dbg & (16 | 64))) {
                const __half* sk16 = reinterpret_cast<const __half*>(skip);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const int hi_ = jc.h0 + mt * p.th + hl, wi_ = jc.w0 + wl;
                    if (hi_ < ci.hc && wi_ < ci.wc) {
                        const int bd = jc.d * p.out_s[0] + ci.ed, bh = hi_ * p.out_s[1] + ci.eh, bw = wi_ * p.out_s[2] + ci.ew;
                        const long long rowbase = jc.n * p.out_sn + ((static_cast<long long>(bd) * p.out_h + bh) * p.out_w + bw) * p.out_c;
#pragma unroll
                        for (int k0 = 0; k0 < CPH; k0 += 8) {
                            const ColInfo c0 = s_col[col0 + k0];
                            const int q8 = c0.pidx;
                            if (q8 < 8 && bd + (q8 >> 2) < p.out_d && bh + ((q8 >> 1) & 1) < p.out_h && bw + (q8 & 1) < p.out_w) {
                                if (p.dbg & 64) { prefetch_l1(sk16 + rowbase + c0.off); prefetch_l1(sk16 + p.out_lo + rowbase + c0.off); }
                                else { prefetch_l2(sk16 + rowbase + c0.off); prefetch_l2(sk16 + p.out_lo + rowbase + c0.off); }
                            }
                        }
                    }
                }
            }
            float acc0[MT][CPH];
            float acc1[MT][SPLIT ? CPH : 1];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int k = 0; k < CPH; ++k) acc0[mt][k] = 0.f;
                if (SPLIT) {
#pragma unroll
                    for (int k = 0; k < CPH; ++k) acc1[mt][k] = 0.f;
                }
            }
            for (int kb0 = 0; kb0 < nkb; kb0 += chunk_kb) {
              // chunks inside this (group of) stage(s): see the MMA issuer
              const int nsub = chunk_rows < p.gr ? (ci.taps[kb0 / p.ncb].nr + chunk_rows - 1) / chunk_rows : 1;
              for (int sub = 0; sub < nsub; ++sub) {
                mbar_wait(&tmem_full[buf], bphase);
                tc_fence_after();
                constexpr int LW = CPH >= 16 ? 16 : 8;   // columns per tcgen05.ld
                constexpr int BW = (CPH >= 64 && SPLIT) ? 16 : (CPH >= 32 ? 32 : CPH);   // columns in flight per wait (64 + 64 register
                                                         // accumulators leave room for 16 + 16 loaded values): all loads of a batch are issued
                                                         // before the single tcgen05.wait::ld (the drain is latency-bound)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                          static_cast<uint32_t>(buf * kBufCols + mt * kAccCols + col0);
#pragma unroll
                    for (int b0 = 0; b0 < CPH; b0 += BW) {
                        uint32_t v0[BW], v1[SPLIT ? BW : 1];
                        if (p.dbg & 1) {
#pragma unroll
                            for (int k = 0; k < BW; ++k) { v0[k] = 0; if (SPLIT) v1[k] = 0; }
                        } else {
#pragma unroll
                            for (int c0 = 0; c0 < BW; c0 += LW) {
                                tmem_ld<LW>(trow + b0 + c0, *reinterpret_cast<uint32_t(*)[LW]>(&v0[c0]));
                                if (SPLIT) tmem_ld<LW>(trow + kCoutPad + b0 + c0, *reinterpret_cast<uint32_t(*)[LW]>(&v1[c0]));
                            }
                            tmem_ld_wait();
                        }
                        add_pairs<BW>(&acc0[mt][b0], reinterpret_cast<const float*>(v0));
                        if (SPLIT) add_pairs<BW>(&acc1[mt][b0], reinterpret_cast<const float*>(v1));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                if (++buf == kNumBuf) { buf = 0; bphase ^= 1; }
              }
            }
            if (SPLIT) {                                 // result = D0 + 2^-11 D1; the D1 registers are dead from here on
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int k = 0; k < CPH; k += 2)
                        upk2(fma2(pk2(acc1[mt][k], acc1[mt][k + 1]), bc2(1.f / 2048.f), pk2(acc0[mt][k], acc0[mt][k + 1])), acc0[mt][k], acc0[mt][k + 1]);
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
            const int hi_ = jc.h0 + mt * p.th + hl, wi_ = jc.w0 + wl;
            if (hi_ < ci.hc && wi_ < ci.wc && !(p.dbg & 8)) {
                const int bd = jc.d * p.out_s[0] + ci.ed, bh = hi_ * p.out_s[1] + ci.eh, bw = wi_ * p.out_s[2] + ci.ew;
                const long long rowbase = p.out_split
                    ? jc.n * p.out_sn + ((static_cast<long long>(bd) * p.out_h + bh) * p.out_w + bw) * p.out_c
                    : jc.n * p.out_sn + bd * p.out_sd + static_cast<long long>(bh) * p.out_w + bw;
                // which of the (up to 8) merged output parities of this row fall inside the output
                uint32_t vmask = 0;
#pragma unroll
                for (int q8 = 0; q8 < 8; ++q8) {
                    const bool ok = bd + (q8 >> 2) < p.out_d && bh + ((q8 >> 1) & 1) < p.out_h && bw + (q8 & 1) < p.out_w;
                    vmask |= (ok ? 1u : 0u) << q8;
                }
                if (p.out_split) {
                    // RT_LAYOUT_SPLIT16 output: a batch of 8 columns = 8 consecutive channels of one output position
                    // (guaranteed by the column order chosen on the host) -> one 16-byte store per fp16 plane.
                    __half* oh16 = reinterpret_cast<__half*>(out);
                    const __half* sk16 = reinterpret_cast<const __half*>(skip);
                    // every skip vector of this row first (the accumulators' D1 half is dead here, registers are free): one
                    // exposed memory latency per tile instead of one per batch of 8 columns
                    constexpr bool kUpFront = CPH <= 32;          // 64 columns per thread: the vectors would not fit in registers
                    uint4 sh[kUpFront ? CPH / 8 : 1], sl[kUpFront ? CPH / 8 : 1];
                    const bool up_front = kUpFront && !(p.dbg & 32);
                    if (skip && up_front) {
#pragma unroll
                        for (int k0 = 0; k0 < CPH; k0 += 8) {
                            const ColInfo c0 = s_col[col0 + k0];
                            if ((vmask >> c0.pidx) & 1u) {
                                sh[k0 / 8] = __ldg(reinterpret_cast<const uint4*>(sk16 + rowbase + c0.off));
                                sl[k0 / 8] = __ldg(reinterpret_cast<const uint4*>(sk16 + p.out_lo + rowbase + c0.off));
                            }
                        }
                    }
#pragma unroll
                    for (int k0 = 0; k0 < CPH; k0 += 8) {
                        const ColInfo c0 = s_col[col0 + k0];
                        if ((vmask >> c0.pidx) & 1u) {
                            const long long idx = rowbase + c0.off;
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                v[j] = acc0[mt][k0 + j] + s_bias[c0.ch + j];
                            }
                            if (skip) {
                                if (!up_front) {
                                    sh[kUpFront ? k0 / 8 : 0] = __ldg(reinterpret_cast<const uint4*>(sk16 + idx));
                                    sl[kUpFront ? k0 / 8 : 0] = __ldg(reinterpret_cast<const uint4*>(sk16 + p.out_lo + idx));
                                }
                                const __half* hh = reinterpret_cast<const __half*>(&sh[kUpFront ? k0 / 8 : 0]);
                                const __half* ll = reinterpret_cast<const __half*>(&sl[kUpFront ? k0 / 8 : 0]);
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] += fmaf(__half2float(ll[j]), 1.f / 2048.f, __half2float(hh[j]));
                            }
                            if (p.fuse_act) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int ch = c0.ch + j;
                                    v[j] = fmaf(fmaxf(fmaf(v[j], s_act[ch], s_act[128 + ch]), 0.f), s_act[256 + ch], s_act[384 + ch]);
                                }
                            }
                            if (p.fuse_elu) {
#pragma unroll
                                for (int j = 0; j < 8; j += 2) elu1_x2(v[j], v[j + 1]);
                            }
                            uint4 hv, lv;
                            split8_packed(v, hv, lv);
                            *reinterpret_cast<uint4*>(oh16 + idx) = hv;
                            *reinterpret_cast<uint4*>(oh16 + p.out_lo + idx) = lv;
                        }
                    }
                } else
                // Batches of 8 columns: all skip-tensor loads of a batch are issued before any is consumed.
#pragma unroll
                for (int k0 = 0; k0 < CPH; k0 += 8) {
                    long long idx[8];
                    float sk[8];
                    int ch[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const ColInfo ci2 = s_col[col0 + k0 + j];
                        const bool ok = (vmask >> ci2.pidx) & 1u;          // pidx == 8 -> never set
                        ch[j] = ci2.ch;
                        idx[j] = ok ? rowbase + ci2.off : -1;
                        sk[j] = (ok && skip) ? __ldg(skip + idx[j]) : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (idx[j] >= 0) {
                            float val = acc0[mt][k0 + j] + s_bias[ch[j]] + sk[j];
                            if (p.fuse_act) val = fmaf(fmaxf(fmaf(val, s_act[ch[j]], s_act[128 + ch[j]]), 0.f), s_act[256 + ch[j]], s_act[384 + ch[j]]);
                            if (p.fuse_elu) val = elu1(val);
                            out[idx[j]] = val;
                        }
                    }
                }
            }
            }   // mt
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// Host side.
// ---------------------------------------------------------------------------------------------------------------
struct TcPlan {
    TcParams p{};
    __half* w_dev = nullptr;          // packed weights [taps][ncb][nb][kc]
    CUtensorMap map_w{};
    int cin = 0;                      // channels of the packed activations (K-block multiple)
    int cin_src = 0;                  // channels of the caller's tensor (<= cin; the rest is zero padding)
    int in_d = 0, in_h = 0, in_w = 0; // input spatial extent
    long long in_sn = 0, in_sd = 0, in_sc = 0;   // dense fp32 input strides (sample, depth, channel)
    size_t in_elems = 0;              // per sample, = D*H*W*C
    int smem_bytes = 0;
    bool in_split = false;            // x arrives as RT_LAYOUT_SPLIT16 (no pack pass)
    ColInfo* d_cols = nullptr;        // device copy of the column table
    float* act_dev = nullptr;         // [4][cout] S-ReLU parameters (fuse_act)
};

uint16_t f2h_bits(float f) {
    __half h = __float2half_rn(f);
    uint16_t b;
    memcpy(&b, &h, 2);
    return b;
}
float h2f_bits(uint16_t b) {
    __half h;
    memcpy(&h, &b, 2);
    return __half2float(h);
}

}  // namespace

// Input channels as the kernel sees them: 16, 32 or a multiple of 64.
static int tc_padded_cin(int cin) { return cin <= 16 ? 16 : (cin <= 32 ? 32 : (cin + 63) / 64 * 64); }

// Coverage of the tensor-core tiles (shape only; no device work).
bool tc_shape_supported(const rt_conv3d_desc& d) {
    const int cin = d.transposed ? d.k : d.c, cout = d.transposed ? d.c : d.k;
    // K blocks are 16, 32 or 64 channels wide.  A dense fp32 input is re-laid out by the pack pass anyway, which zero-pads
    // the channels up to the next block size (so the reference's 1..8-channel unit-test tensors run on the tensor cores
    // too); a split16 input is consumed in place and must already have a block-sized channel count.
    if (cin < 1) return false;
    if (d.in_layout == RT_LAYOUT_SPLIT16 && tc_padded_cin(cin) != cin) return false;
    if (cout < 1) return false;
    if (cout > 128 && (d.transposed || cout % 8 != 0)) return false;      // > 128: <= 128-channel parts (forward convs, rt_conv3d_create)
    if (d.v > 3 || d.r > 3 || d.s > 3) return false;
    for (int i = 0; i < 3; ++i)
        if (d.stride[i] > 2 || d.stride[i] < 1) return false;
    if (d.out_layout == RT_LAYOUT_SPLIT16 && cout % 8 != 0) return false;   // 16-byte channel vectors
    if (d.precision == RT_PREC_SIMT) return false;
    return true;
}

int tc_plan_init(rt_conv3d_plan* plan, const std::vector<float>& w, const std::vector<float>& /*bias*/) {
    const rt_conv3d_desc& d = plan->desc;
    const bool tr = d.transposed != 0;
    const int cin_src = plan->cin, cout = plan->cout;
    const int cin = tc_padded_cin(cin_src);            // zero-padded by the pack pass (weights of the padding are zero too)
    if (!tc_shape_supported(d)) return RT_ERR_UNSUPPORTED;
    const bool out_split = d.out_layout == RT_LAYOUT_SPLIT16;
    // This plan may be one <= 128-channel part of a wider convolution: it then writes channels [coff, coff + cout) of ctot.
    const int ctot = plan->out_c_total > 0 ? plan->out_c_total : cout, coff = plan->out_c_total > 0 ? plan->out_c_offset : 0;
    if (tr && plan->out_c_total > 0) return RT_ERR_UNSUPPORTED;
    // Channels per parity class (power of two) and which stride-2 output parities are merged into GEMM-N.
    int cpc = 1;
    while (cpc < cout) cpc *= 2;
    if (cpc > 128) return RT_ERR_UNSUPPORTED;
    int lmerge[3] = {0, 0, 0};
    if (tr) {
        int ncol = cpc;
        // Merged columns are capped at 64: N = 128 (x2 in split mode) leaves only two TMEM buffers and a 256-column
        // epilogue per chunk, which measured slower (deconv3D_2: 0.95 ms at 128 vs 0.73 ms at 64).
        int max_n = 64;
        if (const char* e = getenv("REDTAIL_TC_MAXN")) max_n = atoi(e) >= 16 ? atoi(e) : 64;
        const int order[3] = {2, 1, 0};                  // merge W first (adjacent outputs), then H, then D
        for (int oi = 0; oi < 3; ++oi) {
            const int i = order[oi];
            if (d.stride[i] == 2 && ncol * 2 <= max_n) { lmerge[i] = 1; ncol *= 2; }
        }
    }
    const int ncols = cpc << (lmerge[0] + lmerge[1] + lmerge[2]);
    int cout_pad = 16;
    while (cout_pad < ncols) cout_pad *= 2;
    const bool split = d.precision == RT_PREC_FP32;
    // fp16-exact weights (an fp16 weight file, or fp32 values that happen to be representable): W_lo == 0 everywhere.
    bool wlo = false;
    if (split && !getenv("REDTAIL_TC_NO_WLO_SKIP")) {
        for (float v : w) {
            const float c = v > 65504.f ? 65504.f : (v < -65504.f ? -65504.f : v);
            if (h2f_bits(f2h_bits(c)) != c) { wlo = true; break; }
        }
    } else wlo = split;
    const int nb = (split && wlo) ? 2 * cout_pad : cout_pad;
    if (nb > 256) return RT_ERR_UNSUPPORTED;
    if (d.v > 3 || d.r > 3 || d.s > 3) return RT_ERR_UNSUPPORTED;
    for (int i = 0; i < 3; ++i)
        if (d.stride[i] > 2) return RT_ERR_UNSUPPORTED;
    if (!get_encode_tiled()) return RT_ERR_UNSUPPORTED;

    TcPlan* t = new TcPlan();
    TcParams& p = t->p;
    t->cin = cin;
    t->cin_src = cin_src;
    p.split = split;
    p.wlo = wlo ? 1 : 0;
    p.dbg = getenv("REDTAIL_TC_DEBUG") ? atoi(getenv("REDTAIL_TC_DEBUG")) : 0;
    p.kc = cin >= 64 ? 64 : cin;
    if (const char* e = getenv("REDTAIL_TC_KC")) {           // experiment switch: narrower K blocks = smaller, more numerous stages
        const int kc = atoi(e);
        if ((kc == 16 || kc == 32 || kc == 64) && kc <= p.kc && cin % kc == 0) p.kc = kc;
    }
    p.ncb = cin / p.kc;
    p.cout = cout; p.cout_pad = cout_pad; p.nb = nb;
    p.fuse_elu = d.fuse_elu;
    p.fuse_act = plan->act_host.empty() ? 0 : 1;
    const int kdim[3] = {d.v, d.r, d.s};
    int out_ext[3];
    if (!tr) {
        t->in_d = d.in_dims[0]; t->in_h = d.in_dims[2]; t->in_w = d.in_dims[3];
        t->in_sd = static_cast<long long>(cin_src) * t->in_h * t->in_w;  // input [D,C,H,W]
        t->in_sc = static_cast<long long>(t->in_h) * t->in_w;
        out_ext[0] = d.out_dims[1]; out_ext[1] = d.out_dims[2]; out_ext[2] = d.out_dims[3];
        for (int i = 0; i < 3; ++i) { p.in_s[i] = d.stride[i]; p.out_s[i] = 1; }
        const long long plane = static_cast<long long>(out_ext[1]) * out_ext[2];
        if (d.out_transposed) { p.out_sd = ctot * plane; p.out_sc = plane; }          // [Do,K,Ho,Wo]
        else { p.out_sc = out_ext[0] * plane; p.out_sd = plane; }                       // [K,Do,Ho,Wo]
        p.out_sn = static_cast<long long>(ctot) * out_ext[0] * plane;
    } else {
        t->in_d = d.in_dims[1]; t->in_h = d.in_dims[2]; t->in_w = d.in_dims[3];
        t->in_sc = static_cast<long long>(t->in_d) * t->in_h * t->in_w;  // input [K,D,H,W]
        t->in_sd = static_cast<long long>(t->in_h) * t->in_w;
        out_ext[0] = plan->out_planes; out_ext[1] = d.out_dims[2]; out_ext[2] = d.out_dims[3];
        for (int i = 0; i < 3; ++i) { p.in_s[i] = 1; p.out_s[i] = d.stride[i]; }
        const long long plane = static_cast<long long>(out_ext[1]) * out_ext[2];
        p.out_sd = cout * plane; p.out_sc = plane;                                      // [Dx,C,Hx,Wx]
        p.out_sn = static_cast<long long>(out_ext[0]) * cout * plane;
    }
    t->in_sn = static_cast<long long>(cin_src) * t->in_d * t->in_h * t->in_w;              // caller's tensor
    t->in_elems = static_cast<size_t>(cin) * t->in_d * t->in_h * t->in_w;                   // packed planes
    p.out_d = out_ext[0]; p.out_h = out_ext[1]; p.out_w = out_ext[2];

    p.ncols = ncols;
    p.lc = 0;
    while ((1 << p.lc) < cpc) ++p.lc;
    p.ld = lmerge[0]; p.lh = lmerge[1]; p.lw = lmerge[2];
    // Column order: dense output  -> [pd][ph][channel][pw]  (adjacent W parities leave a thread back to back);
    //               split16 output-> [pd][ph][pw][channel]  (8 consecutive columns = one 16-byte channel vector).
    struct ColDesc { int pd, ph, pw, c; bool valid; };
    std::vector<ColDesc> cold(cout_pad, ColDesc{0, 0, 0, 0, false});
    for (int col = 0; col < ncols; ++col) {
        ColDesc cd{};
        if (!out_split) {
            cd.pw = col & ((1 << p.lw) - 1);
            int tt = col >> p.lw;
            cd.c = tt & (cpc - 1);
            tt >>= p.lc;
            cd.ph = tt & ((1 << p.lh) - 1);
            cd.pd = tt >> p.lh;
        } else {
            cd.c = col & (cpc - 1);
            int tt = col >> p.lc;
            cd.pw = tt & ((1 << p.lw) - 1);
            tt >>= p.lw;
            cd.ph = tt & ((1 << p.lh) - 1);
            cd.pd = tt >> p.lh;
        }
        cd.valid = cd.c < cout;
        cold[col] = cd;
    }

    // Classes (un-merged output parities) and their A-shift tables.
    int ncls[3];
    for (int i = 0; i < 3; ++i) ncls[i] = (tr && !lmerge[i]) ? d.stride[i] : 1;
    p.nclasses = ncls[0] * ncls[1] * ncls[2];
    // GEMM-M lattice extent per dim: outputs / stride (rounded up) for the transposed conv.
    auto lattice = [&](int i, int e) {
        if (!tr) return out_ext[i];
        const int v2 = lmerge[i] ? (out_ext[i] + d.stride[i] - 1) / d.stride[i] : (out_ext[i] - e + d.stride[i] - 1) / d.stride[i];
        return v2 > 0 ? v2 : 0;
    };
    const int cls_h = lattice(1, 0), cls_w = lattice(2, 0);
    // Stride-2 forward convs (conv3D_3ds / 6ds): with 8-wide patches one 8-row UMMA group is exactly one patch row, so an A
    // box that holds every input row lets the three filter rows share it (row i starts at box row i, groups are two box
    // rows apart) -- 9 stages per tile instead of 27.
    p.hs2 = 0;
    if (!tr && d.stride[1] == 2 && d.r == 3 && !getenv("REDTAIL_TC_NOGROUP") && !getenv("REDTAIL_TC_NOHS2")) {
        const int rows = (16 - 1) * 2 + 3;
        const int a_g = (rows * 8 * p.kc * 2 + 1023) & ~1023;
        const int b_g = (nb * p.kc * 2 + 1023) & ~1023;
        if ((196 * 1024) / (a_g * (split ? 2 : 1) + 3 * b_g) >= 2) p.hs2 = 1;
    }
    long long best = -1;
    if (p.hs2) { p.tw = 8; p.th = 16; best = 0; }
    for (int tw = 8; tw <= 128 && !p.hs2; tw *= 2) {
        const int th = kTileM / tw;
        if (!tr && (tw * d.stride[2] > 256 || th * d.stride[1] > 256)) continue;       // TMA box limit
        const long long area = static_cast<long long>((cls_w + tw - 1) / tw) * tw * ((cls_h + th - 1) / th) * th;
        if (best < 0 || area < best) { best = area; p.tw = tw; p.th = th; }
    }
    // For (class parity e, merged parity em, shift off) the filter tap along dim i, or -1 when that parity does not
    // use that shift:  forward conv: tap = off + pad;  transposed: tap = e + pad - stride * off.
    auto tap_of = [&](int i, int e, int off) {
        const int tp = tr ? e + d.pad[i] - d.stride[i] * off : off + d.pad[i];
        return (tp >= 0 && tp < kdim[i]) ? tp : -1;
    };
    struct Tile { int cls; int off[3]; };
    std::vector<Tile> tiles;
    // Row groups (see TapEntry): the H shifts of a class are consecutive (3 filter rows of a stride-1 conv, 2 shifts of
    // a merged-parity transposed conv); they share one A box when >= 2 pipeline stages still fit.  Needs in_s == 1.
    p.gr = p.hs2 ? 3 : 1;
    if (!p.hs2 && p.in_s[0] == 1 && p.in_s[1] == 1 && p.in_s[2] == 1 && !getenv("REDTAIL_TC_NOGROUP")) {
        int want = 1;
        if (!tr) want = d.r;
        else
            for (int e = 0; e < (lmerge[1] ? 1 : d.stride[1]); ++e) {
                int cnt = 0;
                for (int off = -4; off <= 4; ++off) {
                    bool any = false;
                    for (int em = 0; em <= lmerge[1]; ++em) any = any || tap_of(1, lmerge[1] ? em : e, off) >= 0;
                    cnt += any;
                }
                want = cnt > want ? cnt : want;
            }
        if (want > 3) want = 3;
        int min_stages = 2;     // measured: a 2-stage ring with 1/3 of the A traffic beats a deeper per-tap ring (conv3D_4/5: -25 %)
        if (const char* e = getenv("REDTAIL_TC_MINSTAGES")) min_stages = atoi(e) > 0 ? atoi(e) : 2;
        for (; want > 1; --want) {
            const int a_g = (p.th + want - 1) * p.tw * p.kc * 2;
            const int b_g = (nb * p.kc * 2 + 1023) & ~1023;
            if ((196 * 1024) / (a_g * (split ? 2 : 1) + want * b_g) >= min_stages) { p.gr = want; break; }
        }
    }
    // Two H-stacked M tiles per job (forward convs with a row group, narrow N): the second tile reuses the weight tiles
    // and shares the 2 halo rows of the A box -- 25 % less L2 -> SM traffic on the L2-bound 32-channel layers.
    p.mt = 1;
    if (!p.hs2 && p.gr >= 2 && cout_pad <= 32 && cls_h >= 2 * p.th && !getenv("REDTAIL_TC_MT1")) {
        const int a_g = (2 * p.th + p.gr - 1) * p.tw * p.kc * 2;
        const int b_g = (nb * p.kc * 2 + 1023) & ~1023;
        if ((196 * 1024) / (a_g * (split ? 2 : 1) + p.gr * b_g) >= 2) p.mt = 2;
    }
    int job = 0, ci = 0;
    for (int ed = 0; ed < ncls[0]; ++ed)
        for (int eh = 0; eh < ncls[1]; ++eh)
            for (int ew = 0; ew < ncls[2]; ++ew, ++ci) {
                ClassInfo& c = p.cls[ci];
                c.ed = ed; c.eh = eh; c.ew = ew;
                const int e[3] = {ed, eh, ew};
                c.dc = lattice(0, ed); c.hc = lattice(1, eh); c.wc = lattice(2, ew);
                c.tiles_h = (c.hc + p.th * p.mt - 1) / (p.th * p.mt); c.tiles_w = (c.wc + p.tw - 1) / p.tw;
                c.job_begin = job;
                job += c.dc * c.tiles_h * c.tiles_w;
                c.ntaps = 0;
                // candidate shifts per dim: every offset for which at least one (merged) parity has a valid tap
                std::vector<int> offs[3];
                for (int i = 0; i < 3; ++i)
                    for (int off = -4; off <= 4; ++off) {
                        bool any = false;
                        for (int em = 0; em <= lmerge[i]; ++em) any = any || tap_of(i, lmerge[i] ? em : e[i], off) >= 0;
                        if (any) offs[i].push_back(off);
                    }
                for (int od : offs[0])
                    for (size_t hi2 = 0; hi2 < offs[1].size(); hi2 += p.gr)
                        for (int ow : offs[2]) {
                            if (c.ntaps >= kMaxTaps || tiles.size() + p.gr > 255) { delete t; return RT_ERR_UNSUPPORTED; }
                            TapEntry te{};
                            te.dd = static_cast<int8_t>(od); te.dh = static_cast<int8_t>(offs[1][hi2]); te.dw = static_cast<int8_t>(ow);
                            const int nr = static_cast<int>(offs[1].size() - hi2) < p.gr ? static_cast<int>(offs[1].size() - hi2) : p.gr;
                            te.nr = static_cast<uint8_t>(nr);
                            for (int g = 0; g < nr; ++g) {                 // consecutive dh by construction
                                te.widx[g] = static_cast<uint8_t>(tiles.size());
                                tiles.push_back(Tile{ci, {od, offs[1][hi2 + g], ow}});
                            }
                            c.taps[c.ntaps++] = te;
                        }
            }
    p.jobs_per_sample = job;
    p.interleave = 0;
    // Measured (round 2, NVSmall): deconv3D_2 0.530 -> 0.580 ms, deconv3D_1 0.193 -> 0.222 ms WITH the interleaved order, i.e. the
    // input re-read is not what bounds these layers and neighbouring classes writing the same output lines from different SMs
    // costs more than the L2 hits save.  Kept as an experiment switch (REDTAIL_TC_INTERLEAVE=1), off by default.
    if (p.nclasses > 1 && getenv("REDTAIL_TC_INTERLEAVE") && atoi(getenv("REDTAIL_TC_INTERLEAVE")) == 1) {
        p.grid_d = p.grid_th = p.grid_tw = 0;
        for (int c2 = 0; c2 < p.nclasses; ++c2) {
            p.grid_d = p.cls[c2].dc > p.grid_d ? p.cls[c2].dc : p.grid_d;
            p.grid_th = p.cls[c2].tiles_h > p.grid_th ? p.cls[c2].tiles_h : p.grid_th;
            p.grid_tw = p.cls[c2].tiles_w > p.grid_tw ? p.cls[c2].tiles_w : p.grid_tw;
        }
        p.interleave = 1;
        p.jobs_per_sample = p.nclasses * p.grid_d * p.grid_th * p.grid_tw;
    }
    for (int ci = 0; ci < p.nclasses; ++ci) {
        unsigned long long pack = 0;
        for (int t2 = 0; t2 < p.cls[ci].ntaps; ++t2) pack |= static_cast<unsigned long long>(p.cls[ci].taps[t2].nr & 3u) << (2 * t2);
        p.cls[ci].nr_pack = pack;
    }

    // Weight packing: [tile][cb][row][kc] fp16; row = accumulator column (hi), cout_pad + column (lo, fp32 mode).
    const int ntap = static_cast<int>(tiles.size());
    std::vector<uint16_t> pk(static_cast<size_t>(ntap) * p.ncb * nb * p.kc, 0);
    for (int ti = 0; ti < ntap; ++ti) {
        const ClassInfo& c = p.cls[tiles[ti].cls];
        const int e[3] = {c.ed, c.eh, c.ew};
        for (int col = 0; col < ncols; ++col) {
            const int ch = cold[col].c;
            const int em[3] = {cold[col].pd, cold[col].ph, cold[col].pw};
            if (!cold[col].valid) continue;
            int tp[3];
            bool ok = true;
            for (int i = 0; i < 3; ++i) {
                tp[i] = tap_of(i, lmerge[i] ? em[i] : e[i], tiles[ti].off[i]);
                ok = ok && tp[i] >= 0;
            }
            if (!ok) continue;
            for (int kin = 0; kin < cin_src; ++kin) {
                const int kk = tr ? kin : ch, cc = tr ? ch : kin;       // KVCRS indices
                float val = w[(((static_cast<size_t>(kk) * d.v + tp[0]) * d.c + cc) * d.r + tp[1]) * d.s + tp[2]];
                val = val > 65504.f ? 65504.f : (val < -65504.f ? -65504.f : val);
                const size_t base = ((static_cast<size_t>(ti) * p.ncb + kin / p.kc) * nb) * p.kc + (kin % p.kc);
                const uint16_t hb = f2h_bits(val);
                pk[base + static_cast<size_t>(col) * p.kc] = hb;
                if (wlo) pk[base + static_cast<size_t>(cout_pad + col) * p.kc] = f2h_bits((val - h2f_bits(hb)) * 2048.f);
            }
        }
    }
    // Output addressing + device column table.
    p.out_split = out_split ? 1 : 0;
    p.out_c = ctot;
    if (out_split) {
        const long long plane_elems = static_cast<long long>(p.out_d) * p.out_h * p.out_w * ctot;
        p.out_sn = 2 * plane_elems;          // halves per sample: hi plane + lo plane
        p.out_lo = plane_elems;
    }
    {
        std::vector<ColInfo> tab(128, ColInfo{0, 0, 8});
        for (int col = 0; col < cout_pad; ++col) {
            const ColDesc& cd = cold[col];
            ColInfo ci2;
            ci2.ch = cd.c;
            ci2.pidx = cd.valid ? cd.pd * 4 + cd.ph * 2 + cd.pw : 8;
            ci2.off = out_split ? ((static_cast<long long>(cd.pd) * p.out_h + cd.ph) * p.out_w + cd.pw) * ctot + cd.c + coff
                                : cd.pd * p.out_sd + (cd.c + coff) * p.out_sc + static_cast<long long>(cd.ph) * p.out_w + cd.pw;
            tab[col] = ci2;
        }
        if (cudaMalloc(&t->d_cols, tab.size() * sizeof(ColInfo)) != cudaSuccess ||
            cudaMemcpy(t->d_cols, tab.data(), tab.size() * sizeof(ColInfo), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(t->d_cols);
            delete t;
            return static_cast<int>(cudaErrorMemoryAllocation);
        }
    }
    if (p.fuse_act) {
        if (cudaMalloc(&t->act_dev, plan->act_host.size() * sizeof(float)) != cudaSuccess ||
            cudaMemcpy(t->act_dev, plan->act_host.data(), plan->act_host.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(t->act_dev); cudaFree(t->d_cols);
            delete t;
            return static_cast<int>(cudaErrorMemoryAllocation);
        }
    }
    t->in_split = d.in_layout == RT_LAYOUT_SPLIT16;
    if (cudaMalloc(&t->w_dev, pk.size() * 2) != cudaSuccess ||
        cudaMemcpy(t->w_dev, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(t->w_dev);
        delete t;
        return static_cast<int>(cudaErrorMemoryAllocation);
    }
    const CUtensorMapSwizzle swz = p.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(p.kc), static_cast<uint64_t>(ntap) * p.ncb * nb};
        const uint64_t strides[1] = {static_cast<uint64_t>(p.kc) * 2};
        const uint32_t box[2] = {static_cast<uint32_t>(p.kc), static_cast<uint32_t>(nb)};
        const int rc = make_tensor_map(&t->map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, t->w_dev, dims, strides, box, nullptr, swz);
        if (rc != 0) { cudaFree(t->w_dev); delete t; return rc > 0 ? rc : RT_ERR_UNSUPPORTED; }
    }
    // Shared-memory budget.
    p.a_tx = (p.hs2 ? (p.th - 1) * 2 + p.gr : p.th * p.mt + p.gr - 1) * p.tw * p.kc * 2;
    p.a_bytes = (p.a_tx + 1023) & ~1023;
    p.b_tx = nb * p.kc * 2;
    p.b_bytes = (p.b_tx + 1023) & ~1023;
    p.stage_bytes = p.a_bytes * (split ? 2 : 1) + p.gr * p.b_bytes;
    // Split mode: a TMEM accumulation chain is at most ~8 MMA K-steps (K = 16 each) long before the epilogue adds it
    // into fp32 registers; fp16 mode: one chain per tile.
    // Chain length (MMA K-steps accumulated inside the tensor core before a round-to-nearest add in registers): 8 for
    // problems big enough to be throughput-bound; 2 when the whole problem is a few waves of tiles, where the extra
    // TMEM round trips are free and the reference's tightest unit-test tolerances (1e-4 on values ~200) need it.
    const int per_row = p.kc / 16, per_stage = p.gr * per_row;          // MMA K-steps per filter row / per stage
    int chain = p.jobs_per_sample < 4 * 148 ? 2 : 8;
    // Every chunk costs a TMEM drain (tcgen05.ld moves 64 B/clk/SM and does not overlap the MMAs' own TMEM traffic), so the
    // big layers close a chunk once per stage (6 K-steps for <= 32 input channels, 12 for 64-channel blocks) rather than
    // inside it; REDTAIL_TC_CHAIN=<n> forces n K-steps (sub-stage when n is shorter than a stage).
    if (chain == 8 && per_stage > 8 && per_stage <= 12) chain = per_stage;
    if (const char* e = getenv("REDTAIL_TC_CHAIN")) chain = atoi(e) > 0 ? atoi(e) : chain;
    if (!split) { p.chunk_kb = 1 << 30; p.chunk_rows = p.gr; }
    else if (chain >= per_stage) { p.chunk_kb = chain / per_stage; p.chunk_rows = p.gr; }
    else { p.chunk_kb = 1; p.chunk_rows = chain / per_row > 0 ? chain / per_row : 1; }
    p.stages = (196 * 1024) / p.stage_bytes;
    if (p.stages > 8) p.stages = 8;
    if (p.stages < 2) { cudaFree(t->w_dev); delete t; return RT_ERR_UNSUPPORTED; }
    t->smem_bytes = p.stages * p.stage_bytes + 1024 /*align slack*/ + 512 /*barriers*/ + 512 /*bias*/ + 128 * 16 /*column table*/ + 4 * 128 * 4 /*S-ReLU parameters*/;
    if (t->smem_bytes < 120 * 1024) t->smem_bytes = 120 * 1024;   // > half of the SM: one CTA per SM, so the 512-column TMEM grab never contends
    plan->tc = t;
    return RT_OK;
}

void tc_plan_destroy(rt_conv3d_plan* plan) {
    TcPlan* t = static_cast<TcPlan*>(plan->tc);
    if (!t) return;
    cudaFree(t->w_dev);
    cudaFree(t->d_cols);
    cudaFree(t->act_dev);
    delete t;
    plan->tc = nullptr;
}

size_t tc_workspace_size(const rt_conv3d_plan* plan, int max_batch) {
    const TcPlan* t = static_cast<const TcPlan*>(plan->tc);
    if (t->in_split) return 0;
    return t->in_elems * 4 * static_cast<size_t>(max_batch) + 512;     // per sample: hi plane + lo plane (fp16)
}

int tc_conv3d_enqueue(const rt_conv3d_plan* plan, int n, const float* x, const float* skip, float* y, void* workspace,
                      cudaStream_t s) {
    const TcPlan* t = static_cast<const TcPlan*>(plan->tc);
    if (!workspace && !t->in_split) return RT_ERR_ARG;
    TcParams p = t->p;
    p.njobs = p.jobs_per_sample * n;
    if (p.njobs == 0) return RT_OK;
    // Activations as two channels-last fp16 planes per sample: [n][hi | lo][D][H][W][C].
    const __half* hi;
    if (t->in_split) {
        hi = reinterpret_cast<const __half*>(x);
    } else {
        __half* whi = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
        if (static_cast<long long>(n) * t->in_d > 65535 || t->in_h > 65535) return RT_ERR_UNSUPPORTED;
        if (!plan->reuse_pack) {        // parts 1.. of a > 128-channel convolution read the planes part 0 has just written
            dim3 grid((t->in_w + 63) / 64, t->in_h, n * t->in_d);
            const size_t sm = static_cast<size_t>(t->cin) * 65 * sizeof(float);
            if (sm > 48 * 1024) {               // 256 / 512-channel inputs (TrailNet): the transpose tile needs the opt-in limit
                static bool pack_attr[64] = {};
                int dev_ = 0;
                RT_CUDA(cudaGetDevice(&dev_));
                if (sm > 227 * 1024) return RT_ERR_UNSUPPORTED;
                if (dev_ < 0 || dev_ >= 64 || !pack_attr[dev_]) {
                    RT_CUDA(cudaFuncSetAttribute(pack_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
                    if (dev_ >= 0 && dev_ < 64) pack_attr[dev_] = true;
                }
            }
            pack_split_kernel<<<grid, 256, sm, s>>>(x, whi, whi + t->in_elems, t->in_d, t->cin_src, t->cin, t->in_h, t->in_w, t->in_sn,
                                                    t->in_sd, t->in_sc, static_cast<long long>(2 * t->in_elems));
            note_launch("conv3d_pack_split");
            RT_CHECK_LAUNCH();
        }
        hi = whi;
    }
    const __half* lo = hi + t->in_elems;
    // 2. tensor maps over the packed activations: dims (C, W, H, D, N)
    CUtensorMap ma_hi, ma_lo;
    {
        const uint64_t dims[5] = {static_cast<uint64_t>(t->cin), static_cast<uint64_t>(t->in_w), static_cast<uint64_t>(t->in_h),
                                  static_cast<uint64_t>(t->in_d), static_cast<uint64_t>(n)};
        const uint64_t st[4] = {static_cast<uint64_t>(t->cin) * 2, static_cast<uint64_t>(t->cin) * 2 * t->in_w,
                                static_cast<uint64_t>(t->cin) * 2 * t->in_w * t->in_h,
                                static_cast<uint64_t>(t->in_elems) * 4};          // sample stride: hi + lo planes
        const uint32_t box[5] = {static_cast<uint32_t>(p.kc), static_cast<uint32_t>(p.tw * p.in_s[2]),
                                 static_cast<uint32_t>(p.hs2 ? (p.th - 1) * 2 + p.gr : (p.th * p.mt + p.gr - 1) * p.in_s[1]), 1u, 1u};
        const uint32_t es[5] = {1u, static_cast<uint32_t>(p.in_s[2]), static_cast<uint32_t>(p.hs2 ? 1 : p.in_s[1]), 1u, 1u};
        const CUtensorMapSwizzle swz = p.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        int rc = make_tensor_map(&ma_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, hi, dims, st, box, es, swz);
        if (rc == 0 && p.split) rc = make_tensor_map(&ma_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, lo, dims, st, box, es, swz);
        if (rc != 0) return rc > 0 ? rc : RT_ERR_UNSUPPORTED;
        if (!p.split) ma_lo = ma_hi;
    }
    // 3. main kernel
    int grid = num_sms();
    if (grid > p.njobs) grid = p.njobs;
#define RT_LAUNCH_UMMA(CPH, SPL, MTT, EWW)                                                                               \
    do {                                                                                                              \
        static bool attr_set[64] = {};            /* per device: the attribute belongs to the device's copy of the kernel */ \
        int dev_ = 0;                                                                                                 \
        RT_CUDA(cudaGetDevice(&dev_));                                                                                \
        if (dev_ < 0 || dev_ >= 64 || !attr_set[dev_]) {                                                              \
            RT_CUDA(cudaFuncSetAttribute(conv3d_umma_kernel<CPH, SPL, MTT, EWW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
            if (dev_ >= 0 && dev_ < 64) attr_set[dev_] = true;                                                        \
        }                                                                                                             \
        conv3d_umma_kernel<CPH, SPL, MTT, EWW><<<grid, kThreadsOf(EWW, MTT), t->smem_bytes, s>>>(ma_hi, ma_lo, t->map_w, p, plan->bias, t->d_cols, skip, y, t->act_dev); \
    } while (0)
    // 16 epilogue warps (4 column groups) from 32 accumulator columns up; REDTAIL_TC_EW=8 keeps 8.
    static const bool ew8 = getenv("REDTAIL_TC_EW") && atoi(getenv("REDTAIL_TC_EW")) == 8;
    switch (p.cout_pad) {
        case 16:
            if (p.mt == 2) { if (p.split) RT_LAUNCH_UMMA(8, true, 2, 8); else RT_LAUNCH_UMMA(8, false, 2, 8); }
            else { if (p.split) RT_LAUNCH_UMMA(8, true, 1, 8); else RT_LAUNCH_UMMA(8, false, 1, 8); }
            break;
        case 32:
            if (ew8) {
                if (p.mt == 2) { if (p.split) RT_LAUNCH_UMMA(16, true, 2, 8); else RT_LAUNCH_UMMA(16, false, 2, 8); }
                else { if (p.split) RT_LAUNCH_UMMA(16, true, 1, 8); else RT_LAUNCH_UMMA(16, false, 1, 8); }
            } else {
                if (p.mt == 2) { if (p.split) RT_LAUNCH_UMMA(8, true, 2, 16); else RT_LAUNCH_UMMA(8, false, 2, 16); }
                else { if (p.split) RT_LAUNCH_UMMA(8, true, 1, 16); else RT_LAUNCH_UMMA(8, false, 1, 16); }
            }
            break;
        case 64:
            if (ew8) { if (p.split) RT_LAUNCH_UMMA(32, true, 1, 8); else RT_LAUNCH_UMMA(32, false, 1, 8); }
            else { if (p.split) RT_LAUNCH_UMMA(16, true, 1, 16); else RT_LAUNCH_UMMA(16, false, 1, 16); }
            break;
        case 128:           // 128 + 128 accumulator columns: 64 + 64 per thread with 8 warps fit in 200 registers; 16 warps (112) spill
            if (p.split) RT_LAUNCH_UMMA(64, true, 1, 8); else RT_LAUNCH_UMMA(64, false, 1, 8);
            break;
        default: return RT_ERR_UNSUPPORTED;
    }
#undef RT_LAUNCH_UMMA
    note_launch(p.split ? (p.wlo ? "conv3d_umma_fp16x2split" : "conv3d_umma_fp16x2split_w16") : "conv3d_umma_fp16");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

}  // namespace rt

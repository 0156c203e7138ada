// Fused CostVolume(kDefault) -> Conv3D(3x3x3, stride 1, pad 1) [-> Transform] [-> ELU]  (SURVEY.md section 8, row N3).
//
// Reference pair being replaced: CostVolumePlugin::enqueue (lib/cost_volume_plugin.cpp:122-139, kernels lib/kernels.cu:50-97)
// followed by Conv3DPlugin::enqueue (lib/conv3d_plugin.cpp:186-279), Transform and ELU, i.e. `cost_vol` -> `conv3D_1` of every
// generated builder (sample_app/nvsmall_1025x321_net.cpp:151-196).
//
// The volume cv[d', 0:C] = L and cv[d', C:2C, y, x'] = (x' >= d') ? R[y, x'-d'] : 0 is a broadcast / shift of two 2-D maps, so
// the 3-D cross-correlation over it separates.  With v = filter plane (input plane d' = d + v - 1, must lie in [0, D)):
//     left half :  sum_{c,dh,dw} W[k,v,c,dh,dw] L[c,y+dh,x+dw]                         = A_v[k,y,x]        (independent of d)
//     right half:  sum_{c,dh,dw} W[k,v,C+c,dh,dw] [0<=x+dw<w] [x+dw>=d'] R[c,y+dh,x+dw-d'] = C_v[k,y,x-d']
// where A_v, C_v are plain zero-padded 3x3 2-D convolutions of the feature maps (the `x+dw >= d'` mask is exactly the left
// zero padding of R).  Two places differ from the plain 2-D convolution and are patched from a 3x1 "edge" convolution
// E_v[k,y,u] = sum_{c,dh} W[k,v,C+c,dh,dw=+1] R[c,y+dh,u+1]:
//   * x = w-1: the dw=+1 tap is outside the volume but inside the shifted image -> subtract E_v[u],  u = x-d';
//   * u = -1 (x = d'-1): the 2-D map has no column -1, but the dw=+1 tap reads R[.,0] -> C_v[-1] = E_v[-1].
// Work: two 2-D convolutions C -> 3K (on the tcgen05 kernel, fp16x2-split) instead of a 3-D convolution 2C -> K over D planes,
// then one HBM-bound pass that writes the conv3D_1 output (and is the only pass that touches D x H x W data).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "conv3d_internal.h"

struct rt_cvconv_plan {
    rt_costvol_conv3d_desc desc;
    rt_conv3d_plan* conv_l = nullptr;   // L -> A [3K,h,w]
    rt_conv3d_plan* conv_r = nullptr;   // R -> C [3K,h,w]
    float* bias = nullptr;              // [K] (zeros if absent)
    float* w_edge = nullptr;            // [3 v][3 dh][C][K]: the dw = +1 column of the right-half weights
};

namespace rt {
namespace {

constexpr int kXT = 64;   // output positions (x) per CTA of the combine kernel
__device__ const float g_zero = 0.f;

// E[n][v][y][j][k], j in [0, D]: u = w - D + j for j < D, u = -1 for j == D.   One CTA per (y, v, n): the 3 x C x (D+1)
// slab of the right feature map (columns u+1) and the 3 x C x K weight column are staged in shared memory, every thread
// then owns a few (j, k) outputs -- a [D+1, 3C] x [3C, K] product with broadcast / conflict-free shared-memory reads.
__global__ void __launch_bounds__(256)
cvconv_edge_kernel(const float* __restrict__ right, const float* __restrict__ w_edge, float* __restrict__ e,
                   int c, int h, int w, int disp, int k) {
    extern __shared__ float esm[];
    const int y = blockIdx.x, v = blockIdx.y, n = blockIdx.z;
    const int nj = disp + 1;
    float* rs = esm;                       // [3][c][nj]
    float* ws = esm + 3 * c * nj;          // [3][c][k]
    const long long hw = static_cast<long long>(h) * w;
    for (int idx = threadIdx.x; idx < 3 * c * nj; idx += 256) {
        const int j = idx % nj, cc = (idx / nj) % c, dh = idx / (nj * c);
        const int yy = y + dh - 1;
        const int col = j == disp ? 0 : w - disp + j + 1;
        rs[idx] = (yy >= 0 && yy < h && col < w) ? __ldg(right + (static_cast<long long>(n) * c + cc) * hw + static_cast<long long>(yy) * w + col) : 0.f;
    }
    const float* wv = w_edge + static_cast<long long>(v) * 3 * c * k;
    for (int idx = threadIdx.x; idx < 3 * c * k; idx += 256) ws[idx] = __ldg(wv + idx);
    __syncthreads();
    float* eo = e + ((static_cast<long long>(n) * 3 + v) * h + y) * nj * k;
    for (int o = threadIdx.x; o < nj * k; o += 256) {
        const int j = o / k, kk = o % k;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int r = 0;
        for (; r + 4 <= 3 * c; r += 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fmaf(ws[(r + q) * k + kk], rs[(r + q) * nj + j], acc[q]);
        }
        for (; r < 3 * c; ++r) acc[0] = fmaf(ws[r * k + kk], rs[r * nj + j], acc[0]);
        eo[o] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    }
}

// One CTA: kXT consecutive x of one (n, y) row, all K channels, all D planes.  Thread = (x, 8-channel group), channel
// group fastest so that the 16-byte stores of a warp are contiguous in the channels-last output.
//   smem: S[pw][K] position-major, S[u] = C_0[u+1] + C_1[u] + C_2[u-1] for u in [x0-D, x0+kXT] (the interior planes'
//   right-half term); 16-byte chunks XOR-swizzled with the position parity so a quarter warp reads 8 distinct chunks.
// The first / last plane (one filter plane falls outside the volume) are formed in registers straight from global memory.
template <bool kSplitOut>
__global__ void __launch_bounds__(256, 3)
cvconv_combine_kernel(const float* __restrict__ a, const float* __restrict__ cc, const float* __restrict__ e,
                      const float* __restrict__ bias, void* __restrict__ out,
                      int h, int w, int disp, int k, int fuse_elu, int out_transposed) {
    extern __shared__ float4 sm4[];
    const int x0 = blockIdx.x * kXT, y = blockIdx.y, n = blockIdx.z;
    const int nthreads = blockDim.x;
    const long long hw = static_cast<long long>(h) * w;
    const int pw = kXT + disp + 1;          // staged positions: u = x0 - disp + i
    const int k4 = k >> 2;                  // 16-byte chunks per position
    const float* cplane = cc + (static_cast<long long>(n) * 3 * k) * hw + static_cast<long long>(y) * w;
    const float* erow = e + ((static_cast<long long>(n) * 3) * h + y) * (disp + 1) * k;
    // C_v[ch][u] with the two patches of the file header: zero left of u = -1, the edge map at u = -1.
    // (one load from a selected address, no branches, so that the loads of a thread overlap)
    auto load_c = [&](int v, int ch, int u) -> float {
        const float* pc = cplane + (static_cast<long long>(v) * k + ch) * hw + u;
        const float* pe = erow + (static_cast<long long>(v) * h * (disp + 1) + disp) * k + ch;
        const float* ptr = (u >= 0 && u < w) ? pc : (u == -1 ? pe : &g_zero);
        return __ldg(ptr);
    };
    for (int idx = threadIdx.x; idx < k4 * pw; idx += nthreads) {
        const int i = idx % pw, g = idx / pw;
        const int u = x0 - disp + i;
        float4 sv;
        float* sp = reinterpret_cast<float*>(&sv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ch = 4 * g + q;
            sp[q] = load_c(0, ch, u + 1) + load_c(1, ch, u) + load_c(2, ch, u - 1);
        }
        sm4[i * k4 + (g ^ (i & 1))] = sv;
    }
    __syncthreads();
    const int kg_n = k >> 3;
    const int kg = threadIdx.x % kg_n, xl = threadIdx.x / kg_n;
    const int x = x0 + xl;
    if (xl >= kXT || x >= w) return;
    const int c0 = kg * 8;
    const bool last_col = x == w - 1;
    const long long plane = static_cast<long long>(disp) * hw * k;          // elements of one split16 fp16 plane
    auto emit = [&](int d, float (&val)[8]) {
        if (last_col) {
            // the dw=+1 column of the filter is outside the volume at x = w-1: E_v[u], u = x - d - (v-1) -> j_e = D - d - v
            const int v_lo = d == 0 ? 1 : 0, v_hi = d == disp - 1 ? 1 : 2;
            for (int v = v_lo; v <= v_hi; ++v) {
                const float* ep = erow + (static_cast<long long>(v) * h * (disp + 1) + (disp - d - v)) * k + c0;
#pragma unroll
                for (int j = 0; j < 8; ++j) val[j] -= __ldg(ep + j);
            }
        }
        if (fuse_elu) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) elu1_x2(val[j], val[j + 1]);     // the library's one ELU (common.cuh), two lanes per instruction
        }
        if (kSplitOut) {
            __half* hi = static_cast<__half*>(out) + static_cast<long long>(n) * 2 * plane;
            const long long o = ((static_cast<long long>(d) * h + y) * w + x) * k + c0;
            split_store8(val, hi + o, hi + plane + o);
        } else {
            float* o = static_cast<float*>(out);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long idx = out_transposed ? ((static_cast<long long>(n) * disp + d) * k + c0 + j) * hw
                                                     : ((static_cast<long long>(n) * k + c0 + j) * disp + d) * hw;
                o[idx + static_cast<long long>(y) * w + x] = val[j];
            }
        }
    };
    const float* ap = a + (static_cast<long long>(n) * 3 * k + c0) * hw + static_cast<long long>(y) * w + x;
    float base[8];
    {
        float rf[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float b = __ldg(bias + c0 + j);
            const float a0 = __ldg(ap + static_cast<long long>(j) * hw);
            const float a1 = __ldg(ap + static_cast<long long>(k + j) * hw);
            const float a2 = __ldg(ap + static_cast<long long>(2 * k + j) * hw);
            base[j] = b + (a0 + a1 + a2);
            rf[j] = b + ((a1 + a2) + (load_c(1, c0 + j, x) + load_c(2, c0 + j, x - 1)));             // d = 0: planes v = 1, 2
        }
        emit(0, rf);
    }
#pragma unroll 2
    for (int d = 1; d < disp - 1; ++d) {
        const int i = xl + disp - d;                                        // staged index of u = x - d
        const float4 s0 = sm4[i * k4 + ((2 * kg) ^ (i & 1))];
        const float4 s1 = sm4[i * k4 + ((2 * kg + 1) ^ (i & 1))];
        float val[8];
        upk2(add2(pk2(base[0], base[1]), pk2(s0.x, s0.y)), val[0], val[1]);
        upk2(add2(pk2(base[2], base[3]), pk2(s0.z, s0.w)), val[2], val[3]);
        upk2(add2(pk2(base[4], base[5]), pk2(s1.x, s1.y)), val[4], val[5]);
        upk2(add2(pk2(base[6], base[7]), pk2(s1.z, s1.w)), val[6], val[7]);
        emit(d, val);
    }
    {
        float rl[8];
        const int ul = x - (disp - 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float b = __ldg(bias + c0 + j);
            const float a0 = __ldg(ap + static_cast<long long>(j) * hw);
            const float a1 = __ldg(ap + static_cast<long long>(k + j) * hw);
            rl[j] = b + ((a0 + a1) + (load_c(0, c0 + j, ul + 1) + load_c(1, c0 + j, ul)));          // d = D-1: planes v = 0, 1
        }
        emit(disp - 1, rl);
    }
}

size_t combine_smem_impl(int k, int disp) { return static_cast<size_t>(kXT + disp + 1) * k * sizeof(float); }

size_t combine_smem(const rt_costvol_conv3d_desc& d) { return combine_smem_impl(d.k, d.max_disp); }
size_t edge_smem(const rt_costvol_conv3d_desc& d) { return static_cast<size_t>(3) * d.c * (d.max_disp + 1 + d.k) * sizeof(float); }

rt_conv3d_desc inner_desc(const rt_costvol_conv3d_desc& d) {
    rt_conv3d_desc c3{};
    c3.k = 3 * d.k; c3.v = 1; c3.c = d.c; c3.r = 3; c3.s = 3;
    c3.stride[0] = c3.stride[1] = c3.stride[2] = 1;
    c3.pad[0] = 0; c3.pad[1] = 1; c3.pad[2] = 1;
    c3.in_dims[0] = 1; c3.in_dims[1] = d.c; c3.in_dims[2] = d.h; c3.in_dims[3] = d.w;
    c3.out_dims[0] = 3 * d.k; c3.out_dims[1] = 1; c3.out_dims[2] = d.h; c3.out_dims[3] = d.w;
    c3.weights_dtype = RT_F32;
    c3.precision = d.precision;
    c3.in_layout = RT_LAYOUT_DENSE; c3.out_layout = RT_LAYOUT_DENSE;
    return c3;
}

size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

}  // namespace
}  // namespace rt

using namespace rt;

extern "C" {

int rt_costvol_conv3d_supported(const rt_costvol_conv3d_desc* d) {
    if (!d) return 0;
    if (d->c <= 0 || d->h <= 0 || d->w <= 0 || d->k <= 0 || d->max_disp < 2) return 0;
    if (d->k % 8 != 0 || d->k > 32) return 0;            // 8-channel vectors, <= 256 threads per CTA
    if (d->w < d->max_disp) return 0;                     // edge columns assume the right border is not reached by u = -1
    if (d->out_layout != RT_LAYOUT_DENSE && d->out_layout != RT_LAYOUT_SPLIT16) return 0;
    if (d->out_layout == RT_LAYOUT_SPLIT16 && d->out_transposed) return 0;
    if (combine_smem(*d) > 200 * 1024 || edge_smem(*d) > 200 * 1024) return 0;
    if (d->precision != RT_PREC_SIMT) {
        rt_conv3d_desc c3 = inner_desc(*d);
        if (!tc_shape_supported(c3)) return 0;
    }
    return 1;
}

int rt_costvol_conv3d_create(const rt_costvol_conv3d_desc* d, rt_cvconv_plan** out) {
    if (!d || !out || !d->weights) return RT_ERR_ARG;
    if (d->weights_dtype != RT_F32 && d->weights_dtype != RT_F16) return RT_ERR_UNSUPPORTED;
    if (!rt_costvol_conv3d_supported(d)) return RT_ERR_UNSUPPORTED;
    const int c = d->c, k = d->k;
    std::vector<float> w, b;
    host_to_f32(d->weights_dtype, d->weights, static_cast<int64_t>(k) * 3 * 2 * c * 9, w);
    if (d->bias) host_to_f32(d->weights_dtype, d->bias, k, b);
    else b.assign(k, 0.f);
    // W[k][v][c'][r][s] -> WL/WR [(v*K + k)][c][r][s], and the dw=+1 column of the right half [v][dh][c][k].
    std::vector<float> wl(static_cast<size_t>(3) * k * c * 9), wr(wl.size()), we(static_cast<size_t>(9) * c * k);
    for (int kk = 0; kk < k; ++kk)
        for (int v = 0; v < 3; ++v)
            for (int cc = 0; cc < c; ++cc)
                for (int r = 0; r < 3; ++r)
                    for (int s = 0; s < 3; ++s) {
                        const size_t dst = ((static_cast<size_t>(v * k + kk) * c + cc) * 3 + r) * 3 + s;
                        const size_t src_l = (((static_cast<size_t>(kk) * 3 + v) * 2 * c + cc) * 3 + r) * 3 + s;
                        const size_t src_r = (((static_cast<size_t>(kk) * 3 + v) * 2 * c + c + cc) * 3 + r) * 3 + s;
                        wl[dst] = w[src_l];
                        wr[dst] = w[src_r];
                        if (s == 2) we[(static_cast<size_t>(v * 3 + r) * c + cc) * k + kk] = w[src_r];
                    }
    rt_cvconv_plan* p = new rt_cvconv_plan();
    p->desc = *d;
    p->desc.weights = nullptr;
    p->desc.bias = nullptr;
    rt_conv3d_desc c3 = inner_desc(*d);
    c3.weights = wl.data();
    int rc = rt_conv3d_create(&c3, &p->conv_l);
    if (rc == RT_OK) {
        c3.weights = wr.data();
        rc = rt_conv3d_create(&c3, &p->conv_r);
    }
    if (rc == RT_OK) rc = static_cast<int>(cudaMalloc(&p->bias, b.size() * sizeof(float)));
    if (rc == RT_OK) rc = static_cast<int>(cudaMemcpy(p->bias, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (rc == RT_OK) rc = static_cast<int>(cudaMalloc(&p->w_edge, we.size() * sizeof(float)));
    if (rc == RT_OK) rc = static_cast<int>(cudaMemcpy(p->w_edge, we.data(), we.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (rc == RT_OK) {
        // The limit is per kernel, not per plan: always raise it to the bound rt_costvol_conv3d_supported() checks against, so a
        // later plan with a smaller footprint cannot lower it under an earlier plan's needs.
        const int smem = 200 * 1024;
        cudaError_t e1 = cudaFuncSetAttribute(cvconv_combine_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaError_t e2 = cudaFuncSetAttribute(cvconv_combine_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaError_t e3 = cudaFuncSetAttribute(cvconv_edge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) rc = static_cast<int>(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3);
    }
    if (rc != RT_OK) {
        rt_costvol_conv3d_destroy(p);
        return rc;
    }
    *out = p;
    return RT_OK;
}

void rt_costvol_conv3d_destroy(rt_cvconv_plan* p) {
    if (!p) return;
    if (p->conv_l) rt_conv3d_destroy(p->conv_l);
    if (p->conv_r) rt_conv3d_destroy(p->conv_r);
    if (p->bias) cudaFree(p->bias);
    if (p->w_edge) cudaFree(p->w_edge);
    delete p;
}

// workspace: A [n,3K,h,w] | C [n,3K,h,w] | E [n,3,h,D+1,K] | inner conv workspace (the two convs run back to back)
size_t rt_costvol_conv3d_workspace_size(const rt_cvconv_plan* p, int max_batch) {
    if (!p || max_batch < 1) return 0;
    const rt_costvol_conv3d_desc& d = p->desc;
    const size_t maps = align256(static_cast<size_t>(max_batch) * 3 * d.k * d.h * d.w * sizeof(float));
    const size_t edge = align256(static_cast<size_t>(max_batch) * 3 * d.h * (d.max_disp + 1) * d.k * sizeof(float));
    const size_t inner = std::max(rt_conv3d_workspace_size(p->conv_l, max_batch), rt_conv3d_workspace_size(p->conv_r, max_batch));
    return 2 * maps + edge + align256(inner);
}

int rt_costvol_conv3d_enqueue(const rt_cvconv_plan* p, int n, const void* left, const void* right, void* y, void* workspace,
                              void* stream) {
    if (!p || n < 1 || !left || !right || !y || !workspace) return RT_ERR_ARG;
    const rt_costvol_conv3d_desc& d = p->desc;
    cudaStream_t s = as_stream(stream);
    const size_t maps = align256(static_cast<size_t>(n) * 3 * d.k * d.h * d.w * sizeof(float));
    const size_t edge = align256(static_cast<size_t>(n) * 3 * d.h * (d.max_disp + 1) * d.k * sizeof(float));
    char* ws = static_cast<char*>(workspace);
    float* a = reinterpret_cast<float*>(ws);
    float* cc = reinterpret_cast<float*>(ws + maps);
    float* e = reinterpret_cast<float*>(ws + 2 * maps);
    void* inner = ws + 2 * maps + edge;
    int rc = rt_conv3d_enqueue(p->conv_l, n, left, nullptr, a, inner, stream);
    if (rc != RT_OK) return rc;
    rc = rt_conv3d_enqueue(p->conv_r, n, right, nullptr, cc, inner, stream);
    if (rc != RT_OK) return rc;
    {
        const dim3 grid(d.h, 3, n);
        cvconv_edge_kernel<<<grid, 256, edge_smem(d), s>>>(static_cast<const float*>(right), p->w_edge, e, d.c, d.h, d.w, d.max_disp, d.k);
        note_launch("cvconv_edge_kernel");
        RT_CHECK_LAUNCH();
    }
    {
        const dim3 grid(static_cast<unsigned>(ceil_div(d.w, kXT)), d.h, n);
        const int threads = kXT * (d.k / 8);
        const size_t smem = combine_smem(d);
        if (d.out_layout == RT_LAYOUT_SPLIT16)
            cvconv_combine_kernel<true><<<grid, threads, smem, s>>>(a, cc, e, p->bias, y, d.h, d.w, d.max_disp, d.k, d.fuse_elu, 0);
        else
            cvconv_combine_kernel<false><<<grid, threads, smem, s>>>(a, cc, e, p->bias, y, d.h, d.w, d.max_disp, d.k, d.fuse_elu,
                                                                      d.out_transposed);
        note_launch("cvconv_combine_kernel");
        RT_CHECK_LAUNCH();
    }
    return RT_OK;
}

}  // extern "C"

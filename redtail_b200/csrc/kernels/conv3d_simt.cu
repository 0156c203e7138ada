// 3-D convolution / transposed convolution plans (Conv3DPlugin, Conv3DTransposePlugin) -- plan management and the
// fp32 CUDA-core kernels (RT_PREC_SIMT).  The tensor-core kernels live in conv3d_tc.cu; this file is the exact-fp32
// validation path they are checked against on the device, and the path for shapes the tcgen05 tiles do not cover
// (channel counts that are not multiples of 16, e.g. the reference's tiny unit-test tensors).
//
// Semantics (reference): lib/conv3d_plugin.cpp:74-100,187-216 + lib/conv_utils.cpp:14-81 (cuDNN cross-correlation,
// symmetric pad = pad_start, input [D,C,H,W], output [K,Do,Ho,Wo], bias over K);
// lib/conv3d_transpose_plugin.cpp:86-114,205-243 (cudnnConvolutionBackwardData: input [K,Dy,Hy,Wy], output
// [Dx,C,Hx,Wx] of caller-supplied out_dims, bias over C -- lib/kernels.cu:292-308).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "conv3d_internal.h"

namespace rt {
namespace {

constexpr int KT = 16;          // output channels per thread
constexpr int kThreads = 128;   // output x positions per CTA

struct ConvGeom {
    int cin, cout;              // conv: C -> K ; transposed: K -> C
    int v, r, s;
    int sd, sh, sw, pd, ph, pw;
    int di, hi, wi;             // input spatial extent
    int dout, ho, wo;           // output spatial extent actually written
    int fuse_elu, out_transposed;
};

__device__ __forceinline__ float4 ldw(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Forward conv.  x [Di, Cin, Hi, Wi];  w packed [Cout/16][V][Cin][R][S][16];  y [Cout, Do, Ho, Wo] (or [Do, Cout, Ho, Wo]).
__global__ void __launch_bounds__(kThreads)
conv3d_fwd_simt_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       const float* __restrict__ skip, float* __restrict__ y, ConvGeom g, int ktiles) {
    const int wo = blockIdx.x * kThreads + threadIdx.x;
    const int ho = blockIdx.y;
    int z = blockIdx.z;
    const int kt = z % ktiles; z /= ktiles;
    const int dd = z % g.dout;
    const int n = z / g.dout;
    const int64_t in_plane = static_cast<int64_t>(g.hi) * g.wi;
    const float* xn = x + static_cast<int64_t>(n) * g.di * g.cin * in_plane;
    const float* wk = w + static_cast<int64_t>(kt) * g.v * g.cin * g.r * g.s * KT;
    float acc[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) acc[i] = 0.f;
    const bool active = wo < g.wo;
    for (int v = 0; v < g.v; ++v) {
        const int d_in = dd * g.sd + v - g.pd;
        if (d_in < 0 || d_in >= g.di) continue;
        for (int c = 0; c < g.cin; ++c) {
            const float* xp = xn + (static_cast<int64_t>(d_in) * g.cin + c) * in_plane;
            const float* wp = wk + (static_cast<int64_t>(v) * g.cin + c) * g.r * g.s * KT;
            for (int r = 0; r < g.r; ++r) {
                const int h_in = ho * g.sh + r - g.ph;
                if (h_in < 0 || h_in >= g.hi) continue;
                for (int s = 0; s < g.s; ++s) {
                    const int w_in = wo * g.sw + s - g.pw;
                    const float xv = (active && w_in >= 0 && w_in < g.wi) ? __ldg(xp + static_cast<int64_t>(h_in) * g.wi + w_in) : 0.f;
                    const float* wq = wp + (r * g.s + s) * KT;
#pragma unroll
                    for (int q = 0; q < KT / 4; ++q) {
                        const float4 wv = ldw(wq + 4 * q);
                        acc[4 * q + 0] = fmaf(wv.x, xv, acc[4 * q + 0]);
                        acc[4 * q + 1] = fmaf(wv.y, xv, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(wv.z, xv, acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(wv.w, xv, acc[4 * q + 3]);
                    }
                }
            }
        }
    }
    if (!active) return;
    const int64_t out_plane = static_cast<int64_t>(g.ho) * g.wo;
    const int64_t nout = static_cast<int64_t>(n) * g.cout * g.dout * out_plane;
#pragma unroll
    for (int i = 0; i < KT; ++i) {
        const int k = kt * KT + i;
        if (k >= g.cout) break;
        const int64_t idx = nout + (g.out_transposed ? (static_cast<int64_t>(dd) * g.cout + k) : (static_cast<int64_t>(k) * g.dout + dd)) * out_plane +
                            static_cast<int64_t>(ho) * g.wo + wo;
        float vv = acc[i] + bias[k];
        if (skip) vv += skip[idx];
        if (g.fuse_elu) vv = elu1(vv);
        y[idx] = vv;
    }
}

// Transposed conv (gather form of the data gradient).  yin [K, Dy, Hy, Wy];  w packed [C/16][V][K][R][S][16];
// x out [Dx, C, Hx, Wx]:  x[dx,c,hx,wx] = b[c] + sum_{k,v,r,s} w[k,v,c,r,s] * yin[k,(dx+pd-v)/sd,(hx+ph-r)/sh,(wx+pw-s)/sw]
// over the taps where the divisions are exact and in range.
__global__ void __launch_bounds__(kThreads)
conv3d_bwd_simt_kernel(const float* __restrict__ yin, const float* __restrict__ w, const float* __restrict__ bias,
                       const float* __restrict__ skip, float* __restrict__ x, ConvGeom g, int ctiles) {
    const int wx = blockIdx.x * kThreads + threadIdx.x;
    const int hx = blockIdx.y;
    int z = blockIdx.z;
    const int ct = z % ctiles; z /= ctiles;
    const int dx = z % g.dout;
    const int n = z / g.dout;
    const int64_t in_plane = static_cast<int64_t>(g.hi) * g.wi;
    const float* yn = yin + static_cast<int64_t>(n) * g.cin * g.di * in_plane;
    const float* wc = w + static_cast<int64_t>(ct) * g.v * g.cin * g.r * g.s * KT;
    float acc[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) acc[i] = 0.f;
    const bool active = wx < g.wo;
    for (int v = 0; v < g.v; ++v) {
        const int td = dx + g.pd - v;
        if (td < 0 || td % g.sd) continue;
        const int dy = td / g.sd;
        if (dy >= g.di) continue;
        for (int r = 0; r < g.r; ++r) {
            const int th = hx + g.ph - r;
            if (th < 0 || th % g.sh) continue;
            const int hy = th / g.sh;
            if (hy >= g.hi) continue;
            for (int s = 0; s < g.s; ++s) {
                const int tw = wx + g.pw - s;
                const bool ok = active && tw >= 0 && (tw % g.sw) == 0 && (tw / g.sw) < g.wi;
                const int wy = ok ? tw / g.sw : 0;
                for (int k = 0; k < g.cin; ++k) {
                    const float yv = ok ? __ldg(yn + (static_cast<int64_t>(k) * g.di + dy) * in_plane + static_cast<int64_t>(hy) * g.wi + wy) : 0.f;
                    const float* wq = wc + ((static_cast<int64_t>(v) * g.cin + k) * g.r * g.s + r * g.s + s) * KT;
#pragma unroll
                    for (int q = 0; q < KT / 4; ++q) {
                        const float4 wv = ldw(wq + 4 * q);
                        acc[4 * q + 0] = fmaf(wv.x, yv, acc[4 * q + 0]);
                        acc[4 * q + 1] = fmaf(wv.y, yv, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(wv.z, yv, acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(wv.w, yv, acc[4 * q + 3]);
                    }
                }
            }
        }
    }
    if (!active) return;
    const int64_t out_plane = static_cast<int64_t>(g.ho) * g.wo;
    const int64_t nout = static_cast<int64_t>(n) * g.dout * g.cout * out_plane;
#pragma unroll
    for (int i = 0; i < KT; ++i) {
        const int c = ct * KT + i;
        if (c >= g.cout) break;
        const int64_t idx = nout + (static_cast<int64_t>(dx) * g.cout + c) * out_plane + static_cast<int64_t>(hx) * g.wo + wx;
        float vv = acc[i] + bias[c];
        if (skip) vv += skip[idx];
        if (g.fuse_elu) vv = elu1(vv);
        x[idx] = vv;
    }
}

ConvGeom make_geom(const rt_conv3d_plan* p) {
    const rt_conv3d_desc& d = p->desc;
    ConvGeom g{};
    g.v = d.v; g.r = d.r; g.s = d.s;
    g.sd = d.stride[0]; g.sh = d.stride[1]; g.sw = d.stride[2];
    g.pd = d.pad[0]; g.ph = d.pad[1]; g.pw = d.pad[2];
    g.fuse_elu = d.fuse_elu;
    g.out_transposed = d.out_transposed;
    if (!d.transposed) {
        g.cin = d.c; g.cout = d.k;
        g.di = d.in_dims[0]; g.hi = d.in_dims[2]; g.wi = d.in_dims[3];
        g.dout = d.out_dims[1]; g.ho = d.out_dims[2]; g.wo = d.out_dims[3];
    } else {
        g.cin = d.k; g.cout = d.c;
        g.di = d.in_dims[1]; g.hi = d.in_dims[2]; g.wi = d.in_dims[3];
        g.dout = p->out_planes; g.ho = d.out_dims[2]; g.wo = d.out_dims[3];
    }
    return g;
}

}  // namespace

int simt_conv3d_enqueue(const rt_conv3d_plan* p, int n, const float* x, const float* skip, float* y, cudaStream_t s) {
    const ConvGeom g = make_geom(p);
    const int tiles = (g.cout + KT - 1) / KT;
    const int64_t gz = static_cast<int64_t>(n) * g.dout * tiles;
    if (gz > 65535 * 32 || g.ho > 65535) return RT_ERR_UNSUPPORTED;
    // gridDim.z is limited to 65535: fold the overflow into several launches over n.
    const int per_launch_n = gz <= 65535 ? n : static_cast<int>(65535 / (static_cast<int64_t>(g.dout) * tiles));
    if (per_launch_n < 1) return RT_ERR_UNSUPPORTED;
    const int64_t in_elems = static_cast<int64_t>(p->desc.in_dims[0]) * p->desc.in_dims[1] * p->desc.in_dims[2] * p->desc.in_dims[3];
    const int64_t out_elems = static_cast<int64_t>(g.cout) * g.dout * g.ho * g.wo;
    for (int n0 = 0; n0 < n; n0 += per_launch_n) {
        const int nn = (n - n0) < per_launch_n ? (n - n0) : per_launch_n;
        dim3 grid((g.wo + kThreads - 1) / kThreads, g.ho, nn * g.dout * tiles);
        const float* xs = x + n0 * in_elems;
        const float* ss = skip ? skip + n0 * out_elems : nullptr;
        float* ys = y + n0 * out_elems;
        if (!p->desc.transposed)
            conv3d_fwd_simt_kernel<<<grid, kThreads, 0, s>>>(xs, p->w_simt, p->bias, ss, ys, g, tiles);
        else
            conv3d_bwd_simt_kernel<<<grid, kThreads, 0, s>>>(xs, p->w_simt, p->bias, ss, ys, g, tiles);
        note_launch(p->desc.transposed ? "conv3d_transpose_simt" : "conv3d_simt");
        RT_CHECK_LAUNCH();
    }
    return RT_OK;
}

}  // namespace rt

using namespace rt;

namespace {

float half_bits_to_float(uint16_t h) {
    const uint32_t sign = (h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1F, man = h & 0x3FF, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            int e = -1;
            do { man <<= 1; ++e; } while (!(man & 0x400));
            bits = sign | ((127 - 15 - e) << 23) | ((man & 0x3FF) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

}  // namespace

namespace rt {
void host_to_f32(int dtype, const void* src, int64_t count, std::vector<float>& dst) {
    dst.resize(count);
    if (dtype == RT_F32) memcpy(dst.data(), src, count * 4);
    else {
        const uint16_t* h = static_cast<const uint16_t*>(src);
        for (int64_t i = 0; i < count; ++i) dst[i] = half_bits_to_float(h[i]);
    }
}
}  // namespace rt

extern "C" {

static int conv3d_create_part(const rt_conv3d_desc* d, int out_c_total, int out_c_offset, bool reuse_pack, rt_conv3d_plan** out);

int rt_conv3d_create(const rt_conv3d_desc* d, rt_conv3d_plan** out) {
    if (!d || !out || !d->weights) return RT_ERR_ARG;
    // Forward convolutions with more than 128 output channels on the tensor-core path: <= 128-channel parts (conv3d_internal.h).
    if (!d->transposed && d->k > 128 && d->precision != RT_PREC_SIMT && !d->fuse_softargmax && d->k % 8 == 0) {
        rt_conv3d_plan* top = new rt_conv3d_plan();
        top->desc = *d;
        top->desc.weights = nullptr; top->desc.bias = nullptr;
        top->cout = d->k; top->cin = d->c;
        top->out_planes = d->out_dims[1];
        const size_t es = d->weights_dtype == RT_F16 ? 2 : 4;
        const size_t per_k = static_cast<size_t>(d->v) * d->c * d->r * d->s;
        for (int k0 = 0; k0 < d->k; k0 += 128) {
            rt_conv3d_desc pd = *d;
            pd.k = d->k - k0 < 128 ? d->k - k0 : 128;
            pd.out_dims[0] = pd.k;
            pd.weights = static_cast<const char*>(d->weights) + static_cast<size_t>(k0) * per_k * es;
            pd.bias = d->bias ? static_cast<const char*>(d->bias) + static_cast<size_t>(k0) * es : nullptr;
            std::vector<float> act_part;
            if (d->act_params) {                 // this part's slice of the [4][K] activation parameters
                act_part.resize(static_cast<size_t>(4) * pd.k);
                for (int j = 0; j < 4; ++j)
                    for (int k = 0; k < pd.k; ++k) act_part[static_cast<size_t>(j) * pd.k + k] = d->act_params[static_cast<size_t>(j) * d->k + k0 + k];
                pd.act_params = act_part.data();
            }
            rt_conv3d_plan* part = nullptr;
            const int rc = conv3d_create_part(&pd, d->k, k0, k0 > 0, &part);
            if (rc != RT_OK) { rt_conv3d_destroy(top); return rc; }
            top->parts.push_back(part);
        }
        *out = top;
        return RT_OK;
    }
    return conv3d_create_part(d, 0, 0, false, out);
}

static int conv3d_create_part(const rt_conv3d_desc* d, int out_c_total, int out_c_offset, bool reuse_pack, rt_conv3d_plan** out) {
    if (!d || !out || !d->weights) return RT_ERR_ARG;
    if (d->k <= 0 || d->v <= 0 || d->c <= 0 || d->r <= 0 || d->s <= 0) return RT_ERR_ARG;
    for (int i = 0; i < 3; ++i)
        if (d->stride[i] <= 0 || d->pad[i] < 0) return RT_ERR_ARG;
    if (d->weights_dtype != RT_F32 && d->weights_dtype != RT_F16) return RT_ERR_UNSUPPORTED;
    // Shape consistency, as the plugins assert it (conv3d_plugin.cpp:82-100, conv3d_transpose_plugin.cpp:104-108).
    const int* conv_in = d->transposed ? d->out_dims : d->in_dims;     // [D, C, H, W] of the forward conv
    const int* conv_out = d->transposed ? d->in_dims : d->out_dims;    // [K, Do, Ho, Wo]
    if (conv_in[1] != d->c || conv_out[0] != d->k) return RT_ERR_ARG;
    const int kk[3] = {d->v, d->r, d->s};
    if (d->pad_end_d < 0 || (d->transposed && d->pad_end_d != 0)) return RT_ERR_ARG;
    const int sp_in[3] = {conv_in[0] + d->pad_end_d, conv_in[2], conv_in[3]};
    for (int i = 0; i < 3; ++i) {
        const int span = sp_in[i] + 2 * d->pad[i] - kk[i];
        if (span < 0 || span / d->stride[i] + 1 != conv_out[1 + i]) return RT_ERR_ARG;
    }
    if (d->transposed && (d->slice_d < 0 || d->slice_d >= d->out_dims[0])) return RT_ERR_ARG;
    if (!d->transposed && d->slice_d != 0) return RT_ERR_ARG;
    if (d->transposed && d->out_transposed) return RT_ERR_ARG;
    if (d->fuse_softargmax < 0 || d->fuse_softargmax > 2) return RT_ERR_ARG;
    if (d->fuse_softargmax && !dsa_shape_supported(*d)) return RT_ERR_UNSUPPORTED;     // one kernel covers the fused form

    rt_conv3d_plan* p = new rt_conv3d_plan();
    p->desc = *d;
    p->desc.weights = nullptr;
    p->desc.bias = nullptr;
    p->out_c_total = out_c_total; p->out_c_offset = out_c_offset; p->reuse_pack = reuse_pack;
    if (d->act_params) {
        if (d->transposed || d->precision == RT_PREC_SIMT || d->fuse_softargmax) { delete p; return RT_ERR_UNSUPPORTED; }
        p->act_host.assign(d->act_params, d->act_params + static_cast<size_t>(4) * d->k);
    }
    p->desc.act_params = nullptr;
    p->out_planes = d->transposed ? d->out_dims[0] - d->slice_d : d->out_dims[1];
    p->cout = d->transposed ? d->c : d->k;
    p->cin = d->transposed ? d->k : d->c;

    const int64_t wcount = static_cast<int64_t>(d->k) * d->v * d->c * d->r * d->s;
    std::vector<float> w, b;
    host_to_f32(d->weights_dtype, d->weights, wcount, w);
    if (d->bias) host_to_f32(d->weights_dtype, d->bias, p->cout, b);
    else b.assign(p->cout, 0.f);

    // SIMT packing: [ceil(Cout/16)][V][Cin][R][S][16], zero padded.
    const int tiles = (p->cout + 15) / 16;
    std::vector<float> pk(static_cast<size_t>(tiles) * d->v * p->cin * d->r * d->s * 16, 0.f);
    for (int k = 0; k < d->k; ++k)
        for (int v = 0; v < d->v; ++v)
            for (int c = 0; c < d->c; ++c)
                for (int r = 0; r < d->r; ++r)
                    for (int s = 0; s < d->s; ++s) {
                        const float val = w[(((static_cast<int64_t>(k) * d->v + v) * d->c + c) * d->r + r) * d->s + s];
                        const int co = d->transposed ? c : k, ci = d->transposed ? k : c;
                        const int64_t idx = ((((static_cast<int64_t>(co / 16) * d->v + v) * p->cin + ci) * d->r + r) * d->s + s) * 16 + (co % 16);
                        pk[idx] = val;
                    }
    cudaError_t e = cudaMalloc(&p->w_simt, pk.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(p->w_simt, pk.data(), pk.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&p->bias, b.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(p->bias, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        rt_conv3d_destroy(p);
        return static_cast<int>(e);
    }
    if (d->precision == RT_PREC_SIMT && (d->in_layout != RT_LAYOUT_DENSE || d->out_layout != RT_LAYOUT_DENSE)) {
        rt_conv3d_destroy(p);
        return RT_ERR_UNSUPPORTED;      // the CUDA-core validation kernels only speak the dense plugin layouts
    }
    if (d->fuse_softargmax) {
        const int rc = dsa_plan_init(p, w, b);
        if (rc != RT_OK) {
            rt_conv3d_destroy(p);
            return rc;
        }
        *out = p;
        return RT_OK;
    }
    if (d->precision == RT_PREC_SIMT && out_c_total != 0) { rt_conv3d_destroy(p); return RT_ERR_UNSUPPORTED; }
    if (d->precision != RT_PREC_SIMT) {
        const int rc = tc_plan_init(p, w, b);
        if (rc != RT_OK) {           // no silent downgrade of a requested tensor-core precision
            rt_conv3d_destroy(p);
            return rc;
        }
        if (ds_shape_supported(*d)) {
            const int rc2 = ds_plan_init(p, w);
            if (rc2 != RT_OK && rc2 != RT_ERR_UNSUPPORTED) {
                rt_conv3d_destroy(p);
                return rc2;
            }
        }
    }
    *out = p;
    return RT_OK;
}

int rt_conv3d_tc_supported(const rt_conv3d_desc* d) {
    if (!d) return 0;
    if (d->fuse_softargmax) return dsa_shape_supported(*d) ? 1 : 0;     // the fused form has exactly one kernel
    return tc_shape_supported(*d) ? 1 : 0;
}

void rt_conv3d_destroy(rt_conv3d_plan* p) {
    if (!p) return;
    for (rt_conv3d_plan* part : p->parts) rt_conv3d_destroy(part);
    tc_plan_destroy(p);
    ds_plan_destroy(p);
    dsa_plan_destroy(p);
    cudaFree(p->w_simt);
    cudaFree(p->bias);
    delete p;
}

size_t rt_conv3d_workspace_size(const rt_conv3d_plan* p, int max_batch) {
    if (p && !p->parts.empty()) return rt_conv3d_workspace_size(p->parts[0], max_batch);    // same input, same packed planes
    if (!p || !p->tc) return 0;
    return tc_workspace_size(p, max_batch);
}

int rt_conv3d_enqueue(const rt_conv3d_plan* p, int n, const void* x, const void* skip, void* y, void* workspace,
                      void* stream) {
    if (!p || !x || !y || n < 0) return RT_ERR_ARG;
    if (n == 0) return RT_OK;
    if (!p->parts.empty()) {
        for (const rt_conv3d_plan* part : p->parts) {
            const int rc = rt_conv3d_enqueue(part, n, x, skip, y, workspace, stream);
            if (rc != RT_OK) return rc;
        }
        return RT_OK;
    }
    if (p->dsa) return skip ? RT_ERR_ARG : dsa_enqueue(p, n, x, y, as_stream(stream));
    if (p->ds && !skip) return ds_conv3d_enqueue(p, n, x, y, as_stream(stream));
    if (p->tc)
        return tc_conv3d_enqueue(p, n, static_cast<const float*>(x), static_cast<const float*>(skip),
                                 static_cast<float*>(y), workspace, as_stream(stream));
    return simt_conv3d_enqueue(p, n, static_cast<const float*>(x), static_cast<const float*>(skip),
                               static_cast<float*>(y), as_stream(stream));
}

}  // extern "C"

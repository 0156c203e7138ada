// 2-D convolution / deconvolution of the feature towers -- the TensorRT-native IConvolutionLayer / IDeconvolutionLayer
// the generated builders call (sample_app/nvsmall_1025x321_net.cpp:48-165; resnet18_2D_513x257_net.cpp:613).
// 1.1 % of NVSmall's FLOPs (12.9 GFLOP): fp32 CUDA-core direct convolution, one output x per thread, 16 output
// channels per thread in registers, weights broadcast from L1, optional fused ELU (the builders always follow a tower
// conv with an ELU plugin).   x [n, Cin, H, W] -> y [n, Cout, Ho, Wo], fp32.
#include <cstring>
#include <vector>

#include "common.cuh"

struct rt_conv2d_plan {
    rt_conv2d_desc desc;
    int out_h, out_w;
    float* w = nullptr;     // [ceil(Cout/16)][Cin][R][S][16]
    float* bias = nullptr;  // [Cout]
};

namespace rt {
namespace {

constexpr int KT = 16;
constexpr int kThreads = 128;

struct Geom2 {
    int cin, cout, r, s, sh, sw, ph, pw, hi, wi, ho, wo, fuse_elu;
};

__device__ __forceinline__ float4 ldw(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <bool TRANSPOSED>
__global__ void __launch_bounds__(kThreads)
conv2d_simt_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ y, Geom2 g, int tiles) {
    const int wo = blockIdx.x * kThreads + threadIdx.x;
    const int ho = blockIdx.y;
    const int kt = blockIdx.z % tiles, n = blockIdx.z / tiles;
    const int64_t in_plane = static_cast<int64_t>(g.hi) * g.wi;
    const float* xn = x + static_cast<int64_t>(n) * g.cin * in_plane;
    const float* wk = w + static_cast<int64_t>(kt) * g.cin * g.r * g.s * KT;
    float acc[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) acc[i] = 0.f;
    const bool active = wo < g.wo;
    for (int r = 0; r < g.r; ++r) {
        int h_in;
        if (!TRANSPOSED) {
            h_in = ho * g.sh + r - g.ph;
            if (h_in < 0 || h_in >= g.hi) continue;
        } else {
            const int t = ho + g.ph - r;
            if (t < 0 || t % g.sh) continue;
            h_in = t / g.sh;
            if (h_in >= g.hi) continue;
        }
        for (int s = 0; s < g.s; ++s) {
            int w_in;
            bool ok = active;
            if (!TRANSPOSED) {
                w_in = wo * g.sw + s - g.pw;
                ok = ok && w_in >= 0 && w_in < g.wi;
            } else {
                const int t = wo + g.pw - s;
                ok = ok && t >= 0 && (t % g.sw) == 0 && (t / g.sw) < g.wi;
                w_in = t / g.sw;
            }
            if (!ok) w_in = 0;
            const float* xp = xn + static_cast<int64_t>(h_in) * g.wi + w_in;
            const float* wq = wk + (r * g.s + s) * KT;
            for (int c = 0; c < g.cin; ++c) {
                const float xv = ok ? __ldg(xp + c * in_plane) : 0.f;
                const float* wc = wq + static_cast<int64_t>(c) * g.r * g.s * KT;
#pragma unroll
                for (int q = 0; q < KT / 4; ++q) {
                    const float4 wv = ldw(wc + 4 * q);
                    acc[4 * q + 0] = fmaf(wv.x, xv, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(wv.y, xv, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(wv.z, xv, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(wv.w, xv, acc[4 * q + 3]);
                }
            }
        }
    }
    if (!active) return;
    const int64_t out_plane = static_cast<int64_t>(g.ho) * g.wo;
#pragma unroll
    for (int i = 0; i < KT; ++i) {
        const int k = kt * KT + i;
        if (k >= g.cout) break;
        float v = acc[i] + bias[k];
        if (g.fuse_elu) v = elu1(v);
        y[(static_cast<int64_t>(n) * g.cout + k) * out_plane + static_cast<int64_t>(ho) * g.wo + wo] = v;
    }
}

float h2f(uint16_t h) {
    __half_raw r;
    r.x = h;
    return __half2float(__half(r));
}

}  // namespace
}  // namespace rt

using namespace rt;

extern "C" {

int rt_conv2d_create(const rt_conv2d_desc* d, rt_conv2d_plan** out) {
    if (!d || !out || !d->weights || d->cin <= 0 || d->cout <= 0 || d->r <= 0 || d->s <= 0) return RT_ERR_ARG;
    if (d->stride[0] <= 0 || d->stride[1] <= 0 || d->pad[0] < 0 || d->pad[1] < 0 || d->in_h <= 0 || d->in_w <= 0) return RT_ERR_ARG;
    if (d->weights_dtype != RT_F32 && d->weights_dtype != RT_F16) return RT_ERR_UNSUPPORTED;
    rt_conv2d_plan* p = new rt_conv2d_plan();
    p->desc = *d;
    p->desc.weights = p->desc.bias = nullptr;
    if (!d->transposed) {
        p->out_h = (d->in_h + 2 * d->pad[0] - d->r) / d->stride[0] + 1;
        p->out_w = (d->in_w + 2 * d->pad[1] - d->s) / d->stride[1] + 1;
    } else {
        p->out_h = (d->in_h - 1) * d->stride[0] + d->r - 2 * d->pad[0];
        p->out_w = (d->in_w - 1) * d->stride[1] + d->s - 2 * d->pad[1];
    }
    if (p->out_h <= 0 || p->out_w <= 0) { delete p; return RT_ERR_ARG; }
    const int64_t wcount = static_cast<int64_t>(d->cin) * d->cout * d->r * d->s;
    std::vector<float> w(wcount), b(d->cout, 0.f);
    if (d->weights_dtype == RT_F32) {
        memcpy(w.data(), d->weights, wcount * 4);
        if (d->bias) memcpy(b.data(), d->bias, d->cout * 4);
    } else {
        for (int64_t i = 0; i < wcount; ++i) w[i] = h2f(static_cast<const uint16_t*>(d->weights)[i]);
        if (d->bias) for (int i = 0; i < d->cout; ++i) b[i] = h2f(static_cast<const uint16_t*>(d->bias)[i]);
    }
    const int tiles = (d->cout + KT - 1) / KT;
    std::vector<float> pk(static_cast<size_t>(tiles) * d->cin * d->r * d->s * KT, 0.f);
    for (int k = 0; k < d->cout; ++k)
        for (int c = 0; c < d->cin; ++c)
            for (int r = 0; r < d->r; ++r)
                for (int s = 0; s < d->s; ++s) {
                    // conv: KCRS ; deconv: [Cin][Cout][R][S]
                    const int64_t src = d->transposed ? ((static_cast<int64_t>(c) * d->cout + k) * d->r + r) * d->s + s
                                                      : ((static_cast<int64_t>(k) * d->cin + c) * d->r + r) * d->s + s;
                    pk[(((static_cast<int64_t>(k / KT) * d->cin + c) * d->r + r) * d->s + s) * KT + k % KT] = w[src];
                }
    cudaError_t e = cudaMalloc(&p->w, pk.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(p->w, pk.data(), pk.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&p->bias, b.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(p->bias, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { rt_conv2d_destroy(p); return static_cast<int>(e); }
    *out = p;
    return RT_OK;
}

void rt_conv2d_destroy(rt_conv2d_plan* p) {
    if (!p) return;
    cudaFree(p->w);
    cudaFree(p->bias);
    delete p;
}

void rt_conv2d_out_dims(const rt_conv2d_plan* p, int* out_h, int* out_w) {
    if (out_h) *out_h = p->out_h;
    if (out_w) *out_w = p->out_w;
}

int rt_conv2d_enqueue(const rt_conv2d_plan* p, int n, const void* x, void* y, void* stream) {
    if (!p || !x || !y || n < 0) return RT_ERR_ARG;
    if (n == 0) return RT_OK;
    const rt_conv2d_desc& d = p->desc;
    Geom2 g{d.cin, d.cout, d.r, d.s, d.stride[0], d.stride[1], d.pad[0], d.pad[1], d.in_h, d.in_w, p->out_h, p->out_w, d.fuse_elu};
    const int tiles = (d.cout + KT - 1) / KT;
    if (static_cast<int64_t>(n) * tiles > 65535 || p->out_h > 65535) return RT_ERR_UNSUPPORTED;
    dim3 grid((p->out_w + kThreads - 1) / kThreads, p->out_h, n * tiles);
    if (!d.transposed)
        conv2d_simt_kernel<false><<<grid, kThreads, 0, as_stream(stream)>>>(static_cast<const float*>(x), p->w, p->bias, static_cast<float*>(y), g, tiles);
    else
        conv2d_simt_kernel<true><<<grid, kThreads, 0, as_stream(stream)>>>(static_cast<const float*>(x), p->w, p->bias, static_cast<float*>(y), g, tiles);
    note_launch(d.transposed ? "deconv2d_simt" : "conv2d_simt");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

}  // extern "C"

// The non-convolution layers of the TrailNet S-ResNet-18 classifier (models/pretrained/TrailNet_SResNet-18.prototxt, run by
// ros/packages/caffe_ros/src/tensor_net.cpp through TensorRT's Caffe parser): per-channel Scale, ReLU, the S-ReLU chain
// Scale(+b1) -> ReLU -> Scale(+b2) as one pass, Caffe-style pooling (ceil-mode output size, windows clipped at the border),
// InnerProduct and Softmax over channels.  All of them are HBM-bound element-wise / reduction passes over dense fp32 NCHW
// tensors; 99.8 % of the network's FLOPs are in the convolutions, which run on the tcgen05 kernel (conv3d_tc.cu, V = D = 1).
#include "common.cuh"

namespace rt {
namespace {

// y = (x * scale[c] + shift[c]);  RELU_CHAIN: y = max(x * s1[c] + b1[c], 0) * s2[c] + b2[c]   (S-ReLU, prototxt:54-105)
template <bool RELU_CHAIN>
__global__ void __launch_bounds__(256)
scale_channel_kernel(const float* __restrict__ x, float* __restrict__ y, int c, int64_t hw, int64_t total,
                     const float* __restrict__ s1, const float* __restrict__ b1, const float* __restrict__ s2,
                     const float* __restrict__ b2) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>((i / hw) % c);
        float v = x[i];
        v = s1 ? fmaf(v, __ldg(s1 + ch), b1 ? __ldg(b1 + ch) : 0.f) : v + (b1 ? __ldg(b1 + ch) : 0.f);
        if (RELU_CHAIN) {
            v = fmaxf(v, 0.f);
            v = s2 ? fmaf(v, __ldg(s2 + ch), b2 ? __ldg(b2 + ch) : 0.f) : v + (b2 ? __ldg(b2 + ch) : 0.f);
        }
        y[i] = v;
    }
}

__global__ void __launch_bounds__(256) relu_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        y[i] = fmaxf(x[i], 0.f);
}

// Caffe PoolingLayer (pooling_layer.cpp): window [i*stride - pad, +k) clipped to the padded image for the AVE divisor and to
// the image for the values; MAX ignores the padding.  One thread per output element, w fastest.
__global__ void __launch_bounds__(256)
pool2d_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t planes, int h, int w, int oh, int ow, int k, int stride,
              int pad, int is_max) {
    const int64_t total = planes * oh * ow;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(i % ow), oy = static_cast<int>((i / ow) % oh);
        const int64_t pl = i / (static_cast<int64_t>(ow) * oh);
        int hs = oy * stride - pad, ws = ox * stride - pad;
        const int he = min(hs + k, h + pad), we = min(ws + k, w + pad);
        const int area = (he - hs) * (we - ws);
        const int h0 = max(hs, 0), w0 = max(ws, 0), h1 = min(he, h), w1 = min(we, w);
        const float* src = x + pl * h * w;
        float acc = is_max ? -INFINITY : 0.f;
        for (int yy = h0; yy < h1; ++yy)
            for (int xx = w0; xx < w1; ++xx) {
                const float v = __ldg(src + static_cast<int64_t>(yy) * w + xx);
                acc = is_max ? fmaxf(acc, v) : acc + v;
            }
        y[i] = is_max ? acc : acc / static_cast<float>(area);
    }
}

// y[n, m] = b[m] + sum_k x[n, k] * W[m, k]: one CTA per (n, m); the TrailNet heads are 16384 -> 3.
__global__ void __launch_bounds__(256)
fully_connected_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ y,
                       int k, int m) {
    const int n = blockIdx.y, mo = blockIdx.x;
    const float* xr = x + static_cast<int64_t>(n) * k;
    const float* wr = w + static_cast<int64_t>(mo) * k;
    float acc = 0.f;
    for (int i = threadIdx.x; i < k; i += blockDim.x) acc = fmaf(__ldg(xr + i), __ldg(wr + i), acc);
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        y[static_cast<int64_t>(n) * m + mo] = s + (b ? b[mo] : 0.f);
    }
}

// Softmax over the channel dimension of [n, c, inner] (max-subtracted, as Caffe's SoftmaxLayer and cuDNN ACCURATE do).
__global__ void __launch_bounds__(256)
softmax_channels_kernel(const float* __restrict__ x, float* __restrict__ y, int c, int64_t inner, int64_t total) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t n = i / inner, p = i % inner;
        const float* src = x + n * c * inner + p;
        float m = -INFINITY;
        for (int ch = 0; ch < c; ++ch) m = fmaxf(m, src[ch * inner]);
        float s = 0.f;
        for (int ch = 0; ch < c; ++ch) s += expf(src[ch * inner] - m);
        float* dst = y + n * c * inner + p;
        for (int ch = 0; ch < c; ++ch) dst[ch * inner] = expf(src[ch * inner] - m) / s;
    }
}

// im2col for 2-D convolutions with filters larger than 3x3 (TrailNet conv1: 7x7, stride 2, 3 channels): x [n,c,h,w] fp32 ->
// RT_LAYOUT_SPLIT16 matrix [n][hi|lo][1][ho][wo][kp] with K index = (ci * r + ri) * s + si (the KCRS order of the weights), zero
// beyond c*r*s.  The convolution itself is then a 1x1 convolution over kp "channels" on the tcgen05 kernel.
// One CTA = one output row of one sample: the c*r input rows it reads are staged in shared memory once (each input element is
// used by up to r*s/stride^2 outputs), a table maps K -> (offset in the tile, filter column); one thread-item = one output
// position x 8 consecutive K = one 16-byte vector per fp16 plane, K groups fastest so that the stores of a warp are contiguous.
__global__ void __launch_bounds__(256)
im2col_split16_kernel(const float* __restrict__ x, __half* __restrict__ y, int c, int h, int w, int r, int s, int stride, int pad,
                      int ho, int wo, int kp) {
    extern __shared__ float im_tile[];                     // [c*r][w] input rows, then int2 table[kp]
    int2* tab = reinterpret_cast<int2*>(im_tile + ((static_cast<size_t>(c) * r * w + 1) & ~static_cast<size_t>(1)));      // 8-byte aligned
    const int oy = blockIdx.x, n = blockIdx.y;
    const int krs = c * r * s, groups = kp >> 3;
    const float* xn = x + static_cast<int64_t>(n) * c * h * w;
    for (int i = threadIdx.x; i < c * r * w; i += blockDim.x) {
        const int ix = i % w, row = i / w, ri = row % r, ci = row / r;
        const int iy = oy * stride - pad + ri;
        im_tile[i] = (iy >= 0 && iy < h) ? __ldg(xn + (static_cast<int64_t>(ci) * h + iy) * w + ix) : 0.f;
    }
    for (int k = threadIdx.x; k < kp; k += blockDim.x) {
        const int si = k % s, row = k / s;                 // row = ci * r + ri
        tab[k] = k < krs ? make_int2(row * w, si) : make_int2(-1, 0);
    }
    __syncthreads();
    const int64_t plane = static_cast<int64_t>(ho) * wo * kp;          // halves of one fp16 plane per sample
    __half* yn = y + static_cast<int64_t>(n) * 2 * plane + static_cast<int64_t>(oy) * wo * kp;
    for (int i = threadIdx.x; i < wo * groups; i += blockDim.x) {
        const int g = i % groups, ox = i / groups;
        const int x0 = ox * stride - pad;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int2 t = tab[g * 8 + j];
            const int ix = x0 + t.y;
            v[j] = (t.x >= 0 && ix >= 0 && ix < w) ? im_tile[t.x + ix] : 0.f;
        }
        const int64_t o = static_cast<int64_t>(ox) * kp + g * 8;
        split_store8(v, yn + o, yn + plane + o);
    }
}

inline int grid_for(int64_t total) {
    const int64_t blocks = ceil_div(total, 256);
    const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
    return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace
}  // namespace rt

using namespace rt;

extern "C" int rt_scale_channel(const void* x, void* y, int n, int c, int64_t hw, const float* scale, const float* shift, void* stream) {
    if (!x || !y || n < 0 || c <= 0 || hw < 0) return RT_ERR_ARG;
    const int64_t total = static_cast<int64_t>(n) * c * hw;
    if (total == 0) return RT_OK;
    scale_channel_kernel<false><<<grid_for(total), 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), static_cast<float*>(y), c, hw, total,
                                                                              scale, shift, nullptr, nullptr);
    note_launch("scale_channel");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_srelu(const void* x, void* y, int n, int c, int64_t hw, const float* s1, const float* b1, const float* s2,
                        const float* b2, void* stream) {
    if (!x || !y || n < 0 || c <= 0 || hw < 0) return RT_ERR_ARG;
    const int64_t total = static_cast<int64_t>(n) * c * hw;
    if (total == 0) return RT_OK;
    scale_channel_kernel<true><<<grid_for(total), 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), static_cast<float*>(y), c, hw, total,
                                                                             s1, b1, s2, b2);
    note_launch("srelu");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_relu(const void* x, void* y, int64_t count, void* stream) {
    if (!x || !y || count < 0) return RT_ERR_ARG;
    if (count == 0) return RT_OK;
    relu_kernel<<<grid_for(count), 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), static_cast<float*>(y), count);
    note_launch("relu");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_pool2d(const void* x, void* y, int n, int c, int h, int w, int out_h, int out_w, int k, int stride, int pad, int is_max,
                         void* stream) {
    if (!x || !y || n < 0 || c <= 0 || h <= 0 || w <= 0 || out_h <= 0 || out_w <= 0 || k <= 0 || stride <= 0 || pad < 0) return RT_ERR_ARG;
    // every window must start inside the padded image and contain at least one image element
    if ((out_h - 1) * stride - pad >= h || (out_w - 1) * stride - pad >= w) return RT_ERR_ARG;
    const int64_t planes = static_cast<int64_t>(n) * c;
    if (planes == 0) return RT_OK;
    pool2d_kernel<<<grid_for(planes * out_h * out_w), 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), static_cast<float*>(y), planes, h, w,
                                                                                 out_h, out_w, k, stride, pad, is_max);
    note_launch(is_max ? "pool2d_max" : "pool2d_avg");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_fully_connected(const void* x, const float* w, const float* b, void* y, int n, int k, int m, void* stream) {
    if (!x || !w || !y || n < 0 || k <= 0 || m <= 0 || n > 65535) return RT_ERR_ARG;
    if (n == 0) return RT_OK;
    fully_connected_kernel<<<dim3(m, n), 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), w, b, static_cast<float*>(y), k, m);
    note_launch("fully_connected");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_softmax_channels(const void* x, void* y, int n, int c, int64_t inner, void* stream) {
    if (!x || !y || n < 0 || c <= 0 || inner <= 0) return RT_ERR_ARG;
    const int64_t total = static_cast<int64_t>(n) * inner;
    if (total == 0) return RT_OK;
    softmax_channels_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), static_cast<float*>(y), c, inner, total);
    note_launch("softmax_channels");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_im2col_split16(const void* x, void* y, int n, int c, int h, int w, int r, int s, int stride, int pad, int out_h, int out_w,
                                 int kp, void* stream) {
    if (!x || !y || n < 0 || c <= 0 || h <= 0 || w <= 0 || r <= 0 || s <= 0 || stride <= 0 || pad < 0 || out_h <= 0 || out_w <= 0) return RT_ERR_ARG;
    if (kp % 8 != 0 || kp < c * r * s || n > 65535) return RT_ERR_ARG;
    if (n == 0) return RT_OK;
    const size_t smem = ((static_cast<size_t>(c) * r * w + 1) & ~static_cast<size_t>(1)) * sizeof(float) + static_cast<size_t>(kp) * sizeof(int2);
    if (smem > 200 * 1024) return RT_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) RT_CUDA(cudaFuncSetAttribute(im2col_split16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    im2col_split16_kernel<<<dim3(out_h, n), 256, smem, as_stream(stream)>>>(static_cast<const float*>(x), static_cast<__half*>(y), c, h, w, r, s, stride,
                                                                           pad, out_h, out_w, kp);
    note_launch("im2col_split16");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

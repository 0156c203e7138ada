// Element-wise and data-movement ops of the plugin path (all HBM-bound, no reuse):
//   ELU (lib/elu_plugin.cpp:123-135), Padding (lib/padding_plugin.cpp:79-94), Slice (lib/slice_plugin.cpp:80-92),
//   Transform{1,0,2,3} (lib/transform_plugin.cpp:94-108), fp32<->fp16 (lib/kernels.cu:340-375) and the TRT-native
//   scale / element-wise sum / channel concat / sigmoid layers the generated builders call.
// Design: grid-stride kernels sized to a multiple of the SM count, 128-bit accesses on the 16 B-aligned body and a
// scalar head/tail (dense plugin tensors have odd extents, so the alignment of a sub-tensor is arbitrary).
#include "common.cuh"

namespace rt {

std::atomic<uint64_t> g_launches{0};
const char* g_last_kernel = "";

namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t work_items, int per_sm = 8) {
    int64_t blocks = ceil_div(work_items, kThreads);
    int64_t cap = static_cast<int64_t>(num_sms()) * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

enum UnaryOp { kElu = 0, kSigmoid = 1, kScale = 2 };

struct UnaryArgs { float shift, scale, power; };

template <int OP>
__device__ __forceinline__ float apply_unary(float v, const UnaryArgs& a) {
    if (OP == kElu) return elu1(v);
    if (OP == kSigmoid) return 1.f / (1.f + expf(-v));
    float y = v * a.scale + a.shift;
    return a.power == 1.f ? y : powf(y, a.power);
}

// y = f(x); body vectorised as 16-byte words when both pointers are 16 B aligned (checked on the host).
template <typename T, int OP, bool VEC>
__global__ void unary_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t count, UnaryArgs a) {
    constexpr int V = 16 / sizeof(T);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (VEC) {
        const int64_t nvec = count / V;
        for (int64_t v = i; v < nvec; v += stride) {
            uint4 raw = reinterpret_cast<const uint4*>(x)[v];
            T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
            for (int k = 0; k < V; ++k) e[k] = from_f32<T>(apply_unary<OP>(to_f32(e[k]), a));
            reinterpret_cast<uint4*>(y)[v] = raw;
        }
        for (int64_t t = nvec * V + i; t < count; t += stride) y[t] = from_f32<T>(apply_unary<OP>(to_f32(x[t]), a));
    } else {
        for (; i < count; i += stride) y[i] = from_f32<T>(apply_unary<OP>(to_f32(x[i]), a));
    }
}

template <typename T, int OP>
int launch_unary(const void* x, void* y, int64_t count, UnaryArgs a, cudaStream_t s, const char* name) {
    if (count <= 0) return RT_OK;
    constexpr int V = 16 / sizeof(T);
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const int grid = grid_for(vec ? ceil_div(count, V) : count);
    if (vec) unary_kernel<T, OP, true><<<grid, kThreads, 0, s>>>(static_cast<const T*>(x), static_cast<T*>(y), count, a);
    else     unary_kernel<T, OP, false><<<grid, kThreads, 0, s>>>(static_cast<const T*>(x), static_cast<T*>(y), count, a);
    note_launch(name);
    RT_CHECK_LAUNCH();
    return RT_OK;
}

template <int OP>
int dispatch_unary(int dtype, const void* x, void* y, int64_t count, UnaryArgs a, void* stream, const char* name) {
    if (!x || !y || count < 0) return RT_ERR_ARG;
    if (dtype == RT_F32) return launch_unary<float, OP>(x, y, count, a, as_stream(stream), name);
    if (dtype == RT_F16) return launch_unary<__half, OP>(x, y, count, a, as_stream(stream), name);
    return RT_ERR_UNSUPPORTED;
}

template <typename T>
__global__ void sum_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, int64_t count) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride)
        y[i] = from_f32<T>(to_f32(a[i]) + to_f32(b[i]));
}

template <typename S, typename D>
__global__ void convert_kernel(const S* __restrict__ x, D* __restrict__ y, int64_t count) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride)
        y[i] = from_f32<D>(to_f32(x[i]));
}

// Generic strided block copy used by pad / slice / concat:
//   for o in [0,outer): y[o*y_stride + y_off + i] = x[o*x_stride + x_off + i], i in [0,len); optional zero tail.
// Copies 16-byte words when src and dst of a row share alignment, scalars otherwise.
template <typename T>
__global__ void block_copy_kernel(const T* __restrict__ x, T* __restrict__ y, int outer, int64_t x_stride,
                                  int64_t y_stride, int64_t x_off, int64_t y_off, int64_t len, int64_t zero_len) {
    constexpr int V = 16 / sizeof(T);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (int o = 0; o < outer; ++o) {
        const T* src = x + o * x_stride + x_off;
        T* dst = y + o * y_stride + y_off;
        const uintptr_t sa = reinterpret_cast<uintptr_t>(src), da = reinterpret_cast<uintptr_t>(dst);
        if (len > 0) {
            if (((sa ^ da) & 15) == 0) {
                int64_t head = ((16 - (da & 15)) & 15) / sizeof(T);
                if (head > len) head = len;
                const int64_t nvec = (len - head) / V;
                for (int64_t i = tid; i < head; i += stride) dst[i] = src[i];
                const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
                uint4* d4 = reinterpret_cast<uint4*>(dst + head);
                for (int64_t i = tid; i < nvec; i += stride) d4[i] = s4[i];
                for (int64_t i = head + nvec * V + tid; i < len; i += stride) dst[i] = src[i];
            } else {
                for (int64_t i = tid; i < len; i += stride) dst[i] = src[i];
            }
        }
        T* z = dst + len;
        for (int64_t i = tid; i < zero_len; i += stride) z[i] = from_f32<T>(0.f);
    }
}

template <typename T>
int launch_block_copy(const void* x, void* y, int outer, int64_t x_stride, int64_t y_stride, int64_t x_off,
                      int64_t y_off, int64_t len, int64_t zero_len, cudaStream_t s, const char* name) {
    if (outer <= 0 || (len <= 0 && zero_len <= 0)) return RT_OK;
    const int64_t work = (len > zero_len ? len : zero_len) / (16 / sizeof(T)) + 1;
    block_copy_kernel<T><<<grid_for(work), kThreads, 0, s>>>(static_cast<const T*>(x), static_cast<T*>(y), outer,
                                                             x_stride, y_stride, x_off, y_off, len, zero_len);
    note_launch(name);
    RT_CHECK_LAUNCH();
    return RT_OK;
}

// Transform{1,0,2,3}: y[n, b, a, :] = x[n, a, b, :].  Rows of `inner` elements are copied whole, one (a,b) pair per
// block-row; inner is a full H*W plane (82 593 floats for NVSmall) so every access is a long contiguous run.
template <typename T>
__global__ void transpose01_kernel(const T* __restrict__ x, T* __restrict__ y, int d0, int d1, int64_t inner) {
    const int pair = blockIdx.y;                 // a * d1 + b
    const int a = pair / d1, b = pair % d1;
    const int64_t nofs = static_cast<int64_t>(blockIdx.z) * d0 * d1 * inner;
    const T* src = x + nofs + static_cast<int64_t>(pair) * inner;
    T* dst = y + nofs + (static_cast<int64_t>(b) * d0 + a) * inner;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    constexpr int V = 16 / sizeof(T);
    const uintptr_t sa = reinterpret_cast<uintptr_t>(src), da = reinterpret_cast<uintptr_t>(dst);
    if (((sa ^ da) & 15) == 0) {
        int64_t head = ((16 - (da & 15)) & 15) / sizeof(T);
        if (head > inner) head = inner;
        const int64_t nvec = (inner - head) / V;
        for (int64_t i = tid; i < head; i += stride) dst[i] = src[i];
        const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
        uint4* d4 = reinterpret_cast<uint4*>(dst + head);
        for (int64_t i = tid; i < nvec; i += stride) d4[i] = s4[i];
        for (int64_t i = head + nvec * V + tid; i < inner; i += stride) dst[i] = src[i];
    } else {
        for (int64_t i = tid; i < inner; i += stride) dst[i] = src[i];
    }
}

}  // namespace
}  // namespace rt

using namespace rt;

extern "C" {

const char* rt_version(void) { return "redtail_b200 0.1 (sm_100a)"; }
uint64_t rt_launch_count(void) { return g_launches.load(); }
const char* rt_last_kernel(void) { return g_last_kernel; }
void rt_add_launch_count(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int rt_elu(int dtype, const void* x, void* y, int64_t count, void* stream) {
    return dispatch_unary<kElu>(dtype, x, y, count, UnaryArgs{0, 1, 1}, stream, "elu");
}
int rt_sigmoid(int dtype, const void* x, void* y, int64_t count, void* stream) {
    return dispatch_unary<kSigmoid>(dtype, x, y, count, UnaryArgs{0, 1, 1}, stream, "sigmoid");
}
int rt_scale(int dtype, const void* x, void* y, int64_t count, float shift, float scale, float power, void* stream) {
    return dispatch_unary<kScale>(dtype, x, y, count, UnaryArgs{shift, scale, power}, stream, "scale");
}

int rt_eltwise_sum(int dtype, const void* a, const void* b, void* y, int64_t count, void* stream) {
    if (!a || !b || !y || count < 0) return RT_ERR_ARG;
    if (count == 0) return RT_OK;
    if (dtype == RT_F32)
        sum_kernel<float><<<grid_for(count), kThreads, 0, as_stream(stream)>>>(
            static_cast<const float*>(a), static_cast<const float*>(b), static_cast<float*>(y), count);
    else if (dtype == RT_F16)
        sum_kernel<__half><<<grid_for(count), kThreads, 0, as_stream(stream)>>>(
            static_cast<const __half*>(a), static_cast<const __half*>(b), static_cast<__half*>(y), count);
    else
        return RT_ERR_UNSUPPORTED;
    note_launch("eltwise_sum");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

int rt_convert(int src_dtype, const void* x, int dst_dtype, void* y, int64_t count, void* stream) {
    if (!x || !y || count < 0) return RT_ERR_ARG;
    if (count == 0) return RT_OK;
    cudaStream_t s = as_stream(stream);
    const int grid = grid_for(count);
    if (src_dtype == RT_F32 && dst_dtype == RT_F16)
        convert_kernel<float, __half><<<grid, kThreads, 0, s>>>(static_cast<const float*>(x), static_cast<__half*>(y), count);
    else if (src_dtype == RT_F16 && dst_dtype == RT_F32)
        convert_kernel<__half, float><<<grid, kThreads, 0, s>>>(static_cast<const __half*>(x), static_cast<float*>(y), count);
    else if (src_dtype == dst_dtype && (src_dtype == RT_F32 || src_dtype == RT_F16))
        return static_cast<int>(cudaMemcpyAsync(y, x, count * (src_dtype == RT_F32 ? 4 : 2), cudaMemcpyDeviceToDevice, s));
    else
        return RT_ERR_UNSUPPORTED;
    note_launch("convert");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

int rt_pad_planes(int dtype, const void* x, void* y, int n, int planes, int64_t plane_elems, int pad_end, void* stream) {
    if (!x || !y || n < 0 || planes < 0 || plane_elems < 0 || pad_end < 0) return RT_ERR_ARG;
    const int64_t len = planes * plane_elems, zl = pad_end * plane_elems;
    if (dtype == RT_F32) return launch_block_copy<float>(x, y, n, len, len + zl, 0, 0, len, zl, as_stream(stream), "pad_planes");
    if (dtype == RT_F16) return launch_block_copy<__half>(x, y, n, len, len + zl, 0, 0, len, zl, as_stream(stream), "pad_planes");
    return RT_ERR_UNSUPPORTED;
}

int rt_slice_planes(int dtype, const void* x, void* y, int n, int planes, int64_t plane_elems, int start, int end, void* stream) {
    if (!x || !y || n < 0 || start < 0 || end < start || end > planes || plane_elems < 0) return RT_ERR_ARG;
    const int64_t xs = planes * plane_elems, len = (end - start) * plane_elems;
    if (dtype == RT_F32) return launch_block_copy<float>(x, y, n, xs, len, start * plane_elems, 0, len, 0, as_stream(stream), "slice_planes");
    if (dtype == RT_F16) return launch_block_copy<__half>(x, y, n, xs, len, start * plane_elems, 0, len, 0, as_stream(stream), "slice_planes");
    return RT_ERR_UNSUPPORTED;
}

int rt_concat_channels(int dtype, const void* a, int ca, const void* b, int cb, void* y, int n, int64_t inner, void* stream) {
    if (!a || !b || !y || n < 0 || ca < 0 || cb < 0 || inner < 0) return RT_ERR_ARG;
    const int64_t la = ca * inner, lb = cb * inner;
    int rc;
    if (dtype == RT_F32) {
        rc = launch_block_copy<float>(a, y, n, la, la + lb, 0, 0, la, 0, as_stream(stream), "concat");
        if (rc) return rc;
        return launch_block_copy<float>(b, y, n, lb, la + lb, 0, la, lb, 0, as_stream(stream), "concat");
    }
    if (dtype == RT_F16) {
        rc = launch_block_copy<__half>(a, y, n, la, la + lb, 0, 0, la, 0, as_stream(stream), "concat");
        if (rc) return rc;
        return launch_block_copy<__half>(b, y, n, lb, la + lb, 0, la, lb, 0, as_stream(stream), "concat");
    }
    return RT_ERR_UNSUPPORTED;
}

int rt_transpose01(int dtype, const void* x, void* y, int n, int d0, int d1, int64_t inner, void* stream) {
    if (!x || !y || n < 0 || d0 < 0 || d1 < 0 || inner < 0) return RT_ERR_ARG;
    if (n == 0 || d0 == 0 || d1 == 0 || inner == 0) return RT_OK;
    if (static_cast<int64_t>(d0) * d1 > 65535 || n > 65535) return RT_ERR_UNSUPPORTED;
    const int esz = dtype == RT_F32 ? 4 : 2;
    int gx = static_cast<int>(ceil_div(inner, static_cast<int64_t>(kThreads) * (16 / esz) * 4));
    if (gx < 1) gx = 1;
    dim3 grid(gx, d0 * d1, n);
    if (dtype == RT_F32)
        transpose01_kernel<float><<<grid, kThreads, 0, as_stream(stream)>>>(static_cast<const float*>(x), static_cast<float*>(y), d0, d1, inner);
    else if (dtype == RT_F16)
        transpose01_kernel<__half><<<grid, kThreads, 0, as_stream(stream)>>>(static_cast<const __half*>(x), static_cast<__half*>(y), d0, d1, inner);
    else
        return RT_ERR_UNSUPPORTED;
    note_launch("transpose01");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

}  // extern "C"

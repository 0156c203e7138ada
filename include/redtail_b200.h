/*
 * redtail_b200.h -- the C-ABI of libredtail_b200.so: hand-written sm_100a kernels behind the
 * redtail stereoDNN plugin path.  Plain pointers and sizes only; every pointer named x/y/left/
 * right/out/workspace is DEVICE memory, every `stream` is a cudaStream_t passed as void*.
 * All entry points are asynchronous on `stream` and return 0 on success, otherwise the
 * cudaError_t value (or a negative rt_status for argument errors) -- the convention of the
 * reference's `enqueue()` (stereoDNN/lib/ *_plugin.cpp: "return 0 / -1 / cudaError_t").
 *
 * Each group cites the reference interface it replaces (paths relative to stereoDNN/):
 *
 *   rt_cost_volume            CudaKernels::computeCostVolume      lib/internal_utils.h:78-97, lib/kernels.cu:136-161
 *   rt_corr_cost_volume       CudaKernels::computeCorrCostVolume  lib/kernels.cu:252-287
 *   rt_elu                    EluPlugin::enqueue                  lib/elu_plugin.cpp:123-135 (cudnnActivationForward)
 *   rt_softargmax             SoftargmaxPlugin::enqueue           lib/softargmax_plugin.cpp:167-205 (5 cuDNN passes)
 *   rt_pad_planes             PaddingPlugin::enqueue              lib/padding_plugin.cpp:79-94
 *   rt_slice_planes           SlicePlugin::enqueue                lib/slice_plugin.cpp:80-92
 *   rt_transpose01            TransformPlugin::enqueue            lib/transform_plugin.cpp:94-108 (cudnnTransformTensor)
 *   rt_conv3d_*               Conv3DPlugin                        lib/conv3d_plugin.cpp:102-136,187-216 (cudnnConvolutionForward+AddTensor)
 *                             Conv3DTransposePlugin               lib/conv3d_transpose_plugin.cpp:116-151,205-243 (cudnnConvolutionBackwardData
 *                                                                 + CudaKernels::addDBiasTo3DConv lib/kernels.cu:310-334)
 *   rt_conv2d_* rt_scale rt_eltwise_sum rt_concat_channels rt_sigmoid
 *                             the TensorRT-native layers the generated builders call
 *                             (sample_app/nvsmall_1025x321_net.cpp:36-53,350; resnet18_2D_513x257_net.cpp:613,722,766)
 *   rt_convert                CudaKernels::fp32Tofp16 / fp16Tofp32 lib/kernels.cu:340-375
 *   rt_preprocess_bgr8 rt_disparity_to_u16 rt_write_png16
 *                             the OpenCV image handling of the app around the engine (sample_app/main.cpp:83-98,317-330)
 *
 * There is no CPU fallback anywhere behind this header.
 */
#ifndef REDTAIL_B200_H
#define REDTAIL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Data types; values follow nvinfer1::DataType (kFLOAT = 0, kHALF = 1). */
enum { RT_F32 = 0, RT_F16 = 1 };

/* Argument-error codes (CUDA errors are returned as their positive cudaError_t value). */
enum { RT_OK = 0, RT_ERR_ARG = -1, RT_ERR_UNSUPPORTED = -2, RT_ERR_NO_DEVICE = -3 };

/* Tensor-core numerics of the 3-D convolution path (see DESIGN.md "Numerics"):
 *   RT_PREC_FP32   fp32-accurate: every operand split into two fp16 terms (hi + 2^-11 lo), three tcgen05
 *                  kind::f16 products accumulated in fp32 TMEM -- holds the 1e-3 disparity tolerance.
 *   RT_PREC_FP16   single fp16 product, fp32 accumulate -- the reference's fp16 configs (1e-2 tolerance).
 *   RT_PREC_SIMT   plain fp32 FMA on CUDA cores (validation path, no tensor cores).                        */
enum { RT_PREC_FP32 = 0, RT_PREC_FP16 = 1, RT_PREC_SIMT = 2 };

/* Activation layouts of the 3-D convolution path.
 *   RT_LAYOUT_DENSE    fp32, the reference's plugin layouts ([D,C,H,W] / [K,D,H,W], batch leading).
 *   RT_LAYOUT_SPLIT16  engine-internal: per sample two channels-last fp16 planes [D][H][W][C], `hi` then `lo`, with
 *                      x = hi + lo / 2048 (what the tcgen05 kernel consumes; same bytes per element as fp32).  A tensor in
 *                      this layout never needs Transform / Padding / fp32<->fp16 passes between two convolutions.   */
enum { RT_LAYOUT_DENSE = 0, RT_LAYOUT_SPLIT16 = 1 };

const char* rt_version(void);
/* Number of kernels this library has launched since load (bench.py reports it as gpu_launches). */
uint64_t rt_launch_count(void);
/* Adds n to that counter: the engine replays captured CUDA graphs, whose kernel launches do not pass through the library. */
void rt_add_launch_count(uint64_t n);
/* Name of the last kernel variant launched per op family, for tests that assert the tensor-core path ran. */
const char* rt_last_kernel(void);

/* ---- cost volume ------------------------------------------------------------------------------------ */
/* left,right [n,c,h,w] -> out [n,max_disp,2c,h,w]:  out[d,ch]=left[ch]; out[d,c+ch,y,x]=right[ch,y,x-d] | 0. */
int rt_cost_volume(int dtype, const void* left, const void* right, void* out,
                   int n, int c, int h, int w, int max_disp, void* stream);
/* left,right [n,c,h,w] -> out [n,max_disp,h,w]:  out[d,y,x] = sum_ch left[ch,y,x]*right[ch,y,x-d] | 0 (fp32 accumulate). */
int rt_corr_cost_volume(int dtype, const void* left, const void* right, void* out,
                        int n, int c, int h, int w, int max_disp, void* stream);

/* left,right dense fp32 [n,c,h,w] -> out RT_LAYOUT_SPLIT16 [n][hi|lo][max_disp][h][w][2c] (same values as rt_cost_volume). */
int rt_cost_volume_split16(const void* left, const void* right, void* out,
                           int n, int c, int h, int w, int max_disp, void* stream);
/* Layout converters: x dense fp32 [n,d,c,h,w] <-> RT_LAYOUT_SPLIT16 [n][hi|lo][d][h][w][c]  (c % 8 == 0). */
int rt_dense_to_split16(const void* x, void* y, int n, int d, int c, int h, int w, void* stream);
int rt_split16_to_dense(const void* x, void* y, int n, int d, int c, int h, int w, void* stream);

/* ---- element-wise / data movement ------------------------------------------------------------------- */
int rt_elu(int dtype, const void* x, void* y, int64_t count, void* stream);            /* x>0 ? x : expm1(x) */
int rt_sigmoid(int dtype, const void* x, void* y, int64_t count, void* stream);
int rt_scale(int dtype, const void* x, void* y, int64_t count, float shift, float scale, float power, void* stream);
int rt_eltwise_sum(int dtype, const void* a, const void* b, void* y, int64_t count, void* stream);
int rt_convert(int src_dtype, const void* x, int dst_dtype, void* y, int64_t count, void* stream);
/* x [n, planes, plane_elems] -> y [n, planes+pad_end, plane_elems], appended planes zero. */
int rt_pad_planes(int dtype, const void* x, void* y, int n, int planes, int64_t plane_elems, int pad_end, void* stream);
/* x [n, planes, plane_elems] -> y [n, end-start, plane_elems]. */
int rt_slice_planes(int dtype, const void* x, void* y, int n, int planes, int64_t plane_elems, int start, int end, void* stream);
/* x [n, d0, d1, inner] -> y [n, d1, d0, inner]   (Transform{1,0,2,3}). */
int rt_transpose01(int dtype, const void* x, void* y, int n, int d0, int d1, int64_t inner, void* stream);
/* Channel concatenation of two [n,c_i,inner] tensors. */
int rt_concat_channels(int dtype, const void* a, int ca, const void* b, int cb, void* y, int n, int64_t inner, void* stream);

/* ---- soft-argmax ------------------------------------------------------------------------------------ */
/* x [n,d,h*w] -> y [n,h*w]:  sum_d d*softmax_d(is_min ? -x : x), max-subtracted, fp32 math. */
int rt_softargmax(int dtype, int is_min, const void* x, void* y, int n, int d, int64_t hw, void* stream);

/* ---- 3-D convolution / transposed convolution plans -------------------------------------------------- */
typedef struct rt_conv3d_plan rt_conv3d_plan;

typedef struct {
    int transposed;          /* 0: Conv3DPlugin, 1: Conv3DTransposePlugin                                    */
    int k, v, c, r, s;       /* weight dims, KVCRS order (conv: K outputs/C inputs; transposed: K inputs/C outputs) */
    int stride[3];           /* D,H,W                                                                        */
    int pad[3];              /* symmetric pad (= the plugin's pad_start), D,H,W                               */
    int in_dims[4];          /* conv: [D,C,H,W]   transposed: [K,D,H,W]                                       */
    int out_dims[4];         /* conv: [K,Do,Ho,Wo] transposed: [Dx,C,Hx,Wx] (caller-supplied, as the plugin's) */
    int weights_dtype;       /* RT_F32 | RT_F16 (host arrays)                                                */
    const void* weights;     /* host, k*v*c*r*s elements                                                     */
    const void* bias;        /* host, K (conv) or C (transposed) elements, or NULL                           */
    int precision;           /* RT_PREC_*                                                                    */
    int fuse_elu;            /* apply ELU in the epilogue                                                    */
    int out_transposed;      /* conv only: write [Do,K,Ho,Wo] instead of [K,Do,Ho,Wo] (fuses Transform{1,0,2,3}) */
    int slice_d;             /* transposed only: drop this many trailing D planes of out_dims (fuses SlicePlugin) */
    int in_layout;           /* RT_LAYOUT_DENSE (fp32, the plugin layouts above) | RT_LAYOUT_SPLIT16                */
    int out_layout;          /* layout of y; `skip`, when given, uses the same layout as y                          */
    int pad_end_d;           /* conv only: this many virtual zero planes follow the input (fuses PaddingPlugin);
                                in_dims[0] excludes them -- on the tensor-core path they are TMA out-of-bounds fill */
    int fuse_softargmax;     /* transposed, C == 1 only: 0 none, 1 soft-argmin, 2 soft-argmax over the (sliced) output planes
                                (fuses SoftargmaxPlugin, lib/softargmax_plugin.cpp:167-205): y is then [n,Hx,Wx] fp32 and the
                                [Dx,1,Hx,Wx] volume is never written.  Needs RT_PREC_FP32, a split16 input, K == 32.       */
    const float* act_params; /* conv only, tensor-core precisions: host array [4][K] = s1, b1, s2, b2 or NULL.  Applied after bias and
                                skip:  y = max(v * s1[k] + b1[k], 0) * s2[k] + b2[k]  -- the S-ReLU chain Scale -> ReLU -> Scale of the
                                TrailNet model (TrailNet_SResNet-18.prototxt:54-105) fused into the convolution.               */
} rt_conv3d_desc;

int  rt_conv3d_create(const rt_conv3d_desc* desc, rt_conv3d_plan** plan);   /* repacks + uploads weights        */
int  rt_conv3d_tc_supported(const rt_conv3d_desc* desc);   /* 1 when the tcgen05 kernels cover this shape (no allocation) */
void rt_conv3d_destroy(rt_conv3d_plan* plan);
size_t rt_conv3d_workspace_size(const rt_conv3d_plan* plan, int max_batch);
/* x, y dense fp32 in the layouts of rt_conv3d_desc; `skip` (may be NULL) is added before ELU, same layout as y. */
int  rt_conv3d_enqueue(const rt_conv3d_plan* plan, int n, const void* x, const void* skip, void* y,
                       void* workspace, void* stream);

/* ---- fused CostVolume(kDefault) -> Conv3D 3x3x3 / stride 1 / pad 1 [-> Transform{1,0,2,3}] [-> ELU] ------------------
 * Replaces the pair cost_volume_plugin.cpp:122-139 + conv3d_plugin.cpp:186-279 when the engine sees them back to back
 * (SURVEY.md section 8, row N3).  The cost volume cv[d, 0:c] = L, cv[d, c:2c, y, x] = R[y, x-d] is a broadcast / shift of
 * two 2-D maps, so the 3-D convolution over it separates exactly:
 *     out[k,d,y,x] = b[k] + sum_v A_v[k,y,x] + sum_v C_v[k,y,x-(d+v-1)]          (v = filter plane, d+v-1 inside [0,D))
 * with A_v = conv2d(L, W[:,v,0:c]), C_v = conv2d(R, W[:,v,c:2c]) (3x3, pad 1), minus the dw=+1 filter column at x = w-1
 * (the volume ends there, the shifted image does not).  Neither the 1 GB volume nor the 438 GFLOP 3-D convolution are
 * executed; results agree with the unfused pair to fp32 rounding.                                                        */
typedef struct rt_cvconv_plan rt_cvconv_plan;
typedef struct {
    int c, h, w;             /* left/right feature maps, dense fp32 [n,c,h,w]                                  */
    int max_disp;            /* D planes of the (never materialised) cost volume [D,2c,h,w]                    */
    int k;                   /* conv outputs; weights KVCRS [k,3,2c,3,3]                                       */
    int weights_dtype;       /* RT_F32 | RT_F16 (host arrays)                                                  */
    const void* weights;
    const void* bias;        /* host, k elements, or NULL                                                      */
    int precision;           /* RT_PREC_* of the two inner 2-D convolutions                                    */
    int fuse_elu;
    int out_transposed;      /* dense output only: 1 -> [D,K,H,W] (Transform fused), 0 -> [K,D,H,W]            */
    int out_layout;          /* RT_LAYOUT_DENSE | RT_LAYOUT_SPLIT16 ([n][hi|lo][D][h][w][k], k % 8 == 0)        */
} rt_costvol_conv3d_desc;
int  rt_costvol_conv3d_supported(const rt_costvol_conv3d_desc* desc);          /* 1 / 0, no allocation */
int  rt_costvol_conv3d_create(const rt_costvol_conv3d_desc* desc, rt_cvconv_plan** plan);
void rt_costvol_conv3d_destroy(rt_cvconv_plan* plan);
size_t rt_costvol_conv3d_workspace_size(const rt_cvconv_plan* plan, int max_batch);
int  rt_costvol_conv3d_enqueue(const rt_cvconv_plan* plan, int n, const void* left, const void* right, void* y,
                               void* workspace, void* stream);

/* ---- 2-D convolution / deconvolution (TensorRT-native layers of the builders) ------------------------ */
typedef struct rt_conv2d_plan rt_conv2d_plan;
typedef struct {
    int transposed;          /* 0: IConvolutionLayer (weights KCRS), 1: IDeconvolutionLayer (weights [Cin,Cout,R,S]) */
    int cin, cout, r, s;
    int stride[2], pad[2];
    int in_h, in_w;
    int weights_dtype; const void* weights; const void* bias;   /* host */
    int fuse_elu;
} rt_conv2d_desc;
int  rt_conv2d_create(const rt_conv2d_desc* desc, rt_conv2d_plan** plan);
void rt_conv2d_destroy(rt_conv2d_plan* plan);
void rt_conv2d_out_dims(const rt_conv2d_plan* plan, int* out_h, int* out_w);
int  rt_conv2d_enqueue(const rt_conv2d_plan* plan, int n, const void* x, void* y, void* stream);

/* ---- non-convolution layers of the TrailNet classifier (models/pretrained/TrailNet_SResNet-18.prototxt; ------------------
 *      ros/packages/caffe_ros/src/tensor_net.cpp runs it through TensorRT's Caffe parser).  Dense fp32 NCHW tensors. ------ */
/* y = x * scale[c] + shift[c]  (Caffe Scale layer with bias_term; a NULL array means 1 / 0).  x [n,c,hw]. */
int rt_scale_channel(const void* x, void* y, int n, int c, int64_t hw, const float* scale, const float* shift, void* stream);
/* S-ReLU chain of the model (prototxt:54-105): y = max(x * s1[c] + b1[c], 0) * s2[c] + b2[c] in one pass. */
int rt_srelu(const void* x, void* y, int n, int c, int64_t hw, const float* s1, const float* b1, const float* s2, const float* b2,
             void* stream);
int rt_relu(const void* x, void* y, int64_t count, void* stream);
/* Caffe pooling: out_h/out_w as the caller computed them (Caffe: ceil((in + 2 pad - k) / stride) + 1, last window starting
 * inside the image), windows clipped at the border; AVE divides by the window area clipped to the PADDED image. */
int rt_pool2d(const void* x, void* y, int n, int c, int h, int w, int out_h, int out_w, int k, int stride, int pad, int is_max,
              void* stream);
/* InnerProduct: y[n,m] = b[m] + sum_k x[n,k] W[m,k]  (W, b device fp32; b may be NULL). */
int rt_fully_connected(const void* x, const float* w, const float* b, void* y, int n, int k, int m, void* stream);
/* im2col for 2-D filters larger than 3x3 (TrailNet conv1, 7x7 stride 2): x [n,c,h,w] fp32 -> RT_LAYOUT_SPLIT16 [n][hi|lo][1][out_h][out_w][kp],
 * K index = (ci*r + ri)*s + si (KCRS weight order), zero for K >= c*r*s; the convolution is then 1x1 over kp channels on tcgen05. */
int rt_im2col_split16(const void* x, void* y, int n, int c, int h, int w, int r, int s, int stride, int pad, int out_h, int out_w,
                      int kp, void* stream);
/* Softmax over the channel dimension of [n,c,inner], max-subtracted. */
int rt_softmax_channels(const void* x, void* y, int n, int c, int64_t inner, void* stream);

/* ---- image side of the apps (sample_app/main.cpp:83-98 readImgFile, :317-330 PNG output) -------------- */
/* src: n 8-bit BGR images [src_h, src_w, 3] (row pitch src_pitch bytes, images src_pitch*src_h apart), device memory
 * -> dst [n,3,dst_h,dst_w] fp32 RGB in [0,1]: float conversion, cv::resize INTER_AREA (down-scaling or identity only),
 * BGR->RGB, HWC->CHW and the 1/255 scale in one kernel. */
int rt_preprocess_bgr8(const void* src, int n, int src_h, int src_w, int64_t src_pitch, void* dst, int dst_h, int dst_w,
                       void* stream);
/* out[i] = saturate_u16(round(disp[i] * scale)): the 16-bit PNG payload (scale 256; 256*width for ResNet18_2D). */
int rt_disparity_to_u16(const void* disp, void* out, int64_t count, float scale, void* stream);
/* HOST: pixels [height,width] uint16 -> 16-bit greyscale PNG file (what cv::imwrite produces for a CV_16U Mat). */
int rt_write_png16(const char* path, const uint16_t* pixels, int height, int width);

#ifdef __cplusplus
}
#endif
#endif /* REDTAIL_B200_H */

/*
 * redtail_b200_engine.h -- C-ABI over the whole stereo network (libnvstereo_inference.so).
 *
 * What a host application of the reference does by hand in C++ -- read the weight file, call the generated
 * create<Net>Network() builder, buildCudaEngine(), cudaMemcpy in, IExecutionContext::execute(), cudaMemcpy out
 * (stereoDNN/sample_app/main.cpp:176-315; ros/packages/stereo_dnn_ros/src/stereo_dnn_ros_node.cpp:60-103,297-335) --
 * exposed as plain C entry points so that Python (ctypes), bench.py and non-C++ hosts can drive it.  The engine behind
 * it is the same nvinfer1-compatible engine the unchanged reference builders use; the network wiring comes from
 * redtail_b200/csrc/host/nets.cpp, which goes through the same IPluginContainer / add* API.
 */
#ifndef REDTAIL_B200_ENGINE_H
#define REDTAIL_B200_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt_stereo_engine rt_stereo_engine;

/* model: "nvsmall" (towers -> concat cost volume D -> 8 conv3d -> 3 transposed conv3d -> soft-argmin; covers the
 *        reference's NVSmall 1025x321 [max_disp 48] and NVTiny 513x161 [max_disp 24] weight sets).
 * weights_path: the reference's weight-file format (cstring name, u32 count, count x f32|f16).
 * weights_dtype: RT_F32 | RT_F16.  Returns 0 or a negative rt_status / positive cudaError_t. */
int rt_stereo_create(const char* model, int height, int width, int max_disp, const char* weights_path,
                     int weights_dtype, int max_batch, rt_stereo_engine** engine);
void rt_stereo_destroy(rt_stereo_engine* engine);

/* Device buffers: left,right [batch,3,H,W] fp32, disp [batch,H,W] fp32.  Asynchronous on `stream` (cudaStream_t). */
int rt_stereo_enqueue(rt_stereo_engine* engine, int batch, const float* left, const float* right, float* disp, void* stream);
/* Host buffers (pinned for full speed): H2D copies, inference and the D2H copy, synchronous. */
int rt_stereo_execute_host(rt_stereo_engine* engine, int batch, const float* left, const float* right, float* disp);
/* The app's whole loop body (sample_app/main.cpp:83-98,287-330) with the image handling on the GPU: host 8-bit BGR images
 * [batch, src_h, src_w, 3] as cv::imread returns them (src >= network size) -> H2D of the 8-bit data, rt_preprocess_bgr8
 * (float, INTER_AREA resize, BGR->RGB, CHW, /255), inference, D2H of the disparity (disp, may be NULL) and/or of its
 * KITTI-style 16-bit payload round(disp * u16_scale) (disp_u16, may be NULL; rt_write_png16 writes it).  Synchronous. */
int rt_stereo_execute_images(rt_stereo_engine* engine, int batch, const uint8_t* left_bgr, const uint8_t* right_bgr,
                             int src_h, int src_w, float* disp, uint16_t* disp_u16, float u16_scale);
/* Runs once with per-layer CUDA-event timing (IProfiler) and writes "layer name\tms\n" lines into buf. */
int rt_stereo_profile(rt_stereo_engine* engine, int batch, const float* left, const float* right, float* disp,
                      char* buf, size_t buf_len);
/* Engine plan (ICudaEngine::serialize / IRuntime::deserializeCudaEngine with StereoDnnPluginFactory, the flow of
 * sample_app/main.cpp:207-220,270-275 -- which the reference could only use for ResNet18_2D because its Conv3D plugins
 * do not serialise).  rt_stereo_serialize copies the plan into buf when buf_len suffices and always returns the plan
 * size (0 on error); rt_stereo_deserialize rebuilds an engine from it (no weight file needed). */
size_t rt_stereo_serialize(const rt_stereo_engine* engine, void* buf, size_t buf_len);
int rt_stereo_deserialize(const void* plan, size_t plan_size, rt_stereo_engine** engine);
/* Same, with the plan's maximum batch size replaced by max_batch (> 0): a plan is a network description, so the batch it
 * was dumped with is not binding.  This is how the reference's other networks (ResNet-18, ResNet18_2D: plans written by
 * the reference's own generated builders through ICudaEngine::serialize or tools/dropin's host-only dump) are run at the
 * batch sizes of BASELINE.json's configs. */
int rt_stereo_deserialize_batch(const void* plan, size_t plan_size, int max_batch, rt_stereo_engine** engine);
/* Introspection. */
int rt_stereo_num_layers(const rt_stereo_engine* engine);
size_t rt_stereo_device_bytes(const rt_stereo_engine* engine);
const char* rt_stereo_last_error(void);

/* ---- single-input classifier networks from Caffe models (TrailNet S-ResNet-18) ------------------------------------------
 * What ros/packages/caffe_ros/src/tensor_net.cpp does in C++ (loadNetwork :182-260: Caffe parser -> buildCudaEngine ->
 * serialize -> deserializeCudaEngine -> context; forward :262-291: cudaMemcpy in, execute) as plain C entry points.  The
 * model goes through the nvcaffeparser1-compatible parser of include/NvCaffeParser.h. */
typedef struct rt_net_engine rt_net_engine;
/* output_blob is marked as the network output (tensor_net.cpp:144-151).  fp16 = 1: convolutions in the fp16 configuration. */
int rt_caffe_create(const char* prototxt_path, const char* caffemodel_path, const char* input_blob, const char* output_blob,
                    int max_batch, rt_net_engine** engine);
void rt_net_destroy(rt_net_engine* engine);
/* CHW of the input and output bindings (tensor_net.cpp:231-247). */
int rt_net_dims(const rt_net_engine* engine, int in_chw[3], int out_chw[3]);
/* Device buffers in [batch,C,H,W], out [batch,Co,Ho,Wo] fp32; asynchronous on `stream`. */
int rt_net_enqueue(rt_net_engine* engine, int batch, const float* in, float* out, void* stream);
/* Host buffers: H2D copy, inference, D2H copy, synchronous (TensorNet::forward without the OpenCV preprocessing). */
int rt_net_execute_host(rt_net_engine* engine, int batch, const float* in, float* out);
/* One profiled execution: "name\tms\n" per executed step. */
int rt_net_profile(rt_net_engine* engine, int batch, const float* in, float* out, char* buf, size_t buf_len);
/* Engine plan (tensor_net.cpp:168-172, 196-217: the model cache file). */
size_t rt_net_serialize(const rt_net_engine* engine, void* buf, size_t buf_len);
int rt_net_deserialize(const void* plan, size_t plan_size, int max_batch, rt_net_engine** engine);
int rt_net_num_layers(const rt_net_engine* engine);
/* HOST ONLY (no CUDA device needed): the plan rt_net_serialize would write for this model, without building an engine. */
size_t rt_caffe_dump_plan(const char* prototxt_path, const char* caffemodel_path, const char* output_blob, int max_batch, void* buf, size_t buf_len);

#ifdef __cplusplus
}
#endif
#endif

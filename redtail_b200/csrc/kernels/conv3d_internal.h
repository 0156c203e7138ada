// Internal plan object shared by the SIMT validation path (conv3d_simt.cu) and the tcgen05 path (conv3d_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "redtail_b200.h"

struct rt_conv3d_plan {
    rt_conv3d_desc desc;
    // Output geometry actually written (transposed conv: out_dims[0] - slice_d planes).
    int out_planes;
    // ---- SIMT path: weights repacked to [ceil(Cout/16)][V][Cin][R][S][16] fp32, bias fp32 [Cout] (zeros if absent).
    float* w_simt = nullptr;
    float* bias = nullptr;
    int cout = 0, cin = 0;
    // ---- tcgen05 path (filled by rt::tc_plan_init when precision != RT_PREC_SIMT and the shape is supported).
    void* tc = nullptr;
    // ---- depth-stationary tcgen05 kernel (conv3d_ds.cu): 32 -> 32 channel, 3x3x3, stride 1, split16 in/out; preferred over `tc`
    //      when set (a launch with a skip tensor still takes the generic kernel).
    void* ds = nullptr;
    // ---- transposed conv (Cout = 1) + slice + soft-argmax/min as one kernel (deconv_softargmax.cu); set iff desc.fuse_softargmax.
    void* dsa = nullptr;
    // ---- more than 128 output channels (the 256 / 512-channel layers of the TrailNet classifier): the plan is a list of
    //      <= 128-channel parts that write their slice of the one output tensor (channel offset out_c_offset of out_c_total);
    //      the first part's re-layout of a dense input (pack pass) is shared by the others.
    std::vector<rt_conv3d_plan*> parts;
    int out_c_total = 0;        // channels of the output tensor this plan writes into (0: its own cout)
    int out_c_offset = 0;       // first channel this plan writes
    bool reuse_pack = false;    // the packed activations are already in the workspace (written by an earlier part)
    std::vector<float> act_host;   // [4][cout] s1, b1, s2, b2 of a fused S-ReLU (empty: none); uploaded by tc_plan_init
};

namespace rt {
// Host weight arrays (RT_F32 | RT_F16) -> fp32.
void host_to_f32(int dtype, const void* src, int64_t count, std::vector<float>& dst);
bool tc_shape_supported(const rt_conv3d_desc& d);
int simt_conv3d_enqueue(const rt_conv3d_plan* p, int n, const float* x, const float* skip, float* y, cudaStream_t s);
// Returns RT_OK and sets p->tc, or RT_ERR_UNSUPPORTED when the shape is outside what the tensor-core kernels cover
// (the caller then fails loudly -- there is no silent fallback from a requested tensor-core precision).
int tc_plan_init(rt_conv3d_plan* p, const std::vector<float>& w_kvcrs, const std::vector<float>& bias);
void tc_plan_destroy(rt_conv3d_plan* p);
size_t tc_workspace_size(const rt_conv3d_plan* p, int max_batch);
int tc_conv3d_enqueue(const rt_conv3d_plan* p, int n, const float* x, const float* skip, float* y, void* workspace,
                      cudaStream_t s);
// Depth-stationary kernel (conv3d_ds.cu).  ds_plan_init returns RT_ERR_UNSUPPORTED (and leaves p->ds null) outside its shape class.
bool ds_shape_supported(const rt_conv3d_desc& d);
int ds_plan_init(rt_conv3d_plan* p, const std::vector<float>& w_kvcrs);
void ds_plan_destroy(rt_conv3d_plan* p);
int ds_conv3d_enqueue(const rt_conv3d_plan* p, int n, const void* x, void* y, cudaStream_t s);
// Fused Conv3DTranspose(32 -> 1) + Slice + Softargmax (deconv_softargmax.cu); y = [n, Hx, Wx] fp32 disparities.
bool dsa_shape_supported(const rt_conv3d_desc& d);
int dsa_plan_init(rt_conv3d_plan* p, const std::vector<float>& w_kvcrs, const std::vector<float>& bias);
void dsa_plan_destroy(rt_conv3d_plan* p);
int dsa_enqueue(const rt_conv3d_plan* p, int n, const void* x, void* y, cudaStream_t s);
}  // namespace rt

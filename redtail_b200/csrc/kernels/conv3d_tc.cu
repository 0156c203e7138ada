// placeholder, replaced by the tcgen05 implementation
#include "common.cuh"
#include "conv3d_internal.h"
namespace rt {
int tc_plan_init(rt_conv3d_plan*, const std::vector<float>&, const std::vector<float>&) { return RT_ERR_UNSUPPORTED; }
void tc_plan_destroy(rt_conv3d_plan*) {}
size_t tc_workspace_size(const rt_conv3d_plan*, int) { return 0; }
int tc_conv3d_enqueue(const rt_conv3d_plan*, int, const float*, const float*, float*, void*, cudaStream_t) { return RT_ERR_UNSUPPORTED; }
}

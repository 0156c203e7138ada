// Introspection interface the engine uses to recognise redtail plugins (for layer fusion and plan serialisation).
// Every plugin class of plugins.cpp implements it next to nvinfer1::IPlugin / IPluginExt; a third-party IPlugin that
// does not is still executed, through its own enqueue().
#pragma once
#include <NvInfer.h>

#include <string>

#include "redtail_tensorrt_plugins.h"

namespace redtail { namespace tensorrt {

enum class OpKind { kElu, kCostVolume, kConv3D, kConv3DTranspose, kTransform, kPadding, kSlice, kSoftargmax };

struct OpInfo {
    OpKind kind;
    std::string name;
    DataType data_type = DataType::kFLOAT;
    // cost volume / softargmax
    CostVolumeType cv_type = CostVolumeType::kDefault;
    int max_disparity = 0;
    SoftargmaxType sm_type = SoftargmaxType::kMax;
    // conv3d / conv3d transpose
    Dims kernel_dims{};      // KVCRS
    Dims stride{}, pad_start{}, pad_end{};
    Dims out_dims{};         // transpose only
    Weights kernel{DataType::kFLOAT, nullptr, 0}, bias{DataType::kFLOAT, nullptr, 0};
    // transform / padding / slice
    Permutation perm{};
    int pad_end_planes = 0;
    int slice_start = 0, slice_end = 0;
};

class IRedtailOp {
public:
    virtual const OpInfo& opInfo() const = 0;
    // Fused execution hands the plugin's work to an engine-level kernel; the plugin then must not allocate.
    virtual ~IRedtailOp() {}
};

} }

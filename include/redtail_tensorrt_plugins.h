/*
 * redtail_tensorrt_plugins.h -- public operator surface of the stereo-DNN plugin library, B200 edition.
 *
 * Source-compatible with the reference header of the same name (stereoDNN/lib/redtail_tensorrt_plugins.h:18-146):
 * same namespace, enums (and values), IPluginContainer virtuals, add* helpers and StereoDnnPluginFactory, so the
 * generated network builders (sample_app/ *_net.cpp), sample_app/main.cpp and tests/tests_main.cpp build against
 * it unchanged.  What differs is everything behind it: each plugin's enqueue() calls one C-ABI entry point of
 * include/redtail_b200.h, i.e. one hand-written sm_100a kernel -- no TensorRT, no cuDNN.
 */
#ifndef REDTAIL_TENSORRT_PLUGINS_H
#define REDTAIL_TENSORRT_PLUGINS_H

#include <NvInfer.h>

#include <memory>
#include <string>

namespace redtail { namespace tensorrt {

using namespace nvinfer1;

/* Layout convention of the 3-D (transposed) convolution plugins: kTensorFlow = input [D,C,H,W], output [K,D,H,W]
 * (reference: lib/conv_utils.cpp:14-44); kCuDnn = plain NCDHW on both sides. */
enum class Conv3DType { kCuDnn = 0, kTensorFlow = 1 };

/* kDefault: two [C,H,W] maps -> [D,2C,H,W] (copy / shift / concat); kCorrelation: -> [D,H,W] (channel dot product). */
enum class CostVolumeType { kDefault = 0, kCorrelation = 1 };

enum class SoftargmaxType { kMax = 0, kMin = 1 };

/* Owns every plugin it creates for the lifetime of the container (the engine only borrows raw pointers,
 * reference: lib/internal_utils.h:114-168).  Creation is thread-safe. */
class IPluginContainer
{
public:
    virtual ~IPluginContainer() = default;

    virtual IPlugin* createEluPlugin(DataType data_type, std::string name) = 0;
    virtual IPlugin* deserializeEluPlugin(const char* name, const void* data, size_t size) = 0;

    virtual IPlugin* createCostVolumePlugin(DataType data_type, CostVolumeType cv_type, int max_disparity,
                                            std::string name) = 0;
    virtual IPlugin* deserializeCostVolumePlugin(const char* name, const void* data, size_t size) = 0;

    /* kernel_dims: KVCRS; stride/pad: D,H,W.  H/W padding must be symmetric, D may be (p, p) or (p, p+1). */
    virtual IPlugin* createConv3DPlugin(Conv3DType conv_type, Dims kernel_dims,
                                        Dims stride_dims, Dims pad_start_dims, Dims pad_end_dims,
                                        Weights kernel_weights, Weights bias_weights,
                                        std::string name) = 0;

    /* out_dims: [D,C,H,W] of the result (may exceed the TF size by one D plane, to be removed by a Slice). */
    virtual IPlugin* createConv3DTransposePlugin(Conv3DType conv_type, Dims kernel_dims, Dims out_dims,
                                                 Dims stride_dims, Dims pad_start_dims, Dims pad_end_dims,
                                                 Weights kernel_weights, Weights bias_weights,
                                                 std::string name) = 0;

    virtual IPlugin* createTransformPlugin(Permutation permutation, std::string name) = 0;

    virtual IPlugin* createPaddingPlugin(DimsNCHW pad_start, DimsNCHW pad_end,
                                         std::string name) = 0;

    virtual IPlugin* createSlicePlugin(Dims dims, Dims slice_start, Dims slice_end,
                                       std::string name) = 0;

    virtual IPlugin* createSoftargmaxPlugin(DataType data_type, SoftargmaxType sm_type, std::string name) = 0;
    virtual IPlugin* deserializeSoftargmaxPlugin(const char* name, const void* data, size_t size) = 0;

    static std::unique_ptr<IPluginContainer> create(ILogger& log);
};

/* Helpers used by the generated builders: create the plugin in the container and add it to the network. */
ILayer* addElu(IPluginContainer& plugin_factory, INetworkDefinition& network, ITensor& input,
               DataType data_type, const std::string& name);

ILayer* addCostVolume(IPluginContainer& plugin_factory, INetworkDefinition& network,
                      ITensor& left_input, ITensor& right_input,
                      CostVolumeType cv_type, int max_disparity,
                      DataType data_type, const std::string& name);

ILayer* addConv3D(IPluginContainer& plugin_factory, INetworkDefinition& network, ITensor& input,
                  Conv3DType conv_type, Dims kernel_dims, Dims stride_dims,
                  Dims pad_start_dims, Dims pad_end_dims,
                  Weights kernel_weights, Weights bias_weights,
                  const std::string& name);

ILayer* addConv3DTranspose(IPluginContainer& plugin_factory, INetworkDefinition& network, ITensor& input,
                           Conv3DType conv_type, Dims kernel_dims, Dims out_dims,
                           Dims stride_dims, Dims pad_start_dims, Dims pad_end_dims,
                           Weights kernel_weights, Weights bias_weights,
                           const std::string& name);

ILayer* addSlice(IPluginContainer& plugin_factory, INetworkDefinition& network, ITensor& input,
                 Dims dims, Dims slice_start, Dims slice_end,
                 const std::string& name);

ILayer* addTransform(IPluginContainer& plugin_factory, INetworkDefinition& network, ITensor& input,
                     Permutation permutation,
                     const std::string& name);

ILayer* addPad(IPluginContainer& plugin_factory, INetworkDefinition& network, ITensor& input,
               DimsNCHW pad_start, DimsNCHW pad_end,
               const std::string& name);

ILayer* addSoftargmax(IPluginContainer& plugin_factory, INetworkDefinition& network, ITensor& input,
                      SoftargmaxType sm_type, DataType data_type, const std::string& name);

/* Plugin factory for engine deserialisation: dispatches on the leading int32 type tag of the serialised blob
 * (reference: lib/internal_utils.cpp:289-313). */
class StereoDnnPluginFactory: public IPluginFactory
{
public:
    /* 0..2 are the reference's tags (lib/internal_utils.h); 3..7 extend the scheme to the plugins the reference could not
     * serialise (conv3d_plugin.cpp:224-229 asserts), so that NVSmall/NVTiny/ResNet-18 plans can be saved and reloaded too. */
    enum class PluginType { kElu = 0, kCostVolume = 1, kSoftargmax = 2,
                            kConv3D = 3, kConv3DTranspose = 4, kTransform = 5, kPadding = 6, kSlice = 7 };

    StereoDnnPluginFactory(IPluginContainer& container);

    IPlugin* createPlugin(const char* layerName, const void* serialData, size_t serialLength) override;

private:
    IPluginContainer& container_;
};

} }

#endif // REDTAIL_TENSORRT_PLUGINS_H

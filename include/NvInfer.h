/*
 * NvInfer.h -- nvinfer1-compatible interface of the redtail_b200 mini inference engine.
 *
 * The reference's plugin library, generated network builders, sample app and tests are written against the
 * TensorRT 3/4 C++ API (`IPlugin` / `IPluginExt`, `INetworkDefinition::add*`, `IBuilder::buildCudaEngine`,
 * `IExecutionContext::execute`; SURVEY.md 8b lists every symbol they touch).  That API no longer exists in current
 * TensorRT and TensorRT is not part of this product at all: this header re-declares the subset those callers use,
 * with the same names, argument meaning and call order, and libnvstereo_inference.so implements it with a small
 * graph-capture + static-memory-plan executor whose every layer runs a hand-written sm_100a kernel through the
 * C-ABI of include/redtail_b200.h.  The following reference sources compile against it unchanged:
 * stereoDNN/sample_app/{nvsmall_1025x321,nvtiny_513x161,resnet18_1025x321,resnet18_2D_513x257}_net.cpp, networks.h,
 * sample_app/main.cpp and tests/tests_main.cpp (built by tools/dropin/build.sh, run by tests/test_gpu_dropin.py).
 *
 * This is an independent implementation of a published interface; nothing here is derived from TensorRT sources.
 */
#ifndef REDTAIL_B200_NVINFER_H
#define REDTAIL_B200_NVINFER_H

#include <cuda_runtime_api.h>

#include <algorithm>   /* TensorRT's header pulls the standard algorithms in; sample_app/main.cpp:70 relies on it (std::find_if) */
#include <cstddef>
#include <cstdint>

#define NV_TENSORRT_MAJOR 4
#define NV_TENSORRT_MINOR 0
#define NV_TENSORRT_PATCH 0
#define NV_TENSORRT_VERSION ((NV_TENSORRT_MAJOR * 1000) + (NV_TENSORRT_MINOR * 100) + NV_TENSORRT_PATCH)
#define REDTAIL_B200_ENGINE 1

namespace nvinfer1 {

template <typename T>
inline int EnumMax();

enum class DataType : int { kFLOAT = 0, kHALF = 1, kINT8 = 2, kINT32 = 3 };
template <> inline int EnumMax<DataType>() { return 4; }

enum class DimensionType : int { kSPATIAL = 0, kCHANNEL = 1, kINDEX = 2, kSEQUENCE = 3 };
template <> inline int EnumMax<DimensionType>() { return 4; }

// Aggregate on purpose: callers write `Dims{3, {1, 1, 1}}` and `{5, {32, 3, 64, 3, 3}}`.
class Dims {
public:
    static const int MAX_DIMS = 8;
    int nbDims;
    int d[MAX_DIMS];
    DimensionType type[MAX_DIMS];
};

class DimsHW : public Dims {
public:
    DimsHW() : DimsHW(0, 0) {}
    DimsHW(int height, int width) {
        nbDims = 2;
        d[0] = height; d[1] = width;
        for (int i = 2; i < MAX_DIMS; ++i) d[i] = 0;
        for (int i = 0; i < MAX_DIMS; ++i) type[i] = DimensionType::kSPATIAL;
    }
    int& h() { return d[0]; }
    int h() const { return d[0]; }
    int& w() { return d[1]; }
    int w() const { return d[1]; }
};

class DimsCHW : public Dims {
public:
    DimsCHW() : DimsCHW(0, 0, 0) {}
    DimsCHW(int channels, int height, int width) {
        nbDims = 3;
        d[0] = channels; d[1] = height; d[2] = width;
        for (int i = 3; i < MAX_DIMS; ++i) d[i] = 0;
        for (int i = 0; i < MAX_DIMS; ++i) type[i] = DimensionType::kSPATIAL;
        type[0] = DimensionType::kCHANNEL;
    }
    int& c() { return d[0]; }
    int c() const { return d[0]; }
    int& h() { return d[1]; }
    int h() const { return d[1]; }
    int& w() { return d[2]; }
    int w() const { return d[2]; }
};

class DimsNCHW : public Dims {
public:
    DimsNCHW() : DimsNCHW(0, 0, 0, 0) {}
    DimsNCHW(int batchSize, int channels, int height, int width) {
        nbDims = 4;
        d[0] = batchSize; d[1] = channels; d[2] = height; d[3] = width;
        for (int i = 4; i < MAX_DIMS; ++i) d[i] = 0;
        for (int i = 0; i < MAX_DIMS; ++i) type[i] = DimensionType::kSPATIAL;
        type[0] = DimensionType::kINDEX;
        type[1] = DimensionType::kCHANNEL;
    }
    int& n() { return d[0]; }
    int n() const { return d[0]; }
    int& c() { return d[1]; }
    int c() const { return d[1]; }
    int& h() { return d[2]; }
    int h() const { return d[2]; }
    int& w() { return d[3]; }
    int w() const { return d[3]; }
};

class Weights {
public:
    DataType type;
    const void* values;
    int64_t count;
};

struct Permutation {
    int order[Dims::MAX_DIMS];
};

class IHostMemory {
public:
    virtual void* data() const = 0;
    virtual std::size_t size() const = 0;
    virtual DataType type() const = 0;
    virtual void destroy() = 0;
protected:
    virtual ~IHostMemory() {}
};

enum class LayerType : int {
    kCONVOLUTION = 0, kFULLY_CONNECTED = 1, kACTIVATION = 2, kPOOLING = 3, kLRN = 4, kSCALE = 5, kSOFTMAX = 6,
    kDECONVOLUTION = 7, kCONCATENATION = 8, kELEMENTWISE = 9, kPLUGIN = 10, kRNN = 11, kUNARY = 12, kPADDING = 13,
    kSHUFFLE = 14
};
template <> inline int EnumMax<LayerType>() { return 15; }

enum class ActivationType : int { kRELU = 0, kSIGMOID = 1, kTANH = 2 };
template <> inline int EnumMax<ActivationType>() { return 3; }
enum class ScaleMode : int { kUNIFORM = 0, kCHANNEL = 1, kELEMENTWISE = 2 };
template <> inline int EnumMax<ScaleMode>() { return 3; }
enum class ElementWiseOperation : int { kSUM = 0, kPROD = 1, kMAX = 2, kMIN = 3, kSUB = 4, kDIV = 5, kPOW = 6 };
template <> inline int EnumMax<ElementWiseOperation>() { return 7; }

class ILogger {
public:
    enum class Severity { kINTERNAL_ERROR = 0, kERROR = 1, kWARNING = 2, kINFO = 3 };
    virtual void log(Severity severity, const char* msg) = 0;
protected:
    virtual ~ILogger() {}
};
template <> inline int EnumMax<ILogger::Severity>() { return 4; }

class IProfiler {
public:
    virtual void reportLayerTime(const char* layerName, float ms) = 0;
protected:
    virtual ~IProfiler() {}
};

class ITensor {
public:
    virtual void setName(const char* name) = 0;
    virtual const char* getName() const = 0;
    virtual Dims getDimensions() const = 0;
    virtual DataType getType() const = 0;
    virtual bool isNetworkInput() const = 0;
    virtual bool isNetworkOutput() const = 0;
protected:
    virtual ~ITensor() {}
};

class ILayer {
public:
    virtual LayerType getType() const = 0;
    virtual void setName(const char* name) = 0;
    virtual const char* getName() const = 0;
    virtual int getNbInputs() const = 0;
    virtual ITensor* getInput(int index) const = 0;
    virtual int getNbOutputs() const = 0;
    virtual ITensor* getOutput(int index) const = 0;
protected:
    virtual ~ILayer() {}
};

class IConvolutionLayer : public ILayer {
public:
    virtual void setKernelSize(DimsHW kernelSize) = 0;
    virtual DimsHW getKernelSize() const = 0;
    virtual void setNbOutputMaps(int nbOutputMaps) = 0;
    virtual int getNbOutputMaps() const = 0;
    virtual void setStride(DimsHW stride) = 0;
    virtual DimsHW getStride() const = 0;
    virtual void setPadding(DimsHW padding) = 0;
    virtual DimsHW getPadding() const = 0;
    virtual void setKernelWeights(Weights weights) = 0;
    virtual Weights getKernelWeights() const = 0;
    virtual void setBiasWeights(Weights weights) = 0;
    virtual Weights getBiasWeights() const = 0;
protected:
    virtual ~IConvolutionLayer() {}
};

class IDeconvolutionLayer : public ILayer {
public:
    virtual void setKernelSize(DimsHW kernelSize) = 0;
    virtual DimsHW getKernelSize() const = 0;
    virtual void setNbOutputMaps(int nbOutputMaps) = 0;
    virtual int getNbOutputMaps() const = 0;
    virtual void setStride(DimsHW stride) = 0;
    virtual DimsHW getStride() const = 0;
    virtual void setPadding(DimsHW padding) = 0;
    virtual DimsHW getPadding() const = 0;
    virtual void setKernelWeights(Weights weights) = 0;
    virtual Weights getKernelWeights() const = 0;
    virtual void setBiasWeights(Weights weights) = 0;
    virtual Weights getBiasWeights() const = 0;
protected:
    virtual ~IDeconvolutionLayer() {}
};

class IScaleLayer : public ILayer {
public:
    virtual ScaleMode getMode() const = 0;
protected:
    virtual ~IScaleLayer() {}
};

class IElementWiseLayer : public ILayer {
public:
    virtual ElementWiseOperation getOperation() const = 0;
protected:
    virtual ~IElementWiseLayer() {}
};

class IConcatenationLayer : public ILayer {
protected:
    virtual ~IConcatenationLayer() {}
};

class IActivationLayer : public ILayer {
public:
    virtual ActivationType getActivationType() const = 0;
protected:
    virtual ~IActivationLayer() {}
};

enum class PoolingType : int { kMAX = 0, kAVERAGE = 1, kMAX_AVERAGE_BLEND = 2 };
template <> inline int EnumMax<PoolingType>() { return 3; }

// The TrailNet classifier (models/pretrained/TrailNet_SResNet-18.prototxt, run by ros/packages/caffe_ros/src/tensor_net.cpp
// through the Caffe parser) needs three more native layers: pooling, fully connected, soft-max.
class IPoolingLayer : public ILayer {
public:
    virtual void setPoolingType(PoolingType type) = 0;
    virtual PoolingType getPoolingType() const = 0;
    virtual void setWindowSize(DimsHW windowSize) = 0;
    virtual DimsHW getWindowSize() const = 0;
    virtual void setStride(DimsHW stride) = 0;
    virtual DimsHW getStride() const = 0;
    virtual void setPadding(DimsHW padding) = 0;
    virtual DimsHW getPadding() const = 0;
protected:
    virtual ~IPoolingLayer() {}
};

class IFullyConnectedLayer : public ILayer {
public:
    virtual void setNbOutputChannels(int nbOutputs) = 0;
    virtual int getNbOutputChannels() const = 0;
    virtual void setKernelWeights(Weights weights) = 0;
    virtual Weights getKernelWeights() const = 0;
    virtual void setBiasWeights(Weights weights) = 0;
    virtual Weights getBiasWeights() const = 0;
protected:
    virtual ~IFullyConnectedLayer() {}
};

class ISoftMaxLayer : public ILayer {      // over the channel dimension of a CHW tensor
protected:
    virtual ~ISoftMaxLayer() {}
};

// Output extent of a pooling layer; the Caffe parser installs Caffe's rounding (ceil, last window must start inside the image).
class IOutputDimensionsFormula {
public:
    virtual DimsHW compute(DimsHW inputDims, DimsHW kernelSize, DimsHW stride, DimsHW padding, DimsHW dilation, const char* layerName) const = 0;
protected:
    virtual ~IOutputDimensionsFormula() {}
};

class IShuffleLayer : public ILayer {
public:
    virtual void setReshapeDimensions(Dims dimensions) = 0;
    virtual Dims getReshapeDimensions() const = 0;
protected:
    virtual ~IShuffleLayer() {}
};

// ---- plugins (TensorRT 3/4 flavour: IPlugin / IPluginExt / IPluginFactory) --------------------------------------
class IPlugin {
public:
    virtual int getNbOutputs() const = 0;
    virtual Dims getOutputDimensions(int index, const Dims* inputs, int nbInputDims) = 0;
    virtual void configure(const Dims* inputDims, int nbInputs, const Dims* outputDims, int nbOutputs, int maxBatchSize) = 0;
    virtual int initialize() = 0;
    virtual void terminate() = 0;
    virtual std::size_t getWorkspaceSize(int maxBatchSize) const = 0;
    virtual int enqueue(int batchSize, const void* const* inputs, void** outputs, void* workspace, cudaStream_t stream) = 0;
    virtual std::size_t getSerializationSize() = 0;
    virtual void serialize(void* buffer) = 0;
    virtual ~IPlugin() {}
};

enum class PluginFormat : uint8_t { kNCHW = 0, kNC2HW2 = 1, kNHWC8 = 2 };
template <> inline int EnumMax<PluginFormat>() { return 3; }

class IPluginExt : public IPlugin {
public:
    virtual int getTensorRTVersion() const { return NV_TENSORRT_VERSION; }
    virtual bool supportsFormat(DataType type, PluginFormat format) const = 0;
    virtual void configureWithFormat(const Dims* inputDims, int nbInputs, const Dims* outputDims, int nbOutputs,
                                     DataType type, PluginFormat format, int maxBatchSize) = 0;
    virtual ~IPluginExt() {}
protected:
    // IPluginExt plugins are configured through configureWithFormat only.
    void configure(const Dims*, int, const Dims*, int, int) final {}
};

class IPluginLayer : public ILayer {
public:
    virtual IPlugin& getPlugin() = 0;
protected:
    virtual ~IPluginLayer() {}
};

class IPluginFactory {
public:
    virtual IPlugin* createPlugin(const char* layerName, const void* serialData, std::size_t serialLength) = 0;
};

// ---- network definition ------------------------------------------------------------------------------------------
class INetworkDefinition {
public:
    virtual ITensor* addInput(const char* name, DataType type, Dims dimensions) = 0;
    virtual void markOutput(ITensor& tensor) = 0;
    virtual IConvolutionLayer* addConvolution(ITensor& input, int nbOutputMaps, DimsHW kernelSize,
                                              Weights kernelWeights, Weights biasWeights) = 0;
    virtual IDeconvolutionLayer* addDeconvolution(ITensor& input, int nbOutputMaps, DimsHW kernelSize,
                                                  Weights kernelWeights, Weights biasWeights) = 0;
    virtual IActivationLayer* addActivation(ITensor& input, ActivationType type) = 0;
    virtual IScaleLayer* addScale(ITensor& input, ScaleMode mode, Weights shift, Weights scale, Weights power) = 0;
    virtual IConcatenationLayer* addConcatenation(ITensor* const* inputs, int nbInputs) = 0;
    virtual IElementWiseLayer* addElementWise(ITensor& input1, ITensor& input2, ElementWiseOperation op) = 0;
    virtual IShuffleLayer* addShuffle(ITensor& input) = 0;
    virtual IPluginLayer* addPlugin(ITensor* const* inputs, int nbInputs, IPlugin& plugin) = 0;
    virtual IPluginLayer* addPluginExt(ITensor* const* inputs, int nbInputs, IPluginExt& plugin) = 0;
    virtual int getNbLayers() const = 0;
    virtual ILayer* getLayer(int index) const = 0;
    virtual int getNbInputs() const = 0;
    virtual ITensor* getInput(int index) const = 0;
    virtual int getNbOutputs() const = 0;
    virtual ITensor* getOutput(int index) const = 0;
    virtual void destroy() = 0;
    virtual IPoolingLayer* addPooling(ITensor& input, PoolingType type, DimsHW windowSize) = 0;
    virtual IFullyConnectedLayer* addFullyConnected(ITensor& input, int nbOutputs, Weights kernelWeights, Weights biasWeights) = 0;
    virtual ISoftMaxLayer* addSoftMax(ITensor& input) = 0;
    virtual void setPoolingOutputDimensionsFormula(IOutputDimensionsFormula* formula) = 0;   // nullptr: floor ((in + 2 pad - k) / stride) + 1
    virtual IOutputDimensionsFormula& getPoolingOutputDimensionsFormula() const = 0;
protected:
    virtual ~INetworkDefinition() {}
};

// ---- engine / runtime ----------------------------------------------------------------------------------------------
class ICudaEngine;

class IExecutionContext {
public:
    // Synchronous: runs the network on an internal stream and waits (the callers' contract, sample_app/main.cpp:304).
    virtual bool execute(int batchSize, void** bindings) = 0;
    // Asynchronous on `stream`.
    virtual bool enqueue(int batchSize, void** bindings, cudaStream_t stream, cudaEvent_t* inputConsumed) = 0;
    virtual void setDebugSync(bool sync) = 0;
    virtual bool getDebugSync() const = 0;
    virtual void setProfiler(IProfiler*) = 0;
    virtual IProfiler* getProfiler() const = 0;
    virtual const ICudaEngine& getEngine() const = 0;
    virtual void destroy() = 0;
protected:
    virtual ~IExecutionContext() {}
};

class ICudaEngine {
public:
    virtual int getNbBindings() const = 0;
    virtual int getBindingIndex(const char* name) const = 0;
    virtual const char* getBindingName(int bindingIndex) const = 0;
    virtual bool bindingIsInput(int bindingIndex) const = 0;
    virtual Dims getBindingDimensions(int bindingIndex) const = 0;
    virtual DataType getBindingDataType(int bindingIndex) const = 0;
    virtual int getMaxBatchSize() const = 0;
    virtual int getNbLayers() const = 0;
    virtual std::size_t getWorkspaceSize() const = 0;
    virtual IHostMemory* serialize() const = 0;
    virtual IExecutionContext* createExecutionContext() = 0;
    virtual void destroy() = 0;
protected:
    virtual ~ICudaEngine() {}
};

// INT8 calibration interfaces: declared so that ros/packages/caffe_ros (int8_calibrator.h, tensor_net.cpp:293-301) compiles
// unchanged.  This engine has no INT8 path: platformHasFastInt8() is false, so the reference never switches it on; a builder that
// is put into INT8 mode anyway refuses to build (loudly).
enum class CalibrationAlgoType : int { kLEGACY_CALIBRATION = 0, kENTROPY_CALIBRATION = 1 };
class IInt8Calibrator {
public:
    virtual int getBatchSize() const = 0;
    virtual bool getBatch(void* bindings[], const char* names[], int nbBindings) = 0;
    virtual const void* readCalibrationCache(std::size_t& length) = 0;
    virtual void writeCalibrationCache(const void* ptr, std::size_t length) = 0;
    virtual CalibrationAlgoType getAlgorithm() = 0;
    virtual ~IInt8Calibrator() {}
};
class IInt8EntropyCalibrator : public IInt8Calibrator {
public:
    CalibrationAlgoType getAlgorithm() override { return CalibrationAlgoType::kENTROPY_CALIBRATION; }
    virtual ~IInt8EntropyCalibrator() {}
};

class IBuilder {
public:
    virtual INetworkDefinition* createNetwork() = 0;
    virtual void setMaxBatchSize(int batchSize) = 0;
    virtual int getMaxBatchSize() const = 0;
    virtual void setMaxWorkspaceSize(std::size_t workspaceSize) = 0;
    virtual std::size_t getMaxWorkspaceSize() const = 0;
    virtual void setHalf2Mode(bool mode) = 0;
    virtual bool getHalf2Mode() const = 0;
    virtual void setDebugSync(bool sync) = 0;
    virtual bool getDebugSync() const = 0;
    virtual void setMinFindIterations(int minFind) = 0;
    virtual int getMinFindIterations() const = 0;
    virtual void setAverageFindIterations(int avgFind) = 0;
    virtual int getAverageFindIterations() const = 0;
    virtual ICudaEngine* buildCudaEngine(INetworkDefinition& network) = 0;
    virtual bool platformHasFastFp16() const = 0;
    virtual bool platformHasFastInt8() const = 0;
    virtual void destroy() = 0;
    virtual void setInt8Mode(bool mode) = 0;
    virtual bool getInt8Mode() const = 0;
    virtual void setInt8Calibrator(IInt8Calibrator* calibrator) = 0;
protected:
    virtual ~IBuilder() {}
};

class IRuntime {
public:
    virtual ICudaEngine* deserializeCudaEngine(const void* blob, std::size_t size, IPluginFactory* pluginFactory) = 0;
    virtual void destroy() = 0;
protected:
    virtual ~IRuntime() {}
};

}  // namespace nvinfer1

extern "C" void* createInferBuilder_INTERNAL(void* logger, int version);
extern "C" void* createInferRuntime_INTERNAL(void* logger, int version);

namespace nvinfer1 {
namespace {
inline IBuilder* createInferBuilder(ILogger& logger) {
    return static_cast<IBuilder*>(createInferBuilder_INTERNAL(&logger, NV_TENSORRT_VERSION));
}
inline IRuntime* createInferRuntime(ILogger& logger) {
    return static_cast<IRuntime*>(createInferRuntime_INTERNAL(&logger, NV_TENSORRT_VERSION));
}
}  // namespace
}  // namespace nvinfer1

#endif  // REDTAIL_B200_NVINFER_H

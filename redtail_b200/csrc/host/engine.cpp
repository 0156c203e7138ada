// Mini inference engine behind include/NvInfer.h: graph capture (INetworkDefinition), build (shape resolution,
// plugin protocol, peephole fusion, static memory plan) and execution (one stream, optional per-layer profiling).
//
// It plays the role TensorRT plays for the reference (SURVEY.md layer L3): the generated builders call
// network->add*(), buildCudaEngine() walks the layers in insertion order, and execute() runs them -- every step is a
// call into the C-ABI of include/redtail_b200.h (native layers) or a plugin's enqueue().
//
// Fusion (REDTAIL_ENGINE_FUSION=0 disables it; the unfused graph is what the reference's TensorRT would run):
//   Conv2D/Deconv2D -> ELU                                  => one conv2d launch with ELU epilogue
//   Conv3D [-> Transform{1,0,2,3}] [-> ELU]                 => one conv3d launch (layout + ELU in the epilogue)
//   Conv3DTranspose [-> Slice] [-> +skip (kSUM)] [-> ELU]   => one transposed-conv launch
// Intermediate tensors of a fused chain are never materialised.
#include <cuda_runtime_api.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "internal_utils.h"
#include "op_info.h"
#include "redtail_b200.h"

using namespace nvinfer1;
using redtail::tensorrt::DimsUtils;
using redtail::tensorrt::IRedtailOp;
using redtail::tensorrt::OpInfo;
using redtail::tensorrt::OpKind;

namespace {

int rtType(DataType t) { return t == DataType::kHALF ? RT_F16 : RT_F32; }

size_t volume(const Dims& d) { return DimsUtils::getTensorSize(d); }

void logMsg(ILogger& log, ILogger::Severity sev, const std::string& s) { log.log(sev, s.c_str()); }

// Little-endian plan (de)serialisation helpers.
struct PlanWriter {
    std::string buf;
    template <typename T> void put(T v) { buf.append(reinterpret_cast<const char*>(&v), sizeof(T)); }
    void str(const std::string& v) { put<int32_t>(static_cast<int32_t>(v.size())); buf.append(v); }
    void dims(const Dims& d) { put<int32_t>(d.nbDims); for (int i = 0; i < Dims::MAX_DIMS; ++i) put<int32_t>(d.d[i]); }
    void align8() { while (buf.size() % 8) buf.push_back('\0'); }
    void weights(const Weights& w)
    {
        put<int32_t>(static_cast<int32_t>(w.type));
        put<int64_t>(w.values ? w.count : 0);
        align8();
        if (w.values && w.count > 0) buf.append(static_cast<const char*>(w.values), w.count * (w.type == DataType::kHALF ? 2 : 4));
    }
};
struct PlanReader {
    const char* base; size_t size; size_t pos = 0; bool ok = true;
    template <typename T> T get()
    {
        T v{};
        if (pos + sizeof(T) > size) { ok = false; return v; }
        memcpy(&v, base + pos, sizeof(T));
        pos += sizeof(T);
        return v;
    }
    std::string str()
    {
        const int32_t n = get<int32_t>();
        if (!ok || n < 0 || pos + n > size) { ok = false; return std::string(); }
        std::string v(base + pos, n);
        pos += n;
        return v;
    }
    Dims dims()
    {
        Dims d{};
        d.nbDims = get<int32_t>();
        for (int i = 0; i < Dims::MAX_DIMS; ++i) d.d[i] = get<int32_t>();
        // The plan does not store the dimension types; a 3-D tensor of these networks is CHW (DimsCHW), which callers of
        // getBindingDimensions() on a deserialised engine check (ros/packages/caffe_ros/src/tensor_net.cpp:21-28).
        if (d.nbDims == 3) d.type[0] = DimensionType::kCHANNEL;
        return d;
    }
    void align8() { pos = (pos + 7) & ~static_cast<size_t>(7); }
    Weights weights()          // values point into the plan copy the engine keeps alive
    {
        Weights w{DataType::kFLOAT, nullptr, 0};
        w.type = static_cast<DataType>(get<int32_t>());
        w.count = get<int64_t>();
        align8();
        const size_t bytes = static_cast<size_t>(w.count) * (w.type == DataType::kHALF ? 2 : 4);
        if (!ok || w.count < 0 || pos + bytes > size) { ok = false; w.count = 0; return w; }
        w.values = w.count > 0 ? base + pos : nullptr;
        pos += bytes;
        return w;
    }
};
constexpr char kPlanMagic[8] = {'R', 'T', 'B', '2', 'P', 'L', 'A', 'N'};
constexpr int32_t kPlanVersion = 1;

class HostMemoryImpl : public IHostMemory {
public:
    explicit HostMemoryImpl(std::string d) : data_(std::move(d)) {}
    void* data() const override { return const_cast<char*>(data_.data()); }
    std::size_t size() const override { return data_.size(); }
    DataType type() const override { return DataType::kINT8; }
    void destroy() override { delete this; }
private:
    std::string data_;
};

// ------------------------------------------------------------------------------------------------------------------
// Network definition
// ------------------------------------------------------------------------------------------------------------------
class NetworkImpl;

struct TensorImpl : public ITensor {
    NetworkImpl* net = nullptr;
    int id = -1;
    std::string name;
    Dims dims{};
    DataType type = DataType::kFLOAT;
    bool is_input = false, is_output = false;
    int producer = -1;      // layer index, -1 for network inputs

    void setName(const char* n) override { name = n ? n : ""; }
    const char* getName() const override { return name.c_str(); }
    Dims getDimensions() const override;
    DataType getType() const override { return type; }
    bool isNetworkInput() const override { return is_input; }
    bool isNetworkOutput() const override { return is_output; }
};

enum class LKind { kConv, kDeconv, kScale, kEltwise, kConcat, kActivation, kShuffle, kPlugin, kPooling, kFullyConnected, kSoftMax };

struct LayerData {
    LKind kind;
    LayerType type;
    std::string name;
    std::vector<TensorImpl*> in, out;
    int nb_out_maps = 0;
    DimsHW ksize{1, 1}, stride{1, 1}, pad{0, 0};
    Weights kw{DataType::kFLOAT, nullptr, 0}, bw{DataType::kFLOAT, nullptr, 0};
    ScaleMode smode = ScaleMode::kUNIFORM;
    Weights shift{DataType::kFLOAT, nullptr, 0}, scale{DataType::kFLOAT, nullptr, 0}, power{DataType::kFLOAT, nullptr, 0};
    ElementWiseOperation eop = ElementWiseOperation::kSUM;
    ActivationType act = ActivationType::kRELU;
    Dims reshape{};
    bool has_reshape = false;
    IPlugin* plugin = nullptr;
    IPluginExt* plugin_ext = nullptr;
    PoolingType pool = PoolingType::kMAX;
    int pool_oh = 0, pool_ow = 0;      // output extent fixed by a plan (the formula object is not serialisable); 0 = use the formula
};

struct LayerNode {
    virtual ~LayerNode() {}
    LayerData d;
    virtual ILayer* iface() = 0;
};

template <class Iface>
struct LayerT : public Iface, public LayerNode {
    ILayer* iface() override { return this; }
    LayerType getType() const override { return d.type; }
    void setName(const char* n) override { d.name = n ? n : ""; }
    const char* getName() const override { return d.name.c_str(); }
    int getNbInputs() const override { return static_cast<int>(d.in.size()); }
    ITensor* getInput(int i) const override { return i >= 0 && i < getNbInputs() ? d.in[i] : nullptr; }
    int getNbOutputs() const override { return static_cast<int>(d.out.size()); }
    ITensor* getOutput(int i) const override { return i >= 0 && i < getNbOutputs() ? d.out[i] : nullptr; }
};

template <class Iface>
struct ConvLikeLayer : public LayerT<Iface> {
    void setKernelSize(DimsHW k) override { this->d.ksize = k; }
    DimsHW getKernelSize() const override { return this->d.ksize; }
    void setNbOutputMaps(int n) override { this->d.nb_out_maps = n; }
    int getNbOutputMaps() const override { return this->d.nb_out_maps; }
    void setStride(DimsHW s) override { this->d.stride = s; }
    DimsHW getStride() const override { return this->d.stride; }
    void setPadding(DimsHW p) override { this->d.pad = p; }
    DimsHW getPadding() const override { return this->d.pad; }
    void setKernelWeights(Weights w) override { this->d.kw = w; }
    Weights getKernelWeights() const override { return this->d.kw; }
    void setBiasWeights(Weights w) override { this->d.bw = w; }
    Weights getBiasWeights() const override { return this->d.bw; }
};
using ConvLayer = ConvLikeLayer<IConvolutionLayer>;
using DeconvLayer = ConvLikeLayer<IDeconvolutionLayer>;
struct ScaleLayer : public LayerT<IScaleLayer> { ScaleMode getMode() const override { return d.smode; } };
struct EltwiseLayer : public LayerT<IElementWiseLayer> { ElementWiseOperation getOperation() const override { return d.eop; } };
struct ConcatLayer : public LayerT<IConcatenationLayer> {};
struct ActivationLayer : public LayerT<IActivationLayer> { ActivationType getActivationType() const override { return d.act; } };
struct ShuffleLayer : public LayerT<IShuffleLayer> {
    void setReshapeDimensions(Dims dims) override { d.reshape = dims; d.has_reshape = true; }
    Dims getReshapeDimensions() const override { return d.reshape; }
};
struct PluginLayer : public LayerT<IPluginLayer> { IPlugin& getPlugin() override { return *d.plugin; } };
struct PoolingLayer : public LayerT<IPoolingLayer> {
    void setPoolingType(PoolingType t) override { d.pool = t; }
    PoolingType getPoolingType() const override { return d.pool; }
    void setWindowSize(DimsHW k) override { d.ksize = k; }
    DimsHW getWindowSize() const override { return d.ksize; }
    void setStride(DimsHW s) override { d.stride = s; }
    DimsHW getStride() const override { return d.stride; }
    void setPadding(DimsHW p) override { d.pad = p; }
    DimsHW getPadding() const override { return d.pad; }
};
struct FullyConnectedLayer : public LayerT<IFullyConnectedLayer> {
    void setNbOutputChannels(int n) override { d.nb_out_maps = n; }
    int getNbOutputChannels() const override { return d.nb_out_maps; }
    void setKernelWeights(Weights w) override { d.kw = w; }
    Weights getKernelWeights() const override { return d.kw; }
    void setBiasWeights(Weights w) override { d.bw = w; }
    Weights getBiasWeights() const override { return d.bw; }
};
struct SoftMaxLayer : public LayerT<ISoftMaxLayer> {};
// TensorRT's default pooling extent: floor((in + 2 pad - k) / stride) + 1.
struct FloorPoolingFormula : public IOutputDimensionsFormula {
    DimsHW compute(DimsHW in, DimsHW k, DimsHW stride, DimsHW pad, DimsHW, const char*) const override
    {
        return DimsHW((in.h() + 2 * pad.h() - k.h()) / stride.h() + 1, (in.w() + 2 * pad.w() - k.w()) / stride.w() + 1);
    }
};

class NetworkImpl : public INetworkDefinition {
public:
    explicit NetworkImpl(ILogger& log) : log_(log) {}
    ~NetworkImpl() override {}

    ITensor* addInput(const char* name, DataType type, Dims dims) override
    {
        if (dims.nbDims < 1 || dims.nbDims > 4) { logMsg(log_, ILogger::Severity::kERROR, "addInput: rank must be 1..4"); return nullptr; }
        TensorImpl* t = newTensor(name ? name : "");
        t->type = type; t->dims = dims; t->is_input = true;
        inputs_.push_back(t);
        return t;
    }
    void markOutput(ITensor& tensor) override
    {
        TensorImpl* t = static_cast<TensorImpl*>(&tensor);
        t->is_output = true;
        outputs_.push_back(t);
    }
    IConvolutionLayer* addConvolution(ITensor& input, int nbOutputMaps, DimsHW kernelSize, Weights kw, Weights bw) override
    {
        auto* l = add<ConvLayer>(LKind::kConv, LayerType::kCONVOLUTION, {&input}, 1);
        l->d.nb_out_maps = nbOutputMaps; l->d.ksize = kernelSize; l->d.kw = kw; l->d.bw = bw;
        return l;
    }
    IDeconvolutionLayer* addDeconvolution(ITensor& input, int nbOutputMaps, DimsHW kernelSize, Weights kw, Weights bw) override
    {
        auto* l = add<DeconvLayer>(LKind::kDeconv, LayerType::kDECONVOLUTION, {&input}, 1);
        l->d.nb_out_maps = nbOutputMaps; l->d.ksize = kernelSize; l->d.kw = kw; l->d.bw = bw;
        return l;
    }
    IActivationLayer* addActivation(ITensor& input, ActivationType type) override
    {
        auto* l = add<ActivationLayer>(LKind::kActivation, LayerType::kACTIVATION, {&input}, 1);
        l->d.act = type;
        return l;
    }
    IScaleLayer* addScale(ITensor& input, ScaleMode mode, Weights shift, Weights scale, Weights power) override
    {
        auto* l = add<ScaleLayer>(LKind::kScale, LayerType::kSCALE, {&input}, 1);
        l->d.smode = mode; l->d.shift = shift; l->d.scale = scale; l->d.power = power;
        return l;
    }
    IConcatenationLayer* addConcatenation(ITensor* const* inputs, int nbInputs) override
    {
        std::vector<ITensor*> in(inputs, inputs + nbInputs);
        return add<ConcatLayer>(LKind::kConcat, LayerType::kCONCATENATION, in, 1);
    }
    IElementWiseLayer* addElementWise(ITensor& a, ITensor& b, ElementWiseOperation op) override
    {
        auto* l = add<EltwiseLayer>(LKind::kEltwise, LayerType::kELEMENTWISE, {&a, &b}, 1);
        l->d.eop = op;
        return l;
    }
    IShuffleLayer* addShuffle(ITensor& input) override
    {
        return add<ShuffleLayer>(LKind::kShuffle, LayerType::kSHUFFLE, {&input}, 1);
    }
    IPluginLayer* addPlugin(ITensor* const* inputs, int nbInputs, IPlugin& plugin) override
    {
        std::vector<ITensor*> in(inputs, inputs + nbInputs);
        auto* l = add<PluginLayer>(LKind::kPlugin, LayerType::kPLUGIN, in, plugin.getNbOutputs());
        l->d.plugin = &plugin;
        return l;
    }
    IPluginLayer* addPluginExt(ITensor* const* inputs, int nbInputs, IPluginExt& plugin) override
    {
        std::vector<ITensor*> in(inputs, inputs + nbInputs);
        auto* l = add<PluginLayer>(LKind::kPlugin, LayerType::kPLUGIN, in, plugin.getNbOutputs());
        l->d.plugin = &plugin; l->d.plugin_ext = &plugin;
        return l;
    }
    IPoolingLayer* addPooling(ITensor& input, PoolingType type, DimsHW windowSize) override
    {
        auto* l = add<PoolingLayer>(LKind::kPooling, LayerType::kPOOLING, {&input}, 1);
        l->d.pool = type; l->d.ksize = windowSize; l->d.stride = DimsHW(1, 1); l->d.pad = DimsHW(0, 0);
        return l;
    }
    IFullyConnectedLayer* addFullyConnected(ITensor& input, int nbOutputs, Weights kw, Weights bw) override
    {
        auto* l = add<FullyConnectedLayer>(LKind::kFullyConnected, LayerType::kFULLY_CONNECTED, {&input}, 1);
        l->d.nb_out_maps = nbOutputs; l->d.kw = kw; l->d.bw = bw;
        return l;
    }
    ISoftMaxLayer* addSoftMax(ITensor& input) override { return add<SoftMaxLayer>(LKind::kSoftMax, LayerType::kSOFTMAX, {&input}, 1); }
    void setPoolingOutputDimensionsFormula(IOutputDimensionsFormula* f) override { pool_formula_ = f; invalidate(); }
    IOutputDimensionsFormula& getPoolingOutputDimensionsFormula() const override
    {
        return pool_formula_ ? *pool_formula_ : const_cast<FloorPoolingFormula&>(floor_formula_);
    }
    int getNbLayers() const override { return static_cast<int>(layers_.size()); }
    ILayer* getLayer(int i) const override { return i >= 0 && i < getNbLayers() ? layers_[i]->iface() : nullptr; }
    int getNbInputs() const override { return static_cast<int>(inputs_.size()); }
    ITensor* getInput(int i) const override { return i >= 0 && i < getNbInputs() ? inputs_[i] : nullptr; }
    int getNbOutputs() const override { return static_cast<int>(outputs_.size()); }
    ITensor* getOutput(int i) const override { return i >= 0 && i < getNbOutputs() ? outputs_[i] : nullptr; }
    void destroy() override { delete this; }

    // Computes the dims of every tensor not yet resolved (layers are in topological = insertion order).
    bool resolve()
    {
        for (size_t li = resolved_; li < layers_.size(); ++li) {
            if (!resolveLayer(layers_[li]->d)) return false;
        }
        resolved_ = layers_.size();
        return true;
    }
    // A layer's parameters (stride / padding / reshape) may be set after add*(): force re-resolution before build.
    void invalidate()
    {
        resolved_ = 0;
        if (!pool_from_plan_)
            for (auto& l : layers_) if (l->d.kind == LKind::kPooling) { l->d.pool_oh = 0; l->d.pool_ow = 0; }
    }
    bool pool_from_plan_ = false;       // deserialised networks carry the pooling extents the original formula produced

    ILogger& log_;
    IOutputDimensionsFormula* pool_formula_ = nullptr;
    FloorPoolingFormula floor_formula_;
    std::vector<std::unique_ptr<TensorImpl>> tensors_;
    std::vector<std::unique_ptr<LayerNode>> layers_;
    std::vector<TensorImpl*> inputs_, outputs_;

private:
    TensorImpl* newTensor(const std::string& name)
    {
        tensors_.emplace_back(new TensorImpl());
        TensorImpl* t = tensors_.back().get();
        t->net = this; t->id = static_cast<int>(tensors_.size()) - 1; t->name = name;
        return t;
    }
    template <class L>
    L* add(LKind kind, LayerType type, std::vector<ITensor*> in, int nb_out)
    {
        L* l = new L();
        layers_.emplace_back(l);
        l->d.kind = kind; l->d.type = type;
        l->d.name = "(Unnamed Layer* " + std::to_string(layers_.size() - 1) + ")";
        for (ITensor* t : in) l->d.in.push_back(static_cast<TensorImpl*>(t));
        for (int i = 0; i < nb_out; ++i) {
            TensorImpl* t = newTensor(l->d.name + "_output_" + std::to_string(i));
            t->producer = static_cast<int>(layers_.size()) - 1;
            l->d.out.push_back(t);
        }
        return l;
    }
    bool fail(const LayerData& d, const std::string& why)
    {
        logMsg(log_, ILogger::Severity::kERROR, d.name + ": " + why);
        return false;
    }
    bool resolveLayer(LayerData& d)
    {
        for (auto* t : d.in)
            if (t == nullptr) return fail(d, "null input tensor");
        const Dims in0 = d.in[0]->dims;
        switch (d.kind) {
            case LKind::kConv:
            case LKind::kDeconv: {
                if (in0.nbDims != 3) return fail(d, "2-D convolution expects a CHW input");
                const int cin = in0.d[0];
                if (d.kw.count != static_cast<int64_t>(cin) * d.nb_out_maps * d.ksize.h() * d.ksize.w())
                    return fail(d, "kernel weight count does not match the layer shape");
                if (d.bw.count != 0 && d.bw.count != d.nb_out_maps) return fail(d, "bias count mismatch");
                int ho, wo;
                if (d.kind == LKind::kConv) {
                    ho = (in0.d[1] + 2 * d.pad.h() - d.ksize.h()) / d.stride.h() + 1;
                    wo = (in0.d[2] + 2 * d.pad.w() - d.ksize.w()) / d.stride.w() + 1;
                } else {
                    ho = (in0.d[1] - 1) * d.stride.h() + d.ksize.h() - 2 * d.pad.h();
                    wo = (in0.d[2] - 1) * d.stride.w() + d.ksize.w() - 2 * d.pad.w();
                }
                if (ho <= 0 || wo <= 0) return fail(d, "empty output");
                d.out[0]->dims = DimsCHW(d.nb_out_maps, ho, wo);
                break;
            }
            case LKind::kScale:
                if (d.smode == ScaleMode::kCHANNEL) {
                    if (in0.nbDims != 3) return fail(d, "per-channel scale expects a CHW input");
                    for (const Weights* w : {&d.shift, &d.scale, &d.power})
                        if (w->count != 0 && w->count != in0.d[0]) return fail(d, "per-channel scale: weight count != channels");
                } else if (d.smode != ScaleMode::kUNIFORM) return fail(d, "ScaleMode::kELEMENTWISE is not supported");
                d.out[0]->dims = in0;
                break;
            case LKind::kPooling: {
                if (in0.nbDims != 3) return fail(d, "pooling expects a CHW input");
                if (d.pool == PoolingType::kMAX_AVERAGE_BLEND) return fail(d, "PoolingType::kMAX_AVERAGE_BLEND is not supported");
                if (d.ksize.h() != d.ksize.w() || d.stride.h() != d.stride.w() || d.pad.h() != d.pad.w()) return fail(d, "pooling: square windows only");
                DimsHW o(d.pool_oh, d.pool_ow);
                if (d.pool_oh <= 0)
                    o = getPoolingOutputDimensionsFormula().compute(DimsHW(in0.d[1], in0.d[2]), d.ksize, d.stride, d.pad, DimsHW(1, 1), d.name.c_str());
                if (o.h() <= 0 || o.w() <= 0) return fail(d, "empty output");
                d.pool_oh = o.h(); d.pool_ow = o.w();
                d.out[0]->dims = DimsCHW(in0.d[0], o.h(), o.w());
                break;
            }
            case LKind::kFullyConnected:
                if (d.nb_out_maps <= 0 || d.kw.count != static_cast<int64_t>(volume(in0)) * d.nb_out_maps) return fail(d, "fully connected: weight count mismatch");
                if (d.bw.count != 0 && d.bw.count != d.nb_out_maps) return fail(d, "bias count mismatch");
                d.out[0]->dims = DimsCHW(d.nb_out_maps, 1, 1);
                break;
            case LKind::kSoftMax:
                if (in0.nbDims != 3) return fail(d, "soft-max expects a CHW input");
                d.out[0]->dims = in0;
                break;
            case LKind::kActivation:
                d.out[0]->dims = in0;
                break;
            case LKind::kEltwise:
                if (d.eop != ElementWiseOperation::kSUM) return fail(d, "only ElementWiseOperation::kSUM is supported");
                if (!DimsUtils::areEqual(in0, d.in[1]->dims)) return fail(d, "element-wise inputs differ in shape");
                d.out[0]->dims = in0;
                break;
            case LKind::kConcat: {
                Dims o = in0;
                if (o.nbDims != 3) return fail(d, "concatenation expects CHW inputs");
                for (size_t i = 1; i < d.in.size(); ++i) {
                    const Dims& di = d.in[i]->dims;
                    if (di.nbDims != 3 || di.d[1] != o.d[1] || di.d[2] != o.d[2]) return fail(d, "concat inputs differ in H/W");
                    o.d[0] += di.d[0];
                }
                d.out[0]->dims = o;
                break;
            }
            case LKind::kShuffle: {
                Dims o = d.has_reshape ? d.reshape : in0;
                if (volume(o) != volume(in0)) return fail(d, "reshape changes the element count");
                d.out[0]->dims = o;
                break;
            }
            case LKind::kPlugin: {
                std::vector<Dims> ind;
                for (auto* t : d.in) ind.push_back(t->dims);
                for (size_t i = 0; i < d.out.size(); ++i)
                    d.out[i]->dims = d.plugin->getOutputDimensions(static_cast<int>(i), ind.data(), static_cast<int>(ind.size()));
                break;
            }
        }
        return true;
    }
    size_t resolved_ = 0;
};

Dims TensorImpl::getDimensions() const
{
    if (net) net->resolve();
    return dims;
}

// ------------------------------------------------------------------------------------------------------------------
// Engine
// ------------------------------------------------------------------------------------------------------------------
struct TensorSlot {
    std::string name;
    Dims dims{};
    size_t elems = 0;          // per sample
    int binding = -1;          // >= 0: caller-owned buffer
    int alias_of = -1;         // shares storage with another tensor (shuffle)
    int pair_of = -1;          // siamese partner: stored right behind tensor `pair_of` (at + batch * elems) so that one launch
                               // of batch 2N covers both towers
    bool has_partner = false;  // some tensor names this one as its pair_of: the allocation is twice the size
    bool forced_split = false; // engine-made tensor that only exists in RT_LAYOUT_SPLIT16 (the im2col matrix of a large-filter conv)
    size_t offset = 0;         // arena offset (bytes) when neither binding nor alias
    int first = -1, last = -1; // step liveness
    bool used = false;
};

// A fused 3-D (transposed) convolution whose plan is created after the layout pass.
struct ConvStep {
    rt_conv3d_desc desc{};
    rt_conv3d_plan* plan = nullptr;
    int in_id = -1, out_id = -1, skip_id = -1;
    int batch_mul = 1;         // 2: the step also covers its siamese twin (same weights; inputs and outputs stored as pairs)
    std::string name;
    // CostVolume -> Conv3D pair replaced by the separable formulation (rt_costvol_conv3d_*): inputs are the two feature maps.
    bool cvfused = false;
    rt_costvol_conv3d_desc cvdesc{};
    rt_cvconv_plan* cvplan = nullptr;
    int l_id = -1, r_id = -1;
};

struct PadInfo { int in_id; int planes; int layer; };

struct Step {
    std::string name;
    std::vector<int> in, out;  // tensor ids (for liveness)
    size_t workspace = 0;
    ConvStep* conv = nullptr;                                   // set for fused 3-D convolution steps
    int costvol_c = 0, costvol_h = 0, costvol_w = 0, costvol_d = 0;   // set for concat cost-volume steps
    bool is_transform = false;                                  // Transform{1,0,2,3} plugin step
    int softargmax = 0;                                         // SoftargmaxPlugin step: 1 = kMin, 2 = kMax (may fuse into its producer)
    bool dropped = false;
    // ptr(id) resolves a tensor id to its device pointer for this execution.
    std::function<int(int batch, const std::function<void*(int)>& ptr, void* workspace, cudaStream_t)> run;
};

class EngineImpl;

class ContextImpl : public IExecutionContext {
public:
    explicit ContextImpl(EngineImpl* e);
    ~ContextImpl() override
    {
        for (auto& g : graphs_) cudaGraphExecDestroy(g.exec);
        if (stream_) cudaStreamDestroy(stream_);
        if (arena_) cudaFree(arena_);
        if (workspace_) cudaFree(workspace_);
    }
    bool execute(int batchSize, void** bindings) override;
    bool enqueue(int batchSize, void** bindings, cudaStream_t stream, cudaEvent_t* inputConsumed) override;
    void setDebugSync(bool sync) override { debug_sync_ = sync; }
    bool getDebugSync() const override { return debug_sync_; }
    void setProfiler(IProfiler* p) override { profiler_ = p; }
    IProfiler* getProfiler() const override { return profiler_; }
    const ICudaEngine& getEngine() const override;
    void destroy() override { delete this; }

private:
    bool run(int batchSize, void** bindings, cudaStream_t stream, bool profile);
    bool launch(int batchSize, void** bindings, cudaStream_t stream);
    // One inference = ~25 kernel launches of 30-1000 us: the launch sequence of a (batch, bindings) pair is captured into a
    // CUDA graph the second time it is seen and replayed afterwards (REDTAIL_ENGINE_GRAPH=0 keeps eager launches).
    struct GraphEntry { int batch; std::vector<void*> bindings; cudaGraphExec_t exec; uint64_t launches; uint64_t stamp; };
    std::vector<GraphEntry> graphs_;
    std::vector<std::pair<int, std::vector<void*>>> seen_;    // keys executed eagerly once (capture happens on the second use)
    bool graphs_enabled_ = true;
    uint64_t stamp_ = 0;
    EngineImpl* engine_;
    cudaStream_t stream_ = nullptr;
    // Activations and scratch belong to the CONTEXT (as in TensorRT): several contexts of one engine may run concurrently
    // on different streams without touching each other's intermediate tensors.
    void* arena_ = nullptr;
    void* workspace_ = nullptr;
    bool alloc_ok_ = true;
    IProfiler* profiler_ = nullptr;
    bool debug_sync_ = false;
};

class EngineImpl : public ICudaEngine {
public:
    EngineImpl(ILogger& log) : log_(log) {}
    ~EngineImpl() override
    {
        for (auto* p : configured_plugins_) p->terminate();
        for (auto* p : conv2d_plans_) rt_conv2d_destroy(p);
        for (auto* p : conv3d_plans_) rt_conv3d_destroy(p);
        for (auto* p : cvconv_plans_) rt_costvol_conv3d_destroy(p);
        for (void* p : dev_blobs_) cudaFree(p);
    }
    int getNbBindings() const override { return static_cast<int>(bindings_.size()); }
    int getBindingIndex(const char* name) const override
    {
        for (size_t i = 0; i < bindings_.size(); ++i)
            if (slots_[bindings_[i]].name == name) return static_cast<int>(i);
        return -1;
    }
    const char* getBindingName(int i) const override { return i >= 0 && i < getNbBindings() ? slots_[bindings_[i]].name.c_str() : nullptr; }
    bool bindingIsInput(int i) const override { return i >= 0 && i < nb_inputs_; }
    Dims getBindingDimensions(int i) const override { return i >= 0 && i < getNbBindings() ? slots_[bindings_[i]].dims : Dims{}; }
    DataType getBindingDataType(int) const override { return DataType::kFLOAT; }
    int getMaxBatchSize() const override { return max_batch_; }
    int getNbLayers() const override { return static_cast<int>(steps_.size()); }
    size_t getWorkspaceSize() const override { return workspace_bytes_; }
    // The plan is the network description (layers, parameters, weights, the plugins' own serialisation blobs) plus the
    // build settings: building is deterministic and takes milliseconds (no autotuning), so deserialisation replays the
    // description through the same INetworkDefinition calls and builds again -- plugin layers come back through the
    // caller's IPluginFactory exactly as in TensorRT (sample_app/main.cpp:207-220).
    IHostMemory* serialize() const override
    {
        if (plan_.empty()) {
            logMsg(log_, ILogger::Severity::kERROR, "ICudaEngine::serialize: a plugin of this network does not implement serialize().");
            return nullptr;
        }
        return new HostMemoryImpl(plan_);
    }
    IExecutionContext* createExecutionContext() override { return new ContextImpl(this); }
    void destroy() override { delete this; }

    bool build(NetworkImpl& net, int max_batch, bool half2);
    static std::string serializeNetwork(NetworkImpl& net, int max_batch, bool half2);

    std::string plan_;                                  // serialised network (empty if some plugin cannot serialise)
    std::shared_ptr<std::string> plan_storage_;         // deserialised engines: the blob copy the layer weights point into
    ILogger& log_;
    int max_batch_ = 1;
    int nb_inputs_ = 0;
    std::vector<TensorSlot> slots_;
    std::vector<int> bindings_;          // binding index -> tensor id
    std::vector<Step> steps_;
    bool graph_safe_ = true;             // false when a step runs a plugin this library does not know (its enqueue() may synchronise)
    size_t arena_bytes_ = 0;             // per execution context
    size_t workspace_bytes_ = 0;         // per execution context
    std::vector<IPlugin*> configured_plugins_;
    std::vector<rt_conv2d_plan*> conv2d_plans_;
    std::vector<rt_conv3d_plan*> conv3d_plans_;
    std::vector<rt_cvconv_plan*> cvconv_plans_;
    std::vector<std::unique_ptr<ConvStep>> conv_steps_;
    std::vector<void*> dev_blobs_;       // per-channel scale / shift arrays and fully-connected weights of the native layers
    std::vector<std::vector<float>> host_blobs_;   // re-laid-out host weights the conv descriptors point to until the plans exist

    // Host weights (fp32 or fp16) -> device fp32 array of `count` elements (`fill` where the layer has none); nullptr on failure.
    float* deviceArray(const Weights& w, int64_t count, float fill)
    {
        std::vector<float> h(static_cast<size_t>(count), fill);
        if (w.values && w.count == count) {
            if (w.type == DataType::kHALF) {
                const uint16_t* s = static_cast<const uint16_t*>(w.values);
                for (int64_t i = 0; i < count; ++i) {           // fp16 -> fp32 on the host
                    const uint32_t sign = (s[i] >> 15) & 1u, ex = (s[i] >> 10) & 31u, man = s[i] & 1023u;
                    float v;
                    if (ex == 0) v = ldexpf(static_cast<float>(man), -24);
                    else if (ex == 31) v = man ? NAN : INFINITY;
                    else v = ldexpf(static_cast<float>(man | 1024u), static_cast<int>(ex) - 25);
                    h[static_cast<size_t>(i)] = sign ? -v : v;
                }
            } else memcpy(h.data(), w.values, static_cast<size_t>(count) * sizeof(float));
        }
        void* d = nullptr;
        if (cudaMalloc(&d, h.size() * sizeof(float)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
        dev_blobs_.push_back(d);
        return static_cast<float*>(d);
    }

private:
    bool fail(const std::string& s) { logMsg(log_, ILogger::Severity::kERROR, s); return false; }
    bool assignLayoutsAndCreatePlans(bool fusion);
    void pairSiameseSteps();
    bool planMemory();
};

const ICudaEngine& ContextImpl::getEngine() const { return *engine_; }

ContextImpl::ContextImpl(EngineImpl* e) : engine_(e)
{
    alloc_ok_ = cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking) == cudaSuccess;
    if (alloc_ok_ && e->arena_bytes_ > 0) alloc_ok_ = cudaMalloc(&arena_, e->arena_bytes_) == cudaSuccess;
    if (alloc_ok_ && e->workspace_bytes_ > 0) alloc_ok_ = cudaMalloc(&workspace_, e->workspace_bytes_) == cudaSuccess;
    if (!alloc_ok_) logMsg(e->log_, ILogger::Severity::kERROR, "createExecutionContext: cudaMalloc of the activation arena / workspace failed");
    const char* g = getenv("REDTAIL_ENGINE_GRAPH");
    graphs_enabled_ = !(g && g[0] == '0') && e->graph_safe_ && getenv("REDTAIL_ENGINE_TRACE") == nullptr;   // tracing synchronises every step
}

// Eager the first time a (batch, bindings) pair is seen (first-use work such as cudaFuncSetAttribute happens there),
// captured into a graph the second time, replayed from then on.
bool ContextImpl::launch(int batchSize, void** bindings, cudaStream_t stream)
{
    EngineImpl& e = *engine_;
    if (!graphs_enabled_ || debug_sync_ || batchSize < 1 || batchSize > e.max_batch_ || bindings == nullptr)
        return run(batchSize, bindings, stream, false);
    std::vector<void*> key(bindings, bindings + e.getNbBindings());
    for (auto& g : graphs_)
        if (g.batch == batchSize && g.bindings == key) {
            g.stamp = ++stamp_;
            if (cudaGraphLaunch(g.exec, stream) != cudaSuccess) {
                logMsg(e.log_, ILogger::Severity::kERROR, "enqueue: cudaGraphLaunch failed");
                return false;
            }
            rt_add_launch_count(g.launches);
            return true;
        }
    bool second = false;
    for (auto& k : seen_) second = second || (k.first == batchSize && k.second == key);
    if (!second) {
        if (seen_.size() >= 16) seen_.erase(seen_.begin());
        seen_.emplace_back(batchSize, key);
        return run(batchSize, bindings, stream, false);
    }
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)
        return run(batchSize, bindings, stream, false);           // the caller is capturing already: just record into its graph
    if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        (void)cudaGetLastError();
        graphs_enabled_ = false;
        return run(batchSize, bindings, stream, false);
    }
    const uint64_t l0 = rt_launch_count();
    const bool ok = run(batchSize, bindings, stream, false);
    const uint64_t captured = rt_launch_count() - l0;
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(stream, &graph);
    cudaGraphExec_t exec = nullptr;
    if (!ok || ce != cudaSuccess || graph == nullptr || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        (void)cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        graphs_enabled_ = false;                                  // e.g. a third-party plugin that synchronises in enqueue()
        logMsg(e.log_, ILogger::Severity::kWARNING, "engine: CUDA graph capture of the step sequence failed, staying with eager launches");
        return ok ? run(batchSize, bindings, stream, false) : false;
    }
    cudaGraphDestroy(graph);
    if (graphs_.size() >= 8) {                                    // least recently used entry makes room
        size_t lru = 0;
        for (size_t i = 1; i < graphs_.size(); ++i) if (graphs_[i].stamp < graphs_[lru].stamp) lru = i;
        cudaGraphExecDestroy(graphs_[lru].exec);
        graphs_.erase(graphs_.begin() + lru);
    }
    graphs_.push_back(GraphEntry{batchSize, key, exec, captured, ++stamp_});
    if (cudaGraphLaunch(exec, stream) != cudaSuccess) return false;
    return true;                                                  // (the captured launches were counted while capturing)
}

// Weight helpers ---------------------------------------------------------------------------------------------------
float weightScalar(const Weights& w, float dflt)
{
    if (w.count < 1 || w.values == nullptr) return dflt;
    if (w.type == DataType::kFLOAT) return *static_cast<const float*>(w.values);
    // fp16 scalar
    const uint16_t h = *static_cast<const uint16_t*>(w.values);
    const uint32_t sign = (h & 0x8000u) << 16, exp = (h >> 10) & 0x1F, man = h & 0x3FF;
    uint32_t bits;
    if (exp == 0) bits = sign;        // zero / subnormal scalars are not meaningful for scale layers
    else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

bool EngineImpl::build(NetworkImpl& net, int max_batch, bool half2)
{
    max_batch_ = max_batch > 0 ? max_batch : 1;
    net.invalidate();
    if (!net.resolve()) return false;
    if (net.outputs_.empty()) return fail("network has no outputs");

    const char* fusion_env = getenv("REDTAIL_ENGINE_FUSION");
    const bool fusion = !(fusion_env && fusion_env[0] == '0');

    // Tensor table.
    slots_.resize(net.tensors_.size());
    for (size_t i = 0; i < net.tensors_.size(); ++i) {
        const TensorImpl& t = *net.tensors_[i];
        slots_[i].name = t.name; slots_[i].dims = t.dims; slots_[i].elems = volume(t.dims);
    }
    for (auto* t : net.inputs_) { slots_[t->id].binding = static_cast<int>(bindings_.size()); bindings_.push_back(t->id); }
    nb_inputs_ = static_cast<int>(bindings_.size());
    for (auto* t : net.outputs_) {
        if (slots_[t->id].binding >= 0) return fail("tensor '" + t->name + "' bound twice");
        slots_[t->id].binding = static_cast<int>(bindings_.size()); bindings_.push_back(t->id);
    }

    // Consumer counts (for fusion legality).
    const int nl = static_cast<int>(net.layers_.size());
    std::vector<int> consumers(net.tensors_.size(), 0);
    std::vector<std::vector<int>> consumer_layers(net.tensors_.size());
    for (int li = 0; li < nl; ++li)
        for (auto* t : net.layers_[li]->d.in) { consumers[t->id]++; consumer_layers[t->id].push_back(li); }
    auto soleConsumer = [&](TensorImpl* t) -> int {   // layer index, or -1
        if (t->is_output || consumers[t->id] != 1) return -1;
        return consumer_layers[t->id][0];
    };
    auto opOf = [&](int li) -> const OpInfo* {
        const LayerData& d = net.layers_[li]->d;
        if (d.kind != LKind::kPlugin) return nullptr;
        auto* op = dynamic_cast<IRedtailOp*>(d.plugin);
        return op ? &op->opInfo() : nullptr;
    };
    // An ELU plugin declared kHALF (the fp16 builders, resnet18_2D_513x257_net.cpp with data_type = kHALF) fuses as well:
    // the epilogue evaluates it in fp32, which is at least as accurate as the fp16 tensor the plugin would round to.
    auto isFp32Elu = [&](int li) {
        const OpInfo* o = li >= 0 ? opOf(li) : nullptr;
        return o && o->kind == OpKind::kElu && (o->data_type == DataType::kFLOAT || o->data_type == DataType::kHALF);
    };

    std::vector<bool> done(nl, false);
    std::map<int, PadInfo> absorbed_pad;      // conv layer index -> the PaddingPlugin folded into it
    for (int li = 0; li < nl; ++li) {
        if (done[li]) continue;
        LayerData& d = net.layers_[li]->d;
        done[li] = true;
        Step st;
        st.name = d.name;
        for (auto* t : d.in) st.in.push_back(t->id);
        switch (d.kind) {
            case LKind::kConv:
            case LKind::kDeconv: {
                rt_conv2d_desc cd{};
                cd.transposed = d.kind == LKind::kDeconv;
                cd.cin = d.in[0]->dims.d[0]; cd.cout = d.nb_out_maps;
                cd.r = d.ksize.h(); cd.s = d.ksize.w();
                cd.stride[0] = d.stride.h(); cd.stride[1] = d.stride.w();
                cd.pad[0] = d.pad.h(); cd.pad[1] = d.pad.w();
                cd.in_h = d.in[0]->dims.d[1]; cd.in_w = d.in[0]->dims.d[2];
                cd.weights_dtype = rtType(d.kw.type); cd.weights = d.kw.values;
                cd.bias = d.bw.count > 0 ? d.bw.values : nullptr;
                TensorImpl* out = d.out[0];
                int nxt = fusion ? soleConsumer(out) : -1;
                int im2col_in = -1;                  // >= 0: the conv reads this engine-made im2col matrix instead of its input
                // A 2-D (transposed) convolution is a 3-D one with V = D = 1 ([K,1,H,W] == [K,H,W]): the tower convs, the
                // ResNet18_2D encoder / decoder and its deconvolutions run on the same tcgen05 implicit-GEMM kernel as the 3-D
                // stack, with its epilogue fusions (+skip, ELU).  Channel counts that are not a K-block multiple (the 33-channel
                // concat of ResNet18_2D) are zero-padded by the kernel's re-layout pass.  Only the 5x5 first layer stays on the
                // CUDA-core kernel (filter larger than 3x3).
                const char* pe2 = getenv("REDTAIL_CONV3D_PRECISION");
                const bool simt_only = (pe2 && !strcmp(pe2, "simt")) || getenv("REDTAIL_SIMT_TOWERS") != nullptr;
                const bool tr2 = d.kind == LKind::kDeconv;
                const int ho2 = out->dims.d[1], wo2 = out->dims.d[2];
                rt_conv3d_desc c3{};
                c3.transposed = tr2 ? 1 : 0;
                if (!tr2) { c3.k = cd.cout; c3.c = cd.cin; } else { c3.k = cd.cin; c3.c = cd.cout; }   // deconv weights are [Cin,Cout,R,S]
                c3.v = 1; c3.r = cd.r; c3.s = cd.s;
                c3.stride[0] = 1; c3.stride[1] = cd.stride[0]; c3.stride[2] = cd.stride[1];
                c3.pad[0] = 0; c3.pad[1] = cd.pad[0]; c3.pad[2] = cd.pad[1];
                if (!tr2) {
                    c3.in_dims[0] = 1; c3.in_dims[1] = cd.cin; c3.in_dims[2] = cd.in_h; c3.in_dims[3] = cd.in_w;
                    c3.out_dims[0] = cd.cout; c3.out_dims[1] = 1; c3.out_dims[2] = ho2; c3.out_dims[3] = wo2;
                } else {
                    c3.in_dims[0] = cd.cin; c3.in_dims[1] = 1; c3.in_dims[2] = cd.in_h; c3.in_dims[3] = cd.in_w;
                    c3.out_dims[0] = 1; c3.out_dims[1] = cd.cout; c3.out_dims[2] = ho2; c3.out_dims[3] = wo2;
                }
                c3.weights_dtype = cd.weights_dtype; c3.weights = cd.weights; c3.bias = cd.bias;
                c3.precision = pe2 && !strcmp(pe2, "fp16") ? RT_PREC_FP16 : RT_PREC_FP32;
                const char* tw_env = getenv("REDTAIL_ENGINE_TOWER_SPLIT16");
                bool tc2 = fusion && !simt_only && !(tw_env && tw_env[0] == '0') && rt_conv3d_tc_supported(&c3) == 1;
                // A forward convolution with a filter larger than 3x3 and a GEMM-K worth the trip (TrailNet conv1: 7x7 stride 2 over 3
                // channels, K = 147; ros/packages/caffe_ros + TrailNet_SResNet-18.prototxt:31-53) runs as im2col + a 1x1 convolution
                // over K "channels" on the tcgen05 kernel: 30 % of the TrailNet step on the CUDA-core kernel otherwise.  The 5x5 first
                // tower layer of the stereo nets (K = 75, 0.06 ms) stays on CUDA cores.
                const int gemm_k = cd.cin * cd.r * cd.s;
                const char* s16_env = getenv("REDTAIL_ENGINE_SPLIT16");
                const char* i2c_env = getenv("REDTAIL_ENGINE_IM2COL");
                const int min_k = getenv("REDTAIL_ENGINE_IM2COL_MINK") ? atoi(getenv("REDTAIL_ENGINE_IM2COL_MINK")) : 96;
                if (!tc2 && !tr2 && fusion && !simt_only && (cd.r > 3 || cd.s > 3) && gemm_k > min_k && gemm_k <= 512 &&
                    !(s16_env && s16_env[0] == '0') && !(i2c_env && i2c_env[0] == '0') && cd.stride[0] == cd.stride[1] && cd.pad[0] == cd.pad[1]) {
                    const int kp = (gemm_k + 63) / 64 * 64;
                    // weights [cout][cin*r*s] -> [cout][kp], zero padded, fp32 (the engine owns the buffer)
                    host_blobs_.emplace_back(static_cast<size_t>(cd.cout) * kp, 0.f);
                    std::vector<float>& wb = host_blobs_.back();
                    for (int k = 0; k < cd.cout; ++k)
                        for (int j = 0; j < gemm_k; ++j) {
                            const size_t src = static_cast<size_t>(k) * gemm_k + j;
                            float v;
                            if (d.kw.type == DataType::kHALF) {
                                const uint16_t hb = static_cast<const uint16_t*>(d.kw.values)[src];
                                const uint32_t sign = (hb >> 15) & 1u, ex = (hb >> 10) & 31u, man = hb & 1023u;
                                v = ex == 0 ? ldexpf(static_cast<float>(man), -24) : ex == 31 ? (man ? NAN : INFINITY) : ldexpf(static_cast<float>(man | 1024u), static_cast<int>(ex) - 25);
                                if (sign) v = -v;
                            } else v = static_cast<const float*>(d.kw.values)[src];
                            wb[static_cast<size_t>(k) * kp + j] = v;
                        }
                    rt_conv3d_desc g{};
                    g.k = cd.cout; g.v = 1; g.c = kp; g.r = 1; g.s = 1;
                    for (int i = 0; i < 3; ++i) { g.stride[i] = 1; g.pad[i] = 0; }
                    g.in_dims[0] = 1; g.in_dims[1] = kp; g.in_dims[2] = ho2; g.in_dims[3] = wo2;
                    g.out_dims[0] = cd.cout; g.out_dims[1] = 1; g.out_dims[2] = ho2; g.out_dims[3] = wo2;
                    g.weights_dtype = RT_F32; g.weights = wb.data();
                    g.bias = cd.bias; g.precision = c3.precision;
                    if (cd.bias && cd.weights_dtype == RT_F16) {         // bias in the same dtype as the (now fp32) weights
                        host_blobs_.emplace_back(static_cast<size_t>(cd.cout), 0.f);
                        const uint16_t* hb = static_cast<const uint16_t*>(cd.bias);
                        for (int k = 0; k < cd.cout; ++k) {
                            const uint32_t sign = (hb[k] >> 15) & 1u, ex = (hb[k] >> 10) & 31u, man = hb[k] & 1023u;
                            float v = ex == 0 ? ldexpf(static_cast<float>(man), -24) : ex == 31 ? (man ? NAN : INFINITY) : ldexpf(static_cast<float>(man | 1024u), static_cast<int>(ex) - 25);
                            host_blobs_.back()[k] = sign ? -v : v;
                        }
                        g.bias = host_blobs_.back().data();
                    }
                    g.in_layout = RT_LAYOUT_SPLIT16;
                    if (rt_conv3d_tc_supported(&g) == 1) {
                        // the im2col matrix: an engine-made tensor that only exists in the split16 layout
                        TensorSlot ms;
                        ms.name = d.name + "_im2col";
                        ms.dims = DimsCHW(kp, ho2, wo2);
                        ms.elems = static_cast<size_t>(kp) * ho2 * wo2;
                        ms.forced_split = true;
                        const int mid = static_cast<int>(slots_.size());
                        slots_.push_back(ms);
                        Step im;
                        im.name = d.name + "_im2col";
                        im.in.push_back(d.in[0]->id);
                        im.out.push_back(mid);
                        const int xin = d.in[0]->id, cin_ = cd.cin, ih = cd.in_h, iw = cd.in_w, fr = cd.r, fs = cd.s, fst = cd.stride[0], fpd = cd.pad[0];
                        im.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                            return rt_im2col_split16(ptr(xin), ptr(mid), batch, cin_, ih, iw, fr, fs, fst, fpd, ho2, wo2, kp, s);
                        };
                        steps_.push_back(std::move(im));
                        g.in_layout = RT_LAYOUT_DENSE;        // decided (to split16) by the layout pass
                        c3 = g;
                        st.in.clear();
                        st.in.push_back(mid);
                        im2col_in = mid;
                        tc2 = true;
                    }
                }
                int skip2 = -1;
                const char* act_env = getenv("REDTAIL_ENGINE_ACT_FUSION");
                const bool act_fusion = !(act_env && act_env[0] == '0');
                if (tc2 && (tr2 || act_fusion) && nxt >= 0 && net.layers_[nxt]->d.kind == LKind::kEltwise) {
                    // deconv -> + skip (kSUM) [-> ELU]: the ResNet18_2D decoder (resnet18_2D_513x257_net.cpp:700-760);
                    // conv -> + shortcut (kSUM) [-> S-ReLU]: the residual blocks of TrailNet (TrailNet_SResNet-18.prototxt:217-278)
                    LayerData& e = net.layers_[nxt]->d;
                    TensorImpl* other = e.in[0] == out ? e.in[1] : e.in[0];
                    if (other != out && (other->is_input || (other->producer >= 0 && other->producer < li))) {
                        skip2 = other->id;
                        done[nxt] = true; out = e.out[0]; st.name += " + " + e.name;
                        nxt = soleConsumer(out);
                    }
                }
                if (isFp32Elu(nxt)) {
                    cd.fuse_elu = 1;
                    done[nxt] = true;
                    out = net.layers_[nxt]->d.out[0];
                    st.name += " + " + net.layers_[nxt]->d.name;
                }
                // conv [+ shortcut] -> Scale(per channel) -> ReLU -> Scale(per channel): the S-ReLU of TrailNet in the epilogue
                if (tc2 && !tr2 && act_fusion && !cd.fuse_elu && nxt >= 0 && net.layers_[nxt]->d.kind == LKind::kScale &&
                    net.layers_[nxt]->d.smode == ScaleMode::kCHANNEL && net.layers_[nxt]->d.power.count == 0) {
                    const int n0 = nxt, n1 = soleConsumer(net.layers_[n0]->d.out[0]);
                    const int n2 = n1 >= 0 && net.layers_[n1]->d.kind == LKind::kActivation && net.layers_[n1]->d.act == ActivationType::kRELU
                                       ? soleConsumer(net.layers_[n1]->d.out[0]) : -1;
                    if (n2 >= 0 && net.layers_[n2]->d.kind == LKind::kScale && net.layers_[n2]->d.smode == ScaleMode::kCHANNEL &&
                        net.layers_[n2]->d.power.count == 0) {
                        const int co = cd.cout;
                        host_blobs_.emplace_back(static_cast<size_t>(4) * co, 0.f);
                        std::vector<float>& ap = host_blobs_.back();
                        auto fill = [&](int row, const Weights& w, float dflt) {
                            for (int k = 0; k < co; ++k) {
                                float v = dflt;
                                if (w.count == co && w.values) {
                                    if (w.type == DataType::kHALF) {
                                        const uint16_t hb = static_cast<const uint16_t*>(w.values)[k];
                                        const uint32_t sign = (hb >> 15) & 1u, ex = (hb >> 10) & 31u, man = hb & 1023u;
                                        v = ex == 0 ? ldexpf(static_cast<float>(man), -24) : ex == 31 ? (man ? NAN : INFINITY) : ldexpf(static_cast<float>(man | 1024u), static_cast<int>(ex) - 25);
                                        if (sign) v = -v;
                                    } else v = static_cast<const float*>(w.values)[k];
                                }
                                ap[static_cast<size_t>(row) * co + k] = v;
                            }
                        };
                        fill(0, net.layers_[n0]->d.scale, 1.f); fill(1, net.layers_[n0]->d.shift, 0.f);
                        fill(2, net.layers_[n2]->d.scale, 1.f); fill(3, net.layers_[n2]->d.shift, 0.f);
                        c3.act_params = ap.data();
                        done[n0] = done[n1] = done[n2] = true;
                        out = net.layers_[n2]->d.out[0];
                        st.name += " + " + net.layers_[n0]->d.name + " + " + net.layers_[n1]->d.name + " + " + net.layers_[n2]->d.name;
                    }
                }
                if (tc2) {
                    c3.fuse_elu = cd.fuse_elu;
                    // Deferred like the 3-D layers: the layout pass may keep the activations between consecutive convolutions
                    // in RT_LAYOUT_SPLIT16 (no re-layout pass in front of the next conv).
                    conv_steps_.emplace_back(new ConvStep());
                    ConvStep* cs = conv_steps_.back().get();
                    cs->desc = c3; cs->in_id = im2col_in >= 0 ? im2col_in : d.in[0]->id; cs->out_id = out->id; cs->skip_id = skip2; cs->name = d.name;
                    if (skip2 >= 0) st.in.push_back(skip2);
                    st.out.push_back(out->id);
                    st.conv = cs;
                    st.run = [cs](int batch, const std::function<void*(int)>& ptr, void* ws, cudaStream_t s) {
                        return rt_conv3d_enqueue(cs->plan, batch * cs->batch_mul, ptr(cs->in_id), cs->skip_id >= 0 ? ptr(cs->skip_id) : nullptr,
                                                 ptr(cs->out_id), ws, s);
                    };
                    break;
                }
                rt_conv2d_plan* plan = nullptr;
                const int rc = rt_conv2d_create(&cd, &plan);
                if (rc != RT_OK) return fail(d.name + ": rt_conv2d_create failed (" + std::to_string(rc) + ")");
                conv2d_plans_.push_back(plan);
                const int in_id = d.in[0]->id, out_id = out->id;
                st.out.push_back(out_id);
                st.run = [plan, in_id, out_id](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                    return rt_conv2d_enqueue(plan, batch, ptr(in_id), ptr(out_id), s);
                };
                break;
            }
            case LKind::kScale: {
                if (d.smode == ScaleMode::kCHANNEL) {
                    // Caffe Scale layer with per-channel blobs (TrailNet_SResNet-18.prototxt: sub_mean and the S-ReLU chains).
                    const int c = d.in[0]->dims.d[0];
                    const int64_t hw = static_cast<int64_t>(d.in[0]->dims.d[1]) * d.in[0]->dims.d[2];
                    if (d.power.count > 0) {
                        std::vector<float> pw(static_cast<size_t>(c), 1.f);
                        if (d.power.type == DataType::kFLOAT) memcpy(pw.data(), d.power.values, sizeof(float) * c);
                        for (float v : pw) if (v != 1.f) return fail(d.name + ": per-channel scale with power != 1 is not supported");
                    }
                    const int in_id = d.in[0]->id;
                    float* s1 = d.scale.count ? deviceArray(d.scale, c, 1.f) : nullptr;
                    float* b1 = d.shift.count ? deviceArray(d.shift, c, 0.f) : nullptr;
                    if ((d.scale.count && !s1) || (d.shift.count && !b1)) return fail(d.name + ": device allocation failed");
                    // S-ReLU = Scale -> ReLU -> Scale (prototxt:54-105): one pass instead of three.
                    const int n1 = fusion ? soleConsumer(d.out[0]) : -1;
                    if (n1 >= 0 && net.layers_[n1]->d.kind == LKind::kActivation && net.layers_[n1]->d.act == ActivationType::kRELU) {
                        const int n2 = soleConsumer(net.layers_[n1]->d.out[0]);
                        if (n2 >= 0 && net.layers_[n2]->d.kind == LKind::kScale && net.layers_[n2]->d.smode == ScaleMode::kCHANNEL &&
                            net.layers_[n2]->d.power.count == 0) {
                            LayerData& d2 = net.layers_[n2]->d;
                            float* s2 = d2.scale.count ? deviceArray(d2.scale, c, 1.f) : nullptr;
                            float* b2 = d2.shift.count ? deviceArray(d2.shift, c, 0.f) : nullptr;
                            if ((d2.scale.count && !s2) || (d2.shift.count && !b2)) return fail(d2.name + ": device allocation failed");
                            done[n1] = done[n2] = true;
                            st.name += " + " + net.layers_[n1]->d.name + " + " + d2.name;
                            const int out_id = d2.out[0]->id;
                            st.out.push_back(out_id);
                            st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                                return rt_srelu(ptr(in_id), ptr(out_id), batch, c, hw, s1, b1, s2, b2, s);
                            };
                            break;
                        }
                    }
                    const int out_id = d.out[0]->id;
                    st.out.push_back(out_id);
                    st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                        return rt_scale_channel(ptr(in_id), ptr(out_id), batch, c, hw, s1, b1, s);
                    };
                    break;
                }
                const float shift = weightScalar(d.shift, 0.f), scale = weightScalar(d.scale, 1.f), power = weightScalar(d.power, 1.f);
                const int in_id = d.in[0]->id, out_id = d.out[0]->id;
                // (x * 1 + 0)^1: the generator emits this identity in front of every tower (tensorrt_model_builder.py:134-136,
                // nvsmall_1025x321_net.cpp:36-45).  No pass: the output is the input.
                if (fusion && shift == 0.f && scale == 1.f && power == 1.f && slots_[out_id].binding < 0) {
                    slots_[out_id].alias_of = in_id;
                    continue;
                }
                const int64_t elems = static_cast<int64_t>(slots_[in_id].elems);
                st.out.push_back(out_id);
                st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                    return rt_scale(RT_F32, ptr(in_id), ptr(out_id), elems * batch, shift, scale, power, s);
                };
                break;
            }
            case LKind::kPooling: {
                const int in_id = d.in[0]->id, out_id = d.out[0]->id;
                const int c = d.in[0]->dims.d[0], h = d.in[0]->dims.d[1], w = d.in[0]->dims.d[2];
                const int oh = d.pool_oh, ow = d.pool_ow, k = d.ksize.h(), stp = d.stride.h(), pd = d.pad.h();
                const int is_max = d.pool == PoolingType::kMAX ? 1 : 0;
                st.out.push_back(out_id);
                st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                    return rt_pool2d(ptr(in_id), ptr(out_id), batch, c, h, w, oh, ow, k, stp, pd, is_max, s);
                };
                break;
            }
            case LKind::kFullyConnected: {
                const int in_id = d.in[0]->id, out_id = d.out[0]->id;
                const int k = static_cast<int>(volume(d.in[0]->dims)), m = d.nb_out_maps;
                float* wd = deviceArray(d.kw, static_cast<int64_t>(k) * m, 0.f);
                float* bd = d.bw.count ? deviceArray(d.bw, m, 0.f) : nullptr;
                if (!wd || (d.bw.count && !bd)) return fail(d.name + ": device allocation failed");
                st.out.push_back(out_id);
                st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                    return rt_fully_connected(ptr(in_id), wd, bd, ptr(out_id), batch, k, m, s);
                };
                break;
            }
            case LKind::kSoftMax: {
                const int in_id = d.in[0]->id, out_id = d.out[0]->id;
                const int c = d.in[0]->dims.d[0];
                const int64_t inner = static_cast<int64_t>(d.in[0]->dims.d[1]) * d.in[0]->dims.d[2];
                st.out.push_back(out_id);
                st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                    return rt_softmax_channels(ptr(in_id), ptr(out_id), batch, c, inner, s);
                };
                break;
            }
            case LKind::kActivation: {
                if (d.act == ActivationType::kRELU) {
                    const int in_id = d.in[0]->id, out_id = d.out[0]->id;
                    const int64_t elems = static_cast<int64_t>(slots_[in_id].elems);
                    st.out.push_back(out_id);
                    st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                        return rt_relu(ptr(in_id), ptr(out_id), elems * batch, s);
                    };
                    break;
                }
                if (d.act != ActivationType::kSIGMOID) return fail(d.name + ": only ActivationType::kSIGMOID / kRELU are supported");
                const int in_id = d.in[0]->id, out_id = d.out[0]->id;
                const int64_t elems = static_cast<int64_t>(slots_[in_id].elems);
                st.out.push_back(out_id);
                st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                    return rt_sigmoid(RT_F32, ptr(in_id), ptr(out_id), elems * batch, s);
                };
                break;
            }
            case LKind::kEltwise: {
                const int a = d.in[0]->id, b = d.in[1]->id, out_id = d.out[0]->id;
                const int64_t elems = static_cast<int64_t>(slots_[a].elems);
                st.out.push_back(out_id);
                st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                    return rt_eltwise_sum(RT_F32, ptr(a), ptr(b), ptr(out_id), elems * batch, s);
                };
                break;
            }
            case LKind::kConcat: {
                if (d.in.size() != 2) return fail(d.name + ": concatenation of exactly two tensors is supported");
                const int a = d.in[0]->id, b = d.in[1]->id, out_id = d.out[0]->id;
                const int ca = d.in[0]->dims.d[0], cb = d.in[1]->dims.d[0];
                const int64_t inner = static_cast<int64_t>(d.in[0]->dims.d[1]) * d.in[0]->dims.d[2];
                st.out.push_back(out_id);
                st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                    return rt_concat_channels(RT_F32, ptr(a), ca, ptr(b), cb, ptr(out_id), batch, inner, s);
                };
                break;
            }
            case LKind::kShuffle: {
                const int in_id = d.in[0]->id, out_id = d.out[0]->id;
                if (slots_[out_id].binding < 0) {          // pure reshape: alias, no step
                    slots_[out_id].alias_of = in_id;
                    continue;
                }
                const int64_t elems = static_cast<int64_t>(slots_[in_id].elems);
                st.out.push_back(out_id);
                st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                    return static_cast<int>(cudaMemcpyAsync(ptr(out_id), ptr(in_id), elems * batch * sizeof(float), cudaMemcpyDeviceToDevice, s));
                };
                break;
            }
            case LKind::kPlugin: {
                const OpInfo* op = opOf(li);
                // ---- fused 3-D convolution chains --------------------------------------------------------------
                if (fusion && op && (op->kind == OpKind::kConv3D || op->kind == OpKind::kConv3DTranspose)) {
                    rt_conv3d_desc cd{};
                    const bool tr = op->kind == OpKind::kConv3DTranspose;
                    const Dims& k = op->kernel_dims;
                    cd.transposed = tr;
                    cd.k = k.d[0]; cd.v = k.d[1]; cd.c = k.d[2]; cd.r = k.d[3]; cd.s = k.d[4];
                    for (int i = 0; i < 3; ++i) { cd.stride[i] = op->stride.d[i]; cd.pad[i] = op->pad_start.d[i]; }
                    for (int i = 0; i < 4; ++i) { cd.in_dims[i] = d.in[0]->dims.d[i]; cd.out_dims[i] = d.out[0]->dims.d[i]; }
                    cd.weights_dtype = rtType(op->kernel.type); cd.weights = op->kernel.values;
                    cd.bias = op->bias.count > 0 ? op->bias.values : nullptr;
                    const char* pe = getenv("REDTAIL_CONV3D_PRECISION");
                    cd.precision = pe && !strcmp(pe, "fp16") ? RT_PREC_FP16 : pe && !strcmp(pe, "simt") ? RT_PREC_SIMT : RT_PREC_FP32;
                    TensorImpl* out = d.out[0];
                    int skip_id = -1;
                    int nxt = soleConsumer(out);
                    const OpInfo* no = nxt >= 0 ? opOf(nxt) : nullptr;
                    if (!tr) {
                        if (no && no->kind == OpKind::kTransform) {
                            cd.out_transposed = 1;
                            done[nxt] = true; out = net.layers_[nxt]->d.out[0]; st.name += " + " + net.layers_[nxt]->d.name;
                            nxt = soleConsumer(out);
                        }
                    } else {
                        if (no && no->kind == OpKind::kSlice && no->slice_start == 0) {
                            cd.slice_d = cd.out_dims[0] - no->slice_end;
                            done[nxt] = true; out = net.layers_[nxt]->d.out[0]; st.name += " + " + net.layers_[nxt]->d.name;
                            nxt = soleConsumer(out);
                        }
                        if (nxt >= 0 && net.layers_[nxt]->d.kind == LKind::kEltwise) {
                            LayerData& e = net.layers_[nxt]->d;
                            TensorImpl* other = e.in[0] == out ? e.in[1] : e.in[0];
                            // The add moves up to this layer's position, so the skip tensor must exist by then: a network
                            // input, or produced by an EARLIER layer (topological order alone only puts it before the add).
                            if (other != out && (other->is_input || (other->producer >= 0 && other->producer < li))) {
                                skip_id = other->id;
                                done[nxt] = true; out = e.out[0]; st.name += " + " + e.name;
                                nxt = soleConsumer(out);
                            }
                        }
                    }
                    if (isFp32Elu(nxt)) {
                        cd.fuse_elu = 1;
                        done[nxt] = true; out = net.layers_[nxt]->d.out[0]; st.name += " + " + net.layers_[nxt]->d.name;
                    }
                    // A PaddingPlugin in front of this conv was absorbed: read the un-padded tensor, the zero planes are
                    // virtual (TMA out-of-bounds fill on the tensor-core path).
                    int in_id = d.in[0]->id;
                    auto pit = absorbed_pad.find(li);
                    if (pit != absorbed_pad.end()) {
                        in_id = pit->second.in_id;
                        cd.pad_end_d = pit->second.planes;
                        cd.in_dims[0] -= cd.pad_end_d;
                        st.in.clear();
                        st.in.push_back(in_id);
                        st.name = net.layers_[pit->second.layer]->d.name + " + " + st.name;
                    }
                    // The plan is created after the layout pass below (it depends on the tensor layouts).
                    conv_steps_.emplace_back(new ConvStep());
                    ConvStep* cs = conv_steps_.back().get();
                    cs->desc = cd; cs->in_id = in_id; cs->out_id = out->id; cs->skip_id = skip_id; cs->name = d.name;
                    if (skip_id >= 0) st.in.push_back(skip_id);
                    st.out.push_back(out->id);
                    st.conv = cs;
                    st.run = [cs](int batch, const std::function<void*(int)>& ptr, void* ws, cudaStream_t s) {
                        return rt_conv3d_enqueue(cs->plan, batch * cs->batch_mul, ptr(cs->in_id), cs->skip_id >= 0 ? ptr(cs->skip_id) : nullptr,
                                                 ptr(cs->out_id), ws, s);
                    };
                    break;
                }
                // ---- Padding in front of a 3-D convolution: absorbed into the conv (virtual zero planes) --------------
                if (fusion && op && op->kind == OpKind::kPadding) {
                    const int nxt = soleConsumer(d.out[0]);
                    const OpInfo* no = nxt >= 0 ? opOf(nxt) : nullptr;
                    if (no && no->kind == OpKind::kConv3D) {
                        absorbed_pad[nxt] = PadInfo{d.in[0]->id, op->pad_end_planes, li};
                        continue;                                  // no step of its own
                    }
                }
                // ---- cost volume / Transform: remember them, the layout pass may rewrite or drop them ------------------
                if (fusion && op && op->kind == OpKind::kCostVolume && op->cv_type == redtail::tensorrt::CostVolumeType::kDefault) {
                    st.costvol_c = d.in[0]->dims.d[0];
                    st.costvol_h = d.in[0]->dims.d[1];
                    st.costvol_w = d.in[0]->dims.d[2];
                    st.costvol_d = op->max_disparity;
                }
                if (fusion && op && op->kind == OpKind::kTransform) st.is_transform = true;
                if (fusion && op && op->kind == OpKind::kSoftargmax) st.softargmax = op->sm_type == redtail::tensorrt::SoftargmaxType::kMin ? 1 : 2;
                // ---- redtail plugins declared kHALF (the fp16 builders: resnet18_2D_513x257_net.cpp with data_type = kHALF) ----
                // The graph's tensors are fp32; instead of wrapping the plugin in fp32<->fp16 reformat passes (what TensorRT
                // does, and what rounds every activation to fp16) the engine runs the same kernel on the fp32 tensors: one
                // pass instead of three, and the fp16 configuration keeps fp32 activations (only its WEIGHTS are fp16).
                if (fusion && op && op->data_type == DataType::kHALF &&
                    (op->kind == OpKind::kElu || op->kind == OpKind::kCostVolume || op->kind == OpKind::kSoftargmax)) {
                    const int out_id = d.out[0]->id;
                    st.out.push_back(out_id);
                    if (op->kind == OpKind::kElu) {
                        const int in_id = d.in[0]->id;
                        const int64_t elems = static_cast<int64_t>(slots_[in_id].elems);
                        st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                            return rt_elu(RT_F32, ptr(in_id), ptr(out_id), elems * batch, s);
                        };
                    } else if (op->kind == OpKind::kCostVolume) {
                        const int l = d.in[0]->id, r = d.in[1]->id;
                        const int c = d.in[0]->dims.d[0], h = d.in[0]->dims.d[1], w = d.in[0]->dims.d[2], dd = op->max_disparity;
                        const bool corr = op->cv_type == redtail::tensorrt::CostVolumeType::kCorrelation;
                        st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                            return corr ? rt_corr_cost_volume(RT_F32, ptr(l), ptr(r), ptr(out_id), batch, c, h, w, dd, s)
                                        : rt_cost_volume(RT_F32, ptr(l), ptr(r), ptr(out_id), batch, c, h, w, dd, s);
                        };
                    } else {
                        const int in_id = d.in[0]->id;
                        const Dims& id = d.in[0]->dims;
                        const int dd = id.d[0];
                        const int64_t hw = static_cast<int64_t>(slots_[in_id].elems) / dd;
                        const int is_min = op->sm_type == redtail::tensorrt::SoftargmaxType::kMin ? 1 : 0;
                        st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                            return rt_softargmax(RT_F32, is_min, ptr(in_id), ptr(out_id), batch, dd, hw, s);
                        };
                    }
                    break;
                }
                // ---- generic plugin protocol ---------------------------------------------------------------------
                std::vector<Dims> ind, outd;
                for (auto* t : d.in) ind.push_back(t->dims);
                for (auto* t : d.out) outd.push_back(t->dims);
                bool half_io = false;
                if (d.plugin_ext) {
                    DataType dt;
                    if (half2 && d.plugin_ext->supportsFormat(DataType::kHALF, PluginFormat::kNCHW)) dt = DataType::kHALF;
                    else if (d.plugin_ext->supportsFormat(DataType::kFLOAT, PluginFormat::kNCHW)) dt = DataType::kFLOAT;
                    else if (d.plugin_ext->supportsFormat(DataType::kHALF, PluginFormat::kNCHW)) dt = DataType::kHALF;
                    else return fail(d.name + ": plugin supports neither fp32 nor fp16 linear NCHW");
                    half_io = dt == DataType::kHALF;
                    d.plugin_ext->configureWithFormat(ind.data(), static_cast<int>(ind.size()), outd.data(),
                                                      static_cast<int>(outd.size()), dt, PluginFormat::kNCHW, max_batch_);
                } else {
                    d.plugin->configure(ind.data(), static_cast<int>(ind.size()), outd.data(), static_cast<int>(outd.size()), max_batch_);
                }
                if (d.plugin->initialize() != 0) return fail(d.name + ": plugin initialize() failed");
                if (op == nullptr) graph_safe_ = false;
                configured_plugins_.push_back(d.plugin);
                IPlugin* plugin = d.plugin;
                std::vector<int> in_ids, out_ids;
                std::vector<size_t> in_elems, out_elems;
                for (auto* t : d.in) { in_ids.push_back(t->id); in_elems.push_back(slots_[t->id].elems); }
                for (auto* t : d.out) { out_ids.push_back(t->id); out_elems.push_back(slots_[t->id].elems); st.out.push_back(t->id); }
                const size_t plug_ws = plugin->getWorkspaceSize(max_batch_);
                auto align = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
                size_t conv_bytes = 0;
                if (half_io) {
                    for (size_t e : in_elems) conv_bytes += align(e * max_batch_ * 2);
                    for (size_t e : out_elems) conv_bytes += align(e * max_batch_ * 2);
                }
                st.workspace = conv_bytes + plug_ws;
                const int mb = max_batch_;
                st.run = [=](int batch, const std::function<void*(int)>& ptr, void* ws, cudaStream_t s) -> int {
                    std::vector<const void*> in(in_ids.size());
                    std::vector<void*> out(out_ids.size());
                    char* w = static_cast<char*>(ws);
                    if (!half_io) {
                        for (size_t i = 0; i < in_ids.size(); ++i) in[i] = ptr(in_ids[i]);
                        for (size_t i = 0; i < out_ids.size(); ++i) out[i] = ptr(out_ids[i]);
                        return plugin->enqueue(batch, in.data(), out.data(), w, s);
                    }
                    // fp16 plugin inside an fp32 graph: reformat in, run, reformat out (what TensorRT's reformat layers do).
                    for (size_t i = 0; i < in_ids.size(); ++i) {
                        int rc = rt_convert(RT_F32, ptr(in_ids[i]), RT_F16, w, static_cast<int64_t>(in_elems[i]) * batch, s);
                        if (rc) return rc;
                        in[i] = w;
                        w += (in_elems[i] * mb * 2 + 255) & ~static_cast<size_t>(255);
                    }
                    for (size_t i = 0; i < out_ids.size(); ++i) {
                        out[i] = w;
                        w += (out_elems[i] * mb * 2 + 255) & ~static_cast<size_t>(255);
                    }
                    int rc = plugin->enqueue(batch, in.data(), out.data(), w, s);
                    if (rc) return rc;
                    for (size_t i = 0; i < out_ids.size(); ++i) {
                        rc = rt_convert(RT_F16, out[i], RT_F32, ptr(out_ids[i]), static_cast<int64_t>(out_elems[i]) * batch, s);
                        if (rc) return rc;
                    }
                    return 0;
                };
                break;
            }
        }
        steps_.push_back(std::move(st));
    }
    if (!assignLayoutsAndCreatePlans(fusion)) return false;
    if (fusion) pairSiameseSteps();
    if (!planMemory()) return false;
    plan_ = serializeNetwork(net, max_batch_, half2);
    return true;
}

std::string EngineImpl::serializeNetwork(NetworkImpl& net, int max_batch, bool half2)
{
    PlanWriter w;
    w.buf.append(kPlanMagic, 8);
    w.put<int32_t>(kPlanVersion);
    w.put<int32_t>(max_batch);
    w.put<uint8_t>(half2 ? 1 : 0);
    w.put<int32_t>(static_cast<int32_t>(net.inputs_.size()));
    for (auto* t : net.inputs_) { w.put<int32_t>(t->id); w.str(t->name); w.put<int32_t>(static_cast<int32_t>(t->type)); w.dims(t->dims); }
    w.put<int32_t>(static_cast<int32_t>(net.layers_.size()));
    for (auto& ln : net.layers_) {
        const LayerData& d = ln->d;
        w.put<int32_t>(static_cast<int32_t>(d.kind));
        w.str(d.name);
        w.put<int32_t>(static_cast<int32_t>(d.in.size()));
        for (auto* t : d.in) w.put<int32_t>(t->id);
        w.put<int32_t>(static_cast<int32_t>(d.out.size()));
        for (auto* t : d.out) { w.put<int32_t>(t->id); w.str(t->name); }
        switch (d.kind) {
            case LKind::kConv:
            case LKind::kDeconv:
                w.put<int32_t>(d.nb_out_maps);
                w.put<int32_t>(d.ksize.h()); w.put<int32_t>(d.ksize.w());
                w.put<int32_t>(d.stride.h()); w.put<int32_t>(d.stride.w());
                w.put<int32_t>(d.pad.h()); w.put<int32_t>(d.pad.w());
                w.weights(d.kw); w.weights(d.bw);
                break;
            case LKind::kScale:
                w.put<int32_t>(static_cast<int32_t>(d.smode));
                w.weights(d.shift); w.weights(d.scale); w.weights(d.power);
                break;
            case LKind::kActivation: w.put<int32_t>(static_cast<int32_t>(d.act)); break;
            case LKind::kEltwise: w.put<int32_t>(static_cast<int32_t>(d.eop)); break;
            case LKind::kConcat: break;
            case LKind::kShuffle: w.put<uint8_t>(d.has_reshape ? 1 : 0); w.dims(d.reshape); break;
            case LKind::kPooling:
                w.put<int32_t>(static_cast<int32_t>(d.pool));
                w.put<int32_t>(d.ksize.h()); w.put<int32_t>(d.stride.h()); w.put<int32_t>(d.pad.h());
                w.put<int32_t>(d.pool_oh); w.put<int32_t>(d.pool_ow);
                break;
            case LKind::kFullyConnected: w.put<int32_t>(d.nb_out_maps); w.weights(d.kw); w.weights(d.bw); break;
            case LKind::kSoftMax: break;
            case LKind::kPlugin: {
                const size_t n = d.plugin->getSerializationSize();
                if (n == 0) return std::string();          // a third-party plugin without serialisation: no plan
                w.put<int64_t>(static_cast<int64_t>(n));
                w.align8();
                const size_t at = w.buf.size();
                w.buf.resize(at + n);
                d.plugin->serialize(&w.buf[at]);
                break;
            }
        }
    }
    w.put<int32_t>(static_cast<int32_t>(net.outputs_.size()));
    for (auto* t : net.outputs_) w.put<int32_t>(t->id);
    return w.buf;
}

// Decides which activation tensors live in RT_LAYOUT_SPLIT16 (channels-last fp16 hi/lo, what the tcgen05 conv kernel
// consumes and can produce) instead of dense fp32, then creates the convolution plans.
// A tensor is split16 iff it is produced by a tensor-core conv step or by the concat cost volume, and every consumer is
// a tensor-core conv step (as input, or as skip of a conv whose output is split16 too) -- possibly through
// Transform{1,0,2,3} steps, which are no-ops on a channels-last tensor and are dropped.
bool EngineImpl::assignLayoutsAndCreatePlans(bool fusion)
{
    const int ns = static_cast<int>(steps_.size());
    const int nt = static_cast<int>(slots_.size());
    // Debug aid: REDTAIL_SIMT_LAYERS="conv3D_2,deconv3D_3" runs the named fused layers on the exact-fp32 CUDA-core kernels
    // (used to attribute the disparity error to individual layers).
    if (const char* sl = getenv("REDTAIL_SIMT_LAYERS")) {
        const std::string list = std::string(",") + sl + ",";
        for (auto& st : steps_)
            if (st.conv && list.find("," + st.conv->name + ",") != std::string::npos) st.conv->desc.precision = RT_PREC_SIMT;
    }
    // CostVolume(kDefault) whose only consumer is a 3x3x3 / stride-1 / pad-1 Conv3D: the pair collapses into two 2-D
    // convolutions of the feature maps + one combine pass (costvol_conv3d.cu); the volume is never built.
    const char* cv_env = getenv("REDTAIL_ENGINE_CVCONV");
    if (fusion && !(cv_env && cv_env[0] == '0')) {
        for (int si = 0; si < ns; ++si) {
            Step& cv = steps_[si];
            if (cv.costvol_d <= 0 || cv.out.size() != 1 || slots_[cv.out[0]].binding >= 0) continue;
            const int t = cv.out[0];
            int user = -1, nusers = 0;
            for (int sj = 0; sj < ns; ++sj)
                for (int id : steps_[sj].in)
                    if (id == t) { user = sj; ++nusers; }
            if (nusers != 1 || !steps_[user].conv) continue;
            ConvStep* cs = steps_[user].conv;
            const rt_conv3d_desc& d = cs->desc;
            if (cs->in_id != t || cs->skip_id >= 0 || d.transposed || d.pad_end_d != 0) continue;
            if (d.v != 3 || d.r != 3 || d.s != 3 || d.c != 2 * cv.costvol_c) continue;
            bool unit = true;
            for (int i = 0; i < 3; ++i) unit = unit && d.stride[i] == 1 && d.pad[i] == 1;
            if (!unit || d.in_dims[0] != cv.costvol_d || d.in_dims[2] != cv.costvol_h || d.in_dims[3] != cv.costvol_w) continue;
            rt_costvol_conv3d_desc f{};
            f.c = cv.costvol_c; f.h = cv.costvol_h; f.w = cv.costvol_w; f.max_disp = cv.costvol_d; f.k = d.k;
            f.weights_dtype = d.weights_dtype; f.weights = d.weights; f.bias = d.bias;
            f.precision = d.precision; f.fuse_elu = d.fuse_elu; f.out_transposed = d.out_transposed;
            f.out_layout = RT_LAYOUT_DENSE;
            if (!rt_costvol_conv3d_supported(&f)) continue;
            cs->cvfused = true; cs->cvdesc = f; cs->l_id = cv.in[0]; cs->r_id = cv.in[1]; cs->in_id = -1;
            steps_[user].in = cv.in;
            steps_[user].name = cv.name + " + " + steps_[user].name;
            cv.dropped = true;
        }
    }
    std::vector<int> producer(nt, -1);
    std::vector<std::vector<int>> consumers(nt);
    for (int si = 0; si < ns; ++si) {
        if (steps_[si].dropped) continue;
        for (int id : steps_[si].out) producer[id] = si;
        for (int id : steps_[si].in) consumers[id].push_back(si);
    }
    auto cvSplitOk = [&](const ConvStep* cs) {
        rt_costvol_conv3d_desc f = cs->cvdesc;
        f.out_layout = RT_LAYOUT_SPLIT16; f.out_transposed = 0;
        return rt_costvol_conv3d_supported(&f) == 1;
    };
    auto tcOk = [&](const ConvStep* cs, int in_layout, int out_layout) {
        rt_conv3d_desc d = cs->desc;
        d.in_layout = in_layout; d.out_layout = out_layout;
        return rt_conv3d_tc_supported(&d) == 1;
    };
    std::vector<char> split(nt, 0);
    const bool enable = fusion && getenv("REDTAIL_ENGINE_SPLIT16") == nullptr ? true : (fusion && getenv("REDTAIL_ENGINE_SPLIT16")[0] != '0');
    if (enable) {
        // optimistic start: every tensor produced by a capable step ...
        for (int t = 0; t < nt; ++t) {
            if (slots_[t].binding >= 0 || slots_[t].alias_of >= 0 || producer[t] < 0) continue;
            const Step& ps = steps_[producer[t]];
            const bool conv_prod = ps.conv && (ps.conv->cvfused ? cvSplitOk(ps.conv) : tcOk(ps.conv, RT_LAYOUT_DENSE, RT_LAYOUT_SPLIT16));
            const bool cv_prod = ps.costvol_d > 0 && ps.costvol_c % 8 == 0;
            const bool tr_prod = ps.is_transform;              // decided through its input below
            split[t] = (conv_prod || cv_prod || tr_prod || slots_[t].forced_split) ? 1 : 0;
        }
        // ... then demote until every constraint holds.
        for (bool changed = true; changed;) {
            changed = false;
            for (int t = 0; t < nt; ++t) {
                if (!split[t]) continue;
                bool ok = !consumers[t].empty();
                const Step& ps = steps_[producer[t]];
                if (ps.is_transform && !split[ps.in[0]]) ok = false;         // a transform only forwards a split16 tensor
                for (int si : consumers[t]) {
                    const Step& c = steps_[si];
                    if (c.is_transform) { ok = ok && split[c.out[0]]; continue; }
                    if (!c.conv || c.conv->cvfused) { ok = false; continue; }   // the fused pair reads dense feature maps
                    if (c.conv->in_id == t) ok = ok && tcOk(c.conv, RT_LAYOUT_SPLIT16, split[c.conv->out_id] ? RT_LAYOUT_SPLIT16 : RT_LAYOUT_DENSE);
                    if (c.conv->skip_id == t) ok = ok && split[c.conv->out_id];   // skip shares the output's layout
                }
                if (!ok) { split[t] = 0; changed = true; }
            }
            // a conv with a dense skip cannot write split16
            for (int si = 0; si < ns; ++si) {
                const ConvStep* cs = steps_[si].conv;
                if (cs && cs->skip_id >= 0 && split[cs->out_id] && !split[cs->skip_id]) { split[cs->out_id] = 0; changed = true; }
            }
        }
    }
    // Conv3DTranspose (-> Slice) with ONE output channel whose only consumer is the Softargmax plugin (the last two layers of
    // every stereo net, nvsmall_1025x321_net.cpp:398-420): one kernel, the [Dx,1,Hx,Wx] volume is never written
    // (deconv_softargmax.cu).  REDTAIL_ENGINE_DSA=0 keeps the two steps.
    const char* dsa_env = getenv("REDTAIL_ENGINE_DSA");
    if (fusion && !(dsa_env && dsa_env[0] == '0')) {
        for (int si = 0; si < ns; ++si) {
            Step& sm = steps_[si];
            if (!sm.softargmax || sm.dropped || sm.in.size() != 1 || sm.out.size() != 1) continue;
            const int t = sm.in[0];
            if (producer[t] < 0 || consumers[t].size() != 1 || slots_[t].binding >= 0) continue;
            Step& ps = steps_[producer[t]];
            ConvStep* cs = ps.conv;
            if (!cs || cs->cvfused || cs->out_id != t || cs->skip_id >= 0 || cs->batch_mul != 1 || !split[cs->in_id] || split[t]) continue;
            rt_conv3d_desc d = cs->desc;
            d.in_layout = RT_LAYOUT_SPLIT16; d.out_layout = RT_LAYOUT_DENSE; d.fuse_softargmax = sm.softargmax;
            if (rt_conv3d_tc_supported(&d) != 1) continue;
            cs->desc.fuse_softargmax = sm.softargmax;
            cs->out_id = sm.out[0];
            ps.out = sm.out;
            ps.name += " + " + sm.name;
            sm.dropped = true;
        }
    }
    for (int t = 0; t < nt; ++t)
        if (slots_[t].forced_split && !split[t]) return fail(slots_[t].name + ": the im2col matrix needs the split16 layout (REDTAIL_ENGINE_IM2COL=0 keeps the CUDA-core kernel)");
    int nsplit = 0;
    for (int si = 0; si < ns; ++si) {
        Step& st = steps_[si];
        if (st.dropped) continue;
        if (st.conv && st.conv->cvfused) {
            ConvStep* cs = st.conv;
            const bool sp = split[cs->out_id] != 0;
            cs->cvdesc.out_layout = sp ? RT_LAYOUT_SPLIT16 : RT_LAYOUT_DENSE;
            if (sp) cs->cvdesc.out_transposed = 0;
            nsplit += sp;
            const int rc = rt_costvol_conv3d_create(&cs->cvdesc, &cs->cvplan);
            if (rc != RT_OK) return fail(cs->name + ": rt_costvol_conv3d_create failed (" + std::to_string(rc) + ")");
            cvconv_plans_.push_back(cs->cvplan);
            st.workspace = rt_costvol_conv3d_workspace_size(cs->cvplan, max_batch_);
            st.run = [cs](int batch, const std::function<void*(int)>& ptr, void* ws, cudaStream_t s) {
                return rt_costvol_conv3d_enqueue(cs->cvplan, batch, ptr(cs->l_id), ptr(cs->r_id), ptr(cs->out_id), ws, s);
            };
            continue;
        }
        if (st.is_transform && split[st.in[0]] && split[st.out[0]]) {       // no-op on channels-last data
            slots_[st.out[0]].alias_of = st.in[0];
            st.dropped = true;
            continue;
        }
        if (st.costvol_d > 0 && split[st.out[0]]) {
            const int l = st.in[0], r = st.in[1], o = st.out[0];
            const int c = st.costvol_c, h = st.costvol_h, w = st.costvol_w, dd = st.costvol_d;
            st.workspace = 0;
            st.run = [=](int batch, const std::function<void*(int)>& ptr, void*, cudaStream_t s) {
                return rt_cost_volume_split16(ptr(l), ptr(r), ptr(o), batch, c, h, w, dd, s);
            };
        }
        if (st.conv) {
            ConvStep* cs = st.conv;
            cs->desc.in_layout = split[cs->in_id] ? RT_LAYOUT_SPLIT16 : RT_LAYOUT_DENSE;
            cs->desc.out_layout = split[cs->out_id] ? RT_LAYOUT_SPLIT16 : RT_LAYOUT_DENSE;
            nsplit += split[cs->out_id];
            int rc = rt_conv3d_create(&cs->desc, &cs->plan);
            if (rc == RT_ERR_UNSUPPORTED && cs->desc.precision != RT_PREC_SIMT) {
                logMsg(log_, ILogger::Severity::kWARNING, cs->name + ": shape not covered by the tcgen05 kernels, using the fp32 SIMT kernels.");
                cs->desc.precision = RT_PREC_SIMT;
                rc = rt_conv3d_create(&cs->desc, &cs->plan);
            }
            if (rc != RT_OK) return fail(cs->name + ": rt_conv3d_create failed (" + std::to_string(rc) + ")");
            conv3d_plans_.push_back(cs->plan);
            st.workspace = rt_conv3d_workspace_size(cs->plan, max_batch_);
        }
    }
    steps_.erase(std::remove_if(steps_.begin(), steps_.end(), [](const Step& s) { return s.dropped; }), steps_.end());
    logMsg(log_, ILogger::Severity::kINFO, "engine: " + std::to_string(nsplit) + " activation tensors kept in the split16 layout");
    return true;
}

// Siamese towers: the generated builders emit the left and the right tower as two chains of layers with bit-identical
// weights.  Two tensor-core conv steps that differ only in their tensors are run as ONE launch of
// batch 2N: their inputs (and outputs) are stored as a pair -- the twin's tensor right behind the first one's N samples.
void EngineImpl::pairSiameseSteps()
{
    if (const char* e = getenv("REDTAIL_ENGINE_SIAMESE")) if (e[0] == '0') return;
    auto sameDesc = [](const rt_conv3d_desc& a, const rt_conv3d_desc& b) {
        if (a.transposed != b.transposed || a.k != b.k || a.v != b.v || a.c != b.c || a.r != b.r || a.s != b.s) return false;
        for (int i = 0; i < 3; ++i) if (a.stride[i] != b.stride[i] || a.pad[i] != b.pad[i]) return false;
        for (int i = 0; i < 4; ++i) if (a.in_dims[i] != b.in_dims[i] || a.out_dims[i] != b.out_dims[i]) return false;
        if (a.weights_dtype != b.weights_dtype || a.precision != b.precision || a.fuse_elu != b.fuse_elu ||
            a.out_transposed != b.out_transposed || a.slice_d != b.slice_d || a.in_layout != b.in_layout ||
            a.out_layout != b.out_layout || a.pad_end_d != b.pad_end_d || (a.bias == nullptr) != (b.bias == nullptr) ||
            a.fuse_softargmax != b.fuse_softargmax || a.act_params != nullptr || b.act_params != nullptr)
            return false;
        const size_t es = a.weights_dtype == RT_F16 ? 2 : 4;
        const size_t wn = static_cast<size_t>(a.k) * a.v * a.c * a.r * a.s;
        if (a.weights == nullptr || b.weights == nullptr || memcmp(a.weights, b.weights, wn * es) != 0) return false;
        const size_t bn = a.transposed ? a.c : a.k;
        return a.bias == nullptr || memcmp(a.bias, b.bias, bn * es) == 0;
    };
    auto resolved = [&](int id) { while (slots_[id].alias_of >= 0) id = slots_[id].alias_of; return id; };
    auto pairable = [&](int a, int b) {        // may tensor b be stored behind tensor a?
        if (a == b) return false;
        const TensorSlot &x = slots_[a], &y = slots_[b];
        if (y.pair_of == a) return true;                                   // already are
        return x.binding < 0 && y.binding < 0 && x.pair_of < 0 && y.pair_of < 0 && !x.has_partner && !y.has_partner &&
               x.elems == y.elems;
    };
    int paired = 0;
    // The towers are emitted one after the other (all left layers, then all right layers), so the merged launch takes the
    // position of the LATER twin, where both inputs exist.  Walking the first tower backwards keeps that legal: when layer i
    // moves behind layer j, its consumer (layer i+1) has already moved behind layer j+1.
    for (size_t ii = steps_.size(); ii-- > 0;) {
        const size_t i = ii;
        ConvStep* a = steps_[i].conv;
        if (!a || a->cvfused || a->skip_id >= 0 || a->batch_mul != 1 || steps_[i].dropped || a->desc.precision == RT_PREC_SIMT) continue;
        for (size_t j = i + 1; j < steps_.size(); ++j) {
            ConvStep* b = steps_[j].conv;
            if (!b || b->cvfused || b->skip_id >= 0 || b->batch_mul != 1 || steps_[j].dropped) continue;
            if (!sameDesc(a->desc, b->desc)) continue;
            const int ai = resolved(a->in_id), bi = resolved(b->in_id), ao = resolved(a->out_id), bo = resolved(b->out_id);
            if (!pairable(ai, bi) || !pairable(ao, bo) || ai == bo || bi == ao) continue;
            // a's output now appears at step j: nothing up to and including j may read it
            bool legal = true;
            for (size_t k = 0; k <= j && legal; ++k) {
                if (steps_[k].dropped || k == i) continue;
                for (int id : steps_[k].in) if (resolved(id) == ao) legal = false;
            }
            if (!legal) continue;
            if (slots_[bi].pair_of < 0) { slots_[bi].pair_of = ai; slots_[ai].has_partner = true; }
            if (slots_[bo].pair_of < 0) { slots_[bo].pair_of = ao; slots_[ao].has_partner = true; }
            a->batch_mul = 2;
            Step& sj = steps_[j];
            sj.in.push_back(a->in_id);
            sj.out.push_back(a->out_id);
            sj.name = steps_[i].name + " || " + sj.name;
            sj.run = steps_[i].run;                  // launches a's plan over [a's N samples | b's N samples]
            sj.conv = a;
            sj.workspace = rt_conv3d_workspace_size(a->plan, max_batch_ * 2);
            steps_[i].dropped = true;
            ++paired;
            break;
        }
    }
    steps_.erase(std::remove_if(steps_.begin(), steps_.end(), [](const Step& s) { return s.dropped; }), steps_.end());
    if (paired) logMsg(log_, ILogger::Severity::kINFO, "engine: " + std::to_string(paired) + " siamese layer pairs run as one launch of twice the batch");
}

bool EngineImpl::planMemory()
{
    auto root = [&](int id) {
        while (slots_[id].alias_of >= 0 || slots_[id].pair_of >= 0) id = slots_[id].alias_of >= 0 ? slots_[id].alias_of : slots_[id].pair_of;
        return id;
    };
    // Liveness over steps, on alias roots.
    for (size_t si = 0; si < steps_.size(); ++si) {
        for (int id : steps_[si].in) {
            TensorSlot& s = slots_[root(id)];
            if (s.first < 0) s.first = static_cast<int>(si);
            s.last = static_cast<int>(si); s.used = true;
        }
        for (int id : steps_[si].out) {
            TensorSlot& s = slots_[root(id)];
            if (s.first < 0) s.first = static_cast<int>(si);
            s.last = std::max(s.last, static_cast<int>(si)); s.used = true;
        }
    }
    // A binding that aliases (through a reshape) to an arena tensor would need a copy; reject the exotic case.
    for (size_t i = 0; i < slots_.size(); ++i)
        if (slots_[i].alias_of >= 0 && slots_[i].binding >= 0) return fail("reshape output bound as network output");
    // First-fit over tensors sorted by size (descending); two tensors may overlap in memory iff their live ranges do not.
    std::vector<int> order;
    for (size_t i = 0; i < slots_.size(); ++i)
        if (slots_[i].used && slots_[i].binding < 0 && slots_[i].alias_of < 0 && slots_[i].pair_of < 0) order.push_back(static_cast<int>(i));
    auto bytesOf = [&](int id) {
        return (slots_[id].elems * max_batch_ * sizeof(float) * (slots_[id].has_partner ? 2 : 1) + 255) & ~static_cast<size_t>(255);
    };
    std::sort(order.begin(), order.end(), [&](int a, int b) { return bytesOf(a) > bytesOf(b); });
    std::vector<int> placed;
    arena_bytes_ = 0;
    for (int id : order) {
        const size_t sz = bytesOf(id);
        std::vector<std::pair<size_t, size_t>> busy;   // [begin, end) of time-overlapping placed tensors
        for (int p : placed)
            if (!(slots_[p].last < slots_[id].first || slots_[id].last < slots_[p].first))
                busy.emplace_back(slots_[p].offset, slots_[p].offset + bytesOf(p));
        std::sort(busy.begin(), busy.end());
        size_t off = 0;
        for (auto& b : busy) {
            if (off + sz <= b.first) break;
            off = std::max(off, b.second);
        }
        slots_[id].offset = off;
        arena_bytes_ = std::max(arena_bytes_, off + sz);
        placed.push_back(id);
    }
    if (getenv("REDTAIL_ENGINE_DUMP_PLAN")) {      // debugging aid: the memory plan on stderr
        for (int id : order)
            fprintf(stderr, "[redtail] tensor %3d %-40s off %12zu bytes %12zu live [%d, %d]%s\n", id, slots_[id].name.c_str(), slots_[id].offset,
                    bytesOf(id), slots_[id].first, slots_[id].last, slots_[id].has_partner ? " (pair)" : "");
        for (size_t si = 0; si < steps_.size(); ++si) {
            fprintf(stderr, "[redtail] step %3zu %-60s ws %zu in", si, steps_[si].name.substr(0, 60).c_str(), steps_[si].workspace);
            for (int id : steps_[si].in) fprintf(stderr, " %d", id);
            fprintf(stderr, " out");
            for (int id : steps_[si].out) fprintf(stderr, " %d", id);
            fprintf(stderr, "\n");
        }
    }
    workspace_bytes_ = 0;
    for (auto& s : steps_) workspace_bytes_ = std::max(workspace_bytes_, s.workspace);
    logMsg(log_, ILogger::Severity::kINFO, "engine: " + std::to_string(steps_.size()) + " steps, activation arena " +
                                               std::to_string(arena_bytes_ >> 20) + " MiB, workspace " + std::to_string(workspace_bytes_ >> 20) + " MiB");
    return true;
}

bool ContextImpl::run(int batchSize, void** bindings, cudaStream_t stream, bool profile)
{
    EngineImpl& e = *engine_;
    if (batchSize < 1 || batchSize > e.max_batch_ || bindings == nullptr) {
        logMsg(e.log_, ILogger::Severity::kERROR, "execute: invalid batch size or bindings");
        return false;
    }
    if (!alloc_ok_) {
        logMsg(e.log_, ILogger::Severity::kERROR, "execute: this context has no activation memory (allocation failed at creation)");
        return false;
    }
    std::function<void*(int)> ptr = [&](int id) -> void* {
        size_t extra = 0;
        while (e.slots_[id].alias_of >= 0 || e.slots_[id].pair_of >= 0) {
            if (e.slots_[id].alias_of >= 0) { id = e.slots_[id].alias_of; continue; }
            extra += e.slots_[id].elems * static_cast<size_t>(batchSize) * sizeof(float);   // behind the partner's samples
            id = e.slots_[id].pair_of;
        }
        const TensorSlot& s = e.slots_[id];
        if (s.binding >= 0) return static_cast<char*>(bindings[s.binding]) + extra;
        return static_cast<char*>(arena_) + s.offset + extra;
    };
    std::vector<cudaEvent_t> ev;
    if (profile) {
        ev.resize(e.steps_.size() + 1);
        for (auto& x : ev) cudaEventCreate(&x);
        cudaEventRecord(ev[0], stream);
    }
    bool ok = true;
    static const bool trace = getenv("REDTAIL_ENGINE_TRACE") != nullptr;      // debugging aid: name every step on stderr and wait for it
    for (size_t i = 0; i < e.steps_.size(); ++i) {
        if (trace) { fprintf(stderr, "[redtail] step %zu %s ...", i, e.steps_[i].name.c_str()); fflush(stderr); }
        const int rc = e.steps_[i].run(batchSize, ptr, workspace_, stream);
        if (trace) {
            const cudaError_t ce = cudaStreamSynchronize(stream);
            fprintf(stderr, " rc %d, %s\n", rc, cudaGetErrorString(ce));
            fflush(stderr);
        }
        if (rc != 0) {
            logMsg(e.log_, ILogger::Severity::kERROR, e.steps_[i].name + ": enqueue failed with status " + std::to_string(rc));
            ok = false;
            break;
        }
        if (debug_sync_ && cudaStreamSynchronize(stream) != cudaSuccess) { ok = false; break; }
        if (profile) cudaEventRecord(ev[i + 1], stream);
    }
    if (profile) {
        cudaStreamSynchronize(stream);
        for (size_t i = 0; ok && i < e.steps_.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            profiler_->reportLayerTime(e.steps_[i].name.c_str(), ms);
        }
        for (auto& x : ev) cudaEventDestroy(x);
    }
    return ok;
}

bool ContextImpl::execute(int batchSize, void** bindings)
{
    // The reference's callers upload with plain cudaMemcpy (pageable memory: the call may return before the DMA has
    // finished) or queue work on the legacy default stream; a non-blocking stream is not ordered against either, so wait
    // for the legacy stream first (free when nothing is pending).
    if (cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess) {
        logMsg(engine_->log_, ILogger::Severity::kERROR, "execute: pending work on the default stream failed");
        return false;
    }
    const bool ok = profiler_ != nullptr ? run(batchSize, bindings, stream_, true) : launch(batchSize, bindings, stream_);
    const cudaError_t err = cudaStreamSynchronize(stream_);
    if (err != cudaSuccess) {
        logMsg(engine_->log_, ILogger::Severity::kERROR, std::string("execute: ") + cudaGetErrorString(err));
        return false;
    }
    return ok;
}

bool ContextImpl::enqueue(int batchSize, void** bindings, cudaStream_t stream, cudaEvent_t* inputConsumed)
{
    const bool ok = launch(batchSize, bindings, stream);
    if (inputConsumed) cudaEventRecord(*inputConsumed, stream);
    return ok;
}

// ------------------------------------------------------------------------------------------------------------------
// Builder / runtime
// ------------------------------------------------------------------------------------------------------------------
class BuilderImpl : public IBuilder {
public:
    explicit BuilderImpl(ILogger& log) : log_(log) {}
    INetworkDefinition* createNetwork() override { return new NetworkImpl(log_); }
    void setMaxBatchSize(int b) override { max_batch_ = b; }
    int getMaxBatchSize() const override { return max_batch_; }
    void setMaxWorkspaceSize(size_t w) override { max_ws_ = w; }
    size_t getMaxWorkspaceSize() const override { return max_ws_; }
    void setHalf2Mode(bool m) override { half2_ = m; }
    bool getHalf2Mode() const override { return half2_; }
    void setDebugSync(bool s) override { debug_sync_ = s; }
    bool getDebugSync() const override { return debug_sync_; }
    void setMinFindIterations(int n) override { min_find_ = n; }
    int getMinFindIterations() const override { return min_find_; }
    void setAverageFindIterations(int n) override { avg_find_ = n; }
    int getAverageFindIterations() const override { return avg_find_; }
    ICudaEngine* buildCudaEngine(INetworkDefinition& network) override
    {
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            logMsg(log_, ILogger::Severity::kERROR, "buildCudaEngine: no CUDA device (this engine has no CPU path).");
            return nullptr;
        }
        if (int8_) {
            logMsg(log_, ILogger::Severity::kERROR, "buildCudaEngine: INT8 mode is not supported (platformHasFastInt8() is false); use fp32 or fp16.");
            return nullptr;
        }
        auto* e = new EngineImpl(log_);
        if (!e->build(static_cast<NetworkImpl&>(network), max_batch_, half2_)) {
            delete e;
            return nullptr;
        }
        return e;
    }
    bool platformHasFastFp16() const override { return true; }
    bool platformHasFastInt8() const override { return false; }
    void destroy() override { delete this; }
    void setInt8Mode(bool m) override { int8_ = m; }
    bool getInt8Mode() const override { return int8_; }
    void setInt8Calibrator(IInt8Calibrator*) override {}

private:
    ILogger& log_;
    int max_batch_ = 1, min_find_ = 1, avg_find_ = 1;
    size_t max_ws_ = 0;
    bool half2_ = false, debug_sync_ = false, int8_ = false;
};

class RuntimeImpl : public IRuntime {
public:
    explicit RuntimeImpl(ILogger& log) : log_(log) {}
    ICudaEngine* deserializeCudaEngine(const void* blob, size_t size, IPluginFactory* factory) override
    {
        auto fail = [&](const std::string& why) -> ICudaEngine* {
            logMsg(log_, ILogger::Severity::kERROR, "deserializeCudaEngine: " + why);
            return nullptr;
        };
        if (!blob || size < 16 || memcmp(blob, kPlanMagic, 8) != 0) return fail("not a redtail_b200 plan");
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return fail("no CUDA device (this engine has no CPU path).");
        // The layer weights of the rebuilt network point into this copy; the engine keeps it alive.
        auto storage = std::make_shared<std::string>(static_cast<const char*>(blob), size);
        PlanReader r{storage->data(), storage->size()};
        r.pos = 8;
        if (r.get<int32_t>() != kPlanVersion) return fail("unsupported plan version");
        const int max_batch = r.get<int32_t>();
        const bool half2 = r.get<uint8_t>() != 0;
        std::unique_ptr<NetworkImpl> net(new NetworkImpl(log_));
        std::map<int, ITensor*> tensors;                       // plan tensor id -> tensor of the rebuilt network
        const int nin = r.get<int32_t>();
        for (int i = 0; r.ok && i < nin; ++i) {
            const int id = r.get<int32_t>();
            const std::string name = r.str();
            const DataType type = static_cast<DataType>(r.get<int32_t>());
            const Dims dims = r.dims();
            tensors[id] = net->addInput(name.c_str(), type, dims);
        }
        const int nl = r.get<int32_t>();
        for (int li = 0; r.ok && li < nl; ++li) {
            const LKind kind = static_cast<LKind>(r.get<int32_t>());
            const std::string name = r.str();
            std::vector<ITensor*> in;
            const int n_in = r.get<int32_t>();
            for (int i = 0; i < n_in; ++i) {
                auto it = tensors.find(r.get<int32_t>());
                if (it == tensors.end() || it->second == nullptr) return fail(name + ": input tensor not defined yet");
                in.push_back(it->second);
            }
            std::vector<std::pair<int, std::string>> outs;
            const int n_out = r.get<int32_t>();
            for (int i = 0; i < n_out; ++i) { const int id = r.get<int32_t>(); outs.emplace_back(id, r.str()); }
            if (!r.ok || in.empty()) return fail("truncated plan");
            ILayer* layer = nullptr;
            switch (kind) {
                case LKind::kConv:
                case LKind::kDeconv: {
                    const int maps = r.get<int32_t>();
                    const int kh = r.get<int32_t>(), kw_ = r.get<int32_t>();
                    const int sh = r.get<int32_t>(), sw = r.get<int32_t>();
                    const int ph = r.get<int32_t>(), pw = r.get<int32_t>();
                    const Weights kw = r.weights(), bw = r.weights();
                    if (kind == LKind::kConv) {
                        auto* l = net->addConvolution(*in[0], maps, DimsHW(kh, kw_), kw, bw);
                        l->setStride(DimsHW(sh, sw)); l->setPadding(DimsHW(ph, pw));
                        layer = l;
                    } else {
                        auto* l = net->addDeconvolution(*in[0], maps, DimsHW(kh, kw_), kw, bw);
                        l->setStride(DimsHW(sh, sw)); l->setPadding(DimsHW(ph, pw));
                        layer = l;
                    }
                    break;
                }
                case LKind::kScale: {
                    const ScaleMode mode = static_cast<ScaleMode>(r.get<int32_t>());
                    const Weights shift = r.weights(), scale = r.weights(), power = r.weights();
                    layer = net->addScale(*in[0], mode, shift, scale, power);
                    break;
                }
                case LKind::kActivation: layer = net->addActivation(*in[0], static_cast<ActivationType>(r.get<int32_t>())); break;
                case LKind::kEltwise:
                    if (in.size() != 2) return fail(name + ": element-wise layer needs two inputs");
                    layer = net->addElementWise(*in[0], *in[1], static_cast<ElementWiseOperation>(r.get<int32_t>()));
                    break;
                case LKind::kConcat: layer = net->addConcatenation(in.data(), static_cast<int>(in.size())); break;
                case LKind::kShuffle: {
                    const bool has = r.get<uint8_t>() != 0;
                    const Dims rd = r.dims();
                    auto* l = net->addShuffle(*in[0]);
                    if (has) l->setReshapeDimensions(rd);
                    layer = l;
                    break;
                }
                case LKind::kPlugin: {
                    const int64_t n = r.get<int64_t>();
                    r.align8();
                    if (!r.ok || n <= 0 || r.pos + static_cast<size_t>(n) > r.size) return fail(name + ": truncated plugin blob");
                    if (factory == nullptr) return fail(name + ": the plan contains plugin layers but no IPluginFactory was given");
                    IPlugin* plugin = factory->createPlugin(name.c_str(), r.base + r.pos, static_cast<size_t>(n));
                    r.pos += static_cast<size_t>(n);
                    if (plugin == nullptr) return fail(name + ": IPluginFactory::createPlugin returned null");
                    auto* ext = dynamic_cast<IPluginExt*>(plugin);
                    layer = ext ? net->addPluginExt(in.data(), static_cast<int>(in.size()), *ext)
                                : net->addPlugin(in.data(), static_cast<int>(in.size()), *plugin);
                    break;
                }
                case LKind::kPooling: {
                    const PoolingType pt = static_cast<PoolingType>(r.get<int32_t>());
                    const int k = r.get<int32_t>(), st2 = r.get<int32_t>(), pd = r.get<int32_t>();
                    const int oh = r.get<int32_t>(), ow = r.get<int32_t>();
                    auto* l = net->addPooling(*in[0], pt, DimsHW(k, k));
                    l->setStride(DimsHW(st2, st2)); l->setPadding(DimsHW(pd, pd));
                    net->pool_from_plan_ = true;
                    static_cast<PoolingLayer*>(l)->d.pool_oh = oh; static_cast<PoolingLayer*>(l)->d.pool_ow = ow;
                    layer = l;
                    break;
                }
                case LKind::kFullyConnected: {
                    const int maps = r.get<int32_t>();
                    const Weights kw = r.weights(), bw = r.weights();
                    layer = net->addFullyConnected(*in[0], maps, kw, bw);
                    break;
                }
                case LKind::kSoftMax: layer = net->addSoftMax(*in[0]); break;
                default: return fail("unknown layer kind in plan");
            }
            if (!r.ok || layer == nullptr) return fail(name + ": truncated plan");
            layer->setName(name.c_str());
            if (layer->getNbOutputs() != static_cast<int>(outs.size())) return fail(name + ": output count mismatch");
            for (size_t i = 0; i < outs.size(); ++i) {
                layer->getOutput(static_cast<int>(i))->setName(outs[i].second.c_str());
                tensors[outs[i].first] = layer->getOutput(static_cast<int>(i));
            }
        }
        const int nout = r.get<int32_t>();
        for (int i = 0; r.ok && i < nout; ++i) {
            auto it = tensors.find(r.get<int32_t>());
            if (it == tensors.end()) return fail("output tensor not defined");
            net->markOutput(*it->second);
        }
        if (!r.ok) return fail("truncated plan");
        auto* e = new EngineImpl(log_);
        e->plan_storage_ = storage;
        if (!e->build(*net, max_batch, half2)) {
            delete e;
            return nullptr;
        }
        return e;
    }
    void destroy() override { delete this; }

private:
    ILogger& log_;
};

}  // namespace

// Host-only extension (no CUDA device needed): the plan of a network WITHOUT building it -- the layer list, parameters,
// weights and plugin blobs exactly as ICudaEngine::serialize() would write them.  Used by tools/dropin (net_driver dump
// mode) to capture the graphs of the reference's generated builders for the test suite's graph-level checker.
// Returns the plan size (0 if a plugin cannot serialise); copies it when buf_len suffices.
extern "C" size_t redtail_serialize_network(void* network, int max_batch, int half2, void* buf, size_t buf_len)
{
    if (!network) return 0;
    NetworkImpl& net = *static_cast<NetworkImpl*>(static_cast<INetworkDefinition*>(network));
    net.invalidate();
    if (!net.resolve()) return 0;
    const std::string plan = EngineImpl::serializeNetwork(net, max_batch > 0 ? max_batch : 1, half2 != 0);
    if (buf && buf_len >= plan.size()) memcpy(buf, plan.data(), plan.size());
    return plan.size();
}

extern "C" void* createInferBuilder_INTERNAL(void* logger, int)
{
    return static_cast<IBuilder*>(new BuilderImpl(*static_cast<ILogger*>(logger)));
}

extern "C" void* createInferRuntime_INTERNAL(void* logger, int)
{
    return static_cast<IRuntime*>(new RuntimeImpl(*static_cast<ILogger*>(logger)));
}

// The eight stereo-DNN plugins, the plugin container, the add* helpers and the deserialisation factory.
//
// Same operator surface and call protocol as the reference's stereoDNN/lib/*_plugin.cpp + internal_utils.cpp
// (getOutputDimensions -> configure[WithFormat] -> initialize -> enqueue* -> terminate; enqueue returns 0 or an error
// code and is asynchronous on the caller's stream).  Behind enqueue() every plugin calls exactly one entry point of the
// C-ABI in include/redtail_b200.h -- there is no cuDNN, no TensorRT and no host-side arithmetic here.
#include <cstring>
#include <mutex>
#include <sstream>
#include <vector>

#include "internal_utils.h"
#include "op_info.h"
#include "redtail_b200.h"

namespace redtail { namespace tensorrt {

using namespace nvinfer1;

namespace {

int rtType(DataType t) { return t == DataType::kHALF ? RT_F16 : RT_F32; }

// Little-endian POD (de)serialisation helpers (byte formats: SURVEY.md 8b).
struct ByteWriter {
    std::string buf;
    template <typename T> void put(T v) { buf.append(reinterpret_cast<const char*>(&v), sizeof(T)); }
};
// A plan is untrusted input: a short or inconsistent blob clears `ok` (the factory then returns nullptr and
// deserializeCudaEngine reports the error) instead of reading past the buffer or the Dims array.
struct ByteReader {
    const char* p; size_t left; bool ok = true;
    template <typename T> T get() {
        T v{};
        if (left < sizeof(T)) { ok = false; left = 0; return v; }
        memcpy(&v, p, sizeof(T));
        p += sizeof(T); left -= sizeof(T);
        return v;
    }
    void dims(Dims& d) {
        d = Dims{};
        const int32_t n = get<int32_t>();
        if (n < 0 || n > Dims::MAX_DIMS) { ok = false; return; }
        d.nbDims = n;
        for (int i = 0; i < n; i++) d.d[i] = get<int32_t>();
    }
    bool done() const { return ok && left == 0; }
};

void logDims(ILogger& log, const std::string& name, const char* what, Dims d)
{
    log.log(ILogger::Severity::kINFO, (name + ": " + what + DimsUtils::toString(d)).c_str());
}

// ---------------------------------------------------------------------------------------------------------------
// ELU (alpha = 1).  Reference: lib/elu_plugin.cpp:16-213.
// ---------------------------------------------------------------------------------------------------------------
class EluPlugin : public IPluginExt, public IRedtailOp
{
public:
    EluPlugin(DataType data_type, ILogger& log, std::string name) : log_(log)
    {
        assert(data_type == DataType::kFLOAT || data_type == DataType::kHALF);
        info_.kind = OpKind::kElu; info_.name = name; info_.data_type = data_type;
    }
    // Deserialisation: i32 dtype, u8 format, i32 nbDims, i32 d[] (the type tag was consumed by the factory).
    EluPlugin(const char* name, const void* data, size_t size, ILogger& log) : log_(log)
    {
        ByteReader r{static_cast<const char*>(data), size};
        info_.kind = OpKind::kElu; info_.name = name;
        info_.data_type = static_cast<DataType>(r.get<int32_t>());
        format_ = static_cast<PluginFormat>(r.get<uint8_t>());
        r.dims(in_dims_);
        valid_ = r.done() && (info_.data_type == DataType::kFLOAT || info_.data_type == DataType::kHALF);
    }
    bool valid() const { return valid_; }
    const OpInfo& opInfo() const override { return info_; }

    bool supportsFormat(DataType type, PluginFormat format) const override
    {
        return type == info_.data_type && format == PluginFormat::kNCHW;
    }
    int getNbOutputs() const override { return 1; }
    Dims getOutputDimensions(int index, const Dims* inputs, int nbInputDims) override
    {
        assert(index == 0 && nbInputDims == 1);
        UNUSEDR(index); UNUSEDR(nbInputDims);
        in_dims_ = inputs[0];
        return in_dims_;
    }
    void configureWithFormat(const Dims* inputDims, int nbInputs, const Dims* outputDims, int nbOutputs,
                             DataType type, PluginFormat format, int maxBatchSize) override
    {
        assert(nbInputs == 1 && nbOutputs == 1);
        assert(DimsUtils::areEqual(inputDims[0], outputDims[0]));
        assert(type == info_.data_type && format == PluginFormat::kNCHW);
        UNUSEDR(nbInputs); UNUSEDR(nbOutputs); UNUSEDR(outputDims); UNUSEDR(type); UNUSEDR(maxBatchSize);
        in_dims_ = inputDims[0];
        format_ = format;
        logDims(log_, info_.name, "Dims: ", in_dims_);
    }
    int initialize() override { return 0; }
    void terminate() override {}
    size_t getWorkspaceSize(int) const override { return 0; }
    int enqueue(int batchSize, const void* const* inputs, void** outputs, void*, cudaStream_t stream) override
    {
        return rt_elu(rtType(info_.data_type), inputs[0], outputs[0],
                      static_cast<int64_t>(batchSize) * DimsUtils::getTensorSize(in_dims_), stream);
    }
    size_t getSerializationSize() override { return blob().size(); }
    void serialize(void* buffer) override { auto b = blob(); memcpy(buffer, b.data(), b.size()); }

private:
    std::string blob() const
    {
        ByteWriter w;
        w.put<int32_t>(static_cast<int32_t>(StereoDnnPluginFactory::PluginType::kElu));
        w.put<int32_t>(static_cast<int32_t>(info_.data_type));
        w.put<uint8_t>(static_cast<uint8_t>(format_));
        w.put<int32_t>(in_dims_.nbDims);
        for (int i = 0; i < in_dims_.nbDims; i++) w.put<int32_t>(in_dims_.d[i]);
        return w.buf;
    }
    OpInfo info_;
    PluginFormat format_ = PluginFormat::kNCHW;
    Dims in_dims_{};
    bool valid_ = true;       // false: constructed from a malformed serialised blob
    ILogger& log_;
};

// ---------------------------------------------------------------------------------------------------------------
// Cost volume.  Reference: lib/cost_volume_plugin.cpp:16-186.  Unlike the reference (maxBatchSize == 1 asserted,
// :99,118,124) batches are supported: samples are independent, leading N dim.
// ---------------------------------------------------------------------------------------------------------------
class CostVolumePlugin : public IPluginExt, public IRedtailOp
{
public:
    CostVolumePlugin(DataType data_type, CostVolumeType cv_type, int max_disparity, ILogger& log, std::string name) : log_(log)
    {
        assert(data_type == DataType::kFLOAT || data_type == DataType::kHALF);
        assert(max_disparity > 0);
        info_.kind = OpKind::kCostVolume; info_.name = name; info_.data_type = data_type;
        info_.cv_type = cv_type; info_.max_disparity = max_disparity;
    }
    // i32 dtype, u8 format, i32 cv_type, i32 max_disp, i32 n_in, i32 in[], i32 n_out, i32 out[]
    CostVolumePlugin(const char* name, const void* data, size_t size, ILogger& log) : log_(log)
    {
        ByteReader r{static_cast<const char*>(data), size};
        info_.kind = OpKind::kCostVolume; info_.name = name;
        info_.data_type = static_cast<DataType>(r.get<int32_t>());
        format_ = static_cast<PluginFormat>(r.get<uint8_t>());
        info_.cv_type = static_cast<CostVolumeType>(r.get<int32_t>());
        info_.max_disparity = r.get<int32_t>();
        r.dims(in_dims_);
        r.dims(out_dims_);
        valid_ = r.done() && info_.max_disparity > 0 && (info_.data_type == DataType::kFLOAT || info_.data_type == DataType::kHALF);
    }
    bool valid() const { return valid_; }
    const OpInfo& opInfo() const override { return info_; }

    bool supportsFormat(DataType type, PluginFormat format) const override
    {
        return type == info_.data_type && format == PluginFormat::kNCHW;
    }
    int getNbOutputs() const override { return 1; }
    Dims getOutputDimensions(int index, const Dims* inputs, int nbInputDims) override
    {
        assert(index == 0 && nbInputDims == 2);
        assert(inputs[0].nbDims == 3 && DimsUtils::areEqual(inputs[0], inputs[1]));
        UNUSEDR(index); UNUSEDR(nbInputDims);
        in_dims_ = inputs[0];
        if (info_.cv_type == CostVolumeType::kDefault)
            out_dims_ = DimsNCHW(info_.max_disparity, 2 * in_dims_.d[0], in_dims_.d[1], in_dims_.d[2]);
        else
            out_dims_ = DimsCHW(info_.max_disparity, in_dims_.d[1], in_dims_.d[2]);
        return out_dims_;
    }
    void configureWithFormat(const Dims* inputDims, int nbInputs, const Dims* outputDims, int nbOutputs,
                             DataType type, PluginFormat format, int maxBatchSize) override
    {
        assert(nbInputs == 2 && nbOutputs == 1);
        assert(DimsUtils::areEqual(inputDims[0], in_dims_) && DimsUtils::areEqual(inputDims[1], in_dims_));
        assert(DimsUtils::areEqual(outputDims[0], out_dims_));
        assert(type == info_.data_type && format == PluginFormat::kNCHW);
        UNUSEDR(inputDims); UNUSEDR(nbInputs); UNUSEDR(outputDims); UNUSEDR(nbOutputs); UNUSEDR(type); UNUSEDR(maxBatchSize);
        format_ = format;
        logDims(log_, info_.name, "InDims : ", in_dims_);
        logDims(log_, info_.name, "OutDims: ", out_dims_);
    }
    int initialize() override { return 0; }
    void terminate() override {}
    size_t getWorkspaceSize(int) const override { return 0; }
    int enqueue(int batchSize, const void* const* inputs, void** outputs, void*, cudaStream_t stream) override
    {
        if (info_.cv_type == CostVolumeType::kDefault)
            return rt_cost_volume(rtType(info_.data_type), inputs[0], inputs[1], outputs[0], batchSize,
                                  in_dims_.d[0], in_dims_.d[1], in_dims_.d[2], info_.max_disparity, stream);
        return rt_corr_cost_volume(rtType(info_.data_type), inputs[0], inputs[1], outputs[0], batchSize,
                                   in_dims_.d[0], in_dims_.d[1], in_dims_.d[2], info_.max_disparity, stream);
    }
    size_t getSerializationSize() override { return blob().size(); }
    void serialize(void* buffer) override { auto b = blob(); memcpy(buffer, b.data(), b.size()); }

private:
    std::string blob() const
    {
        ByteWriter w;
        w.put<int32_t>(static_cast<int32_t>(StereoDnnPluginFactory::PluginType::kCostVolume));
        w.put<int32_t>(static_cast<int32_t>(info_.data_type));
        w.put<uint8_t>(static_cast<uint8_t>(format_));
        w.put<int32_t>(static_cast<int32_t>(info_.cv_type));
        w.put<int32_t>(info_.max_disparity);
        w.put<int32_t>(in_dims_.nbDims);
        for (int i = 0; i < in_dims_.nbDims; i++) w.put<int32_t>(in_dims_.d[i]);
        w.put<int32_t>(out_dims_.nbDims);
        for (int i = 0; i < out_dims_.nbDims; i++) w.put<int32_t>(out_dims_.d[i]);
        return w.buf;
    }
    OpInfo info_;
    PluginFormat format_ = PluginFormat::kNCHW;
    Dims in_dims_{}, out_dims_{};
    bool valid_ = true;       // false: constructed from a malformed serialised blob
    ILogger& log_;
};

// ---------------------------------------------------------------------------------------------------------------
// Soft-arg{max,min}.  Reference: lib/softargmax_plugin.cpp:17-310.  One fused kernel, no workspace
// (the reference needs 2x the input, :161-165).
// ---------------------------------------------------------------------------------------------------------------
class SoftargmaxPlugin : public IPluginExt, public IRedtailOp
{
public:
    SoftargmaxPlugin(DataType data_type, SoftargmaxType sm_type, ILogger& log, std::string name) : log_(log)
    {
        assert(data_type == DataType::kFLOAT || data_type == DataType::kHALF);
        info_.kind = OpKind::kSoftargmax; info_.name = name; info_.data_type = data_type; info_.sm_type = sm_type;
    }
    // i32 dtype, i32 sm_type, i32 n_in, i32 in[], i32 n_out, i32 out[]
    SoftargmaxPlugin(const char* name, const void* data, size_t size, ILogger& log) : log_(log)
    {
        ByteReader r{static_cast<const char*>(data), size};
        info_.kind = OpKind::kSoftargmax; info_.name = name;
        info_.data_type = static_cast<DataType>(r.get<int32_t>());
        info_.sm_type = static_cast<SoftargmaxType>(r.get<int32_t>());
        r.dims(in_dims_);
        r.dims(out_dims_);
        valid_ = r.done() && (info_.data_type == DataType::kFLOAT || info_.data_type == DataType::kHALF);
    }
    bool valid() const { return valid_; }
    const OpInfo& opInfo() const override { return info_; }

    bool supportsFormat(DataType type, PluginFormat format) const override
    {
        return type == info_.data_type && format == PluginFormat::kNCHW;
    }
    int getNbOutputs() const override { return 1; }
    Dims getOutputDimensions(int index, const Dims* inputs, int nbInputDims) override
    {
        assert(index == 0 && nbInputDims == 1);
        assert(inputs[0].nbDims == 3 || inputs[0].nbDims == 4);
        UNUSEDR(index); UNUSEDR(nbInputDims);
        if (inputs[0].nbDims == 3)
            in_dims_ = inputs[0];
        else {
            assert(inputs[0].d[1] == 1);     // [D,1,H,W]
            in_dims_ = {3, {inputs[0].d[0], inputs[0].d[2], inputs[0].d[3]}};
        }
        out_dims_ = DimsCHW(1, in_dims_.d[1], in_dims_.d[2]);
        return out_dims_;
    }
    void configureWithFormat(const Dims*, int nbInputs, const Dims* outputDims, int nbOutputs,
                             DataType type, PluginFormat format, int) override
    {
        assert(nbInputs == 1 && nbOutputs == 1);
        assert(DimsUtils::areEqual(outputDims[0], out_dims_));
        assert(type == info_.data_type && format == PluginFormat::kNCHW);
        UNUSEDR(nbInputs); UNUSEDR(nbOutputs); UNUSEDR(outputDims); UNUSEDR(type); UNUSEDR(format);
        logDims(log_, info_.name, "InDims : ", in_dims_);
        logDims(log_, info_.name, "OutDims: ", out_dims_);
    }
    int initialize() override { return 0; }
    void terminate() override {}
    size_t getWorkspaceSize(int) const override { return 0; }
    int enqueue(int batchSize, const void* const* inputs, void** outputs, void*, cudaStream_t stream) override
    {
        return rt_softargmax(rtType(info_.data_type), info_.sm_type == SoftargmaxType::kMin, inputs[0], outputs[0],
                             batchSize, in_dims_.d[0], static_cast<int64_t>(in_dims_.d[1]) * in_dims_.d[2], stream);
    }
    size_t getSerializationSize() override { return blob().size(); }
    void serialize(void* buffer) override { auto b = blob(); memcpy(buffer, b.data(), b.size()); }

private:
    std::string blob() const
    {
        ByteWriter w;
        w.put<int32_t>(static_cast<int32_t>(StereoDnnPluginFactory::PluginType::kSoftargmax));
        w.put<int32_t>(static_cast<int32_t>(info_.data_type));
        w.put<int32_t>(static_cast<int32_t>(info_.sm_type));
        w.put<int32_t>(in_dims_.nbDims);
        for (int i = 0; i < in_dims_.nbDims; i++) w.put<int32_t>(in_dims_.d[i]);
        w.put<int32_t>(out_dims_.nbDims);
        for (int i = 0; i < out_dims_.nbDims; i++) w.put<int32_t>(out_dims_.d[i]);
        return w.buf;
    }
    OpInfo info_;
    Dims in_dims_{}, out_dims_{};
    bool valid_ = true;       // false: constructed from a malformed serialised blob
    ILogger& log_;
};

// ---------------------------------------------------------------------------------------------------------------
// 3-D convolution and transposed convolution.  Reference: lib/conv3d_plugin.cpp:19-374,
// lib/conv3d_transpose_plugin.cpp:24-397.  The plan (weight repack + upload) is created in configure(), as the
// reference uploads weights there (conv3d_plugin.cpp:123-131), and released in terminate().
// ---------------------------------------------------------------------------------------------------------------
int precisionFromEnv()
{
    const char* e = getenv("REDTAIL_CONV3D_PRECISION");   // "fp32" (default) | "fp16" | "simt"
    if (!e) return RT_PREC_FP32;
    if (!strcmp(e, "fp16")) return RT_PREC_FP16;
    if (!strcmp(e, "simt")) return RT_PREC_SIMT;
    return RT_PREC_FP32;
}

class Conv3DPluginBase : public IPlugin, public IRedtailOp
{
public:
    Conv3DPluginBase(bool transposed, Conv3DType conv_type, Dims kernel_dims, Dims out_dims, Dims stride_dims,
                     Dims pad_start_dims, Dims pad_end_dims, Weights kernel_weights, Weights bias_weights,
                     ILogger& log, std::string name) : transposed_(transposed), log_(log)
    {
        // Same argument contract as the reference constructors (conv3d_plugin.cpp:22-66).
        assert(conv_type == Conv3DType::kTensorFlow);     // the only flavour the builders use
        assert(kernel_dims.nbDims == 5);
        assert(stride_dims.nbDims == 3 && pad_start_dims.nbDims == 3 && pad_end_dims.nbDims == 3);
        assert(pad_start_dims.d[1] == pad_end_dims.d[1] && pad_start_dims.d[2] == pad_end_dims.d[2]);
        assert(pad_start_dims.d[0] == pad_end_dims.d[0] || pad_start_dims.d[0] == pad_end_dims.d[0] - 1);
        assert(kernel_weights.type == DataType::kFLOAT || kernel_weights.type == DataType::kHALF);
        assert(kernel_weights.count > 0 && kernel_weights.values != nullptr);
        assert((bias_weights.count > 0 && bias_weights.values != nullptr) ||
               (bias_weights.count == 0 && bias_weights.values == nullptr));
        assert(bias_weights.count == 0 || bias_weights.type == kernel_weights.type);
        UNUSEDR(conv_type);
        info_.kind = transposed ? OpKind::kConv3DTranspose : OpKind::kConv3D;
        info_.name = name;
        info_.kernel_dims = kernel_dims; info_.stride = stride_dims;
        info_.pad_start = pad_start_dims; info_.pad_end = pad_end_dims;
        info_.out_dims = out_dims;
        info_.kernel = kernel_weights; info_.bias = bias_weights;
        if (transposed) assert(out_dims.nbDims == 4);
    }
    ~Conv3DPluginBase() override { terminate(); }
    const OpInfo& opInfo() const override { return info_; }

    int getNbOutputs() const override { return 1; }
    Dims getOutputDimensions(int index, const Dims* inputs, int nbInputDims) override
    {
        assert(index == 0 && nbInputDims == 1 && inputs[0].nbDims == 4);
        UNUSEDR(index); UNUSEDR(nbInputDims);
        in_dims_ = DimsNCHW(inputs[0].d[0], inputs[0].d[1], inputs[0].d[2], inputs[0].d[3]);
        const Dims& k = info_.kernel_dims;
        if (!transposed_) {
            // input [D,C,H,W] -> output [K,Do,Ho,Wo], symmetric pad = pad_start (conv_utils.cpp:46-81).
            assert(in_dims_.d[1] == k.d[2]);
            const int sp[3] = {in_dims_.d[0], in_dims_.d[2], in_dims_.d[3]};
            const int kk[3] = {k.d[1], k.d[3], k.d[4]};
            int o[3];
            for (int i = 0; i < 3; i++) o[i] = (sp[i] + 2 * info_.pad_start.d[i] - kk[i]) / info_.stride.d[i] + 1;
            out_dims_ = DimsNCHW(k.d[0], o[0], o[1], o[2]);
        } else {
            // input [K,Dy,Hy,Wy] -> output out_dims [Dx,C,Hx,Wx] (conv3d_transpose_plugin.cpp:86-114).
            assert(in_dims_.d[0] == k.d[0]);
            assert(info_.out_dims.d[1] == k.d[2]);
            out_dims_ = DimsNCHW(info_.out_dims.d[0], info_.out_dims.d[1], info_.out_dims.d[2], info_.out_dims.d[3]);
        }
        return out_dims_;
    }
    void configure(const Dims* inputDims, int nbInputs, const Dims* outputDims, int nbOutputs, int maxBatchSize) override
    {
        assert(nbInputs == 1 && nbOutputs == 1);
        assert(DimsUtils::areEqual(inputDims[0], in_dims_) && DimsUtils::areEqual(outputDims[0], out_dims_));
        UNUSEDR(inputDims); UNUSEDR(nbInputs); UNUSEDR(outputDims); UNUSEDR(nbOutputs);
        max_batch_size_ = maxBatchSize;
        if (plan_ == nullptr) createPlan();
        logDims(log_, info_.name, "InDims  : ", in_dims_);
        logDims(log_, info_.name, "OutDims : ", out_dims_);
    }
    int initialize() override { return plan_ != nullptr ? 0 : -1; }
    void terminate() override
    {
        if (plan_ != nullptr) rt_conv3d_destroy(plan_);
        plan_ = nullptr;
    }
    size_t getWorkspaceSize(int maxBatchSize) const override
    {
        return plan_ ? rt_conv3d_workspace_size(plan_, maxBatchSize) : 0;
    }
    int enqueue(int batchSize, const void* const* inputs, void** outputs, void* workspace, cudaStream_t stream) override
    {
        if (plan_ == nullptr) return -1;
        return rt_conv3d_enqueue(plan_, batchSize, inputs[0], nullptr, outputs[0], workspace, stream);
    }
    // The reference asserts here (conv3d_plugin.cpp:224-229: "not implemented"), which is why only ResNet18_2D plans could
    // be saved.  Blob: tag | kernel dims KVCRS | out dims (transposed) | stride | pad start | pad end | weight type |
    // kernel count | bias count | kernel bytes | bias bytes; StereoDnnPluginFactory::createPlugin is the reader.
    size_t getSerializationSize() override { return blob().size(); }
    void serialize(void* buffer) override { auto b = blob(); memcpy(buffer, b.data(), b.size()); }

    // Deserialised plugins own their weights (the plan blob is only valid during createPlugin()).
    void ownWeights()
    {
        const size_t es = info_.kernel.type == DataType::kHALF ? 2 : 4;
        owned_kernel_.assign(static_cast<const char*>(info_.kernel.values), static_cast<const char*>(info_.kernel.values) + info_.kernel.count * es);
        info_.kernel.values = owned_kernel_.data();
        if (info_.bias.count > 0) {
            owned_bias_.assign(static_cast<const char*>(info_.bias.values), static_cast<const char*>(info_.bias.values) + info_.bias.count * es);
            info_.bias.values = owned_bias_.data();
        }
    }

private:
    std::string blob() const
    {
        ByteWriter w;
        w.put<int32_t>(static_cast<int32_t>(transposed_ ? StereoDnnPluginFactory::PluginType::kConv3DTranspose
                                                         : StereoDnnPluginFactory::PluginType::kConv3D));
        for (int i = 0; i < 5; i++) w.put<int32_t>(info_.kernel_dims.d[i]);
        for (int i = 0; i < 4; i++) w.put<int32_t>(transposed_ ? info_.out_dims.d[i] : 0);
        for (int i = 0; i < 3; i++) w.put<int32_t>(info_.stride.d[i]);
        for (int i = 0; i < 3; i++) w.put<int32_t>(info_.pad_start.d[i]);
        for (int i = 0; i < 3; i++) w.put<int32_t>(info_.pad_end.d[i]);
        w.put<int32_t>(static_cast<int32_t>(info_.kernel.type));
        w.put<int64_t>(info_.kernel.count);
        w.put<int64_t>(info_.bias.count);
        const size_t es = info_.kernel.type == DataType::kHALF ? 2 : 4;
        w.buf.append(static_cast<const char*>(info_.kernel.values), info_.kernel.count * es);
        if (info_.bias.count > 0) w.buf.append(static_cast<const char*>(info_.bias.values), info_.bias.count * es);
        return w.buf;
    }
    std::vector<char> owned_kernel_, owned_bias_;

    void createPlan()
    {
        rt_conv3d_desc d{};
        const Dims& k = info_.kernel_dims;
        d.transposed = transposed_ ? 1 : 0;
        d.k = k.d[0]; d.v = k.d[1]; d.c = k.d[2]; d.r = k.d[3]; d.s = k.d[4];
        for (int i = 0; i < 3; i++) { d.stride[i] = info_.stride.d[i]; d.pad[i] = info_.pad_start.d[i]; }
        for (int i = 0; i < 4; i++) { d.in_dims[i] = in_dims_.d[i]; d.out_dims[i] = out_dims_.d[i]; }
        d.weights_dtype = rtType(info_.kernel.type);
        d.weights = info_.kernel.values;
        d.bias = info_.bias.count > 0 ? info_.bias.values : nullptr;
        assert(info_.kernel.count == static_cast<int64_t>(d.k) * d.v * d.c * d.r * d.s);
        assert(info_.bias.count == 0 || info_.bias.count == (transposed_ ? d.c : d.k));
        d.precision = precisionFromEnv();
        int rc = rt_conv3d_create(&d, &plan_);
        if (rc == RT_ERR_UNSUPPORTED && d.precision != RT_PREC_SIMT) {
            // Channel counts outside the tensor-core tiles (e.g. the 1..8-channel unit-test tensors) run on the
            // fp32 CUDA-core kernels; this is logged, never silent.
            log_.log(ILogger::Severity::kWARNING,
                     (info_.name + ": shape not covered by the tcgen05 kernels, using the fp32 SIMT kernels.").c_str());
            d.precision = RT_PREC_SIMT;
            rc = rt_conv3d_create(&d, &plan_);
        }
        CHECKL(rc, log_);
    }

    bool transposed_;
    OpInfo info_;
    Dims in_dims_{}, out_dims_{};
    int max_batch_size_ = 0;
    rt_conv3d_plan* plan_ = nullptr;
    ILogger& log_;
};

// ---------------------------------------------------------------------------------------------------------------
// Transform (4-D permutation; only {1,0,2,3} exists in the nets).  Reference: lib/transform_plugin.cpp:16-182.
// ---------------------------------------------------------------------------------------------------------------
class TransformPlugin : public IPlugin, public IRedtailOp
{
public:
    TransformPlugin(Permutation permutation, ILogger& log, std::string name) : log_(log)
    {
        info_.kind = OpKind::kTransform; info_.name = name; info_.perm = permutation;
    }
    const OpInfo& opInfo() const override { return info_; }
    int getNbOutputs() const override { return 1; }
    Dims getOutputDimensions(int index, const Dims* inputs, int nbInputDims) override
    {
        assert(index == 0 && nbInputDims == 1 && inputs[0].nbDims == 4);
        UNUSEDR(index); UNUSEDR(nbInputDims);
        in_dims_ = DimsNCHW(inputs[0].d[0], inputs[0].d[1], inputs[0].d[2], inputs[0].d[3]);
        const int* o = info_.perm.order;
        // Only the outer-dims swap is implemented (the reference's stride mapping is itself only right for involutions).
        assert(o[0] == 1 && o[1] == 0 && o[2] == 2 && o[3] == 3);
        out_dims_ = DimsNCHW(in_dims_.d[o[0]], in_dims_.d[o[1]], in_dims_.d[o[2]], in_dims_.d[o[3]]);
        return out_dims_;
    }
    void configure(const Dims* inputDims, int nbInputs, const Dims* outputDims, int nbOutputs, int) override
    {
        assert(nbInputs == 1 && nbOutputs == 1);
        assert(DimsUtils::areEqual(inputDims[0], in_dims_) && DimsUtils::areEqual(outputDims[0], out_dims_));
        UNUSEDR(inputDims); UNUSEDR(nbInputs); UNUSEDR(outputDims); UNUSEDR(nbOutputs);
        logDims(log_, info_.name, "InDims : ", in_dims_);
        logDims(log_, info_.name, "OutDims: ", out_dims_);
    }
    int initialize() override { return 0; }
    void terminate() override {}
    size_t getWorkspaceSize(int) const override { return 0; }
    int enqueue(int batchSize, const void* const* inputs, void** outputs, void*, cudaStream_t stream) override
    {
        return rt_transpose01(RT_F32, inputs[0], outputs[0], batchSize, in_dims_.d[0], in_dims_.d[1],
                              static_cast<int64_t>(in_dims_.d[2]) * in_dims_.d[3], stream);
    }
    size_t getSerializationSize() override { return 5 * sizeof(int32_t); }
    void serialize(void* buffer) override
    {
        ByteWriter w;
        w.put<int32_t>(static_cast<int32_t>(StereoDnnPluginFactory::PluginType::kTransform));
        for (int i = 0; i < 4; i++) w.put<int32_t>(info_.perm.order[i]);
        memcpy(buffer, w.buf.data(), w.buf.size());
    }

private:
    OpInfo info_;
    Dims in_dims_{}, out_dims_{};
    ILogger& log_;
};

// ---------------------------------------------------------------------------------------------------------------
// Padding: zero planes appended on the outermost dim.  Reference: lib/padding_plugin.cpp:15-115.
// ---------------------------------------------------------------------------------------------------------------
class PaddingPlugin : public IPlugin, public IRedtailOp
{
public:
    PaddingPlugin(DimsNCHW pad_start, DimsNCHW pad_end, ILogger& log, std::string name) : log_(log)
    {
        assert(pad_start.n() == 0 && pad_start.c() == 0 && pad_start.h() == 0 && pad_start.w() == 0);
        assert(pad_end.n() >= 0 && pad_end.c() == 0 && pad_end.h() == 0 && pad_end.w() == 0);
        UNUSEDR(pad_start);
        info_.kind = OpKind::kPadding; info_.name = name; info_.pad_end_planes = pad_end.n();
    }
    const OpInfo& opInfo() const override { return info_; }
    int getNbOutputs() const override { return 1; }
    Dims getOutputDimensions(int index, const Dims* inputs, int nbInputDims) override
    {
        assert(index == 0 && nbInputDims == 1 && inputs[0].nbDims == 4);
        UNUSEDR(index); UNUSEDR(nbInputDims);
        in_dims_ = DimsNCHW(inputs[0].d[0], inputs[0].d[1], inputs[0].d[2], inputs[0].d[3]);
        out_dims_ = DimsNCHW(in_dims_.d[0] + info_.pad_end_planes, in_dims_.d[1], in_dims_.d[2], in_dims_.d[3]);
        return out_dims_;
    }
    void configure(const Dims* inputDims, int nbInputs, const Dims* outputDims, int nbOutputs, int) override
    {
        assert(nbInputs == 1 && nbOutputs == 1);
        assert(DimsUtils::areEqual(inputDims[0], in_dims_) && DimsUtils::areEqual(outputDims[0], out_dims_));
        UNUSEDR(inputDims); UNUSEDR(nbInputs); UNUSEDR(outputDims); UNUSEDR(nbOutputs);
        logDims(log_, info_.name, "InDims : ", in_dims_);
        logDims(log_, info_.name, "OutDims: ", out_dims_);
    }
    int initialize() override { return 0; }
    void terminate() override {}
    size_t getWorkspaceSize(int) const override { return 0; }
    int enqueue(int batchSize, const void* const* inputs, void** outputs, void*, cudaStream_t stream) override
    {
        const int64_t plane = static_cast<int64_t>(in_dims_.d[1]) * in_dims_.d[2] * in_dims_.d[3];
        return rt_pad_planes(RT_F32, inputs[0], outputs[0], batchSize, in_dims_.d[0], plane, info_.pad_end_planes, stream);
    }
    size_t getSerializationSize() override { return 2 * sizeof(int32_t); }
    void serialize(void* buffer) override
    {
        ByteWriter w;
        w.put<int32_t>(static_cast<int32_t>(StereoDnnPluginFactory::PluginType::kPadding));
        w.put<int32_t>(info_.pad_end_planes);
        memcpy(buffer, w.buf.data(), w.buf.size());
    }

private:
    OpInfo info_;
    Dims in_dims_{}, out_dims_{};
    ILogger& log_;
};

// ---------------------------------------------------------------------------------------------------------------
// Slice: [start, end) of the outermost dim.  Reference: lib/slice_plugin.cpp:17-111.
// ---------------------------------------------------------------------------------------------------------------
class SlicePlugin : public IPlugin, public IRedtailOp
{
public:
    SlicePlugin(Dims dims, Dims slice_start, Dims slice_end, ILogger& log, std::string name) : in_dims_(dims), log_(log)
    {
        assert(dims.nbDims == 4 && slice_start.nbDims == 4 && slice_end.nbDims == 4);
        assert(0 <= slice_start.d[0] && slice_start.d[0] < slice_end.d[0] && slice_end.d[0] <= dims.d[0]);
        for (int i = 1; i < 4; i++) assert(slice_start.d[i] == 0 && slice_end.d[i] == dims.d[i]);
        info_.kind = OpKind::kSlice; info_.name = name;
        info_.slice_start = slice_start.d[0]; info_.slice_end = slice_end.d[0];
    }
    const OpInfo& opInfo() const override { return info_; }
    int getNbOutputs() const override { return 1; }
    Dims getOutputDimensions(int index, const Dims* inputs, int nbInputDims) override
    {
        assert(index == 0 && nbInputDims == 1 && inputs[0].nbDims == 4);
        UNUSEDR(index); UNUSEDR(nbInputDims);
        out_dims_ = DimsNCHW(info_.slice_end - info_.slice_start, inputs[0].d[1], inputs[0].d[2], inputs[0].d[3]);
        return out_dims_;
    }
    void configure(const Dims* inputDims, int nbInputs, const Dims* outputDims, int nbOutputs, int) override
    {
        assert(nbInputs == 1 && nbOutputs == 1);
        assert(DimsUtils::areEqual(inputDims[0], in_dims_) && DimsUtils::areEqual(outputDims[0], out_dims_));
        UNUSEDR(inputDims); UNUSEDR(nbInputs); UNUSEDR(outputDims); UNUSEDR(nbOutputs);
        logDims(log_, info_.name, "InDims : ", in_dims_);
        logDims(log_, info_.name, "OutDims: ", out_dims_);
    }
    int initialize() override { return 0; }
    void terminate() override {}
    size_t getWorkspaceSize(int) const override { return 0; }
    int enqueue(int batchSize, const void* const* inputs, void** outputs, void*, cudaStream_t stream) override
    {
        const int64_t plane = static_cast<int64_t>(in_dims_.d[1]) * in_dims_.d[2] * in_dims_.d[3];
        return rt_slice_planes(RT_F32, inputs[0], outputs[0], batchSize, in_dims_.d[0], plane,
                               info_.slice_start, info_.slice_end, stream);
    }
    size_t getSerializationSize() override { return 7 * sizeof(int32_t); }
    void serialize(void* buffer) override
    {
        ByteWriter w;
        w.put<int32_t>(static_cast<int32_t>(StereoDnnPluginFactory::PluginType::kSlice));
        for (int i = 0; i < 4; i++) w.put<int32_t>(in_dims_.d[i]);
        w.put<int32_t>(info_.slice_start);
        w.put<int32_t>(info_.slice_end);
        memcpy(buffer, w.buf.data(), w.buf.size());
    }

private:
    OpInfo info_;
    Dims in_dims_{}, out_dims_{};
    ILogger& log_;
};

// ---------------------------------------------------------------------------------------------------------------
// Container.  Reference: lib/internal_utils.h:114-168, lib/internal_utils.cpp:119-255.
// ---------------------------------------------------------------------------------------------------------------
class PluginContainer : public IPluginContainer
{
public:
    explicit PluginContainer(ILogger& log) : log_(log) {}
    ~PluginContainer() override
    {
        for (auto* p : plugins_) delete p;
    }

    IPlugin* createEluPlugin(DataType data_type, std::string name) override
    {
        return keep(new EluPlugin(data_type, log_, name));
    }
    IPlugin* deserializeEluPlugin(const char* name, const void* data, size_t size) override
    {
        auto* p = new EluPlugin(name, data, size, log_);
        if (!p->valid()) {
            log_.log(ILogger::Severity::kERROR, (std::string(name ? name : "") + ": malformed serialised EluPlugin").c_str());
            delete p;
            return nullptr;
        }
        return keep(p);
    }
    IPlugin* createCostVolumePlugin(DataType data_type, CostVolumeType cv_type, int max_disparity, std::string name) override
    {
        return keep(new CostVolumePlugin(data_type, cv_type, max_disparity, log_, name));
    }
    IPlugin* deserializeCostVolumePlugin(const char* name, const void* data, size_t size) override
    {
        auto* p = new CostVolumePlugin(name, data, size, log_);
        if (!p->valid()) {
            log_.log(ILogger::Severity::kERROR, (std::string(name ? name : "") + ": malformed serialised CostVolumePlugin").c_str());
            delete p;
            return nullptr;
        }
        return keep(p);
    }
    IPlugin* createConv3DPlugin(Conv3DType conv_type, Dims kernel_dims, Dims stride_dims, Dims pad_start_dims,
                                Dims pad_end_dims, Weights kernel_weights, Weights bias_weights, std::string name) override
    {
        return keep(new Conv3DPluginBase(false, conv_type, kernel_dims, Dims{}, stride_dims, pad_start_dims, pad_end_dims,
                                         kernel_weights, bias_weights, log_, name));
    }
    IPlugin* createConv3DTransposePlugin(Conv3DType conv_type, Dims kernel_dims, Dims out_dims, Dims stride_dims,
                                         Dims pad_start_dims, Dims pad_end_dims, Weights kernel_weights,
                                         Weights bias_weights, std::string name) override
    {
        return keep(new Conv3DPluginBase(true, conv_type, kernel_dims, out_dims, stride_dims, pad_start_dims, pad_end_dims,
                                         kernel_weights, bias_weights, log_, name));
    }
    IPlugin* createTransformPlugin(Permutation permutation, std::string name) override
    {
        return keep(new TransformPlugin(permutation, log_, name));
    }
    IPlugin* createPaddingPlugin(DimsNCHW pad_start, DimsNCHW pad_end, std::string name) override
    {
        return keep(new PaddingPlugin(pad_start, pad_end, log_, name));
    }
    IPlugin* createSlicePlugin(Dims dims, Dims slice_start, Dims slice_end, std::string name) override
    {
        return keep(new SlicePlugin(dims, slice_start, slice_end, log_, name));
    }
    IPlugin* createSoftargmaxPlugin(DataType data_type, SoftargmaxType sm_type, std::string name) override
    {
        return keep(new SoftargmaxPlugin(data_type, sm_type, log_, name));
    }
    IPlugin* deserializeSoftargmaxPlugin(const char* name, const void* data, size_t size) override
    {
        auto* p = new SoftargmaxPlugin(name, data, size, log_);
        if (!p->valid()) {
            log_.log(ILogger::Severity::kERROR, (std::string(name ? name : "") + ": malformed serialised SoftargmaxPlugin").c_str());
            delete p;
            return nullptr;
        }
        return keep(p);
    }

private:
    IPlugin* keep(IPlugin* p)
    {
        std::lock_guard<std::mutex> lock(lock_);
        plugins_.push_back(p);
        return p;
    }
    ILogger& log_;
    std::mutex lock_;
    std::vector<IPlugin*> plugins_;
};

ILayer* addPluginLayer(INetworkDefinition& network, ITensor* const* inputs, int num_inputs, IPlugin* plugin)
{
    // IPluginExt plugins go through addPluginExt (reference: lib/internal_utils.cpp:127-134).
    auto ext = dynamic_cast<IPluginExt*>(plugin);
    return ext != nullptr ? network.addPluginExt(inputs, num_inputs, *ext) : network.addPlugin(inputs, num_inputs, *plugin);
}

}  // namespace

std::unique_ptr<IPluginContainer> IPluginContainer::create(ILogger& log)
{
    return std::make_unique<PluginContainer>(log);
}

ILayer* addElu(IPluginContainer& f, INetworkDefinition& network, ITensor& input, DataType data_type, const std::string& name)
{
    ITensor* in[] = {&input};
    return addPluginLayer(network, in, 1, f.createEluPlugin(data_type, name));
}

ILayer* addCostVolume(IPluginContainer& f, INetworkDefinition& network, ITensor& left_input, ITensor& right_input,
                      CostVolumeType cv_type, int max_disparity, DataType data_type, const std::string& name)
{
    ITensor* in[] = {&left_input, &right_input};
    return addPluginLayer(network, in, 2, f.createCostVolumePlugin(data_type, cv_type, max_disparity, name));
}

ILayer* addConv3D(IPluginContainer& f, INetworkDefinition& network, ITensor& input, Conv3DType conv_type,
                  Dims kernel_dims, Dims stride_dims, Dims pad_start_dims, Dims pad_end_dims,
                  Weights kernel_weights, Weights bias_weights, const std::string& name)
{
    ITensor* in[] = {&input};
    return addPluginLayer(network, in, 1, f.createConv3DPlugin(conv_type, kernel_dims, stride_dims, pad_start_dims,
                                                               pad_end_dims, kernel_weights, bias_weights, name));
}

ILayer* addConv3DTranspose(IPluginContainer& f, INetworkDefinition& network, ITensor& input, Conv3DType conv_type,
                           Dims kernel_dims, Dims out_dims, Dims stride_dims, Dims pad_start_dims, Dims pad_end_dims,
                           Weights kernel_weights, Weights bias_weights, const std::string& name)
{
    ITensor* in[] = {&input};
    return addPluginLayer(network, in, 1, f.createConv3DTransposePlugin(conv_type, kernel_dims, out_dims, stride_dims,
                                                                        pad_start_dims, pad_end_dims, kernel_weights,
                                                                        bias_weights, name));
}

ILayer* addSlice(IPluginContainer& f, INetworkDefinition& network, ITensor& input, Dims dims, Dims slice_start,
                 Dims slice_end, const std::string& name)
{
    ITensor* in[] = {&input};
    return addPluginLayer(network, in, 1, f.createSlicePlugin(dims, slice_start, slice_end, name));
}

ILayer* addTransform(IPluginContainer& f, INetworkDefinition& network, ITensor& input, Permutation permutation,
                     const std::string& name)
{
    ITensor* in[] = {&input};
    return addPluginLayer(network, in, 1, f.createTransformPlugin(permutation, name));
}

ILayer* addPad(IPluginContainer& f, INetworkDefinition& network, ITensor& input, DimsNCHW pad_start, DimsNCHW pad_end,
               const std::string& name)
{
    ITensor* in[] = {&input};
    return addPluginLayer(network, in, 1, f.createPaddingPlugin(pad_start, pad_end, name));
}

ILayer* addSoftargmax(IPluginContainer& f, INetworkDefinition& network, ITensor& input, SoftargmaxType sm_type,
                      DataType data_type, const std::string& name)
{
    ITensor* in[] = {&input};
    return addPluginLayer(network, in, 1, f.createSoftargmaxPlugin(data_type, sm_type, name));
}

StereoDnnPluginFactory::StereoDnnPluginFactory(IPluginContainer& container) : container_(container) {}

IPlugin* StereoDnnPluginFactory::createPlugin(const char* layerName, const void* serialData, size_t serialLength)
{
    if (serialData == nullptr || serialLength < sizeof(int32_t)) return nullptr;
    int32_t tag;
    memcpy(&tag, serialData, sizeof(tag));
    const char* rest = static_cast<const char*>(serialData) + sizeof(tag);
    const size_t rest_len = serialLength - sizeof(tag);
    switch (static_cast<PluginType>(tag)) {
        case PluginType::kElu:        return container_.deserializeEluPlugin(layerName, rest, rest_len);
        case PluginType::kCostVolume: return container_.deserializeCostVolumePlugin(layerName, rest, rest_len);
        case PluginType::kSoftargmax: return container_.deserializeSoftargmaxPlugin(layerName, rest, rest_len);
        case PluginType::kConv3D:
        case PluginType::kConv3DTranspose: {
            ByteReader r{rest, rest_len};
            const bool tr = static_cast<PluginType>(tag) == PluginType::kConv3DTranspose;
            Dims kd{}; kd.nbDims = 5;
            for (int i = 0; i < 5; i++) kd.d[i] = r.get<int32_t>();
            int od[4];
            for (int i = 0; i < 4; i++) od[i] = r.get<int32_t>();
            int st[3], ps[3], pe[3];
            for (int i = 0; i < 3; i++) st[i] = r.get<int32_t>();
            for (int i = 0; i < 3; i++) ps[i] = r.get<int32_t>();
            for (int i = 0; i < 3; i++) pe[i] = r.get<int32_t>();
            const DataType wt = static_cast<DataType>(r.get<int32_t>());
            const int64_t kcount = r.get<int64_t>(), bcount = r.get<int64_t>();
            const size_t es = wt == DataType::kHALF ? 2 : 4;
            if (!r.ok || (wt != DataType::kHALF && wt != DataType::kFLOAT) || kcount <= 0 || bcount < 0 ||
                r.left != static_cast<size_t>(kcount + bcount) * es)
                return nullptr;
            int64_t kvol = 1;
            for (int i = 0; i < 5; i++) { if (kd.d[i] <= 0) return nullptr; kvol *= kd.d[i]; }
            if (kvol != kcount) return nullptr;
            Weights kw{wt, r.p, kcount};
            Weights bw{wt, bcount > 0 ? r.p + kcount * es : nullptr, bcount};
            IPlugin* p = tr ? container_.createConv3DTransposePlugin(Conv3DType::kTensorFlow, kd, DimsNCHW(od[0], od[1], od[2], od[3]),
                                                                     DimsCHW(st[0], st[1], st[2]), DimsCHW(ps[0], ps[1], ps[2]),
                                                                     DimsCHW(pe[0], pe[1], pe[2]), kw, bw, layerName)
                            : container_.createConv3DPlugin(Conv3DType::kTensorFlow, kd, DimsCHW(st[0], st[1], st[2]),
                                                            DimsCHW(ps[0], ps[1], ps[2]), DimsCHW(pe[0], pe[1], pe[2]), kw, bw, layerName);
            if (auto* c = dynamic_cast<Conv3DPluginBase*>(p)) c->ownWeights();   // serialData is only valid during this call
            return p;
        }
        case PluginType::kTransform: {
            ByteReader r{rest, rest_len};
            Permutation perm{};
            for (int i = 0; i < 4; i++) perm.order[i] = r.get<int32_t>();
            if (!r.done()) return nullptr;
            return container_.createTransformPlugin(perm, layerName);
        }
        case PluginType::kPadding: {
            ByteReader r{rest, rest_len};
            const int planes = r.get<int32_t>();
            if (!r.done() || planes < 0) return nullptr;
            return container_.createPaddingPlugin(DimsNCHW(0, 0, 0, 0), DimsNCHW(planes, 0, 0, 0), layerName);
        }
        case PluginType::kSlice: {
            ByteReader r{rest, rest_len};
            int d[4];
            for (int i = 0; i < 4; i++) d[i] = r.get<int32_t>();
            const int s0 = r.get<int32_t>(), s1 = r.get<int32_t>();
            if (!r.done()) return nullptr;
            return container_.createSlicePlugin(DimsNCHW(d[0], d[1], d[2], d[3]), DimsNCHW(s0, 0, 0, 0),
                                                DimsNCHW(s1, d[1], d[2], d[3]), layerName);
        }
    }
    return nullptr;       // unknown tag
}

} }

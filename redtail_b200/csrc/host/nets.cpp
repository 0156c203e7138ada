// Network wiring of the stereo nets, written against the public plugin API (IPluginContainer + add* helpers) and the
// nvinfer1-compatible INetworkDefinition -- i.e. exactly the calls the reference's generated builders make
// (stereoDNN/sample_app/nvsmall_1025x321_net.cpp:21-427, nvtiny_513x161_net.cpp), but table-driven and
// size-generic instead of one generated file per resolution.  Layer names match the generated builders so weight
// files, profiles and logs line up.  Also holds the C-ABI of include/redtail_b200_engine.h.
#include <cuda_runtime_api.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "internal_utils.h"
#include "redtail_b200.h"
#include "NvCaffeParser.h"
#include "redtail_b200_engine.h"

extern "C" size_t redtail_serialize_network(void* network, int max_batch, int half2, void* buf, size_t buf_len);

using namespace nvinfer1;
using namespace redtail::tensorrt;

namespace {

thread_local std::string g_last_error;

struct WeightStore {
    // Owns the bytes; Weights views point into it (the engine consumes them during build).
    std::unordered_map<std::string, std::vector<uint8_t>> blobs;
    DataType type = DataType::kFLOAT;
    bool has(const std::string& n) const { return blobs.count(n) != 0; }
    Weights get(const std::string& n) const
    {
        auto it = blobs.find(n);
        if (it == blobs.end()) throw std::runtime_error("weights: missing entry '" + n + "'");
        const size_t es = type == DataType::kFLOAT ? 4 : 2;
        return Weights{type, it->second.data(), static_cast<int64_t>(it->second.size() / es)};
    }
    int64_t count(const std::string& n) const { return get(n).count; }
};

// Weight-file reader: cstring name, u32 count, count x (f32|f16)   (format: scripts/tensorrt_model_builder.py:52-60,
// reference reader sample_app/main.cpp:111-134).
bool readWeights(const std::string& path, DataType type, WeightStore& ws, std::string& err)
{
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) { err = "cannot open weights file " + path; return false; }
    ws.type = type;
    const size_t es = type == DataType::kFLOAT ? 4 : 2;
    while (f.peek() != std::ifstream::traits_type::eof()) {
        std::string name;
        std::getline(f, name, '\0');
        uint32_t count = 0;
        f.read(reinterpret_cast<char*>(&count), sizeof(count));
        if (!f || name.empty()) { err = "corrupt weights file " + path; return false; }
        std::vector<uint8_t> data(static_cast<size_t>(count) * es);
        f.read(reinterpret_cast<char*>(data.data()), data.size());
        if (!f) { err = "truncated weights file " + path; return false; }
        if (ws.blobs.count(name)) { err = "duplicate weight entry " + name; return false; }
        ws.blobs[name] = std::move(data);
    }
    return true;
}

class CapiLogger : public ILogger {
public:
    void log(Severity severity, const char* msg) override
    {
        if (severity <= Severity::kERROR) last_error = msg;
        if (severity <= Severity::kWARNING || verbose) fprintf(stderr, "[redtail_b200] %s\n", msg);
    }
    std::string last_error;
    bool verbose = getenv("REDTAIL_VERBOSE") != nullptr;
};

// NVSmall family: siamese 2-D towers (conv1 5x5 s2 + 4x conv 3x3, ELU after all but the last), concat cost volume,
// 3-D encoder conv3D_1,2 | 3ds,4,5 | 6ds,7,8, decoder deconv3D_1..3 with skips from conv3D_5 / conv3D_2, soft-argmin.
INetworkDefinition* buildNVSmallFamily(IBuilder& builder, IPluginContainer& pc, DimsCHW img, int max_disp,
                                       const WeightStore& w, ILogger& log)
{
    INetworkDefinition* net = builder.createNetwork();
    const DataType dt = DataType::kFLOAT;
    ITensor* feat[2];
    const char* sides[2] = {"left", "right"};
    for (int s = 0; s < 2; ++s) {
        const std::string p = sides[s];
        ITensor* x = net->addInput(sides[s], DataType::kFLOAT, img);
        auto* sc = net->addScale(*x, ScaleMode::kUNIFORM, w.get(p + "_scale_shift"), w.get(p + "_scale_scale"), w.get(p + "_scale_power"));
        sc->setName((p + "_scale").c_str());
        x = sc->getOutput(0);
        int cin = img.c();
        for (int i = 1; i <= 5; ++i) {
            const std::string nm = p + "_conv" + std::to_string(i);
            const int k = i == 1 ? 5 : 3;
            const int cout = static_cast<int>(w.count(nm + "_b"));
            if (w.count(nm + "_k") != static_cast<int64_t>(cout) * cin * k * k) throw std::runtime_error(nm + ": unexpected kernel size");
            auto* c = net->addConvolution(*x, cout, DimsHW{k, k}, w.get(nm + "_k"), w.get(nm + "_b"));
            c->setName(nm.c_str());
            c->setStride(i == 1 ? DimsHW{2, 2} : DimsHW{1, 1});
            c->setPadding(DimsHW{k / 2, k / 2});
            x = c->getOutput(0);
            if (i < 5) {
                auto* a = addElu(pc, *net, *x, dt, nm + "_act");
                a->setName((nm + "_act").c_str());
                x = a->getOutput(0);
            }
            cin = cout;
        }
        feat[s] = x;
    }
    auto* cv = addCostVolume(pc, *net, *feat[0], *feat[1], CostVolumeType::kDefault, max_disp, dt, "cost_vol");
    cv->setName("cost_vol");
    ITensor* x = cv->getOutput(0);                                   // [D, 2C, h, w]
    int cin = x->getDimensions().d[1];

    const char* enc[8] = {"1", "2", "3ds", "4", "5", "6ds", "7", "8"};
    ITensor* skip5 = nullptr;
    ITensor* skip2 = nullptr;
    for (int i = 0; i < 8; ++i) {
        const std::string nm = std::string("conv3D_") + enc[i];
        const bool ds = strstr(enc[i], "ds") != nullptr;
        const int cout = static_cast<int>(w.count(nm + "_b"));
        if (w.count(nm + "_k") != static_cast<int64_t>(cout) * 27 * cin) throw std::runtime_error(nm + ": unexpected kernel size");
        if (ds) {   // TF-SAME on an even D with stride 2 pads (0,1): one zero plane at the end (nvsmall_1025x321_net.cpp:212)
            auto* pad = addPad(pc, *net, *x, {0, 0, 0, 0}, {1, 0, 0, 0}, nm + "_pad");
            pad->setName((nm + "_pad").c_str());
            x = pad->getOutput(0);
        }
        auto* c = addConv3D(pc, *net, *x, Conv3DType::kTensorFlow, Dims{5, {cout, 3, cin, 3, 3}},
                            ds ? Dims{3, {2, 2, 2}} : Dims{3, {1, 1, 1}},
                            ds ? Dims{3, {0, 1, 1}} : Dims{3, {1, 1, 1}}, Dims{3, {1, 1, 1}},
                            w.get(nm + "_k"), w.get(nm + "_b"), nm);
        c->setName(nm.c_str());
        x = c->getOutput(0);                                         // [K, D, H, W]
        if (i != 7) {                                                // conv3D_8 feeds the decoder un-transposed (:316-323)
            auto* t = addTransform(pc, *net, *x, {1, 0, 2, 3}, nm + "_tran_transform");
            t->setName((nm + "_tran").c_str());
            x = t->getOutput(0);                                     // [D, K, H, W]
        }
        auto* a = addElu(pc, *net, *x, dt, nm + "_act");
        a->setName((nm + "_act").c_str());
        x = a->getOutput(0);
        if (i == 1) skip2 = x;
        if (i == 4) skip5 = x;
        cin = cout;
    }
    ITensor* skips[3] = {skip5, skip2, nullptr};
    for (int i = 1; i <= 3; ++i) {
        const std::string nm = "deconv3D_" + std::to_string(i);
        const int cout = static_cast<int>(w.count(nm + "_b"));
        if (w.count(nm + "_k") != static_cast<int64_t>(cin) * 27 * cout) throw std::runtime_error(nm + ": unexpected kernel size");
        Dims od;                                                     // [D+1, C, H, W] of the tensor being reconstructed
        if (skips[i - 1]) {
            const Dims sd = skips[i - 1]->getDimensions();
            od = Dims{4, {sd.d[0] + 1, cout, sd.d[2], sd.d[3]}};
        } else {
            od = Dims{4, {2 * max_disp + 1, cout, img.h(), img.w()}};
        }
        auto* dc = addConv3DTranspose(pc, *net, *x, Conv3DType::kTensorFlow, Dims{5, {cin, 3, cout, 3, 3}}, od,
                                      Dims{3, {2, 2, 2}}, Dims{3, {0, 1, 1}}, Dims{3, {0, 1, 1}},
                                      w.get(nm + "_k"), w.get(nm + "_b"), nm);
        dc->setName(nm.c_str());
        auto* sl = addSlice(pc, *net, *dc->getOutput(0), od, {4, {0, 0, 0, 0}}, {4, {od.d[0] - 1, od.d[1], od.d[2], od.d[3]}}, nm + "_slice");
        sl->setName((nm + "_slice_layer").c_str());
        x = sl->getOutput(0);                                        // [D, C, H, W]
        if (skips[i - 1]) {
            auto* add = net->addElementWise(*x, *skips[i - 1], ElementWiseOperation::kSUM);
            add->setName((nm + "_add_skip").c_str());
            auto* a = addElu(pc, *net, *add->getOutput(0), dt, nm + "_act");
            a->setName((nm + "_act").c_str());
            auto* t = addTransform(pc, *net, *a->getOutput(0), {1, 0, 2, 3}, nm + "_transform_transform");
            t->setName((nm + "_transform").c_str());
            x = t->getOutput(0);                                     // [C, D, H, W] for the next transposed conv
        }
        cin = cout;
    }
    auto* disp = addSoftargmax(pc, *net, *x, SoftargmaxType::kMin, dt, "disp_softargmax");
    disp->setName("disp");
    disp->getOutput(0)->setName("disp");
    net->markOutput(*disp->getOutput(0));
    (void)log;
    return net;
}

class LayerTimer : public IProfiler {
public:
    void reportLayerTime(const char* name, float ms) override { rows.emplace_back(name, ms); }
    std::vector<std::pair<std::string, float>> rows;
};

}  // namespace

struct rt_stereo_engine {
    CapiLogger log;
    std::unique_ptr<IPluginContainer> plugins;
    ICudaEngine* engine = nullptr;
    IExecutionContext* context = nullptr;
    int h = 0, w = 0, max_batch = 1;
    float* d_left = nullptr;      // staging for rt_stereo_execute_host / rt_stereo_execute_images
    float* d_right = nullptr;
    float* d_disp = nullptr;
    uint8_t* d_img = nullptr;     // 8-bit source images (left batch, then right batch)
    size_t d_img_bytes = 0;
    uint16_t* d_u16 = nullptr;
    cudaStream_t stream = nullptr;
    size_t device_bytes = 0;
};

extern "C" {

const char* rt_stereo_last_error(void) { return g_last_error.c_str(); }

int rt_stereo_create(const char* model, int height, int width, int max_disp, const char* weights_path,
                     int weights_dtype, int max_batch, rt_stereo_engine** out)
{
    if (!model || !weights_path || !out || height <= 0 || width <= 0 || max_disp <= 0 || max_batch <= 0) {
        g_last_error = "rt_stereo_create: bad argument";
        return RT_ERR_ARG;
    }
    if (strcmp(model, "nvsmall") != 0 && strcmp(model, "nvtiny") != 0) {
        g_last_error = std::string("rt_stereo_create: unknown model '") + model + "' (supported: nvsmall, nvtiny)";
        return RT_ERR_UNSUPPORTED;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_last_error = "rt_stereo_create: no CUDA device -- this engine has no CPU path";
        return RT_ERR_NO_DEVICE;
    }
    std::unique_ptr<rt_stereo_engine> e(new rt_stereo_engine());
    WeightStore ws;
    std::string err;
    if (!readWeights(weights_path, weights_dtype == RT_F16 ? DataType::kHALF : DataType::kFLOAT, ws, err)) {
        g_last_error = err;
        return RT_ERR_ARG;
    }
    e->h = height; e->w = width; e->max_batch = max_batch;
    e->plugins = IPluginContainer::create(e->log);
    IBuilder* builder = createInferBuilder(e->log);
    INetworkDefinition* net = nullptr;
    try {
        net = buildNVSmallFamily(*builder, *e->plugins, DimsCHW{3, height, width}, max_disp, ws, e->log);
    } catch (const std::exception& ex) {
        g_last_error = ex.what();
        builder->destroy();
        return RT_ERR_ARG;
    }
    builder->setMaxBatchSize(max_batch);
    builder->setMaxWorkspaceSize(static_cast<size_t>(1) << 30);
    e->engine = builder->buildCudaEngine(*net);
    net->destroy();
    builder->destroy();
    if (!e->engine) {
        g_last_error = "rt_stereo_create: engine build failed: " + e->log.last_error;
        return RT_ERR_UNSUPPORTED;
    }
    e->context = e->engine->createExecutionContext();
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_last_error = "cudaStreamCreate failed";
        rt_stereo_destroy(e.release());
        return RT_ERR_NO_DEVICE;
    }
    *out = e.release();
    return RT_OK;
}

size_t rt_stereo_serialize(const rt_stereo_engine* e, void* buf, size_t buf_len)
{
    if (!e || !e->engine) return 0;
    IHostMemory* m = e->engine->serialize();
    if (!m) { g_last_error = "rt_stereo_serialize: " + e->log.last_error; return 0; }
    const size_t n = m->size();
    if (buf && buf_len >= n) memcpy(buf, m->data(), n);
    m->destroy();
    return n;
}

int rt_stereo_deserialize(const void* plan, size_t plan_size, rt_stereo_engine** out)
{
    if (!plan || plan_size == 0 || !out) { g_last_error = "rt_stereo_deserialize: bad argument"; return RT_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_last_error = "rt_stereo_deserialize: no CUDA device -- this engine has no CPU path";
        return RT_ERR_NO_DEVICE;
    }
    std::unique_ptr<rt_stereo_engine> e(new rt_stereo_engine());
    e->plugins = IPluginContainer::create(e->log);
    StereoDnnPluginFactory factory(*e->plugins);
    IRuntime* runtime = createInferRuntime(e->log);
    e->engine = runtime->deserializeCudaEngine(plan, plan_size, &factory);
    runtime->destroy();
    if (!e->engine) {
        g_last_error = "rt_stereo_deserialize: " + e->log.last_error;
        return RT_ERR_UNSUPPORTED;
    }
    const int li = e->engine->getBindingIndex("left");
    if (li < 0 || e->engine->getBindingIndex("right") < 0 || e->engine->getBindingIndex("disp") < 0) {
        g_last_error = "rt_stereo_deserialize: plan has no left/right/disp bindings";
        rt_stereo_destroy(e.release());
        return RT_ERR_ARG;
    }
    const Dims d = e->engine->getBindingDimensions(li);
    e->h = d.d[1]; e->w = d.d[2]; e->max_batch = e->engine->getMaxBatchSize();
    e->context = e->engine->createExecutionContext();
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_last_error = "cudaStreamCreate failed";
        rt_stereo_destroy(e.release());
        return RT_ERR_NO_DEVICE;
    }
    *out = e.release();
    return RT_OK;
}

int rt_stereo_deserialize_batch(const void* plan, size_t plan_size, int max_batch, rt_stereo_engine** out)
{
    // plan header: 8-byte magic, int32 version, int32 max batch (engine.cpp, serializeNetwork)
    if (!plan || plan_size < 16 || max_batch <= 0 || !out) { g_last_error = "rt_stereo_deserialize_batch: bad argument"; return RT_ERR_ARG; }
    std::string copy(static_cast<const char*>(plan), plan_size);
    const int32_t mb = max_batch;
    memcpy(&copy[12], &mb, sizeof(mb));
    return rt_stereo_deserialize(copy.data(), copy.size(), out);
}

void rt_stereo_destroy(rt_stereo_engine* e)
{
    if (!e) return;
    if (e->context) e->context->destroy();
    if (e->engine) e->engine->destroy();
    cudaFree(e->d_left); cudaFree(e->d_right); cudaFree(e->d_disp); cudaFree(e->d_img); cudaFree(e->d_u16);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int rt_stereo_enqueue(rt_stereo_engine* e, int batch, const float* left, const float* right, float* disp, void* stream)
{
    if (!e || !left || !right || !disp || batch < 1 || batch > e->max_batch) return RT_ERR_ARG;
    void* bindings[3];
    bindings[e->engine->getBindingIndex("left")] = const_cast<float*>(left);
    bindings[e->engine->getBindingIndex("right")] = const_cast<float*>(right);
    bindings[e->engine->getBindingIndex("disp")] = disp;
    if (!e->context->enqueue(batch, bindings, static_cast<cudaStream_t>(stream), nullptr)) {
        g_last_error = "rt_stereo_enqueue: " + e->log.last_error;
        return RT_ERR_UNSUPPORTED;
    }
    return RT_OK;
}

int rt_stereo_execute_host(rt_stereo_engine* e, int batch, const float* left, const float* right, float* disp)
{
    if (!e || !left || !right || !disp || batch < 1 || batch > e->max_batch) return RT_ERR_ARG;
    const size_t in_bytes = static_cast<size_t>(3) * e->h * e->w * sizeof(float);
    const size_t out_bytes = static_cast<size_t>(e->h) * e->w * sizeof(float);
    if (!e->d_left) {
        cudaError_t err = cudaMalloc(reinterpret_cast<void**>(&e->d_left), in_bytes * e->max_batch);
        if (err == cudaSuccess) err = cudaMalloc(reinterpret_cast<void**>(&e->d_right), in_bytes * e->max_batch);
        if (err == cudaSuccess) err = cudaMalloc(reinterpret_cast<void**>(&e->d_disp), out_bytes * e->max_batch);
        if (err != cudaSuccess) return static_cast<int>(err);
    }
    cudaError_t err = cudaMemcpyAsync(e->d_left, left, in_bytes * batch, cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(e->d_right, right, in_bytes * batch, cudaMemcpyHostToDevice, e->stream);
    if (err != cudaSuccess) return static_cast<int>(err);
    const int rc = rt_stereo_enqueue(e, batch, e->d_left, e->d_right, e->d_disp, e->stream);
    if (rc != RT_OK) return rc;
    err = cudaMemcpyAsync(disp, e->d_disp, out_bytes * batch, cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    return static_cast<int>(err);
}

int rt_stereo_execute_images(rt_stereo_engine* e, int batch, const uint8_t* left, const uint8_t* right, int src_h, int src_w,
                             float* disp, uint16_t* disp_u16, float u16_scale)
{
    if (!e || !left || !right || (!disp && !disp_u16) || batch < 1 || batch > e->max_batch || src_h < e->h || src_w < e->w) return RT_ERR_ARG;
    const size_t in_bytes = static_cast<size_t>(3) * e->h * e->w * sizeof(float);
    const size_t out_elems = static_cast<size_t>(e->h) * e->w;
    const size_t img_bytes = static_cast<size_t>(src_h) * src_w * 3;
    cudaError_t err = cudaSuccess;
    if (!e->d_left) {
        err = cudaMalloc(reinterpret_cast<void**>(&e->d_left), in_bytes * e->max_batch);
        if (err == cudaSuccess) err = cudaMalloc(reinterpret_cast<void**>(&e->d_right), in_bytes * e->max_batch);
        if (err == cudaSuccess) err = cudaMalloc(reinterpret_cast<void**>(&e->d_disp), out_elems * sizeof(float) * e->max_batch);
        if (err != cudaSuccess) return static_cast<int>(err);
    }
    if (e->d_img_bytes < 2 * img_bytes * e->max_batch) {
        cudaFree(e->d_img);
        e->d_img = nullptr; e->d_img_bytes = 0;
        err = cudaMalloc(reinterpret_cast<void**>(&e->d_img), 2 * img_bytes * e->max_batch);
        if (err != cudaSuccess) return static_cast<int>(err);
        e->d_img_bytes = 2 * img_bytes * e->max_batch;
    }
    if (disp_u16 && !e->d_u16) {
        err = cudaMalloc(reinterpret_cast<void**>(&e->d_u16), out_elems * sizeof(uint16_t) * e->max_batch);
        if (err != cudaSuccess) return static_cast<int>(err);
    }
    uint8_t* dl = e->d_img;
    uint8_t* dr = e->d_img + img_bytes * e->max_batch;
    err = cudaMemcpyAsync(dl, left, img_bytes * batch, cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(dr, right, img_bytes * batch, cudaMemcpyHostToDevice, e->stream);
    if (err != cudaSuccess) return static_cast<int>(err);
    int rc = rt_preprocess_bgr8(dl, batch, src_h, src_w, static_cast<int64_t>(3) * src_w, e->d_left, e->h, e->w, e->stream);
    if (rc == RT_OK) rc = rt_preprocess_bgr8(dr, batch, src_h, src_w, static_cast<int64_t>(3) * src_w, e->d_right, e->h, e->w, e->stream);
    if (rc == RT_OK) rc = rt_stereo_enqueue(e, batch, e->d_left, e->d_right, e->d_disp, e->stream);
    if (rc != RT_OK) return rc;
    if (disp) err = cudaMemcpyAsync(disp, e->d_disp, out_elems * sizeof(float) * batch, cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess && disp_u16) {
        rc = rt_disparity_to_u16(e->d_disp, e->d_u16, static_cast<int64_t>(out_elems) * batch, u16_scale, e->stream);
        if (rc != RT_OK) return rc;
        err = cudaMemcpyAsync(disp_u16, e->d_u16, out_elems * sizeof(uint16_t) * batch, cudaMemcpyDeviceToHost, e->stream);
    }
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    return static_cast<int>(err);
}

int rt_stereo_profile(rt_stereo_engine* e, int batch, const float* left, const float* right, float* disp, char* buf, size_t buf_len)
{
    if (!e || !buf || buf_len == 0) return RT_ERR_ARG;
    LayerTimer timer;
    e->context->setProfiler(&timer);
    void* bindings[3];
    bindings[e->engine->getBindingIndex("left")] = const_cast<float*>(left);
    bindings[e->engine->getBindingIndex("right")] = const_cast<float*>(right);
    bindings[e->engine->getBindingIndex("disp")] = disp;
    const bool ok = e->context->execute(batch, bindings);
    e->context->setProfiler(nullptr);
    if (!ok) return RT_ERR_UNSUPPORTED;
    std::ostringstream s;
    for (auto& r : timer.rows) s << r.first << "\t" << r.second << "\n";
    const std::string str = s.str();
    strncpy(buf, str.c_str(), buf_len - 1);
    buf[buf_len - 1] = 0;
    return RT_OK;
}

int rt_stereo_num_layers(const rt_stereo_engine* e) { return e && e->engine ? e->engine->getNbLayers() : 0; }
size_t rt_stereo_device_bytes(const rt_stereo_engine* e) { return e && e->engine ? e->engine->getWorkspaceSize() : 0; }

// ---- Caffe classifier networks (include/redtail_b200_engine.h, rt_net_*) ----------------------------------------------------
}  // extern "C"

struct rt_net_engine {
    CapiLogger log;
    ICudaEngine* engine = nullptr;
    IExecutionContext* context = nullptr;
    int in_idx = 0, out_idx = 1;
    size_t in_elems = 0, out_elems = 0;
    int in_chw[3] = {0, 0, 0}, out_chw[3] = {0, 0, 0};
    int max_batch = 1;
    float* d_in = nullptr;
    float* d_out = nullptr;
    cudaStream_t stream = nullptr;
};

namespace {
int finishNetEngine(std::unique_ptr<rt_net_engine>& e, rt_net_engine** out)
{
    if (e->engine->getNbBindings() != 2 || !e->engine->bindingIsInput(0) || e->engine->bindingIsInput(1)) {
        g_last_error = "rt_net: the network must have exactly one input and one output";
        rt_net_destroy(e.release());
        return RT_ERR_UNSUPPORTED;
    }
    e->in_idx = 0; e->out_idx = 1;
    const Dims id = e->engine->getBindingDimensions(0), od = e->engine->getBindingDimensions(1);
    e->in_elems = 1; e->out_elems = 1;
    for (int i = 0; i < 3; ++i) {
        e->in_chw[i] = i < id.nbDims ? id.d[i] : 1; e->out_chw[i] = i < od.nbDims ? od.d[i] : 1;
        e->in_elems *= static_cast<size_t>(e->in_chw[i]); e->out_elems *= static_cast<size_t>(e->out_chw[i]);
    }
    e->max_batch = e->engine->getMaxBatchSize();
    e->context = e->engine->createExecutionContext();
    if (!e->context || cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_last_error = "rt_net: context / stream creation failed";
        rt_net_destroy(e.release());
        return RT_ERR_NO_DEVICE;
    }
    *out = e.release();
    return RT_OK;
}
}  // namespace

extern "C" {

int rt_caffe_create(const char* prototxt_path, const char* caffemodel_path, const char* input_blob, const char* output_blob,
                    int max_batch, rt_net_engine** out)
{
    if (!prototxt_path || !caffemodel_path || !output_blob || !out || max_batch <= 0) { g_last_error = "rt_caffe_create: bad argument"; return RT_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_last_error = "rt_caffe_create: no CUDA device -- this engine has no CPU path";
        return RT_ERR_NO_DEVICE;
    }
    std::unique_ptr<rt_net_engine> e(new rt_net_engine());
    IBuilder* builder = createInferBuilder(e->log);
    INetworkDefinition* net = builder->createNetwork();
    nvcaffeparser1::ICaffeParser* parser = nvcaffeparser1::createCaffeParser();
    const nvcaffeparser1::IBlobNameToTensor* blobs = parser->parse(prototxt_path, caffemodel_path, *net, DataType::kFLOAT);
    ITensor* ob = blobs ? blobs->find(output_blob) : nullptr;
    if (!blobs || !ob || (input_blob && blobs->find(input_blob) == nullptr)) {
        g_last_error = !blobs ? std::string("rt_caffe_create: could not parse ") + prototxt_path + " / " + caffemodel_path
                              : std::string("rt_caffe_create: blob not found: ") + (ob ? input_blob : output_blob);
        net->destroy(); parser->destroy(); builder->destroy();
        return RT_ERR_ARG;
    }
    net->markOutput(*ob);
    builder->setMaxBatchSize(max_batch);
    builder->setMaxWorkspaceSize(static_cast<size_t>(1) << 30);
    e->engine = builder->buildCudaEngine(*net);
    net->destroy(); parser->destroy(); builder->destroy();
    if (!e->engine) { g_last_error = "rt_caffe_create: engine build failed: " + e->log.last_error; return RT_ERR_UNSUPPORTED; }
    return finishNetEngine(e, out);
}

void rt_net_destroy(rt_net_engine* e)
{
    if (!e) return;
    if (e->context) e->context->destroy();
    if (e->engine) e->engine->destroy();
    cudaFree(e->d_in); cudaFree(e->d_out);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int rt_net_dims(const rt_net_engine* e, int in_chw[3], int out_chw[3])
{
    if (!e || !in_chw || !out_chw) return RT_ERR_ARG;
    for (int i = 0; i < 3; ++i) { in_chw[i] = e->in_chw[i]; out_chw[i] = e->out_chw[i]; }
    return RT_OK;
}

int rt_net_enqueue(rt_net_engine* e, int batch, const float* in, float* out, void* stream)
{
    if (!e || !in || !out || batch < 1 || batch > e->max_batch) return RT_ERR_ARG;
    void* bindings[2];
    bindings[e->in_idx] = const_cast<float*>(in);
    bindings[e->out_idx] = out;
    return e->context->enqueue(batch, bindings, static_cast<cudaStream_t>(stream), nullptr) ? RT_OK : RT_ERR_UNSUPPORTED;
}

int rt_net_execute_host(rt_net_engine* e, int batch, const float* in, float* out)
{
    if (!e || !in || !out || batch < 1 || batch > e->max_batch) return RT_ERR_ARG;
    if (!e->d_in) {
        if (cudaMalloc(reinterpret_cast<void**>(&e->d_in), e->in_elems * sizeof(float) * e->max_batch) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void**>(&e->d_out), e->out_elems * sizeof(float) * e->max_batch) != cudaSuccess) return RT_ERR_NO_DEVICE;
    }
    cudaError_t err = cudaMemcpyAsync(e->d_in, in, e->in_elems * sizeof(float) * batch, cudaMemcpyHostToDevice, e->stream);
    if (err != cudaSuccess) return static_cast<int>(err);
    const int rc = rt_net_enqueue(e, batch, e->d_in, e->d_out, e->stream);
    if (rc != RT_OK) return rc;
    err = cudaMemcpyAsync(out, e->d_out, e->out_elems * sizeof(float) * batch, cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    return static_cast<int>(err);
}

int rt_net_profile(rt_net_engine* e, int batch, const float* in, float* out, char* buf, size_t buf_len)
{
    if (!e || !in || !out || !buf || buf_len == 0 || batch < 1 || batch > e->max_batch) return RT_ERR_ARG;
    LayerTimer timer;
    e->context->setProfiler(&timer);
    void* bindings[2];
    bindings[e->in_idx] = const_cast<float*>(in);
    bindings[e->out_idx] = out;
    const bool ok = e->context->execute(batch, bindings);
    e->context->setProfiler(nullptr);
    if (!ok) return RT_ERR_UNSUPPORTED;
    std::ostringstream s;
    for (auto& r : timer.rows) s << r.first << "\t" << r.second << "\n";
    const std::string str = s.str();
    strncpy(buf, str.c_str(), buf_len - 1);
    buf[buf_len - 1] = 0;
    return RT_OK;
}

size_t rt_net_serialize(const rt_net_engine* e, void* buf, size_t buf_len)
{
    if (!e || !e->engine) return 0;
    IHostMemory* m = e->engine->serialize();
    if (!m) return 0;
    const size_t n = m->size();
    if (buf && buf_len >= n) memcpy(buf, m->data(), n);
    m->destroy();
    return n;
}

int rt_net_deserialize(const void* plan, size_t plan_size, int max_batch, rt_net_engine** out)
{
    // plan header: 8-byte magic, int32 version, int32 max batch (engine.cpp, serializeNetwork)
    if (!plan || plan_size < 16 || !out) { g_last_error = "rt_net_deserialize: bad argument"; return RT_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_last_error = "rt_net_deserialize: no CUDA device"; return RT_ERR_NO_DEVICE; }
    std::string copy(static_cast<const char*>(plan), plan_size);
    if (max_batch > 0) { const int32_t mb = max_batch; memcpy(&copy[12], &mb, 4); }
    std::unique_ptr<rt_net_engine> e(new rt_net_engine());
    IRuntime* rt = createInferRuntime(e->log);
    e->engine = rt->deserializeCudaEngine(copy.data(), copy.size(), nullptr);
    rt->destroy();
    if (!e->engine) { g_last_error = "rt_net_deserialize: " + e->log.last_error; return RT_ERR_ARG; }
    return finishNetEngine(e, out);
}

int rt_net_num_layers(const rt_net_engine* e) { return e && e->engine ? e->engine->getNbLayers() : 0; }

// Host-only (no CUDA device needed): parse the Caffe model and write the network's plan without building an engine -- what
// rt_net_serialize would return.  Lets the CPU test-suite check the parser with its float64 graph-level checker.
size_t rt_caffe_dump_plan(const char* prototxt_path, const char* caffemodel_path, const char* output_blob, int max_batch, void* buf, size_t buf_len)
{
    if (!prototxt_path || !caffemodel_path || !output_blob) return 0;
    CapiLogger log;
    IBuilder* builder = createInferBuilder(log);
    INetworkDefinition* net = builder->createNetwork();
    nvcaffeparser1::ICaffeParser* parser = nvcaffeparser1::createCaffeParser();
    const nvcaffeparser1::IBlobNameToTensor* blobs = parser->parse(prototxt_path, caffemodel_path, *net, DataType::kFLOAT);
    ITensor* ob = blobs ? blobs->find(output_blob) : nullptr;
    size_t n = 0;
    if (ob) {
        net->markOutput(*ob);
        n = redtail_serialize_network(net, max_batch, 0, buf, buf_len);
    } else g_last_error = "rt_caffe_dump_plan: parse failed or output blob not found";
    net->destroy(); parser->destroy(); builder->destroy();
    return n;
}

}  // extern "C"

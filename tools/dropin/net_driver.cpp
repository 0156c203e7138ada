// Drives the reference's UNCHANGED generated network builders (stereoDNN/sample_app/*_net.cpp, declared in the
// reference's networks.h) through this repo's nvinfer1-compatible engine: the same call sequence as
// sample_app/main.cpp:176-315 minus OpenCV (raw CHW float32 .bin images in, raw float32 disparity out).
// Test infrastructure (tools/dropin); built only where /root/reference exists.
//
//   nvstereo_net_driver <nvsmall|nvtiny|resnet18|resnet18_2D> <width> <height> <weights.bin> <left.bin> <right.bin> <out.bin> [profile|plan|dump] [fp16]
//   fp16: <weights.bin> is a trt_weights_fp16.bin; the weights reach the builders as DataType::kHALF, ResNet18_2D gets
//         data_type = kHALF and the builder half2 mode, the other nets keep kFLOAT plugins -- sample_app/main.cpp:224-256.
//   plan: the engine is serialised, destroyed together with the weights and the plugin container, and re-created with
//         IRuntime::deserializeCudaEngine + StereoDnnPluginFactory (sample_app/main.cpp:207-220,270-275) before it runs.
#include <NvInfer.h>
#include <cuda_runtime_api.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "redtail_tensorrt_plugins.h"
#include "networks.h"

using namespace nvinfer1;
using namespace redtail::tensorrt;

extern "C" size_t redtail_serialize_network(void* network, int max_batch, int half2, void* buf, size_t buf_len);

struct Log : public ILogger {
    void log(Severity s, const char* msg) override { if (s <= Severity::kWARNING) std::cerr << "TRT: " << msg << std::endl; }
};
struct Prof : public IProfiler {
    void reportLayerTime(const char* n, float ms) override { printf("%-64.64s %8.3f ms\n", n, ms); total += ms; }
    float total = 0;
};

static std::vector<float> readBin(const char* path)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    const size_t n = f.tellg();
    f.seekg(0);
    std::vector<float> v(n / 4);
    f.read(reinterpret_cast<char*>(v.data()), n);
    return v;
}

static std::unordered_map<std::string, Weights> readWeights(const char* path, std::vector<std::vector<float>>& keep, bool fp16)
{
    std::unordered_map<std::string, Weights> w;
    std::ifstream f(path, std::ios::binary);
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    while (f.peek() != std::ifstream::traits_type::eof()) {
        std::string name;
        std::getline(f, name, '\0');
        uint32_t count = 0;
        f.read(reinterpret_cast<char*>(&count), 4);
        keep.emplace_back(fp16 ? (count + 1) / 2 : count);           // fp16 payloads are stored two per float slot
        f.read(reinterpret_cast<char*>(keep.back().data()), count * (fp16 ? 2ull : 4ull));
        w[name] = Weights{fp16 ? DataType::kHALF : DataType::kFLOAT, keep.back().data(), count};
    }
    return w;
}

int main(int argc, char** argv)
{
    if (argc < 8) { fprintf(stderr, "usage: %s model w h weights left.bin right.bin out.bin [profile]\n", argv[0]); return 1; }
    const std::string model = argv[1];
    const int w = atoi(argv[2]), h = atoi(argv[3]);
    Log log;
    std::vector<std::vector<float>> keep;
    bool fp16 = false;
    for (int i = 8; i < argc; ++i) fp16 = fp16 || !strcmp(argv[i], "fp16");
    const DataType data_type = fp16 ? DataType::kHALF : DataType::kFLOAT;
    auto weights = readWeights(argv[4], keep, fp16);
    auto left = readBin(argv[5]), right = readBin(argv[6]);
    if (left.size() != size_t(3) * h * w || right.size() != left.size()) { fprintf(stderr, "image size mismatch\n"); return 2; }

    auto container = IPluginContainer::create(log);
    IBuilder* builder = createInferBuilder(log);
    INetworkDefinition* net = nullptr;
    if (model == "nvsmall") net = createNVSmall1025x321Network(*builder, *container, DimsCHW{3, h, w}, weights, DataType::kFLOAT, log);
    else if (model == "nvtiny") net = createNVTiny513x161Network(*builder, *container, DimsCHW{3, h, w}, weights, DataType::kFLOAT, log);
    else if (model == "resnet18") net = createResNet18_1025x321Network(*builder, *container, DimsCHW{3, h, w}, weights, DataType::kFLOAT, log);
    else if (model == "resnet18_2D") net = createResNet18_2D_513x257Network(*builder, *container, DimsCHW{3, h, w}, weights, data_type, log);
    else { fprintf(stderr, "unknown model\n"); return 1; }
    if (argc > 8 && !strcmp(argv[8], "dump")) {
        // Host-only: write the plan of the network the reference's builder just described (no engine, no GPU) to <out.bin>.
        const size_t n = redtail_serialize_network(net, 1, fp16 ? 1 : 0, nullptr, 0);
        if (n == 0) { fprintf(stderr, "network is not serialisable\n"); return 5; }
        std::string blob(n, '\0');
        redtail_serialize_network(net, 1, fp16 ? 1 : 0, &blob[0], n);
        std::ofstream(argv[7], std::ios::binary).write(blob.data(), blob.size());
        printf("Network plan: %zu bytes, %d layers\n", n, net->getNbLayers());
        return 0;
    }
    builder->setMaxBatchSize(1);
    builder->setMaxWorkspaceSize(size_t(1) << 30);
    builder->setHalf2Mode(fp16);
    ICudaEngine* engine = builder->buildCudaEngine(*net);
    net->destroy();
    builder->destroy();
    if (!engine) { fprintf(stderr, "engine build failed\n"); return 3; }
    const bool plan_mode = argc > 8 && !strcmp(argv[8], "plan");
    std::unique_ptr<IPluginContainer> container2;
    if (plan_mode) {
        IHostMemory* plan = engine->serialize();
        if (!plan) { fprintf(stderr, "serialize failed\n"); return 5; }
        std::string blob(static_cast<const char*>(plan->data()), plan->size());
        plan->destroy();
        engine->destroy();
        container.reset();                                   // nothing of the build survives but the blob
        keep.clear(); keep.shrink_to_fit(); weights.clear();
        container2 = IPluginContainer::create(log);
        StereoDnnPluginFactory factory(*container2);
        IRuntime* runtime = createInferRuntime(log);
        engine = runtime->deserializeCudaEngine(blob.data(), blob.size(), &factory);
        runtime->destroy();
        if (!engine) { fprintf(stderr, "deserializeCudaEngine failed\n"); return 5; }
        printf("Plan: %zu bytes, engine rebuilt from it\n", blob.size());
    }
    if (engine->getNbBindings() != 3) { fprintf(stderr, "expected 3 bindings\n"); return 3; }
    void* buf[3];
    const int il = engine->getBindingIndex("left"), ir = engine->getBindingIndex("right"), io = engine->getBindingIndex("disp");
    std::vector<float> out(size_t(h) * w);
    cudaMalloc(&buf[il], left.size() * 4); cudaMalloc(&buf[ir], right.size() * 4); cudaMalloc(&buf[io], out.size() * 4);
    cudaMemcpy(buf[il], left.data(), left.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(buf[ir], right.data(), right.size() * 4, cudaMemcpyHostToDevice);
    IExecutionContext* ctx = engine->createExecutionContext();
    Prof prof;
    const bool profile_mode = argc > 8 && !strcmp(argv[8], "profile");
    if (profile_mode) ctx->setProfiler(&prof);
    const auto t0 = std::chrono::high_resolution_clock::now();
    const bool ok = ctx->execute(1, buf);
    const auto t1 = std::chrono::high_resolution_clock::now();
    if (!ok) { fprintf(stderr, "execute failed\n"); return 4; }
    printf("Host time: %.3f ms (%d engine steps)\n", std::chrono::duration<float, std::milli>(t1 - t0).count(), engine->getNbLayers());
    if (profile_mode) printf("All layers: %.3f ms\n", prof.total);
    cudaMemcpy(out.data(), buf[io], out.size() * 4, cudaMemcpyDeviceToHost);
    std::ofstream(argv[7], std::ios::binary).write(reinterpret_cast<const char*>(out.data()), out.size() * 4);
    ctx->destroy();
    engine->destroy();
    for (auto b : buf) cudaFree(b);
    return 0;
}

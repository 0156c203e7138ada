// Drives the reference's UNCHANGED TrailNet runtime -- ros/packages/caffe_ros/src/tensor_net.cpp + int8_calibrator.cpp, compiled
// where they lie against this repo's NvInfer.h / NvCaffeParser.h and the ros / boost / cv stand-ins of tools/dropin/include --
// the way ros/packages/caffe_ros/src/caffe_ros.cpp:88-134 does: loadNetwork(prototxt, caffemodel, "data", "out", fp32), then
// forward(raw 8-bit image, w, h, encoding) and read the six softmax outputs.
//   trailnet_driver preprocess <rgb8.bin> <w> <h> <out.f32>                      host only: tensor_net.cpp's preprocessImage()
//   trailnet_driver run <prototxt> <caffemodel> <rgb8.bin> <w> <h> [out.f32]     needs a GPU
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "caffe_ros/tensor_net.h"

static std::vector<unsigned char> readAll(const std::string& path)
{
    std::ifstream f(path, std::ios::binary);
    return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv)
{
    const std::string mode = argc > 1 ? argv[1] : "";
    if (mode == "preprocess" && argc >= 6) {
        const int w = std::stoi(argv[3]), h = std::stoi(argv[4]);
        auto raw = readAll(argv[2]);
        if (raw.size() != static_cast<size_t>(w) * h * 3) { std::cerr << "bad image size\n"; return 2; }
        cv::Mat img(h, w, CV_8UC3, raw.data());
        // the test node's defaults (caffe_ros.cpp:41-52): input format BGR, scale 1, shift 0; tests.cpp publishes rgb8
        cv::Mat chw = caffe_ros::preprocessImage(img, 320, 180, caffe_ros::InputFormat::BGR, "rgb8", 1.0f, 0.0f);
        std::ofstream(argv[5], std::ios::binary).write(reinterpret_cast<const char*>(chw.ptr<float>(0)), sizeof(float) * 3 * 180 * 320);
        return 0;
    }
    if (mode == "run" && argc >= 7) {
        const int w = std::stoi(argv[5]), h = std::stoi(argv[6]);
        auto raw = readAll(argv[4]);
        if (raw.size() != static_cast<size_t>(w) * h * 3) { std::cerr << "bad image size\n"; return 2; }
        caffe_ros::TensorNet net;
        net.loadNetwork(argv[2], argv[3], "data", "out", nvinfer1::DataType::kFLOAT, false);
        net.forward(raw.data(), w, h, "rgb8");
        const int n = net.getOutChannels() * net.getOutHeight() * net.getOutWidth();
        const float* out = net.getOutput();
        for (int i = 0; i < n; ++i) std::cout << out[i] << (i + 1 < n ? " " : "\n");
        if (argc > 7) std::ofstream(argv[7], std::ios::binary).write(reinterpret_cast<const char*>(out), sizeof(float) * n);
        return n == 6 ? 0 : 4;
    }
    std::cerr << "usage: trailnet_driver preprocess <rgb8.bin> <w> <h> <out.f32> | run <prototxt> <caffemodel> <rgb8.bin> <w> <h> [out.f32]\n";
    return 1;
}

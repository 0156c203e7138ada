// Exercises the cv:: stand-in (tools/dropin/include/opencv2/opencv.hpp) with exactly the call sequences of the reference's
// sample_app/main.cpp:83-98 (readImgFile) and :317-330 (16-bit PNG output), so that tests/test_dropin_shim.py can compare
// them with real OpenCV (cv2).      shim_check <in.png> <w> <h> <out.f32> <out16.png>
#include <opencv2/opencv.hpp>

#include <cassert>
#include <fstream>
#include <string>
#include <vector>

static std::vector<float> readImgFile(const std::string& filename, int w, int h)
{
    auto img = cv::imread(filename);
    assert(img.data != nullptr);
    img.convertTo(img, CV_32F);
    cv::resize(img, img, cv::Size(w, h), 0, 0, cv::INTER_AREA);
    cv::cvtColor(img, img, CV_BGR2RGB);
    cv::Mat res = img.reshape(1, w * h).t();
    res /= 255.0;
    return std::vector<float>(res.ptr<float>(0), res.ptr<float>(0) + w * h * 3);
}

int main(int argc, char** argv)
{
    if (argc < 6) return 1;
    const int w = std::stoi(argv[2]), h = std::stoi(argv[3]);
    auto v = readImgFile(argv[1], w, h);
    std::ofstream(argv[4], std::ios::binary).write(reinterpret_cast<const char*>(v.data()), v.size() * 4);
    // output side: treat the red plane (values in [0,1]) scaled to [0, 300) as a disparity map
    std::vector<float> disp(v.begin(), v.begin() + w * h);
    for (auto& d : disp) d *= 300.f;
    auto img_f = cv::Mat(h, w, CV_32F, disp.data());
    img_f *= 256;
    cv::Mat img_u16;
    img_f.convertTo(img_u16, CV_16U);
    return cv::imwrite(argv[5], img_u16) ? 0 : 2;
}

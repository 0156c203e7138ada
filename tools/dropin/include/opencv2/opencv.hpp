// Minimal cv:: stand-in for the two OpenCV calls of the reference's tests/tests_main.cpp (fp16 host conversion of test
// vectors, :200-202,237-239 -- dead code there, since the harness always feeds fp32).  Test infrastructure only.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#define CV_16S 3
#define CV_32F 5

namespace cv {

typedef unsigned char uchar;

class Mat {
public:
    Mat() {}
    explicit Mat(const std::vector<float>& v) : rows(static_cast<int>(v.size())), cols(1), type_(CV_32F)
    {
        store_.resize(v.size() * 4);
        std::memcpy(store_.data(), v.data(), store_.size());
        data = store_.data();
    }
    Mat(int r, int c, int type, void* ptr) : data(static_cast<uchar*>(ptr)), rows(r), cols(c), type_(type) {}
    int type() const { return type_; }
    size_t total() const { return static_cast<size_t>(rows) * cols; }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        store_.assign(total() * (type == CV_32F ? 4 : 2), 0);
        data = store_.data();
    }
    uchar* data = nullptr;
    int rows = 0, cols = 0;
private:
    int type_ = CV_32F;
    std::vector<uchar> store_;
};

inline uint16_t f2h(float f)
{
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = static_cast<int32_t>((x >> 23) & 0xFF) - 127 + 15;
    uint32_t man = x & 0x7FFFFFu;
    if (exp >= 31) return static_cast<uint16_t>(sign | 0x7C00u);
    if (exp <= 0) {
        if (exp < -10) return static_cast<uint16_t>(sign);
        man |= 0x800000u;
        const int shift = 14 - exp;
        uint32_t h = man >> shift;
        if ((man >> (shift - 1)) & 1u) ++h;
        return static_cast<uint16_t>(sign | h);
    }
    uint32_t h = (static_cast<uint32_t>(exp) << 10) | (man >> 13);
    if (man & 0x1000u) ++h;
    return static_cast<uint16_t>(sign | h);
}
inline float h2f(uint16_t h)
{
    const uint32_t sign = (h & 0x8000u) << 16, exp = (h >> 10) & 0x1F, man = h & 0x3FF;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else { int e = -1; uint32_t m = man; do { m <<= 1; ++e; } while (!(m & 0x400)); bits = sign | ((127 - 15 - e) << 23) | ((m & 0x3FF) << 13); }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    float f; std::memcpy(&f, &bits, 4); return f;
}

inline void convertFp16(const Mat& src, Mat& dst)
{
    if (src.type() == CV_32F) {
        dst.create(src.rows, src.cols, CV_16S);
        for (size_t i = 0; i < src.total(); ++i) reinterpret_cast<uint16_t*>(dst.data)[i] = f2h(reinterpret_cast<const float*>(src.data)[i]);
    } else {
        dst.create(src.rows, src.cols, CV_32F);
        for (size_t i = 0; i < src.total(); ++i) reinterpret_cast<float*>(dst.data)[i] = h2f(reinterpret_cast<const uint16_t*>(src.data)[i]);
    }
}

}  // namespace cv

// cv:: stand-in for the OpenCV calls of the reference's UNCHANGED sources (OpenCV's C++ headers are not in this image):
//   * tests/tests_main.cpp:200-202,237-239 -- cv::Mat over a float vector + cv::convertFp16 (dead code there);
//   * sample_app/main.cpp:83-98            -- readImgFile: imread, convertTo(CV_32F), resize(INTER_AREA), cvtColor(BGR2RGB),
//                                             reshape(1, w*h).t(), /= 255.0, ptr<float>();
//   * sample_app/main.cpp:317-330          -- Mat(h, w, CV_32F, ptr), *= 256, convertTo(CV_16U), imwrite(".png");
//   * ros/packages/caffe_ros/src/tensor_net.cpp:262-336 -- Mat over the caller's 8-bit image, cvtColor (RGB/BGRA -> BGR/RGB), convertTo,
//                                             resize(INTER_CUBIC), *= scale, += shift, reshape(1, w*h).t(), isContinuous, size().area().
// Test / drop-in infrastructure only (tools/dropin); the product's own image path is rt_preprocess_bgr8 /
// rt_disparity_to_u16 / rt_write_png16 (include/redtail_b200.h).  Semantics follow OpenCV 4: INTER_AREA uses OpenCV's area
// tables (same sums, same order -- the restatement is checked against cv2 in tests/test_dropin_shim.py), convertTo
// rounds half to even and saturates, PNGs are 8-bit RGB/RGBA/grey in, 16-bit grey out.  PNG inflate uses zlib (-lz).
#pragma once
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>       // OpenCV's core headers bring the standard containers in; caffe_ros/int8_calibrator.h:31 relies on that
#include <memory>
#include <sstream>     // OpenCV's core headers bring it in; sample_app/main.cpp:214 relies on that
#include <string>
#include <vector>

#define CV_8U 0
#define CV_16U 2
#define CV_16S 3
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_BGRA2BGR 1
#define CV_BGRA2RGB 3
#define CV_BGR2RGB 4
#define CV_RGB2BGR 4

namespace cv {

typedef unsigned char uchar;

enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum ColorConversionCodes { COLOR_BGRA2BGR = 1, COLOR_BGRA2RGB = 3, COLOR_BGR2RGB = 4, COLOR_RGB2BGR = 4 };
enum ImreadModes { IMREAD_UNCHANGED = -1, IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };

struct Size {
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
    int area() const { return width * height; }
    int width = 0, height = 0;
};

class Mat {
public:
    Mat() {}
    explicit Mat(const std::vector<float>& v) : rows(static_cast<int>(v.size())), cols(1), type_(CV_32F)
    {
        alloc();
        std::memcpy(data, v.data(), v.size() * 4);
    }
    // user data, not owned (sample_app/main.cpp:320)
    Mat(int r, int c, int type, void* ptr) : data(static_cast<uchar*>(ptr)), rows(r), cols(c), type_(type) {}
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize1() const { return depth() == CV_8U ? 1 : (depth() == CV_32F ? 4 : 2); }
    size_t total() const { return static_cast<size_t>(rows) * cols; }
    bool empty() const { return data == nullptr || total() == 0; }
    bool isContinuous() const { return true; }        // this stand-in never creates padded rows
    Size size() const { return Size(cols, rows); }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        alloc();
    }
    template <typename T> T* ptr(int row = 0) { return reinterpret_cast<T*>(data + static_cast<size_t>(row) * cols * channels() * elemSize1()); }
    template <typename T> const T* ptr(int row = 0) const { return reinterpret_cast<const T*>(data + static_cast<size_t>(row) * cols * channels() * elemSize1()); }

    // rtype: CV_32F / CV_16U / CV_8U depth (channels kept); alpha/beta as in OpenCV
    void convertTo(Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const
    {
        const int ddepth = rtype < 0 ? depth() : (rtype & 7);
        Mat out;
        out.create(rows, cols, CV_MAKETYPE(ddepth, channels()));
        const size_t n = total() * channels();
        for (size_t i = 0; i < n; ++i) {
            double v;
            switch (depth()) {
                case CV_8U: v = data[i]; break;
                case CV_16U: v = reinterpret_cast<const uint16_t*>(data)[i]; break;
                case CV_32F: v = reinterpret_cast<const float*>(data)[i]; break;
                default: std::abort();
            }
            if (alpha != 1.0 || beta != 0.0) v = static_cast<float>(v) * static_cast<float>(alpha) + static_cast<float>(beta);
            switch (ddepth) {
                case CV_32F: reinterpret_cast<float*>(out.data)[i] = static_cast<float>(v); break;
                case CV_16U: { const double r = std::nearbyint(v); reinterpret_cast<uint16_t*>(out.data)[i] = static_cast<uint16_t>(r < 0 ? 0 : (r > 65535 ? 65535 : r)); break; }
                case CV_8U: { const double r = std::nearbyint(v); out.data[i] = static_cast<uchar>(r < 0 ? 0 : (r > 255 ? 255 : r)); break; }
                default: std::abort();
            }
        }
        dst = out;
    }
    // same data viewed with `cn` channels and `new_rows` rows (continuous matrices only)
    Mat reshape(int cn, int new_rows = 0) const
    {
        Mat m = *this;
        const size_t scalars = total() * channels();
        if (cn == 0) cn = channels();
        if (new_rows == 0) new_rows = rows;
        m.type_ = CV_MAKETYPE(depth(), cn);
        m.rows = new_rows;
        m.cols = static_cast<int>(scalars / (static_cast<size_t>(new_rows) * cn));
        return m;
    }
    Mat t() const          // single-channel transpose
    {
        if (channels() != 1 || depth() != CV_32F) std::abort();
        Mat o;
        o.create(cols, rows, type_);
        const float* s = reinterpret_cast<const float*>(data);
        float* d = reinterpret_cast<float*>(o.data);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) d[static_cast<size_t>(c) * rows + r] = s[static_cast<size_t>(r) * cols + c];
        return o;
    }
    Mat& operator*=(double s)      // OpenCV: convertTo(*this, -1, s) -- fp32 multiply by float(s)
    {
        if (depth() != CV_32F) std::abort();
        float* p = reinterpret_cast<float*>(data);
        const float f = static_cast<float>(s);
        for (size_t i = 0, n = total() * channels(); i < n; ++i) p[i] *= f;
        return *this;
    }
    Mat& operator/=(double s) { return *this *= 1.0 / s; }
    Mat& operator+=(double s)      // OpenCV: add(*this, Scalar(s), *this) -- fp32 add of float(s)
    {
        if (depth() != CV_32F) std::abort();
        float* p = reinterpret_cast<float*>(data);
        const float f = static_cast<float>(s);
        for (size_t i = 0, n = total() * channels(); i < n; ++i) p[i] += f;
        return *this;
    }

    uchar* data = nullptr;
    int rows = 0, cols = 0;

private:
    void alloc()
    {
        store_ = std::make_shared<std::vector<uchar>>(total() * channels() * elemSize1(), 0);
        data = store_->data();
    }
    int type_ = CV_32F;
    std::shared_ptr<std::vector<uchar>> store_;       // shared like cv::Mat's reference-counted buffer
};

// ---- fp16 helpers of the unit-test harness -------------------------------------------------------------------------
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

// ---- PNG in (8-bit grey / RGB / RGBA, non-interlaced) -> CV_8UC3 BGR, like cv::imread's default flag ------------------
inline Mat imread(const std::string& filename, int = 1)
{
    Mat none;
    FILE* f = std::fopen(filename.c_str(), "rb");
    if (!f) return none;
    std::vector<uchar> raw;
    uchar buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) raw.insert(raw.end(), buf, buf + n);
    std::fclose(f);
    static const uchar sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (raw.size() < 33 || std::memcmp(raw.data(), sig, 8) != 0) return none;
    auto be32 = [&](size_t at) { return (uint32_t(raw[at]) << 24) | (uint32_t(raw[at + 1]) << 16) | (uint32_t(raw[at + 2]) << 8) | raw[at + 3]; };
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uchar> idat;
    for (size_t pos = 8; pos + 12 <= raw.size();) {
        const uint32_t len = be32(pos);
        if (pos + 12 + len > raw.size()) return none;
        const char* typ = reinterpret_cast<const char*>(&raw[pos + 4]);
        if (!std::memcmp(typ, "IHDR", 4)) { w = be32(pos + 8); h = be32(pos + 12); depth = raw[pos + 16]; ctype = raw[pos + 17]; interlace = raw[pos + 20]; }
        else if (!std::memcmp(typ, "IDAT", 4)) idat.insert(idat.end(), raw.begin() + pos + 8, raw.begin() + pos + 8 + len);
        pos += 12 + len;
    }
    const int cn = ctype == 0 ? 1 : (ctype == 2 ? 3 : (ctype == 6 ? 4 : 0));
    if (w == 0 || h == 0 || depth != 8 || cn == 0 || interlace != 0) return none;
    const size_t line = static_cast<size_t>(w) * cn;
    std::vector<uchar> px((line + 1) * h);
    uLongf out_len = px.size();
    if (uncompress(px.data(), &out_len, idat.data(), idat.size()) != Z_OK || out_len != px.size()) return none;
    // undo the scanline filters in place
    std::vector<uchar> prev(line, 0), cur(line);
    Mat img;
    img.create(static_cast<int>(h), static_cast<int>(w), CV_8UC3);
    for (uint32_t y = 0; y < h; ++y) {
        const uchar* s = &px[y * (line + 1)];
        const int ft = s[0];
        for (size_t i = 0; i < line; ++i) {
            const int a = i >= static_cast<size_t>(cn) ? cur[i - cn] : 0, b = prev[i], c = i >= static_cast<size_t>(cn) ? prev[i - cn] : 0;
            int pred = 0;
            switch (ft) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                default: return none;
            }
            cur[i] = static_cast<uchar>(s[1 + i] + pred);
        }
        uchar* d = img.ptr<uchar>(static_cast<int>(y));
        for (uint32_t x = 0; x < w; ++x) {
            const uchar* p = &cur[static_cast<size_t>(x) * cn];
            const uchar r = p[0], g = cn >= 3 ? p[1] : p[0], bl = cn >= 3 ? p[2] : p[0];
            d[3 * x] = bl; d[3 * x + 1] = g; d[3 * x + 2] = r;
        }
        prev.swap(cur);
    }
    return img;
}

// ---- cv::resize, INTER_AREA on CV_32F images, down-scaling or identity (modules/imgproc/src/resize.cpp: computeResizeAreaTab,
// ResizeArea_): horizontal reduction per source row, then vertical, fp32 weights, ascending order ---------------------------
namespace detail {
struct AreaEntry { int si; float w; };
inline std::vector<std::vector<AreaEntry>> areaTab(int ssize, int dsize)
{
    const double scale = static_cast<double>(ssize) / dsize;
    std::vector<std::vector<AreaEntry>> tab(dsize);
    for (int dx = 0; dx < dsize; ++dx) {
        const double fsx1 = dx * scale, cell = std::min(scale, ssize - fsx1), fsx2 = fsx1 + cell;
        int sx1 = static_cast<int>(std::ceil(fsx1)), sx2 = static_cast<int>(std::floor(fsx2));
        sx2 = std::min(sx2, ssize - 1);
        sx1 = std::min(sx1, sx2);
        if (sx1 - fsx1 > 1e-3) tab[dx].push_back({sx1 - 1, static_cast<float>((sx1 - fsx1) / cell)});
        for (int sx = sx1; sx < sx2; ++sx) tab[dx].push_back({sx, static_cast<float>(1.0 / cell)});
        if (fsx2 - sx2 > 1e-3) tab[dx].push_back({sx2, static_cast<float>(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell)});
    }
    return tab;
}
}  // namespace detail

// INTER_CUBIC on CV_32F images (modules/imgproc/src/resize.cpp: interpolateCubic with A = -0.75, 4 taps, replicated border;
// horizontal pass per source row, then vertical, all in fp32) -- what tensor_net.cpp:327 applies to the camera frame.
namespace detail {
struct CubicTap { int s[4]; float w[4]; };
inline std::vector<CubicTap> cubicTab(int ssize, int dsize)
{
    const double scale = static_cast<double>(ssize) / dsize;
    std::vector<CubicTap> tab(dsize);
    for (int d = 0; d < dsize; ++d) {
        float fx = static_cast<float>((d + 0.5) * scale - 0.5);
        const int sx = static_cast<int>(std::floor(fx));
        fx -= sx;
        const float A = -0.75f;
        CubicTap t;
        t.w[0] = ((A * (fx + 1) - 5 * A) * (fx + 1) + 8 * A) * (fx + 1) - 4 * A;
        t.w[1] = ((A + 2) * fx - (A + 3)) * fx * fx + 1;
        t.w[2] = ((A + 2) * (1 - fx) - (A + 3)) * (1 - fx) * (1 - fx) + 1;
        t.w[3] = 1.f - t.w[0] - t.w[1] - t.w[2];
        for (int k = 0; k < 4; ++k) t.s[k] = std::min(std::max(sx - 1 + k, 0), ssize - 1);
        tab[d] = t;
    }
    return tab;
}
}  // namespace detail

inline void resizeCubic(const Mat& src, Mat& dst, Size dsize)
{
    const int cn = src.channels(), sh = src.rows, dw = dsize.width, dh = dsize.height;
    const auto xt = detail::cubicTab(src.cols, dw), yt = detail::cubicTab(sh, dh);
    std::vector<float> tmp(static_cast<size_t>(sh) * dw * cn);
    for (int y = 0; y < sh; ++y) {
        const float* s = src.ptr<float>(y);
        float* t = &tmp[static_cast<size_t>(y) * dw * cn];
        for (int dx = 0; dx < dw; ++dx)
            for (int c = 0; c < cn; ++c) {
                const detail::CubicTap& k = xt[dx];
                t[dx * cn + c] = s[k.s[0] * cn + c] * k.w[0] + s[k.s[1] * cn + c] * k.w[1] + s[k.s[2] * cn + c] * k.w[2] + s[k.s[3] * cn + c] * k.w[3];
            }
    }
    Mat out;
    out.create(dh, dw, src.type());
    for (int dy = 0; dy < dh; ++dy) {
        float* d = out.ptr<float>(dy);
        const detail::CubicTap& t = yt[dy];
        const float* r0 = &tmp[static_cast<size_t>(t.s[0]) * dw * cn];
        const float* r1 = &tmp[static_cast<size_t>(t.s[1]) * dw * cn];
        const float* r2 = &tmp[static_cast<size_t>(t.s[2]) * dw * cn];
        const float* r3 = &tmp[static_cast<size_t>(t.s[3]) * dw * cn];
        for (int i = 0; i < dw * cn; ++i) d[i] = r0[i] * t.w[0] + r1[i] * t.w[1] + r2[i] * t.w[2] + r3[i] * t.w[3];
    }
    dst = out;
}

inline void resize(const Mat& src, Mat& dst, Size dsize, double = 0, double = 0, int interpolation = INTER_LINEAR)
{
    if (interpolation == INTER_CUBIC && src.depth() == CV_32F) { resizeCubic(src, dst, dsize); return; }
    if (interpolation != INTER_AREA || src.depth() != CV_32F || dsize.width > src.cols || dsize.height > src.rows) {
        std::fprintf(stderr, "opencv shim: only INTER_AREA down-scaling and INTER_CUBIC of CV_32F images are implemented\n");
        std::abort();
    }
    const int cn = src.channels(), sw = src.cols, sh = src.rows, dw = dsize.width, dh = dsize.height;
    const auto xt = detail::areaTab(sw, dw), yt = detail::areaTab(sh, dh);
    std::vector<float> tmp(static_cast<size_t>(sh) * dw * cn, 0.f);
    for (int y = 0; y < sh; ++y) {
        const float* s = src.ptr<float>(y);
        float* t = &tmp[static_cast<size_t>(y) * dw * cn];
        for (int dx = 0; dx < dw; ++dx)
            for (const auto& e : xt[dx])
                for (int c = 0; c < cn; ++c) t[dx * cn + c] += s[e.si * cn + c] * e.w;
    }
    Mat out;
    out.create(dh, dw, src.type());
    for (int dy = 0; dy < dh; ++dy) {
        float* d = out.ptr<float>(dy);
        for (const auto& e : yt[dy]) {
            const float* t = &tmp[static_cast<size_t>(e.si) * dw * cn];
            for (int i = 0; i < dw * cn; ++i) d[i] += t[i] * e.w;
        }
    }
    dst = out;
}

inline void cvtColor(const Mat& src, Mat& dst, int code)
{
    if ((code == CV_BGRA2BGR || code == CV_BGRA2RGB) && src.channels() == 4) {      // drop alpha (and swap R/B for BGRA2RGB)
        Mat o;
        o.create(src.rows, src.cols, CV_MAKETYPE(src.depth(), 3));
        const size_t es4 = src.elemSize1(), n4 = src.total();
        for (size_t i = 0; i < n4; ++i)
            for (int c = 0; c < 3; ++c)
                std::memcpy(o.data + (3 * i + c) * es4, src.data + (4 * i + (code == CV_BGRA2RGB ? 2 - c : c)) * es4, es4);
        dst = o;
        return;
    }
    if (code != CV_BGR2RGB || src.channels() != 3) std::abort();
    Mat out;
    out.create(src.rows, src.cols, src.type());
    const size_t es = src.elemSize1(), n = src.total();
    for (size_t i = 0; i < n; ++i) {
        std::memcpy(out.data + (3 * i + 0) * es, src.data + (3 * i + 2) * es, es);
        std::memcpy(out.data + (3 * i + 1) * es, src.data + (3 * i + 1) * es, es);
        std::memcpy(out.data + (3 * i + 2) * es, src.data + (3 * i + 0) * es, es);
    }
    dst = out;
}

// ---- PNG out: CV_16U single channel (KITTI-style disparity), stored deflate blocks ---------------------------------
inline bool imwrite(const std::string& filename, const Mat& img)
{
    if (img.depth() != CV_16U || img.channels() != 1) return false;
    std::vector<uchar> png = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    auto put32 = [](std::vector<uchar>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); };
    auto chunk = [&](const char* type, const std::vector<uchar>& d) {
        put32(png, static_cast<uint32_t>(d.size()));
        const size_t at = png.size();
        png.insert(png.end(), type, type + 4);
        png.insert(png.end(), d.begin(), d.end());
        put32(png, static_cast<uint32_t>(crc32(0L, png.data() + at, static_cast<uInt>(png.size() - at))));
    };
    std::vector<uchar> ihdr;
    put32(ihdr, img.cols); put32(ihdr, img.rows);
    ihdr.push_back(16); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk("IHDR", ihdr);
    std::vector<uchar> rawpx;
    for (int y = 0; y < img.rows; ++y) {
        rawpx.push_back(0);
        const uint16_t* r = img.ptr<uint16_t>(y);
        for (int x = 0; x < img.cols; ++x) { rawpx.push_back(r[x] >> 8); rawpx.push_back(r[x] & 0xFF); }
    }
    uLongf zlen = compressBound(rawpx.size());
    std::vector<uchar> z(zlen);
    if (compress2(z.data(), &zlen, rawpx.data(), rawpx.size(), 3) != Z_OK) return false;
    z.resize(zlen);
    chunk("IDAT", z);
    chunk("IEND", {});
    FILE* f = std::fopen(filename.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(png.data(), 1, png.size(), f) == png.size();
    std::fclose(f);
    return ok;
}

}  // namespace cv

// Image side of the stereo path: what stereoDNN/sample_app/main.cpp does on the host with
// OpenCV around every inference, as device kernels.
//
//   rt_preprocess_bgr8     readImgFile (sample_app/main.cpp:83-98): 8-bit BGR HWC image -> float -> cv::resize INTER_AREA
//                          -> BGR->RGB -> HWC->CHW -> * (1/255): one kernel, network input [3,h,w] fp32 in [0,1].
//   rt_disparity_to_u16    the output side (sample_app/main.cpp:317-330): disparity * scale (256, x width for ResNet18_2D),
//                          saturating round to uint16 -- the payload of the KITTI-style 16-bit PNG.
//   rt_write_png16         host: that payload as a 16-bit greyscale PNG (cv::imwrite of a CV_16U Mat), no libpng / zlib:
//                          stored deflate blocks, so the file is 2*h*w bytes plus headers.
//
// INTER_AREA follows OpenCV's area tables exactly (modules/imgproc/src/resize.cpp, computeResizeAreaTab + ResizeArea_):
// for a destination index dx over a source of `ssize` pixels, scale = ssize/dsize, cell = min(scale, ssize - dx*scale),
// the source interval [dx*scale, dx*scale + cell) contributes a fractional first pixel, whole pixels of weight 1/cell and
// a fractional last pixel; rows are reduced horizontally first, then vertically, in fp32, ascending order -- the same
// sums in the same order as cv::resize on a CV_32F image (checked against cv2 in tests/test_gpu_preprocess.py).
// Only down-scaling / identity (scale >= 1 on both axes) is implemented: that is what the apps do (1242x375 -> 1025x321).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace rt {
namespace {

struct AreaSpan { int first; int count; float w_first, w_mid, w_last; int has_first, has_last; };

__host__ __device__ inline AreaSpan area_span(int dx, int ssize, double scale) {
    const double fsx1 = dx * scale;
    const double cell = fmin(scale, ssize - fsx1);
    const double fsx2 = fsx1 + cell;
    int sx1 = static_cast<int>(ceil(fsx1)), sx2 = static_cast<int>(floor(fsx2));
    sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
    sx1 = sx1 < sx2 ? sx1 : sx2;
    AreaSpan s;
    s.has_first = (sx1 - fsx1 > 1e-3) ? 1 : 0;
    s.has_last = (fsx2 - sx2 > 1e-3) ? 1 : 0;
    s.first = sx1;
    s.count = sx2 - sx1;
    s.w_first = static_cast<float>((sx1 - fsx1) / cell);
    s.w_mid = static_cast<float>(1.0 / cell);
    s.w_last = static_cast<float>(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell);
    return s;
}

// One thread per destination pixel (all three channels).  Source rows are 8-bit and tiny next to the network's tensors
// (1.4 MB per 1242x375 image): the kernel is latency-, not bandwidth-bound, and takes a few microseconds.
__global__ void __launch_bounds__(256)
preprocess_bgr8_kernel(const uint8_t* __restrict__ src, int sh, int sw, long long pitch, long long img_stride,
                       float* __restrict__ dst, int dh, int dw, double scale_x, double scale_y) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    const int dy = blockIdx.y;
    const int n = blockIdx.z;
    if (dx >= dw) return;
    const AreaSpan xs = area_span(dx, sw, scale_x), ys = area_span(dy, sh, scale_y);
    const uint8_t* img = src + n * img_stride;
    float acc[3] = {0.f, 0.f, 0.f};
    // vertical reduction of horizontally reduced rows, both in OpenCV's order; __fmul_rn/__fadd_rn keep the products and
    // sums separately rounded like the CPU code (no FMA contraction)
    auto hrow = [&](int sy, float (&r)[3]) {
        const uint8_t* row = img + sy * pitch;
        r[0] = r[1] = r[2] = 0.f;
        if (xs.has_first)
            for (int c = 0; c < 3; ++c) r[c] = __fadd_rn(r[c], __fmul_rn(static_cast<float>(row[(xs.first - 1) * 3 + c]), xs.w_first));
        for (int k = 0; k < xs.count; ++k)
            for (int c = 0; c < 3; ++c) r[c] = __fadd_rn(r[c], __fmul_rn(static_cast<float>(row[(xs.first + k) * 3 + c]), xs.w_mid));
        if (xs.has_last)
            for (int c = 0; c < 3; ++c) r[c] = __fadd_rn(r[c], __fmul_rn(static_cast<float>(row[(xs.first + xs.count) * 3 + c]), xs.w_last));
    };
    float r[3];
    if (ys.has_first) {
        hrow(ys.first - 1, r);
        for (int c = 0; c < 3; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(r[c], ys.w_first));
    }
    for (int k = 0; k < ys.count; ++k) {
        hrow(ys.first + k, r);
        for (int c = 0; c < 3; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(r[c], ys.w_mid));
    }
    if (ys.has_last) {
        hrow(ys.first + ys.count, r);
        for (int c = 0; c < 3; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(r[c], ys.w_last));
    }
    const float inv = static_cast<float>(1.0 / 255.0);          // `res /= 255.0` = convertTo(.., 1/255.0) in fp32
    const long long plane = static_cast<long long>(dh) * dw;
    float* out = dst + n * 3 * plane + static_cast<long long>(dy) * dw + dx;
    out[0] = __fmul_rn(acc[2], inv);                            // R  (source order is B, G, R)
    out[plane] = __fmul_rn(acc[1], inv);
    out[2 * plane] = __fmul_rn(acc[0], inv);
}

__global__ void __launch_bounds__(256)
disparity_to_u16_kernel(const float* __restrict__ disp, uint16_t* __restrict__ out, long long count, float scale) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float v = rintf(disp[i] * scale);                  // cv::Mat::convertTo(CV_16U): saturate_cast<ushort>(cvRound(v))
        out[i] = static_cast<uint16_t>(fminf(fmaxf(v, 0.f), 65535.f));
    }
}

uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return crc;
}

void put_be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back(static_cast<uint8_t>(x >> 24)); v.push_back(static_cast<uint8_t>(x >> 16));
    v.push_back(static_cast<uint8_t>(x >> 8)); v.push_back(static_cast<uint8_t>(x));
}

void put_chunk(std::vector<uint8_t>& png, const char* type, const std::vector<uint8_t>& data) {
    put_be32(png, static_cast<uint32_t>(data.size()));
    const size_t at = png.size();
    png.insert(png.end(), type, type + 4);
    png.insert(png.end(), data.begin(), data.end());
    put_be32(png, crc32_update(0xFFFFFFFFu, png.data() + at, png.size() - at) ^ 0xFFFFFFFFu);
}

}  // namespace
}  // namespace rt

using namespace rt;

extern "C" {

int rt_preprocess_bgr8(const void* src, int n, int src_h, int src_w, int64_t src_pitch, void* dst, int dst_h, int dst_w,
                       void* stream) {
    if (!src || !dst || n < 0 || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || src_pitch < 3LL * src_w) return RT_ERR_ARG;
    if (dst_h > src_h || dst_w > src_w) return RT_ERR_UNSUPPORTED;      // INTER_AREA up-scaling is a different (bilinear) rule
    if (n == 0) return RT_OK;
    if (dst_h > 65535 || n > 65535) return RT_ERR_UNSUPPORTED;
    dim3 grid((dst_w + 255) / 256, dst_h, n);
    preprocess_bgr8_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const uint8_t*>(src), src_h, src_w, src_pitch,
                                                               src_pitch * src_h, static_cast<float*>(dst), dst_h, dst_w,
                                                               static_cast<double>(src_w) / dst_w, static_cast<double>(src_h) / dst_h);
    note_launch("preprocess_bgr8_area");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

int rt_disparity_to_u16(const void* disp, void* out, int64_t count, float scale, void* stream) {
    if (!disp || !out || count < 0) return RT_ERR_ARG;
    if (count == 0) return RT_OK;
    const int64_t blocks = ceil_div(count, 256);
    const int grid = static_cast<int>(blocks < 8LL * num_sms() ? blocks : 8LL * num_sms());
    disparity_to_u16_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(disp), static_cast<uint16_t*>(out), count, scale);
    note_launch("disparity_to_u16");
    RT_CHECK_LAUNCH();
    return RT_OK;
}

int rt_write_png16(const char* path, const uint16_t* pixels, int height, int width) {
    if (!path || !pixels || height <= 0 || width <= 0) return RT_ERR_ARG;
    std::vector<uint8_t> png = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, static_cast<uint32_t>(width));
    put_be32(ihdr, static_cast<uint32_t>(height));
    ihdr.push_back(16); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 16-bit greyscale
    put_chunk(png, "IHDR", ihdr);
    // raw scanlines: filter byte 0 + big-endian samples
    const size_t line = 1 + static_cast<size_t>(width) * 2;
    std::vector<uint8_t> raw(line * height);
    for (int y = 0; y < height; ++y) {
        uint8_t* r = raw.data() + y * line;
        r[0] = 0;
        for (int x = 0; x < width; ++x) {
            const uint16_t v = pixels[static_cast<size_t>(y) * width + x];
            r[1 + 2 * x] = static_cast<uint8_t>(v >> 8);
            r[2 + 2 * x] = static_cast<uint8_t>(v & 0xFF);
        }
    }
    // zlib stream of stored blocks (<= 65535 bytes each) + adler32
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (size_t i = 0; i < raw.size(); ++i) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
    for (size_t off = 0; off < raw.size(); off += 65535) {
        const size_t len = raw.size() - off < 65535 ? raw.size() - off : 65535;
        z.push_back(off + len == raw.size() ? 1 : 0);
        z.push_back(static_cast<uint8_t>(len & 0xFF)); z.push_back(static_cast<uint8_t>(len >> 8));
        z.push_back(static_cast<uint8_t>(~len & 0xFF)); z.push_back(static_cast<uint8_t>((~len >> 8) & 0xFF));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + len);
    }
    put_be32(z, (b << 16) | a);
    put_chunk(png, "IDAT", z);
    put_chunk(png, "IEND", std::vector<uint8_t>());
    FILE* f = fopen(path, "wb");
    if (!f) return RT_ERR_ARG;
    const size_t wr = fwrite(png.data(), 1, png.size(), f);
    fclose(f);
    return wr == png.size() ? RT_OK : RT_ERR_ARG;
}

}  // extern "C"

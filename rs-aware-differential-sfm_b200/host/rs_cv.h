// rs_cv.h -- the OpenCV types of the reference's hot-path signatures (cv::Mat 8UC3 / 8UC1 images,
// cv::Mat_<cv::Point_<double>> flow fields, cv::Vec3b / cv::Vec3f; camera.h, rsframe.h).
// A real OpenCV is used when present; otherwise this minimal, ref-counted stand-in with the same
// layout (row-major, interleaved channels) and the members the callers use (rows, cols, clone(),
// at<T>(y,x), data).  Image file IO: cv::imwrite / cv::imread handle 8-bit grey and BGR PNG files whose
// zlib stream uses stored blocks only -- which is what the reference asks OpenCV for
// (CV_IMWRITE_PNG_COMPRESSION 0; errorMeasure.cpp:201-206, main.cc:391-394, :549-553) and what this
// writer produces; decoding compressed PNGs (the shipped example frames) stays with the caller.
#pragma once

#if defined(RSDSFM_USE_REAL_OPENCV) || (defined(__has_include) && __has_include(<opencv2/core.hpp>))
#include <opencv2/core.hpp>
#else

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC3 21
#define CV_64FC2 14

namespace cv {

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<double> Point2d;

template <typename T, int N>
struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    Vec(T a, T b, T c) { static_assert(N == 3, "size"); val[0] = a; val[1] = b; val[2] = c; }
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
    bool operator==(const Vec &o) const { for (int i = 0; i < N; ++i) if (val[i] != o.val[i]) return false; return true; }
    bool operator!=(const Vec &o) const { return !(*this == o); }
};
typedef Vec<unsigned char, 3> Vec3b;
typedef Vec<float, 3> Vec3f;

struct Scalar { double v[4]; Scalar(double a = 0) { v[0] = a; v[1] = v[2] = v[3] = 0; } };

// Row-major, continuous, reference-counted pixel buffer (shallow copies like cv::Mat).
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;
    Mat() {}
    Mat(int r, int c, int type, const Scalar &s = Scalar(0)) { create(r, c, type); std::memset(data, (int)s.v[0], total_bytes()); }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        elem_ = (type == CV_8UC1) ? 1 : (type == CV_8UC3) ? 3 : (type == CV_32FC3) ? 12 : 16;
        buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c * elem_);
        data = buf_->data();
    }
    Mat clone() const { Mat m; m.create(rows, cols, type_); if (data) std::memcpy(m.data, data, total_bytes()); return m; }
    int type() const { return type_; }
    size_t elemSize() const { return (size_t)elem_; }
    size_t total_bytes() const { return (size_t)rows * cols * elem_; }
    bool empty() const { return data == nullptr; }
    template <typename T> T &at(int y, int x) { return *reinterpret_cast<T *>(data + ((size_t)y * cols + x) * elem_); }
    template <typename T> const T &at(int y, int x) const { return *reinterpret_cast<const T *>(data + ((size_t)y * cols + x) * elem_); }
    Mat &operator*=(double s) { if (s == 0.0 && data) std::memset(data, 0, total_bytes()); return *this; }   // `gs_image *= 0`
protected:
    int type_ = CV_8UC3, elem_ = 3;
    std::shared_ptr<std::vector<unsigned char>> buf_;
};

template <typename T>
class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) { rows = r; cols = c; type_ = -1; elem_ = (int)sizeof(T); buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c * sizeof(T)); data = buf_->data(); }
    T &operator()(int y, int x) { return *reinterpret_cast<T *>(data + ((size_t)y * cols + x) * sizeof(T)); }
    const T &operator()(int y, int x) const { return *reinterpret_cast<const T *>(data + ((size_t)y * cols + x) * sizeof(T)); }
};

// ---- PNG with stored (uncompressed) deflate blocks: 8-bit grey (8UC1) and BGR (8UC3, written as RGB)
namespace rs_png {
inline uint32_t crc32(const unsigned char *p, size_t n, uint32_t crc = 0)
{
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        ready = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return ~crc;
}
inline void put32(std::vector<unsigned char> &o, uint32_t v) { for (int s = 24; s >= 0; s -= 8) o.push_back((unsigned char)(v >> s)); }
inline uint32_t get32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void chunk(std::vector<unsigned char> &o, const char *tag, const std::vector<unsigned char> &body)
{
    put32(o, (uint32_t)body.size());
    const size_t at = o.size();
    o.insert(o.end(), tag, tag + 4);
    o.insert(o.end(), body.begin(), body.end());
    put32(o, crc32(o.data() + at, 4 + body.size()));
}
}  // namespace rs_png

inline bool imwrite(const std::string &path, const Mat &img, const std::vector<int> & = std::vector<int>())
{
    const int ch = (img.type() == CV_8UC1) ? 1 : (img.type() == CV_8UC3) ? 3 : 0;
    if (!ch || img.empty()) return false;
    // filtered scanlines (filter type 0), BGR -> RGB
    std::vector<unsigned char> raw;
    raw.reserve((size_t)img.rows * ((size_t)img.cols * ch + 1));
    for (int y = 0; y < img.rows; ++y) {
        raw.push_back(0);
        const unsigned char *row = img.data + (size_t)y * img.cols * ch;
        for (int x = 0; x < img.cols; ++x)
            for (int c = 0; c < ch; ++c) raw.push_back(row[(size_t)x * ch + (ch == 3 ? 2 - c : c)]);
    }
    // zlib container around stored blocks of at most 65535 bytes
    std::vector<unsigned char> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (size_t at = 0; at < raw.size() || at == 0;) {
        const size_t len = raw.size() - at < 65535 ? raw.size() - at : 65535;
        const bool last = at + len >= raw.size();
        z.push_back(last ? 1 : 0);
        z.push_back((unsigned char)(len & 0xff)); z.push_back((unsigned char)(len >> 8));
        z.push_back((unsigned char)(~len & 0xff)); z.push_back((unsigned char)((~len >> 8) & 0xff));
        for (size_t i = 0; i < len; ++i) { a = (a + raw[at + i]) % 65521u; b = (b + a) % 65521u; }
        z.insert(z.end(), raw.begin() + (std::ptrdiff_t)at, raw.begin() + (std::ptrdiff_t)(at + len));
        at += len;
        if (last) break;
    }
    rs_png::put32(z, (b << 16) | a);
    std::vector<unsigned char> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a}, hdr;
    rs_png::put32(hdr, (uint32_t)img.cols); rs_png::put32(hdr, (uint32_t)img.rows);
    hdr.push_back(8); hdr.push_back(ch == 3 ? 2 : 0); hdr.push_back(0); hdr.push_back(0); hdr.push_back(0);
    rs_png::chunk(out, "IHDR", hdr);
    rs_png::chunk(out, "IDAT", z);
    rs_png::chunk(out, "IEND", std::vector<unsigned char>());
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok;
}

// Reads back what imwrite wrote (8-bit grey / RGB, stored blocks, filter type 0); an empty Mat otherwise.
inline Mat imread(const std::string &path, int = 1)
{
    Mat none;
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return none;
    std::vector<unsigned char> d;
    unsigned char buf[65536];
    for (size_t n; (n = std::fread(buf, 1, sizeof buf, f)) > 0;) d.insert(d.end(), buf, buf + n);
    std::fclose(f);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (d.size() < 8 || std::memcmp(d.data(), sig, 8) != 0) return none;
    int w = 0, h = 0, ch = 0;
    std::vector<unsigned char> z;
    for (size_t at = 8; at + 12 <= d.size();) {
        const uint32_t len = rs_png::get32(&d[at]);
        if (at + 12 + len > d.size()) return none;
        const unsigned char *tag = &d[at + 4], *body = &d[at + 8];
        if (!std::memcmp(tag, "IHDR", 4) && len == 13) {
            w = (int)rs_png::get32(body); h = (int)rs_png::get32(body + 4);
            if (body[8] != 8 || (body[9] != 0 && body[9] != 2) || body[12] != 0) return none;
            ch = body[9] == 2 ? 3 : 1;
        } else if (!std::memcmp(tag, "IDAT", 4)) z.insert(z.end(), body, body + len);
        at += 12 + len;
    }
    if (!ch || z.size() < 6) return none;
    std::vector<unsigned char> raw;
    for (size_t at = 2; at + 5 <= z.size();) {
        const unsigned char head = z[at];
        if (head & 6) return none;                                  // a compressed block
        const size_t len = z[at + 1] | ((size_t)z[at + 2] << 8);
        if (at + 5 + len > z.size()) return none;
        raw.insert(raw.end(), z.begin() + (std::ptrdiff_t)(at + 5), z.begin() + (std::ptrdiff_t)(at + 5 + len));
        at += 5 + len;
        if (head & 1) break;
    }
    if (raw.size() != (size_t)h * ((size_t)w * ch + 1)) return none;
    Mat img(h, w, ch == 3 ? CV_8UC3 : CV_8UC1);
    for (int y = 0; y < h; ++y) {
        const unsigned char *row = &raw[(size_t)y * ((size_t)w * ch + 1)];
        if (row[0] != 0) return none;                               // a filtered scanline
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < ch; ++c) img.data[((size_t)y * w + x) * ch + (ch == 3 ? 2 - c : c)] = row[1 + (size_t)x * ch + c];
    }
    return img;
}

}  // namespace cv
#endif

// rs_cv.h -- the OpenCV types of the reference's hot-path signatures (cv::Mat 8UC3 / 8UC1 images,
// cv::Mat_<cv::Point_<double>> flow fields, cv::Vec3b / cv::Vec3f; camera.h, rsframe.h).
// A real OpenCV is used when present; otherwise this minimal, ref-counted stand-in with the same
// layout (row-major, interleaved channels) and the members the callers use (rows, cols, clone(),
// at<T>(y,x), data).  Image file IO (imread / imwrite) is outside the hot path and not provided.
#pragma once

#if defined(RSDSFM_USE_REAL_OPENCV) || (defined(__has_include) && __has_include(<opencv2/core.hpp>))
#include <opencv2/core.hpp>
#else

#include <cstdint>
#include <cstring>
#include <memory>
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

}  // namespace cv
#endif

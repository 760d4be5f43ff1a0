// plslam_cvlite.h — the handful of OpenCV / Eigen types the reference's front-end signatures mention.
//
// Built with -DPLSLAM_WITH_OPENCV the real <opencv2/core.hpp> types are used and the mirrors below have exactly the
// reference's signatures.  Without OpenCV (this image has no OpenCV C++ SDK) layout-compatible stand-ins are used so
// that the host layer still compiles and can be tested: cv::KeyPoint (28 bytes), cv::line_descriptor::KeyLine (68 bytes),
// a minimal reference-counted cv::Mat for CV_8UC1 planes, and Eigen::Vector3d as three doubles.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/plslam_c.h"

#ifdef PLSLAM_WITH_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/line_descriptor.hpp>
#include <Eigen/Core>
#else
namespace cv {
struct Point2f { float x = 0, y = 0; };
struct KeyPoint {  // same field order and size as cv::KeyPoint
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};
static_assert(sizeof(KeyPoint) == sizeof(pl_keypoint), "cv::KeyPoint layout");
enum { CV_8U = 0, CV_8UC1 = 0, CV_32F = 5, CV_32FC1 = 5, CV_64F = 6, CV_64FC1 = 6 };
class Mat {  // single-channel matrix (8U, 32F or 64F), shared ownership, optional ROI (data offset + step in bytes)
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uint8_t* data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t st = 0) : rows(r), cols(c), data((uint8_t*)ext), type_(type) { step = st ? st : (size_t)c * elemSize(); }
    void create(int r, int c, int type) {
        if (r == rows && c == cols && buf_ && type == type_ && step == (size_t)c * elemSize()) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
        const size_t bytes = (size_t)r * step;
        buf_ = std::shared_ptr<uint8_t>(new uint8_t[bytes > 0 ? bytes : 1](), std::default_delete<uint8_t[]>());
        data = buf_.get();
    }
    void release() { rows = cols = 0; step = 0; data = nullptr; buf_.reset(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_32F ? 4 : (type_ == CV_64F ? 8 : 1); }
    uint8_t* ptr(int r = 0) { return data + (size_t)r * step; }
    const uint8_t* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
    template <typename T> T& at(int i) { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }
    template <typename T> const T& at(int i) const { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }
    Mat row(int r) const {
        Mat m = *this;
        m.data = data + (size_t)r * step;
        m.rows = 1;
        return m;
    }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize());
        return m;
    }
    Mat roi(int x, int y, int w, int h) const {
        Mat m = *this;
        m.data = data + (size_t)y * step + (size_t)x * elemSize();
        m.rows = h; m.cols = w;
        return m;
    }
private:
    std::shared_ptr<uint8_t> buf_;
    int type_ = CV_8UC1;
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;
namespace line_descriptor {
struct KeyLine {  // same field order and size as cv::line_descriptor::KeyLine
    float angle = 0;
    int class_id = -1, octave = 0;
    Point2f pt;
    float response = 0, size = 0;
    float startPointX = 0, startPointY = 0, endPointX = 0, endPointY = 0;
    float sPointInOctaveX = 0, sPointInOctaveY = 0, ePointInOctaveX = 0, ePointInOctaveY = 0;
    float lineLength = 0;
    int numOfPixels = 0;
};
static_assert(sizeof(KeyLine) == sizeof(pl_keyline), "cv::line_descriptor::KeyLine layout");
}  // namespace line_descriptor
}  // namespace cv
namespace Eigen {
struct Vector3d {
    double v[3] = {0, 0, 0};
    Vector3d() {}
    Vector3d(double x, double y, double z) { v[0] = x; v[1] = y; v[2] = z; }
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    const double* data() const { return v; }
};
}  // namespace Eigen
#endif

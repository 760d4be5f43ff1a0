// cv_standin.hpp — a stand-in for the handful of OpenCV types and functions the reference's own sources use on the hot path
// (TEST INFRASTRUCTURE: it exists so that /root/reference/src/ORBextractor.cc and the vendored DBoW2 compile here, where there
// is no OpenCV C++ SDK, into oracle/_ref/libplref.so — see oracle/ref_shim/Makefile).  The image-processing functions forward to
// the oracle's restatements, each of which is pinned bit-exactly to cv2 4.13 by tests/test_oracle_orb.py; everything else in the
// compiled code is the reference's.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <list>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../oracle.h"

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5  // (only DBoW2's unused FORB::toMat32F asks for it)
#define CV_PI 3.1415926535897932384626433832795
#define CV_Assert(x) assert(x)

static inline int cvRound(double v) { return (int)lrint(v); }  // SSE2 cvtsd2si: round half to even
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {
using ::cvRound;
using ::cvFloor;
using ::cvCeil;
enum { INTER_LINEAR = 1, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U>
    Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct KeyPoint {  // field order of cv::KeyPoint
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

// 8-bit single-channel matrix: shared buffer + (data, step) view, the only kind the front-end uses
class Mat {
public:
    int rows, cols;
    size_t step;
    uchar* data;
    size_t esz = 1;   // element size: 1 (CV_8U) or 4 (CV_32F: the 4x4 poses and 3x1 points of the matchers)
    Mat() : rows(0), cols(0), step(0), data(nullptr) {}
    Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr) { create(r, c, type); }
    Mat(Size sz, int type) : rows(0), cols(0), step(0), data(nullptr) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, void* ext, size_t st = 0)
        : rows(r), cols(c), step(st ? st : (size_t)c * (type == CV_32F ? 4 : 1)), data((uchar*)ext), esz(type == CV_32F ? 4 : 1) {}
    void create(int r, int c, int type) {
        if (r == rows && c == cols && data && esz == (size_t)(type == CV_32F ? 4 : 1)) return;  // (as cv::Mat::create: a matrix of the right size, ROI or not, is kept)
        esz = type == CV_32F ? 4 : 1;
        rows = r; cols = c; step = (size_t)c * esz;
        buf_ = std::shared_ptr<uchar>(new uchar[(size_t)r * step > 0 ? (size_t)r * step : 1], std::default_delete<uchar[]>());
        data = buf_.get();
    }
    void release() { rows = cols = 0; step = 0; data = nullptr; buf_.reset(); }
    // Mat::zeros is a matrix EXPRESSION: assigned to a matrix of the same size it fills that matrix in place (the reference's
    // computeDescriptors relies on it: `descriptors = Mat::zeros(...)` clears the rows of the output it was handed)
    struct ZerosExpr { int r, c, type; };
    static ZerosExpr zeros(int r, int c, int type) { return ZerosExpr{r, c, type}; }
    Mat(const ZerosExpr& e) : rows(0), cols(0), step(0), data(nullptr) { *this = e; }
    Mat& operator=(const ZerosExpr& e) {
        create(e.r, e.c, e.type);
        for (int y = 0; y < rows; y++) std::memset(ptr(y), 0, (size_t)cols * (e.type == CV_32F ? 4 : 1));
        return *this;
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    size_t step1() const { return step; }
    Size size() const { return Size(cols, rows); }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + x * sizeof(T)); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    Mat rowRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat colRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * esz; m.cols = b - a; return m; }
    Mat col(int c) const { return colRange(c, c + 1); }
    double dot(const Mat& o) const;
    // cv::Mat::push_back(row) of an 8-bit matrix (LineMatcher collects the descriptors of the projected lines that way)
    void push_back(const Mat& r) {
        assert(esz == 1 && r.esz == 1 && (rows == 0 || cols == r.cols));
        Mat m(rows + r.rows, r.cols, CV_8UC1);
        for (int y = 0; y < rows; y++) std::memcpy(m.ptr(y), ptr(y), (size_t)cols);
        for (int y = 0; y < r.rows; y++) std::memcpy(m.ptr(rows + y), r.ptr(y), (size_t)r.cols);
        *this = m;
    }
    // single-index access of a vector (3x1 or 1x3)
    template <typename T> T& at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    template <typename T> const T& at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    // ---- CV_32F matrix expressions, as far as the matchers write them: -A.t()*b, A*b, A*b+c.  cv::MatExpr turns each of them into
    // ONE gemm; for CV_32F cv::gemm accumulates the products in double and rounds alpha*sum + beta*c to float once
    // (GEMMSingleMul<float, double>): the stand-in evaluates them the same way. ----
    struct TExpr { const Mat* m; double alpha; };
    struct MulExpr;
    TExpr t() const { return TExpr{this, 1.0}; }
    Mat(const MulExpr& e);
    Mat operator()(const Rect& r) const { Mat m = *this; m.data = data + (size_t)r.y * step + r.x; m.rows = r.height; m.cols = r.width; return m; }
    Mat clone() const {
        Mat m(rows, cols, esz == 4 ? CV_32F : CV_8UC1);
        for (int y = 0; y < rows; y++) std::memcpy(m.ptr(y), ptr(y), (size_t)cols * esz);
        return m;
    }
    void copyTo(Mat& o) const {
        if (o.rows != rows || o.cols != cols) o.create(rows, cols, CV_8UC1);
        for (int y = 0; y < rows; y++) std::memmove(o.ptr(y), ptr(y), (size_t)cols);
    }
private:
    std::shared_ptr<uchar> buf_;
};
struct Mat::MulExpr { Mat a, b; double alpha; bool ta; };
inline Mat::TExpr operator-(const Mat::TExpr& e) { return Mat::TExpr{e.m, -e.alpha}; }
inline Mat::MulExpr operator*(const Mat::TExpr& e, const Mat& b) { return Mat::MulExpr{*e.m, b, e.alpha, true}; }
inline Mat::MulExpr operator*(const Mat& a, const Mat& b) { return Mat::MulExpr{a, b, 1.0, false}; }
inline Mat gemm_f32(const Mat::MulExpr& e, const Mat* c) {
    const int M = e.ta ? e.a.cols : e.a.rows, K = e.ta ? e.a.rows : e.a.cols, N = e.b.cols;
    assert(e.a.esz == 4 && e.b.esz == 4 && e.b.rows == K && (!c || (c->rows == M && c->cols == N && c->esz == 4)));
    Mat d(M, N, CV_32F);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            double s = 0;
            for (int k = 0; k < K; k++) s += (double)(e.ta ? e.a.at<float>(k, i) : e.a.at<float>(i, k)) * (double)e.b.at<float>(k, j);
            d.at<float>(i, j) = c ? (float)(s * e.alpha + (double)c->at<float>(i, j) * 1.0) : (float)(s * e.alpha);
        }
    return d;
}
inline Mat::Mat(const MulExpr& e) : rows(0), cols(0), step(0), data(nullptr) { *this = gemm_f32(e, nullptr); }
inline Mat operator+(const Mat::MulExpr& e, const Mat& c) { return gemm_f32(e, &c); }
// cv::Mat_<float>(r, c) << a, b, c: the comma initialiser
template <typename T>
struct Mat_ : Mat {
    Mat_(int r, int c) : Mat(r, c, CV_32F) {}
};
template <typename T>
struct MatCommaInit {
    Mat m;
    int idx;
    MatCommaInit& operator,(T v) { m.at<T>(idx / m.cols, idx % m.cols) = v; idx++; return *this; }
    operator Mat() const { return m; }
};
template <typename T>
inline MatCommaInit<T> operator<<(const Mat_<T>& m, T v) {
    MatCommaInit<T> ci{m, 0};
    ci, v;
    return ci;
}
// s * A, s * A.t() (scaled copies, the scale applied in float) and -A * b (one gemm with alpha = -1) of CV_32F matrices
inline Mat operator*(double s, const Mat& a) {
    assert(a.esz == 4);
    const float alpha = (float)s;
    Mat d(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) d.at<float>(i, j) = a.at<float>(i, j) * alpha;
    return d;
}
inline Mat operator*(double s, const Mat::TExpr& e) {
    const Mat& a = *e.m;
    assert(a.esz == 4);
    const float alpha = (float)(s * e.alpha);
    Mat d(a.cols, a.rows, CV_32F);
    for (int i = 0; i < a.cols; i++)
        for (int j = 0; j < a.rows; j++) d.at<float>(i, j) = a.at<float>(j, i) * alpha;
    return d;
}
struct NegExpr { Mat m; };
inline NegExpr operator-(const Mat& a) { return NegExpr{a}; }
inline Mat::MulExpr operator*(const NegExpr& e, const Mat& b) { return Mat::MulExpr{e.m, b, -1.0, false}; }
// a - b of two CV_32F matrices (element-wise float subtraction) and cv::norm (L2) of a CV_32F matrix: products accumulated in double
inline Mat operator-(const Mat& a, const Mat& b) {
    assert(a.esz == 4 && b.esz == 4 && a.rows == b.rows && a.cols == b.cols);
    Mat d(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) d.at<float>(i, j) = a.at<float>(i, j) - b.at<float>(i, j);
    return d;
}
// A / s of a CV_32F matrix (cv::MatExpr: a scaled copy, the scale 1 / s applied in float) and Mat::dot (products accumulated in double)
inline Mat operator/(const Mat& a, double s) {
    assert(a.esz == 4);
    const float alpha = (float)(1.0 / s);
    Mat d(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) d.at<float>(i, j) = a.at<float>(i, j) * alpha;
    return d;
}
inline double dot_f32(const Mat& a, const Mat& b) {
    assert(a.esz == 4 && b.esz == 4 && a.rows * a.cols == b.rows * b.cols);
    double s = 0;
    const int n = a.rows * a.cols;
    for (int i = 0; i < n; i++) s += (double)a.at<float>(i / a.cols, i % a.cols) * (double)b.at<float>(i / b.cols, i % b.cols);
    return s;
}
inline double Mat::dot(const Mat& o) const { return dot_f32(*this, o); }
inline double norm(const Mat& a) {
    assert(a.esz == 4);
    double s = 0;
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) s += (double)a.at<float>(i, j) * (double)a.at<float>(i, j);
    return std::sqrt(s);
}

// InputArray / OutputArray: thin handles on a Mat
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat& m) { m_ = &m; }
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void release() const { m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline Mat noArray() { return Mat(); }

// ---- functions: forwarded to the oracle's restatements (pinned to cv2 4.13) ----
inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }
inline void resize(const Mat& src, Mat& dst, Size dsize, double = 0, double = 0, int interpolation = INTER_LINEAR) {
    assert(interpolation == INTER_LINEAR);
    if (dst.rows != dsize.height || dst.cols != dsize.width) dst.create(dsize.height, dsize.width, CV_8UC1);
    orc_resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}
// (the reference calls it with dst = the bordered buffer whose interior src is: BORDER_ISOLATED, in-place safe)
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType) {
    assert(top == bottom && left == right && top == left && (borderType & 15) == BORDER_REFLECT_101);
    (void)bottom; (void)right; (void)borderType;
    if (dst.rows != src.rows + 2 * top || dst.cols != src.cols + 2 * left) dst.create(src.rows + 2 * top, src.cols + 2 * left, CV_8UC1);
    Mat tmp = src.clone();
    orc_border_reflect101_u8(tmp.data, tmp.cols, tmp.rows, tmp.step, dst.data, dst.step, top);
}
inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sx, double sy = 0, int borderType = BORDER_DEFAULT) {
    assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && (sy == 2 || sy == 0) && borderType == BORDER_REFLECT_101);
    (void)ksize; (void)sx; (void)sy; (void)borderType;
    Mat tmp = src.clone();
    if (dst.rows != src.rows || dst.cols != src.cols) dst.create(src.rows, src.cols, CV_8UC1);
    orc_gaussian_blur7_u8(tmp.data, tmp.cols, tmp.rows, tmp.step, dst.data, dst.step);
}
inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nonmax = true) {
    const int cap = img.rows * img.cols + 1;
    std::vector<float> xs(cap), ys(cap), rs(cap);
    const int n = orc_fast_detect(img.data, img.cols, img.rows, img.step, threshold, nonmax ? 1 : 0, xs.data(), ys.data(), rs.data(), cap);
    kps.clear();
    for (int i = 0; i < n; i++) kps.push_back(KeyPoint(xs[i], ys[i], 7.f, -1.f, rs[i]));
}
// cv::line_descriptor::KeyLine (opencv_contrib): the field order of pl_keyline
namespace line_descriptor {
struct KeyLine {
    float angle;
    int class_id, octave;
    Point2f pt;
    float response, size, startPointX, startPointY, endPointX, endPointY, sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY,
        lineLength;
    int numOfPixels;
};
}  // namespace line_descriptor
// cv::LineIterator(img, pt1, pt2).count (8-connected, clipped to the image): the oracle's cv2-pinned restatement
class LineIterator {
public:
    int count;
    LineIterator(const Mat& img, Point2f a, Point2f b) : count(orc_line_iterator_count(a.x, a.y, b.x, b.y, img.cols, img.rows)) {}
};
// cv::FileStorage: DBoW2's YAML save / load (virtual members of the vocabulary template, so they must compile; never called —
// the vocabulary is loaded with the reference's own loadFromTextFile)
class FileNode {
public:
    FileNode operator[](const char*) const { assert(!"cv::FileStorage is not on the path"); return FileNode(); }
    FileNode operator[](const std::string&) const { assert(!"cv::FileStorage is not on the path"); return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const char*, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

struct KeyPointsFilter {  // (only the reference's unused ComputeKeyPointsOld calls it)
    static void retainBest(std::vector<KeyPoint>&, int) { assert(!"KeyPointsFilter::retainBest is not on the path"); }
};
}  // namespace cv

// LineExtractor.h — host-side mirror of ORB_SLAM2::LineExtractor (reference include/LineExtractor.h:21-61,
// src/LineExtractor.cpp:12-70) on top of the C ABI.  Frame::ExtractLine (src/Frame.cc:326-328) compiles unchanged.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "plslam_cvlite.h"

namespace ORB_SLAM2 {
using namespace cv::line_descriptor;

class LineExtractor {
public:
    explicit LineExtractor(int maxCols = 1280, int maxRows = 1024, int device = 0) {
        int rc = pl_line_create(&h_, device, maxCols, maxRows, 1);
        if (rc != PL_OK) throw std::runtime_error(std::string("LineExtractor (CUDA): ") + pl_last_error());
    }
    ~LineExtractor() { pl_line_destroy(h_); }
    LineExtractor(const LineExtractor&) = delete;
    LineExtractor& operator=(const LineExtractor&) = delete;

    // scale / num_octaves are accepted for signature compatibility; the reference always runs 1 octave at scale 1
    // (`int scale = 1.2` truncates to 1, LineExtractor.h:29-30)
    void ExtractLineSegment(const cv::Mat& img, std::vector<KeyLine>& key_lines, cv::Mat& line_descriptor,
                            std::vector<Eigen::Vector3d>& keyline_coefficients, int /*scale*/ = 1.2, int /*num_octaves*/ = 1) {
        const int nums_lineFeature = 80;  // LineExtractor.cpp:23
        pl_keyline kls[80];
        uint8_t desc[80 * 32];
        double co[80 * 3];
        int n = 0;
        key_lines.clear();
        if (img.empty()) return;
        int rc = pl_line_extract(h_, img.ptr(0), img.rows, img.cols, (size_t)img.step, nums_lineFeature, kls, desc, co, &n);
        if (rc != PL_OK) throw std::runtime_error(std::string("LineExtractor (CUDA): ") + pl_last_error());
        key_lines.resize(n);
        static_assert(sizeof(KeyLine) == sizeof(pl_keyline), "layout");
        if (n) std::memcpy((void*)key_lines.data(), kls, sizeof(pl_keyline) * (size_t)n);
#ifdef PLSLAM_WITH_OPENCV
        line_descriptor.create(n, 32, CV_8UC1);
        for (int i = 0; i < n; i++) std::memcpy(line_descriptor.ptr(i), desc + 32 * i, 32);
#else
        line_descriptor.create(n, 32, cv::CV_8UC1);
        if (n) std::memcpy(line_descriptor.data, desc, (size_t)n * 32);
#endif
        for (int i = 0; i < n; i++) {  // the reference appends (push_back) to the caller's vector (:60-69)
            Eigen::Vector3d c;
            c(0) = co[3 * i]; c(1) = co[3 * i + 1]; c(2) = co[3 * i + 2];
            keyline_coefficients.push_back(c);
        }
    }
    pl_line* handle() { return h_; }

private:
    pl_line* h_ = nullptr;
};

}  // namespace ORB_SLAM2

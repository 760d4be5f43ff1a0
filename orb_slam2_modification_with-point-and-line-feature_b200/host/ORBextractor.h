// ORBextractor.h — host-side mirror of ORB_SLAM2::ORBextractor on top of the C ABI (include/plslam_c.h).
//
// Same class name, constructor, operator(), getters and public mvImagePyramid as the reference
// (reference include/ORBextractor.h:44-107), so src/Frame.cc:318-324 (Frame::ExtractORB), the getters at
// Frame.cc:79-85,143-149 and Tracking.cc:119-125 compile against it unchanged.  All work happens in the CUDA library;
// there is no CPU path: a missing device or library makes the constructor throw.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "plslam_cvlite.h"

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    // max image size / device are the only additions; they have defaults so the reference call sites are unchanged
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int maxCols = 1280, int maxRows = 1024,
                 int device = 0)
        : nlevels_(nlevels) {
        int rc = pl_orb_create(&h_, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device, maxCols, maxRows, 1);
        if (rc != PL_OK) throw std::runtime_error(std::string("ORBextractor (CUDA): ") + pl_last_error());
        mvImagePyramid.resize(nlevels);
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        pl_orb_scale_factors(h_, mvScaleFactor.data());
        pl_orb_inv_scale_factors(h_, mvInvScaleFactor.data());
        pl_orb_level_sigma2(h_, mvLevelSigma2.data());
        pl_orb_inv_level_sigma2(h_, mvInvLevelSigma2.data());
        kps_.resize(pl_orb_max_keypoints(h_));
    }
    ~ORBextractor() { pl_orb_destroy(h_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image.  Mask is ignored, as in the reference.
    void operator()(cv::InputArray image_, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors_) {
#ifdef PLSLAM_WITH_OPENCV
        cv::Mat image = image_.getMat();
#else
        const cv::Mat& image = image_;
#endif
        if (image.empty()) return;  // ORBextractor.cc:1046-1047
        const int cap = (int)kps_.size();
        desc_.resize((size_t)cap * 32);
        int n = 0;
        int rc = pl_orb_extract(h_, image.ptr(0), image.rows, image.cols, (size_t)image.step, kps_.data(), desc_.data(), cap, &n);
        if (rc != PL_OK) throw std::runtime_error(std::string("ORBextractor (CUDA): ") + pl_last_error());
        keypoints.resize(n);
        static_assert(sizeof(cv::KeyPoint) == sizeof(pl_keypoint), "layout");
        if (n) std::memcpy((void*)keypoints.data(), kps_.data(), sizeof(pl_keypoint) * (size_t)n);
#ifdef PLSLAM_WITH_OPENCV
        if (n == 0) descriptors_.release();
        else {
            descriptors_.create(n, 32, CV_8U);
            cv::Mat d = descriptors_.getMat();
            for (int i = 0; i < n; i++) std::memcpy(d.ptr(i), desc_.data() + 32 * (size_t)i, 32);
        }
#else
        if (n == 0) descriptors_.release();
        else {
            descriptors_.create(n, 32, cv::CV_8U);
            std::memcpy(descriptors_.data, desc_.data(), (size_t)n * 32);
        }
#endif
        if (mbFillPyramid) {  // the reference exposes the pyramid as a public member (read by Frame::ComputeStereoMatches)
            for (int l = 0; l < nlevels_; l++) {
                int r = 0, c = 0;
                pl_orb_pyramid_dims(h_, l, &r, &c);
                if (bordered_.size() != (size_t)nlevels_) bordered_.resize(nlevels_);
#ifdef PLSLAM_WITH_OPENCV
                bordered_[l].create(r + 38, c + 38, CV_8UC1);
                pl_orb_pyramid_read(h_, 0, l, bordered_[l].ptr(0), bordered_[l].step);
                mvImagePyramid[l] = bordered_[l](cv::Rect(19, 19, c, r));
#else
                bordered_[l].create(r + 38, c + 38, cv::CV_8UC1);
                pl_orb_pyramid_read(h_, 0, l, bordered_[l].ptr(0), bordered_[l].step);
                mvImagePyramid[l] = bordered_[l].roi(19, 19, c, r);
#endif
            }
        }
    }

    int inline GetLevels() { return nlevels_; }
    float inline GetScaleFactor() { return pl_orb_scale_factor(h_); }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    std::vector<cv::Mat> mvImagePyramid;  // ROI at (19,19) of the bordered plane, as in ORBextractor.cc:1115
    bool mbFillPyramid = true;            // set false when nothing reads mvImagePyramid (saves ~1.2 MB of D2H per frame)

    pl_orb* handle() { return h_; }

protected:
    pl_orb* h_ = nullptr;
    int nlevels_;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<pl_keypoint> kps_;
    std::vector<uint8_t> desc_;
    std::vector<cv::Mat> bordered_;
};

}  // namespace ORB_SLAM2

// FrameGlue.h — host-side entry points for the pieces of ORB_SLAM2::Frame that sit between the extractors and the matchers
// (SURVEY.md §8(f) ranks 1 and 4): Frame::ComputeBoW / ORBVocabulary, UndistortKeyPoints, ComputeStereoFromRGBD, UnprojectStereo,
// IsInFrustum.  Same idea as Matchers.h: the reference's method names, POD arguments, the C ABI underneath.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "plslam_cvlite.h"

namespace DBoW2 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
typedef std::map<WordId, WordValue> BowVector;                      // Thirdparty/DBoW2/DBoW2/BowVector.h:56-58
typedef std::map<NodeId, std::vector<unsigned int>> FeatureVector;  // Thirdparty/DBoW2/DBoW2/FeatureVector.h:21-22
}  // namespace DBoW2

namespace ORB_SLAM2 {

// ORBVocabulary (include/ORBVocabulary.h:31-32 = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>): the two calls the
// front-end makes — loadFromTextFile (System.cc:67) and transform (Frame.cc:721-735)
class ORBVocabulary {
public:
    explicit ORBVocabulary(int device = 0) : device_(device) {}
    ~ORBVocabulary() { pl_voc_destroy(h_); }
    ORBVocabulary(const ORBVocabulary&) = delete;
    ORBVocabulary& operator=(const ORBVocabulary&) = delete;
    bool loadFromTextFile(const std::string& filename) {
        pl_voc_destroy(h_);
        h_ = nullptr;
        return pl_voc_load_text(&h_, device_, filename.c_str()) == PL_OK;
    }
    bool empty() const {
        int n = 0;
        return !h_ || pl_voc_info(h_, nullptr, nullptr, &n, nullptr) != PL_OK || n <= 1;
    }
    // transform(features, v, fv, levelsup) with the descriptors as the rows of mDescriptors (n x 32, CV_8U)
    void transform(const uint8_t* descriptors, int n, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const {
        v.clear();
        fv.clear();
        if (!h_ || n <= 0) return;
        const int off[2] = {0, n};
        int nw = 0, nn = 0;
        std::vector<unsigned int> wid(n), nid(n), fi(n);
        std::vector<double> wv(n);
        std::vector<int> noff(n + 1);
        if (pl_voc_transform_batch(h_, 1, off, descriptors, levelsup, &nw, wid.data(), wv.data(), &nn, nid.data(), noff.data(), fi.data()) != PL_OK)
            throw std::runtime_error(std::string("ORBVocabulary (CUDA): ") + pl_last_error());
        for (int k = 0; k < nw; k++) v.emplace_hint(v.end(), wid[k], wv[k]);
        for (int k = 0; k < nn; k++) fv.emplace_hint(fv.end(), nid[k], std::vector<unsigned int>(fi.begin() + noff[k], fi.begin() + noff[k + 1]));
    }
    pl_voc* handle() { return h_; }

private:
    int device_;
    pl_voc* h_ = nullptr;
};

// the element-wise maps of Frame (Frame.cc:737-765, 1065-1134, 345-401) over POD arrays; see include/plslam_c.h (F rows)
class FrameGlue {
public:
    explicit FrameGlue(int device = 0) {
        if (pl_match_create(&h_, device) != PL_OK) throw std::runtime_error(std::string("FrameGlue (CUDA): ") + pl_last_error());
    }
    ~FrameGlue() { pl_match_destroy(h_); }
    FrameGlue(const FrameGlue&) = delete;
    FrameGlue& operator=(const FrameGlue&) = delete;
    // Frame::UndistortKeyPoints: mvKeysUn = mvKeys with undistorted pt
    void UndistortKeyPoints(const std::vector<cv::KeyPoint>& mvKeys, float fx, float fy, float cx, float cy, const float mDistCoef[5],
                            std::vector<cv::KeyPoint>& mvKeysUn) {
        const int n = (int)mvKeys.size();
        std::vector<float> xy(2 * (size_t)n), out(2 * (size_t)n);
        for (int i = 0; i < n; i++) { xy[2 * i] = mvKeys[i].pt.x; xy[2 * i + 1] = mvKeys[i].pt.y; }
        check(pl_frame_undistort_points(h_, xy.data(), n, fx, fy, cx, cy, mDistCoef, out.data()));
        mvKeysUn = mvKeys;
        for (int i = 0; i < n; i++) { mvKeysUn[i].pt.x = out[2 * i]; mvKeysUn[i].pt.y = out[2 * i + 1]; }
    }
    // Frame::ComputeStereoFromRGBD for one frame: imDepth = rows x cols CV_32F (row stride step bytes)
    void ComputeStereoFromRGBD(const float* imDepth, int rows, int cols, size_t step, const std::vector<cv::KeyPoint>& mvKeys,
                               const std::vector<cv::KeyPoint>& mvKeysUn, float mbf, std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
        const int n = (int)mvKeys.size();
        const int off[2] = {0, n};
        std::vector<float> xy(2 * (size_t)n), xu(n);
        for (int i = 0; i < n; i++) { xy[2 * i] = mvKeys[i].pt.x; xy[2 * i + 1] = mvKeys[i].pt.y; xu[i] = mvKeysUn[i].pt.x; }
        mvuRight.assign(n, -1.f);
        mvDepth.assign(n, -1.f);
        if (n) check(pl_frame_stereo_from_rgbd_batch(h_, 1, imDepth, 0, rows, cols, step, step * (size_t)rows, off, xy.data(), xu.data(), mbf, mvDepth.data(),
                                                     mvuRight.data()));
    }
    // Frame::IsInFrustum for every map point of a snapshot against one frame: fills what SearchByProjection reads (pl_mappoint_view)
    void IsInFrustum(const float Tcw[12], const float Ow[3], float fx, float fy, float cx, float cy, float mbf, const float bounds[4], int nLevels,
                     float logScaleFactor, int m, const float* worldPos, const float* normal, const float* minDistInv, const float* maxDistInv,
                     const float* maxDist, float viewingCosLimit, std::vector<uint8_t>& inView, std::vector<float>& projX, std::vector<float>& projY,
                     std::vector<float>& projXR, std::vector<int>& scaleLevel, std::vector<float>& viewCos) {
        inView.assign(m, 0); projX.assign(m, 0.f); projY.assign(m, 0.f); projXR.assign(m, 0.f); scaleLevel.assign(m, 0); viewCos.assign(m, 0.f);
        if (m) check(pl_frame_is_in_frustum_batch(h_, 1, Tcw, Ow, fx, fy, cx, cy, mbf, bounds, nLevels, logScaleFactor, m, worldPos, normal, minDistInv,
                                                  maxDistInv, maxDist, viewingCosLimit, inView.data(), projX.data(), projY.data(), projXR.data(),
                                                  scaleLevel.data(), viewCos.data()));
    }
    pl_match* handle() { return h_; }

private:
    void check(int rc) { if (rc != PL_OK) throw std::runtime_error(std::string("FrameGlue (CUDA): ") + pl_last_error()); }
    pl_match* h_ = nullptr;
};

}  // namespace ORB_SLAM2

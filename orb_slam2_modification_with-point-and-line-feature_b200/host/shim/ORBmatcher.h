// shim/ORBmatcher.h — ORB_SLAM2::ORBmatcher with the reference's own signatures (include/ORBmatcher.h:41-213), on the GPU.
//
// Drop-in for include/ORBmatcher.h + src/ORBmatcher.cc: same class name, constructor, public methods, argument types and
// return values; the searches run through the C ABI (include/plslam_c.h) via the POD-view mirror of host/Matchers.h, and the
// results are written back into mvpMapPoints / vpMatches exactly where the reference writes them.  Build inside the
// reference tree with -DPLSLAM_WITH_REFERENCE -DPLSLAM_WITH_OPENCV (real Frame / KeyFrame / MapPoint), or stand-alone against
// the stand-ins of shim/RefTypes.h (tests/cpp/matcher_shim_test.cpp).
#pragma once
#include <set>
#include <utility>
#include <vector>

#include "RefTypes.h"
#define PLSLAM_VIEW_NS plslam_views
#include "../Matchers.h"

namespace ORB_SLAM2 {

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);                                            // ORBmatcher.h:50
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);                                 // :58
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3);   // :74
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);  // :94
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist);  // :108
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th);    // :122
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);               // :141
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);              // :142
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10);  // :156
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t>>& vMatchedPairs, const bool bOnlyStereo);  // :170
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12, const float th);  // :186
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0);          // :196
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint);  // :208

    static const int TH_LOW;        // ORBmatcher.h:212-214, ORBmatcher.cc:49-51
    static const int TH_HIGH;
    static const int HISTO_LENGTH;

protected:
    float mfNNratio;
    bool mbCheckOrientation;
    plslam_views::ORBmatcher gpu_;
};

}  // namespace ORB_SLAM2

// RefTypes.h — the reference classes the matcher shims (shim/ORBmatcher.h, shim/LineMatcher.h) take as arguments.
//
// Built inside the reference tree (-DPLSLAM_WITH_REFERENCE -DPLSLAM_WITH_OPENCV, include path = /root/reference/include) the
// real Frame.h / KeyFrame.h / MapPoint.h / MapLine.h are used and the shims replace src/ORBmatcher.cc / src/LineMatcher.cpp.
// Without the reference headers (this image has no OpenCV / Eigen / DBoW2 SDK) the stand-ins below expose the same public
// members with the same names and meaning — exactly the ones the reference matchers read or write — so that the shims compile
// unchanged and tests/cpp/matcher_shim_test.cpp can drive them like Tracking.cc does.
//   Frame      include/Frame.h:49-330      KeyFrame  include/KeyFrame.h:50-330
//   MapPoint   include/MapPoint.h:41-230   MapLine   include/MapLine.h:27-225
#pragma once
#include <cmath>
#include <map>
#include <set>
#include <vector>

#include "../plslam_cvlite.h"

#ifdef PLSLAM_WITH_REFERENCE
#include "Frame.h"
#include "KeyFrame.h"
#include "MapLine.h"
#include "MapPoint.h"
#else
namespace DBoW2 {
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {};  // Thirdparty/DBoW2/DBoW2/FeatureVector.h:21-22
}  // namespace DBoW2

namespace ORB_SLAM2 {
using cv::line_descriptor::KeyLine;
class KeyFrame;
class Frame;

class MapPoint {
public:
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    cv::Mat GetNormal() { return mNormalVector.clone(); }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    int Observations() { return nObs; }
    bool isBad() { return mbBad; }
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    int GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }  // MapPoint.cc:395-399
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }  // MapPoint.cc:401-405
    void AddObservation(KeyFrame* pKF, size_t idx) {
        if (!mObservations.count(pKF)) nObs++;
        mObservations[pKF] = idx;
    }
    void Replace(MapPoint* pMP) { mbBad = true; mpReplaced = pMP; }
    // Frame::isInFrustum leaves these on the point (MapPoint.h:160-176)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    float mTrackViewCos = 0;
    long unsigned int mnLastFrameSeen = 0;
    // (protected in the reference; the test fills them)
    cv::Mat mWorldPos, mNormalVector, mDescriptor;
    std::map<KeyFrame*, size_t> mObservations;
    int nObs = 0;
    bool mbBad = false;
    MapPoint* mpReplaced = nullptr;
    float mfMinDistance = 0, mfMaxDistance = 0;
};

class MapLine {
public:
    int Observations() { return nObs; }
    bool isBad() { return mbBad; }
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    std::map<KeyFrame*, size_t> GetObservations() { return mObservations; }
    void Replace(MapLine* pML) { mbBad = true; mpReplaced = pML; }
    Eigen::Vector3d mStart3d, mEnd3d;  // MapLine.h:201-202
    cv::Mat mLineDescriptor;           // MapLine.h:212
    bool mbTrackInView = false;
    std::map<KeyFrame*, size_t> mObservations;
    int nObs = 0;
    bool mbBad = false;
    MapLine* mpReplaced = nullptr;
};

// what Frame and KeyFrame share for the matchers
struct FrameCommon {
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    int NL = 0;
    std::vector<KeyLine> mvKeyLines, mvKeyLinesUn;
    cv::Mat mLineDescriptors;
    std::vector<MapLine*> mvpMapLines;
    DBoW2::FeatureVector mFeatVec;
    float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0, mbf = 0, mb = 0;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
    int mnScaleLevels = 0;
    float mfScaleFactor = 0, mfLogScaleFactor = 0;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
};

class Frame : public FrameCommon {
public:
    cv::Mat mTcw;  // 4 x 4, CV_32F
    std::vector<bool> mvbOutlier, mvbLineOutlier;
    cv::Mat im_gray_;  // only its size is read (UpdateKeyLineData, LineMatcher.cpp:1614-1623)
    long unsigned int mnId = 0;
    cv::Mat GetCameraCenter() {  // mOw = -Rcw^T tcw (Frame.cc:331-343)
        cv::Mat ow(3, 1, cv::CV_32F);
        for (int r = 0; r < 3; r++) {
            double a = 0;
            for (int k = 0; k < 3; k++) a += (double)mTcw.at<float>(k, r) * (double)mTcw.at<float>(k, 3);
            ow.at<float>(r) = (float)-a;
        }
        return ow;
    }
};

class KeyFrame : public FrameCommon {
public:
    cv::Mat Tcw;  // 4 x 4, CV_32F
    cv::Mat GetPose() { return Tcw.clone(); }
    cv::Mat GetRotation() { return Tcw.roi(0, 0, 3, 3).clone(); }
    cv::Mat GetTranslation() { return Tcw.roi(3, 0, 1, 3).clone(); }
    cv::Mat GetCameraCenter() {
        cv::Mat ow(3, 1, cv::CV_32F);
        for (int r = 0; r < 3; r++) {
            double a = 0;
            for (int k = 0; k < 3; k++) a += (double)Tcw.at<float>(k, r) * (double)Tcw.at<float>(k, 3);
            ow.at<float>(r) = (float)-a;
        }
        return ow;
    }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
    std::vector<MapLine*> GetMapLineMatches() { return mvpMapLines; }
    MapLine* GetMapLine(const size_t& idx) { return mvpMapLines[idx]; }
    cv::Mat im_gray_;
    long unsigned int mnId = 0;
};

}  // namespace ORB_SLAM2
#endif

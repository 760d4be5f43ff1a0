// shim/LineMatcher.h — ORB_SLAM2::LineMatcher with the reference's own signatures (include/LineMatcher.h:27-130), on the GPU.
//
// Drop-in for include/LineMatcher.h + src/LineMatcher.cpp (see shim/ORBmatcher.h for how it is built).  The two overloads the
// reference declares but never defines (relocalisation and Sim3 variants, LineMatcher.h:74-79) are not declared here either.
#pragma once
#include <utility>
#include <vector>

#include "RefTypes.h"
#define PLSLAM_VIEW_NS plslam_views
#include "../Matchers.h"

namespace ORB_SLAM2 {

class LineMatcher {
public:
    typedef cv::line_descriptor::KeyLine KeyLine;
    LineMatcher(float nnratio = 0.6, bool checkOri = true);                                            // LineMatcher.h:35
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);                                 // :43
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame);                               // :48
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, std::vector<KeyLine>& new_kls, std::vector<std::pair<int, int>>& match_indices);  // :50
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* RefFrame, std::vector<MapLine*>& vpMapLineMatches);  // :55
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* RefFrame);                                   // :56
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* RefFrame, std::vector<KeyLine>& new_kls, std::vector<std::pair<int, int>>& match_indices);  // :58
    int SearchByProjection(Frame& F, const std::vector<MapLine*>& vpMapLines);                         // :63
    int SearchByProjection(Frame& F, const std::vector<MapLine*>& vpMapLines, std::vector<KeyLine>& new_kls, std::vector<std::pair<int, int>>& match_indices);  // :65
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<std::pair<size_t, size_t>>& vMatchedPairs, const bool bOnlyStereo);  // :81
    int Fuse(KeyFrame* pKF, const std::vector<MapLine*>& vpMapLines);                                  // :86

    float mfNNratio;
    bool mbCheckOrientation;
    // LineMatcher.h:94-98 (the CUDA predicate uses the same constants)
    double angle_threshold_ = 15.0 * 3.14159265358979323846 / 180.0;
    double length_threshold_ = 0.45;
    double overlap_threshold_ = 0.5;
    double desc_dist_threshold_ = 45;
    double reproj_error_threshold_ = 45;

protected:
    // what the three projection searches share: offer `lines` (gate `valid`) to the current frame, write the matches back
    int Search(Frame& Cur, const std::vector<MapLine*>& lines, const std::vector<KeyLine>& src_kl, const std::vector<uint8_t>& valid,
               std::vector<KeyLine>* new_kls, std::vector<std::pair<int, int>>* match_indices);
    plslam_views::LineMatcher gpu_;
};

}  // namespace ORB_SLAM2

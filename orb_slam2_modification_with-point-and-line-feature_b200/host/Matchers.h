// Matchers.h — host-side entry points for the Hamming searches of ORB_SLAM2::ORBmatcher / LineMatcher.
//
// The reference matchers take Frame / KeyFrame / MapPoint / MapLine objects (include/ORBmatcher.h:74-213,
// include/LineMatcher.h:49-84).  Those classes are outside the hot-path scope, so the mirrors here keep the method
// names and argument meaning but take the POD views of include/plslam_c.h; INTEGRATION.md shows the ten-line adapters
// that fill a view from a reference Frame and write the result back into mvpMapPoints / mvpMapLines.
#pragma once
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "plslam_cvlite.h"

// The shims with the reference's own signatures (host/shim/ORBmatcher.h, LineMatcher.h) are the classes called
// ORB_SLAM2::ORBmatcher / LineMatcher there; they include this header with PLSLAM_VIEW_NS set to another namespace.
#ifndef PLSLAM_VIEW_NS
#define PLSLAM_VIEW_NS ORB_SLAM2
#endif

namespace PLSLAM_VIEW_NS {

class ORBmatcher {
public:
    static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;  // ORBmatcher.cc:49-51
    ORBmatcher(float nnratio = 0.6, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {
        if (pl_match_create(&h_, device) != PL_OK) throw std::runtime_error(std::string("ORBmatcher (CUDA): ") + pl_last_error());
    }
    ~ORBmatcher() { pl_match_destroy(h_); }
    ORBmatcher(const ORBmatcher&) = delete;
    ORBmatcher& operator=(const ORBmatcher&) = delete;

    // ORBmatcher::DescriptorDistance(a, b) for n row pairs (ORBmatcher.cc:2083-2103)
    void DescriptorDistance(const uint8_t* a, const uint8_t* b, int n, int* dist) { check(pl_hamming_pairs(h_, a, b, n, dist)); }
    // SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th) (ORBmatcher.cc:72-183)
    int SearchByProjection(const pl_frame_view& F, const pl_mappoint_view& vpMapPoints, float th, std::vector<int>& match_of_feature) {
        int n = 0;
        match_of_feature.resize(F.n > 0 ? F.n : 1);
        check(pl_orb_search_local_points(h_, &F, &vpMapPoints, th, mfNNratio, match_of_feature.data(), &n));
        match_of_feature.resize(F.n);
        return n;
    }
    // SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) (ORBmatcher.cc:1710-1879)
    int SearchByProjection(const pl_frame_view& CurrentFrame, const pl_lastframe_view& LastFrame, float th, bool bMono,
                           std::vector<int>& match_of_feature) {
        int n = 0;
        match_of_feature.resize(CurrentFrame.n > 0 ? CurrentFrame.n : 1);
        check(pl_orb_search_last_frame(h_, &CurrentFrame, &LastFrame, th, bMono ? 1 : 0, mbCheckOrientation ? 1 : 0, match_of_feature.data(), &n));
        match_of_feature.resize(CurrentFrame.n);
        return n;
    }
    // SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1891-2024)
    // KFpoints = pKF->GetMapPointMatches() as a pl_posepoint_view (valid = pMP && !isBad && !sAlreadyFound.count(pMP));
    // Ow = CurrentFrame.mOw, logScaleFactor = CurrentFrame.mfLogScaleFactor; CurrentFrame.claimed[i] = (mvpMapPoints[i] != NULL)
    int SearchByProjection(const pl_frame_view& CurrentFrame, const pl_posepoint_view& KFpoints, const float Ow[3], float logScaleFactor, float th,
                           int ORBdist, std::vector<int>& match_of_feature) {
        int n = 0;
        match_of_feature.resize(CurrentFrame.n > 0 ? CurrentFrame.n : 1);
        int* mo[1] = {match_of_feature.data()};
        check(pl_orb_search_keyframe_points_batch(h_, 1, &CurrentFrame, &KFpoints, Ow, &logScaleFactor, th, ORBdist, mbCheckOrientation ? 1 : 0, mo, &n));
        match_of_feature.resize(CurrentFrame.n);
        return n;
    }
    // SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, vector<MapPoint*> &vpMatched, int th) (ORBmatcher.cc:423-554)
    // KF.tcw = Rcw | tcw with the scale divided out (:435-438), Ow = -Rcw^T tcw (:439), KF.claimed[i] = (vpMatched[i] != NULL)
    int SearchByProjection(const pl_frame_view& KF, const float Ow[3], float logScaleFactor, const pl_posepoint_view& vpPoints, int th,
                           std::vector<int>& match_of_feature) {
        int n = 0;
        match_of_feature.resize(KF.n > 0 ? KF.n : 1);
        int* mo[1] = {match_of_feature.data()};
        check(pl_orb_search_sim3_points_batch(h_, 1, &KF, &vpPoints, Ow, &logScaleFactor, th, mo, &n));
        match_of_feature.resize(KF.n);
        return n;
    }
    // SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches) (ORBmatcher.cc:247-410):
    // match_of_F_feature[j] = index of the key-frame feature whose map point goes to vpMapPointMatches[j], or -1
    int SearchByBoW(const pl_bow_view& KF, const pl_bow_view& F, std::vector<int>& match_of_F_feature) {
        int n = 0;
        match_of_F_feature.resize(F.n > 0 ? F.n : 1);
        int* mo[1] = {match_of_F_feature.data()};
        check(pl_orb_search_bow_batch(h_, 1, &KF, &F, 0, mfNNratio, mbCheckOrientation ? 1 : 0, mo, &n));
        match_of_F_feature.resize(F.n);
        return n;
    }
    // SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12) (ORBmatcher.cc:729-872):
    // match12[idx1] = index of the KF2 feature whose map point goes to vpMatches12[idx1], or -1
    int SearchByBoW(const pl_bow_view& KF1, const pl_bow_view& KF2, std::vector<int>& match12, bool /*keyframe pair*/) {
        int n = 0;
        match12.resize(KF1.n > 0 ? KF1.n : 1);
        int* mo[1] = {match12.data()};
        check(pl_orb_search_bow_batch(h_, 1, &KF1, &KF2, 1, mfNNratio, mbCheckOrientation ? 1 : 0, mo, &n));
        match12.resize(KF1.n);
        return n;
    }
    // Fuse(KeyFrame *pKF, const vector<MapPoint*> &vpMapPoints, th) (ORBmatcher.cc:1107-1277): best_idx[i] = key-frame feature the map point i
    // fuses with, or -1; the caller walks it in order for the Replace / AddObservation bookkeeping (:1251-1272)
    int Fuse(const pl_frame_view& KF, const pl_posepoint_view& vpMapPoints, const float Ow[3], float logScaleFactor, const float* invLevelSigma2,
             float th, std::vector<int>& best_idx) {
        int n = 0;
        best_idx.assign(vpMapPoints.n > 0 ? vpMapPoints.n : 1, -1);
        int* bi[1] = {best_idx.data()};
        const float* sg[1] = {invLevelSigma2};
        check(pl_orb_fuse_candidates_batch(h_, 1, &KF, &vpMapPoints, Ow, &logScaleFactor, sg, th, 0, bi, nullptr, &n));
        best_idx.resize(vpMapPoints.n);
        return n;
    }
    // Fuse(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, float th, vector<MapPoint*> &vpReplacePoint) (ORBmatcher.cc:1290-1427);
    // KF.tcw = Rcw | tcw with the scale divided out, Ow = -Rcw^T tcw (:1299-1303)
    int Fuse(const pl_frame_view& KF, const float Ow[3], float logScaleFactor, const pl_posepoint_view& vpPoints, float th, std::vector<int>& best_idx) {
        int n = 0;
        best_idx.assign(vpPoints.n > 0 ? vpPoints.n : 1, -1);
        int* bi[1] = {best_idx.data()};
        check(pl_orb_fuse_candidates_batch(h_, 1, &KF, &vpPoints, Ow, &logScaleFactor, nullptr, th, 1, bi, nullptr, &n));
        best_idx.resize(vpPoints.n);
        return n;
    }
    // SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12, s12, R12, t12, th) (ORBmatcher.cc:1441-1692):
    // T21 = sR21 | t21 and T12 = sR12 | t12 (:1457-1460); match12[i1] = KF2 feature whose map point goes to vpMatches12[i1], or -1
    int SearchBySim3(const pl_frame_view& KF1, const pl_frame_view& KF2, const pl_posepoint_view& vpMapPoints1, const pl_posepoint_view& vpMapPoints2,
                     const float T21[12], const float T12[12], float logScaleFactor1, float logScaleFactor2, float th, std::vector<int>& match12) {
        int n = 0;
        match12.assign(KF1.n > 0 ? KF1.n : 1, -1);
        check(pl_orb_search_by_sim3(h_, &KF1, &KF2, &vpMapPoints1, &vpMapPoints2, T21, T12, logScaleFactor1, logScaleFactor2, th, match12.data(), &n));
        match12.resize(KF1.n);
        return n;
    }
    // SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched, vector<int> &vnMatches12, int windowSize) (ORBmatcher.cc:573-717)
    int SearchForInitialization(const pl_frame_view& F1, const pl_frame_view& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                int windowSize = 10) {
        int n = 0;
        vnMatches12.assign(F1.n > 0 ? F1.n : 1, -1);
        static_assert(sizeof(cv::Point2f) == 8, "cv::Point2f is two floats");
        check(pl_orb_search_for_initialization(h_, &F1, &F2, (float*)vbPrevMatched.data(), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0,
                                               vnMatches12.data(), &n));
        vnMatches12.resize(F1.n);
        return n;
    }
    // SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, vector<pair<size_t,size_t>> &vMatchedPairs, bOnlyStereo) (ORBmatcher.cc:884-1095)
    int SearchForTriangulation(const pl_triang_view& KF1, const pl_triang_view& KF2, const float F12[9], const float Cw1[3], const float Tcw2[12],
                               float fx2, float fy2, float cx2, float cy2, const float* scaleFactors2, const float* levelSigma2_2, int nLevels2,
                               std::vector<std::pair<size_t, size_t>>& vMatchedPairs, bool bOnlyStereo) {
        int n = 0;
        std::vector<int> pairs(2 * (size_t)(KF1.bow.n > 0 ? KF1.bow.n : 1));
        check(pl_orb_search_for_triangulation(h_, &KF1, &KF2, F12, Cw1, Tcw2, fx2, fy2, cx2, cy2, scaleFactors2, levelSigma2_2, nLevels2,
                                              bOnlyStereo ? 1 : 0, mbCheckOrientation ? 1 : 0, pairs.data(), &n));
        vMatchedPairs.clear();
        for (int i = 0; i < n; i++) vMatchedPairs.push_back(std::make_pair((size_t)pairs[2 * i], (size_t)pairs[2 * i + 1]));
        return n;
    }
    // MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:256-321) / MapLine twin (MapLine.cpp:269-330) for many map points at once
    void ComputeDistinctiveDescriptors(const uint8_t* vDescriptors, const std::vector<int>& group_off, std::vector<int>& best_row) {
        const int g = (int)group_off.size() - 1;
        best_row.assign(g > 0 ? g : 1, -1);
        if (g > 0) check(pl_distinctive_descriptors(h_, vDescriptors, group_off.data(), g, best_row.data()));
        best_row.resize(g > 0 ? g : 0);
    }
    pl_match* handle() { return h_; }

protected:
    void check(int rc) { if (rc != PL_OK) throw std::runtime_error(std::string("ORBmatcher (CUDA): ") + pl_last_error()); }
    float mfNNratio;
    bool mbCheckOrientation;
    pl_match* h_ = nullptr;
};

class LineMatcher {
public:
    LineMatcher(float nnratio = 0.6, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {
        if (pl_match_create(&h_, device) != PL_OK) throw std::runtime_error(std::string("LineMatcher (CUDA): ") + pl_last_error());
    }
    ~LineMatcher() { pl_match_destroy(h_); }
    LineMatcher(const LineMatcher&) = delete;
    LineMatcher& operator=(const LineMatcher&) = delete;

    // the two halves of SearchByProjection(Frame&, const Frame&) / (Frame&, KeyFrame*) / (Frame&, const vector<MapLine*>&)
    // (LineMatcher.cpp:72-269, 527-721, 755-952): projection + clipping of the 3-D lines, then all-pairs LineMatching
    int ProjectLines(const double* start3d, const double* end3d, const cv::line_descriptor::KeyLine* src, const uint8_t* valid, int n,
                     const float Tcw[12], float fx, float fy, float cx, float cy, const float bounds[4], int cols, int rows,
                     std::vector<cv::line_descriptor::KeyLine>& new_KeyLines, std::vector<int>& new_kl_index) {
        new_KeyLines.resize(n > 0 ? n : 1);
        new_kl_index.resize(n > 0 ? n : 1);
        int m = 0;
        check(pl_line_project(h_, start3d, end3d, (const pl_keyline*)src, valid, n, Tcw, fx, fy, cx, cy, bounds[0], bounds[1], bounds[2], bounds[3],
                              cols, rows, (pl_keyline*)new_KeyLines.data(), new_kl_index.data(), &m));
        new_KeyLines.resize(m);
        new_kl_index.resize(m);
        return m;
    }
    int MatchLines(const std::vector<cv::line_descriptor::KeyLine>& new_KeyLines, const uint8_t* new_descriptors,
                   const std::vector<cv::line_descriptor::KeyLine>& cur, const uint8_t* cur_desc, const uint8_t* cur_claimed,
                   std::vector<int>& match_of_line, bool* used_relaxed = nullptr) {
        int n = 0, rel = 0;
        match_of_line.assign(cur.size() ? cur.size() : 1, -1);
        check(pl_line_match_pairs(h_, (const pl_keyline*)new_KeyLines.data(), new_descriptors, (int)new_KeyLines.size(), (const pl_keyline*)cur.data(),
                                  cur_desc, cur_claimed, (int)cur.size(), match_of_line.data(), &n, &rel));
        match_of_line.resize(cur.size());
        if (used_relaxed) *used_relaxed = rel != 0;
        return n;
    }
    // SearchByProjection(Frame &CurrentFrame, KeyFrame *RefFrame, vector<MapLine*> &vpMapLineMatches) (LineMatcher.cpp:489-525):
    // match_of_line[j] = index of the reference line whose MapLine goes to vpMapLineMatches[j], or -1
    int SearchByProjection(const uint8_t* ref_desc, int n_ref, const uint8_t* cur_desc, int n_cur, std::vector<int>& match_of_line) {
        int n = 0;
        match_of_line.assign(n_cur > 0 ? n_cur : 1, -1);
        check(pl_line_match_knn_ratio(h_, ref_desc, n_ref, cur_desc, n_cur, match_of_line.data(), &n));
        match_of_line.resize(n_cur);
        return n;
    }
    // SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, vector<pair<size_t,size_t>> &vMatchedPairs, bOnlyStereo) (LineMatcher.cpp:1174-1204)
    int SearchForTriangulation(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, std::vector<std::pair<size_t, size_t>>& vMatchedPairs) {
        int n = 0;
        std::vector<int> pairs(2 * (size_t)(n1 > 0 ? n1 : 1));
        check(pl_line_search_for_triangulation(h_, desc1, n1, desc2, n2, pairs.data(), &n, nullptr, nullptr));
        vMatchedPairs.clear();
        for (int i = 0; i < n; i++) vMatchedPairs.push_back(std::make_pair((size_t)pairs[2 * i], (size_t)pairs[2 * i + 1]));
        return n;
    }
    // Fuse(KeyFrame *pKF, const vector<MapLine*> &vpMapLines), descriptor half (LineMatcher.cpp:1296-1330): tdx[i] = key-frame line
    // the caller hands to MapLine::Replace, or -1
    int FuseCandidates(const uint8_t* ml_desc, const uint8_t* valid, int n, const uint8_t* kf_desc, int n_kf, std::vector<int>& tdx) {
        int nf = 0;
        tdx.assign(n > 0 ? n : 1, -1);
        check(pl_line_fuse_candidates(h_, ml_desc, valid, n, kf_desc, n_kf, tdx.data(), &nf));
        tdx.resize(n);
        return nf;
    }
    void KnnMatch2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx, int* dist) { check(pl_hamming_knn2(h_, q, nq, t, nt, idx, dist)); }

    // thresholds of LineMatcher.h:94-98 (used inside the CUDA predicate)
    double angle_threshold_ = 15.0 * 3.14159265358979323846 / 180.0, length_threshold_ = 0.45, overlap_threshold_ = 0.5,
           desc_dist_threshold_ = 45, reproj_error_threshold_ = 45;
    pl_match* handle() { return h_; }

protected:
    void check(int rc) { if (rc != PL_OK) throw std::runtime_error(std::string("LineMatcher (CUDA): ") + pl_last_error()); }
    float mfNNratio;
    bool mbCheckOrientation;
    pl_match* h_ = nullptr;
};

}  // namespace PLSLAM_VIEW_NS

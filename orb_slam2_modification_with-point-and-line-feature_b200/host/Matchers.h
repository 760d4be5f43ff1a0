// Matchers.h — host-side entry points for the Hamming searches of ORB_SLAM2::ORBmatcher / LineMatcher.
//
// The reference matchers take Frame / KeyFrame / MapPoint / MapLine objects (include/ORBmatcher.h:74-213,
// include/LineMatcher.h:49-84).  Those classes are outside the hot-path scope, so the mirrors here keep the method
// names and argument meaning but take the POD views of include/plslam_c.h; INTEGRATION.md shows the ten-line adapters
// that fill a view from a reference Frame and write the result back into mvpMapPoints / mvpMapLines.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "plslam_cvlite.h"

namespace ORB_SLAM2 {

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
    // SearchByProjection(Frame&, KeyFrame*, vector<MapLine*>&) brute-force variant (LineMatcher.cpp:492-525): kNN-2
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

}  // namespace ORB_SLAM2

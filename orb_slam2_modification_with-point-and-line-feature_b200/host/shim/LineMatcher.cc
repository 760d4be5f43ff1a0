// shim/LineMatcher.cc — replaces src/LineMatcher.cpp: the reference's LineMatcher methods, the searches on the GPU.
// (Same pattern as shim/ORBmatcher.cc: gather the fields the reference loop reads, one C ABI call, write back.)
#include "LineMatcher.h"

#include <cstring>

namespace ORB_SLAM2 {

LineMatcher::LineMatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri), gpu_(nnratio, checkOri) {}

// LineMatcher.cpp:20-39
int LineMatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    const uint32_t* pa = a.ptr<uint32_t>();
    const uint32_t* pb = b.ptr<uint32_t>();
    int dist = 0;
    for (int i = 0; i < 8; i++) dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}

// Projection, clipping (LiangBarsky), UpdateKeyLineData and the all-pairs LineMatching with its relaxed retry
// (LineMatcher.cpp:96-261 and its two twins) in one call; then the reference's assignments.
int LineMatcher::Search(Frame& Cur, const std::vector<MapLine*>& lines, const std::vector<KeyLine>& src_kl, const std::vector<uint8_t>& valid,
                        std::vector<KeyLine>* new_kls, std::vector<std::pair<int, int>>* match_indices) {
    const int n = (int)lines.size(), nn = n > 0 ? n : 1, NL = Cur.NL, nl = NL > 0 ? NL : 1;
    std::vector<double> s3(3 * (size_t)nn, 0.0), e3(3 * (size_t)nn, 0.0);
    std::vector<uint8_t> desc(32 * (size_t)nn, 0), claimed(nl, 0);
    std::vector<KeyLine> kl(nn);
    for (int i = 0; i < n; i++) {
        if (!valid[i]) continue;
        MapLine* p = lines[i];
        for (int k = 0; k < 3; k++) { s3[3 * i + k] = p->mStart3d[k]; e3[3 * i + k] = p->mEnd3d[k]; }
        std::memcpy(&desc[32 * (size_t)i], p->mLineDescriptor.data, 32);
        kl[i] = src_kl[i];
    }
    static_assert(sizeof(KeyLine) == sizeof(pl_keyline), "cv::line_descriptor::KeyLine layout");
    pl_mapline_view lv = {n, s3.data(), e3.data(), reinterpret_cast<const pl_keyline*>(kl.data()), desc.data(), valid.data()};
    for (int j = 0; j < NL; j++) claimed[j] = Cur.mvpMapLines[j] && Cur.mvpMapLines[j]->Observations() > 0;  // :217-219
    pl_lineframe_view cv_;
    cv_.n = NL;
    cv_.kl = reinterpret_cast<const pl_keyline*>(Cur.mvKeyLinesUn.data());
    cv_.desc = Cur.mLineDescriptors.data;
    cv_.claimed = claimed.data();
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) cv_.tcw[4 * r + c] = Cur.mTcw.at<float>(r, c);  // :77-93
    cv_.fx = Cur.fx; cv_.fy = Cur.fy; cv_.cx = Cur.cx; cv_.cy = Cur.cy;
    cv_.min_x = Cur.mnMinX; cv_.min_y = Cur.mnMinY; cv_.max_x = Cur.mnMaxX; cv_.max_y = Cur.mnMaxY;  // :95-96
    cv_.cols = Cur.im_gray_.cols; cv_.rows = Cur.im_gray_.rows;
    std::vector<int> match(nl, -1), new_idx(nn, -1);
    std::vector<KeyLine> proj(nn);
    int* mo[1] = {match.data()};
    pl_keyline* nk[1] = {reinterpret_cast<pl_keyline*>(proj.data())};
    int* ni[1] = {new_idx.data()};
    int n_matches = 0, relaxed = 0, n_proj = 0;
    if (pl_line_search_by_projection_batch(gpu_.handle(), 1, &cv_, &lv, mo, &n_matches, &relaxed, nk, ni, &n_proj) != PL_OK)
        throw std::runtime_error(std::string("LineMatcher (CUDA): ") + pl_last_error());
    if (relaxed) std::fill(Cur.mvpMapLines.begin(), Cur.mvpMapLines.end(), static_cast<MapLine*>(NULL));  // :239
    for (int j = 0; j < NL; j++)
        if (match[j] >= 0) Cur.mvpMapLines[j] = lines[match[j]];  // :226-228
    if (new_kls) new_kls->assign(proj.begin(), proj.begin() + n_proj);
    if (match_indices) {  // (the test-only overloads: pairs (projected line, current line) of the final assignment)
        match_indices->clear();
        std::vector<int> pos(nn, -1);
        for (int i = 0; i < n_proj; i++) pos[new_idx[i]] = i;
        for (int j = 0; j < NL; j++)
            if (match[j] >= 0) match_indices->push_back(std::make_pair(pos[match[j]], j));
    }
    return n_matches;
}

// LineMatcher.cpp:72-269 (Tracking::TrackWithMotionModel)
int LineMatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame) {
    const size_t n = LastFrame.mvpMapLines.size();
    std::vector<uint8_t> valid(n ? n : 1, 0);
    for (size_t i = 0; i < n; i++) {
        MapLine* p = LastFrame.mvpMapLines[i];
        valid[i] = p && !LastFrame.mvbLineOutlier[i] && !p->isBad();  // :101-107
    }
    return Search(CurrentFrame, LastFrame.mvpMapLines, LastFrame.mvKeyLinesUn, valid, nullptr, nullptr);
}
// LineMatcher.cpp:272-487 (the unit-test twin of the above)
int LineMatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, std::vector<KeyLine>& new_kls, std::vector<std::pair<int, int>>& match_indices) {
    const size_t n = LastFrame.mvpMapLines.size();
    std::vector<uint8_t> valid(n ? n : 1, 0);
    for (size_t i = 0; i < n; i++) {
        MapLine* p = LastFrame.mvpMapLines[i];
        valid[i] = p && !LastFrame.mvbLineOutlier[i] && !p->isBad();
    }
    return Search(CurrentFrame, LastFrame.mvpMapLines, LastFrame.mvKeyLinesUn, valid, &new_kls, &match_indices);
}

// LineMatcher.cpp:492-525 (Tracking::TrackReferenceKeyFrame): knnMatch(ref, cur, 2) + the 0.75 ratio
int LineMatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* RefFrame, std::vector<MapLine*>& vpMapLineMatches) {
    vpMapLineMatches = std::vector<MapLine*>(CurrentFrame.NL, static_cast<MapLine*>(NULL));  // :497
    const std::vector<MapLine*> vpMapLinesKF = RefFrame->GetMapLineMatches();
    std::vector<int> match;
    const int n = gpu_.SearchByProjection(RefFrame->mLineDescriptors.data, RefFrame->mLineDescriptors.rows, CurrentFrame.mLineDescriptors.data,
                                          CurrentFrame.mLineDescriptors.rows, match);
    for (int j = 0; j < CurrentFrame.NL && j < (int)match.size(); j++)
        if (match[j] >= 0) vpMapLineMatches[j] = vpMapLinesKF[match[j]];  // :511
    return n;
}

// LineMatcher.cpp:527-721
int LineMatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* RefFrame) {
    const size_t n = RefFrame->mvpMapLines.size();
    std::vector<uint8_t> valid(n ? n : 1, 0);
    for (size_t i = 0; i < n; i++) valid[i] = RefFrame->mvpMapLines[i] != nullptr;  // :563-565
    return Search(CurrentFrame, RefFrame->mvpMapLines, RefFrame->mvKeyLinesUn, valid, nullptr, nullptr);
}
int LineMatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* RefFrame, std::vector<KeyLine>& new_kls, std::vector<std::pair<int, int>>& match_indices) {
    const size_t n = RefFrame->mvpMapLines.size();
    std::vector<uint8_t> valid(n ? n : 1, 0);
    for (size_t i = 0; i < n; i++) valid[i] = RefFrame->mvpMapLines[i] != nullptr;
    return Search(CurrentFrame, RefFrame->mvpMapLines, RefFrame->mvKeyLinesUn, valid, &new_kls, &match_indices);
}

// LineMatcher.cpp:755-952 (Tracking::SearchLocalLines).  A projected local map line starts from a default KeyLine (`KeyLine proj_kl;`,
// :834, :855, :883): every field LineMatching reads is set by UpdateKeyLineData.
int LineMatcher::SearchByProjection(Frame& F, const std::vector<MapLine*>& vpMapLines) {
    const size_t n = vpMapLines.size();
    std::vector<uint8_t> valid(n ? n : 1, 0);
    std::vector<KeyLine> src(n ? n : 1);
    for (size_t i = 0; i < n; i++) valid[i] = vpMapLines[i]->mbTrackInView && !vpMapLines[i]->isBad();  // :790-800
    return Search(F, vpMapLines, src, valid, nullptr, nullptr);
}
int LineMatcher::SearchByProjection(Frame& F, const std::vector<MapLine*>& vpMapLines, std::vector<KeyLine>& new_kls, std::vector<std::pair<int, int>>& match_indices) {
    const size_t n = vpMapLines.size();
    std::vector<uint8_t> valid(n ? n : 1, 0);
    std::vector<KeyLine> src(n ? n : 1);
    for (size_t i = 0; i < n; i++) valid[i] = vpMapLines[i]->mbTrackInView && !vpMapLines[i]->isBad();
    return Search(F, vpMapLines, src, valid, &new_kls, &match_indices);
}

// LineMatcher.cpp:1174-1204 (LocalMapping::CreateNewMapLines) with KeyFrame::lineDescriptorMAD (KeyFrame.cc:773-797)
int LineMatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<std::pair<size_t, size_t>>& vMatchedPairs, const bool /*bOnlyStereo*/) {
    return gpu_.SearchForTriangulation(pKF1->mLineDescriptors.data, pKF1->mLineDescriptors.rows, pKF2->mLineDescriptors.data, pKF2->mLineDescriptors.rows,
                                       vMatchedPairs);
}

// LineMatcher.cpp:1207-1379 (LocalMapping::SearchInNeighbors), the active branch (:1296-1330)
int LineMatcher::Fuse(KeyFrame* pKF, const std::vector<MapLine*>& vpMapLines) {
    const int n = (int)vpMapLines.size(), nn = n > 0 ? n : 1;
    std::vector<uint8_t> desc(32 * (size_t)nn, 0), valid(nn, 0);
    for (int i = 0; i < n; i++) {
        MapLine* p = vpMapLines[i];
        valid[i] = p != nullptr;  // (isBad / IsInKeyFrame change while the loop below runs: they are evaluated there, :1226-1230)
        if (p) std::memcpy(&desc[32 * (size_t)i], p->mLineDescriptor.data, 32);
    }
    std::vector<int> tdx;
    gpu_.FuseCandidates(desc.data(), valid.data(), n, pKF->mLineDescriptors.data, pKF->mLineDescriptors.rows, tdx);
    int nFused = 0;
    for (int i = 0; i < n; i++) {
        MapLine* pML = vpMapLines[i];
        if (!pML) continue;
        if (pML->isBad() || pML->IsInKeyFrame(pKF)) continue;
        if (tdx[i] < 0) continue;
        MapLine* pMLinKF = pKF->GetMapLine(tdx[i]);  // :1304-1311
        if (pMLinKF) {
            if (!pMLinKF->isBad()) {
                if (pMLinKF->Observations() > pML->Observations()) pML->Replace(pMLinKF);
                else pMLinKF->Replace(pML);
            }
        }
        nFused++;
    }
    return nFused;
}

}  // namespace ORB_SLAM2

// shim/ORBmatcher.cc — replaces src/ORBmatcher.cc: the reference's ORBmatcher methods, the searches on the GPU.
//
// Every method (1) gathers what the reference loop reads from the objects into the POD views of include/plslam_c.h,
// (2) calls the C ABI through plslam_views::ORBmatcher (host/Matchers.h), (3) writes the result where the reference writes it.
// The gather is the part of the reference loop that touches Frame / KeyFrame / MapPoint objects (pointer chasing under their
// mutexes); the candidate search, the Hamming distances and the ordered claims are what runs on the device.
#include "ORBmatcher.h"

#include <cmath>
#include <cstring>

namespace ORB_SLAM2 {

const int ORBmatcher::TH_HIGH = 100;     // ORBmatcher.cc:49-51
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri), gpu_(nnratio, checkOri) {}

// ORBmatcher.cc:2083-2103 (one pair: on the host, it is a handful of popcounts; the batched form is gpu_.DescriptorDistance)
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    const uint32_t* pa = a.ptr<uint32_t>();
    const uint32_t* pb = b.ptr<uint32_t>();
    int dist = 0;
    for (int i = 0; i < 8; i++) dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}

namespace {

// the arrays behind a pl_frame_view
struct FrameArrays {
    std::vector<int> claimed;
    pl_frame_view v;
};
template <class F>
void fill_frame_common(const F& f, FrameArrays& a) {
    std::memset(&a.v, 0, sizeof(a.v));
    a.v.n = f.N;
    static_assert(sizeof(cv::KeyPoint) == sizeof(pl_keypoint), "cv::KeyPoint layout");
    a.v.keys_un = reinterpret_cast<const pl_keypoint*>(f.mvKeysUn.data());
    a.v.desc = f.mDescriptors.data;
    a.v.u_right = f.mvuRight.data();
    a.claimed.assign(f.N > 0 ? f.N : 1, 0);
    a.v.claimed = a.claimed.data();
    a.v.min_x = f.mnMinX; a.v.min_y = f.mnMinY; a.v.max_x = f.mnMaxX; a.v.max_y = f.mnMaxY;
    a.v.fx = f.fx; a.v.fy = f.fy; a.v.cx = f.cx; a.v.cy = f.cy; a.v.bf = f.mbf; a.v.b = f.mb;
    a.v.n_levels = f.mnScaleLevels;
    a.v.scale_factors = f.mvScaleFactors.data();
}
void set_tcw(pl_frame_view& v, const cv::Mat& T) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) v.tcw[4 * r + c] = T.at<float>(r, c);
}
// -R^T t of a 3 x 4 row-major pose, evaluated like the cv::Mat expression `-Rcw.t()*tcw` (double accumulation, one rounding)
void camera_centre(const float tcw[12], float ow[3]) {
    for (int r = 0; r < 3; r++) {
        double a = 0;
        for (int k = 0; k < 3; k++) a += (double)tcw[4 * k + r] * (double)tcw[4 * k + 3];
        ow[r] = (float)-a;
    }
}
// mfMaxDistance from GetMaxDistanceInvariance() = 1.2f * mfMaxDistance (MapPoint.cc:401-405; the member itself is protected)
float max_distance_of(float max_inv) {
    const float q = max_inv / 1.2f;
    if (1.2f * q == max_inv) return q;
    const float lo = std::nextafter(q, 0.0f), hi = std::nextafter(q, 3.0e38f);
    if (1.2f * lo == max_inv) return lo;
    if (1.2f * hi == max_inv) return hi;
    return q;
}
// the arrays behind a pl_posepoint_view
struct PosePoints {
    std::vector<uint8_t> valid, desc;
    std::vector<float> pos, mind, maxd, maxraw, angle, normal;
    pl_posepoint_view v;
    explicit PosePoints(size_t n) : valid(n ? n : 1, 0), desc(32 * (n ? n : 1), 0), pos(3 * (n ? n : 1), 0.f), mind(n ? n : 1, 0.f), maxd(n ? n : 1, 0.f),
                                    maxraw(n ? n : 1, 0.f), angle(n ? n : 1, 0.f), normal(3 * (n ? n : 1), 0.f) {
        v.n = (int)n;
        v.valid = valid.data(); v.world_pos = pos.data(); v.desc = desc.data(); v.min_dist_inv = mind.data(); v.max_dist_inv = maxd.data();
        v.max_dist = maxraw.data(); v.angle = angle.data(); v.normal = normal.data();
    }
    void fill(size_t i, MapPoint* p, bool want_normal) {
        valid[i] = 1;
        const cv::Mat w = p->GetWorldPos();
        for (int k = 0; k < 3; k++) pos[3 * i + k] = w.at<float>(k);
        const cv::Mat d = p->GetDescriptor();
        std::memcpy(&desc[32 * i], d.data, 32);
        mind[i] = p->GetMinDistanceInvariance();
        maxd[i] = p->GetMaxDistanceInvariance();
        maxraw[i] = max_distance_of(maxd[i]);
        if (want_normal) {
            const cv::Mat nv = p->GetNormal();
            for (int k = 0; k < 3; k++) normal[3 * i + k] = nv.at<float>(k);
        }
    }
};
// Scw -> Rcw | tcw with the scale divided out, as ORBmatcher.cc:435-438 / :1299-1302 evaluate it with cv::Mat expressions:
// scw = float(sqrt(row0 . row0)) (double dot product), A / scw = float(double(a) * (1.0 / double(scw)))
void remove_scale(const cv::Mat& Scw, float tcw[12]) {
    double dot = 0;
    for (int c = 0; c < 3; c++) dot += (double)Scw.at<float>(0, c) * (double)Scw.at<float>(0, c);
    const float scw = (float)std::sqrt(dot);
    const double inv = 1.0 / (double)scw;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) tcw[4 * r + c] = (float)((double)Scw.at<float>(r, c) * inv);
}
// a DBoW2::FeatureVector flattened in key order (plslam_c.h, pl_bow_view)
struct BowArrays {
    std::vector<float> angle;
    std::vector<uint8_t> valid;
    std::vector<unsigned int> node_id, feat_idx;
    std::vector<int> node_off;
    pl_bow_view v;
    template <class FV>
    void flatten(const FV& fv) {
        node_off.push_back(0);
        for (typename FV::const_iterator it = fv.begin(); it != fv.end(); ++it) {
            node_id.push_back((unsigned int)it->first);
            feat_idx.insert(feat_idx.end(), it->second.begin(), it->second.end());
            node_off.push_back((int)feat_idx.size());
        }
        if (feat_idx.empty()) feat_idx.push_back(0);
        if (node_id.empty()) node_id.push_back(0);
        v.n_nodes = (int)node_off.size() - 1;
        v.node_id = node_id.data();
        v.node_off = node_off.data();
        v.feat_idx = feat_idx.data();
    }
};
template <class F>
void fill_bow(F& f, const std::vector<cv::KeyPoint>& keys, const std::vector<MapPoint*>* pts, bool valid_is_null, BowArrays& a) {
    a.v.n = f.N;
    a.angle.resize(f.N > 0 ? f.N : 1);
    for (int i = 0; i < f.N; i++) a.angle[i] = keys[i].angle;
    a.v.angle = a.angle.data();
    a.v.desc = f.mDescriptors.data;
    a.v.valid = nullptr;
    if (pts) {
        a.valid.assign(f.N > 0 ? f.N : 1, 0);
        for (int i = 0; i < f.N; i++) {
            MapPoint* p = (*pts)[i];
            a.valid[i] = valid_is_null ? (p == nullptr) : (p && !p->isBad());
        }
        a.v.valid = a.valid.data();
    }
    a.flatten(f.mFeatVec);
}

}  // namespace

// ORBmatcher.cc:72-193 (Tracking::SearchLocalPoints)
int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    FrameArrays fa;
    fill_frame_common(F, fa);
    for (int i = 0; i < F.N; i++) fa.claimed[i] = F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0;  // :128-130
    const size_t m = vpMapPoints.size(), mm = m ? m : 1;
    std::vector<uint8_t> desc(32 * mm), inview(mm, 0), hasobs(mm, 0);
    std::vector<float> px(mm, 0.f), py(mm, 0.f), pxr(mm, 0.f), vcos(mm, 0.f);
    std::vector<int> lvl(mm, 0);
    for (size_t i = 0; i < m; i++) {
        MapPoint* p = vpMapPoints[i];
        inview[i] = p->mbTrackInView && !p->isBad();  // :83-88
        if (!inview[i]) continue;
        px[i] = p->mTrackProjX; py[i] = p->mTrackProjY; pxr[i] = p->mTrackProjXR;
        lvl[i] = p->mnTrackScaleLevel; vcos[i] = p->mTrackViewCos;
        const cv::Mat d = p->GetDescriptor();
        std::memcpy(&desc[32 * i], d.data, 32);
        hasobs[i] = p->Observations() > 0;
    }
    pl_mappoint_view mv = {(int)m, desc.data(), inview.data(), px.data(), py.data(), pxr.data(), lvl.data(), vcos.data(), hasobs.data()};
    std::vector<int> match;
    const int n = gpu_.SearchByProjection(fa.v, mv, th, match);
    for (int i = 0; i < F.N; i++)
        if (match[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[match[i]];  // :172
    return n;
}

// ORBmatcher.cc:1710-1879 (Tracking::TrackWithMotionModel; the caller clears mvpMapPoints first, Tracking.cc:1244)
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    FrameArrays fa;
    fill_frame_common(CurrentFrame, fa);
    set_tcw(fa.v, CurrentFrame.mTcw);
    std::vector<uint8_t> had(CurrentFrame.N > 0 ? CurrentFrame.N : 1, 0);
    for (int i = 0; i < CurrentFrame.N; i++) {
        MapPoint* p = CurrentFrame.mvpMapPoints[i];
        fa.claimed[i] = p && p->Observations() > 0;  // :1807-1809
        had[i] = p != nullptr;
    }
    const int m = LastFrame.N, mm = m > 0 ? m : 1;
    std::vector<uint8_t> valid(mm, 0), hasobs(mm, 0), desc(32 * (size_t)mm);
    std::vector<float> X(3 * (size_t)mm, 0.f), ang(mm, 0.f);
    std::vector<int> oct(mm, 0);
    for (int i = 0; i < m; i++) {
        MapPoint* p = LastFrame.mvpMapPoints[i];
        valid[i] = p && !LastFrame.mvbOutlier[i];  // :1748-1752
        if (!valid[i]) continue;
        const cv::Mat w = p->GetWorldPos();
        for (int k = 0; k < 3; k++) X[3 * i + k] = w.at<float>(k);
        const cv::Mat d = p->GetDescriptor();
        std::memcpy(&desc[32 * (size_t)i], d.data, 32);
        oct[i] = LastFrame.mvKeys[i].octave;     // :1781
        ang[i] = LastFrame.mvKeysUn[i].angle;    // :1851
        hasobs[i] = p->Observations() > 0;
    }
    pl_lastframe_view lv;
    lv.n = m; lv.valid = valid.data(); lv.world_pos = X.data(); lv.desc = desc.data(); lv.octave = oct.data(); lv.angle = ang.data();
    lv.has_observations = hasobs.data();
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) lv.tcw[4 * r + c] = LastFrame.mTcw.at<float>(r, c);
    std::vector<int> match;
    const int n = gpu_.SearchByProjection(fa.v, lv, th, bMono, match);
    for (int i2 = 0; i2 < CurrentFrame.N; i2++) {
        if (match[i2] >= 0) CurrentFrame.mvpMapPoints[i2] = LastFrame.mvpMapPoints[match[i2]];  // :1844
        else if (!had[i2]) CurrentFrame.mvpMapPoints[i2] = static_cast<MapPoint*>(NULL);        // (a match the rotation check removed, :1868-1871)
    }
    return n;
}

// ORBmatcher.cc:1891-2024 (Tracking::Relocalization)
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist) {
    FrameArrays fa;
    fill_frame_common(CurrentFrame, fa);
    set_tcw(fa.v, CurrentFrame.mTcw);
    for (int i = 0; i < CurrentFrame.N; i++) fa.claimed[i] = CurrentFrame.mvpMapPoints[i] != nullptr;  // :1965
    float ow[3];
    camera_centre(fa.v.tcw, ow);  // :1897
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    PosePoints pp(vpMPs.size());
    for (size_t i = 0; i < vpMPs.size(); i++) {
        MapPoint* p = vpMPs[i];
        if (p && !p->isBad() && !sAlreadyFound.count(p)) {  // :1913-1918
            pp.fill(i, p, false);
            pp.angle[i] = pKF->mvKeysUn[i].angle;  // :1976
        }
    }
    pp.v.normal = nullptr;
    std::vector<int> match;
    const int n = gpu_.SearchByProjection(fa.v, pp.v, ow, CurrentFrame.mfLogScaleFactor, th, ORBdist, match);
    for (int i2 = 0; i2 < CurrentFrame.N; i2++)
        if (match[i2] >= 0) CurrentFrame.mvpMapPoints[i2] = vpMPs[match[i2]];  // :1970
    return n;
}

// ORBmatcher.cc:423-554 (LoopClosing::ComputeSim3)
int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th) {
    FrameArrays fa;
    fill_frame_common(*pKF, fa);
    remove_scale(Scw, fa.v.tcw);  // :435-438
    float ow[3];
    camera_centre(fa.v.tcw, ow);  // :439
    for (int i = 0; i < pKF->N; i++) fa.claimed[i] = vpMatched[i] != nullptr;  // :521
    std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());  // :442-443
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
    PosePoints pp(vpPoints.size());
    for (size_t i = 0; i < vpPoints.size(); i++) {
        MapPoint* p = vpPoints[i];
        if (!p->isBad() && !spAlreadyFound.count(p)) pp.fill(i, p, true);  // :455-459
    }
    pp.v.angle = nullptr;
    std::vector<int> match;
    const int n = gpu_.SearchByProjection(fa.v, ow, pKF->mfLogScaleFactor, pp.v, th, match);
    for (int idx = 0; idx < pKF->N; idx++)
        if (match[idx] >= 0) vpMatched[idx] = vpPoints[match[idx]];  // :545
    return n;
}

// ORBmatcher.cc:247-410 (Tracking::TrackReferenceKeyFrame, Relocalization)
int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));  // :251
    BowArrays a, b;
    fill_bow(*pKF, pKF->mvKeysUn, &vpMapPointsKF, false, a);  // :301-305, angle :349
    fill_bow(F, F.mvKeys, nullptr, false, b);                 // angle :349
    std::vector<int> match;
    const int n = gpu_.SearchByBoW(a.v, b.v, match);
    for (int j = 0; j < F.N; j++)
        if (match[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[match[j]];  // :344
    return n;
}

// ORBmatcher.cc:729-872 (LoopClosing::ComputeSim3)
int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = std::vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));  // :746
    BowArrays a, b;
    fill_bow(*pKF1, pKF1->mvKeysUn, &vpMapPoints1, false, a);  // :775-780
    fill_bow(*pKF2, pKF2->mvKeysUn, &vpMapPoints2, false, b);  // :797-802
    std::vector<int> match;
    const int n = gpu_.SearchByBoW(a.v, b.v, match, true);
    for (int i1 = 0; i1 < pKF1->N; i1++)
        if (match[i1] >= 0) vpMatches12[i1] = vpMapPoints2[match[i1]];  // :820
    return n;
}

// ORBmatcher.cc:573-717 (Tracking::MonocularInitialization)
int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize) {
    FrameArrays a, b;
    fill_frame_common(F1, a);
    fill_frame_common(F2, b);
    return gpu_.SearchForInitialization(a.v, b.v, vbPrevMatched, vnMatches12, windowSize);
}

// ORBmatcher.cc:884-1095 (LocalMapping::CreateNewMapPoints)
int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t>>& vMatchedPairs,
                                       const bool bOnlyStereo) {
    const std::vector<MapPoint*> mp1 = pKF1->GetMapPointMatches(), mp2 = pKF2->GetMapPointMatches();
    BowArrays a, b;
    fill_bow(*pKF1, pKF1->mvKeysUn, &mp1, true, a);  // :938-940
    fill_bow(*pKF2, pKF2->mvKeysUn, &mp2, true, b);  // :964-966
    pl_triang_view t1 = {a.v, reinterpret_cast<const pl_keypoint*>(pKF1->mvKeysUn.data()), pKF1->mvuRight.data()};
    pl_triang_view t2 = {b.v, reinterpret_cast<const pl_keypoint*>(pKF2->mvKeysUn.data()), pKF2->mvuRight.data()};
    float f12[9], cw1[3], tcw2[12];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) f12[3 * r + c] = F12.at<float>(r, c);
    const cv::Mat Cw = pKF1->GetCameraCenter();  // :890
    for (int k = 0; k < 3; k++) cw1[k] = Cw.at<float>(k);
    const cv::Mat R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();  // :891-892
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) tcw2[4 * r + c] = R2w.at<float>(r, c);
        tcw2[4 * r + 3] = t2w.at<float>(r);
    }
    return gpu_.SearchForTriangulation(t1, t2, f12, cw1, tcw2, pKF2->fx, pKF2->fy, pKF2->cx, pKF2->cy, pKF2->mvScaleFactors.data(),
                                       pKF2->mvLevelSigma2.data(), pKF2->mnScaleLevels, vMatchedPairs, bOnlyStereo);
}

// ORBmatcher.cc:1441-1692 (LoopClosing::ComputeSim3)
int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12,
                             const float th) {
    FrameArrays k1, k2;
    fill_frame_common(*pKF1, k1);
    fill_frame_common(*pKF2, k2);
    const cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();  // :1450-1454
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) { k1.v.tcw[4 * r + c] = R1w.at<float>(r, c); k2.v.tcw[4 * r + c] = R2w.at<float>(r, c); }
        k1.v.tcw[4 * r + 3] = t1w.at<float>(r);
        k2.v.tcw[4 * r + 3] = t2w.at<float>(r);
    }
    // :1457-1460 as cv::Mat evaluates them: s * A = float(double(a) * double(s)); -A * b = float(-(sum of double products))
    float T12[12], T21[12];
    const double s = (double)s12, inv_s = 1.0 / (double)s12;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            T12[4 * r + c] = (float)((double)R12.at<float>(r, c) * s);
            T21[4 * r + c] = (float)((double)R12.at<float>(c, r) * inv_s);
        }
    for (int r = 0; r < 3; r++) {
        T12[4 * r + 3] = t12.at<float>(r);
        double acc = 0;
        for (int k = 0; k < 3; k++) acc += (double)T21[4 * r + k] * (double)t12.at<float>(k);
        T21[4 * r + 3] = (float)-acc;
    }
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
    std::vector<bool> m1(N1, false), m2(N2, false);  // :1470-1489
    for (int i = 0; i < N1; i++) {
        MapPoint* p = vpMatches12[i];
        if (p) {
            m1[i] = true;
            const int idx2 = p->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) m2[idx2] = true;
        }
    }
    PosePoints p1(N1), p2(N2);
    for (int i = 0; i < N1; i++) {
        MapPoint* p = vpMapPoints1[i];
        if (p && !m1[i] && !p->isBad()) p1.fill(i, p, false);  // :1500-1507
    }
    for (int i = 0; i < N2; i++) {
        MapPoint* p = vpMapPoints2[i];
        if (p && !m2[i] && !p->isBad()) p2.fill(i, p, false);  // :1576-1582
    }
    p1.v.angle = p2.v.angle = nullptr;
    p1.v.normal = p2.v.normal = nullptr;
    std::vector<int> match12;
    const int n = gpu_.SearchBySim3(k1.v, k2.v, p1.v, p2.v, T21, T12, pKF1->mfLogScaleFactor, pKF2->mfLogScaleFactor, th, match12);
    for (int i1 = 0; i1 < N1; i1++)
        if (match12[i1] >= 0) vpMatches12[i1] = vpMapPoints2[match12[i1]];  // :1684
    return n;
}

// ORBmatcher.cc:1107-1277 (LocalMapping::SearchInNeighbors)
int ORBmatcher::Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    FrameArrays fa;
    fill_frame_common(*pKF, fa);
    const cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), Ow = pKF->GetCameraCenter();  // :1110-1122
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) fa.v.tcw[4 * r + c] = Rcw.at<float>(r, c);
        fa.v.tcw[4 * r + 3] = tcw.at<float>(r);
    }
    const float ow[3] = {Ow.at<float>(0), Ow.at<float>(1), Ow.at<float>(2)};
    PosePoints pp(vpMapPoints.size());
    for (size_t i = 0; i < vpMapPoints.size(); i++) {
        MapPoint* p = vpMapPoints[i];
        if (p && !p->isBad() && !p->IsInKeyFrame(pKF)) pp.fill(i, p, true);  // :1128-1136
    }
    pp.v.angle = nullptr;
    std::vector<int> best;
    gpu_.Fuse(fa.v, pp.v, ow, pKF->mfLogScaleFactor, pKF->mvInvLevelSigma2.data(), th, best);
    int nFused = 0;
    for (size_t i = 0; i < vpMapPoints.size(); i++) {  // :1249-1272, in the reference's order; an earlier Replace may have changed a later point
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP || best[i] < 0 || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        MapPoint* pMPinKF = pKF->GetMapPoint(best[i]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, best[i]);
            pKF->AddMapPoint(pMP, best[i]);
        }
        nFused++;
    }
    return nFused;
}

// ORBmatcher.cc:1290-1427 (LoopClosing::SearchAndFuse)
int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint) {
    FrameArrays fa;
    fill_frame_common(*pKF, fa);
    remove_scale(Scw, fa.v.tcw);  // :1299-1302
    float ow[3];
    camera_centre(fa.v.tcw, ow);  // :1303
    const std::set<MapPoint*> spAlreadyFound = [&] {  // :1306
        const std::vector<MapPoint*> v = pKF->GetMapPointMatches();
        return std::set<MapPoint*>(v.begin(), v.end());
    }();
    PosePoints pp(vpPoints.size());
    for (size_t i = 0; i < vpPoints.size(); i++) {
        MapPoint* p = vpPoints[i];
        if (!p->isBad() && !spAlreadyFound.count(p)) pp.fill(i, p, true);  // :1317-1320
    }
    pp.v.angle = nullptr;
    std::vector<int> best;
    gpu_.Fuse(fa.v, ow, pKF->mfLogScaleFactor, pp.v, th, best);
    int nFused = 0;
    for (size_t i = 0; i < vpPoints.size(); i++) {  // :1404-1420
        if (best[i] < 0) continue;
        MapPoint* pMP = vpPoints[i];
        MapPoint* pMPinKF = pKF->GetMapPoint(best[i]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) vpReplacePoint[i] = pMPinKF;
        } else {
            pMP->AddObservation(pKF, best[i]);
            pKF->AddMapPoint(pMP, best[i]);
        }
        nFused++;
    }
    return nFused;
}

}  // namespace ORB_SLAM2

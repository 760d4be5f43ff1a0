// Drives the matcher shims (package host/shim/ORBmatcher.h, LineMatcher.h: the reference's own signatures) the way
// Tracking.cc does — SearchByProjection(CurrentFrame, LastFrame, th, bMono) (:1244), SearchByProjection(F, vpMapPoints, th) (:1812),
// LineMatcher::SearchByProjection(CurrentFrame, LastFrame) (:1247) — on Frame / MapPoint / MapLine objects built from the raw
// arrays tests/test_matcher_shim_gpu.py wrote, and prints what the calls left in mvpMapPoints / mvpMapLines.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../orb_slam2_modification_with-point-and-line-feature_b200/host/shim/LineMatcher.h"
#include "../../orb_slam2_modification_with-point-and-line-feature_b200/host/shim/ORBmatcher.h"

using namespace ORB_SLAM2;

static std::string g_dir;
template <typename T>
static std::vector<T> load(const char* name) {
    const std::string p = g_dir + "/" + name + ".bin";
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", p.c_str()); exit(2); }
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<T> v((size_t)sz / sizeof(T));
    if (sz && fread(v.data(), 1, (size_t)sz, f) != (size_t)sz) { fprintf(stderr, "short read %s\n", p.c_str()); exit(2); }
    fclose(f);
    return v;
}
static cv::Mat mat_f(const float* p, int r, int c) {
    cv::Mat m(r, c, cv::CV_32F);
    for (int i = 0; i < r; i++)
        for (int j = 0; j < c; j++) m.at<float>(i, j) = p[i * c + j];
    return m;
}
static cv::Mat mat_u8(const uint8_t* p, int r, int c) {
    cv::Mat m(r, c, cv::CV_8U);
    if (r * c) memcpy(m.data, p, (size_t)r * c);
    return m;
}
static void camera(Frame& F, const std::vector<float>& K, const std::vector<float>& sf) {
    F.fx = K[0]; F.fy = K[1]; F.cx = K[2]; F.cy = K[3]; F.mbf = K[4]; F.mb = K[5];
    F.mnMinX = K[6]; F.mnMinY = K[7]; F.mnMaxX = K[8]; F.mnMaxY = K[9];
    F.mnScaleLevels = (int)sf.size();
    F.mvScaleFactors = sf;
    F.mfScaleFactor = sf.size() > 1 ? sf[1] : 1.f;
    F.mfLogScaleFactor = logf(F.mfScaleFactor);
}
template <typename P>
static void print_idx(const char* tag, int n, const std::vector<P*>& v, const P* base) {
    printf("%s %d", tag, n);
    for (size_t i = 0; i < v.size(); i++) printf(" %d", v[i] ? (int)(v[i] - base) : -1);
    printf("\n");
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s dir th_c3 [th_c2]\n", argv[0]); return 2; }
    g_dir = argv[1];
    const float th3 = (float)atof(argv[2]), th2 = argc > 3 ? (float)atof(argv[3]) : 3.f;
    const std::vector<float> K = load<float>("K"), sf = load<float>("scale_factors");
    // ---- the map ----
    const std::vector<float> pw = load<float>("pool_pos"), pcos = load<float>("pool_view_cos"), ppx = load<float>("pool_proj_x"),
                             ppy = load<float>("pool_proj_y"), ppxr = load<float>("pool_proj_xr");
    const std::vector<uint8_t> pd = load<uint8_t>("pool_desc"), pbad = load<uint8_t>("pool_bad"), pview = load<uint8_t>("pool_in_view");
    const std::vector<int> pobs = load<int>("pool_nobs"), plvl = load<int>("pool_level");
    const int P = (int)pobs.size();
    std::vector<MapPoint> pool(P > 0 ? P : 1);
    for (int i = 0; i < P; i++) {
        pool[i].mWorldPos = mat_f(&pw[3 * i], 3, 1);
        pool[i].mDescriptor = mat_u8(&pd[32 * i], 1, 32);
        pool[i].nObs = pobs[i];
        pool[i].mbBad = pbad[i] != 0;
        pool[i].mbTrackInView = pview[i] != 0;
        pool[i].mTrackProjX = ppx[i]; pool[i].mTrackProjY = ppy[i]; pool[i].mTrackProjXR = ppxr[i];
        pool[i].mnTrackScaleLevel = plvl[i];
        pool[i].mTrackViewCos = pcos[i];
    }
    const std::vector<double> ls3 = load<double>("lpool_start"), le3 = load<double>("lpool_end");
    const std::vector<uint8_t> ld = load<uint8_t>("lpool_desc"), lbad = load<uint8_t>("lpool_bad");
    const std::vector<int> lobs = load<int>("lpool_nobs");
    const int PL = (int)lobs.size();
    std::vector<MapLine> lpool(PL > 0 ? PL : 1);
    for (int i = 0; i < PL; i++) {
        lpool[i].mStart3d = Eigen::Vector3d(ls3[3 * i], ls3[3 * i + 1], ls3[3 * i + 2]);
        lpool[i].mEnd3d = Eigen::Vector3d(le3[3 * i], le3[3 * i + 1], le3[3 * i + 2]);
        lpool[i].mLineDescriptor = mat_u8(&ld[32 * i], 1, 32);
        lpool[i].nObs = lobs[i];
        lpool[i].mbBad = lbad[i] != 0;
    }
    // ---- the two frames ----
    Frame Cur, Last;
    camera(Cur, K, sf);
    camera(Last, K, sf);
    {
        const std::vector<cv::KeyPoint> ku = load<cv::KeyPoint>("cur_keys_un");
        const std::vector<uint8_t> d = load<uint8_t>("cur_desc");
        const std::vector<int> mp = load<int>("cur_mp");
        Cur.N = (int)ku.size();
        Cur.mvKeys = Cur.mvKeysUn = ku;
        Cur.mDescriptors = mat_u8(d.data(), Cur.N, 32);
        Cur.mvuRight = load<float>("cur_u_right");
        Cur.mvpMapPoints.assign(Cur.N, nullptr);
        for (int i = 0; i < Cur.N; i++) Cur.mvpMapPoints[i] = mp[i] >= 0 ? &pool[mp[i]] : nullptr;
        Cur.mvbOutlier.assign(Cur.N, false);
        Cur.mTcw = mat_f(load<float>("cur_tcw").data(), 4, 4);
        const std::vector<int> img = load<int>("img_size");
        Cur.im_gray_ = cv::Mat(img[1], img[0], cv::CV_8U);
        const std::vector<KeyLine> kl = load<KeyLine>("cur_kl_un");
        const std::vector<uint8_t> cld = load<uint8_t>("cur_ldesc");
        const std::vector<int> ml = load<int>("cur_ml");
        Cur.NL = (int)kl.size();
        Cur.mvKeyLines = Cur.mvKeyLinesUn = kl;
        Cur.mLineDescriptors = mat_u8(cld.data(), Cur.NL, 32);
        Cur.mvpMapLines.assign(Cur.NL, nullptr);
        for (int j = 0; j < Cur.NL; j++) Cur.mvpMapLines[j] = ml[j] >= 0 ? &lpool[ml[j]] : nullptr;
        Cur.mvbLineOutlier.assign(Cur.NL, false);
    }
    {
        const std::vector<cv::KeyPoint> k = load<cv::KeyPoint>("last_keys"), ku = load<cv::KeyPoint>("last_keys_un");
        const std::vector<int> mp = load<int>("last_mp");
        const std::vector<uint8_t> out = load<uint8_t>("last_outlier");
        Last.N = (int)k.size();
        Last.mvKeys = k;
        Last.mvKeysUn = ku;
        Last.mvpMapPoints.assign(Last.N, nullptr);
        Last.mvbOutlier.assign(Last.N, false);
        for (int i = 0; i < Last.N; i++) {
            Last.mvpMapPoints[i] = mp[i] >= 0 ? &pool[mp[i]] : nullptr;
            Last.mvbOutlier[i] = out[i] != 0;
        }
        Last.mTcw = mat_f(load<float>("last_tcw").data(), 4, 4);
        const std::vector<KeyLine> kl = load<KeyLine>("last_kl_un");
        const std::vector<int> ml = load<int>("last_ml");
        const std::vector<uint8_t> lo = load<uint8_t>("last_line_outlier");
        Last.NL = (int)kl.size();
        Last.mvKeyLinesUn = kl;
        Last.mvpMapLines.assign(Last.NL, nullptr);
        Last.mvbLineOutlier.assign(Last.NL, false);
        for (int i = 0; i < Last.NL; i++) {
            Last.mvpMapLines[i] = ml[i] >= 0 ? &lpool[ml[i]] : nullptr;
            Last.mvbLineOutlier[i] = lo[i] != 0;
        }
    }
    // ---- Tracking::TrackWithMotionModel (Tracking.cc:1244-1247) ----
    ORBmatcher matcher(0.9f, true);
    int n3 = matcher.SearchByProjection(Cur, Last, th3, false);
    print_idx("c3", n3, Cur.mvpMapPoints, pool.data());
    LineMatcher lmatcher(0.9f, true);
    int nl = lmatcher.SearchByProjection(Cur, Last);
    print_idx("d3", nl, Cur.mvpMapLines, lpool.data());
    // ---- Tracking::SearchLocalPoints (Tracking.cc:1805-1812) ----
    const std::vector<int> local = load<int>("local_points");
    std::vector<MapPoint*> vpMapPoints;
    for (int i : local) vpMapPoints.push_back(&pool[i]);
    ORBmatcher matcher2(0.8f);
    int n2 = matcher2.SearchByProjection(Cur, vpMapPoints, th2);
    print_idx("c2", n2, Cur.mvpMapPoints, pool.data());
    // ---- Tracking::SearchLocalLines (Tracking.cc:1863) ----
    const std::vector<int> llocal = load<int>("local_lines");
    const std::vector<uint8_t> lview = load<uint8_t>("lpool_in_view");
    std::vector<MapLine*> vpMapLines;
    for (int i : llocal) { lpool[i].mbTrackInView = lview[i] != 0; vpMapLines.push_back(&lpool[i]); }
    LineMatcher lmatcher2(0.8f);
    int nl5 = lmatcher2.SearchByProjection(Cur, vpMapLines);
    print_idx("d5", nl5, Cur.mvpMapLines, lpool.data());
    printf("dist %d\n", ORBmatcher::DescriptorDistance(pool[0].mDescriptor, Cur.N ? Cur.mDescriptors.row(0) : pool[0].mDescriptor));
    return 0;
}

// frame_oracle.cpp — CPU restatement of the per-feature maps of ORB_SLAM2::Frame that sit between the extractors and the
// matchers.  TEST INFRASTRUCTURE ONLY: used by tests/, smoke() and bench.py's CPU legs as the checker
// of the CUDA path, never by the product.
//
//   Frame::UndistortKeyPoints        src/Frame.cc:737-765   -> cv::undistortPoints (OpenCV calib3d/imgproc undistort.cpp,
//                                                              un-vendored; restated here and PINNED bit-exactly against
//                                                              python cv2 4.13 by tests/test_oracle_frame.py)
//   Frame::ComputeStereoFromRGBD     src/Frame.cc:1065-1117
//   Frame::UnprojectStereo           src/Frame.cc:1120-1134 (and the KeyLine end-point twins :1140-1205)
//   Frame::IsInFrustum(MapPoint*)    src/Frame.cc:345-401, MapPoint::PredictScale src/MapPoint.cc:416-431
//   Frame::IsInFrustum(MapLine*)     src/Frame.cc:403-430
// ComputeStereoFromRGBD, UnprojectStereo and IsInFrustum(MapPoint*) are PINNED bit-exactly to the reference's own functions
// (oracle/_ref, tests/test_oracle_ref.py); the MapLine twin and ComputeStereoMatches are restated only.
#include <cmath>
#include <cstdint>
#include <cstring>

#include "oracle.h"

// cv::undistortPoints(src, dst, K, D, noArray(), P = K): cvUndistortPointsInternal with TermCriteria(MAX_ITER, 5, 0.01):
// five fixed-point iterations of the inverse Brown model in double, then re-projection with P
extern "C" int orc_frame_undistort_points(const float* xy, int n, float fx_, float fy_, float cx_, float cy_, const float* dist_coef, float* xy_out) {
    if (dist_coef[0] == 0.0f) {  // Frame.cc:739-743
        memcpy(xy_out, xy, (size_t)n * 8);
        return 0;
    }
    const double fx = fx_, fy = fy_, cx = cx_, cy = cy_, ifx = 1. / fx, ify = 1. / fy;
    double k[12] = {0};
    for (int i = 0; i < 5; i++) k[i] = dist_coef[i];  // k1 k2 p1 p2 k3 ; k4..k6, s1..s4 = 0
    for (int i = 0; i < n; i++) {
        double x = xy[2 * i], y = xy[2 * i + 1];
        const double u = x, v = y;
        x = (x - cx) * ifx;
        y = (y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) {
                x = (u - cx) * ifx;
                y = (v - cy) * ify;
                break;
            }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        // RR = P * I = K
        const double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
        xy_out[2 * i] = (float)(xx * ww);
        xy_out[2 * i + 1] = (float)(yy * ww);
    }
    return 0;
}

extern "C" int orc_frame_stereo_from_rgbd_batch(int n_frames, const float* depth, int rows, int cols, size_t step_bytes, size_t frame_stride_bytes,
                                                const int* off, const float* xy, const float* x_un, float bf, float* depth_out, float* u_right_out) {
    (void)rows; (void)cols;
    for (int f = 0; f < n_frames; f++) {
        const uint8_t* img = (const uint8_t*)depth + (size_t)f * frame_stride_bytes;
        for (int i = off[f]; i < off[f + 1]; i++) {
            depth_out[i] = -1;
            u_right_out[i] = -1;
            const float v = xy[2 * i + 1], u = xy[2 * i];
            const float d = *(const float*)(img + (size_t)(int)v * step_bytes + (size_t)(int)u * 4);  // imDepth.at<float>(v, u): int(v), int(u)
            if (d > 0) {
                depth_out[i] = d;
                u_right_out[i] = x_un[i] - bf / d;
            }
        }
    }
    return 0;
}

namespace {
// cv::Mat (CV_32F) R*x + t: one gemm, double accumulation, rounded once
inline void gemm3(const float* R, int ldr, const float* t, const float* x, float* out) {
    for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double)R[ldr * r + k] * (double)x[k];
        out[r] = (float)(s * 1.0 + (double)t[r] * 1.0);
    }
}
}  // namespace

extern "C" int orc_frame_unproject_batch(int n_frames, const int* off, const float* xy_un, const float* z, const float* rwc, const float* ow, float fx,
                                         float fy, float cx, float cy, float* world, uint8_t* valid) {
    const float invfx = 1.0f / fx, invfy = 1.0f / fy;  // Frame.cc:180-181
    for (int f = 0; f < n_frames; f++)
        for (int i = off[f]; i < off[f + 1]; i++) {
            world[3 * i] = world[3 * i + 1] = world[3 * i + 2] = 0;
            valid[i] = 0;
            const float zz = z[i];
            if (zz > 0) {
                const float u = xy_un[2 * i], v = xy_un[2 * i + 1];
                const float x3Dc[3] = {(u - cx) * zz * invfx, (v - cy) * zz * invfy, zz};
                gemm3(rwc + 9 * (size_t)f, 3, ow + 3 * (size_t)f, x3Dc, world + 3 * (size_t)i);
                valid[i] = 1;
            }
        }
    return 0;
}

extern "C" int orc_frame_is_in_frustum_batch(int n_frames, const float* tcw, const float* ow, float fx, float fy, float cx, float cy, float bf,
                                             const float* bounds, int n_levels, float log_scale_factor, int m, const float* world_pos,
                                             const float* normal, const float* min_dist_inv, const float* max_dist_inv, const float* max_dist,
                                             float viewing_cos_limit, uint8_t* in_view, float* proj_x, float* proj_y, float* proj_xr,
                                             int* scale_level, float* view_cos) {
    for (int f = 0; f < n_frames; f++) {
        const float* T = tcw + 12 * (size_t)f;
        const float t[3] = {T[3], T[7], T[11]};
        const float* Ow = ow + 3 * (size_t)f;
        for (int j = 0; j < m; j++) {
            const size_t o = (size_t)f * m + j;
            in_view[o] = 0;
            proj_x[o] = proj_y[o] = proj_xr[o] = view_cos[o] = 0;
            scale_level[o] = 0;
            const float* P = world_pos + 3 * (size_t)j;
            float Pc[3];
            gemm3(T, 4, t, P, Pc);
            if (Pc[2] < 0.0f) continue;
            const float invz = 1.0f / Pc[2];
            const float u = fx * Pc[0] * invz + cx;
            const float v = fy * Pc[1] * invz + cy;
            if (u < bounds[0] || u > bounds[2]) continue;
            if (v < bounds[1] || v > bounds[3]) continue;
            const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
            double ss = 0;
            for (int k = 0; k < 3; k++) ss += (double)PO[k] * PO[k];
            const float dist = (float)std::sqrt(ss);  // cv::norm
            if (dist < min_dist_inv[j] || dist > max_dist_inv[j]) continue;
            double dot = 0;
            for (int k = 0; k < 3; k++) dot += (double)PO[k] * normal[3 * (size_t)j + k];
            const float viewCos = (float)(dot / dist);
            if (viewCos < viewing_cos_limit) continue;
            const float ratio = max_dist[j] / dist;  // MapPoint::PredictScale
            int nScale = (int)std::ceil(std::log(ratio) / log_scale_factor);
            if (nScale < 0) nScale = 0;
            else if (nScale >= n_levels) nScale = n_levels - 1;
            in_view[o] = 1;
            proj_x[o] = u;
            proj_xr[o] = u - bf * invz;
            proj_y[o] = v;
            scale_level[o] = nScale;
            view_cos[o] = viewCos;
        }
    }
    return 0;
}

extern "C" int orc_frame_lines_in_frustum_batch(int n_frames, const float* tcw, int m, const double* start3d, const double* end3d, uint8_t* in_view) {
    for (int f = 0; f < n_frames; f++) {
        const float* T = tcw + 12 * (size_t)f;
        const float t[3] = {T[3], T[7], T[11]};
        for (int j = 0; j < m; j++) {
            const float s[3] = {(float)start3d[3 * (size_t)j], (float)start3d[3 * (size_t)j + 1], (float)start3d[3 * (size_t)j + 2]};
            const float e[3] = {(float)end3d[3 * (size_t)j], (float)end3d[3 * (size_t)j + 1], (float)end3d[3 * (size_t)j + 2]};
            float ps[3], pe[3];
            gemm3(T, 4, t, s, ps);
            gemm3(T, 4, t, e, pe);
            in_view[(size_t)f * m + j] = (ps[2] < 0.0f && pe[2] < 0.0f) ? 0 : 1;
        }
    }
    return 0;
}

// Frame::ComputeStereoMatches — src/Frame.cc:888-1062.  left / right = the two extractors after
// operator() on the rectified pair (their mvImagePyramid is read, Frame.cc:895,985,997,1002).
#include <algorithm>
#include <climits>
#include <utility>
#include <vector>
extern "C" int orc_frame_compute_stereo_matches(const orc_orb* left, const orc_orb* right, const pl_keypoint* keysL, const uint8_t* descL, int N,
                                                const pl_keypoint* keysR, const uint8_t* descR, int Nr, float mbf, float mb, float* mvuRight,
                                                float* mvDepth) {
    float sf[32], invsf[32], s2[32], is2[32];
    int perLevel[32], umax[32];
    orc_orb_tables(left, sf, invsf, s2, is2, perLevel, umax);
    struct Lvl { int w, h; std::vector<uint8_t> px; };  // bordered plane, the image is the ROI at (19, 19)
    std::vector<Lvl> pl, pr;
    for (int side = 0; side < 2; side++) {
        const orc_orb* o = side ? right : left;
        std::vector<Lvl>& v = side ? pr : pl;
        for (int l = 0;; l++) {
            int w, h;
            if (orc_orb_level_dims(o, l, &w, &h) != 0) break;
            Lvl L{w, h, std::vector<uint8_t>((size_t)(w + 38) * (h + 38))};
            orc_orb_level_bordered(o, l, L.px.data());
            v.push_back(std::move(L));
        }
    }
    auto at = [](const Lvl& L, int r, int c) -> float { return (float)L.px[(size_t)(r + 19) * (L.w + 38) + (c + 19)]; };
    for (int i = 0; i < N; i++) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; }
    if (pl.empty() || pr.empty()) return 0;
    const int thOrbDist = (100 + 50) / 2;
    const int nRows = pl[0].h;
    std::vector<std::vector<size_t>> vRowIndices(nRows);
    for (int iR = 0; iR < Nr; iR++) {
        const float kpY = keysR[iR].y;
        const float r = 2.0f * sf[keysR[iR].octave];
        const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);  // the reference writes out of bounds otherwise
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < N; iL++) {
        const pl_keypoint& kpL = keysL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        if (!(vL >= 0 && (int)vL < nRows)) continue;
        const std::vector<size_t>& vCandidates = vRowIndices[(size_t)vL];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = 100;
        size_t bestIdxR = 0;
        const uint8_t* dL = descL + 32 * (size_t)iL;
        for (size_t iR : vCandidates) {
            const pl_keypoint& kpR = keysR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = orc_descriptor_distance(dL, descR + 32 * iR);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = keysR[bestIdxR].x;
            const float scaleFactor = invsf[kpL.octave];
            const float scaleduL = std::round(kpL.x * scaleFactor);
            const float scaledvL = std::round(kpL.y * scaleFactor);
            const float scaleduR0 = std::round(uR0 * scaleFactor);
            const int w = 5;
            const Lvl& IL = pl[kpL.octave];
            const Lvl& IRl = pr[kpL.octave];
            const int r0 = (int)(scaledvL - w), c0 = (int)(scaleduL - w);
            const float cL = at(IL, r0 + w, c0 + w);
            int bestDistS = INT_MAX, bestincR = 0;
            const int L = 5;
            float vDists[2 * 5 + 1];
            const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= IRl.w) continue;
            for (int incR = -L; incR <= +L; incR++) {
                const int cr0 = (int)(scaleduR0 + incR - w);
                const float cR = at(IRl, r0 + w, cr0 + w);
                double acc = 0;  // cv::norm(IL, IR, NORM_L1) of CV_32F: double accumulation of |a - b|
                for (int y = 0; y < 2 * w + 1; y++)
                    for (int x = 0; x < 2 * w + 1; x++) acc += std::fabs((at(IL, r0 + y, c0 + x) - cL) - (at(IRl, r0 + y, cr0 + x) - cR));
                const float dist = (float)acc;
                if (dist < bestDistS) { bestDistS = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = sf[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                mvDepth[iL] = mbf / disparity;
                mvuRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
            }
        }
    }
    if (vDistIdx.empty()) return 0;  // the reference indexes an empty vector here
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        mvuRight[vDistIdx[i].second] = -1;
        mvDepth[vDistIdx[i].second] = -1;
    }
    return 0;
}

// match_oracle.cpp — CPU ORACLE for the descriptor searches.  TEST INFRASTRUCTURE ONLY (see orb_oracle.cpp header).
//
// Restates, sequentially and literally:
//   ORBmatcher::DescriptorDistance (reference src/ORBmatcher.cc:2083-2103) == LineMatcher::DescriptorDistance
//     (src/LineMatcher.cpp:20-39): SWAR popcount over eight 32-bit words;
//   cv::BFMatcher(NORM_HAMMING, crossCheck=false).knnMatch(q, t, 2) as called at src/LineMatcher.cpp:496-503 and
//     :1179-1185 — pinned to python cv2 4.13 (ties resolve to the lowest train index) by tests/test_oracle_match.py
//     and tests/golden/knn_*.npz.
//
// PINNED TO REFERENCE CODE (oracle/_ref/libplref.so = the reference's own function definitions, cut out of /root/reference at build
// time and compiled against stand-in Frame / KeyFrame / MapPoint / MapLine classes; tests/test_oracle_ref.py): DescriptorDistance,
// ComputeThreeMaxima, every ORBmatcher search C2 - C7 (SearchByProjection x4, SearchByBoW x2), SearchForInitialization, both Fuse
// overloads, SearchBySim3, MapPoint::ComputeDistinctiveDescriptors, and the LineMatcher searches D3 / D4 / D5 with the LineMatching predicate,
// LiangBarsky and UpdateKeyLineData.  Still restated only: SearchForTriangulation, the kNN-ratio / MAD
// rules of D6 (their kNN itself is pinned to cv2).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "oracle.h"

static inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    const int32_t* pa = reinterpret_cast<const int32_t*>(a);
    const int32_t* pb = reinterpret_cast<const int32_t*>(b);
    int dist = 0;
    for (int i = 0; i < 8; i++, pa++, pb++) {
        unsigned int v = *pa ^ *pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

extern "C" int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }
namespace { void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3); }
// lab hook: ComputeThreeMaxima on bins of the given sizes (tests/test_oracle_ref.py compares it with the reference's own)
extern "C" void orc_compute_three_maxima(const int* counts, int L, int* ind) {
    std::vector<std::vector<int>> histo(L);
    for (int i = 0; i < L; i++) histo[i].assign(counts[i], 0);
    ind[0] = ind[1] = ind[2] = -1;
    three_maxima(histo.data(), L, ind[0], ind[1], ind[2]);
}

extern "C" void orc_hamming_pairs(const uint8_t* a, const uint8_t* b, int n, int* dist) {
    for (int i = 0; i < n; i++) dist[i] = descriptor_distance(a + 32 * (size_t)i, b + 32 * (size_t)i);
}

static void knn2_range(const uint8_t* q, int q0, int q1, const uint8_t* t, int nt, int* idx, int* dist) {
    for (int i = q0; i < q1; i++) {
        int b1 = -1, d1 = 1 << 30, b2 = -1, d2 = 1 << 30;
        for (int j = 0; j < nt; j++) {
            int d = descriptor_distance(q + 32 * (size_t)i, t + 32 * (size_t)j);
            if (d < d1) { b2 = b1; d2 = d1; b1 = j; d1 = d; }
            else if (d < d2) { b2 = j; d2 = d; }
        }
        idx[2 * i] = b1; dist[2 * i] = b1 < 0 ? -1 : d1;
        idx[2 * i + 1] = b2; dist[2 * i + 1] = b2 < 0 ? -1 : d2;
    }
}

// threads <= 1: scalar single thread.  The reference's BFMatcher call is single-threaded per call site; the
// multi-thread mode only exists for the "all host cores" CPU baseline of bench.py.
extern "C" void orc_hamming_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx, int* dist, int threads) {
    if (threads <= 1) { knn2_range(q, 0, nq, t, nt, idx, dist); return; }
    std::vector<std::thread> th;
    int per = (nq + threads - 1) / threads;
    for (int k = 0; k < threads; k++) {
        int a = k * per, b = std::min(nq, a + per);
        if (a >= b) break;
        th.emplace_back(knn2_range, q, a, b, t, nt, idx, dist);
    }
    for (auto& x : th) x.join();
}

extern "C" void orc_hamming_candidates(const uint8_t* q, int nq, const uint8_t* t, const int* off, const int* cidx, int* dist) {
    for (int i = 0; i < nq; i++)
        for (int k = off[i]; k < off[i + 1]; k++) dist[k] = descriptor_distance(q + 32 * (size_t)i, t + 32 * (size_t)cidx[k]);
}

// =================================================================================================================
// Projection searches (ORBmatcher) and line matching (LineMatcher) over the POD views of include/plslam_c.h
// =================================================================================================================
#include <climits>
#include <cmath>
#include <utility>

#include "../include/plslam_c.h"

namespace {
const int FRAME_GRID_ROWS = 48, FRAME_GRID_COLS = 64;  // include/Frame.h
const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;  // src/ORBmatcher.cc:49-51

// Frame::AssignFeaturesToGrid (Frame.cc:265-287) + PosInGrid (:527-538)
struct Grid {
    std::vector<int> cell[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    float invW, invH;
    explicit Grid(const pl_frame_view& F) {
        invW = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(F.max_x - F.min_x);
        invH = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(F.max_y - F.min_y);
        for (int i = 0; i < F.n; i++) {
            const pl_keypoint& kp = F.keys_un[i];
            int posX = (int)std::round((kp.x - F.min_x) * invW);
            int posY = (int)std::round((kp.y - F.min_y) * invH);
            if (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) continue;
            cell[posX][posY].push_back(i);
        }
    }
    // Frame::GetFeaturesInArea — Frame.cc:432-485
    void in_area(const pl_frame_view& F, float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
        out.clear();
        const int nMinCellX = std::max(0, (int)std::floor((x - F.min_x - r) * invW));
        if (nMinCellX >= FRAME_GRID_COLS) return;
        const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)std::ceil((x - F.min_x + r) * invW));
        if (nMaxCellX < 0) return;
        const int nMinCellY = std::max(0, (int)std::floor((y - F.min_y - r) * invH));
        if (nMinCellY >= FRAME_GRID_ROWS) return;
        const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)std::ceil((y - F.min_y + r) * invH));
        if (nMaxCellY < 0) return;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
                for (int idx : cell[ix][iy]) {
                    const pl_keypoint& kp = F.keys_un[idx];
                    if (bCheckLevels) {
                        if (kp.octave < minLevel) continue;
                        if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                    }
                    const float distx = kp.x - x, disty = kp.y - y;
                    if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(idx);
                }
    }
};

// ORBmatcher::ComputeThreeMaxima — ORBmatcher.cc:2035-2077
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// cv::Mat (CV_32F) expression R*x + t as OpenCV evaluates it: one gemm with double accumulation, rounded once
inline void mat_rx_plus_t(const float* T /*3x4 row-major*/, const float* X, float* out) {
    for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double)T[4 * r + k] * (double)X[k];
        out[r] = (float)(s * 1.0 + (double)T[4 * r + 3] * 1.0);
    }
}
}  // namespace

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) — ORBmatcher.cc:72-183
extern "C" int orc_orb_search_local_points(const pl_frame_view* Fp, const pl_mappoint_view* M, float th, float nn_ratio,
                                           int* match_of_feature, int* n_matches) {
    const pl_frame_view& F = *Fp;
    Grid grid(F);
    std::vector<uint8_t> claimed(F.n);
    for (int i = 0; i < F.n; i++) { claimed[i] = F.claimed ? (F.claimed[i] != 0) : 0; match_of_feature[i] = -1; }
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<int> vIndices;
    for (int iMP = 0; iMP < M->n; iMP++) {
        if (!M->track_in_view[iMP]) continue;
        const int nPredictedLevel = M->scale_level[iMP];
        float r = M->view_cos[iMP] > 0.998 ? 2.5f : 4.0f;  // RadiusByViewingCos :186-193
        if (bFactor) r *= th;
        grid.in_area(F, M->proj_x[iMP], M->proj_y[iMP], r * F.scale_factors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, vIndices);
        if (vIndices.empty()) continue;
        const uint8_t* d = M->desc + 32 * (size_t)iMP;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : vIndices) {
            if (claimed[idx]) continue;
            if (F.u_right[idx] > 0) {
                const float er = std::fabs(M->proj_xr[iMP] - F.u_right[idx]);
                if (er > r * F.scale_factors[nPredictedLevel]) continue;
            }
            const int dist = descriptor_distance(d, F.desc + 32 * (size_t)idx);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F.keys_un[idx].octave; bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = F.keys_un[idx].octave; bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nn_ratio * bestDist2) continue;
            match_of_feature[bestIdx] = iMP;
            claimed[bestIdx] = M->has_observations ? (M->has_observations[iMP] != 0) : 1;
            nmatches++;
        }
    }
    *n_matches = nmatches;
    return 0;
}

// ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) — ORBmatcher.cc:1710-1879
extern "C" int orc_orb_search_last_frame(const pl_frame_view* Cp, const pl_lastframe_view* L, float th, int mono, int check_orientation,
                                         int* match_of_feature, int* n_matches) {
    const pl_frame_view& C = *Cp;
    Grid grid(C);
    std::vector<uint8_t> claimed(C.n);
    for (int i = 0; i < C.n; i++) { claimed[i] = C.claimed ? (C.claimed[i] != 0) : 0; match_of_feature[i] = -1; }
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = HISTO_LENGTH / 360.0f;
    // twc = -Rcw^T * tcw ; tlc = Rlw*twc + tlw   (cv::Mat float expressions: gemm with double accumulation)
    float twc[3], tlc[3];
    for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double)C.tcw[4 * k + r] * (double)C.tcw[4 * k + 3];
        twc[r] = (float)(s * -1.0);
    }
    mat_rx_plus_t(L->tcw, twc, tlc);
    const bool bForward = tlc[2] > C.b && !mono;
    const bool bBackward = -tlc[2] > C.b && !mono;
    std::vector<int> vIndices2;
    for (int i = 0; i < L->n; i++) {
        if (!L->valid[i]) continue;
        float x3Dc[3];
        mat_rx_plus_t(C.tcw, L->world_pos + 3 * (size_t)i, x3Dc);
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = (float)(1.0 / x3Dc[2]);
        if (invzc < 0) continue;
        float u = C.fx * xc * invzc + C.cx;
        float v = C.fy * yc * invzc + C.cy;
        if (u < C.min_x || u > C.max_x) continue;
        if (v < C.min_y || v > C.max_y) continue;
        const int nLastOctave = L->octave[i];
        const float radius = th * C.scale_factors[nLastOctave];
        if (bForward) grid.in_area(C, u, v, radius, nLastOctave, -1, vIndices2);
        else if (bBackward) grid.in_area(C, u, v, radius, 0, nLastOctave, vIndices2);
        else grid.in_area(C, u, v, radius, nLastOctave - 1, nLastOctave + 1, vIndices2);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = L->desc + 32 * (size_t)i;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (claimed[i2]) continue;
            if (C.u_right[i2] > 0) {
                const float ur = u - C.bf * invzc;
                const float er = std::fabs(ur - C.u_right[i2]);
                if (er > radius) continue;
            }
            const int dist = descriptor_distance(dMP, C.desc + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            match_of_feature[bestIdx2] = i;
            claimed[bestIdx2] = L->has_observations ? (L->has_observations[i] != 0) : 1;
            nmatches++;
            if (check_orientation) {
                float rot = L->angle[i] - C.keys_un[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    match_of_feature[rotHist[i][j]] = -1;
                    nmatches--;
                }
    }
    *n_matches = nmatches;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// LineMatcher
// ---------------------------------------------------------------------------------------------------------------
namespace {
// LineMatcher::LiangBarsky — LineMatcher.cpp:1389-1460 (bugs and all: round() of the deltas, the horizontal-line test)
bool liang_barsky(const double line[4], double out[4], const float bounds[4]) {
    const double sx = line[0], sy = line[1], ex = line[2], ey = line[3];
    double p[4], q[4];
    p[0] = sx - ex; p[1] = ex - sx; p[2] = sy - ey; p[3] = ey - sy;
    q[0] = sx - bounds[0]; q[1] = bounds[2] - sx; q[2] = sy - bounds[1]; q[3] = bounds[3] - sy;
    if (p[0] == 0) { if (q[0] <= 0 || q[2] <= 0) return false; }
    if (p[2] == 0) { if (q[2] >= 0 || q[3] >= 0) return false; }
    double u[4];
    for (int i = 0; i < 4; i++) u[i] = q[i] / p[i];
    double u_min = 0, u_max = 1;
    for (int i = 0; i < 4; i++) {
        if (p[i] < 0) { if (u_min < u[i]) u_min = u[i]; }
        else { if (u_max > u[i]) u_max = u[i]; }
    }
    if (u_max >= u_min) {
        out[0] = sx + std::round(u_min * (ex - sx));
        out[1] = sy + std::round(u_min * (ey - sy));
        out[2] = sx + std::round(u_max * (ex - sx));
        out[3] = sy + std::round(u_max * (ey - sy));
        return true;
    }
    return false;
}

// cv::LineIterator(img, Point2f a, Point2f b).count (8-connected; points are cvRound'ed; the segment is clipped to the
// image first — cv::clipLine; count 0 when it lies completely outside)
int clip_code(long x, long y, long right, long bottom) { return (x < 0) + (x > right) * 2 + (y < 0) * 4 + (y > bottom) * 8; }
int line_iterator_count(float fx0, float fy0, float fx1, float fy1, int cols, int rows) {
    long x1 = lrintf(fx0), y1 = lrintf(fy0), x2 = lrintf(fx1), y2 = lrintf(fy1);
    const long right = cols - 1, bottom = rows - 1;
    if (cols <= 0 || rows <= 0) return 0;
    int c1 = clip_code(x1, y1, right, bottom), c2 = clip_code(x2, y2, right, bottom);
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {  // cv::clipLine (imgproc/src/drawing.cpp), 64-bit arithmetic
        long a;
        if (c1 & 12) { a = c1 < 8 ? 0 : bottom; x1 += (long)((double)(a - y1) * (x2 - x1) / (y2 - y1)); y1 = a; c1 = (x1 < 0) + (x1 > right) * 2; }
        if (c2 & 12) { a = c2 < 8 ? 0 : bottom; x2 += (long)((double)(a - y2) * (x2 - x1) / (y2 - y1)); y2 = a; c2 = (x2 < 0) + (x2 > right) * 2; }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) { a = c1 == 1 ? 0 : right; y1 += (long)((double)(a - x1) * (y2 - y1) / (x2 - x1)); x1 = a; c1 = 0; }
            if (c2) { a = c2 == 1 ? 0 : right; y2 += (long)((double)(a - x2) * (y2 - y1) / (x2 - x1)); x2 = a; c2 = 0; }
        }
    }
    if ((c1 | c2) != 0) return 0;
    const long dx = std::labs(x2 - x1), dy = std::labs(y2 - y1);
    return (int)std::max(dx, dy) + 1;
}

// LineMatcher::UpdateKeyLineData — LineMatcher.cpp:1601-1624
void update_keyline(const double nl[4], pl_keyline& k, int cols, int rows) {
    k.sx = (float)nl[0]; k.sy = (float)nl[1]; k.ex = (float)nl[2]; k.ey = (float)nl[3];
    k.sx_oct = (float)nl[0]; k.sy_oct = (float)nl[1]; k.ex_oct = (float)nl[2]; k.ey_oct = (float)nl[3];
    k.pt_x = (k.ex + k.sx) / 2; k.pt_y = (k.ey + k.sy) / 2;
    k.length = float(std::sqrt(std::pow(k.sx - k.ex, 2) + std::pow(k.sy - k.ey, 2)));
    k.num_pixels = line_iterator_count(k.sx, k.sy, k.ex, k.ey, cols, rows);
    k.angle = atan2f((k.ey - k.sy), (k.ex - k.sx));  // std::atan2(float, float) under `using namespace std`
    k.size = (k.ex - k.sx) * (k.ey - k.sy);
    k.response = k.length / std::max(cols, rows);
}

// LineMatcher::LineOverLap — LineMatcher.cpp:1508-1559
bool line_overlap(const pl_keyline& a, const pl_keyline& b, double threshold) {
    double d1_x = std::fabs(a.sx - a.ex), d2_x = std::fabs(b.sx - b.ex);
    double min_x = std::min(std::min(a.sx, a.ex), std::min(b.sx, b.ex));
    double max_x = std::max(std::max(a.sx, a.ex), std::max(b.sx, b.ex));
    double d1_y = std::fabs(a.sy - a.ey), d2_y = std::fabs(b.sy - b.ey);
    double min_y = std::min(std::min(a.sy, a.ey), std::min(b.sy, b.ey));
    double max_y = std::max(std::max(a.sy, a.ey), std::max(b.sy, b.ey));
    if (d1_x == 0 || d2_x == 0) { if ((d1_y + d2_y - max_y + min_y) / std::min(d1_y, d2_y) >= threshold) return true; }
    if (d1_y == 0 || d2_y == 0) { if ((d1_x + d2_x - max_x + min_x) / std::min(d1_x, d2_x) >= threshold) return true; }
    if ((d1_x + d2_x - max_x + min_x) / std::min(d1_x, d2_x) >= threshold) {
        if (d1_y + d2_y + min_y >= max_y) return true;
        else if (max_y - min_y - d1_y - d2_y < 0.3 * std::min(d1_y, d2_y)) return true;
    } else if ((d1_x + d2_x - max_x + min_x) / std::min(d1_x, d2_x) < threshold && (max_x - min_x - d1_x - d2_x) < 0.3 * std::min(d1_x, d2_x)) {
        if ((d1_y + d2_y - max_y + min_y) / std::min(d1_y, d2_y) >= threshold) return true;
    }
    return false;
}

// LineMatcher::ReprojectionError — LineMatcher.cpp:1579-1596
double reprojection_error(const pl_keyline& l1, const pl_keyline& l2) {
    const double a[3] = {l1.sx, l1.sy, 1}, b[3] = {l1.ex, l1.ey, 1};
    const double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    const double den = std::sqrt(std::pow(c[0], 2) + std::pow(c[1], 2));
    const double ds = (l2.sx * c[0] + l2.sy * c[1] + 1.0 * c[2]) / den;
    const double de = (l2.ex * c[0] + l2.ey * c[1] + 1.0 * c[2]) / den;
    return std::sqrt(ds * ds + de * de);
}

// LineMatcher::LineMatching — LineMatcher.cpp:1463-1504; thresholds LineMatcher.h:94-98
bool line_matching(const pl_keyline& kl1, const pl_keyline& kl2, const uint8_t* d1, const uint8_t* d2, const double off[5]) {
    const double angle_threshold = 15.0 * M_PI / 180.0, length_threshold = 0.45, overlap_threshold = 0.5, desc_dist_threshold = 45,
                 reproj_error_threshold = 45;
    if (descriptor_distance(d1, d2) > desc_dist_threshold + off[3]) return false;
    if (std::fabs(kl1.angle - kl2.angle) > angle_threshold + off[0] * M_PI / 180.0) return false;
    if (std::min(kl1.length, kl2.length) / std::max(kl1.length, kl2.length) < length_threshold + off[1]) return false;
    if (!line_overlap(kl1, kl2, overlap_threshold + off[2])) return false;
    if (reprojection_error(kl1, kl2) > reproj_error_threshold) return false;
    return true;
}
}  // namespace

extern "C" int orc_line_iterator_count(float x0, float y0, float x1, float y1, int cols, int rows) {
    return line_iterator_count(x0, y0, x1, y1, cols, rows);
}

// front half of LineMatcher::SearchByProjection (all three overloads): LineMatcher.cpp:96-212 / :558-… / :790-…
extern "C" int orc_line_project(const double* start3d, const double* end3d, const pl_keyline* src_kl, const uint8_t* valid, int n,
                                const float tcw[12], float fx, float fy, float cx, float cy, float min_x, float min_y, float max_x,
                                float max_y, int img_cols, int img_rows, pl_keyline* out_kl, int* out_index, int* n_out) {
    double T[12];
    for (int i = 0; i < 12; i++) T[i] = tcw[i];  // Eigen::Isometry3d filled from the float Tcw (:77-92)
    const float bounds[4] = {min_x, min_y, max_x, max_y};
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (!valid[i]) continue;
        const double* Xs = start3d + 3 * (size_t)i;
        const double* Xe = end3d + 3 * (size_t)i;
        double cs[3], ce[3];
        for (int r = 0; r < 3; r++) {  // Isometry3d * Vector3d: linear part times vector, plus translation
            cs[r] = T[4 * r] * Xs[0] + T[4 * r + 1] * Xs[1] + T[4 * r + 2] * Xs[2] + T[4 * r + 3];
            ce[r] = T[4 * r] * Xe[0] + T[4 * r + 1] * Xe[1] + T[4 * r + 2] * Xe[2] + T[4 * r + 3];
        }
        if (cs[2] < 0 && ce[2] < 0) continue;
        double proj[4], clipped[4];
        bool have = false, ok = false;
        if (cs[2] < 0.0 || ce[2] < 0.0) {
            const double lambda = -1.0 * cs[2] / (cs[2] - ce[2]);
            const double x_c_cross = cs[0] + lambda * (cs[0] - ce[0]);
            const double y_c_cross = cs[1] + lambda * (cs[1] - ce[1]);
            if (cs[2] < 0.0) {
                const float u_end = (float)(fx * ce[0] / ce[2] + cx), v_end = (float)(fy * ce[1] / ce[2] + cy);
                proj[0] = x_c_cross; proj[1] = y_c_cross; proj[2] = u_end; proj[3] = v_end;
                have = true;
            } else if (ce[2] < 0.0) {
                const float u_start = (float)(fx * cs[0] / cs[2] + cx), v_start = (float)(fy * cs[1] / cs[2] + cy);
                proj[0] = u_start; proj[1] = v_start; proj[2] = x_c_cross; proj[3] = y_c_cross;
                have = true;
            }
            if (have) {
                ok = liang_barsky(proj, clipped, bounds);
                if (ok) {
                    pl_keyline k = src_kl[i];
                    update_keyline(clipped, k, img_cols, img_rows);
                    out_kl[m] = k; out_index[m] = i; m++;
                }
            }
        }
        if (cs[2] > 0.0 && ce[2] > 0.0) {
            const float u_start = (float)(fx * cs[0] / cs[2] + cx), v_start = (float)(fy * cs[1] / cs[2] + cy);
            const float u_end = (float)(fx * ce[0] / ce[2] + cx), v_end = (float)(fy * ce[1] / ce[2] + cy);
            proj[0] = u_start; proj[1] = v_start; proj[2] = u_end; proj[3] = v_end;
            if (liang_barsky(proj, clipped, bounds)) {
                pl_keyline k = src_kl[i];
                update_keyline(clipped, k, img_cols, img_rows);
                out_kl[m] = k; out_index[m] = i; m++;
            }
        }
    }
    *n_out = m;
    return 0;
}

// back half: all-pairs LineMatching with "last matching i wins, every hit counts" and the relaxed retry (:215-261)
extern "C" int orc_line_match_pairs(const pl_keyline* proj, const uint8_t* proj_desc, int n_proj, const pl_keyline* cur, const uint8_t* cur_desc,
                                    const uint8_t* cur_claimed, int n_cur, int* match_of_line, int* n_matches, int* used_relaxed) {
    const double zero[5] = {0, 0, 0, 0, 0}, relaxed[5] = {10.0, -0.1, -0.1, 5, 10};
    int cnt = 0;
    for (int j = 0; j < n_cur; j++) match_of_line[j] = -1;
    for (int j = 0; j < n_cur; j++) {
        if (cur_claimed && cur_claimed[j]) continue;
        for (int i = 0; i < n_proj; i++)
            if (line_matching(proj[i], cur[j], proj_desc + 32 * (size_t)i, cur_desc + 32 * (size_t)j, zero)) { match_of_line[j] = i; cnt++; }
    }
    *used_relaxed = 0;
    if (cnt * 1.0 / n_cur < 0.2) {
        *used_relaxed = 1;
        cnt = 0;
        for (int j = 0; j < n_cur; j++) match_of_line[j] = -1;
        for (int j = 0; j < n_cur; j++)
            for (int i = 0; i < n_proj; i++)
                if (line_matching(proj[i], cur[j], proj_desc + 32 * (size_t)i, cur_desc + 32 * (size_t)j, relaxed)) { match_of_line[j] = i; cnt++; }
    }
    *n_matches = cnt;
    return 0;
}

// =================================================================================================================
// C4 / C5 / C6 / C7 / D6
// =================================================================================================================
namespace {
// MapPoint::PredictScale — src/MapPoint.cc:397-431 (log(float) resolves to logf: `using namespace std` reaches the
// file through Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:36)
inline int predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels) {
    const float ratio = max_distance / current_dist;
    int nScale = (int)std::ceil(std::log(ratio) / log_scale_factor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= n_levels) nScale = n_levels - 1;
    return nScale;
}
// cv::norm(PO) of a 3x1 CV_32F matrix: double accumulation of exact products, sqrt, returned as double
inline float norm3(const float* v) {
    double s = 0;
    for (int k = 0; k < 3; k++) s += (double)v[k] * v[k];
    return (float)std::sqrt(s);
}
}  // namespace

// ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist) — ORBmatcher.cc:1891-2024
extern "C" int orc_orb_search_keyframe_points(const pl_frame_view* Cp, const pl_posepoint_view* P, const float* ow, float log_scale_factor,
                                              float th, int orb_dist, int check_orientation, int* match_of_feature, int* n_matches) {
    const pl_frame_view& C = *Cp;
    Grid grid(C);
    std::vector<uint8_t> claimed(C.n);
    for (int i = 0; i < C.n; i++) { claimed[i] = C.claimed ? (C.claimed[i] != 0) : 0; match_of_feature[i] = -1; }
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = HISTO_LENGTH / 360.0f;
    std::vector<int> vIndices2;
    for (int i = 0; i < P->n; i++) {
        if (!P->valid[i]) continue;
        const float* x3Dw = P->world_pos + 3 * (size_t)i;
        float x3Dc[3];
        mat_rx_plus_t(C.tcw, x3Dw, x3Dc);
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = (float)(1.0 / x3Dc[2]);
        const float u = C.fx * xc * invzc + C.cx;
        const float v = C.fy * yc * invzc + C.cy;
        if (u < C.min_x || u > C.max_x) continue;
        if (v < C.min_y || v > C.max_y) continue;
        const float PO[3] = {x3Dw[0] - ow[0], x3Dw[1] - ow[1], x3Dw[2] - ow[2]};
        const float dist3D = norm3(PO);
        const float maxDistance = P->max_dist_inv[i], minDistance = P->min_dist_inv[i];
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const int nPredictedLevel = predict_scale(P->max_dist[i], dist3D, log_scale_factor, C.n_levels);
        const float radius = th * C.scale_factors[nPredictedLevel];
        grid.in_area(C, u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1, vIndices2);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = P->desc + 32 * (size_t)i;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (claimed[i2]) continue;
            const int dist = descriptor_distance(dMP, C.desc + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= orb_dist) {
            match_of_feature[bestIdx2] = i;
            claimed[bestIdx2] = 1;
            nmatches++;
            if (check_orientation) {
                float rot = P->angle[i] - C.keys_un[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    match_of_feature[rotHist[i][j]] = -1;
                    nmatches--;
                }
    }
    *n_matches = nmatches;
    return 0;
}

// ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th) — ORBmatcher.cc:423-554.
// K->tcw = Rcw | tcw with the scale already divided out; KeyFrame::GetFeaturesInArea (KeyFrame.cc:642-681) has no
// level filter, the level gate of :527-531 is applied with it here (same set, same order).
extern "C" int orc_orb_search_sim3_points(const pl_frame_view* Kp, const pl_posepoint_view* P, const float* ow, float log_scale_factor, int th,
                                          int* match_of_feature, int* n_matches) {
    const pl_frame_view& K = *Kp;
    Grid grid(K);
    std::vector<uint8_t> matched(K.n);
    for (int i = 0; i < K.n; i++) { matched[i] = K.claimed ? (K.claimed[i] != 0) : 0; match_of_feature[i] = -1; }
    int nmatches = 0;
    std::vector<int> vIndices;
    for (int iMP = 0; iMP < P->n; iMP++) {
        if (!P->valid[iMP]) continue;
        const float* p3Dw = P->world_pos + 3 * (size_t)iMP;
        float p3Dc[3];
        mat_rx_plus_t(K.tcw, p3Dw, p3Dc);
        if (p3Dc[2] < 0.0) continue;
        const float invz = 1 / p3Dc[2];
        const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
        const float u = K.fx * x + K.cx, v = K.fy * y + K.cy;
        if (!(u >= K.min_x && u < K.max_x && v >= K.min_y && v < K.max_y)) continue;  // KeyFrame::IsInImage (KeyFrame.cc:683-686)
        const float maxDistance = P->max_dist_inv[iMP], minDistance = P->min_dist_inv[iMP];
        const float PO[3] = {p3Dw[0] - ow[0], p3Dw[1] - ow[1], p3Dw[2] - ow[2]};
        const float dist = norm3(PO);
        if (dist < minDistance || dist > maxDistance) continue;
        const float* Pn = P->normal + 3 * (size_t)iMP;
        double dot = 0;  // cv::Mat::dot of CV_32F: double accumulation
        for (int k = 0; k < 3; k++) dot += (double)PO[k] * Pn[k];
        if (dot < 0.5 * dist) continue;
        const int nPredictedLevel = predict_scale(P->max_dist[iMP], dist, log_scale_factor, K.n_levels);
        const float radius = th * K.scale_factors[nPredictedLevel];
        grid.in_area(K, u, v, radius, -1, -1, vIndices);
        if (vIndices.empty()) continue;
        const uint8_t* dMP = P->desc + 32 * (size_t)iMP;
        int bestDist = 256, bestIdx = -1;
        for (int idx : vIndices) {
            if (matched[idx]) continue;
            const int kpLevel = K.keys_un[idx].octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const int d = descriptor_distance(dMP, K.desc + 32 * (size_t)idx);
            if (d < bestDist) { bestDist = d; bestIdx = idx; }
        }
        if (bestDist <= TH_LOW) {
            match_of_feature[bestIdx] = iMP;
            matched[bestIdx] = 1;
            nmatches++;
        }
    }
    *n_matches = nmatches;
    return 0;
}

// ORBmatcher::SearchByBoW — mode 0: (KeyFrame*, Frame&, vpMapPointMatches) ORBmatcher.cc:247-410;
//                           mode 1: (KeyFrame*, KeyFrame*, vpMatches12) :729-872
extern "C" int orc_orb_search_bow(const pl_bow_view* A, const pl_bow_view* B, int mode, float nn_ratio, int check_orientation, int* match_out,
                                  int* n_matches) {
    const int n_out = mode == 0 ? B->n : A->n;
    for (int i = 0; i < n_out; i++) match_out[i] = -1;
    std::vector<uint8_t> matchedB(B->n, 0);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = HISTO_LENGTH / 360.0f;
    int ka = 0, kb = 0;
    while (ka < A->n_nodes && kb < B->n_nodes) {
        if (A->node_id[ka] == B->node_id[kb]) {
            for (int pa = A->node_off[ka]; pa < A->node_off[ka + 1]; pa++) {
                const int idxA = (int)A->feat_idx[pa];
                if (A->valid && !A->valid[idxA]) continue;
                const uint8_t* dA = A->desc + 32 * (size_t)idxA;
                int bestDist1 = 256, bestIdxB = -1, bestDist2 = 256;
                for (int pb = B->node_off[kb]; pb < B->node_off[kb + 1]; pb++) {
                    const int idxB = (int)B->feat_idx[pb];
                    if (matchedB[idxB]) continue;
                    if (mode == 1 && B->valid && !B->valid[idxB]) continue;
                    const int dist = descriptor_distance(dA, B->desc + 32 * (size_t)idxB);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxB = idxB; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                const bool pass = mode == 0 ? bestDist1 <= TH_LOW : bestDist1 < TH_LOW;
                if (pass && static_cast<float>(bestDist1) < nn_ratio * static_cast<float>(bestDist2)) {
                    matchedB[bestIdxB] = 1;
                    if (mode == 0) match_out[bestIdxB] = idxA;
                    else match_out[idxA] = bestIdxB;
                    if (check_orientation) {
                        float rot = A->angle[idxA] - B->angle[bestIdxB];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(mode == 0 ? bestIdxB : idxA);
                    }
                    nmatches++;
                }
            }
            ka++;
            kb++;
        } else if (A->node_id[ka] < B->node_id[kb]) {
            while (ka < A->n_nodes && A->node_id[ka] < B->node_id[kb]) ka++;  // lower_bound
        } else {
            while (kb < B->n_nodes && B->node_id[kb] < A->node_id[ka]) kb++;
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) {
                match_out[rotHist[i][j]] = -1;
                nmatches--;
            }
        }
    }
    *n_matches = nmatches;
    return 0;
}

// LineMatcher::SearchByProjection(Frame&, KeyFrame*, vector<MapLine*>&) — LineMatcher.cpp:489-525
extern "C" int orc_line_match_knn_ratio(const uint8_t* ref_desc, int n_ref, const uint8_t* cur_desc, int n_cur, int* match_of_line, int* n_matches) {
    for (int j = 0; j < n_cur; j++) match_of_line[j] = -1;
    *n_matches = 0;
    if (n_ref <= 0 || n_cur < 2) return 0;
    std::vector<int> idx(2 * (size_t)n_ref), dist(2 * (size_t)n_ref);
    orc_hamming_knn2(ref_desc, n_ref, cur_desc, n_cur, idx.data(), dist.data(), 1);
    int cnt = 0;
    for (int i = 0; i < n_ref; i++) {
        const float best = (float)dist[2 * i], better = (float)dist[2 * i + 1];
        const float distanceRatio = best / better;
        if (distanceRatio < 0.75) {
            match_of_line[idx[2 * i]] = i;
            cnt++;
        }
    }
    *n_matches = cnt;
    return 0;
}

// LineMatcher::SearchForTriangulation — LineMatcher.cpp:1174-1204, KeyFrame::lineDescriptorMAD — KeyFrame.cc:773-797
extern "C" int orc_line_search_for_triangulation(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, int* pairs, int* n_matches,
                                                 double* nn_mad_out, double* nn12_mad_out) {
    *n_matches = 0;
    if (nn_mad_out) *nn_mad_out = 0;
    if (nn12_mad_out) *nn12_mad_out = 0;
    if (n1 <= 0 || n2 < 2) return 0;
    std::vector<int> idx(2 * (size_t)n1), dist(2 * (size_t)n1);
    orc_hamming_knn2(desc1, n1, desc2, n2, idx.data(), dist.data(), 1);
    std::vector<float> d0(n1), d1(n1);
    for (int i = 0; i < n1; i++) { d0[i] = (float)dist[2 * i]; d1[i] = (float)dist[2 * i + 1]; }
    // nn
    std::vector<float> a(d0);
    std::sort(a.begin(), a.end());
    const double nn_dist_median = a[n1 / 2];
    for (int i = 0; i < n1; i++) a[i] = fabsf((float)(d0[i] - nn_dist_median));
    std::sort(a.begin(), a.end());
    const double nn_mad = 1.4826 * a[n1 / 2];
    // nn12
    std::vector<float> e(n1);
    for (int i = 0; i < n1; i++) e[i] = d1[i] - d0[i];
    std::sort(e.begin(), e.end());
    const double nn12_dist_median = e[n1 / 2];
    std::vector<float> b(n1);
    for (int i = 0; i < n1; i++) b[i] = fabsf((float)(d1[i] - d0[i] - nn12_dist_median));
    std::sort(b.begin(), b.end());
    double nn12_mad = 1.4826 * b[n1 / 2];
    if (nn_mad_out) *nn_mad_out = nn_mad;
    if (nn12_mad_out) *nn12_mad_out = nn12_mad;
    const double nn12_dist_th = nn12_mad * 0.1;
    int cnt = 0;
    for (int i = 0; i < n1; i++) {
        const double dist_12 = d1[i] - d0[i];
        if (dist_12 > nn12_dist_th) {
            pairs[2 * cnt] = i;
            pairs[2 * cnt + 1] = idx[2 * i];
            cnt++;
        }
    }
    *n_matches = cnt;
    return 0;
}

// LineMatcher::Fuse, active branch — LineMatcher.cpp:1296-1330
extern "C" int orc_line_fuse_candidates(const uint8_t* ml_desc, const uint8_t* valid, int n, const uint8_t* kf_desc, int n_kf, int* tdx,
                                        int* n_fused) {
    int fused = 0;
    for (int i = 0; i < n; i++) {
        tdx[i] = -1;
        if ((valid && !valid[i]) || n_kf <= 0) continue;
        int best = 257, bi = -1;
        for (int j = 0; j < n_kf; j++) {
            const int d = descriptor_distance(ml_desc + 32 * (size_t)i, kf_desc + 32 * (size_t)j);
            if (d < best) { best = d; bi = j; }
        }
        double min_dist = 100;
        const double dist = (float)best;
        if (dist < min_dist) min_dist = dist;
        if ((float)best < 1.5 * min_dist) {
            tdx[i] = bi;
            fused++;
        }
    }
    *n_fused = fused;
    return 0;
}

// =================================================================================================================
// E rows: the remaining ORBmatcher entry points and ComputeDistinctiveDescriptors
// =================================================================================================================
namespace {
// KeyFrame::IsInImage — KeyFrame.cc:683-686
inline bool kf_in_image(const pl_frame_view& K, float x, float y) { return x >= K.min_x && x < K.max_x && y >= K.min_y && y < K.max_y; }

// the window search shared by Fuse / SearchBySim3: KeyFrame::GetFeaturesInArea(u, v, radius) + the level gate
// [lvl-1, lvl]; chi2 != nullptr adds the reprojection gates of ORBmatcher.cc:1217-1235
inline void best_in_window(const pl_frame_view& K, const Grid& grid, float u, float v, float ur, float radius, int lvl, const uint8_t* dMP,
                           const float* chi2_inv_sigma2, std::vector<int>& scratch, int& bestDist, int& bestIdx) {
    bestDist = 256;  // the reference starts from 256 (:1199) or INT_MAX (:1377, :1520): the same for a <= TH test
    bestIdx = -1;
    grid.in_area(K, u, v, radius, -1, -1, scratch);
    for (int idx : scratch) {
        const pl_keypoint& kp = K.keys_un[idx];
        const int kpLevel = kp.octave;
        if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
        if (chi2_inv_sigma2) {
            if (K.u_right[idx] >= 0) {
                const float ex = u - kp.x, ey = v - kp.y, er = ur - K.u_right[idx];
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * chi2_inv_sigma2[kpLevel] > 7.8) continue;
            } else {
                const float ex = u - kp.x, ey = v - kp.y;
                const float e2 = ex * ex + ey * ey;
                if (e2 * chi2_inv_sigma2[kpLevel] > 5.99) continue;
            }
        }
        const int dist = descriptor_distance(dMP, K.desc + 32 * (size_t)idx);
        if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
    }
}
}  // namespace

// ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) — ORBmatcher.cc:1107-1277 (variant 0)
// ORBmatcher::Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint) — ORBmatcher.cc:1290-1427 (variant 1)
extern "C" int orc_orb_fuse_candidates(const pl_frame_view* Kp, const pl_posepoint_view* P, const float* ow, float log_scale_factor,
                                       const float* inv_level_sigma2, float th, int variant, int* best_idx, int* best_dist, int* n_fused) {
    const pl_frame_view& K = *Kp;
    Grid grid(K);
    std::vector<int> scratch;
    int nFused = 0;
    for (int i = 0; i < P->n; i++) {
        best_idx[i] = -1;
        if (best_dist) best_dist[i] = 256;
        if (!P->valid[i]) continue;
        const float* p3Dw = P->world_pos + 3 * (size_t)i;
        float p3Dc[3];
        mat_rx_plus_t(K.tcw, p3Dw, p3Dc);
        if (p3Dc[2] < 0.0f) continue;
        const float invz = variant == 0 ? 1 / p3Dc[2] : (float)(1.0 / p3Dc[2]);  // :1145 `1/z` (float) vs :1335 `1.0/z` (double, rounded)
        const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
        const float u = K.fx * x + K.cx, v = K.fy * y + K.cy;
        if (!kf_in_image(K, u, v)) continue;
        const float ur = u - K.bf * invz;
        const float maxDistance = P->max_dist_inv[i], minDistance = P->min_dist_inv[i];
        const float PO[3] = {p3Dw[0] - ow[0], p3Dw[1] - ow[1], p3Dw[2] - ow[2]};
        const float dist3D = norm3(PO);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const float* Pn = P->normal + 3 * (size_t)i;
        double dot = 0;
        for (int k = 0; k < 3; k++) dot += (double)PO[k] * Pn[k];
        if (dot < 0.5 * dist3D) continue;
        const int lvl = predict_scale(P->max_dist[i], dist3D, log_scale_factor, K.n_levels);
        const float radius = th * K.scale_factors[lvl];
        int bd, bi;
        best_in_window(K, grid, u, v, ur, radius, lvl, P->desc + 32 * (size_t)i, variant == 0 ? inv_level_sigma2 : nullptr, scratch, bd, bi);
        if (best_dist) best_dist[i] = bd;
        if (bd <= TH_LOW) {
            best_idx[i] = bi;
            nFused++;
        }
    }
    *n_fused = nFused;
    return 0;
}

// ORBmatcher::SearchBySim3 — ORBmatcher.cc:1441-1692
namespace {
void sim3_half(const pl_frame_view& Ka, const pl_frame_view& Kb, const pl_posepoint_view& P, const float* Tba, float log_sf_b, float th,
               const pl_frame_view& Kintr, std::vector<int>& out) {
    // points of key frame a (world -> camera a with Ka.tcw, -> camera b with Tba = sRba | tba), searched in key frame b
    Grid grid(Kb);
    std::vector<int> scratch;
    out.assign(P.n, -1);
    for (int i = 0; i < P.n; i++) {
        if (!P.valid[i]) continue;
        float pa[3], pb[3];
        mat_rx_plus_t(Ka.tcw, P.world_pos + 3 * (size_t)i, pa);
        mat_rx_plus_t(Tba, pa, pb);
        if (pb[2] < 0.0) continue;
        const float invz = (float)(1.0 / pb[2]);
        const float x = pb[0] * invz, y = pb[1] * invz;
        const float u = Kintr.fx * x + Kintr.cx, v = Kintr.fy * y + Kintr.cy;
        if (!kf_in_image(Kb, u, v)) continue;
        const float dist3D = norm3(pb);
        if (dist3D < P.min_dist_inv[i] || dist3D > P.max_dist_inv[i]) continue;
        const int lvl = predict_scale(P.max_dist[i], dist3D, log_sf_b, Kb.n_levels);
        const float radius = th * Kb.scale_factors[lvl];
        int bd, bi;
        best_in_window(Kb, grid, u, v, 0.f, radius, lvl, P.desc + 32 * (size_t)i, nullptr, scratch, bd, bi);
        if (bd <= TH_HIGH) out[i] = bi;
    }
}
}  // namespace
extern "C" int orc_orb_search_by_sim3(const pl_frame_view* K1, const pl_frame_view* K2, const pl_posepoint_view* P1, const pl_posepoint_view* P2,
                                      const float* t21, const float* t12, float log_sf1, float log_sf2, float th, int* match12, int* n_found) {
    std::vector<int> vnMatch1, vnMatch2;
    sim3_half(*K1, *K2, *P1, t21, log_sf2, th, *K1, vnMatch1);
    sim3_half(*K2, *K1, *P2, t12, log_sf1, th, *K1, vnMatch2);
    int nFound = 0;
    for (int i1 = 0; i1 < P1->n; i1++) {
        match12[i1] = -1;
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && idx2 < P2->n) {
            if (vnMatch2[idx2] == i1) { match12[i1] = idx2; nFound++; }
        }
    }
    *n_found = nFound;
    return 0;
}

// ORBmatcher::SearchForInitialization — ORBmatcher.cc:573-717
extern "C" int orc_orb_search_for_initialization(const pl_frame_view* F1p, const pl_frame_view* F2p, float* prev_matched, int window_size,
                                                 float nn_ratio, int check_orientation, int* matches12, int* n_matches) {
    const pl_frame_view &F1 = *F1p, &F2 = *F2p;
    Grid grid(F2);
    int nmatches = 0;
    for (int i = 0; i < F1.n; i++) matches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = HISTO_LENGTH / 360.0f;
    std::vector<int> vMatchedDistance(F2.n, INT_MAX), vnMatches21(F2.n, -1), vIndices2;
    for (int i1 = 0; i1 < F1.n; i1++) {
        const pl_keypoint& kp1 = F1.keys_un[i1];
        const int level1 = kp1.octave;
        if (level1 > 0) continue;
        grid.in_area(F2, prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)window_size, level1, level1, vIndices2);
        if (vIndices2.empty()) continue;
        const uint8_t* d1 = F1.desc + 32 * (size_t)i1;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            const int dist = descriptor_distance(d1, F2.desc + 32 * (size_t)i2);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nn_ratio) {
                if (vnMatches21[bestIdx2] >= 0) {
                    matches12[vnMatches21[bestIdx2]] = -1;
                    nmatches--;
                }
                matches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (check_orientation) {
                    float rot = F1.keys_un[i1].angle - F2.keys_un[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)std::round(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(i1);
                }
            }
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) {
                const int idx1 = rotHist[i][j];
                if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
            }
        }
    }
    for (int i1 = 0; i1 < F1.n; i1++)
        if (matches12[i1] >= 0) {
            prev_matched[2 * i1] = F2.keys_un[matches12[i1]].x;
            prev_matched[2 * i1 + 1] = F2.keys_un[matches12[i1]].y;
        }
    *n_matches = nmatches;
    return 0;
}

// ORBmatcher::SearchForTriangulation — ORBmatcher.cc:884-1095, CheckDistEpipolarLine — :205-232
extern "C" int orc_orb_search_for_triangulation(const pl_triang_view* A, const pl_triang_view* B, const float* F12, const float* cw1,
                                                const float* kf2_tcw, float fx2, float fy2, float cx2, float cy2, const float* scale_factors2,
                                                const float* level_sigma2_2, int n_levels2, int only_stereo, int check_orientation, int* pairs,
                                                int* n_matches) {
    (void)n_levels2;
    float C2[3];
    mat_rx_plus_t(kf2_tcw, cw1, C2);
    const float invz = 1.0f / C2[2];
    const float ex = fx2 * C2[0] * invz + cx2;
    const float ey = fy2 * C2[1] * invz + cy2;
    int nmatches = 0;
    const pl_bow_view &a = A->bow, &b = B->bow;
    std::vector<uint8_t> vbMatched2(b.n, 0);
    std::vector<int> vMatches12(a.n, -1);
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = HISTO_LENGTH / 360.0f;
    int ka = 0, kb = 0;
    while (ka < a.n_nodes && kb < b.n_nodes) {
        if (a.node_id[ka] == b.node_id[kb]) {
            for (int pa = a.node_off[ka]; pa < a.node_off[ka + 1]; pa++) {
                const int idx1 = (int)a.feat_idx[pa];
                if (a.valid && !a.valid[idx1]) continue;  // pMP1 exists
                const bool bStereo1 = A->u_right[idx1] >= 0;
                if (only_stereo && !bStereo1) continue;
                const pl_keypoint& kp1 = A->keys_un[idx1];
                const uint8_t* d1 = a.desc + 32 * (size_t)idx1;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int pb = b.node_off[kb]; pb < b.node_off[kb + 1]; pb++) {
                    const int idx2 = (int)b.feat_idx[pb];
                    if (vbMatched2[idx2] || (b.valid && !b.valid[idx2])) continue;
                    const bool bStereo2 = B->u_right[idx2] >= 0;
                    if (only_stereo && !bStereo2) continue;
                    const int dist = descriptor_distance(d1, b.desc + 32 * (size_t)idx2);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const pl_keypoint& kp2 = B->keys_un[idx2];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - kp2.x, distey = ey - kp2.y;
                        if (distex * distex + distey * distey < 100 * scale_factors2[kp2.octave]) continue;
                    }
                    // CheckDistEpipolarLine
                    const float la = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
                    const float lb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
                    const float lc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
                    const float num = la * kp2.x + lb * kp2.y + lc;
                    const float den = la * la + lb * lb;
                    if (den == 0) continue;
                    const float dsqr = num * num / den;
                    if (dsqr < 3.84 * level_sigma2_2[kp2.octave]) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    vMatches12[idx1] = bestIdx2;
                    vbMatched2[bestIdx2] = 1;
                    nmatches++;
                    if (check_orientation) {
                        float rot = kp1.angle - B->keys_un[bestIdx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                }
            }
            ka++;
            kb++;
        } else if (a.node_id[ka] < b.node_id[kb]) {
            while (ka < a.n_nodes && a.node_id[ka] < b.node_id[kb]) ka++;
        } else {
            while (kb < b.n_nodes && b.node_id[kb] < a.node_id[ka]) kb++;
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) {
                vMatches12[rotHist[i][j]] = -1;
                nmatches--;
            }
        }
    }
    int np = 0;
    for (int i = 0; i < a.n; i++)
        if (vMatches12[i] >= 0) { pairs[2 * np] = i; pairs[2 * np + 1] = vMatches12[i]; np++; }
    *n_matches = nmatches;
    return 0;
}

// MapPoint::ComputeDistinctiveDescriptors — MapPoint.cc:256-321 (== MapLine.cpp:269-330)
extern "C" int orc_distinctive_descriptors(const uint8_t* desc, const int* group_off, int n_groups, int* best_row) {
    for (int g = 0; g < n_groups; g++) {
        const int N = group_off[g + 1] - group_off[g];
        best_row[g] = -1;
        if (N <= 0) continue;
        const uint8_t* d = desc + 32 * (size_t)group_off[g];
        std::vector<float> Distances((size_t)N * N, 0.f);
        for (int i = 0; i < N; i++)
            for (int j = i + 1; j < N; j++) {
                const int distij = descriptor_distance(d + 32 * (size_t)i, d + 32 * (size_t)j);
                Distances[(size_t)i * N + j] = (float)distij;
                Distances[(size_t)j * N + i] = (float)distij;
            }
        int BestMedian = INT_MAX, BestIdx = 0;
        for (int i = 0; i < N; i++) {
            std::vector<int> vDists(Distances.begin() + (size_t)i * N, Distances.begin() + (size_t)(i + 1) * N);
            std::sort(vDists.begin(), vDists.end());
            const int median = vDists[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        best_row[g] = BestIdx;
    }
    return 0;
}

// Frame::UndistortKeyLines — src/Frame.cc:767-845 (end points through cv::undistortPoints, then the same field updates as
// LineMatcher::UpdateKeyLineData) and Frame::AssignFeaturesToGrid — src/Frame.cc:265-287
extern "C" int orc_frame_undistort_keylines(const pl_keyline* kls, int n, float fx, float fy, float cx, float cy, const float* dist_coef, int img_cols,
                                            int img_rows, pl_keyline* out) {
    if (dist_coef[0] == 0.0f) {
        for (int i = 0; i < n; i++) out[i] = kls[i];
        return 0;
    }
    for (int i = 0; i < n; i++) {
        pl_keyline k = kls[i];
        const float in[4] = {k.sx, k.sy, k.ex, k.ey};
        float un[4];
        orc_frame_undistort_points(in, 2, fx, fy, cx, cy, dist_coef, un);
        const double nl[4] = {un[0], un[1], un[2], un[3]};
        update_keyline(nl, k, img_cols, img_rows);
        out[i] = k;
    }
    return 0;
}
extern "C" int orc_frame_assign_features_to_grid(const pl_keypoint* keys_un, int n, const float* bounds, int* cell_start, int* sorted_idx) {
    pl_frame_view F;
    memset(&F, 0, sizeof(F));
    F.n = n;
    F.keys_un = keys_un;
    F.min_x = bounds[0]; F.min_y = bounds[1]; F.max_x = bounds[2]; F.max_y = bounds[3];
    Grid grid(F);
    int k = 0;
    for (int x = 0; x < FRAME_GRID_COLS; x++)
        for (int y = 0; y < FRAME_GRID_ROWS; y++) {
            cell_start[x * FRAME_GRID_ROWS + y] = k;
            for (int idx : grid.cell[x][y]) sorted_idx[k++] = idx;
        }
    cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = k;
    return 0;
}

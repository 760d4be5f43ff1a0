// search_kernels.cu — the projection searches of ORBmatcher and LineMatcher on sm_100a, batched over independent
// search instances (one instance = one reference call, e.g. one frame pair): the CUDA path behind pl_orb_search_* and
// pl_line_* (include/plslam_c.h).
//
// Reference functions replaced:
//   ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)       src/ORBmatcher.cc:72-183      (C2)
//   ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)   src/ORBmatcher.cc:1710-1879   (C3)
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea                src/Frame.cc:265-287,527-538,432-485
//   LineMatcher::SearchByProjection (last frame / key frame / local map)       src/LineMatcher.cpp:72-269,527-721,755-952
//   LineMatcher::LiangBarsky / UpdateKeyLineData / LineMatching / LineOverLap / ReprojectionError   :1389-1624
//
// The reference walks the candidates of every map point sequentially and lets later points see the features claimed by
// earlier ones (ORBmatcher.cc:128-130,177 / :1807-1809,1834).  Here the expensive part — window lookup in the 64x48
// feature grid, the static gates and the 256-bit distances — runs warp-per-point in parallel and keeps the reference's
// candidate order (CSR lists); one warp per instance then replays the claims in point order out of shared memory.
// Every input array of a call travels in ONE packed pinned buffer (one H2D copy), every result in one D2H copy.
#include "match_common.cuh"

namespace pl {

constexpr int kGridCols = 64, kGridRows = 48, kGridCells = kGridCols * kGridRows;  // include/Frame.h
constexpr int kThHigh = 100, kThLow = 50, kHistoLen = 30;                           // ORBmatcher.cc:49-51

struct FrameDev {
    int n;
    const pl_keypoint* keys;
    const uint4* desc;
    const float* u_right;
    const int* claimed;
    float min_x, min_y, max_x, max_y, fx, fy, cx, cy, bf, b;
    float tcw[12];
    float sf[kMaxLevels];
    int n_levels;
    float inv_w, inv_h;  // mfGridElementWidthInv / HeightInv (Frame.cc:184-185)
};
// one search instance: a frame and the points projected into it (Last-frame features for C3, local map points for C2)
struct SearchDev {
    FrameDev F;
    int np;
    const uint8_t* valid;      // C3: pMP && !outlier ; C2: mbTrackInView && !isBad
    const uint4* pdesc;
    const int* level;          // C3: mvKeys[i].octave ; C2: mnTrackScaleLevel
    const uint8_t* has_obs;
    const float* world_pos;    // C3
    const float* angle;        // C3
    const float *proj_x, *proj_y, *proj_xr, *view_cos;  // C2
    int forward, backward;     // C3
    const float *min_inv, *max_inv, *max_raw, *normal;  // C4 / C5: distance invariance, mfMaxDistance, viewing normal (C5)
    float ow[3], log_sf;                                  // C4 / C5: camera centre, mfLogScaleFactor
    int pt_base, feat_base, sort_base, n2;  // offsets into the batch-wide scratch arrays
    // E rows (Fuse / SearchBySim3): second transform applied after F.tcw (sR21 | t21), chi-square gates of Fuse
    float t2[12];
    float inv_sigma2[kMaxLevels];
    int use_t2, use_chi2;
};

// Frame::AssignFeaturesToGrid (Frame.cc:265-287): features sorted by (cell, index); cell = posX*48 + posY with
// PosInGrid's round() (Frame.cc:527-538).  One CTA per instance; bitonic sort of unique keys in shared memory.
__global__ void __launch_bounds__(1024) k_frame_grid(const SearchDev* __restrict__ SD, int* __restrict__ sorted_idx, int* __restrict__ cell_start) {
    extern __shared__ unsigned int s_keys[];
    const SearchDev& S = SD[blockIdx.x];
    const FrameDev& F = S.F;
    const int n2 = S.n2, tid = threadIdx.x;
    int* sidx = sorted_idx + S.sort_base;
    int* cst = cell_start + (size_t)blockIdx.x * (kGridCells + 1);
    for (int i = tid; i < n2; i += blockDim.x) {
        unsigned key = 0xFFFFFFFFu;
        if (i < F.n) {
            const pl_keypoint kp = F.keys[i];
            const int px = (int)roundf(__fmul_rn(__fsub_rn(kp.x, F.min_x), F.inv_w));
            const int py = (int)roundf(__fmul_rn(__fsub_rn(kp.y, F.min_y), F.inv_h));
            if (px >= 0 && px < kGridCols && py >= 0 && py < kGridRows) key = ((unsigned)(px * kGridRows + py) << 16) | (unsigned)i;
        }
        s_keys[i] = key;
    }
    for (int i = tid; i <= kGridCells; i += blockDim.x) cst[i] = 0;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    const unsigned a = s_keys[i], b = s_keys[ixj];
                    if (up ? (a > b) : (a < b)) { s_keys[i] = b; s_keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    // cell_start[c] = first position with cell >= c (invalid keys sort to the end)
    for (int i = tid; i < n2; i += blockDim.x) {
        const unsigned key = s_keys[i];
        const int c = key == 0xFFFFFFFFu ? kGridCells : (int)(key >> 16);
        const int prev = i == 0 ? -1 : (s_keys[i - 1] == 0xFFFFFFFFu ? kGridCells : (int)(s_keys[i - 1] >> 16));
        if (key != 0xFFFFFFFFu) sidx[i] = (int)(key & 0xFFFFu);
        for (int cc = prev + 1; cc <= c; cc++) cst[cc] = i;
    }
    if (tid == 0) {
        const unsigned last = s_keys[n2 - 1];
        const int c = last == 0xFFFFFFFFu ? kGridCells : (int)(last >> 16);
        for (int cc = c + 1; cc <= kGridCells; cc++) cst[cc] = n2;
    }
}

// Frame::GetFeaturesInArea (Frame.cc:432-485) + static gates (+ distances when `out` is given) for one point; one warp.
// Candidates are visited in the reference's order; returns their number.  out[k] = idx | dist << 16.
__device__ int gather_candidates(const FrameDev& F, const int* __restrict__ sorted_idx, const int* __restrict__ cell_start, float x,
                                 float y, float r, int minLevel, int maxLevel, float ur, float gate_r, uint4 q0, uint4 q1,
                                 unsigned int* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, F.min_x), r), F.inv_w)));
    if (nMinCellX >= kGridCols) return 0;
    const int nMaxCellX = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, F.min_x), r), F.inv_w)));
    if (nMaxCellX < 0) return 0;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, F.min_y), r), F.inv_h)));
    if (nMinCellY >= kGridRows) return 0;
    const int nMaxCellY = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, F.min_y), r), F.inv_h)));
    if (nMaxCellY < 0) return 0;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    int count = 0;
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
        // cells (ix, nMinCellY..nMaxCellY) are contiguous in the sorted order
        const int beg = cell_start[ix * kGridRows + nMinCellY], end = cell_start[ix * kGridRows + nMaxCellY + 1];
        for (int base = beg; base < end; base += 32) {
            const int k = base + lane;
            bool ok = false;
            int idx = 0;
            if (k < end) {
                idx = sorted_idx[k];
                const pl_keypoint kp = F.keys[idx];
                ok = true;
                if (bCheckLevels) {
                    if (kp.octave < minLevel) ok = false;
                    if (maxLevel >= 0 && kp.octave > maxLevel) ok = false;
                }
                const float distx = __fsub_rn(kp.x, x), disty = __fsub_rn(kp.y, y);
                if (!(fabsf(distx) < r && fabsf(disty) < r)) ok = false;
                if (ok) {
                    const float uR = F.u_right[idx];
                    if (uR > 0 && fabsf(__fsub_rn(ur, uR)) > gate_r) ok = false;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok && out) {
                const int dist = hamming256(q0, q1, F.desc[2 * (size_t)idx], F.desc[2 * (size_t)idx + 1]);
                out[count + __popc(m & ((1u << lane) - 1u))] = (unsigned)idx | ((unsigned)dist << 16);
            }
            count += __popc(m);
        }
    }
    return count;
}

// cv::Mat (CV_32F) expression R*x + t: one gemm, double accumulation, rounded once (core/src/matmul: GEMMSingleMul)
__device__ __forceinline__ float mat_row(const float* T, int r, float X, float Y, float Z) {
    double s = 0;
    s = __dadd_rn(s, __dmul_rn((double)T[4 * r], (double)X));
    s = __dadd_rn(s, __dmul_rn((double)T[4 * r + 1], (double)Y));
    s = __dadd_rn(s, __dmul_rn((double)T[4 * r + 2], (double)Z));
    return (float)__dadd_rn(s, (double)T[4 * r + 3]);
}

// phase A, run twice: FILL == false counts the candidates of every point (cand_n), FILL == true writes them at the
// CSR offsets.  MODE 0 = C3 (ORBmatcher.cc:1746-1830 without the claim check), 1 = C2 (ORBmatcher.cc:84-149),
// 2 = C4 (ORBmatcher.cc:1909-1960), 3 = C5 (ORBmatcher.cc:451-531), 4 = SearchForInitialization (ORBmatcher.cc:598-606).
template <int MODE, bool FILL>
__global__ void __launch_bounds__(256) k_candidates(const SearchDev* __restrict__ SD, float th, const int* __restrict__ sorted_idx,
                                                    const int* __restrict__ cell_start, int* __restrict__ cand_n,
                                                    const int* __restrict__ cand_off, const int* __restrict__ inst_base,
                                                    unsigned int* __restrict__ cand) {
    const SearchDev& S = SD[blockIdx.y];
    const FrameDev& F = S.F;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= S.np) return;
    const int* sidx = sorted_idx + S.sort_base;
    const int* cst = cell_start + (size_t)blockIdx.y * (kGridCells + 1);
    unsigned int* out = nullptr;
    if (FILL) {
        if (cand_n[S.pt_base + i] == 0) return;
        out = cand + inst_base[blockIdx.y] + cand_off[S.pt_base + i];
    }
    int count = 0;
    if (S.valid[i]) {
        const uint4 q0 = S.pdesc[2 * (size_t)i], q1 = S.pdesc[2 * (size_t)i + 1];
        if (MODE == 0) {
            const float X = S.world_pos[3 * (size_t)i], Y = S.world_pos[3 * (size_t)i + 1], Z = S.world_pos[3 * (size_t)i + 2];
            const float xc = mat_row(F.tcw, 0, X, Y, Z), yc = mat_row(F.tcw, 1, X, Y, Z), zc = mat_row(F.tcw, 2, X, Y, Z);
            const float invzc = (float)(1.0 / (double)zc);
            if (!(invzc < 0)) {
                const float u = __fadd_rn(__fmul_rn(__fmul_rn(F.fx, xc), invzc), F.cx);
                const float v = __fadd_rn(__fmul_rn(__fmul_rn(F.fy, yc), invzc), F.cy);
                if (!(u < F.min_x || u > F.max_x) && !(v < F.min_y || v > F.max_y)) {
                    const int oct = S.level[i];
                    const float radius = __fmul_rn(th, F.sf[oct]);
                    int minL, maxL;
                    if (S.forward) { minL = oct; maxL = -1; }
                    else if (S.backward) { minL = 0; maxL = oct; }
                    else { minL = oct - 1; maxL = oct + 1; }
                    const float ur = __fsub_rn(u, __fmul_rn(F.bf, invzc));
                    count = gather_candidates(F, sidx, cst, u, v, radius, minL, maxL, ur, radius, q0, q1, out);
                }
            }
        } else if (MODE == 2 || MODE == 3) {
            const float X = S.world_pos[3 * (size_t)i], Y = S.world_pos[3 * (size_t)i + 1], Z = S.world_pos[3 * (size_t)i + 2];
            const float xc = mat_row(F.tcw, 0, X, Y, Z), yc = mat_row(F.tcw, 1, X, Y, Z), zc = mat_row(F.tcw, 2, X, Y, Z);
            float u, v;
            bool ok;
            if (MODE == 2) {
                const float invzc = (float)(1.0 / (double)zc);
                u = __fadd_rn(__fmul_rn(__fmul_rn(F.fx, xc), invzc), F.cx);
                v = __fadd_rn(__fmul_rn(__fmul_rn(F.fy, yc), invzc), F.cy);
                ok = !(u < F.min_x || u > F.max_x) && !(v < F.min_y || v > F.max_y);
            } else {
                ok = !(zc < 0.0f);
                const float invz = __fdiv_rn(1.0f, zc);
                u = __fadd_rn(__fmul_rn(F.fx, __fmul_rn(xc, invz)), F.cx);
                v = __fadd_rn(__fmul_rn(F.fy, __fmul_rn(yc, invz)), F.cy);
                ok = ok && (u >= F.min_x && u < F.max_x && v >= F.min_y && v < F.max_y);  // KeyFrame::IsInImage
            }
            if (ok) {
                const float px = __fsub_rn(X, S.ow[0]), py = __fsub_rn(Y, S.ow[1]), pz = __fsub_rn(Z, S.ow[2]);
                // cv::norm of a 3x1 CV_32F: exact products accumulated in double, sqrt
                double ss = __dmul_rn((double)px, (double)px);
                ss = __dadd_rn(ss, __dmul_rn((double)py, (double)py));
                ss = __dadd_rn(ss, __dmul_rn((double)pz, (double)pz));
                const float dist3D = (float)sqrt(ss);
                ok = !(dist3D < S.min_inv[i] || dist3D > S.max_inv[i]);
                if (ok && MODE == 3) {
                    const float* nrm = S.normal + 3 * (size_t)i;
                    double dot = __dmul_rn((double)px, (double)nrm[0]);  // cv::Mat::dot: double accumulation
                    dot = __dadd_rn(dot, __dmul_rn((double)py, (double)nrm[1]));
                    dot = __dadd_rn(dot, __dmul_rn((double)pz, (double)nrm[2]));
                    ok = !(dot < __dmul_rn(0.5, (double)dist3D));
                }
                if (ok) {
                    // MapPoint::PredictScale (MapPoint.cc:397-431)
                    const float ratio = __fdiv_rn(S.max_raw[i], dist3D);
                    int lvl = (int)ceilf(__fdiv_rn(glibc_logf(ratio), S.log_sf));
                    if (lvl < 0) lvl = 0;
                    else if (lvl >= F.n_levels) lvl = F.n_levels - 1;
                    const float radius = __fmul_rn(th, F.sf[lvl]);
                    // C4: window lvl-1..lvl+1; C5: KeyFrame::GetFeaturesInArea + the level gate of :527-531
                    count = gather_candidates(F, sidx, cst, u, v, radius, lvl - 1, MODE == 2 ? lvl + 1 : lvl, 0.f, __int_as_float(0x7f800000), q0, q1, out);
                }
            }
        } else if (MODE == 4) {
            // SearchForInitialization (ORBmatcher.cc:598-606): window around vbPrevMatched[i1], level 0 only; th = windowSize
            count = gather_candidates(F, sidx, cst, S.proj_x[i], S.proj_y[i], th, 0, 0, 0.f, __int_as_float(0x7f800000), q0, q1, out);
        } else {
            const int lvl = S.level[i];
            float r = S.view_cos[i] > 0.998 ? 2.5f : 4.0f;  // RadiusByViewingCos (ORBmatcher.cc:186-193)
            if (th != 1.0f) r = __fmul_rn(r, th);
            const float rs = __fmul_rn(r, F.sf[lvl]);
            count = gather_candidates(F, sidx, cst, S.proj_x[i], S.proj_y[i], rs, lvl - 1, lvl, S.proj_xr[i], rs, q0, q1, out);
        }
    }
    if (!FILL && lane == 0) cand_n[S.pt_base + i] = count;
}

// per instance: exclusive scan of cand_n over its points -> cand_off (relative to the instance), totals[inst]
__global__ void __launch_bounds__(1024) k_cand_scan(const SearchDev* __restrict__ SD, const int* __restrict__ cand_n, int* __restrict__ cand_off,
                                                    int* __restrict__ totals) {
    __shared__ int s_warp[33];
    __shared__ int s_carry;
    const SearchDev& S = SD[blockIdx.x];
    const int tid = threadIdx.x;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < S.np; base += 1024) {
        const int i = base + tid;
        const int v = i < S.np ? cand_n[S.pt_base + i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += t;
        }
        if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            const int w0 = s_warp[tid];
            int w = w0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (tid >= o) w += t;
            }
            s_warp[tid] = w - w0;
            if (tid == 31) s_warp[32] = w;
        }
        __syncthreads();
        const int carry = s_carry;
        if (i < S.np) cand_off[S.pt_base + i] = carry + s_warp[tid >> 5] + incl - v;
        __syncthreads();
        if (tid == 0) s_carry = carry + s_warp[32];
        __syncthreads();
    }
    if (tid == 0) totals[blockIdx.x] = s_carry;
}

// ---- phase B: ordered replay out of shared memory ----
constexpr int kResThreads = 256;
constexpr int kResChunkCand = 8 * 1024;  // packed candidates per chunk (32 KB: six replay CTAs fit on an SM, which matters when the LSD grower leaves only a few SMs free)

// finds the end p1 of the chunk starting at p0 (largest p1 with off[p1]-off[p0] <= kResChunkCand) and copies its
// candidates into shared memory.  A point with more than kResChunkCand candidates is flagged and skipped.
__device__ int resolve_load_chunk(unsigned int* s_cand, const unsigned int* __restrict__ cand, const int* __restrict__ off, const int* __restrict__ cnt,
                                  int p0, int np, int total, int* s_p1, int* s_overflow) {
    const int tid = threadIdx.x;
    __syncthreads();
    if (tid == 0) {
        const int o0 = off[p0];
        int p = p0;
        // exponential + binary search on the monotone offsets
        int lo = p0, hi = np;  // invariant: chunk [p0, lo) fits
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            const int endo = mid < np ? off[mid] : total;
            if (endo - o0 <= kResChunkCand) lo = mid; else hi = mid - 1;
        }
        p = lo;
        if (p == p0) { *s_overflow = 1; p = p0 + 1; *s_p1 = -p; }  // single oversized point: skip it
        else *s_p1 = p;
    }
    __syncthreads();
    int p1 = *s_p1;
    if (p1 < 0) return -p1 | (1 << 30);  // flagged skip
    const int o0 = off[p0], o1 = p1 < np ? off[p1] : total;
    for (int k = tid; k < o1 - o0; k += kResThreads) s_cand[k] = cand[o0 + k];
    (void)cnt;
    __syncthreads();
    return p1;
}

// C3 phase B (ORBmatcher.cc:1802-1876); also C4 (:1962-2021, th_dist = ORBdist) and C5 (:519-547, th_dist = TH_LOW, no
// orientation check): best unclaimed candidate per point, claims in point order, rotation histogram
__global__ void __launch_bounds__(kResThreads) k_last_frame_resolve(const SearchDev* __restrict__ SD, int check_orientation, int th_dist,
                                                                    const unsigned int* __restrict__ cand_all, const int* __restrict__ cand_n,
                                                                    const int* __restrict__ cand_off, const int* __restrict__ inst_base,
                                                                    const int* __restrict__ totals, int* __restrict__ match_all,
                                                                    int* __restrict__ rec_all, int* __restrict__ out /* [inst][2] */) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ int s_hist[kHistoLen];
    __shared__ int s_p1, s_overflow, s_nm, s_nrec;
    const SearchDev& S = SD[blockIdx.x];
    const FrameDev& C = S.F;
    unsigned int* s_cand = (unsigned int*)s_raw;
    uint8_t* s_claimed = s_raw + (size_t)kResChunkCand * 4;
    const unsigned int* cand = cand_all + inst_base[blockIdx.x];
    const int* off = cand_off + S.pt_base;
    const int* cnt = cand_n + S.pt_base;
    const int total = totals[blockIdx.x];
    int* match = match_all + S.feat_base;
    int* rec = rec_all + 2 * (size_t)S.pt_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < C.n; i += kResThreads) {
        s_claimed[i] = C.claimed ? (C.claimed[i] != 0) : 0;
        match[i] = -1;
    }
    if (tid < kHistoLen) s_hist[tid] = 0;
    if (tid == 0) { s_overflow = 0; s_nm = 0; s_nrec = 0; }
    const float factor = kHistoLen / 360.0f;
    int p0 = 0;
    while (p0 < S.np) {
        int p1 = resolve_load_chunk(s_cand, cand, off, cnt, p0, S.np, total, &s_p1, &s_overflow);
        if (p1 & (1 << 30)) { p0 = p1 & ~(1 << 30); continue; }
        if (warp == 0) {
            // The replay is one dependent chain per point: nothing in it may wait for global memory.  The counts, offsets and
            // observation flags of 32 points are fetched by the lanes at once; the rotation bins (which decide nothing here) are
            // formed afterwards by the whole CTA from the (feature, point) pairs this loop records.
            int nmatches = s_nm, nrec = s_nrec;
            const int o0 = off[p0];
            for (int g0 = p0; g0 < p1; g0 += 32) {
                int my_c = 0, my_o = 0, my_obs = 1;
                if (g0 + lane < p1) {
                    my_c = cnt[g0 + lane];
                    my_o = off[g0 + lane] - o0;
                    if (S.has_obs) my_obs = S.has_obs[g0 + lane] != 0;
                }
                const int gn = min(32, p1 - g0);
                for (int j = 0; j < gn; j++) {
                    const int c = __shfl_sync(0xffffffffu, my_c, j);
                    if (c == 0) continue;
                    const int i = g0 + j;
                    const unsigned int* cd = s_cand + __shfl_sync(0xffffffffu, my_o, j);
                    unsigned best = 0xFFFFFFFFu;
                    for (int base = 0; base < c; base += 32) {
                        unsigned key = 0xFFFFFFFFu;
                        if (base + lane < c) {
                            const unsigned e = cd[base + lane];
                            if (!s_claimed[e & 0xFFFFu]) key = ((e >> 16) << 16) | (unsigned)(base + lane);  // dist, then candidate position
                        }
#pragma unroll
                        for (int sft = 16; sft > 0; sft >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, sft));
                        best = min(best, key);
                    }
                    if (best == 0xFFFFFFFFu) continue;
                    const int bestDist = (int)(best >> 16);
                    if (bestDist <= th_dist) {
                        const int bestIdx2 = (int)(cd[best & 0xFFFFu] & 0xFFFFu);
                        const int obs = __shfl_sync(0xffffffffu, my_obs, j);
                        if (lane == 0) {
                            match[bestIdx2] = i;
                            s_claimed[bestIdx2] = (uint8_t)obs;
                            if (check_orientation) {
                                rec[2 * nrec] = bestIdx2;
                                rec[2 * nrec + 1] = i;
                            }
                        }
                        nmatches++;
                        nrec += check_orientation != 0;
                        __syncwarp();
                    }
                }
            }
            if (lane == 0) { s_nm = nmatches; s_nrec = nrec; }
        }
        p0 = p1;
    }
    __syncthreads();
    int nmatches = s_nm;
    const int nrec = s_nrec;
    if (check_orientation) {  // rotation bin of every match (ORBmatcher.cc:1847-1857), all threads
        for (int k = tid; k < nrec; k += kResThreads) {
            const int idx = rec[2 * k], i = rec[2 * k + 1];
            float rot = __fsub_rn(S.angle[i], C.keys[idx].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, factor));
            if (bin == kHistoLen) bin = 0;
            rec[2 * k + 1] = bin;
            atomicAdd(&s_hist[bin], 1);
        }
        __syncthreads();
    }
    if (warp == 0) {
        if (check_orientation) {
            __threadfence_block();
            // ComputeThreeMaxima (ORBmatcher.cc:2035-2077)
            int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
            for (int i = 0; i < kHistoLen; i++) {
                const int sz = s_hist[i];
                if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
                else if (sz > max3) { max3 = sz; ind3 = i; }
            }
            if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
            int removed = 0;
            for (int k = lane; k < nrec; k += 32) {
                const int bin = rec[2 * k + 1];
                if (bin != ind1 && bin != ind2 && bin != ind3) {
                    match[rec[2 * k]] = -1;
                    removed++;
                }
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, sft);
            nmatches -= removed;
        }
        if (lane == 0) { out[2 * blockIdx.x] = nmatches; out[2 * blockIdx.x + 1] = s_overflow; }
    }
}

// C2 phase B (ORBmatcher.cc:123-178): best / second best with their octaves, ratio test, claims in point order.
// The reference's sequential two-minimum update is evaluated in closed form: with pb = first position of the minimum
// distance, the second best is the running minimum r of the prefix before pb (it is what the update displaces when
// pb arrives) unless a later candidate s is strictly smaller than it; first occurrences win.
__global__ void __launch_bounds__(kResThreads) k_local_points_resolve(const SearchDev* __restrict__ SD, float nn_ratio,
                                                                      const unsigned int* __restrict__ cand_all, const int* __restrict__ cand_n,
                                                                      const int* __restrict__ cand_off, const int* __restrict__ inst_base,
                                                                      const int* __restrict__ totals, int* __restrict__ match_all,
                                                                      int* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ int s_p1, s_overflow, s_nm;
    const SearchDev& S = SD[blockIdx.x];
    const FrameDev& F = S.F;
    unsigned int* s_cand = (unsigned int*)s_raw;
    uint8_t* s_claimed = s_raw + (size_t)kResChunkCand * 4;
    const unsigned int* cand = cand_all + inst_base[blockIdx.x];
    const int* off = cand_off + S.pt_base;
    const int* cnt = cand_n + S.pt_base;
    const int total = totals[blockIdx.x];
    int* match = match_all + S.feat_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // per feature: bit 7 = claimed, low bits = octave (the ratio test compares octaves: no global read inside the replay)
    for (int i = tid; i < F.n; i += kResThreads) {
        s_claimed[i] = (uint8_t)(((F.claimed && F.claimed[i] != 0) ? 0x80 : 0) | (F.keys[i].octave & 0x7f));
        match[i] = -1;
    }
    if (tid == 0) { s_overflow = 0; s_nm = 0; }
    int p0 = 0;
    while (p0 < S.np) {
        int p1 = resolve_load_chunk(s_cand, cand, off, cnt, p0, S.np, total, &s_p1, &s_overflow);
        if (p1 & (1 << 30)) { p0 = p1 & ~(1 << 30); continue; }
        if (warp == 0) {
            int nmatches = s_nm;
            const int o0 = off[p0];
            for (int g0 = p0; g0 < p1; g0 += 32) {
              int my_c = 0, my_o = 0, my_obs = 1;  // counts, offsets, observation flags of 32 points, fetched at once
              if (g0 + lane < p1) {
                  my_c = cnt[g0 + lane];
                  my_o = off[g0 + lane] - o0;
                  if (S.has_obs) my_obs = S.has_obs[g0 + lane] != 0;
              }
              const int gn = min(32, p1 - g0);
              for (int j = 0; j < gn; j++) {
                const int c = __shfl_sync(0xffffffffu, my_c, j);
                if (c == 0) continue;
                const int i = g0 + j;
                const unsigned int* cd = s_cand + __shfl_sync(0xffffffffu, my_o, j);
                // pass 1: first position of the minimum distance among the unclaimed candidates
                unsigned best = 0xFFFFFFFFu;
                for (int base = 0; base < c; base += 32) {
                    unsigned key = 0xFFFFFFFFu;
                    if (base + lane < c) {
                        const unsigned e = cd[base + lane];
                        if (!(s_claimed[e & 0xFFFFu] & 0x80)) key = ((e >> 16) << 16) | (unsigned)(base + lane);
                    }
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, sft));
                    best = min(best, key);
                }
                if (best == 0xFFFFFFFFu) continue;  // every candidate already claimed: bestDist stays 256
                const int pb = (int)(best & 0xFFFFu), bestDist = (int)(best >> 16);
                if (bestDist > kThHigh) continue;
                // pass 2: minimum before pb (r) and minimum after pb (s), first occurrences
                unsigned rk = 0xFFFFFFFFu, sk = 0xFFFFFFFFu;
                for (int base = 0; base < c; base += 32) {
                    unsigned kr = 0xFFFFFFFFu, ks = 0xFFFFFFFFu;
                    const int pos = base + lane;
                    if (pos < c && pos != pb) {
                        const unsigned e = cd[pos];
                        if (!(s_claimed[e & 0xFFFFu] & 0x80)) {
                            const unsigned key = ((e >> 16) << 16) | (unsigned)pos;
                            if (pos < pb) kr = key; else ks = key;
                        }
                    }
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) {
                        kr = min(kr, __shfl_xor_sync(0xffffffffu, kr, sft));
                        ks = min(ks, __shfl_xor_sync(0xffffffffu, ks, sft));
                    }
                    rk = min(rk, kr);
                    sk = min(sk, ks);
                }
                const int bestIdx = (int)(cd[pb] & 0xFFFFu);
                const int bestLevel = s_claimed[bestIdx] & 0x7f;
                int bestDist2 = 256, bestLevel2 = -1;
                if (rk != 0xFFFFFFFFu) {  // displaced running best of the prefix
                    bestDist2 = (int)(rk >> 16);
                    bestLevel2 = s_claimed[cd[rk & 0xFFFFu] & 0xFFFFu] & 0x7f;
                }
                if (sk != 0xFFFFFFFFu && (int)(sk >> 16) < bestDist2) {
                    bestDist2 = (int)(sk >> 16);
                    bestLevel2 = s_claimed[cd[sk & 0xFFFFu] & 0xFFFFu] & 0x7f;
                }
                if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nn_ratio, (float)bestDist2)) continue;
                const int obs = __shfl_sync(0xffffffffu, my_obs, j);
                if (lane == 0) {
                    match[bestIdx] = i;
                    if (obs) s_claimed[bestIdx] |= 0x80;
                }
                nmatches++;
                __syncwarp();
              }
            }
            if (lane == 0) s_nm = nmatches;
        }
        p0 = p1;
    }
    __syncthreads();
    if (tid == 0) { out[2 * blockIdx.x] = s_nm; out[2 * blockIdx.x + 1] = s_overflow; }
}

// ---------------------------------------------------------------------------------------------------------------
// E rows: searches without claims between points (Fuse, SearchBySim3) and SearchForInitialization's replay
// ---------------------------------------------------------------------------------------------------------------
// KeyFrame::GetFeaturesInArea (KeyFrame.cc:642-681) + the level gate [lvl-1, lvl] + (optionally) the chi-square gates of
// ORBmatcher.cc:1217-1235, reduced to the first minimum distance in visiting order.  One warp.
// Returns dist << 36 | ordinal << 16 | idx  (all ones = no candidate).
__device__ unsigned long long gather_best(const FrameDev& F, const int* __restrict__ sorted_idx, const int* __restrict__ cell_start, float x,
                                          float y, float r, int lvl, float ur, const float* inv_sigma2, uint4 q0, uint4 q1) {
    const int lane = threadIdx.x & 31;
    const unsigned long long none = ~0ull;
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, F.min_x), r), F.inv_w)));
    if (nMinCellX >= kGridCols) return none;
    const int nMaxCellX = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, F.min_x), r), F.inv_w)));
    if (nMaxCellX < 0) return none;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, F.min_y), r), F.inv_h)));
    if (nMinCellY >= kGridRows) return none;
    const int nMaxCellY = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, F.min_y), r), F.inv_h)));
    if (nMaxCellY < 0) return none;
    unsigned long long best = none;
    unsigned ordinal = 0;
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
        const int beg = cell_start[ix * kGridRows + nMinCellY], end = cell_start[ix * kGridRows + nMaxCellY + 1];
        for (int base = beg; base < end; base += 32, ordinal += 32) {
            const int k = base + lane;
            if (k >= end) continue;
            const int idx = sorted_idx[k];
            const pl_keypoint kp = F.keys[idx];
            const float distx = __fsub_rn(kp.x, x), disty = __fsub_rn(kp.y, y);
            if (!(fabsf(distx) < r && fabsf(disty) < r)) continue;
            if (kp.octave < lvl - 1 || kp.octave > lvl) continue;
            if (inv_sigma2) {
                const float ex = __fsub_rn(x, kp.x), ey = __fsub_rn(y, kp.y);
                float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                const float kpr = F.u_right[idx];
                double lim = 5.99;
                if (kpr >= 0) {
                    const float er = __fsub_rn(ur, kpr);
                    e2 = __fadd_rn(e2, __fmul_rn(er, er));
                    lim = 7.8;
                }
                if ((double)__fmul_rn(e2, inv_sigma2[kp.octave]) > lim) continue;
            }
            const unsigned dist = (unsigned)hamming256(q0, q1, F.desc[2 * (size_t)idx], F.desc[2 * (size_t)idx + 1]);
            const unsigned long long key = ((unsigned long long)dist << 36) | ((unsigned long long)(ordinal + lane) << 16) | (unsigned)idx;
            best = key < best ? key : best;
        }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, sft);
        best = o < best ? o : best;
    }
    return best;
}

// VARIANT 0 = Fuse(pKF, vpMapPoints, th) ORBmatcher.cc:1124-1249; 1 = Fuse(pKF, Scw, ...) :1314-1404;
// 2 = one direction of SearchBySim3 :1472-1546 / :1549-1617.  One warp per point, no interaction between points.
template <int VARIANT>
__global__ void __launch_bounds__(256) k_point_best(const SearchDev* __restrict__ SD, float th, int th_dist, const int* __restrict__ sorted_idx,
                                                    const int* __restrict__ cell_start, int* __restrict__ best_idx, int* __restrict__ best_dist,
                                                    int* __restrict__ hits) {
    const SearchDev& S = SD[blockIdx.y];
    const FrameDev& F = S.F;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= S.np) return;
    const int* sidx = sorted_idx + S.sort_base;
    const int* cst = cell_start + (size_t)blockIdx.y * (kGridCells + 1);
    unsigned long long best = ~0ull;
    if (S.valid[i]) {
        const float X = S.world_pos[3 * (size_t)i], Y = S.world_pos[3 * (size_t)i + 1], Z = S.world_pos[3 * (size_t)i + 2];
        float xc = mat_row(F.tcw, 0, X, Y, Z), yc = mat_row(F.tcw, 1, X, Y, Z), zc = mat_row(F.tcw, 2, X, Y, Z);
        if (VARIANT == 2) {
            const float x2 = mat_row(S.t2, 0, xc, yc, zc), y2 = mat_row(S.t2, 1, xc, yc, zc), z2 = mat_row(S.t2, 2, xc, yc, zc);
            xc = x2; yc = y2; zc = z2;
        }
        if (!(zc < 0.0f)) {
            const float invz = VARIANT == 0 ? __fdiv_rn(1.0f, zc) : (float)(1.0 / (double)zc);
            const float u = __fadd_rn(__fmul_rn(F.fx, __fmul_rn(xc, invz)), F.cx);
            const float v = __fadd_rn(__fmul_rn(F.fy, __fmul_rn(yc, invz)), F.cy);
            if (u >= F.min_x && u < F.max_x && v >= F.min_y && v < F.max_y) {  // KeyFrame::IsInImage
                const float ur = __fsub_rn(u, __fmul_rn(F.bf, invz));
                float px, py, pz;
                if (VARIANT == 2) { px = xc; py = yc; pz = zc; }  // cv::norm(p3Dc2) (:1503)
                else { px = __fsub_rn(X, S.ow[0]); py = __fsub_rn(Y, S.ow[1]); pz = __fsub_rn(Z, S.ow[2]); }
                double ss = __dmul_rn((double)px, (double)px);
                ss = __dadd_rn(ss, __dmul_rn((double)py, (double)py));
                ss = __dadd_rn(ss, __dmul_rn((double)pz, (double)pz));
                const float dist3D = (float)sqrt(ss);
                bool ok = !(dist3D < S.min_inv[i] || dist3D > S.max_inv[i]);
                if (ok && VARIANT != 2) {
                    const float* nrm = S.normal + 3 * (size_t)i;
                    double dot = __dmul_rn((double)px, (double)nrm[0]);
                    dot = __dadd_rn(dot, __dmul_rn((double)py, (double)nrm[1]));
                    dot = __dadd_rn(dot, __dmul_rn((double)pz, (double)nrm[2]));
                    ok = !(dot < __dmul_rn(0.5, (double)dist3D));
                }
                if (ok) {
                    const float ratio = __fdiv_rn(S.max_raw[i], dist3D);  // MapPoint::PredictScale (MapPoint.cc:397-431)
                    int lvl = (int)ceilf(__fdiv_rn(glibc_logf(ratio), S.log_sf));
                    if (lvl < 0) lvl = 0;
                    else if (lvl >= F.n_levels) lvl = F.n_levels - 1;
                    const float radius = __fmul_rn(th, F.sf[lvl]);
                    const uint4 q0 = S.pdesc[2 * (size_t)i], q1 = S.pdesc[2 * (size_t)i + 1];
                    best = gather_best(F, sidx, cst, u, v, radius, lvl, ur, (VARIANT == 0 && S.use_chi2) ? S.inv_sigma2 : nullptr, q0, q1);
                }
            }
        }
    }
    if (lane == 0) {
        const int bd = best == ~0ull ? 256 : (int)(best >> 36);
        const bool hit = bd <= th_dist;
        best_idx[S.pt_base + i] = hit ? (int)(best & 0xFFFFu) : -1;
        best_dist[S.pt_base + i] = bd;
        if (hit) atomicAdd(&hits[blockIdx.y], 1);
    }
}

// SearchBySim3's agreement step (ORBmatcher.cc:1620-1647): vnMatch1 / vnMatch2 are the two k_point_best results
__global__ void __launch_bounds__(256) k_sim3_agree(const int* __restrict__ m1, int n1, const int* __restrict__ m2, int n2, int* __restrict__ match12,
                                                    int* __restrict__ n_found) {
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    if (i1 < n1) {
        const int idx2 = m1[i1];
        hit = idx2 >= 0 && idx2 < n2 && m2[idx2] == i1;
        match12[i1] = hit ? idx2 : -1;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_found, __popc(m));
}

// SearchForInitialization, phase B (ORBmatcher.cc:609-708): points (features of F1) in order; a feature of F2 that already
// has a match is only offered at a strictly smaller distance (vMatchedDistance, :626), taking it un-matches its previous
// owner (:650-654).  Best / second best in closed form as in k_local_points_resolve.
constexpr int kInitMaxFeat = 20000;
__global__ void __launch_bounds__(kResThreads) k_init_resolve(const SearchDev* __restrict__ SD, float nn_ratio, int check_orientation,
                                                              const unsigned int* __restrict__ cand, const int* __restrict__ cand_n,
                                                              const int* __restrict__ cand_off, const int* __restrict__ totals,
                                                              const pl_keypoint* __restrict__ keys1, int* __restrict__ match12,
                                                              int* __restrict__ rec, float* __restrict__ prev_matched, int* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ int s_hist[kHistoLen];
    __shared__ int s_p1, s_overflow, s_nm, s_nrec;
    const SearchDev& S = SD[0];
    const FrameDev& F2 = S.F;
    unsigned int* s_cand = (unsigned int*)s_raw;
    int* s_m21 = (int*)(s_raw + (size_t)kResChunkCand * 4);
    unsigned short* s_mdist = (unsigned short*)(s_m21 + F2.n);
    const int* off = cand_off;
    const int* cnt = cand_n;
    const int total = totals[0];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < F2.n; i += kResThreads) { s_m21[i] = -1; s_mdist[i] = 0xFFFFu; }
    for (int i = tid; i < S.np; i += kResThreads) match12[i] = -1;
    if (tid < kHistoLen) s_hist[tid] = 0;
    if (tid == 0) { s_overflow = 0; s_nm = 0; s_nrec = 0; }
    const float factor = kHistoLen / 360.0f;
    int p0 = 0;
    while (p0 < S.np) {
        int p1 = resolve_load_chunk(s_cand, cand, off, cnt, p0, S.np, total, &s_p1, &s_overflow);
        if (p1 & (1 << 30)) { p0 = p1 & ~(1 << 30); continue; }
        if (warp == 0) {
            int nmatches = s_nm, nrec = s_nrec;
            const int o0 = off[p0];
            for (int i1 = p0; i1 < p1; i1++) {
                const int c = cnt[i1];
                if (c == 0) continue;
                const unsigned int* cd = s_cand + (off[i1] - o0);
                unsigned best = 0xFFFFFFFFu;
                for (int base = 0; base < c; base += 32) {
                    unsigned key = 0xFFFFFFFFu;
                    if (base + lane < c) {
                        const unsigned e = cd[base + lane];
                        if ((unsigned)s_mdist[e & 0xFFFFu] > (e >> 16)) key = ((e >> 16) << 16) | (unsigned)(base + lane);
                    }
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, sft));
                    best = min(best, key);
                }
                if (best == 0xFFFFFFFFu) continue;
                const int pb = (int)(best & 0xFFFFu), bestDist = (int)(best >> 16);
                if (bestDist > kThLow) continue;
                unsigned second = 0xFFFFFFFFu;
                for (int base = 0; base < c; base += 32) {
                    unsigned key = 0xFFFFFFFFu;
                    const int pos = base + lane;
                    if (pos < c && pos != pb) {
                        const unsigned e = cd[pos];
                        if ((unsigned)s_mdist[e & 0xFFFFu] > (e >> 16)) key = e >> 16;
                    }
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, sft));
                    second = min(second, key);
                }
                const float d2 = second == 0xFFFFFFFFu ? 2147483648.0f : (float)second;  // (float)INT_MAX
                if (!((float)bestDist < __fmul_rn(d2, nn_ratio))) continue;
                const int bestIdx2 = (int)(cd[pb] & 0xFFFFu);
                const int prev = s_m21[bestIdx2];
                if (prev >= 0) nmatches--;
                __syncwarp();
                if (lane == 0) {
                    if (prev >= 0) match12[prev] = -1;
                    match12[i1] = bestIdx2;
                    s_m21[bestIdx2] = i1;
                    s_mdist[bestIdx2] = (unsigned short)bestDist;
                }
                nmatches++;
                if (check_orientation) {
                    float rot = __fsub_rn(keys1[i1].angle, F2.keys[bestIdx2].angle);
                    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                    int bin = (int)roundf(__fmul_rn(rot, factor));
                    if (bin == kHistoLen) bin = 0;
                    if (lane == 0) {
                        rec[2 * nrec] = i1;
                        rec[2 * nrec + 1] = bin;
                        s_hist[bin]++;
                    }
                    nrec++;
                }
                __syncwarp();
            }
            if (lane == 0) { s_nm = nmatches; s_nrec = nrec; }
        }
        p0 = p1;
    }
    __threadfence_block();
    __syncthreads();
    int nmatches = s_nm;
    const int nrec = s_nrec;
    if (warp == 0 && check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < kHistoLen; i++) {
            const int sz = s_hist[i];
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
        int removed = 0;
        for (int k = lane; k < nrec; k += 32) {
            const int bin = rec[2 * k + 1], idx1 = rec[2 * k];
            if (bin != ind1 && bin != ind2 && bin != ind3 && match12[idx1] >= 0) {  // a point is recorded at most once
                match12[idx1] = -1;
                removed++;
            }
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, sft);
        nmatches -= removed;
    }
    __threadfence_block();
    __syncthreads();
    // :711-714 — vbPrevMatched follows the surviving matches
    for (int i = tid; i < S.np; i += kResThreads) {
        const int m = match12[i];
        if (m >= 0) { prev_matched[2 * i] = F2.keys[m].x; prev_matched[2 * i + 1] = F2.keys[m].y; }
    }
    if (tid == 0) { out[0] = nmatches; out[1] = s_overflow; }
}

// ---------------------------------------------------------------------------------------------------------------
// LineMatcher
// ---------------------------------------------------------------------------------------------------------------
// LineMatcher::LiangBarsky (LineMatcher.cpp:1389-1460), reproduced with its round() of the deltas and its
// horizontal-line rejection
__device__ bool liang_barsky(const double line[4], double out[4], float bx0, float by0, float bx1, float by1) {
    const double sx = line[0], sy = line[1], ex = line[2], ey = line[3];
    double p[4], q[4];
    p[0] = sx - ex; p[1] = ex - sx; p[2] = sy - ey; p[3] = ey - sy;
    q[0] = sx - (double)bx0; q[1] = (double)bx1 - sx; q[2] = sy - (double)by0; q[3] = (double)by1 - sy;
    if (p[0] == 0) { if (q[0] <= 0 || q[2] <= 0) return false; }
    if (p[2] == 0) { if (q[2] >= 0 || q[3] >= 0) return false; }
    double u_min = 0, u_max = 1;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double u = q[i] / p[i];
        if (p[i] < 0) { if (u_min < u) u_min = u; }
        else { if (u_max > u) u_max = u; }
    }
    if (u_max >= u_min) {
        out[0] = sx + round(u_min * (ex - sx));
        out[1] = sy + round(u_min * (ey - sy));
        out[2] = sx + round(u_max * (ex - sx));
        out[3] = sy + round(u_max * (ey - sy));
        return true;
    }
    return false;
}

// cv::LineIterator(...).count for 8-connectivity after cv::clipLine (imgproc/src/drawing.cpp)
__device__ int line_iterator_count(float fx0, float fy0, float fx1, float fy1, int cols, int rows) {
    long long x1 = (long long)cv_round(fx0), y1 = (long long)cv_round(fy0), x2 = (long long)cv_round(fx1), y2 = (long long)cv_round(fy1);
    const long long right = cols - 1, bottom = rows - 1;
    if (cols <= 0 || rows <= 0) return 0;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) { a = c1 < 8 ? 0 : bottom; x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1)); y1 = a; c1 = (x1 < 0) + (x1 > right) * 2; }
        if (c2 & 12) { a = c2 < 8 ? 0 : bottom; x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1)); y2 = a; c2 = (x2 < 0) + (x2 > right) * 2; }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) { a = c1 == 1 ? 0 : right; y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1)); x1 = a; c1 = 0; }
            if (c2) { a = c2 == 1 ? 0 : right; y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1)); x2 = a; c2 = 0; }
        }
    }
    if ((c1 | c2) != 0) return 0;
    const long long dx = x2 > x1 ? x2 - x1 : x1 - x2, dy = y2 > y1 ? y2 - y1 : y1 - y2;
    return (int)(dx > dy ? dx : dy) + 1;
}

// LineMatcher::UpdateKeyLineData (LineMatcher.cpp:1601-1624)
__device__ void update_keyline(const double nl[4], pl_keyline& k, int cols, int rows) {
    k.sx = (float)nl[0]; k.sy = (float)nl[1]; k.ex = (float)nl[2]; k.ey = (float)nl[3];
    k.sx_oct = k.sx; k.sy_oct = k.sy; k.ex_oct = k.ex; k.ey_oct = k.ey;
    k.pt_x = __fdiv_rn(__fadd_rn(k.ex, k.sx), 2.f);
    k.pt_y = __fdiv_rn(__fadd_rn(k.ey, k.sy), 2.f);
    const double ddx = (double)__fsub_rn(k.sx, k.ex), ddy = (double)__fsub_rn(k.sy, k.ey);
    k.length = (float)sqrt(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)));
    k.num_pixels = line_iterator_count(k.sx, k.sy, k.ex, k.ey, cols, rows);
    k.angle = (float)atan2((double)__fsub_rn(k.ey, k.sy), (double)__fsub_rn(k.ex, k.sx));
    k.size = __fmul_rn(__fsub_rn(k.ex, k.sx), __fsub_rn(k.ey, k.sy));
    k.response = __fdiv_rn(k.length, (float)max(cols, rows));
}


struct LineSearchDev {  // one LineMatcher::SearchByProjection call
    float tcw[12];
    float fx, fy, cx, cy, min_x, min_y, max_x, max_y;
    int cols, rows;
    int n_lines;                 // 3-D map lines offered
    const double *s3, *e3;
    const pl_keyline* src;
    const uint4* ldesc;          // their descriptors (MapLine::mLineDescriptor)
    const uint8_t* valid;
    int n_cur;
    const pl_keyline* cur;
    const uint4* cur_desc;
    const uint8_t* cur_claimed;
    int line_base, cur_base;     // offsets into batch-wide outputs
};

// front half (LineMatcher.cpp:96-212): projection, behind-camera handling, clipping, UpdateKeyLineData, ordered
// compaction into new_KeyLines / new_kl_index.  One CTA per instance.
__global__ void __launch_bounds__(256) k_line_project(const LineSearchDev* __restrict__ LD, pl_keyline* __restrict__ out_kl_all,
                                                      int* __restrict__ out_index_all, int* __restrict__ n_out) {
    __shared__ int s_warp[9];
    __shared__ int s_base;
    const LineSearchDev& P = LD[blockIdx.x];
    pl_keyline* out_kl = out_kl_all + P.line_base;
    int* out_index = out_index_all + P.line_base;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, n = P.n_lines;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 256) {
        const int i = base + tid;
        bool ok = false;
        pl_keyline k;
        if (i < n && P.valid[i]) {
            double T[12];
#pragma unroll
            for (int q = 0; q < 12; q++) T[q] = (double)P.tcw[q];
            const double* Xs = P.s3 + 3 * (size_t)i;
            const double* Xe = P.e3 + 3 * (size_t)i;
            double cs[3], ce[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                cs[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4 * r], Xs[0]), __dmul_rn(T[4 * r + 1], Xs[1])), __dmul_rn(T[4 * r + 2], Xs[2])), T[4 * r + 3]);
                ce[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4 * r], Xe[0]), __dmul_rn(T[4 * r + 1], Xe[1])), __dmul_rn(T[4 * r + 2], Xe[2])), T[4 * r + 3]);
            }
            double proj[4], clipped[4];
            bool have = false;
            if (!(cs[2] < 0 && ce[2] < 0)) {
                if (cs[2] < 0.0 || ce[2] < 0.0) {
                    const double lambda = -1.0 * cs[2] / (cs[2] - ce[2]);
                    const double xcr = cs[0] + lambda * (cs[0] - ce[0]);
                    const double ycr = cs[1] + lambda * (cs[1] - ce[1]);
                    if (cs[2] < 0.0) {
                        proj[0] = xcr; proj[1] = ycr;
                        proj[2] = (double)(float)((double)P.fx * ce[0] / ce[2] + (double)P.cx);
                        proj[3] = (double)(float)((double)P.fy * ce[1] / ce[2] + (double)P.cy);
                        have = true;
                    } else if (ce[2] < 0.0) {
                        proj[0] = (double)(float)((double)P.fx * cs[0] / cs[2] + (double)P.cx);
                        proj[1] = (double)(float)((double)P.fy * cs[1] / cs[2] + (double)P.cy);
                        proj[2] = xcr; proj[3] = ycr;
                        have = true;
                    }
                } else if (cs[2] > 0.0 && ce[2] > 0.0) {
                    proj[0] = (double)(float)((double)P.fx * cs[0] / cs[2] + (double)P.cx);
                    proj[1] = (double)(float)((double)P.fy * cs[1] / cs[2] + (double)P.cy);
                    proj[2] = (double)(float)((double)P.fx * ce[0] / ce[2] + (double)P.cx);
                    proj[3] = (double)(float)((double)P.fy * ce[1] / ce[2] + (double)P.cy);
                    have = true;
                }
            }
            if (have && liang_barsky(proj, clipped, P.min_x, P.min_y, P.max_x, P.max_y)) {
                k = P.src[i];
                update_keyline(clipped, k, P.cols, P.rows);
                ok = true;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_warp[w] = __popc(m);
        __syncthreads();
        int off = s_base;
        for (int q = 0; q < w; q++) off += s_warp[q];
        if (ok) {
            const int pos = off + __popc(m & ((1u << lane) - 1u));
            out_kl[pos] = k;
            out_index[pos] = i;
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int q = 0; q < 8; q++) t += s_warp[q];
            s_base += t;
        }
        __syncthreads();
    }
    if (tid == 0) n_out[blockIdx.x] = s_base;
}

// LineMatcher::LineOverLap (LineMatcher.cpp:1508-1559)
__device__ bool line_overlap(const pl_keyline& a, const pl_keyline& b, double threshold) {
    const double d1_x = (double)fabsf(__fsub_rn(a.sx, a.ex)), d2_x = (double)fabsf(__fsub_rn(b.sx, b.ex));
    const double min_x = (double)fminf(fminf(a.sx, a.ex), fminf(b.sx, b.ex)), max_x = (double)fmaxf(fmaxf(a.sx, a.ex), fmaxf(b.sx, b.ex));
    const double d1_y = (double)fabsf(__fsub_rn(a.sy, a.ey)), d2_y = (double)fabsf(__fsub_rn(b.sy, b.ey));
    const double min_y = (double)fminf(fminf(a.sy, a.ey), fminf(b.sy, b.ey)), max_y = (double)fmaxf(fmaxf(a.sy, a.ey), fmaxf(b.sy, b.ey));
    const double ovx = (d1_x + d2_x - max_x + min_x) / fmin(d1_x, d2_x);
    const double ovy = (d1_y + d2_y - max_y + min_y) / fmin(d1_y, d2_y);
    if (d1_x == 0 || d2_x == 0) { if (ovy >= threshold) return true; }
    if (d1_y == 0 || d2_y == 0) { if (ovx >= threshold) return true; }
    if (ovx >= threshold) {
        if (d1_y + d2_y + min_y >= max_y) return true;
        else if (max_y - min_y - d1_y - d2_y < 0.3 * fmin(d1_y, d2_y)) return true;
    } else if (ovx < threshold && (max_x - min_x - d1_x - d2_x) < 0.3 * fmin(d1_x, d2_x)) {
        if (ovy >= threshold) return true;
    }
    return false;
}
// LineMatcher::ReprojectionError (LineMatcher.cpp:1579-1596)
__device__ double reprojection_error(const pl_keyline& l1, const pl_keyline& l2) {
    const double ax = l1.sx, ay = l1.sy, bx = l1.ex, by = l1.ey;
    const double c0 = __dsub_rn(ay, by), c1 = __dsub_rn(bx, ax), c2 = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
    const double den = sqrt(__dadd_rn(__dmul_rn(c0, c0), __dmul_rn(c1, c1)));
    const double ds = __dadd_rn(__dadd_rn(__dmul_rn((double)l2.sx, c0), __dmul_rn((double)l2.sy, c1)), c2) / den;
    const double de = __dadd_rn(__dadd_rn(__dmul_rn((double)l2.ex, c0), __dmul_rn((double)l2.ey, c1)), c2) / den;
    return sqrt(__dadd_rn(__dmul_rn(ds, ds), __dmul_rn(de, de)));
}
// LineMatcher::LineMatching (LineMatcher.cpp:1463-1504), thresholds LineMatcher.h:94-98
__device__ bool line_matching(const pl_keyline& kl1, const pl_keyline& kl2, int hamming, const double* off) {
    const double kPi = 3.14159265358979323846;
    if ((double)hamming > 45.0 + off[3]) return false;
    if ((double)fabsf(__fsub_rn(kl1.angle, kl2.angle)) > 15.0 * kPi / 180.0 + off[0] * kPi / 180.0) return false;
    if ((double)__fdiv_rn(fminf(kl1.length, kl2.length), fmaxf(kl1.length, kl2.length)) < 0.45 + off[1]) return false;
    if (!line_overlap(kl1, kl2, 0.5 + off[2])) return false;
    if (reprojection_error(kl1, kl2) > 45.0) return false;
    return true;
}


// back half (LineMatcher.cpp:215-261): thread per current line j walks all projected lines i; the last hit wins and
// every hit counts; relaxed retry when fewer than 20 % of the current lines matched.  One CTA per instance.
// match[j] = index into the PROJECTED list (new_KeyLines) or -1.
__global__ void __launch_bounds__(256) k_line_match_pairs(const LineSearchDev* __restrict__ LD, const pl_keyline* __restrict__ proj_all,
                                                          const int* __restrict__ proj_index_all, const int* __restrict__ n_proj_all,
                                                          int* __restrict__ match_all, int* __restrict__ out /* [inst][2] */) {
    __shared__ int s_cnt;
    const LineSearchDev& P = LD[blockIdx.x];
    const pl_keyline* proj = proj_all + P.line_base;
    const int* pidx = proj_index_all + P.line_base;
    const int n_proj = n_proj_all[blockIdx.x], n_cur = P.n_cur;
    int* match = match_all + P.cur_base;
    const int tid = threadIdx.x;
    const double zero[5] = {0, 0, 0, 0, 0}, relaxed[5] = {10.0, -0.1, -0.1, 5, 10};
    if (n_cur == 0) {
        if (tid == 0) { out[2 * blockIdx.x] = 0; out[2 * blockIdx.x + 1] = 0; }
        return;
    }
    for (int pass = 0; pass < 2; pass++) {
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        const double* off = pass == 0 ? zero : relaxed;
        int local = 0;
        for (int j = tid; j < n_cur; j += 256) {
            int m = -1;
            if (!(pass == 0 && P.cur_claimed && P.cur_claimed[j])) {
                const pl_keyline kj = P.cur[j];
                const uint4 d0 = P.cur_desc[2 * (size_t)j], d1 = P.cur_desc[2 * (size_t)j + 1];
                for (int i = 0; i < n_proj; i++) {
                    const int src = pidx[i];
                    const int hd = hamming256(P.ldesc[2 * (size_t)src], P.ldesc[2 * (size_t)src + 1], d0, d1);
                    if (line_matching(proj[i], kj, hd, off)) { m = i; local++; }
                }
            }
            match[j] = m;
        }
        if (local) atomicAdd(&s_cnt, local);
        __syncthreads();
        const int cnt = s_cnt;
        if (pass == 0 && !((double)cnt * 1.0 / (double)n_cur < 0.2)) {
            if (tid == 0) { out[2 * blockIdx.x] = cnt; out[2 * blockIdx.x + 1] = 0; }
            return;
        }
        if (pass == 1 && tid == 0) { out[2 * blockIdx.x] = cnt; out[2 * blockIdx.x + 1] = 1; }
        __syncthreads();
    }
}

}  // namespace pl

// =================================================================================================================
// host side
// =================================================================================================================
using namespace pl;

namespace {

inline size_t padb(size_t b) { return PlStage::pad(b); }

size_t frame_bytes(const pl_frame_view& F) {
    const size_t n = (size_t)std::max(F.n, 0);
    return padb(n * sizeof(pl_keypoint)) + padb(n * 32) + padb(n * 4) + padb(n * 4);
}

int check_frame(const pl_frame_view* F) {
    PL_CHECK_ARG(F && F->n >= 0 && F->n <= 65535 && F->scale_factors);
    if (F->n > 32768) {  // k_frame_grid sorts (cell, index) pairs in shared memory: 4 bytes x the next power of two
        set_error("a frame view with %d key points exceeds the 32768 the grid kernel can sort in shared memory", F->n);
        return PL_ERR_CAPACITY;
    }
    PL_CHECK_ARG(F->n == 0 || (F->keys_un && F->desc && F->u_right));
    PL_CHECK_ARG(F->n_levels >= 1 && F->n_levels <= kMaxLevels && F->max_x > F->min_x && F->max_y > F->min_y);
    return PL_OK;
}

void put_frame(PlStage& st, const pl_frame_view& F, FrameDev& D) {
    D.n = F.n;
    D.keys = st.put(F.keys_un, (size_t)F.n);
    D.desc = (const uint4*)st.put(F.desc, (size_t)F.n * 32);
    D.u_right = st.put(F.u_right, (size_t)F.n);
    D.claimed = F.claimed ? st.put(F.claimed, (size_t)F.n) : nullptr;
    D.min_x = F.min_x; D.min_y = F.min_y; D.max_x = F.max_x; D.max_y = F.max_y;
    D.fx = F.fx; D.fy = F.fy; D.cx = F.cx; D.cy = F.cy; D.bf = F.bf; D.b = F.b;
    for (int i = 0; i < 12; i++) D.tcw[i] = F.tcw[i];
    for (int i = 0; i < kMaxLevels; i++) D.sf[i] = i < F.n_levels ? F.scale_factors[i] : 0.f;
    D.n_levels = F.n_levels;
    D.inv_w = (float)kGridCols / (F.max_x - F.min_x);
    D.inv_h = (float)kGridRows / (F.max_y - F.min_y);
}

// map points in view per instance (mnMatchesInliers bookkeeping of the caller: Tracking.cc:1792 counts them as visible)
__global__ void __launch_bounds__(256) k_count_in_view(const SearchDev* __restrict__ SD, int* __restrict__ cnt) {
    __shared__ int s;
    const SearchDev& S = SD[blockIdx.x];
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    int c = 0;
    for (int i = threadIdx.x; i < S.np; i += blockDim.x) c += S.valid[i] != 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) cnt[blockIdx.x] = s;
}

struct BatchScratch {
    int *sorted_idx, *cell_start, *cand_n, *cand_off, *totals, *inst_base, *match, *rec, *out;
    unsigned int* cand;
};

// the common pipeline of C2 / C3 once the SearchDev array is packed: grid -> count -> scan -> fill -> resolve
template <int MODE>
int run_search(pl_match* h, const std::vector<SearchDev>& host_sd, const SearchDev* d_sd, int n, int total_feats, int total_pts, int total_n2,
               int max_n2, int max_np, int max_feat, float th, float nn_ratio, int check_ori, int* const* match_out, int* n_matches,
               int th_dist = kThHigh) {
    cudaStream_t st = h->stream;
    BatchScratch B;
    int rc;
    void* p;
    if ((rc = match_scratch(h, 6, (size_t)std::max(total_n2, 1) * 4, &p)) != PL_OK) return rc; B.sorted_idx = (int*)p;
    if ((rc = match_scratch(h, 7, (size_t)n * (kGridCells + 1) * 4, &p)) != PL_OK) return rc; B.cell_start = (int*)p;
    if ((rc = match_scratch(h, 8, (size_t)std::max(total_pts, 1) * 4, &p)) != PL_OK) return rc; B.cand_n = (int*)p;
    if ((rc = match_scratch(h, 9, (size_t)std::max(total_pts, 1) * 4, &p)) != PL_OK) return rc; B.cand_off = (int*)p;
    if ((rc = match_scratch(h, 10, (size_t)n * 4, &p)) != PL_OK) return rc; B.totals = (int*)p;
    if ((rc = match_scratch(h, 11, (size_t)n * 4, &p)) != PL_OK) return rc; B.inst_base = (int*)p;
    if ((rc = match_scratch(h, 12, (size_t)std::max(total_feats, 1) * 4, &p)) != PL_OK) return rc; B.match = (int*)p;
    if ((rc = match_scratch(h, 13, (size_t)std::max(total_pts, 1) * 8, &p)) != PL_OK) return rc; B.rec = (int*)p;
    if ((rc = match_scratch(h, 14, (size_t)n * 8, &p)) != PL_OK) return rc; B.out = (int*)p;
    if ((size_t)max_n2 * 4 > 48 * 1024) PL_CUDA_TRY(cudaFuncSetAttribute(k_frame_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, max_n2 * 4));
    k_frame_grid<<<n, 1024, (size_t)max_n2 * 4, st>>>(d_sd, B.sorted_idx, B.cell_start);
    h->last_launches++;
    const dim3 cgrid((std::max(max_np, 1) * 32 + 255) / 256, n);
    k_candidates<MODE, false><<<cgrid, 256, 0, st>>>(d_sd, th, B.sorted_idx, B.cell_start, B.cand_n, nullptr, nullptr, nullptr);
    k_cand_scan<<<n, 1024, 0, st>>>(d_sd, B.cand_n, B.cand_off, B.totals);
    h->last_launches += 2;
    // CSR bases of the instances (one small D2H + H2D; the candidate buffer is sized from the grand total)
    if ((rc = h->res.reserve(padb((size_t)n * 4) * 2 + padb((size_t)std::max(total_feats, 1) * 4) + padb((size_t)n * 8))) != PL_OK) return rc;
    int* h_tot;
    h->res.out<int>((size_t)n, &h_tot);
    int* h_base;
    h->res.out<int>((size_t)n, &h_base);
    PL_CUDA_TRY(cudaMemcpyAsync(h_tot, B.totals, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    long long grand = 0;
    for (int i = 0; i < n; i++) { h_base[i] = (int)grand; grand += h_tot[i]; }
    if (grand > 0x7fffffffLL) { set_error("too many candidates in one batch"); return PL_ERR_CAPACITY; }
    if ((rc = match_scratch(h, 15, (size_t)std::max<long long>(grand, 1) * 4, &p)) != PL_OK) return rc;
    B.cand = (unsigned int*)p;
    PL_CUDA_TRY(cudaMemcpyAsync(B.inst_base, h_base, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    k_candidates<MODE, true><<<cgrid, 256, 0, st>>>(d_sd, th, B.sorted_idx, B.cell_start, B.cand_n, B.cand_off, B.inst_base, B.cand);
    const size_t rsm = (size_t)kResChunkCand * 4 + (size_t)std::max(max_feat, 1) + 16;
    if (MODE != 1) {
        PL_CUDA_TRY(cudaFuncSetAttribute(k_last_frame_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm));
        k_last_frame_resolve<<<n, kResThreads, rsm, st>>>(d_sd, check_ori, th_dist, B.cand, B.cand_n, B.cand_off, B.inst_base, B.totals, B.match, B.rec, B.out);
    } else {
        PL_CUDA_TRY(cudaFuncSetAttribute(k_local_points_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm));
        k_local_points_resolve<<<n, kResThreads, rsm, st>>>(d_sd, nn_ratio, B.cand, B.cand_n, B.cand_off, B.inst_base, B.totals, B.match, B.out);
    }
    h->last_launches += 2;
    PL_CUDA_TRY(cudaGetLastError());
    int *h_match, *h_out;
    h->res.out<int>((size_t)std::max(total_feats, 1), &h_match);
    h->res.out<int>((size_t)n * 2, &h_out);
    if (total_feats) PL_CUDA_TRY(cudaMemcpyAsync(h_match, B.match, (size_t)total_feats * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(cudaMemcpyAsync(h_out, B.out, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    bool overflow = false;
    for (int i = 0; i < n; i++) {
        if (host_sd[i].F.n) memcpy(match_out[i], h_match + host_sd[i].feat_base, (size_t)host_sd[i].F.n * 4);
        n_matches[i] = h_out[2 * i];
        overflow |= h_out[2 * i + 1] != 0;
    }
    if (overflow) {
        set_error("a map point had more than %d candidate features in its search window", kResChunkCand);
        return PL_ERR_CAPACITY;
    }
    return PL_OK;
}

}  // namespace

extern "C" {

PL_API int pl_orb_search_last_frame_batch(pl_match* h, int n, const pl_frame_view* cur, const pl_lastframe_view* last, float th, int mono,
                                          int check_orientation, int* const* match_of_feature, int* n_matches) {
    PL_CHECK_ARG(h && n >= 0 && (n == 0 || (cur && last && match_of_feature && n_matches)));
    if (n == 0) return PL_OK;
    size_t bytes = padb(sizeof(SearchDev) * (size_t)n);
    int total_feats = 0, total_pts = 0, total_n2 = 0, max_n2 = 1, max_np = 0, max_feat = 0;
    for (int i = 0; i < n; i++) {
        int rc = check_frame(&cur[i]);
        if (rc != PL_OK) return rc;
        const pl_lastframe_view& L = last[i];
        PL_CHECK_ARG(L.n >= 0 && (L.n == 0 || (L.valid && L.world_pos && L.desc && L.octave && L.angle)));
        for (int k = 0; k < L.n; k++) PL_CHECK_ARG(!L.valid[k] || (L.octave[k] >= 0 && L.octave[k] < cur[i].n_levels));
        PL_CHECK_ARG(match_of_feature[i] != nullptr || cur[i].n == 0);
        bytes += frame_bytes(cur[i]) + padb((size_t)L.n) * 2 + padb((size_t)L.n * 12) + padb((size_t)L.n * 32) + padb((size_t)L.n * 4) * 2;
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int rc = h->in.reserve(bytes);
    if (rc != PL_OK) return rc;
    std::vector<SearchDev> sd(n);
    for (int i = 0; i < n; i++) {
        SearchDev& S = sd[i];
        memset(&S, 0, sizeof(S));
        put_frame(h->in, cur[i], S.F);
        const pl_lastframe_view& L = last[i];
        S.np = L.n;
        S.valid = h->in.put(L.valid, (size_t)L.n);
        S.world_pos = h->in.put(L.world_pos, (size_t)L.n * 3);
        S.pdesc = (const uint4*)h->in.put(L.desc, (size_t)L.n * 32);
        S.level = h->in.put(L.octave, (size_t)L.n);
        S.angle = h->in.put(L.angle, (size_t)L.n);
        S.has_obs = L.has_observations ? h->in.put(L.has_observations, (size_t)L.n) : nullptr;
        // forward / backward decision (ORBmatcher.cc:1727-1741): twc = -Rcw^T tcw ; tlc = Rlw twc + tlw (cv::Mat float gemm)
        float twc[3], tlc[3];
        for (int r = 0; r < 3; r++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += (double)cur[i].tcw[4 * k + r] * (double)cur[i].tcw[4 * k + 3];
            twc[r] = (float)(s * -1.0);
        }
        for (int r = 0; r < 3; r++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += (double)L.tcw[4 * r + k] * (double)twc[k];
            tlc[r] = (float)(s * 1.0 + (double)L.tcw[4 * r + 3] * 1.0);
        }
        S.forward = (tlc[2] > cur[i].b && !mono) ? 1 : 0;
        S.backward = (-tlc[2] > cur[i].b && !mono) ? 1 : 0;
        int n2 = 1;
        while (n2 < std::max(cur[i].n, 1)) n2 <<= 1;
        S.n2 = n2;
        S.pt_base = total_pts; S.feat_base = total_feats; S.sort_base = total_n2;
        total_pts += L.n; total_feats += cur[i].n; total_n2 += n2;
        max_n2 = std::max(max_n2, n2); max_np = std::max(max_np, L.n); max_feat = std::max(max_feat, cur[i].n);
    }
    const SearchDev* d_sd = h->in.put(sd.data(), (size_t)n);
    { int urc = h->in.upload(h->stream); if (urc != PL_OK) return urc; }
    return run_search<0>(h, sd, d_sd, n, total_feats, total_pts, total_n2, max_n2, max_np, max_feat, th, 0.f, check_orientation, match_of_feature,
                         n_matches);
}

PL_API int pl_orb_search_local_points_batch(pl_match* h, int n, const pl_frame_view* F, const pl_mappoint_view* mps, float th, float nn_ratio,
                                            int* const* match_of_feature, int* n_matches) {
    PL_CHECK_ARG(h && n >= 0 && (n == 0 || (F && mps && match_of_feature && n_matches)));
    if (n == 0) return PL_OK;
    size_t bytes = padb(sizeof(SearchDev) * (size_t)n);
    int total_feats = 0, total_pts = 0, total_n2 = 0, max_n2 = 1, max_np = 0, max_feat = 0;
    for (int i = 0; i < n; i++) {
        int rc = check_frame(&F[i]);
        if (rc != PL_OK) return rc;
        const pl_mappoint_view& M = mps[i];
        PL_CHECK_ARG(M.n >= 0 && (M.n == 0 || (M.desc && M.track_in_view && M.proj_x && M.proj_y && M.proj_xr && M.scale_level && M.view_cos)));
        for (int k = 0; k < M.n; k++) PL_CHECK_ARG(!M.track_in_view[k] || (M.scale_level[k] >= 0 && M.scale_level[k] < F[i].n_levels));
        PL_CHECK_ARG(match_of_feature[i] != nullptr || F[i].n == 0);
        bytes += frame_bytes(F[i]) + padb((size_t)M.n) * 2 + padb((size_t)M.n * 32) + padb((size_t)M.n * 4) * 5;
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int rc = h->in.reserve(bytes);
    if (rc != PL_OK) return rc;
    std::vector<SearchDev> sd(n);
    for (int i = 0; i < n; i++) {
        SearchDev& S = sd[i];
        memset(&S, 0, sizeof(S));
        put_frame(h->in, F[i], S.F);
        const pl_mappoint_view& M = mps[i];
        S.np = M.n;
        S.valid = h->in.put(M.track_in_view, (size_t)M.n);
        S.pdesc = (const uint4*)h->in.put(M.desc, (size_t)M.n * 32);
        S.level = h->in.put(M.scale_level, (size_t)M.n);
        S.proj_x = h->in.put(M.proj_x, (size_t)M.n);
        S.proj_y = h->in.put(M.proj_y, (size_t)M.n);
        S.proj_xr = h->in.put(M.proj_xr, (size_t)M.n);
        S.view_cos = h->in.put(M.view_cos, (size_t)M.n);
        S.has_obs = M.has_observations ? h->in.put(M.has_observations, (size_t)M.n) : nullptr;
        int n2 = 1;
        while (n2 < std::max(F[i].n, 1)) n2 <<= 1;
        S.n2 = n2;
        S.pt_base = total_pts; S.feat_base = total_feats; S.sort_base = total_n2;
        total_pts += M.n; total_feats += F[i].n; total_n2 += n2;
        max_n2 = std::max(max_n2, n2); max_np = std::max(max_np, M.n); max_feat = std::max(max_feat, F[i].n);
    }
    const SearchDev* d_sd = h->in.put(sd.data(), (size_t)n);
    { int urc = h->in.upload(h->stream); if (urc != PL_OK) return urc; }
    return run_search<1>(h, sd, d_sd, n, total_feats, total_pts, total_n2, max_n2, max_np, max_feat, th, nn_ratio, 0, match_of_feature, n_matches);
}

PL_API int pl_orb_search_local_map_batch(pl_match* h, int n, const pl_frame_view* F, const float* ow, int n_maps, const pl_localmap_view* maps,
                                         const int* map_of_frame, float viewing_cos_limit, float log_scale_factor, float th, float nn_ratio,
                                         int* const* match_of_feature, int* n_matches, int* n_in_view) {
    PL_CHECK_ARG(h && n >= 0 && n_maps >= 0 && log_scale_factor > 0.f);
    PL_CHECK_ARG(n == 0 || (F && ow && maps && map_of_frame && match_of_feature && n_matches && n_maps > 0));
    if (n == 0) return PL_OK;
    size_t bytes = padb(sizeof(SearchDev) * (size_t)n) + padb((size_t)n * 48) + padb((size_t)n * 12);
    for (int k = 0; k < n_maps; k++) {
        const pl_localmap_view& M = maps[k];
        PL_CHECK_ARG(M.n >= 0 && (M.n == 0 || (M.world_pos && M.normal && M.desc && M.min_dist_inv && M.max_dist_inv && M.max_dist)));
        const size_t m = (size_t)M.n;
        bytes += padb(m * 12) * 2 + padb(m * 32) + padb(m * 4) * 3 + padb(m);
    }
    int total_feats = 0, total_n2 = 0, max_n2 = 1, max_np = 0, max_feat = 0;
    long long total_pts_ll = 0;
    for (int i = 0; i < n; i++) {
        int rc = check_frame(&F[i]);
        if (rc != PL_OK) return rc;
        PL_CHECK_ARG(map_of_frame[i] >= 0 && map_of_frame[i] < n_maps);
        PL_CHECK_ARG(match_of_feature[i] != nullptr || F[i].n == 0);
        bytes += frame_bytes(F[i]);
        total_pts_ll += maps[map_of_frame[i]].n;
    }
    if (total_pts_ll > 0x7fffffffLL / 4) { set_error("too many (frame, map point) pairs in one batch"); return PL_ERR_CAPACITY; }
    const int total_pts = (int)total_pts_ll;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int rc = h->in.reserve(bytes);
    if (rc != PL_OK) return rc;
    // what IsInFrustum leaves on the map points, per (frame, map point): device only
    const size_t tp = (size_t)std::max(total_pts, 1), plane4 = padb(tp * 4);
    void* p;
    if ((rc = match_scratch(h, 16, padb(tp) + plane4 * 5, &p)) != PL_OK) return rc;
    uint8_t* d_iv = (uint8_t*)p;
    float* d_x = (float*)(d_iv + padb(tp));
    float* d_y = (float*)((uint8_t*)d_x + plane4);
    float* d_xr = (float*)((uint8_t*)d_y + plane4);
    int* d_lvl = (int*)((uint8_t*)d_xr + plane4);
    float* d_vc = (float*)((uint8_t*)d_lvl + plane4);
    if ((rc = match_scratch(h, 17, (size_t)n * 4, &p)) != PL_OK) return rc;
    int* d_cnt = (int*)p;
    struct MapDev { const float *pos, *normal, *mi, *ma, *mr; const uint4* desc; const uint8_t* obs; };
    std::vector<MapDev> md(n_maps);
    for (int k = 0; k < n_maps; k++) {
        const pl_localmap_view& M = maps[k];
        const size_t m = (size_t)M.n;
        md[k].pos = h->in.put(M.world_pos, m * 3);
        md[k].normal = h->in.put(M.normal, m * 3);
        md[k].desc = (const uint4*)h->in.put(M.desc, m * 32);
        md[k].mi = h->in.put(M.min_dist_inv, m);
        md[k].ma = h->in.put(M.max_dist_inv, m);
        md[k].mr = h->in.put(M.max_dist, m);
        md[k].obs = M.has_observations ? h->in.put(M.has_observations, m) : nullptr;
    }
    std::vector<float> tcw((size_t)n * 12);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 12; k++) tcw[(size_t)i * 12 + k] = F[i].tcw[k];
    const float* d_tcw = h->in.put(tcw.data(), tcw.size());
    const float* d_ow = h->in.put(ow, (size_t)n * 3);
    std::vector<SearchDev> sd(n);
    int pt = 0;
    for (int i = 0; i < n; i++) {
        SearchDev& S = sd[i];
        memset(&S, 0, sizeof(S));
        put_frame(h->in, F[i], S.F);
        const MapDev& D = md[map_of_frame[i]];
        const int m = maps[map_of_frame[i]].n;
        S.np = m;
        S.valid = d_iv + pt;
        S.pdesc = D.desc;
        S.level = d_lvl + pt;
        S.proj_x = d_x + pt;
        S.proj_y = d_y + pt;
        S.proj_xr = d_xr + pt;
        S.view_cos = d_vc + pt;
        S.has_obs = D.obs;
        int n2 = 1;
        while (n2 < std::max(F[i].n, 1)) n2 <<= 1;
        S.n2 = n2;
        S.pt_base = pt; S.feat_base = total_feats; S.sort_base = total_n2;
        pt += m; total_feats += F[i].n; total_n2 += n2;
        max_n2 = std::max(max_n2, n2); max_np = std::max(max_np, m); max_feat = std::max(max_feat, F[i].n);
    }
    const SearchDev* d_sd = h->in.put(sd.data(), (size_t)n);
    cudaStream_t st = h->stream;
    { int urc = h->in.upload(st); if (urc != PL_OK) return urc; }
    // Frame::IsInFrustum: consecutive frames that look at the same snapshot with the same camera are one launch
    int launches = 0;
    for (int i = 0; i < n;) {
        int j = i + 1;
        const pl_frame_view& A = F[i];
        while (j < n && map_of_frame[j] == map_of_frame[i] && F[j].fx == A.fx && F[j].fy == A.fy && F[j].cx == A.cx && F[j].cy == A.cy &&
               F[j].bf == A.bf && F[j].min_x == A.min_x && F[j].min_y == A.min_y && F[j].max_x == A.max_x && F[j].max_y == A.max_y &&
               F[j].n_levels == A.n_levels && j - i < 65535)
            j++;
        const MapDev& D = md[map_of_frame[i]];
        const float bounds[4] = {A.min_x, A.min_y, A.max_x, A.max_y};
        const int b = sd[i].pt_base;
        if ((rc = pl::launch_is_in_frustum(st, j - i, d_tcw + (size_t)i * 12, d_ow + (size_t)i * 3, A.fx, A.fy, A.cx, A.cy, A.bf, bounds, log_scale_factor,
                                           viewing_cos_limit, A.n_levels, maps[map_of_frame[i]].n, D.pos, D.normal, D.mi, D.ma, D.mr, d_iv + b, d_x + b,
                                           d_y + b, d_xr + b, d_lvl + b, d_vc + b)) != PL_OK)
            return rc;
        launches += maps[map_of_frame[i]].n > 0;
        i = j;
    }
    k_count_in_view<<<n, 256, 0, st>>>(d_sd, d_cnt);
    launches++;
    rc = run_search<1>(h, sd, d_sd, n, total_feats, total_pts, total_n2, max_n2, max_np, max_feat, th, nn_ratio, 0, match_of_feature, n_matches);
    h->last_launches += launches;
    if (rc != PL_OK) return rc;
    if (n_in_view) PL_CUDA_TRY(cudaMemcpy(n_in_view, d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost));  // (the stream is idle: run_search waited for it)
    return PL_OK;
}

PL_API int pl_orb_search_last_frame(pl_match* h, const pl_frame_view* Cur, const pl_lastframe_view* Last, float th, int mono,
                                    int check_orientation, int* match_of_feature, int* n_matches) {
    PL_CHECK_ARG(Cur && Last && match_of_feature && n_matches);
    int* mo[1] = {match_of_feature};
    return pl_orb_search_last_frame_batch(h, 1, Cur, Last, th, mono, check_orientation, mo, n_matches);
}

PL_API int pl_orb_search_local_points(pl_match* h, const pl_frame_view* F, const pl_mappoint_view* mps, float th, float nn_ratio,
                                      int* match_of_feature, int* n_matches) {
    PL_CHECK_ARG(F && mps && match_of_feature && n_matches);
    int* mo[1] = {match_of_feature};
    return pl_orb_search_local_points_batch(h, 1, F, mps, th, nn_ratio, mo, n_matches);
}

}  // extern "C"

// C4 / C5 share the packing: pose-based projection of map points described by pl_posepoint_view
namespace {
template <int MODE>
int pose_points_search(pl_match* h, int n, const pl_frame_view* fr, const pl_posepoint_view* pts, const float* ow, const float* log_sf, float th,
                       int th_dist, int check_orientation, int* const* match_of_feature, int* n_matches) {
    PL_CHECK_ARG(h && n >= 0 && (n == 0 || (fr && pts && ow && log_sf && match_of_feature && n_matches)));
    if (n == 0) return PL_OK;
    size_t bytes = padb(sizeof(SearchDev) * (size_t)n);
    int total_feats = 0, total_pts = 0, total_n2 = 0, max_n2 = 1, max_np = 0, max_feat = 0;
    for (int i = 0; i < n; i++) {
        int rc = check_frame(&fr[i]);
        if (rc != PL_OK) return rc;
        const pl_posepoint_view& P = pts[i];
        PL_CHECK_ARG(P.n >= 0 && (P.n == 0 || (P.valid && P.world_pos && P.desc && P.min_dist_inv && P.max_dist_inv && P.max_dist)));
        PL_CHECK_ARG(P.n == 0 || (MODE == 2 ? (!check_orientation || P.angle) : P.normal != nullptr));
        PL_CHECK_ARG(log_sf[i] > 0.f);
        PL_CHECK_ARG(match_of_feature[i] != nullptr || fr[i].n == 0);
        bytes += frame_bytes(fr[i]) + padb((size_t)P.n) + padb((size_t)P.n * 12) * 2 + padb((size_t)P.n * 32) + padb((size_t)P.n * 4) * 4;
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int rc = h->in.reserve(bytes);
    if (rc != PL_OK) return rc;
    std::vector<SearchDev> sd(n);
    for (int i = 0; i < n; i++) {
        SearchDev& S = sd[i];
        memset(&S, 0, sizeof(S));
        put_frame(h->in, fr[i], S.F);
        const pl_posepoint_view& P = pts[i];
        S.np = P.n;
        S.valid = h->in.put(P.valid, (size_t)P.n);
        S.world_pos = h->in.put(P.world_pos, (size_t)P.n * 3);
        S.pdesc = (const uint4*)h->in.put(P.desc, (size_t)P.n * 32);
        S.min_inv = h->in.put(P.min_dist_inv, (size_t)P.n);
        S.max_inv = h->in.put(P.max_dist_inv, (size_t)P.n);
        S.max_raw = h->in.put(P.max_dist, (size_t)P.n);
        S.angle = (MODE == 2 && P.angle) ? h->in.put(P.angle, (size_t)P.n) : nullptr;
        S.normal = MODE == 3 ? h->in.put(P.normal, (size_t)P.n * 3) : nullptr;
        S.has_obs = nullptr;  // a feature that received a map point is taken (:1965, :521)
        for (int k = 0; k < 3; k++) S.ow[k] = ow[3 * (size_t)i + k];
        S.log_sf = log_sf[i];
        int n2 = 1;
        while (n2 < std::max(fr[i].n, 1)) n2 <<= 1;
        S.n2 = n2;
        S.pt_base = total_pts; S.feat_base = total_feats; S.sort_base = total_n2;
        total_pts += P.n; total_feats += fr[i].n; total_n2 += n2;
        max_n2 = std::max(max_n2, n2); max_np = std::max(max_np, P.n); max_feat = std::max(max_feat, fr[i].n);
    }
    const SearchDev* d_sd = h->in.put(sd.data(), (size_t)n);
    { int urc = h->in.upload(h->stream); if (urc != PL_OK) return urc; }
    return run_search<MODE>(h, sd, d_sd, n, total_feats, total_pts, total_n2, max_n2, max_np, max_feat, th, 0.f, check_orientation,
                            match_of_feature, n_matches, th_dist);
}
}  // namespace

extern "C" {

PL_API int pl_orb_search_keyframe_points_batch(pl_match* h, int n, const pl_frame_view* cur, const pl_posepoint_view* pts, const float* ow,
                                               const float* log_scale_factor, float th, int orb_dist, int check_orientation,
                                               int* const* match_of_feature, int* n_matches) {
    return pose_points_search<2>(h, n, cur, pts, ow, log_scale_factor, th, orb_dist, check_orientation, match_of_feature, n_matches);
}

PL_API int pl_orb_search_sim3_points_batch(pl_match* h, int n, const pl_frame_view* kf, const pl_posepoint_view* pts, const float* ow,
                                           const float* log_scale_factor, int th, int* const* match_of_feature, int* n_matches) {
    return pose_points_search<3>(h, n, kf, pts, ow, log_scale_factor, (float)th, kThLow, 0, match_of_feature, n_matches);
}

// ---- LineMatcher ----
PL_API int pl_line_search_by_projection_batch(pl_match* h, int n, const pl_lineframe_view* cur, const pl_mapline_view* lines,
                                              int* const* match_of_line, int* n_matches, int* used_relaxed, pl_keyline* const* new_keylines,
                                              int* const* new_kl_index, int* n_projected) {
    PL_CHECK_ARG(h && n >= 0 && (n == 0 || (cur && lines && match_of_line && n_matches && used_relaxed)));
    if (n == 0) return PL_OK;
    size_t bytes = padb(sizeof(LineSearchDev) * (size_t)n);
    int total_lines = 0, total_cur = 0;
    for (int i = 0; i < n; i++) {
        const pl_mapline_view& L = lines[i];
        const pl_lineframe_view& Cv = cur[i];
        PL_CHECK_ARG(L.n >= 0 && (L.n == 0 || (L.start3d && L.end3d && L.kl && L.desc && L.valid)));
        PL_CHECK_ARG(Cv.n >= 0 && (Cv.n == 0 || (Cv.kl && Cv.desc && match_of_line[i])));
        bytes += padb((size_t)L.n * 24) * 2 + padb((size_t)L.n * sizeof(pl_keyline)) + padb((size_t)L.n * 32) + padb((size_t)L.n) +
                 padb((size_t)Cv.n * sizeof(pl_keyline)) + padb((size_t)Cv.n * 32) + padb((size_t)Cv.n);
        total_lines += L.n;
        total_cur += Cv.n;
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int rc = h->in.reserve(bytes);
    if (rc != PL_OK) return rc;
    std::vector<LineSearchDev> ld(n);
    int lb = 0, cb = 0;
    for (int i = 0; i < n; i++) {
        LineSearchDev& D = ld[i];
        memset(&D, 0, sizeof(D));
        const pl_mapline_view& L = lines[i];
        const pl_lineframe_view& Cv = cur[i];
        for (int k = 0; k < 12; k++) D.tcw[k] = Cv.tcw[k];
        D.fx = Cv.fx; D.fy = Cv.fy; D.cx = Cv.cx; D.cy = Cv.cy;
        D.min_x = Cv.min_x; D.min_y = Cv.min_y; D.max_x = Cv.max_x; D.max_y = Cv.max_y;
        D.cols = Cv.cols; D.rows = Cv.rows;
        D.n_lines = L.n;
        D.s3 = h->in.put(L.start3d, (size_t)L.n * 3);
        D.e3 = h->in.put(L.end3d, (size_t)L.n * 3);
        D.src = h->in.put(L.kl, (size_t)L.n);
        D.ldesc = (const uint4*)h->in.put(L.desc, (size_t)L.n * 32);
        D.valid = h->in.put(L.valid, (size_t)L.n);
        D.n_cur = Cv.n;
        D.cur = h->in.put(Cv.kl, (size_t)Cv.n);
        D.cur_desc = (const uint4*)h->in.put(Cv.desc, (size_t)Cv.n * 32);
        D.cur_claimed = Cv.claimed ? h->in.put(Cv.claimed, (size_t)Cv.n) : nullptr;
        D.line_base = lb; D.cur_base = cb;
        lb += L.n; cb += Cv.n;
    }
    const LineSearchDev* d_ld = h->in.put(ld.data(), (size_t)n);
    cudaStream_t st = h->stream;
    { int urc = h->in.upload(st); if (urc != PL_OK) return urc; }
    void* p;
    if ((rc = match_scratch(h, 6, (size_t)std::max(total_lines, 1) * sizeof(pl_keyline), &p)) != PL_OK) return rc;
    pl_keyline* d_pk = (pl_keyline*)p;
    if ((rc = match_scratch(h, 7, (size_t)std::max(total_lines, 1) * 4, &p)) != PL_OK) return rc;
    int* d_pi = (int*)p;
    if ((rc = match_scratch(h, 8, (size_t)n * 4, &p)) != PL_OK) return rc;
    int* d_np = (int*)p;
    if ((rc = match_scratch(h, 9, (size_t)std::max(total_cur, 1) * 4, &p)) != PL_OK) return rc;
    int* d_match = (int*)p;
    if ((rc = match_scratch(h, 10, (size_t)n * 8, &p)) != PL_OK) return rc;
    int* d_out = (int*)p;
    k_line_project<<<n, 256, 0, st>>>(d_ld, d_pk, d_pi, d_np);
    k_line_match_pairs<<<n, 256, 0, st>>>(d_ld, d_pk, d_pi, d_np, d_match, d_out);
    h->last_launches += 2;
    PL_CUDA_TRY(cudaGetLastError());
    const bool want_proj = new_keylines && new_kl_index && n_projected;
    if ((rc = h->res.reserve(padb((size_t)std::max(total_cur, 1) * 4) + padb((size_t)n * 8) + padb((size_t)n * 4) +
                             padb((size_t)std::max(total_lines, 1) * sizeof(pl_keyline)) + padb((size_t)std::max(total_lines, 1) * 4))) != PL_OK)
        return rc;
    int *h_match, *h_out, *h_np, *h_pi;
    pl_keyline* h_pk;
    h->res.out<int>((size_t)std::max(total_cur, 1), &h_match);
    h->res.out<int>((size_t)n * 2, &h_out);
    h->res.out<int>((size_t)n, &h_np);
    h->res.out<pl_keyline>((size_t)std::max(total_lines, 1), &h_pk);
    h->res.out<int>((size_t)std::max(total_lines, 1), &h_pi);
    if (total_cur) PL_CUDA_TRY(cudaMemcpyAsync(h_match, d_match, (size_t)total_cur * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(cudaMemcpyAsync(h_out, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(cudaMemcpyAsync(h_np, d_np, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(cudaMemcpyAsync(h_pi, d_pi, (size_t)std::max(total_lines, 1) * 4, cudaMemcpyDeviceToHost, st));
    if (want_proj && total_lines) PL_CUDA_TRY(cudaMemcpyAsync(h_pk, d_pk, (size_t)total_lines * sizeof(pl_keyline), cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    for (int i = 0; i < n; i++) {
        const LineSearchDev& D = ld[i];
        // match index: position in new_KeyLines -> index of the original map line (new_kl_index), as the reference
        // assigns LastFrame.mvpMapLines[new_kl_index[i]] (LineMatcher.cpp:226-228)
        for (int j = 0; j < D.n_cur; j++) {
            const int m = h_match[D.cur_base + j];
            match_of_line[i][j] = m < 0 ? -1 : h_pi[D.line_base + m];
        }
        n_matches[i] = h_out[2 * i];
        used_relaxed[i] = h_out[2 * i + 1];
        if (want_proj) {
            n_projected[i] = h_np[i];
            if (new_keylines[i] && h_np[i]) memcpy(new_keylines[i], h_pk + D.line_base, (size_t)h_np[i] * sizeof(pl_keyline));
            if (new_kl_index[i] && h_np[i]) memcpy(new_kl_index[i], h_pi + D.line_base, (size_t)h_np[i] * 4);
        }
    }
    return PL_OK;
}

// the two halves as separate calls (kept for callers that need new_KeyLines / new_kl_index, e.g. the reference's
// test-only overloads LineMatcher.cpp:272-487)
PL_API int pl_line_project(pl_match* h, const double* start3d, const double* end3d, const pl_keyline* src_kl, const uint8_t* valid, int n,
                           const float tcw[12], float fx, float fy, float cx, float cy, float min_x, float min_y, float max_x, float max_y,
                           int img_cols, int img_rows, pl_keyline* out_kl, int* out_index, int* n_out) {
    PL_CHECK_ARG(h && tcw && out_kl && out_index && n_out && n >= 0);
    *n_out = 0;
    if (n == 0) return PL_OK;
    PL_CHECK_ARG(start3d && end3d && src_kl && valid);
    std::vector<uint8_t> zdesc((size_t)n * 32, 0);
    pl_mapline_view L;
    L.n = n; L.start3d = start3d; L.end3d = end3d; L.kl = src_kl; L.desc = zdesc.data(); L.valid = valid;
    pl_lineframe_view Cv;
    memset(&Cv, 0, sizeof(Cv));
    for (int i = 0; i < 12; i++) Cv.tcw[i] = tcw[i];
    Cv.fx = fx; Cv.fy = fy; Cv.cx = cx; Cv.cy = cy; Cv.min_x = min_x; Cv.min_y = min_y; Cv.max_x = max_x; Cv.max_y = max_y;
    Cv.cols = img_cols; Cv.rows = img_rows;
    int nm = 0, rel = 0, dummy = 0;
    int* mo[1] = {&dummy};
    pl_keyline* ok[1] = {out_kl};
    int* oi[1] = {out_index};
    return pl_line_search_by_projection_batch(h, 1, &Cv, &L, mo, &nm, &rel, ok, oi, n_out);
}

}  // extern "C"

namespace pl {
// all-pairs matching of already projected lines (identity projection list)
__global__ void __launch_bounds__(32) k_iota(int* p, int n) {
    for (int i = threadIdx.x; i < n; i += 32) p[i] = i;
}
}  // namespace pl

extern "C" {
PL_API int pl_line_match_pairs(pl_match* h, const pl_keyline* proj, const uint8_t* proj_desc, int n_proj, const pl_keyline* cur,
                               const uint8_t* cur_desc, const uint8_t* cur_claimed, int n_cur, int* match_of_line, int* n_matches,
                               int* used_relaxed) {
    PL_CHECK_ARG(h && match_of_line && n_matches && used_relaxed && n_proj >= 0 && n_cur >= 0);
    PL_CHECK_ARG((n_proj == 0 || (proj && proj_desc)) && (n_cur == 0 || (cur && cur_desc)));
    *n_matches = 0;
    *used_relaxed = 0;
    if (n_cur == 0) return PL_OK;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int rc = h->in.reserve(padb(sizeof(LineSearchDev)) + padb((size_t)n_proj * sizeof(pl_keyline)) + padb((size_t)n_proj * 32) +
                           padb((size_t)n_cur * sizeof(pl_keyline)) + padb((size_t)n_cur * 32) + padb((size_t)n_cur));
    if (rc != PL_OK) return rc;
    LineSearchDev D;
    memset(&D, 0, sizeof(D));
    D.n_lines = n_proj;
    const pl_keyline* d_proj = h->in.put(proj, (size_t)n_proj);
    D.ldesc = (const uint4*)h->in.put(proj_desc, (size_t)n_proj * 32);
    D.n_cur = n_cur;
    D.cur = h->in.put(cur, (size_t)n_cur);
    D.cur_desc = (const uint4*)h->in.put(cur_desc, (size_t)n_cur * 32);
    D.cur_claimed = cur_claimed ? h->in.put(cur_claimed, (size_t)n_cur) : nullptr;
    const LineSearchDev* d_ld = h->in.put(&D, 1);
    cudaStream_t st = h->stream;
    { int urc = h->in.upload(st); if (urc != PL_OK) return urc; }
    void* p;
    if ((rc = match_scratch(h, 7, (size_t)std::max(n_proj, 1) * 4, &p)) != PL_OK) return rc;
    int* d_pi = (int*)p;
    if ((rc = match_scratch(h, 8, 16, &p)) != PL_OK) return rc;
    int* d_np = (int*)p;
    if ((rc = match_scratch(h, 9, (size_t)n_cur * 4, &p)) != PL_OK) return rc;
    int* d_match = (int*)p;
    if ((rc = match_scratch(h, 10, 16, &p)) != PL_OK) return rc;
    int* d_out = (int*)p;
    k_iota<<<1, 32, 0, st>>>(d_pi, n_proj);
    PL_CUDA_TRY(cudaMemcpyAsync(d_np, &n_proj, 4, cudaMemcpyHostToDevice, st));
    k_line_match_pairs<<<1, 256, 0, st>>>(d_ld, d_proj, d_pi, d_np, d_match, d_out);
    h->last_launches += 2;
    PL_CUDA_TRY(cudaGetLastError());
    int res[2] = {0, 0};
    PL_CUDA_TRY(cudaMemcpyAsync(match_of_line, d_match, (size_t)n_cur * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(cudaMemcpyAsync(res, d_out, 8, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    *n_matches = res[0];
    *used_relaxed = res[1];
    return PL_OK;
}
}  // extern "C"

// ---- E rows: Fuse / SearchBySim3 / SearchForInitialization ----
namespace {
int check_posepoints(const pl_posepoint_view& P, bool need_normal) {
    PL_CHECK_ARG(P.n >= 0 && (P.n == 0 || (P.valid && P.world_pos && P.desc && P.min_dist_inv && P.max_dist_inv && P.max_dist)));
    PL_CHECK_ARG(P.n == 0 || !need_normal || P.normal != nullptr);
    return PL_OK;
}
size_t posepoint_bytes(const pl_posepoint_view& P) {
    return padb((size_t)P.n) + padb((size_t)P.n * 12) * 2 + padb((size_t)P.n * 32) + padb((size_t)P.n * 4) * 3;
}
void put_posepoints(PlStage& st, const pl_posepoint_view& P, SearchDev& S, bool with_normal) {
    S.np = P.n;
    S.valid = st.put(P.valid, (size_t)P.n);
    S.world_pos = st.put(P.world_pos, (size_t)P.n * 3);
    S.pdesc = (const uint4*)st.put(P.desc, (size_t)P.n * 32);
    S.min_inv = st.put(P.min_dist_inv, (size_t)P.n);
    S.max_inv = st.put(P.max_dist_inv, (size_t)P.n);
    S.max_raw = st.put(P.max_dist, (size_t)P.n);
    S.normal = with_normal ? st.put(P.normal, (size_t)P.n * 3) : nullptr;
}
struct IndepLayout { int total_pts = 0, total_n2 = 0, max_n2 = 1, max_np = 0; };
void place(SearchDev& S, int nfeat, IndepLayout& L) {
    int n2 = 1;
    while (n2 < std::max(nfeat, 1)) n2 <<= 1;
    S.n2 = n2;
    S.pt_base = L.total_pts; S.feat_base = 0; S.sort_base = L.total_n2;
    L.total_pts += S.np; L.total_n2 += n2;
    L.max_n2 = std::max(L.max_n2, n2); L.max_np = std::max(L.max_np, S.np);
}
// grid + k_point_best for n packed instances; results stay on the device (best_idx / best_dist / hits in scratch 12 / 13 / 14)
template <int VARIANT>
int run_point_best(pl_match* h, const SearchDev* d_sd, int n, const IndepLayout& L, float th, int th_dist, int** d_idx, int** d_dist, int** d_hits) {
    cudaStream_t st = h->stream;
    int rc;
    void* p;
    int *sorted_idx, *cell_start;
    if ((rc = match_scratch(h, 6, (size_t)std::max(L.total_n2, 1) * 4, &p)) != PL_OK) return rc; sorted_idx = (int*)p;
    if ((rc = match_scratch(h, 7, (size_t)n * (kGridCells + 1) * 4, &p)) != PL_OK) return rc; cell_start = (int*)p;
    if ((rc = match_scratch(h, 12, (size_t)std::max(L.total_pts, 1) * 4, &p)) != PL_OK) return rc; *d_idx = (int*)p;
    if ((rc = match_scratch(h, 13, (size_t)std::max(L.total_pts, 1) * 8, &p)) != PL_OK) return rc; *d_dist = (int*)p;
    if ((rc = match_scratch(h, 14, (size_t)n * 8, &p)) != PL_OK) return rc; *d_hits = (int*)p;
    PL_CUDA_TRY(cudaMemsetAsync(*d_hits, 0, (size_t)n * 8, st));
    if ((size_t)L.max_n2 * 4 > 48 * 1024) PL_CUDA_TRY(cudaFuncSetAttribute(k_frame_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, L.max_n2 * 4));
    k_frame_grid<<<n, 1024, (size_t)L.max_n2 * 4, st>>>(d_sd, sorted_idx, cell_start);
    h->last_launches++;
    if (L.max_np > 0) {
        const dim3 grid((L.max_np * 32 + 255) / 256, n);
        k_point_best<VARIANT><<<grid, 256, 0, st>>>(d_sd, th, th_dist, sorted_idx, cell_start, *d_idx, *d_dist, *d_hits);
        h->last_launches++;
    }
    PL_CUDA_TRY(cudaGetLastError());
    return PL_OK;
}
}  // namespace

extern "C" {

PL_API int pl_orb_fuse_candidates_batch(pl_match* h, int n, const pl_frame_view* kf, const pl_posepoint_view* pts, const float* ow,
                                        const float* log_scale_factor, const float* const* inv_level_sigma2, float th, int variant,
                                        int* const* best_idx, int* const* best_dist, int* n_fused) {
    PL_CHECK_ARG(h && n >= 0 && (variant == 0 || variant == 1) && (n == 0 || (kf && pts && ow && log_scale_factor && best_idx && n_fused)));
    PL_CHECK_ARG(n == 0 || variant == 1 || inv_level_sigma2);
    if (n == 0) return PL_OK;
    size_t bytes = padb(sizeof(SearchDev) * (size_t)n);
    for (int i = 0; i < n; i++) {
        int rc = check_frame(&kf[i]);
        if (rc != PL_OK) return rc;
        if ((rc = check_posepoints(pts[i], true)) != PL_OK) return rc;
        PL_CHECK_ARG(log_scale_factor[i] > 0.f && (best_idx[i] != nullptr || pts[i].n == 0));
        PL_CHECK_ARG(variant == 1 || inv_level_sigma2[i] != nullptr);
        bytes += frame_bytes(kf[i]) + posepoint_bytes(pts[i]);
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int rc = h->in.reserve(bytes);
    if (rc != PL_OK) return rc;
    std::vector<SearchDev> sd(n);
    IndepLayout L;
    for (int i = 0; i < n; i++) {
        SearchDev& S = sd[i];
        memset(&S, 0, sizeof(S));
        put_frame(h->in, kf[i], S.F);
        put_posepoints(h->in, pts[i], S, true);
        for (int k = 0; k < 3; k++) S.ow[k] = ow[3 * (size_t)i + k];
        S.log_sf = log_scale_factor[i];
        if (variant == 0) {
            S.use_chi2 = 1;
            for (int k = 0; k < kf[i].n_levels; k++) S.inv_sigma2[k] = inv_level_sigma2[i][k];
        }
        place(S, kf[i].n, L);
    }
    const SearchDev* d_sd = h->in.put(sd.data(), (size_t)n);
    if ((rc = h->in.upload(h->stream)) != PL_OK) return rc;
    int *d_idx, *d_dist, *d_hits;
    rc = variant == 0 ? run_point_best<0>(h, d_sd, n, L, th, kThLow, &d_idx, &d_dist, &d_hits)
                      : run_point_best<1>(h, d_sd, n, L, th, kThLow, &d_idx, &d_dist, &d_hits);
    if (rc != PL_OK) return rc;
    const size_t np = (size_t)std::max(L.total_pts, 1);
    if ((rc = h->res.reserve(padb(np * 4) * 2 + padb((size_t)n * 4))) != PL_OK) return rc;
    int *h_idx, *h_dist, *h_hits;
    h->res.out<int>(np, &h_idx);
    h->res.out<int>(np, &h_dist);
    h->res.out<int>((size_t)n, &h_hits);
    cudaStream_t st = h->stream;
    if (L.total_pts) {
        PL_CUDA_TRY(cudaMemcpyAsync(h_idx, d_idx, (size_t)L.total_pts * 4, cudaMemcpyDeviceToHost, st));
        PL_CUDA_TRY(cudaMemcpyAsync(h_dist, d_dist, (size_t)L.total_pts * 4, cudaMemcpyDeviceToHost, st));
    }
    PL_CUDA_TRY(cudaMemcpyAsync(h_hits, d_hits, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    for (int i = 0; i < n; i++) {
        if (pts[i].n) {
            memcpy(best_idx[i], h_idx + sd[i].pt_base, (size_t)pts[i].n * 4);
            if (best_dist && best_dist[i]) memcpy(best_dist[i], h_dist + sd[i].pt_base, (size_t)pts[i].n * 4);
        }
        n_fused[i] = h_hits[i];
    }
    return PL_OK;
}

PL_API int pl_orb_search_by_sim3(pl_match* h, const pl_frame_view* kf1, const pl_frame_view* kf2, const pl_posepoint_view* pts1,
                                 const pl_posepoint_view* pts2, const float t21[12], const float t12[12], float log_scale_factor1,
                                 float log_scale_factor2, float th, int* match12, int* n_found) {
    PL_CHECK_ARG(h && kf1 && kf2 && pts1 && pts2 && t21 && t12 && n_found && log_scale_factor1 > 0.f && log_scale_factor2 > 0.f);
    int rc;
    if ((rc = check_frame(kf1)) != PL_OK || (rc = check_frame(kf2)) != PL_OK) return rc;
    if ((rc = check_posepoints(*pts1, false)) != PL_OK || (rc = check_posepoints(*pts2, false)) != PL_OK) return rc;
    PL_CHECK_ARG(pts1->n == kf1->n && pts2->n == kf2->n && (match12 || kf1->n == 0));  // GetMapPointMatches(): one slot per feature
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    if ((rc = h->in.reserve(padb(sizeof(SearchDev) * 2) + frame_bytes(*kf1) + frame_bytes(*kf2) + posepoint_bytes(*pts1) + posepoint_bytes(*pts2))) != PL_OK)
        return rc;
    SearchDev sd[2];
    IndepLayout L;
    for (int d = 0; d < 2; d++) {
        SearchDev& S = sd[d];
        memset(&S, 0, sizeof(S));
        const pl_frame_view& from = d == 0 ? *kf1 : *kf2;  // owner of the points
        const pl_frame_view& into = d == 0 ? *kf2 : *kf1;  // searched key frame
        put_frame(h->in, into, S.F);
        put_posepoints(h->in, d == 0 ? *pts1 : *pts2, S, false);
        for (int k = 0; k < 12; k++) { S.F.tcw[k] = from.tcw[k]; S.t2[k] = (d == 0 ? t21 : t12)[k]; }
        S.F.fx = kf1->fx; S.F.fy = kf1->fy; S.F.cx = kf1->cx; S.F.cy = kf1->cy;  // :1445-1448: pKF1's intrinsics both ways
        S.use_t2 = 1;
        S.log_sf = d == 0 ? log_scale_factor2 : log_scale_factor1;
        place(S, into.n, L);
    }
    const SearchDev* d_sd = h->in.put(sd, 2);
    if ((rc = h->in.upload(h->stream)) != PL_OK) return rc;
    int *d_idx, *d_dist, *d_hits;
    if ((rc = run_point_best<2>(h, d_sd, 2, L, th, kThHigh, &d_idx, &d_dist, &d_hits)) != PL_OK) return rc;
    cudaStream_t st = h->stream;
    // d_dist (scratch 13, 8 bytes per point) is reused: the first half holds distances, the second half receives match12
    int* d_m12 = d_dist + std::max(L.total_pts, 1);
    int* d_found = d_hits + 2;
    if (kf1->n) {
        k_sim3_agree<<<(kf1->n + 255) / 256, 256, 0, st>>>(d_idx + sd[0].pt_base, kf1->n, d_idx + sd[1].pt_base, kf2->n, d_m12, d_found);
        h->last_launches++;
        PL_CUDA_TRY(cudaGetLastError());
    }
    if ((rc = h->res.reserve(padb((size_t)std::max(kf1->n, 1) * 4) + padb(16))) != PL_OK) return rc;
    int *h_m12, *h_found;
    h->res.out<int>((size_t)std::max(kf1->n, 1), &h_m12);
    h->res.out<int>(4, &h_found);
    if (kf1->n) PL_CUDA_TRY(cudaMemcpyAsync(h_m12, d_m12, (size_t)kf1->n * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(cudaMemcpyAsync(h_found, d_found, 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    if (kf1->n) memcpy(match12, h_m12, (size_t)kf1->n * 4);
    *n_found = h_found[0];
    return PL_OK;
}

PL_API int pl_orb_search_for_initialization(pl_match* h, const pl_frame_view* F1, const pl_frame_view* F2, float* prev_matched, int window_size,
                                            float nn_ratio, int check_orientation, int* matches12, int* n_matches) {
    PL_CHECK_ARG(h && F1 && F2 && n_matches && window_size > 0);
    int rc;
    if ((rc = check_frame(F1)) != PL_OK || (rc = check_frame(F2)) != PL_OK) return rc;
    PL_CHECK_ARG(F1->n == 0 || (prev_matched && matches12));
    if (F2->n > kInitMaxFeat) { set_error("SearchForInitialization: F2 has %d features, the limit is %d", F2->n, kInitMaxFeat); return PL_ERR_CAPACITY; }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    *n_matches = 0;
    if (F1->n == 0) return PL_OK;
    const size_t n1 = (size_t)F1->n;
    if ((rc = h->in.reserve(padb(sizeof(SearchDev)) + frame_bytes(*F2) + padb(n1 * sizeof(pl_keypoint)) + padb(n1 * 32) + padb(n1) + padb(n1 * 4) * 2 +
                            padb(n1 * 8))) != PL_OK)
        return rc;
    SearchDev S;
    memset(&S, 0, sizeof(S));
    put_frame(h->in, *F2, S.F);
    // the points are the level-0 features of F1 (:592-596), searched around vbPrevMatched
    std::vector<uint8_t> valid(n1);
    std::vector<float> px(n1), py(n1);
    for (size_t i = 0; i < n1; i++) {
        valid[i] = F1->keys_un[i].octave > 0 ? 0 : 1;
        px[i] = prev_matched[2 * i];
        py[i] = prev_matched[2 * i + 1];
    }
    S.np = F1->n;
    S.valid = h->in.put(valid.data(), n1);
    S.pdesc = (const uint4*)h->in.put(F1->desc, n1 * 32);
    S.proj_x = h->in.put(px.data(), n1);
    S.proj_y = h->in.put(py.data(), n1);
    const pl_keypoint* d_keys1 = h->in.put(F1->keys_un, n1);
    float* d_prev = const_cast<float*>(h->in.put(prev_matched, n1 * 2));
    IndepLayout L;
    place(S, F2->n, L);
    const SearchDev* d_sd = h->in.put(&S, 1);
    if ((rc = h->in.upload(h->stream)) != PL_OK) return rc;
    cudaStream_t st = h->stream;
    void* p;
    int *sorted_idx, *cell_start, *cand_n, *cand_off, *totals, *match, *rec, *out;
    if ((rc = match_scratch(h, 6, (size_t)S.n2 * 4, &p)) != PL_OK) return rc; sorted_idx = (int*)p;
    if ((rc = match_scratch(h, 7, (size_t)(kGridCells + 1) * 4, &p)) != PL_OK) return rc; cell_start = (int*)p;
    if ((rc = match_scratch(h, 8, n1 * 4, &p)) != PL_OK) return rc; cand_n = (int*)p;
    if ((rc = match_scratch(h, 9, n1 * 4, &p)) != PL_OK) return rc; cand_off = (int*)p;
    if ((rc = match_scratch(h, 10, 4, &p)) != PL_OK) return rc; totals = (int*)p;
    if ((rc = match_scratch(h, 12, n1 * 4, &p)) != PL_OK) return rc; match = (int*)p;
    if ((rc = match_scratch(h, 13, n1 * 8, &p)) != PL_OK) return rc; rec = (int*)p;
    if ((rc = match_scratch(h, 14, 8, &p)) != PL_OK) return rc; out = (int*)p;
    if ((size_t)S.n2 * 4 > 48 * 1024) PL_CUDA_TRY(cudaFuncSetAttribute(k_frame_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, S.n2 * 4));
    k_frame_grid<<<1, 1024, (size_t)S.n2 * 4, st>>>(d_sd, sorted_idx, cell_start);
    const dim3 cgrid((F1->n * 32 + 255) / 256, 1);
    const float thw = (float)window_size;
    k_candidates<4, false><<<cgrid, 256, 0, st>>>(d_sd, thw, sorted_idx, cell_start, cand_n, nullptr, nullptr, nullptr);
    k_cand_scan<<<1, 1024, 0, st>>>(d_sd, cand_n, cand_off, totals);
    h->last_launches += 3;
    int h_tot = 0;
    PL_CUDA_TRY(pl::stream_sync(st));  // a copy into pageable memory blocks inside the runtime until the stream gets there
    PL_CUDA_TRY(cudaMemcpyAsync(&h_tot, totals, 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    if ((rc = match_scratch(h, 15, (size_t)std::max(h_tot, 1) * 4, &p)) != PL_OK) return rc;
    unsigned int* cand = (unsigned int*)p;
    if ((rc = match_scratch(h, 11, 4, &p)) != PL_OK) return rc;
    int* inst_base = (int*)p;
    PL_CUDA_TRY(cudaMemsetAsync(inst_base, 0, 4, st));
    k_candidates<4, true><<<cgrid, 256, 0, st>>>(d_sd, thw, sorted_idx, cell_start, cand_n, cand_off, inst_base, cand);
    const size_t rsm = (size_t)kResChunkCand * 4 + (size_t)std::max(F2->n, 1) * 6 + 16;
    PL_CUDA_TRY(cudaFuncSetAttribute(k_init_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm));
    k_init_resolve<<<1, kResThreads, rsm, st>>>(d_sd, nn_ratio, check_orientation, cand, cand_n, cand_off, totals, d_keys1, match, rec, d_prev, out);
    h->last_launches += 2;
    PL_CUDA_TRY(cudaGetLastError());
    if ((rc = h->res.reserve(padb(n1 * 4) + padb(n1 * 8) + padb(8))) != PL_OK) return rc;
    int *h_match, *h_out;
    float* h_prev;
    h->res.out<int>(n1, &h_match);
    h->res.out<float>(n1 * 2, &h_prev);
    h->res.out<int>(2, &h_out);
    PL_CUDA_TRY(cudaMemcpyAsync(h_match, match, n1 * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(cudaMemcpyAsync(h_prev, d_prev, n1 * 8, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(cudaMemcpyAsync(h_out, out, 8, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    memcpy(matches12, h_match, n1 * 4);
    memcpy(prev_matched, h_prev, n1 * 8);
    *n_matches = h_out[0];
    if (h_out[1]) {
        set_error("a feature had more than %d candidates in its search window", kResChunkCand);
        return PL_ERR_CAPACITY;
    }
    return PL_OK;
}

}  // extern "C"

// ---- Frame::UndistortKeyLines / Frame::AssignFeaturesToGrid ----
namespace pl {
__global__ void __launch_bounds__(256) k_undistort_keylines(const pl_keyline* __restrict__ kls, int n, double fx, double fy, double cx, double cy,
                                                            double k0, double k1, double k2, double k3, double k4, int cols, int rows,
                                                            pl_keyline* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pl_keyline k = kls[i];
    const float2 s = undistort_point(k.sx, k.sy, fx, fy, cx, cy, k0, k1, k2, k3, k4);
    const float2 e = undistort_point(k.ex, k.ey, fx, fy, cx, cy, k0, k1, k2, k3, k4);
    const double nl[4] = {(double)s.x, (double)s.y, (double)e.x, (double)e.y};
    update_keyline(nl, k, cols, rows);  // the same field updates as LineMatcher::UpdateKeyLineData (Frame.cc:823-842 == LineMatcher.cpp:1601-1624)
    out[i] = k;
}
}  // namespace pl

extern "C" {

PL_API int pl_frame_undistort_keylines(pl_match* h, const pl_keyline* kls, int n, float fx, float fy, float cx, float cy, const float dist_coef[5],
                                       int img_cols, int img_rows, pl_keyline* out) {
    PL_CHECK_ARG(h && n >= 0 && dist_coef && (n == 0 || (kls && out)) && fx != 0.f && fy != 0.f && img_cols > 0 && img_rows > 0);
    if (n == 0) return PL_OK;
    if (dist_coef[0] == 0.0f) {  // Frame.cc:768-771: mvKeyLinesUn = mvKeyLines
        memmove(out, kls, (size_t)n * sizeof(pl_keyline));
        return PL_OK;
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int rc = h->in.reserve(padb((size_t)n * sizeof(pl_keyline)) * 2);
    if (rc != PL_OK) return rc;
    const pl_keyline* d_in = h->in.put(kls, (size_t)n);
    pl_keyline* h_out;
    pl_keyline* d_out = h->in.out<pl_keyline>((size_t)n, &h_out);
    cudaStream_t st = h->stream;
    if ((rc = h->in.upload(st)) != PL_OK) return rc;
    k_undistort_keylines<<<(n + 255) / 256, 256, 0, st>>>(d_in, n, (double)fx, (double)fy, (double)cx, (double)cy, (double)dist_coef[0],
                                                          (double)dist_coef[1], (double)dist_coef[2], (double)dist_coef[3], (double)dist_coef[4],
                                                          img_cols, img_rows, d_out);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(h_out, d_out, (size_t)n * sizeof(pl_keyline), cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    memcpy(out, h_out, (size_t)n * sizeof(pl_keyline));
    return PL_OK;
}

PL_API int pl_frame_assign_features_to_grid(pl_match* h, const pl_keypoint* keys_un, int n, const float bounds[4], int* cell_start, int* sorted_idx) {
    PL_CHECK_ARG(h && n >= 0 && n <= 65535 && bounds && cell_start && (n == 0 || (keys_un && sorted_idx)) && bounds[2] > bounds[0] && bounds[3] > bounds[1]);
    if (n == 0) {
        for (int c = 0; c <= kGridCells; c++) cell_start[c] = 0;
        return PL_OK;
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    int rc = h->in.reserve(padb(sizeof(SearchDev)) + padb((size_t)n * sizeof(pl_keypoint)) + padb((size_t)n2 * 4) + padb((size_t)(kGridCells + 1) * 4));
    if (rc != PL_OK) return rc;
    SearchDev S;
    memset(&S, 0, sizeof(S));
    S.F.n = n;
    S.F.keys = h->in.put(keys_un, (size_t)n);
    S.F.min_x = bounds[0]; S.F.min_y = bounds[1]; S.F.max_x = bounds[2]; S.F.max_y = bounds[3];
    S.F.inv_w = (float)kGridCols / (bounds[2] - bounds[0]);  // mfGridElementWidthInv (Frame.cc:184)
    S.F.inv_h = (float)kGridRows / (bounds[3] - bounds[1]);
    S.n2 = n2;
    int *h_sorted, *h_cst;
    int* d_sorted = h->in.out<int>((size_t)n2, &h_sorted);
    int* d_cst = h->in.out<int>((size_t)kGridCells + 1, &h_cst);
    const SearchDev* d_sd = h->in.put(&S, 1);
    cudaStream_t st = h->stream;
    if ((rc = h->in.upload(st)) != PL_OK) return rc;
    if ((size_t)n2 * 4 > 48 * 1024) PL_CUDA_TRY(cudaFuncSetAttribute(k_frame_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, n2 * 4));
    k_frame_grid<<<1, 1024, (size_t)n2 * 4, st>>>(d_sd, d_sorted, d_cst);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(h_sorted, d_sorted, (size_t)n2 * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(cudaMemcpyAsync(h_cst, d_cst, (size_t)(kGridCells + 1) * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    memcpy(cell_start, h_cst, (size_t)(kGridCells + 1) * 4);
    memcpy(sorted_idx, h_sorted, (size_t)h_cst[kGridCells] * 4);
    return PL_OK;
}

}  // extern "C"

// match_kernels.cu — descriptor search on sm_100a: the CUDA path behind pl_hamming_* / pl_orb_search_* /
// pl_line_match_* (include/plslam_c.h).
//
// Reference functions replaced:
//   ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2083-2103) == LineMatcher::DescriptorDistance
//     (src/LineMatcher.cpp:20-39)                                          -> hamming256()
//   cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) (LineMatcher.cpp:496-503,1179-1185)  -> k_knn2_partial + k_knn2_merge
//   candidate loops of SearchByProjection / SearchByBoW (ORBmatcher.cc:123-164, 306-336, 1802-1829)
//                                                                          -> k_hamming_candidates
#include <algorithm>
#include <vector>

#include "pl_common.cuh"

namespace pl {

// 256-bit Hamming distance of two 32-byte rows held as 2 x uint4
__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
           __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void __launch_bounds__(256) k_hamming_pairs(const uint4* __restrict__ a, const uint4* __restrict__ b, int n,
                                                       int* __restrict__ dist) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dist[i] = hamming256(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}

// ---------------------------------------------------------------------------------------------------------------
// brute-force kNN-2.  Packed key = dist << 20 | train index: the two smallest keys are the two nearest
// neighbours with ties resolved to the lowest train index (cv::BFMatcher behaviour).
// grid = (ceil(nq/128), n_chunks): every CTA scans one chunk of the train set through shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kKnnThreads = 128;
constexpr uint32_t kKeyNone = 0xFFFFFFFFu;

__global__ void __launch_bounds__(kKnnThreads) k_knn2_partial(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t,
                                                              int nt, int chunk, uint32_t* __restrict__ part, int n_chunks) {
    __shared__ uint4 s_t[kKnnThreads * 2];
    const int tid = threadIdx.x;
    const int qi = blockIdx.x * kKnnThreads + tid;
    const int c0 = blockIdx.y * chunk, c1 = min(nt, c0 + chunk);
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    if (qi < nq) { q0 = q[2 * qi]; q1 = q[2 * qi + 1]; }
    uint32_t b1 = kKeyNone, b2 = kKeyNone;
    for (int base = c0; base < c1; base += kKnnThreads) {
        const int n = min(kKnnThreads, c1 - base);
        __syncthreads();
        if (tid < n) {
            s_t[2 * tid] = t[2 * (size_t)(base + tid)];
            s_t[2 * tid + 1] = t[2 * (size_t)(base + tid) + 1];
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < n; j++) {
            const uint32_t key = ((uint32_t)hamming256(q0, q1, s_t[2 * j], s_t[2 * j + 1]) << 20) | (uint32_t)(base + j);
            const uint32_t lo = min(key, b1);
            b2 = min(b2, max(key, b1));
            b1 = lo;
        }
    }
    if (qi < nq) {
        part[((size_t)qi * n_chunks + blockIdx.y) * 2] = b1;
        part[((size_t)qi * n_chunks + blockIdx.y) * 2 + 1] = b2;
    }
}

__global__ void __launch_bounds__(256) k_knn2_merge(const uint32_t* __restrict__ part, int nq, int n_chunks, int* __restrict__ idx,
                                                    int* __restrict__ dist) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    uint32_t b1 = kKeyNone, b2 = kKeyNone;
    for (int c = 0; c < 2 * n_chunks; c++) {
        const uint32_t key = part[(size_t)qi * n_chunks * 2 + c];
        const uint32_t lo = min(key, b1);
        b2 = min(b2, max(key, b1));
        b1 = lo;
    }
    idx[2 * qi] = b1 == kKeyNone ? -1 : (int)(b1 & 0xFFFFFu);
    dist[2 * qi] = b1 == kKeyNone ? -1 : (int)(b1 >> 20);
    idx[2 * qi + 1] = b2 == kKeyNone ? -1 : (int)(b2 & 0xFFFFFu);
    dist[2 * qi + 1] = b2 == kKeyNone ? -1 : (int)(b2 >> 20);
}

// one warp per query row: distances to its CSR candidate list, in list order
__global__ void __launch_bounds__(256) k_hamming_candidates(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t,
                                                            const int* __restrict__ off, const int* __restrict__ cidx,
                                                            int* __restrict__ dist) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nq) return;
    const uint4 q0 = q[2 * w], q1 = q[2 * w + 1];
    const int b = off[w], e = off[w + 1];
    for (int k = b + lane; k < e; k += 32) {
        const int j = cidx[k];
        dist[k] = hamming256(q0, q1, t[2 * (size_t)j], t[2 * (size_t)j + 1]);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Projection searches.  The reference walks the candidates of every map point sequentially and lets later points see
// the features claimed by earlier ones (ORBmatcher.cc:128-130,177 / :1807-1809,1834).  Here the expensive part —
// window lookup in the 64x48 feature grid, the static gates and the 256-bit distances — runs warp-per-point in
// parallel and keeps the reference's candidate order; a single warp then replays the claims in point order.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGridCols = 64, kGridRows = 48, kGridCells = kGridCols * kGridRows;  // include/Frame.h
constexpr int kCandCap = 512;  // candidates kept per point (more -> PL_ERR_CAPACITY)
constexpr int kThHigh = 100, kHistoLen = 30;

struct FrameDev {
    int n;
    const pl_keypoint* keys;
    const uint4* desc;
    const float* u_right;
    const int* claimed;
    float min_x, min_y, max_x, max_y, fx, fy, cx, cy, bf, b;
    float tcw[12];
    float sf[kMaxLevels];
    float inv_w, inv_h;   // mfGridElementWidthInv / HeightInv (Frame.cc:184-185)
};

// Frame::AssignFeaturesToGrid (Frame.cc:265-287): features sorted by (cell, index); cell = posX*48 + posY with
// PosInGrid's round() (Frame.cc:527-538).  One CTA; bitonic sort of unique keys in shared memory.
__global__ void __launch_bounds__(1024) k_frame_grid(FrameDev F, int n2, int* __restrict__ sorted_idx, int* __restrict__ cell_start) {
    extern __shared__ unsigned int s_keys[];
    const int tid = threadIdx.x;
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
    for (int i = tid; i <= kGridCells; i += blockDim.x) cell_start[i] = 0;
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
    // cell_start[c] = first position with cell >= c  (positions of invalid keys are at the end)
    for (int i = tid; i < n2; i += blockDim.x) {
        const unsigned key = s_keys[i];
        const int c = key == 0xFFFFFFFFu ? kGridCells : (int)(key >> 16);
        const int prev = i == 0 ? -1 : (s_keys[i - 1] == 0xFFFFFFFFu ? kGridCells : (int)(s_keys[i - 1] >> 16));
        if (key != 0xFFFFFFFFu) sorted_idx[i] = (int)(key & 0xFFFFu);
        for (int cc = prev + 1; cc <= c; cc++) cell_start[cc] = i;
    }
    if (tid == 0) {
        const unsigned last = s_keys[n2 - 1];
        const int c = last == 0xFFFFFFFFu ? kGridCells : (int)(last >> 16);
        for (int cc = c + 1; cc <= kGridCells; cc++) cell_start[cc] = n2;
    }
}

// Frame::GetFeaturesInArea (Frame.cc:432-485) + static gates + distances for one point; one warp.
// Appends (idx | dist<<16) in the reference's candidate order to cand[0..kCandCap); returns the count (may exceed cap).
__device__ int gather_candidates(const FrameDev& F, const int* __restrict__ sorted_idx, const int* __restrict__ cell_start, float x,
                                 float y, float r, int minLevel, int maxLevel, bool gate_right, float ur, float gate_r, uint4 q0, uint4 q1,
                                 unsigned int* __restrict__ cand) {
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
            int idx = 0, dist = 0;
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
                if (ok && gate_right) {
                    const float uR = F.u_right[idx];
                    if (uR > 0 && fabsf(__fsub_rn(ur, uR)) > gate_r) ok = false;
                }
                if (ok) dist = hamming256(q0, q1, F.desc[2 * (size_t)idx], F.desc[2 * (size_t)idx + 1]);
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int pos = count + __popc(m & ((1u << lane) - 1u));
                if (pos < kCandCap) cand[pos] = (unsigned)idx | ((unsigned)dist << 16);
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

struct LastDev {
    int n;
    const uint8_t* valid;
    const float* world_pos;
    const uint4* desc;
    const int* octave;
    const float* angle;
    const uint8_t* has_obs;
};

// C3 phase A: one warp per Last feature (ORBmatcher.cc:1746-1830 without the claim check)
__global__ void __launch_bounds__(256) k_last_frame_candidates(FrameDev C, LastDev L, float th, int forward, int backward,
                                                               const int* __restrict__ sorted_idx, const int* __restrict__ cell_start,
                                                               unsigned int* __restrict__ cand, int* __restrict__ cand_n) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= L.n) return;
    int count = 0;
    if (L.valid[i]) {
        const float X = L.world_pos[3 * (size_t)i], Y = L.world_pos[3 * (size_t)i + 1], Z = L.world_pos[3 * (size_t)i + 2];
        const float xc = mat_row(C.tcw, 0, X, Y, Z), yc = mat_row(C.tcw, 1, X, Y, Z), zc = mat_row(C.tcw, 2, X, Y, Z);
        const float invzc = (float)(1.0 / (double)zc);
        if (!(invzc < 0)) {
            const float u = __fadd_rn(__fmul_rn(__fmul_rn(C.fx, xc), invzc), C.cx);
            const float v = __fadd_rn(__fmul_rn(__fmul_rn(C.fy, yc), invzc), C.cy);
            if (!(u < C.min_x || u > C.max_x) && !(v < C.min_y || v > C.max_y)) {
                const int oct = L.octave[i];
                const float radius = __fmul_rn(th, C.sf[oct]);
                int minL, maxL;
                if (forward) { minL = oct; maxL = -1; }
                else if (backward) { minL = 0; maxL = oct; }
                else { minL = oct - 1; maxL = oct + 1; }
                const float ur = __fsub_rn(u, __fmul_rn(C.bf, invzc));
                count = gather_candidates(C, sorted_idx, cell_start, u, v, radius, minL, maxL, true, ur, radius, L.desc[2 * (size_t)i],
                                          L.desc[2 * (size_t)i + 1], cand + (size_t)i * kCandCap);
            }
        }
    }
    if (lane == 0) cand_n[i] = count;
}

// C3 phase B: claims replayed in order by one warp (ORBmatcher.cc:1802-1876)
__global__ void __launch_bounds__(32) k_last_frame_resolve(FrameDev C, LastDev L, int check_orientation, const unsigned int* __restrict__ cand,
                                                           const int* __restrict__ cand_n, int* __restrict__ match, int* __restrict__ rec,
                                                           int* __restrict__ out /* [0]=nmatches [1]=overflow */) {
    extern __shared__ uint8_t s_claimed[];
    __shared__ int s_hist[kHistoLen];
    const int lane = threadIdx.x;
    for (int i = lane; i < C.n; i += 32) {
        s_claimed[i] = C.claimed ? (C.claimed[i] != 0) : 0;
        match[i] = -1;
    }
    if (lane < kHistoLen) s_hist[lane] = 0;
    __syncwarp();
    int nmatches = 0, nrec = 0, overflow = 0;
    const float factor = kHistoLen / 360.0f;
    for (int i = 0; i < L.n; i++) {
        const int cnt = cand_n[i];
        if (cnt == 0) continue;
        if (cnt > kCandCap) { overflow = 1; continue; }
        unsigned best = 0xFFFFFFFFu;
        const unsigned int* cd = cand + (size_t)i * kCandCap;
        for (int base = 0; base < cnt; base += 32) {
            unsigned key = 0xFFFFFFFFu;
            if (base + lane < cnt) {
                const unsigned e = cd[base + lane];
                if (!s_claimed[e & 0xFFFFu]) key = ((e >> 16) << 16) | (unsigned)(base + lane);  // dist, then candidate position
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
            best = min(best, key);
        }
        if (best == 0xFFFFFFFFu) continue;
        const int bestDist = (int)(best >> 16);
        if (bestDist <= kThHigh) {
            const int bestIdx2 = (int)(cd[best & 0xFFFFu] & 0xFFFFu);
            if (lane == 0) {
                match[bestIdx2] = i;
                s_claimed[bestIdx2] = L.has_obs ? (L.has_obs[i] != 0) : 1;
            }
            nmatches++;
            if (check_orientation) {
                float rot = __fsub_rn(L.angle[i], C.keys[bestIdx2].angle);
                if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                int bin = (int)roundf(__fmul_rn(rot, factor));
                if (bin == kHistoLen) bin = 0;
                if (lane == 0) {
                    rec[2 * nrec] = bestIdx2;
                    rec[2 * nrec + 1] = bin;
                    s_hist[bin]++;
                }
                nrec++;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    if (check_orientation) {
        // ComputeThreeMaxima (ORBmatcher.cc:2035-2077)
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < kHistoLen; i++) {
            const int s = s_hist[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
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
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nmatches -= removed;
    }
    if (lane == 0) { out[0] = nmatches; out[1] = overflow; }
}

struct MapPtsDev {
    int n;
    const uint4* desc;
    const uint8_t* in_view;
    const float *proj_x, *proj_y, *proj_xr, *view_cos;
    const int* level;
    const uint8_t* has_obs;
};

// C2 phase A: one warp per local map point (ORBmatcher.cc:84-149 without the claim check)
__global__ void __launch_bounds__(256) k_local_points_candidates(FrameDev F, MapPtsDev M, float th, const int* __restrict__ sorted_idx,
                                                                 const int* __restrict__ cell_start, unsigned int* __restrict__ cand,
                                                                 int* __restrict__ cand_n) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= M.n) return;
    int count = 0;
    if (M.in_view[i]) {
        const int lvl = M.level[i];
        float r = M.view_cos[i] > 0.998 ? 2.5f : 4.0f;  // RadiusByViewingCos (double literal compare as in the reference)
        if (th != 1.0f) r = __fmul_rn(r, th);
        const float rs = __fmul_rn(r, F.sf[lvl]);
        count = gather_candidates(F, sorted_idx, cell_start, M.proj_x[i], M.proj_y[i], rs, lvl - 1, lvl, true, M.proj_xr[i], rs,
                                  M.desc[2 * (size_t)i], M.desc[2 * (size_t)i + 1], cand + (size_t)i * kCandCap);
    }
    if (lane == 0) cand_n[i] = count;
}

// C2 phase B (ORBmatcher.cc:123-178): best / second best with their octaves, ratio test, claims in point order
__global__ void __launch_bounds__(32) k_local_points_resolve(FrameDev F, MapPtsDev M, float nn_ratio, const unsigned int* __restrict__ cand,
                                                             const int* __restrict__ cand_n, int* __restrict__ match, int* __restrict__ out) {
    extern __shared__ uint8_t s_claimed[];
    const int lane = threadIdx.x;
    for (int i = lane; i < F.n; i += 32) {
        s_claimed[i] = F.claimed ? (F.claimed[i] != 0) : 0;
        match[i] = -1;
    }
    __syncwarp();
    int nmatches = 0, overflow = 0;
    for (int i = 0; i < M.n; i++) {
        const int cnt = cand_n[i];
        if (cnt == 0) continue;
        if (cnt > kCandCap) { overflow = 1; continue; }
        const unsigned int* cd = cand + (size_t)i * kCandCap;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int base = 0; base < cnt; base += 32) {
            unsigned e = 0;
            int oct = 0;
            bool live = false;
            if (base + lane < cnt) {
                e = cd[base + lane];
                live = !s_claimed[e & 0xFFFFu];
                oct = F.keys[e & 0xFFFFu].octave;
            }
            unsigned m = __ballot_sync(0xffffffffu, live);
            while (m) {  // sequential two-minimum update, in candidate order
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const unsigned ej = __shfl_sync(0xffffffffu, e, j);
                const int oj = __shfl_sync(0xffffffffu, oct, j);
                const int dist = (int)(ej >> 16);
                if (dist < bestDist) {
                    bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = oj; bestIdx = (int)(ej & 0xFFFFu);
                } else if (dist < bestDist2) {
                    bestLevel2 = oj; bestDist2 = dist;
                }
            }
        }
        if (bestDist <= kThHigh) {
            if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nn_ratio, (float)bestDist2)) continue;
            if (lane == 0) {
                match[bestIdx] = i;
                s_claimed[bestIdx] = M.has_obs ? (M.has_obs[i] != 0) : 1;
            }
            nmatches++;
            __syncwarp();
        }
    }
    if (lane == 0) { out[0] = nmatches; out[1] = overflow; }
}


// ---------------------------------------------------------------------------------------------------------------
// LineMatcher: projection of 3-D map lines + all-pairs LineMatching
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

struct LineProjParams {
    float tcw[12];
    float fx, fy, cx, cy, min_x, min_y, max_x, max_y;
    int cols, rows;
};

// front half of LineMatcher::SearchByProjection (LineMatcher.cpp:96-212): one CTA, ordered compaction
__global__ void __launch_bounds__(256) k_line_project(LineProjParams P, const double* __restrict__ s3, const double* __restrict__ e3,
                                                      const pl_keyline* __restrict__ src, const uint8_t* __restrict__ valid, int n,
                                                      pl_keyline* __restrict__ out_kl, int* __restrict__ out_index, int* __restrict__ n_out) {
    __shared__ int s_warp[9];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 256) {
        const int i = base + tid;
        bool ok = false;
        pl_keyline k;
        if (i < n && valid[i]) {
            double T[12];
#pragma unroll
            for (int q = 0; q < 12; q++) T[q] = (double)P.tcw[q];
            const double* Xs = s3 + 3 * (size_t)i;
            const double* Xe = e3 + 3 * (size_t)i;
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
                k = src[i];
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
    if (tid == 0) *n_out = s_base;
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
// every hit counts; relaxed retry when fewer than 20 % of the current lines matched.  One CTA.
__global__ void __launch_bounds__(256) k_line_match_pairs(const pl_keyline* __restrict__ proj, const uint4* __restrict__ proj_desc, int n_proj,
                                                          const pl_keyline* __restrict__ cur, const uint4* __restrict__ cur_desc,
                                                          const uint8_t* __restrict__ cur_claimed, int n_cur, int* __restrict__ match,
                                                          int* __restrict__ out /* [0]=count [1]=used_relaxed */) {
    __shared__ int s_cnt;
    const int tid = threadIdx.x;
    const double zero[5] = {0, 0, 0, 0, 0}, relaxed[5] = {10.0, -0.1, -0.1, 5, 10};
    for (int pass = 0; pass < 2; pass++) {
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        const double* off = pass == 0 ? zero : relaxed;
        int local = 0;
        for (int j = tid; j < n_cur; j += 256) {
            int m = -1;
            if (!(pass == 0 && cur_claimed && cur_claimed[j])) {
                const pl_keyline kj = cur[j];
                const uint4 d0 = cur_desc[2 * (size_t)j], d1 = cur_desc[2 * (size_t)j + 1];
                for (int i = 0; i < n_proj; i++) {
                    const int hd = hamming256(proj_desc[2 * (size_t)i], proj_desc[2 * (size_t)i + 1], d0, d1);
                    if (line_matching(proj[i], kj, hd, off)) { m = i; local++; }
                }
            }
            match[j] = m;
        }
        if (local) atomicAdd(&s_cnt, local);
        __syncthreads();
        const int cnt = s_cnt;
        if (pass == 0 && !((double)cnt * 1.0 / (double)n_cur < 0.2)) {
            if (tid == 0) { out[0] = cnt; out[1] = 0; }
            return;
        }
        if (pass == 1 && tid == 0) { out[0] = cnt; out[1] = 1; }
        __syncthreads();
    }
}

}  // namespace pl

using namespace pl;

struct pl_match {
    int device = 0;
    cudaStream_t stream = nullptr;
    int last_launches = 0;
    // growable device scratch
    uint8_t* d_buf[24] = {nullptr};
    size_t d_cap[24] = {0};
    int sm_count = 148;
};

namespace {
int scratch(pl_match* h, int slot, size_t bytes, void** out) {
    if (h->d_cap[slot] < bytes) {
        if (h->d_buf[slot]) cudaFree(h->d_buf[slot]);
        h->d_buf[slot] = nullptr;
        h->d_cap[slot] = 0;
        size_t want = std::max(bytes, (size_t)1 << 16);
        PL_CUDA_TRY(cudaMalloc((void**)&h->d_buf[slot], want));
        h->d_cap[slot] = want;
    }
    *out = h->d_buf[slot];
    return PL_OK;
}

int knn2_launch(pl_match* h, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int* d_idx, int* d_dist) {
    // enough CTAs for >= 2 waves when the problem allows it; chunks are multiples of the 128-row smem tile
    const int qblocks = (nq + kKnnThreads - 1) / kKnnThreads;
    int want_chunks = std::max(1, (2 * h->sm_count + qblocks - 1) / qblocks);
    int max_chunks = std::max(1, (nt + kKnnThreads - 1) / kKnnThreads);
    int n_chunks = std::min(want_chunks, max_chunks);
    int chunk = ((nt + n_chunks - 1) / n_chunks + kKnnThreads - 1) / kKnnThreads * kKnnThreads;
    if (chunk < kKnnThreads) chunk = kKnnThreads;
    n_chunks = std::max(1, (nt + chunk - 1) / chunk);
    void* part = nullptr;
    int rc = scratch(h, 0, (size_t)nq * n_chunks * 2 * sizeof(uint32_t), &part);
    if (rc != PL_OK) return rc;
    k_knn2_partial<<<dim3(qblocks, n_chunks), kKnnThreads, 0, h->stream>>>((const uint4*)d_q, nq, (const uint4*)d_t, nt, chunk,
                                                                            (uint32_t*)part, n_chunks);
    k_knn2_merge<<<(nq + 255) / 256, 256, 0, h->stream>>>((const uint32_t*)part, nq, n_chunks, d_idx, d_dist);
    h->last_launches += 2;
    PL_CUDA_TRY(cudaGetLastError());
    return PL_OK;
}
}  // namespace

extern "C" {

PL_API int pl_match_create(pl_match** out, int device) {
    PL_CHECK_ARG(out != nullptr);
    *out = nullptr;
    int ndev = 0;
    PL_CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) {
        set_error("device %d not available (%d CUDA devices); this library has no CPU fallback", device, ndev);
        return PL_ERR_CUDA;
    }
    PL_CUDA_TRY(cudaSetDevice(device));
    pl_match* h = new pl_match();
    h->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_error("pl_match_create: %s", cudaGetErrorString(e));
        delete h;
        return PL_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = h;
    return PL_OK;
}

PL_API void pl_match_destroy(pl_match* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int i = 0; i < 24; i++)
        if (h->d_buf[i]) cudaFree(h->d_buf[i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

PL_API int pl_match_sync(pl_match* h) {
    PL_CHECK_ARG(h);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    PL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PL_OK;
}
PL_API void* pl_match_stream(pl_match* h) { return h ? (void*)h->stream : nullptr; }
PL_API int pl_match_last_launches(const pl_match* h) { return h ? h->last_launches : 0; }

PL_API int pl_hamming_pairs(pl_match* h, const uint8_t* a, const uint8_t* b, int n, int* dist) {
    PL_CHECK_ARG(h && a && b && dist && n >= 0);
    if (n == 0) return PL_OK;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    void *da, *db, *dd;
    int rc;
    if ((rc = scratch(h, 1, (size_t)n * 32, &da)) != PL_OK) return rc;
    if ((rc = scratch(h, 2, (size_t)n * 32, &db)) != PL_OK) return rc;
    if ((rc = scratch(h, 3, (size_t)n * 4, &dd)) != PL_OK) return rc;
    PL_CUDA_TRY(cudaMemcpyAsync(da, a, (size_t)n * 32, cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(db, b, (size_t)n * 32, cudaMemcpyHostToDevice, h->stream));
    k_hamming_pairs<<<(n + 255) / 256, 256, 0, h->stream>>>((const uint4*)da, (const uint4*)db, n, (int*)dd);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(dist, dd, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PL_OK;
}

PL_API int pl_hamming_knn2_dev(pl_match* h, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int* d_idx, int* d_dist) {
    PL_CHECK_ARG(h && d_q && d_idx && d_dist && nq >= 0 && nt >= 0 && nt < (1 << 20));
    PL_CHECK_ARG(((uintptr_t)d_q & 15) == 0 && ((uintptr_t)d_t & 15) == 0);
    if (nq == 0) return PL_OK;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    return knn2_launch(h, d_q, nq, d_t, nt, d_idx, d_dist);
}

PL_API int pl_hamming_knn2(pl_match* h, const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx, int* dist) {
    PL_CHECK_ARG(h && q && idx && dist && nq >= 0 && nt >= 0 && nt < (1 << 20) && (t || nt == 0));
    if (nq == 0) return PL_OK;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    void *dq, *dt, *di, *dd;
    int rc;
    if ((rc = scratch(h, 1, (size_t)nq * 32, &dq)) != PL_OK) return rc;
    if ((rc = scratch(h, 2, (size_t)std::max(nt, 1) * 32, &dt)) != PL_OK) return rc;
    if ((rc = scratch(h, 3, (size_t)nq * 8, &di)) != PL_OK) return rc;
    if ((rc = scratch(h, 4, (size_t)nq * 8, &dd)) != PL_OK) return rc;
    PL_CUDA_TRY(cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, h->stream));
    if (nt) PL_CUDA_TRY(cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, h->stream));
    rc = knn2_launch(h, (const uint8_t*)dq, nq, (const uint8_t*)dt, nt, (int*)di, (int*)dd);
    if (rc != PL_OK) return rc;
    PL_CUDA_TRY(cudaMemcpyAsync(idx, di, (size_t)nq * 8, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(dist, dd, (size_t)nq * 8, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PL_OK;
}

PL_API int pl_hamming_candidates(pl_match* h, const uint8_t* q, int nq, const uint8_t* t, int nt, const int* cand_off,
                                 const int* cand_idx, int* dist_out) {
    PL_CHECK_ARG(h && q && t && cand_off && nq >= 0 && nt > 0);
    if (nq == 0) return PL_OK;
    const int total = cand_off[nq];
    PL_CHECK_ARG(total >= 0 && (total == 0 || (cand_idx && dist_out)));
    if (total == 0) return PL_OK;
    for (int i = 0; i < total; i++) PL_CHECK_ARG(cand_idx[i] >= 0 && cand_idx[i] < nt);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    void *dq, *dt, *doff, *dci, *dd;
    int rc;
    if ((rc = scratch(h, 1, (size_t)nq * 32, &dq)) != PL_OK) return rc;
    if ((rc = scratch(h, 2, (size_t)nt * 32, &dt)) != PL_OK) return rc;
    if ((rc = scratch(h, 3, (size_t)(nq + 1) * 4, &doff)) != PL_OK) return rc;
    if ((rc = scratch(h, 4, (size_t)total * 4, &dci)) != PL_OK) return rc;
    if ((rc = scratch(h, 5, (size_t)total * 4, &dd)) != PL_OK) return rc;
    PL_CUDA_TRY(cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(doff, cand_off, (size_t)(nq + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(dci, cand_idx, (size_t)total * 4, cudaMemcpyHostToDevice, h->stream));
    k_hamming_candidates<<<(nq * 32 + 255) / 256, 256, 0, h->stream>>>((const uint4*)dq, nq, (const uint4*)dt, (const int*)doff,
                                                                       (const int*)dci, (int*)dd);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(dist_out, dd, (size_t)total * 4, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PL_OK;
}

}  // extern "C"

// ---- projection searches ----
namespace {
template <typename T>
int upload(pl_match* h, int slot, const T* src, size_t n, const T** dst) {
    void* p = nullptr;
    int rc = scratch(h, slot, std::max<size_t>(n, 1) * sizeof(T), &p);
    if (rc != PL_OK) return rc;
    if (n && src) PL_CUDA_TRY(cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    *dst = (const T*)p;
    return PL_OK;
}

// slots 6..11: frame arrays; returns the device view and builds the grid (slots 12,13)
int upload_frame(pl_match* h, const pl_frame_view* F, FrameDev* D, const int** sorted_idx, const int** cell_start) {
    PL_CHECK_ARG(F && F->n >= 0 && F->n <= 65535 && F->keys_un && F->desc && F->u_right && F->scale_factors);
    PL_CHECK_ARG(F->n_levels >= 1 && F->n_levels <= kMaxLevels && F->max_x > F->min_x && F->max_y > F->min_y);
    int rc;
    D->n = F->n;
    if ((rc = upload(h, 6, F->keys_un, (size_t)F->n, &D->keys)) != PL_OK) return rc;
    const uint8_t* dd;
    if ((rc = upload(h, 7, F->desc, (size_t)F->n * 32, &dd)) != PL_OK) return rc;
    D->desc = (const uint4*)dd;
    if ((rc = upload(h, 8, F->u_right, (size_t)F->n, &D->u_right)) != PL_OK) return rc;
    D->claimed = nullptr;
    if (F->claimed && (rc = upload(h, 9, F->claimed, (size_t)F->n, &D->claimed)) != PL_OK) return rc;
    D->min_x = F->min_x; D->min_y = F->min_y; D->max_x = F->max_x; D->max_y = F->max_y;
    D->fx = F->fx; D->fy = F->fy; D->cx = F->cx; D->cy = F->cy; D->bf = F->bf; D->b = F->b;
    for (int i = 0; i < 12; i++) D->tcw[i] = F->tcw[i];
    for (int i = 0; i < kMaxLevels; i++) D->sf[i] = i < F->n_levels ? F->scale_factors[i] : 0.f;
    D->inv_w = (float)kGridCols / (F->max_x - F->min_x);
    D->inv_h = (float)kGridRows / (F->max_y - F->min_y);
    int n2 = 1;
    while (n2 < std::max(F->n, 1)) n2 <<= 1;
    void *si, *cs;
    if ((rc = scratch(h, 12, (size_t)n2 * 4, &si)) != PL_OK) return rc;
    if ((rc = scratch(h, 13, (size_t)(kGridCells + 1) * 4, &cs)) != PL_OK) return rc;
    if ((size_t)n2 * 4 > 48 * 1024) PL_CUDA_TRY(cudaFuncSetAttribute(k_frame_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, n2 * 4));
    k_frame_grid<<<1, 1024, (size_t)n2 * 4, h->stream>>>(*D, n2, (int*)si, (int*)cs);
    h->last_launches++;
    *sorted_idx = (const int*)si;
    *cell_start = (const int*)cs;
    return PL_OK;
}
}  // namespace

extern "C" {

PL_API int pl_orb_search_last_frame(pl_match* h, const pl_frame_view* Cur, const pl_lastframe_view* Last, float th, int mono,
                                    int check_orientation, int* match_of_feature, int* n_matches) {
    PL_CHECK_ARG(h && Cur && Last && match_of_feature && n_matches && Last->n >= 0);
    PL_CHECK_ARG(Last->n == 0 || (Last->valid && Last->world_pos && Last->desc && Last->octave && Last->angle));
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    FrameDev C;
    const int *sorted_idx, *cell_start;
    int rc = upload_frame(h, Cur, &C, &sorted_idx, &cell_start);
    if (rc != PL_OK) return rc;
    LastDev L;
    L.n = Last->n;
    const uint8_t* dd;
    if ((rc = upload(h, 14, Last->valid, (size_t)L.n, &L.valid)) != PL_OK) return rc;
    if ((rc = upload(h, 15, Last->world_pos, (size_t)L.n * 3, &L.world_pos)) != PL_OK) return rc;
    if ((rc = upload(h, 16, Last->desc, (size_t)L.n * 32, &dd)) != PL_OK) return rc;
    L.desc = (const uint4*)dd;
    if ((rc = upload(h, 17, Last->octave, (size_t)L.n, &L.octave)) != PL_OK) return rc;
    if ((rc = upload(h, 18, Last->angle, (size_t)L.n, &L.angle)) != PL_OK) return rc;
    L.has_obs = nullptr;
    if (Last->has_observations && (rc = upload(h, 19, Last->has_observations, (size_t)L.n, &L.has_obs)) != PL_OK) return rc;
    for (int i = 0; i < L.n; i++) PL_CHECK_ARG(!Last->valid[i] || (Last->octave[i] >= 0 && Last->octave[i] < Cur->n_levels));
    // forward / backward decision (ORBmatcher.cc:1727-1741): twc = -Rcw^T tcw ; tlc = Rlw twc + tlw (cv::Mat float gemm)
    float twc[3], tlc[3];
    for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double)Cur->tcw[4 * k + r] * (double)Cur->tcw[4 * k + 3];
        twc[r] = (float)(s * -1.0);
    }
    for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double)Last->tcw[4 * r + k] * (double)twc[k];
        tlc[r] = (float)(s * 1.0 + (double)Last->tcw[4 * r + 3] * 1.0);
    }
    const int forward = (tlc[2] > Cur->b && !mono) ? 1 : 0, backward = (-tlc[2] > Cur->b && !mono) ? 1 : 0;
    void *cand, *cand_n, *match, *rec, *out;
    if ((rc = scratch(h, 20, (size_t)std::max(L.n, 1) * kCandCap * 4, &cand)) != PL_OK) return rc;
    if ((rc = scratch(h, 21, (size_t)std::max(L.n, 1) * 4, &cand_n)) != PL_OK) return rc;
    if ((rc = scratch(h, 22, (size_t)std::max(C.n, 1) * 4, &match)) != PL_OK) return rc;
    if ((rc = scratch(h, 23, (size_t)std::max(L.n, 1) * 8 + 16, &rec)) != PL_OK) return rc;
    if ((rc = scratch(h, 5, 16, &out)) != PL_OK) return rc;
    if (L.n)
        k_last_frame_candidates<<<(L.n * 32 + 255) / 256, 256, 0, h->stream>>>(C, L, th, forward, backward, sorted_idx, cell_start,
                                                                                (unsigned int*)cand, (int*)cand_n);
    k_last_frame_resolve<<<1, 32, (size_t)std::max(C.n, 1), h->stream>>>(C, L, check_orientation, (const unsigned int*)cand, (const int*)cand_n,
                                                                        (int*)match, (int*)rec, (int*)out);
    h->last_launches += 2;
    PL_CUDA_TRY(cudaGetLastError());
    int res[2] = {0, 0};
    PL_CUDA_TRY(cudaMemcpyAsync(match_of_feature, match, (size_t)C.n * 4, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(res, out, 8, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    *n_matches = res[0];
    if (res[1]) {
        set_error("a map point had more than %d candidate features in its search window", kCandCap);
        return PL_ERR_CAPACITY;
    }
    return PL_OK;
}

PL_API int pl_orb_search_local_points(pl_match* h, const pl_frame_view* F, const pl_mappoint_view* mps, float th, float nn_ratio,
                                      int* match_of_feature, int* n_matches) {
    PL_CHECK_ARG(h && F && mps && match_of_feature && n_matches && mps->n >= 0);
    PL_CHECK_ARG(mps->n == 0 || (mps->desc && mps->track_in_view && mps->proj_x && mps->proj_y && mps->proj_xr && mps->scale_level && mps->view_cos));
    for (int i = 0; i < mps->n; i++) PL_CHECK_ARG(!mps->track_in_view[i] || (mps->scale_level[i] >= 0 && mps->scale_level[i] < F->n_levels));
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    FrameDev D;
    const int *sorted_idx, *cell_start;
    int rc = upload_frame(h, F, &D, &sorted_idx, &cell_start);
    if (rc != PL_OK) return rc;
    MapPtsDev M;
    M.n = mps->n;
    const uint8_t* dd;
    if ((rc = upload(h, 14, mps->desc, (size_t)M.n * 32, &dd)) != PL_OK) return rc;
    M.desc = (const uint4*)dd;
    if ((rc = upload(h, 15, mps->track_in_view, (size_t)M.n, &M.in_view)) != PL_OK) return rc;
    if ((rc = upload(h, 16, mps->proj_x, (size_t)M.n, &M.proj_x)) != PL_OK) return rc;
    if ((rc = upload(h, 17, mps->proj_y, (size_t)M.n, &M.proj_y)) != PL_OK) return rc;
    if ((rc = upload(h, 18, mps->proj_xr, (size_t)M.n, &M.proj_xr)) != PL_OK) return rc;
    if ((rc = upload(h, 19, mps->view_cos, (size_t)M.n, &M.view_cos)) != PL_OK) return rc;
    if ((rc = upload(h, 10, mps->scale_level, (size_t)M.n, &M.level)) != PL_OK) return rc;
    M.has_obs = nullptr;
    if (mps->has_observations && (rc = upload(h, 11, mps->has_observations, (size_t)M.n, &M.has_obs)) != PL_OK) return rc;
    void *cand, *cand_n, *match, *out;
    if ((rc = scratch(h, 20, (size_t)std::max(M.n, 1) * kCandCap * 4, &cand)) != PL_OK) return rc;
    if ((rc = scratch(h, 21, (size_t)std::max(M.n, 1) * 4, &cand_n)) != PL_OK) return rc;
    if ((rc = scratch(h, 22, (size_t)std::max(D.n, 1) * 4, &match)) != PL_OK) return rc;
    if ((rc = scratch(h, 5, 16, &out)) != PL_OK) return rc;
    if (M.n)
        k_local_points_candidates<<<(M.n * 32 + 255) / 256, 256, 0, h->stream>>>(D, M, th, sorted_idx, cell_start, (unsigned int*)cand,
                                                                                  (int*)cand_n);
    k_local_points_resolve<<<1, 32, (size_t)std::max(D.n, 1), h->stream>>>(D, M, nn_ratio, (const unsigned int*)cand, (const int*)cand_n,
                                                                          (int*)match, (int*)out);
    h->last_launches += 2;
    PL_CUDA_TRY(cudaGetLastError());
    int res[2] = {0, 0};
    PL_CUDA_TRY(cudaMemcpyAsync(match_of_feature, match, (size_t)D.n * 4, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(res, out, 8, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    *n_matches = res[0];
    if (res[1]) {
        set_error("a map point had more than %d candidate features in its search window", kCandCap);
        return PL_ERR_CAPACITY;
    }
    return PL_OK;
}

// ---- line matching ----
PL_API int pl_line_project(pl_match* h, const double* start3d, const double* end3d, const pl_keyline* src_kl, const uint8_t* valid, int n,
                           const float tcw[12], float fx, float fy, float cx, float cy, float min_x, float min_y, float max_x, float max_y,
                           int img_cols, int img_rows, pl_keyline* out_kl, int* out_index, int* n_out) {
    PL_CHECK_ARG(h && tcw && out_kl && out_index && n_out && n >= 0);
    *n_out = 0;
    if (n == 0) return PL_OK;
    PL_CHECK_ARG(start3d && end3d && src_kl && valid);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int rc;
    const double *ds, *de;
    const pl_keyline* dk;
    const uint8_t* dv;
    if ((rc = upload(h, 6, start3d, (size_t)n * 3, &ds)) != PL_OK) return rc;
    if ((rc = upload(h, 7, end3d, (size_t)n * 3, &de)) != PL_OK) return rc;
    if ((rc = upload(h, 8, src_kl, (size_t)n, &dk)) != PL_OK) return rc;
    if ((rc = upload(h, 9, valid, (size_t)n, &dv)) != PL_OK) return rc;
    void *ok, *oi, *on;
    if ((rc = scratch(h, 10, (size_t)n * sizeof(pl_keyline), &ok)) != PL_OK) return rc;
    if ((rc = scratch(h, 11, (size_t)n * 4, &oi)) != PL_OK) return rc;
    if ((rc = scratch(h, 5, 16, &on)) != PL_OK) return rc;
    LineProjParams P;
    for (int i = 0; i < 12; i++) P.tcw[i] = tcw[i];
    P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy; P.min_x = min_x; P.min_y = min_y; P.max_x = max_x; P.max_y = max_y;
    P.cols = img_cols; P.rows = img_rows;
    k_line_project<<<1, 256, 0, h->stream>>>(P, ds, de, dk, dv, n, (pl_keyline*)ok, (int*)oi, (int*)on);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    int m = 0;
    PL_CUDA_TRY(cudaMemcpyAsync(&m, on, 4, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    *n_out = m;
    if (m) {
        PL_CUDA_TRY(cudaMemcpyAsync(out_kl, ok, (size_t)m * sizeof(pl_keyline), cudaMemcpyDeviceToHost, h->stream));
        PL_CUDA_TRY(cudaMemcpyAsync(out_index, oi, (size_t)m * 4, cudaMemcpyDeviceToHost, h->stream));
        PL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return PL_OK;
}

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
    int rc;
    const pl_keyline *dp, *dc;
    const uint8_t *dpd, *dcd, *dcl = nullptr;
    if ((rc = upload(h, 6, proj, (size_t)n_proj, &dp)) != PL_OK) return rc;
    if ((rc = upload(h, 7, proj_desc, (size_t)n_proj * 32, &dpd)) != PL_OK) return rc;
    if ((rc = upload(h, 8, cur, (size_t)n_cur, &dc)) != PL_OK) return rc;
    if ((rc = upload(h, 9, cur_desc, (size_t)n_cur * 32, &dcd)) != PL_OK) return rc;
    if (cur_claimed && (rc = upload(h, 10, cur_claimed, (size_t)n_cur, &dcl)) != PL_OK) return rc;
    void *dm, *dout;
    if ((rc = scratch(h, 11, (size_t)n_cur * 4, &dm)) != PL_OK) return rc;
    if ((rc = scratch(h, 5, 16, &dout)) != PL_OK) return rc;
    k_line_match_pairs<<<1, 256, 0, h->stream>>>(dp, (const uint4*)dpd, n_proj, dc, (const uint4*)dcd, dcl, n_cur, (int*)dm, (int*)dout);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    int res[2] = {0, 0};
    PL_CUDA_TRY(cudaMemcpyAsync(match_of_line, dm, (size_t)n_cur * 4, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(res, dout, 8, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(cudaStreamSynchronize(h->stream));
    *n_matches = res[0];
    *used_relaxed = res[1];
    return PL_OK;
}

}  // extern "C"

// bow_kernels.cu — ORBmatcher::SearchByBoW (C6, C7) and the brute-force line matchers of LineMatcher (D6) on sm_100a:
// the CUDA path behind pl_orb_search_bow_batch, pl_line_match_knn_ratio, pl_line_search_for_triangulation and
// pl_line_fuse_candidates (include/plslam_c.h).
//
// Reference functions replaced:
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)       src/ORBmatcher.cc:247-410   (C6)
//   ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)    src/ORBmatcher.cc:729-872   (C7)
//   LineMatcher::SearchByProjection(Frame&, KeyFrame*, vector<MapLine*>&) src/LineMatcher.cpp:489-525
//   LineMatcher::SearchForTriangulation + KeyFrame::lineDescriptorMAD    src/LineMatcher.cpp:1174-1204, src/KeyFrame.cc:773-797
//   LineMatcher::Fuse (descriptor half of the active branch)            src/LineMatcher.cpp:1296-1330
//   ORBmatcher::SearchForTriangulation + CheckDistEpipolarLine          src/ORBmatcher.cc:884-1095, :205-232 (E4)
//   MapPoint / MapLine::ComputeDistinctiveDescriptors                   src/MapPoint.cc:256-321, src/MapLine.cpp:269-330 (E5)
//
// SearchByBoW: the merge-join of the two FeatureVectors (<= ~100 level-4 nodes per image) is done on the host while the
// inputs are packed; it yields, in the reference's processing order, one query per valid feature of side A inside a
// common node, with the features of side B in that node as its ordered candidates.  The 256-bit distances run warp per
// query; one warp per call then replays the best / second-best test and the "already matched" flags in query order.
#include "match_common.cuh"

namespace pl {

constexpr int kBowThLow = 50, kBowHisto = 30;  // ORBmatcher.cc:49-51

struct BowDev {  // one SearchByBoW call
    int nq, nA, nB, n_out;
    const int* q_idx;        // side-A feature of query k
    const int* q_off;        // nq + 1 offsets into cand
    unsigned int* cand;      // in: side-B feature index; after k_bow_dist: index | distance << 16
    const uint4 *descA, *descB;
    const float *angA, *angB;
    int* match;              // n_out
    int* rec;                // 2 * nq: (slot in match, histogram bin) per accepted match
    int* out;                // [0] = matches
    // groups of queries that share no feature with any other group (the common nodes of the two feature vectors when every
    // feature sits in one node): the ordered replay is per group; n_groups == 0: one group with every query
    const int* g_off;        // n_groups + 1 offsets into the queries
    int n_groups;
};

__global__ void __launch_bounds__(256) k_bow_dist(const BowDev* __restrict__ BD) {
    const BowDev& D = BD[blockIdx.y];
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= D.nq) return;
    const int a = D.q_idx[q];
    const uint4 a0 = D.descA[2 * (size_t)a], a1 = D.descA[2 * (size_t)a + 1];
    for (int k = D.q_off[q] + lane; k < D.q_off[q + 1]; k += 32) {
        const unsigned b = D.cand[k];
        const int dist = hamming256(a0, a1, D.descB[2 * (size_t)b], D.descB[2 * (size_t)b + 1]);
        D.cand[k] = b | ((unsigned)dist << 16);
    }
}

// mode 0: C6 (:296-347), mode 1: C7 (:770-836).  The reference's sequential two-minimum update is evaluated in closed
// form: with pb = first position of the minimum among the unmatched candidates, the second best is the smaller of the
// minimum before pb and the minimum after pb (the update displaces the running best when pb arrives).
constexpr int kBowWarps = 16;
__global__ void __launch_bounds__(32 * kBowWarps) k_bow_resolve(const BowDev* __restrict__ BD, int mode, float nn_ratio, int check_orientation) {
    extern __shared__ uint8_t s_matched[];  // side-B feature already matched in this call (vpMapPointMatches / vbMatched2)
    __shared__ int s_hist[kBowHisto];
    __shared__ int s_nrec, s_nmatches;
    const BowDev& D = BD[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < D.nB; i += blockDim.x) s_matched[i] = 0;
    for (int i = threadIdx.x; i < D.n_out; i += blockDim.x) D.match[i] = -1;
    if (threadIdx.x < kBowHisto) s_hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_nrec = s_nmatches = 0;
    __syncthreads();
    const float factor = kBowHisto / 360.0f;
    // The reference walks the queries in order and a feature of side B can be taken once: an ordered process — but only among the
    // queries of one node, because a feature belongs to one node.  One warp replays a group (node); the groups run side by side.
    const int ng = D.n_groups > 0 ? D.n_groups : 1;
    int nmatches = 0;
    for (int gi = warp; gi < ng; gi += kBowWarps) {
        const int q0 = D.n_groups > 0 ? D.g_off[gi] : 0, q1 = D.n_groups > 0 ? D.g_off[gi + 1] : D.nq;
        for (int q = q0; q < q1; q++) {
            const int o0 = D.q_off[q], c = D.q_off[q + 1] - o0;
            if (c == 0) continue;
            const unsigned int* cd = D.cand + o0;
            unsigned best = 0xFFFFFFFFu;
            for (int base = 0; base < c; base += 32) {
                unsigned key = 0xFFFFFFFFu;
                if (base + lane < c) {
                    const unsigned e = cd[base + lane];
                    if (!s_matched[e & 0xFFFFu]) key = ((e >> 16) << 16) | (unsigned)(base + lane);
                }
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, sft));
                best = min(best, key);
            }
            if (best == 0xFFFFFFFFu) continue;
            const int pb = (int)(best & 0xFFFFu), bestDist1 = (int)(best >> 16);
            if (mode == 0 ? bestDist1 > kBowThLow : bestDist1 >= kBowThLow) continue;
            unsigned second = 0xFFFFFFFFu;
            for (int base = 0; base < c; base += 32) {
                unsigned key = 0xFFFFFFFFu;
                const int pos = base + lane;
                if (pos < c && pos != pb) {
                    const unsigned e = cd[pos];
                    if (!s_matched[e & 0xFFFFu]) key = e >> 16;
                }
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, sft));
                second = min(second, key);
            }
            const int bestDist2 = second == 0xFFFFFFFFu ? 256 : (int)second;
            if (!((float)bestDist1 < __fmul_rn(nn_ratio, (float)bestDist2))) continue;
            const int idxB = (int)(cd[pb] & 0xFFFFu), idxA = D.q_idx[q];
            const int slot = mode == 0 ? idxB : idxA;
            if (lane == 0) {
                s_matched[idxB] = 1;
                D.match[slot] = mode == 0 ? idxA : idxB;
            }
            if (check_orientation) {
                float rot = __fsub_rn(D.angA[idxA], D.angB[idxB]);
                if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                int bin = (int)roundf(__fmul_rn(rot, factor));
                if (bin == kBowHisto) bin = 0;
                if (lane == 0) {
                    const int k = atomicAdd(&s_nrec, 1);  // (the records are a set: the histogram rule below does not look at their order)
                    D.rec[2 * k] = slot;
                    D.rec[2 * k + 1] = bin;
                    atomicAdd(&s_hist[bin], 1);
                }
            }
            nmatches++;
            __syncwarp();
        }
    }
    if (lane == 0 && nmatches) atomicAdd(&s_nmatches, nmatches);
    __syncthreads();
    if (warp != 0) return;
    nmatches = s_nmatches;
    if (check_orientation) {
        const int nrec = s_nrec;
        // ComputeThreeMaxima (ORBmatcher.cc:2035-2077)
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < kBowHisto; i++) {
            const int sz = s_hist[i];
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
        int removed = 0;
        for (int k = lane; k < nrec; k += 32) {
            const int bin = D.rec[2 * k + 1];
            if (bin != ind1 && bin != ind2 && bin != ind3) {
                D.match[D.rec[2 * k]] = -1;
                removed++;
            }
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, sft);
        nmatches -= removed;
    }
    if (lane == 0) D.out[0] = nmatches;
}

// ---- D6 ----
// one thread per query: best/second of k_knn2 -> the ratio rule of LineMatcher.cpp:504-514; the "last query wins"
// overwrite of vpMapLineMatches[trainIdx] is an atomicMax on the query index
__global__ void __launch_bounds__(256) k_knn_ratio(const int* __restrict__ idx, const int* __restrict__ dist, int nq, int* __restrict__ match,
                                                   int* __restrict__ n_matches) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const float best = (float)dist[2 * i], better = (float)dist[2 * i + 1];
    const float ratio = __fdiv_rn(best, better);
    if ((double)ratio < 0.75) {
        atomicMax(match + idx[2 * i], i);
        atomicAdd(n_matches, 1);
    }
}

// KeyFrame::lineDescriptorMAD (KeyFrame.cc:773-797) + the acceptance of SearchForTriangulation (LineMatcher.cpp:1190-1201).
// All distances are small integers held as floats, so the medians are found exactly by rank counting; one CTA.
__device__ float median_by_rank(const float* v, int n, float* s_red) {
    // the element of rank n/2 in sorted order: value x with #(v < x) <= n/2 < #(v <= x)
    const int k = n / 2;
    float result = 0.f;
    __shared__ float s_res;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float x = v[i];
        int less = 0, leq = 0;
        for (int j = 0; j < n; j++) {
            less += v[j] < x;
            leq += v[j] <= x;
        }
        if (less <= k && k < leq) s_res = x;  // every thread that qualifies writes the same value
    }
    (void)s_red;
    __syncthreads();
    result = s_res;
    __syncthreads();
    return result;
}
__global__ void __launch_bounds__(1024) k_triangulation_mad(const int* __restrict__ idx, const int* __restrict__ dist, int n, float* __restrict__ tmp,
                                                            int* __restrict__ pairs, int* __restrict__ n_matches, double* __restrict__ mads) {
    extern __shared__ float s_v[];  // n floats
    __shared__ int s_cnt[33];
    const int tid = threadIdx.x;
    // nn
    for (int i = tid; i < n; i += blockDim.x) s_v[i] = (float)dist[2 * i];
    __syncthreads();
    const double nn_median = (double)median_by_rank(s_v, n, tmp);
    for (int i = tid; i < n; i += blockDim.x) s_v[i] = fabsf((float)((double)(float)dist[2 * i] - nn_median));
    __syncthreads();
    const double nn_mad = 1.4826 * (double)median_by_rank(s_v, n, tmp);
    // nn12
    for (int i = tid; i < n; i += blockDim.x) s_v[i] = __fsub_rn((float)dist[2 * i + 1], (float)dist[2 * i]);
    __syncthreads();
    const double nn12_median = (double)median_by_rank(s_v, n, tmp);
    for (int i = tid; i < n; i += blockDim.x)
        s_v[i] = fabsf((float)((double)__fsub_rn((float)dist[2 * i + 1], (float)dist[2 * i]) - nn12_median));
    __syncthreads();
    const double nn12_mad = 1.4826 * (double)median_by_rank(s_v, n, tmp);
    const double th = nn12_mad * 0.1;
    if (tid == 0) { mads[0] = nn_mad; mads[1] = nn12_mad; }
    // accepted pairs in query order: block-wide ordered compaction
    int base_count = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        bool ok = false;
        if (i < n) ok = (double)__fsub_rn((float)dist[2 * i + 1], (float)dist[2 * i]) > th;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if ((tid & 31) == 0) s_cnt[tid >> 5] = __popc(m);
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) { const int c = s_cnt[w]; s_cnt[w] = acc; acc += c; }
            s_cnt[32] = acc;
        }
        __syncthreads();
        if (ok) {
            const int pos = base_count + s_cnt[tid >> 5] + __popc(m & ((1u << (tid & 31)) - 1u));
            pairs[2 * pos] = i;
            pairs[2 * pos + 1] = idx[2 * i];
        }
        base_count += s_cnt[32];
        __syncthreads();
    }
    if (tid == 0) *n_matches = base_count;
}

// Fuse: nearest key-frame line of every valid map line, then dist < 1.5 * min(100, dist) (LineMatcher.cpp:1300-1312)
__global__ void __launch_bounds__(256) k_fuse_rule(const int* __restrict__ idx, const int* __restrict__ dist, const uint8_t* __restrict__ valid, int n,
                                                   int* __restrict__ tdx, int* __restrict__ n_fused) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int t = -1;
    if (!valid || valid[i]) {
        const float d = (float)dist[2 * i];
        double min_dist = 100;
        if ((double)d < min_dist) min_dist = (double)d;
        if ((double)d < 1.5 * min_dist) t = idx[2 * i];
    }
    tdx[i] = t;
    if (t >= 0) atomicAdd(n_fused, 1);
}


// ---- E4: SearchForTriangulation ----
struct TriDev {
    BowDev B;                       // queries / candidates of the merge-join (q_idx = idx1, cand = idx2)
    const pl_keypoint *keysA, *keysB;
    const float *urA, *urB;
    float f12[9], ex, ey;
    float sf2[kMaxLevels], sigma2[kMaxLevels];
};

// distances + the gates that do not depend on earlier matches (ORBmatcher.cc:978-1013): TH_LOW, the epipole distance for
// monocular pairs, CheckDistEpipolarLine.  A candidate that fails is given the distance 511.
__global__ void __launch_bounds__(256) k_triang_dist(const TriDev* __restrict__ TD) {
    const TriDev& T = TD[0];
    const BowDev& D = T.B;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= D.nq) return;
    const int a = D.q_idx[q];
    const uint4 a0 = D.descA[2 * (size_t)a], a1 = D.descA[2 * (size_t)a + 1];
    const pl_keypoint kp1 = T.keysA[a];
    const bool stereo1 = T.urA[a] >= 0;
    // CheckDistEpipolarLine (:209-211): line l = x1' F12
    const float la = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, T.f12[0]), __fmul_rn(kp1.y, T.f12[3])), T.f12[6]);
    const float lb = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, T.f12[1]), __fmul_rn(kp1.y, T.f12[4])), T.f12[7]);
    const float lc = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, T.f12[2]), __fmul_rn(kp1.y, T.f12[5])), T.f12[8]);
    const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
    for (int k = D.q_off[q] + lane; k < D.q_off[q + 1]; k += 32) {
        const unsigned b = D.cand[k];
        int dist = hamming256(a0, a1, D.descB[2 * (size_t)b], D.descB[2 * (size_t)b + 1]);
        if (dist <= kBowThLow) {
            const pl_keypoint kp2 = T.keysB[b];
            bool ok = true;
            if (!stereo1 && !(T.urB[b] >= 0)) {
                const float distex = __fsub_rn(T.ex, kp2.x), distey = __fsub_rn(T.ey, kp2.y);
                if (__fadd_rn(__fmul_rn(distex, distex), __fmul_rn(distey, distey)) < __fmul_rn(100.f, T.sf2[kp2.octave])) ok = false;
            }
            if (ok) {
                const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, kp2.x), __fmul_rn(lb, kp2.y)), lc);
                if (den == 0) ok = false;
                else {
                    const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
                    ok = (double)dsqr < __dmul_rn(3.84, (double)T.sigma2[kp2.octave]);
                }
            }
            if (!ok) dist = 511;
        } else {
            dist = 511;
        }
        D.cand[k] = b | ((unsigned)dist << 16);
    }
}

// the ordered part (:956-1046): per query the LAST candidate of minimum distance among those that passed the gates and are
// not matched yet (`dist > bestDist` skips, so an equal distance replaces); vbMatched2 in query order; rotation histogram
__global__ void __launch_bounds__(32) k_triang_resolve(const TriDev* __restrict__ TD, int check_orientation) {
    extern __shared__ uint8_t s_matched[];
    __shared__ int s_hist[kBowHisto];
    const BowDev& D = TD[0].B;
    const int lane = threadIdx.x;
    for (int i = lane; i < D.nB; i += 32) s_matched[i] = 0;
    for (int i = lane; i < D.n_out; i += 32) D.match[i] = -1;
    if (lane < kBowHisto) s_hist[lane] = 0;
    __syncwarp();
    const float factor = kBowHisto / 360.0f;
    int nmatches = 0, nrec = 0;
    for (int q = 0; q < D.nq; q++) {
        const int o0 = D.q_off[q], c = D.q_off[q + 1] - o0;
        if (c == 0) continue;
        const unsigned int* cd = D.cand + o0;
        unsigned best = 0xFFFFFFFFu;
        for (int base = 0; base < c; base += 32) {
            unsigned key = 0xFFFFFFFFu;
            const int pos = base + lane;
            if (pos < c) {
                const unsigned e = cd[pos];
                if ((e >> 16) <= (unsigned)kBowThLow && !s_matched[e & 0xFFFFu]) key = ((e >> 16) << 16) | (unsigned)(0xFFFF - pos);
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, sft));
            best = min(best, key);
        }
        if (best == 0xFFFFFFFFu) continue;
        const int pb = 0xFFFF - (int)(best & 0xFFFFu);
        const int idxB = (int)(cd[pb] & 0xFFFFu), idxA = D.q_idx[q];
        if (lane == 0) {
            s_matched[idxB] = 1;
            D.match[idxA] = idxB;
        }
        if (check_orientation) {
            float rot = __fsub_rn(D.angA[idxA], D.angB[idxB]);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, factor));
            if (bin == kBowHisto) bin = 0;
            if (lane == 0) {
                D.rec[2 * nrec] = idxA;
                D.rec[2 * nrec + 1] = bin;
                s_hist[bin]++;
            }
            nrec++;
        }
        nmatches++;
        __syncwarp();
    }
    if (check_orientation) {
        __syncwarp();
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < kBowHisto; i++) {
            const int sz = s_hist[i];
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
        int removed = 0;
        for (int k = lane; k < nrec; k += 32) {
            const int bin = D.rec[2 * k + 1];
            if (bin != ind1 && bin != ind2 && bin != ind3) {
                D.match[D.rec[2 * k]] = -1;
                removed++;
            }
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, sft);
        nmatches -= removed;
    }
    if (lane == 0) D.out[0] = nmatches;
}

// ---- E5: ComputeDistinctiveDescriptors ----
// one CTA per map point / map line; a warp per row: the N distances of the row go into a 257-bin histogram, the median
// (sorted[int(0.5*(N-1))]) is read off its prefix sums; the row with the least median wins, first one on ties
constexpr int kDistWarps = 4, kDistBins = 288;  // 257 bins rounded up to 9 per lane
__global__ void __launch_bounds__(kDistWarps * 32) k_distinctive(const uint4* __restrict__ desc, const int* __restrict__ group_off,
                                                                 int* __restrict__ best_row) {
    __shared__ int s_hist[kDistWarps][kDistBins];
    __shared__ unsigned s_best[kDistWarps];
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int o = group_off[g], N = group_off[g + 1] - o;
    const int kth = (int)(0.5 * (double)(N - 1));
    unsigned best = 0xFFFFFFFFu;
    for (int i = warp; i < N; i += kDistWarps) {
        for (int b = lane; b < kDistBins; b += 32) s_hist[warp][b] = 0;
        __syncwarp();
        const uint4 a0 = desc[2 * (size_t)(o + i)], a1 = desc[2 * (size_t)(o + i) + 1];
        for (int j = lane; j < N; j += 32) {
            const int d = j == i ? 0 : hamming256(a0, a1, desc[2 * (size_t)(o + j)], desc[2 * (size_t)(o + j) + 1]);
            atomicAdd(&s_hist[warp][d], 1);
        }
        __syncwarp();
        // lane owns bins [9*lane, 9*lane+9)
        int local = 0;
#pragma unroll
        for (int b = 0; b < 9; b++) local += s_hist[warp][9 * lane + b];
        int incl = local;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, sft);
            if (lane >= sft) incl += t;
        }
        const int excl = incl - local;
        int median = -1;
        if (kth >= excl && kth < incl) {
            int acc = excl;
#pragma unroll
            for (int b = 0; b < 9; b++) {
                acc += s_hist[warp][9 * lane + b];
                if (median < 0 && kth < acc) median = 9 * lane + b;
            }
        }
        const unsigned who = __ballot_sync(0xffffffffu, median >= 0);
        median = __shfl_sync(0xffffffffu, median, __ffs(who) - 1);
        best = min(best, ((unsigned)median << 16) | (unsigned)i);
        __syncwarp();
    }
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned b = s_best[0];
        for (int w = 1; w < kDistWarps; w++) b = min(b, s_best[w]);
        best_row[g] = N > 0 ? (int)(b & 0xFFFFu) : -1;
    }
}

}  // namespace pl

using namespace pl;

// from match_kernels.cu
int pl_knn2_launch_dev(pl_match* h, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int* d_idx, int* d_dist);

namespace {
inline size_t padb(size_t b) { return PlStage::pad(b); }
}

extern "C" {

PL_API int pl_orb_search_bow_batch(pl_match* h, int n, const pl_bow_view* a, const pl_bow_view* b, int mode, float nn_ratio, int check_orientation,
                                   int* const* match_out, int* n_matches) {
    PL_CHECK_ARG(h && n >= 0 && (mode == 0 || mode == 1) && (n == 0 || (a && b && match_out && n_matches)));
    if (n == 0) return PL_OK;
    // merge-join of the feature vectors (ORBmatcher.cc:276-381 / :758-851): queries and candidates in processing order
    std::vector<std::vector<int>> q_idx(n), q_off(n), g_off(n);
    std::vector<std::vector<unsigned>> cand(n);
    std::vector<uint8_t> seenA, seenB;
    size_t bytes = padb(sizeof(BowDev) * (size_t)n);
    int max_nq = 0, max_nB = 0;
    for (int i = 0; i < n; i++) {
        const pl_bow_view &A = a[i], &B = b[i];
        PL_CHECK_ARG(A.n >= 0 && B.n >= 0 && A.n <= 65535 && B.n <= 65535 && A.n_nodes >= 0 && B.n_nodes >= 0);
        PL_CHECK_ARG((A.n == 0 || (A.desc && A.angle)) && (B.n == 0 || (B.desc && B.angle)));
        PL_CHECK_ARG((A.n_nodes == 0 || (A.node_id && A.node_off && A.feat_idx)) && (B.n_nodes == 0 || (B.node_id && B.node_off && B.feat_idx)));
        PL_CHECK_ARG(match_out[i] != nullptr || (mode == 0 ? B.n : A.n) == 0);
        for (int k = 0; k + 1 < A.n_nodes; k++) PL_CHECK_ARG(A.node_id[k] < A.node_id[k + 1]);
        for (int k = 0; k + 1 < B.n_nodes; k++) PL_CHECK_ARG(B.node_id[k] < B.node_id[k + 1]);
        q_off[i].push_back(0);
        g_off[i].push_back(0);
        // the nodes can be replayed side by side only if no feature of either side sits in two of the joined nodes (true for a
        // DBoW2 feature vector; the views are caller data, so it is checked)
        bool disjoint = true;
        seenA.assign((size_t)A.n, 0);
        seenB.assign((size_t)B.n, 0);
        int ka = 0, kb = 0;
        while (ka < A.n_nodes && kb < B.n_nodes) {
            if (A.node_id[ka] == B.node_id[kb]) {
                for (int pb = B.node_off[kb]; pb < B.node_off[kb + 1]; pb++) {
                    const unsigned ib = B.feat_idx[pb];
                    PL_CHECK_ARG(ib < (unsigned)B.n);
                    if (seenB[ib]) disjoint = false;
                    seenB[ib] = 1;
                }
                for (int pa = A.node_off[ka]; pa < A.node_off[ka + 1]; pa++) {
                    const unsigned ia = A.feat_idx[pa];
                    PL_CHECK_ARG(ia < (unsigned)A.n);
                    if (seenA[ia]) disjoint = false;
                    seenA[ia] = 1;
                    if (A.valid && !A.valid[ia]) continue;
                    for (int pb = B.node_off[kb]; pb < B.node_off[kb + 1]; pb++) {
                        const unsigned ib = B.feat_idx[pb];
                        if (mode == 1 && B.valid && !B.valid[ib]) continue;  // !pMP2 || isBad (:797-802)
                        cand[i].push_back(ib);
                    }
                    q_idx[i].push_back((int)ia);
                    q_off[i].push_back((int)cand[i].size());
                }
                if ((int)q_idx[i].size() > g_off[i].back()) g_off[i].push_back((int)q_idx[i].size());
                ka++;
                kb++;
            } else if (A.node_id[ka] < B.node_id[kb]) {
                ka = (int)(std::lower_bound(A.node_id + ka, A.node_id + A.n_nodes, B.node_id[kb]) - A.node_id);
            } else {
                kb = (int)(std::lower_bound(B.node_id + kb, B.node_id + B.n_nodes, A.node_id[ka]) - B.node_id);
            }
        }
        if (!disjoint) g_off[i].assign(1, 0);  // one group with every query: the plain ordered replay
        const size_t nq = q_idx[i].size(), n_out = (size_t)(mode == 0 ? B.n : A.n);
        bytes += padb(g_off[i].size() * 4) + padb(nq * 4) + padb((nq + 1) * 4) + padb(cand[i].size() * 4) + padb((size_t)A.n * 32) + padb((size_t)B.n * 32) + padb((size_t)A.n * 4) +
                 padb((size_t)B.n * 4) + padb(n_out * 4) + padb(nq * 8 + 8) + padb(8);
        max_nq = std::max(max_nq, (int)nq);
        max_nB = std::max(max_nB, B.n);
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int rc = h->in.reserve(bytes);
    if (rc != PL_OK) return rc;
    std::vector<BowDev> bd(n);
    std::vector<int*> h_match(n), h_out(n);
    for (int i = 0; i < n; i++) {
        const pl_bow_view &A = a[i], &B = b[i];
        BowDev& D = bd[i];
        D.nq = (int)q_idx[i].size();
        D.nA = A.n;
        D.nB = B.n;
        D.n_out = mode == 0 ? B.n : A.n;
        D.q_idx = h->in.put(q_idx[i].data(), q_idx[i].size());
        D.q_off = h->in.put(q_off[i].data(), q_off[i].size());
        D.g_off = h->in.put(g_off[i].data(), g_off[i].size());
        D.n_groups = (int)g_off[i].size() - 1;
        D.cand = const_cast<unsigned int*>(h->in.put(cand[i].data(), cand[i].size()));
        D.descA = (const uint4*)h->in.put(A.desc, (size_t)A.n * 32);
        D.descB = (const uint4*)h->in.put(B.desc, (size_t)B.n * 32);
        D.angA = h->in.put(A.angle, (size_t)A.n);
        D.angB = h->in.put(B.angle, (size_t)B.n);
        D.match = h->in.out<int>((size_t)D.n_out, &h_match[i]);
        D.rec = h->in.out<int>((size_t)D.nq * 2 + 2);
        D.out = h->in.out<int>(2, &h_out[i]);
    }
    const BowDev* d_bd = h->in.put(bd.data(), (size_t)n);
    cudaStream_t st = h->stream;
    { int urc = h->in.upload(st); if (urc != PL_OK) return urc; }
    if (max_nq > 0) {
        k_bow_dist<<<dim3((max_nq * 32 + 255) / 256, n), 256, 0, st>>>(d_bd);
        h->last_launches++;
    }
    const size_t sm = (size_t)std::max(max_nB, 1);
    if (sm > 48 * 1024) PL_CUDA_TRY(cudaFuncSetAttribute(k_bow_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_bow_resolve<<<n, 32 * kBowWarps, sm, st>>>(d_bd, mode, nn_ratio, check_orientation);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    // results live in the same packed buffer: one copy back
    PL_CUDA_TRY(cudaMemcpyAsync(h->in.h, h->in.d, h->in.cur, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    for (int i = 0; i < n; i++) {
        if (bd[i].n_out) memcpy(match_out[i], h_match[i], (size_t)bd[i].n_out * 4);
        n_matches[i] = h_out[i][0];
    }
    return PL_OK;
}

// common front half of the D6 matchers: upload, knnMatch(k = 2) on the device
static int d6_knn(pl_match* h, const uint8_t* q, int nq, const uint8_t* t, int nt, int** d_idx, int** d_dist) {
    void *dq, *dt, *di, *dd;
    int rc;
    if ((rc = match_scratch(h, 1, (size_t)nq * 32, &dq)) != PL_OK) return rc;
    if ((rc = match_scratch(h, 2, (size_t)std::max(nt, 1) * 32, &dt)) != PL_OK) return rc;
    if ((rc = match_scratch(h, 3, (size_t)nq * 8, &di)) != PL_OK) return rc;
    if ((rc = match_scratch(h, 4, (size_t)nq * 8, &dd)) != PL_OK) return rc;
    PL_CUDA_TRY(cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, h->stream));
    rc = pl_knn2_launch_dev(h, (const uint8_t*)dq, nq, (const uint8_t*)dt, nt, (int*)di, (int*)dd);
    *d_idx = (int*)di;
    *d_dist = (int*)dd;
    return rc;
}

PL_API int pl_line_match_knn_ratio(pl_match* h, const uint8_t* ref_desc, int n_ref, const uint8_t* cur_desc, int n_cur, int* match_of_line,
                                   int* n_matches) {
    PL_CHECK_ARG(h && n_ref >= 0 && n_cur >= 0 && n_matches && (n_cur == 0 || (cur_desc && match_of_line)) && (n_ref == 0 || ref_desc));
    for (int j = 0; j < n_cur; j++) match_of_line[j] = -1;
    *n_matches = 0;
    if (n_ref == 0 || n_cur < 2) return PL_OK;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int *d_idx, *d_dist;
    int rc = d6_knn(h, ref_desc, n_ref, cur_desc, n_cur, &d_idx, &d_dist);
    if (rc != PL_OK) return rc;
    void* dm;
    if ((rc = match_scratch(h, 5, (size_t)(n_cur + 1) * 4, &dm)) != PL_OK) return rc;
    PL_CUDA_TRY(cudaMemsetAsync(dm, 0xff, (size_t)n_cur * 4, h->stream));
    PL_CUDA_TRY(cudaMemsetAsync((int*)dm + n_cur, 0, 4, h->stream));
    k_knn_ratio<<<(n_ref + 255) / 256, 256, 0, h->stream>>>(d_idx, d_dist, n_ref, (int*)dm, (int*)dm + n_cur);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    std::vector<int> tmp((size_t)n_cur + 1);
    PL_CUDA_TRY(cudaMemcpyAsync(tmp.data(), dm, (size_t)(n_cur + 1) * 4, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    memcpy(match_of_line, tmp.data(), (size_t)n_cur * 4);
    *n_matches = tmp[n_cur];
    return PL_OK;
}

PL_API int pl_line_search_for_triangulation(pl_match* h, const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, int* pairs, int* n_matches,
                                            double* nn_mad, double* nn12_mad) {
    PL_CHECK_ARG(h && n1 >= 0 && n2 >= 0 && n_matches && (n1 == 0 || (desc1 && pairs)) && (n2 == 0 || desc2) && n1 <= 12000);
    *n_matches = 0;
    if (nn_mad) *nn_mad = 0;
    if (nn12_mad) *nn12_mad = 0;
    if (n1 == 0 || n2 < 2) return PL_OK;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int *d_idx, *d_dist;
    int rc = d6_knn(h, desc1, n1, desc2, n2, &d_idx, &d_dist);
    if (rc != PL_OK) return rc;
    void *dp, *dmisc;
    if ((rc = match_scratch(h, 5, (size_t)n1 * 8, &dp)) != PL_OK) return rc;
    if ((rc = match_scratch(h, 6, 64, &dmisc)) != PL_OK) return rc;
    const size_t sm = (size_t)n1 * 4;
    if (sm > 48 * 1024) PL_CUDA_TRY(cudaFuncSetAttribute(k_triangulation_mad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_triangulation_mad<<<1, 1024, sm, h->stream>>>(d_idx, d_dist, n1, nullptr, (int*)dp, (int*)((uint8_t*)dmisc + 16), (double*)dmisc);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    uint8_t misc[24];
    PL_CUDA_TRY(cudaMemcpyAsync(misc, dmisc, 24, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(pairs, dp, (size_t)n1 * 8, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    double m[2];
    memcpy(m, misc, 16);
    memcpy(n_matches, misc + 16, 4);
    if (nn_mad) *nn_mad = m[0];
    if (nn12_mad) *nn12_mad = m[1];
    return PL_OK;
}

PL_API int pl_line_fuse_candidates(pl_match* h, const uint8_t* ml_desc, const uint8_t* valid, int n, const uint8_t* kf_desc, int n_kf, int* tdx,
                                   int* n_fused) {
    PL_CHECK_ARG(h && n >= 0 && n_kf >= 0 && n_fused && (n == 0 || (ml_desc && tdx)) && (n_kf == 0 || kf_desc));
    *n_fused = 0;
    for (int i = 0; i < n; i++) tdx[i] = -1;
    if (n == 0 || n_kf == 0) return PL_OK;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int *d_idx, *d_dist;
    int rc = d6_knn(h, ml_desc, n, kf_desc, n_kf, &d_idx, &d_dist);
    if (rc != PL_OK) return rc;
    void *dt, *dv = nullptr;
    if ((rc = match_scratch(h, 5, (size_t)(n + 1) * 4, &dt)) != PL_OK) return rc;
    if (valid) {
        if ((rc = match_scratch(h, 6, (size_t)n, &dv)) != PL_OK) return rc;
        PL_CUDA_TRY(cudaMemcpyAsync(dv, valid, (size_t)n, cudaMemcpyHostToDevice, h->stream));
    }
    PL_CUDA_TRY(cudaMemsetAsync((int*)dt + n, 0, 4, h->stream));
    k_fuse_rule<<<(n + 255) / 256, 256, 0, h->stream>>>(d_idx, d_dist, (const uint8_t*)dv, n, (int*)dt, (int*)dt + n);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    std::vector<int> tmp((size_t)n + 1);
    PL_CUDA_TRY(cudaMemcpyAsync(tmp.data(), dt, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    memcpy(tdx, tmp.data(), (size_t)n * 4);
    *n_fused = tmp[n];
    return PL_OK;
}

}  // extern "C"

extern "C" {

PL_API int pl_orb_search_for_triangulation(pl_match* h, const pl_triang_view* a, const pl_triang_view* b, const float f12[9], const float cw1[3],
                                           const float kf2_tcw[12], float fx2, float fy2, float cx2, float cy2, const float* scale_factors2,
                                           const float* level_sigma2_2, int n_levels2, int only_stereo, int check_orientation, int* pairs,
                                           int* n_matches) {
    PL_CHECK_ARG(h && a && b && f12 && cw1 && kf2_tcw && scale_factors2 && level_sigma2_2 && n_matches && n_levels2 >= 1 && n_levels2 <= kMaxLevels);
    const pl_bow_view &A = a->bow, &B = b->bow;
    PL_CHECK_ARG(A.n >= 0 && B.n >= 0 && A.n <= 65535 && B.n <= 65535 && A.n_nodes >= 0 && B.n_nodes >= 0);
    PL_CHECK_ARG((A.n == 0 || (A.desc && A.angle && a->keys_un && a->u_right && pairs)) && (B.n == 0 || (B.desc && B.angle && b->keys_un && b->u_right)));
    PL_CHECK_ARG((A.n_nodes == 0 || (A.node_id && A.node_off && A.feat_idx)) && (B.n_nodes == 0 || (B.node_id && B.node_off && B.feat_idx)));
    for (int k = 0; k + 1 < A.n_nodes; k++) PL_CHECK_ARG(A.node_id[k] < A.node_id[k + 1]);
    for (int k = 0; k + 1 < B.n_nodes; k++) PL_CHECK_ARG(B.node_id[k] < B.node_id[k + 1]);
    for (int i = 0; i < B.n; i++) PL_CHECK_ARG(b->keys_un[i].octave >= 0 && b->keys_un[i].octave < n_levels2);
    *n_matches = 0;
    // merge-join (:925-1061) with the gates that only depend on one side: pMP1 / pMP2, bOnlyStereo
    std::vector<int> q_idx, q_off(1, 0);
    std::vector<unsigned> cand;
    int ka = 0, kb = 0;
    while (ka < A.n_nodes && kb < B.n_nodes) {
        if (A.node_id[ka] == B.node_id[kb]) {
            for (int pa = A.node_off[ka]; pa < A.node_off[ka + 1]; pa++) {
                const unsigned ia = A.feat_idx[pa];
                PL_CHECK_ARG(ia < (unsigned)A.n);
                if (A.valid && !A.valid[ia]) continue;
                if (only_stereo && !(a->u_right[ia] >= 0)) continue;
                for (int pb = B.node_off[kb]; pb < B.node_off[kb + 1]; pb++) {
                    const unsigned ib = B.feat_idx[pb];
                    PL_CHECK_ARG(ib < (unsigned)B.n);
                    if (B.valid && !B.valid[ib]) continue;
                    if (only_stereo && !(b->u_right[ib] >= 0)) continue;
                    cand.push_back(ib);
                }
                q_idx.push_back((int)ia);
                q_off.push_back((int)cand.size());
            }
            ka++;
            kb++;
        } else if (A.node_id[ka] < B.node_id[kb]) {
            ka = (int)(std::lower_bound(A.node_id + ka, A.node_id + A.n_nodes, B.node_id[kb]) - A.node_id);
        } else {
            kb = (int)(std::lower_bound(B.node_id + kb, B.node_id + B.n_nodes, A.node_id[ka]) - B.node_id);
        }
    }
    const size_t nq = q_idx.size();
    if (A.n == 0) return PL_OK;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    const size_t bytes = padb(sizeof(TriDev)) + padb(nq * 4) + padb((nq + 1) * 4) + padb(cand.size() * 4) + padb((size_t)A.n * 32) + padb((size_t)B.n * 32) +
                         (padb((size_t)A.n * 4) + padb((size_t)B.n * 4)) * 2 + padb((size_t)A.n * sizeof(pl_keypoint)) +
                         padb((size_t)B.n * sizeof(pl_keypoint)) + padb((size_t)A.n * 4) + padb(nq * 8 + 8) + padb(8);
    int rc = h->in.reserve(bytes);
    if (rc != PL_OK) return rc;
    TriDev T;
    memset(&T, 0, sizeof(T));
    BowDev& D = T.B;
    int *h_match, *h_out;
    D.nq = (int)nq; D.nA = A.n; D.nB = B.n; D.n_out = A.n;
    D.q_idx = h->in.put(q_idx.data(), nq);
    D.q_off = h->in.put(q_off.data(), q_off.size());
    D.cand = const_cast<unsigned int*>(h->in.put(cand.data(), cand.size()));
    D.descA = (const uint4*)h->in.put(A.desc, (size_t)A.n * 32);
    D.descB = (const uint4*)h->in.put(B.desc, (size_t)B.n * 32);
    D.angA = h->in.put(A.angle, (size_t)A.n);
    D.angB = h->in.put(B.angle, (size_t)B.n);
    T.keysA = h->in.put(a->keys_un, (size_t)A.n);
    T.keysB = h->in.put(b->keys_un, (size_t)B.n);
    T.urA = h->in.put(a->u_right, (size_t)A.n);
    T.urB = h->in.put(b->u_right, (size_t)B.n);
    D.match = h->in.out<int>((size_t)A.n, &h_match);
    D.rec = h->in.out<int>(nq * 2 + 2);
    D.out = h->in.out<int>(2, &h_out);
    for (int k = 0; k < 9; k++) T.f12[k] = f12[k];
    for (int k = 0; k < n_levels2; k++) { T.sf2[k] = scale_factors2[k]; T.sigma2[k] = level_sigma2_2[k]; }
    // the epipole (:897-903): C2 = R2w*Cw + t2w as one cv::Mat gemm (double accumulation, rounded once); the float
    // expressions that follow are kept un-contracted (volatile: the host compiler must not fuse them)
    float C2[3];
    for (int r = 0; r < 3; r++) {
        double sacc = 0;
        for (int k = 0; k < 3; k++) sacc += (double)kf2_tcw[4 * r + k] * (double)cw1[k];
        C2[r] = (float)(sacc * 1.0 + (double)kf2_tcw[4 * r + 3] * 1.0);
    }
    {
        volatile float invz = 1.0f / C2[2];
        volatile float tx = fx2 * C2[0], ty = fy2 * C2[1];
        volatile float txz = tx * invz, tyz = ty * invz;
        T.ex = txz + cx2;
        T.ey = tyz + cy2;
    }
    const TriDev* d_t = h->in.put(&T, 1);
    cudaStream_t st = h->stream;
    if ((rc = h->in.upload(st)) != PL_OK) return rc;
    if (nq > 0) {
        k_triang_dist<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, st>>>(d_t);
        h->last_launches++;
    }
    const size_t sm = (size_t)std::max(B.n, 1);
    if (sm > 48 * 1024) PL_CUDA_TRY(cudaFuncSetAttribute(k_triang_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_triang_resolve<<<1, 32, sm, st>>>(d_t, check_orientation);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(h->in.h, h->in.d, h->in.cur, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    int np = 0;
    for (int i = 0; i < A.n; i++)
        if (h_match[i] >= 0) { pairs[2 * np] = i; pairs[2 * np + 1] = h_match[i]; np++; }
    *n_matches = h_out[0];
    return PL_OK;
}

PL_API int pl_distinctive_descriptors(pl_match* h, const uint8_t* desc, const int* group_off, int n_groups, int* best_row) {
    PL_CHECK_ARG(h && n_groups >= 0 && (n_groups == 0 || (group_off && best_row)));
    if (n_groups == 0) return PL_OK;
    PL_CHECK_ARG(group_off[0] == 0);
    for (int g = 0; g < n_groups; g++) PL_CHECK_ARG(group_off[g + 1] >= group_off[g] && group_off[g + 1] - group_off[g] <= 65535);
    const size_t rows = (size_t)group_off[n_groups];
    PL_CHECK_ARG(rows == 0 || desc);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int rc = h->in.reserve(padb(rows * 32) + padb((size_t)(n_groups + 1) * 4) + padb((size_t)n_groups * 4));
    if (rc != PL_OK) return rc;
    const uint4* d_desc = (const uint4*)h->in.put(desc, rows * 32);
    const int* d_off = h->in.put(group_off, (size_t)n_groups + 1);
    int* h_best;
    int* d_best = h->in.out<int>((size_t)n_groups, &h_best);
    cudaStream_t st = h->stream;
    if ((rc = h->in.upload(st)) != PL_OK) return rc;
    k_distinctive<<<n_groups, kDistWarps * 32, 0, st>>>(d_desc, d_off, d_best);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(h_best, d_best, (size_t)n_groups * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    memcpy(best_row, h_best, (size_t)n_groups * 4);
    return PL_OK;
}

}  // extern "C"

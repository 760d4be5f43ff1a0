// stereo_kernels.cu — Frame::ComputeStereoMatches (src/Frame.cc:888-1062) on sm_100a: the CUDA path
// behind pl_frame_compute_stereo_matches (include/plslam_c.h).
//
// Every left key point is independent until the final median rule, so one warp takes one left key point:
//   1. the row-band search (:897-973): the reference buckets the right key points by image row (vRowIndices); here every lane
//      tests one right key point against the band, the octave window and the disparity range, and the warp keeps the first
//      minimum of (distance, iR) — the order the bucket would have been walked in;
//   2. the SAD refinement (:976-1029): 11 x 11 window, 11 shifts, read straight from the two extractors' pyramid planes in HBM
//      (the reason mvImagePyramid is exposed, Frame.cc:985,1002); all sums are integers, so the float arithmetic of
//      cv::norm(IL, IR, NORM_L1) is exact; lane = pixel, shuffle reduction per shift;
//   3. the parabola fit and the disparity gate (:1032-1055) on lane 0 with separately rounded float operations.
// A second, single-CTA kernel applies the outlier rule (:1059-1071): median of the SAD scores by a shared-memory bitonic sort.
#include "match_common.cuh"

namespace pl {

struct StereoLevels {
    const uint8_t* l[kMaxLevels];
    const uint8_t* r[kMaxLevels];
    size_t lp[kMaxLevels], rp[kMaxLevels];
    int rcols[kMaxLevels];
    float sf[kMaxLevels], invsf[kMaxLevels];
    int n_levels, n_rows0;
};

__global__ void __launch_bounds__(256) k_stereo_match(StereoLevels S, const pl_keypoint* __restrict__ kl, const uint4* __restrict__ dl, int n,
                                                      const pl_keypoint* __restrict__ kr, const uint4* __restrict__ dr, int nr, float bf, float b,
                                                      float* __restrict__ u_right, float* __restrict__ depth, int* __restrict__ sad) {
    const int iL = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (iL >= n) return;
    const pl_keypoint kpL = kl[iL];
    float out_ur = -1.0f, out_d = -1.0f;
    int out_sad = -1;
    const float minD = 0.f, maxD = __fdiv_rn(bf, b);
    const float uL = kpL.x, vL = kpL.y;
    const int row = (int)vL;
    const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, minD);
    unsigned best = 0xFFFFFFFFu;
    if (vL >= 0.f && row < S.n_rows0 && !(maxU < 0)) {
        const uint4 q0 = dl[2 * (size_t)iL], q1 = dl[2 * (size_t)iL + 1];
        for (int base = 0; base < nr; base += 32) {
            const int iR = base + lane;
            unsigned key = 0xFFFFFFFFu;
            if (iR < nr) {
                const pl_keypoint kpR = kr[iR];
                const float r = __fmul_rn(2.0f, S.sf[kpR.octave]);
                const int maxr = (int)ceilf(__fadd_rn(kpR.y, r)), minr = (int)floorf(__fsub_rn(kpR.y, r));
                if (row >= minr && row <= maxr && !(kpR.octave < kpL.octave - 1 || kpR.octave > kpL.octave + 1) && kpR.x >= minU && kpR.x <= maxU) {
                    const int dist = hamming256(q0, q1, dr[2 * (size_t)iR], dr[2 * (size_t)iR + 1]);
                    if (dist < 100) key = ((unsigned)dist << 16) | (unsigned)iR;  // bestDist starts at TH_HIGH, strict < (:933, :962)
                }
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, sft));
            best = min(best, key);
        }
    }
    if (best != 0xFFFFFFFFu && (int)(best >> 16) < 75) {  // thOrbDist = (TH_HIGH + TH_LOW) / 2
        const float uR0 = kr[best & 0xFFFFu].x;
        const int oct = kpL.octave;
        const float scaleFactor = S.invsf[oct];
        const float scaleduL = roundf(__fmul_rn(kpL.x, scaleFactor));
        const float scaledvL = roundf(__fmul_rn(kpL.y, scaleFactor));
        const float scaleduR0 = roundf(__fmul_rn(uR0, scaleFactor));
        const int w = 5, L = 5;
        const float iniu = __fsub_rn(__fadd_rn(scaleduR0, (float)L), (float)w), endu = __fadd_rn(__fadd_rn(__fadd_rn(scaleduR0, (float)L), (float)w), 1.f);
        if (!(iniu < 0 || endu >= (float)S.rcols[oct])) {
            const int r0 = (int)__fsub_rn(scaledvL, (float)w), c0 = (int)__fsub_rn(scaleduL, (float)w);
            const uint8_t* pl_ = S.l[oct];
            const uint8_t* pr_ = S.r[oct];
            const size_t lp = S.lp[oct], rp = S.rp[oct];
            const int cL = pl_[(ptrdiff_t)(r0 + w) * (ptrdiff_t)lp + (c0 + w)];
            int cr0[11], cR[11], acc[11];
#pragma unroll
            for (int s = 0; s < 11; s++) {
                cr0[s] = (int)__fsub_rn(__fadd_rn(scaleduR0, (float)(s - L)), (float)w);
                cR[s] = pr_[(ptrdiff_t)(r0 + w) * (ptrdiff_t)rp + (cr0[s] + w)];
                acc[s] = 0;
            }
            for (int p = lane; p < 121; p += 32) {
                const int y = p / 11, x = p - 11 * y;
                const int a = (int)pl_[(ptrdiff_t)(r0 + y) * (ptrdiff_t)lp + (c0 + x)] - cL;
                const uint8_t* rrow = pr_ + (ptrdiff_t)(r0 + y) * (ptrdiff_t)rp;
#pragma unroll
                for (int s = 0; s < 11; s++) acc[s] += abs(a - ((int)rrow[cr0[s] + x] - cR[s]));
            }
#pragma unroll
            for (int s = 0; s < 11; s++)
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], sft);
            // first minimum over the shifts (dist < bestDist, :1010-1014)
            int bestDist = 0x7fffffff, bestinc = 0;
#pragma unroll
            for (int s = 0; s < 11; s++)
                if (acc[s] < bestDist) { bestDist = acc[s]; bestinc = s - L; }
            if (!(bestinc == -L || bestinc == L)) {
                float dist1 = 0, dist2 = 0, dist3 = 0;
#pragma unroll
                for (int s = 1; s < 10; s++)
                    if (s - L == bestinc) { dist1 = (float)acc[s - 1]; dist2 = (float)acc[s]; dist3 = (float)acc[s + 1]; }
                const float num = __fsub_rn(dist1, dist3);
                const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2)));
                const float deltaR = __fdiv_rn(num, den);
                if (!(deltaR < -1 || deltaR > 1)) {
                    float bestuR = __fmul_rn(S.sf[oct], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
                    float disparity = __fsub_rn(uL, bestuR);
                    if (disparity >= minD && disparity < maxD) {
                        if (disparity <= 0) {
                            disparity = 0.01f;
                            bestuR = (float)((double)uL - 0.01);
                        }
                        out_d = __fdiv_rn(bf, disparity);
                        out_ur = bestuR;
                        out_sad = bestDist;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        u_right[iL] = out_ur;
        depth[iL] = out_d;
        sad[iL] = out_sad;
    }
}

// :1059-1071 — median of the SAD scores (sorted by (score, iL)), then everything at or above 1.5 * 1.4 * median is dropped
__global__ void __launch_bounds__(1024) k_stereo_outliers(int n, int n2, const int* __restrict__ sad, float* __restrict__ u_right, float* __restrict__ depth) {
    extern __shared__ unsigned s_keys[];
    __shared__ int s_cnt;
    const int tid = threadIdx.x;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < n2; i += blockDim.x) {
        unsigned k = 0xFFFFFFFFu;
        if (i < n && sad[i] >= 0) { k = ((unsigned)sad[i] << 14) | (unsigned)i; local++; }  // score <= 121 * 510 < 2^16, i < 2^14
        s_keys[i] = k;
    }
    if (local) atomicAdd(&s_cnt, local);
    __syncthreads();
    for (int a = 2; a <= n2; a <<= 1)
        for (int j = a >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & a) == 0;
                    const unsigned x = s_keys[i], y = s_keys[ixj];
                    if (up ? (x > y) : (x < y)) { s_keys[i] = y; s_keys[ixj] = x; }
                }
            }
            __syncthreads();
        }
    const int cnt = s_cnt;
    if (cnt == 0) return;  // the reference indexes an empty vector here
    const float median = (float)(int)(s_keys[cnt / 2] >> 14);
    const float thDist = __fmul_rn(1.5f * 1.4f, median);
    for (int i = tid; i < n; i += blockDim.x) {
        const int s = sad[i];
        if (s >= 0 && !((float)s < thDist)) { u_right[i] = -1.f; depth[i] = -1.f; }
    }
}

}  // namespace pl

using namespace pl;

namespace {
inline size_t padb(size_t b) { return PlStage::pad(b); }
}

extern "C" {

PL_API int pl_frame_compute_stereo_matches(pl_match* h, pl_orb* left, pl_orb* right, int frame, const pl_keypoint* keys_left, const uint8_t* desc_left,
                                           int n_left, const pl_keypoint* keys_right, const uint8_t* desc_right, int n_right, float bf, float b,
                                           float* u_right, float* depth) {
    PL_CHECK_ARG(h && left && right && n_left >= 0 && n_right >= 0 && n_left <= 16384 && n_right <= 16384 && b > 0.f);
    PL_CHECK_ARG((n_left == 0 || (keys_left && desc_left && u_right && depth)) && (n_right == 0 || (keys_right && desc_right)));
    if (n_left == 0) return PL_OK;
    StereoLevels S;
    memset(&S, 0, sizeof(S));
    S.n_levels = pl_orb_levels(left);
    PL_CHECK_ARG(S.n_levels >= 1 && S.n_levels <= kMaxLevels && pl_orb_levels(right) == S.n_levels);
    int rc;
    if ((rc = pl_orb_scale_factors(left, S.sf)) != PL_OK || (rc = pl_orb_inv_scale_factors(left, S.invsf)) != PL_OK) return rc;
    for (int l = 0; l < S.n_levels; l++) {
        int rows = 0, cols = 0, rrows = 0;
        if ((rc = pl_orb_pyramid_dev(left, frame, l, &S.l[l], &S.lp[l], &rows, &cols)) != PL_OK) return rc;
        if (l == 0) S.n_rows0 = rows;
        if ((rc = pl_orb_pyramid_dev(right, frame, l, &S.r[l], &S.rp[l], &rrows, &S.rcols[l])) != PL_OK) return rc;
        PL_CHECK_ARG(rrows == rows && S.rcols[l] == cols);  // a rectified pair: the SAD windows index both pyramids with the same rows
    }
    int w0 = 0, h0 = 0;
    pl_orb_pyramid_dims(left, 0, &h0, &w0);
    for (int i = 0; i < n_left; i++)  // inside the image: the SAD window around the scaled position then stays inside the bordered plane
        PL_CHECK_ARG(keys_left[i].octave >= 0 && keys_left[i].octave < S.n_levels && keys_left[i].x >= 0.f && keys_left[i].x < (float)w0 &&
                     keys_left[i].y >= 0.f && keys_left[i].y < (float)h0);
    for (int i = 0; i < n_right; i++) PL_CHECK_ARG(keys_right[i].octave >= 0 && keys_right[i].octave < S.n_levels);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    // the pyramids must be complete.  pl_orb_sync reports (and clears) the sticky capacity word of an extractor: a frame that
    // overflowed emitted no key points, so the matches are computed for what the caller passed, but the caller is told
    bool overflowed = false;
    if ((rc = pl_orb_sync(left)) != PL_OK && rc != PL_ERR_CAPACITY) return rc;
    overflowed |= rc == PL_ERR_CAPACITY;
    if ((rc = pl_orb_sync(right)) != PL_OK && rc != PL_ERR_CAPACITY) return rc;
    overflowed |= rc == PL_ERR_CAPACITY;
    const size_t nl = (size_t)n_left, nrr = (size_t)n_right;
    if ((rc = h->in.reserve(padb(nl * sizeof(pl_keypoint)) + padb(nl * 32) + padb(nrr * sizeof(pl_keypoint)) + padb(nrr * 32) + padb(nl * 4) * 3)) != PL_OK)
        return rc;
    const pl_keypoint* d_kl = h->in.put(keys_left, nl);
    const uint4* d_dl = (const uint4*)h->in.put(desc_left, nl * 32);
    const pl_keypoint* d_kr = h->in.put(keys_right, nrr);
    const uint4* d_dr = (const uint4*)h->in.put(desc_right, nrr * 32);
    float *h_ur, *h_d;
    float* d_ur = h->in.out<float>(nl, &h_ur);
    float* d_d = h->in.out<float>(nl, &h_d);
    int* d_sad = h->in.out<int>(nl);
    cudaStream_t st = h->stream;
    if ((rc = h->in.upload(st)) != PL_OK) return rc;
    k_stereo_match<<<(unsigned)((nl * 32 + 255) / 256), 256, 0, st>>>(S, d_kl, d_dl, n_left, d_kr, d_dr, n_right, bf, b, d_ur, d_d, d_sad);
    int n2 = 1;
    while (n2 < n_left) n2 <<= 1;
    if ((size_t)n2 * 4 > 48 * 1024) PL_CUDA_TRY(cudaFuncSetAttribute(k_stereo_outliers, cudaFuncAttributeMaxDynamicSharedMemorySize, n2 * 4));
    k_stereo_outliers<<<1, 1024, (size_t)n2 * 4, st>>>(n_left, n2, d_sad, d_ur, d_d);
    h->last_launches += 2;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(h_ur, d_ur, nl * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(cudaMemcpyAsync(h_d, d_d, nl * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    memcpy(u_right, h_ur, nl * 4);
    memcpy(depth, h_d, nl * 4);
    if (overflowed) {
        pl::set_error("pl_frame_compute_stereo_matches: an extractor reported a frame over its key point capacity (outputs are filled)");
        return PL_ERR_CAPACITY;
    }
    return PL_OK;
}

}  // extern "C"

// orb_kernels.cu — ORB extraction on sm_100a: the CUDA path behind pl_orb_* (include/plslam_c.h).
//
// Replaces ORBextractor::operator() (reference src/ORBextractor.cc:1043-1105) and everything it calls:
//   ComputePyramid (:1107-1132)          -> k_pyr_base, k_pyr_resize          (chained per level)
//   per-cell cv::FAST + fallback (:765-829) -> k_fast_cells                   (one CTA per 30-px cell and frame)
//   DistributeOctTree (:539-763)         -> k_octree                          (one CTA per level and frame)
//   GaussianBlur 7x7 (:1085-1086)        -> k_blur7                           (smem-tiled separable fixed point)
//   IC_Angle (:77-104) + computeOrbDescriptor (:108-147) + pt*=scale (:1094-1101) -> k_orient_brief (warp per kp)
//
// Data layout (per handle, B = frames per chunk): all intermediates are level-major, frame-minor planes with
// 128-byte-aligned row pitch; a chunk is sized so that pyramid + blurred planes + candidate lists stay resident
// in the 126 MB L2 between the producing and the consuming kernel.
#include <math.h>
#include <stdarg.h>

#include <string.h>

#include <algorithm>
#include <vector>

#include "pl_common.cuh"

namespace pl {

static const int8_t h_pattern[1024] = {
#include "../../include/pl_brief_pattern.inc"
};
__constant__ int8_t c_pattern[1024];

// ---------------------------------------------------------------------------------------------------------------
// geometry tables (host-built, device-resident)
// ---------------------------------------------------------------------------------------------------------------
struct LevelGeom {
    int w, h;                // un-bordered level size (ORBextractor.cc:1112)
    int pitch, bpitch;       // row pitch of the bordered plane / blurred plane
    unsigned long long plane_off, plane_size, blur_off, blur_size;  // byte offsets; frame f at off + f*size*... see idx
    int cell_base, n_cells;  // range in the cell table
    int quota;               // mnFeaturesPerLevel[level]
    int out_cap, out_base;   // kept-keypoint slots per frame
    int cand_cap, cand_base; // candidate slots per frame
    int n_ini;               // DistributeOctTree roots (:543)
    float hx;                // root width (:545)
    int bw, bh;              // maxBorder-minBorder extents
    float scale, kp_size;    // mvScaleFactor[level], (float)(int)(31*scale)
    int xtab_off, ytab_off;  // resize tables (entries)
    int tile_base, n_tiles_x, n_tiles_y;  // blur tiles
};
struct OrbGeom {
    int nlevels, total_cells, total_tiles, cand_per_frame, out_per_frame, ini_th, min_th, node_cap, max_cells_level;
    int umax[16];
    LevelGeom lv[kMaxLevels];
};
struct Cell {
    short level, x0, y0, x1, y1;  // FAST window [x0,x1) x [y0,y1) in un-bordered level coordinates (:791-806)
    short pad;
};
struct ResizeTab {  // cv::resize fixed-point coefficients (INTER_RESIZE_COEF_BITS = 11)
    short s;        // source index
    short a0, a1;   // weights of s and s+1
    short pad;
};

// ---------------------------------------------------------------------------------------------------------------
// pyramid
// ---------------------------------------------------------------------------------------------------------------
// level 0: copyMakeBorder(image, temp, 19.., BORDER_REFLECT_101)  (:1127).  One thread = 4 output bytes.
__global__ void __launch_bounds__(256) k_pyr_base(const OrbGeom* __restrict__ g, const uint8_t* __restrict__ in,
                                                  size_t in_step, size_t in_frame_stride, uint8_t* __restrict__ pyr,
                                                  int frames_cap) {
    const LevelGeom& L = g->lv[0];
    const int bx = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int by = blockIdx.y;
    const int f = blockIdx.z;
    if (bx >= L.pitch) return;
    const uint8_t* src = in + (size_t)f * in_frame_stride + (size_t)reflect101(by - kEdge, L.h) * in_step;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int x = bx + i;
        uint32_t b = x < L.w + 2 * kEdge ? __ldg(src + reflect101(x - kEdge, L.w)) : 0u;
        v |= b << (8 * i);
    }
    uint8_t* dst = pyr + L.plane_off + (size_t)f * L.plane_size + (size_t)by * L.pitch + bx;
    *reinterpret_cast<uint32_t*>(dst) = v;
}

// level k>0: resize(pyr[k-1] -> pyr[k], INTER_LINEAR) then copyMakeBorder REFLECT_101 (:1120-1122), fused: every
// pixel of the bordered plane evaluates the bilinear sample at its reflected coordinate.
__global__ void __launch_bounds__(256) k_pyr_resize(const OrbGeom* __restrict__ g, int level,
                                                    const ResizeTab* __restrict__ tabs, uint8_t* __restrict__ pyr) {
    const LevelGeom& L = g->lv[level];
    const LevelGeom& P = g->lv[level - 1];
    const int bx = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int by = blockIdx.y;
    const int f = blockIdx.z;
    if (bx >= L.pitch) return;
    const ResizeTab ty = tabs[L.ytab_off + reflect101(by - kEdge, L.h)];
    const int sy0 = ty.s, sy1 = min(ty.s + 1, P.h - 1);
    const uint8_t* p0 = pyr + P.plane_off + (size_t)f * P.plane_size + (size_t)(sy0 + kEdge) * P.pitch + kEdge;
    const uint8_t* p1 = pyr + P.plane_off + (size_t)f * P.plane_size + (size_t)(sy1 + kEdge) * P.pitch + kEdge;
    const ResizeTab* tx = tabs + L.xtab_off;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int x = bx + i;
        uint32_t b = 0;
        if (x < L.w + 2 * kEdge) {
            const ResizeTab t = tx[reflect101(x - kEdge, L.w)];
            const int sx0 = t.s, sx1 = min(t.s + 1, P.w - 1);
            int S0 = p0[sx0] * t.a0 + p0[sx1] * t.a1;
            int S1 = p1[sx0] * t.a0 + p1[sx1] * t.a1;
            b = (uint32_t)(((((int)ty.a0 * (S0 >> 4)) >> 16) + (((int)ty.a1 * (S1 >> 4)) >> 16) + 2) >> 2) & 0xffu;
        }
        v |= b << (8 * i);
    }
    uint8_t* dst = pyr + L.plane_off + (size_t)f * L.plane_size + (size_t)by * L.pitch + bx;
    *reinterpret_cast<uint32_t*>(dst) = v;
}

// exact-2x special case of cv::resize (INTER_LINEAR with integer scale 2 runs the INTER_AREA fast path)
__global__ void __launch_bounds__(256) k_pyr_half(const OrbGeom* __restrict__ g, int level, uint8_t* __restrict__ pyr) {
    const LevelGeom& L = g->lv[level];
    const LevelGeom& P = g->lv[level - 1];
    const int bx = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int by = blockIdx.y;
    const int f = blockIdx.z;
    if (bx >= L.pitch) return;
    const int y = reflect101(by - kEdge, L.h);
    const uint8_t* p0 = pyr + P.plane_off + (size_t)f * P.plane_size + (size_t)(2 * y + kEdge) * P.pitch + kEdge;
    const uint8_t* p1 = p0 + P.pitch;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int xx = bx + i;
        uint32_t b = 0;
        if (xx < L.w + 2 * kEdge) {
            int x = reflect101(xx - kEdge, L.w);
            b = (uint32_t)((p0[2 * x] + p0[2 * x + 1] + p1[2 * x] + p1[2 * x + 1] + 2) >> 2);
        }
        v |= b << (8 * i);
    }
    uint8_t* dst = pyr + L.plane_off + (size_t)f * L.plane_size + (size_t)by * L.pitch + bx;
    *reinterpret_cast<uint32_t*>(dst) = v;
}

// ---------------------------------------------------------------------------------------------------------------
// FAST-9-16 per cell with threshold fallback
// ---------------------------------------------------------------------------------------------------------------
constexpr int kFastThreads = 128;
constexpr int kMaxWin = 66;   // wCell = ceil(width/floor(width/30)) < 60, window = wCell + 6
constexpr int kWinPitch = 68;
constexpr int kMaxInt = kMaxWin - 6;
constexpr int kMaxSurv = ((kMaxInt + 1) / 2) * ((kMaxInt + 1) / 2);
constexpr int kFastBitWords = 128;  // >= kMaxInt * kMaxInt / 32 (one bit per interior pixel of a cell), a multiple of 32
static_assert(kFastBitWords * 32 >= kMaxInt * kMaxInt && kFastBitWords % 32 == 0 && kWinPitch % 4 == 0, "FAST cell bitmap");

// Packed ring differences: for ring value r and centre constant A = (v+256) + ((256-v)<<16),
// r*0xFFFF + A = (v-r+256) | ((r-v+256)<<16): both 16-bit halves are biased by +256 and positive, so one
// VIMNMX3.S16x2 evaluates the "darker" (d) and the "brighter" (-d) arc tests at once.
__device__ __forceinline__ uint32_t ring_d2(const uint8_t* c, int off, uint32_t A) { return (uint32_t)c[off] * 0xFFFFu + A; }

// cornerScore<16> (OpenCV fast_score.cpp): max over the 16 arcs of 9 contiguous ring pixels of
// max(min_arc(v - ring), min_arc(ring - v)) - 1 == largest threshold for which the pixel is a FAST-9 corner.
__device__ __forceinline__ int fast_score_full(const uint8_t* c, uint32_t A) {
    constexpr int P = kWinPitch;
    const int off[16] = {3 * P, 3 * P + 1, 2 * P + 2, P + 3, 3, -P + 3, -2 * P + 2, -3 * P + 1,
                         -3 * P, -3 * P - 1, -2 * P - 2, -P - 3, -3, P - 3, 2 * P - 2, 3 * P - 1};
    uint32_t d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = ring_d2(c, off[k], A);
    uint32_t t[16];
#pragma unroll
    for (int k = 0; k < 16; k++) t[k] = __vimin3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    uint32_t best = 0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        uint32_t m0 = __vimin3_s16x2(t[k], t[(k + 3) & 15], t[(k + 6) & 15]);
        uint32_t m1 = __vimin3_s16x2(t[k + 1], t[(k + 4) & 15], t[(k + 7) & 15]);
        best = __vimax3_s16x2(best, m0, m1);
    }
    int lo = (int)(best & 0xffffu), hi = (int)(best >> 16);
    return max(lo, hi) - 257;
}

__global__ void __launch_bounds__(kFastThreads) k_fast_cells(const OrbGeom* __restrict__ g, const Cell* __restrict__ cells,
                                                             const uint8_t* __restrict__ pyr, uint32_t* __restrict__ cand,
                                                             int* __restrict__ cell_off, int* __restrict__ cell_cnt,
                                                             int* __restrict__ lvl_count, int* __restrict__ flags) {
    __shared__ __align__(16) uint8_t s_win[kMaxWin * kWinPitch + 16];
    __shared__ __align__(16) uint8_t s_score[((kMaxInt + 2) * (kMaxInt + 2) + 15) & ~15];
    __shared__ uint16_t s_queue[kMaxInt * kMaxInt];   // survivors of the high-speed test, packed py << 8 | px
    __shared__ uint32_t s_list[kMaxSurv];             // local maxima: (py * iw + px) << 8 | score
    __shared__ uint32_t s_bits[kFastBitWords];        // kept maxima, one bit per interior pixel (raster order)
    __shared__ int s_pref[kFastBitWords];             // kept maxima before each word
    __shared__ int s_qn, s_ln, s_has_ini, s_base, s_keep;

    const int cid = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const Cell c = cells[cid];
    const LevelGeom& L = g->lv[c.level];
    const int ww = c.x1 - c.x0, wh = c.y1 - c.y0;
    const int iw = ww - 6, ih = wh - 6;
    int* my_off = cell_off + (size_t)f * g->total_cells + cid;
    int* my_cnt = cell_cnt + (size_t)f * g->total_cells + cid;
    if (iw <= 0 || ih <= 0) {
        if (tid == 0) { *my_off = 0; *my_cnt = 0; }
        return;
    }
    const int ini_th = g->ini_th, min_th = g->min_th;
    const int low_th = min(ini_th, min_th);
    if (tid == 0) { s_qn = 0; s_ln = 0; s_has_ini = 0; }
    // window -> shared memory, four pixels per thread and step: two aligned 32-bit loads + a funnel shift per word (the window
    // starts at an arbitrary byte of the plane); the (row, word) of a thread advances without divisions
    const int nw = (ww + 3) >> 2;  // words per window row (<= kWinPitch / 4)
    const int dq = kFastThreads / nw, dr = kFastThreads - dq * nw;
    const uint8_t* src = pyr + L.plane_off + (size_t)f * L.plane_size + (size_t)(c.y0 + kEdge) * L.pitch + c.x0 + kEdge;
    uint32_t* s_win32 = reinterpret_cast<uint32_t*>(s_win);
    {
        int r = tid / nw, k = tid - r * nw;
        while (r < wh) {
            const uint8_t* a = src + (size_t)r * L.pitch + 4 * k;
            const uint32_t* b = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(a) & ~(uintptr_t)3);
            const unsigned sh = ((unsigned)reinterpret_cast<uintptr_t>(a) & 3u) * 8u;
            const uint32_t w0 = __ldg(b), w1 = sh ? __ldg(b + 1) : 0u;
            s_win32[r * (kWinPitch / 4) + k] = __funnelshift_r(w0, w1, sh);
            k += dr; r += dq;
            if (k >= nw) { k -= nw; r++; }
        }
    }
    const int sp = iw + 2;
    {
        uint32_t* sc32 = reinterpret_cast<uint32_t*>(s_score);
        for (int i = tid; i < ((ih + 2) * sp + 3) >> 2; i += kFastThreads) sc32[i] = 0;
    }
    if (tid < kFastBitWords) s_bits[tid] = 0;
    __syncthreads();

    // phase 1: high-speed test on the two opposite pairs (0,8) and (4,12) — any arc of 9 contains one pixel of each pair —
    // for four horizontally adjacent pixels per step with byte-wise SIMD: saturating c +- t, unsigned per-byte compares.
    // Word k of window row py + 3 holds the centres at window columns 4k .. 4k+3; the interior is columns 3 .. 3 + iw - 1.
    {
        const uint32_t T4 = (uint32_t)low_th * 0x01010101u;
        int py = tid / nw, k = tid - py * nw;
        while (py < ih) {
            const uint32_t* row = s_win32 + (py + 3) * (kWinPitch / 4);
            const uint32_t cw = row[k];
            const uint32_t up = row[k - 3 * (kWinPitch / 4)], dn = row[k + 3 * (kWinPitch / 4)];
            const uint32_t wl = k > 0 ? row[k - 1] : 0u, wr = row[k + 1];
            const uint32_t le = __funnelshift_r(wl, cw, 8);    // window columns 4k+j-3
            const uint32_t ri = __funnelshift_r(cw, wr, 24);   // window columns 4k+j+3
            const uint32_t hi = __vaddus4(cw, T4), lo = __vsubus4(cw, T4);
            const uint32_t bright = (__vcmpgtu4(up, hi) | __vcmpgtu4(dn, hi)) & (__vcmpgtu4(le, hi) | __vcmpgtu4(ri, hi));
            const uint32_t dark = (__vcmpltu4(up, lo) | __vcmpltu4(dn, lo)) & (__vcmpltu4(le, lo) | __vcmpltu4(ri, lo));
            uint32_t m = bright | dark;
            while (m) {
                const int j = (__ffs(m) - 1) >> 3;
                m &= ~(0xFFu << (8 * j));
                const int px = 4 * k + j - 3;
                if (px >= 0 && px < iw) {
                    const int q = atomicAdd(&s_qn, 1);
                    s_queue[q] = (uint16_t)((py << 8) | px);
                }
            }
            k += dr; py += dq;
            if (k >= nw) { k -= nw; py++; }
        }
    }
    __syncthreads();
    const int qn = s_qn;
    // phase 2: full score for the surviving pixels (dense over the queue)
    for (int q = tid; q < qn; q += kFastThreads) {
        const int e = s_queue[q], py = e >> 8, px = e & 0xff;
        const uint8_t* cp = s_win + (py + 3) * kWinPitch + px + 3;
        uint32_t v = cp[0];
        uint32_t A = (v + 256u) + ((256u - v) << 16);
        int s = fast_score_full(cp, A);
        if (s >= low_th) s_score[(py + 1) * sp + px + 1] = (uint8_t)s;
    }
    __syncthreads();
    // phase 3: 3x3 non-max suppression inside the cell (strict >, everything outside the cell interior counts 0)
    for (int q = tid; q < qn; q += kFastThreads) {
        const int e = s_queue[q], py = e >> 8, px = e & 0xff;
        const uint8_t* sc = s_score + (py + 1) * sp + px + 1;
        int s = sc[0];
        if (s == 0) continue;
        if (s > sc[-1] && s > sc[1] && s > sc[-sp - 1] && s > sc[-sp] && s > sc[-sp + 1] && s > sc[sp - 1] && s > sc[sp] &&
            s > sc[sp + 1]) {
            int k = atomicAdd(&s_ln, 1);
            s_list[k] = ((uint32_t)(py * iw + px) << 8) | (uint32_t)s;
            if (s >= ini_th) s_has_ini = 1;
        }
    }
    __syncthreads();
    // phase 4: threshold fallback (:809-816) and ordered emission (raster order inside the cell): the kept maxima set their
    // bit in a raster bitmap; a maximum's rank is the number of bits before its own (word prefix + popc)
    const int ln = s_ln;
    const int keep_th = s_has_ini ? ini_th : min_th;
    for (int k = tid; k < ln; k += kFastThreads) {
        const uint32_t e = s_list[k];
        if ((int)(e & 0xffu) >= keep_th) atomicOr(&s_bits[e >> 13], 1u << ((e >> 8) & 31u));
    }
    __syncthreads();
    if (tid < 32) {  // exclusive prefix of the word popcounts: kFastBitWords / 32 words per lane
        constexpr int kPer = kFastBitWords / 32;
        int cnt[kPer], local = 0;
#pragma unroll
        for (int j = 0; j < kPer; j++) { cnt[j] = __popc(s_bits[lane * kPer + j]); local += cnt[j]; }
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int run = incl - local;
#pragma unroll
        for (int j = 0; j < kPer; j++) { s_pref[lane * kPer + j] = run; run += cnt[j]; }
        if (lane == 31) s_keep = incl;
    }
    __syncthreads();
    const int n_keep = s_keep;
    if (tid == 0) {
        int base = 0;
        if (n_keep > 0) {
            base = atomicAdd(lvl_count + (size_t)f * g->nlevels + c.level, n_keep);
            if (base + n_keep > L.cand_cap) {
                atomicOr(flags + f, 1);
                base = -1;
            }
        }
        s_base = base;
        *my_off = base < 0 ? 0 : base;
        *my_cnt = base < 0 ? 0 : n_keep;
    }
    __syncthreads();
    const int base = s_base;
    if (base < 0 || n_keep == 0) return;
    uint32_t* out = cand + (size_t)f * g->cand_per_frame + L.cand_base + base;
    for (int k = tid; k < ln; k += kFastThreads) {
        uint32_t e = s_list[k];
        if ((int)(e & 0xffu) < keep_th) continue;
        const int i = (int)(e >> 8);
        const int rank = s_pref[i >> 5] + __popc(s_bits[i >> 5] & ((1u << (i & 31)) - 1u));
        const int py = i / iw, px = i - py * iw;
        // coordinates relative to minBorder (=16): cell-relative FAST coordinate + j*wCell (:822-823)
        uint32_t x = (uint32_t)(c.x0 + 3 + px - 16), y = (uint32_t)(c.y0 + 3 + py - 16);
        out[rank] = x | (y << 12) | ((e & 0xffu) << 24);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// DistributeOctTree as an ordered, level-synchronous process
// ---------------------------------------------------------------------------------------------------------------
constexpr int kOctThreads = 256;

// exclusive scan of a[0..n) in shared memory, in place; returns the total.  All threads of the block must call.
__device__ int block_excl_scan(int* a, int n, int* s_warp /* >= 9 ints */) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = (n + nt - 1) / nt;
    const int beg = min(tid * per, n), end = min(beg + per, n);
    int sum = 0;
    for (int i = beg; i < end; i++) sum += a[i];
    // block exclusive scan of per-thread sums
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        int nw = nt >> 5;
        int v = tid < nw ? s_warp[tid] : 0;
        int w = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, w, o);
            if (tid >= o) w += u;
        }
        if (tid < nw) s_warp[tid] = w - v;  // exclusive warp offsets
        if (tid == nw - 1) s_warp[32] = w;  // total
    }
    __syncthreads();
    int run = s_warp[tid >> 5] + incl - sum;
    for (int i = beg; i < end; i++) {
        int v = a[i];
        a[i] = run;
        run += v;
    }
    int total = s_warp[32];
    __syncthreads();
    return total;
}

// in-place bitonic sort (descending) of keys[0..n2) with payload, n2 = power of two
__device__ void block_bitonic_desc(uint32_t* keys, int* vals, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    bool desc = (i & k) == 0;
                    uint32_t a = keys[i], b = keys[ixj];
                    if (desc ? (a < b) : (a > b)) {
                        keys[i] = b; keys[ixj] = a;
                        int t = vals[i]; vals[i] = vals[ixj]; vals[ixj] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
}

struct OctSmem {
    short4* box[2];
    int* cnt[2];
    int* ccnt;      // 4*NC child counts
    int* cpos;      // 4*NC scan
    int* kpos;      // NC
    int* eidx;      // NC: node position -> index in E (or -1)
    int* rank_of;   // NC: E index -> processing rank
    uint32_t* ekey[2];  // NC2
    int* eval[2];       // NC2
    int* order;     // NC2: E index per processing rank
    int* gain;      // NC2: list growth per processing rank
    int* cellpre;   // max_cells_level + 1
    int* warp;      // 40
};

__device__ __forceinline__ int quad_of(short4 b, int x, int y) {
    const int hx = (b.z - b.x + 1) >> 1, hy = (b.w - b.y + 1) >> 1;  // ceil(float(d)/2) for d >= 0 (:483-484)
    return (x >= b.x + hx ? 1 : 0) + (y >= b.y + hy ? 2 : 0);          // n1=TL n2=TR n3=BL n4=BR (:515-526)
}
__device__ __forceinline__ short4 child_box(short4 b, int q) {
    const int hx = (b.z - b.x + 1) >> 1, hy = (b.w - b.y + 1) >> 1;
    short4 c;
    c.x = (q & 1) ? b.x + hx : b.x;
    c.z = (q & 1) ? b.z : b.x + hx;
    c.y = (q & 2) ? b.y + hy : b.y;
    c.w = (q & 2) ? b.w : b.y + hy;
    return c;
}

__global__ void __launch_bounds__(kOctThreads) k_octree(const OrbGeom* __restrict__ g, const uint32_t* __restrict__ cand,
                                                       const int* __restrict__ cell_off, const int* __restrict__ cell_cnt,
                                                       const int* __restrict__ lvl_count, uint32_t* __restrict__ ord,
                                                       uint16_t* __restrict__ node_of, uint32_t* __restrict__ lvl_kp,
                                                       int* __restrict__ lvl_n, int* __restrict__ flags) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int level = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const LevelGeom& L = g->lv[level];
    const int NC = g->node_cap;
    int NC2 = 1;
    while (NC2 < NC) NC2 <<= 1;
    OctSmem S;
    {
        uint8_t* p = smem_raw;
        S.box[0] = (short4*)p; p += sizeof(short4) * NC;
        S.box[1] = (short4*)p; p += sizeof(short4) * NC;
        S.cnt[0] = (int*)p; p += 4 * NC;
        S.cnt[1] = (int*)p; p += 4 * NC;
        S.ccnt = (int*)p; p += 16 * NC;
        S.cpos = (int*)p; p += 16 * NC;
        S.kpos = (int*)p; p += 4 * NC;
        S.eidx = (int*)p; p += 4 * NC;
        S.rank_of = (int*)p; p += 4 * NC;
        S.ekey[0] = (uint32_t*)p; p += 4 * NC2;
        S.ekey[1] = (uint32_t*)p; p += 4 * NC2;
        S.eval[0] = (int*)p; p += 4 * NC2;
        S.eval[1] = (int*)p; p += 4 * NC2;
        S.order = (int*)p; p += 4 * NC2;
        S.gain = (int*)p; p += 4 * NC2;
        S.cellpre = (int*)p; p += 4 * (g->max_cells_level + 1);
        S.warp = (int*)p;
    }
    __shared__ int s_ne[2], s_t, s_nexp;

    int M = lvl_count[(size_t)f * g->nlevels + level];
    int* out_n = lvl_n + (size_t)f * g->nlevels + level;
    if (flags[f] != 0 || M > L.cand_cap) M = 0;  // overflowed frame: emit nothing, the host reports PL_ERR_CAPACITY
    if (M == 0) {
        if (tid == 0) *out_n = 0;
        return;
    }
    const uint32_t* cnd = cand + (size_t)f * g->cand_per_frame + L.cand_base;
    uint32_t* od = ord + (size_t)f * g->cand_per_frame + L.cand_base;
    uint16_t* nd = node_of + (size_t)f * g->cand_per_frame + L.cand_base;
    const int* coff = cell_off + (size_t)f * g->total_cells + L.cell_base;
    const int* ccn = cell_cnt + (size_t)f * g->total_cells + L.cell_base;

    // ---- order the candidates: cells row-major, raster inside a cell (:789-829) ----
    for (int i = tid; i < L.n_cells; i += kOctThreads) S.cellpre[i] = ccn[i];
    __syncthreads();
    block_excl_scan(S.cellpre, L.n_cells, S.warp);
    {
        const int warp = tid >> 5, lane = tid & 31;
        for (int cidx = warp; cidx < L.n_cells; cidx += kOctThreads / 32) {
            const int n = ccn[cidx], so = coff[cidx], dofs = S.cellpre[cidx];
            for (int k = lane; k < n; k += 32) od[dofs + k] = cnd[so + k];
        }
    }
    __syncthreads();

    // ---- roots (:543-585) ----
    const int N = L.quota;
    int cur = 0;
    const int nIni = L.n_ini;
    for (int i = tid; i < nIni; i += kOctThreads) S.ccnt[i] = 0;
    __syncthreads();
    for (int k = tid; k < M; k += kOctThreads) {
        uint32_t e = od[k];
        int r = (int)__fdiv_rn((float)(e & 0xfffu), L.hx);
        r = min(r, nIni - 1);
        nd[k] = (uint16_t)r;
        atomicAdd(&S.ccnt[r], 1);
    }
    __syncthreads();
    for (int i = tid; i < nIni; i += kOctThreads) S.kpos[i] = S.ccnt[i] > 0;
    __syncthreads();
    int Sz = block_excl_scan(S.kpos, nIni, S.warp);
    for (int i = tid; i < nIni; i += kOctThreads) {
        if (S.ccnt[i] > 0) {
            short4 b;
            b.x = (short)(int)__fmul_rn(L.hx, (float)i);
            b.z = (short)(int)__fmul_rn(L.hx, (float)(i + 1));
            b.y = 0;
            b.w = (short)L.bh;
            S.box[cur][S.kpos[i]] = b;
            S.cnt[cur][S.kpos[i]] = S.ccnt[i];
        }
    }
    __syncthreads();
    for (int k = tid; k < M; k += kOctThreads) nd[k] = (uint16_t)S.kpos[nd[k]];
    __syncthreads();

    bool finish = false, partial = false;
    int ecur = 0;  // which E buffer holds the expandable set
    int nE = 0;
    bool overflow = false;
    while (!finish) {
        if (!partial) {
            // ---------------- full round (:594-665) ----------------
            const int S0 = Sz;
            for (int i = tid; i < 4 * S0; i += kOctThreads) S.ccnt[i] = 0;
            __syncthreads();
            for (int k = tid; k < M; k += kOctThreads) {
                int p = nd[k];
                if (S.cnt[cur][p] > 1) {
                    uint32_t e = od[k];
                    int q = quad_of(S.box[cur][p], e & 0xfffu, (e >> 12) & 0xfffu);
                    atomicAdd(&S.ccnt[4 * p + q], 1);
                }
            }
            __syncthreads();
            for (int i = tid; i < 4 * S0; i += kOctThreads) S.cpos[i] = S.ccnt[i] > 0;
            for (int i = tid; i < S0; i += kOctThreads) S.kpos[i] = S.cnt[cur][i] == 1;
            if (tid == 0) { s_ne[ecur ^ 1] = 0; s_nexp = 0; }
            __syncthreads();
            const int C = block_excl_scan(S.cpos, 4 * S0, S.warp);
            const int K = block_excl_scan(S.kpos, S0, S.warp);
            const int S1 = C + K;
            if (S1 > NC) { overflow = true; break; }
            const int nxt = cur ^ 1;
            for (int i = tid; i < 4 * S0; i += kOctThreads) {
                int c = S.ccnt[i];
                if (c > 0) {
                    int pos = C - 1 - S.cpos[i];
                    S.box[nxt][pos] = child_box(S.box[cur][i >> 2], i & 3);
                    S.cnt[nxt][pos] = c;
                    if (c > 1) {
                        int e = atomicAdd(&s_ne[ecur ^ 1], 1);
                        S.ekey[ecur ^ 1][e] = ((uint32_t)c << 16) | (uint32_t)S.cpos[i];
                        S.eval[ecur ^ 1][e] = pos;
                    }
                }
            }
            for (int i = tid; i < S0; i += kOctThreads) {
                if (S.cnt[cur][i] == 1) {
                    int pos = C + S.kpos[i];
                    S.box[nxt][pos] = S.box[cur][i];
                    S.cnt[nxt][pos] = 1;
                }
            }
            for (int k = tid; k < M; k += kOctThreads) {
                int p = nd[k];
                if (S.cnt[cur][p] > 1) {
                    uint32_t e = od[k];
                    int q = quad_of(S.box[cur][p], e & 0xfffu, (e >> 12) & 0xfffu);
                    nd[k] = (uint16_t)(C - 1 - S.cpos[4 * p + q]);
                } else {
                    nd[k] = (uint16_t)(C + S.kpos[p]);
                }
            }
            __syncthreads();
            cur = nxt;
            ecur ^= 1;
            nE = s_ne[ecur];
            Sz = S1;
            if (S1 >= N || S1 == S0) finish = true;
            else if (S1 + 3 * nE > N) partial = true;
            __syncthreads();
        } else {
            // ---------------- partial round (:673-738) ----------------
            const int S0 = Sz;
            if (nE == 0) { finish = true; break; }  // size unchanged -> bFinish
            int n2 = 1;
            while (n2 < nE) n2 <<= 1;
            for (int i = tid; i < S0; i += kOctThreads) S.eidx[i] = -1;
            for (int i = tid; i < 4 * nE; i += kOctThreads) S.ccnt[i] = 0;
            for (int i = nE + tid; i < n2; i += kOctThreads) { S.ekey[ecur][i] = 0; S.eval[ecur][i] = -1; }
            __syncthreads();
            // E index = position in the (still unsorted) E arrays; remember it in the payload
            for (int i = tid; i < nE; i += kOctThreads) S.eidx[S.eval[ecur][i]] = i;
            __syncthreads();
            for (int k = tid; k < M; k += kOctThreads) {
                int p = nd[k];
                int e = S.eidx[p];
                if (e >= 0) {
                    uint32_t c = od[k];
                    int q = quad_of(S.box[cur][p], c & 0xfffu, (c >> 12) & 0xfffu);
                    atomicAdd(&S.ccnt[4 * e + q], 1);
                }
            }
            // sort E by (count, creation sequence) descending == the reference's ascending sort walked from the back;
            // payload = E index
            int* order = S.order;
            int* gain = S.gain;
            int* cflag = S.cpos;  // 4*NC ints, free in a partial round; 4*T <= 4*NC
            for (int i = tid; i < n2; i += kOctThreads) order[i] = i < nE ? i : -1;
            __syncthreads();
            block_bitonic_desc(S.ekey[ecur], order, n2);
            // gain of processing rank i = (#non-empty children - 1); first rank at which the list reaches N stops (:729)
            for (int i = tid; i < nE; i += kOctThreads) {
                int e = order[i];
                int ne = (S.ccnt[4 * e] > 0) + (S.ccnt[4 * e + 1] > 0) + (S.ccnt[4 * e + 2] > 0) + (S.ccnt[4 * e + 3] > 0);
                gain[i] = ne - 1;
                S.rank_of[e] = i;
            }
            if (tid == 0) { s_t = nE; s_ne[ecur ^ 1] = 0; }
            __syncthreads();
            block_excl_scan(gain, nE, S.warp);  // gain[i] = sum of gains of ranks < i
            for (int i = tid; i < nE; i += kOctThreads) {
                int e = order[i];
                int ne = (S.ccnt[4 * e] > 0) + (S.ccnt[4 * e + 1] > 0) + (S.ccnt[4 * e + 2] > 0) + (S.ccnt[4 * e + 3] > 0);
                if (S0 + gain[i] + ne - 1 >= N) atomicMin(&s_t, i + 1);
            }
            __syncthreads();
            const int T = s_t;  // ranks [0,T) are split
            for (int i = tid; i < 4 * T; i += kOctThreads) cflag[i] = S.ccnt[4 * order[i >> 2] + (i & 3)] > 0;
            for (int i = tid; i < S0; i += kOctThreads) {
                int e = S.eidx[i];
                S.kpos[i] = !(e >= 0 && S.rank_of[e] < T);
            }
            __syncthreads();
            const int C = block_excl_scan(cflag, 4 * T, S.warp);
            const int K = block_excl_scan(S.kpos, S0, S.warp);
            const int S1 = C + K;
            if (S1 > NC) { overflow = true; break; }
            const int nxt = cur ^ 1;
            for (int i = tid; i < 4 * T; i += kOctThreads) {
                int e = order[i >> 2];
                int c = S.ccnt[4 * e + (i & 3)];
                if (c > 0) {
                    int pos = C - 1 - cflag[i];
                    S.box[nxt][pos] = child_box(S.box[cur][S.eval[ecur][e]], i & 3);
                    S.cnt[nxt][pos] = c;
                    if (c > 1) {
                        int ee = atomicAdd(&s_ne[ecur ^ 1], 1);
                        S.ekey[ecur ^ 1][ee] = ((uint32_t)c << 16) | (uint32_t)cflag[i];
                        S.eval[ecur ^ 1][ee] = pos;
                    }
                }
            }
            for (int i = tid; i < S0; i += kOctThreads) {
                int e = S.eidx[i];
                if (!(e >= 0 && S.rank_of[e] < T)) {
                    int pos = C + S.kpos[i];
                    S.box[nxt][pos] = S.box[cur][i];
                    S.cnt[nxt][pos] = S.cnt[cur][i];
                }
            }
            for (int k = tid; k < M; k += kOctThreads) {
                int p = nd[k];
                int e = S.eidx[p];
                if (e >= 0 && S.rank_of[e] < T) {
                    uint32_t c = od[k];
                    int q = quad_of(S.box[cur][p], c & 0xfffu, (c >> 12) & 0xfffu);
                    nd[k] = (uint16_t)(C - 1 - cflag[4 * S.rank_of[e] + q]);
                } else {
                    nd[k] = (uint16_t)(C + S.kpos[p]);
                }
            }
            __syncthreads();
            cur = nxt;
            ecur ^= 1;
            nE = s_ne[ecur];
            Sz = S1;
            if (S1 >= N || S1 == S0) finish = true;
            __syncthreads();
        }
    }
    if (overflow || Sz > L.out_cap) {
        if (tid == 0) { atomicOr(flags + f, 2); *out_n = 0; }
        return;
    }
    // ---- keep the best response per node, first wins (:742-760) ----
    unsigned long long* best = (unsigned long long*)S.ccnt;  // 16*NC bytes >= 8*NC
    for (int i = tid; i < Sz; i += kOctThreads) best[i] = 0ull;
    __syncthreads();
    for (int k = tid; k < M; k += kOctThreads) {
        unsigned long long v = ((unsigned long long)(od[k] >> 24) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)k);
        atomicMax(&best[nd[k]], v);
    }
    __syncthreads();
    uint32_t* okp = lvl_kp + (size_t)f * g->out_per_frame + L.out_base;
    for (int i = tid; i < Sz; i += kOctThreads) {
        uint32_t k = 0xFFFFFFFFu - (uint32_t)(best[i] & 0xFFFFFFFFull);
        okp[i] = od[k];
    }
    if (tid == 0) *out_n = Sz;
}

// ---------------------------------------------------------------------------------------------------------------
// GaussianBlur 7x7 sigma 2, 8U fixed point (OpenCV >= 4): kernel {18,34,48,56,48,34,18}/256
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBlurTW = 128, kBlurTH = 28;  // one warp blurs a 128-wide, 28-tall strip: 4 pixels per lane, rows streamed
struct BlurTile { short level, tx, ty, pad; };

// horizontal pass of four neighbouring pixels x..x+3: p points at image column x - 3 (plane column x + 16, word aligned
// because the image starts at plane column kEdge = 19), so the 12 bytes b0..b11 hold columns x-3..x+8 and pixel x+j reads
// b[j..j+6].  Two pixels are evaluated per 32-bit operation: P(k) = b[k] | b[k+1] << 16, and the weighted sum of a 16-bit
// half is at most 256 * 255, so nothing carries into the other half.
__device__ __forceinline__ void blur7_hrow(const uint8_t* __restrict__ p, unsigned (&out)[4]) {
    const unsigned w0 = __ldg((const unsigned*)p), w1 = __ldg((const unsigned*)(p + 4)), w2 = __ldg((const unsigned*)(p + 8));
    const unsigned v1 = __funnelshift_r(w0, w1, 8), v5 = __funnelshift_r(w1, w2, 8);  // bytes b1..b4, b5..b8
    const unsigned P0 = __byte_perm(w0, 0, 0x4140), P2 = __byte_perm(w0, 0, 0x4342);
    const unsigned P1 = __byte_perm(v1, 0, 0x4140), P3 = __byte_perm(v1, 0, 0x4342);
    const unsigned P4 = __byte_perm(w1, 0, 0x4140), P6 = __byte_perm(w1, 0, 0x4342);
    const unsigned P5 = __byte_perm(v5, 0, 0x4140), P7 = __byte_perm(v5, 0, 0x4342);
    const unsigned P8 = __byte_perm(w2, 0, 0x4140);
    const unsigned A = 18u * (P0 + P6) + 34u * (P1 + P5) + 48u * (P2 + P4) + 56u * P3;  // pixels x, x+1
    const unsigned B = 18u * (P2 + P8) + 34u * (P3 + P7) + 48u * (P4 + P6) + 56u * P5;  // pixels x+2, x+3
    out[0] = A & 0xffffu;
    out[1] = A >> 16;
    out[2] = B & 0xffffu;
    out[3] = B >> 16;
}

// GaussianBlur 7x7 sigma 2 of OpenCV's 8-bit path: separable 8.8 fixed point {18,34,48,56,48,34,18}, horizontal pass exact,
// vertical pass (sum + 0x8000) >> 16.  One warp per strip; a lane owns 4 columns and keeps the horizontal sums of the last
// seven rows in registers, so every input byte is loaded once per strip (plus the halo) and there is no shared memory.
__global__ void __launch_bounds__(256) k_blur7(const OrbGeom* __restrict__ g, const BlurTile* __restrict__ tiles, int n_items,
                                               const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur) {
    const int item = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (item >= n_items) return;
    const BlurTile t = tiles[item];
    const int f = blockIdx.y;
    const LevelGeom& L = g->lv[t.level];
    const int x = t.tx * kBlurTW + lane * 4, y0 = t.ty * kBlurTH;
    if (x >= L.w) return;
    const int rows = min(kBlurTH, L.h - y0);
    // the bordered plane already holds BORDER_REFLECT_101 of the image itself (19 >= 3), which is exactly the
    // border GaussianBlur applies to the un-bordered clone (:1085-1086); image column x - 3 is plane column x + 16
    const uint8_t* src = pyr + L.plane_off + (size_t)f * L.plane_size + (size_t)(y0 - 3 + kEdge) * L.pitch + (x + kEdge - 3);
    uint8_t* dst = blur + L.blur_off + (size_t)f * L.blur_size + (size_t)y0 * L.bpitch + x;
    unsigned ring[7][4];
#pragma unroll
    for (int k = 0; k < 6; k++) blur7_hrow(src + (size_t)k * L.pitch, ring[k]);
    for (int r0 = 0; r0 < rows; r0 += 7) {
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int r = r0 + k;
            if (r < rows) {
                blur7_hrow(src + (size_t)(r + 6) * L.pitch, ring[(k + 6) % 7]);
                unsigned o = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const unsigned acc = 18u * (ring[k % 7][j] + ring[(k + 6) % 7][j]) + 34u * (ring[(k + 1) % 7][j] + ring[(k + 5) % 7][j]) +
                                         48u * (ring[(k + 2) % 7][j] + ring[(k + 4) % 7][j]) + 56u * ring[(k + 3) % 7][j];
                    o |= ((acc + 0x8000u) >> 16) << (8 * j);
                }
                *(unsigned*)(dst + (size_t)r * L.bpitch) = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// IC_Angle + steered BRIEF + final keypoint assembly: one warp per kept keypoint
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_orient_brief(const OrbGeom* __restrict__ g, const uint8_t* __restrict__ pyr,
                                                      const uint8_t* __restrict__ blur, const uint32_t* __restrict__ lvl_kp,
                                                      const int* __restrict__ lvl_n, pl_keypoint* __restrict__ kps,
                                                      uint8_t* __restrict__ desc, int cap, int* __restrict__ n_out,
                                                      int* __restrict__ flags) {
    __shared__ int8_t s_pat[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_pat[i] = c_pattern[i];
    __syncthreads();
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nl = g->nlevels;
    const int* ln = lvl_n + (size_t)f * nl;
    int total = 0;
    for (int l = 0; l < nl; l++) total += ln[l];
    if (slot == 0 && lane == 0) {
        n_out[f] = total;
        if (total > cap) atomicOr(flags + f, 4);
    }
    if (slot >= g->out_per_frame || total > cap) return;
    int level = 0, offset = 0;
    for (int l = 0; l < nl; l++) {
        if (slot >= g->lv[l].out_base) level = l;
    }
    for (int l = 0; l < level; l++) offset += ln[l];
    const LevelGeom& L = g->lv[level];
    const int idx = slot - L.out_base;
    if (idx >= ln[level]) return;
    const uint32_t e = lvl_kp[(size_t)f * g->out_per_frame + slot];
    const int cx = (int)(e & 0xfffu) + 16, cy = (int)((e >> 12) & 0xfffu) + 16;  // + minBorder (:841-842)
    // --- IC_Angle on the un-blurred bordered plane (:77-104); lane = u + 15 ---
    const uint8_t* ctr = pyr + L.plane_off + (size_t)f * L.plane_size + (size_t)(cy + kEdge) * L.pitch + cx + kEdge;
    int m10 = 0, m01 = 0;
    const int u = lane - kHalfPatch;
    for (int v = -kHalfPatch; v <= kHalfPatch; v++) {
        const int d = g->umax[v < 0 ? -v : v];
        if (lane < 31 && u >= -d && u <= d) {
            int val = ctr[(ptrdiff_t)v * L.pitch + u];
            m10 += u * val;
            m01 += v * val;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    // --- steered BRIEF on the blurred plane (:108-147); lane = descriptor byte ---
    const float factorPI = (float)(3.141592653589793238462643383279502884 / 180.f);
    const float ang = __fmul_rn(angle, factorPI);
    const float a = glibc_sincosf(ang, 1), b = glibc_sincosf(ang, 0);
    const uint8_t* bc = blur + L.blur_off + (size_t)f * L.blur_size + (size_t)cy * L.bpitch + cx;
    const int8_t* pat = s_pat + lane * 32;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float x0 = (float)pat[4 * k], y0 = (float)pat[4 * k + 1], x1 = (float)pat[4 * k + 2], y1 = (float)pat[4 * k + 3];
        const int r0 = cv_round(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = cv_round(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = cv_round(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = cv_round(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = bc[(ptrdiff_t)r0 * L.bpitch + c0], t1 = bc[(ptrdiff_t)r1 * L.bpitch + c1];
        val |= (t0 < t1) << k;
    }
    const size_t o = (size_t)f * cap + offset + idx;
    desc[o * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        pl_keypoint kp;
        kp.x = level ? __fmul_rn((float)cx, L.scale) : (float)cx;
        kp.y = level ? __fmul_rn((float)cy, L.scale) : (float)cy;
        kp.size = L.kp_size;
        kp.angle = angle;
        kp.response = (float)(e >> 24);
        kp.octave = level;
        kp.class_id = -1;
        kps[o] = kp;
    }
}

}  // namespace pl

// =================================================================================================================
// host side: handle, geometry, launches
// =================================================================================================================
using namespace pl;

struct pl_orb {
    int device = 0;
    cudaStream_t stream = nullptr;
    int nfeatures = 0, nlevels = 0, ini_th = 0, min_th = 0;
    double scale_factor_d = 0;  // ORBextractor.h:98 (double member)
    float scale_factor_f = 0;
    std::vector<float> sf, invsf, sigma2, invsigma2;
    std::vector<int> per_level;
    int umax[16];
    int max_cols = 0, max_rows = 0, max_batch = 0;
    // geometry of the current image size
    int rows = 0, cols = 0;
    OrbGeom geom;
    std::vector<Cell> cells;
    std::vector<ResizeTab> tabs;
    std::vector<BlurTile> tiles;
    std::vector<int> half_level;  // 1 if level uses the exact-2x path
    size_t pyr_bytes_per_frame = 0, blur_bytes_per_frame = 0;
    // device buffers
    OrbGeom* d_geom = nullptr;
    Cell* d_cells = nullptr;
    ResizeTab* d_tabs = nullptr;
    BlurTile* d_tiles = nullptr;
    uint8_t *d_in = nullptr, *d_pyr = nullptr, *d_blur = nullptr;
    uint32_t *d_cand = nullptr, *d_ord = nullptr, *d_lvl_kp = nullptr;
    uint16_t* d_node = nullptr;
    int *d_cell_off = nullptr, *d_cell_cnt = nullptr, *d_lvl_count = nullptr, *d_lvl_n = nullptr, *d_flags = nullptr;
    pl_keypoint* d_kps = nullptr;
    uint8_t* d_desc = nullptr;
    int* d_nout = nullptr;
    int out_cap_alloc = 0;
    size_t in_alloc = 0, pyr_alloc = 0, blur_alloc = 0, cand_alloc = 0, cell_alloc = 0, lvlkp_alloc = 0;
    size_t cells_alloc = 0, tabs_alloc = 0, tiles_alloc = 0;
    int* h_flags = nullptr;  // pinned
    int* d_sticky = nullptr; // capacity flags of the device-pointer API since the last pl_orb_sync
    int last_batch = 0;      // frames of the last chunk (for debug reads)
    int staged_frames = 0;   // frames of the caller's images in d_in after a host-pointer call (its last chunk), 0 = none
    cudaEvent_t ev_staged = nullptr;  // recorded behind the staging copy of pl_orb_stage_batch (pl_orb_stream_wait_staged)
    bool staged_pending = false;      // pl_orb_stage_batch staged a batch that pl_orb_extract_staged has not processed yet
    int staged_rows = 0, staged_cols = 0;
    int last_launches = 0;
    size_t oct_smem = 0;
    // optional per-stage timing (CUDA events on the launching stream)
    bool profiling = false;
    cudaEvent_t ev[8] = {nullptr};
    float stage_ms[8] = {0};
    int stage_chunks = 0;
};

namespace {

inline int cvRoundF(float v) { return (int)lrintf(v); }
inline short satShort(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

// fills tabs[off..off+dn) for one axis of cv::resize INTER_LINEAR (resize.cpp)
void build_resize_axis(ResizeTab* out, int sn, int dn) {
    double inv_scale = (double)dn / sn;
    double scale = 1. / inv_scale;
    for (int d = 0; d < dn; d++) {
        float fx = (float)((d + 0.5) * scale - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sn - 1) { fx = 0; sx = sn - 1; }
        out[d].s = (short)sx;
        out[d].a0 = satShort(cvRoundF((1.f - fx) * 2048));
        out[d].a1 = satShort(cvRoundF(fx * 2048));
        out[d].pad = 0;
    }
}

template <typename T>
int ensure(T** p, size_t* have, size_t need_elems) {
    if (*have >= need_elems && *p) return PL_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    PL_CUDA_TRY(cudaMalloc((void**)p, need_elems * sizeof(T)));
    *have = need_elems;
    return PL_OK;
}

int build_geometry(pl_orb* h, int rows, int cols) {
    if (rows == h->rows && cols == h->cols) return PL_OK;
    OrbGeom& G = h->geom;
    memset(&G, 0, sizeof(G));
    G.nlevels = h->nlevels;
    G.ini_th = h->ini_th;
    G.min_th = h->min_th;
    for (int i = 0; i < 16; i++) G.umax[i] = h->umax[i];
    h->cells.clear();
    h->tabs.clear();
    h->tiles.clear();
    h->half_level.assign(h->nlevels, 0);
    size_t plane_off = 0, blur_off = 0;
    int cand_base = 0, out_base = 0, max_cells_level = 1, node_cap = 8;
    const float W = 30;
    for (int l = 0; l < h->nlevels; l++) {
        LevelGeom& L = G.lv[l];
        float scale = h->invsf[l];
        L.w = cvRoundF((float)cols * scale);
        L.h = cvRoundF((float)rows * scale);
        if (L.w < 2 * kEdge + 1 + 6 || L.h < 2 * kEdge + 1 + 6) {
            set_error("image %dx%d too small for %d pyramid levels (level %d is %dx%d)", cols, rows, h->nlevels, l, L.w, L.h);
            return PL_ERR_ARG;
        }
        L.pitch = (int)align_up((size_t)L.w + 2 * kEdge, 128);
        L.bpitch = (int)align_up((size_t)L.w, 128);
        L.plane_size = (size_t)L.pitch * (L.h + 2 * kEdge);
        L.blur_size = (size_t)L.bpitch * L.h;
        L.plane_off = plane_off;
        L.blur_off = blur_off;
        plane_off += L.plane_size * h->max_batch;
        blur_off += L.blur_size * h->max_batch;
        L.scale = h->sf[l];
        L.kp_size = (float)(int)(kPatch * h->sf[l]);
        L.quota = h->per_level[l];
        // cells (:773-806)
        const int minB = kEdge - 3, maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
        const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        L.cell_base = (int)h->cells.size();
        if (nCols >= 1 && nRows >= 1) {
            const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
            for (int i = 0; i < nRows; i++) {
                const float iniY = (float)(minB + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBY - 3) continue;
                if (maxY > maxBY) maxY = (float)maxBY;
                for (int j = 0; j < nCols; j++) {
                    const float iniX = (float)(minB + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBX - 6) continue;
                    if (maxX > maxBX) maxX = (float)maxBX;
                    Cell c;
                    c.level = (short)l;
                    c.x0 = (short)(int)iniX; c.y0 = (short)(int)iniY;
                    c.x1 = (short)(int)maxX; c.y1 = (short)(int)maxY;
                    c.pad = 0;
                    if (c.x1 - c.x0 > kMaxWin || c.y1 - c.y0 > kMaxWin) {
                        set_error("internal: FAST cell window %dx%d exceeds %d", c.x1 - c.x0, c.y1 - c.y0, kMaxWin);
                        return PL_ERR_ARG;
                    }
                    h->cells.push_back(c);
                }
            }
        }
        L.n_cells = (int)h->cells.size() - L.cell_base;
        max_cells_level = std::max(max_cells_level, L.n_cells);
        // octree roots (:543-545)
        L.bw = maxBX - minB;
        L.bh = maxBY - minB;
        L.n_ini = (int)roundf((float)(maxBX - minB) / (float)(maxBY - minB));
        if (L.n_ini < 1) {
            set_error("aspect ratio of level %d (%dx%d) gives zero quadtree roots (reference divides by zero here)", l, L.w, L.h);
            return PL_ERR_ARG;
        }
        L.hx = (float)(maxBX - minB) / (float)L.n_ini;
        // capacities
        const int worst = ((L.bw + 1) / 2) * ((L.bh + 1) / 2);
        L.cand_cap = std::min(std::min(worst, 32 * L.quota + 4096), 65535);  // counts are packed into 16 bits in the quadtree keys
        L.cand_base = cand_base;
        cand_base += L.cand_cap;
        L.out_cap = std::max(L.quota + 3, 4 * L.n_ini) + 1;
        L.out_base = out_base;
        out_base += L.out_cap;
        node_cap = std::max(node_cap, L.out_cap + 4);
        // resize tables
        L.xtab_off = (int)h->tabs.size();
        h->tabs.resize(h->tabs.size() + L.w);
        L.ytab_off = (int)h->tabs.size();
        h->tabs.resize(h->tabs.size() + L.h);
        if (l > 0) {
            const LevelGeom& P = G.lv[l - 1];
            build_resize_axis(h->tabs.data() + L.xtab_off, P.w, L.w);
            build_resize_axis(h->tabs.data() + L.ytab_off, P.h, L.h);
            double sx = 1. / ((double)L.w / P.w), sy = 1. / ((double)L.h / P.h);
            int isx = (int)lrint(sx), isy = (int)lrint(sy);
            if (fabs(sx - isx) < 2.2204460492503131e-16 && fabs(sy - isy) < 2.2204460492503131e-16 && isx == 2 && isy == 2)
                h->half_level[l] = 1;
        }
        // blur tiles
        L.tile_base = (int)h->tiles.size();
        L.n_tiles_x = (L.w + kBlurTW - 1) / kBlurTW;
        L.n_tiles_y = (L.h + kBlurTH - 1) / kBlurTH;
        for (int ty = 0; ty < L.n_tiles_y; ty++)
            for (int tx = 0; tx < L.n_tiles_x; tx++) h->tiles.push_back(BlurTile{(short)l, (short)tx, (short)ty, 0});
    }
    G.total_cells = (int)h->cells.size();
    G.total_tiles = (int)h->tiles.size();
    G.cand_per_frame = cand_base;
    G.out_per_frame = out_base;
    G.node_cap = node_cap;
    G.max_cells_level = max_cells_level;
    h->pyr_bytes_per_frame = plane_off / h->max_batch;
    h->blur_bytes_per_frame = blur_off / h->max_batch;
    // octree shared memory
    int nc2 = 1;
    while (nc2 < node_cap) nc2 <<= 1;
    h->oct_smem = (size_t)node_cap * (8 * 2 + 4 * 2 + 16 + 16 + 4 + 4 + 4) + (size_t)nc2 * 24 + 4 * (size_t)(max_cells_level + 1) + 4 * 48;
    if (h->oct_smem > 200 * 1024) {
        set_error("nfeatures too large for the shared-memory quadtree (needs %zu bytes)", h->oct_smem);
        return PL_ERR_ARG;
    }
    PL_CUDA_TRY(cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->oct_smem));
    // device buffers
    const size_t B = h->max_batch;
    int rc;
    if ((rc = ensure(&h->d_pyr, &h->pyr_alloc, plane_off + 256)) != PL_OK) return rc;
    if ((rc = ensure(&h->d_blur, &h->blur_alloc, blur_off + 256)) != PL_OK) return rc;
    size_t cand_elems = B * (size_t)G.cand_per_frame;
    if (h->cand_alloc < cand_elems) {
        if (h->d_cand) cudaFree(h->d_cand);
        if (h->d_ord) cudaFree(h->d_ord);
        if (h->d_node) cudaFree(h->d_node);
        h->d_cand = h->d_ord = nullptr; h->d_node = nullptr; h->cand_alloc = 0;
        PL_CUDA_TRY(cudaMalloc((void**)&h->d_cand, cand_elems * 4));
        PL_CUDA_TRY(cudaMalloc((void**)&h->d_ord, cand_elems * 4));
        PL_CUDA_TRY(cudaMalloc((void**)&h->d_node, cand_elems * 2));
        h->cand_alloc = cand_elems;
    }
    size_t cell_elems = B * (size_t)G.total_cells;
    if (h->cell_alloc < cell_elems) {
        if (h->d_cell_off) cudaFree(h->d_cell_off);
        if (h->d_cell_cnt) cudaFree(h->d_cell_cnt);
        h->d_cell_off = h->d_cell_cnt = nullptr; h->cell_alloc = 0;
        PL_CUDA_TRY(cudaMalloc((void**)&h->d_cell_off, cell_elems * 4));
        PL_CUDA_TRY(cudaMalloc((void**)&h->d_cell_cnt, cell_elems * 4));
        h->cell_alloc = cell_elems;
    }
    if ((rc = ensure(&h->d_lvl_kp, &h->lvlkp_alloc, B * (size_t)G.out_per_frame)) != PL_OK) return rc;
    if ((rc = ensure(&h->d_cells, &h->cells_alloc, h->cells.size())) != PL_OK) return rc;
    if ((rc = ensure(&h->d_tabs, &h->tabs_alloc, h->tabs.size())) != PL_OK) return rc;
    if ((rc = ensure(&h->d_tiles, &h->tiles_alloc, h->tiles.size())) != PL_OK) return rc;
    PL_CUDA_TRY(cudaMemcpyAsync(h->d_geom, &G, sizeof(G), cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(h->d_cells, h->cells.data(), h->cells.size() * sizeof(Cell), cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(h->d_tabs, h->tabs.data(), h->tabs.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(h->d_tiles, h->tiles.data(), h->tiles.size() * sizeof(BlurTile), cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    h->rows = rows;
    h->cols = cols;
    return PL_OK;
}

// one chunk (<= max_batch frames), all device pointers; asynchronous
int launch_chunk(pl_orb* h, const uint8_t* d_gray, int nf, size_t step, size_t frame_stride, pl_keypoint* d_kps,
                 uint8_t* d_desc, int cap, int* d_nout) {
    const OrbGeom& G = h->geom;
    cudaStream_t st = h->stream;
    int launches = 0;
    PL_CUDA_TRY(cudaMemsetAsync(h->d_lvl_count, 0, sizeof(int) * (size_t)nf * G.nlevels, st));
    PL_CUDA_TRY(cudaMemsetAsync(h->d_flags, 0, sizeof(int) * (size_t)nf, st));
    const bool prof = h->profiling;
    if (prof) cudaEventRecord(h->ev[0], st);
    {
        const LevelGeom& L = G.lv[0];
        dim3 grid((L.pitch / 4 + 255) / 256, L.h + 2 * kEdge, nf);
        k_pyr_base<<<grid, 256, 0, st>>>(h->d_geom, d_gray, step, frame_stride, h->d_pyr, nf);
        launches++;
    }
    for (int l = 1; l < G.nlevels; l++) {
        const LevelGeom& L = G.lv[l];
        dim3 grid((L.pitch / 4 + 255) / 256, L.h + 2 * kEdge, nf);
        if (h->half_level[l]) k_pyr_half<<<grid, 256, 0, st>>>(h->d_geom, l, h->d_pyr);
        else k_pyr_resize<<<grid, 256, 0, st>>>(h->d_geom, l, h->d_tabs, h->d_pyr);
        launches++;
    }
    if (prof) cudaEventRecord(h->ev[1], st);
    k_fast_cells<<<dim3(G.total_cells, nf), kFastThreads, 0, st>>>(h->d_geom, h->d_cells, h->d_pyr, h->d_cand, h->d_cell_off,
                                                                  h->d_cell_cnt, h->d_lvl_count, h->d_flags);
    launches++;
    if (prof) cudaEventRecord(h->ev[2], st);
    k_octree<<<dim3(G.nlevels, nf), kOctThreads, h->oct_smem, st>>>(h->d_geom, h->d_cand, h->d_cell_off, h->d_cell_cnt,
                                                                   h->d_lvl_count, h->d_ord, h->d_node, h->d_lvl_kp, h->d_lvl_n,
                                                                   h->d_flags);
    launches++;
    if (prof) cudaEventRecord(h->ev[3], st);
    k_blur7<<<dim3((G.total_tiles + 7) / 8, nf), 256, 0, st>>>(h->d_geom, h->d_tiles, G.total_tiles, h->d_pyr, h->d_blur);
    launches++;
    if (prof) cudaEventRecord(h->ev[4], st);
    k_orient_brief<<<dim3((G.out_per_frame + 7) / 8, nf), 256, 0, st>>>(h->d_geom, h->d_pyr, h->d_blur, h->d_lvl_kp, h->d_lvl_n,
                                                                        d_kps, d_desc, cap, d_nout, h->d_flags);
    launches++;
    PL_CUDA_TRY(cudaGetLastError());
    if (prof) {
        cudaEventRecord(h->ev[5], st);
        PL_CUDA_TRY(cudaEventSynchronize(h->ev[5]));
        for (int i = 0; i < 5; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
            h->stage_ms[i] += ms;
        }
        h->stage_chunks++;
    }
    h->last_batch = nf;
    h->last_launches += launches;
    return PL_OK;
}

// device-pointer API: the per-frame capacity flags of a chunk are folded into one sticky word that pl_orb_sync reports
__global__ void __launch_bounds__(32) k_or_flags(const int* __restrict__ flags, int n, int* __restrict__ sticky) {
    int v = 0;
    for (int i = threadIdx.x; i < n; i += 32) v |= flags[i];
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0 && v) atomicOr(sticky, v);
}

int check_flags(pl_orb* h, int nf) {
    PL_CUDA_TRY(cudaMemcpyAsync(h->h_flags, h->d_flags, sizeof(int) * nf, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    for (int i = 0; i < nf; i++)
        if (h->h_flags[i]) {
            set_error("frame %d of the chunk exceeded a capacity (flags=%d: 1=FAST candidates, 2=quadtree nodes, 4=caller cap)", i,
                      h->h_flags[i]);
            return PL_ERR_CAPACITY;
        }
    return PL_OK;
}

}  // namespace

extern "C" {

PL_API int pl_orb_create(pl_orb** out, int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast,
                         int device, int max_cols, int max_rows, int max_batch) {
    PL_CHECK_ARG(out != nullptr);
    *out = nullptr;
    PL_CHECK_ARG(nfeatures > 0 && nlevels >= 1 && nlevels <= kMaxLevels && scale_factor > 1.0f);
    PL_CHECK_ARG(ini_th_fast >= 1 && min_th_fast >= 1 && ini_th_fast <= 254 && min_th_fast <= 254);
    PL_CHECK_ARG(max_cols > 0 && max_rows > 0 && max_cols <= 4000 && max_rows <= 4000 && max_batch >= 1);
    int ndev = 0;
    PL_CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) {
        set_error("device %d not available (%d CUDA devices); this library has no CPU fallback", device, ndev);
        return PL_ERR_CUDA;
    }
    PL_CUDA_TRY(cudaSetDevice(device));
    pl_orb* h = new pl_orb();
    h->device = device;
    h->nfeatures = nfeatures; h->nlevels = nlevels; h->ini_th = ini_th_fast; h->min_th = min_th_fast;
    h->scale_factor_f = scale_factor;
    h->scale_factor_d = scale_factor;
    h->max_cols = max_cols; h->max_rows = max_rows; h->max_batch = max_batch;
    // ORBextractor::ORBextractor (:410-470)
    h->sf.resize(nlevels); h->sigma2.resize(nlevels); h->invsf.resize(nlevels); h->invsigma2.resize(nlevels);
    h->sf[0] = 1.0f; h->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        h->sf[i] = (float)(h->sf[i - 1] * h->scale_factor_d);
        h->sigma2[i] = h->sf[i] * h->sf[i];
    }
    for (int i = 0; i < nlevels; i++) {
        h->invsf[i] = 1.0f / h->sf[i];
        h->invsigma2[i] = 1.0f / h->sigma2[i];
    }
    h->per_level.resize(nlevels);
    float factor = (float)(1.0f / h->scale_factor_d);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        h->per_level[l] = cvRoundF(nDesired);
        sum += h->per_level[l];
        nDesired *= factor;
    }
    h->per_level[nlevels - 1] = std::max(nfeatures - sum, 0);
    {
        int v, v0, vmax = (int)floorf(kHalfPatch * sqrtf(2.f) / 2 + 1);
        int vmin = (int)ceilf(kHalfPatch * sqrtf(2.f) / 2);
        const double hp2 = kHalfPatch * kHalfPatch;
        for (v = 0; v < 16; v++) h->umax[v] = 0;
        for (v = 0; v <= vmax; ++v) h->umax[v] = (int)lrint(sqrt(hp2 - v * v));
        for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
            while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
            h->umax[v] = v0;
            ++v0;
        }
    }
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_geom, sizeof(OrbGeom));
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_lvl_count, sizeof(int) * (size_t)max_batch * kMaxLevels);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_lvl_n, sizeof(int) * (size_t)max_batch * kMaxLevels);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_flags, sizeof(int) * (size_t)max_batch);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_sticky, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(h->d_sticky, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_nout, sizeof(int) * (size_t)max_batch);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&h->h_flags, sizeof(int) * (size_t)max_batch);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_pattern, h_pattern, sizeof(h_pattern));
    if (e != cudaSuccess) {
        set_error("pl_orb_create: %s", cudaGetErrorString(e));
        pl_orb_destroy(h);
        return PL_ERR_CUDA;
    }
    int rc = build_geometry(h, max_rows, max_cols);
    if (rc != PL_OK) {
        pl_orb_destroy(h);
        return rc;
    }
    *out = h;
    return PL_OK;
}

PL_API void pl_orb_destroy(pl_orb* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) pl::stream_sync(h->stream);
    void* bufs[] = {h->d_geom, h->d_cells, h->d_tabs, h->d_tiles, h->d_in, h->d_pyr, h->d_blur, h->d_cand, h->d_ord, h->d_lvl_kp,
                    h->d_node, h->d_cell_off, h->d_cell_cnt, h->d_lvl_count, h->d_lvl_n, h->d_flags, h->d_sticky, h->d_kps, h->d_desc, h->d_nout};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    for (int i = 0; i < 8; i++)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->ev_staged) cudaEventDestroy(h->ev_staged);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

PL_API int pl_orb_levels(const pl_orb* h) { return h ? h->nlevels : 0; }
PL_API float pl_orb_scale_factor(const pl_orb* h) { return h ? (float)h->scale_factor_d : 0.f; }
static int copy_vec(const pl_orb* h, const std::vector<float>& v, float* out) {
    PL_CHECK_ARG(h && out);
    for (int i = 0; i < h->nlevels; i++) out[i] = v[i];
    return PL_OK;
}
PL_API int pl_orb_scale_factors(const pl_orb* h, float* out) { return copy_vec(h, h->sf, out); }
PL_API int pl_orb_inv_scale_factors(const pl_orb* h, float* out) { return copy_vec(h, h->invsf, out); }
PL_API int pl_orb_level_sigma2(const pl_orb* h, float* out) { return copy_vec(h, h->sigma2, out); }
PL_API int pl_orb_inv_level_sigma2(const pl_orb* h, float* out) { return copy_vec(h, h->invsigma2, out); }
PL_API int pl_orb_features_per_level(const pl_orb* h, int* out) {
    PL_CHECK_ARG(h && out);
    for (int i = 0; i < h->nlevels; i++) out[i] = h->per_level[i];
    return PL_OK;
}
PL_API int pl_orb_max_keypoints(const pl_orb* h) { return h ? h->geom.out_per_frame : 0; }
PL_API void* pl_orb_stream(pl_orb* h) { return h ? (void*)h->stream : nullptr; }
PL_API int pl_orb_last_launches(const pl_orb* h) { return h ? h->last_launches : 0; }

PL_API int pl_orb_set_profiling(pl_orb* h, int on) {
    PL_CHECK_ARG(h);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    if (on && !h->ev[0])
        for (int i = 0; i < 8; i++) PL_CUDA_TRY(cudaEventCreate(&h->ev[i]));
    h->profiling = on != 0;
    for (int i = 0; i < 8; i++) h->stage_ms[i] = 0;
    h->stage_chunks = 0;
    return PL_OK;
}
PL_API int pl_orb_stage_ms(pl_orb* h, float* out5, int* chunks) {
    PL_CHECK_ARG(h && out5);
    for (int i = 0; i < 5; i++) out5[i] = h->stage_ms[i];
    if (chunks) *chunks = h->stage_chunks;
    return PL_OK;
}
PL_API int pl_orb_bytes_per_frame(const pl_orb* h, long long* pyr, long long* blur, long long* level_px, long long* bordered_px) {
    PL_CHECK_ARG(h);
    long long P = 0, Pb = 0;
    for (int l = 0; l < h->nlevels; l++) {
        P += (long long)h->geom.lv[l].w * h->geom.lv[l].h;
        Pb += (long long)(h->geom.lv[l].w + 2 * kEdge) * (h->geom.lv[l].h + 2 * kEdge);
    }
    if (pyr) *pyr = (long long)h->pyr_bytes_per_frame;
    if (blur) *blur = (long long)h->blur_bytes_per_frame;
    if (level_px) *level_px = P;
    if (bordered_px) *bordered_px = Pb;
    return PL_OK;
}

PL_API int pl_orb_sync(pl_orb* h) {
    PL_CHECK_ARG(h);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    int sticky = 0;
    PL_CUDA_TRY(pl::stream_sync(h->stream));  // a copy into pageable memory blocks inside the runtime until the stream gets there
    PL_CUDA_TRY(cudaMemcpyAsync(&sticky, h->d_sticky, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    if (sticky) {
        PL_CUDA_TRY(cudaMemsetAsync(h->d_sticky, 0, sizeof(int), h->stream));
        set_error("a frame extracted through the device-pointer API exceeded a capacity (flags=%d: 1=FAST candidates, 2=quadtree nodes, "
                  "4=caller cap); it produced no keypoints", sticky);
        return PL_ERR_CAPACITY;
    }
    return PL_OK;
}

PL_API int pl_orb_extract_batch_dev(pl_orb* h, const uint8_t* d_gray, int n_frames, int rows, int cols, size_t step,
                                    size_t frame_stride, pl_keypoint* d_kps, uint8_t* d_desc, int cap, int* d_n_out) {
    PL_CHECK_ARG(h && d_kps && d_desc && d_n_out && cap > 0);
    if (!d_gray || rows <= 0 || cols <= 0 || n_frames <= 0) {
        set_error("empty image");
        return PL_ERR_EMPTY;
    }
    PL_CHECK_ARG(cols <= h->max_cols && rows <= h->max_rows && step >= (size_t)cols);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    int rc = build_geometry(h, rows, cols);
    if (rc != PL_OK) return rc;
    h->last_launches = 0;
    for (int f0 = 0; f0 < n_frames; f0 += h->max_batch) {
        int nf = std::min(h->max_batch, n_frames - f0);
        rc = launch_chunk(h, d_gray + (size_t)f0 * frame_stride, nf, step, frame_stride, d_kps + (size_t)f0 * cap,
                          d_desc + (size_t)f0 * cap * 32, cap, d_n_out + f0);
        if (rc != PL_OK) return rc;
        k_or_flags<<<1, 32, 0, h->stream>>>(h->d_flags, nf, h->d_sticky);
    }
    return PL_OK;
}

PL_API int pl_orb_staged_images_dev(pl_orb* h, const uint8_t** d_images, int* n_frames, int* rows, int* cols, size_t* step, size_t* frame_stride) {
    PL_CHECK_ARG(h && d_images && n_frames && rows && cols && step && frame_stride);
    if (h->staged_frames <= 0 || !h->d_in) {
        set_error("no images staged: pl_orb_staged_images_dev follows a host-pointer extract call");
        return PL_ERR_STATE;
    }
    *d_images = h->d_in;
    *n_frames = h->staged_frames;
    *rows = h->staged_rows;
    *cols = h->staged_cols;
    *step = align_up((size_t)h->staged_cols, 16);
    *frame_stride = *step * (size_t)h->staged_rows;
    return PL_OK;
}

}  // extern "C"

namespace {
// the staging buffer of the host-pointer calls (max_batch frames at the staging pitch)
int ensure_staging(pl_orb* h, int rows, int cols) {
    const size_t need_in = (size_t)h->max_batch * align_up((size_t)cols, 16) * rows;
    if (h->in_alloc < need_in) {
        if (h->d_in) cudaFree(h->d_in);
        h->d_in = nullptr; h->in_alloc = 0;
        PL_CUDA_TRY(cudaMalloc((void**)&h->d_in, need_in));
        h->in_alloc = need_in;
    }
    return PL_OK;
}
// nf frames from host memory into the staging buffer, asynchronously on the handle's stream: one copy when the caller's frames are
// dense and already have the staging pitch, else one strided copy per frame
int stage_frames(pl_orb* h, const uint8_t* gray, int nf, int rows, int cols, size_t step, size_t frame_stride) {
    const size_t in_pitch = align_up((size_t)cols, 16);
    if (step == in_pitch && frame_stride == in_pitch * (size_t)rows)
        PL_CUDA_TRY(cudaMemcpyAsync(h->d_in, gray, (size_t)nf * frame_stride, cudaMemcpyHostToDevice, h->stream));
    else
        for (int f = 0; f < nf; f++)
            PL_CUDA_TRY(cudaMemcpy2DAsync(h->d_in + (size_t)f * in_pitch * rows, in_pitch, gray + (size_t)f * frame_stride, step, cols, rows,
                                          cudaMemcpyHostToDevice, h->stream));
    h->staged_frames = nf; h->staged_rows = rows; h->staged_cols = cols;
    return PL_OK;
}
int extract_batch_impl(pl_orb* h, const uint8_t* gray, int n_frames, int rows, int cols, size_t step, size_t frame_stride, pl_keypoint* kps,
                       uint8_t* desc, int cap, int* n_out, bool already_staged);
}  // namespace

extern "C" {

PL_API int pl_orb_extract_batch(pl_orb* h, const uint8_t* gray, int n_frames, int rows, int cols, size_t step,
                                size_t frame_stride, pl_keypoint* kps, uint8_t* desc, int cap, int* n_out) {
    PL_CHECK_ARG(h && kps && desc && n_out && cap > 0);
    if (!gray || rows <= 0 || cols <= 0 || n_frames <= 0) {
        set_error("empty image");
        return PL_ERR_EMPTY;
    }
    PL_CHECK_ARG(cols <= h->max_cols && rows <= h->max_rows && step >= (size_t)cols);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->staged_pending = false;
    return extract_batch_impl(h, gray, n_frames, rows, cols, step, frame_stride, kps, desc, cap, n_out, false);
}

// The two halves of pl_orb_extract_batch for ONE chunk (n_frames <= max_batch).  pl_orb_stage_batch only enqueues the copy of the
// frames into the staging buffer and returns; pl_orb_stream_wait_staged makes another stream (the line extractor's) wait for that
// copy, so that pl_line_extract_batch_from_dev can start on the staged frames while pl_orb_extract_staged — the rest of the call:
// kernels, read-back, capacity check — is still running.  One cv::Mat serves both extractors in the reference (Frame.cc:152-155).
PL_API int pl_orb_stage_batch(pl_orb* h, const uint8_t* gray, int n_frames, int rows, int cols, size_t step, size_t frame_stride) {
    PL_CHECK_ARG(h != nullptr);
    if (!gray || rows <= 0 || cols <= 0 || n_frames <= 0) {
        set_error("empty image");
        return PL_ERR_EMPTY;
    }
    PL_CHECK_ARG(cols <= h->max_cols && rows <= h->max_rows && step >= (size_t)cols);
    if (n_frames > h->max_batch) {
        set_error("pl_orb_stage_batch stages one chunk: %d frames exceed max_batch %d", n_frames, h->max_batch);
        return PL_ERR_CAPACITY;
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    int rc = build_geometry(h, rows, cols);
    if (rc != PL_OK) return rc;
    if ((rc = ensure_staging(h, rows, cols)) != PL_OK) return rc;
    if (!h->ev_staged) PL_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_staged, cudaEventDisableTiming));
    if ((rc = stage_frames(h, gray, n_frames, rows, cols, step, frame_stride)) != PL_OK) return rc;
    PL_CUDA_TRY(cudaEventRecord(h->ev_staged, h->stream));
    h->staged_pending = true;
    return PL_OK;
}

PL_API int pl_orb_stream_wait_staged(pl_orb* h, void* stream) {
    PL_CHECK_ARG(h != nullptr);
    if (!h->staged_pending || !h->ev_staged) {
        set_error("nothing staged: pl_orb_stream_wait_staged follows pl_orb_stage_batch");
        return PL_ERR_STATE;
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    PL_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_staged, 0));
    return PL_OK;
}

PL_API int pl_orb_extract_staged(pl_orb* h, pl_keypoint* kps, uint8_t* desc, int cap, int* n_out) {
    PL_CHECK_ARG(h && kps && desc && n_out && cap > 0);
    if (!h->staged_pending || h->staged_frames <= 0) {
        set_error("nothing staged: pl_orb_extract_staged follows pl_orb_stage_batch");
        return PL_ERR_STATE;
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    const size_t in_pitch = align_up((size_t)h->staged_cols, 16);
    const int rc = extract_batch_impl(h, nullptr, h->staged_frames, h->staged_rows, h->staged_cols, in_pitch, in_pitch * (size_t)h->staged_rows, kps, desc,
                                      cap, n_out, true);
    h->staged_pending = false;
    return rc;
}

}  // extern "C"

namespace {
int extract_batch_impl(pl_orb* h, const uint8_t* gray, int n_frames, int rows, int cols, size_t step, size_t frame_stride, pl_keypoint* kps,
                       uint8_t* desc, int cap, int* n_out, bool already_staged) {
    int rc = build_geometry(h, rows, cols);
    if (rc != PL_OK) return rc;
    const size_t B = h->max_batch;
    const size_t in_pitch = align_up((size_t)cols, 16);
    if ((rc = ensure_staging(h, rows, cols)) != PL_OK) return rc;
    if (h->out_cap_alloc < cap) {
        if (h->d_kps) cudaFree(h->d_kps);
        if (h->d_desc) cudaFree(h->d_desc);
        h->d_kps = nullptr; h->d_desc = nullptr; h->out_cap_alloc = 0;
        PL_CUDA_TRY(cudaMalloc((void**)&h->d_kps, B * (size_t)cap * sizeof(pl_keypoint)));
        PL_CUDA_TRY(cudaMalloc((void**)&h->d_desc, B * (size_t)cap * 32));
        h->out_cap_alloc = cap;
    }
    h->last_launches = 0;
    for (int f0 = 0; f0 < n_frames; f0 += h->max_batch) {
        const int nf = std::min(h->max_batch, n_frames - f0);
        if (!already_staged && (rc = stage_frames(h, gray + (size_t)f0 * frame_stride, nf, rows, cols, step, frame_stride)) != PL_OK) return rc;
        rc = launch_chunk(h, h->d_in, nf, in_pitch, in_pitch * rows, h->d_kps, h->d_desc, cap, h->d_nout);
        if (rc != PL_OK) return rc;
        PL_CUDA_TRY(cudaMemcpyAsync(n_out + f0, h->d_nout, sizeof(int) * nf, cudaMemcpyDeviceToHost, h->stream));
        if (nf > 1) {
            // a batch: the chunk's key points and descriptors in one copy each (the rows of a frame beyond its count are whatever the
            // device buffer held: the caller reads n_out[f] of them) instead of two small copies per frame
            PL_CUDA_TRY(cudaMemcpyAsync(kps + (size_t)f0 * cap, h->d_kps, sizeof(pl_keypoint) * (size_t)nf * cap, cudaMemcpyDeviceToHost, h->stream));
            PL_CUDA_TRY(cudaMemcpyAsync(desc + (size_t)f0 * cap * 32, h->d_desc, (size_t)nf * cap * 32, cudaMemcpyDeviceToHost, h->stream));
        }
        rc = check_flags(h, nf);  // synchronises
        if (rc != PL_OK) return rc;
        if (nf > 1) {  // rows beyond a frame's count stay zero, as the per-frame prefix copies left them
            for (int f = 0; f < nf; f++) {
                const int n = std::max(0, std::min(n_out[f0 + f], cap));
                memset(kps + (size_t)(f0 + f) * cap + n, 0, sizeof(pl_keypoint) * (size_t)(cap - n));
                memset(desc + ((size_t)(f0 + f) * cap + n) * 32, 0, (size_t)(cap - n) * 32);
            }
        }
        if (nf == 1) {  // a single frame: only the filled prefix
            const int n = n_out[f0];
            if (n > 0) {
                PL_CUDA_TRY(cudaMemcpyAsync(kps + (size_t)f0 * cap, h->d_kps, sizeof(pl_keypoint) * n, cudaMemcpyDeviceToHost, h->stream));
                PL_CUDA_TRY(cudaMemcpyAsync(desc + (size_t)f0 * cap * 32, h->d_desc, (size_t)n * 32, cudaMemcpyDeviceToHost, h->stream));
                PL_CUDA_TRY(pl::stream_sync(h->stream));
            }
        }
    }
    return PL_OK;
}
}  // namespace

extern "C" {

PL_API int pl_orb_extract(pl_orb* h, const uint8_t* gray, int rows, int cols, size_t step, pl_keypoint* kps, uint8_t* desc,
                          int cap, int* n_out) {
    return pl_orb_extract_batch(h, gray, 1, rows, cols, step, step * (size_t)(rows > 0 ? rows : 0), kps, desc, cap, n_out);
}

PL_API int pl_orb_pyramid_dims(const pl_orb* h, int level, int* rows, int* cols) {
    PL_CHECK_ARG(h && rows && cols && level >= 0 && level < h->nlevels);
    *rows = h->geom.lv[level].h;
    *cols = h->geom.lv[level].w;
    return PL_OK;
}

PL_API int pl_orb_pyramid_read(pl_orb* h, int frame, int level, uint8_t* out, size_t out_step) {
    PL_CHECK_ARG(h && out && level >= 0 && level < h->nlevels);
    if (frame < 0 || frame >= h->last_batch) {
        set_error("no extract result for frame %d", frame);
        return PL_ERR_STATE;
    }
    const LevelGeom& L = h->geom.lv[level];
    PL_CHECK_ARG(out_step >= (size_t)L.w + 2 * kEdge);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    PL_CUDA_TRY(cudaMemcpy2DAsync(out, out_step, h->d_pyr + L.plane_off + (size_t)frame * L.plane_size, L.pitch, L.w + 2 * kEdge,
                                  L.h + 2 * kEdge, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    return PL_OK;
}

PL_API int pl_orb_pyramid_dev(pl_orb* h, int frame, int level, const uint8_t** d_image, size_t* pitch, int* rows, int* cols) {
    PL_CHECK_ARG(h && d_image && pitch && level >= 0 && level < h->nlevels);
    if (frame < 0 || frame >= h->last_batch) {
        set_error("no extract result for frame %d", frame);
        return PL_ERR_STATE;
    }
    const LevelGeom& L = h->geom.lv[level];
    *d_image = h->d_pyr + L.plane_off + (size_t)frame * L.plane_size + (size_t)kEdge * L.pitch + kEdge;
    *pitch = L.pitch;
    if (rows) *rows = L.h;
    if (cols) *cols = L.w;
    return PL_OK;
}

PL_API int pl_orb_blurred_read(pl_orb* h, int frame, int level, uint8_t* out, size_t out_step) {
    PL_CHECK_ARG(h && out && level >= 0 && level < h->nlevels);
    if (frame < 0 || frame >= h->last_batch) {
        set_error("no extract result for frame %d", frame);
        return PL_ERR_STATE;
    }
    const LevelGeom& L = h->geom.lv[level];
    PL_CHECK_ARG(out_step >= (size_t)L.w);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    PL_CUDA_TRY(cudaMemcpy2DAsync(out, out_step, h->d_blur + L.blur_off + (size_t)frame * L.blur_size, L.bpitch, L.w, L.h,
                                  cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    return PL_OK;
}

PL_API int pl_orb_candidates_read(pl_orb* h, int frame, int level, float* xs, float* ys, float* responses, int cap, int* n_out) {
    PL_CHECK_ARG(h && xs && ys && responses && n_out && level >= 0 && level < h->nlevels);
    if (frame < 0 || frame >= h->last_batch) {
        set_error("no extract result for frame %d", frame);
        return PL_ERR_STATE;
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    const LevelGeom& L = h->geom.lv[level];
    int n = 0;
    PL_CUDA_TRY(pl::stream_sync(h->stream));  // a copy into pageable memory blocks inside the runtime until the stream gets there
    PL_CUDA_TRY(cudaMemcpyAsync(&n, h->d_lvl_count + (size_t)frame * h->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    *n_out = n;
    if (n > cap || n > L.cand_cap) return PL_ERR_CAPACITY;
    std::vector<uint32_t> tmp(n);
    if (n) {
        PL_CUDA_TRY(cudaMemcpyAsync(tmp.data(), h->d_ord + (size_t)frame * h->geom.cand_per_frame + L.cand_base, sizeof(uint32_t) * n,
                                    cudaMemcpyDeviceToHost, h->stream));
        PL_CUDA_TRY(pl::stream_sync(h->stream));
    }
    for (int i = 0; i < n; i++) {
        xs[i] = (float)(tmp[i] & 0xfffu);
        ys[i] = (float)((tmp[i] >> 12) & 0xfffu);
        responses[i] = (float)(tmp[i] >> 24);
    }
    return PL_OK;
}

}  // extern "C"

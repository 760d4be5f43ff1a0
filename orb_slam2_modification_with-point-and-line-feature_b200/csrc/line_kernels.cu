// line_kernels.cu — line extraction on sm_100a: the CUDA path behind pl_line_* (include/plslam_c.h).
//
// Replaces LineExtractor::ExtractLineSegment (reference src/LineExtractor.cpp:12-70) and the OpenCV code it calls:
//   LSDDetector::detect -> cv::LineSegmentDetector(LSD_REFINE_ADV)  (imgproc/src/lsd.cpp, OpenCV 4.13 behaviour)
//       k_lsd_scale      GaussianBlur 7x7 sigma 0.75 (8.8 fixed point) + resize 0.8 INTER_LINEAR_EXACT, fused, smem tile
//       k_lsd_grad       2x2 gradient, level-line angle (fastAtan2), |grad|^2, per-frame max        (streaming)
//       k_lsd_bin_count / k_lsd_bin_scan / k_lsd_scatter   stable counting sort of the seeds by 1024 gradient bins
//       k_lsd_grow       ordered region growing + rect fit + refine (speculative grower warps, persistent CTAs)
//   top-80 by response (:23-35) + KeyLine fields (LSDDetector.cpp)  -> k_line_finalize
//   BinaryDescriptor::compute (binary_descriptor.cpp)               -> k_lbd_sobel (blur 5x5 sigma 1 + Sobel, fused),
//                                                                     k_lbd_rows, k_lbd_finish
//   line coefficients (:60-69)                                      -> k_lbd_finish
//
// Region growing is inherently ordered (greedy by gradient bin, shared USED map, running region angle).  It keeps the
// reference's semantics exactly and runs as a window of speculative transactions with in-order commit (see k_lsd_grow);
// it is a latency-bound stage and is reported as time, not as an HBM fraction.
#include <math.h>

#include <algorithm>
#include <vector>

#include <cuda.h>  // CUtensorMap (the driver entry point is looked up at run time: no link against libcuda)

#include "pl_common.cuh"
#include "pl_glibc_sincos.cuh"

namespace pl {

constexpr double kPiD = 3.14159265358979323846;
constexpr double kDegToRad = kPiD / 180;
constexpr double k2Pi = 2 * kPiD;
constexpr double k32Pi = (3 * kPiD) / 2;
constexpr float kNotDefDeg = -1024.0f;  // stored angle of pixels with undefined gradient
constexpr int kBins = 1024;
constexpr int kTileRows = 8;  // rows per sort tile (one warp walks a tile in raster order)
// sin / cos of k / 128 for the host-exact double sin / cos (pl_glibc_sincos.cuh)
__device__ const double g_sincostab[440] = {
#include "pl_sincostab.inc"
};

struct LineGeom {
    int cols, rows;        // input image
    int W, H;              // LSD works on the 0.8-scaled image
    int in_pitch;          // staged input pitch
    int spitch;            // scaled image pitch (bytes)
    int n_tiles;           // sort tiles per frame
    int seg_cap;           // max LSD segments per frame
    int reg_cap;           // max region size (= W*H)
    int min_reg_size;
    double log_nt;
    int max_lines;
    int dpitch;            // Sobel plane pitch in elements
};
struct ExactTab {  // INTER_LINEAR_EXACT coefficients (8.8 fixed point)
    short ofs, c0, c1, pad;
};

// ---------------------------------------------------------------------------------------------------------------
// k_lsd_scale: GaussianBlur(7x7, 0.75) -> 8-bit, then resize(0.8, INTER_LINEAR_EXACT)  (lsd.cpp flsd())
// fixed-point blur kernel {0,4,56,136,56,4,0}/256 (the outer taps are zero: a 5-tap filter)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kScTW = 64, kScTH = 16;                    // output tile
constexpr int kScSW = kScTW * 5 / 4 + 4, kScSH = kScTH * 5 / 4 + 4;  // source tile bound (scale 1.25 + bilinear + slack)

// Every stage works on four neighbouring pixels per thread and on the whole (constant-size) tile: no index divisions by run-time
// widths, 32 / 64-bit shared-memory accesses, one 32-bit store per four output pixels.  Tile pixels outside the span an output
// tile needs are computed and never read.
__global__ void __launch_bounds__(256) k_lsd_scale(LineGeom g, const uint8_t* __restrict__ in, size_t in_pitch,
                                                   size_t in_frame_stride, const ExactTab* __restrict__ xtab,
                                                   const ExactTab* __restrict__ ytab, uint8_t* __restrict__ scaled,
                                                   size_t scaled_frame_stride) {
    constexpr int SP = kScSW + 4;                    // source tile + 2 px blur halo, bytes per row (a multiple of 4)
    constexpr int SR = kScSH + 4;
    static_assert(SP % 4 == 0 && kScSW % 4 == 0, "tile rows are processed four pixels at a time");
    __shared__ __align__(16) uint8_t s_src[SR * SP];
    __shared__ __align__(16) uint16_t s_h[SR * kScSW];      // after the horizontal pass
    __shared__ __align__(16) uint8_t s_b[kScSH * kScSW];    // blurred 8-bit
    const int f = blockIdx.z, tid = threadIdx.x;
    const int ox0 = blockIdx.x * kScTW, oy0 = blockIdx.y * kScTH;
    const int sx0 = xtab[ox0].ofs, sy0 = ytab[oy0].ofs;   // first source column / row of this output tile
    const uint8_t* src = in + (size_t)f * in_frame_stride;
    // ---- source tile (REFLECT_101 at the image border) ----
    for (int i = tid; i < SR * (SP / 4); i += 256) {
        const int r = i / (SP / 4), q = i - r * (SP / 4);
        const uint8_t* row = src + (size_t)reflect101(sy0 + r - 2, g.rows) * in_pitch;
        const int x = sx0 + 4 * q - 2;
        uint32_t w;
        if (x >= 0 && x + 3 < g.cols) {
            w = (uint32_t)__ldg(row + x) | ((uint32_t)__ldg(row + x + 1) << 8) | ((uint32_t)__ldg(row + x + 2) << 16) | ((uint32_t)__ldg(row + x + 3) << 24);
        } else {
            w = (uint32_t)__ldg(row + reflect101(x, g.cols)) | ((uint32_t)__ldg(row + reflect101(x + 1, g.cols)) << 8) |
                ((uint32_t)__ldg(row + reflect101(x + 2, g.cols)) << 16) | ((uint32_t)__ldg(row + reflect101(x + 3, g.cols)) << 24);
        }
        reinterpret_cast<uint32_t*>(s_src)[r * (SP / 4) + q] = w;
    }
    __syncthreads();
    // ---- horizontal pass: h[c] = 4 (p[c] + p[c+4]) + 56 (p[c+1] + p[c+3]) + 136 p[c+2], 16 bits ----
    for (int i = tid; i < SR * (kScSW / 4); i += 256) {
        const int r = i / (kScSW / 4), q = i - r * (kScSW / 4);
        const uint32_t w0 = reinterpret_cast<const uint32_t*>(s_src)[r * (SP / 4) + q], w1 = reinterpret_cast<const uint32_t*>(s_src)[r * (SP / 4) + q + 1];
        const uint32_t p0 = w0 & 255u, p1 = (w0 >> 8) & 255u, p2 = (w0 >> 16) & 255u, p3 = w0 >> 24;
        const uint32_t p4 = w1 & 255u, p5 = (w1 >> 8) & 255u, p6 = (w1 >> 16) & 255u, p7 = w1 >> 24;
        const uint32_t h0 = 4u * (p0 + p4) + 56u * (p1 + p3) + 136u * p2, h1 = 4u * (p1 + p5) + 56u * (p2 + p4) + 136u * p3;
        const uint32_t h2 = 4u * (p2 + p6) + 56u * (p3 + p5) + 136u * p4, h3 = 4u * (p3 + p7) + 56u * (p4 + p6) + 136u * p5;
        reinterpret_cast<uint2*>(s_h)[r * (kScSW / 4) + q] = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
    }
    __syncthreads();
    // ---- vertical pass, rounded to 8 bits ----
    for (int i = tid; i < kScSH * (kScSW / 4); i += 256) {
        const int r = i / (kScSW / 4), q = i - r * (kScSW / 4);
        const uint2* col = reinterpret_cast<const uint2*>(s_h) + r * (kScSW / 4) + q;
        const uint2 v0 = col[0], v1 = col[kScSW / 4], v2 = col[2 * (kScSW / 4)], v3 = col[3 * (kScSW / 4)], v4 = col[4 * (kScSW / 4)];
        auto tap = [](uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4) { return (4u * (a0 + a4) + 56u * (a1 + a3) + 136u * a2 + 0x8000u) >> 16; };
        const uint32_t b0 = tap(v0.x & 0xffffu, v1.x & 0xffffu, v2.x & 0xffffu, v3.x & 0xffffu, v4.x & 0xffffu);
        const uint32_t b1 = tap(v0.x >> 16, v1.x >> 16, v2.x >> 16, v3.x >> 16, v4.x >> 16);
        const uint32_t b2 = tap(v0.y & 0xffffu, v1.y & 0xffffu, v2.y & 0xffffu, v3.y & 0xffffu, v4.y & 0xffffu);
        const uint32_t b3 = tap(v0.y >> 16, v1.y >> 16, v2.y >> 16, v3.y >> 16, v4.y >> 16);
        reinterpret_cast<uint32_t*>(s_b)[r * (kScSW / 4) + q] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
    __syncthreads();
    // ---- resize: thread = four output pixels of a row ----
    static_assert(kScTW * kScTH == 4 * 256, "one item per thread");
    const int r = tid / (kScTW / 4), oy = oy0 + r, oxq = ox0 + 4 * (tid - r * (kScTW / 4));
    if (oy >= g.H || oxq >= g.W) return;
    const ExactTab ty = ytab[oy];
    const int ay = ty.ofs - sy0, by = min(ty.ofs + 1, g.rows - 1) - sy0;
    uint32_t outw = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int ox = oxq + k;
        if (ox < g.W) {
            const ExactTab tx = xtab[ox];
            const int ax = tx.ofs - sx0, bx = min(tx.ofs + 1, g.cols - 1) - sx0;
            const uint32_t r0 = s_b[ay * kScSW + ax] * tx.c0 + s_b[ay * kScSW + bx] * tx.c1;
            const uint32_t r1 = s_b[by * kScSW + ax] * tx.c0 + s_b[by * kScSW + bx] * tx.c1;
            outw |= (((r0 * ty.c0 + r1 * ty.c1 + 32768u) >> 16) & 255u) << (8 * k);
        }
    }
    // (the pitch is a multiple of 16 and the tile starts at a multiple of 64: an aligned word; columns beyond W are padding)
    *reinterpret_cast<uint32_t*>(scaled + (size_t)f * scaled_frame_stride + (size_t)oy * g.spitch + oxq) = outw;
}

// ---------------------------------------------------------------------------------------------------------------
// k_lsd_grad: ll_angle() — gradient with a 2x2 mask, level-line angle, squared modulus, per-frame max
// ---------------------------------------------------------------------------------------------------------------
// also stores cosf/sinf of the angle (what region_grow adds to its running sums) and the double-precision cos/sin
// rounded to float (what a seed starts its sums from), so the ordered grower only loads them
__global__ void __launch_bounds__(256) k_lsd_grad(LineGeom g, const uint8_t* __restrict__ scaled, size_t scaled_frame_stride,
                                                  float* __restrict__ angdeg, int* __restrict__ g2, float4* __restrict__ rec,
                                                  float2* __restrict__ cs0, size_t plane, double rho, int* __restrict__ max_g2) {
    const int f = blockIdx.z;
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    int my = -1;
    if (x < g.W && y < g.H) {
        float a = kNotDefDeg;
        float ca = 0.f, sa = 0.f;
        int q = 0;
        if (x < g.W - 1 && y < g.H - 1) {
            const uint8_t* r0 = scaled + (size_t)f * scaled_frame_stride + (size_t)y * g.spitch + x;
            const uint8_t* r1 = r0 + g.spitch;
            const int DA = (int)r1[1] - (int)r0[0], BC = (int)r0[1] - (int)r1[0];
            const int gx = DA + BC, gy = DA - BC;
            q = gx * gx + gy * gy;
            const double norm = sqrt((double)q / 4.0);
            if (norm > rho) {
                a = fast_atan2_deg((float)gx, (float)(-gy));
                my = q;
                // cos(float(angle)) / sin(float(angle)) of the reference's region_grow resolve to cosf / sinf
                const float af = (float)((double)a * kDegToRad);
                ca = glibc_sincosf(af, 1);
                sa = glibc_sincosf(af, 0);
                // a seed starts its sums from float(std::cos(double angle)), float(std::sin(double angle))
                const double ad = (double)a * kDegToRad;
                const GlibcSinCos sc{g_sincostab};
                cs0[(size_t)f * plane + (size_t)y * g.W + x] = make_float2((float)sc.cos(ad), (float)sc.sin(ad));
            }
        }
        angdeg[(size_t)f * plane + (size_t)y * g.W + x] = a;
        g2[(size_t)f * plane + (size_t)y * g.W + x] = q;
        // the grower's record of the pixel, one 16-byte load: angle | claim stamp (0xffff = none) | cosf | sinf
        rec[(size_t)f * plane + (size_t)y * g.W + x] = make_float4(a, __uint_as_float(0xffffu), ca, sa);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my = max(my, __shfl_xor_sync(0xffffffffu, my, o));
    if ((threadIdx.x & 31) == 0 && my >= 0) atomicMax(max_g2 + f, my);
}

// gradient bin of a pixel (ll_angle(): int(norm * bin_coef)), or -1 when the gradient is undefined
__device__ __forceinline__ int lsd_bin(float a, int q, double bin_coef) {
    if (a == kNotDefDeg) return -1;
    return (int)(sqrt((double)q / 4.0) * bin_coef);
}
__device__ __forceinline__ double lsd_bin_coef(int maxq) {
    if (maxq < 0) return 0.0;
    const double max_grad = sqrt((double)maxq / 4.0);
    return max_grad > 0 ? (double)(kBins - 1) / max_grad : 0.0;
}

// per-tile histograms: one warp walks kTileRows rows
// (also leaves every pixel's bin, 0xffff = no gradient, in `bins` for k_lsd_scatter: the bin is a double sqrt + multiply per pixel)
__global__ void __launch_bounds__(256) k_lsd_bin_count(LineGeom g, const float* __restrict__ angdeg, const int* __restrict__ g2,
                                                       size_t plane, const int* __restrict__ max_g2,
                                                       unsigned short* __restrict__ tile_hist, unsigned short* __restrict__ bins) {
    __shared__ int s_hist[8][kBins];
    const int f = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * 8 + w;
    for (int i = lane; i < kBins; i += 32) s_hist[w][i] = 0;
    __syncwarp();
    if (tile < g.n_tiles) {
        const double coef = lsd_bin_coef(max_g2[f]);
        const int y0 = tile * kTileRows, y1 = min(y0 + kTileRows, g.H - 1);
        for (int y = y0; y < y1; y++) {
            const size_t row = (size_t)f * plane + (size_t)y * g.W;
            for (int x = lane; x < g.W - 1; x += 32) {
                int b = lsd_bin(angdeg[row + x], g2[row + x], coef);
                bins[row + x] = (unsigned short)b;   // -1 -> 0xffff (kBins <= 0xffff)
                if (b >= 0) atomicAdd(&s_hist[w][b], 1);
            }
        }
        __syncwarp();
        unsigned short* out = tile_hist + ((size_t)f * g.n_tiles + tile) * kBins;
        for (int i = lane; i < kBins; i += 32) out[i] = (unsigned short)s_hist[w][i];
    }
}

// per frame: exclusive offsets over (bin descending, tile ascending); tile_off[tile][bin]; n_seeds[f] = total
__global__ void __launch_bounds__(kBins) k_lsd_bin_scan(LineGeom g, const unsigned short* __restrict__ tile_hist,
                                                        int* __restrict__ tile_off, int* __restrict__ n_seeds) {
    __shared__ int s_warp[33];
    const int f = blockIdx.x, t = threadIdx.x;
    const int bin = kBins - 1 - t;  // thread t owns the t-th bin in visiting order
    const unsigned short* h = tile_hist + (size_t)f * g.n_tiles * kBins;
    int tot = 0;
    for (int k = 0; k < g.n_tiles; k++) tot += h[(size_t)k * kBins + bin];
    // block exclusive scan over t
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    if ((t & 31) == 31) s_warp[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        int v = s_warp[t], w = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, w, o);
            if (t >= o) w += u;
        }
        s_warp[t] = w - v;
        if (t == 31) s_warp[32] = w;
    }
    __syncthreads();
    int run = s_warp[t >> 5] + incl - tot;
    int* o = tile_off + (size_t)f * g.n_tiles * kBins;
    for (int k = 0; k < g.n_tiles; k++) {
        o[(size_t)k * kBins + bin] = run;
        run += h[(size_t)k * kBins + bin];
    }
    if (t == 0) n_seeds[f] = s_warp[32];
}

// stable scatter: seeds[f][pos] = raster index (y*W+x); order = bin descending, raster ascending inside a bin
__global__ void __launch_bounds__(256) k_lsd_scatter(LineGeom g, const unsigned short* __restrict__ bins, size_t plane,
                                                     const int* __restrict__ tile_off, unsigned int* __restrict__ seeds) {
    __shared__ int s_off[8][kBins];
    const int f = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * 8 + w;
    if (tile >= g.n_tiles) return;
    const int* off = tile_off + ((size_t)f * g.n_tiles + tile) * kBins;
    for (int i = lane; i < kBins; i += 32) s_off[w][i] = off[i];
    __syncwarp();
    unsigned int* out = seeds + (size_t)f * plane;
    const int y0 = tile * kTileRows, y1 = min(y0 + kTileRows, g.H - 1);
    const unsigned lt = (1u << lane) - 1u;
    for (int y = y0; y < y1; y++) {
        const size_t row = (size_t)f * plane + (size_t)y * g.W;
        for (int xb = 0; xb < g.W - 1; xb += 32) {
            const int x = xb + lane;
            int b = -1;
            if (x < g.W - 1) {
                const unsigned short bs = bins[row + x];
                b = bs == 0xffffu ? -1 : (int)bs;
            }
            const unsigned valid = __ballot_sync(0xffffffffu, b >= 0);
            if (b >= 0) {
                const unsigned peers = __match_any_sync(valid, b);
                const int rank = __popc(peers & lt);
                const int leader = __ffs(peers) - 1;
                int base = 0;
                if (lane == leader) {
                    base = s_off[w][b];
                    s_off[w][b] = base + __popc(peers);
                }
                base = __shfl_sync(peers, base, leader);
                out[base + rank] = (unsigned)(y * g.W + x);
            }
            __syncwarp();
        }
    }
}

}  // namespace pl

// =================================================================================================================
// k_lsd_grow — the ordered part of LSD
// =================================================================================================================
namespace pl {

struct LsdSeg {  // one detected segment, in detection order (== cv::LineSegmentDetector::detect outputs)
    float x1, y1, x2, y2;
    double width, p, nfa;
};
struct LsdRect {
    double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p;
};
constexpr int kRegRing = 256;    // most recent region points kept in shared memory (the BFS frontier reads them)
constexpr int kSpecCap = 16384;  // region / touched-list capacity of a speculative grower (larger regions run exclusively)
// What the grower reads of a pixel, in ONE 16-byte load (the three separate planes it used to read cost three sectors per
// neighbour): the level-line angle (degrees, or kNotDefDeg), the claim stamp (low 16 bits of the ticket that accepted the pixel
// last, 0xffff = none; a hint only), and cosf / sinf of the angle (what an accepted pixel adds to the region's running sums).
struct LsdPix {
    float ang;
    unsigned int claim;
    float ca, sa;
};
struct LsdFrame {  // per-frame (and per-grower) device views
    const float* ang;   // level-line angle in degrees or kNotDefDeg
    const int* g2;      // gx^2 + gy^2
    LsdPix* rec;        // per-pixel records
    const float2* cs0;  // (float(cos(double angle)), float(sin(double angle))): a seed's initial sums
    float2* sval;       // shared-memory staging of one warp, 36 entries
    const unsigned int* used_bits;  // the committed USED map: one bit per pixel, shared memory
    unsigned int* reg;  // region points, packed y<<16 | x
    unsigned int* ring; // shared-memory copy of reg[n - kRegRing .. n)
    double* scratch;    // shared memory of the warp, 3 x 32 doubles: the ordered sums of region2rect / refine
    int W, H;
    // The pixels a grower marks stay private until its region is committed.  Speculative growers keep them in a
    // sparse bitmap in shared memory (a directory of 32x32-pixel tiles and a small pool of tile bitmaps); the
    // commit-time re-growth, which has no size limit, uses a full bitmap in global memory.
    bool sparse;
    unsigned char* dir;   // [tiles] pool slot of the tile, 0xff = no pixel marked in it
    unsigned short* rev;  // [pool_tiles] tile of a pool slot
    unsigned int* pool;   // [pool_tiles][32] one word per tile row
    int* ntiles;          // pool slots in use
    int tw, pool_tiles;
    unsigned int* bits;   // full private bitmap (W*H bits)
    unsigned int* touched;  // log of accepted pixels (packed), capacity touched_cap; nullptr = not logged (a first growth: the log is the region)
    unsigned int* touched_buf;  // where the log goes once a re-growth (refine) starts
    int reg_cap;            // capacity of reg
    int touched_cap;        // capacity of touched
    // in-flight claims (a hint that saves wasted growth, never needed for correctness): every accepted pixel is
    // stamped with the low 16 bits of the grower's ticket; a grower that is about to accept a pixel stamped by an
    // earlier ticket that is still uncommitted gives up (the seed is re-grown when its turn to commit comes).
    int ticket;
    const volatile int* commit_head;
};
__device__ __forceinline__ bool lsd_committed(const LsdFrame& F, unsigned o) { return (F.used_bits[o >> 5] >> (o & 31)) & 1u; }
__device__ __forceinline__ bool lsd_priv_test(const LsdFrame& F, int x, int y, unsigned o) {
    if (!F.sparse) return (F.bits[o >> 5] >> (o & 31)) & 1u;
    const unsigned d = F.dir[(y >> 5) * F.tw + (x >> 5)];
    return d != 0xffu && ((F.pool[d * 32 + (y & 31)] >> (x & 31)) & 1u);
}
// the tile must exist (lsd_priv_alloc)
__device__ __forceinline__ void lsd_mark(const LsdFrame& F, int x, int y, unsigned o) {
    if (!F.sparse) atomicOr(&F.bits[o >> 5], 1u << (o & 31));
    else atomicOr(&F.pool[(unsigned)F.dir[(y >> 5) * F.tw + (x >> 5)] * 32 + (y & 31)], 1u << (x & 31));
}
// only called for pixels that are marked
__device__ __forceinline__ void lsd_unmark(const LsdFrame& F, int x, int y, unsigned o) {
    if (!F.sparse) atomicAnd(&F.bits[o >> 5], ~(1u << (o & 31)));
    else atomicAnd(&F.pool[(unsigned)F.dir[(y >> 5) * F.tw + (x >> 5)] * 32 + (y & 31)], ~(1u << (x & 31)));
}
// warp-collective: makes sure the tiles of the lanes with `want` exist; false when the pool is exhausted
__device__ __noinline__ bool lsd_priv_alloc_tiles(unsigned int* pool, unsigned char* dir, unsigned short* rev, int* ntiles, int pool_tiles,
                                                  int t, unsigned need) {
    const int lane = threadIdx.x & 31;
    while (need) {
        const int tj = __shfl_sync(0xffffffffu, t, __ffs(need) - 1);
        const int k = *(volatile int*)ntiles;
        if (k >= pool_tiles) return false;
        pool[k * 32 + lane] = 0;
        __syncwarp();
        if (lane == 0) {
            dir[tj] = (unsigned char)k;
            rev[k] = (unsigned short)tj;
            *(volatile int*)ntiles = k + 1;
        }
        __syncwarp();
        need &= ~__ballot_sync(0xffffffffu, t == tj);
    }
    return true;
}
__device__ __forceinline__ bool lsd_priv_alloc(const LsdFrame& F, bool want, int x, int y) {
    if (!F.sparse) return true;
    const int t = want ? (y >> 5) * F.tw + (x >> 5) : -1;
    const unsigned need = __ballot_sync(0xffffffffu, want && F.dir[max(t, 0)] == 0xffu);
    if (__builtin_expect(!need, 1)) return true;
    return lsd_priv_alloc_tiles(F.pool, F.dir, F.rev, F.ntiles, F.pool_tiles, t, need);
}
// back to "nothing marked"; nt = entries of the touched log (every pixel ever marked is in it)
__device__ __forceinline__ void lsd_priv_reset(const LsdFrame& F, int nt) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (F.sparse) {
        const int k = *(volatile int*)F.ntiles;
        #pragma unroll 1
        for (int s2 = lane; s2 < k; s2 += 32) F.dir[F.rev[s2]] = 0xffu;
        __syncwarp();
        if (lane == 0) *(volatile int*)F.ntiles = 0;
    } else {
        #pragma unroll 1
        for (int i = lane; i < nt; i += 32) {
            const unsigned pp = F.touched[i];
            const unsigned o = (pp >> 16) * (unsigned)F.W + (pp & 0xffffu);
            atomicAnd(&F.bits[o >> 5], ~(1u << (o & 31)));
        }
    }
    __syncwarp();
}

__device__ __forceinline__ double lsd_angle_diff_signed(double a, double b) {
    double diff = a - b;
    while (diff <= -kPiD) diff += k2Pi;
    while (diff > kPiD) diff -= k2Pi;
    return diff;
}
// isAligned() without the bounds / NOTDEF checks: |theta - a| folded, <= prec
__device__ __forceinline__ bool lsd_aligned(double theta, double a, double prec) {
    double n = theta - a;
    if (n < 0) n = -n;
    if (n > k32Pi) {
        n -= k2Pi;
        if (n < 0) n = -n;
    }
    return n <= prec;
}
__device__ __forceinline__ double lsd_dist_sq(double x1, double y1, double x2, double y2) {
    return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1);
}

// (rare path, kept out of line: the grower is bound by instruction fetch)
__device__ __noinline__ bool lsd_aligned_cold(float th, float adeg, double prec) {
    return lsd_aligned((double)th * kDegToRad, (double)adeg * kDegToRad, prec);
}
// isAligned() on the float degree values the reference converts to double radians: decided in float when the
// distance is not within 2e-3 degrees of the tolerance or of the fold point, else with the reference's own formula.
__device__ __forceinline__ bool lsd_aligned_deg(float th, float adeg, float precdeg, double prec) {
    const float t = fabsf(th - adeg);
    const float tf = t > 270.f ? fabsf(t - 360.f) : t;
    if (__builtin_expect(fabsf(tf - precdeg) > 2e-3f && fabsf(t - 270.f) > 2e-3f, 1)) return tf <= precdeg;
    return lsd_aligned_cold(th, adeg, prec);
}

// ---- shared-memory accessors by 32-bit shared-space address ----
// The views hold generic pointers; a generic load of shared memory is a long-scoreboard access with 64-bit address arithmetic.
// The grower's hot loop converts them once and uses ld.shared / st.shared / atom.shared.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned lds_u32(unsigned a) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ unsigned lds_u8(unsigned a) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds_f2(unsigned a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f2(unsigned a, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void reds_or(unsigned a, unsigned v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// region_grow(): returns the region size; reg_angle (radians) is returned through *out_angle.
//
// The 32 lanes hold the 8-neighbourhoods of up to four consecutive region points in the reference's visiting order
// (point, then row, then column; the centre is always USED and is skipped).  Acceptances must be resolved in that
// order with the RUNNING region angle, which changes after every accepted pixel.  A batch is resolved by
// hypothesis and verification:
//   1. hypothesis H: the lanes that pass with a hint angle (the most recent angle known);
//   2. every lane forms, in parallel, the float sums the reference would hold when it reaches that lane IF H is what
//      happened before it (the accepted lanes' cos/sin are staged in shared memory and added in lane order), takes
//      fastAtan2 of them and evaluates its own exact test;
//   3. up to and including the first lane whose exact verdict differs from H, every verdict is the sequential one
//      (induction over the lanes); those lanes are committed, the rest goes through another pass.
// Most batches take one pass, whose serial chain is one fastAtan2.  The global loads of the next batch are issued
// before the current one is resolved.
// `nt` counts the entries of the touched log (speculative mode); returns -1 when a capacity is exceeded, -2 when the region
// runs into a pixel stamped by an earlier uncommitted ticket.
// This is the instruction stream the grower warps live in: it is written for a small footprint (the kernel was bound by
// instruction fetch) — no unrolling, shared memory by 32-bit address, one call-free body; kSparse selects the private marks
// (speculative growers: tile directory in shared memory; commit-time re-growth: full bitmap in global memory).
__device__ __forceinline__ bool lsd_claim_hit(int ticket, int commit_head, unsigned claim) {
    const unsigned d = (unsigned)(ticket - (int)claim) & 0xffffu;  // tickets between the stamp and this grower
    return d != 0 && d <= (unsigned)(ticket - commit_head);
}
__device__ __noinline__ float fast_atan2_deg_ool(float y, float x) { return fast_atan2_deg(y, x); }

template <bool kSparse>
__device__ __noinline__ int lsd_region_grow_t(const LsdFrame& Fin, int sx, int sy, double prec, double* out_angle, int& nt_io) {
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
    const unsigned a_used = smem_u32(Fin.used_bits), a_sval = smem_u32(Fin.sval), a_ring = smem_u32(Fin.ring);
    const unsigned a_head = smem_u32((const void*)Fin.commit_head);
    const unsigned a_dir = kSparse ? smem_u32(Fin.dir) : 0u, a_pool = kSparse ? smem_u32(Fin.pool) : 0u;
    LsdPix* const rec = Fin.rec;
    unsigned int* const reg = Fin.reg;
    unsigned int* const touched = Fin.touched;
    unsigned int* const bits = Fin.bits;
    const int W = Fin.W, H = Fin.H, tw = Fin.tw, ticket = Fin.ticket;
    const int reg_cap = Fin.reg_cap, touched_cap = Fin.touched_cap;
    int nt = nt_io;
    const unsigned so = (unsigned)sy * (unsigned)W + (unsigned)sx;
    const float seed_deg = Fin.ang[so];
    const float2 c0 = Fin.cs0[so];
    float sumdx = c0.x, sumdy = c0.y;
    float hint = seed_deg;
    const float precdeg = (float)(prec * (180.0 / kPiD));
    if (touched && nt >= touched_cap) return -1;
    // private mark of a pixel: the tile must exist first (sparse)
    auto tile_alloc = [&](bool want, int x, int y) -> bool {  // warp-collective; false when the pool is exhausted
        if (!kSparse) return true;
        const int t = want ? (y >> 5) * tw + (x >> 5) : -1;
        unsigned need = __ballot_sync(FULL, want && lds_u8(a_dir + (unsigned)max(t, 0)) == 0xffu);
        if (__builtin_expect(need == 0, 1)) return true;
        const unsigned a_rev = smem_u32(Fin.rev), a_nt = smem_u32(Fin.ntiles);
        #pragma unroll 1
        while (need) {
            const int tj = __shfl_sync(FULL, t, __ffs(need) - 1);
            const int k = (int)lds_u32(a_nt);
            if (k >= Fin.pool_tiles) return false;
            sts_u32(a_pool + (unsigned)(k * 32 + lane) * 4u, 0u);
            __syncwarp();
            if (lane == 0) {
                asm volatile("st.shared.u8 [%0], %1;" ::"r"(a_dir + (unsigned)tj), "r"(k) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(a_rev + (unsigned)k * 2u), "r"(tj) : "memory");
                sts_u32(a_nt, (unsigned)(k + 1));
            }
            __syncwarp();
            need &= ~__ballot_sync(FULL, t == tj);
        }
        return true;
    };
    auto mark = [&](int x, int y, unsigned o) {
        if (kSparse) reds_or(a_pool + (lds_u8(a_dir + (unsigned)((y >> 5) * tw + (x >> 5))) * 32u + (unsigned)(y & 31)) * 4u, 1u << (x & 31));
        else atomicOr(&bits[o >> 5], 1u << (o & 31));
    };
    if (!tile_alloc(lane == 0, sx, sy)) return -1;
    if (lane == 0) {
        const unsigned pk = ((unsigned)sy << 16) | (unsigned)sx;
        reg[0] = pk;
        sts_u32(a_ring, pk);
        mark(sx, sy, so);
        if (touched) touched[nt] = pk;
        rec[so].claim = (unsigned)ticket & 0xffffu;
    }
    if (touched) nt++;
    __syncwarp();
    int n = 1, i = 0, have = 0;
    bool any = false;
    const int b = lane >> 3, k8 = lane & 7, kk = k8 + (k8 >= 4 ? 1 : 0);
    const int ddy = kk / 3 - 1, ddx = kk - (kk / 3) * 3 - 1;
    // a candidate: packed position, pixel offset, and its record (angle | claim | cosf | sinf); angle kNotDefDeg = not a candidate
    unsigned c_pk = 0, c_o = 0, n_pk = 0, n_o = 0;
    float4 c_r = make_float4(kNotDefDeg, 0.f, 0.f, 0.f), n_r = c_r;
#define PL_LOAD_CAND(ri, pk_, o_, r_)                                                                                          \
    {                                                                                                                           \
        const unsigned p_ = (n - (ri) <= kRegRing) ? lds_u32(a_ring + (unsigned)((ri) & (kRegRing - 1)) * 4u) : reg[ri];        \
        const int xx_ = (int)(p_ & 0xffffu) + ddx, yy_ = (int)(p_ >> 16) + ddy;                                                 \
        pk_ = ((unsigned)yy_ << 16) | ((unsigned)xx_ & 0xffffu);                                                                \
        o_ = 0;                                                                                                                 \
        r_.x = kNotDefDeg;                                                                                                      \
        if ((unsigned)xx_ < (unsigned)W && (unsigned)yy_ < (unsigned)H) {                                                       \
            o_ = (unsigned)yy_ * (unsigned)W + (unsigned)xx_;                                                                   \
            r_ = *reinterpret_cast<const float4*>(rec + o_);                                                                    \
        }                                                                                                                       \
    }
    #pragma unroll 1
    while (i < n) {
        const int nb = min(4, n - i);
        // the frontier points of this batch that were not loaded ahead, then those of the next batch that exist already (their loads
        // overlap the resolution of this one)
        if (b < have) { c_pk = n_pk; c_o = n_o; c_r = n_r; }
        if (have < nb) {
            if (b >= have && b < nb) PL_LOAD_CAND(i + b, c_pk, c_o, c_r)
        }
        if (b >= nb) c_r.x = kNotDefDeg;
        const int have_next = max(0, min(4, n - (i + 4)));
        if (have_next > 0) {
            if (b < have_next) PL_LOAD_CAND(i + 4 + b, n_pk, n_o, n_r)
        }
        const int cx = (int)(c_pk & 0xffffu), cy = (int)(c_pk >> 16);
        bool cand = false;
        if (c_r.x != kNotDefDeg) {
            cand = ((lds_u32(a_used + (c_o >> 5) * 4u) >> (c_o & 31)) & 1u) == 0;
            if (cand) {
                if (kSparse) {
                    const unsigned d = lds_u8(a_dir + (unsigned)((cy >> 5) * tw + (cx >> 5)));
                    if (d != 0xffu) cand = ((lds_u32(a_pool + (d * 32u + (unsigned)(cy & 31)) * 4u) >> (cx & 31)) & 1u) == 0;
                } else {
                    cand = ((bits[c_o >> 5] >> (c_o & 31)) & 1u) == 0;
                }
            }
        }
        unsigned rem = __ballot_sync(FULL, cand);
        if (rem) {
            // lanes on the same pixel (two frontier points share a neighbour)
            const unsigned grp = nb > 1 ? __match_any_sync(FULL, cand ? c_o : (0x80000000u | (unsigned)lane)) : (1u << lane);
            const unsigned claim = __float_as_uint(c_r.y) & 0xffffu;
            const int commit_head = (int)lds_u32(a_head);
            #pragma unroll 1
            while (rem) {
                const bool inrem = (rem >> lane) & 1u;
                // 1. hypothesis
                float t = fabsf(hint - c_r.x);
                if (t > 270.f) t = fabsf(t - 360.f);
                const unsigned H0 = __ballot_sync(FULL, inrem && t <= precdeg);
                const bool inH = ((H0 >> lane) & 1u) && !(grp & H0 & lt);  // a pixel seen by several lanes is accepted by the first
                const unsigned H2 = __ballot_sync(FULL, inH);
                const int c = __popc(H2 & lt), cmax = __popc(H2);
                if (inH) sts_f2(a_sval + (unsigned)c * 8u, c_r.z, c_r.w);
                __syncwarp();
                // 2. the sums the reference holds on reaching this lane, if H2 happened
                float px = sumdx, py = sumdy;
                #pragma unroll 1
                for (int k0 = 0; k0 < cmax; k0 += 2) {
                    const float2 v0 = lds_f2(a_sval + (unsigned)k0 * 8u), v1 = lds_f2(a_sval + (unsigned)k0 * 8u + 8u);
                    if (k0 < c) { px = __fadd_rn(px, v0.x); py = __fadd_rn(py, v0.y); }
                    if (k0 + 1 < c) { px = __fadd_rn(px, v1.x); py = __fadd_rn(py, v1.y); }
                }
                __syncwarp();
                // the region angle is only defined by the sums after the first acceptance; before it, it is the seed's
                const float th = (any || c > 0) ? fast_atan2_deg(py, px) : seed_deg;
                bool v = false;
                if (inrem && !(grp & H2 & lt)) {
                    // isAligned() on the float degree values the reference converts to double radians: decided in float when the
                    // distance is not within 2e-3 degrees of the tolerance or of the fold point, else with the reference's own formula
                    const float t2 = fabsf(th - c_r.x);
                    const float tf = t2 > 270.f ? fabsf(t2 - 360.f) : t2;
                    if (__builtin_expect(fabsf(tf - precdeg) > 2e-3f && fabsf(t2 - 270.f) > 2e-3f, 1)) v = tf <= precdeg;
                    else v = lsd_aligned((double)th * kDegToRad, (double)c_r.x * kDegToRad, prec);
                }
                // 3. first lane whose verdict contradicts the hypothesis
                const unsigned M = __ballot_sync(FULL, inrem && v != inH);
                unsigned T = H2, resolved = FULL;
                int hl = 31 - __clz(rem);
                if (__builtin_expect(M != 0, 0)) {
                    const int ls = __ffs(M) - 1;
                    const unsigned below = (1u << ls) - 1u;
                    T = (H2 & below) | (((H2 >> ls) & 1u) ? 0u : (1u << ls));
                    resolved = (2u << ls) - 1u;
                    hl = ls;
                }
                hint = __shfl_sync(FULL, th, hl);
                const int cnt = __popc(T);
                if (cnt) {
                    const bool mine = (T >> lane) & 1u;
                    if (__builtin_expect(n + cnt > reg_cap || (touched && nt + cnt > touched_cap), 0)) return -1;
                    if (__builtin_expect(__any_sync(FULL, mine && lsd_claim_hit(ticket, commit_head, claim)), 0)) return -2;
                    if (__builtin_expect(!tile_alloc(mine, cx, cy), 0)) return -1;
                    if (mine) {
                        rec[c_o].claim = (unsigned)ticket & 0xffffu;
                        const int r = __popc(T & lt);
                        reg[n + r] = c_pk;
                        sts_u32(a_ring + (unsigned)((n + r) & (kRegRing - 1)) * 4u, c_pk);
                        mark(cx, cy, c_o);
                        if (touched) touched[nt + r] = c_pk;
                    }
                    // sums after the last accepted lane: its own prefix plus its own pixel
                    const int L = 31 - __clz(T);
                    sumdx = __shfl_sync(FULL, __fadd_rn(px, c_r.z), L);
                    sumdy = __shfl_sync(FULL, __fadd_rn(py, c_r.w), L);
                    n += cnt;
                    if (touched) nt += cnt;
                    any = true;
                    rem &= ~__ballot_sync(FULL, (grp & T) != 0);
                }
                rem &= ~resolved;
            }
        }
        __syncwarp();
        i += nb;
        have = have_next;
    }
#undef PL_LOAD_CAND
    nt_io = nt;
    *out_angle = any ? (double)fast_atan2_deg_ool(sumdy, sumdx) * kDegToRad : (double)seed_deg * kDegToRad;
    return n;
}
__device__ __forceinline__ int lsd_region_grow(const LsdFrame& Fin, int sx, int sy, double prec, double* out_angle, int& nt) {
    return Fin.sparse ? lsd_region_grow_t<true>(Fin, sx, sy, prec, out_angle, nt) : lsd_region_grow_t<false>(Fin, sx, sy, prec, out_angle, nt);
}

// sequential (reference-order) accumulation helper: every lane loads one region point, then all lanes replay the
// 32 points in order through shuffles, so floating-point sums are formed exactly as the scalar loop forms them.
struct RegPt { double x, y, w; float adeg; };
__device__ __forceinline__ RegPt lsd_load_pt(const LsdFrame& F, int idx, int n) {
    RegPt r;
    r.x = r.y = r.w = 0;
    r.adeg = 0;
    if (idx < n) {
        const unsigned p = F.reg[idx];
        const int x = (int)(p & 0xffffu), y = (int)(p >> 16);
        r.x = (double)x;
        r.y = (double)y;
        const size_t o = (size_t)y * F.W + x;
        r.w = sqrt((double)F.g2[o] / 4.0);  // modgrad
        r.adeg = F.ang[o];
    }
    return r;
}

// Ordered sums of region2rect / refine.  The reference adds the terms of a region point by point, so every sum is a serial chain of
// rounded additions — but the (up to three) sums of a pass are independent chains.  The 32 lanes compute the terms of 32 points and
// park them in shared memory (scratch[k][lane]); then lane k alone walks chain k: one shared-memory load and one addition per point
// instead of a shuffle per term and lane.  `acc` is the running sum of chain `lane` (lanes >= 3 carry nothing).
__device__ __forceinline__ void lsd_ordered_add32(unsigned a_scratch, int lane, double t0, double t1, double t2, double& acc) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a_scratch + (unsigned)lane * 8u), "d"(t0) : "memory");
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a_scratch + 256u + (unsigned)lane * 8u), "d"(t1) : "memory");
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a_scratch + 512u + (unsigned)lane * 8u), "d"(t2) : "memory");
    __syncwarp();
    if (lane < 3) {
        const unsigned a = a_scratch + (unsigned)lane * 256u;
        #pragma unroll 1
        for (int j = 0; j < 32; j += 4) {
            double v0, v1, v2, v3;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(a + (unsigned)j * 8u) : "memory");
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v2), "=d"(v3) : "r"(a + (unsigned)j * 8u + 16u) : "memory");
            acc = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(acc, v0), v1), v2), v3);
        }
    }
    __syncwarp();
}

// region2rect() + get_theta()
__device__ __noinline__ void lsd_region2rect(const LsdFrame& Fin, int n, double reg_angle, double prec, double p, LsdRect& rec) {
    const LsdFrame F = Fin;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const unsigned a_scratch = smem_u32(F.scratch);
    double acc = 0;  // lane 0: sum x * w, lane 1: sum y * w, lane 2: sum w
    #pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        const RegPt pt = lsd_load_pt(F, base + lane, n);  // lanes beyond n hold zeros: adding +0.0 changes nothing
        lsd_ordered_add32(a_scratch, lane, __dmul_rn(pt.x, pt.w), __dmul_rn(pt.y, pt.w), pt.w, acc);
    }
    const double sum = __shfl_sync(FULL, acc, 2);
    const double x = __shfl_sync(FULL, acc, 0) / sum, y = __shfl_sync(FULL, acc, 1) / sum;
    // get_theta
    acc = 0;  // lane 0: Ixx, lane 1: Iyy, lane 2: Ixy (the reference subtracts its terms: a - b == a + (-b))
    #pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        const RegPt pt = lsd_load_pt(F, base + lane, n);
        double txx = 0, tyy = 0, txy = 0;
        if (base + lane < n) {
            const double dx = __dsub_rn(pt.x, x), dy = __dsub_rn(pt.y, y);
            txx = __dmul_rn(__dmul_rn(dy, dy), pt.w);
            tyy = __dmul_rn(__dmul_rn(dx, dx), pt.w);
            txy = -__dmul_rn(__dmul_rn(dx, dy), pt.w);
        }
        lsd_ordered_add32(a_scratch, lane, txx, tyy, txy, acc);
    }
    const double Ixx = __shfl_sync(FULL, acc, 0), Iyy = __shfl_sync(FULL, acc, 1), Ixy = __shfl_sync(FULL, acc, 2);
    const double dI = __dsub_rn(Ixx, Iyy);
    const double lambda = __dmul_rn(0.5, __dsub_rn(__dadd_rn(Ixx, Iyy), sqrt(__dadd_rn(__dmul_rn(dI, dI), __dmul_rn(__dmul_rn(4.0, Ixy), Ixy)))));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_deg((float)__dsub_rn(lambda, Ixx), (float)Ixy)
                                           : (double)fast_atan2_deg((float)Ixy, (float)__dsub_rn(lambda, Iyy));
    theta *= kDegToRad;
    if (fabs(lsd_angle_diff_signed(theta, reg_angle)) > prec) theta += kPiD;
    double dx, dy;
    // the host libm's own arithmetic: the last bit of dx / dy decides which pixels the rectangle's edges include
    {
        const GlibcSinCos sc{g_sincostab};
        dx = sc.cos(theta);
        dy = sc.sin(theta);
    }
    double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
#pragma unroll 1
    for (int idx = lane; idx < n; idx += 32) {
        const unsigned pp = F.reg[idx];
        const double rdx = __dsub_rn((double)(int)(pp & 0xffffu), x), rdy = __dsub_rn((double)(int)(pp >> 16), y);
        const double l = __dadd_rn(__dmul_rn(rdx, dx), __dmul_rn(rdy, dy));
        const double w = __dadd_rn(__dmul_rn(-rdx, dy), __dmul_rn(rdy, dx));
        l_max = fmax(l_max, l); l_min = fmin(l_min, l);
        w_max = fmax(w_max, w); w_min = fmin(w_min, w);
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        l_max = fmax(l_max, __shfl_xor_sync(FULL, l_max, o));
        l_min = fmin(l_min, __shfl_xor_sync(FULL, l_min, o));
        w_max = fmax(w_max, __shfl_xor_sync(FULL, w_max, o));
        w_min = fmin(w_min, __shfl_xor_sync(FULL, w_min, o));
    }
    rec.x1 = __dadd_rn(x, __dmul_rn(l_min, dx)); rec.y1 = __dadd_rn(y, __dmul_rn(l_min, dy));
    rec.x2 = __dadd_rn(x, __dmul_rn(l_max, dx)); rec.y2 = __dadd_rn(y, __dmul_rn(l_max, dy));
    rec.width = __dsub_rn(w_max, w_min);
    rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
    if (rec.width < 1.0) rec.width = 1.0;
}

__device__ __forceinline__ double lsd_density(int n, const LsdRect& rec) {
    return (double)n / (sqrt(lsd_dist_sq(rec.x1, rec.y1, rec.x2, rec.y2)) * rec.width);
}

// reduce_region_radius(): sequential swap-with-last removal keeps the reference's point order
__device__ __noinline__ bool lsd_reduce_region_radius(const LsdFrame& Fin, int& n, double reg_angle, double prec, double p, LsdRect& rec,
                                         double density, double density_th) {
    const LsdFrame F = Fin;
    const int lane = threadIdx.x & 31;
    const unsigned p0 = F.reg[0];
    const double xc = (double)(int)(p0 & 0xffffu), yc = (double)(int)(p0 >> 16);
    const double r1 = lsd_dist_sq(xc, yc, rec.x1, rec.y1), r2 = lsd_dist_sq(xc, yc, rec.x2, rec.y2);
    double radSq = r1 > r2 ? r1 : r2;
    while (density < density_th) {
        radSq *= 0.75 * 0.75;
        if (lane == 0) {
            int m = n;
            for (int i = 0; i < m; ++i) {
                const unsigned q = F.reg[i];
                const int qx = (int)(q & 0xffffu), qy = (int)(q >> 16);
                if (lsd_dist_sq(xc, yc, (double)qx, (double)qy) > radSq) {
                    lsd_unmark(F, qx, qy, (unsigned)qy * (unsigned)F.W + (unsigned)qx);
                    F.reg[i] = F.reg[m - 1];
                    F.reg[m - 1] = q;
                    --m;
                    --i;
                }
            }
            n = m;
        }
        n = __shfl_sync(0xffffffffu, n, 0);
        __syncwarp();
        if (n < 2) return false;
        lsd_region2rect(Fin, n, reg_angle, prec, p, rec);
        density = lsd_density(n, rec);
    }
    return true;
}

// refine()
// returns 1 = refined, 0 = rejected, -1 = capacity exceeded (speculative growers)
// F: register copy for the inlined loops; Fm: the same view in (shared) memory, handed to the out-of-line callees
__device__ int lsd_refine(const LsdFrame& F, const LsdFrame& Fm, int& n, double& reg_angle, double prec, double p, LsdRect& rec, double density_th, int& nt) {
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    double density = lsd_density(n, rec);
    if (density >= density_th) return 1;
    const unsigned p0 = F.reg[0];
    const int sx = (int)(p0 & 0xffffu), sy = (int)(p0 >> 16);
    const double xc = (double)sx, yc = (double)sy;
    const double ang_c = (double)F.ang[(size_t)sy * F.W + sx] * kDegToRad;
    const unsigned a_scratch = smem_u32(F.scratch);
    double acc = 0;  // lane 0: sum of the angle differences, lane 1: sum of their squares
    int cnt_in = 0;
    #pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        const RegPt pt = lsd_load_pt(F, base + lane, n);
        bool inside = false;
        double ang_d = 0, ang_d2 = 0;
        if (base + lane < n) {
            lsd_unmark(F, (int)pt.x, (int)pt.y, (unsigned)(int)pt.y * (unsigned)F.W + (unsigned)(int)pt.x);
            inside = sqrt(lsd_dist_sq(xc, yc, pt.x, pt.y)) < rec.width;
            if (inside) {
                ang_d = lsd_angle_diff_signed((double)pt.adeg * kDegToRad, ang_c);
                ang_d2 = __dmul_rn(ang_d, ang_d);
            }
        }
        cnt_in += __popc(__ballot_sync(FULL, inside));
        lsd_ordered_add32(a_scratch, lane, ang_d, ang_d2, 0.0, acc);  // points outside contribute +0.0
    }
    __syncwarp();
    const double sum = __shfl_sync(FULL, acc, 0), s_sum = __shfl_sync(FULL, acc, 1);
    const double mean_angle = sum / (double)cnt_in;
    const double tau = 2.0 * sqrt(__dadd_rn(__dsub_rn(s_sum, __dmul_rn(__dmul_rn(2.0, mean_angle), sum)) / (double)cnt_in,
                                            __dmul_rn(mean_angle, mean_angle)));
    if (F.touched == nullptr) {
        // the first growth was not logged (its log is the region): the region list is about to be overwritten, so the log starts
        // now, with the region as its first n entries
        if (n > F.touched_cap) return -1;  // (kStCapacity)
        #pragma unroll 1
        for (int i = lane; i < n; i += 32) F.touched_buf[i] = F.reg[i];
        nt = n;
        if (lane == 0) const_cast<LsdFrame&>(Fm).touched = F.touched_buf;
        __syncwarp();
    }
    n = lsd_region_grow(Fm, sx, sy, tau, &reg_angle, nt);
    if (n < 0) return n;
    if (n < 2) return 0;
    lsd_region2rect(Fm, n, reg_angle, prec, p, rec);
    density = lsd_density(n, rec);
    if (density < density_th) return lsd_reduce_region_radius(Fm, n, reg_angle, prec, p, rec, density, density_th) ? 1 : 0;
    return 1;
}

// ---- NFA ----
// log_gamma(x) of lsd.cpp for integer-valued x = m + 1 comes from a host-built table (computed with the same
// Windschitl / Lanczos formulas through the host libm, i.e. the exact doubles the CPU path uses); only pixel counts
// beyond the table fall back to the device formula.
constexpr int kLgMax = 16384;
constexpr int kPMax = 12;  // p = 0.125 * 2^-j, j < kPMax
struct NfaTabs {
    const double* lgam;  // lgam[m] = log_gamma(m + 1)
    double logp[kPMax], log1mp[kPMax], log10p[kPMax];
};
__device__ __noinline__ double lsd_log_gamma_direct(double x) {
    if (x > 15.0) {  // Windschitl
        return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
    }
    const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
    double a = (x + 0.5) * log(x + 5.5) - (x + 5.5);
    double b = 0;
    for (int n = 0; n < 7; ++n) {
        a -= log(x + (double)n);
        b += q[n] * pow(x, (double)n);
    }
    return a + log(b);
}
__device__ __forceinline__ double lsd_lgam1(const NfaTabs& T, int m) {  // log_gamma(double(m) + 1)
    return m < kLgMax ? T.lgam[m] : lsd_log_gamma_direct((double)m + 1);
}
__device__ __forceinline__ bool lsd_double_equal(double a, double b) {
    if (a == b) return true;
    const double abs_diff = fabs(a - b), aa = fabs(a), bb = fabs(b);
    double abs_max = aa > bb ? aa : bb;
    if (abs_max < 2.2250738585072014e-308) abs_max = 2.2250738585072014e-308;
    return (abs_diff / abs_max) <= (100.0 * 2.2204460492503131e-16);
}
// nfa(): pj indexes p = 0.125 * 2^-pj
__device__ __noinline__ double lsd_nfa(const NfaTabs& T, int n, int k, double p, int pj, double log_nt) {
    if (n == 0 || k == 0) return -log_nt;
    const double lp = pj < kPMax ? T.logp[pj] : log(p), l1p = pj < kPMax ? T.log1mp[pj] : log(1.0 - p);
    if (n == k) return -log_nt - (double)n * (pj < kPMax ? T.log10p[pj] : log10(p));
    const double p_term = p / (1 - p);
    const double log1term = lsd_lgam1(T, n) - lsd_lgam1(T, k) - lsd_lgam1(T, n - k) + (double)k * lp + (double)(n - k) * l1p;
    double term = exp(log1term);
    if (lsd_double_equal(term, 0)) {
        if (k > n * p) return -log1term / 2.30258509299404568402 - log_nt;
        return -log_nt;
    }
    double bin_tail = term;
    const double tolerance = 0.1;
    for (int i = k + 1; i <= n; ++i) {
        const double bin_term = (double)(n - i + 1) / (double)i;
        const double mult_term = bin_term * p_term;
        term *= mult_term;
        bin_tail += term;
        if (bin_term < 1) {
            // The reference evaluates pow() and log10() in every one of these iterations.  Both are decided without them
            // almost always, with the same outcome:
            //  * p <= 1/8, so mult_term < 1/7 here and pow(mult_term, m) < 2^-56 for m >= 20: 1 - pow(...) rounds to exactly 1;
            //  * the stopping test compares err with 0.1 * |-log10(bin_tail) - log_nt| * bin_tail: a bracket of log10 that is
            //    1e-7 wide (exponent + lg2.approx of the mantissa) settles it unless err is within 1e-5 of the bound, and only
            //    then is the bound evaluated as the reference writes it.
            const int m = n - i + 1;
            const double one_minus_pow = m >= 20 ? 1.0 : 1 - pow(mult_term, (double)m);
            const double err = term * (one_minus_pow / (1 - mult_term) - 1);
            const int hi = __double2hiint(bin_tail), ex = (hi >> 20) & 0x7ff;
            if (ex != 0 && ex != 0x7ff) {
                const float mant = (float)__hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(bin_tail));   // [1, 2]
                const double l10 = ((double)(ex - 1023) + (double)__log2f(mant)) * 0.30102999566398120;
                const double a = fabs(-l10 - log_nt), base = tolerance * bin_tail;
                if (a > 2e-5) {
                    if (err < base * (a - 1e-5)) break;
                    if (err > base * (a + 1e-5)) continue;
                }
            }
            if (err < tolerance * fabs(-log10(bin_tail) - log_nt) * bin_tail) break;
        }
    }
    return -log10(bin_tail) - log_nt;
}

__device__ __forceinline__ double lsd_slope(double px, double py, double qx, double qy) {
    return ((int)ceil(py) == (int)ceil(qy)) ? 0.0 : (qx - px) / (qy - py);
}

// The pixel scan of rect_nfa() (OpenCV >= 4.5: real-valued corner scan).  `grp` lanes starting at lane `g0` scan the
// rectangle (rows round-robin over the group); every lane returns its partial counts: total points and, for each
// of the `np` precisions, aligned points.
struct ScanCounts { int total; int alg[5]; };
// (not inlined: rect_improve calls it from six places, and the kernel is bound by instruction fetch)
__device__ __noinline__ ScanCounts lsd_rect_scan(const LsdFrame& F, const LsdRect& rec, const double* precs, int np, int sub, int grp) {
    ScanCounts c;
    c.total = 0;
#pragma unroll
    for (int t = 0; t < 5; t++) c.alg[t] = 0;
    const double half_width = rec.width / 2.0;
    const double dyhw = rec.dy * half_width, dxhw = rec.dx * half_width;
    const double vx[4] = {rec.x1 - dyhw, rec.x2 - dyhw, rec.x2 + dyhw, rec.x1 + dyhw};
    const double vy[4] = {rec.y1 + dxhw, rec.y2 + dxhw, rec.y2 - dxhw, rec.y1 - dxhw};
    int offset = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (vy[i] < vy[offset] || (vy[i] == vy[offset] && vx[i] < vx[offset])) offset = i;
    double ox[4], oy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ox[i] = vx[(i + offset) & 3];
        oy[i] = vy[(i + offset) & 3];
    }
    const double flstep = lsd_slope(ox[0], oy[0], ox[1], oy[1]), slstep = lsd_slope(ox[1], oy[1], ox[2], oy[2]);
    const double frstep = lsd_slope(ox[0], oy[0], ox[3], oy[3]), srstep = lsd_slope(ox[3], oy[3], ox[2], oy[2]);
    const int y_begin = (int)ceil(oy[0]), y_end = (int)ceil(oy[2]);
    const int c1 = (int)ceil(oy[1]), c3 = (int)ceil(oy[3]);
    const int nrows = y_end - y_begin + 1;
    // few rows: the group walks along x inside each row; many rows: one row per lane
    const bool by_rows = nrows >= grp;
    for (int yb = y_begin; yb <= y_end; yb += (by_rows ? grp : 1)) {
        const int y = by_rows ? yb + sub : yb;
        if (y > y_end || y < 0 || y >= F.H) continue;
        const double left_limit = (y <= c1) ? ox[0] + ((double)y - oy[0]) * flstep : ox[1] + ((double)y - oy[1]) * slstep;
        const double right_limit = (y < c3) ? ox[0] + ((double)y - oy[0]) * frstep : ox[3] + ((double)y - oy[3]) * srstep;
        int xs = (int)ceil(left_limit), xe = (int)right_limit;
        xs = max(xs, 0);
        xe = min(xe, F.W - 1);
        if (xe < xs) continue;
        const float* arow = F.ang + (size_t)y * F.W;
        if (by_rows || sub == 0) c.total += xe - xs + 1;
        for (int x = xs + (by_rows ? 0 : sub); x <= xe; x += (by_rows ? 1 : grp)) {
            const float a = arow[x];
            if (a == kNotDefDeg) continue;
            double n = rec.theta - (double)a * kDegToRad;
            if (n < 0) n = -n;
            if (n > k32Pi) {
                n -= k2Pi;
                if (n < 0) n = -n;
            }
#pragma unroll
            for (int t = 0; t < 5; t++)
                if (t < np) c.alg[t] += (n <= precs[t]);
        }
    }
    return c;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// rect_nfa() of one rectangle with the whole warp
__device__ double lsd_rect_nfa(const LsdFrame& F, const NfaTabs& T, const LsdRect& rec, int pj, double log_nt) {
    const int lane = threadIdx.x & 31;
    const double pr[1] = {rec.prec};
    ScanCounts c = lsd_rect_scan(F, rec, pr, 1, lane, 32);
    const int total = warp_sum(c.total), alg = warp_sum(c.alg[0]);
    return lsd_nfa(T, total, alg, rec.p, pj, log_nt);
}

// rect_improve().  The five trials of a stage are known up front (the trial rectangle evolves regardless of
// acceptance), so a stage evaluates them concurrently — precision stages with one scan counting all five
// tolerances, geometry stages with five lane groups — and then replays the reference's sequential acceptance.
// The stages are separate, non-inlined functions: the kernel is bound by instruction fetch, not by call overhead.
struct ImproveState {
    LsdRect rec;
    int pj;  // rec.p == 0.125 * 2^-pj
    double log_nfa;
};
// precision stage (used twice): r.p halves five times
__device__ __noinline__ void lsd_improve_precision(const LsdFrame& F, const NfaTabs& T, ImproveState& S, double log_nt, bool need_width) {
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const double delta = 0.5;
    LsdRect r = S.rec;
    const int rj = S.pj;
    if (need_width && !((r.width - delta) >= 0.5)) return;  // the last stage tries only while the width allows it
    double ps[5], precs[5];
#pragma unroll
    for (int n = 0; n < 5; ++n) {
        r.p /= 2;
        ps[n] = r.p;
        precs[n] = r.p * kPiD;
    }
    ScanCounts c = lsd_rect_scan(F, S.rec, precs, 5, lane, 32);
    const int total = warp_sum(c.total);
    int alg[5];
#pragma unroll
    for (int n = 0; n < 5; ++n) alg[n] = warp_sum(c.alg[n]);
    // lanes 0..4 evaluate the five NFAs concurrently
    double v = 0;
    {
        int myk = 0;
        double myp = 0;
#pragma unroll
        for (int n = 0; n < 5; ++n)
            if (lane == n) { myk = alg[n]; myp = ps[n]; }
        if (lane < 5) v = lsd_nfa(T, total, myk, myp, rj + lane + 1, log_nt);
    }
    for (int n = 0; n < 5; ++n) {
        const double vn = __shfl_sync(FULL, v, n);
        if (vn > S.log_nfa) {
            S.log_nfa = vn;
            S.rec.p = ps[n];
            S.rec.prec = precs[n];
            S.pj = rj + n + 1;
        }
    }
}
// geometry stage: mode 0 = reduce width, 1 = reduce one side, 2 = reduce the other side
__device__ __noinline__ void lsd_improve_geometry(const LsdFrame& F, const NfaTabs& T, ImproveState& S, double log_nt, int mode) {
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const double delta = 0.5, delta_2 = delta / 2.0;
    LsdRect r = S.rec;
    LsdRect mine = S.rec;  // trial rectangle of this lane's group
    const int grp = lane / 6, sub = lane - grp * 6;  // groups 0..4 (lanes 30,31 idle)
    int ntrial = 0;
    for (int n = 0; n < 5; ++n) {
        if ((r.width - delta) >= 0.5) {
            if (mode == 1) {
                r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2;
                r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2;
            } else if (mode == 2) {
                r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2;
                r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2;
            }
            r.width -= delta;
            if (grp == ntrial) mine = r;
            ntrial++;
        }
    }
    if (ntrial == 0) return;
    const double pr[1] = {S.rec.prec};
    ScanCounts c;
    c.total = 0;
    c.alg[0] = 0;
    if (grp < ntrial) c = lsd_rect_scan(F, mine, pr, 1, sub, 6);
    // reduce inside each 6-lane group: both counts of a lane in one word (a rectangle has far fewer than 2^31 pixels and the sums
    // are formed in 64 bits), three shuffles to the group's first lane, then lane t fetches the sums of group t
    unsigned long long both = ((unsigned long long)(unsigned)c.total << 32) | (unsigned)c.alg[0];
    both += __shfl_down_sync(FULL, both, 3);                                                 // sub 0..2 hold pairs (sub, sub + 3)
    both += __shfl_down_sync(FULL, both, 1) + __shfl_down_sync(FULL, both, 2);               // sub 0: all six
    both = __shfl_sync(FULL, both, (lane < 5 ? lane : 0) * 6);
    const int myn = (int)(both >> 32), myk = (int)(unsigned)both;
    double v = 0;
    if (lane < ntrial) v = lsd_nfa(T, myn, myk, S.rec.p, S.pj, log_nt);
    // replay: trial t's rectangle lives in group t
    for (int t = 0; t < ntrial; ++t) {
        const double vt = __shfl_sync(FULL, v, t);
        if (vt > S.log_nfa) {
            S.log_nfa = vt;
            const int src = t * 6;
            S.rec.x1 = __shfl_sync(FULL, mine.x1, src); S.rec.y1 = __shfl_sync(FULL, mine.y1, src);
            S.rec.x2 = __shfl_sync(FULL, mine.x2, src); S.rec.y2 = __shfl_sync(FULL, mine.y2, src);
            S.rec.width = __shfl_sync(FULL, mine.width, src);
        }
    }
}
__device__ double lsd_rect_improve(const LsdFrame& F, const NfaTabs& T, LsdRect& rec, double log_nt, double log_eps) {
    ImproveState S;
    S.rec = rec;
    S.pj = 0;
    S.log_nfa = lsd_rect_nfa(F, T, rec, 0, log_nt);
    if (S.log_nfa > log_eps) return S.log_nfa;
    lsd_improve_precision(F, T, S, log_nt, false);
    for (int mode = 0; mode < 3 && !(S.log_nfa > log_eps); mode++) lsd_improve_geometry(F, T, S, log_nt, mode);
    if (!(S.log_nfa > log_eps)) lsd_improve_precision(F, T, S, log_nt, true);
    rec = S.rec;
    return S.log_nfa;
}

// flsd() main loop.  Persistent CTAs take frames from a counter; the G warps of a CTA are region growers.
//
// Region growing is ordered (seeds by gradient bin, shared USED map), but regions that do not touch the same pixels
// commute.  The kernel runs the ordered loop as a window of speculative transactions with in-order commit:
//   * select (one warp at a time): the next seed that is unused in the committed map gets the next ticket and a
//     slot of the window (kSlots tickets may be uncommitted at once).  A seed stamped by an uncommitted ticket is
//     most likely being swallowed by that region: it gets its ticket but is not grown ("deferred");
//   * grow (any number of warps at once): the warp grows its seed reading the committed USED map (global bytes),
//     keeps the pixels it marks in its private bitmap in shared memory and logs every pixel it ever accepts (initial
//     growth and refine re-growth) in the touched list of a buffer from the CTA's pool; the final region, the fitted
//     rectangle and the log stay in the buffer, the private bitmap is cleared, and the warp goes for the next seed.
//     A grower that runs into a pixel stamped by an earlier uncommitted ticket gives up (deferred);
//   * commit (one warp at a time, strictly in ticket order): seed already committed -> the seed is void (an earlier
//     region swallowed it);  deferred, or a logged pixel already committed (the growth depended on a region that
//     was committed after it read the map) -> the committing warp grows the seed now, when everything before it is
//     committed, with the frame-sized buffers;  otherwise the growth is exactly what the sequential algorithm would
//     have done.  The region is written to the USED map and its rectangle is queued.
// Committed pixels are never released, tickets follow the seed order, and a grower only ever sees pixels of
// earlier tickets in the map, so the result is the sequential one whatever the stamps say.  rect_improve only reads
// the angle map and does not influence later regions: it runs afterwards in k_lsd_nfa, and accepted segments are
// compacted in seed order.
struct LsdQueueItem { LsdRect rec; };
constexpr int kNfaChunk = 16;         // rectangles per work item of the tail helpers
constexpr int kSmBusySlots = 512;     // per-SM counters behind nfa_ctl (indexed by %smid)
constexpr int kNfaChunksPerFrame = 2048;  // item = frame * kNfaChunksPerFrame + chunk
constexpr uint8_t kNfaTodo = 0xff;    // qvalid: rectangle not validated yet
// rect_improve of rectangle t of frame f -> qres / qvalid (one warp)
__device__ __noinline__ void lsd_nfa_one(const LineGeom& g, const float* __restrict__ ang, const LsdQueueItem* __restrict__ queue,
                                         LsdSeg* __restrict__ qres, uint8_t* __restrict__ qvalid, const NfaTabs& T, int f, int t) {
    const int lane = threadIdx.x & 31;
    LsdFrame F;
    F.ang = ang;
    F.g2 = nullptr; F.rec = nullptr; F.cs0 = nullptr; F.sval = nullptr; F.used_bits = nullptr; F.reg = nullptr; F.ring = nullptr; F.scratch = nullptr;
    F.W = g.W; F.H = g.H;
    F.sparse = false; F.dir = nullptr; F.rev = nullptr; F.pool = nullptr; F.ntiles = nullptr; F.tw = 0; F.pool_tiles = 0;
    F.bits = nullptr; F.touched = nullptr; F.touched_buf = nullptr; F.reg_cap = 0; F.touched_cap = 0;
    F.ticket = 0; F.commit_head = nullptr;
    const double log_eps = 0.0;
    // (the tail helpers read rectangles another SM queued during the same launch, chunk by chunk: a neighbour of an earlier chunk may
    // sit in this SM's L1 from before it was written, so the words come from L2)
    LsdRect rec;
    {
        const unsigned int* src = reinterpret_cast<const unsigned int*>(&queue[(size_t)f * g.seg_cap + t].rec);
        unsigned int* dst = reinterpret_cast<unsigned int*>(&rec);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(LsdRect) / 4); i++) dst[i] = __ldcg(src + i);
    }
    const double log_nfa = lsd_rect_improve(F, T, rec, g.log_nt, log_eps);
    if (lane == 0) {
        LsdSeg sg;
        sg.x1 = (float)((rec.x1 + 0.5) / 0.8); sg.y1 = (float)((rec.y1 + 0.5) / 0.8);
        sg.x2 = (float)((rec.x2 + 0.5) / 0.8); sg.y2 = (float)((rec.y2 + 0.5) / 0.8);
        sg.width = rec.width / 0.8;
        sg.p = rec.p;
        sg.nfa = log_nfa;
        qres[(size_t)f * g.seg_cap + t] = sg;
        qvalid[(size_t)f * g.seg_cap + t] = log_nfa > log_eps;
    }
    __syncwarp();
}

constexpr int kSmall = 64;         // regions up to this size (and log length) are parked in their slot's small buffer instead
constexpr int kTiny = 16;          // regions without a rectangle up to this size never leave shared memory: the points sit in their ticket slot
constexpr int kSvalEntries = 36;
constexpr int kMaxPoolTiles = 64, kMinPoolTiles = 24;
// ticket slot (shared memory, int4): x = seed pixel, y = region size, z = touched-log size,
// w = state | (status + 2) << 8 | (buffer + 1) << 16   (buffer -1 with status >= 0: the slot's small buffer)
enum { kSlotFree = 0, kSlotReady = 1, kSlotGrowing = 2, kSlotDone = 3 };
enum { kStDeferred = -2, kStCapacity = -1, kStNoRect = 0, kStRect = 1 };
__device__ __forceinline__ int slot_pack(int state, int status, int buf) { return state | ((status + 2) << 8) | ((buf + 1) << 16); }
struct GrowResult {
    int status, n, nt;
    LsdRect rec;
};
__device__ __forceinline__ int pool_pop(unsigned long long* mask, int lane) {
    int b = -1;
    if (lane == 0) {
        while (true) {
            const unsigned long long m = *(volatile unsigned long long*)mask;
            if (!m) break;
            const int c = __ffsll((long long)m) - 1;
            if (atomicCAS(mask, m, m & ~(1ull << c)) == m) { b = c; break; }
        }
    }
    return __shfl_sync(0xffffffffu, b, 0);
}
// grow + fit + refine one seed into the buffers of F; leaves the private marks clean
// Fin and out live in shared memory (one per warp): the out-of-line callees read the view from there instead of
// every thread keeping (and spilling) its own copy in local memory.
// region2rect + refine of a region of at least min_reg_size pixels (n, reg_angle, nt: in / out); returns the status and leaves
// the rectangle in out->rec.  Out of line and apart from the growth: only one region in seven gets here.
__device__ __noinline__ int lsd_fit_refine(const LsdFrame& Fin, int* n_io, double reg_angle, int* nt_io, GrowResult* out) {
    const LsdFrame F = Fin;  // register copy for the inlined code
    const double prec = kPiD * 22.5 / 180, p = 22.5 / 180;
    const double density_th = 0.7;
    int n = *n_io, nt = *nt_io;
    LsdRect rec;
    lsd_region2rect(Fin, n, reg_angle, prec, p, rec);
    const int status = lsd_refine(F, Fin, n, reg_angle, prec, p, rec, density_th, nt);
    if ((threadIdx.x & 31) == 0 && status == kStRect) out->rec = rec;
    *n_io = n;
    *nt_io = nt;
    return status;
}
__device__ __noinline__ void lsd_grow_seed(const LsdFrame& Fin, int pix, int min_reg_size, GrowResult* out) {
    const int lane = threadIdx.x & 31;
    const double prec = kPiD * 22.5 / 180;
    const int sx = pix % Fin.W, sy = pix / Fin.W;
    double reg_angle;
    int nt = 0, status = kStNoRect;
    int n = lsd_region_grow(Fin, sx, sy, prec, &reg_angle, nt);
    if (n < 0) status = n;
    else if (n >= min_reg_size) status = lsd_fit_refine(Fin, &n, reg_angle, &nt, out);
    lsd_priv_reset(Fin, nt);
    if (lane == 0) {
        out->status = status;
        out->n = n;
        out->nt = nt;
    }
    __syncwarp();
}

// device buffers of k_lsd_grow
struct GrowBufs {
    const float* angdeg;
    const int* g2;
    LsdPix* rec;
    const float2* cs0;
    const unsigned int* seeds;
    const int* n_seeds;
    unsigned int* big_reg;       // [frame][plane]
    unsigned int* big_touched;   // [cta][frame slot][2 * plane]
    unsigned int* big_bits;      // [cta][frame slot][bits_words], all-zero between uses
    unsigned int* pool_reg;      // [cta][kPool][kSpecCap]
    unsigned int* pool_touched;  // [cta][kPool][kSpecCap]
    LsdRect* pool_rect;          // [cta][kPool]
    unsigned int* small_buf;     // [cta][frame slot][kSlots][reg kSmall | touched kSmall]
    LsdRect* small_rect;         // [cta][frame slot][kSlots]
    LsdQueueItem* queue;         // [frame][seg_cap]
    int* n_rects;
    int* flags;
    long long* phase_cycles;
    int* frame_counter;
    size_t plane;
    // tail helpers: finished frames publish chunks of their rectangles; CTAs without frames validate them
    LsdSeg* qres;
    uint8_t* qvalid;
    unsigned int* nfa_items;  // [nf * kNfaChunksPerFrame] 0xffffffff = not published yet
    int* nfa_ctl;             // [0] items published (tail), [1] items taken (cursor), [2] frames finished, [4 + smid] CTAs of that SM that still grow
    NfaTabs nfa_tabs;
};

}  // namespace pl
#include "lsd_grow2.cuh"
namespace pl {

// NFA validation of the fitted rectangles (rect_improve): it only reads the angle map and does not influence any
// other region, so every rectangle of every frame is independent — one warp per rectangle.  Rectangles are validated
// in two places: grower CTAs that have run out of frames take chunks of kNfaChunk rectangles of finished frames from a
// queue while the last frames are still being grown (the tail of k_lsd_grow would otherwise leave most SMs idle), and
// k_lsd_nfa afterwards does whatever is left (qvalid == kNfaTodo).
constexpr int kNfaBlocksPerFrame = 32, kNfaThreads = 256;
__global__ void __launch_bounds__(kNfaThreads, 4) k_lsd_nfa(LineGeom g, const float* __restrict__ angdeg, size_t plane,
                                                         const LsdQueueItem* __restrict__ queue, const int* __restrict__ n_rects,
                                                         LsdSeg* __restrict__ qres, uint8_t* __restrict__ qvalid, NfaTabs T, int* __restrict__ next) {
    // rect_improve costs between one and 26 rectangle scans: the warps of a frame take rectangles from a counter (four at a time)
    // instead of a fixed stride, so that no warp is left with the expensive ones
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int n = min(n_rects[f], g.seg_cap);
    while (true) {
        int t0 = 0;
        if (lane == 0) t0 = atomicAdd(next + f, 4);
        t0 = __shfl_sync(0xffffffffu, t0, 0);
        if (t0 >= n) break;
        for (int t = t0; t < min(t0 + 4, n); t++)
            if (qvalid[(size_t)f * g.seg_cap + t] == kNfaTodo) lsd_nfa_one(g, angdeg + (size_t)f * plane, queue, qres, qvalid, T, f, t);
    }
}

}  // namespace pl

// =================================================================================================================
// KeyLine assembly + top-N by response, LBD
// =================================================================================================================
namespace pl {

// LSDDetector::detectImpl: clamp extremes, fill KeyLine (opencv_contrib LSDDetector.cpp)
__device__ __forceinline__ pl_keyline make_keyline(const LsdSeg& s, int cols, int rows, int class_id) {
    float e0 = s.x1, e1 = s.y1, e2 = s.x2, e3 = s.y2;
    if (e0 < 0) e0 = 0;
    if (e0 >= cols) e0 = (float)cols - 1.0f;
    if (e2 < 0) e2 = 0;
    if (e2 >= cols) e2 = (float)cols - 1.0f;
    if (e1 < 0) e1 = 0;
    if (e1 >= rows) e1 = (float)rows - 1.0f;
    if (e3 < 0) e3 = 0;
    if (e3 >= rows) e3 = (float)rows - 1.0f;
    pl_keyline kl;
    kl.sx = e0; kl.sy = e1; kl.ex = e2; kl.ey = e3;  // octaveScale = 1
    kl.sx_oct = e0; kl.sy_oct = e1; kl.ex_oct = e2; kl.ey_oct = e3;
    const double ddx = (double)__fsub_rn(e0, e2), ddy = (double)__fsub_rn(e1, e3);
    kl.length = (float)sqrt(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)));
    const int x0 = cv_round(e0), y0 = cv_round(e1), x1 = cv_round(e2), y1 = cv_round(e3);
    kl.num_pixels = max(abs(x1 - x0), abs(y1 - y0)) + 1;  // cv::LineIterator::count, 8-connected, inside the image
    kl.angle = (float)atan2((double)__fsub_rn(kl.ey, kl.sy), (double)__fsub_rn(kl.ex, kl.sx));
    kl.class_id = class_id;
    kl.octave = 0;
    kl.size = __fmul_rn(__fsub_rn(kl.ex, kl.sx), __fsub_rn(kl.ey, kl.sy));
    kl.response = __fdiv_rn(kl.length, (float)max(cols, rows));
    kl.pt_x = __fdiv_rn(__fadd_rn(kl.ex, kl.sx), 2.f);
    kl.pt_y = __fdiv_rn(__fadd_rn(kl.ey, kl.sy), 2.f);
    return kl;
}

// LineExtractor.cpp:23-35: if more than max_lines, keep the max_lines largest responses (ties: detection order),
// ordered by response descending; otherwise keep detection order.  One CTA per frame.
constexpr int kFinThreads = 256;
constexpr int kMaxKeep = 512;
__global__ void __launch_bounds__(kFinThreads) k_line_finalize(LineGeom g, const LsdSeg* __restrict__ qres, const uint8_t* __restrict__ qvalid,
                                                               const int* __restrict__ n_rects, LsdSeg* __restrict__ segs,
                                                               int* __restrict__ n_segs, float* __restrict__ resp_scratch,
                                                               pl_keyline* __restrict__ kls, int* __restrict__ n_out, int cap) {
    __shared__ int s_hist[4096];
    __shared__ unsigned long long s_key[kMaxKeep];
    __shared__ int s_sel, s_need, s_bin, s_nseg;
    const int f = blockIdx.x, tid = threadIdx.x;
    // accepted segments (NFA > 0) compacted in seed order == cv::LineSegmentDetector::detect output order
    if (tid < 32) {
        const int total = min(n_rects[f], g.seg_cap);
        const LsdSeg* qr = qres + (size_t)f * g.seg_cap;
        const uint8_t* qv = qvalid + (size_t)f * g.seg_cap;
        LsdSeg* o = segs + (size_t)f * g.seg_cap;
        int cnt = 0;
        for (int base = 0; base < total; base += 32) {
            const bool v = (base + tid < total) && qv[base + tid];
            const unsigned m = __ballot_sync(0xffffffffu, v);
            if (v) o[cnt + __popc(m & ((1u << tid) - 1u))] = qr[base + tid];
            cnt += __popc(m);
        }
        if (tid == 0) { n_segs[f] = cnt; s_nseg = cnt; }
    }
    __syncthreads();
    const int n = s_nseg;
    const LsdSeg* sg = segs + (size_t)f * g.seg_cap;
    pl_keyline* out = kls + (size_t)f * cap;
    const int keep = min(g.max_lines, cap);
    if (n <= keep) {
        for (int i = tid; i < n; i += kFinThreads) out[i] = make_keyline(sg[i], g.cols, g.rows, i);
        if (tid == 0) n_out[f] = n;
        return;
    }
    float* resp = resp_scratch + (size_t)f * g.seg_cap;
    for (int i = tid; i < n; i += kFinThreads) resp[i] = make_keyline(sg[i], g.cols, g.rows, i).response;
    __syncthreads();
    // radix select of the keep-th largest response on the float bits (non-negative floats order like unsigned ints)
    unsigned prefix = 0, pmask = 0;
    int need = keep;  // how many elements are still to be taken from the current candidate set
    const int shifts[3] = {20, 8, 0};
    const int widths[3] = {12, 12, 8};
    for (int lvl = 0; lvl < 3; lvl++) {
        const int nb = 1 << widths[lvl];
        for (int i = tid; i < nb; i += kFinThreads) s_hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += kFinThreads) {
            const unsigned b = __float_as_uint(resp[i]);
            if ((b & pmask) == prefix) atomicAdd(&s_hist[(b >> shifts[lvl]) & (nb - 1)], 1);
        }
        __syncthreads();
        if (tid < 32) {
            // the bin in which the running count from the top reaches `need`: every lane sums a block of bins, a warp scan finds the
            // block, its lane walks it (one thread over 4096 bins was a third of this kernel's time)
            const int per = nb / 32, hi = nb - 1 - tid * per;  // lane 0 owns the top block
            int sum = 0;
            for (int b = 0; b < per; b++) sum += s_hist[hi - b];
            int above = sum;  // inclusive scan over the lanes (lane order = descending bins)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, above, o);
                if (tid >= o) above += v;
            }
            const int before = above - sum;  // elements in the blocks above this lane's
            if (before < need && above >= need) {  // exactly one lane (need >= 1 and the total is >= need)
                int acc = before, bin = hi;
                for (; bin > hi - per; bin--) {
                    if (acc + s_hist[bin] >= need) break;
                    acc += s_hist[bin];
                }
                s_bin = bin;
                s_need = need - acc;
            }
        }
        __syncthreads();
        prefix |= (unsigned)s_bin << shifts[lvl];
        pmask |= (unsigned)(nb - 1) << shifts[lvl];
        need = s_need;
        __syncthreads();
    }
    // prefix == bits of the keep-th largest response; take all strictly larger + the first `need` equal ones (by index)
    const unsigned thr = prefix;
    if (tid == 0) s_sel = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kFinThreads) {
        const unsigned b = __float_as_uint(resp[i]);
        if (b > thr) {
            const int k = atomicAdd(&s_sel, 1);
            s_key[k] = ((unsigned long long)(0xFFFFFFFFu - b) << 32) | (unsigned)i;
        }
    }
    __syncthreads();
    // the first `need` responses equal to the threshold, in index order.  Almost always there are exactly `need` of them (one): all
    // threads look, and only a real tie between more than `need` equal responses is walked by one thread
    const int n_greater = s_sel;
    __syncthreads();
    if (tid == 0) s_need = 0;  // (reused as the count of equal responses)
    __syncthreads();
    for (int i = tid; i < n; i += kFinThreads)
        if (__float_as_uint(resp[i]) == thr) {
            const int k = atomicAdd(&s_need, 1);
            if (k < need) s_key[n_greater + k] = ((unsigned long long)(0xFFFFFFFFu - thr) << 32) | (unsigned)i;
        }
    __syncthreads();
    if (tid == 0) {
        if (s_need > need) {  // exact float ties: lowest indices first
            int k = n_greater, taken = 0;
            for (int i = 0; i < n && taken < need; i++)
                if (__float_as_uint(resp[i]) == thr) {
                    s_key[k++] = ((unsigned long long)(0xFFFFFFFFu - thr) << 32) | (unsigned)i;
                    taken++;
                }
        }
        s_sel = n_greater + need;
    }
    __syncthreads();
    const int m = s_sel;  // == keep
    int m2 = 1;
    while (m2 < m) m2 <<= 1;
    for (int i = m + tid; i < m2; i += kFinThreads) s_key[i] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= m2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < m2; i += kFinThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a = s_key[i], b = s_key[ixj];
                    if (up ? (a > b) : (a < b)) { s_key[i] = b; s_key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < m; i += kFinThreads) {
        const int idx = (int)(s_key[i] & 0xFFFFFFFFull);
        out[i] = make_keyline(sg[idx], g.cols, g.rows, idx);
    }
    if (tid == 0) n_out[f] = m;
}

// GaussianBlur(5x5, sigma 1) in 8.8 fixed point {14,62,104,62,14} (BinaryDescriptor::computeGaussianPyramid)
__global__ void __launch_bounds__(256) k_blur5(LineGeom g, const uint8_t* __restrict__ in, size_t in_pitch, size_t in_frame_stride,
                                               uint8_t* __restrict__ out) {
    constexpr int TW = 64, TH = 32;
    __shared__ uint8_t s_in[(TH + 4) * (TW + 4)];
    __shared__ uint16_t s_h[(TH + 4) * TW];
    const int f = blockIdx.z, tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const uint8_t* src = in + (size_t)f * in_frame_stride;
    for (int i = tid; i < (TH + 4) * (TW + 4); i += 256) {
        const int r = i / (TW + 4), c = i - r * (TW + 4);
        s_in[i] = __ldg(src + (size_t)reflect101(y0 + r - 2, g.rows) * in_pitch + reflect101(x0 + c - 2, g.cols));
    }
    __syncthreads();
    for (int i = tid; i < (TH + 4) * TW; i += 256) {
        const int r = i / TW, c = i - r * TW;
        const uint8_t* p = s_in + r * (TW + 4) + c;
        s_h[i] = (uint16_t)(14 * (p[0] + p[4]) + 62 * (p[1] + p[3]) + 104 * p[2]);
    }
    __syncthreads();
    uint8_t* dst = out + (size_t)f * g.in_pitch * g.rows;
    for (int i = tid; i < TH * TW; i += 256) {
        const int r = i / TW, c = i - r * TW;
        const int x = x0 + c, y = y0 + r;
        if (x < g.cols && y < g.rows) {
            const uint16_t* p = s_h + r * TW + c;
            const uint32_t acc = 14u * (p[0] + p[4 * TW]) + 62u * (p[TW] + p[3 * TW]) + 104u * p[2 * TW];
            dst[(size_t)y * g.in_pitch + x] = (uint8_t)((acc + 0x8000u) >> 16);
        }
    }
}

// cv::Sobel(img, CV_16S, 1,0,3) and (0,1,3), BORDER_REFLECT_101 (BinaryDescriptor::computeSobel)
__global__ void __launch_bounds__(256) k_sobel3(LineGeom g, const uint8_t* __restrict__ blur, short* __restrict__ dxo, short* __restrict__ dyo) {
    const int f = blockIdx.z;
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= g.cols || y >= g.rows) return;
    const uint8_t* b = blur + (size_t)f * g.in_pitch * g.rows;
    const uint8_t* r0 = b + (size_t)reflect101(y - 1, g.rows) * g.in_pitch;
    const uint8_t* r1 = b + (size_t)y * g.in_pitch;
    const uint8_t* r2 = b + (size_t)reflect101(y + 1, g.rows) * g.in_pitch;
    const int xm = reflect101(x - 1, g.cols), xp = reflect101(x + 1, g.cols);
    const int gx = (r0[xp] + 2 * r1[xp] + r2[xp]) - (r0[xm] + 2 * r1[xm] + r2[xm]);
    const int gy = (r2[xm] + 2 * r2[x] + r2[xp]) - (r0[xm] + 2 * r0[x] + r0[xp]);
    const size_t o = (size_t)f * g.dpitch * g.rows + (size_t)y * g.dpitch + x;
    dxo[o] = (short)gx;
    dyo[o] = (short)gy;
}

// ---- the two kernels above fused, the source tile staged by TMA ----
// GaussianBlur(5x5) and both Sobel(3x3) derivatives in one pass: a CTA owns a 64 x 32 tile of the image; one elected thread issues a
// cp.async.bulk.tensor load of the (64 + 6) x (32 + 6) source tile (box 96 x 38 x 1 of the (cols, rows, frames) tensor: out-of-range
// elements arrive as zeros) and everybody waits on the mbarrier; the edge CTAs then mirror the border into the halo
// (BORDER_REFLECT_101 of the blur; the Sobel's own reflection of the BLURRED image is the blur of the reflected source because
// the kernel is symmetric, so the same halo serves both); horizontal and vertical blur passes through shared memory; each thread
// writes eight dx and eight dy values with one 16-byte store each.  The blurred image never goes to HBM: 1 byte read and 4 bytes
// written per pixel instead of 2 + 5.
// (the box starts at a multiple of 16 bytes in x — the TMA unit faults on an unaligned inner coordinate — so the tile carries 16
// columns of halo on the left, of which 3 are used)
constexpr int kFsTW = 64, kFsTH = 32, kFsBoxW = 96, kFsXOff = 16, kFsBoxH = kFsTH + 6;
__device__ __forceinline__ void mbar_wait_parity(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__global__ void __launch_bounds__(256) k_blur5_sobel3_tma(LineGeom g, const __grid_constant__ CUtensorMap tmap, short* __restrict__ dxo,
                                                          short* __restrict__ dyo) {
    __shared__ __align__(128) uint8_t s_in[kFsBoxH * kFsBoxW];      // source tile, origin (x0 - 16, y0 - 3)
    __shared__ uint16_t s_h[kFsBoxH * (kFsTW + 4)];                 // horizontal pass, columns x0 - 1 .. x0 + 64
    __shared__ __align__(16) uint8_t s_b[(kFsTH + 2) * (kFsTW + 8)];  // blurred tile, origin (x0 - 1, y0 - 1)
    __shared__ __align__(8) unsigned long long s_bar;
    const int f = blockIdx.z, tid = threadIdx.x;
    const int x0 = blockIdx.x * kFsTW, y0 = blockIdx.y * kFsTH;
    const unsigned bar = smem_u32(&s_bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kFsBoxH * kFsBoxW) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(s_in)),
                     "l"(&tmap), "r"(bar), "r"(x0 - kFsXOff), "r"(y0 - 3), "r"(f)
                     : "memory");
    }
    __syncthreads();  // (the barrier's initialisation is visible to everybody before anybody waits on it)
    mbar_wait_parity(bar, 0);
    // BORDER_REFLECT_101: only the CTAs on the image's edge have anything to mirror
    const bool edge_x = x0 == 0 || x0 + kFsTW + 3 > g.cols, edge_y = y0 == 0 || y0 + kFsTH + 3 > g.rows;
    if (edge_x) {
        for (int i = tid; i < kFsBoxH * (kFsTW + 6); i += 256) {
            const int r = i / (kFsTW + 6), c = i - r * (kFsTW + 6);
            const int x = x0 - 3 + c;
            if (x < 0 || x >= g.cols) s_in[r * kFsBoxW + kFsXOff - 3 + c] = s_in[r * kFsBoxW + (reflect101(x, g.cols) - (x0 - kFsXOff))];
        }
        __syncthreads();
    }
    if (edge_y) {
        for (int i = tid; i < kFsBoxH * (kFsTW + 6); i += 256) {
            const int r = i / (kFsTW + 6), c = i - r * (kFsTW + 6);
            const int y = y0 - 3 + r;
            if (y < 0 || y >= g.rows) s_in[r * kFsBoxW + kFsXOff - 3 + c] = s_in[(reflect101(y, g.rows) - (y0 - 3)) * kFsBoxW + kFsXOff - 3 + c];
        }
        __syncthreads();
    }
    // horizontal pass {14, 62, 104, 62, 14}: two outputs per step from one 8-byte window
    for (int i = tid; i < kFsBoxH * ((kFsTW + 2) / 2); i += 256) {
        const int r = i / ((kFsTW + 2) / 2), c = (i - r * ((kFsTW + 2) / 2)) * 2;
        const uint8_t* p = s_in + r * kFsBoxW + kFsXOff - 3 + c;
        const unsigned a0 = 14u * (p[0] + p[4]) + 62u * (p[1] + p[3]) + 104u * p[2];
        const unsigned a1 = 14u * (p[1] + p[5]) + 62u * (p[2] + p[4]) + 104u * p[3];
        *reinterpret_cast<unsigned*>(s_h + r * (kFsTW + 4) + c) = a0 | (a1 << 16);
    }
    __syncthreads();
    // vertical pass -> blurred bytes (rounded as cv::GaussianBlur's fixed-point path rounds)
    for (int i = tid; i < (kFsTH + 2) * ((kFsTW + 2) / 2); i += 256) {
        const int r = i / ((kFsTW + 2) / 2), c = (i - r * ((kFsTW + 2) / 2)) * 2;
        const uint16_t* p = s_h + r * (kFsTW + 4) + c;
        constexpr int P = kFsTW + 4;
        const unsigned v0 = *reinterpret_cast<const unsigned*>(p), v1 = *reinterpret_cast<const unsigned*>(p + P),
                       v2 = *reinterpret_cast<const unsigned*>(p + 2 * P), v3 = *reinterpret_cast<const unsigned*>(p + 3 * P),
                       v4 = *reinterpret_cast<const unsigned*>(p + 4 * P);
        const unsigned lo = 14u * ((v0 & 0xffffu) + (v4 & 0xffffu)) + 62u * ((v1 & 0xffffu) + (v3 & 0xffffu)) + 104u * (v2 & 0xffffu);
        const unsigned hi = 14u * ((v0 >> 16) + (v4 >> 16)) + 62u * ((v1 >> 16) + (v3 >> 16)) + 104u * (v2 >> 16);
        uint8_t* q = s_b + r * (kFsTW + 8) + c;
        q[0] = (uint8_t)((lo + 0x8000u) >> 16);
        q[1] = (uint8_t)((hi + 0x8000u) >> 16);
    }
    __syncthreads();
    // Sobel: thread = eight consecutive pixels of one row
    {
        const int r = tid >> 3, c = (tid & 7) * 8;
        const int y = y0 + r, x = x0 + c;
        if (y < g.rows && x < g.cols) {
            const uint8_t* b0 = s_b + r * (kFsTW + 8) + c;  // blurred (x - 1, y - 1)
            const uint8_t *b1 = b0 + (kFsTW + 8), *b2 = b1 + (kFsTW + 8);
            short gx[8], gy[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                gx[k] = (short)((b0[k + 2] + 2 * b1[k + 2] + b2[k + 2]) - (b0[k] + 2 * b1[k] + b2[k]));
                gy[k] = (short)((b2[k] + 2 * b2[k + 1] + b2[k + 2]) - (b0[k] + 2 * b0[k + 1] + b0[k + 2]));
            }
            const size_t o = (size_t)f * g.dpitch * g.rows + (size_t)y * g.dpitch + x;
            if (x + 8 <= g.cols) {
                *reinterpret_cast<int4*>(dxo + o) = *reinterpret_cast<const int4*>(gx);
                *reinterpret_cast<int4*>(dyo + o) = *reinterpret_cast<const int4*>(gy);
            } else {
                for (int k = 0; k < 8 && x + k < g.cols; k++) {
                    dxo[o + k] = gx[k];
                    dyo[o + k] = gy[k];
                }
            }
        }
    }
}

constexpr int kLbdBands = 9, kLbdBandW = 7, kLbdRows = kLbdBands * kLbdBandW;  // 63
struct LbdTabs { float gaussL[kLbdBandW * 3]; float gaussG[kLbdRows]; };

// computeLBD(), part 1: one thread per (line, row of the line support region): projected-gradient row sums,
// accumulated in the reference's order (left to right along the line)
__global__ void __launch_bounds__(64) k_lbd_rows(LineGeom g, LbdTabs tabs, const pl_keyline* __restrict__ kls, const int* __restrict__ n_lines,
                                                 int cap, const short* __restrict__ dxi, const short* __restrict__ dyi,
                                                 float* __restrict__ rowsum) {
    const int f = blockIdx.y, line = blockIdx.x, hID = threadIdx.x;
    if (line >= n_lines[f] || hID >= kLbdRows) return;
    const pl_keyline kl = kls[(size_t)f * cap + line];
    const short* pdx = dxi + (size_t)f * g.dpitch * g.rows;
    const short* pdy = dyi + (size_t)f * g.dpitch * g.rows;
    const short heightOfLSP = (short)kLbdRows;
    const short halfHeight = (heightOfLSP - 1) / 2;
    const short imageWidth = (short)(g.cols - 1), imageHeight = (short)(g.rows - 1);
    const short lengthOfLSP = (short)kl.num_pixels;
    const short halfWidth = (lengthOfLSP - 1) / 2;
    const float midX = __fdiv_rn(__fadd_rn(kl.sx_oct, kl.ex_oct), 2.f), midY = __fdiv_rn(__fadd_rn(kl.sy_oct, kl.ey_oct), 2.f);
    const float dL0 = (float)cos((double)kl.angle), dL1 = (float)sin((double)kl.angle);
    const float dO0 = -dL1, dO1 = dL0;
    float sCorX0 = __fadd_rn(__fadd_rn(__fmul_rn(-dL0, (float)halfWidth), __fmul_rn(dL1, (float)halfHeight)), midX);
    float sCorY0 = __fadd_rn(__fsub_rn(__fmul_rn(-dL1, (float)halfWidth), __fmul_rn(dL0, (float)halfHeight)), midY);
    for (int h = 0; h < hID; h++) {
        sCorX0 = __fsub_rn(sCorX0, dL1);
        sCorY0 = __fadd_rn(sCorY0, dL0);
    }
    float sCorX = sCorX0, sCorY = sCorY0;
    float pgdL = 0, ngdL = 0, pgdO = 0, ngdO = 0;
    for (short wID = 0; wID < lengthOfLSP; wID++) {
        short t = (short)(int)round((double)sCorX);
        const short xCor = (t < 0) ? 0 : (t > imageWidth) ? imageWidth : t;
        t = (short)(int)round((double)sCorY);
        const short yCor = (t < 0) ? 0 : (t > imageHeight) ? imageHeight : t;
        const float ddx = (float)pdx[(size_t)yCor * g.dpitch + xCor], ddy = (float)pdy[(size_t)yCor * g.dpitch + xCor];
        const float gDL = __fadd_rn(__fmul_rn(ddx, dL0), __fmul_rn(ddy, dL1));
        const float gDO = __fadd_rn(__fmul_rn(ddx, dO0), __fmul_rn(ddy, dO1));
        if (gDL > 0) pgdL = __fadd_rn(pgdL, gDL); else ngdL = __fsub_rn(ngdL, gDL);
        if (gDO > 0) pgdO = __fadd_rn(pgdO, gDO); else ngdO = __fsub_rn(ngdO, gDO);
        sCorX = __fadd_rn(sCorX, dL0);
        sCorY = __fadd_rn(sCorY, dL1);
    }
    const float coef = tabs.gaussG[hID];
    float* o = rowsum + (((size_t)f * cap + line) * kLbdRows + hID) * 4;
    o[0] = __fmul_rn(coef, pgdL);
    o[1] = __fmul_rn(coef, ngdL);
    o[2] = __fmul_rn(coef, pgdO);
    o[3] = __fmul_rn(coef, ngdO);
}

__constant__ int c_lbd_comb[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7},
                                      {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8}, {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

// computeLBD(), part 2 + binaryConversion + line coefficients (LineExtractor.cpp:60-69): one thread per line
__global__ void __launch_bounds__(128) k_lbd_finish(LineGeom g, LbdTabs tabs, const pl_keyline* __restrict__ kls, const int* __restrict__ n_lines,
                                                    int cap, const float* __restrict__ rowsum, uint8_t* __restrict__ desc,
                                                    float* __restrict__ fdesc, double* __restrict__ coeffs) {
    const int f = blockIdx.y, line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_lines[f]) return;
    const float* rs = rowsum + ((size_t)f * cap + line) * kLbdRows * 4;
    float band[kLbdBands][8];
#pragma unroll
    for (int b = 0; b < kLbdBands; b++)
#pragma unroll
        for (int k = 0; k < 8; k++) band[b][k] = 0.f;
    for (int hID = 0; hID < kLbdRows; hID++) {
        const float pgdL = rs[hID * 4], ngdL = rs[hID * 4 + 1], pgdO = rs[hID * 4 + 2], ngdO = rs[hID * 4 + 3];
        const float pgdL2 = __fmul_rn(pgdL, pgdL), ngdL2 = __fmul_rn(ngdL, ngdL), pgdO2 = __fmul_rn(pgdO, pgdO), ngdO2 = __fmul_rn(ngdO, ngdO);
        const int bandID = hID / kLbdBandW, r = hID % kLbdBandW;
#pragma unroll
        for (int t = 0; t < 3; t++) {  // current band, band above, band below — in the reference's order
            const int bb = t == 0 ? bandID : (t == 1 ? bandID - 1 : bandID + 1);
            if (bb < 0 || bb >= kLbdBands) continue;
            const float c = tabs.gaussL[t == 0 ? r + kLbdBandW : (t == 1 ? r + 2 * kLbdBandW : r)];
            const float cc = __fmul_rn(c, c);
            float* B = band[bb];
            B[0] = __fadd_rn(B[0], __fmul_rn(c, pgdL));
            B[1] = __fadd_rn(B[1], __fmul_rn(c, ngdL));
            B[2] = __fadd_rn(B[2], __fmul_rn(cc, pgdL2));
            B[3] = __fadd_rn(B[3], __fmul_rn(cc, ngdL2));
            B[4] = __fadd_rn(B[4], __fmul_rn(c, pgdO));
            B[5] = __fadd_rn(B[5], __fmul_rn(c, ngdO));
            B[6] = __fadd_rn(B[6], __fmul_rn(cc, pgdO2));
            B[7] = __fadd_rn(B[7], __fmul_rn(cc, ngdO2));
        }
    }
    float d[kLbdBands * 8];
    const float invN2 = (float)(1.0 / (kLbdBandW * 2.0)), invN3 = (float)(1.0 / (kLbdBandW * 3.0));
#pragma unroll
    for (int b = 0; b < kLbdBands; b++) {
        const float invN = (b == 0 || b == kLbdBands - 1) ? invN2 : invN3;
        const float* B = band[b];
        float temp = __fmul_rn(B[0], invN);
        d[b * 8] = temp;
        d[b * 8 + 4] = sqrtf(__fsub_rn(__fmul_rn(B[2], invN), __fmul_rn(temp, temp)));
        temp = __fmul_rn(B[1], invN);
        d[b * 8 + 1] = temp;
        d[b * 8 + 5] = sqrtf(__fsub_rn(__fmul_rn(B[3], invN), __fmul_rn(temp, temp)));
        temp = __fmul_rn(B[4], invN);
        d[b * 8 + 2] = temp;
        d[b * 8 + 6] = sqrtf(__fsub_rn(__fmul_rn(B[6], invN), __fmul_rn(temp, temp)));
        temp = __fmul_rn(B[5], invN);
        d[b * 8 + 3] = temp;
        d[b * 8 + 7] = sqrtf(__fsub_rn(__fmul_rn(B[7], invN), __fmul_rn(temp, temp)));
    }
    float tempM = 0, tempS = 0;
#pragma unroll
    for (int b = 0; b < kLbdBands; b++) {
#pragma unroll
        for (int k = 0; k < 4; k++) tempM = __fadd_rn(tempM, __fmul_rn(d[b * 8 + k], d[b * 8 + k]));
#pragma unroll
        for (int k = 4; k < 8; k++) tempS = __fadd_rn(tempS, __fmul_rn(d[b * 8 + k], d[b * 8 + k]));
    }
    tempM = __fdiv_rn(1.f, sqrtf(tempM));
    tempS = __fdiv_rn(1.f, sqrtf(tempS));
#pragma unroll
    for (int b = 0; b < kLbdBands; b++) {
#pragma unroll
        for (int k = 0; k < 4; k++) d[b * 8 + k] = __fmul_rn(d[b * 8 + k], tempM);
#pragma unroll
        for (int k = 4; k < 8; k++) d[b * 8 + k] = __fmul_rn(d[b * 8 + k], tempS);
    }
#pragma unroll
    for (int i = 0; i < kLbdBands * 8; i++)
        if ((double)d[i] > 0.4) d[i] = (float)0.4;
    float temp = 0;
#pragma unroll
    for (int i = 0; i < kLbdBands * 8; i++) temp = __fadd_rn(temp, __fmul_rn(d[i], d[i]));
    temp = __fdiv_rn(1.f, sqrtf(temp));
#pragma unroll
    for (int i = 0; i < kLbdBands * 8; i++) d[i] = __fmul_rn(d[i], temp);
    uint8_t* od = desc + ((size_t)f * cap + line) * 32;
    for (int c = 0; c < 32; c++) {
        const float* f1 = d + 8 * c_lbd_comb[c][0];
        const float* f2 = d + 8 * c_lbd_comb[c][1];
        unsigned r = 0;
        for (int i = 0; i < 8; i++) r += (f1[i] > f2[i]) ? (1u << i) : 0u;
        od[c] = (uint8_t)r;
    }
    if (fdesc) {
        float* of = fdesc + ((size_t)f * cap + line) * 72;
        for (int i = 0; i < 72; i++) of[i] = d[i];
    }
    const pl_keyline kl = kls[(size_t)f * cap + line];
    const double sx = kl.sx, sy = kl.sy, ex = kl.ex, ey = kl.ey;
    double l0 = __dsub_rn(sy, ey), l1 = __dsub_rn(ex, sx), l2 = __dsub_rn(__dmul_rn(sx, ey), __dmul_rn(sy, ex));
    const double nrm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(l0, l0), __dmul_rn(l1, l1)), __dmul_rn(l2, l2)));
    if (nrm > 0) { l0 /= nrm; l1 /= nrm; l2 /= nrm; }
    double* oc = coeffs + ((size_t)f * cap + line) * 3;
    oc[0] = l0; oc[1] = l1; oc[2] = l2;
}

}  // namespace pl

// =================================================================================================================
// host side
// =================================================================================================================
using namespace pl;

struct pl_line {
    int device = 0;
    cudaStream_t stream = nullptr;
    int max_cols = 0, max_rows = 0, max_batch = 0;
    int rows = 0, cols = 0, max_lines = 0;
    LineGeom geom;
    LbdTabs tabs;
    size_t plane = 0, scaled_stride = 0;
    // device buffers (sized for max_cols x max_rows x max_batch at creation)
    uint8_t *d_in = nullptr, *d_scaled = nullptr, *d_blur5 = nullptr;
    float* d_ang = nullptr;
    LsdPix* d_rec = nullptr;
    float2* d_cs0 = nullptr;
    int* d_nrects = nullptr;
    int* d_g2 = nullptr;
    unsigned int *d_reg = nullptr, *d_seeds = nullptr;
    int *d_maxg2 = nullptr, *d_tile_off = nullptr, *d_nseeds = nullptr, *d_nsegs = nullptr, *d_flags = nullptr, *d_nout = nullptr;
    unsigned short* d_tile_hist = nullptr;
    LsdSeg* d_segs = nullptr;
    LsdSeg* d_qres = nullptr;
    LsdQueueItem* d_queue = nullptr;
    uint8_t* d_qvalid = nullptr;
    unsigned int *d_spec_reg = nullptr, *d_spec_touched = nullptr, *d_big_touched = nullptr;
    LsdRect *d_pool_rect = nullptr, *d_small_rect = nullptr;
    unsigned int* d_small_buf = nullptr;
    int* d_frame_counter = nullptr;
    int* d_nfa_ctl = nullptr;
    unsigned int* d_nfa_items = nullptr;
    int* d_nfa_next = nullptr;   // per frame: the next rectangle k_lsd_nfa hands out
    cudaEvent_t ev_grow = nullptr;  // recorded where the streaming stages of a chunk end and its region grower is launched
    int* d_sticky = nullptr;  // capacity flags of the device-pointer API since the last pl_line_sync
    int bits_words = 0, num_sms = 0, grow_tiles = 0, grow_window = 128;
    int tail_nfa = 2;  // 0: k_lsd_nfa validates everything afterwards; 1: CTAs that ran out of frames help; 2: ... but only from SMs on which nothing grows any more
    // k_lsd_grow2 (role-specialised grower, one frame per CTA): shape for up to one frame per SM / for more frames than SMs
    struct Grow2Cfg { int threads = 0, occ = 0, pool_tiles = 0, pool_n = 0, split = 0; size_t smem = 0; } g2_few, g2_many;
    int lookahead = 4;
    // single-frame (tracking) mode: the chunk's launches as one CUDA graph, re-used while the call's parameters stay the same
    bool use_graph = false;   // pl_line_set_graph / PLSLAM_LINE_GRAPH=1
    bool capturing = false;
    cudaGraphExec_t graph_exec = nullptr;
    struct GraphKey {
        const void *gray, *kls, *desc, *coef, *nout;
        size_t step, frame_stride;
        int nf, cap, rows, cols, max_lines;
        bool operator==(const GraphKey& o) const {
            return gray == o.gray && kls == o.kls && desc == o.desc && coef == o.coef && nout == o.nout && step == o.step &&
                   frame_stride == o.frame_stride && nf == o.nf && cap == o.cap && rows == o.rows && cols == o.cols && max_lines == o.max_lines;
        }
    } graph_key{};
    int graph_launches = 0;   // kernel launches inside the captured graph
    bool use_tma = true;      // PLSLAM_LINE_TMA=0: the two-kernel blur + Sobel path also for aligned inputs (test hook)
    void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled, through cudaGetDriverEntryPoint
    bool force_many = false;  // test hook (PLSLAM_LSD_FORCE_MANY): the many-frames shape also for small batches
    int poll_ns = 400;
    int reserved_sms = 0;          // SMs the region grower leaves to the kernels of other streams (pl_line_set_reserved_sms)
    unsigned int* d_big_bits = nullptr;
    float *d_resp = nullptr, *d_rowsum = nullptr, *d_fdesc = nullptr;
    short *d_dx = nullptr, *d_dy = nullptr;
    ExactTab *d_xtab = nullptr, *d_ytab = nullptr;
    pl_keyline* d_kls = nullptr;
    uint8_t* d_desc = nullptr;
    double* d_coef = nullptr;
    int out_cap = 0;  // capacity (lines per frame) of d_kls / d_desc / d_coef / d_rowsum
    int seg_cap_alloc = 0, tiles_alloc = 0;
    size_t plane_alloc = 0, in_alloc = 0;
    int* h_flags = nullptr;
    double* d_lgam = nullptr;
    long long* d_phase = nullptr;
    NfaTabs nfa_tabs;
    int last_batch = 0, last_launches = 0;
    bool profiling = false;
    cudaEvent_t ev[8] = {nullptr};
    float stage_ms[8] = {0};
    int stage_chunks = 0;
};

namespace {

inline int cvRoundD(double v) { return (int)lrint(v); }

void build_exact_axis(std::vector<ExactTab>& t, int dn, int sn, double inv_scale) {
    t.assign(dn, ExactTab{0, 256, 0, 0});
    const double scale = 1.0 / inv_scale;
    for (int v = 0; v < dn; v++) {
        double fv = scale * (v + 0.5) - 0.5;
        int i = (int)floor(fv);
        if (i >= 0 && sn > 1) {
            if (i < sn - 1) {
                t[v].ofs = (short)i;
                t[v].c1 = (short)cvRoundD((fv - i) * 256);
                t[v].c0 = (short)(256 - t[v].c1);
            } else {
                t[v].ofs = (short)(sn - 1);
            }
        }
    }
}

template <typename T>
int dev_alloc(T** p, size_t n) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    PL_CUDA_TRY(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    return PL_OK;
}

int line_geometry(pl_line* h, int rows, int cols, int max_lines) {
    if (rows == h->rows && cols == h->cols && max_lines == h->max_lines) return PL_OK;
    LineGeom& G = h->geom;
    G.cols = cols; G.rows = rows;
    G.W = cvRoundD(cols * 0.8);
    G.H = cvRoundD(rows * 0.8);
    if (G.W < 8 || G.H < 8) {
        set_error("image %dx%d too small for LSD", cols, rows);
        return PL_ERR_ARG;
    }
    G.in_pitch = (int)align_up((size_t)cols, 16);
    G.spitch = (int)align_up((size_t)G.W, 16);
    G.n_tiles = (G.H - 1 + kTileRows - 1) / kTileRows;
    G.reg_cap = G.W * G.H;
    G.log_nt = 5 * (log10((double)G.W) + log10((double)G.H)) / 2 + log10(11.0);
    const double p = 22.5 / 180;
    G.min_reg_size = (int)(size_t)(-G.log_nt / log10(p));
    G.seg_cap = std::min(G.W * G.H / std::max(G.min_reg_size, 1) + 1, h->seg_cap_alloc);
    G.max_lines = max_lines;
    G.dpitch = (int)align_up((size_t)cols, 8);
    h->plane = (size_t)G.W * G.H;
    h->scaled_stride = (size_t)G.spitch * G.H;
    std::vector<ExactTab> xt, yt;
    build_exact_axis(xt, G.W, cols, 0.8);
    build_exact_axis(yt, G.H, rows, 0.8);
    PL_CUDA_TRY(cudaMemcpyAsync(h->d_xtab, xt.data(), xt.size() * sizeof(ExactTab), cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(h->d_ytab, yt.data(), yt.size() * sizeof(ExactTab), cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    h->rows = rows; h->cols = cols; h->max_lines = max_lines;
    return PL_OK;
}

int line_ensure_out(pl_line* h, int cap) {
    if (cap <= h->out_cap) return PL_OK;
    const size_t B = h->max_batch;
    int rc;
    if ((rc = dev_alloc(&h->d_kls, B * cap)) != PL_OK) return rc;
    if ((rc = dev_alloc(&h->d_desc, B * cap * 32)) != PL_OK) return rc;
    if ((rc = dev_alloc(&h->d_coef, B * cap * 3)) != PL_OK) return rc;
    if ((rc = dev_alloc(&h->d_rowsum, B * cap * kLbdRows * 4)) != PL_OK) return rc;
    if ((rc = dev_alloc(&h->d_fdesc, B * cap * 72)) != PL_OK) return rc;
    h->out_cap = cap;
    return PL_OK;
}

// TMA descriptor of the caller's image batch as a (cols, rows, frames) tensor of bytes, box kFsBoxW x kFsBoxH x 1; false when the
// layout cannot be described (strides and base must be multiples of 16 bytes)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool line_make_tmap(pl_line* h, CUtensorMap* out, const uint8_t* d_gray, int cols, int rows, int nf, size_t step, size_t frame_stride) {
    if (!h->encode_tiled || (step & 15) || (frame_stride & 15) || ((uintptr_t)d_gray & 15)) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)nf};
    const cuuint64_t strides[2] = {(cuuint64_t)step, (cuuint64_t)frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)kFsBoxW, (cuuint32_t)kFsBoxH, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = ((EncodeTiledFn)h->encode_tiled)(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)d_gray, dims, strides, box, estr,
                                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// one chunk; every pointer is a device pointer
int line_launch_chunk_direct(pl_line* h, const uint8_t* d_gray, int nf, size_t step, size_t frame_stride, pl_keyline* d_kls, uint8_t* d_desc,
                             double* d_coef, int cap, int* d_nout) {
    const LineGeom& G = h->geom;
    cudaStream_t st = h->stream;
    const size_t plane = h->plane;
    int launches = 0;
    const bool prof = h->profiling;
    PL_CUDA_TRY(cudaMemsetAsync(h->d_maxg2, 0xff, sizeof(int) * nf, st));  // -1
    PL_CUDA_TRY(cudaMemsetAsync(h->d_flags, 0, sizeof(int) * nf, st));
    if (prof) cudaEventRecord(h->ev[0], st);
    {
        dim3 grid((G.W + kScTW - 1) / kScTW, (G.H + kScTH - 1) / kScTH, nf);
        k_lsd_scale<<<grid, 256, 0, st>>>(G, d_gray, step, frame_stride, h->d_xtab, h->d_ytab, h->d_scaled, h->scaled_stride);
        launches++;
    }
    {
        const double rho = 2.0 / sin(kPiD * 22.5 / 180);
        dim3 grid((G.W + 63) / 64, (G.H + 3) / 4, nf);
        k_lsd_grad<<<grid, 256, 0, st>>>(G, h->d_scaled, h->scaled_stride, h->d_ang, h->d_g2, reinterpret_cast<float4*>(h->d_rec), h->d_cs0, plane, rho, h->d_maxg2);
        launches++;
    }
    if (prof) cudaEventRecord(h->ev[1], st);
    // (the Sobel planes are written after the region grower: until then d_dx holds the per-pixel bins, W x H <= cols x rows of them per frame)
    unsigned short* d_bins = reinterpret_cast<unsigned short*>(h->d_dx);
    k_lsd_bin_count<<<dim3((G.n_tiles + 7) / 8, nf), 256, 0, st>>>(G, h->d_ang, h->d_g2, plane, h->d_maxg2, h->d_tile_hist, d_bins);
    k_lsd_bin_scan<<<nf, kBins, 0, st>>>(G, h->d_tile_hist, h->d_tile_off, h->d_nseeds);
    k_lsd_scatter<<<dim3((G.n_tiles + 7) / 8, nf), 256, 0, st>>>(G, d_bins, plane, h->d_tile_off, h->d_seeds);
    launches += 3;
    if (prof) cudaEventRecord(h->ev[2], st);
    {
        // One frame per CTA (k_lsd_grow2), several CTAs per SM when there are more frames than SMs.  The CTAs are persistent (frames
        // come from a counter) but do not fill an SM: kernels of other streams run next to them on what is left.  reserved_sms only
        // shrinks the grid.
        const int sms = std::max(1, h->num_sms - (nf > h->num_sms - h->reserved_sms ? h->reserved_sms : 0));
        PL_CUDA_TRY(cudaMemsetAsync(h->d_frame_counter, 0, sizeof(int), st));
        GrowBufs gb;
        gb.angdeg = h->d_ang; gb.g2 = h->d_g2; gb.rec = h->d_rec; gb.cs0 = h->d_cs0; gb.seeds = h->d_seeds; gb.n_seeds = h->d_nseeds;
        gb.big_reg = h->d_reg; gb.big_touched = h->d_big_touched; gb.big_bits = h->d_big_bits;
        gb.pool_reg = h->d_spec_reg; gb.pool_touched = h->d_spec_touched; gb.pool_rect = h->d_pool_rect;
        gb.small_buf = h->d_small_buf; gb.small_rect = h->d_small_rect; gb.queue = h->d_queue; gb.n_rects = h->d_nrects;
        gb.flags = h->d_flags; gb.phase_cycles = prof ? h->d_phase : nullptr; gb.frame_counter = h->d_frame_counter; gb.plane = plane;
        gb.qres = h->d_qres; gb.qvalid = h->d_qvalid; gb.nfa_items = h->d_nfa_items; gb.nfa_ctl = h->d_nfa_ctl; gb.nfa_tabs = h->nfa_tabs;
        PL_CUDA_TRY(cudaMemsetAsync(h->d_nfa_ctl, 0, (4 + kSmBusySlots) * sizeof(int), st));
        PL_CUDA_TRY(cudaMemsetAsync(h->d_nfa_next, 0, sizeof(int) * nf, st));
        PL_CUDA_TRY(cudaMemsetAsync(h->d_nfa_items, 0xff, sizeof(unsigned int) * (size_t)nf * kNfaChunksPerFrame, st));
        PL_CUDA_TRY(cudaMemsetAsync(h->d_qvalid, 0xff, (size_t)nf * G.seg_cap, st));
        if (h->ev_grow && !h->capturing) PL_CUDA_TRY(cudaEventRecord(h->ev_grow, st));  // (an event recorded inside a capture cannot be waited for from outside)
        const bool many2 = (nf > sms || h->force_many) && h->g2_many.threads > 0;
        const pl_line::Grow2Cfg& c2 = many2 ? h->g2_many : h->g2_few;
        Grow2Smem g2{h->grow_tiles, c2.pool_tiles, std::min(h->grow_window, kSlots2), h->lookahead, h->bits_words, h->tail_nfa, h->poll_ns, c2.pool_n, c2.split};
        const int ctas2 = std::min(nf, sms * std::max(1, c2.occ));
        if (c2.threads <= 256 && c2.occ >= 2) k_lsd_grow2<256, 3><<<ctas2, c2.threads, c2.smem, st>>>(G, g2, nf, gb);
        else if (c2.threads <= 384 && c2.occ >= 2) k_lsd_grow2<384, 2><<<ctas2, c2.threads, c2.smem, st>>>(G, g2, nf, gb);
        else if (c2.threads <= 512) k_lsd_grow2<512, 1><<<ctas2, c2.threads, c2.smem, st>>>(G, g2, nf, gb);
        else k_lsd_grow2<1024, 1><<<ctas2, c2.threads, c2.smem, st>>>(G, g2, nf, gb);
    }
    {
        // a warp per rectangle: few frames get more CTAs each, so that a single frame's ~1500 rectangles are one wave on the whole GPU
        const int per_frame = std::max(kNfaBlocksPerFrame, std::min(256, (4 * h->num_sms + nf - 1) / nf));
        if (prof) cudaEventRecord(h->ev[6], st);  // (k_lsd_nfa alone: stage_ms[5])
        k_lsd_nfa<<<dim3(per_frame, nf), kNfaThreads, 0, st>>>(G, h->d_ang, plane, h->d_queue, h->d_nrects, h->d_qres, h->d_qvalid, h->nfa_tabs, h->d_nfa_next);
    }
    launches += 2;
    if (prof) cudaEventRecord(h->ev[3], st);
    k_line_finalize<<<nf, kFinThreads, 0, st>>>(G, h->d_qres, h->d_qvalid, h->d_nrects, h->d_segs, h->d_nsegs, h->d_resp, d_kls, d_nout, cap);
    launches++;
    {
        dim3 grid((G.cols + 63) / 64, (G.rows + 31) / 32, nf);
        CUtensorMap tmap;
        if (h->use_tma && G.cols >= 8 && G.rows >= 8 && line_make_tmap(h, &tmap, d_gray, G.cols, G.rows, nf, step, frame_stride)) {
            k_blur5_sobel3_tma<<<grid, 256, 0, st>>>(G, tmap, h->d_dx, h->d_dy);
            launches += 1;
        } else {  // the caller's pitch or base address is not 16-byte aligned: a tensor map cannot describe it
            k_blur5<<<grid, 256, 0, st>>>(G, d_gray, step, frame_stride, h->d_blur5);
            dim3 grid2((G.cols + 63) / 64, (G.rows + 3) / 4, nf);
            k_sobel3<<<grid2, 256, 0, st>>>(G, h->d_blur5, h->d_dx, h->d_dy);
            launches += 2;
        }
    }
    if (prof) cudaEventRecord(h->ev[4], st);
    const int keep = std::min(G.max_lines, cap);
    k_lbd_rows<<<dim3(keep, nf), 64, 0, st>>>(G, h->tabs, d_kls, d_nout, cap, h->d_dx, h->d_dy, h->d_rowsum);
    k_lbd_finish<<<dim3((keep + 127) / 128, nf), 128, 0, st>>>(G, h->tabs, d_kls, d_nout, cap, h->d_rowsum, d_desc, h->d_fdesc, d_coef);
    launches += 2;
    PL_CUDA_TRY(cudaGetLastError());
    if (prof) {
        cudaEventRecord(h->ev[5], st);
        PL_CUDA_TRY(cudaEventSynchronize(h->ev[5]));
        for (int i = 0; i < 5; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
            h->stage_ms[i] += ms;
        }
        {
            float ms = 0;
            cudaEventElapsedTime(&ms, h->ev[6], h->ev[3]);
            h->stage_ms[5] += ms;
        }
        h->stage_chunks++;
    }
    h->last_batch = nf;
    h->last_launches += launches;
    return PL_OK;
}

// The launches of a small chunk are a fixed sequence of ~16 stream operations whose parameters only depend on the call's pointers
// and sizes: in tracking mode (one frame per call, the same staging buffers every time) they are captured once and replayed as ONE
// graph launch.  Anything that changes the parameters re-captures; profiling and large batches launch directly.
int line_launch_chunk(pl_line* h, const uint8_t* d_gray, int nf, size_t step, size_t frame_stride, pl_keyline* d_kls, uint8_t* d_desc,
                      double* d_coef, int cap, int* d_nout) {
    if (!h->use_graph || h->profiling || nf > 4) return line_launch_chunk_direct(h, d_gray, nf, step, frame_stride, d_kls, d_desc, d_coef, cap, d_nout);
    const pl_line::GraphKey key{d_gray, d_kls, d_desc, d_coef, d_nout, step, frame_stride, nf, cap, h->geom.rows, h->geom.cols, h->geom.max_lines};
    if (h->graph_exec && key == h->graph_key) {
        PL_CUDA_TRY(cudaGraphLaunch(h->graph_exec, h->stream));
        h->last_batch = nf;
        h->last_launches += h->graph_launches;
        return PL_OK;
    }
    if (h->graph_exec) {
        cudaGraphExecDestroy(h->graph_exec);
        h->graph_exec = nullptr;
    }
    PL_CUDA_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    h->capturing = true;
    const int before = h->last_launches;
    const int rc = line_launch_chunk_direct(h, d_gray, nf, step, frame_stride, d_kls, d_desc, d_coef, cap, d_nout);
    h->capturing = false;
    cudaGraph_t graph = nullptr;
    const cudaError_t ee = cudaStreamEndCapture(h->stream, &graph);
    if (rc != PL_OK || ee != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        (void)cudaGetLastError();
        if (rc != PL_OK) return rc;
        // the sequence could not be captured on this driver: launch it directly from now on
        h->use_graph = false;
        h->last_launches = before;
        return line_launch_chunk_direct(h, d_gray, nf, step, frame_stride, d_kls, d_desc, d_coef, cap, d_nout);
    }
    const cudaError_t ei = cudaGraphInstantiate(&h->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) {
        (void)cudaGetLastError();
        h->graph_exec = nullptr;
        h->use_graph = false;
        h->last_launches = before;
        return line_launch_chunk_direct(h, d_gray, nf, step, frame_stride, d_kls, d_desc, d_coef, cap, d_nout);
    }
    h->graph_key = key;
    h->graph_launches = h->last_launches - before;
    PL_CUDA_TRY(cudaGraphLaunch(h->graph_exec, h->stream));
    return PL_OK;
}

__global__ void __launch_bounds__(32) k_line_or_flags(const int* __restrict__ flags, int n, int* __restrict__ sticky) {
    int v = 0;
    for (int i = threadIdx.x; i < n; i += 32) v |= flags[i];
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0 && v) atomicOr(sticky, v);
}

int line_check_flags(pl_line* h, int nf) {
    PL_CUDA_TRY(cudaMemcpyAsync(h->h_flags, h->d_flags, sizeof(int) * nf, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    for (int i = 0; i < nf; i++)
        if (h->h_flags[i]) {
            set_error("frame %d of the chunk exceeded an LSD capacity (flags=%d: 1=segments, 2=region size)", i, h->h_flags[i]);
            return PL_ERR_CAPACITY;
        }
    return PL_OK;
}

}  // namespace

__global__ void __launch_bounds__(256) k_test_sincos(const double* __restrict__ x, int n, double* __restrict__ s, double* __restrict__ c) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        const pl::GlibcSinCos sc{pl::g_sincostab};
        s[i] = sc.sin(x[i]);
        c[i] = sc.cos(x[i]);
    }
}

extern "C" {

PL_API int pl_test_sincos(const double* x, int n, double* s, double* c) {
    PL_CHECK_ARG(x && s && c && n >= 0);
    if (n == 0) return PL_OK;
    double *dx = nullptr, *ds = nullptr, *dc = nullptr;
    cudaError_t e = cudaMalloc((void**)&dx, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ds, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dc, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_test_sincos<<<(n + 255) / 256, 256>>>(dx, n, ds, dc);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(s, ds, sizeof(double) * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(c, dc, sizeof(double) * n, cudaMemcpyDeviceToHost);
    cudaFree(dx);
    cudaFree(ds);
    cudaFree(dc);
    if (e != cudaSuccess) {
        set_error("pl_test_sincos: %s", cudaGetErrorString(e));
        return PL_ERR_CUDA;
    }
    return PL_OK;
}

PL_API int pl_line_create(pl_line** out, int device, int max_cols, int max_rows, int max_batch) {
    PL_CHECK_ARG(out != nullptr);
    *out = nullptr;
    PL_CHECK_ARG(max_cols >= 16 && max_rows >= 16 && max_cols <= 4000 && max_rows <= 4000 && max_batch >= 1);
    int ndev = 0;
    PL_CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) {
        set_error("device %d not available (%d CUDA devices); this library has no CPU fallback", device, ndev);
        return PL_ERR_CUDA;
    }
    PL_CUDA_TRY(cudaSetDevice(device));
    pl_line* h = new pl_line();
    h->device = device;
    h->max_cols = max_cols; h->max_rows = max_rows; h->max_batch = max_batch;
    // BinaryDescriptor constructor: local / global Gaussian weights (integer divisions as in the original)
    {
        double u = (kLbdBandW * 3 - 1) / 2, sigma = (kLbdBandW * 2 + 1) / 2, inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < kLbdBandW * 3; i++) h->tabs.gaussL[i] = (float)exp((i - u) * (i - u) * inv);
        u = (kLbdRows - 1) / 2; sigma = kLbdRows / 2; inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < kLbdRows; i++) h->tabs.gaussG[i] = (float)exp((i - u) * (i - u) * inv);
    }
    const size_t B = max_batch;
    const int W = cvRoundD(max_cols * 0.8) + 1, H = cvRoundD(max_rows * 0.8) + 1;
    const size_t plane = (size_t)W * H;
    const size_t in_pitch = align_up((size_t)max_cols, 16);
    const int tiles = (H + kTileRows - 1) / kTileRows + 1;
    const int seg_cap = (int)(plane / 8) + 1;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_grow, cudaEventDisableTiming);
    auto A = [&](auto** p, size_t n) {
        if (e == cudaSuccess) e = cudaMalloc((void**)p, n * sizeof(**p));
    };
    A(&h->d_in, B * in_pitch * max_rows);
    A(&h->d_blur5, B * in_pitch * max_rows);
    A(&h->d_scaled, B * align_up((size_t)W, 16) * H);
    A(&h->d_ang, B * plane);
    A(&h->d_rec, B * plane);
    A(&h->d_cs0, B * plane);
    A(&h->d_nrects, B);
    A(&h->d_g2, B * plane);
    A(&h->d_reg, B * plane);
    A(&h->d_seeds, B * plane);
    A(&h->d_maxg2, B);
    A(&h->d_nseeds, B);
    A(&h->d_nsegs, B);
    A(&h->d_flags, B);
    A(&h->d_nout, B);
    A(&h->d_tile_hist, B * tiles * kBins);
    A(&h->d_tile_off, B * tiles * kBins);
    A(&h->d_segs, B * seg_cap);
    A(&h->d_qres, B * seg_cap);
    A(&h->d_queue, B * seg_cap);
    A(&h->d_qvalid, B * seg_cap);
    h->bits_words = (int)((plane + 31) / 32);
    {
        cudaDeviceProp prop;
        if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
        if (e == cudaSuccess) {
            h->num_sms = prop.multiProcessorCount;
            // shared memory: per frame slot the ticket slots + committed bitmap, per grower the ring, the staging area
            // and the sparse private bitmap (a pool of 32x32-pixel tiles: the more, the fewer regions overflow)
            const int tiles = ((W + 31) / 32) * ((H + 31) / 32);
            h->grow_tiles = tiles;
            if (const char* ev = getenv("PLSLAM_LSD_TAIL_NFA")) h->tail_nfa = std::max(0, std::min(2, atoi(ev)));
            if (const char* ev = getenv("PLSLAM_LSD_POLL_NS")) h->poll_ns = std::max(20, std::min(100000, atoi(ev)));
            if (const char* ev = getenv("PLSLAM_LSD_RESERVE_SMS")) h->reserved_sms = std::max(0, std::min(h->num_sms - 1, atoi(ev)));
            h->grow_window = 128;
            if (const char* ev = getenv("PLSLAM_LSD_WINDOW")) h->grow_window = std::max(1, std::min(kSlots2, atoi(ev)));
            // ---- k_lsd_grow2: warps per CTA, CTAs per SM, private tile pool ----
            if (const char* ev = getenv("PLSLAM_LINE_TMA")) h->use_tma = atoi(ev) != 0;
            if (const char* ev = getenv("PLSLAM_LINE_GRAPH")) h->use_graph = atoi(ev) != 0;
            {
                cudaDriverEntryPointQueryResult qres;
                void* fn = nullptr;
                if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
                    h->encode_tiled = fn;
                else
                    (void)cudaGetLastError();
            }
            if (const char* ev = getenv("PLSLAM_LSD_FORCE_MANY")) h->force_many = atoi(ev) != 0;
            if (const char* ev = getenv("PLSLAM_LSD_LOOKAHEAD")) h->lookahead = std::max(1, std::min(kSlots2, atoi(ev)));
            auto allow_smem = [&](const void* fn) {  // dynamic shared memory up to what the kernel's static part leaves
                cudaFuncAttributes fa;
                if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, fn);
                if (e == cudaSuccess)
                    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(prop.sharedMemPerBlockOptin - fa.sharedSizeBytes));
            };
            allow_smem((const void*)k_lsd_grow2<256, 3>);
            allow_smem((const void*)k_lsd_grow2<384, 2>);
            allow_smem((const void*)k_lsd_grow2<512, 1>);
            allow_smem((const void*)k_lsd_grow2<1024, 1>);
            const size_t sm_smem = prop.sharedMemPerMultiprocessor;
            // threads = 32 * (1 sequencer + growers); the largest tile pool that still gives the wanted CTAs per SM
            auto choose2 = [&](int threads, int occ, pl_line::Grow2Cfg* c) {
                const size_t fixed = sizeof(Grow2Shared) + 64;  // static shared memory
                const size_t per_cta = std::min<size_t>(sm_smem / occ > 1024 + fixed ? sm_smem / occ - 1024 - fixed : 0, prop.sharedMemPerBlockOptin - fixed);
                for (int pN = kMaxPoolTiles; pN >= 8; pN -= 4) {
                    Grow2Smem g2{tiles, pN, 0, 0, h->bits_words, 0, 0, 0};
                    if (g2.total(threads / 32) <= per_cta) {
                        c->threads = threads;
                        c->occ = occ;
                        c->pool_tiles = pN;
                        c->pool_n = std::min(kMaxPool2, threads / 32 - 1 + 17);
                        c->smem = g2.total(threads / 32);
                        return true;
                    }
                }
                return false;
            };
            int t_few = 512, t_many = 256, occ_many = 3;
            if (const char* ev = getenv("PLSLAM_LSD_GROW2")) {  // tuning override: "<threads few>,<threads many>,<CTAs per SM>"
                int a = 0, b2 = 0, c2 = 0;
                if (sscanf(ev, "%d,%d,%d", &a, &b2, &c2) == 3) {
                    if (a >= 64 && a <= 1024 && a % 32 == 0) t_few = a;
                    if (b2 >= 64 && b2 <= 1024 && b2 % 32 == 0) t_many = b2;
                    if (c2 >= 1 && c2 <= 8) occ_many = c2;
                }
            }
            for (int t = t_few; t >= 64 && !h->g2_few.threads; t -= 32) choose2(t, 1, &h->g2_few);
            if (t_many > 384) occ_many = 1;
            else if (t_many > 256) occ_many = std::min(occ_many, 2);
            for (int oc = occ_many; oc >= 1 && !h->g2_many.threads; oc--)
                for (int t = t_many; t >= 64 && !h->g2_many.threads; t -= 32) {
                    if (!choose2(t, oc, &h->g2_many)) continue;
                    int occ_real = 0;  // what the device really gives (registers count too)
                    cudaError_t eo = oc < 2 ? cudaSuccess
                                     : t <= 256 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_real, k_lsd_grow2<256, 3>, t, h->g2_many.smem)
                                                : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_real, k_lsd_grow2<384, 2>, t, h->g2_many.smem);
                    if (oc >= 2 && (eo != cudaSuccess || occ_real < oc)) h->g2_many = pl_line::Grow2Cfg{};
                }
            // a frame that is alone on its SM is bound by the sequencer: committing and issuing on a warp each
            h->g2_few.split = h->g2_few.threads >= 128 ? 1 : 0;
            if (const char* ev = getenv("PLSLAM_LSD_SPLIT")) {  // tuning override: "<few>,<many>"
                int a = 0, b2 = 0;
                if (sscanf(ev, "%d,%d", &a, &b2) == 2) {
                    h->g2_few.split = a != 0 && h->g2_few.threads >= 96;
                    h->g2_many.split = b2 != 0 && h->g2_many.threads >= 96;
                }
            }
            if (const char* ev = getenv("PLSLAM_LSD_POOL_TILES")) {
                const int pt = atoi(ev);
                if (pt >= 1 && pt <= kMaxPoolTiles) {
                    h->g2_few.pool_tiles = std::min(h->g2_few.pool_tiles, pt);
                    h->g2_many.pool_tiles = std::min(h->g2_many.pool_tiles, pt);
                }
            }
            if (h->g2_few.threads < 64) {
                set_error("pl_line_create: a %dx%d image does not leave shared memory for a sequencer and a grower warp", max_cols, max_rows);
                pl_line_destroy(h);
                return PL_ERR_CAPACITY;
            }
        }
    }
    // the buffers of k_lsd_grow2 are indexed by CTA
    const size_t c_few = std::min<size_t>(B, (size_t)std::max(h->num_sms, 1));
    const size_t c_many = std::min<size_t>(B, (size_t)std::max(h->num_sms, 1) * std::max(1, h->g2_many.occ));
    const size_t pool_rects = std::max(c_few * h->g2_few.pool_n, c_many * h->g2_many.pool_n), pool_words = pool_rects * (size_t)kSpecCap2;
    const size_t max_fs = std::max(c_few, c_many), small_slots = max_fs * kSlots2;
    A(&h->d_spec_reg, pool_words);
    A(&h->d_spec_touched, pool_words);
    A(&h->d_pool_rect, pool_rects);
    A(&h->d_small_buf, small_slots * 2 * (size_t)kSmall);
    A(&h->d_small_rect, small_slots);
    A(&h->d_big_touched, max_fs * 2 * plane);
    A(&h->d_big_bits, max_fs * (size_t)h->bits_words);
    if (e == cudaSuccess) e = cudaMemset(h->d_big_bits, 0, max_fs * (size_t)h->bits_words * sizeof(unsigned int));
    A(&h->d_frame_counter, 1);
    A(&h->d_nfa_ctl, 4 + kSmBusySlots);
    A(&h->d_nfa_items, B * kNfaChunksPerFrame);
    A(&h->d_nfa_next, B);
    A(&h->d_sticky, 1);
    if (e == cudaSuccess) e = cudaMemset(h->d_sticky, 0, sizeof(int));
    A(&h->d_resp, B * seg_cap);
    A(&h->d_dx, B * align_up((size_t)max_cols, 8) * max_rows);
    A(&h->d_dy, B * align_up((size_t)max_cols, 8) * max_rows);
    A(&h->d_xtab, (size_t)W);
    A(&h->d_ytab, (size_t)H);
    A(&h->d_lgam, (size_t)kLgMax);
    A(&h->d_phase, B * 16);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&h->h_flags, sizeof(int) * B);
    if (e != cudaSuccess) {
        set_error("pl_line_create: %s", cudaGetErrorString(e));
        pl_line_destroy(h);
        return PL_ERR_CUDA;
    }
    h->seg_cap_alloc = seg_cap;
    {
        // log_gamma table of lsd.cpp (#define log_gamma(x) ((x)>15.0?log_gamma_windschitl(x):log_gamma_lanczos(x)))
        std::vector<double> lg(kLgMax);
        for (int m = 0; m < kLgMax; m++) {
            const double x = (double)m + 1;
            if (x > 15.0) {
                lg[m] = 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
            } else {
                static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
                double a = (x + 0.5) * log(x + 5.5) - (x + 5.5), b = 0;
                for (int n = 0; n < 7; ++n) {
                    a -= log(x + double(n));
                    b += q[n] * pow(x, double(n));
                }
                lg[m] = a + log(b);
            }
        }
        if (cudaMemcpy(h->d_lgam, lg.data(), sizeof(double) * kLgMax, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("pl_line_create: table upload failed");
            pl_line_destroy(h);
            return PL_ERR_CUDA;
        }
        h->nfa_tabs.lgam = h->d_lgam;
        double p = 22.5 / 180;
        for (int j = 0; j < kPMax; j++) {
            h->nfa_tabs.logp[j] = log(p);
            h->nfa_tabs.log1mp[j] = log(1.0 - p);
            h->nfa_tabs.log10p[j] = log10(p);
            p /= 2;
        }
    }
    *out = h;
    return PL_OK;
}

PL_API void pl_line_destroy(pl_line* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) pl::stream_sync(h->stream);
    if (h->ev_grow) cudaEventDestroy(h->ev_grow);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    void* bufs[] = {h->d_in, h->d_scaled, h->d_big_bits, h->d_blur5, h->d_ang, h->d_g2, h->d_reg, h->d_seeds, h->d_maxg2, h->d_tile_off,
                    h->d_nseeds, h->d_nsegs, h->d_flags, h->d_nout, h->d_tile_hist, h->d_segs, h->d_resp, h->d_rowsum, h->d_fdesc,
                    h->d_dx, h->d_dy, h->d_xtab, h->d_ytab, h->d_kls, h->d_desc, h->d_coef, h->d_lgam, h->d_phase, h->d_qres, h->d_queue, h->d_qvalid, h->d_spec_reg, h->d_spec_touched, h->d_big_touched, h->d_pool_rect, h->d_small_buf, h->d_small_rect, h->d_frame_counter, h->d_nfa_ctl, h->d_nfa_items, h->d_nfa_next, h->d_sticky, h->d_rec, h->d_cs0, h->d_nrects};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    for (int i = 0; i < 8; i++)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

PL_API int pl_line_sync(pl_line* h) {
    PL_CHECK_ARG(h);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    int sticky = 0;
    PL_CUDA_TRY(pl::stream_sync(h->stream));  // a copy into pageable memory blocks inside the runtime until the stream gets there
    PL_CUDA_TRY(cudaMemcpyAsync(&sticky, h->d_sticky, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    if (sticky) {
        PL_CUDA_TRY(cudaMemsetAsync(h->d_sticky, 0, sizeof(int), h->stream));
        set_error("a frame extracted through the device-pointer API exceeded an LSD capacity (flags=%d: 1=segments, 2=region size)", sticky);
        return PL_ERR_CAPACITY;
    }
    return PL_OK;
}
PL_API void* pl_line_stream(pl_line* h) { return h ? (void*)h->stream : nullptr; }
PL_API int pl_line_set_graph(pl_line* h, int on) {
    PL_CHECK_ARG(h);
    h->use_graph = on != 0;
    return PL_OK;
}
PL_API int pl_line_stream_wait_grow_start(pl_line* h, void* stream) {
    PL_CHECK_ARG(h);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    if (!h->ev_grow) PL_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_grow, cudaEventDisableTiming));
    PL_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_grow, 0));
    return PL_OK;
}
PL_API int pl_line_set_reserved_sms(pl_line* h, int n) {
    PL_CHECK_ARG(h && n >= 0 && n < std::max(h->num_sms, 1));
    h->reserved_sms = n;
    return PL_OK;
}
PL_API int pl_line_last_launches(const pl_line* h) { return h ? h->last_launches : 0; }

PL_API int pl_line_set_profiling(pl_line* h, int on) {
    PL_CHECK_ARG(h);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    if (on && !h->ev[0])
        for (int i = 0; i < 8; i++) PL_CUDA_TRY(cudaEventCreate(&h->ev[i]));
    h->profiling = on != 0;
    for (int i = 0; i < 8; i++) h->stage_ms[i] = 0;
    h->stage_chunks = 0;
    return PL_OK;
}
PL_API int pl_line_nfa_ms(pl_line* h, float* ms) {
    PL_CHECK_ARG(h && ms);
    *ms = h->stage_ms[5];
    return PL_OK;
}

PL_API int pl_line_stage_ms(pl_line* h, float* out5, int* chunks) {
    PL_CHECK_ARG(h && out5);
    for (int i = 0; i < 5; i++) out5[i] = h->stage_ms[i];
    if (chunks) *chunks = h->stage_chunks;
    return PL_OK;
}

static int line_check_common(pl_line* h, const void* gray, int n_frames, int rows, int cols, size_t step, int max_lines) {
    if (!gray || rows <= 0 || cols <= 0 || n_frames <= 0) {
        set_error("empty image");
        return PL_ERR_EMPTY;
    }
    PL_CHECK_ARG(cols <= h->max_cols && rows <= h->max_rows && step >= (size_t)cols);
    PL_CHECK_ARG(max_lines >= 1 && max_lines <= kMaxKeep);
    return PL_OK;
}

PL_API int pl_line_extract_batch_dev(pl_line* h, const uint8_t* d_gray, int n_frames, int rows, int cols, size_t step,
                                     size_t frame_stride, int max_lines, pl_keyline* d_kls, uint8_t* d_desc, double* d_coeffs,
                                     int* d_n_out) {
    PL_CHECK_ARG(h && d_kls && d_desc && d_coeffs && d_n_out);
    int rc = line_check_common(h, d_gray, n_frames, rows, cols, step, max_lines);
    if (rc != PL_OK) return rc;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    if ((rc = line_geometry(h, rows, cols, max_lines)) != PL_OK) return rc;
    if ((rc = line_ensure_out(h, max_lines)) != PL_OK) return rc;
    h->last_launches = 0;
    for (int f0 = 0; f0 < n_frames; f0 += h->max_batch) {
        const int nf = std::min(h->max_batch, n_frames - f0);
        rc = line_launch_chunk(h, d_gray + (size_t)f0 * frame_stride, nf, step, frame_stride, d_kls + (size_t)f0 * max_lines,
                               d_desc + (size_t)f0 * max_lines * 32, d_coeffs + (size_t)f0 * max_lines * 3, max_lines, d_n_out + f0);
        if (rc != PL_OK) return rc;
        k_line_or_flags<<<1, 32, 0, h->stream>>>(h->d_flags, nf, h->d_sticky);
    }
    return PL_OK;
}

// host outputs; the images come from host memory (staged per chunk) or are already on the device
static int line_extract_to_host(pl_line* h, const uint8_t* gray, bool gray_on_device, int n_frames, int rows, int cols, size_t step, size_t frame_stride,
                                int max_lines, pl_keyline* kls, uint8_t* desc, double* coeffs, int* n_out) {
    PL_CHECK_ARG(h && kls && desc && coeffs && n_out);
    int rc = line_check_common(h, gray, n_frames, rows, cols, step, max_lines);
    if (rc != PL_OK) return rc;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    if ((rc = line_geometry(h, rows, cols, max_lines)) != PL_OK) return rc;
    if ((rc = line_ensure_out(h, max_lines)) != PL_OK) return rc;
    const size_t in_pitch = h->geom.in_pitch;
    h->last_launches = 0;
    for (int f0 = 0; f0 < n_frames; f0 += h->max_batch) {
        const int nf = std::min(h->max_batch, n_frames - f0);
        if (gray_on_device) {
            rc = line_launch_chunk(h, gray + (size_t)f0 * frame_stride, nf, step, frame_stride, h->d_kls, h->d_desc, h->d_coef, max_lines, h->d_nout);
        } else {
            // one copy per chunk when the caller's frames are dense and already have the staging pitch, else one strided copy per frame
            if (step == in_pitch && frame_stride == in_pitch * (size_t)rows)
                PL_CUDA_TRY(cudaMemcpyAsync(h->d_in, gray + (size_t)f0 * frame_stride, (size_t)nf * frame_stride, cudaMemcpyHostToDevice, h->stream));
            else
                for (int f = 0; f < nf; f++)
                    PL_CUDA_TRY(cudaMemcpy2DAsync(h->d_in + (size_t)f * in_pitch * rows, in_pitch, gray + (size_t)(f0 + f) * frame_stride, step,
                                                  cols, rows, cudaMemcpyHostToDevice, h->stream));
            rc = line_launch_chunk(h, h->d_in, nf, in_pitch, in_pitch * rows, h->d_kls, h->d_desc, h->d_coef, max_lines, h->d_nout);
        }
        if (rc != PL_OK) return rc;
        PL_CUDA_TRY(cudaMemcpyAsync(n_out + f0, h->d_nout, sizeof(int) * nf, cudaMemcpyDeviceToHost, h->stream));
        if ((rc = line_check_flags(h, nf)) != PL_OK) return rc;
        PL_CUDA_TRY(cudaMemcpyAsync(kls + (size_t)f0 * max_lines, h->d_kls, sizeof(pl_keyline) * (size_t)nf * max_lines,
                                    cudaMemcpyDeviceToHost, h->stream));
        PL_CUDA_TRY(cudaMemcpyAsync(desc + (size_t)f0 * max_lines * 32, h->d_desc, (size_t)nf * max_lines * 32, cudaMemcpyDeviceToHost,
                                    h->stream));
        PL_CUDA_TRY(cudaMemcpyAsync(coeffs + (size_t)f0 * max_lines * 3, h->d_coef, sizeof(double) * (size_t)nf * max_lines * 3,
                                    cudaMemcpyDeviceToHost, h->stream));
        PL_CUDA_TRY(pl::stream_sync(h->stream));
    }
    return PL_OK;
}

PL_API int pl_line_extract_batch(pl_line* h, const uint8_t* gray, int n_frames, int rows, int cols, size_t step, size_t frame_stride,
                                 int max_lines, pl_keyline* kls, uint8_t* desc, double* coeffs, int* n_out) {
    return line_extract_to_host(h, gray, false, n_frames, rows, cols, step, frame_stride, max_lines, kls, desc, coeffs, n_out);
}

PL_API int pl_line_extract_batch_from_dev(pl_line* h, const uint8_t* d_gray, int n_frames, int rows, int cols, size_t step, size_t frame_stride,
                                          int max_lines, pl_keyline* kls, uint8_t* desc, double* coeffs, int* n_out) {
    return line_extract_to_host(h, d_gray, true, n_frames, rows, cols, step, frame_stride, max_lines, kls, desc, coeffs, n_out);
}

PL_API int pl_line_extract(pl_line* h, const uint8_t* gray, int rows, int cols, size_t step, int max_lines, pl_keyline* kls,
                           uint8_t* desc, double* coeffs, int* n_out) {
    return pl_line_extract_batch(h, gray, 1, rows, cols, step, step * (size_t)(rows > 0 ? rows : 0), max_lines, kls, desc, coeffs, n_out);
}

PL_API int pl_line_lsd_read(pl_line* h, int frame, float* xyxy, double* width, double* prec, double* nfa, int cap, int* n_out) {
    PL_CHECK_ARG(h && xyxy && n_out);
    if (frame < 0 || frame >= h->last_batch) {
        set_error("no extract result for frame %d", frame);
        return PL_ERR_STATE;
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    int n = 0;
    PL_CUDA_TRY(pl::stream_sync(h->stream));  // a copy into pageable memory blocks inside the runtime until the stream gets there
    PL_CUDA_TRY(cudaMemcpyAsync(&n, h->d_nsegs + frame, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    *n_out = n;
    if (n > cap) return PL_ERR_CAPACITY;
    std::vector<LsdSeg> tmp(n);
    if (n) {
        PL_CUDA_TRY(cudaMemcpyAsync(tmp.data(), h->d_segs + (size_t)frame * h->geom.seg_cap, sizeof(LsdSeg) * n, cudaMemcpyDeviceToHost, h->stream));
        PL_CUDA_TRY(pl::stream_sync(h->stream));
    }
    for (int i = 0; i < n; i++) {
        xyxy[4 * i] = tmp[i].x1; xyxy[4 * i + 1] = tmp[i].y1; xyxy[4 * i + 2] = tmp[i].x2; xyxy[4 * i + 3] = tmp[i].y2;
        if (width) width[i] = tmp[i].width;
        if (prec) prec[i] = tmp[i].p;
        if (nfa) nfa[i] = tmp[i].nfa;
    }
    return PL_OK;
}

/* with profiling on: cycles of frame `frame` of the last chunk spent in {seed scan, grow, rect fit, refine, NFA} and
 * the number of regions tried / regions that reached the minimum size (k_lsd_grow's own clock64 accounting) */
PL_API int pl_line_grow_phases(pl_line* h, int frame, long long* out16) {
    PL_CHECK_ARG(h && out16 && frame >= 0 && frame < h->last_batch);
    // out16 receives 16 values, see plslam_c.h
    PL_CUDA_TRY(cudaSetDevice(h->device));
    const size_t stride = 16;
    for (int i = 8; i < 16; i++) out16[i] = 0;
    PL_CUDA_TRY(cudaMemcpyAsync(out16, h->d_phase + (size_t)frame * stride, sizeof(long long) * stride, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    return PL_OK;
}

/* test hooks: the 0.8-scaled 8-bit image, the level-line angle map (degrees, -1024 = undefined) and the float LBD
 * descriptors (n x 72) of the last call */
PL_API int pl_line_scaled_dims(const pl_line* h, int* rows, int* cols) {
    PL_CHECK_ARG(h && rows && cols && h->rows > 0);
    *rows = h->geom.H;
    *cols = h->geom.W;
    return PL_OK;
}
PL_API int pl_line_scaled_read(pl_line* h, int frame, uint8_t* out, size_t out_step) {
    PL_CHECK_ARG(h && out && frame >= 0 && frame < h->last_batch && out_step >= (size_t)h->geom.W);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    PL_CUDA_TRY(cudaMemcpy2DAsync(out, out_step, h->d_scaled + (size_t)frame * h->scaled_stride, h->geom.spitch, h->geom.W, h->geom.H,
                                  cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    return PL_OK;
}
PL_API int pl_line_angles_read(pl_line* h, int frame, float* out) {
    PL_CHECK_ARG(h && out && frame >= 0 && frame < h->last_batch);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    PL_CUDA_TRY(cudaMemcpyAsync(out, h->d_ang + (size_t)frame * h->plane, sizeof(float) * h->plane, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    return PL_OK;
}
PL_API int pl_line_fdesc_read(pl_line* h, int frame, float* out, int n) {
    PL_CHECK_ARG(h && out && frame >= 0 && frame < h->last_batch && n >= 0 && n <= h->out_cap);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    PL_CUDA_TRY(cudaMemcpyAsync(out, h->d_fdesc + (size_t)frame * h->max_lines * 72, sizeof(float) * 72 * n, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    return PL_OK;
}

}  // extern "C"

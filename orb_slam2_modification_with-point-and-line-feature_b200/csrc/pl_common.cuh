// pl_common.cuh — shared host/device helpers of the sm_100a front-end library.
//
// Floating-point discipline (DESIGN.md "exactness"): the whole library is compiled with -fmad=false and the
// expressions whose rounding matters are additionally written with explicit round-to-nearest intrinsics, so no
// FMA contraction can change a cvRound() decision relative to the reference CPU path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/plslam_c.h"

namespace pl {

void set_error(const char* fmt, ...);

#define PL_CUDA_TRY(expr)                                                                              \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            pl::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));  \
            return PL_ERR_CUDA;                                                                        \
        }                                                                                              \
    } while (0)

#define PL_CHECK_ARG(cond)                                                          \
    do {                                                                            \
        if (!(cond)) {                                                              \
            pl::set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);    \
            return PL_ERR_ARG;                                                      \
        }                                                                           \
    } while (0)

#define PL_STR2(x) #x
#define PL_STR(x) PL_STR2(x)

constexpr int kEdge = 19;        // EDGE_THRESHOLD   (ORBextractor.cc:73)
constexpr int kHalfPatch = 15;   // HALF_PATCH_SIZE  (ORBextractor.cc:72)
constexpr int kPatch = 31;       // PATCH_SIZE       (ORBextractor.cc:71)
constexpr int kMaxLevels = 16;

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Waits for a stream.  NB for every caller: an asynchronous D2H copy into PAGEABLE host memory is not asynchronous — the call
// blocks inside the runtime until the stream reaches the copy, and while it does the other host threads of the process get
// no work through the runtime either (measured: pl_line_sync used to read its capacity flag into a stack variable; a matcher
// call issued from a second thread then did not start until the line extractor had finished, 35 ms later).  So: wait for the
// stream first, copy into pageable memory afterwards (or copy into pinned memory).
inline cudaError_t stream_sync(cudaStream_t st) { return cudaStreamSynchronize(st); }

#ifdef __CUDACC__
// BORDER_REFLECT_101 index (one reflection is enough for |overshoot| < len; loop kept for tiny images)
__host__ __device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// cvRound(float): round-half-to-even
__device__ __forceinline__ int cv_round(float v) { return __float2int_rn(v); }

// cv::fastAtan2 scalar path (atan_f32), degrees in [0,360).  Every product/sum is a separately rounded float op.
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.141592653589793238462643383279502884);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16;
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// sinf/cosf with glibc's algorithm (sysdeps/ieee754/flt-32/s_sincosf.h, from ARM optimized-routines): the
// reference's std::cos(float)/std::sin(float) at ORBextractor.cc:113 resolve to these.  Double-precision
// polynomial after a fast range reduction; valid for |y| < 120 (the path only feeds [0, 2*pi)).
// Verified on the CPU (tests/test_sincos.py + an exhaustive sweep of all 1,086,918,650 floats in [0, 6.2832],
// DESIGN.md) to return bit-identical floats to glibc 2.39 cosf/sinf, with and without FMA contraction.
struct SinCosTab {
    double hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3;
};
__host__ __device__ __forceinline__ float sincosf_poly(double x, double x2, bool neg_cos, int n) {
    const double c0 = neg_cos ? -0x1p0 : 0x1p0;
    const double c1 = neg_cos ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2;
    const double c2 = neg_cos ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5;
    const double c3 = neg_cos ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10;
    const double c4 = neg_cos ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1_ = s2 + x2 * s3;
        double x7 = x3 * x2;
        double s = x + x3 * s1;
        return (float)(s + x7 * s1_);
    } else {
        double x4 = x2 * x2;
        double c2_ = c3 + x2 * c4;
        double c1_ = c0 + x2 * c1;
        double x6 = x4 * x2;
        double c = c1_ + x4 * c2;
        return (float)(c + x6 * c2_);
    }
}
__host__ __device__ __forceinline__ uint32_t f32_abstop12(float x) {
#ifdef __CUDA_ARCH__
    return (__float_as_uint(x) >> 20) & 0x7ff;
#else
    uint32_t u;
    memcpy(&u, &x, 4);
    return (u >> 20) & 0x7ff;
#endif
}
// is_cos: 1 -> cosf, 0 -> sinf
__host__ __device__ __forceinline__ float glibc_sincosf(float y, int is_cos) {
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    double x = y;
    if (f32_abstop12(y) < f32_abstop12(0x1.921FB6p-1f)) {
        double x2 = x * x;
        if (f32_abstop12(y) < f32_abstop12(0x1p-12f)) return is_cos ? 1.0f : y;
        return sincosf_poly(x, x2, false, is_cos);
    }
    double r = x * hpi_inv;
    int n = ((int32_t)r + 0x800000) >> 24;
    x = x - n * hpi;
    const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return sincosf_poly(x * sgn, x * x, (n & 2) != 0, n ^ is_cos);
}

// glibc logf (sysdeps/ieee754/flt-32/e_logf.c, the ARM optimized-routines algorithm) for normal positive finite x:
// table of 16 (1/c, log c) pairs + cubic in double.  Checked against the host libm on all 2,130,706,432 normal positive
// floats, with and without FMA contraction (tests/test_sincos.py).
__host__ __device__ __forceinline__ float glibc_logf(float x) {
    const double T[16][2] = {
        {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2}, {0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2},
        {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3}, {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3},
        {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4}, {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5},
        {0x1p+0, 0x0p+0},                              {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4},
        {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3},  {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},
        {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
    const double Ln2 = 0x1.62e42fefa39efp-1, A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
    uint32_t ix;
#ifdef __CUDA_ARCH__
    ix = __float_as_uint(x);
#else
    memcpy(&ix, &x, 4);
#endif
    if (ix == 0x3f800000u) return 0.f;
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & (0x1ffu << 23));
    float zf;
#ifdef __CUDA_ARCH__
    zf = __uint_as_float(iz);
    const double z = (double)zf;
    const double r = __dadd_rn(__dmul_rn(z, T[i][0]), -1.0);
    const double y0 = __dadd_rn(T[i][1], __dmul_rn((double)k, Ln2));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(A1, r), A2);
    y = __dadd_rn(__dmul_rn(A0, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
#else
    memcpy(&zf, &iz, 4);
    const double z = (double)zf;
    const double r = z * T[i][0] - 1.0;
    const double y0 = T[i][1] + (double)k * Ln2;
    const double r2 = r * r;
    double y = A1 * r + A2;
    y = A0 * r2 + y;
    y = y * r2 + (y0 + r);
#endif
    return (float)y;
}
#endif  // __CUDACC__

}  // namespace pl

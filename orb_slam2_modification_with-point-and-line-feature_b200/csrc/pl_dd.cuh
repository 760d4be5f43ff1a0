// pl_dd.cuh — correctly rounded double-precision sin / cos for the places where one ulp decides a discrete outcome.
//
// region2rect() of lsd.cpp sets rec.dx = cos(theta), rec.dy = sin(theta) (double) and the rectangle's edges then pass exactly
// through the extreme pixels of the region: which rows and columns rect_nfa() counts is decided by the last bit of dx / dy.  The
// host libm (glibc >= 2.28: < 0.55 ulp, i.e. correctly rounded except when the exact value lies within 0.05 ulp of a rounding
// boundary) and CUDA's sin / cos (up to 2 ulp) disagree on that bit for a few inputs per thousand.  Here both values are evaluated in
// double-double arithmetic (about 100 bits: Cody-Waite reduction with a three-term pi / 2, Taylor series with double-double
// coefficients) and rounded once: the correctly rounded result.  tests/cpp/dd_sincos_check.cu compares it with the host libm.
#pragma once
#include <math.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace pl {

struct dd {
    double hi, lo;
};
#ifdef __CUDA_ARCH__
#define PL_FMA(a, b, c) __fma_rn(a, b, c)
#define PL_ADD(a, b) __dadd_rn(a, b)
#define PL_SUB(a, b) __dsub_rn(a, b)
#define PL_MUL(a, b) __dmul_rn(a, b)
#else
#define PL_FMA(a, b, c) fma(a, b, c)
#define PL_ADD(a, b) ((a) + (b))
#define PL_SUB(a, b) ((a) - (b))
#define PL_MUL(a, b) ((a) * (b))
#endif
__host__ __device__ __forceinline__ dd dd_two_sum(double a, double b) {
    const double s = PL_ADD(a, b), bb = PL_SUB(s, a);
    return {s, PL_ADD(PL_SUB(a, PL_SUB(s, bb)), PL_SUB(b, bb))};
}
__host__ __device__ __forceinline__ dd dd_quick_two_sum(double a, double b) {  // |a| >= |b|
    const double s = PL_ADD(a, b);
    return {s, PL_SUB(b, PL_SUB(s, a))};
}
__host__ __device__ __forceinline__ dd dd_two_prod(double a, double b) {
    const double p = PL_MUL(a, b);
    return {p, PL_FMA(a, b, -p)};
}
__host__ __device__ __forceinline__ dd dd_add(dd a, dd b) {
    dd s = dd_two_sum(a.hi, b.hi);
    const dd t = dd_two_sum(a.lo, b.lo);
    s.lo = PL_ADD(s.lo, t.hi);
    s = dd_quick_two_sum(s.hi, s.lo);
    s.lo = PL_ADD(s.lo, t.lo);
    return dd_quick_two_sum(s.hi, s.lo);
}
__host__ __device__ __forceinline__ dd dd_mul(dd a, dd b) {
    dd p = dd_two_prod(a.hi, b.hi);
    p.lo = PL_ADD(p.lo, PL_ADD(PL_MUL(a.hi, b.lo), PL_MUL(a.lo, b.hi)));
    return dd_quick_two_sum(p.hi, p.lo);
}
__host__ __device__ __forceinline__ dd dd_mul_d(dd a, double b) {
    dd p = dd_two_prod(a.hi, b);
    p.lo = PL_ADD(p.lo, PL_MUL(a.lo, b));
    return dd_quick_two_sum(p.hi, p.lo);
}

// sin(x) and cos(x), correctly rounded, for |x| < 2^20 (the callers pass angles of a few radians)
__host__ __device__ inline void sincos_cr(double x, double* s_out, double* c_out) {
    // (-1)^k / (2k + 1)!  and  (-1)^k / (2k)!,  k = 1 .. 14, as double-double
    const double S[14][2] = {
        {-0.16666666666666666, -9.2518585385429707e-18},  {0.0083333333333333332, 1.1564823173178714e-19},
        {-0.00019841269841269841, -1.7209558293420705e-22}, {2.7557319223985893e-06, -1.8583932740464721e-22},
        {-2.505210838544172e-08, 1.448814070935912e-24},  {1.6059043836821613e-10, 1.2585294588752098e-26},
        {-7.6471637318198164e-13, -7.03872877733453e-30}, {2.8114572543455206e-15, 1.6508842730861433e-31},
        {-8.2206352466243295e-18, -2.2141894119604265e-34}, {1.9572941063391263e-20, -1.3643503830087908e-36},
        {-3.8681701706306841e-23, 8.8431776554823438e-40}, {6.4469502843844736e-26, -1.9330404233703465e-42},
        {-9.183689863795546e-29, -1.4303150396787322e-45}, {1.1309962886447716e-31, 1.0498015412959506e-47}};
    const double C[14][2] = {
        {-0.5, 0},                                        {0.041666666666666664, 2.3129646346357427e-18},
        {-0.0013888888888888889, 5.3005439543735771e-20}, {2.4801587301587302e-05, 2.1511947866775882e-23},
        {-2.7557319223985888e-07, -2.3767714622250297e-23}, {2.08767569878681e-09, -1.20734505911326e-25},
        {-1.1470745597729725e-11, -2.0655512752830745e-28}, {4.7794773323873853e-14, 4.3992054858340813e-31},
        {-1.5619206968586225e-16, -1.1910679660273754e-32}, {4.1103176233121648e-19, 1.4412973378659527e-36},
        {-8.8967913924505741e-22, 7.9114026148723762e-38}, {1.6117375710961184e-24, -3.6846573564509766e-41},
        {-2.4795962632247976e-27, 1.2953730964765229e-43}, {3.2798892370698378e-30, 1.5117542744029879e-46}};
    const double P1 = 1.5707963267948966, P2 = 6.123233995736766e-17, P3 = -1.4973849048591698e-33;
    const double kd = rint(PL_MUL(x, 0.63661977236758138));
    const long long k = (long long)kd;
    // r = x - k * pi / 2 in double-double (k * P1 is formed exactly as a two-product)
    dd r = {x, 0.0};
    dd t = dd_two_prod(kd, P1);
    r = dd_add(r, dd{-t.hi, -t.lo});
    t = dd_two_prod(kd, P2);
    r = dd_add(r, dd{-t.hi, -t.lo});
    r = dd_add(r, dd{-PL_MUL(kd, P3), 0.0});
    const dd z = dd_mul(r, r);
    dd ps = {S[13][0], S[13][1]}, pc = {C[13][0], C[13][1]};
#pragma unroll 1
    for (int i = 12; i >= 0; i--) {
        ps = dd_add(dd_mul(ps, z), dd{S[i][0], S[i][1]});
        pc = dd_add(dd_mul(pc, z), dd{C[i][0], C[i][1]});
    }
    const dd sr = dd_add(r, dd_mul(dd_mul(ps, z), r));     // r + r^3 * S(r^2)
    const dd cr = dd_add(dd{1.0, 0.0}, dd_mul(pc, z));     // 1 + r^2 * C(r^2)
    const int q = (int)(k & 3);
    const dd sv = (q & 1) ? cr : sr, cv = (q & 1) ? sr : cr;
    const double s = PL_ADD(sv.hi, sv.lo), c = PL_ADD(cv.hi, cv.lo);
    *s_out = (q == 2 || q == 3) ? -s : s;
    *c_out = (q == 1 || q == 2) ? -c : c;
}

}  // namespace pl

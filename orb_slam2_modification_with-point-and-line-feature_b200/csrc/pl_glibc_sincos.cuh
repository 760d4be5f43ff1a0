// pl_glibc_sincos.cuh — double-precision sin / cos with the arithmetic of glibc's __sin / __cos (sysdeps/ieee754/dbl-64/s_sin.c,
// glibc >= 2.28, the FMA build the dynamic loader selects on every x86-64 host with FMA3: sysdeps/x86_64/fpu/multiarch).
//
// Why: region2rect() of lsd.cpp sets rec.dx = cos(theta), rec.dy = sin(theta) and the rectangle's edges then pass exactly through
// the extreme pixels of the region, so the last bit of dx / dy decides which pixels rect_nfa() counts and, near log_nfa = 0, whether
// a segment exists.  CUDA's sin / cos differ from the host's in that bit for a few inputs per hundred, a correctly rounded
// implementation for 1.4 per thousand (glibc's result is within 0.55 ulp, not correctly rounded).  This file restates glibc's
// algorithm operation by operation — table look-up of sin / cos at multiples of 1/128 (__sincostab, rebuilt here from its
// definition: the double nearest to the value and the double nearest to the remainder) plus short polynomials — with every fused
// multiply-add where GCC contracts the source expressions (checked against the disassembly of the same expressions).
// tests/cpp/glibc_sincos_check.cpp compares it with the host libm on 10^8 arguments: no difference on an x86-64 host with FMA.
// Valid for |x| < 105414350 (the callers pass angles of a few radians).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDA_ARCH__
#define PLG_FMA(a, b, c) __fma_rn(a, b, c)
#define PLG_ADD(a, b) __dadd_rn(a, b)
#define PLG_SUB(a, b) __dsub_rn(a, b)
#define PLG_MUL(a, b) __dmul_rn(a, b)
#define PLG_BITS(d) __double_as_longlong(d)
#else
#define PLG_FMA(a, b, c) fma(a, b, c)
#define PLG_ADD(a, b) ((a) + (b))
#define PLG_SUB(a, b) ((a) - (b))
#define PLG_MUL(a, b) ((a) * (b))
static inline long long plg_bits(double d) { long long u; memcpy(&u, &d, 8); return u; }
#define PLG_BITS(d) plg_bits(d)
#endif
#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace pl {

// tab: 440 doubles, entry k = {sn, ssn, cs, ccs} with sn + ssn = sin(k / 128), cs + ccs = cos(k / 128)  (built by the host)
struct GlibcSinCos {
    const double* tab;
    __host__ __device__ inline double do_cos(double x, double dx) const {
        const double big = 52776558133248.0;  // 1.5 * 2^45
        if (x < 0) dx = -dx;
        const double ax = fabs(x), ux = PLG_ADD(big, ax);
        const int k = (int)(unsigned int)PLG_BITS(ux) * 4;
        x = PLG_ADD(PLG_SUB(ax, PLG_SUB(ux, big)), dx);
        const double xx = PLG_MUL(x, x);
        const double p = PLG_FMA(xx, 8.33333214285722277379541354343671E-03, -1.66666666666664880952546298448555E-01);
        const double s = PLG_FMA(PLG_MUL(x, xx), p, x);
        double q = PLG_FMA(xx, 1.38888874007937613028114285595617E-03, -4.16666666666664434524222570944589E-02);
        q = PLG_FMA(xx, q, 4.99999999999999999999950396842453E-01);
        const double c = PLG_MUL(xx, q);
        const double sn = tab[k], ssn = tab[k + 1], cs = tab[k + 2], ccs = tab[k + 3];
        double cor = PLG_FMA(-s, ssn, ccs);
        cor = PLG_FMA(-c, cs, cor);
        cor = PLG_FMA(-s, sn, cor);
        return PLG_ADD(cs, cor);
    }
    __host__ __device__ inline double do_sin(double x, double dx) const {
        const double big = 52776558133248.0;
        const double xold = x;
        if (fabs(x) < 0.126) {
            const double xx = PLG_MUL(x, x);
            double p = PLG_FMA(xx, -2.5022014848318398e-08, 2.755729806860771e-06);  // s5, s4
            p = PLG_FMA(xx, p, -0.00019841269834414642);                             // s3
            p = PLG_FMA(xx, p, 0.0083333333333323288);                               // s2
            p = PLG_FMA(xx, p, -0.16666666666666666);                                // s1
            const double t = PLG_FMA(PLG_FMA(p, x, -PLG_MUL(0.5, dx)), xx, dx);
            return PLG_ADD(x, t);
        }
        if (x <= 0) dx = -dx;
        const double ax = fabs(x), ux = PLG_ADD(big, ax);
        const int k = (int)(unsigned int)PLG_BITS(ux) * 4;
        x = PLG_SUB(ax, PLG_SUB(ux, big));
        const double xx = PLG_MUL(x, x);
        const double p = PLG_FMA(xx, 8.33333214285722277379541354343671E-03, -1.66666666666664880952546298448555E-01);
        const double s = PLG_ADD(x, PLG_FMA(PLG_MUL(x, xx), p, dx));
        double q = PLG_FMA(xx, 1.38888874007937613028114285595617E-03, -4.16666666666664434524222570944589E-02);
        q = PLG_FMA(xx, q, 4.99999999999999999999950396842453E-01);
        const double c = PLG_FMA(x, dx, PLG_MUL(xx, q));
        const double sn = tab[k], ssn = tab[k + 1], cs = tab[k + 2], ccs = tab[k + 3];
        double cor = PLG_FMA(ccs, s, ssn);
        cor = PLG_FMA(-c, sn, cor);
        cor = PLG_FMA(s, cs, cor);
        return copysign(PLG_ADD(sn, cor), xold);
    }
    __host__ __device__ inline int reduce(double x, double* a, double* da) const {
        const double hpinv = 0.63661977236758138, toint = 6755399441055744.0;
        const double mp1 = 1.5707963407039642333984375, mp2 = -1.3909067564377153e-08, pp3 = -4.9789962314799099e-17, pp4 = -1.9034889620193266e-25;
        const double t = PLG_FMA(x, hpinv, toint), xn = PLG_SUB(t, toint);
        double y = PLG_FMA(-xn, mp1, x);
        y = PLG_FMA(-xn, mp2, y);
        const int n = (int)(unsigned int)PLG_BITS(t) & 3;
        const double t2 = PLG_FMA(-xn, pp3, y);
        double db = PLG_FMA(-pp3, xn, PLG_SUB(y, t2));
        const double b = PLG_FMA(-xn, pp4, t2);
        db = PLG_ADD(db, PLG_FMA(-xn, pp4, PLG_SUB(t2, b)));
        *a = b;
        *da = db;
        return n;
    }
    __host__ __device__ inline double do_sincos(double a, double da, int n) const {
        const double r = (n & 1) ? do_cos(a, da) : do_sin(a, da);
        return (n & 2) ? -r : r;
    }
    __host__ __device__ inline double sin(double x) const {
        const unsigned int k = (unsigned int)((unsigned long long)PLG_BITS(x) >> 32) & 0x7fffffffu;
        if (k < 0x3e500000u) return x;
        if (k < 0x3feb6000u) return do_sin(x, 0.0);
        if (k < 0x400368fdu) return copysign(do_cos(PLG_SUB(1.5707963267948966, fabs(x)), 6.123233995736766e-17), x);
        double a, da;
        const int n = reduce(x, &a, &da);
        return do_sincos(a, da, n);
    }
    __host__ __device__ inline double cos(double x) const {
        const unsigned int k = (unsigned int)((unsigned long long)PLG_BITS(x) >> 32) & 0x7fffffffu;
        if (k < 0x3e400000u) return 1.0;
        if (k < 0x3feb6000u) return do_cos(x, 0.0);
        if (k < 0x400368fdu) {
            const double hp1 = 6.123233995736766e-17;
            const double y = PLG_SUB(1.5707963267948966, fabs(x)), a = PLG_ADD(y, hp1), da = PLG_ADD(PLG_SUB(y, a), hp1);
            return do_sin(a, da);
        }
        double a, da;
        const int n = reduce(x, &a, &da);
        return do_sincos(a, da, n + 1);
    }
};

}  // namespace pl

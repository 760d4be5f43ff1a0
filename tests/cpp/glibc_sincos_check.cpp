// Host-side check of the double-precision sin / cos restatement the line path uses where one ulp decides a discrete outcome
// (csrc/pl_glibc_sincos.cuh, table csrc/pl_sincostab.inc): compiled WITHOUT floating-point contraction (every fused multiply-add
// of the restatement is explicit) and compared with the host libm's sin / cos, which is what the reference computes with.
// Sample: LSD's own arguments (float degrees of fastAtan2 converted to radians, with and without the + pi of region2rect) and
// uniform doubles in [-20, 20].
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../orb_slam2_modification_with-point-and-line-feature_b200/csrc/pl_glibc_sincos.cuh"

static const double kTab[440] = {
#include "../../orb_slam2_modification_with-point-and-line-feature_b200/csrc/pl_sincostab.inc"
};

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 20000000;
    const pl::GlibcSinCos G{kTab};
    std::mt19937_64 rng(11);
    const double kPi = 3.14159265358979323846, kDegToRad = kPi / 180;
    long bad = 0;
    for (long i = 0; i < n; i++) {
        double x;
        if (i & 1) {
            const float deg = (float)((rng() % 3600000) / 10000.0);
            x = (double)deg * kDegToRad;
            if (i & 2) x += kPi;
            if (i & 4) x = -x;
        } else {
            x = ((double)(rng() >> 11) / 9007199254740992.0) * 40.0 - 20.0;
        }
        volatile double vx = x;  // keep the compiler from folding the libm calls
        if (G.sin(x) != sin(vx) || G.cos(x) != cos(vx)) {
            if (bad < 5) printf("x=%.17g: sin %a vs libm %a, cos %a vs libm %a\n", x, G.sin(x), sin(vx), G.cos(x), cos(vx));
            bad++;
        }
    }
    printf("%ld arguments, %ld differ from the host libm\n", n, bad);
    return bad != 0;
}

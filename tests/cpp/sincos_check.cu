// Host-side check of the glibc sinf/cosf restatement used by the BRIEF kernel and of the logf restatement used by the
// PredictScale of the projection searches (csrc/pl_common.cuh): compiled by nvcc as
// a plain host program (the function is __host__ __device__) and compared with libm over a dense sample of [0, 2*pi].
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../orb_slam2_modification_with-point-and-line-feature_b200/csrc/pl_common.cuh"
namespace pl { void set_error(const char*, ...) {} }

int main(int argc, char** argv) {
    const unsigned stride = argc > 1 ? (unsigned)atoi(argv[1]) : 211u;
    float hi = 6.2832f;
    unsigned b;
    memcpy(&b, &hi, 4);
    long bad = 0, tot = 0;
    for (unsigned u = 0; u <= b; u += stride) {
        float f;
        memcpy(&f, &u, 4);
        if (pl::glibc_sincosf(f, 1) != cosf(f) || pl::glibc_sincosf(f, 0) != sinf(f)) bad++;
        tot++;
    }
    const float probes[] = {0.f, 1e-5f, 0x1p-12f, 0.78539f, 0.7853982f, 0.785399f, 1.5707963f, 3.1415927f, 4.712389f, 6.2831855f};
    for (float f : probes)
        if (pl::glibc_sincosf(f, 1) != cosf(f) || pl::glibc_sincosf(f, 0) != sinf(f)) bad++;
    // glibc logf restatement (MapPoint::PredictScale, src/MapPoint.cc:407,424): a strided sweep over all normal positive floats
    for (unsigned u = 0x00800000u; u < 0x7f800000u; u += stride * 2u + 1u) {
        float f;
        memcpy(&f, &u, 4);
        volatile float vf = f;
        if (pl::glibc_logf(f) != logf(vf)) bad++;
        tot++;
    }
    printf("checked %ld mismatches %ld\n", tot, bad);
    return bad != 0;
}

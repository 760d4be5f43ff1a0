#include <cstdio>
#include <quadmath.h>
int main() {
    printf("// sin / cos of k / 128, k = 0 .. 109, as double-double {sn, ssn, cs, ccs}: the double nearest to the value and the double nearest to\n");
    printf("// the remainder (the layout and definition of glibc's __sincostab; generated with libquadmath by tools/gen_sincostab.cpp)\n");
    for (int k = 0; k < 110; k++) {
        __float128 x = (__float128)k / 128, s = sinq(x), c = cosq(x);
        double sn = (double)s, cs = (double)c;
        printf("%a, %a, %a, %a,\n", sn, (double)(s - (__float128)sn), cs, (double)(c - (__float128)cs));
    }
}

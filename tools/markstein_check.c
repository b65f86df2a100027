/* CPU evidence for div_den (offshore-sph_b200/csrc/step.cu): q = v * RN(1/den); r = fma(-q, den, v); q + r * RN(1/den)
   equals v / den bit for bit.  gcc -O2 -ffp-contract=off tools/markstein_check.c -lm && ./a.out  ->  0 mismatches of 660000000 */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static uint64_t s = 88172645463325252ULL;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main(void) {
    double dens[] = {1.025, 1.0, 1.05, 1.5, 1.0000000001, 1.25, 1.3333333333333333, 1.9999, 1.1, 2.0, 1.7};
    long bad = 0, n = 0;
    for (unsigned k = 0; k < sizeof dens / sizeof *dens; k++) {
        volatile double den = dens[k];
        double rden = 1.0 / den;
        for (long i = 0; i < 60000000; i++) {
            uint64_t u = rnd();
            double v;
            if (i & 1) {            /* random bit patterns with moderate exponents */
                uint64_t e = 1023 - 40 + (u % 80);
                uint64_t bits = (u & 0x800fffffffffffffULL) | (e << 52);
                memcpy(&v, &bits, 8);
            } else v = ((double)(u >> 11) / 9007199254740992.0 - 0.5) * 400.0;
            double ref = v / den;
            double q = v * rden;
            double r = fma(-q, den, v);
            double got = fma(r, rden, q);
            n++;
            if (got != ref) { if (bad < 5) printf("den %.17g v %.17g ref %.17g got %.17g\n", den, v, ref, got); bad++; }
        }
    }
    printf("%ld mismatches of %ld\n", bad, n);
    return 0;
}

/* Exhaustive proof that the 3-op quotient used by the CUDA quantiser,
 *     q0 = x*r;  e = fma(-q0, d, x);  q1 = fma(e, r, q0)      (r = RN(1/d))
 * equals the IEEE-754 binary32 division x/d (what pixo's quantize_block performs,
 * src/jpeg/quantize.rs:99-105) for every divisor d in 1..255 and EVERY binary32 significand
 * of x (the sequence is scale-invariant in x away from under/overflow; DCT outputs are far
 * from both).  Also checks round-half-away via the RZ trick used on the device:
 *     n = trunc(RZ(|q| + 0.5)) with sign of q.
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp verify_div.c -lm -o verify_div
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <fenv.h>

static inline float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main(void)
{
    long bad = 0, bad_round = 0;
#pragma omp parallel for reduction(+:bad,bad_round) schedule(dynamic,1)
    for (int d = 1; d <= 255; d++) {
        volatile float df = (float)d;
        volatile float r = 1.0f / df;
        for (uint32_t m = 0; m < (1u << 23); m++) {
            for (int e = 0; e < 2; e++) { /* two binades: [1,2) and [2^7,2^8) (quotient exponent wraps) */
                float x = asf((e ? 0x43000000u : 0x3F800000u) | m);
                float want = x / df;
                float q0 = x * r;
                float er = fmaf(-q0, df, x);
                float q1 = fmaf(er, r, q0);
                if (q1 != want) bad++;
            }
        }
    }
    printf("divisor sweep: mismatches = %ld (of %ld)\n", bad, 255L * (1L << 24));
    /* rounding trick: for a sample of quotients incl. all k+0.5 ties and neighbours */
    fesetround(FE_TOWARDZERO);
    for (int k = 0; k < 40000; k++) {
        for (int j = -2; j <= 2; j++) {
            float base = (float)k + 0.5f;
            uint32_t u; memcpy(&u, &base, 4); u += (uint32_t)j; float q = asf(u);
            volatile float a = q + 0.5f; /* RZ */
            long n = (long)a;            /* trunc */
            fesetround(FE_TONEAREST);
            long want = (long)roundf(q);
            fesetround(FE_TOWARDZERO);
            if (n != want) bad_round++;
        }
    }
    fesetround(FE_TONEAREST);
    printf("round trick: mismatches = %ld\n", bad_round);
    return (bad || bad_round) ? 1 : 0;
}

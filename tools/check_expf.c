/* check_expf.c -- exhaustive comparison of include/mkf_expf.h with the host libm's expf.
 *   gcc -O2 -ffp-contract=off -mfma -fopenmp -o /tmp/check_expf tools/check_expf.c -lm && /tmp/check_expf
 * Walks every float bit pattern in [lo, hi] (default: all finite floats, NaNs and infinities included) and prints
 * the number of arguments whose result differs in any bit.  Used once per libm version; tests/test_expf.py runs
 * the sampled version of the same comparison. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mkf_expf.h"

static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main(int argc, char** argv)
{
    uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
    uint64_t bad = 0, n = 0, n_path = 0;
#pragma omp parallel for reduction(+ : bad, n, n_path) schedule(static)
    for (uint64_t b = 0; b < (1ull << 32); b += stride) {
        const float x = u2f((uint32_t)b);
        const float a = expf(x), m = mkf_expf(x);
        n++;
        if (x <= 0.0f && x >= -104.0f) n_path++; /* the range src/pf2D.cpp:108 can produce (-q/2, q >= 0) */
        if (f2u(a) != f2u(m) && !(a != a && m != m)) {
            bad++;
            if (bad < 10) fprintf(stderr, "x=%a libm=%a mkf=%a\n", x, a, m);
        }
    }
    printf("{\"arguments\": %llu, \"in_[-104,0]\": %llu, \"mismatches\": %llu, \"stride\": %llu}\n",
           (unsigned long long)n, (unsigned long long)n_path, (unsigned long long)bad, (unsigned long long)stride);
    return bad != 0;
}

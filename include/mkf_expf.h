/* mkf_expf.h -- single-precision exp with ONE defined result per argument, shared by the CUDA kernels
 * (nvcc, device side) and the CPU oracle (gcc).
 *
 * Why it exists.  The legacy particle filter evaluates its GMM prior through `expf(float(q))`
 * (src/pf2D.cpp:105-109, quirk B12).  `expf` is libm's, and libm's result is not correctly rounded: the
 * resampled particle indices of src/pf2D.cpp:225-268 depend on the last bit of every weight, so "the
 * reference's expf" has to be pinned to an algorithm, not to "some float exp".  The algorithm pinned here
 * is the one glibc ships since 2.27 (sysdeps/ieee754/flt-32/e_expf.c, the ARM optimized-routines
 * expf: x*N/ln2 = k + r, 2^(k/N) from a 32-entry table, a cubic in r, everything in double, one final
 * rounding to float).  It is restated from the published description of that algorithm; the table is
 * T[i] = bits(2^(i/32)) - (i << 47), regenerated with mpmath by tools/make_expf_table.py.
 * Which build of it: on x86-64 glibc selects its FMA build of e_expf.c on every CPU that has FMA (all the
 * hosts this runs on), and in that build the reduction r = x*N/ln2 - k is one fused multiply-add; the
 * other operations round individually (whether they are fused makes no difference to any result: all 16
 * combinations were run over every float).  That build is the one restated here.  tools/check_expf.c
 * compares it with the host libm's expf over ALL 2^32 float bit patterns: 0 mismatches against glibc 2.39
 * in this image (the non-FMA build differs for two arguments, -0x1.f8cbb2p+5 and 0x1.04845ep+5, by one
 * ulp).  tests/test_expf.py repeats the comparison on a dense sample (exhaustive with MKF_EXPF_EXHAUSTIVE=1).
 *
 * Every operation is an explicitly rounded IEEE double operation (_rn intrinsics on the device; plain
 * operators under -ffp-contract=off plus fma() on the host) and the conversions are exact, hence the
 * result is bit-identical on CPU and GPU.
 */
#ifndef MKF_EXPF_H
#define MKF_EXPF_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MKF_EXPF_HD __host__ __device__ __forceinline__
#else
#define MKF_EXPF_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define MKF_EMUL(a, b) __dmul_rn((a), (b))
#define MKF_EADD(a, b) __dadd_rn((a), (b))
#define MKF_ESUB(a, b) __dsub_rn((a), (b))
#define MKF_EFMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define MKF_EMUL(a, b) ((a) * (b))
#define MKF_EADD(a, b) ((a) + (b))
#define MKF_ESUB(a, b) ((a) - (b))
#define MKF_EFMA(a, b, c) fma((a), (b), (c))
#endif

/* bits(2^(i/32)) - (i << 47) */
#define MKF_EXPF_TABLE                                                                                  \
    {0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,        \
     0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,        \
     0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,        \
     0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,        \
     0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,        \
     0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,        \
     0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,        \
     0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull}
static const uint64_t mkf_expf_T_host[32] = MKF_EXPF_TABLE;
#if defined(__CUDACC__)
static __device__ const uint64_t mkf_expf_T_dev[32] = MKF_EXPF_TABLE; /* read through the L1 (divergent index) */
#endif

MKF_EXPF_HD uint64_t mkf_expf_tab(int i)
{
#if defined(__CUDA_ARCH__)
    return __ldg(mkf_expf_T_dev + i);
#else
    return mkf_expf_T_host[i];
#endif
}

MKF_EXPF_HD double mkf_expf_asdouble(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}
MKF_EXPF_HD uint64_t mkf_expf_asuint64(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}

/* expf(x).  NaN -> NaN, +inf / overflow -> +inf, -inf / underflow below 2^-150 -> +0. */
MKF_EXPF_HD float mkf_expf(float x)
{
    const double xd = (double)x;
    if (!(x == x)) return x + x;
    if (x > 0x1.62e42ep6f) return __builtin_huge_valf(); /* x > log(2^128)  ~  88.72 */
    if (x < -0x1.9fe368p6f) return 0.0f;                 /* x < log(2^-150) ~ -103.97 */
    /* z = x * 32/ln2 = k + r, |r| <= 1/2, k by the add-and-subtract-2^52*1.5 rounding (ties to even) */
    const double SHIFT = 0x1.8p+52;
    const double z = MKF_EMUL(0x1.71547652b82fep+5 /* 32 / ln 2 */, xd);
    double kd = MKF_EADD(z, SHIFT);
    const uint64_t ki = mkf_expf_asuint64(kd);
    kd = MKF_ESUB(kd, SHIFT);
    const double r = MKF_EFMA(0x1.71547652b82fep+5, xd, -kd); /* fused, as glibc's FMA build does */
    /* exp(x) = 2^(k/32) * 2^(r/32) ~= s * (C0 r^3 + C1 r^2 + C2 r + 1) */
    const uint64_t t = mkf_expf_tab((int)(ki & 31u)) + (ki << 47);
    const double s = mkf_expf_asdouble(t);
    const double C0 = 0x1.c6af84b912394p-20; /* 0x1.c6af84b912394p-5 / 32^3 */
    const double C1 = 0x1.ebfce50fac4f3p-13; /* 0x1.ebfce50fac4f3p-3 / 32^2 */
    const double C2 = 0x1.62e42ff0c52d6p-6;  /* 0x1.62e42ff0c52d6p-1 / 32   */
    const double p = MKF_EADD(MKF_EMUL(C0, r), C1);
    const double r2 = MKF_EMUL(r, r);
    double y = MKF_EADD(MKF_EMUL(C2, r), 1.0);
    y = MKF_EADD(MKF_EMUL(p, r2), y);
    y = MKF_EMUL(y, s);
    return (float)y;
}

#endif /* MKF_EXPF_H */

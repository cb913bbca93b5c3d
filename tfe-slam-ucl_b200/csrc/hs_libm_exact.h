/* sinf / cosf / expf that reproduce glibc 2.39's results bit for bit, usable on host and device.
 *
 * Why: the reference (hector_slam_lib, CPU) gets its transcendental values from the host libm
 * (sin/cos at HSL/map/OccGridMapUtil.h:70-71 and inside Eigen::Rotation2Df; exp at
 * HSL/map/GridMapLogOdds.h:165). A beam endpoint cell is (int)(x + 0.5f) of a product with cosf/sinf
 * (HSL/map/OccGridMapBase.h:148-155), so a 1-ulp difference in the trig value can move an endpoint to
 * the neighbouring cell. CUDA's sinf/cosf/expf are not glibc's, so "bit-exact Bresenham cell sets" needs
 * the same values. glibc >= 2.28 computes these three functions in double precision with short
 * polynomials (the ARM optimized-routines algorithms): sysdeps/ieee754/flt-32/{s_sinf.c,s_cosf.c,
 * sincosf.h,sincosf_data.c,e_expf.c,e_exp2f_data.c}. glibc is a third-party dependency of the reference
 * that is not under /root/reference; its published algorithm is restated here, with the coefficient
 * tables checked against the constants found in this image's /lib/x86_64-linux-gnu/libm.so.6
 * (tests/test_libm_exact.py compares results with the host libm). IEEE double add/mul/fma behave
 * identically on sm_100a and x86-64. On FMA-capable x86-64 hosts glibc's ifunc selects the *_fma builds
 * of these functions (the same C compiled with -mfma, so every a*b+c is one fused operation); the
 * explicit hs_fma calls below reproduce that contraction pattern. Measured here with
 * tests/tools/libm_exact_check.c over ALL floats in the domain (2.2e9 sinf, 2.2e9 cosf, 4.3e9 expf
 * arguments): 0 mismatches against this image's glibc 2.39 with the fused pattern; 12 / 22 / 2
 * mismatches (1 ulp, near multiples of pi/2) with the unfused one. Define HS_LIBM_NO_FMA to get the
 * unfused flavour (what glibc computes on a CPU without FMA).
 *
 * Domain: sinf/cosf follow glibc's fast paths for |x| < 120 (poses are normalised to (-pi, pi] so the
 * hot path never leaves it); beyond that they fall back to double sin/cos rounded to float.
 */
#ifndef HS_LIBM_EXACT_H
#define HS_LIBM_EXACT_H

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define HS_HD __host__ __device__ __forceinline__
#else
#define HS_HD static inline
#endif

HS_HD uint32_t hs_asuint(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, sizeof u); return u;
#endif
}
HS_HD uint64_t hs_asuint64(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u; memcpy(&u, &d, sizeof u); return u;
#endif
}
HS_HD double hs_asdouble(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double d; memcpy(&d, &u, sizeof d); return d;
#endif
}
/* Explicit double ops: the build uses -fmad=false / -ffp-contract=off, so fusion happens only where
 * hs_fma says so. */
HS_HD double hs_fma(double a, double b, double c) {
#if defined(HS_LIBM_NO_FMA)
#if defined(__CUDA_ARCH__)
  return __dadd_rn(__dmul_rn(a, b), c);
#else
  return a * b + c;
#endif
#else
  return fma(a, b, c);
#endif
}
HS_HD double hs_dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
HS_HD double hs_dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}

/* abstop12 (sincosf.h): top 12 bits of |x| */
HS_HD uint32_t hs_abstop12(float x) { return (hs_asuint(x) >> 20) & 0x7ff; }

/* __sincosf_table (sincosf_data.c), !TOINT_INTRINSICS flavour: hpi_inv = 2/pi * 2^24. */
#define HS_SC_HPI_INV 0x1.45F306DC9C883p+23
#define HS_SC_HPI 0x1.921FB54442D18p0
#define HS_SC_C0 0x1p0
#define HS_SC_C1 -0x1.ffffffd0c621cp-2
#define HS_SC_C2 0x1.55553e1068f19p-5
#define HS_SC_C3 -0x1.6c087e89a359dp-10
#define HS_SC_C4 0x1.99343027bf8c3p-16
#define HS_SC_S1 -0x1.555545995a603p-3
#define HS_SC_S2 0x1.1107605230bc4p-7
#define HS_SC_S3 -0x1.994eb3774cf24p-13

/* sinf_poly (sincosf.h): `neg` selects __sincosf_table[1], whose cosine coefficients are negated. */
HS_HD float hs_sinf_poly(double x, double x2, int neg, int n) {
  if ((n & 1) == 0) {
    double x3 = hs_dmul(x, x2);
    double s1 = hs_fma(x2, HS_SC_S3, HS_SC_S2);
    double x7 = hs_dmul(x3, x2);
    double s = hs_fma(x3, HS_SC_S1, x);
    return (float)hs_fma(x7, s1, s);
  } else {
    double sg = neg ? -1.0 : 1.0;
    double x4 = hs_dmul(x2, x2);
    double c2 = hs_fma(x2, sg * HS_SC_C4, sg * HS_SC_C3);
    double c1 = hs_fma(x2, sg * HS_SC_C1, sg * HS_SC_C0);
    double x6 = hs_dmul(x4, x2);
    double c = hs_fma(x4, sg * HS_SC_C2, c1);
    return (float)hs_fma(x6, c2, c);
  }
}

/* reduce_fast (sincosf.h), !TOINT_INTRINSICS: quadrant from a scaled, explicitly rounded conversion. */
HS_HD double hs_reduce_fast(double x, int* np) {
  double r = hs_dmul(x, HS_SC_HPI_INV);
  int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return hs_fma(-(double)n, HS_SC_HPI, x);
}

/* s_sinf.c */
HS_HD float hs_sinf(float y) {
  double x = (double)y;
  if (hs_abstop12(y) < hs_abstop12(0x1.921FB6p-1f)) {
    double s = hs_dmul(x, x);
    if (hs_abstop12(y) < hs_abstop12(0x1p-12f)) return y;
    return hs_sinf_poly(x, s, 0, 0);
  } else if (hs_abstop12(y) < hs_abstop12(120.0f)) {
    int n;
    x = hs_reduce_fast(x, &n);
    double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0; /* sign[4] = {1,-1,-1,1} */
    return hs_sinf_poly(hs_dmul(x, s), hs_dmul(x, x), (n & 2) != 0, n);
  }
  return (float)sin(x);
}

/* s_cosf.c */
HS_HD float hs_cosf(float y) {
  double x = (double)y;
  if (hs_abstop12(y) < hs_abstop12(0x1.921FB6p-1f)) {
    if (hs_abstop12(y) < hs_abstop12(0x1p-12f)) return 1.0f;
    return hs_sinf_poly(x, hs_dmul(x, x), 0, 1);
  } else if (hs_abstop12(y) < hs_abstop12(120.0f)) {
    int n;
    x = hs_reduce_fast(x, &n);
    double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return hs_sinf_poly(hs_dmul(x, s), hs_dmul(x, x), (n & 2) != 0, n ^ 1);
  }
  return (float)cos(x);
}

/* __exp2f_data.tab (e_exp2f_data.c): tab[i] = asuint64(2^(i/32)) - (i << (52-5)) */
#if defined(__CUDACC__)
#define HS_TAB_QUAL __device__ __constant__
#else
#define HS_TAB_QUAL static const
#endif
#define HS_EXP2F_TAB_INIT { \
  0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, \
  0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, \
  0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL, \
  0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, \
  0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, \
  0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, \
  0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, \
  0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL }
#if defined(__CUDACC__)
static __device__ __constant__ uint64_t hs_exp2f_tab_dev[32] = HS_EXP2F_TAB_INIT;
#endif
static const uint64_t hs_exp2f_tab_host[32] = HS_EXP2F_TAB_INIT;

#define HS_EXPF_INVLN2N 0x1.71547652b82fep+5 /* 32 / ln 2 */
#define HS_EXPF_SHIFT 0x1.8p+52
#define HS_EXPF_C0 0x1.c6af84b912394p-20 /* poly_scaled */
#define HS_EXPF_C1 0x1.ebfce50fac4f3p-13
#define HS_EXPF_C2 0x1.62e42ff0c52d6p-6

/* e_expf.c; `tab` = the 32-entry __exp2f_data.tab (callers on the device may keep a copy in shared memory: lanes of a
 * warp look up different entries, which a __constant__ bank serialises) */
HS_HD float hs_expf_tab(float x, const uint64_t* tab) {
  uint32_t abstop = (hs_asuint(x) >> 20) & 0x7ff;
  if (abstop >= ((hs_asuint(88.0f) >> 20) & 0x7ff)) {
    if (hs_asuint(x) == hs_asuint(-INFINITY)) return 0.0f;
    if (abstop >= ((hs_asuint(INFINITY) >> 20) & 0x7ff)) return x + x;
    if (x > 0x1.62e42ep6f) return INFINITY;  /* x > log(0x1p128) */
    if (x < -0x1.9fe368p6f) return 0.0f;     /* x < log(0x1p-150) */
  }
  double xd = (double)x;
  /* z = InvLn2N * xd feeds only the two sums below, so the fused build folds the product into both */
  double kd = hs_fma(HS_EXPF_INVLN2N, xd, HS_EXPF_SHIFT);
  uint64_t ki = hs_asuint64(kd);
  kd = hs_dadd(kd, -HS_EXPF_SHIFT);
  double r = hs_fma(HS_EXPF_INVLN2N, xd, -kd);
  double z;
  uint64_t t = tab[ki % 32];
  t += ki << (52 - 5);
  double s = hs_asdouble(t);
  z = hs_fma(HS_EXPF_C0, r, HS_EXPF_C1);
  double r2 = hs_dmul(r, r);
  double y = hs_fma(HS_EXPF_C2, r, 1.0);
  y = hs_fma(z, r2, y);
  y = hs_dmul(y, s);
  return (float)y;
}

HS_HD float hs_expf(float x) {
#if defined(__CUDA_ARCH__)
  return hs_expf_tab(x, hs_exp2f_tab_dev);
#else
  return hs_expf_tab(x, hs_exp2f_tab_host);
#endif
}

#endif /* HS_LIBM_EXACT_H */

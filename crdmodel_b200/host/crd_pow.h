/* crd_pow.h — x^y for x > 0 as ONE fixed sequence of IEEE-754 double operations (+, -, *, / and exact scalings), so that the
 * host (crd_ark.cpp: the step controller of the explicit RK driver) and the device (crd_resident.cu: the same controller inside
 * the persistent step-loop kernel) compute the same bits.  libm's pow and the CUDA pow differ in the last place; with this the
 * device-resident loop picks exactly the step sizes of the host-driven loop and of the CPU run.
 * ~1e-15 relative accuracy (the controller needs far less).  Translation units that include this are compiled without
 * floating-point contraction (-ffp-contract=off on the host; explicit _rn intrinsics on the device). */
#ifndef CRD_POW_H
#define CRD_POW_H

#if defined(__CUDA_ARCH__)
#define CRD_POW_FN __device__ __forceinline__
#define CRD_PMUL(a, b) __dmul_rn((a), (b))
#define CRD_PADD(a, b) __dadd_rn((a), (b))
#define CRD_PSUB(a, b) __dsub_rn((a), (b))
#define CRD_PDIV(a, b) __ddiv_rn((a), (b))
#define CRD_PBITS(x) ((unsigned long long)__double_as_longlong(x))
#define CRD_PFROMBITS(u) __longlong_as_double((long long)(u))
#else
#include <string.h>
#define CRD_POW_FN static inline
#define CRD_PMUL(a, b) ((a) * (b))
#define CRD_PADD(a, b) ((a) + (b))
#define CRD_PSUB(a, b) ((a) - (b))
#define CRD_PDIV(a, b) ((a) / (b))
static inline unsigned long long crd_pow_bits(double x) { unsigned long long u; memcpy(&u, &x, 8); return u; }
static inline double crd_pow_frombits(unsigned long long u) { double x; memcpy(&x, &u, 8); return x; }
#define CRD_PBITS(x) crd_pow_bits(x)
#define CRD_PFROMBITS(u) crd_pow_frombits(u)
#endif

/* x: positive, normal.  |y log x| < 700. */
CRD_POW_FN double crd_pow_pos(double x, double y) {
  /* x = m 2^e with m in [sqrt(1/2), sqrt(2)) */
  unsigned long long u = CRD_PBITS(x);
  int e = (int)((u >> 52) & 0x7ffULL) - 1023;
  double m = CRD_PFROMBITS((u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);   /* [1, 2) */
  if (m >= 1.4142135623730951) { m = CRD_PMUL(m, 0.5); e += 1; }
  /* log m = 2 atanh(s), s = (m - 1) / (m + 1), |s| <= 0.1716: odd series to s^25 */
  const double s = CRD_PDIV(CRD_PSUB(m, 1.0), CRD_PADD(m, 1.0));
  const double s2 = CRD_PMUL(s, s);
  double p = 1.0 / 25.0;
  p = CRD_PADD(CRD_PMUL(p, s2), 1.0 / 23.0);
  p = CRD_PADD(CRD_PMUL(p, s2), 1.0 / 21.0);
  p = CRD_PADD(CRD_PMUL(p, s2), 1.0 / 19.0);
  p = CRD_PADD(CRD_PMUL(p, s2), 1.0 / 17.0);
  p = CRD_PADD(CRD_PMUL(p, s2), 1.0 / 15.0);
  p = CRD_PADD(CRD_PMUL(p, s2), 1.0 / 13.0);
  p = CRD_PADD(CRD_PMUL(p, s2), 1.0 / 11.0);
  p = CRD_PADD(CRD_PMUL(p, s2), 1.0 / 9.0);
  p = CRD_PADD(CRD_PMUL(p, s2), 1.0 / 7.0);
  p = CRD_PADD(CRD_PMUL(p, s2), 1.0 / 5.0);
  p = CRD_PADD(CRD_PMUL(p, s2), 1.0 / 3.0);
  const double logm = CRD_PMUL(2.0, CRD_PADD(s, CRD_PMUL(s, CRD_PMUL(p, s2))));
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double logx = CRD_PADD(CRD_PMUL((double)e, ln2_hi), CRD_PADD(logm, CRD_PMUL((double)e, ln2_lo)));
  /* exp(t), t = y log x = k ln2 + r, |r| <= 0.35: Taylor to r^14 */
  const double t = CRD_PMUL(y, logx);
  const double kf = CRD_PMUL(t, 1.44269504088896338700e+00);
  const int k = (int)(kf >= 0.0 ? CRD_PADD(kf, 0.5) : CRD_PSUB(kf, 0.5));   /* round half away from zero (truncating conversion) */
  const double r = CRD_PSUB(CRD_PSUB(t, CRD_PMUL((double)k, ln2_hi)), CRD_PMUL((double)k, ln2_lo));
  double q = 1.0 / 87178291200.0;                 /* 1/14! */
  q = CRD_PADD(CRD_PMUL(q, r), 1.0 / 6227020800.0);
  q = CRD_PADD(CRD_PMUL(q, r), 1.0 / 479001600.0);
  q = CRD_PADD(CRD_PMUL(q, r), 1.0 / 39916800.0);
  q = CRD_PADD(CRD_PMUL(q, r), 1.0 / 3628800.0);
  q = CRD_PADD(CRD_PMUL(q, r), 1.0 / 362880.0);
  q = CRD_PADD(CRD_PMUL(q, r), 1.0 / 40320.0);
  q = CRD_PADD(CRD_PMUL(q, r), 1.0 / 5040.0);
  q = CRD_PADD(CRD_PMUL(q, r), 1.0 / 720.0);
  q = CRD_PADD(CRD_PMUL(q, r), 1.0 / 120.0);
  q = CRD_PADD(CRD_PMUL(q, r), 1.0 / 24.0);
  q = CRD_PADD(CRD_PMUL(q, r), 1.0 / 6.0);
  q = CRD_PADD(CRD_PMUL(q, r), 0.5);
  q = CRD_PADD(CRD_PMUL(q, r), 1.0);
  q = CRD_PADD(CRD_PMUL(q, r), 1.0);
  return CRD_PMUL(q, CRD_PFROMBITS((unsigned long long)(k + 1023) << 52));   /* exact scaling by 2^k (|k| < 1022) */
}

#endif /* CRD_POW_H */

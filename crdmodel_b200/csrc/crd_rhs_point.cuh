// crd_rhs_point.cuh — per-point arithmetic of the right-hand side (EXACT / FAST stencil, FHN / Goldbeter kinetics)
// and the state access helpers (plain vector, or the stage combination sum_j c_j x_j formed on the fly).  Shared by
// the launch-per-evaluation kernels (crd_rhs_kernels.cuh) and the device-resident stepper (crd_resident.cu), so both
// produce the same bits for the same state.
#pragma once
#include "crd_fused.cuh"
#include "crd_grid.cuh"

using namespace crd;

namespace {

constexpr double kEps = 0.36;   // EPSILON, FHNmodel_torus.cpp:68
constexpr double kPI = 3.1415926535897932;  // FHNmodel_torus.cpp:63
// Goldbeter constants, GoldbeterModel_torus.cpp:67-78
constexpr double G_v0 = 1.0, G_k = 10.0, G_kf = 1.0, G_v1 = 7.3, G_VM2 = 65.0, G_VM3 = 500.0;
constexpr double G_K2 = 1.0, G_KR = 2.0, G_KA = 0.9, G_m = 2.0, G_n = 2.0, G_p = 4.0;

__host__ __device__ constexpr bool is_torus(int model) { return model == CRD_FHN_TORUS || model == CRD_GOLDBETER_TORUS; }
__host__ __device__ constexpr bool is_fhn(int model) { return model == CRD_FHN_TORUS || model == CRD_FHN_FLAT; }

// ---- per-point arithmetic ---------------------------------------------------------------------------
// Correctly rounded a / c for a divisor known in advance, rc = RN(1/c) from the host's IEEE division.
// q0 = RN(a*rc) is within 1.5 ulp of a/c; one residual step makes it faithful, and by Markstein's
// theorem (q faithful, rc = RN(1/c), r = a - c*q exact through FMA  =>  RN(q + r*rc) = RN(a/c)) the
// second step is the correctly rounded quotient: 5 FP64 issues instead of the ~20 of a general
// division.  Outside the safely normal range (tiny, huge, inf/nan) the caller falls back to the IEEE
// division so subnormals also match.
// Straight-line form used by the stencil (so the three divisions of a point and the points of a thread
// interleave and hide the FP64 latency).  Needs c > 0.  The residual is formed as r' = q*c - a and
// subtracted, which makes a zero numerator come out as the correctly signed zero with no special case:
//   a = -0: q0 = -0, r' = fma(-0, c, +0) = +0, q = fma(-(+0), rc, -0) = -0;   a = +0: likewise +0.
__device__ __forceinline__ double div_const_line(double a, double c, double rc) {
  const double q0 = __dmul_rn(a, rc);
  double r = __fma_rn(q0, c, -a);
  const double q1 = __fma_rn(-r, rc, q0);
  r = __fma_rn(q1, c, -a);
  return __fma_rn(-r, rc, q1);
}
// true when the numerator is outside the range where div_const_line is proven (|n| in [2^-800, 2^800),
// divisor within 2^+-90, checked on the host) and is not an exact zero; integer tests only
__device__ __forceinline__ bool div_needs_ieee(double n) {
  const unsigned hi = (unsigned)__double2hiint(n) & 0x7fffffffu;
  const bool inrange = (hi - 0x0DF00000u) < 0x64000000u;   // biased exponent in [223, 1823)
  return !inrange && (hi | (unsigned)__double2loint(n)) != 0u;
}

// 1/x to ~1 ulp without the IEEE slow path (FAST arithmetic only; x is a sum of positive terms here)
__device__ __forceinline__ double rcp_fast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// the same sum with IEEE divisions; kept out of line so the unrolled hot loop does not carry 3 divisions per row
__device__ __noinline__ double stencil_sum_ieee(double n1, double n2, double n3, double c1, double c2, double c3) {
  return __dadd_rn(__dadd_rn(__ddiv_rn(n1, c1), __ddiv_rn(n2, c2)), __ddiv_rn(n3, c3));
}

// all three numerators at once: min / max of the high words decide the common case in 8 integer instructions
__device__ __forceinline__ bool div3_needs_ieee(double n1, double n2, double n3) {
  const unsigned t1 = (unsigned)__double2hiint(n1) & 0x7fffffffu, t2 = (unsigned)__double2hiint(n2) & 0x7fffffffu,
                 t3 = (unsigned)__double2hiint(n3) & 0x7fffffffu;
  const unsigned mn = min(t1, min(t2, t3)), mx = max(t1, max(t2, t3));
  if (mn >= 0x0DF00000u && mx < 0x71F00000u) return false;        // every |n| in [2^-800, 2^800)
  return div_needs_ieee(n1) || div_needs_ieee(n2) || div_needs_ieee(n3);   // zeros are fine, the rest is not
}

// EXACT: the reference's expression tree with separately rounded operations.
template <int MODEL>
__device__ __forceinline__ double stencil_exact(const RhsConst &k, double a1, double a3, double uC, double uW,
                                                double uE, double uS, double uN) {
  if (is_torus(MODEL)) {
    // :535-537   Diff*(a1*(uE-uW))/(2dx) + Diff*((1/r^2)*(uE-2uC+uW))/(dx*dx) + Diff*(a3*(uN-2uC+uS))/(dy*dy)
    const double two_uC = __dmul_rn(2.0, uC);
    const double n1 = __dmul_rn(k.Diff, __dmul_rn(a1, __dsub_rn(uE, uW)));
    const double n2 = __dmul_rn(k.Diff, __dmul_rn(k.inv_rr, __dadd_rn(__dsub_rn(uE, two_uC), uW)));
    const double n3 = __dmul_rn(k.Diff, __dmul_rn(a3, __dadd_rn(__dsub_rn(uN, two_uC), uS)));
    double T1 = div_const_line(n1, k.twodx, k.r_twodx);
    double T2 = div_const_line(n2, k.dxdx, k.r_dxdx);
    double T3 = div_const_line(n3, k.dydy, k.r_dydy);
    if (!k.div_safe || div3_needs_ieee(n1, n2, n3))
      return stencil_sum_ieee(n1, n2, n3, k.twodx, k.dxdx, k.dydy);  // tiny / huge / non-finite numerator: rare, out of line
    return __dadd_rn(__dadd_rn(T1, T2), T3);
  } else {
    // FHNmodel_flat.cpp:496-498   cu1*(uW+uE) + cu2*(uS+uN) + cu3*uC
    return __dadd_rn(__dadd_rn(__dmul_rn(k.cu1, __dadd_rn(uW, uE)), __dmul_rn(k.cu2, __dadd_rn(uS, uN))),
                     __dmul_rn(k.cu3, uC));
  }
}

// Same sum, but instead of branching per point it ORs "this point needs the IEEE path" into `bad`; the caller
// redoes the flagged thread's rows afterwards.  Keeps the marched rows free of control flow so they interleave.
template <int MODEL>
__device__ __forceinline__ double stencil_exact_acc(const RhsConst &k, double a1, double a3, double uC, double uW,
                                                    double uE, double uS, double uN, bool &bad) {
  if (is_torus(MODEL)) {
    const double two_uC = __dmul_rn(2.0, uC);
    const double n1 = __dmul_rn(k.Diff, __dmul_rn(a1, __dsub_rn(uE, uW)));
    const double n2 = __dmul_rn(k.Diff, __dmul_rn(k.inv_rr, __dadd_rn(__dsub_rn(uE, two_uC), uW)));
    const double n3 = __dmul_rn(k.Diff, __dmul_rn(a3, __dadd_rn(__dsub_rn(uN, two_uC), uS)));
    bad = bad | div_needs_ieee(n1) | div_needs_ieee(n2) | div_needs_ieee(n3);
    return __dadd_rn(__dadd_rn(div_const_line(n1, k.twodx, k.r_twodx), div_const_line(n2, k.dxdx, k.r_dxdx)),
                     div_const_line(n3, k.dydy, k.r_dydy));
  } else {
    return stencil_exact<MODEL>(k, a1, a3, uC, uW, uE, uS, uN);
  }
}

template <int MODEL>
__device__ __forceinline__ double stencil_fast(const RhsConst &k, double c1, double c3, double uC, double uW,
                                               double uE, double uS, double uN) {
  if (is_torus(MODEL)) {
    const double m2 = -2.0 * uC;
    return c1 * (uE - uW) + k.c2 * ((uE + m2) + uW) + c3 * ((uN + m2) + uS);
  } else {
    return k.cu1 * (uW + uE) + k.cu2 * (uS + uN) + k.cu3 * uC;
  }
}

// x^4 rounded once (libm's pow(x, 4.0) is correctly rounded for nearly every argument, (x*x)*(x*x) is not)
__device__ __forceinline__ double pow4_rn(double x, double x2) {
  const double e2 = __fma_rn(x, x, -x2);          // x*x = x2 + e2 exactly
  const double p = __dmul_rn(x2, x2);
  const double pe = __fma_rn(x2, x2, -p);         // x2*x2 = p + pe exactly
  return __dadd_rn(p, __fma_rn(__dmul_rn(2.0, x2), e2, pe));
}

template <int MODEL, bool EXACT>
__device__ __forceinline__ void react(const RhsConst &k, double b, double u, double v, double &du, double &dv) {
  if (is_fhn(MODEL)) {
    if (EXACT) {
      // :657  ydot_u += 3u - u*u*u - v      :660  ydot_v += EPSILON*(u + b)
      du = __dadd_rn(du, __dsub_rn(__dsub_rn(__dmul_rn(3.0, u), __dmul_rn(__dmul_rn(u, u), u)), v));
      // ydot_v starts at 0.0 (N_VConst :506): 0.0 + x differs from x only for x = -0, i.e. u = b = -0
      dv = __dmul_rn(kEps, __dadd_rn(u, b));
      if (k.dv_plus0) dv = __dadd_rn(0.0, dv);
    } else {
      du += (3.0 * u - u * u * u) - v;
      dv = kEps * (u + b);
    }
  } else {
    // GoldbeterModel_torus.cpp:694-695,715-716; b carries v0 + v1*beta(phi)
    const double Z = u, Y = v;
    if (EXACT) {
      const double z2 = __dmul_rn(Z, Z), y2 = __dmul_rn(Y, Y);
      const double z4 = pow4_rn(Z, z2);
      const double v2 = __ddiv_rn(__dmul_rn(G_VM2, z2), __dadd_rn(k.k2n, z2));
      const double v3 = __ddiv_rn(__dmul_rn(__dmul_rn(G_VM3, y2), z4),
                                  __dmul_rn(__dadd_rn(k.krm, y2), __dadd_rn(k.kap, z4)));
      const double kfY = (G_kf == 1.0) ? Y : __dmul_rn(G_kf, Y);   // 1.0 * Y is Y, bit for bit
      du = __dadd_rn(du, __dsub_rn(__dadd_rn(__dadd_rn(__dsub_rn(b, v2), v3), kfY), __dmul_rn(G_k, Z)));
      dv = __dsub_rn(__dsub_rn(v2, v3), kfY);   // never -0 (v2 - v3 is +0 when it vanishes), so 0.0 + dv == dv
    } else {
      // w = v2 - v3 = A/B - C/D with one reciprocal: (A*D - C*B) / (B*D)
      const double z2 = Z * Z, y2 = Y * Y, z4 = z2 * z2;
      const double A = G_VM2 * z2, B = k.k2n + z2;
      const double Cn = (G_VM3 * y2) * z4, Dn = (k.krm + y2) * (k.kap + z4);
      const double w = (A * Dn - Cn * B) * rcp_fast(B * Dn);
      du += ((b - w) + Y) - G_k * Z;
      dv = w - Y;
    }
  }
}

// ---- state access: plain vector, or sum_j c_j x_j formed on the fly ------------------------------------------------
// SEQ (= EXACT grids): the combination is rounded like the op-by-op stage assembly of the RK driver (crd_fused.cuh);
// otherwise a chain of fused multiply-adds in the order of lincomb_kernel.
template <bool LC, bool SEQ>
__device__ __forceinline__ double2 state2(const RhsArgs &a, long long p) {
  if constexpr (!LC) {
    return reinterpret_cast<const double2 *>(a.y)[p];
  } else {
    double2 v[kMaxLc];
#pragma unroll
    for (int j = 0; j < kMaxLc; ++j)
      v[j] = (j < a.nlc) ? reinterpret_cast<const double2 *>(a.lc_x[j])[p] : make_double2(0.0, 0.0);
    double vx[kMaxLc], vy[kMaxLc];
#pragma unroll
    for (int j = 0; j < kMaxLc; ++j) { vx[j] = v[j].x; vy[j] = v[j].y; }
    return make_double2(lc_value<SEQ, kMaxLc>(a.lc_c, vx, a.nlc), lc_value<SEQ, kMaxLc>(a.lc_c, vy, a.nlc));
  }
}
template <bool LC, bool SEQ>
__device__ __forceinline__ double stateu(const RhsArgs &a, long long p) {
  if constexpr (!LC) {
    return a.y[2 * p];
  } else {
    double v[kMaxLc];
#pragma unroll
    for (int j = 0; j < kMaxLc; ++j) v[j] = (j < a.nlc) ? a.lc_x[j][2 * p] : 0.0;
    return lc_value<SEQ, kMaxLc>(a.lc_c, v, a.nlc);
  }
}
// u of the row below / above the launch's rows at column i
template <bool LC, bool SEQ>
__device__ __forceinline__ double ghost_u(const RhsArgs &a, const double *ptr, long long off, long long i) {
  if (!LC || ptr) return ptr[2 * i];
  return stateu<true, SEQ>(a, off + i);
}

}  // namespace

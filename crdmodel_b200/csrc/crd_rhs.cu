// crd_rhs.cu — the fused stencil + reaction right-hand side, its grid descriptor and the halo ring.
//
// Reference being replaced: f() of the four programs
//   src/FHNmodel_torus.cpp:504-667, src/GoldbeterModel_torus.cpp:547-724,
//   src/FHNmodel_flat.cpp:469-616,  src/GoldbeterModel_flat.cpp:515-689
// which make three sweeps over memory per evaluation (N_VConst, stencil, reaction) and re-evaluate
// sin/cos per point.  Here: ONE pass, 16 B read + 16 B written per grid point, metric coefficients
// from a per-theta table computed once on the host with the reference's own expressions.
//
// Kernel shape (HBM-bound, no tensor cores: nothing here is a contraction): every thread owns one
// theta column of RY consecutive phi rows.  It issues all its loads first (RY 16-byte centre loads,
// the east/west u values, one u above and one below), then computes RY points and stores RY 16-byte
// results, so each state row is fetched from HBM once and the row above/below comes from L2/L1.
#include <cmath>
#include <vector>

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
      du = __dadd_rn(du, __dsub_rn(__dadd_rn(__dadd_rn(__dsub_rn(b, v2), v3), __dmul_rn(G_kf, Y)), __dmul_rn(G_k, Z)));
      dv = __dsub_rn(__dsub_rn(v2, v3), __dmul_rn(G_kf, Y));   // never -0 (v2 - v3 is +0 when it vanishes), so 0.0 + dv == dv
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

// ---- state access: plain vector, or sum_j c_j x_j formed on the fly (same operation order as lincomb_kernel) ---
template <bool LC>
__device__ __forceinline__ double2 state2(const RhsArgs &a, long long p) {
  if constexpr (!LC) {
    return reinterpret_cast<const double2 *>(a.y)[p];
  } else {
    double2 v[kMaxLc];
#pragma unroll
    for (int j = 0; j < kMaxLc; ++j)
      v[j] = (j < a.nlc) ? reinterpret_cast<const double2 *>(a.lc_x[j])[p] : make_double2(0.0, 0.0);
    double2 s = make_double2(a.lc_c[0] * v[0].x, a.lc_c[0] * v[0].y);
#pragma unroll
    for (int j = 1; j < kMaxLc; ++j)
      if (j < a.nlc) { s.x = fma(a.lc_c[j], v[j].x, s.x); s.y = fma(a.lc_c[j], v[j].y, s.y); }
    return s;
  }
}
template <bool LC>
__device__ __forceinline__ double stateu(const RhsArgs &a, long long p) {
  if constexpr (!LC) {
    return a.y[2 * p];
  } else {
    double v[kMaxLc];
#pragma unroll
    for (int j = 0; j < kMaxLc; ++j) v[j] = (j < a.nlc) ? a.lc_x[j][2 * p] : 0.0;
    double s = a.lc_c[0] * v[0];
#pragma unroll
    for (int j = 1; j < kMaxLc; ++j)
      if (j < a.nlc) s = fma(a.lc_c[j], v[j], s);
    return s;
  }
}
// u of the row below / above the launch's rows at column i
template <bool LC>
__device__ __forceinline__ double ghost_u(const RhsArgs &a, const double *ptr, long long off, long long i) {
  if (!LC || ptr) return ptr[2 * i];
  return stateu<true>(a, off + i);
}

// ---- the fused kernel ---------------------------------------------------------------------------------
// work item = (row group jg, column i); rows j0 = jg*RY .. j0+RY-1 of the slab described by `a`.
template <int MODEL, bool EXACT, int RY, int MINB, bool LC>
__global__ void __launch_bounds__(256, MINB) rhs_kernel(const RhsArgs a) {
  const long long nx = a.nx, nyl = a.nyl;
  const long long w = blockIdx.x * 256LL + threadIdx.x;
  const long long ngroups = (nyl + RY - 1) / RY;
  if (w >= nx * ngroups) return;
  // w / nx: multiply-shift when the work count fits 31 bits (host-computed magic), else 64-bit division
  const long long jg = a.div_shift >= 0 ? (long long)((__umulhi(a.div_magic, (unsigned)w) + (unsigned)w) >> a.div_shift) : w / nx;
  const long long i = w - jg * nx;
  const long long j0 = jg * RY;
  const long long iw = (i == 0) ? nx - 1 : i - 1;
  const long long ie = (i == nx - 1) ? 0 : i + 1;
  const int nrows = (nyl - j0 < RY) ? (int)(nyl - j0) : RY;

  double2 c[RY];
  double uw[RY], ue[RY];
  double uu[RY + 2];  // u of rows j0-1 .. j0+RY
#pragma unroll
  for (int r = 0; r < RY; ++r) {
    if (r < nrows) {
      const long long row = (j0 + r) * nx;
      c[r] = state2<LC>(a, row + i);
      uw[r] = stateu<LC>(a, row + iw);
      ue[r] = stateu<LC>(a, row + ie);
    } else {
      c[r] = make_double2(0.0, 0.0); uw[r] = 0.0; ue[r] = 0.0;
    }
  }
  uu[0] = (j0 == 0) ? ghost_u<LC>(a, a.south, a.south_off, i) : stateu<LC>(a, (j0 - 1) * nx + i);
  {
    const long long jn = j0 + nrows;  // row above the last one this thread computes
    uu[RY + 1] = (jn == nyl) ? ghost_u<LC>(a, a.north, a.north_off, i) : stateu<LC>(a, jn * nx + i);
  }
#pragma unroll
  for (int r = 0; r < RY; ++r) uu[r + 1] = c[r].x;

  double t1 = 0.0, t3 = 0.0;
  if (is_torus(MODEL)) {
    const double2 tc = reinterpret_cast<const double2 *>(a.cth)[i];
    t1 = tc.x; t3 = tc.y;
  }
  double2 *__restrict__ out = reinterpret_cast<double2 *>(a.ydot);
#pragma unroll
  for (int r = 0; r < RY; ++r) {
    if (r < nrows) {
      const long long jl = j0 + r;
      const double uN = (r + 1 == nrows) ? uu[RY + 1] : uu[r + 2];
      const double uS = uu[r];
      double du = EXACT ? stencil_exact<MODEL>(a.k, t1, t3, c[r].x, uw[r], ue[r], uS, uN)
                        : stencil_fast<MODEL>(a.k, t1, t3, c[r].x, uw[r], ue[r], uS, uN);
      double dv = 0.0;
      if (a.react) {
        const bool frozen = (a.freeze_north && jl == nyl - 1) || (a.freeze_south && jl == 0);
        if (frozen) { du = 0.0; dv = 0.0; }
        else react<MODEL, EXACT>(a.k, a.brow[jl], c[r].x, c[r].y, du, dv);
      }
      out[jl * nx + i] = make_double2(du, dv);
    }
  }
}

// Out-of-line recomputation of one thread's tile column with IEEE divisions (numerators in the subnormal /
// huge / non-finite range somewhere in the column).  col points at the centre of the thread's first row.
template <int MODEL>
__device__ __noinline__ void redo_column_ieee(double Diff, double inv_rr, double twodx, double dxdx, double dydy, double k2n,
                                              double krm, double kap, int react_on, const double2 *col, int pitch, int nrows,
                                              double a1, double a3, const double *brow, double2 *out, long long nx) {
  // scalars by value: taking the address of the kernel parameter block would force every thread to spill it
  RhsConst k;
  k.Diff = Diff; k.inv_rr = inv_rr; k.twodx = twodx; k.dxdx = dxdx; k.dydy = dydy; k.k2n = k2n; k.krm = krm; k.kap = kap;
  for (int r = 0; r < nrows; ++r) {
    const double2 cc = col[r * pitch];
    const double uW = col[r * pitch - 1].x, uE = col[r * pitch + 1].x, uS = col[(r - 1) * pitch].x, uN = col[(r + 1) * pitch].x;
    const double two_uC = __dmul_rn(2.0, cc.x);
    const double n1 = __dmul_rn(Diff, __dmul_rn(a1, __dsub_rn(uE, uW)));
    const double n2 = __dmul_rn(Diff, __dmul_rn(inv_rr, __dadd_rn(__dsub_rn(uE, two_uC), uW)));
    const double n3 = __dmul_rn(Diff, __dmul_rn(a3, __dadd_rn(__dsub_rn(uN, two_uC), uS)));
    double du = __dadd_rn(__dadd_rn(__ddiv_rn(n1, twodx), __ddiv_rn(n2, dxdx)), __ddiv_rn(n3, dydy));
    double dv = 0.0;
    if (react_on) react<MODEL, true>(k, brow[r], cc.x, cc.y, du, dv);
    out[r * nx] = make_double2(du, dv);
  }
}

// ---- the tiled kernel: row segments staged in shared memory by 1-D TMA bulk copies ---------------------
// One CTA = one tile of TX theta columns x TY phi rows.  An elected thread arms an mbarrier with the tile's
// byte count and issues one cp.async.bulk (UBLKCP) per tile row: (TX+2) points of rows j0-1 .. j0+TY, the
// wrap columns and the ghost rows as separate small copies.  No thread computes a global load address for
// the state; the stencil reads its neighbours from shared memory at compile-time offsets, the column's
// previous/next row stay in registers while it marches, results leave as coalesced 16-byte stores.
// Resident CTAs of the same SM overlap each other's load / compute / store phases.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

template <int MODEL, bool EXACT, int TX, int TY, int MINB, bool ACC, bool LC, int NG>
__global__ void __launch_bounds__(256, MINB) rhs_tile_kernel(const RhsArgs a) {
  constexpr int PITCH = TX + 2;            // points per staged row (west halo + TX + east halo)
  constexpr int RPT = TY * TX / 256;       // rows marched by one thread
  // NG: row groups the tile arrives in (one mbarrier each)
  constexpr int GR = (TY + 2 + NG - 1) / NG;
  static_assert(256 % TX == 0 && (TY * TX) % 256 == 0, "tile shape");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2 *tile = reinterpret_cast<double2 *>(smem_raw);                      // [TY+2][PITCH]
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)(TY + 2) * PITCH * 16);

  const long long nx = a.nx, nyl = a.nyl;
  const long long tiles_x = (nx + TX - 1) / TX;
  const long long ty = blockIdx.x / tiles_x;
  const long long tx = blockIdx.x - ty * tiles_x;
  const long long i0 = tx * TX, j0 = ty * TY;
  const int w = (nx - i0 < TX) ? (int)(nx - i0) : TX;            // valid columns of this tile
  const int h = (nyl - j0 < TY) ? (int)(nyl - j0) : TY;          // valid rows of this tile
  const unsigned bar = smem_u32(mbar);

  if (!LC) {
  if (threadIdx.x == 0) {
    // the staged rows arrive in NG groups, each on its own mbarrier, so the march starts when the first
    // group has landed instead of waiting for the whole tile
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8u * gi) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const bool west_in = i0 > 0, east_in = i0 + w < nx;           // halo column contiguous with the tile?
    const unsigned row_bytes = (unsigned)(w + (west_in ? 1 : 0) + (east_in ? 1 : 0)) * 16u;
    const double2 *y2 = reinterpret_cast<const double2 *>(a.y);
    for (int gi = 0; gi < NG; ++gi) {
      const int ra = gi * GR, rb = (ra + GR < h + 2) ? ra + GR : h + 2;
      const unsigned gbytes = (rb > ra) ? (unsigned)(rb - ra) * (unsigned)(w + 2) * 16u : 0u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar + 8u * gi), "r"(gbytes) : "memory");
      for (int r = ra; r < rb; ++r) {
        const long long jr = j0 - 1 + r;
        const double2 *row = (jr < 0) ? reinterpret_cast<const double2 *>(a.south)
                           : (jr >= nyl) ? reinterpret_cast<const double2 *>(a.north) : y2 + jr * nx;
        const unsigned dst = smem_u32(tile + r * PITCH);
        bulk_g2s(dst + (west_in ? 0u : 16u), row + i0 - (west_in ? 1 : 0), row_bytes, bar + 8u * gi);
        if (!west_in) bulk_g2s(dst, row + (nx - 1), 16u, bar + 8u * gi);                    // theta wrap: column nx-1
        if (!east_in) bulk_g2s(dst + (unsigned)(w + 1) * 16u, row, 16u, bar + 8u * gi);     // theta wrap: column 0
      }
    }
  }
  __syncthreads();   // barrier initialised before anyone polls it
  } else {
    // fused stage assembly: the tile is the combination sum_j c_j x_j, formed while it is staged (the stage
    // state is never written to HBM).  Plain coalesced 16-byte loads; each thread stages its own column, nine
    // rows per pass (9 loads per input vector in flight per thread), then the few halo-column entries.
    auto generic = [&](int r, int sc) {
      const long long jr = j0 - 1 + r;
      const long long col = (sc == 0) ? (i0 == 0 ? nx - 1 : i0 - 1) : (sc == w + 1) ? (i0 + w == nx ? 0 : i0 + w) : i0 + sc - 1;
      double2 v;
      if (jr < 0) v = a.south ? reinterpret_cast<const double2 *>(a.south)[col] : state2<true>(a, a.south_off + col);
      else if (jr >= nyl) v = a.north ? reinterpret_cast<const double2 *>(a.north)[col] : state2<true>(a, a.north_off + col);
      else v = state2<true>(a, jr * nx + col);
      tile[r * PITCH + sc] = v;
    };
    const int r_lo = (j0 == 0) ? 1 : 0, r_hi = (j0 + h == nyl) ? h + 1 : h + 2;   // tile rows that are rows of this launch
    constexpr int R = 9;   // (TY + 2) = 18 rows in two passes, 9 loads per input vector in flight per thread
    for (int tcol = threadIdx.x; tcol < w; tcol += 256) {
      const long long p0 = (j0 - 1) * nx + i0 + tcol;   // point offset of tile row 0 in this column
      for (int rb = r_lo; rb < r_hi; rb += R) {
        double2 acc[R], v[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr)
          v[rr] = (rb + rr < r_hi) ? reinterpret_cast<const double2 *>(a.lc_x[0])[p0 + (rb + rr) * nx] : make_double2(0.0, 0.0);
#pragma unroll
        for (int rr = 0; rr < R; ++rr) acc[rr] = make_double2(a.lc_c[0] * v[rr].x, a.lc_c[0] * v[rr].y);
#pragma unroll
        for (int j = 1; j < kMaxLc; ++j) {
          if (j < a.nlc) {
#pragma unroll
            for (int rr = 0; rr < R; ++rr)
              v[rr] = (rb + rr < r_hi) ? reinterpret_cast<const double2 *>(a.lc_x[j])[p0 + (rb + rr) * nx] : make_double2(0.0, 0.0);
#pragma unroll
            for (int rr = 0; rr < R; ++rr) { acc[rr].x = fma(a.lc_c[j], v[rr].x, acc[rr].x); acc[rr].y = fma(a.lc_c[j], v[rr].y, acc[rr].y); }
          }
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr)
          if (rb + rr < r_hi) tile[(rb + rr) * PITCH + tcol + 1] = acc[rr];
      }
      if (r_lo == 1) generic(0, tcol + 1);          // ghost row below the slab
      if (r_hi == h + 1) generic(h + 1, tcol + 1);  // ghost row above the slab
    }
    for (int e = threadIdx.x; e < 2 * (h + 2); e += 256) generic(e >> 1, (e & 1) ? w + 1 : 0);   // halo columns
    __syncthreads();
  }

  const int c = (TX == 256) ? threadIdx.x : threadIdx.x % TX;   // column inside the tile
  const int g0 = (TX == 256) ? 0 : (threadIdx.x / TX) * RPT; // first tile row of this thread (256 threads per CTA)
  double t1 = 0.0, t3 = 0.0;
  const bool active = c < w && g0 < h;
  if (is_torus(MODEL) && active) {
    const double2 tc = reinterpret_cast<const double2 *>(a.cth)[i0 + c];
    t1 = tc.x; t3 = tc.y;
  }
  auto wait_group = [&](int gi) {   // phase 0 of group gi's barrier
    unsigned ok = 0;
    while (!ok) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(bar + 8u * gi) : "memory");
    }
  };
  // rows g0, g0+1, g0+2 of the tile are needed before the first output row
  if (!LC) {
    for (int gi = 0; gi <= (g0 + 2) / GR; ++gi) wait_group(gi);
  }
  if (!active) return;

  const double2 *col = tile + (g0 + 1) * PITCH + (c + 1);   // centre of this thread's first row
  double uS = col[-PITCH].x;
  double2 cc = col[0];
  double2 *out = reinterpret_cast<double2 *>(a.ydot) + (j0 + g0) * nx + (i0 + c);
  const int nrows = (h - g0 < RPT) ? (h - g0) : RPT;
  const double *__restrict__ brow = a.brow + (j0 + g0);
  const int react_on = a.react;
  double2 *const out0 = out;
  bool bad = (ACC && EXACT && is_torus(MODEL)) ? (a.k.div_safe == 0) : false;
  auto row = [&](int r) {
    if (!LC && r > 0 && (g0 + r + 2) % GR == 0 && (g0 + r + 2) / GR < NG) wait_group((g0 + r + 2) / GR);   // north row enters a new group
    const double2 nn = col[(r + 1) * PITCH];
    const double uW = col[r * PITCH - 1].x, uE = col[r * PITCH + 1].x;
    double du = !EXACT ? stencil_fast<MODEL>(a.k, t1, t3, cc.x, uW, uE, uS, nn.x)
                : ACC  ? stencil_exact_acc<MODEL>(a.k, t1, t3, cc.x, uW, uE, uS, nn.x, bad)
                       : stencil_exact<MODEL>(a.k, t1, t3, cc.x, uW, uE, uS, nn.x);
    double dv = 0.0;
    if (react_on) react<MODEL, EXACT>(a.k, __ldg(brow + r), cc.x, cc.y, du, dv);
    *out = make_double2(du, dv);
    out += nx;
    uS = cc.x;
    cc = nn;
  };
  if (nrows == RPT) {   // full tile: straight-line code, rows interleave
#pragma unroll
    for (int r = 0; r < RPT; ++r) row(r);
  } else {
    for (int r = 0; r < nrows; ++r) row(r);
  }
  // rare fix-ups, after the marched rows (same thread, same addresses: program order)
  if (ACC && EXACT && is_torus(MODEL) && bad)
    redo_column_ieee<MODEL>(a.k.Diff, a.k.inv_rr, a.k.twodx, a.k.dxdx, a.k.dydy, a.k.k2n, a.k.krm, a.k.kap, react_on,
                            tile + (g0 + 1) * PITCH + (c + 1), PITCH, nrows, t1, t3, brow, out0, nx);
  if (react_on) {
    // frozen rows while t < tBoundary (:643-653): only slab row 0 / nyl-1 can be one
    if (a.freeze_south && j0 + g0 == 0) out0[0] = make_double2(0.0, 0.0);
    const long long rn = nyl - 1 - (j0 + g0);
    if (a.freeze_north && rn >= 0 && rn < nrows) out0[rn * nx] = make_double2(0.0, 0.0);
  }
}

template <int MODEL, bool EXACT, int TX, int TY, int MINB, bool ACC, bool LC, int NG>
int launch_tile_lc(crd_grid *g, const RhsArgs &a, cudaStream_t st) {
  const long long tiles = ((a.nx + TX - 1) / TX) * ((a.nyl + TY - 1) / TY);
  if (tiles <= 0) return 0;
  if (tiles > 2147483647LL) { set_error("slab too large for one launch"); return -1; }
  const size_t smem = (size_t)(TY + 2) * (TX + 2) * 16 + 8 * NG + 8;
  auto kern = rhs_tile_kernel<MODEL, EXACT, TX, TY, MINB, ACC, LC, NG>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set = true;
  }
  kern<<<(unsigned)tiles, 256, smem, st>>>(a);
  return check_launch(g->ctx, "rhs_tile_kernel");
}
template <int MODEL, bool EXACT, int TX, int TY, int MINB, bool ACC, int NG = 3>
int launch_tile(crd_grid *g, const RhsArgs &a, cudaStream_t st) {
  return a.nlc > 0 ? launch_tile_lc<MODEL, EXACT, TX, TY, MINB, ACC, true, 1>(g, a, st)
                   : launch_tile_lc<MODEL, EXACT, TX, TY, MINB, ACC, false, NG>(g, a, st);
}

// ---- the streaming kernel: persistent CTAs, rows flow through a shared-memory ring ------------------------
// 2 CTAs per SM stay resident and walk over (strip of 256 columns) x (segment of rows) units.  Warp 8 is the
// producer: one lane issues a 1-D TMA bulk copy per row and input vector into the next free ring slot and arms
// that slot's "full" mbarrier with the byte count.  Warps 0..7 are consumers: a thread owns one column, waits for
// the slot, reads (and, for a fused stage, combines sum_j c_j x_j of) its point and the two neighbours' u, releases
// the slot on the "empty" mbarrier, and with the previous two rows still in registers computes the row before.
// Every state row is fetched once (plus one halo row per segment end); no CTA start-up per tile.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}

template <int MODEL, bool EXACT, int NV, bool PLAIN, int RB>
__global__ void __launch_bounds__(288, 2) rhs_stream_kernel(const RhsArgs a, int seg_rows, int S) {
  constexpr int TX = 256, PITCH = TX + 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2 *ring = reinterpret_cast<double2 *>(smem_raw);                       // [S][RB][NV][PITCH]
  const unsigned bars = smem_u32(smem_raw + (size_t)S * RB * NV * PITCH * 16); // full[S], empty[S]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nx = a.nx, nyl = a.nyl;
  const long long strips = (nx + TX - 1) / TX, segs = (nyl + seg_rows - 1) / seg_rows, units = strips * segs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8u * s) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(bars + 8u * (S + s)) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 8) {   // ---------------- producer ----------------
    if (lane != 0) return;
    long long it = 0;   // stage counter
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const long long strip = u % strips, seg = u / strips;
      const long long i0 = strip * TX, jA = seg * seg_rows, jB = (jA + seg_rows < nyl) ? jA + seg_rows : nyl;
      const int w = (nx - i0 < TX) ? (int)(nx - i0) : TX;
      const bool west_in = i0 > 0, east_in = i0 + w < nx;
      const unsigned row_bytes = (unsigned)(w + (west_in ? 1 : 0) + (east_in ? 1 : 0)) * 16u;
      for (long long j0 = jA - 1; j0 <= jB; j0 += RB, ++it) {   // a stage = rows j0 .. j0+RB-1 (clipped at jB)
        const int s = (int)(it % S);
        mbar_wait(bars + 8u * (S + s), (unsigned)((it / S) & 1) ^ 1u);
        const int nr = (jB - j0 + 1 < RB) ? (int)(jB - j0 + 1) : RB;
        unsigned bytes = 0;
        for (int rr = 0; rr < nr; ++rr) {
          const long long jr = j0 + rr;
          const bool ext = (jr < 0 && a.south) || (jr >= nyl && a.north);
          bytes += (unsigned)((PLAIN || ext) ? 1 : NV) * (unsigned)(w + 2) * 16u;
        }
        const unsigned full = bars + 8u * s;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(bytes) : "memory");
        for (int rr = 0; rr < nr; ++rr) {
          const long long jr = j0 + rr;
          const bool ext = (jr < 0 && a.south) || (jr >= nyl && a.north);   // an already combined ghost row
          const int nvec = (PLAIN || ext) ? 1 : NV;
          for (int v = 0; v < nvec; ++v) {
            const double2 *row;
            if (ext) row = reinterpret_cast<const double2 *>(jr < 0 ? a.south : a.north);
            else if (PLAIN) row = reinterpret_cast<const double2 *>(a.y) + jr * nx;
            else {
              const long long off = (jr < 0) ? a.south_off : (jr >= nyl) ? a.north_off : jr * nx;
              row = reinterpret_cast<const double2 *>(a.lc_x[v]) + off;
            }
            const unsigned dst = smem_u32(ring + (((size_t)s * RB + rr) * NV + v) * PITCH);
            bulk_g2s(dst + (west_in ? 0u : 16u), row + i0 - (west_in ? 1 : 0), row_bytes, full);
            if (!west_in) bulk_g2s(dst, row + (nx - 1), 16u, full);
            if (!east_in) bulk_g2s(dst + (unsigned)(w + 1) * 16u, row, 16u, full);
          }
        }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  const int c = threadIdx.x;   // 0..255: column inside the strip
  const int react_on = a.react;
  long long it = 0;
  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    const long long strip = u % strips, seg = u / strips;
    const long long i0 = strip * TX, jA = seg * seg_rows, jB = (jA + seg_rows < nyl) ? jA + seg_rows : nyl;
    const int w = (nx - i0 < TX) ? (int)(nx - i0) : TX;
    const bool active = c < w;
    double t1 = 0.0, t3 = 0.0;
    if (is_torus(MODEL) && active) {
      const double2 tc = reinterpret_cast<const double2 *>(a.cth)[i0 + c];
      t1 = tc.x; t3 = tc.y;
    }
    double2 *out = reinterpret_cast<double2 *>(a.ydot) + jA * nx + (i0 + c);
    double uS = 0.0, cW = 0.0, cE = 0.0;
    double2 cc = make_double2(0.0, 0.0);
    // one row: fetch (and combine) the arriving row jr, then emit row jr-1 from the three rows in registers
    auto step = [&](const double2 *slot, long long jr) {
      const bool ext = (jr < 0 && a.south) || (jr >= nyl && a.north);
      double2 nn;
      double nW, nE;
      if (PLAIN || ext) {
        nn = slot[c + 1]; nW = slot[c].x; nE = slot[c + 2].x;
      } else {
        double2 v[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = slot[j * PITCH + c + 1];
        nn = make_double2(a.lc_c[0] * v[0].x, a.lc_c[0] * v[0].y);
#pragma unroll
        for (int j = 1; j < NV; ++j) { nn.x = fma(a.lc_c[j], v[j].x, nn.x); nn.y = fma(a.lc_c[j], v[j].y, nn.y); }
        nW = __shfl_up_sync(0xffffffffu, nn.x, 1);
        nE = __shfl_down_sync(0xffffffffu, nn.x, 1);
        if (lane == 0 || lane == 31) {   // the neighbour lives in another warp: combine its u from the slot
          const int q = (lane == 0) ? c : c + 2;
          double e = a.lc_c[0] * slot[q].x;
#pragma unroll
          for (int j = 1; j < NV; ++j) e = fma(a.lc_c[j], slot[j * PITCH + q].x, e);
          if (lane == 0) nW = e; else nE = e;
        }
      }
      if (jr > jA) {   // rows jr-2 (uS), jr-1 (cc) and jr (nn) are here: output row jr-1
        const long long jl = jr - 1;
        double du = EXACT ? stencil_exact<MODEL>(a.k, t1, t3, cc.x, cW, cE, uS, nn.x)
                          : stencil_fast<MODEL>(a.k, t1, t3, cc.x, cW, cE, uS, nn.x);
        double dv = 0.0;
        if (react_on) {
          react<MODEL, EXACT>(a.k, __ldg(a.brow + jl), cc.x, cc.y, du, dv);
          const bool frozen = (a.freeze_north && jl == nyl - 1) || (a.freeze_south && jl == 0);
          du = frozen ? 0.0 : du;
          dv = frozen ? 0.0 : dv;
        }
        if (active) *out = make_double2(du, dv);
        out += nx;
      }
      uS = cc.x; cc = nn; cW = nW; cE = nE;
    };
    for (long long j0 = jA - 1; j0 <= jB; j0 += RB, ++it) {
      const int s = (int)(it % S);
      mbar_wait(bars + 8u * s, (unsigned)((it / S) & 1));
      const double2 *stage = ring + (size_t)s * RB * NV * PITCH;
      const int nr = (jB - j0 + 1 < RB) ? (int)(jB - j0 + 1) : RB;
      if (nr == RB) {
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) step(stage + (size_t)rr * NV * PITCH, j0 + rr);
      } else {
        for (int rr = 0; rr < nr; ++rr) step(stage + (size_t)rr * NV * PITCH, j0 + rr);
      }
      __syncwarp();   // every lane has read the stage (the values it still needs are in registers)
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8u * (S + s)) : "memory");
    }
  }
}

template <int MODEL, bool EXACT, int NV, bool PLAIN>
int launch_stream_nv(crd_grid *g, const RhsArgs &a, cudaStream_t st) {
  constexpr int RB = (NV == 1) ? 4 : (NV <= 3 ? 2 : 1);   // rows per ring stage
  const int seg_rows = 128;
  const long long strips = (a.nx + 255) / 256, segs = (a.nyl + seg_rows - 1) / seg_rows, units = strips * segs;
  if (units <= 0) return 0;
  const size_t stage_bytes = (size_t)RB * NV * 258 * 16;
  int S = (int)(100000 / stage_bytes);
  if (S > 8) S = 8;
  if (S < 3) S = 3;
  const size_t smem = (size_t)S * stage_bytes + (size_t)2 * S * 8;
  auto kern = rhs_stream_kernel<MODEL, EXACT, NV, PLAIN, RB>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110000);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set = true;
  }
  const long long ctas = units < 2LL * kSMs ? units : 2LL * kSMs;
  kern<<<(unsigned)ctas, 288, smem, st>>>(a, seg_rows, S);
  return check_launch(g->ctx, "rhs_stream_kernel");
}

template <int MODEL, bool EXACT>
int launch_stream(crd_grid *g, const RhsArgs &a, cudaStream_t st) {
  switch (a.nlc) {
    case 0: return launch_stream_nv<MODEL, EXACT, 1, true>(g, a, st);
    case 2: return launch_stream_nv<MODEL, EXACT, 2, false>(g, a, st);
    case 3: return launch_stream_nv<MODEL, EXACT, 3, false>(g, a, st);
    case 5: return launch_stream_nv<MODEL, EXACT, 5, false>(g, a, st);
    default: return 1;   // other counts: caller falls back to the tiled kernel
  }
}

template <int MODEL, bool EXACT>
int launch_model(crd_grid *g, const RhsArgs &a_in, cudaStream_t st) {
  // variant 0 = automatic.  Large slabs (>= 4 Mi points, HBM-bound): the TMA-tiled kernel wherever a tile row is
  // reasonably full.  Small slabs (the reference's default 400x1600 / 100x400 grids live in L2 and are bound by
  // launch latency and by how many CTAs a partial wave gets): the direct kernel with 2 rows per thread.
  // explicit: direct kernel (rows per thread, min CTAs/SM) 1 (2,4) | 2 (8,2) | 3 (1,4) | 4 (4,3) | 5 (4,4)
  //           tiled kernel (TX, TY, min CTAs/SM) 10 (128,16,4) | 11 (128,32,3) | 12 (64,32,4) | 13 (256,16,3) | 14 (128,16,3)
  //           15 = 13 with flag-and-redo instead of a branch per point
  int variant = g->variant;
  if (variant == 0) {
    const long long pts = a_in.nx * a_in.nyl;
    if (pts < (4LL << 20)) variant = 1;
    else variant = (a_in.nx >= 192) ? ((EXACT && !is_fhn(MODEL)) ? 15 : 13) : (a_in.nx >= 96) ? 10 : 5;   // measured: profiles/README.md
  }
  // measured (profiles/README.md): the streaming kernel wins only for the widest fused stage (5 input vectors,
  // where its once-per-row fetch beats the tiled kernel's register-staged tiles); the tiled kernel everywhere else
  if (g->variant == 0 && a_in.nlc == 5 && a_in.nx >= 192 && a_in.nx * a_in.nyl >= (4LL << 20)) variant = 20;
  if (variant == 20) {   // streaming kernel (persistent CTAs, shared-memory row ring)
    const int r = launch_stream<MODEL, EXACT>(g, a_in, st);
    if (r <= 0) return r;
    variant = 13;
  }
  switch (variant) {
    case 10: return launch_tile<MODEL, EXACT, 128, 16, 4, false>(g, a_in, st);
    case 11: return launch_tile<MODEL, EXACT, 128, 32, 3, false>(g, a_in, st);
    case 12: return launch_tile<MODEL, EXACT, 64, 32, 4, false>(g, a_in, st);
    case 13: return launch_tile<MODEL, EXACT, 256, 16, 3, false>(g, a_in, st);
    case 14: return launch_tile<MODEL, EXACT, 128, 16, 3, false>(g, a_in, st);
    case 15: return launch_tile<MODEL, EXACT, 256, 16, 3, true>(g, a_in, st);   // flag-and-redo instead of a branch per point
    case 16: return launch_tile<MODEL, EXACT, 256, 16, 3, false, 6>(g, a_in, st);  // tile arrives in 6 row groups
    case 17: return launch_tile<MODEL, EXACT, 256, 16, 3, false, 2>(g, a_in, st);  // ... in 2
    case 18: return launch_tile<MODEL, EXACT, 256, 16, 3, false, 9>(g, a_in, st);  // ... in 9
    default: break;
  }
  const int RY = (variant == 1) ? 2 : (variant == 2) ? 8 : (variant == 3) ? 1 : 4;
  const long long ngroups = (a_in.nyl + RY - 1) / RY;
  const long long work = a_in.nx * ngroups;
  const long long blocks = (work + 255) / 256;
  if (blocks <= 0) return 0;
  RhsArgs a = a_in;
  a.div_shift = -1; a.div_magic = 0;
  if (work < (1LL << 31)) {  // Granlund-Montgomery: n / d = (mulhi(m, n) + n) >> l for n < 2^31
    int l = 0;
    while ((1LL << l) < a.nx) ++l;
    a.div_shift = l;
    a.div_magic = (unsigned)((((1ULL << l) - (unsigned long long)a.nx) << 32) / (unsigned long long)a.nx + 1ULL);
  }
  if (blocks > 2147483647LL) { set_error("slab too large for one launch"); return -1; }
  const unsigned nb = (unsigned)blocks;
  const bool lc = a.nlc > 0;
#define CRD_DIRECT(RY_, MB_)                                                   \
  do {                                                                         \
    if (lc) rhs_kernel<MODEL, EXACT, RY_, MB_, true><<<nb, 256, 0, st>>>(a);   \
    else rhs_kernel<MODEL, EXACT, RY_, MB_, false><<<nb, 256, 0, st>>>(a);     \
  } while (0)
  switch (variant) {
    case 1: CRD_DIRECT(2, 4); break;
    case 2: CRD_DIRECT(8, 2); break;
    case 3: CRD_DIRECT(1, 4); break;
    case 4: CRD_DIRECT(4, 3); break;
    default: CRD_DIRECT(4, 4); break;
  }
#undef CRD_DIRECT
  return check_launch(g->ctx, "rhs_kernel");
}

int launch_rhs(crd_grid *g, const RhsArgs &a, cudaStream_t st) {
  const bool exact = g->p.arith == CRD_ARITH_EXACT;
  switch (g->p.model) {
    case CRD_FHN_TORUS: return exact ? launch_model<CRD_FHN_TORUS, true>(g, a, st) : launch_model<CRD_FHN_TORUS, false>(g, a, st);
    case CRD_GOLDBETER_TORUS: return exact ? launch_model<CRD_GOLDBETER_TORUS, true>(g, a, st) : launch_model<CRD_GOLDBETER_TORUS, false>(g, a, st);
    case CRD_FHN_FLAT: return exact ? launch_model<CRD_FHN_FLAT, true>(g, a, st) : launch_model<CRD_FHN_FLAT, false>(g, a, st);
    case CRD_GOLDBETER_FLAT: return exact ? launch_model<CRD_GOLDBETER_FLAT, true>(g, a, st) : launch_model<CRD_GOLDBETER_FLAT, false>(g, a, st);
  }
  set_error("unknown model %d", g->p.model);
  return -1;
}

// Arguments for rows [r0, r1) of the slab.  The rows just outside that range are given either as an external
// pointer (ghost row) or as a row index of the slab itself.
struct RowRef { const double *ptr; long long row; };
inline RowRef ext_row(const double *p) { return RowRef{p, 0}; }
inline RowRef slab_row(long long r) { return RowRef{nullptr, r}; }

RhsArgs make_args(const crd_grid *g, double t, const StateRef &S, double *ydot, long long r0, long long r1, RowRef south,
                  RowRef north) {
  RhsArgs a;
  const long long nx = g->nx;
  a.nlc = S.n;
  a.south_off = a.north_off = 0;
  if (S.n == 0) {
    a.y = S.y + 2 * r0 * nx;
    a.south = south.ptr ? south.ptr : S.y + 2 * south.row * nx;
    a.north = north.ptr ? north.ptr : S.y + 2 * north.row * nx;
    for (int j = 0; j < kMaxLc; ++j) { a.lc_x[j] = nullptr; a.lc_c[j] = 0.0; }
  } else {
    a.y = nullptr;
    for (int j = 0; j < kMaxLc; ++j) {
      a.lc_x[j] = j < S.n ? S.x[j] + 2 * r0 * nx : nullptr;
      a.lc_c[j] = j < S.n ? S.c[j] : 0.0;
    }
    a.south = south.ptr; a.south_off = (south.row - r0) * nx;
    a.north = north.ptr; a.north_off = (north.row - r0) * nx;
  }
  a.ydot = ydot + 2 * r0 * nx;
  a.cth = g->cth;
  a.brow = g->brow + r0;
  a.nx = nx;
  a.nyl = r1 - r0;
  const bool tb = t < g->p.t_boundary;
  a.freeze_south = (tb && g->js == 0 && r0 == 0) ? 1 : 0;
  a.freeze_north = (tb && g->je == g->ny - 1 && r1 == g->nyl) ? 1 : 0;
  a.react = (is_fhn(g->p.model) || g->p.just_diffusion == 0) ? 1 : 0;
  a.div_shift = -1; a.div_magic = 0;
  a.k = g->k;
  return a;
}

// ---- halo ring: push first/last row into the neighbours' ghost blocks, flag the epoch ------------------
template <bool LC>
__global__ void __launch_bounds__(256) halo_push_kernel(const RhsArgs a, double *prev_north, double *next_south,
                                                        unsigned long long *prev_flag, unsigned long long *next_flag,
                                                        unsigned long long *ticket, unsigned long long epoch) {
  const long long nx = a.nx, nyl = a.nyl;
  const long long stride = (long long)gridDim.x * blockDim.x;
  double2 *pn = reinterpret_cast<double2 *>(prev_north), *ns = reinterpret_cast<double2 *>(next_south);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nx; i += stride) {
    pn[i] = state2<LC>(a, i);                     // my row js is the row above prev's je
    ns[i] = state2<LC>(a, (nyl - 1) * nx + i);    // my row je is the row below next's js
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long done = atomicAdd(ticket, 1ULL) + 1ULL;
    if (done == gridDim.x * epoch) {  // last block of this epoch (ticket is never reset)
      __threadfence_system();
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(prev_flag), "l"(epoch) : "memory");
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(next_flag), "l"(epoch) : "memory");
    }
  }
}

__global__ void halo_wait_kernel(const unsigned long long *flag_south, const unsigned long long *flag_north,
                                 unsigned long long epoch, int *err) {
  const unsigned long long *f = threadIdx.x == 0 ? flag_south : flag_north;
  const long long t0 = clock64();
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
    if (v >= epoch) break;
    if (clock64() - t0 > 6000000000LL) {  // ~3 s: a neighbour never posted; report instead of hanging
      *err = 100 + (int)threadIdx.x;
      __threadfence_system();
      break;
    }
    __nanosleep(200);
  }
}

constexpr int kPushBlocks = 16;

// ---- initial conditions -----------------------------------------------------------------------------------
struct IcArgs {
  int model, vary_beta, wave_inside, ic_type;
  long long nx, nyl, js;
  double dx, dy, xmin, ymin;
  double wave_length, wave_width, wave_xmin, wave_xmax;
  double s0, s1, p0, p1;  // steady state, perturbed state
};

__global__ void __launch_bounds__(256) ic_kernel(const IcArgs a, double2 *__restrict__ y) {
  const long long w = blockIdx.x * 256LL + threadIdx.x;
  if (w >= a.nx * a.nyl) return;
  const long long j = w / a.nx, i = w - j * a.nx;
  const double yy = __dadd_rn(a.ymin, __dmul_rn((double)(a.js + j), a.dy));
  const double xx = __dadd_rn(a.xmin, __dmul_rn((double)i, a.dx));
  double2 v;
  const bool fhn = is_fhn(a.model), torus = is_torus(a.model);
  if (a.vary_beta == 0) {
    bool in;
    if (torus) {
      const bool ybox = yy >= a.wave_length && yy <= __dmul_rn(2.0, a.wave_length);
      if (a.wave_inside == 1) in = xx >= a.wave_xmin && xx <= a.wave_xmax && ybox;
      else in = (xx >= a.wave_xmin || xx <= a.wave_xmax) && ybox;
    } else if (fhn) {
      in = xx >= a.wave_xmin && xx <= a.wave_xmax && yy >= a.wave_length && yy <= __dmul_rn(2.0, a.wave_length);
    } else {
      in = xx >= a.wave_xmin && xx <= a.wave_xmax && yy >= __dmul_rn(2.0, a.wave_length) && yy <= __dmul_rn(3.0, a.wave_length);
    }
    v = in ? make_double2(a.p0, a.p1) : make_double2(a.s0, a.s1);
  } else if (fhn) {
    v = make_double2(1.0, 1.0);
  } else {
    v = make_double2(0.4, 1.6);
    if (a.ic_type == 1) {
      const bool in = xx >= a.wave_xmin && xx <= a.wave_xmax && yy >= __dmul_rn(2.0, a.wave_length) && yy <= __dmul_rn(3.0, a.wave_length);
      if (in) v = make_double2(1.4, 2.6);
    }
  }
  y[w] = v;
}

}  // namespace

// ============================================================================================================
extern "C" {

int crd_decomp_phi(int64_t ny, int nranks, int rank, int64_t *js, int64_t *je) {
  if (nranks < 1 || rank < 0 || rank >= nranks || ny < nranks) { set_error("crd_decomp_phi: bad arguments"); return -1; }
  *js = ny * rank / nranks;
  *je = ny * (rank + 1) / nranks - 1;
  return 0;
}

crd_grid *crd_grid_create(crd_ctx *ctx, const crd_params *p) {
  if (!ctx || !p) { set_error("crd_grid_create: null argument"); return nullptr; }
  if (p->model < 0 || p->model > 3) { set_error("crd_grid_create: unknown model %d", p->model); return nullptr; }
  if (p->nx < 2 || p->ny < 2 || p->js < 0 || p->je < p->js || p->je >= p->ny) {
    set_error("crd_grid_create: bad extents nx=%lld ny=%lld js=%lld je=%lld", (long long)p->nx, (long long)p->ny,
              (long long)p->js, (long long)p->je);
    return nullptr;
  }
  if (use(ctx)) return nullptr;
  crd_grid *g = new crd_grid;
  g->ctx = ctx; g->p = *p;
  g->nx = p->nx; g->ny = p->ny; g->js = p->js; g->je = p->je; g->nyl = p->je - p->js + 1;
  const bool torus = is_torus(p->model);
  // geometry exactly as main() computes it (FHNmodel_torus.cpp:188-189,233-234; FHNmodel_flat.cpp:173-176,229-230)
  if (torus) {
    g->xmin = 0.0; g->xmax = 2.0 * kPI; g->ymin = 0.0; g->ymax = 2.0 * kPI;
    g->r = p->surface_width / (2.0 * kPI);
    g->R = p->surface_length / (2.0 * kPI);
  } else {
    g->xmin = 0.0; g->xmax = p->surface_width - g->xmin; g->ymin = 0.0; g->ymax = p->surface_length - g->ymin;
  }
  g->dx = (g->xmax - g->xmin) / (1.0 * g->nx - 1.0);
  g->dy = (g->ymax - g->ymin) / (1.0 * g->ny - 1.0);
  const double Diff = p->diff, dx = g->dx, dy = g->dy, R = g->R, r = g->r;
  RhsConst &k = g->k;
  k.Diff = Diff;
  k.inv_rr = torus ? (1 / (r * r)) : 0.0;
  k.twodx = 2 * dx; k.dxdx = dx * dx; k.dydy = dy * dy;
  k.r_twodx = 1.0 / k.twodx; k.r_dxdx = 1.0 / k.dxdx; k.r_dydy = 1.0 / k.dydy;
  {
    // the reciprocal-refinement division is proven for positive divisors of moderate magnitude only
    k.div_safe = 1;
    for (double c : {k.twodx, k.dxdx, k.dydy})
      if (!(c > 0x1p-90 && c < 0x1p90)) k.div_safe = 0;
  }
  k.c2 = torus ? Diff * k.inv_rr / k.dxdx : 0.0;
  k.cu1 = Diff / dx / dx; k.cu2 = Diff / dy / dy; k.cu3 = -2.0 * (k.cu1 + k.cu2);
  k.dv_plus0 = 0;
  k.k2n = std::pow(G_K2, G_n); k.krm = std::pow(G_KR, G_m); k.kap = std::pow(G_KA, G_p);

  // per-theta metric table (host libm, the reference's expressions :531-537)
  std::vector<double> cth((size_t)2 * g->nx, 0.0);
  if (torus) {
    for (long long i = 0; i < g->nx; ++i) {
      const double xx = g->xmin + (i) * (dx);
      const double a1 = (-sin(xx) / (r * (R + r * cos(xx))));
      const double a3 = (1 / (((R + r * cos(xx))) * ((R + r * cos(xx)))));
      if (p->arith == CRD_ARITH_EXACT) { cth[2 * i] = a1; cth[2 * i + 1] = a3; }
      else { cth[2 * i] = Diff * a1 / k.twodx; cth[2 * i + 1] = Diff * a3 / k.dydy; }
    }
  }
  // per-phi beta (:623-632); Goldbeter rows carry v0 + v1*b (:715)
  std::vector<double> brow((size_t)g->nyl);
  const bool fhn = is_fhn(p->model);
  for (long long j = 0; j < g->nyl; ++j) {
    const double yy = g->ymin + (g->js + j) * (dy);
    double b = p->beta;
    const bool vary = fhn ? (p->vary_beta != 0) : (p->vary_beta == 1);
    if (vary) b = p->beta_min + yy * (p->beta_max - p->beta_min) / (g->ymax - g->ymin);
    brow[j] = fhn ? b : (G_v0 + G_v1 * b);
    if (fhn && b == 0.0 && std::signbit(b)) k.dv_plus0 = 1;
  }
  cudaError_t e;
  if ((e = cudaMalloc(&g->cth, cth.size() * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(&g->brow, brow.size() * sizeof(double))) != cudaSuccess ||
      (e = cudaMemcpy(g->cth, cth.data(), cth.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(g->brow, brow.data(), brow.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) {
    set_error("crd_grid_create: %s", cudaGetErrorString(e));
    crd_grid_destroy(g);
    return nullptr;
  }
  // ghost block + push ticket
  HaloLayout L{g->nx};
  if ((e = cudaMalloc(&g->halo_local, L.bytes())) != cudaSuccess ||
      (e = cudaMemset(g->halo_local, 0, L.bytes())) != cudaSuccess ||
      (e = cudaMalloc(&g->push_ticket, sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaMemset(g->push_ticket, 0, sizeof(unsigned long long))) != cudaSuccess) {
    set_error("crd_grid_create: %s", cudaGetErrorString(e));
    crd_grid_destroy(g);
    return nullptr;
  }
  return g;
}

void crd_grid_destroy(crd_grid *g) {
  if (!g) return;
  cudaSetDevice(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream);
  if (g->prev_ipc && g->halo_prev) cudaIpcCloseMemHandle(g->halo_prev);
  if (g->next_ipc && g->halo_next && g->halo_next != g->halo_prev) cudaIpcCloseMemHandle(g->halo_next);
  cudaFree(g->cth); cudaFree(g->brow); cudaFree(g->halo_local); cudaFree(g->push_ticket);
  if (g->stage_y) cudaFree(g->stage_y);
  if (g->stage_ydot) cudaFree(g->stage_ydot);
  if (g->s_in) cudaStreamDestroy(g->s_in);
  if (g->s_aux) { cudaStreamSynchronize(g->s_aux); cudaStreamDestroy(g->s_aux); cudaEventDestroy(g->ev_y); cudaEventDestroy(g->ev_b); }
  if (g->s_out) cudaStreamDestroy(g->s_out);
  for (int i = 0; i < g->n_chunks; ++i) { cudaEventDestroy(g->ev_in[i]); cudaEventDestroy(g->ev_k[i]); }
  delete[] g->ev_in; delete[] g->ev_k;
  delete g;
}

int crd_grid_params(const crd_grid *g, crd_params *out) { if (!g || !out) return -1; *out = g->p; return 0; }
int64_t crd_grid_local_length(const crd_grid *g) { return g ? 2 * g->nx * g->nyl : 0; }
int64_t crd_grid_global_length(const crd_grid *g) { return g ? 2 * g->nx * g->ny : 0; }
double crd_grid_dx(const crd_grid *g) { return g ? g->dx : 0.0; }
double crd_grid_dy(const crd_grid *g) { return g ? g->dy : 0.0; }
int64_t crd_grid_rhs_count(const crd_grid *g) { return g ? g->rhs_count : 0; }
int crd_grid_set_variant(crd_grid *g, int variant) { if (!g) return -1; g->variant = variant; return 0; }
int crd_grid_set_overlap(crd_grid *g, int on) { if (!g) return -1; g->overlap = on != 0; return 0; }

int crd_grid_halo_handle(crd_grid *g, unsigned char handle[CRD_HALO_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == CRD_HALO_HANDLE_BYTES, "handle size");
  if (!g) return -1;
  if (use(g->ctx)) return -1;
  cudaIpcMemHandle_t h;
  CRD_CUDA(cudaIpcGetMemHandle(&h, g->halo_local));
  std::memcpy(handle, &h, sizeof h);
  return 0;
}

int crd_grid_halo_connect_ipc(crd_grid *g, const unsigned char prev_handle[CRD_HALO_HANDLE_BYTES],
                              const unsigned char next_handle[CRD_HALO_HANDLE_BYTES]) {
  if (!g) return -1;
  if (use(g->ctx)) return -1;
  cudaIpcMemHandle_t hp, hn;
  std::memcpy(&hp, prev_handle, sizeof hp);
  std::memcpy(&hn, next_handle, sizeof hn);
  void *pp = nullptr, *pn = nullptr;
  CRD_CUDA(cudaIpcOpenMemHandle(&pp, hp, cudaIpcMemLazyEnablePeerAccess));
  if (std::memcmp(&hp, &hn, sizeof hp) == 0) pn = pp;  // two ranks: both neighbours are the same block
  else CRD_CUDA(cudaIpcOpenMemHandle(&pn, hn, cudaIpcMemLazyEnablePeerAccess));
  g->halo_prev = (char *)pp; g->halo_next = (char *)pn;
  g->prev_ipc = g->next_ipc = true;
  g->connected = true;
  return 0;
}

int crd_grid_halo_connect_local(crd_grid *g, crd_grid *prev, crd_grid *next) {
  if (!g || !prev || !next) return -1;
  if (prev->nx != g->nx || next->nx != g->nx) { set_error("halo_connect_local: theta mesh differs"); return -1; }
  if (use(g->ctx)) return -1;
  for (crd_grid *o : {prev, next}) {
    if (o->ctx->device != g->ctx->device) {
      int can = 0;
      CRD_CUDA(cudaDeviceCanAccessPeer(&can, g->ctx->device, o->ctx->device));
      if (!can) { set_error("device %d cannot access device %d", g->ctx->device, o->ctx->device); return -1; }
      cudaError_t e = cudaDeviceEnablePeerAccess(o->ctx->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return -1; }
      cudaGetLastError();
    }
  }
  g->halo_prev = prev->halo_local; g->halo_next = next->halo_local;
  g->prev_ipc = g->next_ipc = false;
  g->connected = true;
  return 0;
}

static int launch_push(crd_grid *g, const StateRef &S, cudaStream_t st) {
  g->epoch++;
  HaloLayout L{g->nx};
  const int par = (int)(g->epoch & 1ULL);
  RhsArgs a = make_args(g, 0.0, S, nullptr, 0, g->nyl, slab_row(0), slab_row(0));
  double *pn = (double *)(g->halo_prev + L.ghost_off(par, 1)), *ns = (double *)(g->halo_next + L.ghost_off(par, 0));
  unsigned long long *pf = (unsigned long long *)(g->halo_prev + L.flag_off(1)), *nf = (unsigned long long *)(g->halo_next + L.flag_off(0));
  if (S.n > 0) halo_push_kernel<true><<<kPushBlocks, 256, 0, st>>>(a, pn, ns, pf, nf, g->push_ticket, g->epoch);
  else halo_push_kernel<false><<<kPushBlocks, 256, 0, st>>>(a, pn, ns, pf, nf, g->push_ticket, g->epoch);
  return check_launch(g->ctx, "halo_push_kernel");
}

static int launch_wait(crd_grid *g, cudaStream_t st) {
  HaloLayout L{g->nx};
  halo_wait_kernel<<<1, 2, 0, st>>>((const unsigned long long *)(g->halo_local + L.flag_off(0)),
                                    (const unsigned long long *)(g->halo_local + L.flag_off(1)), g->epoch, g->ctx->err_dev);
  return check_launch(g->ctx, "halo_wait_kernel");
}

// Ring evaluation, overlapped: the boundary rows travel on the grid's auxiliary stream while the interior rows
// (which need no neighbour data) are computed on the main stream.
//   main: --ev_y--> [ interior rows B .. nyl-B ] ------------------------- wait ev_b --> (caller's next work)
//   aux : wait ev_y, [push rows 0 / nyl-1 to the neighbours], [wait for theirs], [rows 0..B), [rows nyl-B..nyl), ev_b
// B = kEdgeRows (a whole number of tile rows).  Slabs too thin to split run everything on the main stream.
constexpr long long kEdgeRows = 32;

static int ensure_aux(crd_grid *g) {
  if (g->s_aux) return 0;
  // highest priority: the few CTAs of push / wait / edge bands must be scheduled as soon as slots free up,
  // not after the 65k interior CTAs of the main stream have all been dispatched
  int prio_lo = 0, prio_hi = 0;
  CRD_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CRD_CUDA(cudaStreamCreateWithPriority(&g->s_aux, cudaStreamNonBlocking, prio_hi));
  CRD_CUDA(cudaEventCreateWithFlags(&g->ev_y, cudaEventDisableTiming));
  CRD_CUDA(cudaEventCreateWithFlags(&g->ev_b, cudaEventDisableTiming));
  return 0;
}

static int post_state(crd_grid *g, const StateRef &S) {
  if (!g->connected) return 0;  // single rank: the slab wraps onto itself
  if (g->epoch != g->computed) { set_error("crd_rhs_post_halo: previous epoch was posted but never computed"); return -1; }
  g->split = g->overlap && g->nyl >= 4 * kEdgeRows;
  if (!g->split) return launch_push(g, S, g->ctx->stream);
  if (ensure_aux(g)) return -1;
  CRD_CUDA(cudaEventRecord(g->ev_y, g->ctx->stream));      // the state is complete at this point of the main stream
  CRD_CUDA(cudaStreamWaitEvent(g->s_aux, g->ev_y, 0));
  return launch_push(g, S, g->s_aux);
}

static int compute_state(crd_grid *g, double t, const StateRef &S, double *ydot) {
  cudaStream_t st = g->ctx->stream;
  const long long nyl = g->nyl, B = kEdgeRows;
  if (!g->connected) {
    RhsArgs a = make_args(g, t, S, ydot, 0, nyl, slab_row(nyl - 1), slab_row(0));
    if (launch_rhs(g, a, st)) return -1;
    g->rhs_count++;
    return 0;
  }
  if (g->epoch == g->computed) { set_error("crd_rhs_compute: no halo posted for this evaluation"); return -1; }
  HaloLayout L{g->nx};
  const int par = (int)(g->epoch & 1ULL);
  const double *gs = (const double *)(g->halo_local + L.ghost_off(par, 0));
  const double *gn = (const double *)(g->halo_local + L.ghost_off(par, 1));
  if (!g->split) {
    if (launch_wait(g, st)) return -1;
    RhsArgs a = make_args(g, t, S, ydot, 0, nyl, ext_row(gs), ext_row(gn));
    if (launch_rhs(g, a, st)) return -1;
  } else {
    // interior rows on the main stream: their neighbours are rows of this slab
    RhsArgs ai = make_args(g, t, S, ydot, B, nyl - B, slab_row(B - 1), slab_row(nyl - B));
    if (launch_rhs(g, ai, st)) return -1;
    // edge rows on the auxiliary stream, once the neighbours' rows of this epoch have landed
    if (launch_wait(g, g->s_aux)) return -1;
    RhsArgs as = make_args(g, t, S, ydot, 0, B, ext_row(gs), slab_row(B));
    RhsArgs an = make_args(g, t, S, ydot, nyl - B, nyl, slab_row(nyl - B - 1), ext_row(gn));
    if (launch_rhs(g, as, g->s_aux) || launch_rhs(g, an, g->s_aux)) return -1;
    CRD_CUDA(cudaEventRecord(g->ev_b, g->s_aux));
    CRD_CUDA(cudaStreamWaitEvent(st, g->ev_b, 0));
  }
  g->computed = g->epoch;
  g->rhs_count++;
  return 0;
}

static StateRef plain_state(const double *y) { StateRef S; S.y = y; S.n = 0; return S; }

int crd_rhs_post_halo(crd_grid *g, const double *y) {
  if (!g || !y) { set_error("crd_rhs_post_halo: null argument"); return -1; }
  if (use(g->ctx)) return -1;
  return post_state(g, plain_state(y));
}

int crd_rhs_compute(crd_grid *g, double t, const double *y, double *ydot) {
  if (!g || !y || !ydot) { set_error("crd_rhs_compute: null argument"); return -1; }
  if (use(g->ctx)) return -1;
  return compute_state(g, t, plain_state(y), ydot);
}

int crd_rhs(crd_grid *g, double t, const double *y, double *ydot) {
  if (crd_rhs_post_halo(g, y)) return -1;
  return crd_rhs_compute(g, t, y, ydot);
}

// ydot = f(t, sum_j c[j]*X[j]) without materialising the combination (explicit RK stage assembly fused into the
// evaluation).  n <= 5; the vectors must not alias ydot.
int crd_rhs_lincomb(crd_grid *g, double t, int n, const double *c, const double *const *X_dev, double *ydot) {
  if (!g || !c || !X_dev || !ydot || n < 1 || n > kMaxLc) { set_error("crd_rhs_lincomb: bad arguments"); return -1; }
  if (use(g->ctx)) return -1;
  StateRef S;
  S.n = n;
  for (int j = 0; j < n; ++j) {
    if (!X_dev[j] || X_dev[j] == ydot) { set_error("crd_rhs_lincomb: null or aliased vector"); return -1; }
    S.x[j] = X_dev[j]; S.c[j] = c[j];
  }
  if (post_state(g, S)) return -1;
  return compute_state(g, t, S, ydot);
}

int crd_f_lincomb(realtype t, int n, const realtype *c, N_Vector *X, N_Vector ydot, void *user_data) {
  crd_grid *g = (crd_grid *)user_data;
  if (!g || !X || !ydot || n < 1 || n > kMaxLc) return -1;
  const double *xs[kMaxLc];
  for (int j = 0; j < n; ++j) {
    xs[j] = N_VGetDeviceArrayPointer_Crd(X[j]);
    if (!xs[j] || N_VGetLocalLength_Crd(X[j]) != crd_grid_local_length(g)) { set_error("crd_f_lincomb: vector does not match the grid"); return -1; }
  }
  return crd_rhs_lincomb(g, t, n, c, xs, N_VGetDeviceArrayPointer_Crd(ydot)) == 0 ? 0 : -1;
}

int crd_f(realtype t, N_Vector y, N_Vector ydot, void *user_data) {
  crd_grid *g = (crd_grid *)user_data;
  if (!g || !y || !ydot) return -1;
  const double *yd = N_VGetDeviceArrayPointer_Crd(y);
  double *fd = N_VGetDeviceArrayPointer_Crd(ydot);
  if (!yd || !fd) return -1;
  if (N_VGetLocalLength_Crd(y) != crd_grid_local_length(g)) { set_error("crd_f: vector length does not match the grid"); return -1; }
  return crd_rhs(g, t, yd, fd) == 0 ? 0 : -1;
}

// Host-buffer entry: stream the slab in row chunks, H2D / kernel / D2H on three streams.
int crd_rhs_host(crd_grid *g, double t, const double *y_host, double *ydot_host) {
  if (!g || !y_host || !ydot_host) { set_error("crd_rhs_host: null argument"); return -1; }
  if (g->connected && g->epoch != g->computed) { set_error("crd_rhs_host: previous epoch was posted but never computed"); return -1; }
  if (use(g->ctx)) return -1;
  const long long nx = g->nx, nyl = g->nyl;
  const size_t row_bytes = (size_t)2 * nx * sizeof(double);
  if (!g->stage_y) {
    CRD_CUDA(cudaMalloc(&g->stage_y, row_bytes * nyl));
    CRD_CUDA(cudaMalloc(&g->stage_ydot, row_bytes * nyl));
    CRD_CUDA(cudaStreamCreateWithFlags(&g->s_in, cudaStreamNonBlocking));
    CRD_CUDA(cudaStreamCreateWithFlags(&g->s_out, cudaStreamNonBlocking));
    // chunks of ~64 MiB, at least 1 row, at most 64 chunks
    long long rows_per = (long long)((64ull << 20) / row_bytes);
    if (rows_per < 1) rows_per = 1;
    long long n = (nyl + rows_per - 1) / rows_per;
    if (n > 64) n = 64;
    if (n < 1) n = 1;
    g->n_chunks = (int)n;
    g->ev_in = new cudaEvent_t[n]; g->ev_k = new cudaEvent_t[n];
    for (int i = 0; i < n; ++i) {
      CRD_CUDA(cudaEventCreateWithFlags(&g->ev_in[i], cudaEventDisableTiming));
      CRD_CUDA(cudaEventCreateWithFlags(&g->ev_k[i], cudaEventDisableTiming));
    }
  }
  const int C = g->n_chunks;
  cudaStream_t sk = g->ctx->stream;
  auto r_begin = [&](int c) { return nyl * c / C; };
  // the periodic wrap makes chunk 0 need the last row: send it first
  CRD_CUDA(cudaMemcpyAsync(g->stage_y + 2 * (nyl - 1) * nx, y_host + 2 * (nyl - 1) * nx, row_bytes, cudaMemcpyHostToDevice, g->s_in));
  const double *ghost_s = nullptr, *ghost_n = nullptr;
  if (g->connected) {
    // ring: the neighbours need this slab's first and last row before anything else
    CRD_CUDA(cudaMemcpyAsync(g->stage_y, y_host, row_bytes, cudaMemcpyHostToDevice, g->s_in));
    CRD_CUDA(cudaEventRecord(g->ev_k[0], g->s_in));
    CRD_CUDA(cudaStreamWaitEvent(sk, g->ev_k[0], 0));
    if (launch_push(g, plain_state(g->stage_y), sk)) return -1;
    if (launch_wait(g, sk)) return -1;
    HaloLayout L{g->nx};
    const int par = (int)(g->epoch & 1ULL);
    ghost_s = (const double *)(g->halo_local + L.ghost_off(par, 0));
    ghost_n = (const double *)(g->halo_local + L.ghost_off(par, 1));
    g->computed = g->epoch;
  }
  for (int c = 0; c < C; ++c) {
    const long long r0 = r_begin(c), r1 = r_begin(c + 1);
    CRD_CUDA(cudaMemcpyAsync(g->stage_y + 2 * r0 * nx, y_host + 2 * r0 * nx, row_bytes * (r1 - r0), cudaMemcpyHostToDevice, g->s_in));
    CRD_CUDA(cudaEventRecord(g->ev_in[c], g->s_in));
  }
  for (int c = 0; c < C; ++c) {
    const long long r0 = r_begin(c), r1 = r_begin(c + 1);
    if (r1 == r0) continue;
    CRD_CUDA(cudaStreamWaitEvent(sk, g->ev_in[c + 1 < C ? c + 1 : c], 0));
    const RowRef south = (ghost_s && r0 == 0) ? ext_row(ghost_s) : slab_row((r0 == 0 ? nyl : r0) - 1);
    const RowRef north = (ghost_n && r1 == nyl) ? ext_row(ghost_n) : slab_row(r1 == nyl ? 0 : r1);
    RhsArgs a = make_args(g, t, plain_state(g->stage_y), g->stage_ydot, r0, r1, south, north);
    if (launch_rhs(g, a, sk)) return -1;
    CRD_CUDA(cudaEventRecord(g->ev_k[c], sk));
    CRD_CUDA(cudaStreamWaitEvent(g->s_out, g->ev_k[c], 0));
    CRD_CUDA(cudaMemcpyAsync(ydot_host + 2 * r0 * nx, g->stage_ydot + 2 * r0 * nx, row_bytes * (r1 - r0), cudaMemcpyDeviceToHost, g->s_out));
  }
  CRD_CUDA(cudaStreamSynchronize(g->s_out));
  CRD_CUDA(cudaStreamSynchronize(sk));
  g->rhs_count++;
  return 0;
}

int crd_fill_initial_conditions(crd_grid *g, const crd_ic_params *ic, double *y_dev) {
  if (!g || !ic || !y_dev) { set_error("crd_fill_initial_conditions: null argument"); return -1; }
  if (use(g->ctx)) return -1;
  IcArgs a;
  a.model = g->p.model; a.vary_beta = g->p.vary_beta; a.wave_inside = ic->wave_inside; a.ic_type = ic->ic_type;
  a.nx = g->nx; a.nyl = g->nyl; a.js = g->js;
  a.dx = g->dx; a.dy = g->dy; a.xmin = g->xmin; a.ymin = g->ymin;
  // FHNmodel_torus.cpp:199-200,285-296 ; FHNmodel_flat.cpp:274-276
  a.wave_length = (g->ymax - g->ymin) * ic->wave_length;
  a.wave_width = (g->xmax - g->xmin) * ic->wave_width;
  const bool torus = is_torus(g->p.model), fhn = is_fhn(g->p.model);
  double mid;
  if (torus) {
    if (ic->wave_inside == 1) { mid = kPI; a.wave_xmin = mid - a.wave_width / 2.0; a.wave_xmax = mid + a.wave_width / 2.0; }
    else { mid = 0.0; a.wave_xmin = mid - a.wave_width / 2.0 + (g->xmax - g->xmin); a.wave_xmax = mid + a.wave_width / 2.0; }
  } else {
    mid = g->p.surface_width / 2.0; a.wave_xmin = mid - a.wave_width / 2.0; a.wave_xmax = mid + a.wave_width / 2.0;
  }
  a.s0 = ic->s0; a.s1 = ic->s1;
  a.p0 = fhn ? ic->s0 + 2 : ic->s0 + 1;      // Us + 2 | Zs + 1
  a.p1 = fhn ? ic->s1 + 1.5 : ic->s1 + 1;    // Vs + 1.5 | Ys + 1
  if (!fhn && g->p.vary_beta == 1 && ic->ic_type == 2) {
    set_error("icType = 2 (unseeded rand(), GoldbeterModel_flat.cpp:373-374) must be generated on the host");
    return -1;
  }
  const long long work = g->nx * g->nyl;
  ic_kernel<<<(unsigned)((work + 255) / 256), 256, 0, g->ctx->stream>>>(a, reinterpret_cast<double2 *>(y_dev));
  return check_launch(g->ctx, "ic_kernel");
}

}  // extern "C"

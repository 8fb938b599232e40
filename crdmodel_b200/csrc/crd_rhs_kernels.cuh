// crd_rhs_kernels.cuh — device code of the right-hand side: per-point arithmetic (EXACT / FAST), the direct,
// tiled (TMA bulk-copy) and streaming (persistent, mbarrier row ring) kernels, the halo push / wait kernels, the
// initial-condition kernel, and their launchers.  Included once, by crd_rhs.cu, which holds the C ABI.
#pragma once
#include <cmath>
#include <cstdlib>
#include <vector>

#include "crd_fused.cuh"
#include "crd_grid.cuh"
#include "crd_rhs_point.cuh"

using namespace crd;

namespace {

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
      c[r] = state2<LC, EXACT>(a, row + i);
      uw[r] = stateu<LC, EXACT>(a, row + iw);
      ue[r] = stateu<LC, EXACT>(a, row + ie);
    } else {
      c[r] = make_double2(0.0, 0.0); uw[r] = 0.0; ue[r] = 0.0;
    }
  }
  uu[0] = (j0 == 0) ? ghost_u<LC, EXACT>(a, a.south, a.south_off, i) : stateu<LC, EXACT>(a, (j0 - 1) * nx + i);
  {
    const long long jn = j0 + nrows;  // row above the last one this thread computes
    uu[RY + 1] = (jn == nyl) ? ghost_u<LC, EXACT>(a, a.north, a.north_off, i) : stateu<LC, EXACT>(a, jn * nx + i);
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

// ---- halo exchange inside a launch (crd_grid.cuh: HaloSync) ------------------------------------------------------------
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// one thread: until the strip's flag has reached this evaluation's epoch (acquire, system scope: the rows were stored by
// another GPU) or the timeout has passed — then the error word is raised and the launch goes on with stale rows; the host
// sees the word at its next wait for the stream and fails the context (crd_common.cuh: device_failed)
__device__ __noinline__ void halo_acquire(const unsigned long long *flag, unsigned long long epoch, long long timeout_ns, int *err, int code) {
  unsigned long long t0 = 0;
  for (unsigned spin = 0;; ++spin) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= epoch) break;
    if ((spin & 15u) == 15u) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if ((long long)(now - t0) > timeout_ns) { *err = code; __threadfence_system(); break; }
      __nanosleep(200);
    }
  }
  // rows fetched next by the async proxy (TMA bulk copies) must be ordered after the acquire as well
  asm volatile("fence.proxy.async;" ::: "memory");
}
// 256 threads: strip `strip` of this slab's first row -> prev's north ghost row, of its last row -> next's south ghost row
template <bool LC, bool SEQ>
__device__ __forceinline__ void halo_push_strip(const RhsArgs &a, long long strip, int t) {
  const long long col = strip * kHaloStrip + t;
  if (col < a.nx) {
    reinterpret_cast<double2 *>(a.hs.push_prev)[col] = state2<LC, SEQ>(a, col);
    reinterpret_cast<double2 *>(a.hs.push_next)[col] = state2<LC, SEQ>(a, (a.nyl - 1) * a.nx + col);
  }
  __threadfence_system();
}
// one thread, after a barrier over the pushing threads: the strip is complete over there
__device__ __forceinline__ void halo_publish(const RhsArgs &a, long long strip) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.hs.flag_prev + strip), "l"(a.hs.epoch) : "memory");
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.hs.flag_next + strip), "l"(a.hs.epoch) : "memory");
}
// work that touches a ghost row goes last (edge_last): index b of n_outer x n_inner items -> (outer, inner), the first and the
// last outer index after all the others
__device__ __forceinline__ void edge_last_coords(long long b, long long n_outer, long long n_inner, int edge_last, long long &outer, long long &inner) {
  if (edge_last && n_outer >= 3) {
    const long long mid = (n_outer - 2) * n_inner;
    if (b < mid) { outer = 1 + b / n_inner; inner = b - (outer - 1) * n_inner; }
    else { const long long e = b - mid; outer = (e < n_inner) ? 0 : n_outer - 1; inner = (e < n_inner) ? e : e - n_inner; }
  } else {
    outer = b / n_inner; inner = b - outer * n_inner;
  }
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
  constexpr size_t kTileBytes = (size_t)(TY + 2) * PITCH * 16;
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw + kTileBytes);

  const long long nx = a.nx, nyl = a.nyl;
  const long long tiles_x = (nx + TX - 1) / TX;
  long long ty, tx;
  edge_last_coords(blockIdx.x, (nyl + TY - 1) / TY, tiles_x, a.hs.edge_last, ty, tx);
  const long long i0 = tx * TX, j0 = ty * TY;
  const int w = (nx - i0 < TX) ? (int)(nx - i0) : TX;            // valid columns of this tile
  const int h = (nyl - j0 < TY) ? (int)(nyl - j0) : TY;          // valid rows of this tile
  const unsigned bar = smem_u32(mbar);

  // phi-split grid, exchange inside this launch: the first CTAs push the slab's boundary rows to the neighbours before
  // anything else (so a CTA never waits before it has pushed); a tile that touches a ghost row acquires that strip's flag
  if (a.hs.push_prev) {
    const long long nstrips = (nx + kHaloStrip - 1) / kHaloStrip;
    for (long long sp = blockIdx.x; sp < nstrips; sp += gridDim.x) {
      halo_push_strip<LC, EXACT>(a, sp, threadIdx.x);
      __syncthreads();
      if (threadIdx.x == 0) halo_publish(a, sp);
    }
  }
  {
    const bool need_s = a.hs.wait_south != nullptr && j0 == 0, need_n = a.hs.wait_north != nullptr && j0 + h == nyl;
    if (need_s || need_n) {
      if (threadIdx.x == 0) {
        if (need_s) halo_acquire(a.hs.wait_south + i0 / kHaloStrip, a.hs.epoch, a.hs.timeout_ns, a.hs.err, 100);
        if (need_n) halo_acquire(a.hs.wait_north + i0 / kHaloStrip, a.hs.epoch, a.hs.timeout_ns, a.hs.err, 101);
      }
      __syncthreads();
    }
  }

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
      if (jr < 0) v = a.south ? reinterpret_cast<const double2 *>(a.south)[col] : state2<true, EXACT>(a, a.south_off + col);
      else if (jr >= nyl) v = a.north ? reinterpret_cast<const double2 *>(a.north)[col] : state2<true, EXACT>(a, a.north_off + col);
      else v = state2<true, EXACT>(a, jr * nx + col);
      tile[r * PITCH + sc] = v;
    };
    const int r_lo = (j0 == 0) ? 1 : 0, r_hi = (j0 + h == nyl) ? h + 1 : h + 2;   // tile rows that are rows of this launch
    constexpr int R = 9;   // (TY + 2) = 18 rows in two passes, 9 loads per input vector in flight per thread
    for (int tcol = threadIdx.x; tcol < w; tcol += 256) {
      const long long p0 = (j0 - 1) * nx + i0 + tcol;   // point offset of tile row 0 in this column
      for (int rb = r_lo; rb < r_hi; rb += R) {
        double2 acc[R], v[R];
        auto load = [&](int j) {
#pragma unroll
          for (int rr = 0; rr < R; ++rr)
            v[rr] = (rb + rr < r_hi) ? reinterpret_cast<const double2 *>(a.lc_x[j])[p0 + (rb + rr) * nx] : make_double2(0.0, 0.0);
        };
        if constexpr (!EXACT) {
          load(0);
#pragma unroll
          for (int rr = 0; rr < R; ++rr) acc[rr] = make_double2(a.lc_c[0] * v[rr].x, a.lc_c[0] * v[rr].y);
#pragma unroll
          for (int j = 1; j < kMaxLc; ++j) {
            if (j < a.nlc) {
              load(j);
#pragma unroll
              for (int rr = 0; rr < R; ++rr) { acc[rr].x = fma(a.lc_c[j], v[rr].x, acc[rr].x); acc[rr].y = fma(a.lc_c[j], v[rr].y, acc[rr].y); }
            }
          }
        } else {
          // the order of lc_value<true>: the increments first (sdata), the vector they are added to last
          if (a.nlc > 1) {
            load(1);
#pragma unroll
            for (int rr = 0; rr < R; ++rr) acc[rr] = make_double2(__dadd_rn(__dmul_rn(a.lc_c[1], v[rr].x), 0.0), __dadd_rn(__dmul_rn(a.lc_c[1], v[rr].y), 0.0));
#pragma unroll
            for (int j = 2; j < kMaxLc; ++j) {
              if (j < a.nlc) {
                load(j);
#pragma unroll
                for (int rr = 0; rr < R; ++rr) { acc[rr].x = __dadd_rn(__dmul_rn(a.lc_c[j], v[rr].x), acc[rr].x); acc[rr].y = __dadd_rn(__dmul_rn(a.lc_c[j], v[rr].y), acc[rr].y); }
              }
            }
            load(0);
#pragma unroll
            for (int rr = 0; rr < R; ++rr) { acc[rr].x = __dadd_rn(__dmul_rn(a.lc_c[0], v[rr].x), acc[rr].x); acc[rr].y = __dadd_rn(__dmul_rn(a.lc_c[0], v[rr].y), acc[rr].y); }
          } else {
            load(0);
#pragma unroll
            for (int rr = 0; rr < R; ++rr) acc[rr] = make_double2(__dmul_rn(a.lc_c[0], v[rr].x), __dmul_rn(a.lc_c[0], v[rr].y));
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
  static bool attr_set[64] = {};   // the attribute is per device
  const int dev = g->ctx->device & 63;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set[dev] = true;
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
// 2 or 3 CTAs per SM (MINB; the ring of the latter is smaller) stay resident and walk over (strip of 256 columns) x (segment
// of rows) units.  Warp 8 is the
// producer: one lane issues a 1-D TMA bulk copy per row and input vector into the next free ring slot and arms
// that slot's "full" mbarrier with the byte count.  Warps 0..7 are consumers: a thread owns one column, waits for
// the slot, reads (and, for a fused stage, combines sum_j c_j x_j of) its point and the two neighbours' u, releases
// the slot on the "empty" mbarrier, and with the previous two rows still in registers computes the row before.
// Every state row is fetched once (plus one halo row per segment end); no CTA start-up per tile.
// a shared-memory load the compiler may not merge with an earlier load of the same address (FIN 2 reads a slot twice on purpose)
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}

// FIN (the last stage of an explicit RK step, NV = s vectors yn, F_0 .. F_{s-2}): the stage derivative F_{s-1} is not stored;
// instead, while it is in registers next to the raw yn and F_j of the same point, the consumer forms ynew = yn + sum hb_j F_j
// (written to a.ydot), err = sum hd_j F_j and the two weighted square sums of erk_finish_kernel (crd_fused.cuh arithmetic, so
// ynew has the same bits; the weights' reciprocals are branch-free, ~1 ulp).  Saves the finish kernel's pass over s + 1 vectors:
// 112 of 528 B per point and step.
struct StageFin {
  double hb[kMaxLc], hd[kMaxLc];
  double rtol, atol;
  unsigned hb_nz;    // bit j: hb[j] != 0
  double y2_bound;   // >= 0: the second sum is skipped and this bound reported instead (crd_fused.cuh: finish_y2_bound)
  double *partial;   // [3][kRedBlocks] per-CTA sums: error sum hi | sum (ynew w')^2 | error sum lo
};

// FIN = 2: the same arithmetic with a smaller register footprint (fits 3 CTAs per SM): instead of the raw vectors of the row
// that goes out next, a thread carries yn and the partial sums yn + sum_{j<NV-1} hb_j F_j, sum_{j<NV-1} hd_j F_j (same operation
// order, the last term is added when F_{NV-1} exists), formed from a second read of the ring slot after the row before has gone out.
template <int MODEL, bool EXACT, int NV, bool PLAIN, int RB, int FIN, int MINB = 2>
__global__ void __launch_bounds__(288, MINB) rhs_stream_kernel(const RhsArgs a, int seg_rows, int S, const StageFin fz) {
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
      long long strip, seg;
      edge_last_coords(u, segs, strips, a.hs.edge_last, seg, strip);
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
          if (ext) {   // exchange inside this launch: the neighbour's strip of that row must have landed
            if (jr < 0 && a.hs.wait_south) halo_acquire(a.hs.wait_south + strip, a.hs.epoch, a.hs.timeout_ns, a.hs.err, 102);
            if (jr >= nyl && a.hs.wait_north) halo_acquire(a.hs.wait_north + strip, a.hs.epoch, a.hs.timeout_ns, a.hs.err, 103);
          }
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
  // A thread owns one column and walks the rows of the unit as they arrive: row jr comes in (loaded, combined, its west /
  // east u taken from the neighbouring lanes), row jr-1 goes out from the three rows held in registers.  Only the first two
  // and the last row of a unit can be anything special (an already combined ghost row, no output yet, a frozen boundary row):
  // they take the checked path; every other row takes the lean one (the kernel is issue-bound: profiles/README.md).
  const int c = threadIdx.x;   // 0..255: column inside the strip
  const int react_on = a.react;
  const int nxi = (int)nx, nyli = (int)nyl;
  int slot_i = 0;
  unsigned slot_par = 0;
  FinAcc<EXACT> facc, faccy;   // FIN: this thread's share of sum (err w)^2 (double-double when EXACT) and sum (ynew w')^2, one
                               // accumulator per component so that the two dependency chains of a point run side by side
  const bool want_y2 = FIN && fz.y2_bound < 0.0, nz_last = (fz.hb_nz >> (NV - 1)) & 1u;
  double lcc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) lcc[j] = a.lc_c[j];
  const bool edge_lane = (lane == 0) || (lane == 31);
  const int edge_q = (lane == 0) ? c : c + 2;   // slot index of the neighbour column that lives in another warp
  // phi-split grid, exchange inside this launch: before anything else the consumers of the first CTAs push the slab's
  // boundary rows (of the stage state when this is a fused stage) to the neighbours
  if (a.hs.push_prev) {
    const long long nstrips = (nx + kHaloStrip - 1) / kHaloStrip;
    for (long long sp = blockIdx.x; sp < nstrips; sp += gridDim.x) {
      halo_push_strip<!PLAIN, EXACT>(a, sp, c);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (c == 0) halo_publish(a, sp);
    }
  }
  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    long long strip_l, seg_l;
    edge_last_coords(u, segs, strips, a.hs.edge_last, seg_l, strip_l);
    const int strip = (int)strip_l, seg = (int)seg_l;
    const int i0 = strip * TX, jA = seg * seg_rows, jB = (jA + seg_rows < nyli) ? jA + seg_rows : nyli;
    const int w = (nxi - i0 < TX) ? (nxi - i0) : TX;
    const bool active = c < w;
    double t1 = 0.0, t3 = 0.0;
    if (is_torus(MODEL) && active) {
      const double2 tc = reinterpret_cast<const double2 *>(a.cth)[i0 + c];
      t1 = tc.x; t3 = tc.y;
    }
    double2 *out = reinterpret_cast<double2 *>(a.ydot) + (long long)jA * nx + (i0 + c);
    const double *bp = a.brow + jA;   // beta of the next row to go out
    double uS = 0.0, cW = 0.0, cE = 0.0;
    double2 cc = make_double2(0.0, 0.0);
    double2 pv[FIN == 1 ? NV : 1] = {};   // FIN 1: the raw vectors (yn, F_0 ..) of the row that goes out next (FIN 2: yn)
    double2 psum = make_double2(0.0, 0.0), perr = make_double2(0.0, 0.0);   // FIN 2: the partial sums of that row
    // the arriving row: nn (combined centre), nW / nE (combined u of the neighbours), v (raw vectors, FIN)
    auto fetch = [&](const double2 *slot, bool ext, double2 &nn, double &nW, double &nE, double2 (&v)[NV]) {
      if (PLAIN || ext) {
        nn = slot[c + 1]; nW = slot[c].x; nE = slot[c + 2].x;
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = nn;
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = slot[j * PITCH + c + 1];
        double vx[NV], vy[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) { vx[j] = v[j].x; vy[j] = v[j].y; }
        nn = make_double2(lc_value_n<EXACT, NV>(lcc, vx), lc_value_n<EXACT, NV>(lcc, vy));
        nW = __shfl_up_sync(0xffffffffu, nn.x, 1);
        nE = __shfl_down_sync(0xffffffffu, nn.x, 1);
        if (edge_lane) {   // the neighbour lives in another warp: combine its u from the slot
          double ve[NV];
#pragma unroll
          for (int j = 0; j < NV; ++j) ve[j] = slot[j * PITCH + edge_q].x;
          const double e = lc_value_n<EXACT, NV>(lcc, ve);
          if (lane == 0) nW = e; else nE = e;
        }
      }
    };
    // row jl = (the row before the arriving one) goes out; frozen: it is a boundary row held at zero while t < tBoundary
    auto emit = [&](double uN, bool frozen) {
      double du = EXACT ? stencil_exact<MODEL>(a.k, t1, t3, cc.x, cW, cE, uS, uN)
                        : stencil_fast<MODEL>(a.k, t1, t3, cc.x, cW, cE, uS, uN);
      double dv = 0.0;
      if (react_on) {
        react<MODEL, EXACT>(a.k, __ldg(bp), cc.x, cc.y, du, dv);
        du = frozen ? 0.0 : du;
        dv = frozen ? 0.0 : dv;
      }
      if constexpr (!FIN) {
        if (active) *out = make_double2(du, dv);
      } else if constexpr (FIN == 2) {
        const double sx = fin_sol_term<EXACT>(fz.hb[NV - 1], du, psum.x, nz_last), ex = fin_err_term<EXACT>(fz.hd[NV - 1], du, perr.x);
        const double sy = fin_sol_term<EXACT>(fz.hb[NV - 1], dv, psum.y, nz_last), ey = fin_err_term<EXACT>(fz.hd[NV - 1], dv, perr.y);
        if (active) {
          *out = make_double2(sx, sy);
          finish_tail<EXACT>(fz.rtol, fz.atol, pv[0].x, sx, ex, facc, want_y2);
          finish_tail<EXACT>(fz.rtol, fz.atol, pv[0].y, sy, ey, faccy, want_y2);
        }
      } else {
        // ynew = yn + sum_j hb_j F_j, err = sum_j hd_j F_j with F_{NV-1} = (du, dv): the operation order of finish_elem
        double sx = pv[0].x, sy = pv[0].y, ex = 0.0, ey = 0.0;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const double2 f = (j == NV - 1) ? make_double2(du, dv) : pv[j + 1 < NV ? j + 1 : 0];
          const bool nz = (j < NV - 1) || nz_last;   // the stored stages' weights are non-zero (checked on the host)
          sx = fin_sol_term<EXACT>(fz.hb[j], f.x, sx, nz); ex = fin_err_term<EXACT>(fz.hd[j], f.x, ex);
          sy = fin_sol_term<EXACT>(fz.hb[j], f.y, sy, nz); ey = fin_err_term<EXACT>(fz.hd[j], f.y, ey);
        }
        if (active) {
          *out = make_double2(sx, sy);
          finish_tail<EXACT>(fz.rtol, fz.atol, pv[0].x, sx, ex, facc, want_y2);
          finish_tail<EXACT>(fz.rtol, fz.atol, pv[0].y, sy, ey, faccy, want_y2);
        }
      }
      out += nx;
      ++bp;
    };
    auto rotate = [&](const double2 &nn, double nW, double nE, const double2 (&v)[NV], const double2 *slot, bool ext) {
      uS = cc.x; cc = nn; cW = nW; cE = nE;
      if constexpr (FIN == 1) {
#pragma unroll
        for (int j = 0; j < NV; ++j) pv[j] = v[j];
      } else if constexpr (FIN == 2) {
        if (!ext) {   // (a ghost row never goes out)
          const unsigned sa = smem_u32(slot + c + 1);
          pv[0] = lds_f64x2(sa);
          psum = pv[0]; perr = make_double2(0.0, 0.0);
#pragma unroll
          for (int j = 0; j < NV - 1; ++j) {
            const double2 f = lds_f64x2(sa + (unsigned)((j + 1) * PITCH * 16));
            // (the stored stages' weights are non-zero: checked on the host before the launch)
            psum.x = fin_sol_term<EXACT>(fz.hb[j], f.x, psum.x, true); perr.x = fin_err_term<EXACT>(fz.hd[j], f.x, perr.x);
            psum.y = fin_sol_term<EXACT>(fz.hb[j], f.y, psum.y, true); perr.y = fin_err_term<EXACT>(fz.hd[j], f.y, perr.y);
          }
        }
      }
    };
    const bool ext_s = (jA == 0) && a.south != nullptr, ext_n = (jB == nyli) && a.north != nullptr;
    const bool frz_first = react_on && a.freeze_south && jA == 0, frz_last = react_on && a.freeze_north && jB == nyli;
    for (int j0 = jA - 1; j0 <= jB; j0 += RB) {
      mbar_wait(bars + 8u * slot_i, slot_par);
      const double2 *stage = ring + (size_t)slot_i * RB * NV * PITCH;
      const int nr = (jB - j0 + 1 < RB) ? (jB - j0 + 1) : RB;
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        if (rr < nr) {
          const int jr = j0 + rr;
          const double2 *slot = stage + (size_t)rr * NV * PITCH;
          double2 nn, v[NV];
          double nW, nE;
          if (jr > jA + 1 && jr < jB) {          // steady state: plain row in, plain row out
            fetch(slot, false, nn, nW, nE, v);
            emit(nn.x, false);
            rotate(nn, nW, nE, v, slot, false);
          } else {                               // the unit's first two rows and its last one
            const bool ext = (jr < jA && ext_s) || (jr == jB && ext_n);
            fetch(slot, ext, nn, nW, nE, v);
            if (jr > jA) emit(nn.x, (jr == jA + 1 && frz_first) || (jr == jB && frz_last));
            rotate(nn, nW, nE, v, slot, ext);
          }
        }
      }
      __syncwarp();   // every lane has read the stage (the values it still needs are in registers)
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8u * (S + slot_i)) : "memory");
      if (++slot_i == S) { slot_i = 0; slot_par ^= 1u; }
    }
  }
  if (FIN) {
    // per-CTA sums in a fixed order: shuffle tree, then the eight consumer warps in order (the producer warp has left:
    // named barrier over the 256 consumer threads); the error sum travels as a double-double pair
    __shared__ double fin_red[3][8];
    dd_merge(facc.e_hi, facc.e_lo, faccy.e_hi, faccy.e_lo);
    double eh = facc.e_hi, el = facc.e_lo, fy2 = facc.y2 + faccy.y2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { dd_shfl_down(eh, el, o); fy2 += __shfl_down_sync(0xffffffffu, fy2, o); }
    if (lane == 0) { fin_red[0][warp] = eh; fin_red[1][warp] = fy2; fin_red[2][warp] = el; }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x == 0) {
      double sh = fin_red[0][0], sy = fin_red[1][0], sl = fin_red[2][0];
      for (int w = 1; w < 8; ++w) { dd_merge(sh, sl, fin_red[0][w], fin_red[2][w]); sy += fin_red[1][w]; }
      if (!want_y2) sy = blockIdx.x == 0 ? fz.y2_bound : 0.0;   // the launch's bound, once
      fz.partial[blockIdx.x] = sh;
      fz.partial[kRedBlocks + blockIdx.x] = sy;
      fz.partial[2 * kRedBlocks + blockIdx.x] = sl;
    }
  }
}

// adds the per-CTA sums of a FIN stage (up to three launches, one region of [3][kRedBlocks] each) in a fixed order, the error
// sum in double-double; result[0] = hi, result[1] = sum (ynew w')^2, result[2] = lo go to mapped pinned host memory
// on a phi-split grid with the device-side allreduce wired (T != nullptr) the ranks' sums are exchanged here as well
__global__ void __launch_bounds__(256) fin_reduce_kernel(const double *partial, int nregions, int n0, int n1, int n2, double *result,
                                                         const CommTab *T, unsigned long long seq) {
  __shared__ double sm[3][8];
  __shared__ double gv[kCommVals];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double sh = 0.0, sl = 0.0, sy = 0.0;
  for (int r = 0; r < nregions; ++r) {
    const double *p = partial + (size_t)r * 3 * kRedBlocks;
    const int n = r == 0 ? n0 : r == 1 ? n1 : n2;
    for (int b = threadIdx.x; b < n; b += 256) { dd_merge(sh, sl, p[b], p[2 * kRedBlocks + b]); sy += p[kRedBlocks + b]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { dd_shfl_down(sh, sl, o); sy += __shfl_down_sync(0xffffffffu, sy, o); }
  if (lane == 0) { sm[0][warp] = sh; sm[1][warp] = sy; sm[2][warp] = sl; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sh = sm[0][0]; sy = sm[1][0]; sl = sm[2][0];
    for (int w = 1; w < 8; ++w) { dd_merge(sh, sl, sm[0][w], sm[2][w]); sy += sm[1][w]; }
    gv[0] = sh; gv[1] = sy; gv[2] = sl; gv[3] = 0.0;
  }
  __syncthreads();
  if (T) comm_allreduce_block(T, seq, COMM_SUM_DD, kCommVals, gv);
  if (threadIdx.x == 0) {
    result[0] = gv[0]; result[1] = gv[1]; result[2] = gv[2];
    __threadfence_system();
  }
}

// MINB = CTAs per SM: 2 (ring of ~100 KB, up to 113 registers) or 3 (ring of ~66 KB, up to 75 registers: 24 consumer warps per SM)
template <int MODEL, bool EXACT, int NV, bool PLAIN, int FIN = 0, int MINB = 2>
int launch_stream_nv(crd_grid *g, const RhsArgs &a, cudaStream_t st, const StageFin *fin = nullptr, int *nblocks = nullptr) {
  constexpr int RB = (NV == 1) ? 4 : (NV <= 3 ? 2 : 1);   // rows per ring stage
  const long long strips = (a.nx + 255) / 256;
  if (strips <= 0 || a.nyl <= 0) return 0;
  int seg_rows = stream_seg_rows(a.nyl, strips, (long long)MINB * g->ctx->sms);
  static const int seg_override = [] { const char *e = std::getenv("CRD_STREAM_SEG_ROWS"); return e ? std::atoi(e) : 0; }();   // profiling
  if (seg_override >= 1) seg_rows = seg_override;
  const long long segs = (a.nyl + seg_rows - 1) / seg_rows, units = strips * segs;
  const size_t stage_bytes = (size_t)RB * NV * 258 * 16;
  int S = (int)((MINB == 3 ? 68000 : 100000) / stage_bytes);
  if (S > 8) S = 8;
  if (S < 3) S = 3;
  const size_t smem = (size_t)S * stage_bytes + (size_t)2 * S * 8;
  auto kern = rhs_stream_kernel<MODEL, EXACT, NV, PLAIN, RB, FIN, MINB>;
  static bool attr_set[64] = {};   // the attribute is per device
  const int dev = g->ctx->device & 63;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110000);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set[dev] = true;
  }
  const long long ctas = units < (long long)MINB * g->ctx->sms ? units : (long long)MINB * g->ctx->sms;
  StageFin fz;
  if (fin) fz = *fin; else std::memset(&fz, 0, sizeof fz);
  if (nblocks) *nblocks = (int)ctas;
  kern<<<(unsigned)ctas, 288, smem, st>>>(a, seg_rows, S, fz);
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

// the last stage of a 5-stage method fused with the step finish (one launch of it: a whole slab, or the interior / an edge band
// of a phi-split one, ghost rows included); returns 1 when it does not apply
int launch_stage_finish(crd_grid *g, const RhsArgs &a, const StageFin &fin, cudaStream_t st, int *nblocks) {
  if (a.nlc != 5) return 1;
  const bool exact = g->p.arith == CRD_ARITH_EXACT;
  // default (and grid variant 22): the partial-sum form of the finish (FIN 2) with 3 CTAs per SM, 24 consumer warps: 4.58 vs
  // 4.90 ms at 16384^2 (EXACT; FAST 4.16 vs 4.79), same bits.  Variant 23: FIN 2 with 2 CTAs per SM; 24: FIN 1 (the raw vectors
  // in registers) with 2 CTAs per SM, the former default.  profiles/README.md
#define CRD_FIN_(M, F, B) (exact ? launch_stream_nv<M, true, 5, false, F, B>(g, a, st, &fin, nblocks) : launch_stream_nv<M, false, 5, false, F, B>(g, a, st, &fin, nblocks))
#ifdef CRD_PROFILING_VARIANTS
#define CRD_FIN(M) (g->variant == 24 ? CRD_FIN_(M, 1, 2) : g->variant == 23 ? CRD_FIN_(M, 2, 2) : CRD_FIN_(M, 2, 3))
#else
#define CRD_FIN(M) CRD_FIN_(M, 2, 3)
#endif
  // Goldbeter in EXACT arithmetic is bound by FP64 work and issue slots, not by bytes in flight: 2 CTAs per SM with the
  // registers that gives (no spills) are 6-8 % ahead of 3 (tools/prof_fin.py: 1.32-1.39 vs 1.44-1.48 ms at 8192 x 8192)
#define CRD_FIN_GB(M) ((exact && g->variant == 0) ? launch_stream_nv<M, true, 5, false, 2, 2>(g, a, st, &fin, nblocks) : CRD_FIN(M))
  switch (g->p.model) {
    case CRD_FHN_TORUS: return CRD_FIN(CRD_FHN_TORUS);
    case CRD_GOLDBETER_TORUS: return CRD_FIN_GB(CRD_GOLDBETER_TORUS);
    case CRD_FHN_FLAT: return CRD_FIN(CRD_FHN_FLAT);
    case CRD_GOLDBETER_FLAT: return CRD_FIN_GB(CRD_GOLDBETER_FLAT);
  }
#undef CRD_FIN_GB
#undef CRD_FIN
#undef CRD_FIN_
  return 1;
}

// Which kernel evaluates a launch of nx x nyl points with nlc input vectors (0 = a plain state) on this grid:
//   1..5 the direct kernel (rows per thread, min CTAs/SM): 1 (2,4) | 2 (8,2) | 3 (1,4) | 4 (4,3) | 5 (4,4)
//   10, 13, 15 the tiled kernel (TX, TY, min CTAs/SM): 10 (128,16,4) | 13 (256,16,3) | 15 = 13 with flag-and-redo instead of a
//              branch per point;   20, 21 the streaming kernel with 2 | 3 CTAs per SM.
// Grid variant 0 = automatic.  Large slabs (>= 2 Mi points, beyond L2 with their result): the TMA-tiled kernel wherever a tile row is reasonably
// full; the streaming kernel for the fused stages (its once-per-row fetch beats the tiled kernel's register-staged tiles: 3
// CTAs/SM for 2 or 3 input vectors, 2 for 5).  Small slabs (the reference's default 400 x 1600 / 100 x 400 grids live in L2 and
// are bound by launch latency and by how many CTAs a partial wave gets): the direct kernel with 2 rows per thread.  Measured:
// profiles/README.md.  Tilings that measured slower (11, 12, 14, 16-18; 23 / 24 for the fused finish) are compiled only with
// -DCRD_PROFILING_VARIANTS.  The tiled and the streaming kernels carry the halo exchange inside the launch (kernel_has_halo).
inline int resolved_variant(const crd_grid *g, long long nx, long long nyl, int nlc) {
  const bool exact = g->p.arith == CRD_ARITH_EXACT;
  int variant = g->variant;
  const bool big = nx * nyl >= (2LL << 20), mid = nx * nyl >= (1LL << 20);
  if (variant == 0) {
    const int tiled = (exact && !is_fhn(g->p.model)) ? 15 : 13;
    if (!big) variant = (mid && nlc > 0 && nx >= 192) ? tiled : 1;   // 1-2 Mi points: a fused stage is already faster tiled
    else variant = (nx >= 192) ? tiled : (nx >= 96) ? 10 : 5;
    if (nlc == 5 && nx >= 192 && big) variant = 20;
    if ((nlc == 2 || nlc == 3) && nx >= 192 && big) variant = 21;
  }
#ifndef CRD_PROFILING_VARIANTS
  if (variant >= 10 && variant != 10 && variant != 13 && variant != 15 && variant != 20 && variant != 21) variant = 13;   // not in this build
#endif
  if (variant < 1) variant = 1;
  const bool stream_ok = nlc == 0 || nlc == 2 || nlc == 3 || nlc == 5;
  if (variant == 21 && !(nlc == 0 || nlc == 2 || nlc == 3)) variant = 20;
  if (variant == 20 && !stream_ok) variant = 13;
  return variant;
}
inline bool kernel_has_halo(int variant) { return variant >= 10; }

template <int MODEL, bool EXACT>
int launch_model(crd_grid *g, const RhsArgs &a_in, cudaStream_t st) {
  const int variant = resolved_variant(g, a_in.nx, a_in.nyl, a_in.nlc);
  if (variant == 21) {   // streaming kernel with 3 CTAs per SM (plain state, 2 or 3 input vectors)
    if (a_in.nlc == 2) return launch_stream_nv<MODEL, EXACT, 2, false, false, 3>(g, a_in, st);
    if (a_in.nlc == 3) return launch_stream_nv<MODEL, EXACT, 3, false, false, 3>(g, a_in, st);
    return launch_stream_nv<MODEL, EXACT, 1, true, false, 3>(g, a_in, st);
  }
  if (variant == 20) return launch_stream<MODEL, EXACT>(g, a_in, st);   // streaming kernel (persistent CTAs, shared-memory row ring)
  switch (variant) {
    case 10: return launch_tile<MODEL, EXACT, 128, 16, 4, false>(g, a_in, st);
    case 13: return launch_tile<MODEL, EXACT, 256, 16, 3, false>(g, a_in, st);
    case 15: return launch_tile<MODEL, EXACT, 256, 16, 3, true>(g, a_in, st);   // flag-and-redo instead of a branch per point
#ifdef CRD_PROFILING_VARIANTS   // tilings that measured slower (profiles/README.md); not part of the shipped library
    case 11: return launch_tile<MODEL, EXACT, 128, 32, 3, false>(g, a_in, st);
    case 12: return launch_tile<MODEL, EXACT, 64, 32, 4, false>(g, a_in, st);
    case 14: return launch_tile<MODEL, EXACT, 128, 16, 3, false>(g, a_in, st);
    case 16: return launch_tile<MODEL, EXACT, 256, 16, 3, false, 6>(g, a_in, st);  // tile arrives in 6 row groups
    case 17: return launch_tile<MODEL, EXACT, 256, 16, 3, false, 2>(g, a_in, st);  // ... in 2
    case 18: return launch_tile<MODEL, EXACT, 256, 16, 3, false, 9>(g, a_in, st);  // ... in 9
#endif
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

// ---- halo exchange as launches of its own (kernels without the in-launch exchange; crd_rhs_post_halo) ------------------
template <bool LC, bool SEQ>
__global__ void __launch_bounds__(256) halo_push_kernel(const RhsArgs a) {
  const long long nstrips = (a.nx + kHaloStrip - 1) / kHaloStrip;
  for (long long sp = blockIdx.x; sp < nstrips; sp += gridDim.x) {
    halo_push_strip<LC, SEQ>(a, sp, threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) halo_publish(a, sp);
  }
}

// two boundary rows per side of a plain state (the pass that forms two evaluations at once reads two rows beyond the slab)
__global__ void __launch_bounds__(256) halo_push2_kernel(const RhsArgs a) {
  const long long nstrips = (a.nx + kHaloStrip - 1) / kHaloStrip;
  const double2 *y = reinterpret_cast<const double2 *>(a.y);
  for (long long sp = blockIdx.x; sp < nstrips; sp += gridDim.x) {
    const long long col = sp * kHaloStrip + threadIdx.x;
    if (col < a.nx) {
      reinterpret_cast<double2 *>(a.hs.push_prev)[col] = y[col];
      reinterpret_cast<double2 *>(a.hs.push2_prev)[col] = y[a.nx + col];
      reinterpret_cast<double2 *>(a.hs.push_next)[col] = y[(a.nyl - 1) * a.nx + col];
      reinterpret_cast<double2 *>(a.hs.push2_next)[col] = y[(a.nyl - 2) * a.nx + col];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) halo_publish(a, sp);
  }
}

__global__ void __launch_bounds__(256) halo_wait_kernel(const RhsArgs a) {
  const long long nstrips = (a.nx + kHaloStrip - 1) / kHaloStrip;
  for (long long q = threadIdx.x; q < 2 * nstrips; q += blockDim.x) {
    const bool north = q >= nstrips;
    halo_acquire((north ? a.hs.wait_north : a.hs.wait_south) + (north ? q - nstrips : q), a.hs.epoch, a.hs.timeout_ns, a.hs.err, north ? 105 : 104);
  }
}

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

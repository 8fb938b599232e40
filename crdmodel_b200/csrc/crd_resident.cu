// crd_resident.cu — the adaptive explicit RK step loop as ONE persistent cooperative kernel whose state lives in
// shared memory.
//
// What it replaces: the time loop inside ARKode() (reference call site src/FHNmodel_torus.cpp:423) as driven by
// crd_ark.cpp — per step 4 fused stage evaluations + the fused finish + f(tn, yn) = 6 dependent kernel launches and
// one host round trip for the error norm.  The reference's own meshes (400 x 1600 FHN, 100 x 400 Goldbeter,
// data/*.ini) are 10 MB / 0.6 MB per vector: a launch is a few us of work, the step is bound by launch latency, the
// host round trip and L2 round trips, not by HBM.  Here:
//   * the phi rows are cut into one BAND per SM (one CTA each, all co-resident: cooperative launch); a band's share
//     of the integrator's vectors (yn / ynew ping-pong, stage derivatives) is kept in that SM's shared memory for the
//     whole call — as many vectors as fit in 227 KB next to the stage tile, most-used first; the rest stay in the
//     band's slice of the global N_Vector arrays, which only this CTA touches (coalesced 16-byte streams, no halos);
//   * a pass = one stage evaluation F_i = f(tn + c_i h, yn + h sum_j A_ij F_j) in two phases.  Phase 1 forms the
//     stage state of every point of the band once (independent streaming loads, 4 points in flight per thread) into a
//     shared-memory TILE and copies the u of the band's first and last row into an exchange buffer — the only data
//     that crosses SMs.  One grid barrier.  Phase 2 is the stencil + reaction with every neighbour a shared-memory
//     read (the rows just outside the band come from the neighbours' exchange rows); the last stage's phase 2 also
//     forms ynew, the error estimate and the two weighted square sums (crd_fused.cuh) while F_s is in registers;
//   * nothing waits between phase 2 of one pass and phase 1 of the next (a thread reads back the points it wrote);
//   * error norm: per-CTA partials ride on the next barrier and are added by every CTA in the same fixed order, so all
//     CTAs compute the same dsm and run the error test and the PID step controller redundantly and identically.  The
//     pass after the last stage is fnew = f(tn + h, ynew) (dense output; also stage 1 of the next step): its phase 1
//     runs speculatively before the verdict, so an accepted step costs no extra barrier; a rejected one rebuilds the
//     tile with the smaller step.
// 5 passes and 5 barriers per step, no launch and no host involvement until tn has passed tout.  Per-point
// arithmetic, combination order and finish arithmetic are the shared device functions of crd_rhs_point.cuh /
// crd_fused.cuh: for the same step size the new state has the same bits as on the launch-per-stage path.
// Single GPU only (a phi-split run keeps the host-driven loop and its halo ring).
#include <cfloat>
#include <cmath>

#include "crd_fused.cuh"
#include "crd_rhs_point.cuh"
#include "../host/crd_pow.h"

namespace {

constexpr int kResStages = kMaxLc;   // yn + 4 stage derivatives in one combination: methods with s <= 5
// buffers handed in by the integrator
enum { R_YN = 0, R_YOLD = 1, R_YCUR = 2, R_FNEW = 3, R_FOLD = 4, R_F1 = 5, R_NBUF = 5 + kResStages - 1 };
// storages the loop uses, in the order they get shared memory (accesses per step: 3, 3, 3, 3, 2, 1.5, 1.5)
enum { ST_Y0 = 0, ST_Y1 = 1, ST_F1 = 2, ST_F2 = 3, ST_F3 = 4, ST_FA = 5, ST_FB = 6, ST_N = 7 };
__host__ __device__ constexpr int storage_home(int st) {
  return st == ST_Y0 ? R_YN : st == ST_Y1 ? R_YCUR : st == ST_FA ? R_FNEW : st == ST_FB ? R_FOLD : R_F1 + (st - ST_F1);
}

struct ResOut {            // mapped pinned host memory, written once by CTA 0 when the loop ends
  double tn, next_h, hold, eta, etamax, eh0, eh1, ynorm_sq, h_failed;
  long long nst, attempts, nfe, netf;
  int flag, err;
  int yi, fi;              // which of the ping-pong storages hold yn / fnew
  int done;
  long long cyc[6];        // CTA 0's cycles in: phase 1, interior rows, barrier wait, edge rows, rest, total
};

struct ResArgs {
  // grid
  long long nx, nyl;
  const double *cth, *brow;
  RhsConst k;
  int react, freeze_rows;   // freeze_rows: this slab owns global rows 0 and ny-1 (always, on one GPU)
  double t_boundary;
  // vectors (global homes) and where the band's share of each storage lives
  double *buf[R_NBUF];
  int nslots;               // storages 0 .. nslots-1 live in shared memory (after the stage tile)
  unsigned slot_bytes;      // bytes of one band-sized array (tile and storages alike)
  // method
  int s, p;
  double A[kResStages][kResStages], b[kResStages], d[kResStages], c[kResStages];
  // tolerances, controller
  double rtol, atol, k1, k2, k3, bias, safety, growth, etamxf, etamin, lbound, ubound;
  int small_nef, maxnef;
  double nglobal;
  // request and incoming state
  double tout;
  int itask;
  long long max_steps;
  double tn, next_h, hold, eta, etamax, eh0, eh1, ynorm_sq;
  // synchronisation, exchange, results
  unsigned long long *bar;   // [0] arrival counter (monotonic within a launch), [16] abort flag
  double *partial;           // [3][gridDim.x]: error sum hi | sum (ynew w')^2 | error sum lo
  double *xch;               // [2 parities][gridDim.x bands][2 sides][nx]: u of the band's first / last row
  ResOut *out;
};

struct ResLoop {   // evolving scalars; every CTA holds an identical copy in shared memory
  double tn, h, next_h, hold, eta, etamax, eh0, eh1, ynorm_sq, h_failed;
  long long nst, attempts, nfe, netf;
  int yi, fi;              // current yn = storage ST_Y0 + yi, current fnew = ST_FA + fi
  int nef, status, flag;   // status: 0 retry the step, 1 accepted, 2 stop
  int stop;                // set when the accepted step ends this call
  unsigned pass;           // passes so far: parity of the exchange buffer
};

struct ResFinish {
  const double2 *yn;
  const double2 *F[kResStages];   // F[0 .. s-2]; the last stage's derivative is still in registers
  double2 *ynew;
  double hb[kResStages], hd[kResStages];
  double rtol, atol;
  int s;
};

// ---- grid barrier -----------------------------------------------------------------------------------------------
// All CTAs are co-resident (cooperative launch).  Split in two so that work which needs nothing from other SMs (the
// band's interior rows) runs between them.  arrive: after a bar.sync has ordered the CTA's writes before it, thread 0
// adds 1 with release semantics (cumulative at gpu scope).  wait: thread 0 spins on an acquire load, which also drops
// the SM's stale L1 lines, and the bar.sync extends that to the whole CTA.  A wait that does not complete within ~2 s
// raises the abort flag, which every spinning CTA honours: the kernel ends with an error instead of hanging the GPU.
__device__ __forceinline__ void grid_arrive(unsigned long long *bar, unsigned long long &target) {   // thread 0, after bar.sync
  target += gridDim.x;
  asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(bar) : "memory");
}
__device__ __forceinline__ bool grid_wait(unsigned long long *bar, unsigned long long target, int *s_ok) {
  if (threadIdx.x == 0) {
    int ok = 1;
    long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
      unsigned long long v;
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if ((spin & 63u) == 63u) {
        unsigned long long ab;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(ab) : "l"(bar + 16) : "memory");
        if (ab != 0ULL) { ok = 0; break; }
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000LL) { atomicExch(bar + 16, 1ULL); ok = 0; break; }
      }
    }
    *s_ok = ok;
  }
  __syncthreads();
  return *s_ok != 0;
}

// ---- the stage state as a combination of storages ------------------------------------------------------------------
struct Comb {
  int n;
  const double2 *x[kMaxLc];
  double c[kMaxLc];
};
// h-independent description of each stage's combination, built once per launch
struct StageTab {
  int n[kResStages];
  int st[kResStages][kMaxLc];     // storage id; -1 = current yn, -2 = current fnew (F_0)
  double A[kResStages][kMaxLc];
};

// what a pass needs besides the combination (32-bit indices: a band is far below 2^31 points).  Threads are laid out
// as (row group ty, column tx): a thread keeps its column — metric coefficients and neighbour offsets are per-thread
// constants — and walks rows ty, ty + G, ...; meshes wider than the CTA loop over columns as well.
struct PassCfg {
  int nx, rows, n;                     // columns, rows and points of the band
  int ncol, G, tx, ty;                 // threads along theta, row groups, this thread's place (ty >= G: idle)
  const double2 *cth;                  // [nx] metric coefficients, shared-memory copy
  const double *brow;                  // [rows] beta row values of the band, shared-memory copy
  int react, freeze_south, freeze_north;
};

// phase 1: the stage state of every point of the band -> tile; its first / last row's u -> exchange buffer
template <bool SEQ, int N, int U>   // U: rows in flight per thread
__device__ __forceinline__ void phase1_n(const Comb &cb, const PassCfg &c, double2 *tile, double *xch_mine) {
  if (c.ty >= c.G) return;
  const int nx = c.nx, rows = c.rows, stride = c.G * nx;
  for (int i = c.tx; i < nx; i += c.ncol) {
    for (int r0 = c.ty; r0 < rows; r0 += c.G * U) {
      const int p0 = r0 * nx + i;
      double2 v[U][N];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = (r0 + u * c.G < rows) ? p0 + u * stride : p0;
#pragma unroll
        for (int j = 0; j < N; ++j) v[u][j] = cb.x[j][p];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = r0 + u * c.G;
        // sum_j c_j x_j in the operation order of state2<true, SEQ> (crd_fused.cuh)
        double cc[N], vx[N], vy[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { cc[j] = cb.c[j]; vx[j] = v[u][j].x; vy[j] = v[u][j].y; }
        const double2 s = make_double2(lc_value_n<SEQ, N>(cc, vx), lc_value_n<SEQ, N>(cc, vy));
        if (r < rows) {
          tile[p0 + u * stride] = s;
          if (r == 0) xch_mine[i] = s.x;
          if (r == rows - 1) xch_mine[nx + i] = s.x;
        }
      }
    }
  }
}
template <bool SEQ, int U>
__device__ __forceinline__ void phase1(const Comb &cb, const PassCfg &c, double2 *tile, double *xch_mine) {
  switch (cb.n) {
    case 1: phase1_n<SEQ, 1, U>(cb, c, tile, xch_mine); break;
    case 2: phase1_n<SEQ, 2, U>(cb, c, tile, xch_mine); break;
    case 3: phase1_n<SEQ, 3, U>(cb, c, tile, xch_mine); break;
    case 4: phase1_n<SEQ, 4, (U > 2 ? 2 : U)>(cb, c, tile, xch_mine); break;    // wide combinations: fewer rows in flight (registers)
    default: phase1_n<SEQ, 5, (U > 2 ? 2 : U)>(cb, c, tile, xch_mine); break;
  }
}

// phase 2 over rows r = r_first + k * r_step, k = 0 .. nr-1, of the band (row k goes to row group k % G): stencil +
// reaction with every neighbour a shared-memory read; the rows just outside the band are the neighbours' exchange
// rows.  Writes F to `out`, or (FIN: last stage) finishes the step with the last stage's derivative still in registers
// (crd_fused.cuh arithmetic).
template <int MODEL, bool EXACT, bool FIN>
__device__ __forceinline__ void phase2_rows(const PassCfg &c, const RhsConst &k_in, const double2 *tile, const double *xs, const double *xn,
                                            double2 *out, const ResFinish *fin, int r_first, int r_step, int nr, FinAcc<EXACT> &facc) {
  if (c.ty >= c.G) return;
  const int nx = c.nx, rows = c.rows;
  const int S = FIN ? fin->s : 0;
  const bool want_y2 = FIN && !(fin->rtol > 2.220446049250313e-16);   // (crd_fused.cuh: finish_y2_bound)
  const RhsConst k = k_in;   // registers: a generic store may alias shared memory, which would force a reload per row
  for (int i = c.tx; i < nx; i += c.ncol) {
    const int iw = (i == 0) ? nx - 1 : i - 1, ie = (i == nx - 1) ? 0 : i + 1;   // theta wraps
    double t1 = 0.0, t3 = 0.0;
    if (is_torus(MODEL)) {
      const double2 tc = c.cth[i];
      t1 = tc.x; t3 = tc.y;
    }
#pragma unroll 1
    for (int q = c.ty; q < nr; q += c.G) {
      const int r = r_first + q * r_step;
      const int row = r * nx, p = row + i;
      const double2 cc = tile[p];
      const double uW = tile[row + iw].x, uE = tile[row + ie].x;
      const double uS = (r == 0) ? __ldcg(xs + i) : tile[p - nx].x;
      const double uN = (r == rows - 1) ? __ldcg(xn + i) : tile[p + nx].x;
      double du = EXACT ? stencil_exact<MODEL>(k, t1, t3, cc.x, uW, uE, uS, uN) : stencil_fast<MODEL>(k, t1, t3, cc.x, uW, uE, uS, uN);
      double dv = 0.0;
      if (c.react) {
        const bool frozen = (c.freeze_north && r == rows - 1) || (c.freeze_south && r == 0);
        if (frozen) { du = 0.0; dv = 0.0; }
        else react<MODEL, EXACT>(k, c.brow[r], cc.x, cc.y, du, dv);
      }
      if (!FIN) {
        out[p] = make_double2(du, dv);
      } else {
        const double2 y0 = fin->yn[p];
        double sx = y0.x, sy = y0.y, ex = 0.0, ey = 0.0;
#pragma unroll
        for (int j = 0; j < kResStages; ++j) {
          if (j < S) {
            const double2 fj = (j == S - 1) ? make_double2(du, dv) : fin->F[j][p];
            const bool nz = fin->hb[j] != 0.0;
            sx = fin_sol_term<EXACT>(fin->hb[j], fj.x, sx, nz); ex = fin_err_term<EXACT>(fin->hd[j], fj.x, ex);
            sy = fin_sol_term<EXACT>(fin->hb[j], fj.y, sy, nz); ey = fin_err_term<EXACT>(fin->hd[j], fj.y, ey);
          }
        }
        fin->ynew[p] = make_double2(sx, sy);
        finish_tail<EXACT>(fin->rtol, fin->atol, y0.x, sx, ex, facc, want_y2);
        finish_tail<EXACT>(fin->rtol, fin->atol, y0.y, sy, ey, facc, want_y2);
      }
    }
  }
}

// PID controller + bounds: the arithmetic of adapt_eta() in crd_ark.cpp.  Called by a whole warp: the three pow() run
// on lanes 0..2 at the same time (one pow latency instead of three); every lane returns the same value.
__device__ double res_adapt_eta(const ResArgs &P, double hcur, double dsm, double eh0, double eh1, double etamax, int lane) {
  const double k = (double)P.p;
  const double base = lane == 0 ? fmax(P.bias * dsm, 1.0e-10) : lane == 1 ? fmax(eh0, 1.0e-10) : fmax(eh1, 1.0e-10);
  const double expo = lane == 0 ? -P.k1 / k : lane == 1 ? P.k2 / k : -P.k3 / k;
  const double pw = crd_pow_pos(base, expo);   // the host controller's own sequence of operations (crd_pow.h): same bits
  const double p1 = __shfl_sync(0xffffffffu, pw, 0), p2 = __shfl_sync(0xffffffffu, pw, 1), p3 = __shfl_sync(0xffffffffu, pw, 2);
  double h_acc = hcur * p1 * p2 * p3;
  const double int_dir = hcur / fabs(hcur);
  h_acc *= P.safety;
  h_acc = int_dir * fmin(fabs(h_acc), fabs(etamax * hcur));
  h_acc = int_dir * fmax(fabs(h_acc), fabs(P.etamin * hcur));
  if (fabs(h_acc) > fabs(hcur * P.lbound * 0.999999) && fabs(h_acc) < fabs(hcur * P.ubound * 1.000001)) h_acc = hcur;
  return h_acc / hcur;
}

template <int MODEL, bool EXACT, int NT, bool TICKS>
__global__ void __launch_bounds__(NT, 1) erk_resident_kernel(const ResArgs P) {
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) unsigned char smem_dyn[];   // [stage tile][storage 0] .. [storage nslots-1]
  __shared__ StageTab T;
  __shared__ RhsConst sk;
  __shared__ ResFinish sf;
  __shared__ ResLoop L;
  __shared__ double2 *S[ST_N];    // band-local base of each storage (generic address: shared or global)
  __shared__ double s_red[3][NW];
  __shared__ int s_ok;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = gridDim.x, b = blockIdx.x;
  const long long nx = P.nx;
  const long long j0 = P.nyl * b / nb, j1 = P.nyl * (b + 1) / nb;   // this band's rows
  unsigned long long bar_target = 0;
  double2 *tile = reinterpret_cast<double2 *>(smem_dyn);

  // after the band-sized arrays: metric table and beta rows of the band (read every pass: keep them out of L2's way)
  double2 *cth_s = reinterpret_cast<double2 *>(smem_dyn + (size_t)(P.nslots + 1) * P.slot_bytes);
  double *brow_s = reinterpret_cast<double *>(cth_s + nx);

  PassCfg cfg;
  cfg.nx = (int)nx; cfg.rows = (int)(j1 - j0); cfg.n = cfg.rows * cfg.nx;
  cfg.ncol = cfg.nx < NT ? cfg.nx : NT; cfg.G = NT / cfg.ncol;
  cfg.ty = (int)threadIdx.x / cfg.ncol; cfg.tx = (int)threadIdx.x - cfg.ty * cfg.ncol;
  cfg.cth = cth_s; cfg.brow = brow_s; cfg.react = P.react; cfg.freeze_south = 0; cfg.freeze_north = 0;
  const int n_edge = cfg.rows > 1 ? 2 : 1;   // the band's first and last row: need the neighbours' exchange rows

  // exchange rows: [parity][band][first | last][nx]
  auto xch_mine = [&](unsigned par) { return P.xch + ((size_t)(par & 1u) * nb + b) * 2 * nx; };
  auto halo_south = [&](unsigned par) { return P.xch + ((size_t)(par & 1u) * nb + (b == 0 ? nb - 1 : b - 1)) * 2 * nx + nx; };
  auto halo_north = [&](unsigned par) { return P.xch + ((size_t)(par & 1u) * nb + (b == nb - 1 ? 0 : b + 1)) * 2 * nx; };
  // the stage-is state yn + h sum_j A[is][j] F_j (every thread builds its own copy: no serial section per pass)
  auto stage_comb = [&](int is, double h) {
    Comb cb;
    cb.n = T.n[is];
#pragma unroll
    for (int j = 0; j < kMaxLc; ++j) {
      if (j < cb.n) {
        const int code = T.st[is][j];
        cb.x[j] = S[code == -1 ? ST_Y0 + L.yi : code == -2 ? ST_FA + L.fi : code];
        cb.c[j] = (j == 0) ? 1.0 : __dmul_rn(h, T.A[is][j]);
      } else { cb.x[j] = nullptr; cb.c[j] = 0.0; }
    }
    return cb;
  };
  auto identity_comb = [&](int st) {   // 1.0 * y is y: a plain vector through the same code
    Comb cb;
    cb.n = 1; cb.x[0] = S[st]; cb.c[0] = 1.0;
#pragma unroll
    for (int j = 1; j < kMaxLc; ++j) { cb.x[j] = nullptr; cb.c[j] = 0.0; }
    return cb;
  };
  auto set_freeze = [&](double t) {
    const int tb = (t < P.t_boundary) && P.freeze_rows;
    cfg.freeze_south = tb && (j0 == 0); cfg.freeze_north = tb && (j1 == P.nyl);
  };
  // tests before a step (same tests, same order as the host loop in crd_ark.cpp); thread 0
  auto step_top = [&]() {
    L.status = 0; L.nef = 0;
    if (P.max_steps > 0 && L.nst >= P.max_steps) { L.flag = ARK_TOO_MUCH_WORK; L.status = 2; }
    else if (L.ynorm_sq >= 0.0 && DBL_EPSILON * sqrt(L.ynorm_sq / P.nglobal) > 1.0) { L.flag = ARK_TOO_MUCH_ACC; L.status = 2; }
    L.h = L.next_h;
  };
  // phase 1 of a pass, then this CTA's arrival at the pass's barrier
  // where the time goes, as seen by thread 0 (reported for CTA 0): cycles since the previous tick are booked on `k`
  long long cyc[5] = {0, 0, 0, 0, 0}, t_last = clock64();
  const long long t_begin = t_last;
  auto tick = [&](int k) {   // TICKS = false (the default kernel): nothing — the counters cost 14 registers and 3 % of a step
    if (TICKS && threadIdx.x == 0) { const long long now = clock64(); cyc[k] += now - t_last; t_last = now; }
  };
  if (threadIdx.x == 0) {
    L.tn = P.tn; L.next_h = P.next_h; L.h = P.next_h; L.hold = P.hold; L.eta = P.eta; L.etamax = P.etamax;
    L.eh0 = P.eh0; L.eh1 = P.eh1; L.ynorm_sq = P.ynorm_sq; L.h_failed = 0.0;
    L.nst = 0; L.attempts = 0; L.nfe = 0; L.netf = 0;
    L.yi = 0; L.fi = 0;
    L.nef = 0; L.status = 0; L.flag = ARK_SUCCESS; L.stop = 0; L.pass = 0;
    for (int st = 0; st < ST_N; ++st)
      S[st] = (st < P.nslots) ? reinterpret_cast<double2 *>(smem_dyn + (size_t)(st + 1) * P.slot_bytes)
                              : reinterpret_cast<double2 *>(P.buf[storage_home(st)]) + j0 * nx;
    for (int is = 0; is < kResStages; ++is) {
      int n = 0;
      T.st[is][n] = -1; T.A[is][n] = 0.0; ++n;
      for (int j = 0; j < is && is < P.s; ++j)
        if (P.A[is][j] != 0.0) { T.st[is][n] = (j == 0) ? -2 : ST_F1 + j - 1; T.A[is][n] = P.A[is][j]; ++n; }
      T.n[is] = n;
      for (int j = n; j < kMaxLc; ++j) { T.st[is][j] = -1; T.A[is][j] = 0.0; }
    }
    sk = P.k;
    sf.rtol = P.rtol; sf.atol = P.atol; sf.s = P.s;
    step_top();
  }
  for (int i = threadIdx.x; i < cfg.nx; i += NT) cth_s[i] = reinterpret_cast<const double2 *>(P.cth)[i];
  for (int r = threadIdx.x; r < cfg.rows; r += NT) brow_s[r] = P.brow[j0 + r];
  __syncthreads();
  // the band's share of yn and fnew moves into shared memory (storages that stay global are already in place)
  for (int q = 0; q < 2; ++q) {
    const int st = q == 0 ? (int)ST_Y0 : (int)ST_FA;
    if (st < P.nslots) {
      const double2 *src = reinterpret_cast<const double2 *>(P.buf[storage_home(st)]) + j0 * nx;
      for (int e = threadIdx.x; e < cfg.n; e += NT) S[st][e] = src[e];
    }
  }
  __syncthreads();

  // ---- the passes: stage 1 .. s-1 of an attempt, then fnew = f(tn + h, ynew) ------------------------------------------
  // One code path for every pass (a single inlined copy of each phase keeps the kernel inside the instruction cache):
  //   phase 1 of the pass's state, arrive | interior rows | wait | [after the last stage: error test] | edge rows
  bool alive = true;
  int is = 1;                       // the pass: stage `is` for is < s, is == s: fnew of the state the last stage produced
  FinAcc<EXACT> facc;
  while (L.status != 2) {
    const int s_ = P.s;
    const bool fnew_pass = (is == s_), last = (is == s_ - 1);
    const double h = L.h;
    tick(4);
    // phase 1: this pass's state into the tile, the band's edge rows into the exchange buffer; then arrive
    phase1<EXACT, 4>(fnew_pass ? identity_comb(ST_Y0 + (L.yi ^ 1)) : stage_comb(is, h), cfg, tile, xch_mine(L.pass));
    if (last && threadIdx.x == 0) {
      sf.yn = S[ST_Y0 + L.yi]; sf.ynew = S[ST_Y0 + (L.yi ^ 1)];
      for (int j = 0; j < s_; ++j) {
        sf.F[j] = (j == 0) ? S[ST_FA + L.fi] : S[ST_F1 + (j < s_ - 1 ? j : 1) - 1];
        sf.hb[j] = __dmul_rn(h, P.b[j]); sf.hd[j] = __dmul_rn(h, P.d[j]);
      }
    }
    __syncthreads();                       // the tile, the CTA's exchange rows and (after a last stage) its error partials are written
    if (threadIdx.x == 0) grid_arrive(P.bar, bar_target);
    tick(0);
    const double *xs = halo_south(L.pass), *xn = halo_north(L.pass);
    double2 *out = fnew_pass ? S[ST_FA + (L.fi ^ 1)] : last ? nullptr : S[ST_F1 + is - 1];
    if (!fnew_pass) set_freeze(__dadd_rn(L.tn, __dmul_rn(P.c[is], h)));
    bool rejected = false;
    for (int part = 0; part < 2; ++part) {
      int r_first, r_step, nr;
      if (part == 0) {
        // interior rows need nothing from the neighbours: the barrier's latency hides behind them (the fnew pass has
        // to know the verdict first)
        r_first = 1; r_step = 1; nr = fnew_pass ? 0 : cfg.rows - 2;
      } else {
        if (!grid_wait(P.bar, bar_target, &s_ok)) { alive = false; break; }
        tick(2);
        if (fnew_pass) {
          // ---- error norm: every CTA adds all partials in the same order, then the same test and controller ----
          if (warp == 0) {
            double se = 0.0, sl = 0.0, sy = 0.0;   // the error sum as a double-double pair, merged in a fixed order, rounded once
            for (int q = lane; q < nb; q += 32) { dd_merge(se, sl, __ldcg(P.partial + q), __ldcg(P.partial + 2 * nb + q)); sy += __ldcg(P.partial + nb + q); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { dd_shfl_down(se, sl, o); sy += __shfl_down_sync(0xffffffffu, sy, o); }
            se = __shfl_sync(0xffffffffu, __dadd_rn(se, sl), 0);
            const double dsm = sqrt(se / P.nglobal);
            const double eta_a = res_adapt_eta(P, L.h, dsm, L.eh0, L.eh1, L.etamax, lane);
            const double eta_r = res_adapt_eta(P, L.h, dsm, L.eh0, L.eh1, 1.0, lane);   // after a failure etamax is 1
            if (lane == 0) {
              L.eta = eta_a;
              if (dsm <= 1.0) {
                L.eh1 = L.eh0; L.eh0 = dsm * P.bias;
                L.ynorm_sq = P.rtol > 2.220446049250313e-16 ? P.nglobal / (P.rtol * P.rtol) : sy;   // the bound when the sum is skipped
                L.status = 1;
                // complete the step: yn <- ynew (ping-pong); fnew goes into the other F storage
                L.yi ^= 1;
                L.hold = L.h;
                L.tn += L.h;
                L.nst++;
                L.etamax = P.growth;
                L.next_h = L.h * L.eta;
                L.stop = (P.itask == ARK_ONE_STEP) || ((L.tn - P.tout) * L.h >= 0.0);
              } else {
                L.status = 0;
                L.nef++; L.netf++;
                L.etamax = 1.0;
                if (L.nef == P.maxnef) { L.flag = ARK_ERR_FAILURE; L.h_failed = L.h; L.status = 2; }
                else {
                  double eta = fmin(eta_r, 1.0);
                  if (L.nef >= P.small_nef) eta = fmin(eta, P.etamxf);
                  L.eta = eta;
                  L.h *= eta;
                  if (fabs(L.h) <= 0.0 || L.tn + L.h == L.tn) { L.flag = ARK_ERR_FAILURE; L.h_failed = L.h; L.status = 2; }
                }
              }
            }
          }
          __syncthreads();
          if (L.status != 1) { rejected = true; break; }
          set_freeze(L.tn);
          r_first = 0; r_step = 1; nr = cfg.rows;
        } else {
          r_first = 0; r_step = cfg.rows - 1; nr = n_edge;
        }
      }
      if (last) phase2_rows<MODEL, EXACT, true>(cfg, sk, tile, xs, xn, out, &sf, r_first, r_step, nr, facc);
      else phase2_rows<MODEL, EXACT, false>(cfg, sk, tile, xs, xn, out, nullptr, r_first, r_step, nr, facc);
      tick(part == 0 ? 1 : 3);
    }
    if (!alive) break;
    if (rejected) {
      // the tile holds a ynew nobody wants: retry from stage 1 with the smaller step (or stop: status 2)
      is = 1;
      continue;
    }
    if (last) {
      // CTA partial sums, fixed order: shuffle tree, then the warps in order
      double eh = facc.e_hi, el = facc.e_lo, y2 = facc.y2;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { dd_shfl_down(eh, el, o); y2 += __shfl_down_sync(0xffffffffu, y2, o); }
      if (lane == 0) { s_red[0][warp] = eh; s_red[1][warp] = y2; s_red[2][warp] = el; }
      facc = FinAcc<EXACT>();
    }
    __syncthreads();          // every read of the tile is done; the band's F_is / ynew / fnew is complete
    if (threadIdx.x == 0) {
      if (last) {
        double se = s_red[0][0], sy = s_red[1][0], sl = s_red[2][0];
        for (int w = 1; w < NW; ++w) { dd_merge(se, sl, s_red[0][w], s_red[2][w]); sy += s_red[1][w]; }
        P.partial[b] = se;
        P.partial[nb + b] = sy;
        P.partial[2 * nb + b] = sl;
      }
      if (is == 1) L.attempts++;
      L.nfe++;
      L.pass++;
      if (fnew_pass) {
        L.fi ^= 1;
        if (L.stop) L.status = 2;
        else step_top();
      }
    }
    __syncthreads();
    is = fnew_pass ? 1 : is + 1;
  }

  // ---- back to the global homes: yn, fnew and (after at least one step) the previous step's state for dense output ----
  __syncthreads();
  {
    const int sts[4] = {ST_Y0 + L.yi, ST_FA + L.fi, ST_Y0 + (L.yi ^ 1), ST_FA + (L.fi ^ 1)};
    for (int q = 0; q < 4; ++q) {
      const int st = sts[q];
      if (st >= P.nslots) continue;            // lives in its home already
      if (q >= 2 && L.nst == 0) continue;      // nothing accepted: the homes still hold the caller's yold / fold
      double2 *dst = reinterpret_cast<double2 *>(P.buf[storage_home(st)]) + j0 * nx;
      for (int e = threadIdx.x; e < cfg.n; e += NT) dst[e] = S[st][e];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ResOut *o = P.out;
    o->tn = L.tn; o->next_h = L.next_h; o->hold = L.hold; o->eta = L.eta; o->etamax = L.etamax;
    o->eh0 = L.eh0; o->eh1 = L.eh1; o->ynorm_sq = L.ynorm_sq; o->h_failed = L.h_failed;
    o->nst = L.nst; o->attempts = L.attempts; o->nfe = L.nfe; o->netf = L.netf;
    o->flag = L.flag; o->err = alive ? 0 : 1;
    o->yi = L.yi; o->fi = L.fi;
    tick(4);
    for (int q = 0; q < 5; ++q) o->cyc[q] = cyc[q];
    o->cyc[5] = clock64() - t_begin;
    __threadfence_system();
    o->done = 1;
    __threadfence_system();
  }
}

struct ResKernel {
  const void *fn;
  int threads;
  int static_smem[64];   // per device: 0 = not asked yet, -1 = cannot run here, else static shared bytes + 1
};

template <int MODEL, bool EXACT, int NT, bool TICKS>
ResKernel *res_kernel_entry() {
  static ResKernel k = {(const void *)erk_resident_kernel<MODEL, EXACT, NT, TICKS>, NT, {}};
  return &k;
}

// 512 threads per CTA, 128 registers per thread.  Measured on the 400 x 1600 mesh: 1024 threads x 64 registers shorten
// the stencil phase by 20 % but spill, lengthen every other phase and lose 5 % per step.
// ticks: the instantiation with per-phase cycle counters (crd_grid_resident_cycles), selected by grid variant 150
template <int MODEL, bool EXACT>
ResKernel *res_pick_nt(int ticks) {
#ifdef CRD_PROFILING_VARIANTS
  if (ticks) return res_kernel_entry<MODEL, EXACT, 512, true>();
#endif
  (void)ticks;
  return res_kernel_entry<MODEL, EXACT, 512, false>();
}

ResKernel *res_pick(int model, bool exact, int ticks) {
  switch (model) {
    case CRD_FHN_TORUS: return exact ? res_pick_nt<CRD_FHN_TORUS, true>(ticks) : res_pick_nt<CRD_FHN_TORUS, false>(ticks);
    case CRD_GOLDBETER_TORUS: return exact ? res_pick_nt<CRD_GOLDBETER_TORUS, true>(ticks) : res_pick_nt<CRD_GOLDBETER_TORUS, false>(ticks);
    case CRD_FHN_FLAT: return exact ? res_pick_nt<CRD_FHN_FLAT, true>(ticks) : res_pick_nt<CRD_FHN_FLAT, false>(ticks);
    case CRD_GOLDBETER_FLAT: return exact ? res_pick_nt<CRD_GOLDBETER_FLAT, true>(ticks) : res_pick_nt<CRD_GOLDBETER_FLAT, false>(ticks);
  }
  return nullptr;
}

// meshes up to this many points run the resident loop by default: its working set (8 storages of 16 B per point) then lives in
// shared memory and the 126 MB L2; beyond it every pass between two grid barriers streams from HBM, which the launch-per-stage
// kernels do faster (500 x 2000: 80 vs 122 us per step resident / launch-per-stage; 550 x 2200: 169 vs 117; profiles/README.md)
constexpr long long kResidentAutoPoints = 1LL << 20;

}  // namespace

extern "C" {

int crd_grid_set_resident(crd_grid *g, int mode) {
  if (!g || mode < -1 || mode > 1) { set_error("crd_grid_set_resident: bad arguments"); return -1; }
  g->resident = mode;
  return 0;
}

int64_t crd_grid_resident_launches(const crd_grid *g) { return g ? g->resident_launches : 0; }

int crd_grid_resident_cycles(const crd_grid *g, int64_t out[6]) {
  if (!g || !out) return -1;
  for (int q = 0; q < 6; ++q) out[q] = g->res_cycles[q];
  return 0;
}

int crd_erk_evolve(struct crd_erk_state *st, void *user_data) {
  crd_grid *g = (crd_grid *)user_data;
  if (!st || !g) { set_error("crd_erk_evolve: null argument"); return -1; }
  crd_ctx *ctx = g->ctx;
  // applicability: one slab that wraps onto itself, a method whose widest stage fits one combination
  if (g->resident < 0) return 1;
  if (g->connected || ctx->nranks > 1) return 1;
  if (g->resident == 0 && g->nx * g->nyl > kResidentAutoPoints) return 1;
  if (g->nx * g->nyl >= (1LL << 31)) return 1;
  if (st->s < 2 || st->s > kResStages) return 1;
  for (int i = 1; i < st->s; ++i) {
    int n = 1;
    for (int j = 0; j < i; ++j) n += st->A[i][j] != 0.0;
    if (n > kMaxLc) return 1;
  }
  if (st->itask != ARK_NORMAL && st->itask != ARK_ONE_STEP) return 1;
  if (!(st->next_h != 0.0) || !std::isfinite(st->next_h)) return 1;
  if (use(ctx)) return -1;

  const long long len = crd_grid_local_length(g);
  N_Vector in[R_NBUF] = {st->yn, st->yold, st->ycur, st->fnew, st->fold};
  for (int j = 1; j < kResStages; ++j) in[R_F1 + j - 1] = st->F[j];
  ResArgs P;
  std::memset(&P, 0, sizeof P);
  for (int b = 0; b < R_F1 + st->s - 1; ++b) {
    if (!in[b] || N_VGetLocalLength_Crd(in[b]) != len) { set_error("crd_erk_evolve: vector does not match the grid"); return -1; }
    P.buf[b] = N_VGetDeviceArrayPointer_Crd(in[b]);
    if (!P.buf[b] || ((uintptr_t)P.buf[b] & 15)) { set_error("crd_erk_evolve: vectors must be 16-byte aligned device arrays"); return -1; }
  }
  for (int b = R_F1 + st->s - 1; b < R_NBUF; ++b) P.buf[b] = P.buf[R_F1];   // storages a shorter method never touches
  P.nx = g->nx; P.nyl = g->nyl; P.cth = g->cth; P.brow = g->brow; P.k = g->k;
  P.react = (is_fhn(g->p.model) || g->p.just_diffusion == 0) ? 1 : 0;
  P.freeze_rows = (g->js == 0 && g->je == g->ny - 1) ? 1 : 0;
  P.t_boundary = g->p.t_boundary;
  P.s = st->s; P.p = st->p;
  for (int i = 0; i < st->s; ++i) {
    for (int j = 0; j < st->s; ++j) P.A[i][j] = st->A[i][j];
    P.b[i] = st->b[i]; P.d[i] = st->d[i]; P.c[i] = st->c[i];
  }
  P.rtol = st->rtol; P.atol = st->atol;
  P.k1 = st->k1; P.k2 = st->k2; P.k3 = st->k3; P.bias = st->bias; P.safety = st->safety; P.growth = st->growth;
  P.etamxf = st->etamxf; P.etamin = st->etamin; P.lbound = st->lbound; P.ubound = st->ubound;
  P.small_nef = st->small_nef; P.maxnef = st->maxnef;
  P.nglobal = (double)st->nglobal;
  P.tout = st->tout; P.itask = st->itask; P.max_steps = st->max_steps;
  P.tn = st->tn; P.next_h = st->next_h; P.hold = st->hold; P.eta = st->eta; P.etamax = st->etamax;
  P.eh0 = st->ehist[0]; P.eh1 = st->ehist[1]; P.ynorm_sq = st->ynorm_sq;

  ResKernel *K = res_pick(g->p.model, g->p.arith == CRD_ARITH_EXACT, g->variant == 150 ? 1 : 0);
  if (!K) { set_error("crd_erk_evolve: unknown model"); return -1; }
  const int dev = ctx->device & 63;
  int sms = 0, smem_optin = 0;
  CRD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
  CRD_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  if (K->static_smem[dev] == 0) {
    int coop = 0;
    CRD_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
    cudaFuncAttributes fa;
    CRD_CUDA(cudaFuncGetAttributes(&fa, K->fn));
    K->static_smem[dev] = coop ? (int)fa.sharedSizeBytes + 1 : -1;
    if (coop) CRD_CUDA(cudaFuncSetAttribute(K->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - (int)fa.sharedSizeBytes));
  }
  if (K->static_smem[dev] < 0) return 1;
  // one band of phi rows per SM; as many storages in shared memory as fit (most used first)
  const long long nb = g->nyl < sms ? g->nyl : sms;
  const long long max_rows = (g->nyl + nb - 1) / nb;
  const size_t slot_bytes = (size_t)max_rows * (size_t)g->nx * 16;   // one band-sized array
  const size_t tables = (size_t)g->nx * 16 + (size_t)max_rows * 8 + 16;              // metric table + beta rows of a band
  const size_t reserve = (size_t)(K->static_smem[dev] - 1) + 1024 + tables;              // 1 KB kept for the runtime
  if ((size_t)smem_optin < reserve + slot_bytes) return 1;   // not even the stage tile of a band fits: the launch-per-stage path streams this mesh
  int nslots = (int)(((size_t)smem_optin - reserve) / slot_bytes) - 1;
  if (nslots > ST_N) nslots = ST_N;

  if (g->variant >= 120 && g->variant <= 120 + ST_N && nslots > g->variant - 120) nslots = g->variant - 120;   // 120 + n: at most n storages in shared memory
  P.nslots = nslots;
  P.slot_bytes = (unsigned)slot_bytes;
  const size_t dyn_smem = (size_t)(nslots + 1) * slot_bytes + tables;
  int fit = 0;
  CRD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, K->fn, K->threads, dyn_smem));
  if (fit < 1) return 1;

  // scratch: barrier words, partials, exchange rows, result block (kept with the grid)
  if (!g->res_bar) {
    CRD_CUDA(cudaMalloc(&g->res_bar, 32 * sizeof(unsigned long long)));
    CRD_CUDA(cudaMalloc(&g->res_partial, sizeof(double) * (4 * (size_t)sms + 4 * (size_t)sms * (size_t)g->nx)));
    CRD_CUDA(cudaHostAlloc(&g->res_out_host, sizeof(ResOut), cudaHostAllocMapped));
    CRD_CUDA(cudaHostGetDevicePointer(&g->res_out_dev, g->res_out_host, 0));
  }
  P.bar = g->res_bar; P.partial = g->res_partial; P.xch = g->res_partial + 4 * (size_t)sms; P.out = (ResOut *)g->res_out_dev;
  ResOut *out = (ResOut *)g->res_out_host;
  std::memset(out, 0, sizeof *out);
  CRD_CUDA(cudaMemsetAsync(g->res_bar, 0, 32 * sizeof(unsigned long long), ctx->stream));
  void *kargs[] = {(void *)&P};
  cudaError_t e = cudaLaunchCooperativeKernel(K->fn, dim3((unsigned)nb), dim3(K->threads), kargs, dyn_smem, ctx->stream);
  if (e != cudaSuccess) {
    // nothing has run: the integrator can carry on with its launch-per-stage loop (e.g. the CTAs cannot all be
    // co-resident because another context occupies SMs); do not ask again for this grid
    set_error("crd_erk_evolve: cooperative launch failed (%s): falling back to the host-driven loop", cudaGetErrorString(e));
    cudaGetLastError();
    g->resident = -1;
    return 1;
  }
  ctx->launches++;
  g->resident_launches++;
  if (sync_stream(ctx, "crd_erk_evolve")) return -1;
  if (!out->done || out->err) { set_error("crd_erk_evolve: the device step loop did not complete (grid barrier timed out)"); return -1; }

  st->tn = out->tn; st->next_h = out->next_h; st->hold = out->hold; st->eta = out->eta; st->etamax = out->etamax;
  st->ehist[0] = out->eh0; st->ehist[1] = out->eh1; st->ynorm_sq = out->ynorm_sq; st->h_failed = out->h_failed;
  st->nst += (long)out->nst; st->nst_attempts += (long)out->attempts; st->nfe += (long)out->nfe; st->netf += (long)out->netf;
  g->rhs_count += out->nfe;
  for (int q = 0; q < 6; ++q) g->res_cycles[q] = out->cyc[q];
  if (out->nst > 0) {
    // the loop ping-pongs between (yn, ycur) and (fnew, fold); the caller's yold array is free from the first accepted step on
    st->yn = in[storage_home(ST_Y0 + out->yi)]; st->yold = in[storage_home(ST_Y0 + (out->yi ^ 1))]; st->ycur = in[R_YOLD];
    st->fnew = in[storage_home(ST_FA + out->fi)]; st->fold = in[storage_home(ST_FA + (out->fi ^ 1))];
  }
  st->flag = out->flag;
  return 0;
}

}  // extern "C"

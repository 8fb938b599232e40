// crd_rhs_pair.cuh — two evaluations in ONE pass over the state:  f1 = f(t1, y),  f2 = f(t2, y + c*f1).
//
// What it replaces in the explicit RK loop: when a step has been accepted the integrator evaluates f(tn, ynew) (dense output;
// stage 1 of the next step) and, right at the start of the next step, stage 2 = f(tn + c2 h, ynew + h a21 F1) — whose state needs
// nothing but ynew and F1.  As two launches that is 32 + 48 B per point (ynew is read twice, F1 written and read back); here ynew
// is read once, F1 and F2 are written, and the stage state never exists in memory: 48 B per point.
//
// Same organisation as rhs_stream_kernel (persistent CTAs, a TMA producer warp feeding a shared-memory ring of row stages,
// consumer threads that own a column and walk the rows with the previous rows in registers), one row deeper: as row r of y
// arrives a thread forms F1 of row r-1, from it the stage state z of row r-1, and with the z of rows r-3 .. r-1 in registers F2
// of row r-2.  East / west neighbours of z come from the neighbouring lanes; to keep warps independent a warp covers 32 columns
// and writes the inner 30 (lanes 0 and 31 only supply F1 / z of the columns next to them), so a strip of 8 warps is 240
// columns wide and its ring rows carry two extra columns of y on either side.  A segment of rows fetches two extra rows of y
// above and below.  The per-point arithmetic is the shared device code of crd_rhs_point.cuh / crd_fused.cuh: F1, z and F2 have
// the bits of crd_rhs followed by crd_rhs_lincomb.
//
// Applies to a slab that wraps onto itself (one rank); a phi-split grid keeps the two separate evaluations.
//
// Measured at 16384 x 16384 (profiles/README.md): FAST arithmetic 2.41 ms against 3.35-3.48 ms for the two launches (x1.4: the
// pass is bound by memory again); EXACT arithmetic 3.2-3.5 ms against 3.4-3.5 ms — two evaluations' worth of separately rounded
// operations, exact constant divisions and range checks make one pass issue-bound, so nothing is gained and the integrator
// (crd_f_pair) keeps the two launches on EXACT grids; the kernel is there and tested bit for bit all the same (crd_rhs_pair).
#pragma once
#include "crd_rhs_kernels.cuh"

namespace {

struct PairArgs {
  const double *y;        // [nyl][nx][2]
  double *f1, *f2;
  const double *cth;      // per-theta metric coefficients (a1, a3)
  const double *brow;     // per-phi beta
  long long nx, nyl;
  RhsConst k;
  double c[2];            // (1, h a21): z = c[0] y + c[1] f1 in the arithmetic of the fused stage assembly
  int react;
  int frz1, frz2;         // first / second evaluation: bit 0 = row 0, bit 1 = row nyl-1 held at zero (t < tBoundary)
};

constexpr int kPairCols = 240;            // columns a strip writes (8 warps x 30)
constexpr int kPairPitch = kPairCols + 4; // points per ring row: two columns of y on either side
constexpr int kPairRB = 4;                // rows per ring stage

template <int MODEL, bool EXACT, int MINB>
__global__ void __launch_bounds__(288, MINB) rhs_pair_kernel(const PairArgs a, int seg_rows, int S) {
  constexpr int PITCH = kPairPitch, RB = kPairRB;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2 *ring = reinterpret_cast<double2 *>(smem_raw);                    // [S][RB][PITCH]
  const unsigned bars = smem_u32(smem_raw + (size_t)S * RB * PITCH * 16);   // full[S], empty[S]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nx = a.nx, nyl = a.nyl;
  const long long strips = (nx + kPairCols - 1) / kPairCols, segs = (nyl + seg_rows - 1) / seg_rows, units = strips * segs;

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
    long long it = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const long long seg = u / strips, strip = u - seg * strips;
      const long long i0 = strip * kPairCols, jA = seg * seg_rows, jB = (jA + seg_rows < nyl) ? jA + seg_rows : nyl;
      const int w = (nx - i0 < kPairCols) ? (int)(nx - i0) : kPairCols;
      // ring row = columns i0-2 .. i0+w+1 (periodic): a main piece inside [0, nx) and up to two wrapped pieces
      const long long m0 = (i0 >= 2) ? i0 - 2 : 0, m1 = (i0 + w + 2 <= nx) ? i0 + w + 2 : nx;
      const int lwrap = (int)(m0 - (i0 - 2)), rwrap = (int)(i0 + w + 2 - m1);   // 0 or 2 | 0, 1 or 2 columns
      const unsigned row_bytes = (unsigned)(w + 4) * 16u;
      for (long long j0 = jA - 2; j0 <= jB + 1; j0 += RB, ++it) {   // a stage = rows j0 .. j0+RB-1 (clipped at jB+1)
        const int s = (int)(it % S);
        mbar_wait(bars + 8u * (S + s), (unsigned)((it / S) & 1) ^ 1u);
        const int nr = (jB + 1 - j0 + 1 < RB) ? (int)(jB + 1 - j0 + 1) : RB;
        const unsigned full = bars + 8u * s;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(row_bytes * (unsigned)nr) : "memory");
        for (int rr = 0; rr < nr; ++rr) {
          long long jr = j0 + rr;
          jr = jr < 0 ? jr + nyl : jr >= nyl ? jr - nyl : jr;
          const double2 *row = reinterpret_cast<const double2 *>(a.y) + jr * nx;
          const unsigned dst = smem_u32(ring + ((size_t)s * RB + rr) * PITCH);
          bulk_g2s(dst + (unsigned)lwrap * 16u, row + m0, (unsigned)(m1 - m0) * 16u, full);
          if (lwrap) bulk_g2s(dst, row + (nx - lwrap), (unsigned)lwrap * 16u, full);
          if (rwrap) bulk_g2s(dst + (unsigned)(w + 4 - rwrap) * 16u, row, (unsigned)rwrap * 16u, full);
        }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  const int q = 30 * warp + lane;        // ring index of the column west of this thread's; its own is q + 1, east q + 2
  const bool inner = (lane >= 1) && (lane <= 30);
  const int react_on = a.react;
  const int nyli = (int)nyl;
  int slot_i = 0;
  unsigned slot_par = 0;
  const double cz[2] = {a.c[0], a.c[1]};
  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    const long long seg = u / strips, strip = u - seg * strips;
    const int i0 = (int)(strip * kPairCols), jA = (int)(seg * seg_rows), jB = (jA + seg_rows < nyli) ? jA + seg_rows : nyli;
    const int w = (nx - i0 < kPairCols) ? (int)(nx - i0) : kPairCols;
    const int col = i0 + q - 1;          // global column (may be -1 or nx: a periodic neighbour, never written)
    const bool writes = inner && (q - 1 < w);
    double t1 = 0.0, t3 = 0.0;
    if (is_torus(MODEL) && q - 1 < w + 1) {
      const long long ci = col < 0 ? col + nx : col >= nx ? col - nx : col;
      const double2 tc = reinterpret_cast<const double2 *>(a.cth)[ci];
      t1 = tc.x; t3 = tc.y;
    }
    double2 *o1 = reinterpret_cast<double2 *>(a.f1) + (long long)jA * nx + col;
    double2 *o2 = reinterpret_cast<double2 *>(a.f2) + (long long)jA * nx + col;
    // rows in registers: y of the row before the arriving one (centre, west / east u) and the u of the row before that;
    // z of the three rows before that: zn (latest, u only), zc (centre + west / east u), zS (u only).  F2 lags F1 by two rows, so
    // that the two evaluations of an iteration do not depend on each other and interleave.
    double2 yc = make_double2(0.0, 0.0), zc = make_double2(0.0, 0.0);
    double yW = 0.0, yE = 0.0, yS = 0.0, zW = 0.0, zE = 0.0, zS = 0.0, zN = 0.0;
    // F2 of row j2 from the z of rows j2-1 (zS), j2 (zc, zW, zE) and j2+1 (zN)
    auto second = [&](int j2) {
      double du = EXACT ? stencil_exact<MODEL>(a.k, t1, t3, zc.x, zW, zE, zS, zN) : stencil_fast<MODEL>(a.k, t1, t3, zc.x, zW, zE, zS, zN);
      double dv = 0.0;
      if (react_on) {
        react<MODEL, EXACT>(a.k, __ldg(a.brow + j2), zc.x, zc.y, du, dv);
        const bool frozen = ((a.frz2 & 1) && j2 == 0) || ((a.frz2 & 2) && j2 == nyli - 1);
        du = frozen ? 0.0 : du;
        dv = frozen ? 0.0 : dv;
      }
      if (writes) o2[(long long)(j2 - jA) * nx] = make_double2(du, dv);
    };
    double2 zl = make_double2(0.0, 0.0);   // z of the row before the arriving one's predecessor (centre), with its west / east u
    double zlW = 0.0, zlE = 0.0;
    for (int j0 = jA - 2; j0 <= jB + 1; j0 += RB) {
      mbar_wait(bars + 8u * slot_i, slot_par);
      const double2 *stage = ring + (size_t)slot_i * RB * PITCH;
      const int nr = (jB + 1 - j0 + 1 < RB) ? (jB + 1 - j0 + 1) : RB;
      if (j0 >= jA + 4 && j0 + RB <= jB) {
        // steady state (the kernel is bound by issue slots: profiles/README.md): a full stage whose F1 rows j0-1 .. j0+RB-2 and
        // F2 rows j0-3 .. j0+RB-4 all lie strictly inside the segment and the slab — every row is written, none is frozen
        double2 *p1 = o1 + (long long)(j0 - 1 - jA) * nx, *p2 = o2 + (long long)(j0 - 3 - jA) * nx;
        const double *b1 = a.brow + (j0 - 1), *b2 = a.brow + (j0 - 3);
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
          const double2 *slot = stage + (size_t)rr * PITCH;
          const double2 nn = slot[q + 1];
          const double nW = slot[q].x, nE = slot[q + 2].x;
          {
            double du = EXACT ? stencil_exact<MODEL>(a.k, t1, t3, zc.x, zW, zE, zS, zl.x) : stencil_fast<MODEL>(a.k, t1, t3, zc.x, zW, zE, zS, zl.x);
            double dv = 0.0;
            if (react_on) react<MODEL, EXACT>(a.k, __ldg(b2 + rr), zc.x, zc.y, du, dv);
            if (writes) p2[(long long)rr * nx] = make_double2(du, dv);
          }
          double du = EXACT ? stencil_exact<MODEL>(a.k, t1, t3, yc.x, yW, yE, yS, nn.x) : stencil_fast<MODEL>(a.k, t1, t3, yc.x, yW, yE, yS, nn.x);
          double dv = 0.0;
          if (react_on) react<MODEL, EXACT>(a.k, __ldg(b1 + rr), yc.x, yc.y, du, dv);
          if (writes) p1[(long long)rr * nx] = make_double2(du, dv);
          const double vx[2] = {yc.x, du}, vy[2] = {yc.y, dv};
          const double2 zn = make_double2(lc_value_n<EXACT, 2>(cz, vx), lc_value_n<EXACT, 2>(cz, vy));
          const double znW = __shfl_up_sync(0xffffffffu, zn.x, 1), znE = __shfl_down_sync(0xffffffffu, zn.x, 1);
          yS = yc.x; yc = nn; yW = nW; yE = nE;
          zS = zc.x; zc = zl; zW = zlW; zE = zlE;
          zl = zn; zlW = znW; zlE = znE;
        }
      } else
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        if (rr < nr) {
          const int jr = j0 + rr;                      // the arriving row of y
          const double2 *slot = stage + (size_t)rr * PITCH;
          const double2 nn = slot[q + 1];
          const double nW = slot[q].x, nE = slot[q + 2].x;
          // F2 of row jr-3: needs z of rows jr-4 (zS), jr-3 (zc), jr-2 (zN = zl.x), all formed in earlier iterations
          if (jr >= jA + 3) { zN = zl.x; second(jr - 3); }
          double2 zn = make_double2(0.0, 0.0);
          if (jr >= jA) {
            // F1 of row jr-1 (rows jA-1 and jB are only needed for z; a row index outside the slab is its periodic image)
            const int j1 = jr - 1;
            const int g1 = j1 < 0 ? j1 + nyli : j1 >= nyli ? j1 - nyli : j1;
            double du = EXACT ? stencil_exact<MODEL>(a.k, t1, t3, yc.x, yW, yE, yS, nn.x) : stencil_fast<MODEL>(a.k, t1, t3, yc.x, yW, yE, yS, nn.x);
            double dv = 0.0;
            if (react_on) {
              react<MODEL, EXACT>(a.k, __ldg(a.brow + g1), yc.x, yc.y, du, dv);
              const bool frozen = ((a.frz1 & 1) && g1 == 0) || ((a.frz1 & 2) && g1 == nyli - 1);
              du = frozen ? 0.0 : du;
              dv = frozen ? 0.0 : dv;
            }
            if (writes && j1 >= jA && j1 < jB) o1[(long long)(j1 - jA) * nx] = make_double2(du, dv);
            const double vx[2] = {yc.x, du}, vy[2] = {yc.y, dv};
            zn = make_double2(lc_value_n<EXACT, 2>(cz, vx), lc_value_n<EXACT, 2>(cz, vy));
          }
          const double znW = __shfl_up_sync(0xffffffffu, zn.x, 1), znE = __shfl_down_sync(0xffffffffu, zn.x, 1);
          yS = yc.x; yc = nn; yW = nW; yE = nE;
          zS = zc.x; zc = zl; zW = zlW; zE = zlE;      // (rows jr-4, jr-3 of the next iteration)
          zl = zn; zlW = znW; zlE = znE;               // row jr-1 now, jr-2 of the next iteration
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8u * (S + slot_i)) : "memory");
      if (++slot_i == S) { slot_i = 0; slot_par ^= 1u; }
    }
    zN = zl.x;          // the last row: F2 of row jB-1 from the z of rows jB-2, jB-1, jB
    second(jB - 1);
  }
}

inline int pair_seg_rows(long long nyl, long long strips, long long ctas) { return stream_seg_rows(nyl, strips, ctas); }

template <int MODEL, bool EXACT, int MINB>
int launch_pair_minb(crd_grid *g, const PairArgs &a, cudaStream_t st) {
  const long long strips = (a.nx + kPairCols - 1) / kPairCols;
  const long long ctas_max = (long long)MINB * g->ctx->sms;
  const int seg_rows = pair_seg_rows(a.nyl, strips, ctas_max);
  const long long segs = (a.nyl + seg_rows - 1) / seg_rows, units = strips * segs;
  const size_t stage_bytes = (size_t)kPairRB * kPairPitch * 16;
  int S = (int)((MINB == 3 ? 68000 : 100000) / stage_bytes);
  if (S > 8) S = 8;
  const size_t smem = (size_t)S * stage_bytes + (size_t)2 * S * 8;
  auto kern = rhs_pair_kernel<MODEL, EXACT, MINB>;
  static bool attr_set[64] = {};   // the attribute is per device
  const int dev = g->ctx->device & 63;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110000);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set[dev] = true;
  }
  const long long ctas = units < ctas_max ? units : ctas_max;
  kern<<<(unsigned)ctas, 288, smem, st>>>(a, seg_rows, S);
  return check_launch(g->ctx, "rhs_pair_kernel");
}

// 3 CTAs per SM (72 registers); 2 CTAs with 96 registers measured the same in EXACT arithmetic (profiles/README.md)
template <int MODEL, bool EXACT>
int launch_pair_model(crd_grid *g, const PairArgs &a, cudaStream_t st) { return launch_pair_minb<MODEL, EXACT, 3>(g, a, st); }

int launch_pair(crd_grid *g, const PairArgs &a, cudaStream_t st) {
  const bool exact = g->p.arith == CRD_ARITH_EXACT;
  switch (g->p.model) {
    case CRD_FHN_TORUS: return exact ? launch_pair_model<CRD_FHN_TORUS, true>(g, a, st) : launch_pair_model<CRD_FHN_TORUS, false>(g, a, st);
    case CRD_GOLDBETER_TORUS: return exact ? launch_pair_model<CRD_GOLDBETER_TORUS, true>(g, a, st) : launch_pair_model<CRD_GOLDBETER_TORUS, false>(g, a, st);
    case CRD_FHN_FLAT: return exact ? launch_pair_model<CRD_FHN_FLAT, true>(g, a, st) : launch_pair_model<CRD_FHN_FLAT, false>(g, a, st);
    case CRD_GOLDBETER_FLAT: return exact ? launch_pair_model<CRD_GOLDBETER_FLAT, true>(g, a, st) : launch_pair_model<CRD_GOLDBETER_FLAT, false>(g, a, st);
  }
  set_error("unknown model %d", g->p.model);
  return -1;
}

}  // namespace

// crd_rhs_pair.cuh — two evaluations in ONE pass over the state:  f1 = f(t1, y),  f2 = f(t2, y + c*f1).
//
// What it replaces in the explicit RK loop: when a step has been accepted the integrator evaluates f(tn, ynew) (dense output;
// stage 1 of the next step) and, right at the start of the next step, stage 2 = f(tn + c2 h, ynew + h a21 F1) — whose state needs
// nothing but ynew and F1.  As two launches that is 32 + 48 B per point (ynew is read twice, F1 written and read back); here ynew
// is read once, F1 and F2 are written, and the stage state never exists in memory: 48 B per point.
//
// One CTA per tile of 240 columns x 12 rows: the tile of y with a frame of two rows / two columns is staged in shared memory by
// bulk copies (TMA, one mbarrier), then every thread marches its column through fully unrolled rows — F1 of a row from the
// staged y, from it the stage state z of that row (registers only), and from the z of three consecutive rows F2 of the middle
// one.  East / west neighbours of z come from the neighbouring lanes; to keep warps independent a warp covers 32 columns and
// writes the inner 30 (lanes 0 and 31 only supply F1 / z of the columns next to them).  Cost of the frame: 14 evaluations of F1
// per 12 rows, 16 rows of y read per 12 (the extra rows are L2 hits: the tile below was staged ~70 tiles earlier).  The per-point
// arithmetic is the shared device code of crd_rhs_point.cuh / crd_fused.cuh: F1, z and F2 have the bits of crd_rhs followed by
// crd_rhs_lincomb.
//
// On a phi-split grid the neighbours deliver TWO rows of y per side before the pass (halo_push2_kernel; the first evaluation on
// the adjacent one of them is repeated here, so F1 needs no exchange of its own): push, wait, pass — three launches, one exchange.
//
// Measured at 16384 x 16384 (profiles/README.md): FAST arithmetic 1.89 ms = 6.8 TB/s at 48 B/point against 3.3-3.4 ms for the two
// launches (x1.75; a first, streaming form of this pass — persistent CTAs, row ring — reached 2.41 ms); EXACT arithmetic (FHN)
// 2.67 ms in a burst, 3.16 ms warm, against 3.45-3.8 ms: two evaluations' worth of separately rounded operations, exact constant
// divisions and range checks per 48 B make the pass latency-bound, not memory-bound.  Goldbeter in EXACT arithmetic is bound
// by FP64 work (2 x 69 operations per point) and stays slower than its two launches (1.2-1.3 against 1.05-1.1 ms at 4096 x
// 16384): the integrator's entry (crd_f_pair) declines there; the kernel is tested bit for bit all the same (crd_rhs_pair).
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
  int frz1, frz2;         // first / second evaluation: rows held at zero (t < tBoundary): bit 0 = row 0, bit 1 = row nyl-1,
                          // bit 2 = row -1, bit 3 = row nyl (a neighbouring rank's boundary row, phi-split grid only)
  // phi-split grid: the two rows beyond the slab on either side, delivered by the neighbours (near = adjacent row); nullptr =
  // the slab wraps onto itself
  const double *south_near, *south_far, *north_near, *north_far;
};

constexpr int kPairCols = 240;            // columns a tile writes (8 warps x 30)
constexpr int kPairPitch = kPairCols + 4; // points per staged row: two columns of y on either side
constexpr int kPairTY = 12;               // rows a FAST tile writes (its 16 staged rows of 244 points are 62 KB: three CTAs per SM)

template <int MODEL, bool EXACT, int MINB, int TY = kPairTY>
__global__ void __launch_bounds__(256, MINB) rhs_pair_tile_kernel(const PairArgs a) {
  constexpr int PITCH = kPairPitch, ROWS = TY + 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2 *tile = reinterpret_cast<double2 *>(smem_raw);                  // [ROWS][PITCH]
  const unsigned bar = smem_u32(smem_raw + (size_t)ROWS * PITCH * 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nx = a.nx, nyl = a.nyl;
  const long long strips = (nx + kPairCols - 1) / kPairCols;
  const long long tj = blockIdx.x / strips, strip = blockIdx.x - tj * strips;
  const int i0 = (int)(strip * kPairCols), jA = (int)(tj * TY), nyli = (int)nyl;
  const int jB = (jA + TY < nyli) ? jA + TY : nyli;
  const int w = (nx - i0 < kPairCols) ? (int)(nx - i0) : kPairCols;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long m0 = (i0 >= 2) ? i0 - 2 : 0, m1 = (i0 + w + 2 <= nx) ? i0 + w + 2 : nx;
    const int lwrap = (int)(m0 - (i0 - 2)), rwrap = (int)(i0 + w + 2 - m1);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(w + 4) * 16u * (unsigned)ROWS) : "memory");
    for (int r = 0; r < ROWS; ++r) {
      long long jr = (long long)jA - 2 + r;
      const double2 *row;
      if (a.south_near && (jr < 0 || jr >= nyl)) {                         // a neighbouring rank's row
        row = reinterpret_cast<const double2 *>(jr == -1 ? a.south_near : jr < 0 ? a.south_far : jr == nyl ? a.north_near : a.north_far);
      } else {
        jr = jr < 0 ? jr + nyl : jr >= nyl ? jr - nyl : jr;
        if (jr >= nyl) jr -= nyl;                                          // (a partial last tile may reach two periods ahead)
        row = reinterpret_cast<const double2 *>(a.y) + jr * nx;
      }
      const unsigned dst = smem_u32(tile + (size_t)r * PITCH);
      bulk_g2s(dst + (unsigned)lwrap * 16u, row + m0, (unsigned)(m1 - m0) * 16u, bar);
      if (lwrap) bulk_g2s(dst, row + (nx - lwrap), (unsigned)lwrap * 16u, bar);
      if (rwrap) bulk_g2s(dst + (unsigned)(w + 4 - rwrap) * 16u, row, (unsigned)rwrap * 16u, bar);
    }
  }
  __syncthreads();
  const int q = 30 * warp + lane;
  const bool inner = (lane >= 1) && (lane <= 30);
  const bool writes = inner && (q - 1 < w);
  const int col = i0 + q - 1;
  double t1 = 0.0, t3 = 0.0;
  if (is_torus(MODEL) && q - 1 < w + 1) {
    const long long ci = col < 0 ? col + nx : col >= nx ? col - nx : col;
    const double2 tc = reinterpret_cast<const double2 *>(a.cth)[ci];
    t1 = tc.x; t3 = tc.y;
  }
  double2 *o1 = reinterpret_cast<double2 *>(a.f1) + (long long)jA * nx + col;
  double2 *o2 = reinterpret_cast<double2 *>(a.f2) + (long long)jA * nx + col;
  const int react_on = a.react;
  const double cz[2] = {a.c[0], a.c[1]};
  mbar_wait(bar, 0u);
  const double2 *my = tile + q;                                            // west at [0], centre at [1], east at [2] of a row
  // interior: every row the tile touches is a row of the slab that is written and never frozen
  const bool interior = jA >= 2 && jA + TY + 1 < nyli - 1;
  double2 zm = make_double2(0.0, 0.0), zc = make_double2(0.0, 0.0);        // z of the two rows before the one being formed
  double zcW = 0.0, zcE = 0.0;
  if (interior) {
    const double *bp = a.brow + (jA - 1);
#pragma unroll
    for (int s = 1; s <= TY + 2; ++s) {                                    // F1 and z of tile row s = slab row jA - 2 + s
      const double2 c = my[s * PITCH + 1];
      const double uW = my[s * PITCH].x, uE = my[s * PITCH + 2].x, uS = my[(s - 1) * PITCH + 1].x, uN = my[(s + 1) * PITCH + 1].x;
      double du = EXACT ? stencil_exact<MODEL>(a.k, t1, t3, c.x, uW, uE, uS, uN) : stencil_fast<MODEL>(a.k, t1, t3, c.x, uW, uE, uS, uN);
      double dv = 0.0;
      if (react_on) react<MODEL, EXACT>(a.k, __ldg(bp + (s - 1)), c.x, c.y, du, dv);
      if (writes && s >= 2 && s <= TY + 1) o1[(long long)(s - 2) * nx] = make_double2(du, dv);
      const double vx[2] = {c.x, du}, vy[2] = {c.y, dv};
      const double2 zn = make_double2(lc_value_n<EXACT, 2>(cz, vx), lc_value_n<EXACT, 2>(cz, vy));
      const double znW = __shfl_up_sync(0xffffffffu, zn.x, 1), znE = __shfl_down_sync(0xffffffffu, zn.x, 1);
      if (s >= 3) {                                                        // F2 of tile row s - 1 from z of rows s - 2, s - 1, s
        double d2u = EXACT ? stencil_exact<MODEL>(a.k, t1, t3, zc.x, zcW, zcE, zm.x, zn.x) : stencil_fast<MODEL>(a.k, t1, t3, zc.x, zcW, zcE, zm.x, zn.x);
        double d2v = 0.0;
        if (react_on) react<MODEL, EXACT>(a.k, __ldg(bp + (s - 2)), zc.x, zc.y, d2u, d2v);
        if (writes) o2[(long long)(s - 3) * nx] = make_double2(d2u, d2v);
      }
      zm = zc; zc = zn; zcW = znW; zcE = znE;
    }
  } else {
#pragma unroll 1
    for (int s = 1; s <= TY + 2; ++s) {
      const int j1 = jA - 2 + s;                                           // slab row (or its periodic image) of F1 / z
      if (j1 > jB) break;
      const int g1 = a.south_near ? j1 : (j1 < 0 ? j1 + nyli : j1 >= nyli ? j1 - nyli : j1);   // (brow[-1], brow[nyl] exist)
      const double2 c = my[s * PITCH + 1];
      const double uW = my[s * PITCH].x, uE = my[s * PITCH + 2].x, uS = my[(s - 1) * PITCH + 1].x, uN = my[(s + 1) * PITCH + 1].x;
      double du = EXACT ? stencil_exact<MODEL>(a.k, t1, t3, c.x, uW, uE, uS, uN) : stencil_fast<MODEL>(a.k, t1, t3, c.x, uW, uE, uS, uN);
      double dv = 0.0;
      if (react_on) {
        react<MODEL, EXACT>(a.k, __ldg(a.brow + g1), c.x, c.y, du, dv);
        const bool frozen = ((a.frz1 & 1) && g1 == 0) || ((a.frz1 & 2) && g1 == nyli - 1) || ((a.frz1 & 4) && g1 == -1) || ((a.frz1 & 8) && g1 == nyli);
        du = frozen ? 0.0 : du;
        dv = frozen ? 0.0 : dv;
      }
      if (writes && j1 >= jA && j1 < jB) o1[(long long)(j1 - jA) * nx] = make_double2(du, dv);
      const double vx[2] = {c.x, du}, vy[2] = {c.y, dv};
      const double2 zn = make_double2(lc_value_n<EXACT, 2>(cz, vx), lc_value_n<EXACT, 2>(cz, vy));
      const double znW = __shfl_up_sync(0xffffffffu, zn.x, 1), znE = __shfl_down_sync(0xffffffffu, zn.x, 1);
      const int j2 = j1 - 1;
      if (s >= 3 && j2 >= jA && j2 < jB) {
        double d2u = EXACT ? stencil_exact<MODEL>(a.k, t1, t3, zc.x, zcW, zcE, zm.x, zn.x) : stencil_fast<MODEL>(a.k, t1, t3, zc.x, zcW, zcE, zm.x, zn.x);
        double d2v = 0.0;
        if (react_on) {
          react<MODEL, EXACT>(a.k, __ldg(a.brow + j2), zc.x, zc.y, d2u, d2v);
          const bool frozen = ((a.frz2 & 1) && j2 == 0) || ((a.frz2 & 2) && j2 == nyli - 1);
          d2u = frozen ? 0.0 : d2u;
          d2v = frozen ? 0.0 : d2v;
        }
        if (writes) o2[(long long)(j2 - jA) * nx] = make_double2(d2u, d2v);
      }
      zm = zc; zc = zn; zcW = znW; zcE = znE;
    }
  }
}

// prepare_only: make sure the kernel is loaded and configured, launch nothing (a phi-split grid does this BEFORE it starts the
// exchange: loading a module can wait for the device, and the neighbours' launches are already spinning on this rank's rows)
template <int MODEL, bool EXACT, int MINB, int TY>
int launch_pair_tile_shape(crd_grid *g, const PairArgs &a, cudaStream_t st, bool prepare_only) {
  const long long strips = (a.nx + kPairCols - 1) / kPairCols, tiles = strips * ((a.nyl + TY - 1) / TY);
  if (tiles > 2147483647LL) { set_error("slab too large for one launch"); return -1; }
  const size_t smem = (size_t)(TY + 4) * kPairPitch * 16 + 16;
  auto kern = rhs_pair_tile_kernel<MODEL, EXACT, MINB, TY>;
  static bool attr_set[64] = {};
  const int dev = g->ctx->device & 63;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 113000);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set[dev] = true;
  }
  if (prepare_only) return 0;
  kern<<<(unsigned)tiles, 256, smem, st>>>(a);
  return check_launch(g->ctx, "rhs_pair_tile_kernel");
}
// FAST arithmetic is bound by memory: 12-row tiles (the smallest frame), three CTAs per SM.  EXACT arithmetic is bound by the
// latency of its dependent FP64 chains: 8-row tiles fit four CTAs per SM at 64 registers, and 32 instead of 24 warps per SM are
// worth more than the larger frame costs (2.67-3.16 against 3.21-3.67 ms at 16384^2; 10-row tiles: 2.73-3.2 ms).
template <int MODEL, bool EXACT>
int launch_pair_tile(crd_grid *g, const PairArgs &a, cudaStream_t st, bool prepare_only) {
  if (EXACT) return launch_pair_tile_shape<MODEL, EXACT, 4, 8>(g, a, st, prepare_only);
  return launch_pair_tile_shape<MODEL, EXACT, 3, kPairTY>(g, a, st, prepare_only);
}

template <int MODEL, bool EXACT>
int launch_pair_model(crd_grid *g, const PairArgs &a, cudaStream_t st, bool prepare_only) { return launch_pair_tile<MODEL, EXACT>(g, a, st, prepare_only); }

int launch_pair(crd_grid *g, const PairArgs &a, cudaStream_t st, bool prepare_only = false) {
  const bool exact = g->p.arith == CRD_ARITH_EXACT;
  switch (g->p.model) {
    case CRD_FHN_TORUS: return exact ? launch_pair_model<CRD_FHN_TORUS, true>(g, a, st, prepare_only) : launch_pair_model<CRD_FHN_TORUS, false>(g, a, st, prepare_only);
    case CRD_GOLDBETER_TORUS: return exact ? launch_pair_model<CRD_GOLDBETER_TORUS, true>(g, a, st, prepare_only) : launch_pair_model<CRD_GOLDBETER_TORUS, false>(g, a, st, prepare_only);
    case CRD_FHN_FLAT: return exact ? launch_pair_model<CRD_FHN_FLAT, true>(g, a, st, prepare_only) : launch_pair_model<CRD_FHN_FLAT, false>(g, a, st, prepare_only);
    case CRD_GOLDBETER_FLAT: return exact ? launch_pair_model<CRD_GOLDBETER_FLAT, true>(g, a, st, prepare_only) : launch_pair_model<CRD_GOLDBETER_FLAT, false>(g, a, st, prepare_only);
  }
  set_error("unknown model %d", g->p.model);
  return -1;
}

}  // namespace

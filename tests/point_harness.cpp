// point_harness.cpp — test infrastructure (tests/test_point_arithmetic_cpu.py): the product's per-point device source
// (crdmodel_b200/csrc/crd_rhs_point.cuh: stencil_exact, react, div_const_line, pow4_rn ...) and its host-side table builder
// (crd_tables.hpp) compiled for the HOST with shims for the CUDA intrinsics, every operation separately rounded
// (-ffp-contract=off), driven over a whole periodic grid.  What the GPU adds to this is only indexing and data movement.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>   // types only (double2, cudaStream_t); nothing of the runtime is called

static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __double2hiint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline int __double2loint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(u & 0xffffffffu); }
using std::max;
using std::min;
#define __noinline__ __attribute__((noinline))
#define asm(...) r = 1.0 / x   /* rcp_fast's PTX seed (FAST arithmetic, not exercised here) */
#include "crd_tables.hpp"
#undef asm

namespace {

template <int MODEL>
void sweep(const crd_grid &g, const std::vector<double> &cth, const std::vector<double> &brow, double t, int react_on,
           const double *y, double *ydot) {
  const long long nx = g.nx, ny = g.ny;
  const bool frozen_now = t < g.p.t_boundary;
  for (long long j = 0; j < ny; ++j) {
    const long long jS = (j == 0) ? ny - 1 : j - 1, jN = (j == ny - 1) ? 0 : j + 1;
    for (long long i = 0; i < nx; ++i) {
      const long long iW = (i == 0) ? nx - 1 : i - 1, iE = (i == nx - 1) ? 0 : i + 1;
      const double uC = y[2 * (j * nx + i)], v = y[2 * (j * nx + i) + 1];
      double du = stencil_exact<MODEL>(g.k, cth[2 * i], cth[2 * i + 1], uC, y[2 * (j * nx + iW)], y[2 * (j * nx + iE)],
                                       y[2 * (jS * nx + i)], y[2 * (jN * nx + i)]);
      double dv = 0.0;
      if (react_on) {
        react<MODEL, true>(g.k, brow[j + 1], uC, v, du, dv);   // (brow[0] / brow[nyl + 1]: the neighbouring ranks' rows)
        if (frozen_now && (j == 0 || j == ny - 1)) { du = 0.0; dv = 0.0; }   // :643-653
      }
      ydot[2 * (j * nx + i)] = du;
      ydot[2 * (j * nx + i) + 1] = dv;
    }
  }
}

}  // namespace

// one EXACT evaluation over a whole (single-slab) grid; returns 0, or -1 for an unknown model
extern "C" int point_harness_rhs(const crd_params *p, double t, int react_on, const double *y, double *ydot, double *dx_dy) {
  crd_grid g;
  g.p = *p;
  g.nx = p->nx; g.ny = p->ny; g.js = 0; g.je = p->ny - 1; g.nyl = p->ny;
  std::vector<double> cth, brow;
  grid_host_tables(&g, cth, brow);
  if (dx_dy) { dx_dy[0] = g.dx; dx_dy[1] = g.dy; }
  switch (p->model) {
    case CRD_FHN_TORUS: sweep<CRD_FHN_TORUS>(g, cth, brow, t, react_on, y, ydot); return 0;
    case CRD_GOLDBETER_TORUS: sweep<CRD_GOLDBETER_TORUS>(g, cth, brow, t, react_on, y, ydot); return 0;
    case CRD_FHN_FLAT: sweep<CRD_FHN_FLAT>(g, cth, brow, t, react_on, y, ydot); return 0;
    case CRD_GOLDBETER_FLAT: sweep<CRD_GOLDBETER_FLAT>(g, cth, brow, t, react_on, y, ydot); return 0;
  }
  return -1;
}

// the per-phi beta table of one rank of a phi split (js .. je of ny rows): out[0] = the row south of the slab, out[1 + j] = local
// row j, out[nyl + 1] = the row north of it (periodic in the global mesh)
extern "C" int point_harness_brow(const crd_params *p, double *out) {
  crd_grid g;
  g.p = *p;
  g.nx = p->nx; g.ny = p->ny; g.js = p->js; g.je = p->je; g.nyl = p->je - p->js + 1;
  std::vector<double> cth, brow;
  grid_host_tables(&g, cth, brow);
  std::copy(brow.begin(), brow.end(), out);
  return (int)brow.size();
}


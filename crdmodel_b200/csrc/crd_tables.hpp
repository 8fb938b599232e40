// crd_tables.hpp — host side of a grid: geometry, the scalars of the RHS kernels and the two coefficient tables, computed
// once per grid with the reference's own expressions and the host libm so that the EXACT kernels reproduce its rounding
// (src/FHNmodel_torus.cpp:188-189,233-234,531-537,623-632; FHNmodel_flat.cpp:173-176,229-230,489-491;
// GoldbeterModel_torus.cpp:694-695,715).  No CUDA calls: crd_grid_create (crd_rhs.cu) uploads the tables, and
// tests/test_point_arithmetic_cpu.py runs the same function next to the device's per-point source on the host.
#pragma once
#include <cmath>
#include <vector>

#include "crd_rhs_point.cuh"

namespace {

// needs g->p, g->nx, g->ny, g->js, g->nyl; fills the geometry members, g->k, cth[nx][2] and brow[nyl + 2] (local rows at 1 .. nyl)
inline void grid_host_tables(crd_grid *g, std::vector<double> &cth, std::vector<double> &brow) {
  const crd_params *p = &g->p;
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
  cth.assign((size_t)2 * g->nx, 0.0);
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
  // (one entry on either side for the rows of the neighbouring ranks, periodic in the global mesh: brow[0] is global row js - 1,
  // brow[1 + j] local row j, brow[nyl + 1] global row je + 1 — the pass that forms two evaluations at once evaluates them)
  brow.assign((size_t)g->nyl + 2, 0.0);
  const bool fhn = is_fhn(p->model);
  for (long long jj = -1; jj <= g->nyl; ++jj) {
    const long long gj = ((g->js + jj) % g->ny + g->ny) % g->ny;
    const long long j = jj + 1;
    const double yy = g->ymin + (gj) * (dy);
    double b = p->beta;
    const bool vary = fhn ? (p->vary_beta != 0) : (p->vary_beta == 1);
    if (vary) b = p->beta_min + yy * (p->beta_max - p->beta_min) / (g->ymax - g->ymin);
    brow[j] = fhn ? b : (G_v0 + G_v1 * b);
    if (fhn && b == 0.0 && std::signbit(b)) k.dv_plus0 = 1;
  }
}

}  // namespace

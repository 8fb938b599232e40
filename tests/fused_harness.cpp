// fused_harness.cpp — test infrastructure (tests/test_fused_arithmetic_cpu.py): the product's per-element arithmetic of the
// fused integrator operations (crdmodel_b200/csrc/crd_fused.cuh: stage combination, solution / error chains, error weights,
// double-double sums, the branch-free reciprocal) compiled for the HOST with shims for the CUDA intrinsics, every operation
// separately rounded (-ffp-contract=off).  What the kernels add to this is indexing, data movement and the reduction trees.
#include <cmath>
#include <cstdint>
#include <cstring>

#include <cuda_runtime.h>   // types and the __device__ / __forceinline__ decorations only; nothing of the runtime is called

static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __double2hiint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline int __double2loint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(u & 0xffffffffu); }
static inline double __shfl_down_sync(unsigned, double v, int) { return v; }
using std::fabs;
using std::fma;
// the hardware seed of the reciprocal (MUFU.RCP64H: the upper 32 bits of an approximation of 1/x, low word zero, relative
// error below 2^-20), modelled as the IEEE quotient perturbed by g_seed_err and truncated to its upper word
static double g_seed_err = 0.0;
static inline double seed_model(double x) {
  double r = (1.0 / x) * (1.0 + g_seed_err);
  uint64_t u; std::memcpy(&u, &r, 8); u &= 0xffffffff00000000ull; std::memcpy(&r, &u, 8);
  return r;
}
#define asm(...) r = seed_model(x)
#define CRD_FUSED_HOST_TEST   // take the device branch of dd_merge (intrinsics = the shims above)
#include "crd_fused.cuh"
#undef asm
#include "crd_grid.cuh"   // stream_seg_rows (host code)

using namespace crd;

extern "C" {

double fh_lc(int seq, int n, const double *c, const double *v) {
  double cc[5] = {0, 0, 0, 0, 0}, vv[5] = {0, 0, 0, 0, 0};
  for (int j = 0; j < n; ++j) { cc[j] = c[j]; vv[j] = v[j]; }
  return seq ? lc_value<true, 5>(cc, vv, n) : lc_value<false, 5>(cc, vv, n);
}

// the finish over n elements: ynew[i], and the three sums (error hi / lo, second norm)
void fh_finish(int seq, int S, const double *hb, const double *hd, double rtol, double atol, long n, const double *yn,
               const double *F /* [S][n] */, double *ynew, double *sums /* [3] */, int want_y2) {
  FinishArgs a;
  for (int j = 0; j < S; ++j) { a.hb[j] = hb[j]; a.hd[j] = hd[j]; }
  a.rtol = rtol; a.atol = atol;
  FinAcc<true> at;
  FinAcc<false> af;
  for (long i = 0; i < n; ++i) {
    double s = yn[i], err = 0.0;
    for (int j = 0; j < S; ++j) {
      const double f = F[(long)j * n + i];
      if (seq) { s = fin_sol_term<true>(a.hb[j], f, s, a.hb[j] != 0.0); err = fin_err_term<true>(a.hd[j], f, err); }
      else { s = fin_sol_term<false>(a.hb[j], f, s, a.hb[j] != 0.0); err = fin_err_term<false>(a.hd[j], f, err); }
    }
    ynew[i] = s;
    if (seq) finish_tail<true>(rtol, atol, yn[i], s, err, at, want_y2 != 0);
    else finish_tail<false>(rtol, atol, yn[i], s, err, af, want_y2 != 0);
  }
  sums[0] = seq ? at.e_hi : af.e_hi; sums[1] = seq ? at.y2 : af.y2; sums[2] = seq ? at.e_lo : af.e_lo;
}

double fh_rcp(double x, double seed_err) { g_seed_err = seed_err; return rcp_rn(x); }
int fh_rcp_in_range(double x) { return rcp_rn_in_range(x) ? 1 : 0; }

// sum of non-negative terms the way the kernels do it: `lanes` interleaved accumulators (dd_add), merged pairwise (dd_merge)
int fh_stream_seg_rows(long nyl, long strips, long ctas) { return stream_seg_rows(nyl, strips, ctas); }

double fh_dd_sum(const double *q, long n, int lanes) {
  double hi[1024], lo[1024];
  if (lanes > 1024) lanes = 1024;
  for (int l = 0; l < lanes; ++l) { hi[l] = 0.0; lo[l] = 0.0; }
  for (long i = 0; i < n; ++i) dd_add(hi[i % lanes], lo[i % lanes], q[i]);
  for (int step = 1; step < lanes; step *= 2)
    for (int l = 0; l + step < lanes; l += 2 * step) dd_merge(hi[l], lo[l], hi[l + step], lo[l + step]);
  return hi[0] + lo[0];
}

}  // extern "C"

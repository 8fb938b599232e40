/* TEST INFRASTRUCTURE — CPU checker, never linked into the product library.
 *
 * Host-memory restatement of SUNDIALS' nvector_parallel (the vector arithmetic the reference
 * programs get from libsundials_nvecparallel; CMake/FindSUNDIALS.cmake:6, call sites
 * src/FHNmodel_torus.cpp:281,303,383,488,506,517-518,786).  SUNDIALS is a third-party dependency
 * that is NOT in /root/reference and NOT pinned by it (API brackets it to 2.6.0-2.7.0); this file
 * restates the published algorithm of nvector_parallel.c: serial loops over the local segment in
 * index order, special-cased N_VLinearSum / N_VScale, one MPI_Allreduce per reduction, WRMS =
 * sqrt(sum_global((x_i w_i)^2) / N_global).  PARITY UNPINNED against SUNDIALS itself: the reference
 * ships no golden vectors for these ops; they are pinned against closed-form numpy results in
 * tests/test_oracle_nvector.py.
 *
 * One deliberate refinement: the weighted square sums (N_VWrmsNorm, N_VWrmsNormMask, N_VWL2Norm) add the terms
 * RN(RN(x_i w_i)^2) — each rounded exactly like nvector_parallel's loop — EXACTLY (Shewchuk's growing expansion, as Python's
 * math.fsum), gather the ranks' (hi, lo) pairs and round once, instead of accumulating in index order.  The result is the
 * correctly rounded sum, independent of the order of summation and of the decomposition, which is also what the device
 * reductions deliver (double-double accumulation): the integrator's error norm, hence its step sequence, is then the
 * same on both sides bit for bit.  CRD_ORACLE_PLAIN_SUM=1 restores the serial `sum += prodi*prodi` (differs in the last
 * bits only).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "nvector/nvector_parallel.h"

#define ZERO 0.0
#define HALF 0.5
#define ONE 1.0
#define ONEPT5 1.5
#define BIG_REAL DBL_MAX

static double allreduce(double d, int op, MPI_Comm comm) {
  double out;
  MPI_Allreduce(&d, &out, 1, MPI_DOUBLE, op, comm);
  return out;
}

/* ---- exact accumulation of doubles (Shewchuk 1997, "msum"): the partials are a non-overlapping expansion of the sum ---- */
typedef struct { double p[64]; int n; } exact_acc;
static void xs_add(exact_acc *a, double x) {
  int i = 0, j;
  for (j = 0; j < a->n; j++) {
    double y = a->p[j], hi, lo;
    if (fabs(x) < fabs(y)) { double t = x; x = y; y = t; }
    hi = x + y;
    lo = y - (hi - x);
    if (lo != 0.0) a->p[i++] = lo;
    x = hi;
  }
  a->p[i++] = x;
  a->n = i;
}
/* (hi, lo): hi = the correctly rounded sum (round-half-even fix-up as in CPython's math.fsum), lo = what the next partials add */
static void xs_result(const exact_acc *a, double *hi_out, double *lo_out) {
  int n = a->n;
  double hi = 0.0, lo = 0.0;
  if (n > 0) {
    hi = a->p[--n];
    while (n > 0) {
      double x = hi, y = a->p[--n];
      hi = x + y;
      lo = y - (hi - x);
      if (lo != 0.0) break;
    }
    if (n > 0 && ((lo < 0.0 && a->p[n - 1] < 0.0) || (lo > 0.0 && a->p[n - 1] > 0.0))) {
      double y = lo * 2.0, x = hi + y, yr = x - hi;
      if (y == yr) { hi = x; lo = 0.0; }   /* exactly half-way and the tail pushes it over */
    }
    /* remainder below hi, for merging with other ranks' sums */
    { double rest = lo; int k; for (k = n - 1; k >= 0; k--) rest += a->p[k]; lo = rest; }
  }
  *hi_out = hi; *lo_out = lo;
}
static int plain_sum(void) {
  static int v = -1;
  if (v < 0) { const char *e = getenv("CRD_ORACLE_PLAIN_SUM"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
}
/* global value of a local exact sum: the ranks' (hi, lo) pairs gathered (a SUM over a vector that is zero except for the
 * rank's own slots is exact) and added in rank order in double-double, rounded once */
static double allreduce_exact(const exact_acc *a, MPI_Comm comm) {
  int size = 1, rank = 0, r;
  double hi, lo;
  xs_result(a, &hi, &lo);
  MPI_Comm_size(comm, &size);
  if (size <= 1) return hi + lo;
  MPI_Comm_rank(comm, &rank);
  {
    double *in = (double *)calloc((size_t)4 * size, sizeof(double)), *out = in + 2 * size;
    exact_acc t;
    double res, rl;
    in[2 * rank] = hi; in[2 * rank + 1] = lo;
    MPI_Allreduce(in, out, 2 * size, MPI_DOUBLE, MPI_SUM, comm);
    t.n = 0;
    for (r = 0; r < 2 * size; r++) xs_add(&t, out[r]);
    xs_result(&t, &res, &rl);
    free(in);
    return res;
  }
}

static N_Vector nvh_clone(N_Vector w);
static N_Vector nvh_cloneempty(N_Vector w);
static void nvh_destroy(N_Vector v);

static void nvh_space(N_Vector v, long int *lrw, long int *liw) {
  int npes;
  MPI_Comm_size(NV_COMM_P(v), &npes);
  *lrw = NV_GLOBLENGTH_P(v);
  *liw = 2 * npes;
}
static realtype *nvh_getarraypointer(N_Vector v) { return NV_DATA_P(v); }
static void nvh_setarraypointer(realtype *d, N_Vector v) {
  if (NV_LOCLENGTH_P(v) > 0) NV_DATA_P(v) = d;
}

static void nvh_linearsum(realtype a, N_Vector x, realtype b, N_Vector y, N_Vector z) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype *xd = NV_DATA_P(x), *yd = NV_DATA_P(y), *zd = NV_DATA_P(z);
  realtype c, *v1, *v2;
  int test;
  if (b == ONE && z == y) { for (i = 0; i < N; i++) yd[i] += a * xd[i]; return; }       /* Vaxpy */
  if (a == ONE && z == x) { for (i = 0; i < N; i++) xd[i] += b * yd[i]; return; }
  if (a == ONE && b == ONE) { for (i = 0; i < N; i++) zd[i] = xd[i] + yd[i]; return; } /* VSum */
  if ((test = (a == ONE && b == -ONE)) || (a == -ONE && b == ONE)) {                    /* VDiff */
    v1 = test ? yd : xd; v2 = test ? xd : yd;
    for (i = 0; i < N; i++) zd[i] = v2[i] - v1[i];
    return;
  }
  if ((test = (a == ONE)) || b == ONE) {                                                /* VLin1 */
    c = test ? b : a; v1 = test ? yd : xd; v2 = test ? xd : yd;
    for (i = 0; i < N; i++) zd[i] = c * v1[i] + v2[i];
    return;
  }
  if ((test = (a == -ONE)) || b == -ONE) {                                              /* VLin2 */
    c = test ? b : a; v1 = test ? yd : xd; v2 = test ? xd : yd;
    for (i = 0; i < N; i++) zd[i] = c * v1[i] - v2[i];
    return;
  }
  if (a == b) { for (i = 0; i < N; i++) zd[i] = a * (xd[i] + yd[i]); return; }          /* VScaleSum */
  if (a == -b) { for (i = 0; i < N; i++) zd[i] = a * (xd[i] - yd[i]); return; }         /* VScaleDiff */
  for (i = 0; i < N; i++) zd[i] = a * xd[i] + b * yd[i];
}

static void nvh_const(realtype c, N_Vector z) {
  long int i, N = NV_LOCLENGTH_P(z);
  realtype *zd = NV_DATA_P(z);
  for (i = 0; i < N; i++) zd[i] = c;
}
static void nvh_prod(N_Vector x, N_Vector y, N_Vector z) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype *xd = NV_DATA_P(x), *yd = NV_DATA_P(y), *zd = NV_DATA_P(z);
  for (i = 0; i < N; i++) zd[i] = xd[i] * yd[i];
}
static void nvh_div(N_Vector x, N_Vector y, N_Vector z) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype *xd = NV_DATA_P(x), *yd = NV_DATA_P(y), *zd = NV_DATA_P(z);
  for (i = 0; i < N; i++) zd[i] = xd[i] / yd[i];
}
static void nvh_scale(realtype c, N_Vector x, N_Vector z) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype *xd = NV_DATA_P(x), *zd = NV_DATA_P(z);
  if (z == x) { for (i = 0; i < N; i++) xd[i] *= c; return; }
  if (c == ONE) { for (i = 0; i < N; i++) zd[i] = xd[i]; return; }
  if (c == -ONE) { for (i = 0; i < N; i++) zd[i] = -xd[i]; return; }
  for (i = 0; i < N; i++) zd[i] = c * xd[i];
}
static void nvh_abs(N_Vector x, N_Vector z) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype *xd = NV_DATA_P(x), *zd = NV_DATA_P(z);
  for (i = 0; i < N; i++) zd[i] = fabs(xd[i]);
}
static void nvh_inv(N_Vector x, N_Vector z) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype *xd = NV_DATA_P(x), *zd = NV_DATA_P(z);
  for (i = 0; i < N; i++) zd[i] = ONE / xd[i];
}
static void nvh_addconst(N_Vector x, realtype b, N_Vector z) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype *xd = NV_DATA_P(x), *zd = NV_DATA_P(z);
  for (i = 0; i < N; i++) zd[i] = xd[i] + b;
}
static realtype nvh_dotprod(N_Vector x, N_Vector y) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype sum = ZERO, *xd = NV_DATA_P(x), *yd = NV_DATA_P(y);
  for (i = 0; i < N; i++) sum += xd[i] * yd[i];
  return allreduce(sum, MPI_SUM, NV_COMM_P(x));
}
static realtype nvh_maxnorm(N_Vector x) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype max = ZERO, *xd = NV_DATA_P(x);
  for (i = 0; i < N; i++) if (fabs(xd[i]) > max) max = fabs(xd[i]);
  return allreduce(max, MPI_MAX, NV_COMM_P(x));
}
static realtype nvh_wrmsnorm(N_Vector x, N_Vector w) {
  long int i, N = NV_LOCLENGTH_P(x), Ng = NV_GLOBLENGTH_P(x);
  realtype sum = ZERO, prodi, *xd = NV_DATA_P(x), *wd = NV_DATA_P(w);
  if (!plain_sum()) {
    exact_acc a; a.n = 0;
    for (i = 0; i < N; i++) { prodi = xd[i] * wd[i]; xs_add(&a, prodi * prodi); }
    return sqrt(allreduce_exact(&a, NV_COMM_P(x)) / Ng);
  }
  for (i = 0; i < N; i++) { prodi = xd[i] * wd[i]; sum += prodi * prodi; }
  return sqrt(allreduce(sum, MPI_SUM, NV_COMM_P(x)) / Ng);
}
static realtype nvh_wrmsnormmask(N_Vector x, N_Vector w, N_Vector id) {
  long int i, N = NV_LOCLENGTH_P(x), Ng = NV_GLOBLENGTH_P(x);
  realtype sum = ZERO, prodi, *xd = NV_DATA_P(x), *wd = NV_DATA_P(w), *idd = NV_DATA_P(id);
  if (!plain_sum()) {
    exact_acc a; a.n = 0;
    for (i = 0; i < N; i++) if (idd[i] > ZERO) { prodi = xd[i] * wd[i]; xs_add(&a, prodi * prodi); }
    return sqrt(allreduce_exact(&a, NV_COMM_P(x)) / Ng);
  }
  for (i = 0; i < N; i++) if (idd[i] > ZERO) { prodi = xd[i] * wd[i]; sum += prodi * prodi; }
  return sqrt(allreduce(sum, MPI_SUM, NV_COMM_P(x)) / Ng);
}
static realtype nvh_min(N_Vector x) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype min = BIG_REAL, *xd = NV_DATA_P(x);
  if (N > 0) { min = xd[0]; for (i = 1; i < N; i++) if (xd[i] < min) min = xd[i]; }
  return allreduce(min, MPI_MIN, NV_COMM_P(x));
}
static realtype nvh_wl2norm(N_Vector x, N_Vector w) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype sum = ZERO, prodi, *xd = NV_DATA_P(x), *wd = NV_DATA_P(w);
  if (!plain_sum()) {
    exact_acc a; a.n = 0;
    for (i = 0; i < N; i++) { prodi = xd[i] * wd[i]; xs_add(&a, prodi * prodi); }
    return sqrt(allreduce_exact(&a, NV_COMM_P(x)));
  }
  for (i = 0; i < N; i++) { prodi = xd[i] * wd[i]; sum += prodi * prodi; }
  return sqrt(allreduce(sum, MPI_SUM, NV_COMM_P(x)));
}
static realtype nvh_l1norm(N_Vector x) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype sum = ZERO, *xd = NV_DATA_P(x);
  for (i = 0; i < N; i++) sum += fabs(xd[i]);
  return allreduce(sum, MPI_SUM, NV_COMM_P(x));
}
static void nvh_compare(realtype c, N_Vector x, N_Vector z) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype *xd = NV_DATA_P(x), *zd = NV_DATA_P(z);
  for (i = 0; i < N; i++) zd[i] = (fabs(xd[i]) >= c) ? ONE : ZERO;
}
static booleantype nvh_invtest(N_Vector x, N_Vector z) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype *xd = NV_DATA_P(x), *zd = NV_DATA_P(z), val = ONE;
  for (i = 0; i < N; i++) { if (xd[i] == ZERO) val = ZERO; else zd[i] = ONE / xd[i]; }
  return allreduce(val, MPI_MIN, NV_COMM_P(x)) == ZERO ? FALSE : TRUE;
}
static booleantype nvh_constrmask(N_Vector c, N_Vector x, N_Vector m) {
  long int i, N = NV_LOCLENGTH_P(x);
  realtype temp = ONE, *cd = NV_DATA_P(c), *xd = NV_DATA_P(x), *md = NV_DATA_P(m);
  for (i = 0; i < N; i++) {
    md[i] = ZERO;
    if (cd[i] == ZERO) continue;
    if (cd[i] > ONEPT5 || cd[i] < -ONEPT5) {
      if (xd[i] * cd[i] <= ZERO) { temp = ZERO; md[i] = ONE; }
      continue;
    }
    if (cd[i] > HALF || cd[i] < -HALF) {
      if (xd[i] * cd[i] < ZERO) { temp = ZERO; md[i] = ONE; }
    }
  }
  return allreduce(temp, MPI_MIN, NV_COMM_P(x)) == ONE ? TRUE : FALSE;
}
static realtype nvh_minquotient(N_Vector num, N_Vector denom) {
  long int i, N = NV_LOCLENGTH_P(num);
  realtype min = BIG_REAL, *nd = NV_DATA_P(num), *dd = NV_DATA_P(denom);
  booleantype notEvenOnce = TRUE;
  for (i = 0; i < N; i++) {
    if (dd[i] == ZERO) continue;
    if (!notEvenOnce) { realtype q = nd[i] / dd[i]; if (q < min) min = q; }
    else { min = nd[i] / dd[i]; notEvenOnce = FALSE; }
  }
  return allreduce(min, MPI_MIN, NV_COMM_P(num));
}

static struct _generic_N_Vector_Ops nvh_ops = {
#ifdef CRD_SUNDIALS_27
    NULL,
#endif
    nvh_clone, nvh_cloneempty, nvh_destroy, nvh_space, nvh_getarraypointer, nvh_setarraypointer,
    nvh_linearsum, nvh_const, nvh_prod, nvh_div, nvh_scale, nvh_abs, nvh_inv, nvh_addconst,
    nvh_dotprod, nvh_maxnorm, nvh_wrmsnorm, nvh_wrmsnormmask, nvh_min, nvh_wl2norm, nvh_l1norm,
    nvh_compare, nvh_invtest, nvh_constrmask, nvh_minquotient};

N_Vector N_VNewEmpty_Parallel(MPI_Comm comm, long int local_length, long int global_length) {
  N_Vector v = (N_Vector)malloc(sizeof *v);
  if (!v) return NULL;
  N_VectorContent_Parallel c = (N_VectorContent_Parallel)malloc(sizeof *c);
  if (!c) { free(v); return NULL; }
  c->local_length = local_length;
  c->global_length = global_length;
  c->comm = comm;
  c->own_data = FALSE;
  c->data = NULL;
  v->content = c;
  v->ops = &nvh_ops;
  return v;
}

N_Vector N_VNew_Parallel(MPI_Comm comm, long int local_length, long int global_length) {
  N_Vector v = N_VNewEmpty_Parallel(comm, local_length, global_length);
  if (!v) return NULL;
  if (local_length > 0) {
    realtype *d = (realtype *)malloc((size_t)local_length * sizeof(realtype));
    if (!d) { N_VDestroy_Parallel(v); return NULL; }
    NV_OWN_DATA_P(v) = TRUE;
    NV_DATA_P(v) = d;
  }
  return v;
}

N_Vector N_VMake_Parallel(MPI_Comm comm, long int local_length, long int global_length, realtype *v_data) {
  N_Vector v = N_VNewEmpty_Parallel(comm, local_length, global_length);
  if (!v) return NULL;
  if (local_length > 0) { NV_OWN_DATA_P(v) = FALSE; NV_DATA_P(v) = v_data; }
  return v;
}

void N_VDestroy_Parallel(N_Vector v) {
  if (!v) return;
  if (NV_OWN_DATA_P(v) == TRUE && NV_DATA_P(v)) { free(NV_DATA_P(v)); NV_DATA_P(v) = NULL; }
  free(v->content);
  free(v);
}

static N_Vector nvh_cloneempty(N_Vector w) {
  return N_VNewEmpty_Parallel(NV_COMM_P(w), NV_LOCLENGTH_P(w), NV_GLOBLENGTH_P(w));
}
static N_Vector nvh_clone(N_Vector w) {
  return N_VNew_Parallel(NV_COMM_P(w), NV_LOCLENGTH_P(w), NV_GLOBLENGTH_P(w));
}
static void nvh_destroy(N_Vector v) { N_VDestroy_Parallel(v); }

/* TEST INFRASTRUCTURE — CPU checker, never linked into the product library.
 *
 * Plain-C restatement of the reference's right-hand side f() for one subdomain with periodic wrap,
 * i.e. what the reference computes at np = 1 (self-exchange, FHNmodel_torus.cpp:854-893):
 *   FHN torus        src/FHNmodel_torus.cpp:504-667        (stencil :527-615, kinetics :621-664)
 *   Goldbeter torus  src/GoldbeterModel_torus.cpp:547-724  (stencil :571-659, kinetics :668-721)
 *   FHN flat         src/FHNmodel_flat.cpp:469-616         (stencil :489-566, kinetics :571-613)
 *   Goldbeter flat   src/GoldbeterModel_flat.cpp:515-689   (stencil :537-616, kinetics :625-686)
 * Operation order follows the C expressions of those lines exactly (left-to-right sums, a*b/c as
 * (a*b)/c, pow() from libm), so that with -O2 -ffp-contract=off and no -march the result is
 * BIT-IDENTICAL to the compiled-in-place reference (oracle/_ref).  Pinned by
 * tests/test_oracle.py against oracle/_ref (when built) and against tests/golden/*.npz, which were
 * generated from oracle/_ref by tests/golden/make_golden.py.
 */
#include <math.h>
#include <stddef.h>

#include "crd_oracle.h"

#define PI 3.1415926535897932 /* FHNmodel_torus.cpp:63 */
#define EPSILON 0.36          /* FHNmodel_torus.cpp:68 */
/* GoldbeterModel_torus.cpp:67-78 */
#define G_v0 1.0
#define G_k 10.0
#define G_kf 1.0
#define G_v1 7.3
#define G_VM2 65.0
#define G_VM3 500.0
#define G_K2 1.0
#define G_KR 2.0
#define G_KA 0.9
#define G_m 2.0
#define G_n 2.0
#define G_p 4.0

/* rows [j0, j1) of the global mesh.  yoff = 0: y is the whole global state.  yoff = j0 - 1 (band form): y holds only the
 * rows j0-1 .. j1 (periodic in the global mesh), i.e. row j of the mesh is row j - yoff of y with the wrapped neighbours of
 * row 0 / ny-1 stored where the band has them (first / last row of y). */
static int rhs_rows_impl(const crd_oracle_params *P, double t, const double *y, double *out, long j0, long j1, int band) {
  const long nx = P->nx, ny = P->ny;
  const int torus = (P->model == CRD_ORACLE_FHN_TORUS || P->model == CRD_ORACLE_GOLDBETER_TORUS);
  const int fhn = (P->model == CRD_ORACLE_FHN_TORUS || P->model == CRD_ORACLE_FHN_FLAT);
  if (nx < 2 || ny < 2 || j0 < 0 || j1 > ny) return -1;
  double XMIN = 0.0, XMAX, YMIN = 0.0, YMAX, R = 0.0, r = 0.0;
  if (torus) {
    XMAX = 2.0 * PI; YMAX = 2.0 * PI;                       /* :73-76 */
    r = P->surface_width / (2.0 * PI);                      /* :188 */
    R = P->surface_length / (2.0 * PI);                     /* :189 */
  } else {
    XMAX = P->surface_width - XMIN;                         /* FHNmodel_flat.cpp:173-176 */
    YMAX = P->surface_length - YMIN;
  }
  const double dx = (XMAX - XMIN) / (1.0 * nx - 1.0);       /* :233 */
  const double dy = (YMAX - YMIN) / (1.0 * ny - 1.0);       /* :234 */
  const double Diff = P->diff;
  const double cu1 = Diff / dx / dx, cu2 = Diff / dy / dy;  /* FHNmodel_flat.cpp:489-491 */
  const double cu3 = -2.0 * (cu1 + cu2);
  const int react = fhn || P->just_diffusion == 0;          /* GoldbeterModel_torus.cpp:668 */

  for (long j = j0; j < j1; ++j) {
    long jS = (j == 0) ? ny - 1 : j - 1, jN = (j == ny - 1) ? 0 : j + 1, jC = j;
    if (band) { jC = j - j0 + 1; jS = jC - 1; jN = jC + 1; }
    const double yy = YMIN + (j) * (dy);                    /* :623 (js = 0) */
    double b = P->beta;
    if (fhn) { if (P->vary_beta != 0) b = P->beta_min + yy * (P->beta_max - P->beta_min) / (YMAX - YMIN); } /* :625-632 */
    else if (P->vary_beta == 1) b = P->beta_min + yy * (P->beta_max - P->beta_min) / (YMAX - YMIN);        /* GB :675-683 */
    /* frozen rows while t < tBoundary (:643-653); north row is tested first */
    const int frozen = react && t < P->t_boundary && (j == ny - 1 || j == 0);
    for (long i = 0; i < nx; ++i) {
      const long iW = (i == 0) ? nx - 1 : i - 1, iE = (i == nx - 1) ? 0 : i + 1;
      const double uC = y[2 * (i + jC * nx)], vC = y[2 * (i + jC * nx) + 1];
      const double uW = y[2 * (iW + jC * nx)], uE = y[2 * (iE + jC * nx)];
      const double uS = y[2 * (i + jS * nx)], uN = y[2 * (i + jN * nx)];
      double du, dv = 0.0;                                  /* N_VConst(0.0, ydot) :506 */
      if (torus) {
        const double xx = XMIN + (i) * (dx);                /* :531 */
        du = Diff * ((-sin(xx) / (r * (R + r * cos(xx)))) * (uE - uW)) / (2 * dx)
           + Diff * ((1 / (r * r)) * (uE - 2 * uC + uW)) / (dx * dx)
           + Diff * ((1 / (((R + r * cos(xx))) * ((R + r * cos(xx))))) * (uN - 2 * uC + uS)) / (dy * dy); /* :535-537 */
      } else {
        du = cu1 * (uW + uE) + cu2 * (uS + uN) + cu3 * uC;  /* FHNmodel_flat.cpp:496-498 */
      }
      if (react) {
        if (frozen) { du = 0; dv = 0; }
        else if (fhn) {
          du += 3.0 * uC - (uC * uC * uC) - vC;             /* :657 */
          dv += EPSILON * (uC + b);                         /* :660 */
        } else {
          const double Z = uC, Y = vC;
          const double v2 = G_VM2 * pow(Z, G_n) / (pow(G_K2, G_n) + pow(Z, G_n));                      /* GB :694 */
          const double v3 = G_VM3 * pow(Y, G_m) * pow(Z, G_p) / ((pow(G_KR, G_m) + pow(Y, G_m)) * (pow(G_KA, G_p) + pow(Z, G_p))); /* :695 */
          du += G_v0 + G_v1 * b - v2 + v3 + G_kf * Y - G_k * Z;                                         /* :715 */
          dv += v2 - v3 - G_kf * Y;                                                                     /* :716 */
        }
      }
      out[2 * (i + (j - j0) * nx)] = du;
      out[2 * (i + (j - j0) * nx) + 1] = dv;
    }
  }
  return 0;
}

int crd_oracle_rhs_rows(const crd_oracle_params *P, double t, const double *y, double *out, long j0, long j1) {
  return rhs_rows_impl(P, t, y, out, j0, j1, 0);
}

/* band form (what one rank of a phi split computes): yband = rows j0-1 .. j0+nrows of the global mesh (nrows + 2 rows) */
int crd_oracle_rhs_band(const crd_oracle_params *P, double t, long j0, long nrows, const double *yband, double *out) {
  return rhs_rows_impl(P, t, yband, out, j0, j0 + nrows, 1);
}

int crd_oracle_rhs(const crd_oracle_params *P, double t, const double *y, double *ydot) {
  return crd_oracle_rhs_rows(P, t, y, ydot, 0, P->ny);
}

/* s <- s*6364136223846793005 + 1442695040888963407 (mod 2^64), u = (s >> 11) * 2^-53,
 * element e uses the (e+1)-th state after `seed`.  FHN: 4u - 2 in [-2, 2); Goldbeter: 1.5u + 0.1. */
void crd_oracle_fill_state(int model, unsigned long long seed, long first_elem, long n_elems, double *out) {
  const unsigned long long A = 6364136223846793005ULL, C = 1442695040888963407ULL;
  /* jump ahead first_elem steps: compose the affine map with itself by squaring */
  unsigned long long accA = 1ULL, accC = 0ULL, curA = A, curC = C;
  for (unsigned long long kk = (unsigned long long)first_elem; kk; kk >>= 1) {
    if (kk & 1ULL) { accA = accA * curA; accC = accC * curA + curC; }
    curC = curC * curA + curC;
    curA = curA * curA;
  }
  unsigned long long s = accA * seed + accC;
  const int fhn = (model == CRD_ORACLE_FHN_TORUS || model == CRD_ORACLE_FHN_FLAT);
  for (long e = 0; e < n_elems; ++e) {
    s = s * A + C;
    double u = (double)(s >> 11) * (1.0 / 9007199254740992.0);
    out[e] = fhn ? 4.0 * u - 2.0 : 1.5 * u + 0.1;
  }
}

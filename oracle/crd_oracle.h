/* TEST INFRASTRUCTURE — interface of the CPU checker (oracle/crd_oracle.c, the plain-C restatement)
 * and of the compiled-in-place reference (oracle/ref_harness.cpp -> oracle/_ref/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 * Never linked into libcrd_b200.so. */
#ifndef CRD_ORACLE_H
#define CRD_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

enum { CRD_ORACLE_FHN_TORUS = 0, CRD_ORACLE_GOLDBETER_TORUS = 1, CRD_ORACLE_FHN_FLAT = 2, CRD_ORACLE_GOLDBETER_FLAT = 3 };

/* The ini-file parameters that reach f() (reference: FHNmodel_torus.cpp:160-174, 227-234). */
typedef struct crd_oracle_params {
  int model;
  long nx, ny;                 /* global mesh: theta/x (fastest index) and phi/y */
  double diff;                 /* Parameters.diffusion */
  double beta;                 /* Parameters.beta */
  double beta_min, beta_max;   /* Parameters.betaMin / betaMax */
  int vary_beta;               /* System.varyBeta */
  int just_diffusion;          /* System.justDiffusion (Goldbeter only) */
  double t_boundary;           /* Parameters.tBoundary */
  double surface_length;       /* Parameters.surfaceLength: major circumference (torus) / length (flat) */
  double surface_width;        /* Parameters.surfaceWidth:  minor circumference (torus) / width  (flat) */
} crd_oracle_params;

/* oracle/crd_oracle.c — restatement, single subdomain with periodic wrap (= the reference at np = 1) */
int crd_oracle_rhs(const crd_oracle_params *P, double t, const double *y, double *ydot);
/* rows [j0, j1) only (ydot points at row j0); used to time a bounded sample */
int crd_oracle_rhs_rows(const crd_oracle_params *P, double t, const double *y, double *ydot_rows, long j0, long j1);
/* band form: rows [j0, j0+nrows) of the global mesh from yband = rows j0-1 .. j0+nrows (periodic in the global mesh) */
int crd_oracle_rhs_band(const crd_oracle_params *P, double t, long j0, long nrows, const double *yband, double *out);
/* synthetic state of SURVEY.md §8(d): 64-bit LCG, seed-addressable by element offset */
void crd_oracle_fill_state(int model, unsigned long long seed, long first_elem, long n_elems, double *out);

/* oracle/ref_harness.cpp — the reference itself */
int crd_ref_kind(void);
int crd_ref_rhs(const crd_oracle_params *P, int nranks, double t, const double *y, double *ydot, int reps, double *seconds);
/* The reference's f() on the band of global phi rows [j0, j0+nrows) of the mesh P describes (theta whole), one rank:
 * yband holds rows j0-1 .. j0+nrows (nrows+2 rows of [nx][2], periodic in the GLOBAL mesh: row -1 is row ny-1), the first
 * and last of which reach Exchange()'s S / N receive buffers as a neighbouring rank's message would; out gets nrows rows. */
int crd_ref_rhs_band(const crd_oracle_params *P, double t, long j0, long nrows, const double *yband, double *out);
int crd_ref_decomp(const crd_oracle_params *P, int nranks, int rank, long out[8]);
int crd_ref_main(const char *ini_path, int nranks);

/* oracle/shim/mpi_shim.cpp */
void crdshim_mpi_set_world(int nranks);
void crdshim_mpi_bind(int rank);
void crdshim_mpi_band_halo(const double *south_row, const double *north_row);

#ifdef __cplusplus
}
#endif
#endif

// TEST INFRASTRUCTURE — compiles ONE of the reference programs *in place* (never copied):
//     #define main ref_main ; #include "/root/reference/src/<Model>.cpp"
// against the shim headers in oracle/shim, and exports the reference's own f()/Exchange()/
// SetupDecomp() (reference: src/FHNmodel_torus.cpp:504-667,708-950 and the same functions of the
// other three programs) as C symbols.  Built by oracle/Makefile into oracle/_ref/libcrd_ref_<kind>.so
// (one library per program because the four sources define the same global names).
// Used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
//
// REF_KIND: 0 FHN torus, 1 Goldbeter torus, 2 FHN flat, 3 Goldbeter flat.  REF_SRC: path of the source.
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

#include "crd_oracle.h"

#define main ref_main
#include REF_SRC
#undef main

// the Goldbeter sources #define one-letter names (GoldbeterModel_torus.cpp:67-78); drop them
#undef v0
#undef k
#undef kf
#undef v1
#undef VM2
#undef VM3
#undef K2
#undef KR
#undef KA
#undef m
#undef n
#undef p

#if REF_KIND == 0 || REF_KIND == 2
#define REF_INIT InitUserData
#define REF_DECOMP SetupDecomp
#define REF_FREE FreeUserData
#else
#define REF_INIT init_user_data
#define REF_DECOMP setup_decomp
#define REF_FREE free_user_data
#endif

namespace {

// file-scope parameters, set the way main() sets them from the ini file
void set_globals(const crd_oracle_params *P) {
  DIFF = P->diff;
  BETA = P->beta;
#if REF_KIND == 0 || REF_KIND == 2
  TBOUNDARY = P->t_boundary;
  VARYBETA = P->vary_beta;
  BETAMIN = P->beta_min;
  BETAMAX = P->beta_max;
#else
  T_BOUNDARY = P->t_boundary;
  VARY_BETA = P->vary_beta;
  BETA_MIN = P->beta_min;   // NB GoldbeterModel_torus.cpp never reads betaMin/Max from the ini (stay 0)
  BETA_MAX = P->beta_max;
  JUST_DIFFUSION = P->just_diffusion;
#endif
#if REF_KIND == 0
  MAJORCIRC = P->surface_length; MINORCIRC = P->surface_width;
#elif REF_KIND == 1
  MAJOR_CIRC = P->surface_length; MINOR_CIRC = P->surface_width;
#elif REF_KIND == 2
  SURFACELENGTH = P->surface_length; SURFACEWIDTH = P->surface_width;
  XMIN = 0.0; XMAX = SURFACEWIDTH - XMIN; YMIN = 0.0; YMAX = SURFACELENGTH - YMIN;   // FHNmodel_flat.cpp:173-176
#else
  SURFACE_LENGTH = P->surface_length; SURFACE_WIDTH = P->surface_width;
  XMIN = 0.0; XMAX = SURFACE_WIDTH - XMIN; YMIN = 0.0; YMAX = SURFACE_LENGTH - YMIN; // GoldbeterModel_flat.cpp:196-199
#endif
}

// fill UserData the way main() does (FHNmodel_torus.cpp:223-238)
int fill_udata(UserData *udata, const crd_oracle_params *P, int rank) {
  int flag = REF_INIT(udata);
  if (flag != 0) return flag;
  long int nx = P->nx, ny = P->ny;
#if REF_KIND == 0
  realtype r = MINORCIRC / (2.0 * PI), R = MAJORCIRC / (2.0 * PI);
  udata->R = R; udata->r = r;
#elif REF_KIND == 1
  realtype r = MINOR_CIRC / (2.0 * PI), R = MAJOR_CIRC / (2.0 * PI);
  udata->R = R; udata->r = r;
#endif
  udata->rank = rank;
  udata->nx = nx;
  udata->ny = ny;
  udata->Diff = P->diff;
  udata->dx = (XMAX - XMIN) / (1.0 * nx - 1.0);
  udata->dy = (YMAX - YMIN) / (1.0 * ny - 1.0);
  return REF_DECOMP(udata);
}

struct RankResult { int flag = 0; double seconds = 0; };

void rank_body(const crd_oracle_params *P, int rank, double t, const double *yg, double *ydotg, int reps,
               RankResult *res) {
  crdshim_mpi_bind(rank);
  UserData *udata = new UserData;
  int flag = fill_udata(udata, P, rank);
  if (flag != 0) { res->flag = flag; delete udata; return; }
  long nxl = udata->nxl, nyl = udata->nyl;
  long N = 2 * nxl * nyl, Ntot = 2 * udata->nx * udata->ny;
  N_Vector y = N_VNew_Parallel(udata->comm, N, Ntot);
  N_Vector ydot = N_VNew_Parallel(udata->comm, N, Ntot);
  double *yl = NV_DATA_P(y), *ydl = NV_DATA_P(ydot);
  for (long j = 0; j < nyl; ++j)
    std::memcpy(yl + 2 * j * nxl, yg + 2 * ((udata->js + j) * udata->nx + udata->is), sizeof(double) * 2 * nxl);
  MPI_Barrier(MPI_COMM_WORLD);
  auto t0 = std::chrono::steady_clock::now();
  for (int it = 0; it < reps; ++it) {
    flag = f(t, y, ydot, (void *)udata);
    if (flag != 0) break;
  }
  MPI_Barrier(MPI_COMM_WORLD);
  auto t1 = std::chrono::steady_clock::now();
  res->flag = flag;
  res->seconds = std::chrono::duration<double>(t1 - t0).count();
  if (ydotg)
    for (long j = 0; j < nyl; ++j)
      std::memcpy(ydotg + 2 * ((udata->js + j) * udata->nx + udata->is), ydl + 2 * j * nxl, sizeof(double) * 2 * nxl);
  N_VDestroy_Parallel(y);
  N_VDestroy_Parallel(ydot);
  REF_FREE(udata);
  delete udata;
}

}  // namespace

extern "C" {

int crd_ref_kind(void) { return REF_KIND; }

// Run the reference f() `reps` times on the global state y (AoS [ny][nx][2]) decomposed over `nranks`
// emulated MPI ranks (MPI_Dims_create decides the 2-D layout, as in the reference).  ydot (may be
// NULL) receives the gathered result of the last evaluation; *seconds the wall time of the reps
// (slowest rank).  Returns the reference's flag (0 = success).
int crd_ref_rhs(const crd_oracle_params *P, int nranks, double t, const double *y, double *ydot, int reps,
                double *seconds) {
  if (nranks < 1) nranks = 1;
  set_globals(P);
  crdshim_mpi_set_world(nranks);
  std::vector<RankResult> res(nranks);
  if (nranks == 1) {
    rank_body(P, 0, t, y, ydot, reps, &res[0]);
  } else {
    std::vector<std::thread> th;
    for (int r = 0; r < nranks; ++r) th.emplace_back(rank_body, P, r, t, y, ydot, reps, &res[r]);
    for (auto &x : th) x.join();
  }
  double s = 0;
  int flag = 0;
  for (auto &r : res) { if (r.seconds > s) s = r.seconds; if (r.flag != 0) flag = r.flag; }
  if (seconds) *seconds = s;
  return flag;
}

// The reference's f() on a band of phi rows of a larger global mesh (what one rank of a 1-D phi split computes): UserData as
// main() + SetupDecomp() fill it for one rank, then js / je / nyl set to the band; the rows outside the band come from the
// caller through the shim's S / N receives.  Exchange(), the stencil, the frozen-row and beta(phi) logic run unmodified.
int crd_ref_rhs_band(const crd_oracle_params *P, double t, long j0, long nrows, const double *yband, double *out) {
  if (nrows < 2 || j0 < 0 || j0 + nrows > P->ny) return -2;   // the reference's face / corner code needs nyl >= 2
  set_globals(P);
  crdshim_mpi_set_world(1);
  crdshim_mpi_bind(0);
  UserData *udata = new UserData;
  int flag = fill_udata(udata, P, 0);   // buffers sized for the whole mesh: large enough for any band
  if (flag != 0) { delete udata; return flag; }
  udata->js = j0; udata->je = j0 + nrows - 1; udata->nyl = nrows;
  const long nx = udata->nx, N = 2 * nx * nrows;
  N_Vector y = N_VNew_Parallel(udata->comm, N, 2 * nx * udata->ny);
  N_Vector ydot = N_VNew_Parallel(udata->comm, N, 2 * nx * udata->ny);
  std::memcpy(NV_DATA_P(y), yband + 2 * nx, sizeof(double) * N);
  crdshim_mpi_band_halo(yband, yband + 2 * nx * (nrows + 1));
  flag = f(t, y, ydot, (void *)udata);
  crdshim_mpi_band_halo(nullptr, nullptr);
  std::memcpy(out, NV_DATA_P(ydot), sizeof(double) * N);
  N_VDestroy_Parallel(y);
  N_VDestroy_Parallel(ydot);
  REF_FREE(udata);
  delete udata;
  return flag;
}

// The reference's extents for rank `rank` of `nranks` (SetupDecomp): out = {is, ie, js, je, nxl, nyl, dims0, dims1}
int crd_ref_decomp(const crd_oracle_params *P, int nranks, int rank, long out[8]) {
  set_globals(P);
  crdshim_mpi_set_world(nranks);
  crdshim_mpi_bind(rank);
  UserData *udata = new UserData;
  int flag = fill_udata(udata, P, rank);
  int dims[2], periods[2], coords[2];
  MPI_Cart_get(udata->comm, 2, dims, periods, coords);
  out[0] = udata->is; out[1] = udata->ie; out[2] = udata->js; out[3] = udata->je;
  out[4] = udata->nxl; out[5] = udata->nyl; out[6] = dims[0]; out[7] = dims[1];
  REF_FREE(udata);
  delete udata;
  crdshim_mpi_bind(0);
  return flag;
}

// The reference's whole program (ini -> integrate -> text files), run with argv = {exe, ini_path}.
// ARKode behind it is the product's own explicit RK driver, MPI and Boost are the shims.
int crd_ref_main(const char *ini_path, int nranks) {
  crdshim_mpi_set_world(nranks < 1 ? 1 : nranks);
  auto body = [&](int rank, int *rc) {
    crdshim_mpi_bind(rank);
    char a0[] = "crd_ref";
    std::vector<char> a1(ini_path, ini_path + std::strlen(ini_path) + 1);
    char *argv[] = {a0, a1.data(), nullptr};
    *rc = ref_main(2, argv);
  };
  std::vector<int> rc(nranks < 1 ? 1 : nranks, 0);
  if (nranks <= 1) body(0, &rc[0]);
  else {
    std::vector<std::thread> th;
    for (int r = 0; r < nranks; ++r) th.emplace_back(body, r, &rc[r]);
    for (auto &x : th) x.join();
  }
  for (int v : rc) if (v != 0) return v;
  return 0;
}

}  // extern "C"

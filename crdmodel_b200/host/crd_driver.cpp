// crd_driver.cpp — host driver of the four programs (compiled once per model with -DCRD_DRIVER_MODEL=0..3
// into bin/FHNmodel_torus, bin/GoldbeterModel_torus, bin/FHNmodel_flat, bin/GoldbeterModel_flat; the names
// of the reference's CMake targets, CMakeLists.txt:37-50).
//
// Drop-in for `mpirun -np P <exe> <ini>` (util/ShellScripts/run*.sh:6): same single argument, same ini keys,
// same banner, same per-subdomain text files
//     <Model>_<surface>_subdomain.RRR.txt   "nx ny is ie js je XMIN XMAX TFINAL"
//     <Model>_<surface>_{u|Z}.RRR.txt       one line per output time, " %.16e" per point, j outer / i inner
//     <Model>_<surface>_{v|Y}.RRR.txt       created always, written when includeAllVars = 1
// (reference main(): src/FHNmodel_torus.cpp:148-497 and the three siblings), but the state lives on B200s:
// device N_Vector, fused stencil+reaction kernel, explicit RK driver through the ARKode-legacy names.
// Ranks are GPUs: `System.gpus = G` (or CRD_GPUS=G) forks G worker processes, one per GPU, each owning a
// phi slab; neighbours' boundary rows travel through CUDA-IPC peer mappings, norms through shared memory.
// New optional keys (absent => the reference's behaviour): System.gpus, System.arith (exact|fast), System.deviceAllreduce (1),
// System.haloTimeout (seconds),
// System.fused (1), System.reuseFirstStage (= fused), System.resident (1), Parameters.phiMesh, Parameters.Zs / Ys,
// System.steadyStateCommand (the reference's SolveGoldbeterODE.py protocol instead of the closed-form steady state).
#include <pthread.h>
#include <sys/mman.h>
#include <unistd.h>


#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <string>
#include <vector>

#include "crd_b200.h"
#include "crd_ini.hpp"
#include "crd_steady.hpp"
#include "crd_workers.hpp"
#include "crd_writer.hpp"

#ifndef CRD_DRIVER_MODEL
#define CRD_DRIVER_MODEL 0
#endif

using std::cerr;
using std::cout;

namespace {

constexpr int kModel = CRD_DRIVER_MODEL;
constexpr bool kTorus = (kModel == CRD_FHN_TORUS || kModel == CRD_GOLDBETER_TORUS);
constexpr bool kFhn = (kModel == CRD_FHN_TORUS || kModel == CRD_FHN_FLAT);
constexpr double PI = 3.1415926535897932;
constexpr int kMaxRanks = 64;

const char *file_stem() {
  switch (kModel) {
    case CRD_FHN_TORUS: return "FHNmodel_torus";
    case CRD_GOLDBETER_TORUS: return "GoldbeterModel_torus";
    case CRD_FHN_FLAT: return "FHNmodel_flat";
    default: return "GoldbeterModel_flat";
  }
}

// check_flag of the reference (FHNmodel_torus.cpp:681-705)
int check_flag(void *flagvalue, const std::string &funcname, int opt) {
  if (opt == 0 && flagvalue == NULL) {
    cerr << "\nSUNDIALS_ERROR: " << funcname << " failed - returned NULL pointer\n\n";
    return 1;
  } else if (opt == 1) {
    int *errflag = (int *)flagvalue;
    if (*errflag < 0) {
      cerr << "\nSUNDIALS_ERROR: " << funcname << " failed with flag = " << *errflag << "\n\n";
      return 1;
    }
  } else if (opt == 2 && flagvalue == NULL) {
    cerr << "\nMEMORY_ERROR: " << funcname << " failed - returned NULL pointer\n\n";
    return 1;
  }
  return 0;
}

// ---- ranks = worker processes sharing one anonymous mapping ------------------------------------------------
struct Shared {
  pthread_barrier_t bar;
  unsigned char handle[kMaxRanks][CRD_HALO_HANDLE_BYTES];
  unsigned char comm[kMaxRanks][CRD_HALO_HANDLE_BYTES];
  double red[kMaxRanks][3 * kMaxRanks];
  int failed;
};
Shared *g_shm = nullptr;
int g_rank = 0, g_nranks = 1;

int shm_allreduce(double *vals, int n, int op, void *) {
  if (g_nranks == 1) return 0;
  if (n > 3 * kMaxRanks) return -1;
  for (int i = 0; i < n; ++i) g_shm->red[g_rank][i] = vals[i];
  pthread_barrier_wait(&g_shm->bar);
  for (int i = 0; i < n; ++i) {
    double acc = g_shm->red[0][i];
    for (int r = 1; r < g_nranks; ++r) {
      const double v = g_shm->red[r][i];
      acc = (op == CRD_SUM) ? acc + v : (op == CRD_MAX) ? (v > acc ? v : acc) : (v < acc ? v : acc);
    }
    vals[i] = acc;
  }
  pthread_barrier_wait(&g_shm->bar);
  return 0;
}

struct Config {
  double DIFF, BETA, SURFACE_LENGTH, SURFACE_WIDTH, WAVE_LENGTH, WAVE_WIDTH, T_BOUNDARY, T_FINAL, BETA_MIN = 0, BETA_MAX = 0;
  int WAVE_INSIDE = 0, OUTPUT_TIMESTEP, NX, INCLUDE_ALL_VARS, VARY_BETA, JUST_DIFFUSION = 0, IC_TYPE = 0;
  long ny;
  int gpus, arith, fused, reuse, resident, dev_allreduce;
  double halo_timeout_s = 0;
  double Zs = 0, Ys = 0;
  bool have_zs = false;
};

Config read_config(const char *path) {
  crd::Ini pt(path);
  Config c;
  // the keys each reference program reads
  c.DIFF = pt.get<double>("Parameters.diffusion");
  c.BETA = pt.get<double>("Parameters.beta");
  c.SURFACE_LENGTH = pt.get<double>("Parameters.surfaceLength");
  c.SURFACE_WIDTH = pt.get<double>("Parameters.surfaceWidth");
  c.WAVE_LENGTH = pt.get<double>("Parameters.waveLength");
  c.WAVE_WIDTH = pt.get<double>("Parameters.waveWidth");
  if (kTorus) c.WAVE_INSIDE = pt.get<int>("Parameters.waveInside");
  c.OUTPUT_TIMESTEP = pt.get<int>("Parameters.outputTimestep");
  c.T_BOUNDARY = pt.get<double>("Parameters.tBoundary");
  c.T_FINAL = pt.get<double>("Parameters.tFinal");
  if (kFhn) {
    // the FHN programs read thetaMesh (FHNmodel_torus.cpp:170) while the shipped data/FHNmodelArgs.ini has xMesh
    c.NX = pt.has("Parameters.thetaMesh") ? pt.get<int>("Parameters.thetaMesh") : pt.get<int>("Parameters.xMesh");
  } else {
    c.NX = pt.get<int>("Parameters.xMesh");
  }
  if (kFhn || kModel == CRD_GOLDBETER_FLAT) {   // GoldbeterModel_torus.cpp never reads betaMin/betaMax (stay 0)
    c.BETA_MIN = pt.get<double>("Parameters.betaMin");
    c.BETA_MAX = pt.get<double>("Parameters.betaMax");
  }
  c.INCLUDE_ALL_VARS = pt.get<int>("System.includeAllVars");
  c.VARY_BETA = pt.get<int>("System.varyBeta");
  if (!kFhn) c.JUST_DIFFUSION = pt.get<int>("System.justDiffusion");
  if (kModel == CRD_GOLDBETER_FLAT) c.IC_TYPE = pt.get<int>("System.icType");  // the torus program leaves it 0
  // mesh in phi / y exactly as main() derives it (FHNmodel_torus.cpp:188-193, FHNmodel_flat.cpp:190-192)
  if (kTorus) {
    const double r = c.SURFACE_WIDTH / (2.0 * PI), R = c.SURFACE_LENGTH / (2.0 * PI);
    const double radiusRatio = R / r;
    c.ny = (long)(c.NX * (radiusRatio));
  } else {
    const long ratio = (long)(c.SURFACE_LENGTH / c.SURFACE_WIDTH);
    c.ny = c.NX * ratio;
  }
  // extensions
  c.ny = pt.get<long>("Parameters.phiMesh", c.ny);
  const char *env = std::getenv("CRD_GPUS");
  c.gpus = pt.get<int>("System.gpus", env ? std::atoi(env) : 1);
  if (c.gpus < 1) c.gpus = 1;
  const std::string ar = pt.has("System.arith") ? pt.str("System.arith") : "exact";
  c.arith = (ar == "fast") ? CRD_ARITH_FAST : CRD_ARITH_EXACT;
  c.fused = pt.get<int>("System.fused", 1);
  c.reuse = pt.get<int>("System.reuseFirstStage", c.fused ? 1 : 0);   // f(tn, yn) is already there from the previous step: same bits
  c.dev_allreduce = pt.get<int>("System.deviceAllreduce", 1);
  c.halo_timeout_s = pt.get<double>("System.haloTimeout", 0.0);   // seconds a rank waits for a neighbour's rows (0: the library's 30 s)
  c.resident = pt.get<int>("System.resident", 1);   // 1: the step loop runs as one persistent kernel when it applies
  if (!kFhn) {
    if (pt.has("Parameters.Zs") && pt.has("Parameters.Ys")) {
      c.Zs = pt.get<double>("Parameters.Zs"); c.Ys = pt.get<double>("Parameters.Ys"); c.have_zs = true;
    } else if (pt.has("System.steadyStateCommand")) {
      // the reference's protocol (GoldbeterModel_torus.cpp:254-261): `<command> <beta>` prints "[Zs] [Ys]"
      const std::string cmd = pt.str("System.steadyStateCommand");
      if (!crd::goldbeter_steady_state_from_command(cmd, pt.str("Parameters.beta"), c.Zs, c.Ys)) {
        cerr << "steady state: `" << cmd << " " << pt.str("Parameters.beta") << "` did not print \"[Zs] [Ys]\"; using the closed form\n";
        crd::goldbeter_steady_state(c.BETA, c.Zs, c.Ys);
      }
    } else {
      crd::goldbeter_steady_state(c.BETA, c.Zs, c.Ys);
    }
  }
  return c;
}

int run(const Config &c, int rank, int nranks) {
  g_rank = rank; g_nranks = nranks;
  time_t start_t = 0, end_t = 0;
  double total_t = 0, eta = 0;
  time(&start_t);

  const double T0 = 0.0, Tf = c.T_FINAL;
  const int Nt = c.OUTPUT_TIMESTEP;
  const long nx = c.NX, ny = c.ny;
  const double rtol = 1.e-5, atol = 1.e-10;
  int flag;

  if (nranks > 1 && nranks > crd_device_count()) {
    // ranks spin on each other's halo flags: never co-schedule two of them on one GPU
    if (rank == 0) cerr << "\nCUDA_ERROR: System.gpus = " << nranks << " but only " << crd_device_count() << " GPU(s) visible\n\n";
    return 1;
  }
  crd_ctx *ctx = crd_ctx_create(rank, NULL);
  if (!ctx) { cerr << "\nCUDA_ERROR: " << crd_last_error() << "\n\n"; return 1; }
  flag = crd_ctx_set_comm(ctx, rank, nranks, nranks > 1 ? shm_allreduce : NULL, NULL);
  if (check_flag(&flag, "crd_ctx_set_comm", 1)) return 1;
  if (c.halo_timeout_s > 0) crd_ctx_set_halo_timeout(ctx, 1e3 * c.halo_timeout_s);

  crd_params p;
  std::memset(&p, 0, sizeof p);
  p.model = kModel; p.arith = c.arith; p.nx = nx; p.ny = ny;
  flag = crd_decomp_phi(ny, nranks, rank, &p.js, &p.je);
  if (flag != 0) { cerr << "SetupDecomp: " << crd_last_error() << "\n"; return 1; }
  p.diff = c.DIFF; p.beta = c.BETA; p.beta_min = c.BETA_MIN; p.beta_max = c.BETA_MAX; p.vary_beta = c.VARY_BETA;
  p.just_diffusion = c.JUST_DIFFUSION; p.t_boundary = c.T_BOUNDARY; p.surface_length = c.SURFACE_LENGTH;
  p.surface_width = c.SURFACE_WIDTH;
  crd_grid *grid = crd_grid_create(ctx, &p);
  if (check_flag((void *)grid, "crd_grid_create", 2)) { cerr << crd_last_error() << "\n"; return 1; }
  const long nxl = nx, nyl = p.je - p.js + 1;
  if (nranks > 1) {
    flag = crd_grid_halo_handle(grid, g_shm->handle[rank]);
    if (check_flag(&flag, "crd_grid_halo_handle", 1)) return 1;
    flag = crd_ctx_comm_handle(ctx, g_shm->comm[rank]);
    if (check_flag(&flag, "crd_ctx_comm_handle", 1)) return 1;
    pthread_barrier_wait(&g_shm->bar);
    flag = crd_grid_halo_connect_ipc(grid, g_shm->handle[(rank + nranks - 1) % nranks], g_shm->handle[(rank + 1) % nranks]);
    if (flag != 0) { cerr << "halo connect: " << crd_last_error() << "\n"; return 1; }
    // the integrator's norms are finished on the devices (mailboxes over NVLink); System.deviceAllreduce = 0 keeps the
    // shared-memory hook of crd_ctx_set_comm instead
    if (c.dev_allreduce) {
      flag = crd_ctx_comm_connect_ipc(ctx, rank, nranks, &g_shm->comm[0][0]);
      if (flag != 0) { cerr << "allreduce connect: " << crd_last_error() << "\n"; return 1; }
    }
    pthread_barrier_wait(&g_shm->bar);
  }

  // steady states (FHNmodel_torus.cpp:242-244 | GoldbeterModel_torus.cpp:254-261)
  double S0, S1;
  if (kFhn) { S0 = -c.BETA; S1 = c.BETA * c.BETA * c.BETA - 3 * c.BETA; }
  else { S0 = c.Zs; S1 = c.Ys; }

  const bool outproc = (rank == 0);
  if (outproc) {
    if (kModel == CRD_FHN_TORUS || kModel == CRD_FHN_FLAT) cout << "\n2D FHN model PDE problem on a torus:\n";
    else if (kModel == CRD_GOLDBETER_TORUS) cout << "\n Goldbeter model PDE problem on a torus:\n";
    else cout << "\n2D Goldbeter model PDE problem on a flat surface:\n";
    cout << "   nprocs = " << nranks << "\n";
    cout << "   nx = " << nx << "\n";
    cout << "   ny = " << ny << "\n";
    cout << "   nxl = " << nxl << "\n";
    cout << "   nyl = " << nyl << "\n";
    cout << "   Diff = " << c.DIFF << "\n";
    cout << "   Tfinal = " << c.T_FINAL << "\n";
    cout << "   Output timesteps = " << c.OUTPUT_TIMESTEP << "\n";
    if (kTorus) {
      cout << "   Major circumference = " << c.SURFACE_LENGTH << "\n";
      cout << "   Minor circumference = " << c.SURFACE_WIDTH << "\n";
    } else {
      cout << "   Surface length = " << c.SURFACE_LENGTH << "\n";
      cout << "   Surface width = " << c.SURFACE_WIDTH << "\n";
    }
    if (kFhn) cout << "   Absorbing boundary turn off time = " << c.T_BOUNDARY << "\n";
    cout << "   Wavelength = " << c.WAVE_LENGTH * 100 << "%\n";
    cout << "   Wavewidth = " << c.WAVE_WIDTH * 100 << "%\n";
    if (kModel == CRD_FHN_TORUS) cout << "   Wave inside = " << c.WAVE_INSIDE << "\n";
    cout << "   rtol = " << rtol << "\n";
    cout << "   atol = " << atol << "\n";
    if (kFhn) {
      cout << "   Include all variables in output = " << c.INCLUDE_ALL_VARS << "\n";
      if (c.VARY_BETA == 0) {
        cout << "   Beta = " << c.BETA << "\n";
        cout << "   Stable state values: U = " << S0 << ", V = " << S1 << "\n\n";
      } else {
        cout << "   Beta varied over torus\n\n";
      }
    } else if (c.JUST_DIFFUSION == 1) {
      cout << "   Diffusion Only\n\n";
    } else {
      cout << "   Include all variables in output = " << c.INCLUDE_ALL_VARS << "\n";
      cout << "   Absorbing boundary turn off time = " << c.T_BOUNDARY << "\n";
      if (c.VARY_BETA == 0) {
        cout << "   Beta = " << c.BETA << "\n";
        cout << "   Stable state values: Z = " << S0 << ", Y = " << S1 << "\n\n";
      } else if (c.VARY_BETA == 1) {
        cout << (kTorus ? "   Beta varied over torus\n" : "   Beta varied over surface\n");
        if (c.IC_TYPE == 0) cout << "   Homogeneous ICs\n\n";
        if (c.IC_TYPE == 1) cout << "   ICs: initial perturbation\n\n";
        if (c.IC_TYPE == 2) cout << "   Random ICs\n\n";
      }
    }
    cout.flush();
  }

  // state vector on the device (replaces N_VNew_Parallel, :281)
  const long N = 2 * nxl * nyl, Ntot = 2 * nx * ny;
  N_Vector y = N_VNew_Crd(ctx, N, Ntot);
  if (check_flag((void *)y, "N_VNew_Crd", 0)) { cerr << crd_last_error() << "\n"; return 1; }
  realtype *ydata = NULL;   // the reference caches N_VGetArrayPointer(y) here (:383); only the host-generated ICs need a host array

  if (kTorus && c.WAVE_INSIDE != 0 && c.WAVE_INSIDE != 1) printf("WaveInside must be 0 or 1");
  if (!kFhn && c.VARY_BETA == 1 && c.IC_TYPE == 2) {
    // icType 2: unseeded rand() through (float), identical on every rank (GoldbeterModel_flat.cpp:373-374)
    ydata = N_VGetArrayPointer(y);     // pinned host mirror
    if (check_flag((void *)ydata, "N_VGetArrayPointer", 0)) return 1;
    for (long j = 0; j < ny; ++j)
      for (long i = 0; i < nx; ++i) {
        const double a = (float)rand() / RAND_MAX * 1.4, b = (float)rand() / RAND_MAX * 1.4;
        if (j >= p.js && j <= p.je) { ydata[2 * (i + (j - p.js) * nx)] = a; ydata[2 * (i + (j - p.js) * nx) + 1] = b; }
      }
    flag = N_VCopyFromHost_Crd(y);
    if (check_flag(&flag, "N_VCopyFromHost_Crd", 1)) return 1;
  } else {
    crd_ic_params ic;
    ic.wave_length = c.WAVE_LENGTH; ic.wave_width = c.WAVE_WIDTH; ic.wave_inside = c.WAVE_INSIDE; ic.ic_type = c.IC_TYPE;
    ic.s0 = S0; ic.s1 = S1;
    flag = crd_fill_initial_conditions(grid, &ic, N_VGetDeviceArrayPointer_Crd(y));
    if (flag != 0) { cerr << crd_last_error() << "\n"; return 1; }
  }

  void *arkode_mem = ARKodeCreate();
  if (check_flag((void *)arkode_mem, "ARKodeCreate", 0)) return 1;
  flag = ARKodeInit(arkode_mem, crd_f, NULL, T0, y);
  if (check_flag(&flag, "ARKodeInit", 1)) return 1;
  flag = ARKodeSStolerances(arkode_mem, rtol, atol);
  if (check_flag(&flag, "ARKodeSStolerances", 1)) return 1;
  flag = ARKodeSetUserData(arkode_mem, (void *)grid);
  if (check_flag(&flag, "ARKodeSetUserData", 1)) return 1;
  flag = ARKodeSetMaxNumSteps(arkode_mem, 200000);
  if (check_flag(&flag, "ARKodeSetMaxNumSteps", 1)) return (1);
  if (c.fused) crd_ARKodeSetFusedOps(arkode_mem, crd_nv_fused_ops_for(grid));   // EXACT grid: the bits of the op-by-op sequence
  crd_ARKodeSetReuseFirstStage(arkode_mem, c.reuse);
  crd_ARKodeSetResident(arkode_mem, c.resident);

  // per-subdomain output files (:375-410)
  const char *stem = file_stem();
  const char *var0 = kFhn ? "u" : "Z", *var1 = kFhn ? "v" : "Y";
  const double XMIN = 0.0, XMAX = kTorus ? 2.0 * PI : c.SURFACE_WIDTH - XMIN;
  char outname[200];
  snprintf(outname, sizeof outname, "%s_subdomain.%03i.txt", stem, rank);
  FILE *UFID = fopen(outname, "w");
  if (!UFID) { cerr << "cannot open " << outname << "\n"; return 1; }
  fprintf(UFID, "%li  %li  %li  %li  %li  %li %f %f %f\n", nx, ny, 0L, nx - 1, (long)p.js, (long)p.je, XMIN, XMAX, c.T_FINAL);
  fclose(UFID);
  snprintf(outname, sizeof outname, "%s_%s.%03i.txt", stem, var0, rank);
  UFID = fopen(outname, "w");
  snprintf(outname, sizeof outname, "%s_%s.%03i.txt", stem, var1, rank);
  FILE *UFID2 = fopen(outname, "w");
  if (!UFID || !UFID2) { cerr << "cannot open output files\n"; return 1; }

  // output off the critical path (crd_writer.hpp): an output is enqueued — gather of the written variables on the device,
  // asynchronous copy into one of three page-locked buffers — and the time loop goes on; a background thread formats
  const bool all_vars = c.INCLUDE_ALL_VARS == 1;
  crd_snapshot *snap = crd_snapshot_create(ctx, (int64_t)nxl * nyl, all_vars ? 2 : 1, 3);
  if (check_flag((void *)snap, "crd_snapshot_create", 2)) { cerr << crd_last_error() << "\n"; return 1; }
  crd::AsyncWriter writer(UFID, UFID2, all_vars, nxl * nyl);
  auto write_state = [&]() -> int {
    int slot;
    while ((slot = crd_snapshot_begin(snap, N_VGetDeviceArrayPointer_Crd(y))) == -2) usleep(200);   // every buffer still being formatted
    if (slot < 0) { cerr << crd_last_error() << "\n"; return 1; }
    writer.submit([snap, slot]() {
      crd::OutputView v;
      if (crd_snapshot_wait(snap, slot, &v.v0, &v.v1) != 0) { v.v0 = nullptr; }
      v.release = [snap, slot]() { crd_snapshot_release(snap, slot); };
      return v;
    });
    return 0;
  };
  if (write_state()) return 1;

  realtype t = T0;
  const realtype dTout = (Tf - T0) / Nt;
  realtype tout = T0 + dTout;
  for (int iout = 0; iout < Nt; iout++) {
    flag = ARKode(arkode_mem, tout, y, &t, ARK_NORMAL);
    if (check_flag(&flag, "ARKode", 1)) break;
    if (flag >= 0) {
      tout += dTout;
      tout = (tout > Tf) ? Tf : tout;
    } else {
      if (outproc) cerr << "Solver failure, stopping integration\n";
      break;
    }
    if (write_state()) return 1;
    // the ranks meet on the host before the next interval: whatever one of them spent here (a full output queue, a slow
    // disk) cannot turn into a neighbour's halo timeout during the next evaluations
    if (nranks > 1) pthread_barrier_wait(&g_shm->bar);
    time(&end_t);
    total_t += difftime(end_t, start_t);
    start_t = end_t;
    eta = (Nt - (iout + 1)) * (total_t / (iout + 1));
    if (outproc) {
      if (iout > 0) for (int b = 0; b < 61; ++b) putchar('\b');
      printf("   %3d %% | %3d min %2d sec elapsed | %3d min %2d sec remaining", 100 * (iout + 1) / Nt, (int)(total_t / 60),
             ((int)total_t % 60), (int)(eta / 60), ((int)eta % 60));
      fflush(stdout);
    }
  }
  writer.finish();
  if (writer.failed()) { cerr << "output writer: " << crd_last_error() << "\n"; flag = -1; }
  crd_snapshot_destroy(snap);
  if (outproc) cout << "\n   ----------------------\n";
  fclose(UFID);
  fclose(UFID2);

  // integrator statistics (the reference never queries them; the north-star asks for them)
  long nst = 0, nfe = 0, nfi = 0, netf = 0, natt = 0;
  ARKodeGetNumSteps(arkode_mem, &nst);
  ARKodeGetNumRhsEvals(arkode_mem, &nfe, &nfi);
  ARKodeGetNumErrTestFails(arkode_mem, &netf);
  ARKodeGetNumStepAttempts(arkode_mem, &natt);
  if (outproc)
    cout << "   Steps = " << nst << " (attempts " << natt << ", error-test failures " << netf << "), RHS evaluations = " << nfe
         << ", kernels launched = " << (long)crd_ctx_launch_count(ctx) << "\n";

  N_VDestroy(y);
  ARKodeFree(&arkode_mem);
  crd_grid_destroy(grid);
  crd_ctx_destroy(ctx);
  return flag < 0 ? 1 : 0;
}

}  // namespace

int main(int argc, char *argv[]) {
  if (argc != 2) {
    std::cerr << "Usage: " << argv[0] << " <Config file path>";
    exit(EXIT_FAILURE);
  }
  Config c;
  try {
    c = read_config(argv[1]);
  } catch (const std::exception &e) {
    // Boost would throw out of main() here (uncaught -> abort); report and fail instead
    std::cerr << "terminate called after throwing an instance of 'ptree_error'\n  what():  " << e.what() << "\n";
    return 134;
  }
  int nranks = c.gpus;
  if (nranks > kMaxRanks) nranks = kMaxRanks;
  if (nranks == 1) return run(c, 0, 1);

  // one worker process per GPU; rendezvous through an anonymous shared mapping (no MPI launcher needed)
  g_shm = (Shared *)mmap(NULL, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (g_shm == MAP_FAILED) { perror("mmap"); return 1; }
  pthread_barrierattr_t ba;
  pthread_barrierattr_init(&ba);
  pthread_barrierattr_setpshared(&ba, PTHREAD_PROCESS_SHARED);
  pthread_barrier_init(&g_shm->bar, &ba, nranks);
  g_shm->failed = 0;
  return crd::run_ranks(nranks, [&](int r) { return run(c, r, nranks); });
}

/* crd_b200.h — C ABI of libcrd_b200.so: the B200 (sm_100a) implementation of CRDModel's
 * time-stepping hot path.  Plain C: pointers, sizes, opaque handles; no CUDA or torch types.
 *
 * What each group replaces in the reference (BlueFern/CRDModel, paths relative to its root):
 *   crd_params / crd_grid_*   the UserData struct + the file-scope ini parameters that reach f()
 *                             src/FHNmodel_torus.cpp:80-122,223-238 ; SetupDecomp :708-772
 *   crd_decomp_phi            SetupDecomp's extent formula :750-755 with dims = {1, nranks}
 *   crd_rhs*, crd_f           the ARKRhsFn callback f()            src/FHNmodel_torus.cpp:504-667
 *                                                                  src/GoldbeterModel_torus.cpp:547-724
 *                                                                  src/FHNmodel_flat.cpp:469-616
 *                                                                  src/GoldbeterModel_flat.cpp:515-689
 *   crd_grid_halo_*           Exchange() (MPI_Isend/Irecv of boundary rows) :775-950 — here: the
 *                             boundary rows are pushed into the neighbour GPU's ghost buffer through
 *                             peer-mapped memory over NVLink, ordered by an epoch flag
 *   N_VNew_Crd, N_V*_Crd      SUNDIALS' nvector_parallel (N_VNew_Parallel :281, N_VConst :506,
 *                             NV_DATA_P :517-518, N_VGetArrayPointer :303,383,786,
 *                             N_VDestroy_Parallel :488) and every op ARKode calls inside :423
 *   crd_nv_fused_ops          (no counterpart in SUNDIALS 2.x) single-pass stage assembly and
 *                             solution+error+norm, used by the explicit RK driver (crd_ark.h)
 * All functions returning int give 0 on success and a negative value on failure (crd_last_error()
 * holds the message); pointer-returning functions give NULL on failure.  There is no CPU fallback:
 * every entry point needs a CUDA device of compute capability 10.x.
 *
 * Threading: a crd_ctx and everything created from it must be used from one host thread at a time.
 * All device work is ordered on the context's stream; only reductions, host copies and crd_ctx_sync
 * wait for it.
 */
#ifndef CRD_B200_H
#define CRD_B200_H

#include <stddef.h>
#include <stdint.h>

#include "crd_ark.h"
#include "crd_sundials_compat.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- models and arithmetic --------------------------------------------------------------------- */
enum { CRD_FHN_TORUS = 0, CRD_GOLDBETER_TORUS = 1, CRD_FHN_FLAT = 2, CRD_GOLDBETER_FLAT = 3 };
/* CRD_ARITH_EXACT: the reference's expression tree, every operation rounded separately (no FMA
 *   contraction, true divisions), metric coefficients from the host libm: FHN results are bit-identical
 *   to the reference's f(); Goldbeter differs only where libm pow(x,2|4) is not correctly rounded.
 * CRD_ARITH_FAST: coefficients folded per theta (D*a(theta)/(2dx) ...), FMA contraction allowed;
 *   within 1e-12 of the reference relative to the magnitude of the summed terms. */
enum { CRD_ARITH_EXACT = 0, CRD_ARITH_FAST = 1 };

/* Everything f() depends on.  Field meaning = the ini keys (data/FHNmodelArgs.ini) / UserData. */
typedef struct crd_params {
  int32_t model;            /* CRD_FHN_TORUS ... */
  int32_t arith;            /* CRD_ARITH_EXACT / CRD_ARITH_FAST */
  int64_t nx, ny;           /* global mesh: theta|x (fastest) and phi|y.  The reference ties
                               ny = nx*(R/r) (FHNmodel_torus.cpp:193); here it is set independently. */
  int64_t js, je;           /* global phi rows owned by this grid, inclusive (crd_decomp_phi) */
  double diff;              /* Parameters.diffusion */
  double beta;              /* Parameters.beta */
  double beta_min, beta_max;/* Parameters.betaMin / betaMax */
  int32_t vary_beta;        /* System.varyBeta */
  int32_t just_diffusion;   /* System.justDiffusion (Goldbeter programs) */
  double t_boundary;        /* Parameters.tBoundary */
  double surface_length;    /* Parameters.surfaceLength (major circumference | length) */
  double surface_width;     /* Parameters.surfaceWidth  (minor circumference | width) */
} crd_params;

/* ---- device context ---------------------------------------------------------------------------- */
typedef struct crd_ctx crd_ctx;
typedef struct crd_grid crd_grid;

/* Host-level reduction across the ranks of a multi-process run: combine vals[0..n) in place with
 * op (CRD_SUM / CRD_MAX / CRD_MIN) over all ranks; return 0.  Supplied by the host program (the C++
 * driver uses shared memory between its worker processes, Python uses torch.distributed). */
enum { CRD_SUM = 1, CRD_MAX = 2, CRD_MIN = 3 };
typedef int (*crd_allreduce_fn)(double *vals, int n, int op, void *user);

const char *crd_last_error(void);
int crd_device_count(void);
/* stream: a cudaStream_t owned by the caller to order all work on (e.g. torch's current stream),
 * or NULL to let the context create its own non-blocking stream. */
crd_ctx *crd_ctx_create(int device, void *stream);
void crd_ctx_destroy(crd_ctx *ctx);
int crd_ctx_set_comm(crd_ctx *ctx, int rank, int nranks, crd_allreduce_fn fn, void *user);
/* Device-side allreduce for the same purpose (the MPI_Allreduce inside nvector_parallel's reductions), without the host:
 * every context exports the handle of its mailbox block, the host program gathers them (handles: nranks x
 * CRD_HALO_HANDLE_BYTES, in rank order) and each rank connects to all.  From then on the finishing block of every reduction
 * stores its values into all ranks' mailboxes through the NVLink peer mappings, waits for everybody's, and combines them in
 * rank order — identical bits on every rank; the hook of crd_ctx_set_comm is no longer called.  All ranks must issue the same
 * sequence of reductions (they do: the integrator is the same program on every rank).  *_local: contexts of one process. */
int crd_ctx_comm_handle(crd_ctx *ctx, unsigned char handle[64]);
int crd_ctx_comm_connect_ipc(crd_ctx *ctx, int rank, int nranks, const unsigned char *handles);
int crd_ctx_comm_connect_local(crd_ctx *ctx, int rank, int nranks, crd_ctx *const *all);
void *crd_ctx_stream(crd_ctx *ctx);
int crd_ctx_device(crd_ctx *ctx);
int crd_ctx_sync(crd_ctx *ctx);
/* Device-side failures.  The only one: on a phi-split grid a neighbour's boundary rows did not arrive within the halo
 * timeout (default 30 s; CRD_HALO_TIMEOUT_MS in the environment, or this call), i.e. a rank died or stalled — the situation in
 * which the reference's MPI_Wait (src/FHNmodel_torus.cpp:904-946) would hang.  The evaluation that ran with stale rows is
 * reported at the next point the host waits for the stream (every reduction, host copy, crd_ctx_sync): that call fails, and
 * from then on EVERY entry point of the context fails (crd_f returns -1, so ARKode ends with ARK_RHSFUNC_FAIL) until
 * crd_ctx_clear_error.  crd_ctx_failed: 0, or the device's error code. */
int crd_ctx_set_halo_timeout(crd_ctx *ctx, double milliseconds);
int crd_ctx_failed(crd_ctx *ctx);
int crd_ctx_clear_error(crd_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t crd_ctx_launch_count(crd_ctx *ctx);
/* device-side elapsed time between two points of the context's stream (CUDA events) */
int crd_timer_start(crd_ctx *ctx);
int crd_timer_stop(crd_ctx *ctx, double *milliseconds); /* synchronises */

/* device / pinned-host memory */
void *crd_malloc(crd_ctx *ctx, size_t bytes);
int crd_free(crd_ctx *ctx, void *dev_ptr);
void *crd_malloc_host(size_t bytes); /* page-locked */
int crd_free_host(void *host_ptr);
int crd_memcpy_h2d(crd_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes); /* stream-ordered, waits */
int crd_memcpy_d2h(crd_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
int crd_memset_zero(crd_ctx *ctx, void *dst_dev, size_t bytes);
/* write `bytes` of junk (> L2) between timed iterations */
int crd_flush_l2(crd_ctx *ctx);

/* ---- decomposition and grid -------------------------------------------------------------------- */
/* js = ny*rank/nranks, je = ny*(rank+1)/nranks - 1  (SetupDecomp :752-753 with dims = {1, nranks}) */
int crd_decomp_phi(int64_t ny, int nranks, int rank, int64_t *js, int64_t *je);

crd_grid *crd_grid_create(crd_ctx *ctx, const crd_params *p);
void crd_grid_destroy(crd_grid *g);
int crd_grid_params(const crd_grid *g, crd_params *out);
int64_t crd_grid_local_length(const crd_grid *g); /* 2*nx*nyl */
int64_t crd_grid_global_length(const crd_grid *g);/* 2*nx*ny */
double crd_grid_dx(const crd_grid *g);
double crd_grid_dy(const crd_grid *g);

/* Halo ring over the phi split.  With one rank the stencil wraps inside the slab and none of this
 * is needed.  With several, every rank exports the handle of its ghost block, the host program
 * gathers them, and each rank connects to its two neighbours (prev owns the rows below js, next the
 * rows above je; periodic).  *_local connects grids living in the same process (any devices). */
#define CRD_HALO_HANDLE_BYTES 64
int crd_grid_halo_handle(crd_grid *g, unsigned char handle[CRD_HALO_HANDLE_BYTES]);
int crd_grid_halo_connect_ipc(crd_grid *g, const unsigned char prev_handle[CRD_HALO_HANDLE_BYTES],
                              const unsigned char next_handle[CRD_HALO_HANDLE_BYTES]);
int crd_grid_halo_connect_local(crd_grid *g, crd_grid *prev, crd_grid *next);

/* ---- right-hand side --------------------------------------------------------------------------- */
/* ydot = f(t, y); y, ydot: device arrays, fp64, interleaved [nyl][nx][2] like the reference's
 * IDX(x,y) = 2*x + 2*y*nxl (FHNmodel_torus.cpp:60).  = crd_rhs_post_halo + crd_rhs_compute. */
int crd_rhs(crd_grid *g, double t, const double *y_dev, double *ydot_dev);
/* phase 1: push this slab's first/last row into the neighbours' ghost buffers (never waits) */
int crd_rhs_post_halo(crd_grid *g, const double *y_dev);
/* phase 2: wait for the neighbours' rows of the same epoch, then the fused stencil+reaction kernel */
int crd_rhs_compute(crd_grid *g, double t, const double *y_dev, double *ydot_dev);
/* Same with HOST arrays: the slab is streamed host->device, evaluated, and streamed back in row
 * chunks on three streams so the copies overlap the kernel (single rank only). */
int crd_rhs_host(crd_grid *g, double t, const double *y_host, double *ydot_host);
/* ARKRhsFn: user_data is the crd_grid*, y and ydot are device N_Vectors (N_VNew_Crd).  Returns 0,
 * or -1 if the evaluation could not be issued (the reference returns -1 when Exchange fails, :522). */
int crd_f(realtype t, N_Vector y, N_Vector ydot, void *user_data);
/* ydot = f(t, sum_j c[j]*X[j]), n <= 5, without materialising the combination: the explicit RK stage assembly
 * (ARKode's N_VLinearSum chain before each stage, inside ARKode() :423) fused into the evaluation.  Rounding of the
 * combination: CRD_ARITH_FAST grids form c[0]*X[0] and add the other terms by fused multiply-adds in index order;
 * CRD_ARITH_EXACT grids form it like the op-by-op stage assembly, s = c[1]*X[1] + 0, s = c[j]*X[j] + s (j = 2..),
 * state = c[0]*X[0] + s, every product and sum rounded separately. */
int crd_rhs_lincomb(crd_grid *g, double t, int n, const double *c, const double *const *X_dev, double *ydot_dev);
int crd_f_lincomb(realtype t, int n, const realtype *c, N_Vector *X, N_Vector ydot, void *user_data);
/* The last stage of a 5-stage explicit RK step fused with the step finish (ARKode's stage evaluation + arkComputeSolutions
 * + the WRMS error norm inside ARKode() :423): X = (yn, F_0 .. F_3), F_4 = f(t, sum_j c[j] X[j]) is never stored,
 * ynew = yn + sum_j hb[j] F_j, err = sum_j hd[j] F_j, out[] = the two weighted square sums of N_VErkFinish_Crd.  Returns
 * 0, or 1 when it does not apply (mesh too small to stream, s != 5): issue the two operations separately.  On a phi-split
 * grid it posts the halo rows of the stage state like crd_rhs_lincomb and sums out[] over the ranks. */
int crd_rhs_lincomb_finish(crd_grid *g, double t, int s, const double *c, const double *hb, const double *hd,
                           const double *const *X_dev, double *ynew_dev, double rtol, double atol, double out[2]);
int crd_f_lincomb_finish(realtype t, int s, const realtype *c, const realtype *hb, const realtype *hd, N_Vector *X, N_Vector ynew,
                         realtype rtol, realtype atol, realtype out[2], void *user_data);
/* Two evaluations in one pass over the state: f1 = f(t1, y), f2 = f(t2, y + c f1) — the derivative at an accepted state
 * (ARKode evaluates it for its dense output; it is also stage 1 of the next step) together with that step's second stage,
 * whose state needs nothing but y and f1 (src/FHNmodel_torus.cpp:423: two of the f() calls inside ARKode()).  y is read once,
 * y + c f1 never exists in memory: 48 instead of 80 B per point.  f1 and f2 have the bits of crd_rhs followed by
 * crd_rhs_lincomb(2, (1, c), (y, f1)).  On a phi-split grid the ranks exchange two rows of y per side first (one exchange for
 * both evaluations; every rank must make the call, and every rank gets the same answer: the size test uses the global mesh and
 * the number of ranks).  Returns 0, or 1 when it does not apply (slabs too small to stream, a forced kernel variant): issue the
 * two evaluations separately. */
int crd_rhs_pair(crd_grid *g, double t1, double t2, double c, const double *y_dev, double *f1_dev, double *f2_dev);
/* the integrator's form (crd_fused_ops.rhs_pair); also declines (1) where the pass is not faster than the two launches:
 * CRD_ARITH_EXACT grids of the Goldbeter programs (bound by FP64 work) */
int crd_f_pair(realtype t1, realtype t2, realtype c, N_Vector y, N_Vector f1, N_Vector f2, void *user_data);
/* count of RHS evaluations issued on this grid */
int64_t crd_grid_rhs_count(const crd_grid *g);
/* kernel variant: 0 = default; others are experimental tilings kept for profiling */
int crd_grid_set_variant(crd_grid *g, int variant);
/* on (default): ring evaluations compute the interior rows while the boundary rows are exchanged */
int crd_grid_set_overlap(crd_grid *g, int on);
/* The adaptive step loop of the explicit integrator as one persistent cooperative kernel (replaces the time loop
 * inside ARKode(), src/FHNmodel_torus.cpp:423, for meshes that live in L2 such as the shipped 400 x 1600 and
 * 100 x 400 grids: no kernel launch and no host round trip per step).  user_data is the crd_grid*; this is the
 * erk_evolve slot of crd_nv_fused_ops() (crd_ark.h).  Returns 0 when it ran, > 0 when it does not apply (several
 * ranks, a mesh beyond the automatic size limit, a method wider than 5 stages), < 0 on failure. */
int crd_erk_evolve(struct crd_erk_state *st, void *user_data);
/* mode: 0 automatic (default: meshes of up to 1 Mi points, whose working set stays in L2), 1 whenever it applies, -1 never */
int crd_grid_set_resident(crd_grid *g, int mode);
/* how many times the resident loop was launched on this grid */
int64_t crd_grid_resident_launches(const crd_grid *g);
/* where the last resident launch spent its SM cycles, as seen by one CTA: phase 1 (stage state), interior rows,
 * grid-barrier wait, edge rows, everything else, total.  Counted only by the profiling instantiation of the kernel
 * (crd_grid_set_variant(g, 150): 14 more registers, 3 % slower); the default kernel reports the total only. */
int crd_grid_resident_cycles(const crd_grid *g, int64_t out[6]);

/* ---- synthetic states and initial conditions --------------------------------------------------- */
/* 64-bit LCG stream, element e of the global vector uses state e+1 after `seed`;
 * FHN 4u-2, Goldbeter 1.5u+0.1.  Fills out_dev[0..n) with elements first_elem .. first_elem+n-1. */
int crd_fill_synthetic(crd_ctx *ctx, int model, uint64_t seed, int64_t first_elem, int64_t n, double *out_dev);
/* The reference's initial conditions (FHNmodel_torus.cpp:285-354, GoldbeterModel_torus.cpp:313-414,
 * FHNmodel_flat.cpp:280-319, GoldbeterModel_flat.cpp:317-379) evaluated on the device for this slab.
 * s0, s1: steady state (Us,Vs | Zs,Ys); ic_type: System.icType (Goldbeter, varyBeta = 1). */
typedef struct crd_ic_params {
  double wave_length, wave_width; /* Parameters.waveLength / waveWidth (fractions) */
  int32_t wave_inside;            /* Parameters.waveInside */
  int32_t ic_type;                /* System.icType */
  double s0, s1;
} crd_ic_params;
int crd_fill_initial_conditions(crd_grid *g, const crd_ic_params *ic, double *y_dev);

/* ---- output snapshots ----------------------------------------------------------------------------- */
/* What main() does after every ARKode() call — walk the state's host array and fprintf variable 0 (and 1 when
 * includeAllVars), src/FHNmodel_torus.cpp:393-410,438-455 — without stalling the time loop: crd_snapshot_begin ENQUEUES an
 * output of the interleaved state as it is at this point of the context's stream (a small gather kernel into contiguous
 * per-variable arrays, then an asynchronous device-to-host copy into one of `nslots` page-locked buffers on a side stream)
 * and returns at once; a consumer thread calls crd_snapshot_wait (blocks until that copy has landed), reads the values and
 * gives the buffer back with crd_snapshot_release.  nvars = 1 captures variable 0 only (8 instead of 16 B/point over PCIe).
 * crd_snapshot_begin returns the slot (>= 0), -1 on failure, -2 when every slot is still held (release one and call again).
 * begin / destroy: the context's thread; wait: any one consumer thread; release: either. */
typedef struct crd_snapshot crd_snapshot;
crd_snapshot *crd_snapshot_create(crd_ctx *ctx, int64_t npoints, int nvars, int nslots);
void crd_snapshot_destroy(crd_snapshot *s);
int crd_snapshot_begin(crd_snapshot *s, const double *y_dev);
int crd_snapshot_wait(crd_snapshot *s, int slot, const double **var0, const double **var1);
int crd_snapshot_release(crd_snapshot *s, int slot);

/* ---- device-resident N_Vector ------------------------------------------------------------------ */
N_Vector N_VNew_Crd(crd_ctx *ctx, long int local_length, long int global_length);
N_Vector N_VNewEmpty_Crd(crd_ctx *ctx, long int local_length, long int global_length);
N_Vector N_VMake_Crd(crd_ctx *ctx, long int local_length, long int global_length, realtype *dev_data);
void N_VDestroy_Crd(N_Vector v);
realtype *N_VGetDeviceArrayPointer_Crd(N_Vector v);
/* host mirror: N_VGetArrayPointer on a device vector returns a pinned host mirror (allocated on first
 * use); it is only current after N_VCopyToHost_Crd and only takes effect after N_VCopyFromHost_Crd.
 * The reference reads its ydata pointer after every ARKode call (:383,438-455) — the driver calls
 * N_VCopyToHost_Crd at those points. */
int N_VCopyToHost_Crd(N_Vector v);
int N_VCopyFromHost_Crd(N_Vector v);
long int N_VGetLocalLength_Crd(N_Vector v);
crd_ctx *N_VGetContext_Crd(N_Vector v);

N_Vector N_VClone_Crd(N_Vector w);
N_Vector N_VCloneEmpty_Crd(N_Vector w);
void N_VSpace_Crd(N_Vector v, long int *lrw, long int *liw);
realtype *N_VGetArrayPointer_Crd(N_Vector v);
void N_VSetArrayPointer_Crd(realtype *dev_data, N_Vector v);
void N_VLinearSum_Crd(realtype a, N_Vector x, realtype b, N_Vector y, N_Vector z);
void N_VConst_Crd(realtype c, N_Vector z);
void N_VProd_Crd(N_Vector x, N_Vector y, N_Vector z);
void N_VDiv_Crd(N_Vector x, N_Vector y, N_Vector z);
void N_VScale_Crd(realtype c, N_Vector x, N_Vector z);
void N_VAbs_Crd(N_Vector x, N_Vector z);
void N_VInv_Crd(N_Vector x, N_Vector z);
void N_VAddConst_Crd(N_Vector x, realtype b, N_Vector z);
realtype N_VDotProd_Crd(N_Vector x, N_Vector y);
realtype N_VMaxNorm_Crd(N_Vector x);
realtype N_VWrmsNorm_Crd(N_Vector x, N_Vector w);
realtype N_VWrmsNormMask_Crd(N_Vector x, N_Vector w, N_Vector id);
realtype N_VMin_Crd(N_Vector x);
realtype N_VWL2Norm_Crd(N_Vector x, N_Vector w);
realtype N_VL1Norm_Crd(N_Vector x);
void N_VCompare_Crd(realtype c, N_Vector x, N_Vector z);
booleantype N_VInvTest_Crd(N_Vector x, N_Vector z);
booleantype N_VConstrMask_Crd(N_Vector c, N_Vector x, N_Vector m);
realtype N_VMinQuotient_Crd(N_Vector num, N_Vector denom);

/* fused operations (see crd_ark.h) */
int N_VLinearCombination_Crd(int n, const realtype *c, N_Vector *X, N_Vector z);
int N_VErkFinish_Crd(int s, const realtype *hb, const realtype *hd, N_Vector yn, N_Vector *F, N_Vector ynew,
                     realtype rtol, realtype atol, realtype out[2]);
/* the same step finish with the bits of the op-by-op sequence: separately rounded chains, IEEE error weights, the error sum
 * accumulated in double-double (order-independent) */
int N_VErkFinishSeq_Crd(int s, const realtype *hb, const realtype *hd, N_Vector yn, N_Vector *F, N_Vector ynew,
                        realtype rtol, realtype atol, realtype out[2]);
/* Tables of fused operations for crd_ARKodeSetFusedOps (crd_ark.h).
 *   crd_nv_fused_ops / crd_nv_fused_vector_ops: chains of fused multiply-adds, approximate reciprocals for the error weights
 *     (the arithmetic that goes with CRD_ARITH_FAST grids).
 *   crd_nv_fused_ops_exact / crd_nv_fused_vector_ops_exact: every fused entry reproduces, bit for bit, what the op-by-op
 *     sequence of N_Vector operations computes (stage assembly, solution, error estimate, error weights), and the error norm is
 *     summed in double-double so that it does not depend on the order of summation: with a CRD_ARITH_EXACT grid the fused loop,
 *     the op-by-op loop, any phi split and the CPU checker take the same steps and produce the same trajectory bit for bit.
 *   crd_nv_fused_ops_for(grid): the table that matches the grid's arithmetic.
 * "vector_ops": stage states are materialised (no rhs_lincomb / rhs_lincomb_finish / erk_evolve). */
const crd_fused_ops *crd_nv_fused_ops(void);
const crd_fused_ops *crd_nv_fused_vector_ops(void);
const crd_fused_ops *crd_nv_fused_ops_exact(void);
const crd_fused_ops *crd_nv_fused_vector_ops_exact(void);
const crd_fused_ops *crd_nv_fused_ops_for(const crd_grid *g);

#ifdef __cplusplus
}
#endif
#endif /* CRD_B200_H */

/* crd_ark.h — extensions of the ARKode-legacy interface (crd_sundials_compat.h) that the device path
 * needs and SUNDIALS 2.x cannot express.
 *
 * The integrator in crdmodel_b200/host/crd_ark.cpp reaches its vectors ONLY through the N_Vector ops
 * table and the ARKRhsFn callback, like ARKode does (reference call sites
 * src/FHNmodel_torus.cpp:356-373,423).  With nothing else set it issues the same op-by-op sequence
 * SUNDIALS 2.x would (N_VLinearSum chains, N_VWrmsNorm, the abs/scale/addconst/inv ewt chain).  A
 * vector implementation may additionally register FUSED operations, which replace those chains by
 * single passes over memory (1360 -> 448 B/point/step):
 *   lincomb     z = sum_j c[j] X[j]                                   (stage assembly, dense output)
 *   rhs_lincomb ydot = f(t, sum_j c[j] X[j])  (stage assembly fused into the right-hand side)
 *   erk_finish  ynew = yn + sum_j hb[j] F[j];  err = sum_j hd[j] F[j];
 *               out[0] = sum_global (err_i  * w_i )^2,  w_i  = 1/(rtol |yn_i|   + atol)
 *               out[1] = sum_global (ynew_i * w'_i)^2,  w'_i = 1/(rtol |ynew_i| + atol): the norm of the new state, which only
 *                        ever feeds the "too much accuracy" test uround * sqrt(out[1] / N) > 1.  Every term is below
 *                        1 / rtol^2, so for rtol > uround the test cannot fire; an implementation may then report the
 *                        bound N / rtol^2 instead of forming the sum (the device kernels do).
 *               (the error weights are a pure function of the state, so no ewt vector is stored).
 *   rhs_lincomb_finish  the LAST stage evaluation and erk_finish in one pass (the stage derivative is never stored)
 *   erk_evolve  the whole adaptive step loop (stages, finish, error test, step controller) inside ONE persistent
 *               cooperative kernel: no launch and no host round trip per step.  For the meshes that live in L2
 *               (the reference's default 400 x 1600 / 100 x 400 grids) a step is bound by launch latency, not by
 *               memory: this is what removes it.
 */
#ifndef CRD_ARK_H
#define CRD_ARK_H
#include "crd_sundials_compat.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CRD_ARK_MAX_LINCOMB 8

struct crd_erk_state;
typedef struct crd_fused_ops {
  int (*lincomb)(int n, const realtype *c, N_Vector *X, N_Vector z);
  int (*erk_finish)(int s, const realtype *hb, const realtype *hd, N_Vector yn, N_Vector *F,
                    N_Vector ynew, realtype rtol, realtype atol, realtype out[2]);
  /* optional: ydot = f(t, sum_j c[j] X[j]) in one pass, the stage state never touching memory (n <= 5);
   * same return convention as ARKRhsFn.  Removes the stage-assembly kernels: 480 instead of 592 B/point/step. */
  int (*rhs_lincomb)(realtype t, int n, const realtype *c, N_Vector *X, N_Vector ydot, void *user_data);
  /* optional: run the adaptive step loop itself on the device (see crd_erk_state).  Returns 0 when it ran
   * (st->flag then holds what ARKode() would return before its dense-output evaluation), > 0 when it does not
   * apply to this problem (the integrator continues with its own host-driven loop), < 0 on a launch failure. */
  int (*erk_evolve)(struct crd_erk_state *st, void *user_data);
  /* optional: the LAST stage and the step finish in one pass.  X = (yn, F_0 .. F_{s-2}); F_{s-1} = f(t, sum_j c[j] X[j]) is
   * not stored; ynew and out[] as erk_finish defines them.  Returns 0, > 0 when it does not apply (the integrator then
   * calls rhs_lincomb and erk_finish), < 0 on failure.  416 instead of 528 B/point/step. */
  int (*rhs_lincomb_finish)(realtype t, int s, const realtype *c, const realtype *hb, const realtype *hd, N_Vector *X,
                            N_Vector ynew, realtype rtol, realtype atol, realtype out[2], void *user_data);
  /* optional: f1 = f(t1, y) and f2 = f(t2, y + c f1) in one pass over y: the derivative at an accepted state (dense output, stage
   * 1 of the next step) together with that step's second stage.  Returns 0, > 0 when it does not apply (the integrator then
   * evaluates them one by one), < 0 on failure.  48 instead of 80 B/point. */
  int (*rhs_pair)(realtype t1, realtype t2, realtype c, N_Vector y, N_Vector f1, N_Vector f2, void *user_data);
} crd_fused_ops;

/* The integrator's state handed to erk_evolve and taken back from it.  The callee advances (tn, yn, fnew =
 * f(tn, yn)) by accepted steps of the explicit method (A, b, d = b - b_embedded, c) under the same error test and
 * PID step controller as the host loop until tn has passed tout (ARK_NORMAL), one step was taken (ARK_ONE_STEP),
 * max_steps steps were taken, or a step failed.  It owns the role of every vector while it runs and returns them
 * permuted (the N_Vector handles are swapped, never the data copied): on return yn / fnew are the accepted state
 * and its derivative, yold / fold those of the step before (dense output), ycur and F[1..s-1] scratch. */
#define CRD_ERK_MAX_STAGES 8
typedef struct crd_erk_state {
  /* method */
  int s, p;                         /* stages, embedding order */
  realtype A[CRD_ERK_MAX_STAGES][CRD_ERK_MAX_STAGES], b[CRD_ERK_MAX_STAGES], d[CRD_ERK_MAX_STAGES], c[CRD_ERK_MAX_STAGES];
  /* tolerances and controller constants (ARKode 1.x defaults: crd_ark.cpp) */
  realtype rtol, atol;
  realtype k1, k2, k3, bias, safety, growth, etamxf, etamin, lbound, ubound;
  int small_nef, maxnef;
  long int nglobal;                 /* global vector length (WRMS denominator) */
  /* request */
  realtype tout;
  int itask;                        /* ARK_NORMAL | ARK_ONE_STEP */
  long int max_steps;               /* steps this call may take (<= 0: no limit) */
  /* evolving state, in/out */
  realtype tn, next_h, hold, eta, etamax, ehist[2];
  realtype ynorm_sq;                /* sum_global (yn_i w_i)^2 of the current state, < 0 when unknown */
  long int nst, nst_attempts, nfe, netf;
  N_Vector yn, yold, ycur, fnew, fold, F[CRD_ERK_MAX_STAGES];
  /* out */
  int flag;                         /* ARK_SUCCESS, ARK_TOO_MUCH_WORK, ARK_TOO_MUCH_ACC, ARK_ERR_FAILURE */
  realtype h_failed;                /* step size of the failed attempt when flag < 0 */
} crd_erk_state;

/* Register fused operations (NULL = op-by-op, the SUNDIALS 2.x sequence). */
int crd_ARKodeSetFusedOps(void *arkode_mem, const crd_fused_ops *ops);
/* Stage 0 of every step is f(tn, yn), which the dense-output bookkeeping has already evaluated at
 * the end of the previous step.  on=1 reuses it (5 instead of 6 RHS evaluations per step, identical
 * results); on=0 (default) re-evaluates like ARKode 1.x. */
int crd_ARKodeSetReuseFirstStage(void *arkode_mem, int on);
/* Device-resident step loop (crd_fused_ops.erk_evolve): on = 1 (default) uses it whenever the registered fused
 * operations offer it and it applies; on = 0 keeps the host-driven loop (one kernel launch per stage and a host
 * round trip per step). */
int crd_ARKodeSetResident(void *arkode_mem, int on);
/* Last stage + step finish in one pass (crd_fused_ops.rhs_lincomb_finish): on = 1 (default) uses it when offered and
 * applicable; on = 0 always issues the stage evaluation and the finish separately. */
int crd_ARKodeSetStageFinish(void *arkode_mem, int on);
/* f(tn, ynew) of an accepted step together with the second stage of the next step in one pass (crd_fused_ops.rhs_pair): on = 1
 * (default) uses it when offered and applicable (needs the reuse of the first stage, an explicit second stage that depends on
 * the first only, adaptive steps); on = 0 evaluates them one by one.  If the next step is not taken with the predicted size
 * (a failed error test, a caller that changes it) the second stage is simply evaluated again. */
int crd_ARKodeSetStagePair(void *arkode_mem, int on);
/* Initial step size (0 = estimate it, the default). */
int crd_ARKodeSetInitStep(void *arkode_mem, realtype hin);
/* Fixed step size (no error test, no adaptivity); 0 switches adaptivity back on. */
int crd_ARKodeSetFixedStep(void *arkode_mem, realtype hfixed);
/* The explicit method in use (what ARKode 1.x reports through ARKodeGetCurrentButcherTables): s stages, order q, embedding
 * order p; A is filled row-major with CRD_ARK_TABLE_DIM columns ([CRD_ARK_TABLE_DIM * CRD_ARK_TABLE_DIM] doubles), c, b and the
 * embedding weights b2 with CRD_ARK_TABLE_DIM entries each. */
#define CRD_ARK_TABLE_DIM 8
int crd_ARKodeGetButcherTable(void *arkode_mem, int *s, int *q, int *p, realtype *A, realtype *c, realtype *b, realtype *b2);

#ifdef __cplusplus
}
#endif
#endif

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
 *               out[1] = sum_global (ynew_i * w'_i)^2,  w'_i = 1/(rtol |ynew_i| + atol)
 *               (the error weights are a pure function of the state, so no ewt vector is stored).
 */
#ifndef CRD_ARK_H
#define CRD_ARK_H
#include "crd_sundials_compat.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CRD_ARK_MAX_LINCOMB 8

typedef struct crd_fused_ops {
  int (*lincomb)(int n, const realtype *c, N_Vector *X, N_Vector z);
  int (*erk_finish)(int s, const realtype *hb, const realtype *hd, N_Vector yn, N_Vector *F,
                    N_Vector ynew, realtype rtol, realtype atol, realtype out[2]);
  /* optional: ydot = f(t, sum_j c[j] X[j]) in one pass, the stage state never touching memory (n <= 5);
   * same return convention as ARKRhsFn.  Removes the stage-assembly kernels: 480 instead of 592 B/point/step. */
  int (*rhs_lincomb)(realtype t, int n, const realtype *c, N_Vector *X, N_Vector ydot, void *user_data);
} crd_fused_ops;

/* Register fused operations (NULL = op-by-op, the SUNDIALS 2.x sequence). */
int crd_ARKodeSetFusedOps(void *arkode_mem, const crd_fused_ops *ops);
/* Stage 0 of every step is f(tn, yn), which the dense-output bookkeeping has already evaluated at
 * the end of the previous step.  on=1 reuses it (5 instead of 6 RHS evaluations per step, identical
 * results); on=0 (default) re-evaluates like ARKode 1.x. */
int crd_ARKodeSetReuseFirstStage(void *arkode_mem, int on);
/* Initial step size (0 = estimate it, the default). */
int crd_ARKodeSetInitStep(void *arkode_mem, realtype hin);
/* Fixed step size (no error test, no adaptivity); 0 switches adaptivity back on. */
int crd_ARKodeSetFixedStep(void *arkode_mem, realtype hfixed);

#ifdef __cplusplus
}
#endif
#endif

/* crd_sundials_compat.h — the SUNDIALS 2.x plug points the CRDModel hot path sits behind.
 *
 * The reference programs bind to SUNDIALS through exactly two interfaces
 * (reference: src/FHNmodel_torus.cpp:49-53,126,281,356-373,423,488,491,506,517):
 *   (1) the generic N_Vector  { void *content; ops *ops; }  with the 2.x operations table, and
 *   (2) the legacy ARKode calls  ARKodeCreate / ARKodeInit(f, NULL, t0, y) / ARKodeSStolerances /
 *       ARKodeSetUserData / ARKodeSetMaxNumSteps / ARKode(..., ARK_NORMAL) / ARKodeFree
 *       with the callback type  ARKRhsFn.
 * SUNDIALS itself is not vendored by the reference (CMake/FindSUNDIALS.cmake:4-6 only searches for an
 * installed copy; the API used pins it to 2.6.0-2.7.0).  This header restates those two interfaces
 * so that (a) the device N_Vector of crd_b200.h can be handed to code written against SUNDIALS 2.x
 * and (b) the explicit adaptive Runge-Kutta driver in crdmodel_b200/host/crd_ark.cpp can be used
 * where ARKode was.  Struct layout and ops order follow nvector/sundials_nvector.h of SUNDIALS 2.6;
 * define CRD_SUNDIALS_27 to get the 2.7 layout (one extra leading slot, nvgetvectorid).
 *
 * Plain C, no CUDA or torch types.
 */
#ifndef CRD_SUNDIALS_COMPAT_H
#define CRD_SUNDIALS_COMPAT_H

#ifdef __cplusplus
extern "C" {
#endif

#ifndef SUNDIALS_DOUBLE_PRECISION
#define SUNDIALS_DOUBLE_PRECISION 1
#endif
typedef double realtype;
#define RCONST(x) x
#ifndef booleantype
#define booleantype int
#endif
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif

/* ---- generic N_Vector (sundials_nvector.h, SUNDIALS 2.6) ------------------------------------ */
typedef struct _generic_N_Vector_Ops *N_Vector_Ops;
typedef struct _generic_N_Vector *N_Vector;

struct _generic_N_Vector_Ops {
#ifdef CRD_SUNDIALS_27
  int         (*nvgetvectorid)(N_Vector);
#endif
  N_Vector    (*nvclone)(N_Vector);
  N_Vector    (*nvcloneempty)(N_Vector);
  void        (*nvdestroy)(N_Vector);
  void        (*nvspace)(N_Vector, long int *, long int *);
  realtype*   (*nvgetarraypointer)(N_Vector);
  void        (*nvsetarraypointer)(realtype *, N_Vector);
  void        (*nvlinearsum)(realtype, N_Vector, realtype, N_Vector, N_Vector);
  void        (*nvconst)(realtype, N_Vector);
  void        (*nvprod)(N_Vector, N_Vector, N_Vector);
  void        (*nvdiv)(N_Vector, N_Vector, N_Vector);
  void        (*nvscale)(realtype, N_Vector, N_Vector);
  void        (*nvabs)(N_Vector, N_Vector);
  void        (*nvinv)(N_Vector, N_Vector);
  void        (*nvaddconst)(N_Vector, realtype, N_Vector);
  realtype    (*nvdotprod)(N_Vector, N_Vector);
  realtype    (*nvmaxnorm)(N_Vector);
  realtype    (*nvwrmsnorm)(N_Vector, N_Vector);
  realtype    (*nvwrmsnormmask)(N_Vector, N_Vector, N_Vector);
  realtype    (*nvmin)(N_Vector);
  realtype    (*nvwl2norm)(N_Vector, N_Vector);
  realtype    (*nvl1norm)(N_Vector);
  void        (*nvcompare)(realtype, N_Vector, N_Vector);
  booleantype (*nvinvtest)(N_Vector, N_Vector);
  booleantype (*nvconstrmask)(N_Vector, N_Vector, N_Vector);
  realtype    (*nvminquotient)(N_Vector, N_Vector);
};

struct _generic_N_Vector {
  void *content;
  struct _generic_N_Vector_Ops *ops;
};

/* Generic dispatchers (sundials_nvector.c): each forwards to v->ops->...  Implemented in
 * crdmodel_b200/host/crd_nvector_generic.c; they work for any N_Vector with a 2.x ops table. */
N_Vector N_VClone(N_Vector w);
void N_VDestroy(N_Vector v);
realtype *N_VGetArrayPointer(N_Vector v);
void N_VLinearSum(realtype a, N_Vector x, realtype b, N_Vector y, N_Vector z);
void N_VConst(realtype c, N_Vector z);
void N_VProd(N_Vector x, N_Vector y, N_Vector z);
void N_VDiv(N_Vector x, N_Vector y, N_Vector z);
void N_VScale(realtype c, N_Vector x, N_Vector z);
void N_VAbs(N_Vector x, N_Vector z);
void N_VInv(N_Vector x, N_Vector z);
void N_VAddConst(N_Vector x, realtype b, N_Vector z);
realtype N_VDotProd(N_Vector x, N_Vector y);
realtype N_VMaxNorm(N_Vector x);
realtype N_VWrmsNorm(N_Vector x, N_Vector w);
realtype N_VWrmsNormMask(N_Vector x, N_Vector w, N_Vector id);
realtype N_VMin(N_Vector x);
realtype N_VWL2Norm(N_Vector x, N_Vector w);
realtype N_VL1Norm(N_Vector x);
void N_VCompare(realtype c, N_Vector x, N_Vector z);
booleantype N_VInvTest(N_Vector x, N_Vector z);
booleantype N_VConstrMask(N_Vector c, N_Vector x, N_Vector m);
realtype N_VMinQuotient(N_Vector num, N_Vector denom);

/* ---- ARKode legacy interface (arkode/arkode.h, SUNDIALS 2.6/2.7) ----------------------------- */
typedef int (*ARKRhsFn)(realtype t, N_Vector y, N_Vector ydot, void *user_data);

#define ARK_NORMAL 1
#define ARK_ONE_STEP 2

#define ARK_SUCCESS 0
#define ARK_TSTOP_RETURN 1
#define ARK_ROOT_RETURN 2
#define ARK_WARNING 99
#define ARK_TOO_MUCH_WORK -1
#define ARK_TOO_MUCH_ACC -2
#define ARK_ERR_FAILURE -3
#define ARK_CONV_FAILURE -4
#define ARK_RHSFUNC_FAIL -8
#define ARK_FIRST_RHSFUNC_ERR -9
#define ARK_REPTD_RHSFUNC_ERR -10
#define ARK_UNREC_RHSFUNC_ERR -11
#define ARK_MEM_FAIL -20
#define ARK_MEM_NULL -21
#define ARK_ILL_INPUT -22
#define ARK_NO_MALLOC -23
#define ARK_BAD_K -24
#define ARK_BAD_T -25
#define ARK_BAD_DKY -26
#define ARK_TOO_CLOSE -27

void *ARKodeCreate(void);
int ARKodeInit(void *arkode_mem, ARKRhsFn fe, ARKRhsFn fi, realtype t0, N_Vector y0);
int ARKodeSStolerances(void *arkode_mem, realtype reltol, realtype abstol);
int ARKodeSetUserData(void *arkode_mem, void *user_data);
int ARKodeSetMaxNumSteps(void *arkode_mem, long int mxsteps);
int ARKode(void *arkode_mem, realtype tout, N_Vector yout, realtype *tret, int itask);
void ARKodeFree(void **arkode_mem);
/* statistics the reference never queries but the north-star asks to be reported */
int ARKodeGetNumSteps(void *arkode_mem, long int *nsteps);
int ARKodeGetNumRhsEvals(void *arkode_mem, long int *nfe_evals, long int *nfi_evals);
int ARKodeGetNumErrTestFails(void *arkode_mem, long int *netfails);
int ARKodeGetNumStepAttempts(void *arkode_mem, long int *nsteps);
int ARKodeGetCurrentStep(void *arkode_mem, realtype *hcur);
int ARKodeGetLastStep(void *arkode_mem, realtype *hlast);
int ARKodeGetCurrentTime(void *arkode_mem, realtype *tcur);

#ifdef __cplusplus
}
#endif
#endif /* CRD_SUNDIALS_COMPAT_H */

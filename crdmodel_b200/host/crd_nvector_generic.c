/* Generic N_Vector dispatchers: N_VXxx(v, ...) forwards to v->ops->nvxxx, exactly what
 * sundials_nvector.c of SUNDIALS 2.x does.  The reference reaches its vectors through these names
 * (src/FHNmodel_torus.cpp:303,383,506,786), so code written against SUNDIALS keeps compiling when
 * the vector behind the handle is the device N_Vector of crd_b200.h.  Host-only, no CUDA. */
#include <stddef.h>

#include "crd_sundials_compat.h"

N_Vector N_VClone(N_Vector w) { return w->ops->nvclone(w); }
void N_VDestroy(N_Vector v) { if (v != NULL) v->ops->nvdestroy(v); }
realtype *N_VGetArrayPointer(N_Vector v) { return v->ops->nvgetarraypointer(v); }
void N_VLinearSum(realtype a, N_Vector x, realtype b, N_Vector y, N_Vector z) { z->ops->nvlinearsum(a, x, b, y, z); }
void N_VConst(realtype c, N_Vector z) { z->ops->nvconst(c, z); }
void N_VProd(N_Vector x, N_Vector y, N_Vector z) { z->ops->nvprod(x, y, z); }
void N_VDiv(N_Vector x, N_Vector y, N_Vector z) { z->ops->nvdiv(x, y, z); }
void N_VScale(realtype c, N_Vector x, N_Vector z) { z->ops->nvscale(c, x, z); }
void N_VAbs(N_Vector x, N_Vector z) { z->ops->nvabs(x, z); }
void N_VInv(N_Vector x, N_Vector z) { z->ops->nvinv(x, z); }
void N_VAddConst(N_Vector x, realtype b, N_Vector z) { z->ops->nvaddconst(x, b, z); }
realtype N_VDotProd(N_Vector x, N_Vector y) { return y->ops->nvdotprod(x, y); }
realtype N_VMaxNorm(N_Vector x) { return x->ops->nvmaxnorm(x); }
realtype N_VWrmsNorm(N_Vector x, N_Vector w) { return x->ops->nvwrmsnorm(x, w); }
realtype N_VWrmsNormMask(N_Vector x, N_Vector w, N_Vector id) { return x->ops->nvwrmsnormmask(x, w, id); }
realtype N_VMin(N_Vector x) { return x->ops->nvmin(x); }
realtype N_VWL2Norm(N_Vector x, N_Vector w) { return x->ops->nvwl2norm(x, w); }
realtype N_VL1Norm(N_Vector x) { return x->ops->nvl1norm(x); }
void N_VCompare(realtype c, N_Vector x, N_Vector z) { z->ops->nvcompare(c, x, z); }
booleantype N_VInvTest(N_Vector x, N_Vector z) { return z->ops->nvinvtest(x, z); }
booleantype N_VConstrMask(N_Vector c, N_Vector x, N_Vector m) { return x->ops->nvconstrmask(c, x, m); }
realtype N_VMinQuotient(N_Vector num, N_Vector denom) { return num->ops->nvminquotient(num, denom); }

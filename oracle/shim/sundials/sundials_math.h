/* TEST INFRASTRUCTURE — stands in for <sundials/sundials_math.h>; the reference includes it
 * (FHNmodel_torus.cpp:53) but uses none of its macros. */
#ifndef CRD_ORACLE_SHIM_SUNDIALS_MATH_H
#define CRD_ORACLE_SHIM_SUNDIALS_MATH_H
#include <math.h>
#include "crd_sundials_compat.h"
#define SUNRabs(x) fabs(x)
#define SUNRsqrt(x) ((x) <= 0.0 ? 0.0 : sqrt(x))
#define SUNMIN(a, b) ((a) < (b) ? (a) : (b))
#define SUNMAX(a, b) ((a) > (b) ? (a) : (b))
#define SUNSQR(a) ((a) * (a))
#endif

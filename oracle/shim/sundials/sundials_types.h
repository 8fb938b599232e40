/* TEST INFRASTRUCTURE — stands in for <sundials/sundials_types.h> (SUNDIALS 2.6/2.7, not vendored by
 * the reference) when compiling /root/reference/src/*.cpp in place. */
#ifndef CRD_ORACLE_SHIM_SUNDIALS_TYPES_H
#define CRD_ORACLE_SHIM_SUNDIALS_TYPES_H
#include "crd_sundials_compat.h"
#endif

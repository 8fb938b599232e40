/* TEST INFRASTRUCTURE — stands in for <arkode/arkode.h> (legacy ARKode 1.x API).  The declarations
 * live in include/crd_sundials_compat.h; the implementation linked behind them is the product's own
 * explicit adaptive Runge-Kutta driver (crdmodel_b200/host/crd_ark.cpp). */
#ifndef CRD_ORACLE_SHIM_ARKODE_H
#define CRD_ORACLE_SHIM_ARKODE_H
#include "crd_sundials_compat.h"
#endif

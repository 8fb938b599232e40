/* TEST INFRASTRUCTURE — the reference includes <arkode/arkode_pcg.h> (FHNmodel_torus.cpp:51) but
 * never calls the PCG solver (ARKodeInit is given fi = NULL). */
#ifndef CRD_ORACLE_SHIM_ARKODE_PCG_H
#define CRD_ORACLE_SHIM_ARKODE_PCG_H
#endif

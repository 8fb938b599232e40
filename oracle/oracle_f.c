/* TEST INFRASTRUCTURE — ARKRhsFn over the plain-C restatement, for host N_Vectors (nvector_host.c),
 * so the explicit RK driver can be run entirely on the CPU as the trajectory checker.
 * user_data is a crd_oracle_params*; single subdomain (np = 1). */
#include "crd_oracle.h"
#include "nvector/nvector_parallel.h"

int crd_oracle_f(realtype t, N_Vector y, N_Vector ydot, void *user_data) {
  const crd_oracle_params *P = (const crd_oracle_params *)user_data;
  return crd_oracle_rhs(P, t, NV_DATA_P(y), NV_DATA_P(ydot)) == 0 ? 0 : -1;
}

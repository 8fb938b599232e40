/* TEST INFRASTRUCTURE — stands in for <nvector/nvector_parallel.h> (SUNDIALS 2.6/2.7).  Content
 * layout and macro names follow the published header; the implementation (oracle/nvector_host.c)
 * is a host-memory restatement of nvector_parallel.c used as the CPU checker for the device
 * N_Vector.  Not a product component. */
#ifndef CRD_ORACLE_SHIM_NVECTOR_PARALLEL_H
#define CRD_ORACLE_SHIM_NVECTOR_PARALLEL_H
#include "crd_sundials_compat.h"
#include "mpi.h"
#ifdef __cplusplus
extern "C" {
#endif

struct _N_VectorContent_Parallel {
  long int local_length;
  long int global_length;
  booleantype own_data;
  realtype *data;
  MPI_Comm comm;
};
typedef struct _N_VectorContent_Parallel *N_VectorContent_Parallel;

#define NV_CONTENT_P(v) ((N_VectorContent_Parallel)(v->content))
#define NV_LOCLENGTH_P(v) (NV_CONTENT_P(v)->local_length)
#define NV_GLOBLENGTH_P(v) (NV_CONTENT_P(v)->global_length)
#define NV_OWN_DATA_P(v) (NV_CONTENT_P(v)->own_data)
#define NV_DATA_P(v) (NV_CONTENT_P(v)->data)
#define NV_COMM_P(v) (NV_CONTENT_P(v)->comm)
#define NV_Ith_P(v, i) (NV_DATA_P(v)[i])

N_Vector N_VNew_Parallel(MPI_Comm comm, long int local_length, long int global_length);
N_Vector N_VNewEmpty_Parallel(MPI_Comm comm, long int local_length, long int global_length);
N_Vector N_VMake_Parallel(MPI_Comm comm, long int local_length, long int global_length, realtype *v_data);
void N_VDestroy_Parallel(N_Vector v);

#ifdef __cplusplus
}
#endif
#endif

/* TEST INFRASTRUCTURE — in-process MPI subset so the reference's own Exchange()/SetupDecomp()
 * (reference: src/FHNmodel_torus.cpp:708-950) compile and run unmodified without an MPI install.
 * Ranks are emulated by threads of one process: every thread announces its rank with
 * crdshim_mpi_bind(rank) before calling into reference code.  Point-to-point messages are buffered
 * copies matched per (source, destination) pair in posting order, which is what MPI's non-overtaking
 * rule gives for the reference's MPI_ANY_TAG receives.  Not a product component.
 */
#ifndef CRD_ORACLE_SHIM_MPI_H
#define CRD_ORACLE_SHIM_MPI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef struct { int kind; int peer; long seq; void *buf; int count; int comm; } MPI_Request;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_ANY_TAG (-1)
#define MPI_DOUBLE 1
#define MPI_FLOAT 2
#define MPI_LONG_DOUBLE 3
#define MPI_LONG 4
#define MPI_INT 5
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Dims_create(int nnodes, int ndims, int dims[]);
int MPI_Cart_create(MPI_Comm old, int ndims, const int dims[], const int periods[], int reorder, MPI_Comm *cart);
int MPI_Cart_get(MPI_Comm comm, int maxdims, int dims[], int periods[], int coords[]);
int MPI_Cart_shift(MPI_Comm comm, int direction, int disp, int *rank_source, int *rank_dest);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Wait(MPI_Request *req, MPI_Status *status);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);

/* shim control (not MPI): world size for subsequent runs; per-thread rank binding */
void crdshim_mpi_set_world(int nranks);
void crdshim_mpi_bind(int rank);
/* band mode (world size 1): Exchange()'s S / N receives deliver these rows instead of the rank's own (NULL = off) */
void crdshim_mpi_band_halo(const double *south_row, const double *north_row);

#ifdef __cplusplus
}
#endif
#endif

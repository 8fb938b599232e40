// crd_common.cuh — internal definitions shared by the sm_100a translation units of libcrd_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "crd_b200.h"

namespace crd {

void set_error(const char *fmt, ...);

#define CRD_CUDA(call)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      crd::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));              \
      return -1;                                                                                         \
    }                                                                                                    \
  } while (0)

#define CRD_CUDA_NULL(call)                                                                              \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      crd::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));              \
      return nullptr;                                                                                    \
    }                                                                                                    \
  } while (0)

constexpr int kSMs = 148;            // B200 (crd_ctx_create checks the device against it: crd_ctx::sms)
constexpr int kRedBlocks = kSMs * 4; // partial sums per reduction (fixed => deterministic order)
constexpr int kRedThreads = 256;
constexpr int kRedSlots = 4;         // values one reduction kernel can return
constexpr int kMaxRanks = 64;        // ranks of a phi split (cross-rank sums gather one slot per rank)

// Device-side allreduce of a few doubles across the ranks of a phi split (the MPI_Allreduce inside nvector_parallel's
// reductions).  Every context owns a mailbox block in device memory (exported by CUDA IPC); a reduction's finishing block
// stores its local values into EVERY rank's mailbox through the peer mappings (its own included), publishes a sequence
// number per destination, waits until every rank's sequence number has arrived in its own block, and combines the values
// in rank order — all ranks therefore compute the same bits, and no host takes part.  Mailboxes are double-buffered by the
// parity of the sequence number (a rank can be at most one reduction ahead: it needs everybody's values to finish one).
constexpr int kCommVals = 4;
struct CommTab {
  double *mail[kMaxRanks];               // rank r's block: double [2 parity][kMaxRanks source][kCommVals]
  unsigned long long *flag[kMaxRanks];   // rank r's flags: [kMaxRanks source] sequence number of the last complete store
  int rank, nranks;
  long long timeout_ns;
  int *err;
};
constexpr size_t kCommMailBytes = sizeof(double) * 2 * kMaxRanks * kCommVals;
constexpr size_t kCommBlockBytes = kCommMailBytes + sizeof(unsigned long long) * kMaxRanks;
enum { COMM_SUM_DD = 0, COMM_MAX = 1, COMM_MIN = 2 };   // SUM_DD: slots (hi, plain, lo): hi/lo merged in double-double, plain added

}  // namespace crd

struct crd_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int rank = 0, nranks = 1;
  crd_allreduce_fn allreduce = nullptr;
  void *allreduce_user = nullptr;
  int64_t launches = 0;
  int sms = crd::kSMs;                // multiProcessorCount of the device (persistent grids are sized from it)
  // a device-side failure (a neighbour's halo rows never arrived) is sticky: once seen, every entry point fails
  bool failed = false;
  long long halo_timeout_ns = 30LL * 1000000000LL;   // CRD_HALO_TIMEOUT_MS / crd_ctx_set_halo_timeout
  // reduction scratch: per-block partials, ticket, and a mapped pinned result the last block writes
  double *red_partial = nullptr;      // [kRedSlots][kRedBlocks]
  unsigned int *red_ticket = nullptr;
  double *red_result_host = nullptr;  // pinned, mapped
  double *red_result_dev = nullptr;   // device alias of red_result_host
  // error word written by device code (halo wait timeout); pinned, mapped
  int *err_host = nullptr;
  int *err_dev = nullptr;
  // device-side allreduce (crd_ctx_comm_*): the local mailbox block, the table of everybody's, the reduction counter
  char *comm_local = nullptr;
  crd::CommTab *comm_tab = nullptr;       // device copy of the table
  void *comm_peer[crd::kMaxRanks] = {};   // IPC mappings to close
  bool dev_comm = false;
  unsigned long long comm_seq = 0;
  double *red_local = nullptr;            // device: a reduction's local result, input of the exchange
  // L2 flush scratch
  void *flush_buf = nullptr;
  size_t flush_bytes = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace crd {
// The error word is mapped host memory the device writes (halo wait timed out): reading it costs nothing.  Reported once
// with its code, then the context stays failed — an evaluation that ran with stale ghost rows must never be integrated on.
inline int device_failed(crd_ctx *c) {
  if (!c->failed && c->err_host && *c->err_host != 0) {
    c->failed = true;
    set_error("device-side error %d: %s did not arrive within %.1f s (rank %d of %d); the context is unusable", *c->err_host,
              *c->err_host == 110 ? "another rank's contribution to a reduction" : "a neighbour's halo rows",
              (double)c->halo_timeout_ns * 1e-9, c->rank, c->nranks);
    return 1;
  }
  return c->failed ? 1 : 0;
}
inline int use(crd_ctx *c) {
  if (device_failed(c)) return -1;
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) { set_error("cudaSetDevice(%d): %s", c->device, cudaGetErrorString(e)); return -1; }
  return 0;
}
// every wait for the stream on the evaluation path goes through here: a timed-out halo wait surfaces at the next one
inline int sync_stream(crd_ctx *c, const char *what) {
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) { set_error("%s: stream synchronize failed: %s", what, cudaGetErrorString(e)); return -1; }
  return device_failed(c) ? -1 : 0;
}
// Host side of an order-independent global sum: every rank holds the double-double (hi, lo) of its local sum and a plain
// second value; the pairs are gathered through the allreduce hook (a SUM over a vector that is zero outside the rank's own
// slots is exact) and merged in rank order, so all ranks — and any phi split — get the same bits.  Returns RN(hi + lo).
int allreduce_dd(crd_ctx *c, double &hi, double &lo, double *plain);
// Where a reduction kernel's finishing block writes: the mapped host result when the value is final (one rank, or the host
// hook finishes it), device scratch when the device-side exchange follows.
inline double *red_target(crd_ctx *c) { return c->dev_comm ? c->red_local : c->red_result_dev; }
// the exchange as a launch of its own (n <= kCommVals values of red_local -> red_result); no-op unless dev_comm
int launch_comm_exchange(crd_ctx *c, int op, int n);

inline int check_launch(crd_ctx *c, const char *what) {
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("launch of %s failed: %s", what, cudaGetErrorString(e)); return -1; }
  return 0;
}
}  // namespace crd

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
    set_error("device-side error %d: the neighbour's halo rows did not arrive within %.1f s (rank %d of %d); the context is unusable",
              *c->err_host, (double)c->halo_timeout_ns * 1e-9, c->rank, c->nranks);
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

inline int check_launch(crd_ctx *c, const char *what) {
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("launch of %s failed: %s", what, cudaGetErrorString(e)); return -1; }
  return 0;
}
}  // namespace crd

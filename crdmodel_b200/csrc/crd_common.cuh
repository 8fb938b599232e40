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

constexpr int kSMs = 148;            // B200
constexpr int kRedBlocks = kSMs * 4; // partial sums per reduction (fixed => deterministic order)
constexpr int kRedThreads = 256;
constexpr int kRedSlots = 4;         // values one reduction kernel can return

}  // namespace crd

struct crd_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int rank = 0, nranks = 1;
  crd_allreduce_fn allreduce = nullptr;
  void *allreduce_user = nullptr;
  int64_t launches = 0;
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
inline int use(const crd_ctx *c) {
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) { set_error("cudaSetDevice(%d): %s", c->device, cudaGetErrorString(e)); return -1; }
  return 0;
}
inline int check_launch(crd_ctx *c, const char *what) {
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("launch of %s failed: %s", what, cudaGetErrorString(e)); return -1; }
  return 0;
}
}  // namespace crd

// crd_snapshot.cu — output snapshots of the state, off the time loop's critical path.
//
// Reference being replaced: after every ARKode() call main() walks the N_Vector's host array and fprintf's variable 0 (and
// variable 1 when includeAllVars) point by point (src/FHNmodel_torus.cpp:393-410,438-455).  With the state on the device
// that would be a blocking copy of the whole interleaved vector per output.  Here an output is ENQUEUED:
//   integrator's stream : [wait: slot's previous copy done] gather kernel: interleaved (u,v) -> contiguous u [and v]   (~32 B/point)
//   side stream         : [wait: gather done] cudaMemcpyAsync device -> page-locked host buffer of the slot, event
// and the time loop continues at once; the formatter thread (host/crd_writer.hpp) waits for the slot's event.  Only the
// variables that are written travel over PCIe (8 B/point instead of 16 when includeAllVars = 0).
#include <atomic>

#include "crd_common.cuh"

using namespace crd;

namespace {
constexpr int kMaxSlots = 4;

__global__ void __launch_bounds__(256) gather_vars_kernel(const double2 *__restrict__ y, double *__restrict__ v0, double *__restrict__ v1, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n2 = n >> 1;   // two points per thread: 16-byte stores
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += stride) {
    const double2 a = y[2 * i], b = y[2 * i + 1];
    reinterpret_cast<double2 *>(v0)[i] = make_double2(a.x, b.x);
    if (v1) reinterpret_cast<double2 *>(v1)[i] = make_double2(a.y, b.y);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const double2 a = y[n - 1];
    v0[n - 1] = a.x;
    if (v1) v1[n - 1] = a.y;
  }
}
}  // namespace

struct crd_snapshot {
  crd_ctx *ctx = nullptr;
  long long n = 0, stride = 0;    // stride: n rounded up to even (the second variable's array stays 16-byte aligned)
  int nvars = 1, nslots = 2;
  double *dev[kMaxSlots] = {};    // [nvars][n]
  double *host[kMaxSlots] = {};   // page-locked, same layout
  cudaEvent_t gathered[kMaxSlots] = {}, copied[kMaxSlots] = {};
  std::atomic<bool> busy[kMaxSlots];   // handed out by begin (the context's thread), released by the consumer thread
  bool used[kMaxSlots] = {};
  cudaStream_t side = nullptr;
  int next = 0;
};

extern "C" {

crd_snapshot *crd_snapshot_create(crd_ctx *ctx, int64_t npoints, int nvars, int nslots) {
  if (!ctx || npoints < 1 || nvars < 1 || nvars > 2 || nslots < 1 || nslots > kMaxSlots) { set_error("crd_snapshot_create: bad arguments"); return nullptr; }
  if (use(ctx)) return nullptr;
  crd_snapshot *s = new crd_snapshot;
  for (auto &b : s->busy) b.store(false);
  s->ctx = ctx; s->n = npoints; s->nvars = nvars; s->nslots = nslots;
  s->stride = (npoints + 1) & ~1LL;
  const size_t bytes = sizeof(double) * (size_t)s->stride * (size_t)nvars;
  cudaError_t e = cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking);
  for (int k = 0; k < nslots && e == cudaSuccess; ++k) {
    if ((e = cudaMalloc(&s->dev[k], bytes)) != cudaSuccess) break;
    if ((e = cudaHostAlloc(&s->host[k], bytes, cudaHostAllocDefault)) != cudaSuccess) break;
    if ((e = cudaEventCreateWithFlags(&s->gathered[k], cudaEventDisableTiming)) != cudaSuccess) break;
    if ((e = cudaEventCreateWithFlags(&s->copied[k], cudaEventDisableTiming)) != cudaSuccess) break;
  }
  if (e != cudaSuccess) {
    set_error("crd_snapshot_create: %s", cudaGetErrorString(e));
    crd_snapshot_destroy(s);
    return nullptr;
  }
  return s;
}

void crd_snapshot_destroy(crd_snapshot *s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  if (s->side) { cudaStreamSynchronize(s->side); cudaStreamDestroy(s->side); }
  for (int k = 0; k < kMaxSlots; ++k) {
    if (s->dev[k]) cudaFree(s->dev[k]);
    if (s->host[k]) cudaFreeHost(s->host[k]);
    if (s->gathered[k]) cudaEventDestroy(s->gathered[k]);
    if (s->copied[k]) cudaEventDestroy(s->copied[k]);
  }
  delete s;
}

// Enqueue an output of the interleaved state y_dev (npoints (u,v) pairs) as it is at this point of the context's stream.
// Returns the slot (>= 0), -1 on failure, -2 when every slot is still held by the consumer (release one and call again).
int crd_snapshot_begin(crd_snapshot *s, const double *y_dev) {
  if (!s || !y_dev) { set_error("crd_snapshot_begin: null argument"); return -1; }
  crd_ctx *c = s->ctx;
  if (use(c)) return -1;
  int k = -1;
  for (int t = 0; t < s->nslots; ++t) {
    const int cand = (s->next + t) % s->nslots;
    if (!s->busy[cand].load(std::memory_order_acquire)) { k = cand; break; }
  }
  if (k < 0) return -2;
  s->next = (k + 1) % s->nslots;
  // the slot's previous copy has left its device buffer before the gather overwrites it (its consumer has released the host side)
  if (s->used[k]) CRD_CUDA(cudaStreamWaitEvent(c->stream, s->copied[k], 0));
  long long blocks = ((s->n + 1) / 2 + 255) / 256;
  if (blocks > (long long)c->sms * 16) blocks = (long long)c->sms * 16;
  if (blocks < 1) blocks = 1;
  gather_vars_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(reinterpret_cast<const double2 *>(y_dev), s->dev[k],
                                                              s->nvars == 2 ? s->dev[k] + s->stride : nullptr, s->n);
  if (check_launch(c, "gather_vars_kernel")) return -1;
  CRD_CUDA(cudaEventRecord(s->gathered[k], c->stream));
  CRD_CUDA(cudaStreamWaitEvent(s->side, s->gathered[k], 0));
  CRD_CUDA(cudaMemcpyAsync(s->host[k], s->dev[k], sizeof(double) * (size_t)s->stride * (size_t)s->nvars, cudaMemcpyDeviceToHost, s->side));
  CRD_CUDA(cudaEventRecord(s->copied[k], s->side));
  s->busy[k].store(true, std::memory_order_release);
  s->used[k] = true;
  return k;
}

// Block until the slot's values are on the host (any host thread); var0 / var1: contiguous [npoints] each (var1 NULL if not captured)
int crd_snapshot_wait(crd_snapshot *s, int slot, const double **var0, const double **var1) {
  if (!s || slot < 0 || slot >= s->nslots || !s->busy[slot].load(std::memory_order_acquire)) { set_error("crd_snapshot_wait: bad slot"); return -1; }
  CRD_CUDA(cudaSetDevice(s->ctx->device));
  CRD_CUDA(cudaEventSynchronize(s->copied[slot]));
  if (device_failed(s->ctx)) return -1;
  if (var0) *var0 = s->host[slot];
  if (var1) *var1 = s->nvars == 2 ? s->host[slot] + s->stride : nullptr;
  return 0;
}

int crd_snapshot_release(crd_snapshot *s, int slot) {
  if (!s || slot < 0 || slot >= s->nslots) return -1;
  s->busy[slot].store(false, std::memory_order_release);
  return 0;
}

}  // extern "C"

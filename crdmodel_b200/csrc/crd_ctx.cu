// crd_ctx.cu — device context, memory helpers, timers, synthetic-state generator.
#include <cstddef>
#include <cstdlib>

#include "crd_common.cuh"
#include "crd_fused.cuh"

namespace crd {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
int allreduce_dd(crd_ctx *c, double &hi, double &lo, double *plain) {
  if (c->nranks <= 1 || c->dev_comm) return 0;   // one rank, or the device-side exchange has already made the values global
  if (c->nranks > kMaxRanks || !c->allreduce) { set_error("allreduce_dd: %d ranks without a usable allreduce hook", c->nranks); return -1; }
  double v[3 * kMaxRanks];
  const int n = 3 * c->nranks;
  for (int i = 0; i < n; ++i) v[i] = 0.0;
  v[3 * c->rank] = hi; v[3 * c->rank + 1] = lo; v[3 * c->rank + 2] = plain ? *plain : 0.0;
  if (c->allreduce(v, n, CRD_SUM, c->allreduce_user) != 0) { set_error("allreduce hook failed"); return -1; }
  double h = v[0], l = v[1], p = v[2];
  for (int r = 1; r < c->nranks; ++r) { dd_merge(h, l, v[3 * r], v[3 * r + 1]); p += v[3 * r + 2]; }
  hi = h; lo = l;
  if (plain) *plain = p;
  return 0;
}
static __global__ void __launch_bounds__(64) comm_exchange_kernel(const CommTab *T, unsigned long long seq, int op, int n, const double *local, double *result) {
  __shared__ double v[kCommVals];
  if (threadIdx.x < kCommVals) v[threadIdx.x] = threadIdx.x < n ? local[threadIdx.x] : 0.0;
  __syncthreads();
  comm_allreduce_block(T, seq, op, kCommVals, v);   // every slot travels (zeros beyond n): nothing stale is ever combined
  if (threadIdx.x == 0) {
    for (int i = 0; i < kCommVals; ++i) result[i] = v[i];
    __threadfence_system();
  }
}

int launch_comm_exchange(crd_ctx *c, int op, int n) {
  if (!c->dev_comm) return 0;
  ++c->comm_seq;
  comm_exchange_kernel<<<1, 64, 0, c->stream>>>(c->comm_tab, c->comm_seq, op, n, c->red_local, c->red_result_dev);
  return check_launch(c, "comm_exchange_kernel");
}
}  // namespace crd

using namespace crd;

extern "C" {

const char *crd_last_error(void) { return g_err; }

int crd_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

crd_ctx *crd_ctx_create(int device, void *stream) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); libcrd_b200 has no CPU path", e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
    return nullptr;
  }
  if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return nullptr; }
  CRD_CUDA_NULL(cudaSetDevice(device));
  cudaDeviceProp prop;
  CRD_CUDA_NULL(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libcrd_b200 is built for sm_100a only", device, prop.major, prop.minor);
    return nullptr;
  }
  if (prop.multiProcessorCount * 3 > kRedBlocks) {
    set_error("device %d has %d SMs; the per-launch partial-sum tables are sized for %d (B200)", device, prop.multiProcessorCount, crd::kSMs);
    return nullptr;
  }
  crd_ctx *c = new crd_ctx;
  c->device = device;
  c->sms = prop.multiProcessorCount;
  if (const char *e = std::getenv("CRD_HALO_TIMEOUT_MS")) {
    const long long ms = std::atoll(e);
    if (ms > 0) c->halo_timeout_ns = ms * 1000000LL;
  }
  // every allocation below is released by crd_ctx_destroy, which tolerates the ones that never happened
  auto setup = [&]() -> int {
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else { CRD_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    CRD_CUDA(cudaMalloc(&c->red_partial, sizeof(double) * kRedSlots * kRedBlocks));
    CRD_CUDA(cudaMalloc(&c->red_ticket, sizeof(unsigned int)));
    CRD_CUDA(cudaMemset(c->red_ticket, 0, sizeof(unsigned int)));
    CRD_CUDA(cudaHostAlloc(&c->red_result_host, sizeof(double) * kRedSlots, cudaHostAllocMapped));
    CRD_CUDA(cudaHostGetDevicePointer(&c->red_result_dev, c->red_result_host, 0));
    CRD_CUDA(cudaMalloc(&c->red_local, sizeof(double) * kRedSlots));
    CRD_CUDA(cudaMalloc(&c->comm_local, kCommBlockBytes));
    CRD_CUDA(cudaMemset(c->comm_local, 0, kCommBlockBytes));
    CRD_CUDA(cudaMalloc(&c->comm_tab, sizeof(CommTab)));
    CRD_CUDA(cudaHostAlloc(&c->err_host, sizeof(int), cudaHostAllocMapped));
    *c->err_host = 0;
    CRD_CUDA(cudaHostGetDevicePointer(&c->err_dev, c->err_host, 0));
    CRD_CUDA(cudaEventCreate(&c->ev0));
    CRD_CUDA(cudaEventCreate(&c->ev1));
    return 0;
  };
  if (setup() != 0) { crd_ctx_destroy(c); return nullptr; }
  return c;
}

void crd_ctx_destroy(crd_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  for (int r = 0; r < kMaxRanks; ++r)
    if (c->comm_peer[r]) cudaIpcCloseMemHandle(c->comm_peer[r]);
  if (c->comm_local) cudaFree(c->comm_local);
  if (c->comm_tab) cudaFree(c->comm_tab);
  if (c->red_local) cudaFree(c->red_local);
  if (c->red_partial) cudaFree(c->red_partial);
  if (c->red_ticket) cudaFree(c->red_ticket);
  if (c->red_result_host) cudaFreeHost(c->red_result_host);
  if (c->err_host) cudaFreeHost(c->err_host);
  if (c->flush_buf) cudaFree(c->flush_buf);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  cudaGetLastError();
  delete c;
}

int crd_ctx_set_comm(crd_ctx *c, int rank, int nranks, crd_allreduce_fn fn, void *user) {
  if (!c || nranks < 1 || rank < 0 || rank >= nranks) { set_error("crd_ctx_set_comm: bad arguments"); return -1; }
  if (nranks > 1 && !fn) { set_error("crd_ctx_set_comm: nranks > 1 needs an allreduce function"); return -1; }
  c->rank = rank; c->nranks = nranks; c->allreduce = fn; c->allreduce_user = user;
  return 0;
}

void *crd_ctx_stream(crd_ctx *c) { return c ? (void *)c->stream : nullptr; }
int crd_ctx_device(crd_ctx *c) { return c ? c->device : -1; }
int64_t crd_ctx_launch_count(crd_ctx *c) { return c ? c->launches : 0; }

int crd_ctx_sync(crd_ctx *c) {
  if (!c) return -1;
  if (device_failed(c)) return -2;
  if (use(c)) return -1;
  return sync_stream(c, "crd_ctx_sync") ? -2 : 0;
}

// ---- device-side allreduce: wiring ------------------------------------------------------------------------------------
static int comm_install(crd_ctx *c, int rank, int nranks, char *const *blocks) {
  CommTab T;
  std::memset(&T, 0, sizeof T);
  for (int r = 0; r < nranks; ++r) {
    T.mail[r] = (double *)blocks[r];
    T.flag[r] = (unsigned long long *)(blocks[r] + kCommMailBytes);
  }
  T.rank = rank; T.nranks = nranks; T.timeout_ns = c->halo_timeout_ns; T.err = c->err_dev;
  CRD_CUDA(cudaMemcpy(c->comm_tab, &T, sizeof T, cudaMemcpyHostToDevice));
  c->rank = rank; c->nranks = nranks;
  c->dev_comm = nranks > 1;
  c->comm_seq = 0;
  return 0;
}

int crd_ctx_comm_handle(crd_ctx *c, unsigned char handle[CRD_HALO_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == CRD_HALO_HANDLE_BYTES, "handle size");
  if (!c || !handle) return -1;
  if (use(c)) return -1;
  cudaIpcMemHandle_t h;
  CRD_CUDA(cudaIpcGetMemHandle(&h, c->comm_local));
  std::memcpy(handle, &h, sizeof h);
  return 0;
}

int crd_ctx_comm_connect_ipc(crd_ctx *c, int rank, int nranks, const unsigned char *handles) {
  if (!c || !handles || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) { set_error("crd_ctx_comm_connect_ipc: bad arguments"); return -1; }
  if (use(c)) return -1;
  char *blocks[kMaxRanks];
  for (int r = 0; r < kMaxRanks; ++r)      // connecting again replaces the earlier mappings
    if (c->comm_peer[r]) { cudaIpcCloseMemHandle(c->comm_peer[r]); c->comm_peer[r] = nullptr; }
  c->dev_comm = false;
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) { blocks[r] = c->comm_local; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + (size_t)r * CRD_HALO_HANDLE_BYTES, sizeof h);
    void *p = nullptr;
    CRD_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->comm_peer[r] = p;
    blocks[r] = (char *)p;
  }
  return comm_install(c, rank, nranks, blocks);
}

int crd_ctx_comm_connect_local(crd_ctx *c, int rank, int nranks, crd_ctx *const *all) {
  if (!c || !all || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks || all[rank] != c) { set_error("crd_ctx_comm_connect_local: bad arguments"); return -1; }
  if (use(c)) return -1;
  char *blocks[kMaxRanks];
  for (int r = 0; r < nranks; ++r) {
    if (!all[r]) { set_error("crd_ctx_comm_connect_local: null context"); return -1; }
    if (all[r]->device != c->device) {
      int can = 0;
      CRD_CUDA(cudaDeviceCanAccessPeer(&can, c->device, all[r]->device));
      if (!can) { set_error("device %d cannot access device %d", c->device, all[r]->device); return -1; }
      cudaError_t e = cudaDeviceEnablePeerAccess(all[r]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return -1; }
      cudaGetLastError();
    }
    blocks[r] = all[r]->comm_local;
  }
  return comm_install(c, rank, nranks, blocks);
}

int crd_ctx_set_halo_timeout(crd_ctx *c, double milliseconds) {
  if (!c || !(milliseconds > 0.0)) { set_error("crd_ctx_set_halo_timeout: bad arguments"); return -1; }
  c->halo_timeout_ns = (long long)(milliseconds * 1.0e6);
  if (c->dev_comm) {   // the exchange's table carries a copy
    if (use(c)) return -1;
    CRD_CUDA(cudaMemcpy((char *)c->comm_tab + offsetof(CommTab, timeout_ns), &c->halo_timeout_ns, sizeof(long long), cudaMemcpyHostToDevice));
  }
  return 0;
}

int crd_ctx_failed(crd_ctx *c) { return c ? (device_failed(c) ? *c->err_host : 0) : -1; }

int crd_ctx_clear_error(crd_ctx *c) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  *c->err_host = 0;
  c->failed = false;
  return 0;
}

int crd_timer_start(crd_ctx *c) {
  if (use(c)) return -1;
  CRD_CUDA(cudaEventRecord(c->ev0, c->stream));
  return 0;
}
int crd_timer_stop(crd_ctx *c, double *ms) {
  if (use(c)) return -1;
  CRD_CUDA(cudaEventRecord(c->ev1, c->stream));
  CRD_CUDA(cudaEventSynchronize(c->ev1));
  float f = 0;
  CRD_CUDA(cudaEventElapsedTime(&f, c->ev0, c->ev1));
  *ms = f;
  return 0;
}

void *crd_malloc(crd_ctx *c, size_t bytes) {
  if (use(c)) return nullptr;
  void *p = nullptr;
  CRD_CUDA_NULL(cudaMalloc(&p, bytes ? bytes : 16));
  return p;
}
int crd_free(crd_ctx *c, void *p) {
  if (use(c)) return -1;
  CRD_CUDA(cudaFree(p));
  return 0;
}
void *crd_malloc_host(size_t bytes) {
  void *p = nullptr;
  CRD_CUDA_NULL(cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault));
  return p;
}
int crd_free_host(void *p) { CRD_CUDA(cudaFreeHost(p)); return 0; }

int crd_memcpy_h2d(crd_ctx *c, void *dst, const void *src, size_t bytes) {
  if (use(c)) return -1;
  CRD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return sync_stream(c, "crd_memcpy_h2d");
}
int crd_memcpy_d2h(crd_ctx *c, void *dst, const void *src, size_t bytes) {
  if (use(c)) return -1;
  CRD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  return sync_stream(c, "crd_memcpy_d2h");
}
int crd_memset_zero(crd_ctx *c, void *dst, size_t bytes) {
  if (use(c)) return -1;
  CRD_CUDA(cudaMemsetAsync(dst, 0, bytes, c->stream));
  return 0;
}

int crd_flush_l2(crd_ctx *c) {
  if (use(c)) return -1;
  if (!c->flush_buf) {
    c->flush_bytes = (size_t)256 << 20;  // 256 MiB > 126 MB L2
    CRD_CUDA(cudaMalloc(&c->flush_buf, c->flush_bytes));
  }
  CRD_CUDA(cudaMemsetAsync(c->flush_buf, 0x5a, c->flush_bytes, c->stream));
  return 0;
}

}  // extern "C"

// ---- synthetic state: jump-ahead LCG, one thread per 32-element run --------------------------------
namespace {
constexpr unsigned long long LCG_A = 6364136223846793005ULL, LCG_C = 1442695040888963407ULL;
constexpr int kRun = 32;

__global__ void __launch_bounds__(256) fill_synthetic_kernel(int fhn, unsigned long long seed, long long first, long long n,
                                                             double *__restrict__ out) {
  long long run = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long e0 = run * kRun;
  if (e0 >= n) return;
  // affine map composed (first + e0) times
  unsigned long long accA = 1ULL, accC = 0ULL, curA = LCG_A, curC = LCG_C;
  for (unsigned long long k = (unsigned long long)(first + e0); k; k >>= 1) {
    if (k & 1ULL) { accA = accA * curA; accC = accC * curA + curC; }
    curC = curC * curA + curC;
    curA = curA * curA;
  }
  unsigned long long s = accA * seed + accC;
  long long e1 = e0 + kRun < n ? e0 + kRun : n;
  for (long long e = e0; e < e1; ++e) {
    s = s * LCG_A + LCG_C;
    double u = __dmul_rn((double)(s >> 11), 1.0 / 9007199254740992.0);
    out[e] = fhn ? __dsub_rn(__dmul_rn(4.0, u), 2.0) : __dadd_rn(__dmul_rn(1.5, u), 0.1);
  }
}
}  // namespace

extern "C" int crd_fill_synthetic(crd_ctx *c, int model, uint64_t seed, int64_t first_elem, int64_t n, double *out_dev) {
  if (use(c)) return -1;
  if (n <= 0) return 0;
  int fhn = (model == CRD_FHN_TORUS || model == CRD_FHN_FLAT);
  long long runs = (n + kRun - 1) / kRun;
  unsigned int blocks = (unsigned int)((runs + 255) / 256);
  fill_synthetic_kernel<<<blocks, 256, 0, c->stream>>>(fhn, seed, first_elem, n, out_dev);
  return check_launch(c, "fill_synthetic_kernel");
}

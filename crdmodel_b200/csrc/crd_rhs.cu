// crd_rhs.cu — C ABI of the right-hand side: grid descriptor, halo ring wiring, evaluation entry points.
//
// Reference being replaced: f() of the four programs
//   src/FHNmodel_torus.cpp:504-667, src/GoldbeterModel_torus.cpp:547-724,
//   src/FHNmodel_flat.cpp:469-616,  src/GoldbeterModel_flat.cpp:515-689
// which make three sweeps over memory per evaluation (N_VConst, stencil, reaction) and re-evaluate
// sin/cos per point, and Exchange() / SetupDecomp() (src/FHNmodel_torus.cpp:708-950).  Here: ONE pass, 16 B read
// + 16 B written per grid point, metric coefficients from a per-theta table computed once on the host with the
// reference's own expressions.  The kernels live in crd_rhs_kernels.cuh; this file builds their arguments:
//   crd_grid_create        geometry exactly as main() derives it, host-libm coefficient tables, ghost block
//   crd_grid_halo_*        export / open the neighbours' ghost blocks (CUDA IPC or same-process peers)
//   crd_rhs*, crd_f*       post the boundary rows, wait, launch (single slab / ring / overlapped ring / host buffers)
#include <cstdlib>

#include "crd_rhs_kernels.cuh"
#include "crd_rhs_pair.cuh"
#include "crd_tables.hpp"


// ============================================================================================================
extern "C" {

int crd_decomp_phi(int64_t ny, int nranks, int rank, int64_t *js, int64_t *je) {
  if (nranks < 1 || rank < 0 || rank >= nranks || ny < nranks) { set_error("crd_decomp_phi: bad arguments"); return -1; }
  *js = ny * rank / nranks;
  *je = ny * (rank + 1) / nranks - 1;
  return 0;
}

crd_grid *crd_grid_create(crd_ctx *ctx, const crd_params *p) {
  if (!ctx || !p) { set_error("crd_grid_create: null argument"); return nullptr; }
  if (p->model < 0 || p->model > 3) { set_error("crd_grid_create: unknown model %d", p->model); return nullptr; }
  if (p->nx < 2 || p->ny < 2 || p->js < 0 || p->je < p->js || p->je >= p->ny) {
    set_error("crd_grid_create: bad extents nx=%lld ny=%lld js=%lld je=%lld", (long long)p->nx, (long long)p->ny,
              (long long)p->js, (long long)p->je);
    return nullptr;
  }
  if (use(ctx)) return nullptr;
  crd_grid *g = new crd_grid;
  g->ctx = ctx; g->p = *p;
  g->nx = p->nx; g->ny = p->ny; g->js = p->js; g->je = p->je; g->nyl = p->je - p->js + 1;
  std::vector<double> cth, brow;   // per-theta metric table, per-phi beta
  grid_host_tables(g, cth, brow);
  cudaError_t e;
  if ((e = cudaMalloc(&g->cth, cth.size() * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(&g->brow_alloc, brow.size() * sizeof(double))) != cudaSuccess ||
      (e = cudaMemcpy(g->cth, cth.data(), cth.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(g->brow_alloc, brow.data(), brow.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) {
    set_error("crd_grid_create: %s", cudaGetErrorString(e));
    crd_grid_destroy(g);
    return nullptr;
  }
  g->brow = g->brow_alloc + 1;   // local row 0; brow[-1] and brow[nyl] are the neighbouring ranks' rows
  // ghost block + push ticket
  HaloLayout L{g->nx};
  if ((e = cudaMalloc(&g->halo_local, L.bytes())) != cudaSuccess ||
      (e = cudaMemset(g->halo_local, 0, L.bytes())) != cudaSuccess) {
    set_error("crd_grid_create: %s", cudaGetErrorString(e));
    crd_grid_destroy(g);
    return nullptr;
  }
  return g;
}

void crd_grid_destroy(crd_grid *g) {
  if (!g) return;
  cudaSetDevice(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream);
  if (g->prev_ipc && g->halo_prev) cudaIpcCloseMemHandle(g->halo_prev);
  if (g->next_ipc && g->halo_next && g->halo_next != g->halo_prev) cudaIpcCloseMemHandle(g->halo_next);
  cudaFree(g->cth); cudaFree(g->brow_alloc); cudaFree(g->halo_local);
  if (g->fin_partial) cudaFree(g->fin_partial);
  if (g->res_bar) cudaFree(g->res_bar);
  if (g->res_partial) cudaFree(g->res_partial);
  if (g->res_out_host) cudaFreeHost(g->res_out_host);
  if (g->stage_y) cudaFree(g->stage_y);
  if (g->stage_ydot) cudaFree(g->stage_ydot);
  if (g->s_in) cudaStreamDestroy(g->s_in);
  if (g->s_out) cudaStreamDestroy(g->s_out);
  for (int i = 0; i < g->n_chunks; ++i) { cudaEventDestroy(g->ev_in[i]); cudaEventDestroy(g->ev_k[i]); }
  delete[] g->ev_in; delete[] g->ev_k;
  delete g;
}

int crd_grid_params(const crd_grid *g, crd_params *out) { if (!g || !out) return -1; *out = g->p; return 0; }
int64_t crd_grid_local_length(const crd_grid *g) { return g ? 2 * g->nx * g->nyl : 0; }
int64_t crd_grid_global_length(const crd_grid *g) { return g ? 2 * g->nx * g->ny : 0; }
double crd_grid_dx(const crd_grid *g) { return g ? g->dx : 0.0; }
double crd_grid_dy(const crd_grid *g) { return g ? g->dy : 0.0; }
int64_t crd_grid_rhs_count(const crd_grid *g) { return g ? g->rhs_count : 0; }
int crd_grid_set_variant(crd_grid *g, int variant) { if (!g) return -1; g->variant = variant; return 0; }
int crd_grid_set_overlap(crd_grid *g, int on) { if (!g) return -1; g->overlap = on != 0; return 0; }

int crd_grid_halo_handle(crd_grid *g, unsigned char handle[CRD_HALO_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == CRD_HALO_HANDLE_BYTES, "handle size");
  if (!g) return -1;
  if (use(g->ctx)) return -1;
  cudaIpcMemHandle_t h;
  CRD_CUDA(cudaIpcGetMemHandle(&h, g->halo_local));
  std::memcpy(handle, &h, sizeof h);
  return 0;
}

int crd_grid_halo_connect_ipc(crd_grid *g, const unsigned char prev_handle[CRD_HALO_HANDLE_BYTES],
                              const unsigned char next_handle[CRD_HALO_HANDLE_BYTES]) {
  if (!g) return -1;
  if (use(g->ctx)) return -1;
  cudaIpcMemHandle_t hp, hn;
  std::memcpy(&hp, prev_handle, sizeof hp);
  std::memcpy(&hn, next_handle, sizeof hn);
  void *pp = nullptr, *pn = nullptr;
  if (g->connected) {   // connecting again replaces the earlier mappings
    if (g->prev_ipc && g->halo_prev) cudaIpcCloseMemHandle(g->halo_prev);
    if (g->next_ipc && g->halo_next && g->halo_next != g->halo_prev) cudaIpcCloseMemHandle(g->halo_next);
    g->halo_prev = g->halo_next = nullptr;
    g->prev_ipc = g->next_ipc = g->connected = false;
  }
  CRD_CUDA(cudaIpcOpenMemHandle(&pp, hp, cudaIpcMemLazyEnablePeerAccess));
  if (std::memcmp(&hp, &hn, sizeof hp) == 0) pn = pp;  // two ranks: both neighbours are the same block
  else {
    cudaError_t e = cudaIpcOpenMemHandle(&pn, hn, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaIpcCloseMemHandle(pp);
      set_error("crd_grid_halo_connect_ipc: cudaIpcOpenMemHandle -> %s", cudaGetErrorString(e));
      return -1;
    }
  }
  g->halo_prev = (char *)pp; g->halo_next = (char *)pn;
  g->prev_ipc = g->next_ipc = true;
  g->connected = true;
  return 0;
}

int crd_grid_halo_connect_local(crd_grid *g, crd_grid *prev, crd_grid *next) {
  if (!g || !prev || !next) return -1;
  if (prev->nx != g->nx || next->nx != g->nx) { set_error("halo_connect_local: theta mesh differs"); return -1; }
  if (use(g->ctx)) return -1;
  for (crd_grid *o : {prev, next}) {
    if (o->ctx->device != g->ctx->device) {
      int can = 0;
      CRD_CUDA(cudaDeviceCanAccessPeer(&can, g->ctx->device, o->ctx->device));
      if (!can) { set_error("device %d cannot access device %d", g->ctx->device, o->ctx->device); return -1; }
      cudaError_t e = cudaDeviceEnablePeerAccess(o->ctx->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return -1; }
      cudaGetLastError();
    }
  }
  g->halo_prev = prev->halo_local; g->halo_next = next->halo_local;
  g->prev_ipc = g->next_ipc = false;
  g->connected = true;
  return 0;
}

// ---- halo exchange of a phi-split grid (what Exchange() does with MPI_Isend / Irecv / Wait, FHNmodel_torus.cpp:775-950) ----
// Every evaluation has an epoch; the ghost rows are double-buffered by its parity (a rank can be at most one evaluation ahead
// of a neighbour: it needs the neighbour's rows of epoch e to finish e).  Two ways to run one:
//   in the launch   (default; tiled and streaming kernels) ONE launch per evaluation: its first CTAs push this slab's boundary
//                   rows into the neighbours' ghost rows and publish them strip by strip, the tiles / row segments that touch
//                   a ghost row are scheduled last and acquire their strip's flag — the exchange overlaps the interior rows;
//   as launches     (direct kernel of the small meshes, crd_grid_set_overlap(g, 0), crd_rhs_post_halo): push kernel, wait
//                   kernel, evaluation.
static HaloSync halo_sync(crd_grid *g, bool push, bool wait) {
  HaloLayout L{g->nx};
  const int par = (int)(g->epoch & 1ULL);
  HaloSync h;
  if (push) {
    h.push_prev = (double *)(g->halo_prev + L.ghost_off(par, 1));      // my row js is the row above prev's je
    h.push_next = (double *)(g->halo_next + L.ghost_off(par, 0));      // my row je is the row below next's js
    h.flag_prev = (unsigned long long *)(g->halo_prev + L.flag_off(1));
    h.flag_next = (unsigned long long *)(g->halo_next + L.flag_off(0));
  }
  if (wait) {
    h.wait_south = (const unsigned long long *)(g->halo_local + L.flag_off(0));
    h.wait_north = (const unsigned long long *)(g->halo_local + L.flag_off(1));
  }
  h.epoch = g->epoch;
  h.timeout_ns = g->ctx->halo_timeout_ns;
  h.err = g->ctx->err_dev;
  return h;
}

static int launch_push(crd_grid *g, const StateRef &S, cudaStream_t st) {
  g->epoch++;
  RhsArgs a = make_args(g, 0.0, S, nullptr, 0, g->nyl, slab_row(0), slab_row(0));
  a.hs = halo_sync(g, true, false);
  const unsigned nb = (unsigned)HaloLayout{g->nx}.nstrips();
  if (S.n > 0 && g->p.arith == CRD_ARITH_EXACT) halo_push_kernel<true, true><<<nb, 256, 0, st>>>(a);
  else if (S.n > 0) halo_push_kernel<true, false><<<nb, 256, 0, st>>>(a);
  else halo_push_kernel<false, false><<<nb, 256, 0, st>>>(a);
  return check_launch(g->ctx, "halo_push_kernel");
}

static int launch_wait(crd_grid *g, cudaStream_t st) {
  RhsArgs a;
  std::memset(&a, 0, sizeof a);
  a.nx = g->nx;
  a.hs = halo_sync(g, false, true);
  halo_wait_kernel<<<1, 256, 0, st>>>(a);
  return check_launch(g->ctx, "halo_wait_kernel");
}

// does the kernel that will evaluate this state carry the exchange inside its launch?
static bool exchange_in_launch(const crd_grid *g, int nlc, bool finish) {
  if (!g->overlap) return false;
  return finish || kernel_has_halo(resolved_variant(g, g->nx, g->nyl, nlc));
}

static int post_state(crd_grid *g, const StateRef &S) {
  if (!g->connected) return 0;  // single rank: the slab wraps onto itself
  if (g->epoch != g->computed) { set_error("crd_rhs_post_halo: previous epoch was posted but never computed"); return -1; }
  return launch_push(g, S, g->ctx->stream);
}

// Last stage fused with the step finish: every launch of the evaluation writes its per-CTA sums into its own region.
struct FinCtx {
  StageFin fin;          // fin.partial = base of region 0
  int nregions = 0;
  int nblocks[3] = {0, 0, 0};
};
constexpr int kFinRegion = 3 * kRedBlocks;   // doubles per region: [3][kRedBlocks]

static int launch_part(crd_grid *g, const RhsArgs &a, cudaStream_t st, FinCtx *fc) {
  if (!fc) return launch_rhs(g, a, st);
  StageFin f = fc->fin;
  f.partial = fc->fin.partial + (size_t)fc->nregions * kFinRegion;
  const int r = launch_stage_finish(g, a, f, st, &fc->nblocks[fc->nregions]);
  if (r != 0) { if (r > 0) set_error("fused stage finish: kernel not available for this launch"); return -1; }
  fc->nregions++;
  return 0;
}

// posted: the boundary rows of this evaluation were pushed by a launch of their own (crd_rhs_post_halo) — otherwise, on a
// connected grid, this call starts the epoch itself
static int compute_state(crd_grid *g, double t, const StateRef &S, double *ydot, bool posted, FinCtx *fc = nullptr) {
  cudaStream_t st = g->ctx->stream;
  const long long nyl = g->nyl;
  if (!g->connected) {
    RhsArgs a = make_args(g, t, S, ydot, 0, nyl, slab_row(nyl - 1), slab_row(0));
    if (launch_part(g, a, st, fc)) return -1;
    g->rhs_count++;
    return 0;
  }
  const bool fused = exchange_in_launch(g, S.n, fc != nullptr);
  if (posted) {
    if (g->epoch == g->computed) { set_error("crd_rhs_compute: no halo posted for this evaluation"); return -1; }
  } else {
    if (g->epoch != g->computed) { set_error("crd_rhs: previous epoch was posted but never computed"); return -1; }
    if (fused) g->epoch++;                              // the launch below pushes
    else if (launch_push(g, S, st)) return -1;
  }
  HaloLayout L{g->nx};
  const int par = (int)(g->epoch & 1ULL);
  const double *gs = (const double *)(g->halo_local + L.ghost_off(par, 0));
  const double *gn = (const double *)(g->halo_local + L.ghost_off(par, 1));
  RhsArgs a = make_args(g, t, S, ydot, 0, nyl, ext_row(gs), ext_row(gn));
  if (fused) {
    a.hs = halo_sync(g, !posted, true);
    a.hs.edge_last = 1;
  } else if (launch_wait(g, st)) return -1;
  if (launch_part(g, a, st, fc)) return -1;
  g->computed = g->epoch;
  g->rhs_count++;
  return 0;
}

static StateRef plain_state(const double *y) { StateRef S; S.y = y; S.n = 0; return S; }

int crd_rhs_post_halo(crd_grid *g, const double *y) {
  if (!g || !y) { set_error("crd_rhs_post_halo: null argument"); return -1; }
  if (use(g->ctx)) return -1;
  return post_state(g, plain_state(y));
}

int crd_rhs_compute(crd_grid *g, double t, const double *y, double *ydot) {
  if (!g || !y || !ydot) { set_error("crd_rhs_compute: null argument"); return -1; }
  if (use(g->ctx)) return -1;
  return compute_state(g, t, plain_state(y), ydot, true);
}

int crd_rhs(crd_grid *g, double t, const double *y, double *ydot) {
  if (!g || !y || !ydot) { set_error("crd_rhs: null argument"); return -1; }
  if (use(g->ctx)) return -1;
  return compute_state(g, t, plain_state(y), ydot, false);
}

// ydot = f(t, sum_j c[j]*X[j]) without materialising the combination (explicit RK stage assembly fused into the
// evaluation).  n <= 5; the vectors must not alias ydot.
int crd_rhs_lincomb(crd_grid *g, double t, int n, const double *c, const double *const *X_dev, double *ydot) {
  if (!g || !c || !X_dev || !ydot || n < 1 || n > kMaxLc) { set_error("crd_rhs_lincomb: bad arguments"); return -1; }
  if (use(g->ctx)) return -1;
  StateRef S;
  S.n = n;
  for (int j = 0; j < n; ++j) {
    if (!X_dev[j] || X_dev[j] == ydot) { set_error("crd_rhs_lincomb: null or aliased vector"); return -1; }
    S.x[j] = X_dev[j]; S.c[j] = c[j];
  }
  return compute_state(g, t, S, ydot, false);
}

int crd_f_lincomb(realtype t, int n, const realtype *c, N_Vector *X, N_Vector ydot, void *user_data) {
  crd_grid *g = (crd_grid *)user_data;
  if (!g || !X || !ydot || n < 1 || n > kMaxLc) return -1;
  const double *xs[kMaxLc];
  for (int j = 0; j < n; ++j) {
    xs[j] = N_VGetDeviceArrayPointer_Crd(X[j]);
    if (!xs[j] || N_VGetLocalLength_Crd(X[j]) != crd_grid_local_length(g)) { set_error("crd_f_lincomb: vector does not match the grid"); return -1; }
  }
  return crd_rhs_lincomb(g, t, n, c, xs, N_VGetDeviceArrayPointer_Crd(ydot)) == 0 ? 0 : -1;
}

// The last stage of an s-stage explicit RK step fused with the step finish: with X = (yn, F_0 .. F_{s-2}) and
// F_{s-1} = f(t, sum_j c[j] X[j]) (never stored),  ynew = yn + sum_j hb[j] F_j,  err = sum_j hd[j] F_j,
// out[0] = sum (err_i w_i)^2, out[1] = sum (ynew_i w'_i)^2 as N_VErkFinish_Crd defines them (global sums on a phi-split grid,
// whose halo exchange rides inside the same launch).  ynew has the bits of crd_rhs_lincomb followed by N_VErkFinish_Crd /
// N_VErkFinishSeq_Crd (FAST / EXACT grid).  Returns 1 when it does not apply (not 5 stages, a mesh too small to stream) — decided
// before anything is sent to the neighbours: the caller then issues the two separate operations.
int crd_rhs_lincomb_finish(crd_grid *g, double t, int s, const double *c, const double *hb, const double *hd,
                           const double *const *X_dev, double *ynew_dev, double rtol, double atol, double out[2]) {
  if (!g || !c || !hb || !hd || !X_dev || !ynew_dev || !out) { set_error("crd_rhs_lincomb_finish: null argument"); return -1; }
  // decided before anything is posted to the neighbours: 5 stages, a mesh the streaming kernel is made for, non-zero solution
  // weights of the stored stages (the kernel adds those terms unconditionally; the op-by-op chain skips zero weights)
  // (on a phi-split grid the size test uses the smallest slab of the split, so that every rank comes to the same answer)
  const long long rows_min = g->connected ? g->ny / (g->ctx->nranks > 1 ? g->ctx->nranks : 1) : g->nyl;
  if (s != kMaxLc || g->nx < 192 || g->nx * rows_min < (1LL << 20)) return 1;
  for (int j = 0; j + 1 < s; ++j)
    if (hb[j] == 0.0) return 1;
  crd_ctx *ctx = g->ctx;
  if (use(ctx)) return -1;
  StateRef S;
  S.n = s;
  for (int j = 0; j < s; ++j) {
    if (!X_dev[j] || X_dev[j] == ynew_dev) { set_error("crd_rhs_lincomb_finish: null or aliased vector"); return -1; }
    S.x[j] = X_dev[j]; S.c[j] = c[j];
  }
  if (!g->fin_partial) CRD_CUDA(cudaMalloc(&g->fin_partial, sizeof(double) * 3 * kFinRegion));
  FinCtx fc;
  for (int j = 0; j < kMaxLc; ++j) { fc.fin.hb[j] = hb[j]; fc.fin.hd[j] = hd[j]; }
  fc.fin.rtol = rtol; fc.fin.atol = atol; fc.fin.partial = g->fin_partial;
  fc.fin.hb_nz = finish_nz_mask(hb, s); fc.fin.y2_bound = finish_y2_bound(rtol, g->nx * g->nyl * 2);
  if (compute_state(g, t, S, ynew_dev, false, &fc)) return -1;
  // add the per-CTA sums in a fixed order
  // ... and, with the device-side allreduce wired, exchange them with the other ranks in the same launch
  const CommTab *tab = ctx->dev_comm ? ctx->comm_tab : nullptr;
  if (tab) ++ctx->comm_seq;
  fin_reduce_kernel<<<1, 256, 0, ctx->stream>>>(g->fin_partial, fc.nregions, fc.nblocks[0], fc.nblocks[1], fc.nblocks[2], ctx->red_result_dev,
                                                tab, ctx->comm_seq);
  if (check_launch(ctx, "fin_reduce_kernel")) return -1;
  if (sync_stream(ctx, "crd_rhs_lincomb_finish")) return -1;
  double hi = ctx->red_result_host[0], y2 = ctx->red_result_host[1], lo = ctx->red_result_host[2];
  if (allreduce_dd(ctx, hi, lo, &y2)) return -1;   // the ranks' pairs merged in rank order: same bits on every rank and for any split
  out[0] = hi + lo;
  out[1] = y2;
  return 0;
}

int crd_f_lincomb_finish(realtype t, int s, const realtype *c, const realtype *hb, const realtype *hd, N_Vector *X, N_Vector ynew,
                         realtype rtol, realtype atol, realtype out[2], void *user_data) {
  crd_grid *g = (crd_grid *)user_data;
  if (!g || !X || !ynew || s < 1 || s > kMaxLc) return -1;
  const double *xs[kMaxLc];
  for (int j = 0; j < s; ++j) {
    xs[j] = N_VGetDeviceArrayPointer_Crd(X[j]);
    if (!xs[j] || N_VGetLocalLength_Crd(X[j]) != crd_grid_local_length(g)) { set_error("crd_f_lincomb_finish: vector does not match the grid"); return -1; }
  }
  if (N_VGetLocalLength_Crd(ynew) != crd_grid_local_length(g)) { set_error("crd_f_lincomb_finish: vector does not match the grid"); return -1; }
  return crd_rhs_lincomb_finish(g, t, s, c, hb, hd, xs, N_VGetDeviceArrayPointer_Crd(ynew), rtol, atol, out);
}

// f1 = f(t1, y) and f2 = f(t2, y + c f1) in one pass over y (crd_rhs_pair.cuh): the derivative at an accepted state together with
// the second stage of the next step.  Bits of crd_rhs followed by crd_rhs_lincomb(2, (1, c), (y, f1)).  Returns 1 when it does
// not apply (a slab too small, a forced kernel variant): the caller issues the two evaluations.  On a phi-split grid the
// neighbours first exchange two rows of y per side (one exchange for both evaluations).
int crd_rhs_pair(crd_grid *g, double t1, double t2, double c, const double *y, double *f1, double *f2) {
  if (!g || !y || !f1 || !f2) { set_error("crd_rhs_pair: null argument"); return -1; }
  // Does it apply?  On a phi-split grid every rank must come to the same answer (the pass makes ONE exchange where the two
  // evaluations make two: ranks that disagreed would wait for rows that never come), so the size test there uses nothing but
  // the global mesh and the number of ranks — the smallest slab of the split, whatever this rank's own share is.
  long long rows = g->nyl;
  if (g->connected) {
    const long long R = g->ctx->nranks > 1 ? g->ctx->nranks : 1;
    rows = g->ny / R;
  }
  if (g->nx < kPairCols || rows < 32 || g->nyl < 2 || g->nx * rows < (1LL << 20) || g->variant != 0) return 1;
  if (y == f1 || y == f2 || f1 == f2) { set_error("crd_rhs_pair: aliased vectors"); return -1; }
  if (use(g->ctx)) return -1;
  if (g->connected && g->epoch != g->computed) { set_error("crd_rhs_pair: previous epoch was posted but never computed"); return -1; }
  PairArgs a;
  a.south_near = a.south_far = a.north_near = a.north_far = nullptr;
  a.nx = g->nx; a.nyl = g->nyl;
  if (g->connected) {
    // one exchange for both evaluations: two rows of y per side (the first evaluation on the adjacent ghost row is repeated
    // locally), as launches of their own in front of the pass.  Every kernel of the sequence is loaded before the first of them
    // is launched: the neighbours' wait kernels spin on this rank's rows, and a module load in between could wait for them.
    if (launch_pair(g, a, g->ctx->stream, true)) return -1;
    {
      static bool loaded[64] = {};
      const int dev = g->ctx->device & 63;
      if (!loaded[dev]) {
        cudaFuncAttributes fa;
        CRD_CUDA(cudaFuncGetAttributes(&fa, halo_push2_kernel));
        CRD_CUDA(cudaFuncGetAttributes(&fa, halo_wait_kernel));
        loaded[dev] = true;
      }
    }
    g->epoch++;
    HaloLayout L{g->nx};
    const int par = (int)(g->epoch & 1ULL);
    RhsArgs h;
    std::memset(&h, 0, sizeof h);
    h.y = y; h.nx = g->nx; h.nyl = g->nyl;
    h.hs = halo_sync(g, true, false);
    h.hs.push2_prev = (double *)(g->halo_prev + L.ghost2_off(par, 1));
    h.hs.push2_next = (double *)(g->halo_next + L.ghost2_off(par, 0));
    halo_push2_kernel<<<(unsigned)L.nstrips(), 256, 0, g->ctx->stream>>>(h);
    if (check_launch(g->ctx, "halo_push2_kernel")) return -1;
    if (launch_wait(g, g->ctx->stream)) return -1;
    a.south_near = (const double *)(g->halo_local + L.ghost_off(par, 0));
    a.south_far = (const double *)(g->halo_local + L.ghost2_off(par, 0));
    a.north_near = (const double *)(g->halo_local + L.ghost_off(par, 1));
    a.north_far = (const double *)(g->halo_local + L.ghost2_off(par, 1));
    g->computed = g->epoch;
  }
  a.y = y; a.f1 = f1; a.f2 = f2;
  a.cth = g->cth; a.brow = g->brow;
  a.nx = g->nx; a.nyl = g->nyl;
  a.k = g->k;
  a.c[0] = 1.0; a.c[1] = c;
  a.react = (is_fhn(g->p.model) || g->p.just_diffusion == 0) ? 1 : 0;
  // rows held at zero while t < tBoundary: global rows 0 and ny - 1 — local rows 0 / nyl - 1 of the first / last rank, and, seen
  // from the other side of the periodic seam, the ghost row north of the last rank / south of the first
  const int south = g->js == 0 ? 1 : 0, north = g->je == g->ny - 1 ? 2 : 0;
  const int ghosts = g->connected ? ((g->js == 0 ? 4 : 0) | (g->je == g->ny - 1 ? 8 : 0)) : 0;
  a.frz1 = t1 < g->p.t_boundary ? (south | north | ghosts) : 0;
  a.frz2 = t2 < g->p.t_boundary ? (south | north) : 0;
  if (launch_pair(g, a, g->ctx->stream)) return -1;
  g->rhs_count += 2;
  return 0;
}

// the integrator's entry (crd_fused_ops.rhs_pair): only where one pass beats the two launches — every FAST grid (x1.45-1.8) and
// EXACT grids of the FHN programs (x1.1-1.35); the Goldbeter kinetics in EXACT arithmetic are bound by FP64 work and the pass is
// slower than the two launches there (crd_rhs_pair.cuh, profiles/README.md)
int crd_f_pair(realtype t1, realtype t2, realtype c, N_Vector y, N_Vector f1, N_Vector f2, void *user_data) {
  crd_grid *g = (crd_grid *)user_data;
  if (!g || !y || !f1 || !f2) return -1;
  if (g->p.arith != CRD_ARITH_FAST && !is_fhn(g->p.model)) return 1;
  const long long len = crd_grid_local_length(g);
  if (N_VGetLocalLength_Crd(y) != len || N_VGetLocalLength_Crd(f1) != len || N_VGetLocalLength_Crd(f2) != len) {
    set_error("crd_f_pair: vector does not match the grid");
    return -1;
  }
  return crd_rhs_pair(g, t1, t2, c, N_VGetDeviceArrayPointer_Crd(y), N_VGetDeviceArrayPointer_Crd(f1), N_VGetDeviceArrayPointer_Crd(f2));
}

int crd_f(realtype t, N_Vector y, N_Vector ydot, void *user_data) {
  crd_grid *g = (crd_grid *)user_data;
  if (!g || !y || !ydot) return -1;
  const double *yd = N_VGetDeviceArrayPointer_Crd(y);
  double *fd = N_VGetDeviceArrayPointer_Crd(ydot);
  if (!yd || !fd) return -1;
  if (N_VGetLocalLength_Crd(y) != crd_grid_local_length(g)) { set_error("crd_f: vector length does not match the grid"); return -1; }
  return crd_rhs(g, t, yd, fd) == 0 ? 0 : -1;
}

// Host-buffer entry: stream the slab in row chunks, H2D / kernel / D2H on three streams.
int crd_rhs_host(crd_grid *g, double t, const double *y_host, double *ydot_host) {
  if (!g || !y_host || !ydot_host) { set_error("crd_rhs_host: null argument"); return -1; }
  if (g->connected && g->epoch != g->computed) { set_error("crd_rhs_host: previous epoch was posted but never computed"); return -1; }
  if (use(g->ctx)) return -1;
  const long long nx = g->nx, nyl = g->nyl;
  const size_t row_bytes = (size_t)2 * nx * sizeof(double);
  if (!g->stage_y) {
    CRD_CUDA(cudaMalloc(&g->stage_y, row_bytes * nyl));
    CRD_CUDA(cudaMalloc(&g->stage_ydot, row_bytes * nyl));
    CRD_CUDA(cudaStreamCreateWithFlags(&g->s_in, cudaStreamNonBlocking));
    CRD_CUDA(cudaStreamCreateWithFlags(&g->s_out, cudaStreamNonBlocking));
    // chunks of ~64 MiB (CRD_HOST_CHUNK_MB), at least 1 row, at most 64 chunks
    unsigned long long chunk_mb = 64;
    if (const char *e = std::getenv("CRD_HOST_CHUNK_MB")) { const long v = std::atol(e); if (v > 0) chunk_mb = (unsigned long long)v; }
    long long rows_per = (long long)((chunk_mb << 20) / row_bytes);
    if (rows_per < 1) rows_per = 1;
    long long n = (nyl + rows_per - 1) / rows_per;
    if (n > 64) n = 64;
    if (n < 1) n = 1;
    g->n_chunks = (int)n;
    g->ev_in = new cudaEvent_t[n]; g->ev_k = new cudaEvent_t[n];
    for (int i = 0; i < n; ++i) {
      CRD_CUDA(cudaEventCreateWithFlags(&g->ev_in[i], cudaEventDisableTiming));
      CRD_CUDA(cudaEventCreateWithFlags(&g->ev_k[i], cudaEventDisableTiming));
    }
  }
  const int C = g->n_chunks;
  cudaStream_t sk = g->ctx->stream;
  auto r_begin = [&](int c) { return nyl * c / C; };
  // the periodic wrap makes chunk 0 need the last row: send it first
  CRD_CUDA(cudaMemcpyAsync(g->stage_y + 2 * (nyl - 1) * nx, y_host + 2 * (nyl - 1) * nx, row_bytes, cudaMemcpyHostToDevice, g->s_in));
  const double *ghost_s = nullptr, *ghost_n = nullptr;
  if (g->connected) {
    // ring: the neighbours need this slab's first and last row before anything else
    CRD_CUDA(cudaMemcpyAsync(g->stage_y, y_host, row_bytes, cudaMemcpyHostToDevice, g->s_in));
    CRD_CUDA(cudaEventRecord(g->ev_k[0], g->s_in));
    CRD_CUDA(cudaStreamWaitEvent(sk, g->ev_k[0], 0));
    if (launch_push(g, plain_state(g->stage_y), sk)) return -1;
    if (launch_wait(g, sk)) return -1;
    HaloLayout L{g->nx};
    const int par = (int)(g->epoch & 1ULL);
    ghost_s = (const double *)(g->halo_local + L.ghost_off(par, 0));
    ghost_n = (const double *)(g->halo_local + L.ghost_off(par, 1));
    g->computed = g->epoch;
  }
  for (int c = 0; c < C; ++c) {
    const long long r0 = r_begin(c), r1 = r_begin(c + 1);
    CRD_CUDA(cudaMemcpyAsync(g->stage_y + 2 * r0 * nx, y_host + 2 * r0 * nx, row_bytes * (r1 - r0), cudaMemcpyHostToDevice, g->s_in));
    CRD_CUDA(cudaEventRecord(g->ev_in[c], g->s_in));
  }
  for (int c = 0; c < C; ++c) {
    const long long r0 = r_begin(c), r1 = r_begin(c + 1);
    if (r1 == r0) continue;
    CRD_CUDA(cudaStreamWaitEvent(sk, g->ev_in[c + 1 < C ? c + 1 : c], 0));
    const RowRef south = (ghost_s && r0 == 0) ? ext_row(ghost_s) : slab_row((r0 == 0 ? nyl : r0) - 1);
    const RowRef north = (ghost_n && r1 == nyl) ? ext_row(ghost_n) : slab_row(r1 == nyl ? 0 : r1);
    RhsArgs a = make_args(g, t, plain_state(g->stage_y), g->stage_ydot, r0, r1, south, north);
    if (launch_rhs(g, a, sk)) return -1;
    CRD_CUDA(cudaEventRecord(g->ev_k[c], sk));
    CRD_CUDA(cudaStreamWaitEvent(g->s_out, g->ev_k[c], 0));
    CRD_CUDA(cudaMemcpyAsync(ydot_host + 2 * r0 * nx, g->stage_ydot + 2 * r0 * nx, row_bytes * (r1 - r0), cudaMemcpyDeviceToHost, g->s_out));
  }
  CRD_CUDA(cudaStreamSynchronize(g->s_out));
  if (sync_stream(g->ctx, "crd_rhs_host")) return -1;   // (also the place where a neighbour's rows that never arrived surface)
  g->rhs_count++;
  return 0;
}

int crd_fill_initial_conditions(crd_grid *g, const crd_ic_params *ic, double *y_dev) {
  if (!g || !ic || !y_dev) { set_error("crd_fill_initial_conditions: null argument"); return -1; }
  if (use(g->ctx)) return -1;
  IcArgs a;
  a.model = g->p.model; a.vary_beta = g->p.vary_beta; a.wave_inside = ic->wave_inside; a.ic_type = ic->ic_type;
  a.nx = g->nx; a.nyl = g->nyl; a.js = g->js;
  a.dx = g->dx; a.dy = g->dy; a.xmin = g->xmin; a.ymin = g->ymin;
  // FHNmodel_torus.cpp:199-200,285-296 ; FHNmodel_flat.cpp:274-276
  a.wave_length = (g->ymax - g->ymin) * ic->wave_length;
  a.wave_width = (g->xmax - g->xmin) * ic->wave_width;
  const bool torus = is_torus(g->p.model), fhn = is_fhn(g->p.model);
  double mid;
  if (torus) {
    if (ic->wave_inside == 1) { mid = kPI; a.wave_xmin = mid - a.wave_width / 2.0; a.wave_xmax = mid + a.wave_width / 2.0; }
    else { mid = 0.0; a.wave_xmin = mid - a.wave_width / 2.0 + (g->xmax - g->xmin); a.wave_xmax = mid + a.wave_width / 2.0; }
  } else {
    mid = g->p.surface_width / 2.0; a.wave_xmin = mid - a.wave_width / 2.0; a.wave_xmax = mid + a.wave_width / 2.0;
  }
  a.s0 = ic->s0; a.s1 = ic->s1;
  a.p0 = fhn ? ic->s0 + 2 : ic->s0 + 1;      // Us + 2 | Zs + 1
  a.p1 = fhn ? ic->s1 + 1.5 : ic->s1 + 1;    // Vs + 1.5 | Ys + 1
  if (!fhn && g->p.vary_beta == 1 && ic->ic_type == 2) {
    set_error("icType = 2 (unseeded rand(), GoldbeterModel_flat.cpp:373-374) must be generated on the host");
    return -1;
  }
  const long long work = g->nx * g->nyl;
  ic_kernel<<<(unsigned)((work + 255) / 256), 256, 0, g->ctx->stream>>>(a, reinterpret_cast<double2 *>(y_dev));
  return check_launch(g->ctx, "ic_kernel");
}

}  // extern "C"

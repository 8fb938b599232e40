// crd_nvector.cu — device-resident N_Vector: the operations SUNDIALS' nvector_parallel gives the
// reference (call sites src/FHNmodel_torus.cpp:281,303,383,488,506,517-518,786 and every vector op
// ARKode performs inside ARKode(), :423), as sm_100a kernels.
//
// Streaming ops: one pass, 16-byte vector accesses, grid sized to the SM count, per-element
// arithmetic rounded exactly like the serial loops of nvector_parallel.c (the special cases of
// N_VLinearSum / N_VScale included, no FMA contraction) so element-wise results are bit-identical to
// the CPU checker.  Reductions: thread-strided partials -> warp shuffle -> block -> the last block
// combines the per-block partials in a fixed order (deterministic), writes the result into mapped
// pinned host memory; the host then applies the cross-rank allreduce hook (one 8-byte value).
#include <cfloat>
#include <cmath>
#include <new>

#include "crd_common.cuh"
#include "crd_fused.cuh"

using namespace crd;

struct crd_nv_content {
  long int local_length;
  long int global_length;
  booleantype own_data;
  realtype *data;       // device
  realtype *host;       // pinned mirror, lazily allocated
  crd_ctx *ctx;
};

namespace {

inline crd_nv_content *NVC(N_Vector v) { return (crd_nv_content *)v->content; }
inline double *D(N_Vector v) { return NVC(v)->data; }
inline long long LEN(N_Vector v) { return NVC(v)->local_length; }
inline crd_ctx *CTX(N_Vector v) { return NVC(v)->ctx; }

inline unsigned int grid_for(long long n_items, int sms = crd::kSMs) {
  long long b = (n_items + 255) / 256;
  const long long cap = (long long)sms * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned int)b;
}

// ---- element-wise --------------------------------------------------------------------------------------
// F: functor with  double operator()(double x, double y) ; x,y may be unused.
template <class F, bool HAS_X, bool HAS_Y>
__global__ void __launch_bounds__(256) ew_kernel(F f, const double *x, const double *y, double *z, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool aligned = (((uintptr_t)x | (uintptr_t)y | (uintptr_t)z) & 15) == 0;
  if (aligned) {
    const long long n2 = n >> 1;
    const double2 *x2 = reinterpret_cast<const double2 *>(x);
    const double2 *y2 = reinterpret_cast<const double2 *>(y);
    double2 *z2 = reinterpret_cast<double2 *>(z);
#pragma unroll 4
    for (long long i = tid; i < n2; i += stride) {
      double2 a = HAS_X ? x2[i] : make_double2(0, 0);
      double2 b = HAS_Y ? y2[i] : make_double2(0, 0);
      z2[i] = make_double2(f(a.x, b.x), f(a.y, b.y));
    }
    if (tid == 0 && (n & 1)) z[n - 1] = f(HAS_X ? x[n - 1] : 0.0, HAS_Y ? y[n - 1] : 0.0);
  } else {
    for (long long i = tid; i < n; i += stride) z[i] = f(HAS_X ? x[i] : 0.0, HAS_Y ? y[i] : 0.0);
  }
}

template <class F, bool HAS_X, bool HAS_Y>
void ew(crd_ctx *c, F f, const double *x, const double *y, double *z, long long n, const char *name) {
  if (n <= 0) return;
  if (use(c)) return;
  ew_kernel<F, HAS_X, HAS_Y><<<grid_for((n + 1) / 2, c->sms), 256, 0, c->stream>>>(f, x, y, z, n);
  check_launch(c, name);
}

struct FConst { double c; __device__ double operator()(double, double) const { return c; } };
struct FCopy { __device__ double operator()(double x, double) const { return x; } };
struct FNeg { __device__ double operator()(double x, double) const { return -x; } };
struct FScale { double c; __device__ double operator()(double x, double) const { return __dmul_rn(c, x); } };
struct FAbs { __device__ double operator()(double x, double) const { return fabs(x); } };
struct FInv { __device__ double operator()(double x, double) const { return __ddiv_rn(1.0, x); } };
struct FAddC { double b; __device__ double operator()(double x, double) const { return __dadd_rn(x, b); } };
struct FProd { __device__ double operator()(double x, double y) const { return __dmul_rn(x, y); } };
struct FDiv { __device__ double operator()(double x, double y) const { return __ddiv_rn(x, y); } };
struct FCmp { double c; __device__ double operator()(double x, double) const { return fabs(x) >= c ? 1.0 : 0.0; } };
// N_VLinearSum cases of nvector_parallel.c
struct FSum { __device__ double operator()(double x, double y) const { return __dadd_rn(x, y); } };
struct FDiff { __device__ double operator()(double x, double y) const { return __dsub_rn(x, y); } };           // x - y
struct FLin1 { double a; __device__ double operator()(double x, double y) const { return __dadd_rn(__dmul_rn(a, x), y); } };   // a*x + y
struct FLin2 { double a; __device__ double operator()(double x, double y) const { return __dsub_rn(__dmul_rn(a, x), y); } };   // a*x - y
struct FScaleSum { double c; __device__ double operator()(double x, double y) const { return __dmul_rn(c, __dadd_rn(x, y)); } };
struct FScaleDiff { double c; __device__ double operator()(double x, double y) const { return __dmul_rn(c, __dsub_rn(x, y)); } };
struct FLinSum { double a, b; __device__ double operator()(double x, double y) const { return __dadd_rn(__dmul_rn(a, x), __dmul_rn(b, y)); } };

// ---- reductions ---------------------------------------------------------------------------------------
enum { OP_SUM = 0, OP_MAX = 1, OP_MIN = 2 };
template <int OP> __device__ __forceinline__ double comb(double a, double b) {
  return OP == OP_SUM ? a + b : (OP == OP_MAX ? fmax(a, b) : fmin(a, b));
}
template <int OP> __device__ __forceinline__ double ident() {
  return OP == OP_SUM ? 0.0 : (OP == OP_MAX ? -DBL_MAX : DBL_MAX);
}

template <int OP, int NV>
__device__ __forceinline__ void block_finish(double (&acc)[NV], double *partial, unsigned int *ticket, double *result) {
  __shared__ double sm[NV][kRedThreads / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[v] = comb<OP>(acc[v], __shfl_down_sync(0xffffffffu, acc[v], o));
    if (lane == 0) sm[v][wid] = acc[v];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double s = sm[v][0];
      for (int w = 1; w < kRedThreads / 32; ++w) s = comb<OP>(s, sm[v][w]);
      partial[v * kRedBlocks + blockIdx.x] = s;
    }
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  // last block: fixed-order combination of the per-block partials
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double s = ident<OP>();
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += kRedThreads) s = comb<OP>(s, __ldcg(&partial[v * kRedBlocks + b]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = comb<OP>(s, __shfl_down_sync(0xffffffffu, s, o));
    if (lane == 0) sm[v][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double s = sm[v][0];
      for (int w = 1; w < kRedThreads / 32; ++w) s = comb<OP>(s, sm[v][w]);
      result[v] = s;   // mapped pinned host memory
    }
    *ticket = 0;
    __threadfence_system();
  }
}

// ---- order-independent sums of squares: double-double accumulation (crd_fused.cuh) -------------------------------------
// partial layout [3][kRedBlocks]: hi | plain second value | lo; result[0] = hi, result[1] = plain, result[2] = lo
template <bool HAS_PLAIN>
__device__ __forceinline__ void block_finish_dd(double hi, double lo, double pl, double *partial, unsigned int *ticket, double *result) {
  __shared__ double sm[3][kRedThreads / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dd_shfl_down(hi, lo, o);
    if (HAS_PLAIN) pl += __shfl_down_sync(0xffffffffu, pl, o);
  }
  if (lane == 0) { sm[0][wid] = hi; sm[1][wid] = pl; sm[2][wid] = lo; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double h = sm[0][0], p = sm[1][0], l = sm[2][0];
    for (int w = 1; w < kRedThreads / 32; ++w) { dd_merge(h, l, sm[0][w], sm[2][w]); p += sm[1][w]; }
    partial[blockIdx.x] = h; partial[kRedBlocks + blockIdx.x] = p; partial[2 * kRedBlocks + blockIdx.x] = l;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  double h = 0.0, l = 0.0, p = 0.0;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += kRedThreads) {
    dd_merge(h, l, __ldcg(&partial[b]), __ldcg(&partial[2 * kRedBlocks + b]));
    p += __ldcg(&partial[kRedBlocks + b]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dd_shfl_down(h, l, o);
    p += __shfl_down_sync(0xffffffffu, p, o);
  }
  if (lane == 0) { sm[0][wid] = h; sm[1][wid] = p; sm[2][wid] = l; }
  __syncthreads();
  if (threadIdx.x == 0) {
    h = sm[0][0]; p = sm[1][0]; l = sm[2][0];
    for (int w = 1; w < kRedThreads / 32; ++w) { dd_merge(h, l, sm[0][w], sm[2][w]); p += sm[1][w]; }
    result[0] = h; result[1] = p; result[2] = l;   // mapped pinned host memory
    *ticket = 0;
    __threadfence_system();
  }
}

// sum_i RN(RN(x_i w_i)^2) [over id_i > 0]: every term rounded like nvector_parallel's loop, the sum in double-double
template <bool HAS_Z>
__global__ void __launch_bounds__(kRedThreads) red_sqw_kernel(const double *x, const double *w, const double *z, long long n,
                                                              double *partial, unsigned int *ticket, double *result) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double hi = 0.0, lo = 0.0;
  auto term = [&](double xv, double wv, double zv) {
    const double p = __dmul_rn(xv, wv);
    if (!HAS_Z || zv > 0.0) dd_add(hi, lo, __dmul_rn(p, p));
  };
  const bool aligned = (((uintptr_t)x | (uintptr_t)w | (uintptr_t)z) & 15) == 0;
  if (aligned) {
    const long long n2 = n >> 1;
    const double2 *x2 = reinterpret_cast<const double2 *>(x), *w2 = reinterpret_cast<const double2 *>(w);
    const double2 *z2 = reinterpret_cast<const double2 *>(z);
#pragma unroll 4
    for (long long i = tid; i < n2; i += stride) {
      const double2 a = x2[i], b = w2[i];
      const double2 c = HAS_Z ? z2[i] : make_double2(0, 0);
      term(a.x, b.x, c.x);
      term(a.y, b.y, c.y);
    }
    if (tid == 0 && (n & 1)) term(x[n - 1], w[n - 1], HAS_Z ? z[n - 1] : 0.0);
  } else {
    for (long long i = tid; i < n; i += stride) term(x[i], w[i], HAS_Z ? z[i] : 0.0);
  }
  block_finish_dd<false>(hi, lo, 0.0, partial, ticket, result);
}

// M: functor  double operator()(double x, double y, double z)
template <int OP, class M, bool HAS_Y, bool HAS_Z>
__global__ void __launch_bounds__(kRedThreads) red_kernel(M m, const double *x, const double *y, const double *z, long long n,
                                                          double *partial, unsigned int *ticket, double *result) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double acc[1] = {ident<OP>()};
  const bool aligned = (((uintptr_t)x | (uintptr_t)y | (uintptr_t)z) & 15) == 0;
  if (aligned) {
    const long long n2 = n >> 1;
    const double2 *x2 = reinterpret_cast<const double2 *>(x);
    const double2 *y2 = reinterpret_cast<const double2 *>(y);
    const double2 *z2 = reinterpret_cast<const double2 *>(z);
#pragma unroll 4
    for (long long i = tid; i < n2; i += stride) {
      double2 a = x2[i];
      double2 b = HAS_Y ? y2[i] : make_double2(0, 0);
      double2 c = HAS_Z ? z2[i] : make_double2(0, 0);
      acc[0] = comb<OP>(acc[0], m(a.x, b.x, c.x));
      acc[0] = comb<OP>(acc[0], m(a.y, b.y, c.y));
    }
    if (tid == 0 && (n & 1)) acc[0] = comb<OP>(acc[0], m(x[n - 1], HAS_Y ? y[n - 1] : 0.0, HAS_Z ? z[n - 1] : 0.0));
  } else {
    for (long long i = tid; i < n; i += stride) acc[0] = comb<OP>(acc[0], m(x[i], HAS_Y ? y[i] : 0.0, HAS_Z ? z[i] : 0.0));
  }
  block_finish<OP, 1>(acc, partial, ticket, result);
}

struct MDot { __device__ double operator()(double x, double y, double) const { return x * y; } };
struct MAbs { __device__ double operator()(double x, double, double) const { return fabs(x); } };
struct MId { __device__ double operator()(double x, double, double) const { return x; } };
struct MQuot { __device__ double operator()(double n, double d, double) const { return d == 0.0 ? DBL_MAX : __ddiv_rn(n, d); } };

// local reduction -> host value -> cross-rank allreduce hook
template <int OP, class M, bool HAS_Y, bool HAS_Z>
double reduce(crd_ctx *c, M m, const double *x, const double *y, const double *z, long long n, const char *name) {
  double v = OP == OP_SUM ? 0.0 : (OP == OP_MAX ? -DBL_MAX : DBL_MAX);
  if (use(c)) return NAN;
  if (n > 0) {
    unsigned int blocks = grid_for((n + 1) / 2, c->sms);
    if (blocks > (unsigned)kRedBlocks) blocks = kRedBlocks;
    red_kernel<OP, M, HAS_Y, HAS_Z><<<blocks, kRedThreads, 0, c->stream>>>(m, x, y, z, n, c->red_partial, c->red_ticket, red_target(c));
    if (check_launch(c, name)) return NAN;
    if (launch_comm_exchange(c, OP == OP_SUM ? COMM_SUM_DD : (OP == OP_MAX ? COMM_MAX : COMM_MIN), 1)) return NAN;
    if (sync_stream(c, name)) return NAN;
    v = c->red_result_host[0];
  }
  if (c->nranks > 1 && !c->dev_comm) {
    const int op = OP == OP_SUM ? CRD_SUM : (OP == OP_MAX ? CRD_MAX : CRD_MIN);
    if (c->allreduce(&v, 1, op, c->allreduce_user) != 0) { set_error("%s: allreduce hook failed", name); return NAN; }
  }
  return v;
}

// sum of weighted squares, order-independent: local double-double -> ranks' pairs merged in rank order -> one rounding
template <bool HAS_Z>
double reduce_sqw(crd_ctx *c, const double *x, const double *w, const double *z, long long n, const char *name) {
  double hi = 0.0, lo = 0.0;
  if (use(c)) return NAN;
  if (n > 0) {
    unsigned int blocks = grid_for((n + 1) / 2, c->sms);
    if (blocks > (unsigned)kRedBlocks) blocks = kRedBlocks;
    red_sqw_kernel<HAS_Z><<<blocks, kRedThreads, 0, c->stream>>>(x, w, z, n, c->red_partial, c->red_ticket, red_target(c));
    if (check_launch(c, name)) return NAN;
    if (launch_comm_exchange(c, COMM_SUM_DD, 3)) return NAN;
    if (sync_stream(c, name)) return NAN;
    hi = c->red_result_host[0]; lo = c->red_result_host[2];
  }
  if (allreduce_dd(c, hi, lo, nullptr)) return NAN;
  return hi + lo;
}

// ---- flag-producing element-wise ops (invtest, constrmask): write z and reduce a MIN flag ---------------
template <class F>
__global__ void __launch_bounds__(kRedThreads) flag_kernel(F f, const double *x, const double *y, double *z, long long n,
                                                           double *partial, unsigned int *ticket, double *result) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  double acc[1] = {1.0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) acc[0] = fmin(acc[0], f(x, y, z, i));
  block_finish<OP_MIN, 1>(acc, partial, ticket, result);
}
struct FInvTest {
  __device__ double operator()(const double *x, const double *, double *z, long long i) const {
    const double v = x[i];
    if (v == 0.0) return 0.0;
    z[i] = __ddiv_rn(1.0, v);
    return 1.0;
  }
};
struct FConstrMask {  // c = x-arg, x = y-arg, m = z
  __device__ double operator()(const double *c, const double *x, double *m, long long i) const {
    const double cv = c[i], xv = x[i];
    double flag = 1.0, mv = 0.0;
    if (cv != 0.0) {
      if (cv > 1.5 || cv < -1.5) { if (xv * cv <= 0.0) { flag = 0.0; mv = 1.0; } }
      else if (cv > 0.5 || cv < -0.5) { if (xv * cv < 0.0) { flag = 0.0; mv = 1.0; } }
    }
    m[i] = mv;
    return flag;
  }
};
template <class F>
double flag_op(crd_ctx *c, F f, const double *x, const double *y, double *z, long long n, const char *name) {
  double v = 1.0;
  if (use(c)) return NAN;
  if (n > 0) {
    unsigned int blocks = grid_for(n);
    if (blocks > (unsigned)kRedBlocks) blocks = kRedBlocks;
    flag_kernel<F><<<blocks, kRedThreads, 0, c->stream>>>(f, x, y, z, n, c->red_partial, c->red_ticket, red_target(c));
    if (check_launch(c, name)) return NAN;
    if (launch_comm_exchange(c, COMM_MIN, 1)) return NAN;
    if (sync_stream(c, name)) return NAN;
    v = c->red_result_host[0];
  }
  if (c->nranks > 1 && !c->dev_comm && c->allreduce(&v, 1, CRD_MIN, c->allreduce_user) != 0) { set_error("%s: allreduce hook failed", name); return NAN; }
  return v;
}

// ---- fused: z = sum_j c_j X_j ----------------------------------------------------------------------------
struct LinCombArgs { const double *x[CRD_ARK_MAX_LINCOMB]; double c[CRD_ARK_MAX_LINCOMB]; };

template <int N>
__global__ void __launch_bounds__(256) lincomb_kernel(const LinCombArgs a, double *z, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n2 = n >> 1;
  double2 *z2 = reinterpret_cast<double2 *>(z);
#pragma unroll 2
  for (long long i = tid; i < n2; i += stride) {
    double2 v[N];
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = reinterpret_cast<const double2 *>(a.x[j])[i];
    double2 s = make_double2(a.c[0] * v[0].x, a.c[0] * v[0].y);
#pragma unroll
    for (int j = 1; j < N; ++j) { s.x = fma(a.c[j], v[j].x, s.x); s.y = fma(a.c[j], v[j].y, s.y); }
    z2[i] = s;
  }
  if (tid == 0 && (n & 1)) {
    double s = a.c[0] * a.x[0][n - 1];
    for (int j = 1; j < N; ++j) s = fma(a.c[j], a.x[j][n - 1], s);
    z[n - 1] = s;
  }
}

// ---- fused: ynew = yn + sum hb_j F_j ; err = sum hd_j F_j ; two weighted square sums (crd_fused.cuh) ----------
// SEQ: the bits of the op-by-op sequence (separately rounded chains, IEEE weights, double-double error sum)
template <int S, bool SEQ>
__global__ void __launch_bounds__(kRedThreads) erk_finish_kernel(const FinishArgs a, long long n, double *partial,
                                                                 unsigned int *ticket, double *result) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n2 = n >> 1;
  FinAcc<SEQ> acc;
  double2 *o2 = reinterpret_cast<double2 *>(a.ynew);
  for (long long i = tid; i < n2; i += stride) {
    double2 f2[S];
#pragma unroll
    for (int j = 0; j < S; ++j) f2[j] = reinterpret_cast<const double2 *>(a.F[j])[i];
    const double2 y0 = reinterpret_cast<const double2 *>(a.yn)[i];
    double fx[S], fy[S];
#pragma unroll
    for (int j = 0; j < S; ++j) { fx[j] = f2[j].x; fy[j] = f2[j].y; }
    double2 o;
    finish_elem<SEQ, S>(a, y0.x, fx, o.x, acc);
    finish_elem<SEQ, S>(a, y0.y, fy, o.y, acc);
    o2[i] = o;
  }
  if (tid == 0 && (n & 1)) {
    double f[S];
    for (int j = 0; j < S; ++j) f[j] = a.F[j][n - 1];
    double o;
    finish_elem<SEQ, S>(a, a.yn[n - 1], f, o, acc);
    a.ynew[n - 1] = o;
  }
  // the second sum, or (skipped: crd_fused.cuh, finish_y2_bound) its bound, contributed once
  const double y2 = a.y2_bound < 0.0 ? acc.y2 : ((blockIdx.x == 0 && threadIdx.x == 0) ? a.y2_bound : 0.0);
  block_finish_dd<true>(acc.e_hi, acc.e_lo, y2, partial, ticket, result);
}

struct _generic_N_Vector_Ops g_ops = {
#ifdef CRD_SUNDIALS_27
    nullptr,
#endif
    N_VClone_Crd, N_VCloneEmpty_Crd, N_VDestroy_Crd, N_VSpace_Crd, N_VGetArrayPointer_Crd, N_VSetArrayPointer_Crd,
    N_VLinearSum_Crd, N_VConst_Crd, N_VProd_Crd, N_VDiv_Crd, N_VScale_Crd, N_VAbs_Crd, N_VInv_Crd, N_VAddConst_Crd,
    N_VDotProd_Crd, N_VMaxNorm_Crd, N_VWrmsNorm_Crd, N_VWrmsNormMask_Crd, N_VMin_Crd, N_VWL2Norm_Crd, N_VL1Norm_Crd,
    N_VCompare_Crd, N_VInvTest_Crd, N_VConstrMask_Crd, N_VMinQuotient_Crd};

// FAST grids: chains of fused multiply-adds.  EXACT grids: every fused entry reproduces the bits of the op-by-op sequence
// (crd_fused.cuh); there is no fused lincomb in that table, so the rare combinations outside a step (dense output) are issued
// op by op as well.
const crd_fused_ops g_fused = {N_VLinearCombination_Crd, N_VErkFinish_Crd, crd_f_lincomb, crd_erk_evolve, crd_f_lincomb_finish, crd_f_pair};
const crd_fused_ops g_fused_ops_only = {N_VLinearCombination_Crd, N_VErkFinish_Crd, nullptr, nullptr, nullptr, nullptr};
const crd_fused_ops g_fused_exact = {nullptr, N_VErkFinishSeq_Crd, crd_f_lincomb, crd_erk_evolve, crd_f_lincomb_finish, crd_f_pair};
const crd_fused_ops g_fused_exact_ops_only = {nullptr, N_VErkFinishSeq_Crd, nullptr, nullptr, nullptr, nullptr};

}  // namespace

extern "C" {

N_Vector N_VNewEmpty_Crd(crd_ctx *ctx, long int local_length, long int global_length) {
  if (!ctx || local_length < 0 || global_length < local_length) { set_error("N_VNewEmpty_Crd: bad arguments"); return nullptr; }
  N_Vector v = new (std::nothrow) _generic_N_Vector;
  crd_nv_content *c = new (std::nothrow) crd_nv_content;
  if (!v || !c) { delete v; delete c; set_error("N_VNewEmpty_Crd: out of host memory"); return nullptr; }
  c->local_length = local_length; c->global_length = global_length;
  c->own_data = FALSE; c->data = nullptr; c->host = nullptr; c->ctx = ctx;
  v->content = c;
  v->ops = &g_ops;
  return v;
}

N_Vector N_VNew_Crd(crd_ctx *ctx, long int local_length, long int global_length) {
  N_Vector v = N_VNewEmpty_Crd(ctx, local_length, global_length);
  if (!v) return nullptr;
  if (local_length > 0) {
    void *p = crd_malloc(ctx, sizeof(double) * (size_t)local_length);
    if (!p) { N_VDestroy_Crd(v); return nullptr; }
    NVC(v)->data = (double *)p;
    NVC(v)->own_data = TRUE;
  }
  return v;
}

N_Vector N_VMake_Crd(crd_ctx *ctx, long int local_length, long int global_length, realtype *dev_data) {
  N_Vector v = N_VNewEmpty_Crd(ctx, local_length, global_length);
  if (!v) return nullptr;
  NVC(v)->data = dev_data;
  return v;
}

void N_VDestroy_Crd(N_Vector v) {
  if (!v) return;
  crd_nv_content *c = NVC(v);
  if (c) {
    if (c->own_data && c->data) crd_free(c->ctx, c->data);
    if (c->host) cudaFreeHost(c->host);
    delete c;
  }
  delete v;
}

realtype *N_VGetDeviceArrayPointer_Crd(N_Vector v) { return v ? D(v) : nullptr; }
long int N_VGetLocalLength_Crd(N_Vector v) { return v ? NVC(v)->local_length : 0; }
crd_ctx *N_VGetContext_Crd(N_Vector v) { return v ? CTX(v) : nullptr; }

realtype *N_VGetArrayPointer_Crd(N_Vector v) {
  crd_nv_content *c = NVC(v);
  if (!c->host && c->local_length > 0) {
    if (cudaHostAlloc(&c->host, sizeof(double) * (size_t)c->local_length, cudaHostAllocDefault) != cudaSuccess) {
      set_error("N_VGetArrayPointer_Crd: cannot allocate the pinned host mirror");
      cudaGetLastError();
      return nullptr;
    }
  }
  return c->host;
}
void N_VSetArrayPointer_Crd(realtype *dev_data, N_Vector v) {
  crd_nv_content *c = NVC(v);
  if (c->own_data && c->data) crd_free(c->ctx, c->data);
  c->data = dev_data; c->own_data = FALSE;
}
int N_VCopyToHost_Crd(N_Vector v) {
  realtype *h = N_VGetArrayPointer_Crd(v);
  if (!h) return LEN(v) == 0 ? 0 : -1;
  return crd_memcpy_d2h(CTX(v), h, D(v), sizeof(double) * (size_t)LEN(v));
}
int N_VCopyFromHost_Crd(N_Vector v) {
  realtype *h = N_VGetArrayPointer_Crd(v);
  if (!h) return LEN(v) == 0 ? 0 : -1;
  return crd_memcpy_h2d(CTX(v), D(v), h, sizeof(double) * (size_t)LEN(v));
}

N_Vector N_VCloneEmpty_Crd(N_Vector w) { return N_VNewEmpty_Crd(CTX(w), NVC(w)->local_length, NVC(w)->global_length); }
N_Vector N_VClone_Crd(N_Vector w) { return N_VNew_Crd(CTX(w), NVC(w)->local_length, NVC(w)->global_length); }
void N_VSpace_Crd(N_Vector v, long int *lrw, long int *liw) { *lrw = NVC(v)->global_length; *liw = 2 * CTX(v)->nranks; }

void N_VLinearSum_Crd(realtype a, N_Vector x, realtype b, N_Vector y, N_Vector z) {
  crd_ctx *c = CTX(z);
  const long long n = LEN(z);
  double *xd = D(x), *yd = D(y), *zd = D(z);
  // the case analysis of nvector_parallel.c, so each element is rounded like the serial loops
  if (b == 1.0 && z == y) { ew<FLin1, true, true>(c, FLin1{a}, xd, yd, yd, n, "nv_axpy"); return; }           // y += a x
  if (a == 1.0 && z == x) { ew<FLin1, true, true>(c, FLin1{b}, yd, xd, xd, n, "nv_axpy"); return; }           // x += b y
  if (a == 1.0 && b == 1.0) { ew<FSum, true, true>(c, FSum{}, xd, yd, zd, n, "nv_sum"); return; }
  if (a == 1.0 && b == -1.0) { ew<FDiff, true, true>(c, FDiff{}, xd, yd, zd, n, "nv_diff"); return; }         // x - y
  if (a == -1.0 && b == 1.0) { ew<FDiff, true, true>(c, FDiff{}, yd, xd, zd, n, "nv_diff"); return; }         // y - x
  if (a == 1.0) { ew<FLin1, true, true>(c, FLin1{b}, yd, xd, zd, n, "nv_lin1"); return; }                     // b y + x
  if (b == 1.0) { ew<FLin1, true, true>(c, FLin1{a}, xd, yd, zd, n, "nv_lin1"); return; }                     // a x + y
  if (a == -1.0) { ew<FLin2, true, true>(c, FLin2{b}, yd, xd, zd, n, "nv_lin2"); return; }                    // b y - x
  if (b == -1.0) { ew<FLin2, true, true>(c, FLin2{a}, xd, yd, zd, n, "nv_lin2"); return; }                    // a x - y
  if (a == b) { ew<FScaleSum, true, true>(c, FScaleSum{a}, xd, yd, zd, n, "nv_scalesum"); return; }
  if (a == -b) { ew<FScaleDiff, true, true>(c, FScaleDiff{a}, xd, yd, zd, n, "nv_scalediff"); return; }
  ew<FLinSum, true, true>(c, FLinSum{a, b}, xd, yd, zd, n, "nv_linearsum");
}
void N_VConst_Crd(realtype cv, N_Vector z) { ew<FConst, false, false>(CTX(z), FConst{cv}, nullptr, nullptr, D(z), LEN(z), "nv_const"); }
void N_VProd_Crd(N_Vector x, N_Vector y, N_Vector z) { ew<FProd, true, true>(CTX(z), FProd{}, D(x), D(y), D(z), LEN(z), "nv_prod"); }
void N_VDiv_Crd(N_Vector x, N_Vector y, N_Vector z) { ew<FDiv, true, true>(CTX(z), FDiv{}, D(x), D(y), D(z), LEN(z), "nv_div"); }
void N_VScale_Crd(realtype cv, N_Vector x, N_Vector z) {
  crd_ctx *c = CTX(z);
  const long long n = LEN(z);
  if (z == x) { ew<FScale, true, false>(c, FScale{cv}, D(x), nullptr, D(x), n, "nv_scale"); return; }
  if (cv == 1.0) { ew<FCopy, true, false>(c, FCopy{}, D(x), nullptr, D(z), n, "nv_copy"); return; }
  if (cv == -1.0) { ew<FNeg, true, false>(c, FNeg{}, D(x), nullptr, D(z), n, "nv_neg"); return; }
  ew<FScale, true, false>(c, FScale{cv}, D(x), nullptr, D(z), n, "nv_scale");
}
void N_VAbs_Crd(N_Vector x, N_Vector z) { ew<FAbs, true, false>(CTX(z), FAbs{}, D(x), nullptr, D(z), LEN(z), "nv_abs"); }
void N_VInv_Crd(N_Vector x, N_Vector z) { ew<FInv, true, false>(CTX(z), FInv{}, D(x), nullptr, D(z), LEN(z), "nv_inv"); }
void N_VAddConst_Crd(N_Vector x, realtype b, N_Vector z) { ew<FAddC, true, false>(CTX(z), FAddC{b}, D(x), nullptr, D(z), LEN(z), "nv_addconst"); }
void N_VCompare_Crd(realtype cv, N_Vector x, N_Vector z) { ew<FCmp, true, false>(CTX(z), FCmp{cv}, D(x), nullptr, D(z), LEN(z), "nv_compare"); }

realtype N_VDotProd_Crd(N_Vector x, N_Vector y) { return reduce<OP_SUM, MDot, true, false>(CTX(x), MDot{}, D(x), D(y), nullptr, LEN(x), "nv_dotprod"); }
realtype N_VMaxNorm_Crd(N_Vector x) {
  double v = reduce<OP_MAX, MAbs, false, false>(CTX(x), MAbs{}, D(x), nullptr, nullptr, LEN(x), "nv_maxnorm");
  return v < 0.0 ? 0.0 : v;
}
realtype N_VWrmsNorm_Crd(N_Vector x, N_Vector w) {
  double s = reduce_sqw<false>(CTX(x), D(x), D(w), nullptr, LEN(x), "nv_wrmsnorm");
  return std::sqrt(s / (double)NVC(x)->global_length);
}
realtype N_VWrmsNormMask_Crd(N_Vector x, N_Vector w, N_Vector id) {
  double s = reduce_sqw<true>(CTX(x), D(x), D(w), D(id), LEN(x), "nv_wrmsnormmask");
  return std::sqrt(s / (double)NVC(x)->global_length);
}
realtype N_VMin_Crd(N_Vector x) { return reduce<OP_MIN, MId, false, false>(CTX(x), MId{}, D(x), nullptr, nullptr, LEN(x), "nv_min"); }
realtype N_VWL2Norm_Crd(N_Vector x, N_Vector w) {
  return std::sqrt(reduce_sqw<false>(CTX(x), D(x), D(w), nullptr, LEN(x), "nv_wl2norm"));
}
realtype N_VL1Norm_Crd(N_Vector x) { return reduce<OP_SUM, MAbs, false, false>(CTX(x), MAbs{}, D(x), nullptr, nullptr, LEN(x), "nv_l1norm"); }
realtype N_VMinQuotient_Crd(N_Vector num, N_Vector denom) {
  return reduce<OP_MIN, MQuot, true, false>(CTX(num), MQuot{}, D(num), D(denom), nullptr, LEN(num), "nv_minquotient");
}
booleantype N_VInvTest_Crd(N_Vector x, N_Vector z) {
  return flag_op(CTX(x), FInvTest{}, D(x), nullptr, D(z), LEN(x), "nv_invtest") == 0.0 ? FALSE : TRUE;
}
booleantype N_VConstrMask_Crd(N_Vector cvec, N_Vector x, N_Vector m) {
  return flag_op(CTX(x), FConstrMask{}, D(cvec), D(x), D(m), LEN(x), "nv_constrmask") == 1.0 ? TRUE : FALSE;
}

int N_VLinearCombination_Crd(int n, const realtype *cf, N_Vector *X, N_Vector z) {
  if (n < 1 || n > CRD_ARK_MAX_LINCOMB || !cf || !X || !z) { set_error("N_VLinearCombination_Crd: bad arguments"); return -1; }
  crd_ctx *c = CTX(z);
  const long long len = LEN(z);
  if (len <= 0) return 0;
  if (use(c)) return -1;
  LinCombArgs a;
  for (int j = 0; j < n; ++j) {
    if (LEN(X[j]) != len) { set_error("N_VLinearCombination_Crd: length mismatch"); return -1; }
    a.x[j] = D(X[j]); a.c[j] = cf[j];
    if ((uintptr_t)a.x[j] & 15) { set_error("N_VLinearCombination_Crd: vectors must be 16-byte aligned"); return -1; }
  }
  if ((uintptr_t)D(z) & 15) { set_error("N_VLinearCombination_Crd: vectors must be 16-byte aligned"); return -1; }
  const unsigned int blocks = grid_for((len + 1) / 2, c->sms);
  double *zd = D(z);
  switch (n) {
    case 1: lincomb_kernel<1><<<blocks, 256, 0, c->stream>>>(a, zd, len); break;
    case 2: lincomb_kernel<2><<<blocks, 256, 0, c->stream>>>(a, zd, len); break;
    case 3: lincomb_kernel<3><<<blocks, 256, 0, c->stream>>>(a, zd, len); break;
    case 4: lincomb_kernel<4><<<blocks, 256, 0, c->stream>>>(a, zd, len); break;
    case 5: lincomb_kernel<5><<<blocks, 256, 0, c->stream>>>(a, zd, len); break;
    case 6: lincomb_kernel<6><<<blocks, 256, 0, c->stream>>>(a, zd, len); break;
    case 7: lincomb_kernel<7><<<blocks, 256, 0, c->stream>>>(a, zd, len); break;
    default: lincomb_kernel<8><<<blocks, 256, 0, c->stream>>>(a, zd, len); break;
  }
  return check_launch(c, "lincomb_kernel");
}

static int erk_finish_impl(bool seq, int s, const realtype *hb, const realtype *hd, N_Vector yn, N_Vector *F, N_Vector ynew,
                           realtype rtol, realtype atol, realtype out[2]) {
  if (s < 1 || s > CRD_ARK_MAX_LINCOMB || !hb || !hd || !yn || !F || !ynew || !out) { set_error("N_VErkFinish_Crd: bad arguments"); return -1; }
  crd_ctx *c = CTX(yn);
  const long long len = LEN(yn);
  if (use(c)) return -1;
  double hi = 0.0, lo = 0.0, y2 = 0.0;
  if (len > 0) {
    FinishArgs a;
    for (int j = 0; j < s; ++j) {
      if (LEN(F[j]) != len) { set_error("N_VErkFinish_Crd: length mismatch"); return -1; }
      a.F[j] = D(F[j]); a.hb[j] = hb[j]; a.hd[j] = hd[j];
      if ((uintptr_t)a.F[j] & 15) { set_error("N_VErkFinish_Crd: vectors must be 16-byte aligned"); return -1; }
    }
    a.yn = D(yn); a.ynew = D(ynew); a.rtol = rtol; a.atol = atol;
    a.hb_nz = finish_nz_mask(hb, s); a.y2_bound = finish_y2_bound(rtol, len);
    if (((uintptr_t)a.yn | (uintptr_t)a.ynew) & 15) { set_error("N_VErkFinish_Crd: vectors must be 16-byte aligned"); return -1; }
    unsigned int blocks = grid_for((len + 1) / 2, c->sms);
    if (blocks > (unsigned)kRedBlocks) blocks = kRedBlocks;
#define CRD_FIN_CASE(S_)                                                                                                     \
  case S_:                                                                                                                   \
    if (seq) erk_finish_kernel<S_, true><<<blocks, kRedThreads, 0, c->stream>>>(a, len, c->red_partial, c->red_ticket, red_target(c)); \
    else erk_finish_kernel<S_, false><<<blocks, kRedThreads, 0, c->stream>>>(a, len, c->red_partial, c->red_ticket, red_target(c));    \
    break;
    switch (s) {
      CRD_FIN_CASE(1) CRD_FIN_CASE(2) CRD_FIN_CASE(3) CRD_FIN_CASE(4) CRD_FIN_CASE(5) CRD_FIN_CASE(6) CRD_FIN_CASE(7)
      default:
        if (seq) erk_finish_kernel<8, true><<<blocks, kRedThreads, 0, c->stream>>>(a, len, c->red_partial, c->red_ticket, red_target(c));
        else erk_finish_kernel<8, false><<<blocks, kRedThreads, 0, c->stream>>>(a, len, c->red_partial, c->red_ticket, red_target(c));
        break;
    }
#undef CRD_FIN_CASE
    if (check_launch(c, "erk_finish_kernel")) return -1;
    if (launch_comm_exchange(c, COMM_SUM_DD, 3)) return -1;
    if (sync_stream(c, "N_VErkFinish_Crd")) return -1;
    hi = c->red_result_host[0]; y2 = c->red_result_host[1]; lo = c->red_result_host[2];
  }
  if (allreduce_dd(c, hi, lo, &y2)) return -1;
  out[0] = hi + lo;
  out[1] = y2;
  return 0;
}

int N_VErkFinish_Crd(int s, const realtype *hb, const realtype *hd, N_Vector yn, N_Vector *F, N_Vector ynew,
                     realtype rtol, realtype atol, realtype out[2]) {
  return erk_finish_impl(false, s, hb, hd, yn, F, ynew, rtol, atol, out);
}
int N_VErkFinishSeq_Crd(int s, const realtype *hb, const realtype *hd, N_Vector yn, N_Vector *F, N_Vector ynew,
                        realtype rtol, realtype atol, realtype out[2]) {
  return erk_finish_impl(true, s, hb, hd, yn, F, ynew, rtol, atol, out);
}

const crd_fused_ops *crd_nv_fused_ops(void) { return &g_fused; }
const crd_fused_ops *crd_nv_fused_vector_ops(void) { return &g_fused_ops_only; }
const crd_fused_ops *crd_nv_fused_ops_exact(void) { return &g_fused_exact; }
const crd_fused_ops *crd_nv_fused_vector_ops_exact(void) { return &g_fused_exact_ops_only; }
const crd_fused_ops *crd_nv_fused_ops_for(const crd_grid *g) {
  crd_params p;
  if (!g || crd_grid_params(g, &p) != 0) return nullptr;
  return p.arith == CRD_ARITH_EXACT ? &g_fused_exact : &g_fused;
}

}  // extern "C"

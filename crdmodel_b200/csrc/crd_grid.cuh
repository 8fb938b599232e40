// crd_grid.cuh — the grid descriptor (device-side equivalent of the reference's UserData,
// src/FHNmodel_torus.cpp:97-122) and the kernel argument block of the fused RHS.
#pragma once
#include "crd_common.cuh"

namespace crd {

// Scalars the RHS kernels need; filled once per grid on the host with the reference's own expressions
// so that the EXACT kernels reproduce its rounding.
struct RhsConst {
  double Diff;
  // torus, exact: T = Diff*(a*(...))/den
  double inv_rr;   // 1/(r*r)
  double twodx;    // 2*dx
  double dxdx;     // dx*dx
  double dydy;     // dy*dy
  double r_twodx, r_dxdx, r_dydy;  // RN(1/x) of the three divisors (host IEEE division)
  int dv_plus0;    // some beta(phi) is -0.0: keep the reference's 0.0 + eps*(u+b)
  int div_safe;    // divisors positive and within 2^+-90: the reciprocal-refinement division applies
  // torus, fast: folded
  double c2;       // Diff*inv_rr/dxdx
  // flat (FHNmodel_flat.cpp:489-491)
  double cu1, cu2, cu3;
  // Goldbeter constants evaluated by the host libm (GoldbeterModel_torus.cpp:694-695)
  double k2n;      // pow(K2, n) = 1
  double krm;      // pow(KR, m) = 4
  double kap;      // pow(KA, p) = 0.6561...
};

constexpr int kMaxLc = 5;   // yn + up to four stage derivatives (Zonneveld stage 5)

// The state an evaluation reads: a vector y, or (fused stage assembly) the combination sum_j c[j]*x[j]
// that is never written to memory.
struct StateRef {
  const double *y = nullptr;
  int n = 0;
  const double *x[kMaxLc] = {};
  double c[kMaxLc] = {};
};

// Halo exchange folded into an evaluation's own launch (phi-split grids).  The boundary rows travel by plain stores through
// peer-mapped memory; ordering is per STRIP of kHaloStrip theta columns: the CTA that has written a strip of a neighbour's
// ghost row release-stores the evaluation's epoch into that strip's flag over there, and whoever needs a ghost-row strip
// acquires its own flag first.  A launch may push (its first CTAs do, before anything else), wait (the tiles / row segments
// that touch a ghost row do, and they are scheduled last), both, or neither.
constexpr int kHaloStrip = 256;
struct HaloSync {
  double *push_prev = nullptr;                    // prev rank's north ghost row of this epoch: receives this slab's row 0
  double *push_next = nullptr;                    // next rank's south ghost row: receives this slab's last row
  double *push2_prev = nullptr, *push2_next = nullptr;   // their second ghost rows: this slab's row 1 / row nyl-2 (pair pass only)
  unsigned long long *flag_prev = nullptr;        // their per-strip flags
  unsigned long long *flag_next = nullptr;
  const unsigned long long *wait_south = nullptr; // this grid's per-strip flags: acquire before the south / north ghost row is read
  const unsigned long long *wait_north = nullptr;
  unsigned long long epoch = 0;
  long long timeout_ns = 0;
  int *err = nullptr;                             // mapped host word: a wait gave up (the launch still completes, with stale rows)
  int edge_last = 0;                              // schedule the work that touches ghost rows last
};

struct RhsArgs {
  const double *y;
  double *ydot;
  // fused stage assembly: nlc >= 1 => state = sum_j lc_c[j]*lc_x[j] (pointers at the launch's first row); y unused
  int nlc;
  const double *lc_x[kMaxLc];
  double lc_c[kMaxLc];
  long long south_off, north_off;  // nlc >= 1 and south/north == nullptr: point offset of that row inside lc_x
  const double *south;   // the row below the slab's first row (global row js-1), same [nx][2] layout
  const double *north;   // the row above the slab's last row  (global row je+1)
  const double *cth;     // [nx][2]: exact (a1, a3) | fast (c1, c3); unused for flat
  const double *brow;    // [nyl]: FHN b(phi) | Goldbeter v0 + v1*b(phi)
  long long nx, nyl;
  int freeze_south;      // t < tBoundary && js == 0       (:649)
  int freeze_north;      // t < tBoundary && je == ny-1    (:643)
  int react;             // 0 only for Goldbeter with justDiffusion (:668)
  int div_shift;         // >= 0: work item / nx by multiply-shift (set at launch)
  unsigned div_magic;
  RhsConst k;
  HaloSync hs;
};

// ghost block, one per grid, cudaMalloc'd so it can be exported with cudaIpcGetMemHandle:
//   double ghost[2 parity][2 side][nx][2]        side 0 = south (row js-1), side 1 = north (row je+1); whole (u,v) rows
//   unsigned long long flag[2 side][nstrips]     per strip of kHaloStrip columns: epoch of the last complete push
struct HaloLayout {
  long long nx;
  __host__ __device__ long long nstrips() const { return (nx + kHaloStrip - 1) / kHaloStrip; }
  __host__ __device__ size_t ghost_off(int parity, int side) const { return (size_t)(parity * 2 + side) * (size_t)nx * 2 * sizeof(double); }
  // the second row beyond the slab on that side (only the pass that forms two evaluations at once exchanges it)
  __host__ __device__ size_t ghost2_off(int parity, int side) const { return (size_t)(4 + parity * 2 + side) * (size_t)nx * 2 * sizeof(double); }
  __host__ __device__ size_t flag_off(int side) const {
    return (size_t)8 * (size_t)nx * 2 * sizeof(double) + 128 + (size_t)side * (((size_t)nstrips() * 8 + 127) / 128 * 128);
  }
  __host__ __device__ size_t bytes() const { return flag_off(2); }
};

// Rows per segment of the streaming kernel.  Its units (256-column strips x row segments) go to the persistent CTAs round-robin,
// so a launch lasts ceil(units / ctas) unit times: pick the segment length that fills whole waves — a unit costs its rows plus
// the two rows above and below it that are fetched on top (and about one row of start-up).  Keeps every CTA busy on meshes of a
// few million points (1024 x 4096 with fixed 128-row segments: 128 units for 444 CTAs) and trims the last partial wave of large ones.
inline int stream_seg_rows(long long nyl, long long strips, long long ctas) {
  int best = 0;
  long long best_cost = 0;
  for (long long w = 1; w <= 256; ++w) {
    long long nseg = ctas * w / strips;
    if (nseg < 1) continue;
    if (nseg > nyl) nseg = nyl;
    const long long rows = (nyl + nseg - 1) / nseg;
    if (rows < 16) break;            // more waves only shorten the segments further
    if (rows > 512) continue;
    nseg = (nyl + rows - 1) / rows;
    const long long waves = (strips * nseg + ctas - 1) / ctas;
    const long long cost = waves * (rows + 3);
    if (best == 0 || cost < best_cost) { best = (int)rows; best_cost = cost; }
  }
  if (best == 0) best = nyl < 16 ? (int)(nyl > 0 ? nyl : 1) : 16;
  return best;
}


}  // namespace crd

struct crd_grid {
  crd_ctx *ctx = nullptr;
  crd_params p{};
  long long nx = 0, ny = 0, js = 0, je = 0, nyl = 0;
  double dx = 0, dy = 0, R = 0, r = 0, xmin = 0, xmax = 0, ymin = 0, ymax = 0;
  crd::RhsConst k{};
  double *cth = nullptr;
  double *brow = nullptr;          // per-phi beta of local row 0 .. nyl-1 (brow[-1], brow[nyl]: the rows of the neighbouring ranks)
  double *brow_alloc = nullptr;    // the allocation: brow - 1
  // halo ring
  char *halo_local = nullptr;
  char *halo_prev = nullptr, *halo_next = nullptr;  // peer-mapped (or local) ghost blocks of the neighbours
  bool prev_ipc = false, next_ipc = false;
  bool connected = false;
  unsigned long long epoch = 0;      // epoch of the last post
  unsigned long long computed = 0;   // epoch of the last compute
  int64_t rhs_count = 0;
  int variant = 0;
  double *fin_partial = nullptr;   // per-CTA sums of a stage fused with the step finish: 3 launches x [2][kRedBlocks]
  // device-resident step loop (crd_resident.cu): -1 off, 0 automatic (meshes that live in L2), 1 always
  int resident = 0;
  int64_t resident_launches = 0;
  int64_t res_cycles[6] = {};   // last launch, CTA 0: phase 1, interior rows, barrier wait, edge rows, rest, total
  unsigned long long *res_bar = nullptr;
  double *res_partial = nullptr;
  void *res_out_host = nullptr, *res_out_dev = nullptr;
  // true (default): the exchange rides inside the evaluation's own launch where the kernel supports it (the tiled and the
  // streaming kernels); false: always the separate push / wait launches
  bool overlap = true;
  // crd_rhs_host staging
  double *stage_y = nullptr, *stage_ydot = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t *ev_in = nullptr, *ev_k = nullptr;
  int n_chunks = 0;
};

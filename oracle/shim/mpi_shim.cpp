// TEST INFRASTRUCTURE — implementation of the in-process MPI subset declared in shim/mpi.h.
// Ranks = threads.  See the header for the matching rule.  Not a product component.
#include "mpi.h"

#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <vector>

namespace {

struct Message { std::vector<char> bytes; };

struct World {
  std::mutex mu;
  std::condition_variable cv;
  int size = 1;
  int dims[2] = {1, 1};
  int periods[2] = {1, 1};
  // per (src,dst): arrived messages keyed by send sequence number
  std::map<std::pair<int, int>, std::map<long, Message>> box;
  std::map<std::pair<int, int>, long> send_seq, recv_seq;
  // allreduce / barrier state
  int arrive = 0;
  long generation = 0;
  std::vector<std::vector<double>> contrib;
  std::vector<double> result;
} W;

thread_local int tl_rank = 0;
// band mode (single rank): rows a neighbouring rank would have sent (see crdshim_mpi_band_halo)
thread_local const double *tl_halo_s = nullptr, *tl_halo_n = nullptr;

size_t dt_size(MPI_Datatype dt) {
  switch (dt) {
    case MPI_DOUBLE: return sizeof(double);
    case MPI_FLOAT: return sizeof(float);
    case MPI_LONG_DOUBLE: return sizeof(long double);
    case MPI_LONG: return sizeof(long);
    default: return sizeof(int);
  }
}

}  // namespace

extern "C" {

void crdshim_mpi_set_world(int nranks) {
  std::lock_guard<std::mutex> lk(W.mu);
  W.size = nranks < 1 ? 1 : nranks;
  W.box.clear();
  W.send_seq.clear();
  W.recv_seq.clear();
  W.arrive = 0;
}

void crdshim_mpi_bind(int rank) { tl_rank = rank; }

// Band mode, world size 1: the calling thread runs the reference's f() on a band of phi rows [js, je] of a larger
// global mesh.  Its Exchange() posts four receives per call in the order W, E, S, N (FHNmodel_torus.cpp:825-846)
// and, alone in the world, would receive its own rows back (periodic self-exchange).  With this set, the S and N
// receives (the 3rd and 4th of each group) deliver the caller's rows js-1 and je+1 instead -- what the neighbouring
// ranks of a phi split would have sent.  NULL, NULL switches it off.
void crdshim_mpi_band_halo(const double *south_row, const double *north_row) { tl_halo_s = south_row; tl_halo_n = north_row; }

int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = tl_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int *size) { *size = W.size; return MPI_SUCCESS; }

// Balanced factorisation, dims in non-increasing order (MPI-3.1 §7.5.2); only ndims == 2 with
// both entries free (0) is needed by the reference (FHNmodel_torus.cpp:718-724).
int MPI_Dims_create(int nnodes, int ndims, int dims[]) {
  if (ndims != 2) return 1;
  if (dims[0] > 0 && dims[1] > 0) return MPI_SUCCESS;
  if (dims[0] > 0) { dims[1] = nnodes / dims[0]; return MPI_SUCCESS; }
  if (dims[1] > 0) { dims[0] = nnodes / dims[1]; return MPI_SUCCESS; }
  int b = 1;
  for (int c = 1; (long)c * c <= nnodes; ++c)
    if (nnodes % c == 0) b = c;
  dims[0] = nnodes / b;
  dims[1] = b;
  return MPI_SUCCESS;
}

int MPI_Cart_create(MPI_Comm, int ndims, const int dims[], const int periods[], int, MPI_Comm *cart) {
  if (ndims != 2) return 1;
  std::lock_guard<std::mutex> lk(W.mu);
  W.dims[0] = dims[0]; W.dims[1] = dims[1];
  W.periods[0] = periods[0]; W.periods[1] = periods[1];
  *cart = 1;
  return MPI_SUCCESS;
}

int MPI_Cart_get(MPI_Comm, int maxdims, int dims[], int periods[], int coords[]) {
  if (maxdims < 2) return 1;
  dims[0] = W.dims[0]; dims[1] = W.dims[1];
  periods[0] = W.periods[0]; periods[1] = W.periods[1];
  coords[0] = tl_rank / W.dims[1];   // row-major rank order
  coords[1] = tl_rank % W.dims[1];
  return MPI_SUCCESS;
}

int MPI_Cart_shift(MPI_Comm, int direction, int disp, int *rank_source, int *rank_dest) {
  int c[2] = {tl_rank / W.dims[1], tl_rank % W.dims[1]};
  auto at = [&](int delta) -> int {
    int cc[2] = {c[0], c[1]};
    int v = cc[direction] + delta;
    int d = W.dims[direction];
    if (W.periods[direction]) v = ((v % d) + d) % d;
    else if (v < 0 || v >= d) return -2;  // MPI_PROC_NULL
    cc[direction] = v;
    return cc[0] * W.dims[1] + cc[1];
  };
  *rank_source = at(-disp);
  *rank_dest = at(+disp);
  return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int, MPI_Comm comm, MPI_Request *req) {
  std::lock_guard<std::mutex> lk(W.mu);
  req->kind = 1; req->peer = source; req->buf = buf; req->count = (int)(count * dt_size(dt));
  req->comm = comm;
  req->seq = W.recv_seq[{source, tl_rank}]++;
  return MPI_SUCCESS;
}

int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int, MPI_Comm comm, MPI_Request *req) {
  size_t nbytes = count * dt_size(dt);
  {
    std::lock_guard<std::mutex> lk(W.mu);
    long s = W.send_seq[{tl_rank, dest}]++;
    Message &m = W.box[{tl_rank, dest}][s];
    m.bytes.assign((const char *)buf, (const char *)buf + nbytes);
    req->kind = 2; req->peer = dest; req->buf = nullptr; req->count = (int)nbytes; req->seq = s;
    req->comm = comm;
  }
  W.cv.notify_all();
  return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *req, MPI_Status *) {
  if (req->kind != 1) return MPI_SUCCESS;  // sends are buffered
  std::unique_lock<std::mutex> lk(W.mu);
  auto key = std::make_pair(req->peer, tl_rank);
  W.cv.wait(lk, [&] {
    auto it = W.box.find(key);
    return it != W.box.end() && it->second.count(req->seq) != 0;
  });
  auto &q = W.box[key];
  Message &m = q[req->seq];
  size_t n = m.bytes.size() < (size_t)req->count ? m.bytes.size() : (size_t)req->count;
  std::memcpy(req->buf, m.bytes.data(), n);
  q.erase(req->seq);
  if (tl_halo_s && req->seq % 4 == 2) std::memcpy(req->buf, tl_halo_s, (size_t)req->count);
  if (tl_halo_n && req->seq % 4 == 3) std::memcpy(req->buf, tl_halo_n, (size_t)req->count);
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm) {
  if (dt != MPI_DOUBLE) return 1;
  const double *in = (const double *)sendbuf;
  double *out = (double *)recvbuf;
  std::unique_lock<std::mutex> lk(W.mu);
  if (W.size == 1) { for (int i = 0; i < count; ++i) out[i] = in[i]; return MPI_SUCCESS; }
  long gen = W.generation;
  if ((int)W.contrib.size() != W.size) W.contrib.assign(W.size, std::vector<double>());
  W.contrib[tl_rank].assign(in, in + count);
  if (++W.arrive == W.size) {
    // combine in rank order so the result does not depend on thread arrival order
    W.result = W.contrib[0];
    for (int r = 1; r < W.size; ++r)
      for (int i = 0; i < count; ++i) {
        double v = W.contrib[r][i];
        if (op == MPI_SUM) W.result[i] += v;
        else if (op == MPI_MAX) W.result[i] = v > W.result[i] ? v : W.result[i];
        else W.result[i] = v < W.result[i] ? v : W.result[i];
      }
    W.arrive = 0;
    ++W.generation;
    W.cv.notify_all();
  } else {
    W.cv.wait(lk, [&] { return W.generation != gen; });
  }
  for (int i = 0; i < count; ++i) out[i] = W.result[i];
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm c) {
  double z = 0, o;
  return MPI_Allreduce(&z, &o, 1, MPI_DOUBLE, MPI_SUM, c);
}

}  // extern "C"

// store.cu -- node store: batched KDInsert and the grid-bucket index.
//
// K1a kd_insert_*: restates KDTree::KDInsert (src/kdtree.cpp:87-136) for a
//   whole batch.  Sequential insertion places node i by descending through the
//   nodes with smaller ids; here every pending node descends ONE level per
//   round in lock step, and an empty child slot contended in a round goes to
//   the smallest id (atomicMin) -- the node that would have arrived first in
//   insertion order.  The resulting parent/child/split arrays are identical
//   to the reference's pointer tree for any batch split.
// K1b index build: bounds -> cell id + rank (histogram) -> exclusive scan ->
//   scatter of 32-byte NodeRec records in cell order.
#include <math.h>

#include <algorithm>

#include "store.cuh"

namespace drrt {

static constexpr int32_t EMPTY = 0x7fffffff;
static constexpr double GATE_NONE = 1.0e300;  // "never pruned"

// ---------------------------------------------------------------------------
__global__ void kd_insert_init(int64_t base, int64_t n, const double *__restrict__ pos,
                               double *x, double *y, double *th, double *gate,
                               int32_t *parent, int32_t *left, int32_t *right,
                               uint8_t *split, int32_t *ins_cur, uint8_t *ins_split,
                               double *ins_gate) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n) return;
  int64_t id = base + k;
  x[id] = pos[3 * k + 0];
  y[id] = pos[3 * k + 1];
  th[id] = pos[3 * k + 2];
  left[id] = EMPTY;
  right[id] = EMPTY;
  parent[id] = -1;
  gate[id] = GATE_NONE;
  if (id == 0) {  // kdtree.cpp:94-99
    split[0] = 0;
    ins_cur[k] = -1;
    return;
  }
  // Descent through the nodes that were in the tree before this batch (kdtree.cpp:104-126):
  // their links never change during the batch, so no lock-step round is needed down to the
  // first child slot that is empty or holds a node of the batch -- the rounds start there.
  int32_t c = 0;
  int s = 0;
  double g = GATE_NONE;
  if (base > 0) {
    const double px = pos[3 * k + 0], py = pos[3 * k + 1], pt = pos[3 * k + 2];
    for (;;) {
      const double mine = s == 0 ? px : (s == 1 ? py : pt);
      const double theirs = s == 0 ? x[c] : (s == 1 ? y[c] : th[c]);
      const bool go_left = mine < theirs;
      const int32_t child = go_left ? left[c] : right[c];
      if (child == EMPTY || child >= base) break;  // kd_insert_step takes over at c
      if (go_left && s == 2 && theirs < g) g = theirs;
      c = child;
      s = (s == 2) ? 0 : s + 1;
    }
  } else {
    c = 0;
  }
  ins_cur[k] = c;
  ins_split[k] = (uint8_t)s;  // bit 7 clear: nothing contended yet
  ins_gate[k] = g;
}

// where the sorted record of an indexed node lives (the float copy of its gate follows the gate)
struct RecPatch {
  NodeRec *rec;
  const int32_t *cell_of, *rank, *cell_start;
  int64_t n_indexed;
};

__device__ __forceinline__ double coord(const double *x, const double *y, const double *th,
                                        int64_t i, int s) {
  return s == 0 ? x[i] : (s == 1 ? y[i] : th[i]);
}

// one lock-step round for pending node k; returns true while still pending
__device__ __forceinline__ bool kd_insert_step(int64_t base, int64_t k, const double *x,
                                               const double *y, const double *th, double *gate,
                                               int32_t *parent, int32_t *left, int32_t *right,
                                               uint8_t *split, int32_t *ins_cur,
                                               uint8_t *ins_split, double *ins_gate,
                                               const RecPatch &R) {
  int32_t c = ins_cur[k];
  if (c < 0) return false;
  const int32_t id = (int32_t)(base + k);
  uint8_t st = ins_split[k];
  int s = st & 3;
  double g = ins_gate[k];
  if (st & 0x80) {
    // resolve last round's contention (kdtree.cpp:107-126)
    bool went_left = coord(x, y, th, id, s) < coord(x, y, th, c, s);
    int32_t v = went_left ? left[c] : right[c];
    if (v == id) {
      parent[id] = c;  // :129-132
      split[id] = (uint8_t)((s == 2) ? 0 : s + 1);
      gate[id] = g;
      if (!went_left && s == 2) {
        // c now has a right child: the walk no longer treats c as a leaf when
        // the query lies to its right (kdtree.cpp:722-727 vs :747-756)
        double tc = th[c];
        if (tc < gate[c]) {
          gate[c] = tc;
          if (c < R.n_indexed) R.rec[R.cell_start[R.cell_of[c]] + R.rank[c]].gatef = (float)tc;
        }
      }
      ins_cur[k] = -1;
      return false;
    }
    c = v;
    s = (s == 2) ? 0 : s + 1;
  }
  bool go_left = coord(x, y, th, id, s) < coord(x, y, th, c, s);  // :104-105
  if (go_left && s == 2) {
    double tc = th[c];
    if (tc < g) g = tc;
  }
  int32_t *slot = go_left ? &left[c] : &right[c];
  if (*slot > id) atomicMin(slot, id);
  ins_cur[k] = c;
  ins_split[k] = (uint8_t)(s | 0x80);
  ins_gate[k] = g;
  return true;
}

__global__ void kd_insert_round(int64_t base, int64_t n, const double *x, const double *y,
                                const double *th, double *gate, int32_t *parent,
                                int32_t *left, int32_t *right, uint8_t *split,
                                int32_t *ins_cur, uint8_t *ins_split, double *ins_gate,
                                int32_t *pending, RecPatch R) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  bool p = false;
  if (k < n)
    p = kd_insert_step(base, k, x, y, th, gate, parent, left, right, split, ins_cur, ins_split,
                       ins_gate, R);
  int c = __syncthreads_count(p);
  if (threadIdx.x == 0 && c) atomicAdd(pending, c);
}

// n <= blockDim.x: all rounds inside one block (the planner inserts one node
// per iteration, mainloop.cpp:256 -> drrt.cpp:239-242)
__global__ void kd_insert_small(int64_t base, int64_t n, const double *x, const double *y,
                                const double *th, double *gate, int32_t *parent,
                                int32_t *left, int32_t *right, uint8_t *split,
                                int32_t *ins_cur, uint8_t *ins_split, double *ins_gate,
                                RecPatch R) {
  int64_t k = threadIdx.x;
  for (;;) {
    bool p = false;
    if (k < n)
      p = kd_insert_step(base, k, x, y, th, gate, parent, left, right, split, ins_cur,
                         ins_split, ins_gate, R);
    if (__syncthreads_count(p) == 0) break;
  }
}

int store_reserve_nodes(Store *S, int64_t need) {
  cudaStream_t st = S->stream;
  size_t keep = (size_t)S->n;
  DRRT_CHECK(S->x.reserve(need, st, true, keep));
  DRRT_CHECK(S->y.reserve(need, st, true, keep));
  DRRT_CHECK(S->th.reserve(need, st, true, keep));
  DRRT_CHECK(S->gate.reserve(need, st, true, keep));
  DRRT_CHECK(S->parent.reserve(need, st, true, keep));
  DRRT_CHECK(S->left.reserve(need, st, true, keep));
  DRRT_CHECK(S->right.reserve(need, st, true, keep));
  DRRT_CHECK(S->split.reserve(need, st, true, keep));
  return DRRT_OK;
}

int store_insert_dev(Store *S, int64_t n, const double *pos_dev, cudaStream_t st,
                     int32_t *first_id) {
  if (n < 0) return set_error(DRRT_ERR_INVALID, "insert: n < 0");
  if (first_id) *first_id = (int32_t)S->n;
  if (n == 0) return DRRT_OK;
  if (S->n + n >= (int64_t)EMPTY) return set_error(DRRT_ERR_INVALID, "insert: id space exhausted");
  if (st != S->stream) DRRT_CUDA(cudaStreamSynchronize(S->stream));
  DRRT_CHECK(store_reserve_nodes(S, S->n + n));
  DRRT_CHECK(S->ins_cur.reserve(n, st));
  DRRT_CHECK(S->ins_split.reserve(n, st));
  DRRT_CHECK(S->ins_gate.reserve(n, st));
  const int64_t base = S->n;
  const int T = 256;
  const unsigned G = (unsigned)((n + T - 1) / T);
  kd_insert_init<<<G, T, 0, st>>>(base, n, pos_dev, S->x.p, S->y.p, S->th.p, S->gate.p,
                                  S->parent.p, S->left.p, S->right.p, S->split.p, S->ins_cur.p,
                                  S->ins_split.p, S->ins_gate.p);
  DRRT_LAUNCHED();
  const RecPatch R{S->rec.p, S->cell_of.p, S->rank_in_cell.p, S->cell_start.p, S->n_indexed};
  if (n <= 1024) {
    kd_insert_small<<<1, 1024, 0, st>>>(base, n, S->x.p, S->y.p, S->th.p, S->gate.p,
                                        S->parent.p, S->left.p, S->right.p, S->split.p,
                                        S->ins_cur.p, S->ins_split.p, S->ins_gate.p, R);
    DRRT_LAUNCHED();
  } else {
    // rounds in groups; the pending count of the LAST round of a group decides
    for (int guard = 0; guard < (1 << 22); guard++) {
      const int group = 8;
      for (int r = 0; r < group; r++) {
        if (r == group - 1) DRRT_CUDA(cudaMemsetAsync(S->d_pending, 0, sizeof(int32_t), st));
        kd_insert_round<<<G, T, 0, st>>>(base, n, S->x.p, S->y.p, S->th.p, S->gate.p,
                                         S->parent.p, S->left.p, S->right.p, S->split.p,
                                         S->ins_cur.p, S->ins_split.p, S->ins_gate.p,
                                         S->d_pending, R);
        DRRT_LAUNCHED();
      }
      DRRT_CUDA(cudaMemcpyAsync(S->h_pending, S->d_pending, sizeof(int32_t),
                                cudaMemcpyDeviceToHost, st));
      DRRT_CUDA(cudaStreamSynchronize(st));
      if (*S->h_pending == 0) break;
    }
  }
  DRRT_CUDA(cudaGetLastError());
  S->n += n;
  return DRRT_OK;
}

// ---------------------------------------------------------------------------
// index build

__device__ __forceinline__ unsigned long long enc_f64(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
static double dec_f64(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double v;
  memcpy(&v, &b, sizeof v);
  return v;
}

__global__ void bounds_kernel(int64_t n, const double *x, const double *y, const double *th,
                              unsigned long long *out /* minx maxx miny maxy mint maxt */) {
  unsigned long long mn[3] = {~0ull, ~0ull, ~0ull}, mx[3] = {0, 0, 0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long k0 = enc_f64(x[i]), k1 = enc_f64(y[i]), k2 = enc_f64(th[i]);
    mn[0] = min(mn[0], k0); mx[0] = max(mx[0], k0);
    mn[1] = min(mn[1], k1); mx[1] = max(mx[1], k1);
    mn[2] = min(mn[2], k2); mx[2] = max(mx[2], k2);
  }
  for (int d = 0; d < 3; d++) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = min(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = max(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&out[2 * d], mn[d]);
      atomicMax(&out[2 * d + 1], mx[d]);
    }
  }
}

__device__ __forceinline__ int cell_coord(double v, double v0, double inv_h, int n) {
  double f = floor((v - v0) * inv_h);
  int c = (f < 0.0) ? 0 : ((f >= (double)n) ? n - 1 : (int)f);
  return c;
}

__global__ void cell_assign_kernel(int64_t n, const double *x, const double *y, const double *th,
                                   GridParams gp, int32_t *cell_of, int32_t *rank,
                                   int32_t *count) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int ix = cell_coord(x[i], gp.x0, gp.inv_hx, gp.nx);
  int iy = cell_coord(y[i], gp.y0, gp.inv_hy, gp.ny);
  int it = cell_coord(theta_key(th[i], gp.period), gp.t0, gp.inv_ht, gp.nt);
  int32_t c = (ix * gp.ny + iy) * gp.nt + it;
  cell_of[i] = c;
  rank[i] = atomicAdd(&count[c], 1);
}

__global__ void scatter_kernel(int64_t n, const double *x, const double *y, const double *th,
                               const double *gate, const int32_t *cell_of, const int32_t *rank,
                               const int32_t *cell_start, NodeRec *rec) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  NodeRec r;
  r.x = x[i];
  r.y = y[i];
  r.th = th[i];
  r.id = (int32_t)i;
  r.gatef = (float)gate[i];
  rec[cell_start[cell_of[i]] + rank[i]] = r;
}

int rebuild_index(Store *S, cudaStream_t st) {
  const int64_t n = S->n;
  if (n == 0) {
    S->n_indexed = 0;
    return DRRT_OK;
  }
  unsigned long long init[6] = {~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull};
  DRRT_CUDA(cudaMemcpyAsync(S->d_bounds, init, sizeof init, cudaMemcpyHostToDevice, st));
  const int T = 256;
  int G = (int)std::min<int64_t>((n + T - 1) / T, 148 * 8);
  bounds_kernel<<<G, T, 0, st>>>(n, S->x.p, S->y.p, S->th.p, (unsigned long long *)S->d_bounds);
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaMemcpyAsync(S->h_bounds, S->d_bounds, 6 * sizeof(double),
                            cudaMemcpyDeviceToHost, st));
  DRRT_CUDA(cudaStreamSynchronize(st));
  double b[6];
  for (int i = 0; i < 6; i++) b[i] = dec_f64(((unsigned long long *)S->h_bounds)[i]);
  for (int i = 0; i < 6; i++)
    if (!std::isfinite(b[i]))
      return set_error(DRRT_ERR_UNSUPPORTED, "index: non-finite node coordinate");

  GridParams gp{};
  gp.period = S->period;
  gp.periodic = S->period > 0.0;
  double Lx = std::max(b[1] - b[0], 1e-9), Ly = std::max(b[3] - b[2], 1e-9);
  double t0 = gp.periodic ? 0.0 : b[4];
  double Lt = gp.periodic ? S->period : std::max(b[5] - b[4], 1e-9);
  // cubic cells holding ~2 nodes on average; the metric weighs x, y and theta
  // equally, so the neighbourhood is a ball in these units
  // (measured on config 2a / 2b, profiles/README.md: 1 or 4 nodes per cell and theta cells half,
  // a quarter or twice as deep as wide are all slower or equal)
  const double occupancy = 2.0;
  double cells = std::max(1.0, (double)n / occupancy);
  double h = cbrt(Lx * Ly * Lt / cells);
  for (;;) {
    gp.nx = (int)std::min(4096.0, std::max(1.0, ceil(Lx / h)));
    gp.ny = (int)std::min(4096.0, std::max(1.0, ceil(Ly / h)));
    gp.nt = (int)std::min(256.0, std::max(1.0, ceil(Lt / h)));
    if ((double)gp.nx * gp.ny * gp.nt <= 3.0e8) break;
    h *= 1.26;
  }
  gp.x0 = b[0];
  gp.y0 = b[2];
  gp.t0 = t0;
  gp.inv_hx = gp.nx / Lx;
  gp.inv_hy = gp.ny / Ly;
  gp.inv_ht = gp.nt / Lt;
  gp.hx = Lx / gp.nx;
  gp.hy = Ly / gp.ny;
  gp.ht = Lt / gp.nt;
  const int64_t ncell = (int64_t)gp.nx * gp.ny * gp.nt;

  DRRT_CHECK(S->rec.reserve(n, st));
  DRRT_CHECK(S->cell_of.reserve(n, st));
  DRRT_CHECK(S->rank_in_cell.reserve(n, st));
  DRRT_CHECK(S->cell_start.reserve(ncell + 1, st));
  DRRT_CUDA(cudaMemsetAsync(S->cell_start.p, 0, (ncell + 1) * sizeof(int32_t), st));
  unsigned GN = (unsigned)((n + T - 1) / T);
  cell_assign_kernel<<<GN, T, 0, st>>>(n, S->x.p, S->y.p, S->th.p, gp, S->cell_of.p,
                                       S->rank_in_cell.p, S->cell_start.p);
  DRRT_LAUNCHED();
  DRRT_CHECK(exclusive_scan_i32(S->cell_start.p, S->cell_start.p, ncell + 1, S->scan_tmp, st));
  scatter_kernel<<<GN, T, 0, st>>>(n, S->x.p, S->y.p, S->th.p, S->gate.p, S->cell_of.p,
                                   S->rank_in_cell.p, S->cell_start.p, S->rec.p);
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaGetLastError());
  S->gp = gp;
  S->ncell = ncell;
  S->n_indexed = n;
  return DRRT_OK;
}

// Rebuild when the unsorted tail would cost more to scan than to fold in.
int ensure_index(Store *S, int64_t nq, cudaStream_t st) {
  int64_t tail = S->n - S->n_indexed;
  if (tail == 0) return DRRT_OK;
  // a tiny store is scanned directly, no index
  if (S->n_indexed == 0 && S->n <= 2048) return DRRT_OK;
  bool rebuild = S->n_indexed == 0 || tail > 4096 || tail * std::max<int64_t>(nq, 1) > (1 << 20) ||
                 tail * 8 > S->n;
  if (!rebuild) return DRRT_OK;
  if (st != S->stream) DRRT_CUDA(cudaStreamSynchronize(S->stream));
  return rebuild_index(S, st);
}

}  // namespace drrt

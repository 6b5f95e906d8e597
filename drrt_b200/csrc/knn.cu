// knn.cu -- K3: batched nearest / k-nearest.
//
// KDTree::KDFindNearest (src/kdtree.cpp:264-307) returns the node of smallest
// metric distance; KDTree::KDFindKNearest (:630-666) has no caller and does
// not terminate as written, so it is built by intent: the k
// smallest metric(q, n), ties by smaller id, ascending (dist, id).
//
// One warp per query.  The warp keeps its k best (dist, id) pairs sorted in
// shared memory, scans the cells of a ball whose radius starts at the size
// expected to hold ~2k nodes and doubles until the k-th best distance is
// inside the scanned ball (nodes outside it are provably farther), and merges
// candidates that beat the current k-th with ballot + warp-cooperative insert.
#include <math.h>

#include <algorithm>

#include "gridscan.cuh"

namespace drrt {

static constexpr int KNN_WARPS = 8;
static constexpr int KNN_KMAX = 128;

struct KnnArgs {
  GridView V;
  const int32_t *qlist;  // null, or [count, q0, q1, ...]: only these queries (knn_select_kernel's leftovers)
  const double *q;
  int64_t nq;
  int k;
  int nearest;  // k == 1 with KDFindNearest's ghost pass
  double r0;
  int32_t *ids;
  double *dist;
};

struct TopK {
  double *d;   // shared, k entries ascending
  int32_t *id;
  int k;
  int count;        // warp-uniform
  double worst_d;   // warp-uniform (count == k), else +inf
  int32_t worst_id;
};

__device__ __forceinline__ bool pair_less(double d0, int32_t i0, double d1, int32_t i1) {
  return d0 < d1 || (d0 == d1 && i0 < i1);
}

// all lanes call with the same (d, id)
__device__ __forceinline__ void topk_insert(TopK &T, double d, int32_t id, int lane) {
  // position = number of kept pairs smaller than (d,id)
  int pos = 0;
  for (int b = 0; b < T.count; b += 32) {
    int i = b + lane;
    bool less = i < T.count && pair_less(T.d[i], T.id[i], d, id);
    pos += __popc(__ballot_sync(0xffffffffu, less));
  }
  int last = (T.count < T.k) ? T.count : T.k - 1;  // slot that receives the shifted tail
  // shift [pos, last) right by one, highest first
  for (int b = ((last - 1) / 32) * 32; b >= 0 && last > pos; b -= 32) {
    int i = b + lane;
    double td = 0.0;
    int32_t ti = 0;
    bool mv = i >= pos && i < last;
    if (mv) { td = T.d[i]; ti = T.id[i]; }
    __syncwarp();
    if (mv) { T.d[i + 1] = td; T.id[i + 1] = ti; }
    __syncwarp();
  }
  if (lane == 0) { T.d[pos] = d; T.id[pos] = id; }
  __syncwarp();
  if (T.count < T.k) T.count++;
  if (T.count == T.k) { T.worst_d = T.d[T.k - 1]; T.worst_id = T.id[T.k - 1]; }
}

template <int METRIC>
// Five CTAs per SM (51 registers) instead of the two that ptxas' own 92 registers allow: both kernels wait
// on L2 round trips, and warps are what hides them -- k = 16: 0.210 -> 0.184 ms, nearest 0.139 -> 0.111 ms per
// 64 K queries (three CTAs 0.210 / 0.138, four 0.186 / 0.119, six 0.180 / 0.111)
#ifndef DRRT_KNN_MINB
#define DRRT_KNN_MINB 5
#endif
__global__ void __launch_bounds__(KNN_WARPS * 32, DRRT_KNN_MINB)
knn_kernel(KnnArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int2 s_runl[KNN_WARPS][RUNS_PER_BATCH + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // plain mode: one query per warp; list mode: the warps share the listed queries
  const int64_t count = A.qlist ? (int64_t)A.qlist[0] : A.nq;
  for (int64_t slot = (int64_t)blockIdx.x * KNN_WARPS + w; slot < count; slot += (int64_t)gridDim.x * KNN_WARPS) {
  const int64_t qi = A.qlist ? (int64_t)A.qlist[slot + 1] : slot;
  const GridView &V = A.V;
  TopK T;
  T.k = A.k;
  T.d = reinterpret_cast<double *>(smem_raw) + (size_t)w * A.k;
  T.id = reinterpret_cast<int32_t *>(smem_raw + sizeof(double) * KNN_WARPS * A.k) + (size_t)w * A.k;
  const double qx = A.q[3 * qi], qy = A.q[3 * qi + 1], qt = A.q[3 * qi + 2];

  double px = qx, py = qy, pt = qt;  // current probe
  auto visit = [&](bool valid, double nx, double ny, double nt, int32_t id, float) {
    bool cand = false;
    double d = 0.0;
    if (valid) {
      double s = metric_sum<METRIC>(px, py, pt, nx, ny, nt);
      d = sqrt(s);
      cand = (T.count < T.k) || pair_less(d, id, T.worst_d, T.worst_id);
    }
    unsigned m = __ballot_sync(0xffffffffu, cand);
    while (m) {
      int src = __ffs(m) - 1;
      m &= m - 1;
      double cd = __shfl_sync(0xffffffffu, d, src);
      int32_t ci = __shfl_sync(0xffffffffu, id, src);
      if ((T.count < T.k) || pair_less(cd, ci, T.worst_d, T.worst_id)) topk_insert(T, cd, ci, lane);
    }
  };
  auto search = [&]() {
    double r = A.r0;
    for (;;) {
      T.count = 0;
      T.worst_d = 1.0e308;
      T.worst_id = 0x7fffffff;
      Query Q;
      bool all = true;
      if (V.n_indexed > 0) {
        setup_query<METRIC, false>(V, px, py, pt, r, Q);
        Q.has_ghost = false;
        all = covers_all(V, Q);
        scan_columns(V, Q, 0, 1, s_runl[w], lane, visit);
      }
      scan_tail(V, 0, 1, lane, visit);
      // every unscanned node is farther than r
      if (all || (T.count == T.k && T.worst_d <= r)) break;
      r *= 2.0;
    }
    __syncwarp();
  };
  search();
  if (A.nearest && V.has_wrap && T.count == 1) {
    // KDFindNearest's ghost pass (kdtree.cpp:277-303): the ghost is searched when
    // it can beat the current nearest distance, and wins only when strictly closer
    double d1 = T.d[0];
    int32_t id1 = T.id[0];
    double gt;
    if (ghost_of<METRIC>(qx, qy, qt, V.wrap_point, d1, &gt)) {
      __syncwarp();
      pt = gt;
      search();
      if (!(T.d[0] < d1)) {
        __syncwarp();
        if (lane == 0) { T.d[0] = d1; T.id[0] = id1; }
      }
    }
  }
  __syncwarp();
  for (int i = lane; i < A.k; i += 32) {
    bool have = i < T.count;
    A.ids[qi * A.k + i] = have ? T.id[i] : -1;
    A.dist[qi * A.k + i] = have ? T.d[i] : DRRT_INF;
  }
  __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// k <= KSEL_KMAX: collect, then select.  The warp gathers every node within the current radius
// (ballot-compacted into its staging area, like a range query), and once at least k are inside
// -- all nodes within r are then known, so the k nearest are among them -- orders them with a
// small bitonic network on (dist, id) and writes the first k.  No per-candidate insertion.
static constexpr int KSEL_KMAX = 64;
static constexpr int KSEL_CAP = 256;  // staged candidates per query; more than that: knn_kernel

// all-ascending bitonic network over n (any n) (d, id) pairs, one warp
__device__ void warp_sort_pairs(double *d, int32_t *id, int n, int lane) {
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  int lk = 1;
  auto cex = [&](int i, int j) {
    if (j < n) {
      const double di = d[i], dj = d[j];
      const int32_t ii = id[i], ij = id[j];
      if (pair_less(dj, ij, di, ii)) { d[i] = dj; id[i] = ij; d[j] = di; id[j] = ii; }
    }
  };
  for (int k = 2; k <= np2; k <<= 1, lk++) {
    const int half = k >> 1;
    for (int p = lane; p < (np2 >> 1); p += 32) {
      const int blk = p >> (lk - 1), o = p & (half - 1);
      cex(blk * k + o, blk * k + k - 1 - o);
    }
    __syncwarp();
    for (int jj = half >> 1; jj > 0; jj >>= 1) {
      for (int p = lane; p < (np2 >> 1); p += 32) {
        const int i = ((p & ~(jj - 1)) << 1) | (p & (jj - 1));
        cex(i, i + jj);
      }
      __syncwarp();
    }
  }
}

template <int METRIC>
__global__ void __launch_bounds__(KNN_WARPS * 32, DRRT_KNN_MINB)
knn_select_kernel(KnnArgs A, int32_t *fallback) {
  __shared__ int2 s_runl[KNN_WARPS][RUNS_PER_BATCH + 1];
  __shared__ double s_d[KNN_WARPS][KSEL_CAP];
  __shared__ int32_t s_id[KNN_WARPS][KSEL_CAP];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t qi = (int64_t)blockIdx.x * KNN_WARPS + w;
  if (qi >= A.nq) return;
  const GridView &V = A.V;
  const double qx = A.q[3 * qi], qy = A.q[3 * qi + 1], qt = A.q[3 * qi + 2];
  const unsigned ltmask = (1u << lane) - 1u;
  double *sd = s_d[w];
  int32_t *sid = s_id[w];
  double r = A.r0;
  bool all = false;
  int cnt = 0;
  auto visit = [&](bool valid, double nx, double ny, double nt, int32_t id, float) {
    const double d = sqrt(metric_sum<METRIC>(qx, qy, qt, nx, ny, nt));
    const bool in = valid && (all || d <= r);
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (!m) return;
    const int pos = cnt;  // warp-uniform register: every lane sees the ballot
    cnt += __popc(m);
    if (in) {
      const int slot = pos + __popc(m & ltmask);
      if (slot < KSEL_CAP) { sd[slot] = d; sid[slot] = id; }
    }
  };
  int n = 0;
  for (;;) {
    cnt = 0;
    __syncwarp();
    all = true;
    if (V.n_indexed > 0) {
      Query Q;
      setup_query<METRIC, false>(V, qx, qy, qt, r, Q);
      Q.has_ghost = false;
      all = covers_all(V, Q);
      scan_columns(V, Q, 0, 1, s_runl[w], lane, visit);
    }
    scan_tail(V, 0, 1, lane, visit);
    __syncwarp();
    n = cnt;
    // every node within r is staged: k of them settle the answer (unscanned nodes are farther)
    if (n >= A.k || all || n > KSEL_CAP) break;
    r *= 2.0;
  }
  if (n > KSEL_CAP) {  // a crowded neighbourhood: leave the query to the insertion kernel
    if (lane == 0) fallback[atomicAdd(&fallback[0], 1) + 1] = (int32_t)qi;
    return;
  }
  warp_sort_pairs(sd, sid, n, lane);
  for (int i = lane; i < A.k; i += 32) {
    const bool have = i < n;
    A.ids[qi * A.k + i] = have ? sid[i] : -1;
    A.dist[qi * A.k + i] = have ? sd[i] : DRRT_INF;
  }
}

int knn_batch_dev(Store *S, int64_t nq, const double *q_dev, int k, int32_t *ids_dev,
                  double *dist_dev, cudaStream_t st, bool nearest) {
  if (nq < 0 || k < 1) return set_error(DRRT_ERR_INVALID, "knn: bad nq/k");
  if (k > KNN_KMAX) return set_error(DRRT_ERR_UNSUPPORTED, "knn: k > %d", KNN_KMAX);
  if (nq == 0) return DRRT_OK;
  DRRT_CHECK(ensure_index(S, nq, st));
  KnnArgs A{};
  A.V = make_view(S);
  A.q = q_dev; A.nq = nq; A.k = k; A.ids = ids_dev; A.dist = dist_dev;
  A.nearest = (nearest && k == 1) ? 1 : 0;
  A.r0 = 1.0;
  if (S->n_indexed > 0) {
    const GridParams &g = S->gp;
    double vol = (g.nx / g.inv_hx) * (g.ny / g.inv_hy) * (g.nt / g.inv_ht);
    double want = 2.0 * k + 8.0;
    A.r0 = cbrt(want * vol / ((double)S->n_indexed * 4.18879));
    if (!(A.r0 > 0.0) || !std::isfinite(A.r0)) A.r0 = 1.0;
  }
  size_t smem = (size_t)KNN_WARPS * k * (sizeof(double) + sizeof(int32_t));
  unsigned nb = (unsigned)((nq + KNN_WARPS - 1) / KNN_WARPS);
  if (k <= KSEL_KMAX && !A.nearest && nq < 0x7ffffff0) {
    // collect-then-select; queries whose ball holds more than KSEL_CAP nodes are listed and
    // answered by the insertion kernel afterwards (none on an evenly filled store)
    DRRT_CHECK(S->i_stage2.reserve((size_t)nq + 2, st));
    int32_t *fb = S->i_stage2.p;
    DRRT_CUDA(cudaMemsetAsync(fb, 0, sizeof(int32_t), st));
    if (S->metric == DRRT_METRIC_EUCLID)
      knn_select_kernel<DRRT_METRIC_EUCLID><<<nb, KNN_WARPS * 32, 0, st>>>(A, fb);
    else
      knn_select_kernel<DRRT_METRIC_R2S1_WRAP><<<nb, KNN_WARPS * 32, 0, st>>>(A, fb);
    DRRT_LAUNCHED();
    A.qlist = fb;  // the list length is only known on the device: a small grid strides over it
    nb = std::min(nb, 592u);
  }
  if (S->metric == DRRT_METRIC_EUCLID)
    knn_kernel<DRRT_METRIC_EUCLID><<<nb, KNN_WARPS * 32, smem, st>>>(A);
  else
    knn_kernel<DRRT_METRIC_R2S1_WRAP><<<nb, KNN_WARPS * 32, smem, st>>>(A);
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaGetLastError());
  return DRRT_OK;
}

}  // namespace drrt

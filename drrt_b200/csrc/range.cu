// range.cu -- K2: batched range-ball query.
//
// Replaces KDTree::KDFindWithinRange / KDFindMoreWithinRange
// (src/kdtree.cpp:794-851), KDFindWithinRangeInSubtree (:703-792),
// AddToRangeList (:670-682) and the ghost loop (src/ghostpoint.cpp:15-75).
//
// Result set of the reference walk, restated without the tree walk
// (DESIGN.md "Range predicate"): with probes P = {q} U {ghost of q},
//   hit(n)  <=>  (n == root and metric(q,n) <= r)
//            or  exists p in P: metric(p,n) < r  and  not (p.theta - gate[n] > r)
// The last clause is the only effect the signed hyperplane prune
// (kdtree.cpp:744-756) can have on the SET: for x/y splits a pruned node is
// provably out of range in IEEE arithmetic; for theta splits the metric wraps
// but the prune does not, and gate[n] (store.cuh) records the smallest theta
// hyperplane that hides n.  The recorded key is the distance to the first
// probe that admits n, as the reference's JList key (AddToRangeList).
//
// Execution: one single pass.  A warp (small neighbourhoods) or a whole CTA
// (large ones) scans the candidate columns of its query in the cell-sorted
// 32-byte NodeRec array -- all columns of a batch flattened so every lane has
// a candidate -- evaluates the f64 metric in the reference's operation order,
// compacts hits with ballots into shared memory, sorts them by node id there,
// and writes them to their final CSR position, which it gets from an
// in-kernel ordered allocation (decoupled look-back over per-CTA hit totals).
#include <math.h>

#include <algorithm>

#include "gridscan.cuh"

namespace drrt {

static constexpr int RANGE_WARPS = 8;
static constexpr int RANGE_T = RANGE_WARPS * 32;
static constexpr int WARP_STAGE = 256;    // hits staged per warp-mode query
static constexpr int BLOCK_STAGE = 5120;  // hits staged per CTA-mode query

struct RangeArgs {
  GridView V;
  const double *q;
  const double *radius;
  double radius_all;
  int64_t nq;
  int64_t *offsets;
  int32_t *ids;
  double *dist;
  int64_t cap;
  unsigned long long *desc;
  int32_t *ticket;
};

// does the reference's walk from probe theta `pt` ever evaluate node n?
__device__ __forceinline__ bool walk_reaches(double pt, double nt, double r, int32_t id,
                                             const double *gate) {
  if (!(pt - nt > r)) return true;  // no theta hyperplane at or above n.theta can prune
  return !(pt - gate[id] > r);
}

template <int METRIC>
__device__ __forceinline__ bool eval_candidate(const Query &Q, double nx, double ny, double nt,
                                               int32_t id, const double *gate, double *dout) {
  double s = metric_sum<METRIC>(Q.x, Q.y, Q.t, nx, ny, nt);
  if (s < Q.s_hi) {
    double d = sqrt(s);
    if (id == 0) {
      if (d <= Q.r) { *dout = d; return true; }  // kdtree.cpp:800-802
    } else if (d < Q.r) {                         // kdtree.cpp:733, :765
      if (METRIC == DRRT_METRIC_EUCLID || walk_reaches(Q.t, nt, Q.r, id, gate)) {
        *dout = d;
        return true;
      }
    }
  }
  if (Q.has_ghost) {  // kdtree.cpp:807-819
    double s2 = metric_sum<METRIC>(Q.x, Q.y, Q.gt, nx, ny, nt);
    if (s2 < Q.s_hi) {
      double d = sqrt(s2);
      if (d < Q.r && (METRIC == DRRT_METRIC_EUCLID || walk_reaches(Q.gt, nt, Q.r, id, gate))) {
        *dout = d;
        return true;
      }
    }
  }
  return false;
}

// Where a scanning warp puts its hits.
struct Sink {
  unsigned long long *keys;  // shared: (id << 32) | slot
  double *sdist;             // shared
  int *counter;              // shared
  int capacity;
  // direct mode (second pass of an overflowed query)
  int32_t *gids;
  double *gdist;
};

__device__ __forceinline__ void emit(const Sink &K, bool hit, int32_t id, double d, int lane) {
  unsigned m = __ballot_sync(0xffffffffu, hit);
  if (!m) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(K.counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (hit) {
    int slot = base + __popc(m & ((1u << lane) - 1));
    if (K.gids) {
      K.gids[slot] = id;
      K.gdist[slot] = d;
    } else if (slot < K.capacity) {
      K.keys[slot] = ((unsigned long long)(unsigned)id << 32) | (unsigned)slot;
      K.sdist[slot] = d;
    }
  }
}

// all-ascending bitonic network over n (any n) keys; `nthreads` threads, rank `tid`
template <bool BLOCK>
__device__ __forceinline__ void group_sync() {
  if (BLOCK) __syncthreads(); else __syncwarp();
}

template <bool BLOCK>
__device__ void sort_keys_shared(unsigned long long *keys, int n, int tid, int nthreads) {
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int k = 2; k <= np2; k <<= 1) {
    const int half = k >> 1;
    for (int p = tid; p < (np2 >> 1); p += nthreads) {
      int blk = p / half, o = p - blk * half;
      int i = blk * k + o, j = blk * k + k - 1 - o;
      if (j < n) {
        unsigned long long a = keys[i], b = keys[j];
        if (a > b) { keys[i] = b; keys[j] = a; }
      }
    }
    group_sync<BLOCK>();
    for (int jj = half >> 1; jj > 0; jj >>= 1) {
      for (int p = tid; p < (np2 >> 1); p += nthreads) {
        int i = 2 * jj * (p / jj) + (p % jj), j = i + jj;
        if (j < n) {
          unsigned long long a = keys[i], b = keys[j];
          if (a > b) { keys[i] = b; keys[j] = a; }
        }
      }
      group_sync<BLOCK>();
    }
  }
}

// same network on (ids, dist) pairs in global memory (overflowed queries only)
template <bool BLOCK>
__device__ void sort_pairs_global(int32_t *ids, double *dist, int64_t n, int tid, int nthreads) {
  int64_t np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int64_t k = 2; k <= np2; k <<= 1) {
    const int64_t half = k >> 1;
    for (int64_t p = tid; p < (np2 >> 1); p += nthreads) {
      int64_t blk = p / half, o = p - blk * half;
      int64_t i = blk * k + o, j = blk * k + k - 1 - o;
      if (j < n && ids[i] > ids[j]) {
        int32_t a = ids[i]; ids[i] = ids[j]; ids[j] = a;
        double b = dist[i]; dist[i] = dist[j]; dist[j] = b;
      }
    }
    group_sync<BLOCK>();
    for (int64_t jj = half >> 1; jj > 0; jj >>= 1) {
      for (int64_t p = tid; p < (np2 >> 1); p += nthreads) {
        int64_t i = 2 * jj * (p / jj) + (p % jj), j = i + jj;
        if (j < n && ids[i] > ids[j]) {
          int32_t a = ids[i]; ids[i] = ids[j]; ids[j] = a;
          double b = dist[i]; dist[i] = dist[j]; dist[j] = b;
        }
      }
      group_sync<BLOCK>();
    }
  }
}

// Ordered allocation: CTA `vb` publishes its hit total and receives the sum of
// all earlier CTAs' totals (decoupled look-back; vb comes from a ticket so
// every predecessor is already running).
static constexpr unsigned long long FLAG_AGG = 1ull << 62, FLAG_PREFIX = 2ull << 62;
static constexpr unsigned long long VAL_MASK = (1ull << 62) - 1;

__device__ long long lookback_exclusive(unsigned long long *desc, int vb, long long agg, int lane) {
  volatile unsigned long long *vd = desc;
  if (vb == 0) {
    if (lane == 0) { __threadfence(); vd[0] = FLAG_PREFIX | (unsigned long long)agg; }
    return 0;
  }
  if (lane == 0) { vd[vb] = FLAG_AGG | (unsigned long long)agg; }
  long long excl = 0;
  int pos = vb - 1;
  for (;;) {
    // each lane inspects one predecessor, most recent first
    int idx = pos - lane;
    unsigned long long v = 0;
    if (idx >= 0) {
      do { v = vd[idx]; } while ((v >> 62) == 0);
    }
    unsigned has_prefix = __ballot_sync(0xffffffffu, idx >= 0 && (v >> 62) == 2);
    int stop = has_prefix ? (__ffs(has_prefix) - 1) : 31;
    long long contrib = (idx >= 0 && lane <= stop) ? (long long)(v & VAL_MASK) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
    excl += contrib;
    if (has_prefix || pos - 32 < 0) break;
    pos -= 32;
  }
  if (lane == 0) { __threadfence(); vd[vb] = FLAG_PREFIX | (unsigned long long)(excl + agg); }
  return excl;
}

template <int METRIC, bool BLOCK>
__global__ void __launch_bounds__(RANGE_T)
range_kernel(RangeArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_vb;
  __shared__ int s_count[RANGE_WARPS];
  __shared__ long long s_base;
  __shared__ int s_runs[RANGE_WARPS][RUNS_PER_BATCH];
  __shared__ int s_pref[RANGE_WARPS][RUNS_PER_BATCH + 1];

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_vb = atomicAdd(A.ticket, 1);
  if (threadIdx.x < RANGE_WARPS) s_count[threadIdx.x] = 0;
  __syncthreads();
  const int vb = s_vb;

  constexpr int STAGE = BLOCK ? BLOCK_STAGE : WARP_STAGE;
  constexpr int NSTAGE = BLOCK ? 1 : RANGE_WARPS;
  unsigned long long *all_keys = reinterpret_cast<unsigned long long *>(smem_raw);
  double *all_dist = reinterpret_cast<double *>(smem_raw + sizeof(unsigned long long) * STAGE * NSTAGE);

  const int64_t qi = BLOCK ? (int64_t)vb : (int64_t)vb * RANGE_WARPS + w;
  const bool active = qi < A.nq;
  Query Q;
  Sink K;
  K.keys = all_keys + (BLOCK ? 0 : w * STAGE);
  K.sdist = all_dist + (BLOCK ? 0 : w * STAGE);
  K.counter = &s_count[BLOCK ? 0 : w];
  K.capacity = STAGE;
  K.gids = nullptr;
  K.gdist = nullptr;
  const GridView &V = A.V;
  auto visit = [&](bool valid, double nx, double ny, double nt, int32_t id) {
    double d = 0.0;
    bool hit = valid && eval_candidate<METRIC>(Q, nx, ny, nt, id, V.gate, &d);
    emit(K, hit, id, d, lane);
  };
  const int first = BLOCK ? w : 0, stride = BLOCK ? RANGE_WARPS : 1;
  if (active) {
    setup_query<METRIC>(V, A.q[3 * qi], A.q[3 * qi + 1], A.q[3 * qi + 2],
                        A.radius ? A.radius[qi] : A.radius_all, Q);
    if (V.n_indexed > 0) scan_columns(V, Q, first, stride, s_runs[w], s_pref[w], lane, visit);
    scan_tail(V, first, stride, lane, visit);
  }
  __syncthreads();

  // ---- ordered allocation of the CTA's output range
  if (w == 0) {
    long long agg = 0;
    if (BLOCK) agg = s_count[0];
    else {
      int c = (lane < RANGE_WARPS) ? s_count[lane] : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      agg = c;
    }
    long long excl = lookback_exclusive(A.desc, vb, agg, lane);
    if (lane == 0) s_base = excl;
  }
  __syncthreads();
  long long my_off = s_base;
  if (!BLOCK)
    for (int i = 0; i < w; i++) my_off += s_count[i];
  const int my_cnt = s_count[BLOCK ? 0 : w];
  if (active && (BLOCK ? threadIdx.x == 0 : lane == 0)) {
    A.offsets[qi] = my_off;
    if (qi == A.nq - 1) A.offsets[A.nq] = my_off + my_cnt;
  }
  if (!active || my_cnt == 0) return;       // uniform per group
  if (my_off + my_cnt > A.cap) return;      // arena too small: host grows it and reruns

  const int tid = BLOCK ? (int)threadIdx.x : lane;
  const int nthr = BLOCK ? RANGE_T : 32;
  int32_t *oid = A.ids + my_off;
  double *od = A.dist + my_off;
  if (my_cnt <= STAGE) {
    sort_keys_shared<BLOCK>(K.keys, my_cnt, tid, nthr);
    for (int i = tid; i < my_cnt; i += nthr) {
      unsigned long long k = K.keys[i];
      oid[i] = (int32_t)(k >> 32);
      od[i] = K.sdist[(unsigned)(k & 0xffffffffu)];
    }
  } else {
    // more hits than the staging area: second pass straight into the arena
    group_sync<BLOCK>();
    if (tid == 0) *K.counter = 0;
    group_sync<BLOCK>();
    K.gids = oid;
    K.gdist = od;
    if (V.n_indexed > 0) scan_columns(V, Q, first, stride, s_runs[w], s_pref[w], lane, visit);
    scan_tail(V, first, stride, lane, visit);
    group_sync<BLOCK>();
    __threadfence_block();
    sort_pairs_global<BLOCK>(oid, od, my_cnt, tid, nthr);
  }
}

template <int METRIC>
static int launch_range(const RangeArgs &A, bool block_mode, cudaStream_t st) {
  if (block_mode) {
    size_t smem = (size_t)BLOCK_STAGE * 16;
    static bool attr_done = false;
    if (!attr_done) {
      DRRT_CUDA(cudaFuncSetAttribute(range_kernel<METRIC, true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_done = true;
    }
    range_kernel<METRIC, true><<<(unsigned)A.nq, RANGE_T, smem, st>>>(A);
  } else {
    size_t smem = (size_t)WARP_STAGE * RANGE_WARPS * 16;
    unsigned nb = (unsigned)((A.nq + RANGE_WARPS - 1) / RANGE_WARPS);
    range_kernel<METRIC, false><<<nb, RANGE_T, smem, st>>>(A);
  }
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaGetLastError());
  return DRRT_OK;
}

__global__ void max_radius_kernel(const double *r, int64_t n, unsigned long long *out) {
  double m = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double v = fabs(r[i]);
    if (v > m) m = v;
  }
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

int range_batch_dev(Store *S, int64_t nq, const double *q_dev, const double *radius_dev,
                    double radius_all, cudaStream_t st, drrt_range_result *out) {
  if (nq < 0) return set_error(DRRT_ERR_INVALID, "range: nq < 0");
  if (nq > 0x7ffffff0) return set_error(DRRT_ERR_INVALID, "range: nq too large");
  S->last_nq = -1;
  DRRT_CHECK(S->r_offsets.reserve(nq + 1, st));
  if (nq == 0 || S->n == 0) {
    DRRT_CUDA(cudaMemsetAsync(S->r_offsets.p, 0, (nq + 1) * sizeof(int64_t), st));
    S->last_nq = nq;
    S->last_total = 0;
    if (out) *out = drrt_range_result{nq, 0, S->r_offsets.p, S->r_ids.p, S->r_dist.p};
    return DRRT_OK;
  }
  DRRT_CHECK(ensure_index(S, nq, st));

  // expected neighbourhood size decides warp- vs CTA-per-query
  double rmax = fabs(radius_all);
  if (radius_dev) {
    DRRT_CUDA(cudaMemsetAsync(S->d_pending, 0, 8, st));
    max_radius_kernel<<<64, 256, 0, st>>>(radius_dev, nq, (unsigned long long *)S->d_pending);
    DRRT_LAUNCHED();
    DRRT_CUDA(cudaMemcpyAsync(S->h_scalar, S->d_pending, 8, cudaMemcpyDeviceToHost, st));
    DRRT_CUDA(cudaStreamSynchronize(st));
    memcpy(&rmax, S->h_scalar, 8);
  }
  double expected = (double)S->n;
  if (S->n_indexed > 0) {
    const GridParams &g = S->gp;
    double vol = (g.nx / g.inv_hx) * (g.ny / g.inv_hy) * (g.nt / g.inv_ht);
    double ball = 4.18879 * rmax * rmax * rmax;
    expected = std::min((double)S->n, (double)S->n * ball / std::max(vol, 1e-300));
  }
  bool block_mode = expected > WARP_STAGE / 2;

  int64_t nvb = block_mode ? nq : (nq + RANGE_WARPS - 1) / RANGE_WARPS;
  DRRT_CHECK(S->r_desc.reserve(nvb, st));
  int64_t guess = (int64_t)std::min(1.0e12, expected * 1.3 * (double)nq) + 1024;
  if ((int64_t)S->r_ids.cap < guess) {
    DRRT_CHECK(S->r_ids.reserve(guess, st));
    DRRT_CHECK(S->r_dist.reserve(guess, st));
  }

  RangeArgs A{};
  A.V = make_view(S);
  A.q = q_dev; A.radius = radius_dev; A.radius_all = radius_all; A.nq = nq;
  A.offsets = S->r_offsets.p; A.desc = S->r_desc.p; A.ticket = S->d_ticket;

  for (int attempt = 0; attempt < 3; attempt++) {
    A.ids = S->r_ids.p; A.dist = S->r_dist.p;
    A.cap = (int64_t)std::min(S->r_ids.cap, S->r_dist.cap);
    DRRT_CUDA(cudaMemsetAsync(S->r_desc.p, 0, nvb * sizeof(unsigned long long), st));
    DRRT_CUDA(cudaMemsetAsync(S->d_ticket, 0, sizeof(int32_t), st));
    if (S->metric == DRRT_METRIC_EUCLID) DRRT_CHECK(launch_range<DRRT_METRIC_EUCLID>(A, block_mode, st));
    else DRRT_CHECK(launch_range<DRRT_METRIC_R2S1_WRAP>(A, block_mode, st));
    DRRT_CUDA(cudaMemcpyAsync(S->h_scalar, S->r_offsets.p + nq, sizeof(int64_t),
                              cudaMemcpyDeviceToHost, st));
    DRRT_CUDA(cudaStreamSynchronize(st));
    int64_t total = S->h_scalar[0];
    if (total <= A.cap) {
      S->last_nq = nq;
      S->last_total = total;
      if (out) *out = drrt_range_result{nq, total, S->r_offsets.p, S->r_ids.p, S->r_dist.p};
      return DRRT_OK;
    }
    DRRT_CHECK(S->r_ids.reserve(total + total / 8 + 1024, st));
    DRRT_CHECK(S->r_dist.reserve(total + total / 8 + 1024, st));
  }
  return set_error(DRRT_ERR_STATE, "range: result arena did not converge");
}

}  // namespace drrt

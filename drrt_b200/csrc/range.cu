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
// Execution: one pass, one kernel.  A warp (small neighbourhoods) or a whole
// CTA (large ones) scans the candidate columns of its query in the cell-sorted
// 32-byte NodeRec array -- runs flattened so every lane holds a candidate --
// evaluates the f64 metric in the reference's operation order and compacts the
// hits with ballots into shared memory, orders them by node id (counting sort
// on id buckets) and writes the row with coalesced stores.  Persistent warps /
// CTAs take SLICES of consecutive queries from a ticket and write their rows
// back to back into the slice's share of the result arena (rows addressed by
// offsets[q] / counts[q]), so nothing is allocated in order across CTAs and
// nobody waits for a predecessor; range_pack_dev makes a plain CSR of it.
#include <math.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "gridscan.cuh"

namespace drrt {

static constexpr int WARP_MODE_WARPS = 8;    // one query per warp
static constexpr int WARP_STAGE = 256;       // hits staged per warp-mode query

struct RangeArgs {
  GridView V;
  const double *q;
  const double *radius;
  double radius_all;
  int64_t nq;
  int64_t *offsets;
  int32_t *ids;
  double *dist;
  int32_t *ticket;
  int32_t *counts;
  // CTA kernel: arena slices of chunk_len (<= CTA_CHUNK) rows
  int chunk_len;
  int64_t nchunks, chunk_cap;
  const int64_t *chunk_base;  // exact slices (rerun) or null: slice c starts at c * chunk_cap
  int64_t *chunk_total;
  unsigned long long *totals;  // [0] hits, [1] slices that overflowed
};

// does the reference's walk from probe theta `pt` ever evaluate node n?
__device__ __forceinline__ bool walk_reaches(double pt, double nt, double r, int32_t id,
                                             const double *gate) {
  if (!(pt - nt > r)) return true;  // no theta hyperplane at or above n.theta can prune
  return !(pt - gate[id] > r);
}

// Why the ghost can be skipped (wrap-aware metric, wrap point == the metric's period W):
// with a = |q.theta - n.theta| the first probe's theta term is min(a, W - a), whose square is
// the squared distance of a to the nearer of {0, W}; the ghost q.theta +- W sees
// a' = |q.theta +- W - n.theta| instead.  For a <= 1.5 W that distance never shrinks
// (it stays equal on the far side of the seam and grows on the near side), so
// metric(ghost, n) >= metric(q, n) in exact arithmetic; in IEEE double each theta term
// carries at most a few roundings of magnitudes <= |q.theta| + |n.theta| + 2W, which the
// band (2^-46 relative to that magnitude, against 2^-52 roundings) covers.  Candidates
// outside |n.theta - q.theta| <= 1.45 W or |n.theta| <= 64 always take the ghost pass, and
// so does a candidate the first probe finds in range but the signed prune hides.
struct GhostBand {
  double s_band;    // first-probe squared sums at or above this cannot be rescued by the ghost
  double nlo, nhi;  // n.theta interval in which that holds
};

template <int METRIC>
__device__ __forceinline__ GhostBand ghost_band(const Query &Q) {
  GhostBand B;
  const double band = Q.r + 1.4210854715202004e-14 * (fabs(Q.t) + 64.0 + 2.0 * DRRT_TWO_PI + fabs(Q.r));
  B.s_band = band * band * (1.0 + 1.4210854715202004e-14);
  B.nlo = fmax(-64.0, Q.t - 1.45 * DRRT_TWO_PI);
  B.nhi = fmin(64.0, Q.t + 1.45 * DRRT_TWO_PI);
  if (METRIC == DRRT_METRIC_EUCLID || !(fabs(Q.t) <= 64.0) || !(B.s_band == B.s_band)) {
    B.s_band = 1.0e308;  // every rejected candidate takes the ghost pass
    B.nlo = 1.0;
    B.nhi = -1.0;
  }
  return B;
}

// walk_reaches' second clause, `!(pt - gate[id] > r)`, from the float copy of the gate that came
// with the record: the copy is within 2^-24 relative of gate[id], the subtraction rounds once more,
// so outside a band of a few float ulps around r the comparison on the copy IS the comparison on
// the double.  Inside the band (or with no copy: NaN) the candidate asks for the exact look-up.
__device__ __forceinline__ void gate_from_copy(double pt, double r, float gf, bool &reached, bool &unsure) {
  const double a = pt - (double)gf;
  // (a node without a gate holds 1e300, +inf as a float: the clamp keeps tol finite and a = -inf reached)
  const double tol = 2.4e-7 * (double)fminf(fabsf(gf), 3.0e38f) + 1e-12 * (1.0 + fabs(pt));
  const bool pruned = a > r + tol;
  reached = a < r - tol;
  unsure = !pruned && !reached;
}

// Candidate test: the reference's predicate with the ghost probe behind ghost_band, for UNR
// candidates at once, written straight-line (selects, no per-candidate branches) so that the
// FP64 chains of the candidates overlap; the exact gate[] look-up and the ghost pass stay behind
// warp-level branches because they are rare.  The decision is taken on the SQUARED sum against
// the exact thresholds of the query (Query::t_lt / t_le, gridscan.cuh), which is the
// reference's `sqrt(s) < range` bit for bit; sout receives the squared sum of the admitting
// probe -- the caller takes the square root when it writes the row, off the scan's
// dependency chain.
template <int METRIC, int UNR>
__device__ __forceinline__ void eval_candidates(const Query &Q, const GhostBand &B, const bool *valid,
                                                const double *nx, const double *ny, const double *nt,
                                                const int32_t *id, const float *gf, const double *gate,
                                                bool *hit, double *sout) {
  double s[UNR];
  bool askgate[UNR];
#pragma unroll
  for (int u = 0; u < UNR; u++) s[u] = metric_sum<METRIC>(Q.x, Q.y, Q.t, nx[u], ny[u], nt[u]);
  bool anygate = false;
#pragma unroll
  for (int u = 0; u < UNR; u++) {
    const bool root = id[u] == 0;
    const bool in = valid[u] && s[u] < (root ? Q.t_le : Q.t_lt);   // kdtree.cpp:800-802, :733, :765
    const bool fast = METRIC == DRRT_METRIC_EUCLID || root || !(Q.t - nt[u] > Q.r);  // walk_reaches, first clause
    hit[u] = in && fast;
    askgate[u] = in && !fast;
    anygate |= askgate[u];
    sout[u] = s[u];
  }
  if (METRIC != DRRT_METRIC_EUCLID && __any_sync(0xffffffffu, anygate)) {
    // second clause from the float copy (no memory access); the exact look-up only inside its band
    bool unsure[UNR], anyunsure = false;
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      bool reached;
      gate_from_copy(Q.t, Q.r, gf[u], reached, unsure[u]);
      unsure[u] = unsure[u] && askgate[u];
      if (askgate[u]) hit[u] = reached;
      anyunsure |= unsure[u];
    }
    if (__any_sync(0xffffffffu, anyunsure)) {
#pragma unroll
      for (int u = 0; u < UNR; u++)
        if (unsure[u]) hit[u] = !(Q.t - gate[id[u]] > Q.r);
    }
  }
  if (Q.has_ghost) {  // kdtree.cpp:807-819; uniform
    bool needg[UNR], any = false;
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      needg[u] = valid[u] && !hit[u] && (s[u] < B.s_band || !(nt[u] >= B.nlo && nt[u] <= B.nhi));
      any |= needg[u];
    }
    if (__any_sync(0xffffffffu, any)) {
      double s2[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++) s2[u] = metric_sum<METRIC>(Q.x, Q.y, Q.gt, nx[u], ny[u], nt[u]);
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        const bool in = needg[u] && s2[u] < Q.t_lt;   // the ghost walk admits with `<` only (:733, :765)
        bool ok = in;
        if (METRIC != DRRT_METRIC_EUCLID && in && (Q.gt - nt[u] > Q.r)) {
          bool reached, unsure;
          gate_from_copy(Q.gt, Q.r, gf[u], reached, unsure);
          ok = unsure ? !(Q.gt - gate[id[u]] > Q.r) : reached;
        }
        if (ok) { hit[u] = true; sout[u] = s2[u]; }
      }
    }
  }
}

template <bool BLOCK>
__device__ __forceinline__ void group_sync() {
  if (BLOCK) __syncthreads(); else __syncwarp();
}

// all-ascending bitonic network over n (any n) keys; `nthreads` threads, rank `tid`
template <bool BLOCK>
__device__ void sort_keys_shared(unsigned long long *keys, int n, int tid, int nthreads) {
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  int lk = 1;
  for (int k = 2; k <= np2; k <<= 1, lk++) {
    const int half = k >> 1;
    for (int p = tid; p < (np2 >> 1); p += nthreads) {
      int blk = p >> (lk - 1), o = p & (half - 1);
      int i = blk * k + o, j = blk * k + k - 1 - o;
      if (j < n) {
        unsigned long long a = keys[i], b = keys[j];
        if (a > b) { keys[i] = b; keys[j] = a; }
      }
    }
    group_sync<BLOCK>();
    for (int jj = half >> 1; jj > 0; jj >>= 1) {
      for (int p = tid; p < (np2 >> 1); p += nthreads) {
        int i = ((p & ~(jj - 1)) << 1) | (p & (jj - 1)), j = i + jj;
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
        int64_t i = ((p & ~(jj - 1)) << 1) | (p & (jj - 1)), j = i + jj;
        if (j < n && ids[i] > ids[j]) {
          int32_t a = ids[i]; ids[i] = ids[j]; ids[j] = a;
          double b = dist[i]; dist[i] = dist[j]; dist[j] = b;
        }
      }
      group_sync<BLOCK>();
    }
  }
}

// ---------------------------------------------------------------------------
// Warp-per-query kernel (small neighbourhoods, config 2b).  Every warp is on its own: it
// takes a slice of consecutive queries from a ticket, answers them one after the other and
// writes their rows back to back into the slice's share of the arena -- no barrier and no
// ordered allocation between warps.  Per query: columns in batches of 32 (scan_columns),
// hits ballot-compacted into the warp's staging area, ordered by a small counting sort on id
// buckets, written with coalesced stores.
#ifndef DRRT_WARP_MINB
#define DRRT_WARP_MINB 4
#endif
// Order by id as in the CTA kernel, per warp: a staged hit bumps a BYTE counter of its id bucket
// (id >> shift, WARP_NB buckets over the ids of the store, four per shared-memory word) and keeps the
// old value as its arrival rank; one prefix sum over the 128 counter words (four per lane) gives every
// bucket's start, a hit's slot in `order` is start + rank, and the buckets that hold several hits are
// put in id order by the lane of their first arrival (an insertion sort of two or three entries).
// No min / max pass over the ids, no floating-point bucket scale, no per-hit loop over bucket mates:
// 604 -> ~260 warp instructions per query of config 2b (of 2 520).  Ids bunched in one bucket
// (WARP_CROWDED arrivals) send the row to the bitonic network.
static constexpr int WARP_NB = 512, WARP_NW = WARP_NB / 4, WARP_CROWDED = 24;
static_assert(WARP_NW == 4 * 32, "four counter words per lane");
static_assert(WARP_STAGE <= 256 && WARP_CROWDED < 256, "byte counters: a crowded bucket is seen before its counter wraps");

template <int METRIC>
__global__ void __launch_bounds__(WARP_MODE_WARPS * 32, DRRT_WARP_MINB) range_warp_kernel(RangeArgs A) {
  constexpr int WARPS = WARP_MODE_WARPS, STAGE = WARP_STAGE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int2 s_runl[WARPS][RUNS_PER_BATCH + 1];
  __shared__ __align__(16) unsigned s_hist[WARPS][WARP_NW];
  __shared__ __align__(16) uint16_t s_wstart[WARPS][WARP_NW];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw) + w * STAGE;  // (id << 32) | arrival rank
  const int32_t *key_id = reinterpret_cast<const int32_t *>(keys) + 1;                     // id of slot s: key_id[2 * s]
  double *sdist = reinterpret_cast<double *>(smem_raw + (size_t)WARPS * STAGE * 8) + w * STAGE;
  uint16_t *order = reinterpret_cast<uint16_t *>(smem_raw + (size_t)WARPS * STAGE * 16) + w * STAGE;
  unsigned *hist = s_hist[w];
  uint16_t *wstart = s_wstart[w];
  const GridView &V = A.V;
  const unsigned ltmask = (1u << lane) - 1u;
  // id bucket = id >> shift, at most WARP_NB buckets over the ids of the store
  int shift = 0;
  {
    unsigned top = (unsigned)(V.n > 0 ? V.n - 1 : 0);
    int bits = 32 - __clz(top | 1u);
    shift = max(0, bits - 9);
  }
  // one radius for the whole batch (the usual call): its exact sqrt thresholds once per CTA
  __shared__ double s_thr[2];
  if (threadIdx.x == 0 && !A.radius) sqrt_thresholds(A.radius_all, s_thr[0], s_thr[1]);
  __syncthreads();

  for (;;) {
    long long ck = 0;
    if (lane == 0) ck = (long long)atomicAdd(A.ticket, 1);
    ck = __shfl_sync(0xffffffffu, ck, 0);
    if (ck >= A.nchunks) break;
    const int64_t base = A.chunk_base ? A.chunk_base[ck] : ck * A.chunk_cap;
    const int64_t cap = A.chunk_base ? A.chunk_base[ck + 1] - base : A.chunk_cap;
    const int64_t q0 = ck * A.chunk_len, q1 = min(A.nq, q0 + (int64_t)A.chunk_len);
    int64_t run = 0;
    bool over = false;
    for (int64_t qi = q0; qi < q1; qi++) {
      Query Q;
      setup_query<METRIC, false>(V, A.q[3 * qi], A.q[3 * qi + 1], A.q[3 * qi + 2],
                                 A.radius ? A.radius[qi] : A.radius_all, Q);
      if (A.radius) sqrt_thresholds(Q.r, Q.t_lt, Q.t_le);
      else { Q.t_lt = s_thr[0]; Q.t_le = s_thr[1]; }
      const GhostBand GB = ghost_band<METRIC>(Q);
      reinterpret_cast<uint4 *>(hist)[lane] = make_uint4(0, 0, 0, 0);
      __syncwarp();
      int32_t *gids = nullptr;
      double *gdist = nullptr;
      int cnt = 0;  // hits of this query so far: warp-uniform, in a register (every lane sees the ballot)
      bool crowd = false;
      auto visit = [&](bool valid, double nx, double ny, double nt, int32_t id, float gf) {
        bool hit;
        double d;  // squared sum of the admitting probe; the key is its square root (taken at the write)
        eval_candidates<METRIC, 1>(Q, GB, &valid, &nx, &ny, &nt, &id, &gf, V.gate, &hit, &d);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (!m) return;
        const int pos = cnt;
        cnt += __popc(m);
        if (hit) {
          const int slot = pos + __popc(m & ltmask);
          if (gids) {
            gids[slot] = id;
            gdist[slot] = sqrt(d);
          } else if (slot < STAGE) {
            const unsigned b = (unsigned)id >> shift;
            const int sh = (b & 3u) * 8;
            const unsigned rank = (atomicAdd(&hist[b >> 2], 1u << sh) >> sh) & 0xffu;
            keys[slot] = ((unsigned long long)(unsigned)id << 32) | rank;
            sdist[slot] = d;
            crowd |= rank >= (unsigned)WARP_CROWDED;
          }
        }
      };
      auto scan_all = [&]() {
        if (V.n_indexed > 0)
          scan_columns(V, Q, 0, 1, s_runl[w], lane, visit);
        scan_tail(V, 0, 1, lane, visit);
      };
      scan_all();
      __syncwarp();
      const int n = cnt;
      const bool staged = n > 0 && n <= STAGE;
      const bool fits = run + n <= cap;
      if (lane == 0) {
        A.offsets[qi] = base + run;
        A.counts[qi] = n;
        if (qi == A.nq - 1) A.offsets[A.nq] = base + run + n;
      }
      int32_t *oid = A.ids + base + run;
      double *od = A.dist + base + run;
      if (!fits) {
        over = true;  // counted, not written: the host reruns with exact slices
      } else if (staged && !__any_sync(0xffffffffu, crowd)) {
        // bucket starts: prefix sum over the byte counters, four words per lane
        const uint4 h4 = reinterpret_cast<const uint4 *>(hist)[lane];
        const int c0 = (int)__dp4a(h4.x, 0x01010101u, 0u), c1 = (int)__dp4a(h4.y, 0x01010101u, 0u);
        const int c2 = (int)__dp4a(h4.z, 0x01010101u, 0u), c3 = (int)__dp4a(h4.w, 0x01010101u, 0u);
        const int sum = (c0 + c1) + (c2 + c3);
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const unsigned e0 = (unsigned)(inc - sum), e1 = e0 + c0, e2 = e1 + c1, e3 = e2 + c2;
        reinterpret_cast<uint2 *>(wstart)[lane] = make_uint2(e0 | (e1 << 16), e2 | (e3 << 16));
        __syncwarp();
        // slot in `order` = bucket start + arrival rank
        for (int i = lane; i < n; i += 32) {
          const unsigned long long k = keys[i];
          const unsigned b = (unsigned)(k >> 32) >> shift;
          const unsigned wd = hist[b >> 2];
          const int sh = (b & 3u) * 8;
          const int s0 = wstart[b >> 2] + (int)__dp4a(wd & ((1u << sh) - 1u), 0x01010101u, 0u);
          order[s0 + (int)(unsigned)(k & 0xffu)] = (uint16_t)i;
        }
        __syncwarp();
        // buckets holding several hits: the lane of the first arrival puts the bucket's entries in id order
        for (int i0 = 0; i0 < n; i0 += 32) {
          const int i = i0 + lane;
          int s0 = 0, c = 0;
          if (i < n) {
            const unsigned long long k = keys[i];
            if ((k & 0xffu) == 0u) {
              const unsigned b = (unsigned)(k >> 32) >> shift;
              const unsigned wd = hist[b >> 2];
              const int sh = (b & 3u) * 8;
              c = (int)((wd >> sh) & 0xffu);
              s0 = wstart[b >> 2] + (int)__dp4a(wd & ((1u << sh) - 1u), 0x01010101u, 0u);
            }
          }
          for (int a2 = 1; a2 < c; a2++) {
            const uint16_t slot = order[s0 + a2];
            const int32_t kid = key_id[2 * slot];
            int j = a2 - 1;
            while (j >= 0) {
              const uint16_t sj = order[s0 + j];
              if (key_id[2 * sj] <= kid) break;
              order[s0 + j + 1] = sj;
              j--;
            }
            order[s0 + j + 1] = slot;
          }
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
          const int slot = order[i];
          oid[i] = key_id[2 * slot];
          od[i] = sqrt(sdist[slot]);
        }
      } else if (staged) {
        // ids bunched in a few buckets: network sort of (id, slot) keys
        for (int i = lane; i < n; i += 32) keys[i] = (keys[i] & 0xffffffff00000000ull) | (unsigned)i;
        __syncwarp();
        sort_keys_shared<false>(keys, n, lane, 32);
        for (int p = lane; p < n; p += 32) {
          const unsigned long long k = keys[p];
          oid[p] = (int32_t)(k >> 32);
          od[p] = sqrt(sdist[(unsigned)(k & 0xffffu)]);
        }
      } else if (n > 0) {
        // more hits than the staging area: second pass straight into the arena
        __syncwarp();
        cnt = 0;
        gids = oid;
        gdist = od;
        scan_all();
        __syncwarp();
        sort_pairs_global<false>(oid, od, n, lane, 32);
      }
      run += n;
      __syncwarp();
    }
    if (lane == 0) {
      A.chunk_total[ck] = run;
      atomicAdd(&A.totals[0], (unsigned long long)run);
      if (over) atomicAdd(&A.totals[1], 1ull);
    }
  }
}

// ---------------------------------------------------------------------------
// CTA-per-query kernel (large neighbourhoods, config 2a).
//
// The kernel is issue- and latency-bound, not HBM-bound (profiles/README.md), so
// everything here is about instructions and barriers per query:
//  * persistent CTAs take CHUNKS of CTA_CHUNK consecutive queries from a ticket
//    and write the chunk's rows back to back into the chunk's own slice of the
//    result arena -- no ordered allocation across CTAs, hence no CTA ever waits
//    for a predecessor (the decoupled look-back of the first version cost ~15 % here).
//    Rows are addressed by offsets[q] / counts[q]; slices are sized from the
//    expected neighbourhood and leave unused slots between chunks.  If a chunk
//    overflows its slice the host reruns with exact slice sizes;
//  * the id sort is fused into the scan: a staged hit bumps a BYTE counter of
//    its id bucket (id >> shift, 16384 buckets, four per shared-memory word) and
//    keeps the old value as its arrival rank; one prefix sum over the 4096 words
//    gives every bucket's start, a hit's position is start + rank, and the few
//    buckets holding several hits (~20 % of the hits) are insertion-sorted;
//  * the ghost probe (kdtree.cpp:807-819) is evaluated only when it can matter
//    (ghost_band below);
//  * two candidates per lane per trip, one shared counter update for both.
static constexpr int CTA_WARPS = 16;
static constexpr int CTA_T = CTA_WARPS * 32;
static constexpr int CTA_STAGE = 4608;  // hits staged per query
static constexpr int CTA_NB = 16384;    // id buckets
static constexpr int CTA_NW = CTA_NB / 4;
#ifndef DRRT_CTA_UNR
#define DRRT_CTA_UNR 2
#endif
static constexpr int CTA_UNR = DRRT_CTA_UNR;
static constexpr int CTA_CROWDED = 32;  // larger buckets are not insertion-sorted (bytes must not wrap)
#ifndef DRRT_CTA_CHUNK
#define DRRT_CTA_CHUNK 8
#endif
static constexpr int CTA_CHUNK = DRRT_CTA_CHUNK;  // consecutive queries per arena slice (fewer when the batch is small)
// dynamic shared memory layout (bytes)
static constexpr int CTA_OFF_DIST = 0;
static constexpr int CTA_OFF_ID = CTA_OFF_DIST + CTA_STAGE * 8;
static constexpr int CTA_OFF_HIST = CTA_OFF_ID + CTA_STAGE * 4;
static constexpr int CTA_OFF_WSTART = CTA_OFF_HIST + CTA_NW * 4;
static constexpr int CTA_OFF_TMP = CTA_OFF_WSTART + CTA_NW * 2;
static constexpr int CTA_OFF_RANK = CTA_OFF_TMP + CTA_STAGE * 2;
static constexpr int CTA_OFF_RUNS = CTA_OFF_RANK + CTA_STAGE;
static constexpr int CTA_RUNS_BYTES = (2 * CTA_T + 1) * 8 + 8;
static constexpr int CTA_CLIST_BYTES = (CTA_STAGE / 2) * 4;  // shared buckets: at most STAGE / 2
static constexpr int CTA_SMEM = CTA_OFF_RUNS + (CTA_RUNS_BYTES > CTA_CLIST_BYTES ? CTA_RUNS_BYTES : CTA_CLIST_BYTES);
static_assert(CTA_OFF_HIST + CTA_STAGE * 8 <= CTA_SMEM, "fallback sort keys alias the sort scratch");
static_assert(CTA_OFF_RUNS % 8 == 0 && CTA_OFF_HIST % 16 == 0 && CTA_OFF_WSTART % 16 == 0, "alignment");
static_assert(CTA_STAGE % CTA_T == 0, "stage is a whole number of CTA trips");

template <int METRIC>
__global__ void __launch_bounds__(CTA_T, 2) range_cta_kernel(RangeArgs A) {
  constexpr int T = CTA_T, WARPS = CTA_WARPS, STAGE = CTA_STAGE, UNR = CTA_UNR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sdist = reinterpret_cast<double *>(smem_raw + CTA_OFF_DIST);
  int32_t *sid = reinterpret_cast<int32_t *>(smem_raw + CTA_OFF_ID);
  unsigned *hist = reinterpret_cast<unsigned *>(smem_raw + CTA_OFF_HIST);
  uint16_t *wstart = reinterpret_cast<uint16_t *>(smem_raw + CTA_OFF_WSTART);
  uint16_t *tmp = reinterpret_cast<uint16_t *>(smem_raw + CTA_OFF_TMP);
  uint8_t *srank = reinterpret_cast<uint8_t *>(smem_raw + CTA_OFF_RANK);
  int2 *runs = reinterpret_cast<int2 *>(smem_raw + CTA_OFF_RUNS);
  unsigned *clist = reinterpret_cast<unsigned *>(smem_raw + CTA_OFF_RUNS);  // after the scan
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw + CTA_OFF_HIST);
  __shared__ long long s_chunk, s_chunk_next;
  __shared__ int s_count, s_crowded, s_nlist;
  __shared__ long long s_tmp[WARPS];
  __shared__ int s_wsum[WARPS];
  __shared__ Query s_query;
  // the probes of a whole slice, set up by CTA_CHUNK threads at once at the top of the slice (ghost,
  // cell ranges, the exact sqrt thresholds of the radius): nothing of it sits on a query's critical path
  __shared__ Query s_slice[CTA_CHUNK];

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const GridView &V = A.V;
  const unsigned ltmask = (1u << lane) - 1u;
  // id bucket = id >> shift, at most CTA_NB buckets over the ids of the store
  int shift = 0;
  {
    unsigned top = (unsigned)(V.n > 0 ? V.n - 1 : 0);
    int bits = 32 - __clz(top | 1u);
    shift = max(0, bits - 14);
  }

  if (tid == 0) s_chunk_next = (long long)atomicAdd(A.ticket, 1);
  for (;;) {
    __syncthreads();  // everyone is done with the previous chunk (and its s_chunk)
    if (tid == 0) s_chunk = s_chunk_next;  // taken while the previous slice's last query was answered
    __syncthreads();
    const int64_t ck = s_chunk;
    if (ck >= A.nchunks) break;
    const int64_t base = A.chunk_base ? A.chunk_base[ck] : ck * A.chunk_cap;
    const int64_t cap = A.chunk_base ? A.chunk_base[ck + 1] - base : A.chunk_cap;
    const int64_t q0 = ck * A.chunk_len, q1 = min(A.nq, q0 + (int64_t)A.chunk_len);
    int64_t run = 0;    // hits of the chunk's earlier rows
    bool over = false;  // some row did not fit the slice
    // the slice's queries in one round of loads (the barrier at the top of the loop ordered the
    // previous slice's reads of s_q)
    if (tid < CTA_CHUNK && q0 + tid < q1) {
      const int64_t qk = q0 + tid;
      Query Qk;
      setup_query<METRIC, true>(V, A.q[3 * qk], A.q[3 * qk + 1], A.q[3 * qk + 2],
                                A.radius ? A.radius[qk] : A.radius_all, Qk);
      s_slice[tid] = Qk;
    }
    __syncthreads();

    for (int64_t qi = q0; qi < q1; qi++) {
      DRRT_PHASE_DECL
      if (tid == 0) { s_count = 0; s_crowded = 0; s_nlist = 0; }
      // the next slice's ticket is requested one query ahead: the round trip of the global atomic
      // is off the slice boundary, and a CTA never holds more than one query's worth of work back
      if (tid == 32 && qi == q1 - 1) s_chunk_next = (long long)atomicAdd(A.ticket, 1);
      for (int i = tid; i < CTA_NW / 4; i += T) reinterpret_cast<uint4 *>(hist)[i] = make_uint4(0, 0, 0, 0);
      if (w == 0) {  // the query's probe to its fixed place (a word per lane; no indexed reads in the scan)
        static_assert(sizeof(Query) % 4 == 0 && sizeof(Query) <= 128, "one 32-bit word per lane");
        if (lane < (int)(sizeof(Query) / 4))
          reinterpret_cast<unsigned *>(&s_query)[lane] = reinterpret_cast<const unsigned *>(&s_slice[qi - q0])[lane];
      }
      __syncthreads();
      DRRT_PHASE(0);
      const Query Q = s_query;
      const GhostBand GB = ghost_band<METRIC>(Q);
      int32_t *gids = nullptr;
      double *gdist = nullptr;

      auto visit_impl = [&](auto direct_tag, const bool *valid, const double *nx, const double *ny,
                            const double *nt, const int32_t *id, const float *gf) {
        constexpr bool DIRECT = decltype(direct_tag)::value;
        bool hit[UNR];
        double d[UNR];
        unsigned m[UNR];
        int tot = 0;
        eval_candidates<METRIC, UNR>(Q, GB, valid, nx, ny, nt, id, gf, V.gate, hit, d);
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          m[u] = __ballot_sync(0xffffffffu, hit[u]);
          tot += __popc(m[u]);
        }
        if (tot == 0) return;
        int pos = 0;
        if (lane == 0) pos = atomicAdd(&s_count, tot);
        pos = __shfl_sync(0xffffffffu, pos, 0);
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          if (hit[u]) {
            const int slot = pos + __popc(m[u] & ltmask);
            if (DIRECT) {
              gids[slot] = id[u];
              gdist[slot] = sqrt(d[u]);
            } else if (slot < STAGE) {
              sid[slot] = id[u];
              sdist[slot] = d[u];
              const unsigned b = (unsigned)id[u] >> shift;
              const int sh = (b & 3u) * 8;
              const unsigned old = atomicAdd(&hist[b >> 2], 1u << sh);
              const unsigned rank = (old >> sh) & 0xffu;
              srank[slot] = (uint8_t)rank;
              if (rank >= CTA_CROWDED) s_crowded = 1;
            }
          }
          pos += __popc(m[u]);
        }
      };
      auto visit = [&](const bool *valid, const double *nx, const double *ny, const double *nt,
                       const int32_t *id, const float *gf) {
        visit_impl(std::false_type{}, valid, nx, ny, nt, id, gf);
      };
      auto visit_direct = [&](const bool *valid, const double *nx, const double *ny,
                              const double *nt, const int32_t *id, const float *gf) {
        visit_impl(std::true_type{}, valid, nx, ny, nt, id, gf);
      };

      if (V.n_indexed > 0) scan_columns_cta<WARPS, UNR>(V, Q, runs, s_tmp, visit DRRT_PHASE_ARG);
      scan_tail_cta<WARPS, UNR>(V, visit);
      __syncthreads();
      DRRT_PHASE(5);

      const int n = s_count;
      const bool staged = n > 0 && n <= STAGE;
      const bool crowded = s_crowded != 0;
      const bool fits = run + n <= cap;
      if (tid == 0) {
        A.offsets[qi] = base + run;
        A.counts[qi] = n;
        if (qi == A.nq - 1) A.offsets[A.nq] = base + run + n;
      }
      int32_t *oid = A.ids + base + run;
      double *od = A.dist + base + run;
      if (!fits) {
        over = true;  // counted, not written: the host reruns with exact slices
      } else if (staged && !crowded) {
        // bucket starts: prefix sum over the byte counters, eight words per thread
        uint4 a = reinterpret_cast<const uint4 *>(hist)[2 * tid];
        uint4 b4 = reinterpret_cast<const uint4 *>(hist)[2 * tid + 1];
        unsigned wv[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
        int cnt[8], sum = 0;
#pragma unroll
        for (int e = 0; e < 8; e++) { cnt[e] = (int)__dp4a(wv[e], 0x01010101u, 0u); sum += cnt[e]; }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (lane == 31) s_wsum[w] = inc;
        __syncthreads();
        int wsv = (lane < WARPS) ? s_wsum[lane] : 0, wsi = wsv;
#pragma unroll
        for (int o = 1; o < WARPS; o <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, wsi, o);
          if (lane >= o) wsi += t;
        }
        int excl = inc - sum + __shfl_sync(0xffffffffu, wsi - wsv, w);
        unsigned pk[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          unsigned lo16 = (unsigned)excl; excl += cnt[e];
          unsigned hi16 = (unsigned)excl; excl += cnt[e + 1];
          pk[e >> 1] = lo16 | (hi16 << 16);
        }
        reinterpret_cast<uint4 *>(wstart)[tid] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        __syncthreads();
        DRRT_PHASE(6);
        // position = bucket start + arrival rank; buckets holding several hits are listed
        // (by their first arrival) and put in id order below
        // four staged hits per thread and trip, straight-line: the four look-up chains (id -> counter
        // word and word start -> position) overlap instead of running one after the other behind
        // `if (i < n)`, and the shared buckets found in a trip are appended with ONE bump of the list
        // counter per warp instead of one per hit row (a query's critical path held seven of those
        // atomic round trips)
        // rows of 32 hits a warp has in a batch (i0 + k T + its 32 lanes, k < 4): the batches are
        // instantiated for one to four live rows, so a warp never runs a row that lies wholly beyond n
        // (128 rows for config 2a's 111 in the fixed four-row form)
        auto batch = [&](auto nk_tag, int i0) {
          constexpr int NK = decltype(nk_tag)::value;
          unsigned ent[NK], m[NK];
          bool listed[NK];
          int tot = 0;
#pragma unroll
          for (int k = 0; k < NK; k++) {
            const int i = i0 + k * T + tid;
            const bool ok = i < n;
            const int ii = ok ? i : 0;  // some staged hit (n > 0): the loads need no guard
            const unsigned b = (unsigned)sid[ii] >> shift;
            const unsigned wd = hist[b >> 2];
            const int sh = (b & 3u) * 8;
            const int s0 = wstart[b >> 2] + (int)__dp4a(wd & ((1u << sh) - 1u), 0x01010101u, 0u);
            const unsigned r = srank[ii];
            if (ok) tmp[s0 + r] = (uint16_t)i;
            const unsigned c = (wd >> sh) & 0xffu;
            ent[k] = (unsigned)s0 | (c << 16);
            listed[k] = ok && c > 1 && r == 0;
          }
#pragma unroll
          for (int k = 0; k < NK; k++) {
            m[k] = __ballot_sync(0xffffffffu, listed[k]);
            tot += __popc(m[k]);
          }
          if (tot) {
            int lb = 0;
            if (lane == 0) lb = atomicAdd(&s_nlist, tot);
            lb = __shfl_sync(0xffffffffu, lb, 0);
#pragma unroll
            for (int k = 0; k < NK; k++) {
              if (listed[k]) clist[lb + __popc(m[k] & ltmask)] = ent[k];
              lb += __popc(m[k]);
            }
          }
        };
        for (int i0 = 0; i0 + w * 32 < n; i0 += 4 * T) {
          const int live = (n - i0 - w * 32 + T - 1) / T;  // rows of this warp that hold a hit (>= 1)
          if (live >= 4) batch(std::integral_constant<int, 4>{}, i0);
          else if (live == 3) batch(std::integral_constant<int, 3>{}, i0);
          else if (live == 2) batch(std::integral_constant<int, 2>{}, i0);
          else batch(std::integral_constant<int, 1>{}, i0);
        }
        __syncthreads();
        DRRT_PHASE(7);
        const int nl = s_nlist;
        for (int e = tid; e < nl; e += T) {
          const unsigned ent = clist[e];
          const int s0 = (int)(ent & 0xffffu), c = (int)(ent >> 16);
          for (int a2 = 1; a2 < c; a2++) {
            const uint16_t slot = tmp[s0 + a2];
            const int32_t kid = sid[slot];
            int j = a2 - 1;
            while (j >= 0) {
              const uint16_t sj = tmp[s0 + j];
              if (sid[sj] <= kid) break;
              tmp[s0 + j + 1] = sj;
              j--;
            }
            tmp[s0 + j + 1] = slot;
          }
        }
        __syncthreads();
        DRRT_PHASE(8);
        // the staged values are squared sums: the keys' square roots are taken here, on
        // independent elements, instead of inside the scan's dependency chain
        for (int p = tid; p < n; p += T) {
          const int slot = tmp[p];
          oid[p] = sid[slot];
          od[p] = sqrt(sdist[slot]);
        }
      } else if (staged) {
        // ids bunched in a few buckets: network sort of (id, slot) keys
        for (int i = tid; i < n; i += T) keys[i] = ((unsigned long long)(unsigned)sid[i] << 32) | (unsigned)i;
        __syncthreads();
        sort_keys_shared<true>(keys, n, tid, T);
        for (int p = tid; p < n; p += T) {
          const unsigned long long k = keys[p];
          oid[p] = (int32_t)(k >> 32);
          od[p] = sqrt(sdist[(unsigned)(k & 0xffffu)]);
        }
      } else if (n > 0) {
        // more hits than the staging area: second pass straight into the arena
        __syncthreads();
        if (tid == 0) s_count = 0;
        __syncthreads();
        gids = oid;
        gdist = od;
        if (V.n_indexed > 0) scan_columns_cta<WARPS, UNR>(V, Q, runs, s_tmp, visit_direct DRRT_PHASE_ARG);
        scan_tail_cta<WARPS, UNR>(V, visit_direct);
        __syncthreads();
        sort_pairs_global<true>(oid, od, n, tid, T);
      }
      DRRT_PHASE(9);   // thread 0's own share of the row write
      run += n;
      __syncthreads();  // staging and scratch are free for the next query
      DRRT_PHASE(10);
    }
    if (tid == 0) {
      A.chunk_total[ck] = run;
      atomicAdd(&A.totals[0], (unsigned long long)run);
      if (over) atomicAdd(&A.totals[1], 1ull);
    }
  }
}

static constexpr int WARP_SMEM = WARP_STAGE * WARP_MODE_WARPS * 18;

// resident CTAs of the persistent kernels (occupancy x SMs).  Cached in the Store: the opt-in to
// > 48 KB of dynamic shared memory (cudaFuncSetAttribute) is a PER-DEVICE attribute and the SM
// count belongs to the store's device, so a process-wide cache would launch the CTA kernel
// without the opt-in on the second device of a multi-GPU process (and race between handles).
template <int METRIC>
static int kernel_grid_of(bool block_mode, int *per_sm) {
  if (block_mode) {
    DRRT_CUDA(cudaFuncSetAttribute(range_cta_kernel<METRIC>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, CTA_SMEM));
    DRRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, range_cta_kernel<METRIC>, CTA_T,
                                                            CTA_SMEM));
  } else {
    DRRT_CUDA(cudaFuncSetAttribute(range_warp_kernel<METRIC>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, WARP_SMEM));
    DRRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, range_warp_kernel<METRIC>,
                                                            WARP_MODE_WARPS * 32, WARP_SMEM));
  }
  return DRRT_OK;
}

static int kernel_grid(Store *S, bool block_mode, int *grid) {
  const int ki = block_mode ? 1 : 0;
  if (!S->range_grid[ki]) {
    int sms = 0, v = 0;
    DRRT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, S->device));
    if (S->metric == DRRT_METRIC_EUCLID) DRRT_CHECK(kernel_grid_of<DRRT_METRIC_EUCLID>(block_mode, &v));
    else DRRT_CHECK(kernel_grid_of<DRRT_METRIC_R2S1_WRAP>(block_mode, &v));
    if (v < 1 || sms < 1) return set_error(DRRT_ERR_CUDA, "range: kernel does not fit an SM");
    S->range_grid[ki] = v * sms;
  }
  *grid = S->range_grid[ki];
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

// ---------------------------------------------------------------------------
// packing: rows of a sliced result -> plain CSR (exclusive prefix sum of counts)

// single CTA; out[0..n] = base + exclusive prefix sums of counts (out[n] = base + total)
__global__ void scan_counts_kernel(const int32_t *counts, int64_t n, int64_t *out, long long base) {
  constexpr int T = 1024;
  __shared__ long long s_w[32];
  __shared__ long long s_carry;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s_carry = base;
  __syncthreads();
  for (int64_t b = 0; b < n; b += T) {
    const int64_t i = b + tid;
    long long v = i < n ? (long long)counts[i] : 0, inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    long long wv = s_w[lane], wi = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    const long long carry = s_carry;
    const long long excl = carry + __shfl_sync(0xffffffffu, wi - wv, w) + inc - v;
    if (i < n) out[i] = excl;
    __syncthreads();
    if (tid == T - 1) s_carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) out[n] = s_carry;
}

// DT: double (the reference's JList key, bit-exact), float (narrow rows: the north star asks
// for distances within 1e-5 relative) or void (ids only)
template <typename DT>
__global__ void pack_rows_kernel(int64_t nq, const int64_t *src_off, const int32_t *counts,
                                 const int64_t *dst_off, int64_t dst_base, const int32_t *ids,
                                 const double *dist, int32_t *ids_out, DT *dist_out) {
  for (int64_t q = blockIdx.x; q < nq; q += gridDim.x) {
    const int64_t s = src_off[q], d = dst_off[q] - dst_base;
    const int c = counts[q];
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
      ids_out[d + i] = ids[s + i];
      if constexpr (!std::is_void<DT>::value) dist_out[d + i] = (DT)dist[s + i];
    }
  }
}

static void fill_result(Store *S, drrt_range_result *out) {
  if (!out) return;
  out->nq = S->last_nq;
  out->total = S->last_total;
  out->offsets_dev = S->last_packed_in_alt ? S->r_offsets2.p : S->r_offsets.p;
  out->ids_dev = S->last_packed_in_alt ? S->r_ids2.p : S->r_ids.p;
  out->dist_dev = S->last_packed_in_alt ? S->r_dist2.p : S->r_dist.p;
  out->counts_dev = S->r_counts.p;
  out->packed = (S->last_packed || S->last_packed_in_alt) ? 1 : 0;
  out->extent = out->packed ? S->last_total : S->last_extent;
}

int range_pack_to(Store *S, cudaStream_t st, int64_t *off_out, int32_t *ids_out, void *dist_out,
                  int dist_kind, int64_t base) {
  if (S->last_nq < 0) return set_error(DRRT_ERR_STATE, "pack: no range batch on this handle");
  const int64_t nq = S->last_nq;
  if (nq == 0) return DRRT_OK;
  scan_counts_kernel<<<1, 1024, 0, st>>>(S->r_counts.p, nq, off_out, (long long)base);
  DRRT_LAUNCHED();
  if (S->last_total > 0) {
    const int32_t *ids = S->last_packed_in_alt ? S->r_ids2.p : S->r_ids.p;
    const double *dist = S->last_packed_in_alt ? S->r_dist2.p : S->r_dist.p;
    const int64_t *off = S->last_packed_in_alt ? S->r_offsets2.p : S->r_offsets.p;
    const unsigned grid = (unsigned)std::min<int64_t>(nq, 148 * 16);
    if (dist_kind == DIST_F64 && dist_out)
      pack_rows_kernel<double><<<grid, 256, 0, st>>>(nq, off, S->r_counts.p, off_out, base, ids, dist,
                                                     ids_out, (double *)dist_out);
    else if (dist_kind == DIST_F32 && dist_out)
      pack_rows_kernel<float><<<grid, 256, 0, st>>>(nq, off, S->r_counts.p, off_out, base, ids, dist,
                                                    ids_out, (float *)dist_out);
    else
      pack_rows_kernel<void><<<grid, 256, 0, st>>>(nq, off, S->r_counts.p, off_out, base, ids, dist,
                                                   ids_out, (void *)nullptr);
    DRRT_LAUNCHED();
  }
  DRRT_CUDA(cudaGetLastError());
  return DRRT_OK;
}

// Makes the last result a plain CSR (rows back to back in query order).
int range_pack_dev(Store *S, cudaStream_t st, drrt_range_result *out) {
  if (S->last_nq < 0) return set_error(DRRT_ERR_STATE, "pack: no range batch on this handle");
  if (!S->last_packed && !S->last_packed_in_alt && S->last_nq > 0) {
    const int64_t nq = S->last_nq;
    DRRT_CHECK(S->r_offsets2.reserve(nq + 1, st));
    DRRT_CHECK(S->r_ids2.reserve(S->last_total + 1, st));
    DRRT_CHECK(S->r_dist2.reserve(S->last_total + 1, st));
    DRRT_CHECK(range_pack_to(S, st, S->r_offsets2.p, S->r_ids2.p, S->r_dist2.p, DIST_F64, 0));
    S->last_packed_in_alt = true;
  }
  fill_result(S, out);
  return DRRT_OK;
}

// The planner's own call -- ONE query per iteration (src/drrt.cpp:213-216, src/mainloop.cpp:204-210)
// -- answered with one kernel launch and one stream wait and NO copy call: the query is read from,
// and the row written straight into, a page-locked host buffer mapped into the device's address
// space (S->h_map; stores of the sorted row cross PCIe as they are issued).  Returns
// RANGE_NOT_HANDLED when the row may not fit the mapped area (or overflowed it): the caller then
// takes the general path.
int range_one_mapped(Store *S, const double *q3, double radius, int64_t *total, const int32_t **ids,
                     const double **dist) {
  cudaStream_t st = S->stream;
  if (S->n == 0) { *total = 0; *ids = nullptr; *dist = nullptr; return DRRT_OK; }
  DRRT_CHECK(ensure_index(S, 1, st));
  DRRT_CHECK(map_reserve(S));
  double expected = (double)S->n;
  if (S->n_indexed > 0) {
    const GridParams &g = S->gp;
    double vol = (g.nx / g.inv_hx) * (g.ny / g.inv_hy) * (g.nt / g.inv_ht);
    double ball = 4.18879 * fabs(radius) * fabs(radius) * fabs(radius);
    expected = std::min((double)S->n, (double)S->n * ball / std::max(vol, 1e-300));
  }
  if (!(expected == expected)) return RANGE_NOT_HANDLED;
  const bool block_mode = expected > WARP_STAGE / 2;
  int64_t rowcap = (int64_t)(expected * 1.25 + 8.0 * sqrt(expected) + 96.0);
  rowcap = std::min<int64_t>(rowcap, S->n);
  // rows beyond the kernels' staging areas are re-scanned and sorted in place in the arena: not over PCIe
  if (rowcap > MAP_ROW_ELEMS || rowcap > (block_mode ? CTA_STAGE : WARP_STAGE)) {
    if (block_mode || rowcap > MAP_ROW_ELEMS) return RANGE_NOT_HANDLED;
  }
  MappedRow *M = reinterpret_cast<MappedRow *>(S->h_map);
  M->q[0] = q3[0]; M->q[1] = q3[1]; M->q[2] = q3[2];
  M->totals[0] = 0; M->totals[1] = 0;
  RangeArgs A{};
  A.V = make_view(S);
  A.q = M->q; A.radius = nullptr; A.radius_all = radius; A.nq = 1;
  A.offsets = M->offsets; A.counts = M->counts; A.ticket = S->d_ticket;
  A.chunk_len = 1; A.nchunks = 1; A.chunk_cap = rowcap; A.chunk_base = nullptr;
  A.chunk_total = M->chunk_total; A.totals = M->totals;
  A.ids = M->ids; A.dist = M->dist;
  int grid = 0;
  DRRT_CHECK(kernel_grid(S, block_mode, &grid));  // (sets the shared-memory opt-in of this device)
  DRRT_CUDA(cudaMemsetAsync(S->d_ticket, 0, sizeof(int32_t), st));
  if (block_mode) {
    if (S->metric == DRRT_METRIC_EUCLID) range_cta_kernel<DRRT_METRIC_EUCLID><<<1, CTA_T, CTA_SMEM, st>>>(A);
    else range_cta_kernel<DRRT_METRIC_R2S1_WRAP><<<1, CTA_T, CTA_SMEM, st>>>(A);
  } else {
    if (S->metric == DRRT_METRIC_EUCLID)
      range_warp_kernel<DRRT_METRIC_EUCLID><<<1, WARP_MODE_WARPS * 32, WARP_SMEM, st>>>(A);
    else
      range_warp_kernel<DRRT_METRIC_R2S1_WRAP><<<1, WARP_MODE_WARPS * 32, WARP_SMEM, st>>>(A);
  }
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaGetLastError());
  DRRT_CUDA(cudaStreamSynchronize(st));
  S->last_nq = -1;  // nothing to fetch or pack afterwards
  if (M->totals[1] != 0 || (int64_t)M->totals[0] > rowcap) return RANGE_NOT_HANDLED;
  *total = (int64_t)M->totals[0];
  *ids = M->ids;
  *dist = M->dist;
  return DRRT_OK;
}

int range_batch_dev(Store *S, int64_t nq, const double *q_dev, const double *radius_dev,
                    double radius_all, cudaStream_t st, drrt_range_result *out) {
  if (nq < 0) return set_error(DRRT_ERR_INVALID, "range: nq < 0");
  if (nq > 0x7ffffff0) return set_error(DRRT_ERR_INVALID, "range: nq too large");
  S->last_nq = -1;
  S->last_packed = true;
  S->last_packed_in_alt = false;
  // the rows of an earlier fused Extend live in the arenas this call overwrites (or frees when
  // they grow): a late drrt_extend_fetch must fail with DRRT_ERR_STATE, not copy from them
  S->last_extend_total = -1;
  S->csr_off = nullptr;
  S->csr_ids = nullptr;
  S->csr_dist = nullptr;
  DRRT_CHECK(S->r_offsets.reserve(nq + 1, st));
  DRRT_CHECK(S->r_counts.reserve(nq + 1, st));
  if (nq == 0 || S->n == 0) {
    DRRT_CUDA(cudaMemsetAsync(S->r_offsets.p, 0, (nq + 1) * sizeof(int64_t), st));
    DRRT_CUDA(cudaMemsetAsync(S->r_counts.p, 0, (nq + 1) * sizeof(int32_t), st));
    S->last_nq = nq;
    S->last_total = 0;
    S->last_extent = 0;
    fill_result(S, out);
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
  const bool block_mode = expected > WARP_STAGE / 2;

  RangeArgs A{};
  A.V = make_view(S);
  A.q = q_dev; A.radius = radius_dev; A.radius_all = radius_all; A.nq = nq;
  A.offsets = S->r_offsets.p; A.counts = S->r_counts.p; A.ticket = S->d_ticket;

  // Both kernels write slices of chunk_len consecutive rows, sized from the expected
  // neighbourhood; a small batch takes shorter slices so that every resident CTA (warp) still
  // gets several of them.
  int grid = 0;
  DRRT_CHECK(kernel_grid(S, block_mode, &grid));
  const int64_t workers = block_mode ? grid : (int64_t)grid * WARP_MODE_WARPS;
  const int chunk_len = (int)std::max<int64_t>(1, std::min<int64_t>(CTA_CHUNK, nq / (4 * workers)));
  const int64_t nchunks = (nq + chunk_len - 1) / chunk_len;
  int64_t rowcap = (int64_t)(expected * 1.1 + 6.0 * sqrt(expected) + 64.0);
  rowcap = std::min<int64_t>(rowcap, S->n);
  DRRT_CHECK(S->r_chunk_total.reserve(nchunks + 1, st));
  A.chunk_len = chunk_len;
  A.nchunks = nchunks;
  A.chunk_cap = rowcap * chunk_len;
  A.chunk_base = nullptr;
  A.chunk_total = S->r_chunk_total.p;
  A.totals = (unsigned long long *)S->d_totals;
  grid = (int)std::min<int64_t>(grid, block_mode ? nchunks : (nchunks + WARP_MODE_WARPS - 1) / WARP_MODE_WARPS);
  int64_t extent = nchunks * A.chunk_cap;
  for (int attempt = 0; attempt < 2; attempt++) {
    DRRT_CHECK(S->r_ids.reserve(extent + 1, st));
    DRRT_CHECK(S->r_dist.reserve(extent + 1, st));
    A.ids = S->r_ids.p; A.dist = S->r_dist.p;
    DRRT_CUDA(cudaMemsetAsync(S->d_ticket, 0, 32, st));  // the ticket (+0) and both totals (+16, +24)
    if (block_mode) {
      if (S->metric == DRRT_METRIC_EUCLID)
        range_cta_kernel<DRRT_METRIC_EUCLID><<<grid, CTA_T, CTA_SMEM, st>>>(A);
      else
        range_cta_kernel<DRRT_METRIC_R2S1_WRAP><<<grid, CTA_T, CTA_SMEM, st>>>(A);
    } else {
      if (S->metric == DRRT_METRIC_EUCLID)
        range_warp_kernel<DRRT_METRIC_EUCLID><<<grid, WARP_MODE_WARPS * 32, WARP_SMEM, st>>>(A);
      else
        range_warp_kernel<DRRT_METRIC_R2S1_WRAP><<<grid, WARP_MODE_WARPS * 32, WARP_SMEM, st>>>(A);
    }
    DRRT_LAUNCHED();
    DRRT_CUDA(cudaGetLastError());
    DRRT_CUDA(cudaMemcpyAsync(S->h_scalar, S->d_totals, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    DRRT_CUDA(cudaStreamSynchronize(st));
    const int64_t total = S->h_scalar[0], overflowed = S->h_scalar[1];
    if (overflowed == 0) {
      S->last_nq = nq;
      S->last_total = total;
      S->last_extent = extent;
      // exact slices are back to back; a lone slice starts at 0 with its rows back to back
      S->last_packed = A.chunk_base != nullptr || nchunks == 1;
      fill_result(S, out);
      return DRRT_OK;
    }
    if (attempt == 1) break;
    // some slice was too small: the counts are exact, rerun with exact slices
    std::vector<int64_t> tot((size_t)nchunks), cb((size_t)nchunks + 1);
    DRRT_CUDA(cudaMemcpyAsync(tot.data(), S->r_chunk_total.p, nchunks * sizeof(int64_t),
                              cudaMemcpyDeviceToHost, st));
    DRRT_CUDA(cudaStreamSynchronize(st));
    cb[0] = 0;
    for (int64_t c = 0; c < nchunks; c++) cb[c + 1] = cb[c] + tot[c];
    DRRT_CHECK(S->r_chunk_base.reserve(nchunks + 1, st));
    DRRT_CUDA(cudaMemcpyAsync(S->r_chunk_base.p, cb.data(), (nchunks + 1) * sizeof(int64_t),
                              cudaMemcpyHostToDevice, st));
    DRRT_CUDA(cudaStreamSynchronize(st));
    A.chunk_base = S->r_chunk_base.p;
    extent = cb[nchunks];
  }
  return set_error(DRRT_ERR_STATE, "range: result slices did not converge");
}

}  // namespace drrt

#ifdef DRRT_PHASE_CLOCKS
extern "C" __attribute__((visibility("default"))) int drrt_debug_phases(unsigned long long *out, int reset) {
  if (cudaMemcpyFromSymbol(out, drrt::g_phase, sizeof(unsigned long long) * 16) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long z[16] = {0};
    if (cudaMemcpyToSymbol(drrt::g_phase, z, sizeof z) != cudaSuccess) return 1;
  }
  return 0;
}
#endif

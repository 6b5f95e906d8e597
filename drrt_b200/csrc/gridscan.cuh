// gridscan.cuh -- candidate enumeration over the cell-sorted node records,
// shared by the range (K2) and k-nearest (K3) kernels.
//
// A query of radius r selects the (x,y) columns whose footprint comes within r
// and, inside each column, one or two contiguous runs of records (theta is the
// fastest cell index; the second run is the periodic wrap).  A warp flattens
// the runs of 32 columns so every lane always has a candidate to evaluate.
// Visitors are called as f(valid, x, y, theta, id, gatef).
#pragma once

#include <type_traits>

#include "store.cuh"

namespace drrt {

// -DDRRT_PHASE_CLOCKS: thread 0 of every CTA adds the cycles between the marks of a query to
// g_phase[k] (an experiment build, benchmarks/phase_clocks.py; the marks compile to nothing otherwise)
#ifdef DRRT_PHASE_CLOCKS
__device__ unsigned long long g_phase[16];
#define DRRT_PHASE_DECL long long ph_last = clock64();
#define DRRT_PHASE_ARG , ph_last
#define DRRT_PHASE_PARAM , long long &ph_last
#define DRRT_PHASE(k)                                                        \
  do {                                                                       \
    if (threadIdx.x == 0) {                                                  \
      const long long ph_now = clock64();                                    \
      atomicAdd(&g_phase[k], (unsigned long long)(ph_now - ph_last));        \
      ph_last = ph_now;                                                      \
    }                                                                        \
  } while (0)
#else
#define DRRT_PHASE_DECL
#define DRRT_PHASE_ARG
#define DRRT_PHASE_PARAM
#define DRRT_PHASE(k) do { } while (0)
#endif

static constexpr int RUNS_PER_BATCH = 64;  // 32 columns x (run + wrapped run)

struct GridView {
  const NodeRec *rec;
  const int32_t *cell_start;
  GridParams gp;
  int64_t n_indexed, n;
  const double *x, *y, *th, *gate;
  int has_wrap;
  double wrap_point;
};

inline GridView make_view(const Store *S) {
  GridView V{};
  V.rec = S->rec.p; V.cell_start = S->cell_start.p; V.gp = S->gp;
  V.n_indexed = S->n_indexed; V.n = S->n;
  V.x = S->x.p; V.y = S->y.p; V.th = S->th.p; V.gate = S->gate.p;
  V.has_wrap = S->n_wraps > 0; V.wrap_point = S->wrap_point;
  return V;
}

// One 32-byte record with a single 256-bit load (sm_100: LDG.E.256); two 128-bit loads fetch
// the same sector twice and double the L1 requests of the scan.  gf = the float copy of gate[id].
__device__ __forceinline__ void load_rec(const NodeRec *r, double &x, double &y, double &t, int32_t &id,
                                         float &gf) {
  unsigned long long a, b, c, d;
  asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(r));
  x = __longlong_as_double((long long)a);
  y = __longlong_as_double((long long)b);
  t = __longlong_as_double((long long)c);
  id = (int32_t)(d & 0xffffffffull);
  gf = __uint_as_float((unsigned)(d >> 32));
}

// candidates that come straight from the SoA arrays (the unsorted tail) carry no float gate: NaN
// fails both of its comparisons and sends the candidate to the exact look-up
#define DRRT_NO_GATEF __uint_as_float(0x7fc00000u)

struct Query {
  double x, y, t, r;
  // Exact thresholds on the SQUARED sum: with d = fl(sqrt(s)) correctly rounded (and therefore
  // monotone in s),  d < r  <=>  s < t_lt  and  d <= r  <=>  s < t_le  (sqrt_thresholds below).
  // The scan decides on s and leaves the sqrt of the hits to the row write.
  double t_lt, t_le;
  double gt;
  bool has_ghost;
  int ix0, ix1, iy0, iy1;
  int ta0, ta1, tb0, tb1;  // theta cell runs of the whole query (b empty when tb0 > tb1)
  double rr, key;          // padded radius; theta key of the probes
  bool all;                // the ball covers the whole indexed box
  bool may_wrap;           // some column can have a second (wrapped) theta run
};

__device__ __forceinline__ int cell_floor(double v, double v0, double inv_h) {
  // one conversion that rounds down and saturates (NaN -> 0, as (int)floor(NaN)), then the clamp
  return max(-1000000000, min(1000000000, __double2int_rd((v - v0) * inv_h)));
}

// Cells of the theta axis within `half` of key: run a = [ta0,ta1], and the
// periodic wrap run b = [tb0,tb1] (empty when tb0 > tb1).
__device__ __forceinline__ void theta_runs(const GridParams &g, double key, double half, int &ta0,
                                           int &ta1, int &tb0, int &tb1) {
  const double slack = 1e-6 * g.ht;
  tb0 = 1; tb1 = 0;
  if (g.periodic) {
    if (!(2.0 * (half + slack) < g.period * (1.0 - 1e-9)) || !(key == key)) {
      ta0 = 0; ta1 = g.nt - 1;
      return;
    }
    int lo = cell_floor(key - half - slack, g.t0, g.inv_ht);
    int hi = cell_floor(key + half + slack, g.t0, g.inv_ht);
    if (hi - lo + 1 >= g.nt) { ta0 = 0; ta1 = g.nt - 1; return; }
    ta0 = max(lo, 0); ta1 = min(hi, g.nt - 1);
    if (lo < 0) { tb0 = lo + g.nt; tb1 = g.nt - 1; }
    else if (hi > g.nt - 1) { tb0 = 0; tb1 = hi - g.nt; }
    if (tb0 <= tb1 && tb0 <= ta1 && ta0 <= tb1) {  // overlap: merge
      ta0 = 0; ta1 = g.nt - 1; tb0 = 1; tb1 = 0;
    }
  } else {
    int lo = cell_floor(key - half - slack, g.t0, g.inv_ht);
    int hi = cell_floor(key + half + slack, g.t0, g.inv_ht);
    ta0 = max(lo, 0); ta1 = min(hi, g.nt - 1);
    if (hi < 0) { ta0 = ta1 = 0; }
    if (lo > g.nt - 1) { ta0 = ta1 = g.nt - 1; }
    if (!(key == key)) { ta0 = 0; ta1 = g.nt - 1; }
  }
}

// t_lt = the smallest double s with fl(sqrt(s)) >= r, t_le = the smallest with fl(sqrt(s)) > r.
// IEEE sqrt is correctly rounded, hence monotone non-decreasing in s, so for every double s
//   fl(sqrt(s)) <  r  <=>  s < t_lt        (kdtree.cpp:733, :765:  d < range)
//   fl(sqrt(s)) <= r  <=>  s < t_le        (kdtree.cpp:800-802: the root, d <= range)
// Both thresholds lie within a few ulps of r*r; a short walk from fl(r*r) finds them, a
// bisection over the bit patterns of the non-negative doubles covers radii whose square
// underflows or overflows.  r <= 0 or NaN: nothing is < r (t_lt = 0); only s == 0 is <= 0.
__device__ __forceinline__ double f64_next(double v) { return __longlong_as_double(__double_as_longlong(v) + 1); }
__device__ __forceinline__ double f64_prev(double v) { return __longlong_as_double(__double_as_longlong(v) - 1); }

static __device__ __noinline__ double sqrt_threshold_bisect(double r, bool strict_gt) {
  // smallest bit pattern b in [0, +inf] with sqrt(b) >= r (or > r)
  long long lo = 0, hi = 0x7ff0000000000000ll;  // pred(hi) holds unless r is +inf / beyond
  auto pred = [&](long long b) {
    const double v = sqrt(__longlong_as_double(b));
    return strict_gt ? v > r : v >= r;
  };
  if (pred(lo)) return __longlong_as_double(lo);
  if (!pred(hi)) return __longlong_as_double(hi);
  while (hi - lo > 1) {
    const long long mid = lo + ((hi - lo) >> 1);
    if (pred(mid)) hi = mid; else lo = mid;
  }
  return __longlong_as_double(hi);
}

__device__ __forceinline__ void sqrt_thresholds(double r, double &t_lt, double &t_le) {
  if (!(r > 0.0)) {
    t_lt = 0.0;
    t_le = (r == 0.0) ? __longlong_as_double(1) : 0.0;  // d <= 0 only for s == 0
    return;
  }
  double s = r * r;
  bool ok = s > 0.0 && s < 1.0e300;
  if (ok) {
    int guard = 0;
    while (sqrt(s) < r && guard < 8) { s = f64_next(s); guard++; }
    while (sqrt(f64_prev(s)) >= r && guard < 16) { s = f64_prev(s); guard++; }
    double u = s;
    while (sqrt(u) <= r && guard < 24) { u = f64_next(u); guard++; }
    ok = guard < 24 && sqrt(s) >= r && sqrt(f64_prev(s)) < r && sqrt(u) > r && sqrt(f64_prev(u)) <= r;
    t_lt = s;
    t_le = u;
  }
  if (!ok) {
    t_lt = sqrt_threshold_bisect(r, false);
    t_le = sqrt_threshold_bisect(r, true);
  }
}

// Fills the probe (ghost) and the candidate cell ranges of a query of radius r:
// every node n with metric(p,n) <= r for a probe p lies in the selected cells.
// THRESH: whether the exact sqrt thresholds are computed here (the range kernels decide on them;
// the CTA kernel computes them for a whole slice of queries at once and passes them in, the
// k-nearest kernels never look at them)
template <int METRIC, bool THRESH = true>
__device__ __forceinline__ void setup_query(const GridView &A, double qx, double qy, double qt,
                                            double r, Query &Q) {
  Q.x = qx;
  Q.y = qy;
  Q.t = qt;
  Q.r = r;
  if (THRESH) sqrt_thresholds(Q.r, Q.t_lt, Q.t_le);
  else { Q.t_lt = 0.0; Q.t_le = 0.0; }
  Q.has_ghost = false;
  Q.gt = 0.0;
  if (A.has_wrap) Q.has_ghost = ghost_of<METRIC>(Q.x, Q.y, Q.t, A.wrap_point, Q.r, &Q.gt);
  const GridParams &g = A.gp;
  double rr = fabs(Q.r) * (1.0 + 1e-9) + 1e-12;
  Q.rr = rr;
  int a = cell_floor(Q.x - rr, g.x0, g.inv_hx), b = cell_floor(Q.x + rr, g.x0, g.inv_hx);
  Q.ix0 = max(a, 0); Q.ix1 = min(b, g.nx - 1);
  if (b < 0) { Q.ix0 = Q.ix1 = 0; }
  if (a > g.nx - 1) { Q.ix0 = Q.ix1 = g.nx - 1; }
  a = cell_floor(Q.y - rr, g.y0, g.inv_hy); b = cell_floor(Q.y + rr, g.y0, g.inv_hy);
  Q.iy0 = max(a, 0); Q.iy1 = min(b, g.ny - 1);
  if (b < 0) { Q.iy0 = Q.iy1 = 0; }
  if (a > g.ny - 1) { Q.iy0 = Q.iy1 = g.ny - 1; }
  // theta cells for the whole query (per column they are narrowed, column_runs)
  Q.key = g.periodic ? theta_key(Q.t, g.period) : Q.t;
  theta_runs(g, Q.key, rr, Q.ta0, Q.ta1, Q.tb0, Q.tb1);
  // the ball swallows the whole indexed box: every column gets its full theta range
  double ex = fmax(fabs(Q.x - g.x0), fabs(Q.x - (g.x0 + g.nx * g.hx)));
  double ey = fmax(fabs(Q.y - g.y0), fabs(Q.y - (g.y0 + g.ny * g.hy)));
  double et = g.periodic ? 0.5 * g.period
                         : fmax(fabs(Q.t - g.t0), fabs(Q.t - (g.t0 + g.nt * g.ht)));
  Q.all = (ex * ex + ey * ey + et * et <= Q.r * Q.r) || !(Q.key == Q.key);
  // a narrower per-column range can wrap even when the query-wide one is the full circle
  Q.may_wrap = g.periodic && !Q.all && (Q.tb0 <= Q.tb1 || (Q.ta0 == 0 && Q.ta1 == g.nt - 1));
}

// whether the scan of Q visits every indexed node
__device__ __forceinline__ bool covers_all(const GridView &A, const Query &Q) {
  (void)A;
  return Q.all;
}

// the one or two contiguous record runs of candidate column `col` (empty when
// the column's footprint cannot come within the query radius)
__device__ __forceinline__ void column_runs(const GridView &A, const Query &Q, int col, int ncol,
                                            int ncy, double cull, int &s0, int &l0, int &s1,
                                            int &l1) {
  const GridParams &g = A.gp;
  s0 = l0 = s1 = l1 = 0;
  if (col >= ncol) return;
  int ix = Q.ix0 + col / ncy, iy = Q.iy0 + col % ncy;
  // closest approach of the column's footprint (edge columns are open-ended)
  double xlo = g.x0 + ((double)ix - 1e-6) * g.hx, xhi = g.x0 + ((double)ix + 1.000001) * g.hx;
  double ylo = g.y0 + ((double)iy - 1e-6) * g.hy, yhi = g.y0 + ((double)iy + 1.000001) * g.hy;
  double dx = 0.0, dy = 0.0;
  if (ix > 0 && Q.x < xlo) dx = xlo - Q.x;
  if (ix < g.nx - 1 && Q.x > xhi) dx = Q.x - xhi;
  if (iy > 0 && Q.y < ylo) dy = ylo - Q.y;
  if (iy < g.ny - 1 && Q.y > yhi) dy = Q.y - yhi;
  double rho2 = dx * dx + dy * dy;
  if (rho2 > cull) return;
  // a node of this column is at least rho from the probe in (x,y): its theta term
  // must stay under sqrt(r^2 - rho^2) to be in range
  int ta0 = Q.ta0, ta1 = Q.ta1, tb0 = Q.tb0, tb1 = Q.tb1;
  if (!Q.all) {
    // an upper bound is all the candidate cells need: single-precision square root of the argument
    // rounded up, itself rounded up (a few instructions where the double one takes ~30)
    double half = (double)__fsqrt_ru(__double2float_ru(fmax(Q.rr * Q.rr - rho2, 0.0))) * (1.0 + 1e-9) +
                  1e-9 * (1.0 + Q.rr);
    theta_runs(g, Q.key, half, ta0, ta1, tb0, tb1);
  }
  int cbase = (ix * g.ny + iy) * g.nt;
  s0 = A.cell_start[cbase + ta0];
  l0 = A.cell_start[cbase + ta1 + 1] - s0;
  if (tb0 <= tb1) {
    s1 = A.cell_start[cbase + tb0];
    l1 = A.cell_start[cbase + tb1 + 1] - s1;
  }
}

// One warp visits column batches first, first+stride, ... of query Q and calls
// f(valid, x, y, theta, id) once per candidate slot with all 32 lanes converged
// (valid == false on padding lanes).  The non-empty runs of a batch are compacted (two ballots)
// into `runs` (RUNS_PER_BATCH + 1 entries of the calling warp: run j covers candidate numbers
// [runs[j-1].x, runs[j].x) and candidate i of it is record i + runs[j].y, sentinel last), and a
// candidate's run comes from the same 32-run register window as in the CTA scan: one REDUX.OR
// marks where the window's runs end, a POPC away is the lane's run -- no per-candidate search.
template <class F>
__device__ __forceinline__ void scan_columns(const GridView &A, const Query &Q, int first,
                                             int stride, int2 *runs, int lane, F f) {
  const int ncy = Q.iy1 - Q.iy0 + 1;
  const int ncol = (Q.ix1 - Q.ix0 + 1) * ncy;
  const double cull = Q.rr * Q.rr * (1.0 + 1e-9);
  const unsigned lt = (1u << lane) - 1u;
  for (int cb = first * 32; cb < ncol; cb += stride * 32) {
    int s0, l0, s1, l1;
    column_runs(A, Q, cb + lane, ncol, ncy, cull, s0, l0, s1, l1);
    // exclusive scan of run lengths, two runs per lane
    int tot = l0 + l1, inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const int excl = inc - tot;
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    const unsigned m0 = __ballot_sync(0xffffffffu, l0 > 0), m1 = __ballot_sync(0xffffffffu, l1 > 0);
    const int nent = __popc(m0) + __popc(m1);
    const int slot = __popc(m0 & lt) + __popc(m1 & lt);  // runs of the lanes before this one
    __syncwarp();
    if (l0 > 0) runs[slot] = make_int2(excl + l0, s0 - excl);
    if (l1 > 0) runs[slot + (l0 > 0)] = make_int2(excl + l0 + l1, s1 - (excl + l0));
    if (lane == 31) runs[nent] = make_int2(0x7fffffff, 0);
    __syncwarp();
    int jl = 0;  // warp-uniform: the run holding the first candidate of the next 32
    for (int base = 0; base < total; base += 32) {
      const int i = base + lane;
      const bool valid = i < total;
      double nx = 0.0, ny = 0.0, nt = 0.0;
      int32_t id = 0;
      float gf = 0.0f;
      const int2 e = runs[min(jl + lane, nent)];
      const unsigned rel = (unsigned)(e.x - base);
      const unsigned ends = __reduce_or_sync(0xffffffffu, rel < 32u ? 1u << rel : 0u);
      const int k = __popc(ends & ((2u << lane) - 1u));  // runs ending at or before candidate i
      const int ey = __shfl_sync(0xffffffffu, e.y, k);
      if (valid) load_rec(A.rec + (i + ey), nx, ny, nt, id, gf);
      jl += __popc(ends) + (__any_sync(0xffffffffu, rel == 32u) ? 1 : 0);
      f(valid, nx, ny, nt, id, gf);
    }
  }
}

// Flattened candidate list of a CTA-wide scan: run j covers candidate numbers
// [runs[j-1].x, runs[j].x) and candidate i of it is record i + runs[j].y.
// A whole CTA of NWARPS warps visits ALL candidate columns of query Q: every
// thread resolves one column, a CTA-wide prefix sum lists the non-empty runs,
// each warp takes an equal contiguous slice of the candidate numbers and every
// lane handles UNR candidates per trip (i, i+32, ...); the run of a candidate is found from a
// 32-run window held in the warp's registers (one REDUX.OR + POPC).  runs holds 2*T+1
// entries, tmp64 NWARPS words.
// f(valid[UNR], x[UNR], y[UNR], theta[UNR], id[UNR], gatef[UNR]) is called converged.
template <int NWARPS, int UNR, class F>
__device__ __forceinline__ void scan_columns_cta(const GridView &A, const Query &Q, int2 *runs,
                                                 long long *tmp64, F f DRRT_PHASE_PARAM) {
  constexpr int T = NWARPS * 32;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int ncy = Q.iy1 - Q.iy0 + 1;
  const int ncol = (Q.ix1 - Q.ix0 + 1) * ncy;
  const double cull = Q.rr * Q.rr * (1.0 + 1e-9);
  for (int cb = 0; cb < ncol; cb += T) {
    int s0, l0, s1, l1;
    column_runs(A, Q, cb + tid, ncol, ncy, cull, s0, l0, s1, l1);
    // candidates before this thread's: a 32-bit shuffle scan; runs before it: two ballots
    const unsigned ltm = (1u << lane) - 1u;
    const unsigned m0 = __ballot_sync(0xffffffffu, l0 > 0), m1 = __ballot_sync(0xffffffffu, l1 > 0);
    int cinc = l0 + l1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, cinc, o);
      if (lane >= o) cinc += t;
    }
    const long long mine = ((long long)((l0 > 0) + (l1 > 0)) << 32) | (unsigned)(l0 + l1);
    const long long inc = ((long long)(__popc(m0 & ltm) + __popc(m1 & ltm) + (l0 > 0) + (l1 > 0)) << 32) | (unsigned)cinc;
    if (lane == 31) tmp64[w] = inc;
    __syncthreads();
    DRRT_PHASE(1);
    // every warp scans the NWARPS warp totals with its first lanes
    static_assert(NWARPS <= 32, "one lane per warp total");
    long long wt = (lane < NWARPS) ? tmp64[lane] : 0, winc = wt;
#pragma unroll
    for (int o = 1; o < NWARPS; o <<= 1) {
      long long t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    const long long all = __shfl_sync(0xffffffffu, winc, NWARPS - 1);
    const long long wbase = __shfl_sync(0xffffffffu, winc - wt, w);
    const long long excl64 = wbase + inc - mine;
    const int excl = (int)(excl64 & 0xffffffffll), slot = (int)(excl64 >> 32);
    const int total = (int)(all & 0xffffffffll), nent = (int)(all >> 32);
    if (l0 > 0) runs[slot] = make_int2(excl + l0, s0 - excl);
    if (l1 > 0) runs[slot + (l0 > 0)] = make_int2(excl + l0 + l1, s1 - (excl + l0));
    if (tid == T - 1) runs[nent] = make_int2(0x7fffffff, 0);
    __syncthreads();
    DRRT_PHASE(2);
#ifdef DRRT_SPLIT_OLD
    const int per = (((total + NWARPS - 1) / NWARPS) + 31) & ~31;
    const int c0 = w * per, c1 = min(total, c0 + per);
#else
    // whole trips (32 * UNR candidates) dealt out in contiguous blocks that differ by at most one trip:
    // only the query's last trip is partly filled (equal shares rounded up to 32 candidates left every
    // warp with a half-empty last trip -- 80 trips for the 71 that config 2a's 4 500 candidates need)
    constexpr int TRIP = 32 * UNR;
    const int ntrips = (total + TRIP - 1) / TRIP;
    const int tq = ntrips / NWARPS, trem = ntrips - tq * NWARPS;
    const int c0 = (w * tq + min(w, trem)) * TRIP;
    const int c1 = min(total, c0 + (tq + (w < trem ? 1 : 0)) * TRIP);
#endif
    if (c0 < c1) {
      // smallest j with runs[j].x > c0 (the sentinel runs[nent] qualifies): a 32-way search, every
      // lane probes one entry per round -- two rounds of shared-memory latency for up to 1024 runs
      // instead of ten dependent ones
      int lo = 0, hi = nent;
      for (;;) {
        const int step = (hi - lo + 32) >> 5;  // ceil((hi - lo + 1) / 32)
        const int pr = min(lo + (lane + 1) * step - 1, hi);
        const unsigned above = __ballot_sync(0xffffffffu, runs[pr].x > c0);
        const int L = __ffs(above) - 1;  // lane 31 probes hi, which qualifies
        hi = min(lo + (L + 1) * step - 1, hi);
        if (step == 1) break;
        lo = lo + L * step;
      }
      int jl = hi;
      // jl is warp-uniform: the run holding the first candidate of the next 32.  Runs are never
      // empty, so those 32 candidates lie in runs jl .. jl+31: every lane fetches one of them, one
      // warp-wide OR marks where runs end, and a lane's run is a population count away (no
      // divergent walk, no chain of dependent shared-memory loads or shuffles).
      // a trip of 32 * UNR candidates; FULL trips (all but the query's last one) load and test without
      // guards -- no zero fill, no predicated copies of the eight record registers per candidate, no
      // `i < c1` in the candidate test: 5.4 % fewer warp instructions, 3.495 -> 3.43 ms (config 2a)
      auto trip = [&](auto full_tag, int base) {
        constexpr bool FULL = decltype(full_tag)::value;
        bool valid[UNR];
        double nx[UNR], ny[UNR], nt[UNR];
        int32_t id[UNR];
        float gf[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          const int i = base + u * 32 + lane;
          valid[u] = FULL || i < c1;
          const int2 e = runs[min(jl + lane, nent)];  // runs[nent] is the sentinel
          // bit b of `ends`: some run of the window ends (exclusively) at candidate i0 + b.  Run jl
          // holds i0, so b >= 1, and runs are never empty, so the bits are distinct.
          const int i0 = base + u * 32;
          const unsigned rel = (unsigned)(e.x - i0);
          const unsigned ends = __reduce_or_sync(0xffffffffu, rel < 32u ? 1u << rel : 0u);
          const int k = __popc(ends & ((2u << lane) - 1u));  // runs ending at or before candidate i
          const int ey = __shfl_sync(0xffffffffu, e.y, k);
          if (FULL) {
            load_rec(A.rec + (i + ey), nx[u], ny[u], nt[u], id[u], gf[u]);
          } else {
            nx[u] = ny[u] = nt[u] = 0.0;
            id[u] = 0;
            gf[u] = 0.0f;
            if (valid[u]) load_rec(A.rec + (i + ey), nx[u], ny[u], nt[u], id[u], gf[u]);
          }
          jl += __popc(ends) + (__any_sync(0xffffffffu, rel == 32u) ? 1 : 0);
        }
        f(valid, nx, ny, nt, id, gf);
      };
      int base = c0;
      for (; base + 32 * UNR <= c1; base += 32 * UNR) trip(std::true_type{}, base);
      if (base < c1) trip(std::false_type{}, base);
    }
    DRRT_PHASE(3);   // thread 0's own trips
    __syncthreads();
    DRRT_PHASE(4);   // waiting for the slowest warp
  }
}

// tail scan of a whole CTA, UNR candidates per lane per trip
template <int NWARPS, int UNR, class F>
__device__ __forceinline__ void scan_tail_cta(const GridView &A, F f) {
  const int tid = threadIdx.x;
  for (int64_t b = A.n_indexed; b < A.n; b += (int64_t)NWARPS * 32 * UNR) {
    bool valid[UNR];
    double nx[UNR], ny[UNR], nt[UNR];
    int32_t id[UNR];
    float gf[UNR];
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      int64_t i = b + (int64_t)u * NWARPS * 32 + tid;
      valid[u] = i < A.n;
      nx[u] = ny[u] = nt[u] = 0.0;
      id[u] = (int32_t)i;
      gf[u] = DRRT_NO_GATEF;
      if (valid[u]) { nx[u] = A.x[i]; ny[u] = A.y[i]; nt[u] = A.th[i]; }
    }
    f(valid, nx, ny, nt, id, gf);
  }
}

// nodes inserted after the last index build: straight from the SoA store
template <class F>
__device__ __forceinline__ void scan_tail(const GridView &A, int first, int stride, int lane, F f) {
  for (int64_t b = A.n_indexed + (int64_t)first * 32; b < A.n; b += (int64_t)stride * 32) {
    int64_t i = b + lane;
    bool valid = i < A.n;
    double nx = 0.0, ny = 0.0, nt = 0.0;
    if (valid) { nx = A.x[i]; ny = A.y[i]; nt = A.th[i]; }
    f(valid, nx, ny, nt, (int32_t)i, DRRT_NO_GATEF);
  }
}

}  // namespace drrt

// edgecheck.cu -- K4/K5: batched edge collision checks, plus the trajectory
// and node-validity companions.
//
// Replaces, fused into one kernel per batch:
//   Edge::NewEdge                     src/dubinsedge.cpp:12-21
//   DubinsEdge::CalculateTrajectory   src/dubinsedge.cpp:172-830
//   DubinsEdge::ExplicitEdgeCheck     src/dubinsedge.cpp:850-898 (analytic call :881-884)
//   ExplicitEdgeCheck(C, edge)        src/obstacle.cpp:879-923
//   ExplicitEdgeCheck2D               src/obstacle.cpp:756-804
//   LineCheck polyline                src/obstacle.cpp:1092-1143
// The verdict is an OR over (segment, obstacle) pairs, so the reference's
// first-hit exit order does not matter.  Only obstacles of kind 1 (ball) and
// kind 3 (polygon) that are used and alive can ever report a hit
// (:761, :781-804); the host packs those into the "active" table.
//
// Dubins edges take two stages.  The solve kernels (one thread per edge; ranked and binned by
// Dubins word for large batches) find the path, decide the edges a pose of which surely collides
// (finish_path) and append a 112-byte path record of every undecided edge to a work list.
// dubins_walk_kernel is persistent: every lane walks one listed edge's 0.1-rad polyline a segment
// per iteration and, when its edge is decided (first hit, or end of path), takes the next
// unclaimed record, so lanes of a warp stay busy although polylines have 3..191 segments.
// Straight (LineCheck) edges have 50 segments each and use one kernel.
// Each segment looks up the cells of a uniform grid its bounding box touches; a
// cell lists the polygon EDGES (and kind-1 balls) that come within the robot
// radius of it, so an edge that is not listed cannot bring SegmentDistSqrd
// under radius^2 for that segment.  A listed edge farther from the segment's
// first point than (radius + segment length) is skipped; the rest get the
// reference's SegmentDistSqrd and, on a hit, its obstacle's bounding reject
// (:765-779) verbatim -- the verdict is the same OR over obstacles.  The grid entries are 32-byte records read with one
// 256-bit ld.global.nc each (L1/L2 resident); staging the table in shared memory measured slower.
#include <math.h>

#include <algorithm>

#include "geom.cuh"
#include "store.cuh"

namespace drrt {

static constexpr int EDGE_T = 128;
static constexpr int64_t RANKED_SOLVE_MIN = 16384;  // smaller Dubins batches take the one-kernel solve
static constexpr int WALK_CLAIM = 64;  // edges a warp claims at a time
static constexpr int WALK_AHEAD = 2;   // points a lane may advance before the warp tests the waiting segments

__device__ __forceinline__ void load_edge(const ObstacleGrid &G, int j, double &ax, double &ay, double &bx, double &by) {
  unsigned long long a, b, c, d;
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(G.erec + j));
  ax = __longlong_as_double((long long)a); ay = __longlong_as_double((long long)b);
  bx = __longlong_as_double((long long)c); by = __longlong_as_double((long long)d);
}

struct ObsView {
  int n;
  const uint8_t *mask;  // null, or per obstacle: 0 = leave it out (subset checks of the sweeps)
  const uint8_t *covers;  // per obstacle: radius_ is at least the largest vertex distance from origin_
  const int32_t *kind;
  const double *ox, *oy, *rad;
  const int32_t *voff;
  const double *vx, *vy;
};

// ExplicitEdgeCheck2D for one active obstacle (kind 1 or 3).  far2 =
// ((radius + |p1-p0|) * (1+1e-9) + 1e-9)^2 : polygon edges farther than that
// from p0 are at more than `radius` from every point of the segment.
template <bool MASKED>
__device__ __forceinline__ bool segment_hits(const ObsView &O, int o, double p0x, double p0y,
                                             double p1x, double p1y, double radius, double far2) {
  if (MASKED && !O.mask[o]) return false;
  double d2 = dist_sqrd_point_segment(O.ox[o], O.oy[o], p0x, p0y, p1x, p1y);  // :769-779
  double reach = radius + O.rad[o];
  if (d2 > reach * reach) return false;
  if (O.kind[o] == 1) return true;  // :781
  int v0 = O.voff[o], v1 = O.voff[o + 1];
  if (v1 - v0 < 2) return false;  // :785
  double ax = O.vx[v1 - 1], ay = O.vy[v1 - 1];
  const double rr = radius * radius;
  for (int v = v0; v < v1; v++) {  // :788-804
    double bx = O.vx[v], by = O.vy[v];
    if (!(dist_sqrd_point_segment(p0x, p0y, ax, ay, bx, by) > far2)) {
      if (segment_dist_sqrd(p0x, p0y, p1x, p1y, ax, ay, bx, by) < rr) return true;
    }
    ax = bx;
    ay = by;
  }
  return false;
}

// the reference's bounding reject for obstacle o (obstacle.cpp:765-779)
__device__ __forceinline__ bool within_reach(const ObsView &O, int o, double p0x, double p0y,
                                             double p1x, double p1y, double radius) {
  double d2 = dist_sqrd_point_segment(O.ox[o], O.oy[o], p0x, p0y, p1x, p1y);
  double reach = radius + O.rad[o];
  return !(d2 > reach * reach);
}

// ExplicitEdgeCheck2D of one polyline segment against every obstacle, through the edge grid:
//   true  <=>  some live ball passes the bounding reject, or some polygon edge has
//              SegmentDistSqrd < radius^2 and its obstacle passes the bounding reject
// which is the OR over obstacles of obstacle.cpp:761-804.
template <bool MASKED>
__device__ __forceinline__ bool segment_vs_obstacles(const ObsView &O, const ObstacleGrid &G,
                                                     double p0x, double p0y, double p1x,
                                                     double p1y, double radius) {
  double dx = p1x - p0x, dy = p1y - p0y;
  double far = (radius + sqrt(dx * dx + dy * dy)) * (1.0 + 1e-9) + 1e-9;
  double far2 = far * far;
  double lox = fmin(p0x, p1x), hix = fmax(p0x, p1x), loy = fmin(p0y, p1y), hiy = fmax(p0y, p1y);
  if (G.nx > 0 && lox <= hix && loy <= hiy) {
    // cells the segment's bounding box touches; beyond the grid nothing is in reach
    double fx0 = floor((lox - G.x0) * G.inv_cell), fx1 = floor((hix - G.x0) * G.inv_cell);
    double fy0 = floor((loy - G.y0) * G.inv_cell), fy1 = floor((hiy - G.y0) * G.inv_cell);
    if (fx1 < 0.0 || fy1 < 0.0 || fx0 > (double)(G.nx - 1) || fy0 > (double)(G.ny - 1)) return false;
    int cx0 = (int)fmax(fx0, 0.0), cx1 = (int)fmin(fx1, (double)(G.nx - 1));
    int cy0 = (int)fmax(fy0, 0.0), cy1 = (int)fmin(fy1, (double)(G.ny - 1));
    const double rr = radius * radius;
    for (int cx = cx0; cx <= cx1; cx++)
      for (int cy = cy0; cy <= cy1; cy++) {
        int c = cx * G.ny + cy;
        for (int e = G.off[c]; e < G.off[c + 1]; e++) {
          int j = G.idx[e];
          if (j >= G.n_edges) {  // ball: obstacle.cpp:781
            const int b = G.ball[j - G.n_edges];
            if (MASKED && !O.mask[b]) continue;
            if (within_reach(O, b, p0x, p0y, p1x, p1y, radius)) return true;
            continue;
          }
          if (MASKED && !O.mask[G.eobs[j]]) continue;
          double ax, ay, bx, by;
          load_edge(G, j, ax, ay, bx, by);
          // an edge farther than radius + |segment| from p0 cannot come within radius of the segment
          if (dist_sqrd_point_segment(p0x, p0y, ax, ay, bx, by) > far2) continue;
          if (segment_dist_sqrd(p0x, p0y, p1x, p1y, ax, ay, bx, by) < rr &&   // :796-803
              within_reach(O, G.eobs[j], p0x, p0y, p1x, p1y, radius))          // :765-779
            return true;
        }
      }
    return false;
  }
  for (int o = 0; o < O.n; o++)
    if (segment_hits<MASKED>(O, o, p0x, p0y, p1x, p1y, radius, far2)) return true;
  return false;
}

// Per-warp scratch of the cooperative segment test.
struct __align__(16) WarpScratch {
  double seg[32][6];   // p0x p0y p1x p1y far2 (+ pad: rows are read with 128-bit loads) of each lane's segment
  int start[32][4];    // up to four grid cells per segment: entry range start / length
  int len[32][4];
  int own[32];         // the lanes that bring entries, in lane order: lane | (pairs before its own) << 5
};

// one grid entry (polygon edge or ball) against one segment: the reference's tests verbatim
template <bool MASKED>
__device__ __forceinline__ bool entry_hits(const ObsView &O, const ObstacleGrid &G, int j,
                                           const double *sg, double radius) {
  if (j >= G.n_edges) {  // ball: obstacle.cpp:781
    const int b = G.ball[j - G.n_edges];
    if (MASKED && !O.mask[b]) return false;
    return within_reach(O, b, sg[0], sg[1], sg[2], sg[3], radius);
  }
  if (MASKED && !O.mask[G.eobs[j]]) return false;
  double ax, ay, bx, by;
  load_edge(G, j, ax, ay, bx, by);
  // an edge farther than radius + |segment| from p0 cannot come within radius of the segment
  if (dist_sqrd_point_segment_sel(sg[0], sg[1], ax, ay, bx, by) > sg[4]) return false;
  return segment_dist_sqrd(sg[0], sg[1], sg[2], sg[3], ax, ay, bx, by) < radius * radius &&  // :796-803
         within_reach(O, G.eobs[j], sg[0], sg[1], sg[2], sg[3], radius);                     // :765-779
}

// Warp-cooperative ExplicitEdgeCheck2D in two halves: every lane brings (at most) one polyline
// segment; the (segment, grid entry) pairs of the whole warp are flattened so that all 32 lanes
// test pairs, whichever lane's segment they belong to.  Most segments meet no entry at all and a
// few meet several, so a per-lane loop runs at ~2 active lanes; this keeps them busy.
// First half, per lane: the grid cells of one segment.  Returns the number of grid entries the
// segment has to be tested against (`wide`: it spans more than four cells and is swept by the whole
// warp instead); when there is anything to test, the lane's cell lists and the segment itself are
// left in the warp scratch for the second half.
__device__ __forceinline__ int lane_segment_cells(const ObstacleGrid &G, double p0x, double p0y, double p1x,
                                                  double p1y, double radius, WarpScratch *ws, int lane,
                                                  bool &wide) {
  int cx0 = 0, cx1 = -1, cy0 = 0, cy1 = -1;  // empty
  // Points that are not finite numbers test nothing -- decided BESIDE the conversions (they saturate,
  // not-a-number gives cell 0) and applied to their result, not in front of them as a branch on a chain
  // of four dependent additions (7.34 -> 7.26 ms per 10 M edges)
  const double fsum = (p0x + p1x) + (p0y + p1y);
  const bool finite = fsum - fsum == 0.0;
  {
    const int ax = __double2int_rd((p0x - G.x0) * G.inv_cell), bx = __double2int_rd((p1x - G.x0) * G.inv_cell);
    const int ay = __double2int_rd((p0y - G.y0) * G.inv_cell), by = __double2int_rd((p1y - G.y0) * G.inv_cell);
    const int x0 = min(ax, bx), x1 = max(ax, bx), y0 = min(ay, by), y1 = max(ay, by);
    if (!(x1 < 0 || y1 < 0 || x0 > G.nx - 1 || y0 > G.ny - 1)) {
      cx0 = max(x0, 0); cx1 = min(x1, G.nx - 1);
      cy0 = max(y0, 0); cy1 = min(y1, G.ny - 1);
    }
  }
  const int ncx = cx1 - cx0 + 1, ncy = cy1 - cy0 + 1;
  const int ncells = (ncx > 0 && ncy > 0 && finite) ? ncx * ncy : 0;
  wide = ncells > 4;
  int total = 0, st[4], ln[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    st[k] = 0; ln[k] = 0;
    if (!wide && k < ncells) {
      const int kx = (k >= ncy) + (k >= 2 * ncy) + (k >= 3 * ncy);  // k / ncy for k < 4
      int c = (cx0 + kx) * G.ny + (cy0 + k - kx * ncy);
      st[k] = G.off[c];
      ln[k] = G.off[c + 1] - st[k];
    }
    total += ln[k];
  }
  if (total > 0 || wide) {
    // (a many-cell segment has no cell lists: its cell box -- cx0, cy0, cells per column, cells -- takes
    // their place; a 40 KB CTA keeps the shared-memory carve-out of four CTAs a step lower, which is
    // worth 4 % through the L1 that is left)
    *reinterpret_cast<int4 *>(ws->start[lane]) = wide ? make_int4(cx0, cy0, ncy, ncells) : make_int4(st[0], st[1], st[2], st[3]);
    *reinterpret_cast<int4 *>(ws->len[lane]) = make_int4(ln[0], ln[1], ln[2], ln[3]);
    // cull distance: radius + an upper bound of the segment's length (|dx| + |dy|, no sqrt)
    const double far = (radius + (fabs(p1x - p0x) + fabs(p1y - p0y))) * (1.0 + 1e-9) + 1e-9;
    *reinterpret_cast<double2 *>(&ws->seg[lane][0]) = make_double2(p0x, p0y);
    *reinterpret_cast<double2 *>(&ws->seg[lane][2]) = make_double2(p1x, p1y);
    ws->seg[lane][4] = far * far;
  }
  return total;
}

// Second half, the whole warp: the (segment, entry) pairs of the lanes whose segment waits in the
// scratch (`total` entries, or `wide`), flattened and tested by all lanes.  Returns whether the
// calling lane's own segment collides.
template <bool MASKED>
__device__ bool warp_pending_vs_obstacles(const ObsView &O, const ObstacleGrid &G, int total, bool wide,
                                          double radius, WarpScratch *ws, int lane) {
  __syncwarp();  // the lanes' scratch entries are written
  int inc = total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  const unsigned ltm = (1u << lane) - 1u;
  const unsigned bring = __ballot_sync(0xffffffffu, total > 0);
  if (total > 0) ws->own[__popc(bring & ltm)] = lane | ((inc - total) << 5);
  __syncwarp();
  const int all = __shfl_sync(0xffffffffu, inc, 31);
  unsigned hitmask = 0;
  int jbase = 0;  // lanes whose entries end before `base`
  for (int base = 0; base < all; base += 32) {
    const int p = base + lane;
    const unsigned rel = (unsigned)(inc - 1 - base);
    const unsigned ends = __reduce_or_sync(0xffffffffu, (total > 0 && rel < 32u) ? 1u << rel : 0u);
    bool h = false;
    int lo = 0;
    if (p < all) {
      const int o = ws->own[jbase + __popc(ends & ltm)];
      lo = o & 31;
      const int4 l4 = *reinterpret_cast<const int4 *>(ws->len[lo]);
      const int4 s4 = *reinterpret_cast<const int4 *>(ws->start[lo]);
      const int q0 = p - (o >> 5), b1 = l4.x, b2 = b1 + l4.y, b3 = b2 + l4.z;
      const int e = q0 < b1 ? s4.x + q0 : q0 < b2 ? s4.y + (q0 - b1) : q0 < b3 ? s4.z + (q0 - b2) : s4.w + (q0 - b3);
      const double2 sa = *reinterpret_cast<const double2 *>(&ws->seg[lo][0]);
      const double2 sb = *reinterpret_cast<const double2 *>(&ws->seg[lo][2]);
      const double sg[5] = {sa.x, sa.y, sb.x, sb.y, ws->seg[lo][4]};
      h = entry_hits<MASKED>(O, G, G.idx[e], sg, radius);
    }
    hitmask |= __reduce_or_sync(0xffffffffu, h ? 1u << lo : 0u);
    jbase += __popc(ends);
  }
  // segments spanning many cells: one at a time, the warp sweeps its cells
  unsigned wmask = __ballot_sync(0xffffffffu, wide);
  while (wmask) {
    const int L = __ffs(wmask) - 1;
    wmask &= wmask - 1;
    const int4 bx = *reinterpret_cast<const int4 *>(ws->start[L]);
    bool h = false;
    for (int k = lane; k < bx.w && !h; k += 32) {
      int c = (bx.x + k / bx.z) * G.ny + (bx.y + k % bx.z);
      for (int e = G.off[c]; e < G.off[c + 1] && !h; e++) h = entry_hits<MASKED>(O, G, G.idx[e], ws->seg[L], radius);
    }
    if (__any_sync(0xffffffffu, h)) hitmask |= 1u << L;
  }
  __syncwarp();  // the scratch is reused by the next round
  return (hitmask >> lane) & 1u;
}

struct EdgeArgs {
  int64_t n;
  const double *start, *end;        // n x 3, or null when ids are given
  const int32_t *sid, *eid;
  const double *sx, *sy, *sth;      // store SoA (ids mode)
  // neighbour-list mode (drrt_extend_batch): edge e = 2*t + dir over the CSR rows of a range
  // result; dir 0 = query -> neighbour (FindBestParent), dir 1 = neighbour -> query (Extend)
  const int64_t *csr_off;
  const int32_t *csr_row;  // row (query) of every pair
  const int32_t *csr_ids;
  const double *csr_q;
  int64_t csr_nq;
  int metric;
  int edge_kind;
  double robot_radius, r_min;
  ObsView obs;
  ObstacleGrid grid;
  uint8_t *verdict;
  double *length;
};

__device__ __forceinline__ void load_pose(const EdgeArgs &A, int64_t e, bool end, double &x,
                                          double &y, double &t) {
  if (A.csr_off) {
    const int64_t pair = e >> 1;
    const bool node_side = ((e & 1) != 0) != end;  // dir 0: start = query; dir 1: start = neighbour
    if (node_side) {
      int32_t id = A.csr_ids[pair];
      x = A.sx[id]; y = A.sy[id]; t = A.sth[id];
    } else {
      const int64_t lo = A.csr_row[pair];  // the query whose neighbour list holds this pair
      x = A.csr_q[3 * lo]; y = A.csr_q[3 * lo + 1]; t = A.csr_q[3 * lo + 2];
    }
    return;
  }
  const int32_t *ids = end ? A.eid : A.sid;
  if (ids) {
    int32_t id = ids[e];
    x = A.sx[id]; y = A.sy[id]; t = A.sth[id];
  } else {
    const double *p = (end ? A.end : A.start) + 3 * e;
    x = p[0]; y = p[1]; t = p[2];
  }
}

// no grid entry (polygon edge or ball within the robot radius) in any cell touching the box
__device__ __forceinline__ bool region_empty(const ObstacleGrid &G, double lox, double loy,
                                             double hix, double hiy) {
  if (G.nx == 0 || !(lox <= hix) || !(loy <= hiy)) return false;
  double fx0 = floor((lox - G.x0) * G.inv_cell), fx1 = floor((hix - G.x0) * G.inv_cell);
  double fy0 = floor((loy - G.y0) * G.inv_cell), fy1 = floor((hiy - G.y0) * G.inv_cell);
  if (fx1 < 0.0 || fy1 < 0.0 || fx0 > (double)(G.nx - 1) || fy0 > (double)(G.ny - 1)) return true;
  int cx0 = (int)fmax(fx0, 0.0), cx1 = (int)fmin(fx1, (double)(G.nx - 1));
  int cy0 = (int)fmax(fy0, 0.0), cy1 = (int)fmin(fy1, (double)(G.ny - 1));
  for (int cx = cx0; cx <= cx1; cx++)  // a row's cells are consecutive in the CSR
    if (G.off[cx * G.ny + cy1 + 1] != G.off[cx * G.ny + cy0]) return false;
  return true;
}

// Is the point (x, y) so close to an obstacle that any polyline segment ending in a point within
// 1e-12 of it collides?  True when a polygon edge listed for the point's grid cell (every edge
// within the robot radius of the cell is) lies closer than radius - 1e-9 and its obstacle's
// radius_ covers its polygon, or the point is that deep inside a ball's reach.  Then
//  * SegmentDistSqrd of the segment against that edge is 0 or at most the endpoint's own
//    DistanceSqrdPointToSegment (distancefunctions.cpp:153-163), i.e. < radius^2 (:796-803), and
//  * the segment passes the obstacle's bounding reject (obstacle.cpp:765-779): the endpoint is
//    within radius - 1e-9 of a polygon point, which is within radius_ of origin_,
// with nine orders of magnitude between the 1e-9 margin and the rounding of either test.
template <bool MASKED>
__device__ __forceinline__ bool point_surely_collides(const ObsView &O, const ObstacleGrid &G, double x,
                                                      double y, double radius) {
  const double rs = radius - 1e-9;
  if (G.nx == 0 || !(rs > 0.0)) return false;
  const double fx = floor((x - G.x0) * G.inv_cell), fy = floor((y - G.y0) * G.inv_cell);
  if (!(fx >= 0.0 && fy >= 0.0 && fx <= (double)(G.nx - 1) && fy <= (double)(G.ny - 1))) return false;
  const int c = (int)fx * G.ny + (int)fy;
  const double rr = rs * rs;
  for (int e = G.off[c]; e < G.off[c + 1]; e++) {
    const int j = G.idx[e];
    if (j >= G.n_edges) {  // ball: obstacle.cpp:781
      const int b = G.ball[j - G.n_edges];
      if (MASKED && !O.mask[b]) continue;
      const double dx = x - O.ox[b], dy = y - O.oy[b], reach = rs + O.rad[b];
      if (reach > 0.0 && dx * dx + dy * dy < reach * reach) return true;
      continue;
    }
    const int o = G.eobs[j];
    if (MASKED && !O.mask[o]) continue;
    if (!O.covers[o]) continue;
    double ax, ay, bx, by;
    load_edge(G, j, ax, ay, bx, by);
    if (dist_sqrd_point_segment(x, y, ax, ay, bx, by) < rr) return true;
  }
  return false;
}

// Decides what it can about a solved path and hands the rest to the walk kernel.  All 32 lanes
// of a warp call it together (`valid`: the lane has an edge from (sx, sy) to (ex, ey)).
//  * parts whose whole circle (or straight piece) lies where no obstacle comes within the robot
//    radius need no per-segment test; a path made only of such parts is free;
//  * every path starts in the start pose and ends in the end pose (to the rounding of one
//    atan2 + sincos, ~1e-14), and the verdict is an OR over its segments: a pose that
//    point_surely_collides decides the edge without walking it.  The planner's nodes are not
//    filtered against the obstacles (the range query returns what the tree holds), so the usual
//    colliding edge is of this kind -- in config 3 ~95 % of them -- and one whose START is
//    clear would otherwise be walked to its very last segment before the hit shows;
//  * an undecided edge gets the next slot of the compact work list (its record carries the edge
//    id in place of dist_, which the walk does not read), so the walk kernel neither loads nor
//    spends a trip on the decided ones.
// work[0] = walk claim counter, work[1] = number of records listed.
template <bool MASKED>
__device__ __forceinline__ void finish_path_impl(const EdgeArgs &A, bool valid, int64_t e, DubinsPath &P,
                                                 double sx, double sy, double ex, double ey,
                                                 DubinsPath *paths, unsigned long long *work) {
  const int lane = threadIdx.x & 31;
  if (valid && A.length) A.length[e] = P.dist;
  if (A.obs.n == 0) return;    // nothing to collide with: the host clears the verdicts
  int decided = valid ? -1 : 0;  // 0 free, 1 collides (a lane without an edge has nothing to decide)
  if (valid) {
    if ((P.kinds & 63) == 0) {
      decided = 0;  // no path: dubinsedge.cpp:897 falls through to false
    } else if (A.grid.nx > 0) {
      int idle = 0, parts = 0, last = 0;
      const double rr = fabs(A.r_min) * (1.0 + 1e-9) + 1e-9;
      for (int k = 0; k < 3; k++) {
        int kind = P.kind(k);
        if (kind == PART_NONE) continue;
        parts++;
        last = k;
        const DubinsPart &p = P.part[k];
        bool empty = (kind == PART_LINE)
                         ? region_empty(A.grid, fmin(p.cx, p.a), fmin(p.cy, p.b), fmax(p.cx, p.a), fmax(p.cy, p.b))
                         : region_empty(A.grid, p.cx - rr, p.cy - rr, p.cx + rr, p.cy + rr);
        if (empty) { idle++; P.kinds |= 1 << (8 + k); }
      }
      if (idle == parts) {
        decided = 0;
      } else if (point_surely_collides<MASKED>(A.obs, A.grid, sx, sy, A.robot_radius)) {
        decided = 1;
      } else {
        // the last arc ends in phi_end (the end pose) unless |dphi| / 0.1 sits on an integer, where
        // the reference's loop may leave the last entry Zero() (PolylineWalker::next)
        bool ends_in_pose = true;
        const int kind = P.kind(last);
        if (kind != PART_LINE) {
          const DubinsPart &p = P.part[last];
          double pe = p.b;
          if (kind == PART_ARC_R) { if (pe > p.a) pe -= 2.0 * DRRT_PI; }
          else { if (pe < p.a) pe += 2.0 * DRRT_PI; }
          if (pe != p.a) {
            const double steps = fabs(pe - p.a) / 0.1, frac = steps - floor(steps);
            ends_in_pose = frac > 1e-9 && frac < 1.0 - 1e-9;
          }
        }
        if (ends_in_pose && point_surely_collides<MASKED>(A.obs, A.grid, ex, ey, A.robot_radius)) decided = 1;
      }
    }
  }
  if (valid && decided >= 0) A.verdict[e] = (uint8_t)decided;
  // warp-aggregated append of the undecided edges
  const unsigned m = __ballot_sync(0xffffffffu, decided < 0);
  if (decided < 0) {
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(&work[1], (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    P.dist = __longlong_as_double((long long)e);
    PolylineWalker::prepare(P);
    paths[base + __popc(m & ((1u << lane) - 1u))] = P;
  }
}

__device__ __forceinline__ void finish_path(const EdgeArgs &A, bool valid, int64_t e, DubinsPath &P,
                                            double sx, double sy, double ex, double ey,
                                            DubinsPath *paths, unsigned long long *work) {
  if (A.obs.mask) finish_path_impl<true>(A, valid, e, P, sx, sy, ex, ey, paths, work);
  else finish_path_impl<false>(A, valid, e, P, sx, sy, ex, ey, paths, work);
}

// K4a: Edge::NewEdge + the candidate solve of CalculateTrajectory, one thread per edge, all
// six candidates in double (small batches; the ranked pipeline below is the fast path)
__global__ void __launch_bounds__(EDGE_T)
dubins_solve_kernel(EdgeArgs A, DubinsPath *paths, unsigned long long *work) {
  const int64_t e = (int64_t)blockIdx.x * EDGE_T + threadIdx.x;
  const bool valid = e < A.n;
  DubinsPath P;
  P.kinds = 0;
  double sx = 0, sy = 0, st = 0, ex = 0, ey = 0, et = 0;
  if (valid) {
    load_pose(A, e, false, sx, sy, st);
    load_pose(A, e, true, ex, ey, et);
    double edge_dist = metric_dist(A.metric, sx, sy, st, ex, ey, et);  // NewEdge :19
    dubins_solve(sx, sy, st, ex, ey, et, A.r_min, edge_dist, P);
  }
  finish_path(A, valid, e, P, sx, sy, ex, ey, paths, work);
}

// K4a, ranked: the double-precision solve of all six Dubins words costs ~5800 instructions per
// edge (22 atan2, 2 acos, 6 sincos) and 142 registers.  Here a float32 pass first lists the
// candidates that can be the shortest (dubins_rank: almost always exactly one), the edges are
// binned by that list (six single-candidate bins + one for everything else), and each bin is
// solved by a kernel specialised for its candidate (~70 registers, a sixth of the work).
static constexpr int RANK_BINS = 7;

__device__ __forceinline__ int rank_bin(int mask) {
  return (__popc(mask) == 1) ? (__ffs(mask) - 1) : (RANK_BINS - 1);
}

__global__ void __launch_bounds__(256)
dubins_rank_kernel(EdgeArgs A, uint8_t *masks, unsigned long long *bin_count) {
  __shared__ unsigned s_hist[RANK_BINS];
  if (threadIdx.x < RANK_BINS) s_hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < A.n) {
    double sx, sy, st, ex, ey, et;
    load_pose(A, e, false, sx, sy, st);
    load_pose(A, e, true, ex, ey, et);
    const int mask = dubins_rank(sx, sy, st, ex, ey, et, A.r_min);
    masks[e] = (uint8_t)mask;
    atomicAdd(&s_hist[rank_bin(mask)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < RANK_BINS && s_hist[threadIdx.x])
    atomicAdd(&bin_count[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
}

// bin_state: [0..7) counts, [8..16) exclusive starts (8 entries), [16..23) scatter cursors
__global__ void dubins_bin_scan_kernel(unsigned long long *bin_state) {
  if (threadIdx.x == 0) {
    unsigned long long acc = 0;
    for (int b = 0; b < RANK_BINS; b++) {
      bin_state[8 + b] = acc;
      bin_state[16 + b] = acc;
      acc += bin_state[b];
    }
    bin_state[8 + RANK_BINS] = acc;
  }
}

__global__ void __launch_bounds__(256)
dubins_bin_scatter_kernel(int64_t n, const uint8_t *masks, unsigned long long *bin_state, uint32_t *perm) {
  __shared__ unsigned s_hist[RANK_BINS];
  __shared__ unsigned long long s_base[RANK_BINS];
  if (threadIdx.x < RANK_BINS) s_hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int bin = 0;
  unsigned local = 0;
  if (e < n) {
    bin = rank_bin(masks[e]);
    local = atomicAdd(&s_hist[bin], 1u);
  }
  __syncthreads();
  if (threadIdx.x < RANK_BINS && s_hist[threadIdx.x])
    s_base[threadIdx.x] = atomicAdd(&bin_state[16 + threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
  __syncthreads();
  if (e < n) perm[s_base[bin] + local] = (uint32_t)e;
}

// MASK > 0: every edge of the bin has exactly that candidate list; MASK == 0: the list is read
// Six CTAs per SM (80 registers, 48 - 80 B of spills in the single-word instances) instead of the five
// that 92 registers allow: the solve is a long dependent FP64 chain per edge (issue slots 42 % busy at
// 20 warps per SM), and the extra warps are worth more than the spills cost -- 10 M edges of config 3
// 7.25 -> 6.96 ms; seven CTAs (72 registers, 170 B) 7.11, eight (64 registers, 240 B) 7.25.
template <int MASK>
__global__ void __launch_bounds__(EDGE_T, 6)
dubins_solve_bin_kernel(EdgeArgs A, DubinsPath *paths, const uint32_t *perm, const uint8_t *masks,
                        const unsigned long long *bin_state, int bin, unsigned long long *work) {
  const unsigned long long lo = bin_state[8 + bin], hi = bin_state[8 + bin + 1];
  // whole warps stay in the loop (finish_path appends with a warp-wide ballot)
  for (unsigned long long w0 = lo + (unsigned long long)blockIdx.x * EDGE_T + (threadIdx.x & ~31u); w0 < hi;
       w0 += (unsigned long long)gridDim.x * EDGE_T) {
    const unsigned long long i = w0 + (threadIdx.x & 31u);
    const bool valid = i < hi;
    int64_t e = 0;
    DubinsPath P;
    P.kinds = 0;
    double sx = 0, sy = 0, st = 0, ex = 0, ey = 0, et = 0;
    if (valid) {
      e = perm[i];
      load_pose(A, e, false, sx, sy, st);
      load_pose(A, e, true, ex, ey, et);
      double edge_dist = metric_dist(A.metric, sx, sy, st, ex, ey, et);  // NewEdge :19
      dubins_solve(sx, sy, st, ex, ey, et, A.r_min, edge_dist, P, MASK ? MASK : (int)masks[e]);
    }
    finish_path(A, valid, e, P, sx, sy, ex, ey, paths, work);
  }
}

// K4b: polyline walk + ExplicitEdgeCheck2D per segment, lanes refill from a global counter
template <bool MASKED>  // MASKED: only the obstacles of a subset take part (obstacle sweeps)
__global__ void __launch_bounds__(EDGE_T)
dubins_walk_kernel(EdgeArgs A, const DubinsPath *paths, unsigned long long *work) {
  const ObsView &O = A.obs;
  const int lane = threadIdx.x & 31;
  __shared__ WarpScratch s_ws[EDGE_T / 32];
  // Every lane works one edge ahead: while it walks an edge, the path record of the edge it
  // will walk next is already on its way into the lane's other slot (cp.async), so a lane that
  // finishes an edge never waits for global memory -- the records were written by the solve
  // kernels and have long left L2 (112 B x 10 M edges).
  __shared__ __align__(16) DubinsPath s_path[2][EDGE_T];
  WarpScratch *ws = &s_ws[threadIdx.x >> 5];
  // the undecided edges, listed by the solve kernels (finish_path); a record carries its edge id
  const long long n_work = (long long)work[1];
  unsigned long long *next_edge = &work[0];
  PolylineWalker W;
  W.part = 3;
  bool e = false;         // this lane is walking a record (it sits in s_path[cur])
  bool en = false;        // the record it walks next is in flight to, or has landed in, s_path[cur ^ 1]
  int cur = 0;
  bool exhausted = false; // no unclaimed edges left
  bool have_prev = false;
  double px = 0.0, py = 0.0;
  long long w_next = 0, w_end = 0;  // the warp's reserve of claimed edges (uniform)
  bool w_done = false;              // the global counter has run past the last edge
  for (;;) {
    if (!e && en) {
      asm volatile("cp.async.wait_all;" ::: "memory");
      e = true;
      en = false;
      cur ^= 1;
      W.start(&s_path[cur][threadIdx.x], A.r_min, true, true);
      have_prev = false;
    }
    // lanes without a next edge take one from the warp's reserve of claimed edges and start its
    // copy; the reserve is refilled WALK_CLAIM edges at a time, so the round trip of the global
    // atomic is paid once per WALK_CLAIM edges instead of in every trip
    unsigned need = __ballot_sync(0xffffffffu, !en && !exhausted);
    if (need) {
      if (w_next == w_end && !w_done) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(next_edge, (unsigned long long)WALK_CLAIM);
        base = __shfl_sync(0xffffffffu, base, 0);
        w_next = (long long)base;
        w_end = min((long long)base + WALK_CLAIM, n_work);
        if (w_next >= n_work) { w_done = true; w_end = w_next; }
      }
      const int avail = (int)(w_end - w_next);
      const int rank = __popc(need & ((1u << lane) - 1));
      if (!en && !exhausted) {
        if (rank < avail) {
          en = true;
          const char *src = reinterpret_cast<const char *>(paths + (w_next + rank));
          const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_path[cur ^ 1][threadIdx.x]);
#pragma unroll
          for (int k = 0; k < (int)sizeof(DubinsPath); k += 16)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + k), "l"(src + k) : "memory");
          asm volatile("cp.async.commit_group;" ::: "memory");
        } else if (w_done) {
          exhausted = true;
        }  // else: served from the next refill
      }
      w_next += min(__popc(need), avail);
    }
    if (__all_sync(0xffffffffu, !e && !en)) break;
    // Lanes run ahead: most segments have no grid entry in their cells, so a lane steps along its
    // polyline -- next point, cells of the new segment -- until it holds a segment that has something
    // to be tested against (or its path ends, or WALK_AHEAD steps have passed); then the warp tests the
    // waiting segments together.  The cooperative machinery runs once per round instead of once per
    // point.  Measured per 10 M edges of config 3: 7.62 ms with one step, 7.43 with two, 7.51 with
    // three, 7.89 / 8.82 / 10.87 with 4 / 8 / 16 -- a lane whose path has ended is refilled only
    // between rounds, and lanes that wait for the others' steps idle.  A lane tests its segments in
    // path order and stops at its first hit, as before: verdicts are unchanged.
    double cx = 0.0, cy = 0.0;
    bool ended = false, pending = false, wide = false;
    int total = 0;
    if (A.grid.nx == 0) {  // no grid (degenerate table): one point per round, per-lane loop over all obstacles
      bool test = false, hit1 = false;
      if (e) {
        bool connects = false;
        if (!W.next(&cx, &cy, &connects)) ended = true;
        else test = have_prev && connects;
      }
      if (test) hit1 = segment_vs_obstacles<MASKED>(O, A.grid, px, py, cx, cy, A.robot_radius);
      if (e) {
        if (ended || hit1) {
          A.verdict[__double_as_longlong(s_path[cur][threadIdx.x].dist)] = hit1 ? 1 : 0;
          e = false;
        } else { px = cx; py = cy; have_prev = true; }
      }
      continue;
    }
    for (int step = 0; step < WALK_AHEAD; step++) {
      const bool stepping = e && !pending && !ended;
      if (!__any_sync(0xffffffffu, stepping)) break;
      if (stepping) {
        bool connects = false;
        if (!W.next(&cx, &cy, &connects)) {
          ended = true;  // end of path, nothing collided (dubinsedge.cpp:897)
        } else {
          if (have_prev && connects) {
            total = lane_segment_cells(A.grid, px, py, cx, cy, A.robot_radius, ws, lane, wide);
            pending = total > 0 || wide;
          }
          if (!pending) { px = cx; py = cy; have_prev = true; }  // nothing near this segment
        }
      }
    }
    bool hit = false;
    if (__any_sync(0xffffffffu, pending))
      hit = warp_pending_vs_obstacles<MASKED>(O, A.grid, pending ? total : 0, pending && wide, A.robot_radius, ws, lane);
    if (e) {
      if (ended || hit) {  // :878-889 first hit decides the edge
        A.verdict[__double_as_longlong(s_path[cur][threadIdx.x].dist)] = hit ? 1 : 0;
        e = false;
      } else if (pending) {
        px = cx;
        py = cy;
        have_prev = true;
      }
    }
  }
}

// K5: LineCheck (obstacle.cpp:1098-1133), one thread per edge.
//
// The reference cuts the straight edge into 50 collinear pieces (repeated +-dist/50, :1108-1133)
// and tests each piece against every obstacle.  The pieces tile the edge (to within ~1e-13 of
// accumulated rounding), so  min_i SegmentDistSqrd(piece_i, Q) = SegmentDistSqrd(edge, Q)  up to
// that rounding, and the OR over the pieces is decided by ONE exact test of the whole edge per
// nearby polygon edge / ball -- unless the distance lies within eps (1e-9 relative) of the
// threshold, or the obstacle's bounding circle (obstacle.cpp:765-779) is not known to contain
// its polygon: those edges take the reference's 50-piece loop verbatim.  A polygon edge within
// radius - eps of the edge is within radius of one piece, and that piece is then within
// radius + O.radius_ of the origin when O.radius_ covers the vertices (`covers`), so the piece
// passes the bounding reject as well.
template <bool MASKED>
__device__ __forceinline__ int line_screen(const ObsView &O, const ObstacleGrid &G, double sx, double sy,
                                           double ex, double ey, double radius) {
  if (G.nx == 0) return 2;
  const double eps = 1e-9 * (1.0 + fmax(fmax(fabs(sx), fabs(sy)), fmax(fabs(ex), fabs(ey))));
  if (!(eps < 0.25 * radius)) return 2;
  const double r_far = (radius + eps) * (radius + eps), r_near = (radius - eps) * (radius - eps);
  const double len = sqrt((ex - sx) * (ex - sx) + (ey - sy) * (ey - sy));
  if (!(len == len) || !(len < 1.0e30)) return 2;
  // walk the edge in stretches of about four cells, looking at the cells each stretch's box touches
  const double cell = 1.0 / G.inv_cell;
  const int pieces = (int)fmin(fmax(ceil(len / (4.0 * cell)), 1.0), 4096.0);
  int state = 0;
  int pcx0 = 1, pcx1 = 0, pcy0 = 1, pcy1 = 0;  // cell box of the previous stretch (empty)
  for (int k = 0; k < pieces; k++) {
    const double t0 = (double)k / pieces, t1 = (double)(k + 1) / pieces;
    const double ax = sx + (ex - sx) * t0, ay = sy + (ey - sy) * t0;
    const double bx = sx + (ex - sx) * t1, by = sy + (ey - sy) * t1;
    const double lox = fmin(ax, bx) - eps, hix = fmax(ax, bx) + eps;
    const double loy = fmin(ay, by) - eps, hiy = fmax(ay, by) + eps;
    const double fx0 = floor((lox - G.x0) * G.inv_cell), fx1 = floor((hix - G.x0) * G.inv_cell);
    const double fy0 = floor((loy - G.y0) * G.inv_cell), fy1 = floor((hiy - G.y0) * G.inv_cell);
    if (fx1 < 0.0 || fy1 < 0.0 || fx0 > (double)(G.nx - 1) || fy0 > (double)(G.ny - 1)) continue;
    const int cx0 = (int)fmax(fx0, 0.0), cx1 = (int)fmin(fx1, (double)(G.nx - 1));
    const int cy0 = (int)fmax(fy0, 0.0), cy1 = (int)fmin(fy1, (double)(G.ny - 1));
    for (int cx = cx0; cx <= cx1; cx++)
      for (int cy = cy0; cy <= cy1; cy++) {
        if (cx >= pcx0 && cx <= pcx1 && cy >= pcy0 && cy <= pcy1) continue;  // seen with the last stretch
        const int c = cx * G.ny + cy;
        for (int q = G.off[c]; q < G.off[c + 1]; q++) {
          const int j = G.idx[q];
          if (j >= G.n_edges) {  // ball: obstacle.cpp:765-781
            const int b = G.ball[j - G.n_edges];
            if (MASKED && !O.mask[b]) continue;
            const double d2 = dist_sqrd_point_segment(O.ox[b], O.oy[b], sx, sy, ex, ey);
            const double hi = radius + O.rad[b] + eps, lo = radius + O.rad[b] - eps;
            if (d2 > hi * hi) continue;
            if (d2 < lo * lo && lo > 0.0) return 1;
            state = 2;
            continue;
          }
          const int o = G.eobs[j];
          if (MASKED && !O.mask[o]) continue;
          double qax, qay, qbx, qby;
          load_edge(G, j, qax, qay, qbx, qby);
          const double sd = segment_dist_sqrd(sx, sy, ex, ey, qax, qay, qbx, qby);
          if (sd > r_far) continue;
          if (sd < r_near && O.covers[o]) return 1;
          state = 2;
        }
      }
    pcx0 = cx0; pcx1 = cx1; pcy0 = cy0; pcy1 = cy1;
  }
  return state;
}

template <bool MASKED>
// eight CTAs per SM (64 registers): 0.784 -> 0.742 ms per 10 M straight edges (four 1.01, six 0.787, ten 0.80)
#ifndef DRRT_LINE_MINB
#define DRRT_LINE_MINB 8
#endif
__global__ void __launch_bounds__(EDGE_T, DRRT_LINE_MINB)
linecheck_kernel(EdgeArgs A) {
  const ObsView &O = A.obs;
  const int64_t e = (int64_t)blockIdx.x * EDGE_T + threadIdx.x;
  if (e >= A.n) return;
  double sx, sy, st, ex, ey, et;
  load_pose(A, e, false, sx, sy, st);
  load_pose(A, e, true, ex, ey, et);
  const double th = atan2(ey - sy, ex - sx);
  const double dist = metric_dist(A.metric, sx, sy, th, ex, ey, th);
  bool hit = false;
  if (O.n > 0) {
    const int state = line_screen<MASKED>(O, A.grid, sx, sy, ex, ey, A.robot_radius);
    hit = state == 1;
    if (state == 2) {
      // too close to call from the whole edge: the reference's 50 pieces
      const double traj_length = 50;
      double x_val = sx, y_val = sy;
      const double x_dist = fabs(ex - x_val), y_dist = fabs(ey - y_val);
      double px = x_val, py = y_val;
      for (int i = 0; i < 50 && !hit; i++) {
        if (ex > sx) x_val += x_dist / traj_length; else x_val -= x_dist / traj_length;
        if (ey > sy) y_val += y_dist / traj_length; else y_val -= y_dist / traj_length;
        hit = segment_vs_obstacles<MASKED>(O, A.grid, px, py, x_val, y_val, A.robot_radius);
        px = x_val;
        py = y_val;
      }
    }
  }
  A.verdict[e] = hit ? 1 : 0;
  if (A.length) A.length[e] = dist;
}

// distance from an axis-aligned rectangle to a segment (0 when they touch)
static double rect_segment_dist(double rx0, double ry0, double rx1, double ry1, double ax, double ay,
                                double bx, double by) {
  // Liang-Barsky: does the segment enter the rectangle?
  double t0 = 0.0, t1 = 1.0, dx = bx - ax, dy = by - ay;
  const double p[4] = {-dx, dx, -dy, dy}, q[4] = {ax - rx0, rx1 - ax, ay - ry0, ry1 - ay};
  bool inside = true;
  for (int k = 0; k < 4 && inside; k++) {
    if (p[k] == 0.0) { if (q[k] < 0.0) inside = false; }
    else {
      double r = q[k] / p[k];
      if (p[k] < 0.0) { if (r > t1) inside = false; else if (r > t0) t0 = r; }
      else { if (r < t0) inside = false; else if (r < t1) t1 = r; }
    }
  }
  if (inside) return 0.0;
  auto pt_seg = [](double px, double py, double sx, double sy, double ex, double ey) {
    double vx = px - sx, vy = py - sy, ux = ex - sx, uy = ey - sy;
    double det = vx * ux + vy * uy, len = ux * ux + uy * uy;
    double t = len > 0.0 ? std::min(1.0, std::max(0.0, det / len)) : 0.0;
    double cx = sx + t * ux - px, cy = sy + t * uy - py;
    return sqrt(cx * cx + cy * cy);
  };
  auto pt_rect = [&](double px, double py) {
    double ex = std::max(std::max(rx0 - px, px - rx1), 0.0), ey = std::max(std::max(ry0 - py, py - ry1), 0.0);
    return sqrt(ex * ex + ey * ey);
  };
  double d = std::min(pt_rect(ax, ay), pt_rect(bx, by));
  d = std::min(d, pt_seg(rx0, ry0, ax, ay, bx, by));
  d = std::min(d, pt_seg(rx1, ry0, ax, ay, bx, by));
  d = std::min(d, pt_seg(rx0, ry1, ax, ay, bx, by));
  d = std::min(d, pt_seg(rx1, ry1, ax, ay, bx, by));
  return d;
}

// Edge grid over the active obstacles (host build; a few thousand edges).  A cell
// lists every polygon edge within pad = radius(1+1e-6)+1e-6 of the cell rectangle,
// and every ball whose reach circle touches it.
static int ensure_obstacle_grid(Store *S, double robot_radius, cudaStream_t st) {
  if (S->ogrid_radius == robot_radius) return DRRT_OK;
  const size_t n = S->h_o_ox.size();
  ObstacleGrid G{};
  auto disable = [&]() {  // nx == 0: kernels test every obstacle
    S->ogrid = G;
    S->ogrid_radius = robot_radius;
    return DRRT_OK;
  };
  if (n == 0) return disable();
  std::vector<double> eax, eay, ebx, eby;
  std::vector<int32_t> eobs, ball;
  for (size_t o = 0; o < n; o++) {
    if (S->h_o_kind[o] == 1) { ball.push_back((int32_t)o); continue; }
    int v0 = S->h_o_voff[o], v1 = S->h_o_voff[o + 1];
    if (v1 - v0 < 2) continue;  // obstacle.cpp:785
    double ax = S->h_o_vx[v1 - 1], ay = S->h_o_vy[v1 - 1];
    for (int v = v0; v < v1; v++) {  // obstacle.cpp:788-804: (last,first), (0,1), ...
      eax.push_back(ax); eay.push_back(ay);
      ebx.push_back(S->h_o_vx[v]); eby.push_back(S->h_o_vy[v]);
      eobs.push_back((int32_t)o);
      ax = S->h_o_vx[v]; ay = S->h_o_vy[v];
    }
  }
  const size_t ne = eax.size(), nb = ball.size();
  if (ne + nb == 0 || ne + nb > 100000000) return disable();
  const double pad = fabs(robot_radius) * (1.0 + 1e-6) + 1e-6;
  double lox = 1e300, loy = 1e300, hix = -1e300, hiy = -1e300;
  for (size_t j = 0; j < ne; j++) {
    lox = std::min(lox, std::min(eax[j], ebx[j]) - pad); hix = std::max(hix, std::max(eax[j], ebx[j]) + pad);
    loy = std::min(loy, std::min(eay[j], eby[j]) - pad); hiy = std::max(hiy, std::max(eay[j], eby[j]) + pad);
  }
  std::vector<double> reach(nb);
  for (size_t k = 0; k < nb; k++) {
    int o = ball[k];
    reach[k] = (fabs(robot_radius) + fabs(S->h_o_rad[o])) * (1.0 + 1e-6) + 1e-6;
    lox = std::min(lox, S->h_o_ox[o] - reach[k]); hix = std::max(hix, S->h_o_ox[o] + reach[k]);
    loy = std::min(loy, S->h_o_oy[o] - reach[k]); hiy = std::max(hiy, S->h_o_oy[o] + reach[k]);
  }
  if (!std::isfinite(lox + loy + hix + hiy)) return disable();
  double extent = std::max(std::max(hix - lox, hiy - loy), 1e-9);
  double cell = std::max(extent / 160.0, pad);
  G.x0 = lox; G.y0 = loy; G.inv_cell = 1.0 / cell;
  G.nx = std::max(1, (int)ceil((hix - lox) / cell) + 1);
  G.ny = std::max(1, (int)ceil((hiy - loy) / cell) + 1);
  G.n_edges = (int32_t)ne;
  const int ncell = G.nx * G.ny;
  std::vector<int32_t> off(ncell + 1, 0), idx;
  for (int pass = 0; pass < 2; pass++) {
    std::vector<int32_t> cur(off.begin(), off.end() - 1);
    auto visit = [&](double bx0, double by0, double bx1, double by1, int entry, double r,
                     const double *seg) {
      int cx0 = std::max(0, (int)floor((bx0 - lox) / cell) - 1), cx1 = std::min(G.nx - 1, (int)floor((bx1 - lox) / cell) + 1);
      int cy0 = std::max(0, (int)floor((by0 - loy) / cell) - 1), cy1 = std::min(G.ny - 1, (int)floor((by1 - loy) / cell) + 1);
      for (int cx = cx0; cx <= cx1; cx++)
        for (int cy = cy0; cy <= cy1; cy++) {
          // the (slightly grown) cell rectangle; device cell = floor((x - x0) * inv_cell)
          double rx0 = lox + (cx - 1e-6) * cell, rx1 = lox + (cx + 1.000001) * cell;
          double ry0 = loy + (cy - 1e-6) * cell, ry1 = loy + (cy + 1.000001) * cell;
          double d = seg ? rect_segment_dist(rx0, ry0, rx1, ry1, seg[0], seg[1], seg[2], seg[3])
                         : rect_segment_dist(rx0, ry0, rx1, ry1, bx0 + r, by0 + r, bx0 + r, by0 + r);
          if (d > r) continue;
          int c = cx * G.ny + cy;
          if (pass == 0) off[c + 1]++;
          else idx[cur[c]++] = entry;
        }
    };
    for (size_t j = 0; j < ne; j++) {
      double seg[4] = {eax[j], eay[j], ebx[j], eby[j]};
      visit(std::min(eax[j], ebx[j]) - pad, std::min(eay[j], eby[j]) - pad,
            std::max(eax[j], ebx[j]) + pad, std::max(eay[j], eby[j]) + pad, (int)j, pad, seg);
    }
    for (size_t k = 0; k < nb; k++) {
      int o = ball[k];
      visit(S->h_o_ox[o] - reach[k], S->h_o_oy[o] - reach[k], S->h_o_ox[o] + reach[k],
            S->h_o_oy[o] + reach[k], (int)(ne + k), reach[k], nullptr);
    }
    if (pass == 0) {
      for (int c = 0; c < ncell; c++) off[c + 1] += off[c];
      idx.resize(off[ncell] + 1);
    }
  }
  DRRT_CUDA(cudaStreamSynchronize(st));
  auto up = [&](auto &buf, const auto &v) -> int {
    DRRT_CHECK(buf.reserve(v.size() + 1, st));
    if (!v.empty())
      DRRT_CUDA(cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice, st));
    return DRRT_OK;
  };
  DRRT_CHECK(up(S->og_off, off)); DRRT_CHECK(up(S->og_idx, idx));
  DRRT_CHECK(up(S->og_eobs, eobs)); DRRT_CHECK(up(S->og_ball, ball));
  std::vector<EdgeRec> erec(ne);
  for (size_t j = 0; j < ne; j++) erec[j] = EdgeRec{eax[j], eay[j], ebx[j], eby[j]};
  DRRT_CHECK(up(S->og_erec, erec));
  DRRT_CUDA(cudaStreamSynchronize(st));
  G.off = S->og_off.p; G.idx = S->og_idx.p;
  G.eobs = S->og_eobs.p; G.ball = S->og_ball.p; G.erec = S->og_erec.p;
  S->ogrid = G;
  S->ogrid_radius = robot_radius;
  return DRRT_OK;
}

int edgecheck_batch_dev(Store *S, int64_t n, const double *start_dev, const double *end_dev,
                        const int32_t *sid_dev, const int32_t *eid_dev, int edge_kind,
                        double robot_radius, double r_min, uint8_t *verdict_dev,
                        double *length_dev, cudaStream_t st) {
  if (n < 0) return set_error(DRRT_ERR_INVALID, "edgecheck: n < 0");
  if (edge_kind != DRRT_EDGE_DUBINS && edge_kind != DRRT_EDGE_STRAIGHT)
    return set_error(DRRT_ERR_INVALID, "edgecheck: edge_kind");
  bool by_id = sid_dev && eid_dev;
  const bool csr = S->csr_q != nullptr;  // neighbour-list mode, set by extend_batch_dev
  if (!csr && !by_id && !(start_dev && end_dev)) return set_error(DRRT_ERR_INVALID, "edgecheck: endpoints");
  if (!verdict_dev) return set_error(DRRT_ERR_INVALID, "edgecheck: verdict");
  if (n == 0) return DRRT_OK;
  EdgeArgs A{};
  A.n = n; A.start = by_id ? nullptr : start_dev; A.end = by_id ? nullptr : end_dev;
  A.sid = by_id ? sid_dev : nullptr; A.eid = by_id ? eid_dev : nullptr;
  A.sx = S->x.p; A.sy = S->y.p; A.sth = S->th.p;
  if (csr) {
    A.csr_off = S->csr_off; A.csr_ids = S->csr_ids; A.csr_q = S->csr_q; A.csr_nq = S->csr_nq;
    A.csr_row = S->x_row.p;
  }
  A.metric = S->metric; A.edge_kind = edge_kind;
  A.robot_radius = robot_radius; A.r_min = r_min;
  A.obs.n = S->has_obstacles ? S->obs.n_active : 0;
  A.obs.mask = S->edge_mask;
  A.obs.covers = S->o_covers.p;
  A.obs.kind = S->obs.kind; A.obs.ox = S->obs.ox; A.obs.oy = S->obs.oy; A.obs.rad = S->obs.rad;
  A.obs.voff = S->obs.vert_off; A.obs.vx = S->obs.vx; A.obs.vy = S->obs.vy;
  A.verdict = verdict_dev; A.length = length_dev;
  const size_t smem = 0;
  if (A.obs.n > 0) {
    DRRT_CHECK(ensure_obstacle_grid(S, robot_radius, st));
    A.grid = S->ogrid;
  }
  const bool masked = A.obs.mask != nullptr;
  unsigned nb = (unsigned)((n + EDGE_T - 1) / EDGE_T);
  if (edge_kind == DRRT_EDGE_STRAIGHT) {
    if (masked) linecheck_kernel<true><<<nb, EDGE_T, smem, st>>>(A);
    else linecheck_kernel<false><<<nb, EDGE_T, smem, st>>>(A);
    DRRT_LAUNCHED();
  } else {
    if (st != S->stream) DRRT_CUDA(cudaStreamSynchronize(S->stream));
    DRRT_CHECK(S->edge_paths.reserve((size_t)n * sizeof(DubinsPath), st));
    DubinsPath *paths = reinterpret_cast<DubinsPath *>(S->edge_paths.p);
    // [0] claim counter of the walk, [1] number of undecided edges listed by the solve kernels
    unsigned long long *work = S->d_edge_counter;
    DRRT_CUDA(cudaMemsetAsync(work, 0, 2 * sizeof(unsigned long long), st));
    if (n < RANKED_SOLVE_MIN || n >= 0x7fffffff) {
      dubins_solve_kernel<<<nb, EDGE_T, 0, st>>>(A, paths, work);
      DRRT_LAUNCHED();
    } else {
      DRRT_CHECK(S->edge_masks.reserve((size_t)n, st));
      DRRT_CHECK(S->edge_perm.reserve((size_t)n, st));
      DRRT_CHECK(S->edge_bins.reserve(32, st));
      unsigned long long *bins = S->edge_bins.p;
      DRRT_CUDA(cudaMemsetAsync(bins, 0, 32 * sizeof(unsigned long long), st));
      const unsigned nb256 = (unsigned)((n + 255) / 256);
      dubins_rank_kernel<<<nb256, 256, 0, st>>>(A, S->edge_masks.p, bins);
      dubins_bin_scan_kernel<<<1, 32, 0, st>>>(bins);
      dubins_bin_scatter_kernel<<<nb256, 256, 0, st>>>(n, S->edge_masks.p, bins, S->edge_perm.p);
      // one specialised solve per single-candidate bin, a generic one for the rest; persistent
      // grids sized for the machine (a bin's extent is only known on the device)
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, S->device);
      const unsigned g = (unsigned)std::min<int64_t>(nb, (int64_t)sms * 12);
      // Two words (RSR, LSL) hold nine edges in ten of a planner's batch and take ~0.7 ms each; the
      // other five bins are small and, launched one after the other, spend most of their ~0.4 ms
      // filling and draining the machine.  They run beside the large ones on two side streams (the
      // bins write disjoint path records and append to the work list atomically).
      if (!S->side_stream[0]) {
        for (int b = 0; b < 2; b++) {
          DRRT_CUDA(cudaStreamCreateWithFlags(&S->side_stream[b], cudaStreamNonBlocking));
          DRRT_CUDA(cudaEventCreateWithFlags(&S->ev_join[b], cudaEventDisableTiming));
        }
        DRRT_CUDA(cudaEventCreateWithFlags(&S->ev_fork, cudaEventDisableTiming));
      }
      cudaStream_t s0 = S->side_stream[0], s1 = S->side_stream[1];
      DRRT_CUDA(cudaEventRecord(S->ev_fork, st));
      DRRT_CUDA(cudaStreamWaitEvent(s0, S->ev_fork, 0));
      DRRT_CUDA(cudaStreamWaitEvent(s1, S->ev_fork, 0));
      dubins_solve_bin_kernel<CAND_RSR><<<g, EDGE_T, 0, st>>>(A, paths, S->edge_perm.p, S->edge_masks.p, bins, 1, work);
      dubins_solve_bin_kernel<CAND_LSR><<<g, EDGE_T, 0, s0>>>(A, paths, S->edge_perm.p, S->edge_masks.p, bins, 3, work);
      dubins_solve_bin_kernel<0><<<g, EDGE_T, 0, s1>>>(A, paths, S->edge_perm.p, S->edge_masks.p, bins, 6, work);
      dubins_solve_bin_kernel<CAND_RSL><<<g, EDGE_T, 0, s0>>>(A, paths, S->edge_perm.p, S->edge_masks.p, bins, 0, work);
      dubins_solve_bin_kernel<CAND_LRL><<<g, EDGE_T, 0, s1>>>(A, paths, S->edge_perm.p, S->edge_masks.p, bins, 5, work);
      dubins_solve_bin_kernel<CAND_LSL><<<g, EDGE_T, 0, st>>>(A, paths, S->edge_perm.p, S->edge_masks.p, bins, 4, work);
      dubins_solve_bin_kernel<CAND_RLR><<<g, EDGE_T, 0, s1>>>(A, paths, S->edge_perm.p, S->edge_masks.p, bins, 2, work);
      DRRT_CUDA(cudaEventRecord(S->ev_join[0], s0));
      DRRT_CUDA(cudaEventRecord(S->ev_join[1], s1));
      DRRT_CUDA(cudaStreamWaitEvent(st, S->ev_join[0], 0));
      DRRT_CUDA(cudaStreamWaitEvent(st, S->ev_join[1], 0));
      for (int k = 0; k < 10; k++) DRRT_LAUNCHED();
    }
    if (A.obs.n == 0) {
      DRRT_CUDA(cudaMemsetAsync(verdict_dev, 0, (size_t)n, st));  // nothing to collide with
    } else {
      // persistent grid: enough CTAs to fill the machine, lanes pull edges until none are left
      int per_sm = 0;
      if (masked)
        DRRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dubins_walk_kernel<true>, EDGE_T, smem));
      else
        DRRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dubins_walk_kernel<false>, EDGE_T, smem));
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, S->device);
      unsigned grid = (unsigned)std::min<int64_t>(nb, (int64_t)std::max(per_sm, 1) * sms);
      if (masked) dubins_walk_kernel<true><<<grid, EDGE_T, smem, st>>>(A, paths, work);
      else dubins_walk_kernel<false><<<grid, EDGE_T, smem, st>>>(A, paths, work);
      DRRT_LAUNCHED();
    }
  }
  DRRT_CUDA(cudaGetLastError());
  return DRRT_OK;
}

// ---------------------------------------------------------------------------
// Fused Extend step: for every neighbour of every new node both Dubins
// edges are checked straight off the range result, and FindBestParent's argmin
// (src/drrt.cpp:388-459: the safe new->near edge minimising near.rrt_LMC_ + edge.dist_) is
// reduced per row on the device.
__global__ void best_parent_kernel(int64_t nq, const int64_t *off, const int32_t *ids,
                                   const uint8_t *unsafe, const double *len, const double *lmc,
                                   int32_t *best, double *best_cost) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= nq) return;
  double c = DRRT_INF;  // new_node->rrt_LMC_ = INF (drrt.cpp:391)
  long long at = -1;
  for (int64_t p = off[row] + lane; p < off[row + 1]; p += 32) {
    if (unsafe[2 * p]) continue;               // :422-431
    double v = lmc[ids[p]] + len[2 * p];       // :444-446
    if (c > v) { c = v; at = p; }
  }
  for (int o = 16; o > 0; o >>= 1) {  // ties: the neighbour listed first (smallest id)
    double oc = __shfl_xor_sync(0xffffffffu, c, o);
    long long oa = __shfl_xor_sync(0xffffffffu, at, o);
    if (oa >= 0 && (at < 0 || oc < c || (oc == c && oa < at))) { c = oc; at = oa; }
  }
  if (lane == 0) {
    best[row] = at >= 0 ? ids[at] : -1;
    best_cost[row] = c;
  }
}

__global__ void fill_rows_kernel(int64_t nq, const int64_t *off, int32_t *row) {
  for (int64_t q = blockIdx.x; q < nq; q += gridDim.x)
    for (int64_t p = off[q] + threadIdx.x; p < off[q + 1]; p += blockDim.x) row[p] = (int32_t)q;
}

int extend_batch_dev(Store *S, int64_t nq, const double *q_dev, const double *radius_dev,
                     double radius_all, double robot_radius, double r_min, const double *lmc_dev,
                     cudaStream_t st, drrt_range_result *rr, uint8_t **unsafe_dev, double **len_dev,
                     int32_t *best_dev, double *best_cost_dev) {
  DRRT_CHECK(range_batch_dev(S, nq, q_dev, radius_dev, radius_all, st, rr));
  // the pair numbering below (2 edges per neighbour) runs over a plain CSR
  DRRT_CHECK(range_pack_dev(S, st, rr));
  S->csr_off = rr->offsets_dev;
  S->csr_ids = rr->ids_dev;
  S->csr_dist = rr->dist_dev;
  const int64_t total = rr->total;
  DRRT_CHECK(S->x_unsafe.reserve((size_t)(2 * total + 1), st));
  DRRT_CHECK(S->x_len.reserve((size_t)(2 * total + 1), st));
  *unsafe_dev = S->x_unsafe.p;
  *len_dev = S->x_len.p;
  if (total > 0) {
    // row index of every pair, written once (the edge kernels would otherwise search the offsets
    // for every pose they load)
    DRRT_CHECK(S->x_row.reserve((size_t)total + 1, st));
    fill_rows_kernel<<<(unsigned)std::min<int64_t>(nq, 148 * 16), 256, 0, st>>>(nq, rr->offsets_dev, S->x_row.p);
    DRRT_LAUNCHED();
    S->csr_q = q_dev;
    S->csr_nq = nq;
    int rc = edgecheck_batch_dev(S, 2 * total, nullptr, nullptr, nullptr, nullptr, DRRT_EDGE_DUBINS,
                                 robot_radius, r_min, S->x_unsafe.p, S->x_len.p, st);
    S->csr_q = nullptr;
    DRRT_CHECK(rc);
  }
  S->ext_lmc = nullptr; S->ext_best = nullptr; S->ext_cost = nullptr; S->rewire_out = -1;
  if (best_dev && nq > 0) {
    if (!lmc_dev) return set_error(DRRT_ERR_INVALID, "extend: best parent needs the lmc array");
    S->ext_lmc = lmc_dev; S->ext_best = best_dev; S->ext_cost = best_cost_dev;  // for drrt_extend_rewire
    best_parent_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(nq, rr->offsets_dev, rr->ids_dev,
                                                                 S->x_unsafe.p, S->x_len.p, lmc_dev,
                                                                 best_dev, best_cost_dev);
    DRRT_LAUNCHED();
    DRRT_CUDA(cudaGetLastError());
  }
  return DRRT_OK;
}

// ---------------------------------------------------------------------------
// edge->trajectory_ (dubinsedge.cpp:808-826), for the robot mover
__global__ void trajectory_kernel(int64_t n, const double *start, const double *end, double r_min,
                                  int metric, double *xy, int32_t *npts, char *type,
                                  double *length) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const double *s = start + 3 * e, *g = end + 3 * e;
  DubinsPath P;
  double edge_dist = metric_dist(metric, s[0], s[1], s[2], g[0], g[1], g[2]);
  dubins_solve(s[0], s[1], s[2], g[0], g[1], g[2], r_min, edge_dist, P);
  if (xy) {
    PolylineWalker W;
    W.start(&P, r_min);
    double x, y;
    int k = 0;
    double *out = xy + (size_t)e * DRRT_TRAJ_MAX * 2;
    while (k < DRRT_TRAJ_MAX && W.next(&x, &y)) {
      out[2 * k] = x;
      out[2 * k + 1] = y;
      k++;
    }
    npts[e] = k;
  }
  if (length) length[e] = P.dist;
  if (type) {
    const char *names = "xxx\0rsl\0rsr\0rlr\0lsr\0lsl\0lrl\0" "0s0\0";
    for (int c = 0; c < 4; c++) type[4 * e + c] = names[4 * P.type + c];
  }
}

int trajectory_batch_dev(Store *S, int64_t n, const double *start_dev, const double *end_dev,
                         double r_min, double *xy_dev, int32_t *npts_dev, char *type_dev,
                         double *length_dev, cudaStream_t st) {
  if (n < 0 || !start_dev || !end_dev || (xy_dev && !npts_dev))
    return set_error(DRRT_ERR_INVALID, "trajectory: arguments");
  if (n == 0) return DRRT_OK;
  trajectory_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, start_dev, end_dev, r_min,
                                                                 S->metric, xy_dev, npts_dev,
                                                                 type_dev, length_dev);
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaGetLastError());
  return DRRT_OK;
}

// ---------------------------------------------------------------------------
// ExplicitPointCheck obstacle.cpp:1062-1085 = QuickCheck (:967-982) over all
// obstacles, then ExplicitPointCheck2D (:984-1060) over all obstacles.
__global__ void pointcheck_kernel(int64_t n, const double *pts, double radius, double min_distance,
                                  int metric, ObstacleTable T, uint8_t *verdict) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double px = pts[3 * i], py = pts[3 * i + 1], pt = pts[3 * i + 2];
  bool hit = false;
  for (int o = 0; o < T.n_all && !hit; o++) {  // QuickCheck2D :925-965
    if (!T.a_live[o]) continue;
    int kind = T.a_kind[o];
    if (1 <= kind && kind <= 5) {
      double dx = T.a_ox[o] - px, dy = T.a_oy[o] - py;
      if (sqrt(dx * dx + dy * dy) > T.a_rad[o]) continue;
    }
    if (kind == 1) hit = true;
    else if (kind == 3) {
      int v0 = T.a_vert_off[o], nv = T.a_vert_off[o + 1] - v0;
      hit = point_in_polygon(px, py, T.a_vx + v0, T.a_vy + v0, nv);
    }
  }
  for (int o = 0; o < T.n_all && !hit; o++) {  // ExplicitPointCheck2D :984-1060
    if (!T.a_live[o]) continue;
    int kind = T.a_kind[o];
    double this_distance = DRRT_INF;
    if (1 <= kind && kind <= 5) {
      // O->origin_(2) = point(2) (:1000): the metric's third term vanishes
      this_distance = metric_dist(metric, T.a_ox[o], T.a_oy[o], pt, px, py, pt) - radius;
      if (this_distance - T.a_rad[o] > min_distance) continue;
    }
    if (kind == 1) {
      this_distance = this_distance - T.a_rad[o];
      if (this_distance < 0.0) hit = true;
    } else if (kind == 3) {
      int v0 = T.a_vert_off[o], nv = T.a_vert_off[o + 1] - v0;
      if (point_in_polygon(px, py, T.a_vx + v0, T.a_vy + v0, nv)) hit = true;
      else if (nv > 0) {
        this_distance = sqrt(dist_to_polygon_sqrd(px, py, T.a_vx + v0, T.a_vy + v0, nv)) - radius;
        if (this_distance < 0.0) hit = true;
      }
    }
  }
  verdict[i] = hit ? 1 : 0;
}

int pointcheck_batch_dev(Store *S, int64_t n, const double *pts_dev, double robot_radius,
                         double collision_distance, uint8_t *verdict_dev, cudaStream_t st) {
  if (n < 0 || !pts_dev || !verdict_dev) return set_error(DRRT_ERR_INVALID, "pointcheck: arguments");
  if (n == 0) return DRRT_OK;
  ObstacleTable T = S->obs;
  if (!S->has_obstacles) T.n_all = 0;
  pointcheck_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, pts_dev, robot_radius,
                                                                 collision_distance, S->metric, T,
                                                                 verdict_dev);
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaGetLastError());
  return DRRT_OK;
}

}  // namespace drrt

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
// One thread per edge: solve the six Dubins candidates, then walk the 0.1-rad
// polyline once.  Each segment looks up the cells of a uniform obstacle grid
// its bounding box touches; a cell lists the obstacles whose reach circle
// (radius_ + robot radius) touches the cell, so an obstacle that is not listed
// fails the reference's own bounding reject (:765-779) for that segment and
// can be skipped exactly.  Listed obstacles get the reference's tests verbatim;
// inside the polygon loop an edge farther from the segment's first point than
// (radius + segment length) cannot bring SegmentDistSqrd under radius^2 and is
// skipped.  The obstacle table is staged in shared memory when it fits.
#include <math.h>

#include <algorithm>

#include "geom.cuh"
#include "store.cuh"

namespace drrt {

static constexpr int EDGE_T = 128;

struct ObsView {
  int n;
  const int32_t *kind;
  const double *ox, *oy, *rad;
  const int32_t *voff;
  const double *vx, *vy;
};

// ExplicitEdgeCheck2D for one active obstacle (kind 1 or 3).  far2 =
// ((radius + |p1-p0|) * (1+1e-9) + 1e-9)^2 : polygon edges farther than that
// from p0 are at more than `radius` from every point of the segment.
__device__ __forceinline__ bool segment_hits(const ObsView &O, int o, double p0x, double p0y,
                                             double p1x, double p1y, double radius, double far2) {
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
    for (int cx = cx0; cx <= cx1; cx++)
      for (int cy = cy0; cy <= cy1; cy++) {
        int c = cx * G.ny + cy;
        for (int e = G.off[c]; e < G.off[c + 1]; e++)
          if (segment_hits(O, G.idx[e], p0x, p0y, p1x, p1y, radius, far2)) return true;
      }
    return false;
  }
  for (int o = 0; o < O.n; o++)
    if (segment_hits(O, o, p0x, p0y, p1x, p1y, radius, far2)) return true;
  return false;
}

struct EdgeArgs {
  int64_t n;
  const double *start, *end;        // n x 3, or null when ids are given
  const int32_t *sid, *eid;
  const double *sx, *sy, *sth;      // store SoA (ids mode)
  int metric;
  int edge_kind;
  double robot_radius, r_min;
  ObsView obs;
  ObstacleGrid grid;
  int obs_smem;                     // stage the obstacle table in shared memory
  uint8_t *verdict;
  double *length;
};

__device__ __forceinline__ void load_pose(const EdgeArgs &A, int64_t e, bool end, double &x,
                                          double &y, double &t) {
  const int32_t *ids = end ? A.eid : A.sid;
  if (ids) {
    int32_t id = ids[e];
    x = A.sx[id]; y = A.sy[id]; t = A.sth[id];
  } else {
    const double *p = (end ? A.end : A.start) + 3 * e;
    x = p[0]; y = p[1]; t = p[2];
  }
}

// copies the active-obstacle table into shared memory and repoints the view
__device__ __forceinline__ void stage_obstacles(ObsView &O, unsigned char *smem) {
  const int n = O.n;
  const int nv = O.voff[n];
  double *s_ox = reinterpret_cast<double *>(smem);
  double *s_oy = s_ox + n, *s_rad = s_oy + n, *s_vx = s_rad + n, *s_vy = s_vx + nv;
  int32_t *s_kind = reinterpret_cast<int32_t *>(s_vy + nv);
  int32_t *s_voff = s_kind + n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s_ox[i] = O.ox[i]; s_oy[i] = O.oy[i]; s_rad[i] = O.rad[i]; s_kind[i] = O.kind[i];
  }
  for (int i = threadIdx.x; i <= n; i += blockDim.x) s_voff[i] = O.voff[i];
  for (int i = threadIdx.x; i < nv; i += blockDim.x) { s_vx[i] = O.vx[i]; s_vy[i] = O.vy[i]; }
  __syncthreads();
  O.ox = s_ox; O.oy = s_oy; O.rad = s_rad; O.vx = s_vx; O.vy = s_vy;
  O.kind = s_kind; O.voff = s_voff;
}

__global__ void __launch_bounds__(EDGE_T)
edgecheck_kernel(EdgeArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ObsView O = A.obs;
  if (A.obs_smem && O.n > 0) stage_obstacles(O, smem_raw);
  const int64_t e = (int64_t)blockIdx.x * EDGE_T + threadIdx.x;
  if (e >= A.n) return;
  double sx, sy, st, ex, ey, et;
  load_pose(A, e, false, sx, sy, st);
  load_pose(A, e, true, ex, ey, et);
  bool hit = false;
  double dist;
  if (A.edge_kind == DRRT_EDGE_DUBINS) {
    DubinsPath P;
    double edge_dist = metric_dist(A.metric, sx, sy, st, ex, ey, et);  // NewEdge :19
    dubins_solve(sx, sy, st, ex, ey, et, A.r_min, edge_dist, P);
    dist = P.dist;
    if (O.n > 0 && P.type != TRAJ_XXX) {
      PolylineWalker W;
      W.start(&P, A.r_min);
      double px, py, cx, cy;
      if (W.next(&px, &py)) {
        while (W.next(&cx, &cy)) {  // dubinsedge.cpp:875-893
          if (segment_vs_obstacles(O, A.grid, px, py, cx, cy, A.robot_radius)) { hit = true; break; }
          px = cx;
          py = cy;
        }
      }
    }
  } else {
    // LineCheck obstacle.cpp:1098-1133
    double th = atan2(ey - sy, ex - sx);
    dist = metric_dist(A.metric, sx, sy, th, ex, ey, th);
    if (O.n > 0) {
      const double traj_length = 50;
      double x_val = sx, y_val = sy;
      const double x_dist = fabs(ex - x_val), y_dist = fabs(ey - y_val);
      double px = x_val, py = y_val;
      for (int i = 0; i < 50; i++) {
        if (ex > sx) x_val += x_dist / traj_length; else x_val -= x_dist / traj_length;
        if (ey > sy) y_val += y_dist / traj_length; else y_val -= y_dist / traj_length;
        if (segment_vs_obstacles(O, A.grid, px, py, x_val, y_val, A.robot_radius)) { hit = true; break; }
        px = x_val;
        py = y_val;
      }
    }
  }
  A.verdict[e] = hit ? 1 : 0;
  if (A.length) A.length[e] = dist;
}

// uniform grid over the active obstacles' reach circles (host build, tiny)
static int ensure_obstacle_grid(Store *S, double robot_radius, cudaStream_t st) {
  if (S->ogrid_radius == robot_radius) return DRRT_OK;
  const size_t n = S->h_o_ox.size();
  ObstacleGrid G{};
  if (n == 0 || n > 65535) {  // nx == 0: kernels test every obstacle
    S->ogrid = G;
    S->ogrid_radius = robot_radius;
    return DRRT_OK;
  }
  std::vector<double> reach(n);
  double lox = 1e300, loy = 1e300, hix = -1e300, hiy = -1e300, mean = 0.0;
  for (size_t i = 0; i < n; i++) {
    reach[i] = (fabs(robot_radius) + fabs(S->h_o_rad[i])) * (1.0 + 1e-9) + 1e-9;
    lox = std::min(lox, S->h_o_ox[i] - reach[i]); hix = std::max(hix, S->h_o_ox[i] + reach[i]);
    loy = std::min(loy, S->h_o_oy[i] - reach[i]); hiy = std::max(hiy, S->h_o_oy[i] + reach[i]);
    mean += reach[i] / n;
  }
  if (!std::isfinite(lox + loy + hix + hiy + mean)) {
    S->ogrid = G;
    S->ogrid_radius = robot_radius;
    return DRRT_OK;
  }
  double extent = std::max(std::max(hix - lox, hiy - loy), 1e-9);
  double cell = std::max(extent / 96.0, 0.5 * mean);
  G.x0 = lox; G.y0 = loy; G.inv_cell = 1.0 / cell;
  G.nx = std::max(1, (int)ceil((hix - lox) / cell) + 1);
  G.ny = std::max(1, (int)ceil((hiy - loy) / cell) + 1);
  const int ncell = G.nx * G.ny;
  std::vector<int32_t> off(ncell + 1, 0);
  std::vector<uint16_t> idx;
  for (int pass = 0; pass < 2; pass++) {
    std::vector<int32_t> cur(off.begin(), off.end() - 1);
    for (size_t i = 0; i < n; i++) {
      const double ox = S->h_o_ox[i], oy = S->h_o_oy[i], r = reach[i];
      int cx0 = std::max(0, (int)floor((ox - r - lox) / cell) - 1), cx1 = std::min(G.nx - 1, (int)floor((ox + r - lox) / cell) + 1);
      int cy0 = std::max(0, (int)floor((oy - r - loy) / cell) - 1), cy1 = std::min(G.ny - 1, (int)floor((oy + r - loy) / cell) + 1);
      for (int cx = cx0; cx <= cx1; cx++)
        for (int cy = cy0; cy <= cy1; cy++) {
          // circle vs (slightly grown) cell rectangle
          double rx0 = lox + (cx - 1e-6) * cell, rx1 = lox + (cx + 1.000001) * cell;
          double ry0 = loy + (cy - 1e-6) * cell, ry1 = loy + (cy + 1.000001) * cell;
          double dx = std::max(std::max(rx0 - ox, ox - rx1), 0.0);
          double dy = std::max(std::max(ry0 - oy, oy - ry1), 0.0);
          if (dx * dx + dy * dy > r * r) continue;
          int c = cx * G.ny + cy;
          if (pass == 0) off[c + 1]++;
          else idx[cur[c]++] = (uint16_t)i;
        }
    }
    if (pass == 0) {
      for (int c = 0; c < ncell; c++) off[c + 1] += off[c];
      idx.resize(off[ncell] + 1);
    }
  }
  DRRT_CUDA(cudaStreamSynchronize(st));
  DRRT_CHECK(S->og_off.reserve(off.size(), st));
  DRRT_CHECK(S->og_idx.reserve(idx.size(), st));
  DRRT_CUDA(cudaMemcpyAsync(S->og_off.p, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  DRRT_CUDA(cudaMemcpyAsync(S->og_idx.p, idx.data(), idx.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
  DRRT_CUDA(cudaStreamSynchronize(st));
  G.off = S->og_off.p;
  G.idx = S->og_idx.p;
  S->ogrid = G;
  S->ogrid_radius = robot_radius;
  return DRRT_OK;
}

static size_t obstacle_smem_bytes(const Store *S, int nverts) {
  size_t n = (size_t)S->obs.n_active;
  return (3 * n + 2 * (size_t)nverts) * sizeof(double) + (2 * n + 1) * sizeof(int32_t) + 16;
}

int edgecheck_batch_dev(Store *S, int64_t n, const double *start_dev, const double *end_dev,
                        const int32_t *sid_dev, const int32_t *eid_dev, int edge_kind,
                        double robot_radius, double r_min, uint8_t *verdict_dev,
                        double *length_dev, cudaStream_t st) {
  if (n < 0) return set_error(DRRT_ERR_INVALID, "edgecheck: n < 0");
  if (edge_kind != DRRT_EDGE_DUBINS && edge_kind != DRRT_EDGE_STRAIGHT)
    return set_error(DRRT_ERR_INVALID, "edgecheck: edge_kind");
  bool by_id = sid_dev && eid_dev;
  if (!by_id && !(start_dev && end_dev)) return set_error(DRRT_ERR_INVALID, "edgecheck: endpoints");
  if (!verdict_dev) return set_error(DRRT_ERR_INVALID, "edgecheck: verdict");
  if (n == 0) return DRRT_OK;
  EdgeArgs A{};
  A.n = n; A.start = by_id ? nullptr : start_dev; A.end = by_id ? nullptr : end_dev;
  A.sid = by_id ? sid_dev : nullptr; A.eid = by_id ? eid_dev : nullptr;
  A.sx = S->x.p; A.sy = S->y.p; A.sth = S->th.p;
  A.metric = S->metric; A.edge_kind = edge_kind;
  A.robot_radius = robot_radius; A.r_min = r_min;
  A.obs.n = S->has_obstacles ? S->obs.n_active : 0;
  A.obs.kind = S->obs.kind; A.obs.ox = S->obs.ox; A.obs.oy = S->obs.oy; A.obs.rad = S->obs.rad;
  A.obs.voff = S->obs.vert_off; A.obs.vx = S->obs.vx; A.obs.vy = S->obs.vy;
  A.verdict = verdict_dev; A.length = length_dev;
  size_t smem = 0;
  if (A.obs.n > 0) {
    DRRT_CHECK(ensure_obstacle_grid(S, robot_radius, st));
    A.grid = S->ogrid;
    size_t need = obstacle_smem_bytes(S, S->obs.n_active_verts);
    if (need <= 96 * 1024) { smem = need; A.obs_smem = 1; }
  }
  static bool attr_done = false;
  if (!attr_done) {
    DRRT_CUDA(cudaFuncSetAttribute(edgecheck_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   96 * 1024));
    attr_done = true;
  }
  unsigned nb = (unsigned)((n + EDGE_T - 1) / EDGE_T);
  edgecheck_kernel<<<nb, EDGE_T, smem, st>>>(A);
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaGetLastError());
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
  if (length) length[e] = P.dist;
  if (type) {
    const char *names = "xxx\0rsl\0rsr\0rlr\0lsr\0lsl\0lrl\0" "0s0\0";
    for (int c = 0; c < 4; c++) type[4 * e + c] = names[4 * P.type + c];
  }
}

int trajectory_batch_dev(Store *S, int64_t n, const double *start_dev, const double *end_dev,
                         double r_min, double *xy_dev, int32_t *npts_dev, char *type_dev,
                         double *length_dev, cudaStream_t st) {
  if (n < 0 || !start_dev || !end_dev || !xy_dev || !npts_dev)
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

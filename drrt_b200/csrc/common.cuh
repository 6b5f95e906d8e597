// common.cuh -- shared declarations of libdrrt_b200 (sm_100a only).
//
// All device arithmetic that decides a result is IEEE double in the
// reference's operation order; the library is compiled with -fmad=false so
// nvcc never contracts a*b+c into an FMA (the reference's x86-64 build has
// none).  Reference constants: include/DRRT/distancefunctions.h:22-25.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>

#include "../../include/drrt_gpu.h"

#define DRRT_PI 3.1415926536          // truncated literal, SURVEY F3
#define DRRT_TWO_PI (2.0 * DRRT_PI)
#define DRRT_INF 1000000000000.0      // integer literal 1e12, SURVEY F3

namespace drrt {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;

int set_error(int code, const char *fmt, ...);

#define DRRT_CUDA(call)                                                       \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess)                                                   \
      return drrt::set_error(DRRT_ERR_CUDA, "%s:%d %s: %s", __FILE__,         \
                             __LINE__, #call, cudaGetErrorString(e__));       \
  } while (0)

#define DRRT_LAUNCHED() (drrt::g_launches.fetch_add(1, std::memory_order_relaxed))

#define DRRT_CHECK(call)                 \
  do {                                   \
    int rc__ = (call);                   \
    if (rc__ != DRRT_OK) return rc__;    \
  } while (0)

// ---------------------------------------------------------------------------
// Sorted node record: one 32-byte sector per candidate.  gatef is gate[id] rounded to float: it
// decides the signed theta prune of a candidate (walk_reaches, range.cu) without a second trip to
// L2 unless the query's bound falls within the rounding of the copy; the insert kernels keep it
// current when a later insert lowers the gate of an indexed node (store.cu, kd_insert_step).
struct __align__(32) NodeRec {
  double x, y, th;
  int32_t id;
  float gatef;
};

// Uniform bucket grid over (x, y, key(theta)); theta is the fastest index so a
// (ix,iy) column's theta range is one contiguous run of records.
struct GridParams {
  double x0, y0, inv_hx, inv_hy;  // cell = floor((x - x0) * inv_h), clamped
  double hx, hy, ht;              // cell edge lengths (1 / inv_h)
  double t0, inv_ht;              // theta key origin / inverse cell size
  double period;                  // > 0: theta key = theta mod period
  int nx, ny, nt;
  int periodic;
};

struct ObstacleTable {
  int32_t n_active;            // obstacles that can ever report a hit (kind 1 / 3)
  const int32_t *kind;         // n_active
  const double *ox, *oy, *rad; // n_active
  const int32_t *vert_off;     // n_active + 1
  int32_t n_active_verts;
  const double *vx, *vy;       // vert_off[n_active]
  // all obstacles (point checks need kinds 2,4,5 too)
  int32_t n_all;
  const int32_t *a_kind;
  const uint8_t *a_live;       // used && life_span > 0
  const double *a_ox, *a_oy, *a_rad;
  const int32_t *a_vert_off;
  const double *a_vx, *a_vy;
};

struct __align__(32) EdgeRec { double ax, ay, bx, by; };

// Broadphase of the edge checks: a uniform grid whose cells list the polygon
// EDGES (and kind-1 balls) that come within the robot radius of the cell.
struct ObstacleGrid {
  double x0, y0, inv_cell;
  int nx, ny;
  const int32_t *off;     // nx*ny + 1
  const int32_t *idx;     // entry < n_edges: polygon edge; else ball (entry - n_edges)
  int32_t n_edges;
  const struct EdgeRec *erec;  // polygon edge end points (A = previous vertex), one 256-bit load each
  const int32_t *eobs;    // owning active obstacle of each edge
  const int32_t *ball;    // active-obstacle index of each kind-1 ball
};

// ---------------------------------------------------------------------------
// Metrics.  std::min(a,b) == (b<a)?b:a ; std::max(a,b) == (a<b)?b:a.

// squared-sum under the sqrt of distance_function (smalltest.cpp:167-175):
//   (dx*dx + dy*dy) + m*m,  m = min(|a2-b2|, min(a2,b2) + 2*PI - max(a2,b2))
__host__ __device__ __forceinline__ double r2s1_sum(double ax, double ay, double at,
                                                    double bx, double by, double bt) {
  double t0 = ax - bx, t1 = ay - by;
  t0 = t0 * t0;
  t1 = t1 * t1;
  double s = t0 + t1;
  double ad = fabs(at - bt);
  // the same two roundings without materialising min and max: at < bt picks (at + 2 PI) - bt, else
  // (bt + 2 PI) - at (bt == at gives either); at + 2 PI is loop-invariant in the scans
  double w1 = (at + DRRT_TWO_PI) - bt, w2 = (bt + DRRT_TWO_PI) - at;
  double wr = (at < bt) ? w1 : w2;
  double m = (wr < ad) ? wr : ad;
  return s + m * m;
}

// DistFunc (thetastar.cpp:10-15): ((t0 + t1) + t2)
__host__ __device__ __forceinline__ double euclid_sum(double ax, double ay, double at,
                                                      double bx, double by, double bt) {
  double t0 = ax - bx, t1 = ay - by, t2 = at - bt;
  t0 = t0 * t0;
  t1 = t1 * t1;
  t2 = t2 * t2;
  return (t0 + t1) + t2;
}

template <int METRIC>
__host__ __device__ __forceinline__ double metric_sum(double ax, double ay, double at,
                                                      double bx, double by, double bt) {
  if (METRIC == DRRT_METRIC_EUCLID) return euclid_sum(ax, ay, at, bx, by, bt);
  return r2s1_sum(ax, ay, at, bx, by, bt);
}

__host__ __device__ __forceinline__ double metric_dist(int metric, double ax, double ay,
                                                       double at, double bx, double by,
                                                       double bt) {
  return sqrt(metric == DRRT_METRIC_EUCLID ? euclid_sum(ax, ay, at, bx, by, bt)
                                           : r2s1_sum(ax, ay, at, bx, by, bt));
}

// theta bucketing key in [0, period): exact fmod, one rounded add at most
__host__ __device__ __forceinline__ double theta_key(double th, double period) {
  if (!(period > 0.0)) return th;
  if (th >= 0.0 && th < period) return th;  // the usual case; fmod is exact but slow
  double k = fmod(th, period);
  if (k < 0.0) k += period;
  if (k >= period) k = 0.0;
  return k;
}

// Ghost of a query for the single wrapped dimension (theta), restating
// GetNextGhostPoint src/ghostpoint.cpp:15-75 for num_wraps_ == 1.
// Returns true when the reference would search the ghost.
template <int METRIC>
__host__ __device__ __forceinline__ bool ghost_of(double qx, double qy, double qt,
                                                  double wrap_value, double best_dist,
                                                  double *gt) {
  double dim_val = qt, dim_closest = 0.0;
  if (qt < wrap_value / 2.0) {
    dim_val += wrap_value;
    dim_closest += wrap_value;
  } else {
    dim_val -= wrap_value;
  }
  // metric(closest_unwrapped_point_, current_ghost_) > best_dist -> skipped (:69-72)
  double d = sqrt(metric_sum<METRIC>(qx, qy, dim_closest, qx, qy, dim_val));
  if (d > best_dist) return false;
  // callers stop on thisGhostPoint.isZero(0) (kdtree.cpp:285, 814)
  if (qx == 0.0 && qy == 0.0 && dim_val == 0.0) return false;
  *gt = dim_val;
  return true;
}

}  // namespace drrt

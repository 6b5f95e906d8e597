// timed.cu -- ExplicitEdgeCheck2D for the time-parametrised obstacle kinds 6 and 7
// (src/obstacle.cpp:805-875; FindIndexBeforeTime src/drrt.cpp:178-189).
//
// An obstacle of these kinds is a disc of radius_ whose centre moves along path_ rows
// (dx, dy, time) relative to origin_; the segment is a robot move (x, y, TIME) -> (x, y, TIME).
// For every path piece that overlaps the move in time the reference solves for the time of
// closest approach of the two uniformly moving centres, clamps it to the common time window
// and reports a collision when the centres are then closer than radius_ + radius.  The planner
// itself never reaches this branch (space_has_time_ == false, src/smalltest.cpp:200; the Dubins
// store has no time axis), so it is offered as what it is: the per-segment test, batched, one
// thread per segment, OR over the timed obstacles, arithmetic in the reference's order
// (-fmad=false).
#include <algorithm>

#include "store.cuh"

namespace drrt {

struct TimedView {
  int n;
  const double *ox, *oy, *rad;  // live kind-6/7 obstacles only
  const int32_t *poff;          // n + 1
  const double *path;           // poff[n] x 3
};

__device__ __forceinline__ int index_before_time(const double *path, int rows, double t) {  // drrt.cpp:178-189
  if (rows < 1) return -1;
  int i = -1;
  while (i + 1 < rows && path[3 * (i + 1) + 2] < t) i += 1;
  return i;
}

__global__ void timed_check_kernel(int64_t n, const double *start, const double *end, double radius, TimedView T,
                                   uint8_t *verdict) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const double *s = start + 3 * e, *g = end + 3 * e;
  const bool fwd = s[2] < g[2];  // :813-819 always past -> future
  const double ex = fwd ? s[0] : g[0], ey = fwd ? s[1] : g[1], et = fwd ? s[2] : g[2];
  const double lx = fwd ? g[0] : s[0], ly = fwd ? g[1] : s[1], lt = fwd ? g[2] : s[2];
  bool hit = false;
  for (int o = 0; o < T.n && !hit; o++) {
    const double *path = T.path + 3 * (size_t)T.poff[o];
    const int rows = T.poff[o + 1] - T.poff[o];
    int first = max(index_before_time(path, rows, et), 0);
    int last = min(index_before_time(path, rows, lt), rows - 1);
    if (last <= first) continue;  // :825
    const double reach = T.rad[o] + radius;
    for (int i0 = first; i0 < last && !hit; i0++) {
      const int i1 = i0 + 1;
      const double x_1 = ex, y_1 = ey, t_1 = et;
      const double x_2 = path[3 * i0] + T.ox[o], y_2 = path[3 * i0 + 1] + T.oy[o], t_2 = path[3 * i0 + 2];
      const double m_x1 = (lx - x_1) / (lt - t_1), m_y1 = (ly - y_1) / (lt - t_1);
      const double m_x2 = (path[3 * i1] + T.ox[o] - x_2) / (path[3 * i1 + 2] - t_2);
      const double m_y2 = (path[3 * i1 + 1] + T.oy[o] - y_2) / (path[3 * i1 + 2] - t_2);
      double t_c = (m_x1 * m_x1 * t_1 + m_x2 * (m_x2 * t_2 + x_1 - x_2) - m_x1 * (m_x2 * (t_1 + t_2) + x_1 - x_2) +
                    (m_y1 - m_y2) * (m_y1 * t_1 - m_y2 * t_2 - y_1 + y_2)) /
                   ((m_x1 - m_x2) * (m_x1 - m_x2) + (m_y1 - m_y2) * (m_y1 - m_y2));
      const double lo = (t_1 < t_2) ? t_2 : t_1;                            // std::max
      const double hi = (path[3 * i1 + 2] < lt) ? path[3 * i1 + 2] : lt;    // std::min
      if (t_c < lo) t_c = lo;
      else if (t_c > hi) t_c = hi;
      const double r_x = m_x1 * (t_c - t_1) + x_1, r_y = m_y1 * (t_c - t_1) + y_1;
      const double o_x = m_x2 * (t_c - t_2) + x_2, o_y = m_y2 * (t_c - t_2) + y_2;
      if ((r_x - o_x) * (r_x - o_x) + (r_y - o_y) * (r_y - o_y) < reach * reach) hit = true;
    }
  }
  verdict[e] = hit ? 1 : 0;
}

static int timed_dev(Store *S, int64_t n, const double *start_dev, const double *end_dev, double radius,
                     uint8_t *verdict_dev, cudaStream_t st) {
  if (n == 0) return DRRT_OK;
  TimedView T{};
  T.n = S->timed_n;
  T.ox = S->tm_ox.p; T.oy = S->tm_oy.p; T.rad = S->tm_rad.p; T.poff = S->tm_poff.p; T.path = S->tm_path.p;
  timed_check_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, start_dev, end_dev, radius, T, verdict_dev);
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaGetLastError());
  return DRRT_OK;
}

template <typename T>
static int put_vec(DevBuf<T> &b, const std::vector<T> &v, cudaStream_t st) {
  DRRT_CHECK(b.reserve(v.size() + 1, st));
  if (!v.empty()) DRRT_CUDA(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return DRRT_OK;
}

}  // namespace drrt

using namespace drrt;

extern "C" {

int drrt_timed_obstacles_set(drrt_store *h, int32_t n_obs, const int32_t *kind, const uint8_t *used,
                             const double *life_span, const double *origin_xy, const double *radius,
                             const int32_t *path_off, const double *path_xyt) {
  ENTER(h);
  if (n_obs < 0) return set_error(DRRT_ERR_INVALID, "timed obstacles: n_obs < 0");
  if (n_obs > 0 && (!kind || !used || !life_span || !origin_xy || !radius || !path_off))
    return set_error(DRRT_ERR_INVALID, "timed obstacles: null array");
  std::vector<double> ox, oy, rad, path;
  std::vector<int32_t> poff(1, 0);
  for (int i = 0; i < n_obs; i++) {
    if (kind[i] != 6 && kind[i] != 7)
      return set_error(DRRT_ERR_INVALID, "timed obstacles: kind %d belongs in drrt_obstacles_set", kind[i]);
    if (path_off[i + 1] < path_off[i]) return set_error(DRRT_ERR_INVALID, "timed obstacles: path_off");
    if (path_off[i + 1] > path_off[i] && !path_xyt) return set_error(DRRT_ERR_INVALID, "timed obstacles: path");
    if (!used[i] || life_span[i] <= 0) continue;  // obstacle.cpp:761: never collides
    ox.push_back(origin_xy[2 * i]); oy.push_back(origin_xy[2 * i + 1]); rad.push_back(radius[i]);
    for (int r = path_off[i]; r < path_off[i + 1]; r++)
      for (int k = 0; k < 3; k++) path.push_back(path_xyt[3 * (size_t)r + k]);
    poff.push_back((int32_t)(path.size() / 3));
  }
  cudaStream_t st = S->stream;
  DRRT_CUDA(cudaStreamSynchronize(st));
  DRRT_CHECK(put_vec(S->tm_ox, ox, st)); DRRT_CHECK(put_vec(S->tm_oy, oy, st));
  DRRT_CHECK(put_vec(S->tm_rad, rad, st)); DRRT_CHECK(put_vec(S->tm_poff, poff, st));
  DRRT_CHECK(put_vec(S->tm_path, path, st));
  DRRT_CUDA(cudaStreamSynchronize(st));  // host vectors die at return
  S->timed_n = (int)ox.size();
  return DRRT_OK;
}

int drrt_edgecheck2d_timed_batch(drrt_store *h, int64_t n, const double *start_xyt, const double *end_xyt,
                                 double radius, uint8_t *verdict) {
  ENTER(h);
  if (n < 0 || (n > 0 && (!start_xyt || !end_xyt || !verdict)))
    return set_error(DRRT_ERR_INVALID, "timed edge check: arguments");
  if (n == 0) return DRRT_OK;
  cudaStream_t st = S->stream;
  DRRT_CHECK(S->d_stage.reserve(3 * n, st));
  DRRT_CHECK(S->d_stage2.reserve(3 * n, st));
  DRRT_CHECK(S->b_stage.reserve(n, st));
  DRRT_CUDA(cudaMemcpyAsync(S->d_stage.p, start_xyt, 3 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  DRRT_CUDA(cudaMemcpyAsync(S->d_stage2.p, end_xyt, 3 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  DRRT_CHECK(timed_dev(S, n, S->d_stage.p, S->d_stage2.p, radius, S->b_stage.p, st));
  DRRT_CUDA(cudaMemcpyAsync(verdict, S->b_stage.p, n, cudaMemcpyDeviceToHost, st));
  DRRT_CUDA(cudaStreamSynchronize(st));
  return DRRT_OK;
}

int drrt_edgecheck2d_timed_batch_dev(drrt_store *h, int64_t n, const double *start_dev, const double *end_dev,
                                     double radius, uint8_t *verdict_dev, void *stream) {
  ENTER(h);
  if (n < 0 || (n > 0 && (!start_dev || !end_dev || !verdict_dev)))
    return set_error(DRRT_ERR_INVALID, "timed edge check: arguments");
  return timed_dev(S, n, start_dev, end_dev, radius, verdict_dev, stream ? (cudaStream_t)stream : S->stream);
}

}  // extern "C"

// capi.cu -- the extern "C" boundary declared in include/drrt_gpu.h.
// Host entry points stage caller buffers through device scratch owned by the
// handle, on the handle's stream, and return after the results are on the host.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "store.cuh"

namespace drrt {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};

int set_error(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

static cudaStream_t pick(Store *S, void *stream) {
  return stream ? (cudaStream_t)stream : S->stream;
}

static int upload(const void *src, void *dst, size_t bytes, cudaStream_t st) {
  if (!bytes) return DRRT_OK;
  DRRT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
  return DRRT_OK;
}
static int download(void *dst, const void *src, size_t bytes, cudaStream_t st) {
  if (!bytes) return DRRT_OK;
  DRRT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
  return DRRT_OK;
}

int map_reserve(Store *S) {
  if (S->h_map) return DRRT_OK;
  cudaError_t e = cudaHostAlloc(&S->h_map, sizeof(MappedRow) + sizeof(MappedSmall),
                                cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) {
    S->h_map = nullptr;
    cudaGetLastError();
    return set_error(DRRT_ERR_NOMEM, "mapped host buffer: %s", cudaGetErrorString(e));
  }
  return DRRT_OK;
}

static MappedSmall *mapped_small(Store *S) {
  return reinterpret_cast<MappedSmall *>(reinterpret_cast<char *>(S->h_map) + sizeof(MappedRow));
}

int knn_host(Store *S, int64_t nq, const double *q, int k, int32_t *ids, double *dist,
                    bool nearest) {
  if (nq < 0 || k < 1 || (nq > 0 && (!q || !ids)))
    return set_error(DRRT_ERR_INVALID, "knn: arguments");
  cudaStream_t st = S->stream;
  if (nq > 0 && nq * k <= MAP_SMALL) {
    // the planner asks for ONE nearest node per iteration (mainloop.cpp:204-210): the kernels read
    // the query from, and write the answer into, mapped host memory -- launches and one wait
    DRRT_CHECK(map_reserve(S));
    MappedSmall *M = mapped_small(S);
    memcpy(M->a, q, 3 * nq * sizeof(double));
    DRRT_CHECK(knn_batch_dev(S, nq, M->a, k, M->out_i, M->out_d, st, nearest));
    DRRT_CUDA(cudaStreamSynchronize(st));
    memcpy(ids, M->out_i, nq * k * sizeof(int32_t));
    if (dist) memcpy(dist, M->out_d, nq * k * sizeof(double));
    return DRRT_OK;
  }
  DRRT_CHECK(S->q_stage.reserve(3 * nq + 1, st));
  DRRT_CHECK(S->i_stage.reserve(nq * k + 1, st));
  DRRT_CHECK(S->d_stage.reserve(nq * k + 1, st));
  DRRT_CHECK(upload(q, S->q_stage.p, 3 * nq * sizeof(double), st));
  DRRT_CHECK(knn_batch_dev(S, nq, S->q_stage.p, k, S->i_stage.p, S->d_stage.p, st, nearest));
  DRRT_CHECK(download(ids, S->i_stage.p, nq * k * sizeof(int32_t), st));
  if (dist) DRRT_CHECK(download(dist, S->d_stage.p, nq * k * sizeof(double), st));
  DRRT_CUDA(cudaStreamSynchronize(st));
  return DRRT_OK;
}

template <typename T>
static int put(DevBuf<T> &b, const std::vector<T> &v, cudaStream_t st) {
  DRRT_CHECK(b.reserve(v.size() + 1, st));
  return upload(v.data(), b.p, v.size() * sizeof(T), st);
}

// Host-buffer leg of the edge checks.  Large pose batches are PIPELINED: the poses of chunk
// k + 1 cross PCIe (48 B per edge, the dominant cost: 480 MB for config 3's 10 M edges against
// 8 ms of kernels) while chunk k is solved and walked, and the verdicts / lengths of chunk k - 1
// leave on a third stream -- the call costs max(H2D, kernels) + one chunk instead of their sum.
#ifndef DRRT_EDGE_CHUNK
#define DRRT_EDGE_CHUNK (1 << 21)
#endif
static constexpr int64_t EDGE_CHUNK = DRRT_EDGE_CHUNK;

static int edge_streams(Store *S) {
  if (S->ev_up[0]) return DRRT_OK;
  if (!S->copy_stream) DRRT_CUDA(cudaStreamCreateWithFlags(&S->copy_stream, cudaStreamNonBlocking));
  if (!S->down_stream) DRRT_CUDA(cudaStreamCreateWithFlags(&S->down_stream, cudaStreamNonBlocking));
  for (int b = 0; b < 2; b++) {
    DRRT_CUDA(cudaEventCreateWithFlags(&S->ev_done[b], cudaEventDisableTiming));
    DRRT_CUDA(cudaEventCreateWithFlags(&S->ev_up[b], cudaEventDisableTiming));
  }
  return DRRT_OK;
}

// edge end ids are validated where they are consumed: an id outside the store is clamped (so the
// edge kernels never read out of bounds) and flagged; the call then fails with DRRT_ERR_INVALID
__global__ void validate_ids_kernel(int32_t *a, int32_t *b, int64_t n, int32_t n_nodes, int32_t *flag) {
  bool bad = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t va = a[i], vb = b[i];
    if (va < 0 || va >= n_nodes) { a[i] = 0; bad = true; }
    if (vb < 0 || vb >= n_nodes) { b[i] = 0; bad = true; }
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *flag = 1;
}

static int validate_ids(Store *S, int32_t *a, int32_t *b, int64_t n, int32_t *flag, cudaStream_t st) {
  if (S->n == 0) return set_error(DRRT_ERR_INVALID, "edgecheck: node ids given, the store is empty");
  validate_ids_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(a, b, n, (int32_t)S->n, flag);
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaGetLastError());
  return DRRT_OK;
}

int edge_host(Store *S, int64_t n, const double *start, const double *end,
              const int32_t *sid, const int32_t *eid, int edge_kind, double robot_radius,
              double r_min, int in_warmup, uint8_t *verdict, double *length) {
  if (n < 0 || (n > 0 && !verdict)) return set_error(DRRT_ERR_INVALID, "edgecheck: arguments");
  bool by_id = sid && eid;
  if (n > 0 && !by_id && !(start && end)) return set_error(DRRT_ERR_INVALID, "edgecheck: endpoints");
  if (n == 0) return DRRT_OK;
  cudaStream_t st = S->stream;
  DRRT_CHECK(map_reserve(S));
  int32_t *flag = &mapped_small(S)->flag;
  *flag = 0;
  // in_warmup_time_ (obstacle.cpp:883): no obstacle is consulted, lengths still computed
  const bool saved = S->has_obstacles;
  struct Restore {
    Store *S; bool v;
    ~Restore() { S->has_obstacles = v; }
  } restore{S, saved};
  if (in_warmup) S->has_obstacles = false;
  DRRT_CHECK(S->b_stage.reserve(n, st));
  if (length) DRRT_CHECK(S->d_stage3.reserve(n, st));

  if (n <= MAP_SMALL) {
    // one edge (or one node's handful) per call, as the unchanged planner asks (drrt.cpp:411-422,
    // 286-293): poses in and verdicts out through mapped host memory, no copy calls
    DRRT_CHECK(map_reserve(S));
    MappedSmall *M = mapped_small(S);
    if (by_id) {
      memcpy(M->ia, sid, n * sizeof(int32_t));
      memcpy(M->ib, eid, n * sizeof(int32_t));
      DRRT_CHECK(validate_ids(S, M->ia, M->ib, n, flag, st));
    } else {
      memcpy(M->a, start, 3 * n * sizeof(double));
      memcpy(M->b, end, 3 * n * sizeof(double));
    }
    DRRT_CHECK(edgecheck_batch_dev(S, n, by_id ? nullptr : M->a, by_id ? nullptr : M->b,
                                   by_id ? M->ia : nullptr, by_id ? M->ib : nullptr, edge_kind,
                                   robot_radius, r_min, M->out_b, length ? M->out_d : nullptr, st));
    DRRT_CUDA(cudaStreamSynchronize(st));
    if (*flag) return set_error(DRRT_ERR_INVALID, "edgecheck: node id out of range");
    memcpy(verdict, M->out_b, n);
    if (length) memcpy(length, M->out_d, n * sizeof(double));
    return DRRT_OK;
  }
  if (n < 2 * EDGE_CHUNK) {
    const double *sd = nullptr, *ed = nullptr;
    const int32_t *si = nullptr, *ei = nullptr;
    if (by_id) {
      DRRT_CHECK(S->i_stage.reserve(n, st));
      DRRT_CHECK(S->i_stage2.reserve(n, st));
      DRRT_CHECK(upload(sid, S->i_stage.p, n * sizeof(int32_t), st));
      DRRT_CHECK(upload(eid, S->i_stage2.p, n * sizeof(int32_t), st));
      DRRT_CHECK(validate_ids(S, S->i_stage.p, S->i_stage2.p, n, flag, st));
      si = S->i_stage.p; ei = S->i_stage2.p;
    } else {
      DRRT_CHECK(S->d_stage.reserve(3 * n, st));
      DRRT_CHECK(S->d_stage2.reserve(3 * n, st));
      DRRT_CHECK(upload(start, S->d_stage.p, 3 * n * sizeof(double), st));
      DRRT_CHECK(upload(end, S->d_stage2.p, 3 * n * sizeof(double), st));
      sd = S->d_stage.p; ed = S->d_stage2.p;
    }
    DRRT_CHECK(edgecheck_batch_dev(S, n, sd, ed, si, ei, edge_kind, robot_radius, r_min, S->b_stage.p,
                                   length ? S->d_stage3.p : nullptr, st));
    DRRT_CHECK(download(verdict, S->b_stage.p, n, st));
    if (length) DRRT_CHECK(download(length, S->d_stage3.p, n * sizeof(double), st));
    DRRT_CUDA(cudaStreamSynchronize(st));
    if (*flag) return set_error(DRRT_ERR_INVALID, "edgecheck: node id out of range");
    return DRRT_OK;
  }

  DRRT_CHECK(edge_streams(S));
  cudaStream_t up = S->copy_stream, down = S->down_stream;
  const int64_t CH = EDGE_CHUNK;
  if (by_id) {
    DRRT_CHECK(S->i_stage.reserve(2 * CH, st));
    DRRT_CHECK(S->i_stage2.reserve(2 * CH, st));
  } else {
    DRRT_CHECK(S->d_stage.reserve(2 * 3 * CH, st));
    DRRT_CHECK(S->d_stage2.reserve(2 * 3 * CH, st));
  }
  DRRT_CUDA(cudaStreamSynchronize(st));  // the staging buffers may have been reallocated
  int rc = DRRT_OK;
  bool used[2] = {false, false};
  auto chunk = [&](int64_t k0, int64_t m, int b) -> int {
    if (used[b]) DRRT_CUDA(cudaStreamWaitEvent(up, S->ev_done[b], 0));  // half b is free again
    const double *sd = nullptr, *ed = nullptr;
    const int32_t *si = nullptr, *ei = nullptr;
    if (by_id) {
      DRRT_CHECK(upload(sid + k0, S->i_stage.p + b * CH, m * sizeof(int32_t), up));
      DRRT_CHECK(upload(eid + k0, S->i_stage2.p + b * CH, m * sizeof(int32_t), up));
      DRRT_CUDA(cudaEventRecord(S->ev_up[b], up));
      DRRT_CUDA(cudaStreamWaitEvent(st, S->ev_up[b], 0));
      DRRT_CHECK(validate_ids(S, S->i_stage.p + b * CH, S->i_stage2.p + b * CH, m, flag, st));
      si = S->i_stage.p + b * CH; ei = S->i_stage2.p + b * CH;
    } else {
      DRRT_CHECK(upload(start + 3 * k0, S->d_stage.p + b * 3 * CH, 3 * m * sizeof(double), up));
      DRRT_CHECK(upload(end + 3 * k0, S->d_stage2.p + b * 3 * CH, 3 * m * sizeof(double), up));
      sd = S->d_stage.p + b * 3 * CH; ed = S->d_stage2.p + b * 3 * CH;
    }
    DRRT_CUDA(cudaEventRecord(S->ev_up[b], up));
    DRRT_CUDA(cudaStreamWaitEvent(st, S->ev_up[b], 0));
    DRRT_CHECK(edgecheck_batch_dev(S, m, sd, ed, si, ei, edge_kind, robot_radius, r_min,
                                   S->b_stage.p + k0, length ? S->d_stage3.p + k0 : nullptr, st));
    DRRT_CUDA(cudaEventRecord(S->ev_done[b], st));
    DRRT_CUDA(cudaStreamWaitEvent(down, S->ev_done[b], 0));
    DRRT_CHECK(download(verdict + k0, S->b_stage.p + k0, m, down));
    if (length) DRRT_CHECK(download(length + k0, S->d_stage3.p + k0, m * sizeof(double), down));
    used[b] = true;
    return DRRT_OK;
  };
  for (int64_t k0 = 0, k = 0; k0 < n && rc == DRRT_OK; k0 += CH, k++)
    rc = chunk(k0, std::min(CH, n - k0), (int)(k & 1));
  // copies into the caller's buffers are always waited for, error or not
  cudaError_t e1 = cudaStreamSynchronize(up), e2 = cudaStreamSynchronize(st), e3 = cudaStreamSynchronize(down);
  if (rc != DRRT_OK) return rc;
  DRRT_CUDA(e1);
  DRRT_CUDA(e2);
  DRRT_CUDA(e3);
  if (*flag) return set_error(DRRT_ERR_INVALID, "edgecheck: node id out of range");
  return DRRT_OK;
}

__global__ void fill_f64_kernel(double *p, int64_t n, double v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

__global__ void scatter_f64_kernel(double *dst, int64_t n, const int32_t *ids, const double *v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[ids[i]] = v[i];
}

// The device mirror of KDTreeNode::rrt_LMC_ covers every node of the store: nodes the caller
// has not written yet hold INF, the value a new node carries until FindBestParent gives it a
// parent (drrt.cpp:391).
int lmc_mirror_cover(Store *S, cudaStream_t st) {
  if (S->lmc_n >= S->n) return DRRT_OK;
  DRRT_CHECK(S->x_lmc.reserve((size_t)S->n + 1, st, true, (size_t)S->lmc_n));
  const int64_t fresh = S->n - S->lmc_n;
  fill_f64_kernel<<<(unsigned)std::min<int64_t>((fresh + 255) / 256, 148 * 8), 256, 0, st>>>(
      S->x_lmc.p + S->lmc_n, fresh, DRRT_INF);
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaGetLastError());
  S->lmc_n = S->n;
  return DRRT_OK;
}

}  // namespace drrt

using namespace drrt;

extern "C" {

const char *drrt_last_error(void) { return g_last_error.c_str(); }
int drrt_abi_version(void) { return DRRT_ABI_VERSION; }
int64_t drrt_kernel_launch_count(void) { return g_launches.load(); }

int drrt_store_create(int device, int dims, int n_wraps, const int *wrap_dims,
                      const double *wrap_points, int metric, int64_t capacity_hint,
                      drrt_store **out) {
  if (!out) return set_error(DRRT_ERR_INVALID, "create: out == NULL");
  *out = nullptr;
  if (dims != 3) return set_error(DRRT_ERR_UNSUPPORTED, "create: dims must be 3 (x, y, theta)");
  if (metric != DRRT_METRIC_R2S1_WRAP && metric != DRRT_METRIC_EUCLID)
    return set_error(DRRT_ERR_INVALID, "create: unknown metric %d", metric);
  if (n_wraps < 0 || n_wraps > 1)
    return set_error(DRRT_ERR_UNSUPPORTED, "create: at most one wrapped dimension");
  if (n_wraps == 1) {
    if (!wrap_dims || !wrap_points) return set_error(DRRT_ERR_INVALID, "create: wrap arrays");
    if (wrap_dims[0] != 2)
      return set_error(DRRT_ERR_UNSUPPORTED, "create: only dimension 2 (theta) may wrap");
    if (!(wrap_points[0] > 0.0)) return set_error(DRRT_ERR_INVALID, "create: wrap point <= 0");
    if (metric == DRRT_METRIC_R2S1_WRAP && wrap_points[0] != DRRT_TWO_PI)
      return set_error(DRRT_ERR_UNSUPPORTED,
                       "create: the R2S1 metric wraps at 2.0*3.1415926536; wrap point differs");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_error(DRRT_ERR_NO_DEVICE, "create: no CUDA device (this library has no CPU path)");
  }
  if (device < 0 || device >= ndev) return set_error(DRRT_ERR_INVALID, "create: device %d", device);
  cudaDeviceProp prop;
  DRRT_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_error(DRRT_ERR_NO_DEVICE, "create: device %d is sm_%d%d; built for sm_100a only",
                     device, prop.major, prop.minor);
  DeviceGuard guard(device);
  if (!guard.ok) return set_error(DRRT_ERR_CUDA, "create: cudaSetDevice");
  Store *S = new (std::nothrow) Store();
  if (!S) return set_error(DRRT_ERR_NOMEM, "create: host allocation");
  S->device = device;
  S->metric = metric;
  S->n_wraps = n_wraps;
  S->wrap_point = n_wraps ? wrap_points[0] : 0.0;
  S->period = (metric == DRRT_METRIC_R2S1_WRAP) ? DRRT_TWO_PI : (n_wraps ? wrap_points[0] : 0.0);
  cudaError_t e = cudaStreamCreateWithFlags(&S->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&S->d_pending, 64);
  if (e == cudaSuccess) e = cudaMalloc(&S->d_ticket, 64);
  if (e == cudaSuccess) S->d_totals = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(S->d_ticket) + 16);
  if (e == cudaSuccess) e = cudaMalloc(&S->d_edge_counter, 64);
  if (e == cudaSuccess) e = cudaMalloc(&S->d_bounds, 6 * sizeof(double));
  if (e == cudaSuccess) e = cudaMallocHost(&S->h_pending, 64);
  if (e == cudaSuccess) e = cudaMallocHost(&S->h_bounds, 6 * sizeof(double));
  if (e == cudaSuccess) e = cudaMallocHost(&S->h_scalar, 8 * sizeof(int64_t));
  if (e != cudaSuccess) {
    delete S;
    return set_error(DRRT_ERR_CUDA, "create: %s", cudaGetErrorString(e));
  }
  if (capacity_hint > 0) {
    // everything that scales with the node count, so a store grown to the hint never
    // reallocates (cudaMalloc/cudaFree stall the device for milliseconds)
    int rc = store_reserve_nodes(S, capacity_hint);
    cudaStream_t st = S->stream;
    if (rc == DRRT_OK) rc = S->rec.reserve(capacity_hint, st);
    if (rc == DRRT_OK) rc = S->cell_of.reserve(capacity_hint, st);
    if (rc == DRRT_OK) rc = S->rank_in_cell.reserve(capacity_hint, st);
    if (rc == DRRT_OK) rc = S->cell_start.reserve(capacity_hint + 4096, st);
    if (rc == DRRT_OK) rc = S->scan_tmp.reserve(capacity_hint / 1024 + 4096, st);
    if (rc != DRRT_OK) { drrt_store_destroy(reinterpret_cast<drrt_store *>(S)); return rc; }
  }
  *out = reinterpret_cast<drrt_store *>(S);
  return DRRT_OK;
}

int drrt_store_destroy(drrt_store *h) {
  if (!h) return DRRT_OK;
  Store *S = reinterpret_cast<Store *>(h);
  {
    DeviceGuard guard(S->device);
    cudaStreamSynchronize(S->stream);
    S->x.release(); S->y.release(); S->th.release(); S->gate.release();
    S->parent.release(); S->left.release(); S->right.release(); S->split.release();
    S->ins_cur.release(); S->ins_split.release(); S->ins_gate.release(); S->stage_pos.release();
    S->rec.release(); S->cell_start.release(); S->cell_of.release(); S->rank_in_cell.release();
    S->scan_tmp.release();
    S->r_offsets.release(); S->r_ids.release(); S->r_dist.release();
    S->r_counts.release(); S->r_chunk_total.release(); S->r_chunk_base.release();
    S->r_offsets2.release(); S->r_ids2.release(); S->r_dist2.release();
    S->q_stage.release(); S->rad_stage.release();
    S->i_stage.release(); S->i_stage2.release(); S->d_stage.release(); S->d_stage2.release();
    S->d_stage3.release(); S->b_stage.release();
    S->o_kind.release(); S->o_voff.release(); S->oa_kind.release(); S->oa_voff.release();
    S->oa_live.release(); S->o_ox.release(); S->o_oy.release(); S->o_rad.release();
    S->o_vx.release(); S->o_vy.release(); S->oa_ox.release(); S->oa_oy.release();
    S->oa_rad.release(); S->oa_vx.release(); S->oa_vy.release(); S->o_mask.release(); S->o_covers.release();
    if (S->d_pending) cudaFree(S->d_pending);
    if (S->d_ticket) cudaFree(S->d_ticket);
    if (S->d_edge_counter) cudaFree(S->d_edge_counter);
    S->x_unsafe.release(); S->x_len.release(); S->x_lmc.release(); S->x_best.release(); S->x_row.release(); S->x_best_cost.release();
    S->edge_paths.release(); S->edge_masks.release(); S->edge_perm.release(); S->edge_bins.release(); S->og_off.release(); S->og_idx.release(); S->og_eobs.release();
    S->og_ball.release(); S->og_erec.release();
    S->e_u.release(); S->e_v.release(); S->node_mark.release(); S->sw_v1.release(); S->sw_v2.release();
    S->sw_out_flag.release(); S->sw_cnt.release(); S->sw_sel.release(); S->sw_u.release(); S->sw_v.release();
    S->sw_pick.release(); S->scan_tmp2.release(); S->sw_out_edge.release(); S->rw_lmc.release();
    S->tm_ox.release(); S->tm_oy.release(); S->tm_rad.release(); S->tm_path.release(); S->tm_poff.release();
    if (S->d_bounds) cudaFree(S->d_bounds);
    if (S->h_pending) cudaFreeHost(S->h_pending);
    if (S->h_bounds) cudaFreeHost(S->h_bounds);
    if (S->h_scalar) cudaFreeHost(S->h_scalar);
    if (S->h_map) cudaFreeHost(S->h_map);
    for (int b = 0; b < 2; b++) {
      S->p_off[b].release(); S->p_ids[b].release(); S->p_dist[b].release();
      if (S->ev_packed[b]) cudaEventDestroy(S->ev_packed[b]);
      if (S->ev_copied[b]) cudaEventDestroy(S->ev_copied[b]);
    }
    for (int b = 0; b < 2; b++) {
      if (S->ev_join[b]) cudaEventDestroy(S->ev_join[b]);
      if (S->side_stream[b]) cudaStreamDestroy(S->side_stream[b]);
      if (S->ev_up[b]) cudaEventDestroy(S->ev_up[b]);
      if (S->ev_done[b]) cudaEventDestroy(S->ev_done[b]);
    }
    if (S->ev_fork) cudaEventDestroy(S->ev_fork);
    if (S->down_stream) cudaStreamDestroy(S->down_stream);
    if (S->copy_stream) cudaStreamDestroy(S->copy_stream);
    if (S->stream) cudaStreamDestroy(S->stream);
  }
  delete S;
  return DRRT_OK;
}

int drrt_store_size(drrt_store *h, int64_t *n) {
  ENTER(h);
  if (!n) return set_error(DRRT_ERR_INVALID, "size: n == NULL");
  *n = S->n;
  return DRRT_OK;
}

void *drrt_store_stream(drrt_store *h) {
  return h ? (void *)reinterpret_cast<Store *>(h)->stream : nullptr;
}

int drrt_store_insert(drrt_store *h, int64_t n, const double *pos, int32_t *first_id) {
  ENTER(h);
  if (n < 0 || (n > 0 && !pos)) return set_error(DRRT_ERR_INVALID, "insert: arguments");
  for (int64_t i = 0; i < 3 * n; i++)
    if (!(pos[i] - pos[i] == 0.0))
      return set_error(DRRT_ERR_UNSUPPORTED, "insert: non-finite coordinate at %lld", (long long)i);
  cudaStream_t st = S->stream;
  DRRT_CHECK(S->stage_pos.reserve(3 * n, st));
  DRRT_CHECK(upload(pos, S->stage_pos.p, 3 * n * sizeof(double), st));
  DRRT_CHECK(store_insert_dev(S, n, S->stage_pos.p, st, first_id));
  DRRT_CUDA(cudaStreamSynchronize(st));
  return DRRT_OK;
}

int drrt_store_insert_dev(drrt_store *h, int64_t n, const double *pos_dev, void *stream,
                          int32_t *first_id) {
  ENTER(h);
  if (n < 0 || (n > 0 && !pos_dev)) return set_error(DRRT_ERR_INVALID, "insert: arguments");
  return store_insert_dev(S, n, pos_dev, pick(S, stream), first_id);
}

int drrt_store_structure(drrt_store *h, int64_t first, int64_t n, int32_t *parent, int32_t *left,
                         int32_t *right, int32_t *split) {
  ENTER(h);
  if (first < 0 || n < 0 || first + n > S->n) return set_error(DRRT_ERR_INVALID, "structure: range");
  cudaStream_t st = S->stream;
  std::vector<uint8_t> sp;
  if (parent) DRRT_CHECK(download(parent, S->parent.p + first, n * sizeof(int32_t), st));
  if (left) DRRT_CHECK(download(left, S->left.p + first, n * sizeof(int32_t), st));
  if (right) DRRT_CHECK(download(right, S->right.p + first, n * sizeof(int32_t), st));
  if (split) {
    sp.resize(n);
    DRRT_CHECK(download(sp.data(), S->split.p + first, n, st));
  }
  DRRT_CUDA(cudaStreamSynchronize(st));
  for (int64_t i = 0; i < n; i++) {
    if (left && left[i] == 0x7fffffff) left[i] = -1;
    if (right && right[i] == 0x7fffffff) right[i] = -1;
    if (split) split[i] = sp[i];
  }
  return DRRT_OK;
}

int drrt_store_positions(drrt_store *h, int64_t first, int64_t n, double *pos) {
  ENTER(h);
  if (first < 0 || n < 0 || first + n > S->n || (n > 0 && !pos))
    return set_error(DRRT_ERR_INVALID, "positions: range");
  cudaStream_t st = S->stream;
  std::vector<double> tmp(3 * (size_t)n);
  DRRT_CHECK(download(tmp.data(), S->x.p + first, n * sizeof(double), st));
  DRRT_CHECK(download(tmp.data() + n, S->y.p + first, n * sizeof(double), st));
  DRRT_CHECK(download(tmp.data() + 2 * n, S->th.p + first, n * sizeof(double), st));
  DRRT_CUDA(cudaStreamSynchronize(st));
  for (int64_t i = 0; i < n; i++) {
    pos[3 * i] = tmp[i];
    pos[3 * i + 1] = tmp[n + i];
    pos[3 * i + 2] = tmp[2 * n + i];
  }
  return DRRT_OK;
}

// ---- range ----------------------------------------------------------------

int drrt_range_batch(drrt_store *h, int64_t nq, const double *q, const double *radius,
                     double radius_all, int32_t *counts, int64_t *total) {
  ENTER(h);
  if (nq < 0 || (nq > 0 && (!q || !counts))) return set_error(DRRT_ERR_INVALID, "range: arguments");
  cudaStream_t st = S->stream;
  DRRT_CHECK(S->q_stage.reserve(3 * nq + 1, st));
  DRRT_CHECK(upload(q, S->q_stage.p, 3 * nq * sizeof(double), st));
  const double *rad_dev = nullptr;
  if (radius) {
    DRRT_CHECK(S->rad_stage.reserve(nq + 1, st));
    DRRT_CHECK(upload(radius, S->rad_stage.p, nq * sizeof(double), st));
    rad_dev = S->rad_stage.p;
  }
  drrt_range_result res{};
  DRRT_CHECK(range_batch_dev(S, nq, S->q_stage.p, rad_dev, radius_all, st, &res));
  if (nq > 0) DRRT_CHECK(download(counts, res.counts_dev, nq * sizeof(int32_t), st));
  DRRT_CUDA(cudaStreamSynchronize(st));
  if (total) *total = res.total;
  return DRRT_OK;
}

int drrt_range_fetch(drrt_store *h, int32_t *ids, double *dist) {
  ENTER(h);
  if (S->last_nq < 0) return set_error(DRRT_ERR_STATE, "fetch: no range batch on this handle");
  cudaStream_t st = S->stream;
  if (S->last_total > 0) {
    // the caller's layout is the plain CSR given by the counts
    drrt_range_result res{};
    DRRT_CHECK(range_pack_dev(S, st, &res));
    if (ids) DRRT_CHECK(download(ids, res.ids_dev, S->last_total * sizeof(int32_t), st));
    if (dist) DRRT_CHECK(download(dist, res.dist_dev, S->last_total * sizeof(double), st));
  }
  DRRT_CUDA(cudaStreamSynchronize(st));
  return DRRT_OK;
}

// Sub-batch of the streamed host query: large enough that the CTA kernel still takes
// full-length slices on every resident CTA, small enough that the first rows leave early.
static constexpr int64_t INTO_SUB = 16384;

}  // extern "C"

namespace drrt {

// The streamed host query behind drrt_range_batch_into[_f32] (and the per-GPU leg of
// drrt_comm_range_batch_into).  dist_kind picks what crosses PCIe per hit beside the 4-byte id:
// the double key (8 B, bit-exact), a float key (4 B) or nothing.
int range_into(Store *S, int64_t nq, const double *q, const double *radius, double radius_all,
               int64_t cap, int64_t *offsets, int32_t *ids, void *dist, int dist_kind,
               int64_t *total) {
  if (nq < 0 || cap < 0 || (nq > 0 && (!q || !offsets)) || (cap > 0 && !ids))
    return set_error(DRRT_ERR_INVALID, "range_into: arguments");
  if (!dist) dist_kind = DIST_NONE;
  const size_t dsz = dist_kind == DIST_F64 ? sizeof(double) : (dist_kind == DIST_F32 ? sizeof(float) : 0);
  cudaStream_t st = S->stream;
  if (nq == 1) {
    int64_t tot = 0;
    const int32_t *mi = nullptr;
    const double *md = nullptr;
    int rc = range_one_mapped(S, q, radius ? radius[0] : radius_all, &tot, &mi, &md);
    if (rc == DRRT_OK) {
      offsets[0] = 0;
      offsets[1] = tot;
      if (total) *total = tot;
      if (tot > cap)
        return set_error(DRRT_ERR_CAPACITY, "range_into: %lld hits, caller buffers hold %lld",
                         (long long)tot, (long long)cap);
      if (tot > 0) {
        memcpy(ids, mi, tot * sizeof(int32_t));
        if (dist_kind == DIST_F64) memcpy(dist, md, tot * sizeof(double));
        else if (dist_kind == DIST_F32)
          for (int64_t i = 0; i < tot; i++) ((float *)dist)[i] = (float)md[i];
      }
      return DRRT_OK;
    }
    if (rc != RANGE_NOT_HANDLED) return rc;
  }
  DRRT_CHECK(S->q_stage.reserve(3 * nq + 1, st));
  DRRT_CHECK(upload(q, S->q_stage.p, 3 * nq * sizeof(double), st));
  if (radius) {
    DRRT_CHECK(S->rad_stage.reserve(nq + 1, st));
    DRRT_CHECK(upload(radius, S->rad_stage.p, nq * sizeof(double), st));
  }
  if (offsets) offsets[0] = 0;
  if (nq <= INTO_SUB + INTO_SUB / 2) {
    // One sub-batch (the planner's own call is nq == 1, src/drrt.cpp:213-216): nothing to overlap,
    // so everything stays on the handle's stream -- one kernel, one completion wait inside
    // range_batch_dev, the copies, one wait.  A lone slice is already a plain CSR and its double
    // rows leave straight from the arena.
    drrt_range_result res{};
    int rc = range_batch_dev(S, nq, S->q_stage.p, radius ? S->rad_stage.p : nullptr, radius_all, st, &res);
    if (rc == DRRT_OK && nq > 0) {
      const bool fits = res.total <= cap;
      if (res.packed && dist_kind != DIST_F32) {
        rc = download(offsets, res.offsets_dev, (nq + 1) * sizeof(int64_t), st);
        if (rc == DRRT_OK && fits && res.total > 0) {
          rc = download(ids, res.ids_dev, res.total * sizeof(int32_t), st);
          if (rc == DRRT_OK && dsz) rc = download(dist, res.dist_dev, res.total * dsz, st);
        }
      } else {
        const size_t need = (size_t)res.total + 1;
        rc = S->p_off[0].reserve(nq + 1, st);
        if (rc == DRRT_OK) rc = S->p_ids[0].reserve(need, st);
        if (rc == DRRT_OK && dsz) rc = S->p_dist[0].reserve(need, st);
        if (rc == DRRT_OK)
          rc = range_pack_to(S, st, S->p_off[0].p, S->p_ids[0].p, S->p_dist[0].p, dist_kind, 0);
        if (rc == DRRT_OK) rc = download(offsets, S->p_off[0].p, (nq + 1) * sizeof(int64_t), st);
        if (rc == DRRT_OK && fits && res.total > 0) {
          rc = download(ids, S->p_ids[0].p, res.total * sizeof(int32_t), st);
          if (rc == DRRT_OK && dsz) rc = download(dist, S->p_dist[0].p, res.total * dsz, st);
        }
      }
    }
    cudaError_t e = cudaStreamSynchronize(st);
    S->last_nq = -1;
    if (rc != DRRT_OK) return rc;
    DRRT_CUDA(e);
    if (total) *total = res.total;
    if (res.total > cap)
      return set_error(DRRT_ERR_CAPACITY, "range_into: %lld hits, caller buffers hold %lld",
                       (long long)res.total, (long long)cap);
    return DRRT_OK;
  }
  // (the copy stream is shared with the pipelined edge checks, which may have created it first:
  // the events are checked on their own)
  if (!S->copy_stream) DRRT_CUDA(cudaStreamCreateWithFlags(&S->copy_stream, cudaStreamNonBlocking));
  for (int b = 0; b < 2; b++) {
    if (!S->ev_packed[b]) DRRT_CUDA(cudaEventCreateWithFlags(&S->ev_packed[b], cudaEventDisableTiming));
    if (!S->ev_copied[b]) DRRT_CUDA(cudaEventCreateWithFlags(&S->ev_copied[b], cudaEventDisableTiming));
  }
  cudaStream_t cs = S->copy_stream;
  const int64_t sub = INTO_SUB;
  int64_t base = 0;
  int rc = DRRT_OK;
  bool used[2] = {false, false};
  // one sub-batch: query, pack into buffer b (offsets already global: the scan starts at `base`),
  // queue the copies; an error leaves the loop but the copies already queued into the caller's
  // buffers are always waited for below
  auto sub_batch = [&](int64_t k0, int64_t nb, int b) -> int {
    drrt_range_result res{};
    DRRT_CHECK(range_batch_dev(S, nb, S->q_stage.p + 3 * k0, radius ? S->rad_stage.p + k0 : nullptr,
                               radius_all, st, &res));
    // buffer b is free once the copies of sub-batch k - 2 have left it (growing it frees the
    // old block, so wait on the host in that case; otherwise the stream waits)
    const size_t need = (size_t)res.total + 1;
    if (used[b]) {
      if (need > S->p_ids[b].cap || (dsz && need > S->p_dist[b].cap) || (size_t)nb + 1 > S->p_off[b].cap)
        DRRT_CUDA(cudaEventSynchronize(S->ev_copied[b]));
      else
        DRRT_CUDA(cudaStreamWaitEvent(st, S->ev_copied[b], 0));
    }
    DRRT_CHECK(S->p_off[b].reserve(nb + 1, st));
    DRRT_CHECK(S->p_ids[b].reserve(need, st));
    if (dsz) DRRT_CHECK(S->p_dist[b].reserve(need, st));
    DRRT_CHECK(range_pack_to(S, st, S->p_off[b].p, S->p_ids[b].p, S->p_dist[b].p, dist_kind, base));
    DRRT_CUDA(cudaEventRecord(S->ev_packed[b], st));
    DRRT_CUDA(cudaStreamWaitEvent(cs, S->ev_packed[b], 0));
    // (the entry a sub-batch shares with its successor is written twice with the same value)
    DRRT_CHECK(download(offsets + k0, S->p_off[b].p, (nb + 1) * sizeof(int64_t), cs));
    if (base + res.total <= cap && res.total > 0) {
      DRRT_CHECK(download(ids + base, S->p_ids[b].p, res.total * sizeof(int32_t), cs));
      if (dsz) DRRT_CHECK(download((char *)dist + base * dsz, S->p_dist[b].p, res.total * dsz, cs));
    }
    DRRT_CUDA(cudaEventRecord(S->ev_copied[b], cs));
    used[b] = true;
    base += res.total;
    return DRRT_OK;
  };
  for (int64_t k0 = 0, k = 0; k0 < nq && rc == DRRT_OK; k0 += sub, k++)
    rc = sub_batch(k0, std::min(sub, nq - k0), (int)(k & 1));
  cudaError_t e1 = cudaStreamSynchronize(cs), e2 = cudaStreamSynchronize(st);
  S->last_nq = -1;  // the handle holds one sub-batch only: nothing to fetch or pack
  if (rc != DRRT_OK) return rc;
  DRRT_CUDA(e1);
  DRRT_CUDA(e2);
  if (total) *total = base;
  if (base > cap)
    return set_error(DRRT_ERR_CAPACITY, "range_into: %lld hits, caller buffers hold %lld",
                     (long long)base, (long long)cap);
  return DRRT_OK;
}

}  // namespace drrt

extern "C" {

int drrt_range_batch_into(drrt_store *h, int64_t nq, const double *q, const double *radius,
                          double radius_all, int64_t cap, int64_t *offsets, int32_t *ids,
                          double *dist, int64_t *total) {
  ENTER(h);
  return range_into(S, nq, q, radius, radius_all, cap, offsets, ids, dist, DIST_F64, total);
}

int drrt_range_batch_into_f32(drrt_store *h, int64_t nq, const double *q, const double *radius,
                              double radius_all, int64_t cap, int64_t *offsets, int32_t *ids,
                              float *dist, int64_t *total) {
  ENTER(h);
  return range_into(S, nq, q, radius, radius_all, cap, offsets, ids, dist, DIST_F32, total);
}

int drrt_range_pack(drrt_store *h, void *stream, drrt_range_result *out) {
  ENTER(h);
  if (!out) return set_error(DRRT_ERR_INVALID, "pack: arguments");
  return range_pack_dev(S, pick(S, stream), out);
}

int drrt_range_batch_dev(drrt_store *h, int64_t nq, const double *q_dev, const double *radius_dev,
                         double radius_all, void *stream, drrt_range_result *out) {
  ENTER(h);
  if (nq < 0 || (nq > 0 && !q_dev)) return set_error(DRRT_ERR_INVALID, "range: arguments");
  return range_batch_dev(S, nq, q_dev, radius_dev, radius_all, pick(S, stream), out);
}

// ---- nearest / knn --------------------------------------------------------

int drrt_nearest_batch(drrt_store *h, int64_t nq, const double *q, int32_t *ids, double *dist) {
  ENTER(h);
  return knn_host(S, nq, q, 1, ids, dist, true);
}
int drrt_nearest_batch_dev(drrt_store *h, int64_t nq, const double *q_dev, int32_t *ids_dev,
                           double *dist_dev, void *stream) {
  ENTER(h);
  if (!q_dev || !ids_dev || !dist_dev) return set_error(DRRT_ERR_INVALID, "nearest: arguments");
  return knn_batch_dev(S, nq, q_dev, 1, ids_dev, dist_dev, pick(S, stream), true);
}
int drrt_knn_batch(drrt_store *h, int64_t nq, const double *q, int k, int32_t *ids, double *dist) {
  ENTER(h);
  return knn_host(S, nq, q, k, ids, dist, false);
}
int drrt_knn_batch_dev(drrt_store *h, int64_t nq, const double *q_dev, int k, int32_t *ids_dev,
                       double *dist_dev, void *stream) {
  ENTER(h);
  if (!q_dev || !ids_dev || !dist_dev) return set_error(DRRT_ERR_INVALID, "knn: arguments");
  return knn_batch_dev(S, nq, q_dev, k, ids_dev, dist_dev, pick(S, stream));
}

// ---- obstacles ------------------------------------------------------------

}  // extern "C"

namespace drrt {

// Derives the device tables from the host copy of the caller's obstacle table (S->t_*): the
// "active" list (live balls and polygons, the only obstacles that can ever answer true,
// obstacle.cpp:761, 781-804), the whole table for the point checks and the sweeps, and the
// per-obstacle `covers` flag; marks the broadphase grid stale.
int obstacles_apply(Store *S) {
  const int n_obs = (int)S->t_kind.size();
  const std::vector<int32_t> &kind = S->t_kind, &vert_off = S->t_voff;
  const int nv_all = n_obs ? vert_off[n_obs] : 0;
  std::vector<int32_t> k_act, vo_act(1, 0), src_act;
  std::vector<uint8_t> live(n_obs), covers_act;
  std::vector<double> ox(n_obs), oy(n_obs), vx(nv_all), vy(nv_all);
  std::vector<double> aox, aoy, arad, avx, avy;
  for (int v = 0; v < nv_all; v++) { vx[v] = S->t_verts[2 * v]; vy[v] = S->t_verts[2 * v + 1]; }
  for (int i = 0; i < n_obs; i++) {
    ox[i] = S->t_origin[2 * i];
    oy[i] = S->t_origin[2 * i + 1];
    live[i] = (S->t_used[i] && !(S->t_life[i] <= 0)) ? 1 : 0;  // obstacle.cpp:761
    // only a live ball or polygon can ever answer true (obstacle.cpp:781-804)
    if (live[i] && (kind[i] == 1 || kind[i] == 3)) {
      src_act.push_back(i);
      k_act.push_back(kind[i]);
      aox.push_back(ox[i]); aoy.push_back(oy[i]); arad.push_back(S->t_radius[i]);
      for (int v = vert_off[i]; v < vert_off[i + 1]; v++) { avx.push_back(vx[v]); avy.push_back(vy[v]); }
      vo_act.push_back((int32_t)avx.size());
      double far = 0.0;  // obstacle.h:134-141 sets radius_ to this; a caller may not
      for (int v = vert_off[i]; v < vert_off[i + 1]; v++)
        far = std::max(far, sqrt((vx[v] - ox[i]) * (vx[v] - ox[i]) + (vy[v] - oy[i]) * (vy[v] - oy[i])));
      covers_act.push_back(S->t_radius[i] * (1.0 + 1e-12) >= far ? 1 : 0);
    }
  }
  cudaStream_t st = S->stream;
  DRRT_CUDA(cudaStreamSynchronize(st));
  DRRT_CHECK(put(S->o_kind, k_act, st)); DRRT_CHECK(put(S->o_voff, vo_act, st));
  DRRT_CHECK(put(S->o_ox, aox, st)); DRRT_CHECK(put(S->o_oy, aoy, st));
  DRRT_CHECK(put(S->o_rad, arad, st)); DRRT_CHECK(put(S->o_vx, avx, st));
  DRRT_CHECK(put(S->o_vy, avy, st));
  DRRT_CHECK(put(S->o_covers, covers_act, st));
  DRRT_CHECK(put(S->oa_kind, S->t_kind, st)); DRRT_CHECK(put(S->oa_voff, S->t_voff, st));
  DRRT_CHECK(put(S->oa_live, live, st));
  DRRT_CHECK(put(S->oa_ox, ox, st)); DRRT_CHECK(put(S->oa_oy, oy, st));
  DRRT_CHECK(put(S->oa_rad, S->t_radius, st)); DRRT_CHECK(put(S->oa_vx, vx, st));
  DRRT_CHECK(put(S->oa_vy, vy, st));
  DRRT_CUDA(cudaStreamSynchronize(st));  // host vectors die at return
  ObstacleTable &T = S->obs;
  T.n_active = (int32_t)k_act.size();
  T.n_active_verts = (int32_t)avx.size();
  T.kind = S->o_kind.p; T.ox = S->o_ox.p; T.oy = S->o_oy.p; T.rad = S->o_rad.p;
  T.vert_off = S->o_voff.p; T.vx = S->o_vx.p; T.vy = S->o_vy.p;
  T.n_all = n_obs;
  T.a_kind = S->oa_kind.p; T.a_live = S->oa_live.p; T.a_ox = S->oa_ox.p; T.a_oy = S->oa_oy.p;
  T.a_rad = S->oa_rad.p; T.a_vert_off = S->oa_voff.p; T.a_vx = S->oa_vx.p; T.a_vy = S->oa_vy.p;
  S->h_o_ox = aox; S->h_o_oy = aoy; S->h_o_rad = arad; S->h_o_vx = avx; S->h_o_vy = avy;
  S->h_o_kind = k_act; S->h_o_voff = vo_act; S->h_o_src = src_act;
  S->h_all_ox = ox; S->h_all_oy = oy; S->h_all_rad = S->t_radius; S->h_all_kind = S->t_kind;
  S->ogrid_radius = -1.0;  // broadphase grid is stale
  S->has_obstacles = true;
  return DRRT_OK;
}

}  // namespace drrt

extern "C" {

int drrt_obstacles_set(drrt_store *h, int32_t n_obs, const int32_t *kind, const uint8_t *used,
                       const double *life_span, const double *origin_xy, const double *radius,
                       const int32_t *vert_off, const double *verts_xy) {
  ENTER(h);
  if (n_obs < 0) return set_error(DRRT_ERR_INVALID, "obstacles: n_obs < 0");
  if (n_obs > 0 && (!kind || !used || !life_span || !origin_xy || !radius || !vert_off))
    return set_error(DRRT_ERR_INVALID, "obstacles: null array");
  for (int i = 0; i < n_obs; i++) {
    if (kind[i] == 6 || kind[i] == 7)
      return set_error(DRRT_ERR_UNSUPPORTED,
                       "obstacles: kind %d (time-parametrised) belongs in drrt_timed_obstacles_set", kind[i]);
    if (vert_off[i + 1] < vert_off[i]) return set_error(DRRT_ERR_INVALID, "obstacles: vert_off");
  }
  const int nv_all = n_obs ? vert_off[n_obs] : 0;
  if (nv_all > 0 && !verts_xy) return set_error(DRRT_ERR_INVALID, "obstacles: verts");
  // the host copy of the caller's table: drrt_obstacles_move / drrt_obstacles_update patch it
  S->t_kind.assign(kind, kind + n_obs);
  S->t_used.assign(used, used + n_obs);
  S->t_life.assign(life_span, life_span + n_obs);
  S->t_origin.assign(origin_xy, origin_xy + 2 * (size_t)n_obs);
  S->t_radius.assign(radius, radius + n_obs);
  S->t_voff.assign(n_obs + 1, 0);
  for (int i = 0; i <= n_obs; i++) S->t_voff[i] = n_obs ? vert_off[i] : 0;
  S->t_verts.assign(verts_xy, verts_xy + 2 * (size_t)nv_all);
  return obstacles_apply(S);
}

// Obstacle::UpdateObstacles -> NextOrigin (src/obstacle.cpp:12-22, 182-195): a moving obstacle's
// origin_ jumps to the next point of its path.  Only the listed entries cross the boundary.
//   translate_polygon == 0: origin_ alone moves -- exactly what the analytic checker then sees
//     (it reads shape_.GetPolygon() as world coordinates, obstacle.cpp:785-794): the
//     bounding reject (:765-779) follows the new origin, the polygon stays;
//   translate_polygon != 0: the vertices move with the origin (the frame GetPosition(),
//     obstacle.cpp:9-10, the Bullet hull and the visualiser use).
int drrt_obstacles_move(drrt_store *h, int32_t n_sel, const int32_t *obstacle_index,
                        const double *new_origin_xy, int translate_polygon) {
  ENTER(h);
  if (n_sel < 0 || (n_sel > 0 && (!obstacle_index || !new_origin_xy)))
    return set_error(DRRT_ERR_INVALID, "obstacles_move: arguments");
  if (!S->has_obstacles) return set_error(DRRT_ERR_STATE, "obstacles_move: no obstacle table");
  const int n_obs = (int)S->t_kind.size();
  for (int i = 0; i < n_sel; i++)
    if (obstacle_index[i] < 0 || obstacle_index[i] >= n_obs)
      return set_error(DRRT_ERR_INVALID, "obstacles_move: obstacle %d", obstacle_index[i]);
  for (int i = 0; i < n_sel; i++) {
    const int o = obstacle_index[i];
    const double nx = new_origin_xy[2 * i], ny = new_origin_xy[2 * i + 1];
    if (!(nx - nx == 0.0) || !(ny - ny == 0.0))
      return set_error(DRRT_ERR_UNSUPPORTED, "obstacles_move: non-finite origin");
    if (translate_polygon) {
      const double dx = nx - S->t_origin[2 * o], dy = ny - S->t_origin[2 * o + 1];
      for (int v = S->t_voff[o]; v < S->t_voff[o + 1]; v++) {
        S->t_verts[2 * v] += dx;
        S->t_verts[2 * v + 1] += dy;
      }
    }
    S->t_origin[2 * o] = nx;
    S->t_origin[2 * o + 1] = ny;
  }
  return n_sel ? obstacles_apply(S) : DRRT_OK;
}

// obstacle_used_ / life_span_ of listed obstacles (the obstacle thread switches obstacles on and
// off, obstacle.cpp:454-499; RemoveObstacle ends with obstacle_used_ = false, :753)
int drrt_obstacles_update(drrt_store *h, int32_t n_sel, const int32_t *obstacle_index,
                          const uint8_t *used, const double *life_span) {
  ENTER(h);
  if (n_sel < 0 || (n_sel > 0 && (!obstacle_index || (!used && !life_span))))
    return set_error(DRRT_ERR_INVALID, "obstacles_update: arguments");
  if (!S->has_obstacles) return set_error(DRRT_ERR_STATE, "obstacles_update: no obstacle table");
  const int n_obs = (int)S->t_kind.size();
  for (int i = 0; i < n_sel; i++)
    if (obstacle_index[i] < 0 || obstacle_index[i] >= n_obs)
      return set_error(DRRT_ERR_INVALID, "obstacles_update: obstacle %d", obstacle_index[i]);
  for (int i = 0; i < n_sel; i++) {
    if (used) S->t_used[obstacle_index[i]] = used[i];
    if (life_span) S->t_life[obstacle_index[i]] = life_span[i];
  }
  return n_sel ? obstacles_apply(S) : DRRT_OK;
}

// ---- edge checks ----------------------------------------------------------

int drrt_edgecheck_batch(drrt_store *h, int64_t n, const double *start, const double *end,
                         int edge_kind, double robot_radius, double r_min, int in_warmup,
                         uint8_t *verdict, double *length) {
  ENTER(h);
  return edge_host(S, n, start, end, nullptr, nullptr, edge_kind, robot_radius, r_min, in_warmup,
                   verdict, length);
}

int drrt_edgecheck_ids_batch(drrt_store *h, int64_t n, const int32_t *start_id,
                             const int32_t *end_id, int edge_kind, double robot_radius,
                             double r_min, int in_warmup, uint8_t *verdict, double *length) {
  ENTER(h);
  if (n > 0 && (!start_id || !end_id)) return set_error(DRRT_ERR_INVALID, "edgecheck: ids");
  return edge_host(S, n, nullptr, nullptr, start_id, end_id, edge_kind, robot_radius, r_min,
                   in_warmup, verdict, length);
}

int drrt_edgecheck_batch_dev(drrt_store *h, int64_t n, const double *start_dev,
                             const double *end_dev, const int32_t *start_id_dev,
                             const int32_t *end_id_dev, int edge_kind, double robot_radius,
                             double r_min, uint8_t *verdict_dev, double *length_dev, void *stream) {
  ENTER(h);
  return edgecheck_batch_dev(S, n, start_dev, end_dev, start_id_dev, end_id_dev, edge_kind,
                             robot_radius, r_min, verdict_dev, length_dev, pick(S, stream));
}

// ---- obstacle sweeps -----------------------------------

// the caller's mask over the obstacle TABLE -> mask over the active (live ball / polygon) list
static int stage_obstacle_mask(Store *S, const uint8_t *mask, int32_t n_obstacles, cudaStream_t st) {
  if (!mask) return set_error(DRRT_ERR_INVALID, "subset check: obstacle_mask == NULL");
  if (!S->has_obstacles || n_obstacles != S->obs.n_all)
    return set_error(DRRT_ERR_INVALID, "subset check: mask has %d entries, the obstacle table %d",
                     n_obstacles, S->has_obstacles ? S->obs.n_all : 0);
  std::vector<uint8_t> act(S->h_o_src.size() + 1, 0);
  for (size_t a = 0; a < S->h_o_src.size(); a++) act[a] = mask[S->h_o_src[a]] ? 1 : 0;
  DRRT_CHECK(S->o_mask.reserve(act.size(), st));
  DRRT_CUDA(cudaMemcpyAsync(S->o_mask.p, act.data(), act.size(), cudaMemcpyHostToDevice, st));
  DRRT_CUDA(cudaStreamSynchronize(st));  // `act` dies at return
  return DRRT_OK;
}

int drrt_edgecheck_subset_batch(drrt_store *h, int64_t n, const double *start, const double *end,
                                const int32_t *start_id, const int32_t *end_id, int edge_kind,
                                double robot_radius, double r_min, const uint8_t *obstacle_mask,
                                int32_t n_obstacles, uint8_t *verdict, double *length) {
  ENTER(h);
  DRRT_CHECK(stage_obstacle_mask(S, obstacle_mask, n_obstacles, S->stream));
  S->edge_mask = S->o_mask.p;
  int rc = edge_host(S, n, start, end, start_id, end_id, edge_kind, robot_radius, r_min, 0, verdict,
                     length);
  S->edge_mask = nullptr;
  return rc;
}

int drrt_edgecheck_subset_batch_dev(drrt_store *h, int64_t n, const double *start_dev,
                                    const double *end_dev, const int32_t *start_id_dev,
                                    const int32_t *end_id_dev, int edge_kind, double robot_radius,
                                    double r_min, const uint8_t *obstacle_mask, int32_t n_obstacles,
                                    uint8_t *verdict_dev, double *length_dev, void *stream) {
  ENTER(h);
  cudaStream_t st = pick(S, stream);
  DRRT_CHECK(stage_obstacle_mask(S, obstacle_mask, n_obstacles, st));
  S->edge_mask = S->o_mask.p;
  int rc = edgecheck_batch_dev(S, n, start_dev, end_dev, start_id_dev, end_id_dev, edge_kind,
                               robot_radius, r_min, verdict_dev, length_dev, st);
  S->edge_mask = nullptr;
  return rc;
}

int drrt_obstacle_conflict_batch(drrt_store *h, int32_t n_sel, const int32_t *obstacle_index,
                                 double robot_radius, int32_t *counts, int64_t *total) {
  ENTER(h);
  if (n_sel < 0 || (n_sel > 0 && (!obstacle_index || !counts)))
    return set_error(DRRT_ERR_INVALID, "conflict: arguments");
  if (n_sel > 0 && !S->has_obstacles) return set_error(DRRT_ERR_STATE, "conflict: no obstacle table");
  // FindPointsInConflictWithObstacle, Dubins space without time (obstacle.cpp:554-559):
  // KDFindWithinRange(robot_radius_ + O->radius_ + PI, (O->origin_(0), O->origin_(1), PI))
  std::vector<double> q(3 * (size_t)n_sel + 3), r((size_t)n_sel + 1);
  for (int i = 0; i < n_sel; i++) {
    const int o = obstacle_index[i];
    if (o < 0 || o >= S->obs.n_all) return set_error(DRRT_ERR_INVALID, "conflict: obstacle %d", o);
    const int kind = S->h_all_kind[o];
    if (kind < 1 || kind > 5)  // :547, :608 "This case has not been coded yet."
      return set_error(DRRT_ERR_UNSUPPORTED, "conflict: obstacle kind %d", kind);
    q[3 * i] = S->h_all_ox[o];
    q[3 * i + 1] = S->h_all_oy[o];
    q[3 * i + 2] = DRRT_PI;
    r[i] = robot_radius + S->h_all_rad[o] + DRRT_PI;
  }
  cudaStream_t st = S->stream;
  DRRT_CHECK(S->q_stage.reserve(3 * (size_t)n_sel + 3, st));
  DRRT_CHECK(S->rad_stage.reserve((size_t)n_sel + 1, st));
  DRRT_CHECK(upload(q.data(), S->q_stage.p, 3 * (size_t)n_sel * sizeof(double), st));
  DRRT_CHECK(upload(r.data(), S->rad_stage.p, (size_t)n_sel * sizeof(double), st));
  DRRT_CUDA(cudaStreamSynchronize(st));  // q, r die at return
  drrt_range_result res{};
  DRRT_CHECK(range_batch_dev(S, n_sel, S->q_stage.p, S->rad_stage.p, 0.0, st, &res));
  if (n_sel > 0) DRRT_CHECK(download(counts, res.counts_dev, (size_t)n_sel * sizeof(int32_t), st));
  DRRT_CUDA(cudaStreamSynchronize(st));
  if (total) *total = res.total;
  return DRRT_OK;
}

int drrt_extend_batch(drrt_store *h, int64_t nq, const double *q, const double *radius,
                      double radius_all, double robot_radius, double r_min, int in_warmup,
                      const double *lmc, int32_t *counts, int64_t *total, int32_t *best_parent,
                      double *best_cost) {
  ENTER(h);
  if (nq < 0 || (nq > 0 && (!q || !counts))) return set_error(DRRT_ERR_INVALID, "extend: arguments");
  cudaStream_t st = S->stream;
  S->last_extend_total = -1;
  DRRT_CHECK(S->q_stage.reserve(3 * nq + 1, st));
  DRRT_CHECK(upload(q, S->q_stage.p, 3 * nq * sizeof(double), st));
  const double *rad_dev = nullptr;
  if (radius) {
    DRRT_CHECK(S->rad_stage.reserve(nq + 1, st));
    DRRT_CHECK(upload(radius, S->rad_stage.p, nq * sizeof(double), st));
    rad_dev = S->rad_stage.p;
  }
  const double *lmc_dev = nullptr;
  if (best_parent) {
    if (lmc) {
      // the whole array from the host (the mirror now equals it)
      DRRT_CHECK(S->x_lmc.reserve(S->n + 1, st));
      DRRT_CHECK(upload(lmc, S->x_lmc.p, S->n * sizeof(double), st));
      S->lmc_n = S->n;
    } else {
      // the device-resident mirror kept current by drrt_lmc_write / drrt_lmc_scatter:
      // nothing store-sized crosses PCIe
      DRRT_CHECK(lmc_mirror_cover(S, st));
    }
    DRRT_CHECK(S->x_best.reserve(nq + 1, st));
    DRRT_CHECK(S->x_best_cost.reserve(nq + 1, st));
    lmc_dev = S->x_lmc.p;
  }
  drrt_range_result res{};
  uint8_t *unsafe_dev = nullptr;
  double *len_dev = nullptr;
  bool saved = S->has_obstacles;
  if (in_warmup) S->has_obstacles = false;  // obstacle.cpp:883
  int rc = extend_batch_dev(S, nq, S->q_stage.p, rad_dev, radius_all, robot_radius, r_min, lmc_dev, st,
                            &res, &unsafe_dev, &len_dev, lmc_dev ? S->x_best.p : nullptr,
                            lmc_dev ? S->x_best_cost.p : nullptr);
  S->has_obstacles = saved;
  DRRT_CHECK(rc);
  if (nq > 0) {
    DRRT_CHECK(download(counts, res.counts_dev, nq * sizeof(int32_t), st));
    if (lmc_dev) {
      DRRT_CHECK(download(best_parent, S->x_best.p, nq * sizeof(int32_t), st));
      if (best_cost) DRRT_CHECK(download(best_cost, S->x_best_cost.p, nq * sizeof(double), st));
    }
  }
  DRRT_CUDA(cudaStreamSynchronize(st));
  if (total) *total = res.total;
  S->last_extend_total = res.total;
  return DRRT_OK;
}

int drrt_extend_fetch(drrt_store *h, int32_t *ids, double *dist, uint8_t *unsafe, double *length) {
  ENTER(h);
  if (S->last_extend_total < 0 || (S->last_extend_total > 0 && (!S->csr_ids || !S->csr_dist)))
    return set_error(DRRT_ERR_STATE, "fetch: no extend batch on this handle (or a later range call "
                                     "replaced its rows)");
  cudaStream_t st = S->stream;
  const int64_t t = S->last_extend_total;
  if (t > 0) {
    if (ids) DRRT_CHECK(download(ids, S->csr_ids, t * sizeof(int32_t), st));
    if (dist) DRRT_CHECK(download(dist, S->csr_dist, t * sizeof(double), st));
    if (unsafe) DRRT_CHECK(download(unsafe, S->x_unsafe.p, 2 * t, st));
    if (length) DRRT_CHECK(download(length, S->x_len.p, 2 * t * sizeof(double), st));
  }
  DRRT_CUDA(cudaStreamSynchronize(st));
  return DRRT_OK;
}

int drrt_extend_batch_dev(drrt_store *h, int64_t nq, const double *q_dev, const double *radius_dev,
                          double radius_all, double robot_radius, double r_min,
                          const double *lmc_dev, void *stream, drrt_range_result *rows,
                          const uint8_t **unsafe_dev, const double **length_dev,
                          int32_t *best_parent_dev, double *best_cost_dev) {
  ENTER(h);
  if (nq < 0 || (nq > 0 && !q_dev) || !rows) return set_error(DRRT_ERR_INVALID, "extend: arguments");
  uint8_t *u = nullptr;
  double *l = nullptr;
  DRRT_CHECK(extend_batch_dev(S, nq, q_dev, radius_dev, radius_all, robot_radius, r_min, lmc_dev,
                              pick(S, stream), rows, &u, &l, best_parent_dev, best_cost_dev));
  if (unsafe_dev) *unsafe_dev = u;
  if (length_dev) *length_dev = l;
  S->last_extend_total = rows->total;  // rows stay valid until the next range / extend call
  return DRRT_OK;
}

// ---- device mirror of rrt_LMC_ ---------------------------------------------

int drrt_lmc_write(drrt_store *h, int64_t first, int64_t n, const double *values) {
  ENTER(h);
  if (first < 0 || n < 0 || first + n > S->n || (n > 0 && !values))
    return set_error(DRRT_ERR_INVALID, "lmc_write: range [%lld, %lld) of a %lld-node store",
                     (long long)first, (long long)(first + n), (long long)S->n);
  cudaStream_t st = S->stream;
  DRRT_CHECK(lmc_mirror_cover(S, st));
  DRRT_CHECK(upload(values, S->x_lmc.p + first, n * sizeof(double), st));
  DRRT_CUDA(cudaStreamSynchronize(st));
  return DRRT_OK;
}

int drrt_lmc_scatter(drrt_store *h, int64_t n, const int32_t *ids, const double *values) {
  ENTER(h);
  if (n < 0 || (n > 0 && (!ids || !values))) return set_error(DRRT_ERR_INVALID, "lmc_scatter: arguments");
  for (int64_t i = 0; i < n; i++)
    if (ids[i] < 0 || ids[i] >= S->n)
      return set_error(DRRT_ERR_INVALID, "lmc_scatter: node id %d at %lld", ids[i], (long long)i);
  if (n == 0) return DRRT_OK;
  cudaStream_t st = S->stream;
  DRRT_CHECK(lmc_mirror_cover(S, st));
  DRRT_CHECK(S->i_stage.reserve(n, st));
  DRRT_CHECK(S->d_stage.reserve(n, st));
  DRRT_CHECK(upload(ids, S->i_stage.p, n * sizeof(int32_t), st));
  DRRT_CHECK(upload(values, S->d_stage.p, n * sizeof(double), st));
  scatter_f64_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(
      S->x_lmc.p, n, S->i_stage.p, S->d_stage.p);
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaGetLastError());
  DRRT_CUDA(cudaStreamSynchronize(st));
  return DRRT_OK;
}

int drrt_lmc_read(drrt_store *h, int64_t first, int64_t n, double *values) {
  ENTER(h);
  if (first < 0 || n < 0 || first + n > S->n || (n > 0 && !values))
    return set_error(DRRT_ERR_INVALID, "lmc_read: range");
  cudaStream_t st = S->stream;
  DRRT_CHECK(lmc_mirror_cover(S, st));
  DRRT_CHECK(download(values, S->x_lmc.p + first, n * sizeof(double), st));
  DRRT_CUDA(cudaStreamSynchronize(st));
  return DRRT_OK;
}

// ---- pinned host memory for callers without a CUDA toolchain (the shim) -------

int drrt_host_alloc(size_t bytes, void **out) {
  if (!out) return set_error(DRRT_ERR_INVALID, "host_alloc: out == NULL");
  *out = nullptr;
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(DRRT_ERR_NOMEM, "host_alloc(%zu): %s", bytes, cudaGetErrorString(e));
  }
  return DRRT_OK;
}

int drrt_host_free(void *p) {
  if (p) DRRT_CUDA(cudaFreeHost(p));
  return DRRT_OK;
}

int drrt_dubins_trajectory_batch(drrt_store *h, int64_t n, const double *start, const double *end,
                                 double r_min, double *xy, int32_t *npts, char *type,
                                 double *length) {
  ENTER(h);
  if (n < 0 || (n > 0 && (!start || !end || (xy && !npts) || (!xy && !type && !length))))
    return set_error(DRRT_ERR_INVALID, "trajectory: arguments");
  if (n == 0) return DRRT_OK;
  cudaStream_t st = S->stream;
  const size_t per = xy ? (size_t)DRRT_TRAJ_MAX * 2 : 0;  // xy == NULL: word and length only
  DRRT_CHECK(S->d_stage.reserve(3 * n, st));
  DRRT_CHECK(S->d_stage2.reserve(3 * n, st));
  DRRT_CHECK(S->d_stage3.reserve(n * per + n, st));
  DRRT_CHECK(S->i_stage.reserve(n, st));
  DRRT_CHECK(S->b_stage.reserve(4 * n, st));
  DRRT_CHECK(upload(start, S->d_stage.p, 3 * n * sizeof(double), st));
  DRRT_CHECK(upload(end, S->d_stage2.p, 3 * n * sizeof(double), st));
  double *len_dev = S->d_stage3.p + n * per;
  DRRT_CHECK(trajectory_batch_dev(S, n, S->d_stage.p, S->d_stage2.p, r_min, xy ? S->d_stage3.p : nullptr,
                                  S->i_stage.p, (char *)S->b_stage.p, len_dev, st));
  if (xy) DRRT_CHECK(download(xy, S->d_stage3.p, n * per * sizeof(double), st));
  if (npts) DRRT_CHECK(download(npts, S->i_stage.p, n * sizeof(int32_t), st));
  if (type) DRRT_CHECK(download(type, S->b_stage.p, 4 * n, st));
  if (length) DRRT_CHECK(download(length, len_dev, n * sizeof(double), st));
  DRRT_CUDA(cudaStreamSynchronize(st));
  return DRRT_OK;
}

int drrt_pointcheck_batch(drrt_store *h, int64_t n, const double *pts, double robot_radius,
                          double collision_distance, uint8_t *verdict) {
  ENTER(h);
  if (n < 0 || (n > 0 && (!pts || !verdict))) return set_error(DRRT_ERR_INVALID, "pointcheck: arguments");
  if (n == 0) return DRRT_OK;
  cudaStream_t st = S->stream;
  if (n <= MAP_SMALL) {  // ExplicitNodeCheck of the one sampled node (mainloop.cpp:226)
    DRRT_CHECK(map_reserve(S));
    MappedSmall *M = mapped_small(S);
    memcpy(M->a, pts, 3 * n * sizeof(double));
    DRRT_CHECK(pointcheck_batch_dev(S, n, M->a, robot_radius, collision_distance, M->out_b, st));
    DRRT_CUDA(cudaStreamSynchronize(st));
    memcpy(verdict, M->out_b, n);
    return DRRT_OK;
  }
  DRRT_CHECK(S->d_stage.reserve(3 * n, st));
  DRRT_CHECK(S->b_stage.reserve(n, st));
  DRRT_CHECK(upload(pts, S->d_stage.p, 3 * n * sizeof(double), st));
  DRRT_CHECK(pointcheck_batch_dev(S, n, S->d_stage.p, robot_radius, collision_distance,
                                  S->b_stage.p, st));
  DRRT_CHECK(download(verdict, S->b_stage.p, n, st));
  DRRT_CUDA(cudaStreamSynchronize(st));
  return DRRT_OK;
}

int drrt_pointcheck_batch_dev(drrt_store *h, int64_t n, const double *pts_dev, double robot_radius,
                              double collision_distance, uint8_t *verdict_dev, void *stream) {
  ENTER(h);
  return pointcheck_batch_dev(S, n, pts_dev, robot_radius, collision_distance, verdict_dev,
                              pick(S, stream));
}

}  // extern "C"

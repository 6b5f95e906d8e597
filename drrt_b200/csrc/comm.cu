// comm.cu -- ONE host process driving N GPUs (drrt_comm_* in include/drrt_gpu.h).
//
// The reference planner is a single host process (src/smalltest.cpp:132-142: every thread of
// it shares one KDTree and one ConfigSpace), so the multi-GPU form of the hot path sits behind
// the same kind of handle: drrt_comm owns one replica of the node store per GPU, one NCCL
// communicator per GPU (ncclCommInitAll, all in this process) and one worker thread per GPU.
//
//   * inserts (and the rrt_LMC_ mirror): the batch crosses PCIe ONCE, to GPU 0, and reaches the
//     other replicas by ncclBroadcast over NVLink / NVSwitch, straight into the staging buffer
//     the insert kernels (K1) read -- N PCIe uploads through one host become one;
//   * query / edge batches: contiguous index ranges, one per GPU, no exchange step inside a
//     query or an edge check; every GPU copies its rows / verdicts STRAIGHT into the caller's
//     (pinned) buffers at its global offset, so nothing is gathered on a device and nothing is
//     re-copied on the host.  The one number a shard needs from the others -- where its rows
//     start in the caller's CSR -- is the sum of the preceding shards' hit totals; the host
//     that enqueues the copies has those totals from the kernels' own completion records, so
//     no device-side all-gather is issued for it (none is invented).
//   * results do not depend on the number of GPUs: ids are global insertion indices, rows are
//     sorted by id, shards are contiguous and concatenated in order.
//
// NCCL is loaded at run time (dlopen of libnccl.so.2, preferring a copy the process already
// holds, e.g. the one PyTorch loaded) so that the single-GPU library has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <functional>
#include <thread>
#include <vector>

#include "store.cuh"

namespace drrt {

struct NcclApi {
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  void *lib = nullptr;
  bool ok = false;
};

static NcclApi &nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      api.lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);  // a copy the process already holds
      if (api.lib) break;
    }
    for (const char *nm : names) {
      if (api.lib) break;
      api.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
    }
    if (!api.lib) return;
    api.CommInitAll = (decltype(api.CommInitAll))dlsym(api.lib, "ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.Broadcast = (decltype(api.Broadcast))dlsym(api.lib, "ncclBroadcast");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))dlsym(api.lib, "ncclGetVersion");
    api.ok = api.CommInitAll && api.CommDestroy && api.Broadcast && api.GetErrorString;
  });
  return api;
}

// one persistent host thread per GPU: keeps the device current, runs the jobs of a call
struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = false, quit = false;
  int rc = DRRT_OK;
  std::string err;

  void loop() {
    for (;;) {
      std::function<int()> j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return has_job || quit; });
        if (quit && !has_job) return;
        j = std::move(job);
        has_job = false;
      }
      int r = j();
      {
        std::lock_guard<std::mutex> lk(mu);
        rc = r;
        err = r == DRRT_OK ? std::string() : g_last_error;
        done = true;
      }
      cv.notify_all();
    }
  }
};

struct Comm {
  int n = 0;
  std::vector<int> devices;
  std::vector<Store *> rep;
  std::vector<ncclComm_t> nccl;
  std::vector<Worker *> workers;
  std::mutex mu;  // one comm call at a time (the reference serialises tree calls under tree_mutex_)
  std::atomic<int64_t> bcast_bytes{0};
};

// contiguous [lo, hi) of n items owned by rank r of w (sizes differ by at most one)
static inline void shard(int64_t n, int r, int w, int64_t *lo, int64_t *hi) {
  const int64_t base = n / w, rem = n % w;
  *lo = r * base + std::min<int64_t>(r, rem);
  *hi = *lo + base + (r < rem ? 1 : 0);
}

// f(rank) on every worker at once; the first failure (lowest rank) is the call's status
template <class F>
static int run_all(Comm *C, F f) {
  for (int r = 0; r < C->n; r++) {
    Worker *w = C->workers[r];
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->job = [f, r]() -> int { return f(r); };
      w->has_job = true;
      w->done = false;
    }
    w->cv.notify_all();
  }
  int rc = DRRT_OK;
  for (int r = 0; r < C->n; r++) {
    Worker *w = C->workers[r];
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return w->done; });
    if (w->rc != DRRT_OK && rc == DRRT_OK) {
      rc = w->rc;
      g_last_error = "gpu " + std::to_string(C->devices[r]) + ": " + w->err;
    }
  }
  return rc;
}

struct Locked {
  std::lock_guard<std::mutex> lock;
  DeviceGuard guard;
  explicit Locked(Store *S) : lock(S->mu), guard(S->device) {}
};
#define LOCKED(S)                                                                  \
  drrt::Locked locked__(S);                                                        \
  if (!locked__.guard.ok) return drrt::set_error(DRRT_ERR_CUDA, "cudaSetDevice failed")

#define DRRT_NCCL(call)                                                                    \
  do {                                                                                     \
    ncclResult_t r__ = (call);                                                             \
    if (r__ != ncclSuccess)                                                                \
      return drrt::set_error(DRRT_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,     \
                             drrt::nccl_api().GetErrorString(r__));                        \
  } while (0)

// `count` doubles from the host into dst(rank) of every replica: one H2D (to rank 0), then
// ncclBroadcast over NVLink; runs on the calling worker thread of rank r, on the store's stream.
static int bcast_doubles(Comm *C, int r, const double *host, size_t count, double *dst) {
  Store *S = C->rep[r];
  cudaStream_t st = S->stream;
  if (count == 0) return DRRT_OK;
  if (r == 0) DRRT_CUDA(cudaMemcpyAsync(dst, host, count * sizeof(double), cudaMemcpyHostToDevice, st));
  if (C->n > 1) {
    DRRT_NCCL(nccl_api().Broadcast(dst, dst, count, ncclDouble, 0, C->nccl[r], st));
    if (r == 0) C->bcast_bytes += (int64_t)(count * sizeof(double)) * (C->n - 1);
  }
  return DRRT_OK;
}

static int bcast_i32(Comm *C, int r, const int32_t *host, size_t count, int32_t *dst) {
  Store *S = C->rep[r];
  cudaStream_t st = S->stream;
  if (count == 0) return DRRT_OK;
  if (r == 0) DRRT_CUDA(cudaMemcpyAsync(dst, host, count * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if (C->n > 1) {
    DRRT_NCCL(nccl_api().Broadcast(dst, dst, count, ncclInt32, 0, C->nccl[r], st));
    if (r == 0) C->bcast_bytes += (int64_t)(count * sizeof(int32_t)) * (C->n - 1);
  }
  return DRRT_OK;
}

__global__ void comm_scatter_f64_kernel(double *dst, int64_t n, const int32_t *ids, const double *v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[ids[i]] = v[i];
}

}  // namespace drrt

using namespace drrt;

#define ENTER_COMM(c)                                                     \
  if (!(c)) return drrt::set_error(DRRT_ERR_INVALID, "null comm handle");  \
  drrt::Comm *C = reinterpret_cast<drrt::Comm *>(c);                      \
  std::lock_guard<std::mutex> comm_lock__(C->mu)

extern "C" {

int drrt_comm_destroy(drrt_comm *c) {
  if (!c) return DRRT_OK;
  Comm *C = reinterpret_cast<Comm *>(c);
  for (Worker *w : C->workers) {
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->quit = true;
    }
    w->cv.notify_all();
    if (w->th.joinable()) w->th.join();
    delete w;
  }
  for (size_t r = 0; r < C->nccl.size(); r++)
    if (C->nccl[r]) nccl_api().CommDestroy(C->nccl[r]);
  for (Store *S : C->rep) drrt_store_destroy(reinterpret_cast<drrt_store *>(S));
  delete C;
  return DRRT_OK;
}

int drrt_comm_init(int n_gpus, const int *devices, int dims, int n_wraps, const int *wrap_dims,
                   const double *wrap_points, int metric, int64_t capacity_hint, drrt_comm **out) {
  if (!out) return set_error(DRRT_ERR_INVALID, "comm_init: out == NULL");
  *out = nullptr;
  if (n_gpus < 1 || n_gpus > 64) return set_error(DRRT_ERR_INVALID, "comm_init: n_gpus = %d", n_gpus);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_error(DRRT_ERR_NO_DEVICE, "comm_init: no CUDA device (this library has no CPU path)");
  }
  Comm *C = new (std::nothrow) Comm();
  if (!C) return set_error(DRRT_ERR_NOMEM, "comm_init: host allocation");
  C->n = n_gpus;
  for (int r = 0; r < n_gpus; r++) {
    const int d = devices ? devices[r] : r;
    if (d < 0 || d >= ndev) {
      delete C;
      return set_error(DRRT_ERR_NO_DEVICE, "comm_init: device %d of %d visible", d, ndev);
    }
    for (int e : C->devices)
      if (e == d) {
        delete C;
        return set_error(DRRT_ERR_INVALID, "comm_init: device %d listed twice", d);
      }
    C->devices.push_back(d);
  }
  int rc = DRRT_OK;
  for (int r = 0; r < n_gpus && rc == DRRT_OK; r++) {
    drrt_store *h = nullptr;
    rc = drrt_store_create(C->devices[r], dims, n_wraps, wrap_dims, wrap_points, metric, capacity_hint, &h);
    if (rc == DRRT_OK) C->rep.push_back(reinterpret_cast<Store *>(h));
  }
  if (rc == DRRT_OK && n_gpus > 1) {
    NcclApi &api = nccl_api();
    if (!api.ok) {
      rc = set_error(DRRT_ERR_UNSUPPORTED, "comm_init: libnccl.so.2 cannot be loaded (%s)",
                     api.lib ? "symbols missing" : "not found");
    } else {
      C->nccl.assign(n_gpus, nullptr);
      ncclResult_t r = api.CommInitAll(C->nccl.data(), n_gpus, C->devices.data());
      if (r != ncclSuccess) {
        C->nccl.clear();
        rc = set_error(DRRT_ERR_CUDA, "comm_init: ncclCommInitAll: %s", api.GetErrorString(r));
      }
    }
  }
  if (rc != DRRT_OK) {
    std::string keep = g_last_error;
    drrt_comm_destroy(reinterpret_cast<drrt_comm *>(C));
    g_last_error = keep;
    return rc;
  }
  for (int r = 0; r < n_gpus; r++) {
    Worker *w = new Worker();
    C->workers.push_back(w);
    w->th = std::thread([w] { w->loop(); });
  }
  *out = reinterpret_cast<drrt_comm *>(C);
  return DRRT_OK;
}

int drrt_comm_size(drrt_comm *c, int *n_gpus) {
  ENTER_COMM(c);
  if (!n_gpus) return set_error(DRRT_ERR_INVALID, "comm_size: out == NULL");
  *n_gpus = C->n;
  return DRRT_OK;
}

drrt_store *drrt_comm_store(drrt_comm *c, int rank) {
  if (!c) return nullptr;
  Comm *C = reinterpret_cast<Comm *>(c);
  return rank >= 0 && rank < C->n ? reinterpret_cast<drrt_store *>(C->rep[rank]) : nullptr;
}

int64_t drrt_comm_broadcast_bytes(drrt_comm *c) {
  return c ? reinterpret_cast<Comm *>(c)->bcast_bytes.load() : 0;
}

// KDTree::KDInsert on every replica: pos crosses PCIe once, ncclBroadcast fills the buffer K1 reads
int drrt_comm_insert(drrt_comm *c, int64_t n, const double *pos, int32_t *first_id) {
  ENTER_COMM(c);
  if (n < 0 || (n > 0 && !pos)) return set_error(DRRT_ERR_INVALID, "comm_insert: arguments");
  for (int64_t i = 0; i < 3 * n; i++)
    if (!(pos[i] - pos[i] == 0.0))
      return set_error(DRRT_ERR_UNSUPPORTED, "comm_insert: non-finite coordinate at %lld", (long long)i);
  std::vector<int32_t> first(C->n, -1);
  int rc = run_all(C, [&](int r) -> int {
    Store *S = C->rep[r];
    LOCKED(S);
    cudaStream_t st = S->stream;
    DRRT_CHECK(S->stage_pos.reserve(3 * n + 1, st));
    DRRT_CHECK(bcast_doubles(C, r, pos, 3 * (size_t)n, S->stage_pos.p));
    DRRT_CHECK(store_insert_dev(S, n, S->stage_pos.p, st, &first[r]));
    DRRT_CUDA(cudaStreamSynchronize(st));
    return DRRT_OK;
  });
  DRRT_CHECK(rc);
  for (int r = 1; r < C->n; r++)
    if (first[r] != first[0]) return set_error(DRRT_ERR_STATE, "comm_insert: replicas disagree on the next id");
  if (first_id) *first_id = first[0];
  return DRRT_OK;
}

// the obstacle table is a few tens of KB and every replica keeps host-side copies of it for its
// broadphase grid: each replica takes it from the host directly
int drrt_comm_obstacles_set(drrt_comm *c, int32_t n_obs, const int32_t *kind, const uint8_t *used,
                            const double *life_span, const double *origin_xy, const double *radius,
                            const int32_t *vert_off, const double *verts_xy) {
  ENTER_COMM(c);
  return run_all(C, [&](int r) -> int {
    return drrt_obstacles_set(reinterpret_cast<drrt_store *>(C->rep[r]), n_obs, kind, used, life_span,
                              origin_xy, radius, vert_off, verts_xy);
  });
}

static int comm_range_into(Comm *C, int64_t nq, const double *q, const double *radius, double radius_all,
                           int64_t cap, int64_t *offsets, int32_t *ids, void *dist, int dist_kind,
                           int64_t *total) {
  if (nq < 0 || cap < 0 || (nq > 0 && (!q || !offsets)) || (cap > 0 && !ids))
    return set_error(DRRT_ERR_INVALID, "comm_range_into: arguments");
  if (!dist) dist_kind = DIST_NONE;
  if (C->n == 1) {  // one GPU: the streamed call itself (rows overlap the next sub-batch)
    Store *S = C->rep[0];
    LOCKED(S);
    return range_into(S, nq, q, radius, radius_all, cap, offsets, ids, dist, dist_kind, total);
  }
  const size_t dsz = dist_kind == DIST_F64 ? sizeof(double) : (dist_kind == DIST_F32 ? sizeof(float) : 0);
  std::vector<int64_t> tot(C->n, 0), base(C->n + 1, 0);
  // phase 1: every GPU answers its shard; rows stay in its arena
  int rc = run_all(C, [&](int r) -> int {
    Store *S = C->rep[r];
    LOCKED(S);
    int64_t lo, hi;
    shard(nq, r, C->n, &lo, &hi);
    const int64_t m = hi - lo;
    cudaStream_t st = S->stream;
    DRRT_CHECK(S->q_stage.reserve(3 * m + 1, st));
    if (m > 0) DRRT_CUDA(cudaMemcpyAsync(S->q_stage.p, q + 3 * lo, 3 * m * sizeof(double), cudaMemcpyHostToDevice, st));
    if (radius) {
      DRRT_CHECK(S->rad_stage.reserve(m + 1, st));
      if (m > 0) DRRT_CUDA(cudaMemcpyAsync(S->rad_stage.p, radius + lo, m * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    drrt_range_result res{};
    DRRT_CHECK(range_batch_dev(S, m, S->q_stage.p, radius ? S->rad_stage.p : nullptr, radius_all, st, &res));
    tot[r] = res.total;
    return DRRT_OK;
  });
  DRRT_CHECK(rc);
  for (int r = 0; r < C->n; r++) base[r + 1] = base[r] + tot[r];
  const int64_t all = base[C->n];
  const bool fits = all <= cap;
  // phase 2: every GPU packs its rows (offsets already global) and copies them straight into the
  // caller's buffers at its base
  rc = run_all(C, [&](int r) -> int {
    Store *S = C->rep[r];
    LOCKED(S);
    int64_t lo, hi;
    shard(nq, r, C->n, &lo, &hi);
    const int64_t m = hi - lo;
    cudaStream_t st = S->stream;
    int rc2 = DRRT_OK;
    if (m > 0) {
      const size_t need = (size_t)tot[r] + 1;
      rc2 = S->p_off[0].reserve(m + 1, st);
      if (rc2 == DRRT_OK) rc2 = S->p_ids[0].reserve(need, st);
      if (rc2 == DRRT_OK && dsz) rc2 = S->p_dist[0].reserve(need, st);
      if (rc2 == DRRT_OK)
        rc2 = range_pack_to(S, st, S->p_off[0].p, S->p_ids[0].p, S->p_dist[0].p, dist_kind, base[r]);
      // rank r owns offsets[lo .. hi); the last rank also writes offsets[nq]
      const int64_t noff = m + (r == C->n - 1 ? 1 : 0);
      if (rc2 == DRRT_OK)
        rc2 = cudaMemcpyAsync(offsets + lo, S->p_off[0].p, noff * sizeof(int64_t), cudaMemcpyDeviceToHost, st) == cudaSuccess
                  ? DRRT_OK : set_error(DRRT_ERR_CUDA, "comm_range_into: offsets copy");
      if (rc2 == DRRT_OK && fits && tot[r] > 0) {
        cudaError_t e = cudaMemcpyAsync(ids + base[r], S->p_ids[0].p, tot[r] * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && dsz)
          e = cudaMemcpyAsync((char *)dist + base[r] * dsz, S->p_dist[0].p, tot[r] * dsz, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc2 = set_error(DRRT_ERR_CUDA, "comm_range_into: row copy: %s", cudaGetErrorString(e));
      }
    }
    cudaError_t e = cudaStreamSynchronize(st);
    S->last_nq = -1;
    DRRT_CHECK(rc2);
    DRRT_CUDA(e);
    return DRRT_OK;
  });
  DRRT_CHECK(rc);
  if (nq == 0 && offsets) offsets[0] = 0;
  // empty trailing shards (nq < n_gpus) leave offsets[nq] to the last non-empty one
  if (nq > 0) offsets[nq] = all;
  if (total) *total = all;
  if (!fits)
    return set_error(DRRT_ERR_CAPACITY, "comm_range_into: %lld hits, caller buffers hold %lld",
                     (long long)all, (long long)cap);
  return DRRT_OK;
}

int drrt_comm_range_batch_into(drrt_comm *c, int64_t nq, const double *q, const double *radius,
                               double radius_all, int64_t cap, int64_t *offsets, int32_t *ids,
                               double *dist, int64_t *total) {
  ENTER_COMM(c);
  return comm_range_into(C, nq, q, radius, radius_all, cap, offsets, ids, dist, DIST_F64, total);
}

int drrt_comm_range_batch_into_f32(drrt_comm *c, int64_t nq, const double *q, const double *radius,
                                   double radius_all, int64_t cap, int64_t *offsets, int32_t *ids,
                                   float *dist, int64_t *total) {
  ENTER_COMM(c);
  return comm_range_into(C, nq, q, radius, radius_all, cap, offsets, ids, dist, DIST_F32, total);
}

int drrt_comm_nearest_batch(drrt_comm *c, int64_t nq, const double *q, int32_t *ids, double *dist) {
  ENTER_COMM(c);
  if (nq < 0 || (nq > 0 && (!q || !ids))) return set_error(DRRT_ERR_INVALID, "comm_nearest: arguments");
  return run_all(C, [&](int r) -> int {
    Store *S = C->rep[r];
    LOCKED(S);
    int64_t lo, hi;
    shard(nq, r, C->n, &lo, &hi);
    if (hi == lo) return DRRT_OK;
    return knn_host(S, hi - lo, q + 3 * lo, 1, ids + lo, dist ? dist + lo : nullptr, true);
  });
}

int drrt_comm_knn_batch(drrt_comm *c, int64_t nq, const double *q, int k, int32_t *ids, double *dist) {
  ENTER_COMM(c);
  if (nq < 0 || k < 1 || (nq > 0 && (!q || !ids))) return set_error(DRRT_ERR_INVALID, "comm_knn: arguments");
  return run_all(C, [&](int r) -> int {
    Store *S = C->rep[r];
    LOCKED(S);
    int64_t lo, hi;
    shard(nq, r, C->n, &lo, &hi);
    if (hi == lo) return DRRT_OK;
    return knn_host(S, hi - lo, q + 3 * lo, k, ids + lo * k, dist ? dist + lo * k : nullptr, false);
  });
}

static int comm_edges(Comm *C, int64_t n, const double *start, const double *end, const int32_t *sid,
                      const int32_t *eid, int edge_kind, double robot_radius, double r_min, int in_warmup,
                      uint8_t *verdict, double *length) {
  if (n < 0 || (n > 0 && !verdict)) return set_error(DRRT_ERR_INVALID, "comm_edgecheck: arguments");
  const bool by_id = sid && eid;
  if (n > 0 && !by_id && !(start && end)) return set_error(DRRT_ERR_INVALID, "comm_edgecheck: endpoints");
  return run_all(C, [&](int r) -> int {
    Store *S = C->rep[r];
    LOCKED(S);
    int64_t lo, hi;
    shard(n, r, C->n, &lo, &hi);
    if (hi == lo) return DRRT_OK;
    return edge_host(S, hi - lo, by_id ? nullptr : start + 3 * lo, by_id ? nullptr : end + 3 * lo,
                     by_id ? sid + lo : nullptr, by_id ? eid + lo : nullptr, edge_kind, robot_radius, r_min,
                     in_warmup, verdict + lo, length ? length + lo : nullptr);
  });
}

int drrt_comm_edgecheck_batch(drrt_comm *c, int64_t n, const double *start, const double *end,
                              int edge_kind, double robot_radius, double r_min, int in_warmup,
                              uint8_t *verdict, double *length) {
  ENTER_COMM(c);
  return comm_edges(C, n, start, end, nullptr, nullptr, edge_kind, robot_radius, r_min, in_warmup, verdict, length);
}

int drrt_comm_edgecheck_ids_batch(drrt_comm *c, int64_t n, const int32_t *start_id, const int32_t *end_id,
                                  int edge_kind, double robot_radius, double r_min, int in_warmup,
                                  uint8_t *verdict, double *length) {
  ENTER_COMM(c);
  if (n > 0 && (!start_id || !end_id)) return set_error(DRRT_ERR_INVALID, "comm_edgecheck: ids");
  return comm_edges(C, n, nullptr, nullptr, start_id, end_id, edge_kind, robot_radius, r_min, in_warmup, verdict, length);
}

int drrt_comm_pointcheck_batch(drrt_comm *c, int64_t n, const double *pts, double robot_radius,
                               double collision_distance, uint8_t *verdict) {
  ENTER_COMM(c);
  if (n < 0 || (n > 0 && (!pts || !verdict))) return set_error(DRRT_ERR_INVALID, "comm_pointcheck: arguments");
  return run_all(C, [&](int r) -> int {
    int64_t lo, hi;
    shard(n, r, C->n, &lo, &hi);
    if (hi == lo) return DRRT_OK;
    return drrt_pointcheck_batch(reinterpret_cast<drrt_store *>(C->rep[r]), hi - lo, pts + 3 * lo, robot_radius,
                                 collision_distance, verdict + lo);
  });
}

// rrt_LMC_ mirror on every replica: one H2D + ncclBroadcast
int drrt_comm_lmc_write(drrt_comm *c, int64_t first, int64_t n, const double *values) {
  ENTER_COMM(c);
  if (first < 0 || n < 0 || first + n > C->rep[0]->n || (n > 0 && !values))
    return set_error(DRRT_ERR_INVALID, "comm_lmc_write: range");
  return run_all(C, [&](int r) -> int {
    Store *S = C->rep[r];
    LOCKED(S);
    cudaStream_t st = S->stream;
    DRRT_CHECK(lmc_mirror_cover(S, st));
    DRRT_CHECK(bcast_doubles(C, r, values, (size_t)n, S->x_lmc.p + first));
    DRRT_CUDA(cudaStreamSynchronize(st));
    return DRRT_OK;
  });
}

int drrt_comm_lmc_scatter(drrt_comm *c, int64_t n, const int32_t *ids, const double *values) {
  ENTER_COMM(c);
  if (n < 0 || (n > 0 && (!ids || !values))) return set_error(DRRT_ERR_INVALID, "comm_lmc_scatter: arguments");
  for (int64_t i = 0; i < n; i++)
    if (ids[i] < 0 || ids[i] >= C->rep[0]->n)
      return set_error(DRRT_ERR_INVALID, "comm_lmc_scatter: node id %d at %lld", ids[i], (long long)i);
  if (n == 0) return DRRT_OK;
  return run_all(C, [&](int r) -> int {
    Store *S = C->rep[r];
    LOCKED(S);
    cudaStream_t st = S->stream;
    DRRT_CHECK(lmc_mirror_cover(S, st));
    DRRT_CHECK(S->i_stage.reserve(n, st));
    DRRT_CHECK(S->d_stage.reserve(n, st));
    DRRT_CHECK(bcast_i32(C, r, ids, (size_t)n, S->i_stage.p));
    DRRT_CHECK(bcast_doubles(C, r, values, (size_t)n, S->d_stage.p));
    comm_scatter_f64_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(
        S->x_lmc.p, n, S->i_stage.p, S->d_stage.p);
    DRRT_LAUNCHED();
    DRRT_CUDA(cudaGetLastError());
    DRRT_CUDA(cudaStreamSynchronize(st));
    return DRRT_OK;
  });
}

// the fused Extend step over sharded new nodes; lmc == NULL uses the replicas' resident mirror
int drrt_comm_extend_batch(drrt_comm *c, int64_t nq, const double *q, const double *radius,
                           double radius_all, double robot_radius, double r_min, int in_warmup,
                           const double *lmc, int32_t *counts, int64_t *total, int32_t *best_parent,
                           double *best_cost) {
  ENTER_COMM(c);
  if (nq < 0 || (nq > 0 && (!q || !counts))) return set_error(DRRT_ERR_INVALID, "comm_extend: arguments");
  if (lmc && best_parent) {
    const int64_t n = C->rep[0]->n;
    int rc = run_all(C, [&](int r) -> int {
      Store *S = C->rep[r];
      LOCKED(S);
      cudaStream_t st = S->stream;
      DRRT_CHECK(lmc_mirror_cover(S, st));
      DRRT_CHECK(bcast_doubles(C, r, lmc, (size_t)n, S->x_lmc.p));
      DRRT_CUDA(cudaStreamSynchronize(st));
      return DRRT_OK;
    });
    DRRT_CHECK(rc);
  }
  std::vector<int64_t> tot(C->n, 0);
  int rc = run_all(C, [&](int r) -> int {
    int64_t lo, hi;
    shard(nq, r, C->n, &lo, &hi);
    if (hi == lo) return DRRT_OK;
    return drrt_extend_batch(reinterpret_cast<drrt_store *>(C->rep[r]), hi - lo, q + 3 * lo,
                             radius ? radius + lo : nullptr, radius_all, robot_radius, r_min, in_warmup,
                             nullptr, counts + lo, &tot[r], best_parent ? best_parent + lo : nullptr,
                             best_cost ? best_cost + lo : nullptr);
  });
  DRRT_CHECK(rc);
  int64_t all = 0;
  for (int64_t t : tot) all += t;
  if (total) *total = all;
  return DRRT_OK;
}

}  // extern "C"

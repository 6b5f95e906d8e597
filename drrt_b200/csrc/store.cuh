// store.cuh -- the GPU-resident node store behind drrt_store.
//
// Device twin of KDTree + KDTreeNode (include/DRRT/kdtree.h:13-181,
// include/DRRT/kdtreenode.h:14-154): only position_ and the KD links go to the
// device; the RRTx graph state stays with the host planner.
//
// HBM layout (N nodes, id = insertion order):
//   x[N] y[N] th[N]            f64 SoA, 24 B/node       (id order)
//   parent[N] left[N] right[N] i32, split[N] u8         insertion-ordered KD tree
//   gate[N]                    f64: smallest theta among the theta-split
//                              ancestors that hold the node in their LEFT
//                              subtree (and the node itself once it has a
//                              right child): decides whether the reference's
//                              signed-hyperplane walk ever evaluates the node
//   rec[N_indexed]             32-B NodeRec sorted by grid cell (the scan target)
//   cell_start[ncell+1]        i32
// Nodes inserted after the last index build form a short unsorted tail that
// queries scan directly; the index is rebuilt when the tail is long enough to
// matter (ensure_index).
#pragma once

#include <vector>

#include "common.cuh"

namespace drrt {

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;  // elements
  int reserve(size_t n, cudaStream_t s, bool keep = false, size_t keep_n = 0);
  void release();
};

struct Store {
  int device = 0;
  int metric = 0;
  int n_wraps = 0;
  double wrap_point = 0.0;
  double period = 0.0;  // theta key period (0: none)
  cudaStream_t stream = nullptr;
  std::mutex mu;

  int64_t n = 0;
  DevBuf<double> x, y, th, gate;
  DevBuf<int32_t> parent, left, right;
  DevBuf<uint8_t> split;

  // insert scratch
  DevBuf<int32_t> ins_cur;
  DevBuf<uint8_t> ins_split;
  DevBuf<double> ins_gate;
  DevBuf<double> stage_pos;  // n x 3 upload staging (device)
  int32_t *d_pending = nullptr;
  int32_t *h_pending = nullptr;  // pinned

  // grid index
  int64_t n_indexed = 0;
  GridParams gp{};
  int64_t ncell = 0;
  DevBuf<NodeRec> rec;
  DevBuf<int32_t> cell_start, cell_of, rank_in_cell, scan_tmp;
  double *d_bounds = nullptr;  // 6 doubles: minx maxx miny maxy mint maxt (ordered-int encoded)
  double *h_bounds = nullptr;  // pinned

  // range results
  DevBuf<int64_t> r_offsets;  // nq+1
  DevBuf<int32_t> r_ids;
  DevBuf<double> r_dist;
  DevBuf<int32_t> r_counts;           // nq
  DevBuf<int64_t> r_chunk_total, r_chunk_base;  // CTA kernel: per-slice hit totals / exact slice starts
  // plain-CSR copy of a sliced result (range_pack_dev)
  DevBuf<int64_t> r_offsets2;
  DevBuf<int32_t> r_ids2;
  DevBuf<double> r_dist2;
  DevBuf<double> q_stage, rad_stage;
  // streamed host query (drrt_range_batch_into): the packed rows of a sub-batch, double
  // buffered, leave on copy_stream while the next sub-batch is computed on stream
  DevBuf<int64_t> p_off[2];
  DevBuf<int32_t> p_ids[2];
  DevBuf<double> p_dist[2];
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_packed[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  // pipelined host edge checks: poses up on copy_stream, kernels on stream, verdicts down on down_stream
  cudaStream_t down_stream = nullptr;
  cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  // ranked Dubins solve: the small single-word bins run beside the two large ones
  cudaStream_t side_stream[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  int32_t *d_ticket = nullptr;
  unsigned long long *d_totals = nullptr;  // inside the d_ticket allocation
  int range_grid[2] = {0, 0};  // resident grid of range_warp_kernel / range_cta_kernel on THIS device
  int64_t last_nq = -1, last_total = 0, last_extent = 0;
  bool last_packed = true;          // the primary arena is a plain CSR
  bool last_packed_in_alt = false;  // a packed copy exists in r_*2
  int64_t *h_scalar = nullptr;  // pinned scratch (8 x i64)
  void *h_map = nullptr;        // MappedRow + MappedSmall (pinned, mapped; map_reserve)

  // generic staging for host entry points
  DevBuf<int32_t> i_stage, i_stage2;
  DevBuf<double> d_stage, d_stage2, d_stage3;
  DevBuf<uint8_t> b_stage;

  // obstacles
  ObstacleTable obs{};
  DevBuf<int32_t> o_kind, o_voff, oa_kind, oa_voff;
  DevBuf<uint8_t> oa_live;
  DevBuf<double> o_ox, o_oy, o_rad, o_vx, o_vy, oa_ox, oa_oy, oa_rad, oa_vx, oa_vy;
  bool has_obstacles = false;
  // broadphase over the active obstacles: uniform grid, cell -> obstacles whose
  // reach circle (radius_ + robot radius) touches the cell; rebuilt when the
  // table or the robot radius changes
  std::vector<double> h_o_ox, h_o_oy, h_o_rad, h_o_vx, h_o_vy;
  std::vector<int32_t> h_o_kind, h_o_voff;
  std::vector<int32_t> h_o_src;  // table index of each active obstacle
  // the whole table (obstacle sweeps: conflict queries, subset checks)
  std::vector<double> h_all_ox, h_all_oy, h_all_rad;
  std::vector<int32_t> h_all_kind;
  // host copy of the caller's table as given to drrt_obstacles_set, patched by drrt_obstacles_move /
  // drrt_obstacles_update (only the changed entries cross the C boundary)
  std::vector<int32_t> t_kind, t_voff;
  std::vector<uint8_t> t_used;
  std::vector<double> t_life, t_origin, t_radius, t_verts;
  DevBuf<uint8_t> o_mask;                // per ACTIVE obstacle: taking part in a subset check
  DevBuf<uint8_t> o_covers;              // per ACTIVE obstacle: radius_ covers every polygon vertex
  const uint8_t *edge_mask = nullptr;    // set for the duration of a subset check
  ObstacleGrid ogrid{};
  double ogrid_radius = -1.0;
  DevBuf<int32_t> og_off, og_idx, og_eobs, og_ball;
  DevBuf<EdgeRec> og_erec;
  // fused Extend: per-neighbour verdicts / lengths (2 per pair), query poses of the CSR rows
  DevBuf<uint8_t> x_unsafe;
  DevBuf<double> x_len, x_lmc;
  DevBuf<int32_t> x_best, x_row;
  DevBuf<double> x_best_cost;
  const double *csr_q = nullptr;
  const int64_t *csr_off = nullptr;
  const int32_t *csr_ids = nullptr;
  const double *csr_dist = nullptr;
  int64_t csr_nq = 0;
  int64_t last_extend_total = -1;
  int64_t lmc_n = 0;  // entries of x_lmc that are current (the rest of the store reads INF once covered)
  // stored neighbour edges of the planner (drrt_edges_append) and the scratch of the obstacle sweeps
  DevBuf<int32_t> e_u, e_v;
  int64_t n_stored_edges = 0;
  DevBuf<uint8_t> node_mark, sw_v1, sw_v2, sw_out_flag;
  DevBuf<int32_t> sw_cnt, sw_sel, sw_u, sw_v, sw_pick, scan_tmp2;
  DevBuf<int64_t> sw_out_edge;
  int64_t sweep_out = -1;
  // time-parametrised obstacles (kinds 6 / 7, drrt_timed_obstacles_set): the live ones
  DevBuf<double> tm_ox, tm_oy, tm_rad, tm_path;
  DevBuf<int32_t> tm_poff;
  int timed_n = 0;
  // rewiring candidates of the last fused Extend (drrt_extend_rewire)
  DevBuf<double> rw_lmc;
  int64_t rewire_out = -1;
  const double *ext_lmc = nullptr, *ext_cost = nullptr;  // the lmc array / best_cost of that Extend (device)
  const int32_t *ext_best = nullptr;
  // solved Dubins paths handed from the solve kernel to the walk kernel
  DevBuf<unsigned char> edge_paths;
  // ranked solve: candidate list per edge, edges permuted by list, bin counters
  DevBuf<uint8_t> edge_masks;
  DevBuf<uint32_t> edge_perm;
  DevBuf<unsigned long long> edge_bins;
  unsigned long long *d_edge_counter = nullptr;
};

// Every extern "C" entry point on a handle: serialise on the handle's mutex and make its device
// current for the duration of the call (restored on return).
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

#define ENTER(h)                                                         \
  if (!(h)) return drrt::set_error(DRRT_ERR_INVALID, "null handle");      \
  drrt::Store *S = reinterpret_cast<drrt::Store *>(h);                    \
  std::lock_guard<std::mutex> lock__(S->mu);                              \
  drrt::DeviceGuard guard__(S->device);                                   \
  if (!guard__.ok) return drrt::set_error(DRRT_ERR_CUDA, "cudaSetDevice failed")

// store.cu
int store_reserve_nodes(Store *S, int64_t need);
int store_insert_dev(Store *S, int64_t n, const double *pos_dev, cudaStream_t st,
                     int32_t *first_id);
int ensure_index(Store *S, int64_t nq, cudaStream_t st);
int rebuild_index(Store *S, cudaStream_t st);

// scan.cu
int exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, DevBuf<int32_t> &tmp,
                       cudaStream_t st);

// range.cu
int range_batch_dev(Store *S, int64_t nq, const double *q_dev, const double *radius_dev,
                    double radius_all, cudaStream_t st, drrt_range_result *out);
int range_pack_dev(Store *S, cudaStream_t st, drrt_range_result *out);
// rows of the last result written back to back into caller-chosen device buffers
// (off_out: nq + 1 entries, base + exclusive prefix sums of the counts); dist_out holds double,
// float or nothing (dist_kind); the handle's result state is untouched
enum { DIST_F64 = 0, DIST_F32 = 1, DIST_NONE = 2 };
int range_pack_to(Store *S, cudaStream_t st, int64_t *off_out, int32_t *ids_out, void *dist_out,
                  int dist_kind, int64_t base);
// Page-locked host memory mapped into the device's address space (cudaHostAllocMapped): the
// small-call fast paths read their inputs from it and write their results into it, so that the
// planner's one-query / one-edge calls cost a kernel launch and a stream wait, no copy calls.
static constexpr int64_t MAP_ROW_ELEMS = 8192;   // hits of the mapped single-query row
static constexpr int64_t MAP_SMALL = 1024;       // items of the other mapped small calls
struct MappedRow {
  double q[4];
  unsigned long long totals[2];
  int64_t offsets[2];
  int64_t chunk_total[2];
  int32_t counts[4];
  int32_t ids[MAP_ROW_ELEMS];
  double dist[MAP_ROW_ELEMS];
};
struct MappedSmall {           // after the MappedRow in the same allocation
  double a[3 * MAP_SMALL];     // queries / start poses / points
  double b[3 * MAP_SMALL];     // end poses
  double out_d[MAP_SMALL];     // distances / lengths
  int32_t ia[MAP_SMALL], ib[MAP_SMALL];  // edge end ids
  int32_t out_i[MAP_SMALL];    // ids
  uint8_t out_b[MAP_SMALL];    // verdicts
  int32_t flag;                // set by validation kernels (e.g. an edge end id outside the store)
};
static constexpr int RANGE_NOT_HANDLED = -1;
int map_reserve(Store *S);
int range_one_mapped(Store *S, const double *q3, double radius, int64_t *total, const int32_t **ids,
                     const double **dist);

// capi.cu: host-buffer legs; the caller holds S->mu and has the store's device current
int range_into(Store *S, int64_t nq, const double *q, const double *radius, double radius_all,
               int64_t cap, int64_t *offsets, int32_t *ids, void *dist, int dist_kind, int64_t *total);
int edge_host(Store *S, int64_t n, const double *start, const double *end, const int32_t *sid,
              const int32_t *eid, int edge_kind, double robot_radius, double r_min, int in_warmup,
              uint8_t *verdict, double *length);
int knn_host(Store *S, int64_t nq, const double *q, int k, int32_t *ids, double *dist, bool nearest);
int lmc_mirror_cover(Store *S, cudaStream_t st);
int obstacles_apply(Store *S);
// knn.cu
int knn_batch_dev(Store *S, int64_t nq, const double *q_dev, int k, int32_t *ids_dev,
                  double *dist_dev, cudaStream_t st, bool nearest = false);
// edgecheck.cu
int edgecheck_batch_dev(Store *S, int64_t n, const double *start_dev, const double *end_dev,
                        const int32_t *sid_dev, const int32_t *eid_dev, int edge_kind,
                        double robot_radius, double r_min, uint8_t *verdict_dev,
                        double *length_dev, cudaStream_t st);
int extend_batch_dev(Store *S, int64_t nq, const double *q_dev, const double *radius_dev,
                     double radius_all, double robot_radius, double r_min, const double *lmc_dev,
                     cudaStream_t st, drrt_range_result *rr, uint8_t **unsafe_dev, double **len_dev,
                     int32_t *best_dev, double *best_cost_dev);
int trajectory_batch_dev(Store *S, int64_t n, const double *start_dev, const double *end_dev,
                         double r_min, double *xy_dev, int32_t *npts_dev, char *type_dev,
                         double *length_dev, cudaStream_t st);
int pointcheck_batch_dev(Store *S, int64_t n, const double *pts_dev, double robot_radius,
                         double collision_distance, uint8_t *verdict_dev, cudaStream_t st);

template <typename T>
int DevBuf<T>::reserve(size_t n, cudaStream_t s, bool keep, size_t keep_n) {
  if (n <= cap) return DRRT_OK;
  size_t ncap = cap ? cap : 256;
  while (ncap < n) ncap = 2 * ncap + 256;
  T *np = nullptr;
  cudaError_t e = cudaMalloc(&np, ncap * sizeof(T));
  if (e != cudaSuccess)
    return set_error(DRRT_ERR_NOMEM, "cudaMalloc(%zu B): %s", ncap * sizeof(T),
                     cudaGetErrorString(e));
  if (p) {
    if (keep && keep_n) {
      e = cudaMemcpyAsync(np, p, keep_n * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) return set_error(DRRT_ERR_CUDA, "grow copy: %s", cudaGetErrorString(e));
    }
    // the old block may still be read by work queued on s
    cudaStreamSynchronize(s);
    cudaFree(p);
  }
  p = np;
  cap = ncap;
  return DRRT_OK;
}

template <typename T>
void DevBuf<T>::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

}  // namespace drrt

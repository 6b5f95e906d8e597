// sweep.cu -- obstacle add / remove sweeps over a DEVICE-RESIDENT list of the planner's stored
// neighbour edges.
//
// AddObstacle (src/obstacle.cpp:612-667) and RemoveObstacle (:669-754) do, for one obstacle O:
//   1. FindPointsInConflictWithObstacle (:540-610): one range query around O;
//   2. for every node found, every stored out-edge (RrtNodeNeighborIterator over the node's
//      neighbour lists) and the parent edge: edge->ExplicitEdgeCheck(O);
//   3. RemoveObstacle only: an edge that collides with O is re-checked against every OTHER
//      obstacle that is used and inside its life window (:717-734).
// Here the stored edges live on the device as an append-only list of (start id, end id) pairs
// (drrt_edges_append, called where the planner links neighbours, drrt.cpp:256-354).  A sweep
// is one call: the conflict query stays on the device, its row marks the nodes, one pass over
// the edge list selects the edges that start in a marked node (in edge order), the Dubins /
// straight edge-check kernels run over the selection with O alone as the obstacle subset (and
// once more with the caller's mask of the other obstacles), and only the indices of the edges
// that collide with O (plus a flag byte) cross PCIe.
#include <algorithm>

#include "store.cuh"

namespace drrt {

static constexpr int SEL_T = 256;
static constexpr int SEL_PER = 4;                    // consecutive items per thread
static constexpr int SEL_TILE = SEL_T * SEL_PER;     // items per block

__global__ void mark_nodes_kernel(const int32_t *ids, int64_t n, uint8_t *mark) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mark[ids[i]] = 1;
}

// PRED 0: the edge starts in a marked node; PRED 1: the item's flag byte is set
template <int PRED>
__device__ __forceinline__ bool sel_pred(int64_t i, const int32_t *e_u, const uint8_t *mark, const uint8_t *flag) {
  return PRED == 0 ? mark[e_u[i]] != 0 : flag[i] != 0;
}

template <int PRED>
__global__ void __launch_bounds__(SEL_T) select_count_kernel(int64_t n, const int32_t *e_u, const uint8_t *mark,
                                                             const uint8_t *flag, int32_t *block_count) {
  const int64_t base = (int64_t)blockIdx.x * SEL_TILE + (int64_t)threadIdx.x * SEL_PER;
  int c = 0;
#pragma unroll
  for (int k = 0; k < SEL_PER; k++)
    if (base + k < n && sel_pred<PRED>(base + k, e_u, mark, flag)) c++;
  __shared__ int s_w[SEL_T / 32];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < SEL_T / 32; w++) t += s_w[w];
    block_count[blockIdx.x] = t;
  }
}

// writes the selected item indices in ascending order (a stable compaction): out[block_base + rank]
template <int PRED>
__global__ void __launch_bounds__(SEL_T) select_write_kernel(int64_t n, const int32_t *e_u, const uint8_t *mark,
                                                             const uint8_t *flag, const int32_t *block_base,
                                                             int32_t *out) {
  const int64_t base = (int64_t)blockIdx.x * SEL_TILE + (int64_t)threadIdx.x * SEL_PER;
  bool p[SEL_PER];
  int c = 0;
#pragma unroll
  for (int k = 0; k < SEL_PER; k++) {
    p[k] = base + k < n && sel_pred<PRED>(base + k, e_u, mark, flag);
    c += p[k];
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __shared__ int s_w[SEL_T / 32];
  if (lane == 31) s_w[w] = inc;
  __syncthreads();
  int pre = inc - c;
  for (int k = 0; k < w; k++) pre += s_w[k];
  int at = block_base[blockIdx.x] + pre;
#pragma unroll
  for (int k = 0; k < SEL_PER; k++)
    if (p[k]) out[at++] = (int32_t)(base + k);
}

__global__ void gather_edges_kernel(int64_t n, const int32_t *sel, const int32_t *e_u, const int32_t *e_v,
                                    int32_t *su, int32_t *sv) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t e = sel[i];
    su[i] = e_u[e];
    sv[i] = e_v[e];
  }
}

__global__ void sweep_out_kernel(int64_t n, const int32_t *pick, const int32_t *sel, const uint8_t *v2,
                                 int64_t *out_edge, uint8_t *out_flag) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t s = pick[i];
    out_edge[i] = sel[s];
    out_flag[i] = (uint8_t)(1u | ((v2 && v2[s]) ? 2u : 0u));
  }
}

__global__ void validate_edge_ids_kernel(const int32_t *a, const int32_t *b, int64_t n, int32_t n_nodes, int32_t *flag) {
  bool bad = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    bad |= a[i] < 0 || a[i] >= n_nodes || b[i] < 0 || b[i] >= n_nodes;
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *flag = 1;
}

// stable compaction of the indices i in [0, n) with pred(i): count per tile, scan, write; *n_out on the host
template <int PRED>
static int select_indices(Store *S, int64_t n, const int32_t *e_u, const uint8_t *mark, const uint8_t *flag,
                          DevBuf<int32_t> &out, int64_t *n_out, cudaStream_t st) {
  *n_out = 0;
  if (n == 0) return DRRT_OK;
  const int64_t nb = (n + SEL_TILE - 1) / SEL_TILE;
  DRRT_CHECK(S->sw_cnt.reserve(nb + 2, st));
  select_count_kernel<PRED><<<(unsigned)nb, SEL_T, 0, st>>>(n, e_u, mark, flag, S->sw_cnt.p);
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaMemsetAsync(S->sw_cnt.p + nb, 0, sizeof(int32_t), st));
  DRRT_CHECK(exclusive_scan_i32(S->sw_cnt.p, S->sw_cnt.p, nb + 1, S->scan_tmp2, st));
  DRRT_CUDA(cudaMemcpyAsync(S->h_scalar, S->sw_cnt.p + nb, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  DRRT_CUDA(cudaStreamSynchronize(st));
  int32_t total = 0;
  memcpy(&total, S->h_scalar, sizeof total);
  *n_out = total;
  if (total == 0) return DRRT_OK;
  DRRT_CHECK(out.reserve((size_t)total + 1, st));
  select_write_kernel<PRED><<<(unsigned)nb, SEL_T, 0, st>>>(n, e_u, mark, flag, S->sw_cnt.p, out.p);
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaGetLastError());
  return DRRT_OK;
}

// Extend's rewiring test for every (new node, neighbour) pair of the last fused Extend
// (src/drrt.cpp:330-333): the near -> new edge is safe, and
//   near.rrt_LMC_ > new.rrt_LMC_ + edge.dist_  &&  new's parent != near  &&
//   new.rrt_LMC_ + edge.dist_ < move_goal.rrt_LMC_
// with new.rrt_LMC_ / new's parent as FindBestParent left them (best_cost / best_parent).
__global__ void rewire_flag_kernel(int64_t total, const int32_t *row, const int32_t *ids, const uint8_t *unsafe,
                                   const double *len, const double *lmc, const int32_t *best,
                                   const double *best_cost, double goal_lmc, uint8_t *flag) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int32_t q = row[t], near = ids[t];
    const double via = best_cost[q] + len[2 * t + 1];   // new_node->rrt_LMC_ + this_edge->dist_
    flag[t] = (!unsafe[2 * t + 1] && lmc[near] > via && best[q] != near && via < goal_lmc) ? 1 : 0;
  }
}

__global__ void rewire_out_kernel(int64_t n, const int32_t *pick, const int32_t *row, const int32_t *ids,
                                  const double *len, const double *best_cost, int32_t *out_q, int32_t *out_near,
                                  double *out_lmc) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t t = pick[i];
    out_q[i] = row[t];
    out_near[i] = ids[t];
    out_lmc[i] = best_cost[row[t]] + len[2 * (int64_t)t + 1];
  }
}

static int stage_mask(Store *S, const uint8_t *mask_over_table, cudaStream_t st) {
  std::vector<uint8_t> act(S->h_o_src.size() + 1, 0);
  for (size_t a = 0; a < S->h_o_src.size(); a++) act[a] = mask_over_table[S->h_o_src[a]] ? 1 : 0;
  DRRT_CHECK(S->o_mask.reserve(act.size(), st));
  DRRT_CUDA(cudaMemcpyAsync(S->o_mask.p, act.data(), act.size(), cudaMemcpyHostToDevice, st));
  DRRT_CUDA(cudaStreamSynchronize(st));  // `act` dies at return
  return DRRT_OK;
}

}  // namespace drrt

using namespace drrt;

extern "C" {

int drrt_edges_append(drrt_store *h, int64_t n, const int32_t *start_id, const int32_t *end_id,
                      int64_t *first_edge) {
  ENTER(h);
  if (n < 0 || (n > 0 && (!start_id || !end_id))) return set_error(DRRT_ERR_INVALID, "edges_append: arguments");
  if (first_edge) *first_edge = S->n_stored_edges;
  if (n == 0) return DRRT_OK;
  if (S->n_stored_edges + n > 0x7ffffff0) return set_error(DRRT_ERR_UNSUPPORTED, "edges_append: more than 2^31 stored edges");
  cudaStream_t st = S->stream;
  const size_t keep = (size_t)S->n_stored_edges;
  DRRT_CHECK(S->e_u.reserve(keep + n, st, true, keep));
  DRRT_CHECK(S->e_v.reserve(keep + n, st, true, keep));
  DRRT_CUDA(cudaMemcpyAsync(S->e_u.p + keep, start_id, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  DRRT_CUDA(cudaMemcpyAsync(S->e_v.p + keep, end_id, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  DRRT_CUDA(cudaMemsetAsync(S->d_pending, 0, sizeof(int32_t), st));
  validate_edge_ids_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(
      S->e_u.p + keep, S->e_v.p + keep, n, (int32_t)S->n, S->d_pending);
  DRRT_LAUNCHED();
  DRRT_CUDA(cudaMemcpyAsync(S->h_pending, S->d_pending, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  DRRT_CUDA(cudaStreamSynchronize(st));
  if (*S->h_pending) return set_error(DRRT_ERR_INVALID, "edges_append: node id outside the store");
  S->n_stored_edges += n;
  S->sweep_out = -1;
  return DRRT_OK;
}

int drrt_edges_count(drrt_store *h, int64_t *n) {
  ENTER(h);
  if (!n) return set_error(DRRT_ERR_INVALID, "edges_count: n == NULL");
  *n = S->n_stored_edges;
  return DRRT_OK;
}

int drrt_edges_clear(drrt_store *h) {
  ENTER(h);
  S->n_stored_edges = 0;
  S->sweep_out = -1;
  return DRRT_OK;
}

int drrt_obstacle_sweep(drrt_store *h, int32_t obstacle_index, const uint8_t *other_mask,
                        int32_t n_obstacles, int edge_kind, double robot_radius, double r_min,
                        int64_t *n_conflict_nodes, int64_t *n_edges_checked, int64_t *n_out) {
  ENTER(h);
  S->sweep_out = -1;
  S->rewire_out = -1;  // shares the selection scratch
  if (!S->has_obstacles) return set_error(DRRT_ERR_STATE, "sweep: no obstacle table");
  const int o = obstacle_index;
  if (o < 0 || o >= S->obs.n_all) return set_error(DRRT_ERR_INVALID, "sweep: obstacle %d", o);
  if (other_mask && n_obstacles != S->obs.n_all)
    return set_error(DRRT_ERR_INVALID, "sweep: mask has %d entries, the obstacle table %d", n_obstacles, S->obs.n_all);
  const int kind = S->h_all_kind[o];
  if (kind < 1 || kind > 5)  // obstacle.cpp:547, :608 "This case has not been coded yet."
    return set_error(DRRT_ERR_UNSUPPORTED, "sweep: obstacle kind %d", kind);
  cudaStream_t st = S->stream;
  if (n_conflict_nodes) *n_conflict_nodes = 0;
  if (n_edges_checked) *n_edges_checked = 0;
  if (n_out) *n_out = 0;
  S->sweep_out = 0;
  if (S->n == 0 || S->n_stored_edges == 0) return DRRT_OK;
  // 1. FindPointsInConflictWithObstacle, Dubins space without time (obstacle.cpp:554-559)
  DRRT_CHECK(map_reserve(S));
  MappedRow *M = reinterpret_cast<MappedRow *>(S->h_map);
  M->q[0] = S->h_all_ox[o]; M->q[1] = S->h_all_oy[o]; M->q[2] = DRRT_PI;
  drrt_range_result res{};
  DRRT_CHECK(range_batch_dev(S, 1, M->q, nullptr, robot_radius + S->h_all_rad[o] + DRRT_PI, st, &res));
  if (n_conflict_nodes) *n_conflict_nodes = res.total;
  if (res.total == 0) return DRRT_OK;
  // 2. mark the nodes, select the stored edges that start in one (edge order)
  DRRT_CHECK(S->node_mark.reserve((size_t)S->n + 1, st));
  DRRT_CUDA(cudaMemsetAsync(S->node_mark.p, 0, (size_t)S->n, st));
  mark_nodes_kernel<<<(unsigned)std::min<int64_t>((res.total + 255) / 256, 148 * 8), 256, 0, st>>>(
      res.ids_dev, res.total, S->node_mark.p);  // nq == 1: the row starts at 0
  DRRT_LAUNCHED();
  int64_t nsel = 0;
  DRRT_CHECK(select_indices<0>(S, S->n_stored_edges, S->e_u.p, S->node_mark.p, nullptr, S->sw_sel, &nsel, st));
  if (n_edges_checked) *n_edges_checked = nsel;
  if (nsel == 0) return DRRT_OK;
  DRRT_CHECK(S->sw_u.reserve((size_t)nsel, st));
  DRRT_CHECK(S->sw_v.reserve((size_t)nsel, st));
  DRRT_CHECK(S->sw_v1.reserve((size_t)nsel, st));
  gather_edges_kernel<<<(unsigned)std::min<int64_t>((nsel + 255) / 256, 148 * 8), 256, 0, st>>>(
      nsel, S->sw_sel.p, S->e_u.p, S->e_v.p, S->sw_u.p, S->sw_v.p);
  DRRT_LAUNCHED();
  // 3. edge->ExplicitEdgeCheck(O) (obstacle.cpp:643, 712): O alone as the obstacle subset
  std::vector<uint8_t> only(S->obs.n_all, 0);
  only[o] = 1;
  DRRT_CHECK(stage_mask(S, only.data(), st));
  S->edge_mask = S->o_mask.p;
  int rc = edgecheck_batch_dev(S, nsel, nullptr, nullptr, S->sw_u.p, S->sw_v.p, edge_kind, robot_radius, r_min,
                               S->sw_v1.p, nullptr, st);
  S->edge_mask = nullptr;
  DRRT_CHECK(rc);
  const uint8_t *v2 = nullptr;
  if (other_mask) {  // RemoveObstacle's re-check against the other live obstacles (:717-734)
    DRRT_CHECK(S->sw_v2.reserve((size_t)nsel, st));
    std::vector<uint8_t> others(other_mask, other_mask + n_obstacles);
    others[o] = 0;
    DRRT_CUDA(cudaStreamSynchronize(st));  // the first check still reads o_mask
    DRRT_CHECK(stage_mask(S, others.data(), st));
    S->edge_mask = S->o_mask.p;
    rc = edgecheck_batch_dev(S, nsel, nullptr, nullptr, S->sw_u.p, S->sw_v.p, edge_kind, robot_radius, r_min,
                             S->sw_v2.p, nullptr, st);
    S->edge_mask = nullptr;
    DRRT_CHECK(rc);
    v2 = S->sw_v2.p;
  }
  // 4. the edges that collide with O, in edge order
  int64_t nhit = 0;
  DRRT_CHECK(select_indices<1>(S, nsel, nullptr, nullptr, S->sw_v1.p, S->sw_pick, &nhit, st));
  if (nhit > 0) {
    DRRT_CHECK(S->sw_out_edge.reserve((size_t)nhit, st));
    DRRT_CHECK(S->sw_out_flag.reserve((size_t)nhit, st));
    sweep_out_kernel<<<(unsigned)std::min<int64_t>((nhit + 255) / 256, 148 * 8), 256, 0, st>>>(
        nhit, S->sw_pick.p, S->sw_sel.p, v2, S->sw_out_edge.p, S->sw_out_flag.p);
    DRRT_LAUNCHED();
    DRRT_CUDA(cudaGetLastError());
  }
  DRRT_CUDA(cudaStreamSynchronize(st));
  S->sweep_out = nhit;
  if (n_out) *n_out = nhit;
  return DRRT_OK;
}

int drrt_obstacle_sweep_fetch(drrt_store *h, int64_t *edge_index, uint8_t *flags) {
  ENTER(h);
  if (S->sweep_out < 0) return set_error(DRRT_ERR_STATE, "sweep_fetch: no sweep on this handle");
  cudaStream_t st = S->stream;
  if (S->sweep_out > 0) {
    if (edge_index)
      DRRT_CUDA(cudaMemcpyAsync(edge_index, S->sw_out_edge.p, S->sweep_out * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (flags) DRRT_CUDA(cudaMemcpyAsync(flags, S->sw_out_flag.p, S->sweep_out, cudaMemcpyDeviceToHost, st));
  }
  DRRT_CUDA(cudaStreamSynchronize(st));
  return DRRT_OK;
}

}  // extern "C"

extern "C" {

int drrt_extend_rewire(drrt_store *h, double move_goal_lmc, int64_t *n_out) {
  ENTER(h);
  S->rewire_out = -1;
  if (S->last_extend_total < 0 || !S->ext_best || !S->ext_cost || !S->ext_lmc)
    return set_error(DRRT_ERR_STATE, "extend_rewire: no fused Extend with a parent reduction on this handle");
  const int64_t total = S->last_extend_total;
  if (n_out) *n_out = 0;
  S->rewire_out = 0;
  if (total == 0) return DRRT_OK;
  if (total > 0x7ffffff0) return set_error(DRRT_ERR_UNSUPPORTED, "extend_rewire: more than 2^31 pairs");
  cudaStream_t st = S->stream;
  DRRT_CHECK(S->sw_v1.reserve((size_t)total, st));
  rewire_flag_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, st>>>(
      total, S->x_row.p, S->csr_ids, S->x_unsafe.p, S->x_len.p, S->ext_lmc, S->ext_best, S->ext_cost,
      move_goal_lmc, S->sw_v1.p);
  DRRT_LAUNCHED();
  int64_t n = 0;
  DRRT_CHECK(select_indices<1>(S, total, nullptr, nullptr, S->sw_v1.p, S->sw_pick, &n, st));
  if (n > 0) {
    DRRT_CHECK(S->sw_u.reserve((size_t)n, st));
    DRRT_CHECK(S->sw_v.reserve((size_t)n, st));
    DRRT_CHECK(S->rw_lmc.reserve((size_t)n, st));
    rewire_out_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(
        n, S->sw_pick.p, S->x_row.p, S->csr_ids, S->x_len.p, S->ext_cost, S->sw_u.p, S->sw_v.p, S->rw_lmc.p);
    DRRT_LAUNCHED();
    DRRT_CUDA(cudaGetLastError());
  }
  DRRT_CUDA(cudaStreamSynchronize(st));
  S->rewire_out = n;
  S->sweep_out = -1;  // the sweep's scratch was reused
  if (n_out) *n_out = n;
  return DRRT_OK;
}

int drrt_extend_rewire_fetch(drrt_store *h, int32_t *query_index, int32_t *near_id, double *new_lmc) {
  ENTER(h);
  if (S->rewire_out < 0) return set_error(DRRT_ERR_STATE, "extend_rewire_fetch: no drrt_extend_rewire on this handle");
  cudaStream_t st = S->stream;
  const int64_t n = S->rewire_out;
  if (n > 0) {
    if (query_index) DRRT_CUDA(cudaMemcpyAsync(query_index, S->sw_u.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (near_id) DRRT_CUDA(cudaMemcpyAsync(near_id, S->sw_v.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (new_lmc) DRRT_CUDA(cudaMemcpyAsync(new_lmc, S->rw_lmc.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  DRRT_CUDA(cudaStreamSynchronize(st));
  return DRRT_OK;
}

}  // extern "C"

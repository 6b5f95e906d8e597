/* drrt_gpu.h -- C ABI of drrt_b200 (libdrrt_b200.so).
 *
 * The drop-in boundary for the data-parallel hot path of cosa0481/DRRT:
 * batched KD-tree neighbourhood queries and batched edge collision checks.
 * The reference has no FFI / plugin API; the boundary is the
 * set of C++ methods cited beside each entry point below (paths relative to
 * the reference root).  drrt_b200/shim/ re-implements those method signatures
 * on top of this header; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *  - plain pointers and sizes only; every call returns an int status
 *    (DRRT_OK == 0) and never throws or exits (the reference prints to stdout
 *    and returns bool; fatal states call exit(-1)).
 *  - node ids are dense insertion indices (id 0 == KDTree::root).
 *  - `*_dev` entry points take DEVICE pointers and a cudaStream_t (as void*;
 *    NULL = the store's own stream); the others take caller-owned HOST
 *    pointers (pinned memory preferred) and include the copies.
 *  - there is no CPU fallback: without a usable sm_100 device every call
 *    fails with DRRT_ERR_NO_DEVICE / DRRT_ERR_CUDA.
 *  - one store may be entered from several host threads (the reference takes
 *    KDTree::tree_mutex_ around every tree call and cspace_mutex_ inside the
 *    edge check); calls on one handle are serialised internally.
 */
#ifndef DRRT_GPU_H
#define DRRT_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRRT_ABI_VERSION 4 /* 4: narrow rows, rrt_LMC_ mirror, drrt_comm_* (N GPUs, one process),
                             stored-edge CSR, obstacle moves; 3: drrt_range_batch_into */

#if defined(__GNUC__)
#define DRRT_API __attribute__((visibility("default")))
#else
#define DRRT_API
#endif

enum {
  DRRT_OK = 0,
  DRRT_ERR_INVALID = 1,     /* bad argument */
  DRRT_ERR_UNSUPPORTED = 2, /* configuration the device path does not cover */
  DRRT_ERR_CUDA = 3,        /* a CUDA call failed; see drrt_last_error() */
  DRRT_ERR_NOMEM = 4,
  DRRT_ERR_NO_DEVICE = 5,
  DRRT_ERR_STATE = 6,       /* call order (e.g. fetch without a batch) */
  DRRT_ERR_CAPACITY = 7     /* caller buffers too small; the sizes needed were returned */
};

/* the reference passes a host function pointer (include/DRRT/kdtree.h:21,47-49);
 * a device cannot call it, so the two metrics in the tree are enumerated */
enum {
  DRRT_METRIC_R2S1_WRAP = 0, /* distance_function  src/smalltest.cpp:167-175 */
  DRRT_METRIC_EUCLID = 1     /* DistFunc           src/thetastar.cpp:10-15   */
};

enum {
  DRRT_EDGE_DUBINS = 0,  /* DubinsEdge::CalculateTrajectory + ExplicitEdgeCheck */
  DRRT_EDGE_STRAIGHT = 1 /* LineCheck  src/obstacle.cpp:1092-1143               */
};

typedef struct drrt_store drrt_store;

DRRT_API const char *drrt_last_error(void);
DRRT_API int drrt_abi_version(void);
/* number of kernels this library has launched in this process (all handles) */
DRRT_API int64_t drrt_kernel_launch_count(void);

/* ---- node store + index ------------------------------------------------ */

/* KDTree::KDTree(int d, VectorXi wraps, VectorXd wrapPoints)
 * include/DRRT/kdtree.h:33-36 + SetDistanceFunction :47-49.
 * Supported: dims == 3; n_wraps == 0, or n_wraps == 1 with wrap_dims[0] == 2
 * (and, for DRRT_METRIC_R2S1_WRAP, wrap_points[0] == 2.0*3.1415926536). */
DRRT_API int drrt_store_create(int device, int dims, int n_wraps, const int *wrap_dims,
                      const double *wrap_points, int metric,
                      int64_t capacity_hint, drrt_store **out);
DRRT_API int drrt_store_destroy(drrt_store *h);
DRRT_API int drrt_store_size(drrt_store *h, int64_t *n); /* KDTree::tree_size_ kdtree.h:23 */
DRRT_API void *drrt_store_stream(drrt_store *h);        /* cudaStream_t of the handle */

/* KDTree::KDInsert  src/kdtree.cpp:87-136, n nodes in order (pos = n x 3 f64,
 * row-major).  first_id receives the id of pos[0]. Builds the same
 * insertion-ordered KD tree (parent / children / split) on the device. */
DRRT_API int drrt_store_insert(drrt_store *h, int64_t n, const double *pos,
                      int32_t *first_id);
DRRT_API int drrt_store_insert_dev(drrt_store *h, int64_t n, const double *pos_dev,
                          void *stream, int32_t *first_id);

/* KDTreeNode::kd_parent_/kd_child_L_/kd_child_R_/kd_split_
 * include/DRRT/kdtreenode.h:17-20,29 for ids [first, first+n): -1 = none.
 * (drrt.cpp:849 and KDTree::GetNodeAt kdtree.cpp:56-85 read them.) */
DRRT_API int drrt_store_structure(drrt_store *h, int64_t first, int64_t n,
                         int32_t *parent, int32_t *left, int32_t *right,
                         int32_t *split);
DRRT_API int drrt_store_positions(drrt_store *h, int64_t first, int64_t n, double *pos);

/* ---- range ball ---------------------------------------------------------
 * KDTree::KDFindWithinRange src/kdtree.cpp:794-820 (and the identical
 * KDFindMoreWithinRange :822-851), ghost loop src/ghostpoint.cpp:15-75.
 * radius: per-query array (nq) or NULL to use radius_all.
 * Result: CSR over queries; each row sorted by node id; dist = the key the
 * reference stores in the JList node (metric to the probe that found it).
 * Two-call sizing: batch -> counts[nq] and *total; fetch -> ids/dist[total]
 * laid out by the exclusive prefix sum of counts. */
DRRT_API int drrt_range_batch(drrt_store *h, int64_t nq, const double *q,
                     const double *radius, double radius_all, int32_t *counts,
                     int64_t *total);
DRRT_API int drrt_range_fetch(drrt_store *h, int32_t *ids, double *dist);

/* The same query in ONE call, for a caller that can bound the result size
 * (e.g. from the previous batch: the RRTx radius shrinks monotonically,
 * src/drrt.cpp:16-27).  ids/dist hold cap elements (dist may be NULL);
 * offsets[nq+1] receives the CSR row starts, *total the number of hits.
 * The queries are answered in sub-batches on the handle's stream and the rows
 * of a finished sub-batch leave for the host on a second stream while the next
 * one is computed, so the PCIe transfer of the rows (12 B per hit) overlaps the
 * kernel instead of following it.  If total > cap the call returns
 * DRRT_ERR_CAPACITY: offsets and *total are complete, rows that did not fit
 * whole are not copied, and a retry with cap >= *total succeeds.  Afterwards
 * the handle holds no result (drrt_range_fetch / drrt_range_pack need a
 * drrt_range_batch[_dev] first). */
DRRT_API int drrt_range_batch_into(drrt_store *h, int64_t nq, const double *q,
                          const double *radius, double radius_all, int64_t cap,
                          int64_t *offsets, int32_t *ids, double *dist,
                          int64_t *total);

/* Narrow rows, opt-in: the same call with the distances rounded to float (ids bit-exact;
 * the north star asks for distances within 1e-5 relative, a float carries 6e-8) -- 8 instead
 * of 12 bytes per hit cross PCIe, which is what bounds the host call.  dist == NULL in either
 * form returns the id sets alone (4 bytes per hit). */
DRRT_API int drrt_range_batch_into_f32(drrt_store *h, int64_t nq, const double *q,
                              const double *radius, double radius_all, int64_t cap,
                              int64_t *offsets, int32_t *ids, float *dist, int64_t *total);

/* pinned (page-locked, portable) host memory for callers without a CUDA toolchain, e.g. the
 * C++ shim: result buffers in it are written by DMA straight from the device */
DRRT_API int drrt_host_alloc(size_t bytes, void **out);
DRRT_API int drrt_host_free(void *p);

/* device-resident variant: queries/radii on the device; results stay in
 * library-owned device buffers valid until the next range call on h.
 * Row q is ids_dev/dist_dev[offsets_dev[q] .. offsets_dev[q] + counts_dev[q]),
 * rows in query order.  Large neighbourhoods are written in slices of a few
 * consecutive rows with unused slots between slices (packed == 0); when
 * packed == 1 the rows are back to back and offsets_dev (nq+1 entries) is the
 * exclusive prefix sum of counts_dev, a plain CSR.  drrt_range_pack turns the
 * last result of h into that form (one device-side copy) and returns it. */
typedef struct {
  int64_t nq;
  int64_t total;               /* sum of counts */
  const int64_t *offsets_dev;  /* nq + 1 */
  const int32_t *ids_dev;
  const double *dist_dev;
  const int32_t *counts_dev;   /* nq */
  int64_t extent;              /* elements of ids_dev / dist_dev the rows span */
  int32_t packed;
} drrt_range_result;
DRRT_API int drrt_range_batch_dev(drrt_store *h, int64_t nq, const double *q_dev,
                         const double *radius_dev, double radius_all,
                         void *stream, drrt_range_result *out);
DRRT_API int drrt_range_pack(drrt_store *h, void *stream, drrt_range_result *out);

/* ---- nearest / k-nearest -------------------------------------------------
 * KDTree::KDFindNearest src/kdtree.cpp:264-307: id and distance of the
 * nearest node (ties: smaller id).
 * KDTree::KDFindKNearest src/kdtree.cpp:630-666 (dead and broken upstream,
 * SURVEY F7): by intent, the k smallest metric(q,n), ties by smaller id,
 * ascending (dist,id); rows padded with id -1 when the store has < k nodes. */
DRRT_API int drrt_nearest_batch(drrt_store *h, int64_t nq, const double *q,
                       int32_t *ids, double *dist);
DRRT_API int drrt_nearest_batch_dev(drrt_store *h, int64_t nq, const double *q_dev,
                           int32_t *ids_dev, double *dist_dev, void *stream);
DRRT_API int drrt_knn_batch(drrt_store *h, int64_t nq, const double *q, int k,
                   int32_t *ids, double *dist);
DRRT_API int drrt_knn_batch_dev(drrt_store *h, int64_t nq, const double *q_dev, int k,
                       int32_t *ids_dev, double *dist_dev, void *stream);

/* ---- obstacles + edge checks --------------------------------------------
 * Flat copy of ConfigSpace::obstacles_ (List of Obstacle,
 * include/DRRT/obstacle.h:16-211) with the fields the analytic checker reads:
 * kind_, obstacle_used_, life_span_, origin_(0..1), radius_,
 * shape_.GetPolygon() rows -- all in the WORLD frame.
 * Kinds 6/7 (time-parametrised) are refused here: they have their own table,
 * drrt_timed_obstacles_set below. */
DRRT_API int drrt_obstacles_set(drrt_store *h, int32_t n_obs, const int32_t *kind,
                       const uint8_t *used, const double *life_span,
                       const double *origin_xy, const double *radius,
                       const int32_t *vert_off, const double *verts_xy);

/* ExplicitEdgeCheck(C, edge) src/obstacle.cpp:879-923 over
 * Edge::NewEdge + DubinsEdge::CalculateTrajectory src/dubinsedge.cpp:12-21,
 * 172-830 + DubinsEdge::ExplicitEdgeCheck :850-898 with the analytic
 * ExplicitEdgeCheck2D src/obstacle.cpp:756-804 per polyline segment
 * (edge_kind DRRT_EDGE_STRAIGHT: the LineCheck polyline instead).
 * start/end: n x 3 f64 poses.  verdict[i] = 1 when edge i collides.
 * length (optional): edge->dist_ after CalculateTrajectory.
 * in_warmup != 0 reproduces ConfigSpace::in_warmup_time_ (:883): all free. */
DRRT_API int drrt_edgecheck_batch(drrt_store *h, int64_t n, const double *start,
                         const double *end, int edge_kind, double robot_radius,
                         double r_min, int in_warmup, uint8_t *verdict,
                         double *length);
/* same, endpoints named by node id in the store (8 B per edge on the wire) */
DRRT_API int drrt_edgecheck_ids_batch(drrt_store *h, int64_t n, const int32_t *start_id,
                             const int32_t *end_id, int edge_kind,
                             double robot_radius, double r_min, int in_warmup,
                             uint8_t *verdict, double *length);
DRRT_API int drrt_edgecheck_batch_dev(drrt_store *h, int64_t n, const double *start_dev,
                             const double *end_dev, const int32_t *start_id_dev,
                             const int32_t *end_id_dev, int edge_kind,
                             double robot_radius, double r_min,
                             uint8_t *verdict_dev, double *length_dev,
                             void *stream);

/* ---- time-parametrised obstacles: ExplicitEdgeCheck2D kinds 6 / 7 -------------------
 * src/obstacle.cpp:805-875 (FindIndexBeforeTime src/drrt.cpp:178-189).  A disc of radius_ whose
 * centre moves along path rows (dx, dy, time) relative to origin_; a segment is a robot move
 * (x, y, TIME) -> (x, y, TIME).  verdict[i] = OR over the obstacles of ExplicitEdgeCheck2D(O,
 * start_i, end_i, radius): for every path piece that overlaps the move in time, the centres at
 * their (clamped) time of closest approach are closer than radius_ + radius.  The reference's
 * planner never reaches this branch (space_has_time_ == false, src/smalltest.cpp:200); it is
 * offered as the per-segment test it is.  path_off: n_obs + 1 row offsets into path_xyt. */
DRRT_API int drrt_timed_obstacles_set(drrt_store *h, int32_t n_obs, const int32_t *kind,
                             const uint8_t *used, const double *life_span,
                             const double *origin_xy, const double *radius,
                             const int32_t *path_off, const double *path_xyt);
DRRT_API int drrt_edgecheck2d_timed_batch(drrt_store *h, int64_t n, const double *start_xyt,
                                 const double *end_xyt, double radius, uint8_t *verdict);
DRRT_API int drrt_edgecheck2d_timed_batch_dev(drrt_store *h, int64_t n, const double *start_dev,
                                     const double *end_dev, double radius,
                                     uint8_t *verdict_dev, void *stream);

/* ---- obstacle add / remove sweeps ------------------------
 * AddObstacle src/obstacle.cpp:612-667 and RemoveObstacle :669-754 re-check the
 * stored edges of every node near an obstacle against THAT obstacle only
 * (`edge->ExplicitEdgeCheck(O)`, i.e. DubinsEdge::ExplicitEdgeCheck
 * src/dubinsedge.cpp:850-898 for one obstacle), and RemoveObstacle then against
 * every OTHER obstacle that is used and inside its life window (:717-734).
 * Both are the edge check above restricted to a subset of the table given to
 * drrt_obstacles_set: obstacle_mask[o] != 0 for the obstacles taking part
 * (n_obstacles must equal the table's length; the host decides the life-window
 * part of the mask).  Edges are n x 3 poses (start/end) or node ids
 * (start_id/end_id); pass the unused pair as NULL.
 * drrt_obstacle_conflict_batch is FindPointsInConflictWithObstacle :540-610
 * for the Dubins space without time (:554-559): one range query per listed
 * obstacle, centre (origin_(0), origin_(1), PI), radius robot_radius +
 * radius_ + PI; counts/total as drrt_range_batch, rows through
 * drrt_range_fetch.  Kinds outside 1..5 -> DRRT_ERR_UNSUPPORTED. */
DRRT_API int drrt_edgecheck_subset_batch(drrt_store *h, int64_t n, const double *start,
                             const double *end, const int32_t *start_id,
                             const int32_t *end_id, int edge_kind, double robot_radius,
                             double r_min, const uint8_t *obstacle_mask,
                             int32_t n_obstacles, uint8_t *verdict, double *length);
DRRT_API int drrt_edgecheck_subset_batch_dev(drrt_store *h, int64_t n, const double *start_dev,
                             const double *end_dev, const int32_t *start_id_dev,
                             const int32_t *end_id_dev, int edge_kind,
                             double robot_radius, double r_min,
                             const uint8_t *obstacle_mask /* host */, int32_t n_obstacles,
                             uint8_t *verdict_dev, double *length_dev, void *stream);
DRRT_API int drrt_obstacle_conflict_batch(drrt_store *h, int32_t n_sel,
                             const int32_t *obstacle_index, double robot_radius,
                             int32_t *counts, int64_t *total);

/* Obstacle::UpdateObstacles -> NextOrigin (src/obstacle.cpp:12-22, 182-195): a moving obstacle's
 * origin_ jumps to the next point of its path.  Only the listed entries cross the boundary; the
 * device table and its broadphase follow.  translate_polygon == 0: origin_ alone moves, which is
 * exactly what the analytic checker then sees (it reads shape_.GetPolygon() as world coordinates,
 * obstacle.cpp:785-794, SURVEY F10: the bounding reject :765-779 follows the new origin, the
 * polygon stays); != 0: the vertices move with the origin (the frame of GetPosition(),
 * obstacle.cpp:9-10).  drrt_obstacles_update switches listed obstacles on / off
 * (obstacle_used_, life_span_; either array may be NULL), as the obstacle thread does
 * (obstacle.cpp:454-499) and as RemoveObstacle ends (:753). */
DRRT_API int drrt_obstacles_move(drrt_store *h, int32_t n_sel, const int32_t *obstacle_index,
                        const double *new_origin_xy, int translate_polygon);
DRRT_API int drrt_obstacles_update(drrt_store *h, int32_t n_sel, const int32_t *obstacle_index,
                          const uint8_t *used, const double *life_span);

/* The sweeps on a DEVICE-RESIDENT list of the planner's stored neighbour edges.
 * drrt_edges_append: the directed edges the planner keeps in its neighbour lists and as parent
 * edges (linked in Extend, src/drrt.cpp:256-354; walked by RrtNodeNeighborIterator in
 * AddObstacle / RemoveObstacle), named by node id; *first_edge receives the index of the first
 * appended edge (edge indices are append order).
 * drrt_obstacle_sweep does AddObstacle's (other_mask == NULL) or RemoveObstacle's work for one
 * obstacle in one call: the conflict query (obstacle.cpp:540-610) stays on the device, every
 * stored edge that STARTS in a node it finds is checked against obstacle_index alone
 * (edge->ExplicitEdgeCheck(O), :643, :651, :712) and, with other_mask (n_obstacles bytes, the
 * obstacles that are used and inside their life window, :717-734; the entry of obstacle_index is
 * ignored), against those others.  n_out = the edges that collide with O;
 * drrt_obstacle_sweep_fetch returns their indices in ascending order and a flag byte each
 * (bit 0: collides with O; bit 1: also collides with another obstacle of the mask, i.e.
 * RemoveObstacle must leave it blocked).  What the planner does with them (dist_ = INF, parent
 * unlinking, RecalculateLMC) stays host code. */
DRRT_API int drrt_edges_append(drrt_store *h, int64_t n, const int32_t *start_id,
                      const int32_t *end_id, int64_t *first_edge);
DRRT_API int drrt_edges_count(drrt_store *h, int64_t *n);
DRRT_API int drrt_edges_clear(drrt_store *h);
DRRT_API int drrt_obstacle_sweep(drrt_store *h, int32_t obstacle_index, const uint8_t *other_mask,
                        int32_t n_obstacles, int edge_kind, double robot_radius, double r_min,
                        int64_t *n_conflict_nodes, int64_t *n_edges_checked, int64_t *n_out);
DRRT_API int drrt_obstacle_sweep_fetch(drrt_store *h, int64_t *edge_index, uint8_t *flags);

/* ---- fused Extend step (the step after the edge check) ----------
 * What Extend does around the hot path for one new node (src/drrt.cpp:194-362):
 * KDFindWithinRange (:213-216), then for every neighbour the new->near edge
 * (FindBestParent :367-460) and the near->new edge (:284-293), each solved and
 * collision-checked, and the parent chosen as the safe new->near edge that
 * minimises near.rrt_LMC_ + edge.dist_ (:444-449).  Here for nq new nodes at
 * once, the neighbour lists never leaving the device between the two stages.
 *   lmc          rrt_LMC_ of every node in the store (host array, store-size entries,
 *                uploaded whole), or NULL: use the device mirror (drrt_lmc_write /
 *                drrt_lmc_scatter); with best_parent == NULL no reduction is done
 *   counts/total as drrt_range_batch
 *   best_parent  node id per query (-1: no safe neighbour), best_cost its
 *                near.lmc + length (INF = 1e12 when none); ties: smaller id
 * drrt_extend_fetch returns the rows (ids, dist as drrt_range_fetch) and, per
 * neighbour t, unsafe[2t], length[2t] for new->near and [2t+1] for near->new. */
/* Device mirror of KDTreeNode::rrt_LMC_ (include/DRRT/kdtreenode.h:35-73; read by
 * FindBestParent src/drrt.cpp:444-446, written by RecalculateLMC / Rewire :570-780).  One
 * entry per node of the store; nodes never written hold INF = 1e12, the value a new node
 * carries until it gets a parent (drrt.cpp:391).  The planner sends the entries it changed
 * (write: a contiguous id range, scatter: listed ids) instead of the whole array per call;
 * drrt_extend_batch with lmc == NULL and best_parent != NULL reduces against the mirror. */
DRRT_API int drrt_lmc_write(drrt_store *h, int64_t first, int64_t n, const double *values);
DRRT_API int drrt_lmc_scatter(drrt_store *h, int64_t n, const int32_t *ids, const double *values);
DRRT_API int drrt_lmc_read(drrt_store *h, int64_t first, int64_t n, double *values);

DRRT_API int drrt_extend_batch(drrt_store *h, int64_t nq, const double *q, const double *radius,
                      double radius_all, double robot_radius, double r_min, int in_warmup,
                      const double *lmc, int32_t *counts, int64_t *total,
                      int32_t *best_parent, double *best_cost);
DRRT_API int drrt_extend_fetch(drrt_store *h, int32_t *ids, double *dist, uint8_t *unsafe,
                      double *length);
/* Extend's rewiring test (src/drrt.cpp:330-333) over the pairs of the last drrt_extend_batch
 * that reduced a parent: near -> new is safe, near.rrt_LMC_ > new.rrt_LMC_ + dist_, new's parent
 * is not near, and new.rrt_LMC_ + dist_ < move_goal.rrt_LMC_ (new.rrt_LMC_ = best_cost, new's
 * parent = best_parent).  Only the candidates come back, in pair order: the index of the new node
 * in the batch, the neighbour's id and the rrt_LMC_ it would get.  MakeParentOf, the queue and
 * the change threshold (:334-346) stay host code. */
DRRT_API int drrt_extend_rewire(drrt_store *h, double move_goal_lmc, int64_t *n_out);
DRRT_API int drrt_extend_rewire_fetch(drrt_store *h, int32_t *query_index, int32_t *near_id,
                             double *new_lmc);
DRRT_API int drrt_extend_batch_dev(drrt_store *h, int64_t nq, const double *q_dev,
                          const double *radius_dev, double radius_all, double robot_radius,
                          double r_min, const double *lmc_dev, void *stream,
                          drrt_range_result *rows, const uint8_t **unsafe_dev,
                          const double **length_dev, int32_t *best_parent_dev,
                          double *best_cost_dev);

/* DubinsEdge::CalculateTrajectory polyline (edge->trajectory_, read by
 * src/moverobot.cpp:107,158): up to DRRT_TRAJ_MAX (x,y) points per edge.
 * xy: n x DRRT_TRAJ_MAX x 2 f64; npts: n; type: n x 4 chars ("rsl\0"...).
 * xy == NULL (and npts == NULL): only the Dubins word and edge->dist_ are returned. */
#define DRRT_TRAJ_MAX 200
DRRT_API int drrt_dubins_trajectory_batch(drrt_store *h, int64_t n, const double *start,
                                 const double *end, double r_min, double *xy,
                                 int32_t *npts, char *type, double *length);

/* ExplicitNodeCheck -> ExplicitPointCheck src/obstacle.cpp:1062-1090
 * (QuickCheck2D :925-965, ExplicitPointCheck2D :984-1060): 1 = in collision */
DRRT_API int drrt_pointcheck_batch(drrt_store *h, int64_t n, const double *pts,
                          double robot_radius, double collision_distance,
                          uint8_t *verdict);
DRRT_API int drrt_pointcheck_batch_dev(drrt_store *h, int64_t n, const double *pts_dev,
                              double robot_radius, double collision_distance,
                              uint8_t *verdict_dev, void *stream);

/* ---- N GPUs behind one handle, one host process --------------------
 * The reference planner is one process whose threads share one KDTree and one ConfigSpace
 * (src/smalltest.cpp:37-39, 132-142).  drrt_comm keeps one REPLICA of the node store (and of the
 * obstacle table and the rrt_LMC_ mirror) per GPU:
 *   - drrt_comm_insert / drrt_comm_lmc_*: the data crosses PCIe once, to the first GPU, and
 *     reaches the others by ncclBroadcast over NVLink (ncclCommInitAll, one communicator per GPU
 *     in this process; libnccl.so.2 is loaded at run time);
 *   - query and edge batches are SHARDED into contiguous index ranges, one per GPU, with no
 *     exchange step inside a query or an edge check; every GPU copies its rows straight into
 *     the caller's buffers at its global offset (the sum of the preceding shards' hit totals).
 * Results are identical for every GPU count (ids are insertion indices, rows sorted by id).
 * devices == NULL: GPUs 0 .. n_gpus-1.  The other arguments are drrt_store_create's.  Calls on
 * one comm are serialised; the entry points mirror the single-store ones of the same name. */
typedef struct drrt_comm drrt_comm;
DRRT_API int drrt_comm_init(int n_gpus, const int *devices, int dims, int n_wraps,
                   const int *wrap_dims, const double *wrap_points, int metric,
                   int64_t capacity_hint, drrt_comm **out);
DRRT_API int drrt_comm_destroy(drrt_comm *c);
DRRT_API int drrt_comm_size(drrt_comm *c, int *n_gpus);
DRRT_API drrt_store *drrt_comm_store(drrt_comm *c, int rank); /* replica `rank` (owned by c) */
DRRT_API int64_t drrt_comm_broadcast_bytes(drrt_comm *c);     /* bytes delivered by ncclBroadcast so far */
DRRT_API int drrt_comm_insert(drrt_comm *c, int64_t n, const double *pos, int32_t *first_id);
DRRT_API int drrt_comm_obstacles_set(drrt_comm *c, int32_t n_obs, const int32_t *kind,
                   const uint8_t *used, const double *life_span, const double *origin_xy,
                   const double *radius, const int32_t *vert_off, const double *verts_xy);
DRRT_API int drrt_comm_range_batch_into(drrt_comm *c, int64_t nq, const double *q,
                   const double *radius, double radius_all, int64_t cap, int64_t *offsets,
                   int32_t *ids, double *dist, int64_t *total);
DRRT_API int drrt_comm_range_batch_into_f32(drrt_comm *c, int64_t nq, const double *q,
                   const double *radius, double radius_all, int64_t cap, int64_t *offsets,
                   int32_t *ids, float *dist, int64_t *total);
DRRT_API int drrt_comm_nearest_batch(drrt_comm *c, int64_t nq, const double *q, int32_t *ids,
                   double *dist);
DRRT_API int drrt_comm_knn_batch(drrt_comm *c, int64_t nq, const double *q, int k, int32_t *ids,
                   double *dist);
DRRT_API int drrt_comm_edgecheck_batch(drrt_comm *c, int64_t n, const double *start,
                   const double *end, int edge_kind, double robot_radius, double r_min,
                   int in_warmup, uint8_t *verdict, double *length);
DRRT_API int drrt_comm_edgecheck_ids_batch(drrt_comm *c, int64_t n, const int32_t *start_id,
                   const int32_t *end_id, int edge_kind, double robot_radius, double r_min,
                   int in_warmup, uint8_t *verdict, double *length);
DRRT_API int drrt_comm_pointcheck_batch(drrt_comm *c, int64_t n, const double *pts,
                   double robot_radius, double collision_distance, uint8_t *verdict);
DRRT_API int drrt_comm_lmc_write(drrt_comm *c, int64_t first, int64_t n, const double *values);
DRRT_API int drrt_comm_lmc_scatter(drrt_comm *c, int64_t n, const int32_t *ids, const double *values);
DRRT_API int drrt_comm_extend_batch(drrt_comm *c, int64_t nq, const double *q, const double *radius,
                   double radius_all, double robot_radius, double r_min, int in_warmup,
                   const double *lmc, int32_t *counts, int64_t *total, int32_t *best_parent,
                   double *best_cost);

#ifdef __cplusplus
}
#endif
#endif /* DRRT_GPU_H */

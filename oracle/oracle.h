/* oracle.h -- CPU restatement of the DRRT hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This directory is the parity oracle for drrt_b200.  Nothing that ships
 * (drrt_b200/, include/, the C-ABI library) may call, link or import it; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the CPU baseline.
 *
 * Every function restates one routine of cosa0481/DRRT and cites the
 * reference file:line it follows (paths relative to /root/reference).
 * All arithmetic is IEEE double in the reference's operation order, compiled
 * with -ffp-contract=off (the reference's x86-64 build has no FMA).
 *
 * PARITY PINNING: the reference has no tests, golden vectors or fixtures for
 * this path, and cannot be built as-is in this container
 * (needs Eigen, Bullet, Pangolin, SceneGraph).  oracle/_ref (see
 * oracle/Makefile, target `ref`) compiles the reference's own kdtree.cpp,
 * ghostpoint.cpp, distancefunctions.cpp, jlist.cpp from /root/reference
 * against a minimal stand-in for the Eigen subset they use; where that build
 * exists the restatement is checked against it (tests/test_oracle_vs_ref.py).
 * The Dubins/obstacle restatement (ref_dubins.c / ref_collide.c) is pinned
 * the same way when dubinsedge.cpp/obstacle.cpp compile; otherwise it is
 * "parity unpinned" beyond the hand-derived known-answer cases in tests/golden.
 */
#ifndef DRRT_ORACLE_H
#define DRRT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* include/DRRT/distancefunctions.h:22-25 */
#define REF_PI 3.1415926536
#define REF_INF 1000000000000.0
#define REF_MAXPATHNODES 1000

#define REF_MAXD 3      /* x, y, theta */
#define REF_MAXW 3

/* metric ids (the reference passes a function pointer, kdtree.h:21) */
#define REF_METRIC_R2S1 0   /* src/smalltest.cpp:167-175 */
#define REF_METRIC_EUCLID 1 /* src/thetastar.cpp:10-15 */

typedef struct ref_tree ref_tree;

double ref_metric(int metric, const double *a, const double *b);

/* ---- KD tree (src/kdtree.cpp) ---- */
ref_tree *ref_tree_create(int dims, int n_wraps, const int *wrap_dims,
                          const double *wrap_points, int metric);
void ref_tree_destroy(ref_tree *t);
int ref_tree_size(const ref_tree *t);
/* KDInsert kdtree.cpp:87-136, n nodes in order; returns id of the first */
int ref_tree_insert(ref_tree *t, int n, const double *pos);
/* structure dump for tests (arrays of length size): -1 = none */
void ref_tree_structure(const ref_tree *t, int *parent, int *left, int *right,
                        int *split);
const double *ref_tree_positions(const ref_tree *t);

/* KDFindWithinRange kdtree.cpp:794-820 (+ ghost loop).  Writes hits in
 * discovery order (the reference's JList holds them in reverse, jlist.cpp:54-69).
 * Returns the number of hits; writes at most cap. Thread-safe per `tid` slot. */
long ref_range(ref_tree *t, const double *q, double range, int *ids,
               double *dist, long cap);
/* batch helpers: counts, then CSR fill sorted by id.  nthreads<=0: all cores. */
void ref_range_batch_count(ref_tree *t, long nq, const double *q,
                           const double *range, int *counts, int nthreads);
void ref_range_batch_fill(ref_tree *t, long nq, const double *q,
                          const double *range, const long long *offsets,
                          int *ids, double *dist, int sort_by_id, int nthreads);
/* number of tree nodes whose distance was evaluated by the last ref_range on
 * this thread (instrumentation for DESIGN.md) */
long ref_range_last_visits(void);

/* KDFindNearest kdtree.cpp:264-307 */
void ref_nearest(ref_tree *t, const double *q, int *id, double *dist);
void ref_nearest_batch(ref_tree *t, long nq, const double *q, int *ids,
                       double *dist, int nthreads);

/* kNN by intent: k smallest metric(q,n), ties by smaller id,
 * ascending (dist,id).  Brute force. */
void ref_knn_batch(ref_tree *t, long nq, const double *q, int k, int *ids,
                   double *dist, int nthreads);

/* independent brute-force range (cross-check of the tree walk): the set
 * { n : some probe p in {q, ghosts} has metric(p,n) < range }  plus root at
 * <=.  Sorted by id.  Returns count. */
long ref_range_brute(ref_tree *t, const double *q, double range, int *ids,
                     double *dist, long cap);

/* ghost iterator (src/ghostpoint.cpp:4-75) exposed for tests: writes up to
 * max_ghosts ghost points (dims doubles each), returns how many */
int ref_ghosts(ref_tree *t, const double *q, double best_dist, double *out,
               int max_ghosts);

/* ---- 2-D geometry (src/distancefunctions.cpp) ---- */
double ref_dist_sqrd_point_segment(const double *pt, const double *s,
                                   const double *e);                /* :50-74 */
double ref_segment_dist_sqrd(const double *PA, const double *PB,
                             const double *QA, const double *QB);   /* :94-164 */
double ref_right_turn_dist(const double *a, const double *b, const double *c,
                           double r);                               /* :24-34 */
double ref_left_turn_dist(const double *a, const double *b, const double *c,
                          double r);                                /* :36-46 */
int ref_point_in_polygon(const double *pt, const double *poly, int nv); /* :166-205 */
double ref_dist_to_polygon_sqrd(const double *pt, const double *poly, int nv); /* :77-92 */

/* ---- obstacles (include/DRRT/obstacle.h:16-211) ---- */
typedef struct {
  int n;
  const int *kind;          /* n */
  const unsigned char *used;/* n   obstacle_used_ */
  const double *life_span;  /* n */
  const double *origin;     /* 2n  origin_(0..1), world frame */
  const double *radius;     /* n   radius_ */
  const int *vert_off;      /* n+1 */
  const double *verts;      /* 2*vert_off[n], world frame rows of shape_ */
} ref_obstacles;

/* ExplicitEdgeCheck2D src/obstacle.cpp:756-804 (kinds 0-5), one obstacle.
 * If clearance!=NULL it is lowered to the smallest |margin| seen:
 *   |sqrt(dist_sqrd)-(radius+O.radius)| and |sqrt(seg_dist_sqrd)-radius|   */
/* time-parametrised obstacles (kinds 6 / 7): path rows (dx, dy, time) relative to origin */
typedef struct {
  int n;
  const int *kind;
  const unsigned char *used;
  const double *life_span;
  const double *origin; /* n x 2 */
  const double *radius;
  const int *path_off;  /* n + 1 */
  const double *path;   /* path_off[n] x 3 */
} ref_timed_obstacles;
int ref_edge_check_2d_timed(const ref_timed_obstacles *O, int o, const double *start, const double *end,
                            double radius, double *clearance);
void ref_edge_check_timed_batch(const ref_timed_obstacles *O, long n, const double *start, const double *end,
                                double radius, unsigned char *verdict, double *clearance);

int ref_edge_check_2d(const ref_obstacles *O, int o, const double *p0,
                      const double *p1, double radius, double *clearance);

/* ---- Dubins edge (src/dubinsedge.cpp) ---- */
#define REF_TRAJ_MAX 200
typedef struct {
  char type[4];     /* "rsl","rsr","rlr","lsr","lsl","lrl","0s0","xxx" */
  double dist;      /* dist_ after CalculateTrajectory (:808) */
  int npts;
  double x[REF_TRAJ_MAX], y[REF_TRAJ_MAX];
} ref_traj;

/* Edge::NewEdge :12-21 + DubinsEdge::CalculateTrajectory :172-830
 * (space_has_time_ == false) */
void ref_dubins_trajectory(const double *start, const double *end, double r_min,
                           int metric, ref_traj *out);
/* LineCheck polyline src/obstacle.cpp:1092-1133 (51 points) */
void ref_line_trajectory(const double *n1, const double *n2, ref_traj *out);

/* ExplicitEdgeCheck(C,edge) obstacle.cpp:879-923 + DubinsEdge::ExplicitEdgeCheck
 * dubinsedge.cpp:850-898 with the analytic test of :881-884.
 * kind: 0 = Dubins, 1 = straight (LineCheck).  verdict 1 = collides.
 * clearance (optional, n) = smallest boundary margin over the WHOLE edge (no
 * early exit) for the 1e-6 band rule; length (optional) = dist_.            */
void ref_edge_check_batch(const ref_obstacles *O, long n, const double *start,
                          const double *end, int edge_kind, double robot_radius,
                          double r_min, int metric, unsigned char *verdict,
                          double *length, double *clearance, int nthreads);

/* QuickCheck + ExplicitPointCheck obstacle.cpp:925-1090 (node validity) */
void ref_point_check_batch(const ref_obstacles *O, long n, const double *pts,
                           double robot_radius, double collision_distance,
                           int metric, unsigned char *verdict, int nthreads);

int ref_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif

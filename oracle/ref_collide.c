/* ref_collide.c -- TEST ORACLE (see oracle.h).  Restatement of the analytic
 * edge-vs-obstacle path of cosa0481/DRRT:
 *   ExplicitEdgeCheck2D      src/obstacle.cpp:756-804 (kinds 0..5)
 *   ExplicitEdgeCheck(C,e)   src/obstacle.cpp:879-923 (obstacle loop, first hit)
 *   DubinsEdge::ExplicitEdgeCheck src/dubinsedge.cpp:850-898 with the analytic
 *       call shown (commented) at :881-884 -- per polyline segment,
 *       radius = robot_radius_.  The active call there is Bullet
 *       (DetectBulletCollision), absent from the reference tree and
 *       unpinned; SURVEY 8c fixes the oracle to the in-tree analytic test.
 *   QuickCheck2D / ExplicitPointCheck2D / ExplicitPointCheck :925-1090
 * Kinds 6/7 (time-parametrised, :805-875): ref_edge_check_2d_timed below -- the
 * planner itself never reaches them (space_has_time_ == false, smalltest.cpp:200).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void note_margin(double *clearance, double m) {
  if (!clearance) return;
  m = fabs(m);
  if (m < *clearance) *clearance = m;
}

/* ExplicitEdgeCheck2D obstacle.cpp:756-804 */
int ref_edge_check_2d(const ref_obstacles *O, int o, const double *p0,
                      const double *p1, double radius, double *clearance) {
  if (!O->used[o] || O->life_span[o] <= 0) return 0; /* :761 */
  int kind = O->kind[o];
  if (1 <= kind && kind <= 5) { /* :765-779 */
    double dist_sqrd = ref_dist_sqrd_point_segment(O->origin + 2 * o, p0, p1);
    double reach = radius + O->radius[o];
    note_margin(clearance, sqrt(dist_sqrd) - reach);
    if (dist_sqrd > reach * reach) return 0;
  }
  if (kind == 1) return 1; /* :781 */
  else if (kind == 2) return 0;
  else if (kind == 3 || kind == 5) { /* :783-804 */
    int nv = O->vert_off[o + 1] - O->vert_off[o];
    const double *poly = O->verts + 2 * (size_t)O->vert_off[o];
    if (nv < 2) return 0;
    const double *A = poly + 2 * (nv - 1);
    int hit = 0;
    for (int i = 0; i < nv; i++) {
      const double *B = poly + 2 * i;
      double seg_dist_sqrd = ref_segment_dist_sqrd(p0, p1, A, B);
      note_margin(clearance, sqrt(seg_dist_sqrd) - radius);
      if (seg_dist_sqrd < radius * radius) {
        if (kind == 5) { /* :800 `if(O->kind_ == 5);` never reports */
        } else {
          if (!clearance) return 1;
          hit = 1;
        }
      }
      A = B;
    }
    return hit;
  }
  return 0;
}

static int edge_vs_all(const ref_obstacles *O, const ref_traj *T, double radius,
                       double *clearance) {
  int hit = 0;
  for (int o = 0; o < O->n; o++) {          /* obstacle.cpp:897-916 */
    for (int i = 1; i < T->npts; i++) {     /* dubinsedge.cpp:875-893 */
      double p0[2] = {T->x[i - 1], T->y[i - 1]};
      double p1[2] = {T->x[i], T->y[i]};
      if (ref_edge_check_2d(O, o, p0, p1, radius, clearance)) {
        if (!clearance) return 1;
        hit = 1;
      }
    }
  }
  return hit;
}

void ref_edge_check_batch(const ref_obstacles *O, long n, const double *start,
                          const double *end, int edge_kind, double robot_radius,
                          double r_min, int metric, unsigned char *verdict,
                          double *length, double *clearance, int nthreads) {
#ifdef _OPENMP
  omp_set_num_threads(nthreads > 0 ? nthreads : omp_get_num_procs());
#endif
#pragma omp parallel for schedule(dynamic, 64)
  for (long e = 0; e < n; e++) {
    ref_traj T;
    const double *s = start + 3 * e, *g = end + 3 * e;
    if (edge_kind == 0) {
      ref_dubins_trajectory(s, g, r_min, metric, &T);
    } else {
      ref_line_trajectory(s, g, &T);
      /* LineCheck :1098-1105: both headings set to atan2(dy,dx), then NewEdge */
      double th = atan2(g[1] - s[1], g[0] - s[0]);
      double a[3] = {s[0], s[1], th}, b[3] = {g[0], g[1], th};
      T.dist = ref_metric(metric, a, b);
    }
    double c = REF_INF;
    int v = edge_vs_all(O, &T, robot_radius, clearance ? &c : NULL);
    verdict[e] = (unsigned char)v;
    if (length) length[e] = T.dist;
    if (clearance) clearance[e] = c;
  }
}

/* QuickCheck2D obstacle.cpp:925-965 (kinds 0..5) */
static int quick_check_2d(const ref_obstacles *O, int o, const double *pt) {
  if (!O->used[o] || O->life_span[o] <= 0) return 0;
  int kind = O->kind[o];
  if (1 <= kind && kind <= 5) {
    double dx = O->origin[2 * o] - pt[0], dy = O->origin[2 * o + 1] - pt[1];
    if (sqrt(dx * dx + dy * dy) > O->radius[o]) return 0; /* EuclideanDistance2D */
  }
  if (kind == 1) return 1;
  else if (kind == 2 || kind == 4) return 0;
  else if (kind == 3) {
    int nv = O->vert_off[o + 1] - O->vert_off[o];
    if (ref_point_in_polygon(pt, O->verts + 2 * (size_t)O->vert_off[o], nv)) return 1;
  }
  return 0;
}

/* ExplicitPointCheck2D obstacle.cpp:984-1060 (kinds 0..5).  The reference
 * copies the query's third coordinate into O->origin_(2) first (:1000) so the
 * metric's theta term vanishes. */
static int explicit_point_check_2d(const ref_obstacles *O, int o, const double *pt,
                                   double radius, double min_distance, int metric) {
  double this_distance = REF_INF;
  if (!O->used[o] || O->life_span[o] <= 0) return 0;
  int kind = O->kind[o];
  int nv = O->vert_off[o + 1] - O->vert_off[o];
  const double *poly = O->verts + 2 * (size_t)O->vert_off[o];
  if (1 <= kind && kind <= 5) {
    double og[3] = {O->origin[2 * o], O->origin[2 * o + 1], pt[2]};
    this_distance = ref_metric(metric, og, pt) - radius;
    if (this_distance - O->radius[o] > min_distance) return 0;
  }
  if (kind == 1) {
    this_distance = this_distance - O->radius[o];
    if (this_distance < 0.0) return 1;
  } else if (kind == 2) return 0;
  else if (kind == 3) {
    if (ref_point_in_polygon(pt, poly, nv)) return 1;
    this_distance = sqrt(ref_dist_to_polygon_sqrd(pt, poly, nv)) - radius;
    if (this_distance < 0.0) return 1;
  }
  return 0;
}

/* ExplicitPointCheck obstacle.cpp:1062-1085 */
void ref_point_check_batch(const ref_obstacles *O, long n, const double *pts,
                           double robot_radius, double collision_distance,
                           int metric, unsigned char *verdict, int nthreads) {
#ifdef _OPENMP
  omp_set_num_threads(nthreads > 0 ? nthreads : omp_get_num_procs());
#endif
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; i++) {
    const double *pt = pts + 3 * i;
    int hit = 0;
    for (int o = 0; o < O->n && !hit; o++) hit = quick_check_2d(O, o, pt);
    for (int o = 0; o < O->n && !hit; o++)
      hit = explicit_point_check_2d(O, o, pt, robot_radius, collision_distance, metric);
    verdict[i] = (unsigned char)hit;
  }
}


/* ---- ExplicitEdgeCheck2D, kinds 6 / 7 (obstacle.cpp:805-875) ------------------------------
 * start / end are (x, y, TIME); the obstacle moves along path rows (dx, dy, time) relative to
 * origin_.  FindIndexBeforeTime: src/drrt.cpp:178-189. */
static int find_index_before_time(const double *path, int rows, double t) {
  if (rows < 1) return -1;
  int i = -1;
  while (i + 1 < rows && path[3 * (i + 1) + 2] < t) i += 1;
  return i;
}

int ref_edge_check_2d_timed(const ref_timed_obstacles *O, int o, const double *start, const double *end,
                            double radius, double *clearance) {
  if (!O->used[o] || O->life_span[o] <= 0) return 0; /* :761 */
  if (O->kind[o] != 6 && O->kind[o] != 7) return 0;
  const double *path = O->path + 3 * (size_t)O->path_off[o];
  const int rows = O->path_off[o + 1] - O->path_off[o];
  const double *early = start, *late = end; /* :813-819 */
  if (!(start[2] < end[2])) { early = end; late = start; }
  int first = find_index_before_time(path, rows, early[2]);
  if (first < 0) first = 0;
  int last = find_index_before_time(path, rows, late[2]);
  if (last > rows - 1) last = rows - 1;
  if (last <= first) return 0; /* :825 */
  int hit = 0;
  for (int i_start = first; i_start < last; i_start++) {
    const int i_end = i_start + 1;
    double x_1 = early[0], y_1 = early[1], t_1 = early[2];
    double x_2 = path[3 * i_start] + O->origin[2 * o];
    double y_2 = path[3 * i_start + 1] + O->origin[2 * o + 1];
    double t_2 = path[3 * i_start + 2];
    double m_x1 = (late[0] - x_1) / (late[2] - t_1);
    double m_y1 = (late[1] - y_1) / (late[2] - t_1);
    double m_x2 = (path[3 * i_end] + O->origin[2 * o] - x_2) / (path[3 * i_end + 2] - t_2);
    double m_y2 = (path[3 * i_end + 1] + O->origin[2 * o + 1] - y_2) / (path[3 * i_end + 2] - t_2);
    double t_c = (m_x1 * m_x1 * t_1 + m_x2 * (m_x2 * t_2 + x_1 - x_2) - m_x1 * (m_x2 * (t_1 + t_2) + x_1 - x_2) +
                  (m_y1 - m_y2) * (m_y1 * t_1 - m_y2 * t_2 - y_1 + y_2)) /
                 ((m_x1 - m_x2) * (m_x1 - m_x2) + (m_y1 - m_y2) * (m_y1 - m_y2));
    double lo = (t_1 < t_2) ? t_2 : t_1;                                     /* std::max */
    double hi = (path[3 * i_end + 2] < late[2]) ? path[3 * i_end + 2] : late[2]; /* std::min */
    if (t_c < lo) t_c = lo;
    else if (t_c > hi) t_c = hi;
    double r_x = m_x1 * (t_c - t_1) + x_1, r_y = m_y1 * (t_c - t_1) + y_1;
    double o_x = m_x2 * (t_c - t_2) + x_2, o_y = m_y2 * (t_c - t_2) + y_2;
    double d2 = (r_x - o_x) * (r_x - o_x) + (r_y - o_y) * (r_y - o_y);
    double reach = O->radius[o] + radius;
    if (clearance && d2 == d2) note_margin(clearance, sqrt(d2) - reach);
    if (d2 < reach * reach) {
      if (!clearance) return 1;
      hit = 1;
    }
  }
  return hit;
}

void ref_edge_check_timed_batch(const ref_timed_obstacles *O, long n, const double *start, const double *end,
                                double radius, unsigned char *verdict, double *clearance) {
  for (long i = 0; i < n; i++) {
    int hit = 0;
    double c = 1e300;
    for (int o = 0; o < O->n && (!hit || clearance); o++)
      if (ref_edge_check_2d_timed(O, o, start + 3 * i, end + 3 * i, radius, clearance ? &c : NULL)) hit = 1;
    verdict[i] = (unsigned char)hit;
    if (clearance) clearance[i] = c;
  }
}

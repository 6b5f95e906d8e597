/* ref_kdtree.c -- TEST ORACLE (see oracle.h).  CPU restatement of
 * src/kdtree.cpp and src/ghostpoint.cpp of cosa0481/DRRT: the unbalanced,
 * insertion-ordered KD tree, its signed-hyperplane range / nearest walks and
 * the ghost-point iterator for wrapped dimensions.  Pointers are replaced by
 * integer ids (id = insertion order); everything else keeps the reference's
 * control flow and floating-point operation order.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

struct ref_tree {
  int dims, n_wraps, metric;
  int wraps[REF_MAXW];
  double wrap_points[REF_MAXW];
  int n, cap;
  double *pos;  /* n*dims */
  int *parent, *left, *right;
  unsigned char *split;
  /* per-thread "in_heap_" flags (kdtreenode.h:24); the reference has one flag
   * per node and is therefore not re-entrant */
  int n_flag_sets;
  unsigned char **flags;
  int *flags_cap;
};

int ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}

static int tid_now(void) {
#ifdef _OPENMP
  return omp_get_thread_num();
#else
  return 0;
#endif
}

/* src/smalltest.cpp:167-175 and src/thetastar.cpp:10-15.
 * std::min(a,b) = (b<a)?b:a ; std::max(a,b) = (a<b)?b:a ; pow(x,2) == x*x. */
double ref_metric(int metric, const double *a, const double *b) {
  if (metric == REF_METRIC_EUCLID) {
    double t0 = a[0] - b[0], t1 = a[1] - b[1], t2 = a[2] - b[2];
    t0 = t0 * t0;
    t1 = t1 * t1;
    t2 = t2 * t2;
    return sqrt((t0 + t1) + t2);
  }
  double t0 = a[0] - b[0], t1 = a[1] - b[1];
  t0 = t0 * t0;
  t1 = t1 * t1;
  double s = t0 + t1;
  double ad = fabs(a[2] - b[2]);
  double mn = (b[2] < a[2]) ? b[2] : a[2];
  double mx = (a[2] < b[2]) ? b[2] : a[2];
  double wr = mn + 2.0 * REF_PI - mx;
  double m = (wr < ad) ? wr : ad;
  return sqrt(s + m * m);
}

ref_tree *ref_tree_create(int dims, int n_wraps, const int *wrap_dims,
                          const double *wrap_points, int metric) {
  if (dims != 3 || n_wraps < 0 || n_wraps > REF_MAXW) return NULL;
  ref_tree *t = (ref_tree *)calloc(1, sizeof(ref_tree));
  t->dims = dims;
  t->n_wraps = n_wraps;
  t->metric = metric;
  for (int i = 0; i < n_wraps; i++) {
    t->wraps[i] = wrap_dims[i];
    t->wrap_points[i] = wrap_points[i];
  }
#ifdef _OPENMP
  t->n_flag_sets = omp_get_num_procs(); /* the most threads a batch call may use */
#else
  t->n_flag_sets = 1;
#endif
  if (t->n_flag_sets < 1) t->n_flag_sets = 1;
  t->flags = (unsigned char **)calloc(t->n_flag_sets, sizeof(unsigned char *));
  t->flags_cap = (int *)calloc(t->n_flag_sets, sizeof(int));
  return t;
}

void ref_tree_destroy(ref_tree *t) {
  if (!t) return;
  for (int i = 0; i < t->n_flag_sets; i++) free(t->flags[i]);
  free(t->flags);
  free(t->flags_cap);
  free(t->pos);
  free(t->parent);
  free(t->left);
  free(t->right);
  free(t->split);
  free(t);
}

int ref_tree_size(const ref_tree *t) { return t->n; }
const double *ref_tree_positions(const ref_tree *t) { return t->pos; }

static void grow(ref_tree *t, int need) {
  if (need <= t->cap) return;
  int cap = t->cap ? t->cap : 1024;
  while (cap < need) cap *= 2;
  t->pos = (double *)realloc(t->pos, sizeof(double) * cap * t->dims);
  t->parent = (int *)realloc(t->parent, sizeof(int) * cap);
  t->left = (int *)realloc(t->left, sizeof(int) * cap);
  t->right = (int *)realloc(t->right, sizeof(int) * cap);
  t->split = (unsigned char *)realloc(t->split, cap);
  t->cap = cap;
}

static unsigned char *flags_for(ref_tree *t, int tid) {
  if (tid >= t->n_flag_sets) tid = 0;
  if (t->flags_cap[tid] < t->n) {
    int cap = t->cap;
    t->flags[tid] = (unsigned char *)realloc(t->flags[tid], cap);
    memset(t->flags[tid] + t->flags_cap[tid], 0, cap - t->flags_cap[tid]);
    t->flags_cap[tid] = cap;
  }
  return t->flags[tid];
}

/* KDTree::KDInsert  src/kdtree.cpp:87-136 */
int ref_tree_insert(ref_tree *t, int n, const double *pos) {
  int first = t->n;
  grow(t, t->n + n);
  const int d = t->dims;
  for (int k = 0; k < n; k++) {
    int id = t->n;
    memcpy(t->pos + (size_t)id * d, pos + (size_t)k * d, sizeof(double) * d);
    t->parent[id] = t->left[id] = t->right[id] = -1;
    if (t->n == 0) {            /* :94-99 */
      t->split[id] = 0;
      t->n = 1;
      continue;
    }
    int parent = 0;
    for (;;) {                  /* :103-127 */
      int s = t->split[parent];
      if (t->pos[(size_t)id * d + s] < t->pos[(size_t)parent * d + s]) {
        if (t->left[parent] < 0) { t->left[parent] = id; break; }
        parent = t->left[parent];
      } else {
        if (t->right[parent] < 0) { t->right[parent] = id; break; }
        parent = t->right[parent];
      }
    }
    t->parent[id] = parent;     /* :129-132 */
    t->split[id] = (t->split[parent] == d - 1) ? 0 : (unsigned char)(t->split[parent] + 1);
    t->n += 1;
  }
  return first;
}

void ref_tree_structure(const ref_tree *t, int *parent, int *left, int *right,
                        int *split) {
  for (int i = 0; i < t->n; i++) {
    parent[i] = t->parent[i];
    left[i] = t->left[i];
    right[i] = t->right[i];
    split[i] = t->split[i];
  }
}

/* ---- ghost iterator  src/ghostpoint.cpp:4-75 ---- */
typedef struct {
  const ref_tree *t;
  double q[REF_MAXD];
  int flags[REF_MAXW]; /* zero-initialised: documented intent, SURVEY F6 */
  int depth;
  double ghost[REF_MAXD];
  double closest[REF_MAXD];
} ghost_iter;

static void ghost_init(ghost_iter *G, const ref_tree *t, const double *q) {
  G->t = t;
  for (int i = 0; i < t->dims; i++) G->q[i] = G->ghost[i] = G->closest[i] = q[i];
  for (int i = 0; i < REF_MAXW; i++) G->flags[i] = 0;
  G->depth = t->n_wraps;
}

/* returns 0 when the reference returns its size-0 sentinel (:28-33) */
static int ghost_next(ghost_iter *G, double best_dist, double *out) {
  const ref_tree *t = G->t;
  for (;;) {
    while (G->depth > 0 && G->flags[G->depth - 1] != 0) G->depth -= 1; /* :23-26 */
    if (G->depth == 0) return 0;
    G->flags[G->depth - 1] = 1;                                         /* :36 */
    double wrap_value = t->wrap_points[G->depth - 1];
    int wrap_dim = t->wraps[G->depth - 1];
    double dim_val = G->q[wrap_dim];
    double dim_closest = 0.0;
    if (G->q[wrap_dim] < wrap_value / 2.0) {  /* :43-52 */
      dim_val += wrap_value;
      dim_closest += wrap_value;
    } else {
      dim_val -= wrap_value;
    }
    G->ghost[wrap_dim] = dim_val;
    G->closest[wrap_dim] = dim_closest;
    while (G->depth < t->n_wraps) {           /* :57-65 */
      G->depth += 1;
      G->flags[G->depth - 1] = 0;
      int wd = t->wraps[G->depth - 1];
      G->ghost[wd] = G->q[wd];
      G->closest[wd] = G->ghost[wd];
    }
    if (ref_metric(t->metric, G->closest, G->ghost) > best_dist) continue; /* :69-72 */
    for (int i = 0; i < t->dims; i++) out[i] = G->ghost[i];
    return 1;
  }
}

/* callers stop on `thisGhostPoint.isZero(0)` (kdtree.cpp:285, 814): a real
 * ghost whose coordinates are all exactly 0 also ends the loop */
static int ghost_is_sentinel(const ref_tree *t, const double *g) {
  for (int i = 0; i < t->dims; i++)
    if (g[i] != 0.0) return 0;
  return 1;
}

int ref_ghosts(ref_tree *t, const double *q, double best_dist, double *out,
               int max_ghosts) {
  ghost_iter G;
  ghost_init(&G, t, q);
  int n = 0;
  double g[REF_MAXD];
  while (n < max_ghosts && ghost_next(&G, best_dist, g)) {
    if (ghost_is_sentinel(t, g)) break;
    for (int i = 0; i < t->dims; i++) out[n * t->dims + i] = g[i];
    n++;
  }
  return n;
}

/* ---- range ---- */
typedef struct {
  ref_tree *t;
  unsigned char *in_list;
  int *ids;
  double *dist;
  long n, cap, visits;
} range_ctx;

static _Thread_local long g_last_visits;
long ref_range_last_visits(void) { return g_last_visits; }

/* AddToRangeList kdtree.cpp:670-682 */
static void add_to_range_list(range_ctx *c, int node, double key) {
  if (c->in_list[node]) return;
  c->in_list[node] = 1;
  if (c->n < c->cap) {
    c->ids[c->n] = node;
    if (c->dist) c->dist[c->n] = key;
  }
  c->n++;
}

/* KDFindWithinRangeInSubtree kdtree.cpp:703-792 */
static void range_in_subtree(range_ctx *c, int root, double range, const double *q) {
  ref_tree *t = c->t;
  const int d = t->dims;
  const double *P = t->pos;
  int parent = root;
  for (;;) { /* :710-730 */
    int s = t->split[parent];
    if (q[s] < P[(size_t)parent * d + s]) {
      if (t->left[parent] < 0) break;
      parent = t->left[parent];
    } else {
      if (t->right[parent] < 0) break;
      parent = t->right[parent];
    }
  }
  double new_dist = ref_metric(t->metric, q, P + (size_t)parent * d); /* :732-735 */
  c->visits++;
  if (new_dist < range) add_to_range_list(c, parent, new_dist);

  for (;;) { /* :738-791 */
    int s = t->split[parent];
    double hyper = q[s] - P[(size_t)parent * d + s]; /* :744-745, signed */
    if (hyper > range) {
      if (parent == root) return;
      parent = t->parent[parent];
      continue;
    }
    if (!c->in_list[parent]) { /* :763-768 */
      new_dist = ref_metric(t->metric, q, P + (size_t)parent * d);
      c->visits++;
      if (new_dist < range) add_to_range_list(c, parent, new_dist);
    }
    if (q[s] < P[(size_t)parent * d + s] && t->right[parent] >= 0) { /* :771-776 */
      range_in_subtree(c, t->right[parent], range, q);
    } else if (P[(size_t)parent * d + s] <= q[s] && t->left[parent] >= 0) { /* :777-783 */
      range_in_subtree(c, t->left[parent], range, q);
    }
    if (parent == root) return;
    parent = t->parent[parent];
  }
}

/* KDFindWithinRange kdtree.cpp:794-820 */
static long range_one(ref_tree *t, int tid, const double *q, double range,
                      int *ids, double *dist, long cap) {
  if (t->n == 0) return 0;
  range_ctx c;
  c.t = t;
  c.in_list = flags_for(t, tid);
  c.ids = ids;
  c.dist = dist;
  c.n = 0;
  c.cap = cap;
  c.visits = 0;
  double dist_to_root = ref_metric(t->metric, q, t->pos); /* :800-802 */
  c.visits++;
  if (dist_to_root <= range) add_to_range_list(&c, 0, dist_to_root);
  range_in_subtree(&c, 0, range, q); /* :805 */
  if (t->n_wraps > 0) {              /* :807-819 */
    ghost_iter G;
    ghost_init(&G, t, q);
    double g[REF_MAXD];
    while (ghost_next(&G, range, g)) {
      if (ghost_is_sentinel(t, g)) break;
      range_in_subtree(&c, 0, range, g);
    }
  }
  g_last_visits = c.visits;
  /* EmptyRangeList kdtree.cpp:692-701: clear the flags of listed nodes.  When
   * the caller's buffer was too small we cannot enumerate them, so wipe all. */
  if (c.n <= cap) {
    for (long i = 0; i < c.n; i++) c.in_list[ids[i]] = 0;
  } else {
    memset(c.in_list, 0, t->n);
  }
  return c.n;
}

long ref_range(ref_tree *t, const double *q, double range, int *ids,
               double *dist, long cap) {
  return range_one(t, tid_now(), q, range, ids, dist, cap);
}

static void set_threads(int nthreads) {
#ifdef _OPENMP
  int np = omp_get_num_procs();
  if (nthreads <= 0 || nthreads > np) nthreads = np; /* one flag set per possible thread */
  omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
}

void ref_range_batch_count(ref_tree *t, long nq, const double *q,
                           const double *range, int *counts, int nthreads) {
  set_threads(nthreads);
  for (int i = 0; i < t->n_flag_sets; i++) flags_for(t, i);
#pragma omp parallel
  {
    int tid = tid_now();
    long cap = 1024;
    int *ids = (int *)malloc(sizeof(int) * cap);
#pragma omp for schedule(dynamic, 16)
    for (long i = 0; i < nq; i++) {
      long n = range_one(t, tid, q + i * t->dims, range[i], ids, NULL, cap);
      if (n > cap) {
        /* flags were wiped by range_one; grow for the next query */
        cap = n * 2;
        ids = (int *)realloc(ids, sizeof(int) * cap);
      }
      counts[i] = (int)n;
    }
    free(ids);
  }
}

typedef struct { int id; double d; } hit_t;
static int hit_cmp(const void *a, const void *b) {
  const hit_t *x = (const hit_t *)a, *y = (const hit_t *)b;
  return (x->id > y->id) - (x->id < y->id);
}

void ref_range_batch_fill(ref_tree *t, long nq, const double *q,
                          const double *range, const long long *offsets,
                          int *ids, double *dist, int sort_by_id, int nthreads) {
  set_threads(nthreads);
  for (int i = 0; i < t->n_flag_sets; i++) flags_for(t, i);
#pragma omp parallel
  {
    int tid = tid_now();
    hit_t *tmp = NULL;
    long tmp_cap = 0;
#pragma omp for schedule(dynamic, 16)
    for (long i = 0; i < nq; i++) {
      long cap = (long)(offsets[i + 1] - offsets[i]);
      int *oi = ids + offsets[i];
      double *od = dist + offsets[i];
      long n = range_one(t, tid, q + i * t->dims, range[i], oi, od, cap);
      if (n > cap) n = cap;
      if (sort_by_id && n > 1) {
        if (n > tmp_cap) {
          tmp_cap = n * 2;
          tmp = (hit_t *)realloc(tmp, sizeof(hit_t) * tmp_cap);
        }
        for (long k = 0; k < n; k++) { tmp[k].id = oi[k]; tmp[k].d = od[k]; }
        qsort(tmp, n, sizeof(hit_t), hit_cmp);
        for (long k = 0; k < n; k++) { oi[k] = tmp[k].id; od[k] = tmp[k].d; }
      }
    }
    free(tmp);
  }
}

/* ---- nearest ---- */
/* KDFindNearestInSubtree kdtree.cpp:140-262 */
static void nearest_in_subtree(ref_tree *t, int *nearest, double *nearest_dist,
                               int root, const double *q, int suggested,
                               double suggested_dist) {
  const int d = t->dims;
  const double *P = t->pos;
  int parent = root;
  int cur = suggested;
  double cur_dist = suggested_dist;
  for (;;) { /* :152-171 */
    int s = t->split[parent];
    if (q[s] < P[(size_t)parent * d + s]) {
      if (t->left[parent] < 0) break;
      parent = t->left[parent];
    } else {
      if (t->right[parent] < 0) break;
      parent = t->right[parent];
    }
  }
  double new_dist = ref_metric(t->metric, q, P + (size_t)parent * d); /* :173-178 */
  if (new_dist < cur_dist) { cur = parent; cur_dist = new_dist; }
  for (;;) { /* :182-261 */
    int s = t->split[parent];
    double hyper = q[s] - P[(size_t)parent * d + s];
    if (hyper > cur_dist) { /* :188-201 */
      if (parent == root) { *nearest = cur; *nearest_dist = cur_dist; return; }
      parent = t->parent[parent];
      continue;
    }
    if (cur != parent) { /* :208-214 */
      new_dist = ref_metric(t->metric, q, P + (size_t)parent * d);
      if (new_dist < cur_dist) { cur = parent; cur_dist = new_dist; }
    }
    if (q[s] < P[(size_t)parent * d + s] && t->right[parent] >= 0) { /* :217-233 */
      int rn; double rd;
      nearest_in_subtree(t, &rn, &rd, t->right[parent], q, cur, cur_dist);
      if (rd < cur_dist) { cur = rn; cur_dist = rd; }
    } else if (P[(size_t)parent * d + s] <= q[s] && t->left[parent] >= 0) { /* :234-251 */
      int ln; double ld;
      nearest_in_subtree(t, &ln, &ld, t->left[parent], q, cur, cur_dist);
      if (ld < cur_dist) { cur = ln; cur_dist = ld; }
    }
    if (parent == root) { *nearest = cur; *nearest_dist = cur_dist; return; } /* :254-259 */
    parent = t->parent[parent];
  }
}

/* KDFindNearest kdtree.cpp:264-307 */
void ref_nearest(ref_tree *t, const double *q, int *id, double *dist) {
  if (t->n == 0) { *id = -1; *dist = REF_INF; return; }
  double dist_to_root = ref_metric(t->metric, q, t->pos);
  int ln; double ld;
  nearest_in_subtree(t, &ln, &ld, 0, q, 0, dist_to_root);
  if (t->n_wraps > 0) {
    ghost_iter G;
    ghost_init(&G, t, q);
    double g[REF_MAXD];
    while (ghost_next(&G, ld, g)) {
      if (ghost_is_sentinel(t, g)) break;
      double dg = ref_metric(t->metric, g, t->pos);
      int tn; double td;
      nearest_in_subtree(t, &tn, &td, 0, g, 0, dg);
      if (td < ld) { ln = tn; ld = td; }
    }
  }
  *id = ln;
  *dist = ld;
}

void ref_nearest_batch(ref_tree *t, long nq, const double *q, int *ids,
                       double *dist, int nthreads) {
  set_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 16)
  for (long i = 0; i < nq; i++) ref_nearest(t, q + i * t->dims, ids + i, dist + i);
}

/* ---- kNN by intent (no working reference implementation: SURVEY F7) ---- */
void ref_knn_batch(ref_tree *t, long nq, const double *q, int k, int *ids,
                   double *dist, int nthreads) {
  set_threads(nthreads);
  const int d = t->dims;
#pragma omp parallel for schedule(dynamic, 4)
  for (long i = 0; i < nq; i++) {
    int *oi = ids + i * k;
    double *od = dist + i * k;
    int m = 0;
    for (int n = 0; n < t->n; n++) {
      double dd = ref_metric(t->metric, q + i * d, t->pos + (size_t)n * d);
      if (m == k && !(dd < od[k - 1])) continue; /* ties keep the smaller id */
      int p = (m < k) ? m : k - 1;
      while (p > 0 && od[p - 1] > dd) { od[p] = od[p - 1]; oi[p] = oi[p - 1]; p--; }
      od[p] = dd;
      oi[p] = n;
      if (m < k) m++;
    }
    for (int p = m; p < k; p++) { oi[p] = -1; od[p] = REF_INF; }
  }
}

/* ---- independent brute-force range ---- */
long ref_range_brute(ref_tree *t, const double *q, double range, int *ids,
                     double *dist, long cap) {
  double ghosts[8 * REF_MAXD];
  int ng = (t->n_wraps > 0) ? ref_ghosts(t, q, range, ghosts, 7) : 0;
  const int d = t->dims;
  long n = 0;
  for (int i = 0; i < t->n; i++) {
    const double *p = t->pos + (size_t)i * d;
    double dd = ref_metric(t->metric, q, p);
    int hit = (i == 0) ? (dd <= range) : (dd < range);
    for (int g = 0; g < ng && !hit; g++) {
      dd = ref_metric(t->metric, ghosts + g * d, p);
      hit = dd < range;
    }
    if (hit) {
      if (n < cap) { ids[n] = i; if (dist) dist[n] = dd; }
      n++;
    }
  }
  return n;
}

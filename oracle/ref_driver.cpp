// ref_driver.cpp -- TEST INFRASTRUCTURE.  Command-line harness around the
// REFERENCE'S OWN sources, compiled where they lie under /root/reference
// (oracle/Makefile target `ref` -> oracle/_ref/drrt_ref; never committed, never
// shipped).  It lets tests/test_oracle_vs_ref.py check the plain-C oracle
// restatement against the reference's compiled kdtree.cpp, ghostpoint.cpp,
// jlist.cpp, heap.cpp, distancefunctions.cpp, dubinsedge.cpp and obstacle.cpp.
//
// Eigen and Bullet are not in this image: oracle/eigen_stub/ holds a
// from-scratch stand-in for the Eigen subset those files use (eager,
// coefficient-wise IEEE double) and type-only Bullet declarations.  The metric
// `distance_function` lives in src/smalltest.cpp next to main(); the Makefile
// extracts lines 167-175 verbatim into _ref/distance_function.inc at build time.
//
// What is glue here (not reference code): file I/O, building KDTreeNode /
// Obstacle objects from flat arrays, and the two loops the reference routes
// through Bullet (dubinsedge.cpp:875-893 calls DetectBulletCollision; the
// analytic call it shows commented at :881-884 is made here instead), exactly
// as SURVEY.md 8c fixes the oracle.
//
//   drrt_ref <command> <in.bin> <out.bin>      (raw little-endian records)
#include <DRRT/drrt.h>
#include <DRRT/thetastar.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <unordered_map>
#include <vector>

#include "_ref/distance_function.inc"  // double distance_function(VectorXd a, VectorXd b)

#ifdef DRRT_SHIM
// the same driver linked against drrt_b200/shim/kdtree_gpu.cpp instead of the
// reference's kdtree.cpp (oracle/_ref/drrt_shim): the reference's own JList /
// Edge / Obstacle objects call into the CUDA library
#include "drrt_gpu_shim.h"
#endif

// src/thetastar.cpp:10-15 (DistFunc) is compiled from the reference too
// obstacle.h:262-264 declares QuickCheck2D with a Vector2d point, obstacle.cpp:925
// defines it with a VectorXd one; this is the definition's prototype
bool QuickCheck2D(std::shared_ptr<ConfigSpace> &C, Eigen::VectorXd point, std::shared_ptr<Obstacle> &O);
double DistFunc(Eigen::VectorXd a, Eigen::VectorXd b);

using std::shared_ptr;
using std::make_shared;

struct Reader {
  std::vector<unsigned char> buf;
  size_t pos = 0;
  explicit Reader(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)n);
    if (n && fread(buf.data(), 1, (size_t)n, f) != (size_t)n) { perror("read"); exit(2); }
    fclose(f);
  }
  template <typename T> T get() { T v; memcpy(&v, buf.data() + pos, sizeof(T)); pos += sizeof(T); return v; }
  template <typename T> std::vector<T> arr(size_t n) {
    std::vector<T> v(n);
    if (n) memcpy(v.data(), buf.data() + pos, n * sizeof(T));
    pos += n * sizeof(T);
    return v;
  }
};

struct Writer {
  FILE *f;
  explicit Writer(const char *path) { f = fopen(path, "wb"); if (!f) { perror(path); exit(2); } }
  ~Writer() { fclose(f); }
  template <typename T> void put(const T &v) { fwrite(&v, sizeof(T), 1, f); }
  template <typename T> void arr(const std::vector<T> &v) { if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f); }
};

static Eigen::VectorXd vec3(const double *p) {
  Eigen::VectorXd v(3);
  v(0) = p[0]; v(1) = p[1]; v(2) = p[2];
  return v;
}

typedef double (*metric_fn)(Eigen::VectorXd, Eigen::VectorXd);
static metric_fn pick_metric(int m) { return m == 1 ? DistFunc : distance_function; }

static shared_ptr<KDTree> make_tree(int metric, int wrap) {
  shared_ptr<KDTree> T;
  if (wrap) {
    Eigen::VectorXi wraps(1);
    wraps(0) = 2;
    Eigen::VectorXd wrap_points(1);
    wrap_points(0) = 2.0 * PI;  // smalltest.cpp:205-208
    T = make_shared<KDTree>(3, wraps, wrap_points);
  } else {
    T = make_shared<KDTree>(3);
  }
  T->SetDistanceFunction(pick_metric(metric));
  return T;
}

// ---- range / nearest ---------------------------------------------------------
static int cmd_tree(const char *cmd, Reader &in, Writer &out) {
  int metric = in.get<int32_t>(), wrap = in.get<int32_t>();
  int64_t n = in.get<int64_t>(), nq = in.get<int64_t>();
  std::vector<double> pos = in.arr<double>(3 * n), q = in.arr<double>(3 * nq), rad = in.arr<double>(nq);
  if (!strcmp(cmd, "recreate")) {
    // ThetaStar builds a fresh local KDTree on every run (thetastar.cpp:50-51, re-entered from
    // obstacle.cpp:526) and lets the old one die: tree A over the first half of the nodes is
    // queried and destroyed, tree B over all of them is built (the allocator usually hands A's
    // address out again) and queried -- B must answer from its own nodes only
    for (int pass = 0; pass < 2; pass++) {
      const int64_t m = pass == 0 ? n / 2 : n;
      shared_ptr<KDTree> T = make_tree(metric, wrap);
      std::unordered_map<KDTreeNode *, int> id_of;
      std::vector<shared_ptr<KDTreeNode>> nodes;
      for (int64_t i = 0; i < m; i++) {
        shared_ptr<KDTreeNode> nd = make_shared<KDTreeNode>(vec3(&pos[3 * i]));
        T->KDInsert(nd);
        id_of[nd.get()] = (int)i;
        nodes.push_back(nd);
      }
      std::vector<int32_t> counts(nq), ids;
      for (int64_t i = 0; i < nq; i++) {
        shared_ptr<JList> S = make_shared<JList>(true);
        T->KDFindWithinRange(S, rad[i], vec3(&q[3 * i]));
        counts[i] = S->length_;
        shared_ptr<KDTreeNode> t = make_shared<KDTreeNode>();
        shared_ptr<double> k = make_shared<double>(0);
        while (S->length_ > 0) {
          T->PopFromRangeList(S, t, k);
          ids.push_back(id_of[t.get()]);
        }
      }
      out.put<int64_t>((int64_t)ids.size());
      out.put<int32_t>(T->root ? id_of[T->root.get()] : -1);
      out.put<int32_t>(T->tree_size_);
      out.arr(counts); out.arr(ids);
    }
    return 0;
  }
  shared_ptr<KDTree> T = make_tree(metric, wrap);
  std::unordered_map<KDTreeNode *, int> id_of;
  std::vector<shared_ptr<KDTreeNode>> nodes;
  for (int64_t i = 0; i < n; i++) {
    shared_ptr<KDTreeNode> nd = make_shared<KDTreeNode>(vec3(&pos[3 * i]));
    T->KDInsert(nd);  // kdtree.cpp:87-136
    id_of[nd.get()] = (int)i;
    nodes.push_back(nd);
  }
  if (!strcmp(cmd, "structure")) {
    std::vector<int32_t> parent(n), left(n), right(n), split(n);
    for (int64_t i = 0; i < n; i++) {
      KDTreeNode *k = nodes[i].get();
      parent[i] = k->kd_parent_exist_ ? id_of[k->kd_parent_.get()] : -1;
      left[i] = k->kd_child_L_exist_ ? id_of[k->kd_child_L_.get()] : -1;
      right[i] = k->kd_child_R_exist_ ? id_of[k->kd_child_R_.get()] : -1;
      split[i] = k->kd_split_;
    }
    out.arr(parent); out.arr(left); out.arr(right); out.arr(split);
    return 0;
  }
  if (!strcmp(cmd, "range")) {
    std::vector<int32_t> counts(nq), ids;
    std::vector<double> dist;
    for (int64_t i = 0; i < nq; i++) {
      shared_ptr<JList> S = make_shared<JList>(true);
      T->KDFindWithinRange(S, rad[i], vec3(&q[3 * i]));  // kdtree.cpp:794-820
      counts[i] = S->length_;
      shared_ptr<KDTreeNode> t = make_shared<KDTreeNode>();
      shared_ptr<double> k = make_shared<double>(0);
      while (S->length_ > 0) {  // PopFromRangeList kdtree.cpp:684-690 (LIFO)
        T->PopFromRangeList(S, t, k);
        ids.push_back(id_of[t.get()]);
        dist.push_back(*k);
      }
    }
    out.put<int64_t>((int64_t)ids.size());
    out.arr(counts); out.arr(ids); out.arr(dist);
    return 0;
  }
  if (!strcmp(cmd, "range_time")) {
    // bench.py --impl reference: the reference's own KDFindWithinRange on its own tree, one call
    // per query followed by the caller's EmptyRangeList (drrt.cpp:233, 356), timed as a whole;
    // `warm` untimed passes over the queries, then `timed` timed ones
    int warm = in.get<int32_t>(), timed = in.get<int32_t>();
    int64_t hits = 0;
    double secs = 0.0;
    for (int rep = 0; rep < warm + timed; rep++) {
      struct timespec t0, t1;
      clock_gettime(CLOCK_MONOTONIC, &t0);
      for (int64_t i = 0; i < nq; i++) {
        shared_ptr<JList> S = make_shared<JList>(true);
        T->KDFindWithinRange(S, rad[i], vec3(&q[3 * i]));  // kdtree.cpp:794-820
        if (rep == warm) hits += S->length_;
        T->EmptyRangeList(S);                              // kdtree.cpp:692-701
      }
      clock_gettime(CLOCK_MONOTONIC, &t1);
      if (rep >= warm) secs += (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    }
    out.put<int64_t>(hits);
    out.put<double>(secs);
    return 0;
  }
  if (!strcmp(cmd, "nearest")) {
    std::vector<int32_t> ids(nq);
    std::vector<double> dist(nq);
    for (int64_t i = 0; i < nq; i++) {
      shared_ptr<KDTreeNode> t = make_shared<KDTreeNode>();
      shared_ptr<double> d = make_shared<double>(0);
      T->KDFindNearest(t, d, vec3(&q[3 * i]));  // kdtree.cpp:264-307
      ids[i] = id_of[t.get()];
      dist[i] = *d;
    }
    out.arr(ids); out.arr(dist);
    return 0;
  }
  if (!strcmp(cmd, "ghosts")) {
    // GetNextGhostPoint ghostpoint.cpp:15-75 with best_dist = rad[i]
    std::vector<int32_t> counts(nq);
    std::vector<double> ghosts;
    for (int64_t i = 0; i < nq; i++) {
      shared_ptr<GhostPointIterator> it = make_shared<GhostPointIterator>(T.get(), vec3(&q[3 * i]));
      // the iterator's flag vector is uninitialised storage upstream;
      // the documented intent (ghostPoint.h:13-18) is all-zero
      it->wrap_dim_flags_.setZero();
      int c = 0;
      while (true) {
        Eigen::VectorXd g = GetNextGhostPoint(it, rad[i]);
        if (g.isZero(0)) break;
        for (int k = 0; k < 3; k++) ghosts.push_back(g(k));
        c++;
      }
      counts[i] = c;
    }
    out.arr(counts); out.arr(ghosts);
    return 0;
  }
  return 1;
}

// ---- obstacles -----------------------------------------------------------------
static shared_ptr<ConfigSpace> make_cspace(int metric, double r_min, double robot_radius,
                                           double collision_distance) {
  Eigen::VectorXd lo(3), hi(3), s(3), g(3);
  lo(0) = 0; lo(1) = 0; lo(2) = 0; hi(0) = 50; hi(1) = 50; hi(2) = 2 * PI;
  s = lo; g = hi;
  shared_ptr<ConfigSpace> C = make_shared<ConfigSpace>(3, lo, hi, s, g);
  C->SetDistanceFunction(pick_metric(metric));
  C->min_turn_radius_ = r_min;       // smalltest.cpp:198
  C->robot_radius_ = robot_radius;   // smalltest.cpp:199
  C->collision_distance_ = collision_distance;
  C->space_has_time_ = false;        // smalltest.cpp:200
  C->space_has_theta_ = true;
  C->in_warmup_time_ = false;
  return C;
}

static std::vector<shared_ptr<Obstacle>> read_obstacles(Reader &in, shared_ptr<ConfigSpace> C) {
  int32_t n = in.get<int32_t>();
  std::vector<int32_t> kind = in.arr<int32_t>(n);
  std::vector<uint8_t> used = in.arr<uint8_t>(n);
  std::vector<double> life = in.arr<double>(n), origin = in.arr<double>(2 * n), radius = in.arr<double>(n);
  std::vector<int32_t> voff = in.arr<int32_t>(n + 1);
  std::vector<double> verts = in.arr<double>(2 * (size_t)(n ? voff[n] : 0));
  std::vector<shared_ptr<Obstacle>> obs;
  for (int i = 0; i < n; i++) {
    int nv = voff[i + 1] - voff[i];
    Eigen::MatrixX2d poly;
    poly.resize(nv, 2);
    for (int v = 0; v < nv; v++) {
      poly(v, 0) = verts[2 * (voff[i] + v)];
      poly(v, 1) = verts[2 * (voff[i] + v) + 1];
    }
    Eigen::VectorXd o(3);
    o(0) = origin[2 * i]; o(1) = origin[2 * i + 1]; o(2) = 0.0;
    Eigen::MatrixXd path;
    Eigen::VectorXd path_times;
    shared_ptr<Obstacle> O = make_shared<Obstacle>(3, poly, true, o, path, path_times);  // obstacle.h:107-144
    O->kind_ = kind[i];
    O->obstacle_used_ = used[i] != 0;
    O->life_span_ = life[i];
    O->radius_ = radius[i];
    O->cspace = C;
    obs.push_back(O);
  }
  return obs;
}

// ---- Dubins trajectories / edge checks -----------------------------------------
static int cmd_edges(const char *cmd, Reader &in, Writer &out) {
  int metric = in.get<int32_t>();
  double r_min = in.get<double>(), robot_radius = in.get<double>();
  shared_ptr<ConfigSpace> C = make_cspace(metric, r_min, robot_radius, 0.1);
  shared_ptr<KDTree> T = make_tree(metric, 1);
  std::vector<shared_ptr<Obstacle>> obs = read_obstacles(in, C);
  int64_t n = in.get<int64_t>();
  std::vector<double> s = in.arr<double>(3 * n), e = in.arr<double>(3 * n);
  std::vector<char> types(4 * n, 0);
  std::vector<double> dist(n), xy;
  std::vector<int32_t> npts(n);
  std::vector<uint8_t> verdict(n, 0);
  bool want_traj = !strcmp(cmd, "traj");
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int64_t i = 0; i < n; i++) {
    shared_ptr<KDTreeNode> a = make_shared<KDTreeNode>(vec3(&s[3 * i]));
    shared_ptr<KDTreeNode> b = make_shared<KDTreeNode>(vec3(&e[3 * i]));
    shared_ptr<Edge> edge = Edge::NewEdge(C, T, a, b);  // dubinsedge.cpp:12-21
    edge->CalculateTrajectory();                         // dubinsedge.cpp:172-830
    strncpy(&types[4 * i], edge->edge_type_.c_str(), 3);
    dist[i] = edge->dist_;
    npts[i] = (int32_t)edge->trajectory_.rows();
    if (want_traj)
      for (int k = 0; k < edge->trajectory_.rows(); k++) {
        xy.push_back(edge->trajectory_(k, 0));
        xy.push_back(edge->trajectory_(k, 1));
      }
    // ExplicitEdgeCheck(C,edge) obstacle.cpp:897-916 over
    // DubinsEdge::ExplicitEdgeCheck dubinsedge.cpp:875-893 with the analytic
    // call of :881-884 (ExplicitEdgeCheck2D obstacle.cpp:756-877)
    bool hit = false;
    for (size_t o = 0; o < obs.size() && !hit; o++)
      for (int k = 1; k < edge->trajectory_.rows() && !hit; k++)
        hit = ExplicitEdgeCheck2D(obs[o], edge->trajectory_.row(k - 1), edge->trajectory_.row(k),
                                  C->robot_radius_);
    verdict[i] = hit ? 1 : 0;
#ifdef DRRT_SHIM
    if (!strcmp(cmd, "edgecheck_gpu")) {
      static bool uploaded = false;
      if (!uploaded) {
        for (size_t o = 0; o < obs.size(); o++) C->obstacles_->ListPush(obs[o]);
        DrrtGpuUploadObstacles(T.get(), C);
        uploaded = true;
      }
      verdict[i] = ExplicitEdgeCheckGpu(C, T.get(), edge) ? 1 : 0;
      dist[i] = edge->dist_;
    }
#endif
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (!strcmp(cmd, "edgecheck_time")) {
    // bench.py: NewEdge + CalculateTrajectory + the analytic edge check per edge, timed as a whole
    int64_t hits = 0;
    for (int64_t i = 0; i < n; i++) hits += verdict[i];
    out.put<int64_t>(hits);
    out.put<double>((double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec));
    return 0;
  }
  out.arr(types); out.arr(dist); out.arr(npts); out.arr(verdict);
  if (want_traj) out.arr(xy);
  return 0;
}

// ---- obstacle file -> list -> edge checks -------------------
// Obstacle::ReadObstaclesFromFile (obstacle.cpp:25-155) fills ConfigSpace::obstacles_; the
// obstacle thread later switches obstacles on (obstacle_used_), done here directly.
static int cmd_obsfile(const char *cmd, Reader &in, Writer &out) {
  int metric = in.get<int32_t>();
  double r_min = in.get<double>(), robot_radius = in.get<double>();
  int32_t plen = in.get<int32_t>();
  std::vector<char> pbuf = in.arr<char>(plen);
  std::string path(pbuf.begin(), pbuf.end());
  int64_t n = in.get<int64_t>();
  std::vector<double> s = in.arr<double>(3 * n), e = in.arr<double>(3 * n);
  shared_ptr<ConfigSpace> C = make_cspace(metric, r_min, robot_radius, 0.1);
  shared_ptr<KDTree> T = make_tree(metric, 1);
  Obstacle::ReadObstaclesFromFile(path, C);
  {
    shared_ptr<ListNode> ln = C->obstacles_->front_;
    for (int i = 0; i < C->obstacles_->length_; i++, ln = ln->child_) ln->obstacle_->obstacle_used_ = true;
  }
#ifdef DRRT_SHIM
  DrrtGpuUploadObstacles(T.get(), C);
#endif
  std::vector<uint8_t> verdict(n, 0);
  std::vector<double> dist(n);
  for (int64_t i = 0; i < n; i++) {
    shared_ptr<KDTreeNode> a = make_shared<KDTreeNode>(vec3(&s[3 * i]));
    shared_ptr<KDTreeNode> b = make_shared<KDTreeNode>(vec3(&e[3 * i]));
    shared_ptr<Edge> edge = Edge::NewEdge(C, T, a, b);
    bool hit = false;
#ifdef DRRT_SHIM
    hit = ExplicitEdgeCheckGpu(C, T.get(), edge);
#else
    edge->CalculateTrajectory();
    shared_ptr<ListNode> ln = C->obstacles_->front_;  // obstacle.cpp:897-916
    for (int o = 0; o < C->obstacles_->length_ && !hit; o++, ln = ln->child_)
      for (int k = 1; k < edge->trajectory_.rows() && !hit; k++)
        hit = ExplicitEdgeCheck2D(ln->obstacle_, edge->trajectory_.row(k - 1), edge->trajectory_.row(k),
                                  C->robot_radius_);
#endif
    verdict[i] = hit ? 1 : 0;
    dist[i] = edge->dist_;
  }
  out.put<int32_t>(C->obstacles_->length_);
  out.arr(verdict); out.arr(dist);
  (void)cmd;
  return 0;
}

// ---- config 1: the planner iteration, timed (bench.py / benchmarks/inloop.py) -----------
// One iteration = what RrtMainLoop + Extend do on the hot path (mainloop.cpp:204-256,
// drrt.cpp:213-293): sample -> KDFindNearest -> Edge::Saturate -> ExplicitNodeCheck ->
// KDFindWithinRange -> for every neighbour NewEdge + CalculateTrajectory + ExplicitEdgeCheck in
// both directions -> KDInsert -> EmptyRangeList, over the reference's own KDTree / JList / Edge
// objects with the parameter block of src/smalltest.cpp:180-220.  drrt_ref runs the reference's
// kdtree.cpp and the analytic checker; drrt_shim runs the SAME loop with kdtree.cpp swapped for
// the shim (every tree call lands in the CUDA library) and the checks through the shim's
// ExplicitEdgeCheckGpu (mode 0: one call per edge, the unchanged planner) or
// ExplicitEdgeCheckBatchGpu (mode 1: the 2|N| edges of an iteration in one call, the
// INTEGRATION.md patch).  The RNG is a fixed xorshift (the reference's sampler is unseedable).
static uint64_t g_rng;
static double urand() {
  g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17;
  return (double)(g_rng >> 11) * (1.0 / 9007199254740992.0);
}

static int cmd_loop(Reader &in, Writer &out) {
  int metric = in.get<int32_t>();
  double r_min = in.get<double>(), robot_radius = in.get<double>();
  shared_ptr<ConfigSpace> C = make_cspace(metric, r_min, robot_radius, 0.1);
  shared_ptr<KDTree> T = make_tree(metric, 1);
  std::vector<shared_ptr<Obstacle>> obs = read_obstacles(in, C);
  int32_t iters = in.get<int32_t>(), mode = in.get<int32_t>();
  g_rng = (uint64_t)in.get<int64_t>() | 1ull;
  double delta = in.get<double>(), gamma = in.get<double>();
  for (size_t o = 0; o < obs.size(); o++) C->obstacles_->ListPush(obs[o]);
#ifdef DRRT_SHIM
  // (ListPush prepends: the device table follows the list order, as the checker walks it)
#endif
  double root[3] = {0.0, 0.0, -3.0 * PI / 4.0};  // smalltest.cpp:190-196
  shared_ptr<KDTreeNode> rn = make_shared<KDTreeNode>(vec3(root));
  T->KDInsert(rn);
#ifdef DRRT_SHIM
  DrrtGpuUploadObstacles(T.get(), C);
#endif
  int64_t n_hits = 0, n_edges = 0, n_coll = 0, n_blocked = 0, id_sum = 0;
  double len_sum = 0.0;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int it = 0; it < iters; it++) {
    const int n = T->tree_size_;
    double radius = gamma * pow(log(1.0 + n) / (double)n, 1.0 / 3.0);  // the intended exponent
    if (radius > delta) radius = delta;
    double sp[3] = {50.0 * urand(), 50.0 * urand(), 2.0 * PI * urand()};
    shared_ptr<KDTreeNode> nn = make_shared<KDTreeNode>(vec3(sp));
    shared_ptr<KDTreeNode> closest = make_shared<KDTreeNode>();
    shared_ptr<double> cd = make_shared<double>(INF);
    T->KDFindNearest(closest, cd, nn->position_);                       // mainloop.cpp:208
    double d0 = T->distanceFunction(nn->position_, closest->position_);
    if (d0 > delta) Edge::Saturate(nn->position_, closest->position_, delta, d0);  // :219
    bool blocked = false;                                               // :226 ExplicitNodeCheck
#ifdef DRRT_SHIM
    blocked = ExplicitNodeCheckGpu(C, T.get(), nn);
#else
    for (size_t o = 0; o < obs.size() && !blocked; o++) blocked = QuickCheck2D(C, nn->position_, obs[obs.size() - 1 - o]);
    for (size_t o = 0; o < obs.size() && !blocked; o++)
      blocked = ExplicitPointCheck2D(C, obs[obs.size() - 1 - o], nn->position_, C->robot_radius_);
#endif
    if (blocked) { n_blocked++; continue; }
    shared_ptr<JList> S = make_shared<JList>(true);
    T->KDFindWithinRange(S, radius, nn->position_);                    // drrt.cpp:213-216
    n_hits += S->length_;
    std::vector<shared_ptr<Edge>> edges;
    shared_ptr<JListNode> ln = S->front_;            // drrt.cpp:251-256: length_ steps from front_
    for (int i = 0; i < S->length_; i++, ln = ln->child_) {
      shared_ptr<KDTreeNode> near = ln->node_;
      edges.push_back(Edge::NewEdge(C, T, nn, near));                  // drrt.cpp:405 (FindBestParent)
      edges.push_back(Edge::NewEdge(C, T, near, nn));                  // drrt.cpp:284
    }
    n_edges += (int64_t)edges.size();
#ifdef DRRT_SHIM
    if (mode == 1) {
      std::vector<unsigned char> unsafe;
      ExplicitEdgeCheckBatchGpu(C, T.get(), edges, unsafe);
      for (size_t e = 0; e < edges.size(); e++) { n_coll += unsafe[e]; len_sum += edges[e]->dist_; }
    } else {
      for (size_t e = 0; e < edges.size(); e++) {
        n_coll += ExplicitEdgeCheckGpu(C, T.get(), edges[e]) ? 1 : 0;
        len_sum += edges[e]->dist_;
      }
    }
#else
    (void)mode;
    for (size_t e = 0; e < edges.size(); e++) {
      edges[e]->CalculateTrajectory();                                  // drrt.cpp:411, 286
      bool hit = false;                                                 // ExplicitEdgeCheck(C, edge) :422, 293
      shared_ptr<ListNode> on = C->obstacles_->front_;
      for (int o = 0; o < C->obstacles_->length_ && !hit; o++, on = on->child_)
        for (int k = 1; k < edges[e]->trajectory_.rows() && !hit; k++)
          hit = ExplicitEdgeCheck2D(on->obstacle_, edges[e]->trajectory_.row(k - 1), edges[e]->trajectory_.row(k),
                                    C->robot_radius_);
      n_coll += hit ? 1 : 0;
      len_sum += edges[e]->dist_;
    }
#endif
    T->KDInsert(nn);                                                    // drrt.cpp:239-242
    ln = S->front_;
    for (int i = 0; i < S->length_; i++, ln = ln->child_) id_sum += (int64_t)(ln->key_ * 1048576.0);
    T->EmptyRangeList(S);                                               // drrt.cpp:356
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  out.put<int64_t>(T->tree_size_);
  out.put<int64_t>(n_hits); out.put<int64_t>(n_edges); out.put<int64_t>(n_coll);
  out.put<int64_t>(n_blocked); out.put<int64_t>(id_sum);
  out.put<double>(len_sum);
  out.put<double>((double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec));
  return 0;
}

// ---- ExplicitEdgeCheck2D kinds 6 / 7 (obstacle.cpp:805-875) ---------------------------------
// time-parametrised obstacles: origin_, path_ rows (dx, dy, time); segments (x, y, time)
static int cmd_timed(Reader &in, Writer &out) {
  double radius = in.get<double>();
  int32_t n_obs = in.get<int32_t>();
  std::vector<int32_t> kind = in.arr<int32_t>(n_obs);
  std::vector<uint8_t> used = in.arr<uint8_t>(n_obs);
  std::vector<double> life = in.arr<double>(n_obs), origin = in.arr<double>(2 * n_obs), rad = in.arr<double>(n_obs);
  std::vector<int32_t> poff = in.arr<int32_t>(n_obs + 1);
  std::vector<double> path = in.arr<double>(3 * (size_t)(n_obs ? poff[n_obs] : 0));
  shared_ptr<ConfigSpace> C = make_cspace(0, 1.0, radius, 0.1);
  std::vector<shared_ptr<Obstacle>> obs;
  for (int i = 0; i < n_obs; i++) {
    Eigen::MatrixX2d poly;
    poly.resize(0, 2);
    Eigen::VectorXd o(3);
    o(0) = origin[2 * i]; o(1) = origin[2 * i + 1]; o(2) = 0.0;
    const int rows = poff[i + 1] - poff[i];
    Eigen::MatrixXd p;
    p.resize(rows, 3);
    for (int r = 0; r < rows; r++)
      for (int k = 0; k < 3; k++) p(r, k) = path[3 * (size_t)(poff[i] + r) + k];
    Eigen::VectorXd times;
    shared_ptr<Obstacle> O = make_shared<Obstacle>(3, poly, true, o, p, times);
    O->kind_ = kind[i];
    O->obstacle_used_ = used[i] != 0;
    O->life_span_ = life[i];
    O->radius_ = rad[i];
    O->path_ = p;
    O->cspace = C;
    obs.push_back(O);
  }
  int64_t n = in.get<int64_t>();
  std::vector<double> s = in.arr<double>(3 * n), e = in.arr<double>(3 * n);
  std::vector<uint8_t> verdict(n, 0);
  for (int64_t i = 0; i < n; i++) {
    bool hit = false;
    for (size_t o = 0; o < obs.size() && !hit; o++)
      hit = ExplicitEdgeCheck2D(obs[o], vec3(&s[3 * i]), vec3(&e[3 * i]), radius);
    verdict[i] = hit ? 1 : 0;
  }
  out.arr(verdict);
  return 0;
}

// ---- node validity ------------------------------------------------------------
static int cmd_points(Reader &in, Writer &out) {
  int metric = in.get<int32_t>();
  double robot_radius = in.get<double>(), collision_distance = in.get<double>();
  shared_ptr<ConfigSpace> C = make_cspace(metric, 1.0, robot_radius, collision_distance);
  std::vector<shared_ptr<Obstacle>> obs = read_obstacles(in, C);
  int64_t n = in.get<int64_t>();
  std::vector<double> p = in.arr<double>(3 * n);
  std::vector<uint8_t> verdict(n, 0);
  for (int64_t i = 0; i < n; i++) {
    Eigen::VectorXd pt = vec3(&p[3 * i]);
    bool hit = false;  // ExplicitPointCheck obstacle.cpp:1062-1085: QuickCheck, then the 2D check
    for (size_t o = 0; o < obs.size() && !hit; o++) hit = QuickCheck2D(C, pt, obs[o]);
    for (size_t o = 0; o < obs.size() && !hit; o++) hit = ExplicitPointCheck2D(C, obs[o], pt, C->robot_radius_);
    verdict[i] = hit ? 1 : 0;
  }
  out.arr(verdict);
  return 0;
}

// ---- geometry --------------------------------------------------------------------
static int cmd_geom(Reader &in, Writer &out) {
  int64_t n = in.get<int64_t>();
  std::vector<double> v = in.arr<double>(8 * n);  // PA PB QA QB (x,y each)
  std::vector<double> seg(n), pt(n);
  for (int64_t i = 0; i < n; i++) {
    Eigen::VectorXd PA(2), PB(2), QA(2), QB(2);
    PA(0) = v[8 * i]; PA(1) = v[8 * i + 1]; PB(0) = v[8 * i + 2]; PB(1) = v[8 * i + 3];
    QA(0) = v[8 * i + 4]; QA(1) = v[8 * i + 5]; QB(0) = v[8 * i + 6]; QB(1) = v[8 * i + 7];
    seg[i] = SegmentDistSqrd(PA, PB, QA, QB);                      // distancefunctions.cpp:94-164
    Eigen::Vector2d a, b;
    a(0) = QA(0); a(1) = QA(1); b(0) = QB(0); b(1) = QB(1);
    pt[i] = DistanceSqrdPointToSegment(PA, a, b);                   // :50-74
  }
  out.arr(seg); out.arr(pt);
  return 0;
}

int main(int argc, char **argv) {
  if (argc != 4) { fprintf(stderr, "usage: drrt_ref <command> <in.bin> <out.bin>\n"); return 2; }
  Reader in(argv[2]);
  Writer out(argv[3]);
  const char *cmd = argv[1];
  if (!strcmp(cmd, "range") || !strcmp(cmd, "range_time") || !strcmp(cmd, "nearest") || !strcmp(cmd, "structure") ||
      !strcmp(cmd, "ghosts") || !strcmp(cmd, "recreate"))
    return cmd_tree(cmd, in, out);
  if (!strcmp(cmd, "traj") || !strcmp(cmd, "edgecheck") || !strcmp(cmd, "edgecheck_gpu") || !strcmp(cmd, "edgecheck_time"))
    return cmd_edges(cmd, in, out);
  if (!strcmp(cmd, "points")) return cmd_points(in, out);
  if (!strcmp(cmd, "loop_time")) return cmd_loop(in, out);
  if (!strcmp(cmd, "timed")) return cmd_timed(in, out);
  if (!strcmp(cmd, "obsfile")) return cmd_obsfile(cmd, in, out);
  if (!strcmp(cmd, "geom")) return cmd_geom(in, out);
  fprintf(stderr, "unknown command %s\n", cmd);
  return 2;
}

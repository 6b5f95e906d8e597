// kdtree_gpu.cpp -- drop-in replacement for the reference's src/kdtree.cpp.
//
// Compiled against the reference's own headers (<DRRT/kdtree.h>); defines every
// KDTree method that src/kdtree.cpp defines, with the same signatures, and
// forwards the tree work to libdrrt_b200.so through include/drrt_gpu.h.  The
// planner (drrt.cpp, mainloop.cpp, moverobot.cpp, obstacle.cpp, thetastar.cpp)
// links against this file instead of kdtree.cpp and runs unchanged.
//
// What stays on the host (the reference's callers read these fields directly):
//   * the viz list nodes_ (visualizer.cpp:210-225),
//   * KDTreeNode::kd_parent_/kd_child_L_/kd_child_R_/kd_split_ (drrt.cpp:849,
//     GetNodeAt, PrintTree) -- filled from the device tree after each insert,
//   * KDTreeNode::in_heap_, the range-list de-duplication flag (AddToRangeList
//     kdtree.cpp:670-682), because callers accumulate several searches into one
//     JList (KDFindMoreWithinRange, drrt.cpp:1063-1066).
// There is no CPU search path here: if the library or the GPU is missing the
// constructor-side create fails and every query prints the error and returns
// an empty result, following the reference's "print and carry on" convention.
#include <DRRT/kdtree.h>

#include <cstdio>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "drrt_gpu.h"

namespace {

struct ShimState {
  drrt_store *h = nullptr;
  std::vector<std::shared_ptr<KDTreeNode>> by_id;  // device id -> host node
  bool failed = false;
  std::mutex row_mu;  // guards the reusable result row below
  // One reusable PINNED result row (drrt_host_alloc): the planner asks one query per
  // iteration (mainloop.cpp:204-210, drrt.cpp:213-216), so the whole call is one
  // drrt_range_batch_into -- one kernel, the row DMA'd straight into this buffer.
  double *q_pin = nullptr;
  int64_t *off_pin = nullptr;
  int32_t *ids_pin = nullptr;
  double *dist_pin = nullptr;
  int64_t cap = 0;

  bool reserve_row(int64_t need) {
    if (need <= cap && q_pin) return true;
    int64_t ncap = cap ? cap : 4096;
    while (ncap < need) ncap *= 2;
    void *a = nullptr, *b = nullptr, *c = nullptr, *d = nullptr;
    if (drrt_host_alloc((size_t)ncap * sizeof(int32_t), &a) != DRRT_OK ||
        drrt_host_alloc((size_t)ncap * sizeof(double), &b) != DRRT_OK ||
        (!q_pin && (drrt_host_alloc(4 * sizeof(double), &c) != DRRT_OK ||
                    drrt_host_alloc(2 * sizeof(int64_t), &d) != DRRT_OK))) {
      drrt_host_free(a); drrt_host_free(b); drrt_host_free(c); drrt_host_free(d);
      return false;
    }
    drrt_host_free(ids_pin);
    drrt_host_free(dist_pin);
    ids_pin = (int32_t *)a;
    dist_pin = (double *)b;
    if (!q_pin) { q_pin = (double *)c; off_pin = (int64_t *)d; }
    cap = ncap;
    return true;
  }
  void release() {
    if (h) drrt_store_destroy(h);
    h = nullptr;
    drrt_host_free(q_pin); drrt_host_free(off_pin); drrt_host_free(ids_pin); drrt_host_free(dist_pin);
    q_pin = nullptr; off_pin = nullptr; ids_pin = nullptr; dist_pin = nullptr;
    cap = 0;
    by_id.clear();
  }
};

std::mutex g_mu;
std::unordered_map<const KDTree *, ShimState> g_state;

// the reference passes a function pointer (kdtree.h:21); map it onto the device enum by
// probing it with values only the wrap-aware metric can answer
int metric_of(double (*f)(Eigen::VectorXd, Eigen::VectorXd)) {
  if (!f) return DRRT_METRIC_R2S1_WRAP;
  Eigen::VectorXd a(3), b(3);
  a(0) = 0; a(1) = 0; a(2) = 0.05;
  b(0) = 0; b(1) = 0; b(2) = 2.0 * PI - 0.05;
  double d = f(a, b);
  if (std::abs(d - 0.1) < 1e-9) return DRRT_METRIC_R2S1_WRAP;     // smalltest.cpp:167-175
  if (std::abs(d - (2.0 * PI - 0.1)) < 1e-9) return DRRT_METRIC_EUCLID;  // thetastar.cpp:10-15
  return -1;
}

ShimState &state_of(KDTree *t) {
  std::lock_guard<std::mutex> lock(g_mu);
  ShimState &s = g_state[t];
  // The state is keyed by the tree's address and the reference's KDTree has no destructor to
  // hook (kdtree.h:13-45).  ThetaStar builds a fresh local tree on every run (thetastar.cpp:50-51,
  // re-entered from obstacle.cpp:526), and the allocator likes to hand the old address out again:
  // a tree that says it is empty while this state still holds nodes is a NEW tree -- drop the old
  // device store and node table instead of answering from them.
  if (t->tree_size_ == 0 && !s.by_id.empty()) {
    s.release();
    s.failed = false;
  }
  if (!s.h && !s.failed) {
    int metric = metric_of(t->distanceFunction);
    int wrap_dims[4] = {0, 0, 0, 0};
    double wrap_points[4] = {0, 0, 0, 0};
    for (int i = 0; i < t->num_wraps_ && i < 4; i++) {
      wrap_dims[i] = t->wraps_(i);
      wrap_points[i] = t->wrap_points_(i);
    }
    int rc = metric < 0 ? DRRT_ERR_UNSUPPORTED
                        : drrt_store_create(0, t->dimensions_, t->num_wraps_, wrap_dims, wrap_points,
                                            metric, 0, &s.h);
    if (rc != DRRT_OK) {
      s.failed = true;
      std::cout << "drrt_gpu: cannot create device store: "
                << (metric < 0 ? "unknown distance function" : drrt_last_error()) << std::endl;
    }
  }
  return s;
}

bool ok(int rc, const char *what) {
  if (rc == DRRT_OK) return true;
  std::cout << "drrt_gpu: " << what << ": " << drrt_last_error() << std::endl;
  return false;
}

}  // namespace

// kdtree.cpp:13-29 -- host only, unchanged behaviour
void KDTree::AddVizNode(std::shared_ptr<KDTreeNode> node) { this->nodes_.push_back(node); }

void KDTree::RemoveVizNode(std::shared_ptr<KDTreeNode> &node) {
  this->nodes_.erase(std::remove(this->nodes_.begin(), this->nodes_.end(), node), this->nodes_.end());
}

// kdtree.cpp:31-54
void KDTree::PrintTree(std::shared_ptr<KDTreeNode> node, int indent, char type) {
  if (indent) std::cout << std::string(indent - 1, ' ') << type;
  std::cout << node->position_(0) << "," << node->position_(1) << ": " << node->rrt_LMC_;
  if (node->rrt_parent_used_) {
    std::cout << " : " << node->rrt_parent_edge_->end_node_->position_(0) << ","
              << node->rrt_parent_edge_->end_node_->position_(1);
    std::cout << " (" << node->rrt_parent_edge_->dist_ << ")";
  }
  if (!node->kd_child_L_exist_ && !node->kd_child_R_exist_) {
    std::cout << " | leaf" << std::endl;
  } else {
    std::cout << std::endl;
    if (node->kd_child_L_exist_) PrintTree(node->kd_child_L_, indent + 4, '<');
    if (node->kd_child_R_exist_) PrintTree(node->kd_child_R_, indent + 4, '>');
  }
}

// kdtree.cpp:56-85: exact-position lookup over the host links
void KDTree::GetNodeAt(Eigen::VectorXd pos, std::shared_ptr<KDTreeNode> &node) {
  std::shared_ptr<KDTreeNode> parent = this->root;
  node = parent;
  while (parent) {
    if (parent->position_ == pos) { node = parent; return; }
    bool left = pos(parent->kd_split_) < parent->position_(parent->kd_split_);
    if (left) {
      if (!parent->kd_child_L_exist_) { std::cout << "no left child" << std::endl; return; }
      parent = parent->kd_child_L_;
    } else {
      if (!parent->kd_child_R_exist_) { std::cout << "no right child" << std::endl; return; }
      parent = parent->kd_child_R_;
    }
  }
}

// kdtree.cpp:87-136
bool KDTree::KDInsert(std::shared_ptr<KDTreeNode> &node) {
  if (node->kd_in_tree_) return false;
  ShimState &s = state_of(this);
  if (s.failed) return false;
  double pos[3] = {node->position_(0), node->position_(1), node->position_(2)};
  int32_t id = -1;
  if (!ok(drrt_store_insert(s.h, 1, pos, &id), "KDInsert")) return false;
  node->kd_in_tree_ = true;
  AddVizNode(node);
  s.by_id.push_back(node);
  // mirror the device tree's links for the host-side readers
  int32_t parent = -1, left = -1, right = -1, split = 0;
  if (ok(drrt_store_structure(s.h, id, 1, &parent, &left, &right, &split), "KDInsert links")) {
    node->kd_split_ = split;
    if (parent >= 0) {
      std::shared_ptr<KDTreeNode> &p = s.by_id[parent];
      node->kd_parent_ = p;
      node->kd_parent_exist_ = true;
      if (node->position_(p->kd_split_) < p->position_(p->kd_split_)) {
        p->kd_child_L_ = node;
        p->kd_child_L_exist_ = true;
      } else {
        p->kd_child_R_ = node;
        p->kd_child_R_exist_ = true;
      }
    }
  }
  if (id == 0) this->root = node;
  this->tree_size_ += 1;
  return true;
}

// kdtree.cpp:853-858
void KDTree::KDInsert(Eigen::VectorXd a) {
  std::shared_ptr<KDTreeNode> N = std::make_shared<KDTreeNode>();
  N->position_ = a;
  this->KDInsert(N);
}

// kdtree.cpp:264-307
bool KDTree::KDFindNearest(std::shared_ptr<KDTreeNode> &nearestNode,
                           std::shared_ptr<double> nearestNodeDist, Eigen::VectorXd queryPoint) {
  ShimState &s = state_of(this);
  if (s.failed) return false;
  double q[3] = {queryPoint(0), queryPoint(1), queryPoint(2)};
  int32_t id = -1;
  double d = 0.0;
  if (!ok(drrt_nearest_batch(s.h, 1, q, &id, &d), "KDFindNearest") || id < 0) return false;
  nearestNode = s.by_id[id];
  *nearestNodeDist = d;
  return true;
}

// kdtree.cpp:140-262: only ever entered with root == this->root (kdtree.cpp:274, 293)
bool KDTree::KDFindNearestInSubtree(std::shared_ptr<KDTreeNode> &nearestNode,
                                    std::shared_ptr<double> nearestNodeDist,
                                    std::shared_ptr<KDTreeNode> &root, Eigen::VectorXd queryPoint,
                                    std::shared_ptr<KDTreeNode> &suggestedClosestNode,
                                    double suggestedClosestDist) {
  if (root != this->root) {
    std::cout << "drrt_gpu: subtree searches are not offered by the device back end" << std::endl;
    return false;
  }
  bool r = KDFindNearest(nearestNode, nearestNodeDist, queryPoint);
  if (r && suggestedClosestDist < *nearestNodeDist) {
    nearestNode = suggestedClosestNode;
    *nearestNodeDist = suggestedClosestDist;
  }
  return r;
}

// kdtree.cpp:309-490: dead upstream (null shared_ptr dereferences); forwards
bool KDTree::KDFindNearestinSubtreeWithGuess(std::shared_ptr<KDTreeNode> nearestNode,
                                             std::shared_ptr<double> nearestNodeDist,
                                             std::shared_ptr<KDTreeNode> root,
                                             Eigen::VectorXd queryPoint,
                                             std::shared_ptr<KDTreeNode> suggestedClosestNode,
                                             double suggestedClosestDist) {
  return KDFindNearestInSubtree(nearestNode, nearestNodeDist, root, queryPoint,
                                suggestedClosestNode, suggestedClosestDist);
}

bool KDTree::KDFindNearestWithGuess(std::shared_ptr<KDTreeNode> nearestNode,
                                    std::shared_ptr<double> nearestNodeDist,
                                    Eigen::VectorXd queryPoint, std::shared_ptr<KDTreeNode> guess) {
  (void)guess;
  if (!nearestNodeDist) return false;
  return KDFindNearest(nearestNode, nearestNodeDist, queryPoint);
}

// kdtree.cpp:493-521 / 523-628: heap plumbing of the CPU k-nearest walk; unused by the device path
std::shared_ptr<KDTreeNode> KDTree::AddToKNNHeap(std::shared_ptr<BinaryHeap> H,
                                                 std::shared_ptr<KDTreeNode> node, double key,
                                                 int k) {
  (void)H; (void)key; (void)k;
  return node;
}

bool KDTree::KDFindKNearestInSubtree(std::shared_ptr<KDTreeNode> farthestNode,
                                     std::shared_ptr<double> farthestNodeDist,
                                     std::shared_ptr<KDTreeNode> root, int k,
                                     Eigen::VectorXd queryPoint,
                                     std::shared_ptr<BinaryHeap> nearestHeap) {
  (void)farthestNode; (void)farthestNodeDist; (void)root; (void)k; (void)queryPoint; (void)nearestHeap;
  std::cout << "drrt_gpu: use KDFindKNearest" << std::endl;
  return false;
}

// kdtree.cpp:630-666 (by intent): the k nearest, nearest first
std::vector<std::shared_ptr<KDTreeNode>> KDTree::KDFindKNearest(int k, Eigen::VectorXd queryPoint) {
  std::vector<std::shared_ptr<KDTreeNode>> out;
  ShimState &s = state_of(this);
  if (s.failed || k < 1) return out;
  double q[3] = {queryPoint(0), queryPoint(1), queryPoint(2)};
  std::vector<int32_t> ids(k);
  std::vector<double> dist(k);
  if (!ok(drrt_knn_batch(s.h, 1, q, k, ids.data(), dist.data()), "KDFindKNearest")) return out;
  for (int i = 0; i < k; i++)
    if (ids[i] >= 0) {
      s.by_id[ids[i]]->dist_ = dist[i];
      out.push_back(s.by_id[ids[i]]);
    }
  return out;
}

// kdtree.cpp:670-701 -- host list plumbing, unchanged behaviour
bool KDTree::AddToRangeList(std::shared_ptr<JList> &S, std::shared_ptr<KDTreeNode> &node,
                            double key) {
  if (node->in_heap_) return false;
  node->in_heap_ = true;
  S->JListPush(node, key);
  return true;
}

void KDTree::PopFromRangeList(std::shared_ptr<JList> &S, std::shared_ptr<KDTreeNode> &t,
                              std::shared_ptr<double> k) {
  S->JListPopKey(t, k);
  t->in_heap_ = false;
}

void KDTree::EmptyRangeList(std::shared_ptr<JList> &S) {
  std::shared_ptr<KDTreeNode> n = std::make_shared<KDTreeNode>();
  std::shared_ptr<double> k = std::make_shared<double>(0);
  while (S->length_ > 0) {
    S->JListPopKey(n, k);
    n->in_heap_ = false;
  }
}

// kdtree.cpp:794-820
void KDTree::KDFindWithinRange(std::shared_ptr<JList> &S, double range, Eigen::VectorXd queryPoint) {
  ShimState &s = state_of(this);
  if (s.failed) return;
  std::lock_guard<std::mutex> row(s.row_mu);
  if (!s.reserve_row(1)) { std::cout << "drrt_gpu: KDFindWithinRange: " << drrt_last_error() << std::endl; return; }
  s.q_pin[0] = queryPoint(0); s.q_pin[1] = queryPoint(1); s.q_pin[2] = queryPoint(2);
  // ONE C-ABI call per query; the row buffer grows (rarely: the RRTx ball shrinks as the
  // tree grows, drrt.cpp:16-27) when the call reports how many hits there are
  for (int attempt = 0; attempt < 2; attempt++) {
    int64_t total = 0;
    int rc = drrt_range_batch_into(s.h, 1, s.q_pin, nullptr, range, s.cap, s.off_pin, s.ids_pin,
                                   s.dist_pin, &total);
    if (rc == DRRT_ERR_CAPACITY && attempt == 0) {
      if (!s.reserve_row(total)) break;
      continue;
    }
    if (!ok(rc, "KDFindWithinRange")) return;
    for (int64_t i = 0; i < total; i++) AddToRangeList(S, s.by_id[s.ids_pin[i]], s.dist_pin[i]);
    return;
  }
  std::cout << "drrt_gpu: KDFindWithinRange: " << drrt_last_error() << std::endl;
}

// kdtree.cpp:822-851: same search, accumulating into the caller's list
void KDTree::KDFindMoreWithinRange(std::shared_ptr<JList> &L, double range,
                                   Eigen::VectorXd queryPoint) {
  KDFindWithinRange(L, range, queryPoint);
}

// kdtree.cpp:703-792: only ever entered with root == this->root (:805, :817)
bool KDTree::KDFindWithinRangeInSubtree(std::shared_ptr<KDTreeNode> &root, double range,
                                        Eigen::VectorXd queryPoint,
                                        std::shared_ptr<JList> &nodeList) {
  if (root != this->root) {
    std::cout << "drrt_gpu: subtree searches are not offered by the device back end" << std::endl;
    return false;
  }
  KDFindWithinRange(nodeList, range, queryPoint);
  return true;
}

// ---------------------------------------------------------------------------
// Batched entry points a GPU-aware caller can use instead of the per-point ones
// (declared in drrt_gpu_shim.h).
drrt_store *DrrtGpuHandle(KDTree *t) { return state_of(t).h; }

std::shared_ptr<KDTreeNode> DrrtGpuNode(KDTree *t, int32_t id) {
  ShimState &s = state_of(t);
  return (id >= 0 && (size_t)id < s.by_id.size()) ? s.by_id[id] : nullptr;
}

void DrrtGpuRelease(KDTree *t) {
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_state.find(t);
  if (it == g_state.end()) return;
  it->second.release();
  g_state.erase(it);
}

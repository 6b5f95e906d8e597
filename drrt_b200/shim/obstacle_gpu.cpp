// obstacle_gpu.cpp -- GPU versions of the reference's edge / node checks
// (src/obstacle.cpp:879-923, 1062-1143), forwarding to include/drrt_gpu.h.
#include "drrt_gpu_shim.h"

#include <iostream>

// Failure convention: the checks FAIL CLOSED.  When the device store is missing or a call
// fails (out of memory, a sticky CUDA fault, a stale obstacle table), every edge / node of the
// call is reported as colliding -- a planner that keeps running must never be told that an
// unchecked edge is free -- and the batch entry points return false.
namespace {
bool ok(int rc, const char *what) {
  if (rc == DRRT_OK) return true;
  std::cout << "drrt_gpu: " << what << ": " << drrt_last_error() << std::endl;
  return false;
}
}  // namespace

bool DrrtGpuUploadObstacles(KDTree *t, std::shared_ptr<ConfigSpace> &C) {
  drrt_store *h = DrrtGpuHandle(t);
  if (!h) return false;
  std::vector<int32_t> kind, vert_off(1, 0);
  std::vector<uint8_t> used;
  std::vector<double> life, origin, radius, verts;
  {
    std::lock_guard<std::mutex> lock(C->cspace_mutex_);  // obstacle.cpp:896
    std::shared_ptr<ListNode> n = C->obstacles_->front_;
    for (int i = 0; i < C->obstacles_->length_; i++, n = n->child_) {
      std::shared_ptr<Obstacle> &O = n->obstacle_;
      kind.push_back(O->kind_);
      used.push_back(O->obstacle_used_ ? 1 : 0);
      life.push_back(O->life_span_);
      origin.push_back(O->origin_(0));
      origin.push_back(O->origin_(1));
      radius.push_back(O->radius_);
      // the analytic checker reads shape_.GetPolygon() as world coordinates
      // (obstacle.cpp:785-794)
      Eigen::MatrixX2d poly = O->shape_.GetPolygon();
      for (int v = 0; v < poly.rows(); v++) {
        verts.push_back(poly(v, 0));
        verts.push_back(poly(v, 1));
      }
      vert_off.push_back((int32_t)(verts.size() / 2));
    }
  }
  return ok(drrt_obstacles_set(h, (int32_t)kind.size(), kind.data(), used.data(), life.data(),
                               origin.data(), radius.data(), vert_off.data(), verts.data()),
            "obstacle upload");
}

bool ExplicitEdgeCheckBatchGpu(std::shared_ptr<ConfigSpace> &C, KDTree *t,
                               std::vector<std::shared_ptr<Edge>> &edges,
                               std::vector<unsigned char> &unsafe) {
  drrt_store *h = DrrtGpuHandle(t);
  const size_t n = edges.size();
  unsafe.assign(n, 1);  // fail closed until the device has answered
  if (!h || n == 0) return h != nullptr;
  std::vector<double> s(3 * n), e(3 * n), len(n);
  for (size_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) {
      s[3 * i + k] = edges[i]->start_node_->position_(k);
      e[3 * i + k] = edges[i]->end_node_->position_(k);
    }
  if (!ok(drrt_edgecheck_batch(h, (int64_t)n, s.data(), e.data(), DRRT_EDGE_DUBINS,
                               C->robot_radius_, C->min_turn_radius_, C->in_warmup_time_ ? 1 : 0,
                               unsafe.data(), len.data()),
          "ExplicitEdgeCheck")) {
    unsafe.assign(n, 1);
    return false;
  }
  for (size_t i = 0; i < n; i++) {  // dubinsedge.cpp:726, 808, 829
    edges[i]->dist_ = len[i];
    edges[i]->w_dist_ = len[i];
    edges[i]->dist_original_ = len[i];
  }
  return true;
}

// Subset re-checks of the obstacle sweeps.  `take(obstacle)` decides, in list order (the
// order DrrtGpuUploadObstacles flattened), which obstacles an edge is tested against.
template <class Take>
static bool edge_check_subset(std::shared_ptr<ConfigSpace> &C, KDTree *t,
                              std::vector<std::shared_ptr<Edge>> &edges,
                              std::vector<unsigned char> &unsafe, Take take) {
  drrt_store *h = DrrtGpuHandle(t);
  const size_t n = edges.size();
  unsafe.assign(n, 1);  // fail closed until the device has answered
  if (!h || n == 0) return h != nullptr;
  std::vector<uint8_t> mask;
  {
    std::shared_ptr<ListNode> node = C->obstacles_->front_;
    for (int i = 0; i < C->obstacles_->length_; i++, node = node->child_)
      mask.push_back(take(node->obstacle_) ? 1 : 0);
  }
  std::vector<double> s(3 * n), e(3 * n);
  for (size_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) {
      s[3 * i + k] = edges[i]->start_node_->position_(k);
      e[3 * i + k] = edges[i]->end_node_->position_(k);
    }
  if (!ok(drrt_edgecheck_subset_batch(h, (int64_t)n, s.data(), e.data(), nullptr, nullptr,
                                      DRRT_EDGE_DUBINS, C->robot_radius_, C->min_turn_radius_,
                                      mask.data(), (int32_t)mask.size(), unsafe.data(), nullptr),
          "ExplicitEdgeCheck(obstacle)")) {
    unsafe.assign(n, 1);
    return false;
  }
  return true;
}

bool ExplicitEdgeCheckObstacleBatchGpu(std::shared_ptr<ConfigSpace> &C, KDTree *t,
                                       std::vector<std::shared_ptr<Edge>> &edges,
                                       std::shared_ptr<Obstacle> &O,
                                       std::vector<unsigned char> &unsafe) {
  // neighbor_edge->ExplicitEdgeCheck(O), obstacle.cpp:643, 651, 712
  return edge_check_subset(C, t, edges, unsafe,
                           [&](const std::shared_ptr<Obstacle> &other) { return other == O; });
}

bool ExplicitEdgeCheckOtherObstaclesBatchGpu(std::shared_ptr<ConfigSpace> &C, KDTree *t,
                                             std::vector<std::shared_ptr<Edge>> &edges,
                                             std::shared_ptr<Obstacle> &O, double time_elapsed,
                                             std::vector<unsigned char> &unsafe) {
  // RemoveObstacle's inner loop, obstacle.cpp:717-734
  return edge_check_subset(C, t, edges, unsafe, [&](const std::shared_ptr<Obstacle> &other) {
    return other != O && other->obstacle_used_ && other->start_time_ <= time_elapsed &&
           time_elapsed <= other->start_time_ + other->life_span_;
  });
}

bool ExplicitEdgeCheckGpu(std::shared_ptr<ConfigSpace> &C, KDTree *t, std::shared_ptr<Edge> &edge) {
  std::vector<std::shared_ptr<Edge>> one(1, edge);
  std::vector<unsigned char> unsafe;
  if (!ExplicitEdgeCheckBatchGpu(C, t, one, unsafe)) return true;  // unchecked == unsafe
  return unsafe.empty() || unsafe[0];
}

bool LineCheckGpu(std::shared_ptr<ConfigSpace> &C, KDTree *t, std::shared_ptr<KDTreeNode> node1,
                  std::shared_ptr<KDTreeNode> node2) {
  drrt_store *h = DrrtGpuHandle(t);
  if (!h) return true;  // unchecked == unsafe
  double s[3] = {node1->position_(0), node1->position_(1), node1->position_(2)};
  double e[3] = {node2->position_(0), node2->position_(1), node2->position_(2)};
  uint8_t unsafe = 1;
  if (!ok(drrt_edgecheck_batch(h, 1, s, e, DRRT_EDGE_STRAIGHT, C->robot_radius_, C->min_turn_radius_,
                               C->in_warmup_time_ ? 1 : 0, &unsafe, nullptr),
          "LineCheck"))
    return true;
  return unsafe != 0;
}

bool ExplicitNodeCheckGpu(std::shared_ptr<ConfigSpace> &C, KDTree *t,
                          std::shared_ptr<KDTreeNode> node) {
  if (C->in_warmup_time_) return false;  // obstacle.cpp:1065
  drrt_store *h = DrrtGpuHandle(t);
  if (!h) return true;  // unchecked == in collision
  double p[3] = {node->position_(0), node->position_(1), node->position_(2)};
  uint8_t hit = 1;
  if (!ok(drrt_pointcheck_batch(h, 1, p, C->robot_radius_, C->collision_distance_, &hit),
          "ExplicitNodeCheck"))
    return true;
  return hit != 0;
}

// drrt_gpu_shim.h -- host-side glue between the reference's C++ objects and the
// C ABI (include/drrt_gpu.h).  Include after <DRRT/drrt.h>.
//
//  * kdtree_gpu.cpp replaces src/kdtree.cpp wholesale (same KDTree methods).
//  * obstacle_gpu.cpp adds GPU versions of the edge / node checks next to the
//    reference's src/obstacle.cpp; INTEGRATION.md shows the call-site patch.
#ifndef DRRT_GPU_SHIM_H
#define DRRT_GPU_SHIM_H

#include <DRRT/drrt.h>

#include <memory>
#include <vector>

#include "drrt_gpu.h"

// the device store behind a KDTree (created lazily on first use), id -> node
drrt_store *DrrtGpuHandle(KDTree *t);
std::shared_ptr<KDTreeNode> DrrtGpuNode(KDTree *t, int32_t id);
void DrrtGpuRelease(KDTree *t);

// Flattens ConfigSpace::obstacles_ (front -> back) into the device table.  Call
// after every AddObstacle / RemoveObstacle / UpdateObstacles (obstacle.cpp:434-538).
bool DrrtGpuUploadObstacles(KDTree *t, std::shared_ptr<ConfigSpace> &C);

// ExplicitEdgeCheck(C, edge) obstacle.cpp:879-923 for one edge / a batch of
// edges given by their end nodes.  Also fills edge->dist_ as CalculateTrajectory
// would (dubinsedge.cpp:808).  Returns true / sets unsafe[i] when the edge collides.
bool ExplicitEdgeCheckGpu(std::shared_ptr<ConfigSpace> &C, KDTree *t, std::shared_ptr<Edge> &edge);
bool ExplicitEdgeCheckBatchGpu(std::shared_ptr<ConfigSpace> &C, KDTree *t,
                               std::vector<std::shared_ptr<Edge>> &edges,
                               std::vector<unsigned char> &unsafe);

// The re-checks of AddObstacle (obstacle.cpp:612-667) and RemoveObstacle (:669-754), for all
// the edges of the nodes in conflict at once: against O alone (`edge->ExplicitEdgeCheck(O)`,
// :643, :651, :712) and against every other used obstacle whose life window holds
// time_elapsed (:717-734).  The obstacle table must be the current list
// (DrrtGpuUploadObstacles after obstacle_used_ changes).  FindPointsInConflictWithObstacle
// (:540-610) itself needs no shim: it calls Tree->KDFindWithinRange.
bool ExplicitEdgeCheckObstacleBatchGpu(std::shared_ptr<ConfigSpace> &C, KDTree *t,
                                       std::vector<std::shared_ptr<Edge>> &edges,
                                       std::shared_ptr<Obstacle> &O,
                                       std::vector<unsigned char> &unsafe);
bool ExplicitEdgeCheckOtherObstaclesBatchGpu(std::shared_ptr<ConfigSpace> &C, KDTree *t,
                                             std::vector<std::shared_ptr<Edge>> &edges,
                                             std::shared_ptr<Obstacle> &O, double time_elapsed,
                                             std::vector<unsigned char> &unsafe);

// LineCheck obstacle.cpp:1092-1143 and ExplicitNodeCheck obstacle.cpp:1087-1090
bool LineCheckGpu(std::shared_ptr<ConfigSpace> &C, KDTree *t, std::shared_ptr<KDTreeNode> node1,
                  std::shared_ptr<KDTreeNode> node2);
bool ExplicitNodeCheckGpu(std::shared_ptr<ConfigSpace> &C, KDTree *t,
                          std::shared_ptr<KDTreeNode> node);

#endif

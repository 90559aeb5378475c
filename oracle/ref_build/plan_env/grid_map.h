// plan_env/grid_map.h — a STAND-IN.  path_searching/dyn_a_star.h (the base class of the hybrid A*) names GridMap only
// to call getInflateOccupancy; the real header pulls in ROS, PCL and OpenCV.  TEST INFRASTRUCTURE (oracle/_ref only).
#ifndef POP_GRID_MAP_STANDIN_H
#define POP_GRID_MAP_STANDIN_H
#include <Eigen/Eigen>
#include <memory>
class GridMap {
 public:
  typedef std::shared_ptr<GridMap> Ptr;
  int getInflateOccupancy(Eigen::Vector3d) { return 0; }
};
#endif

// pcl/filters/passthrough.h — a STAND-IN for pcl::PassThrough on one float field ("x", "y" or "z"): keeps the points whose
// field lies inside [min, max] (PCL: `if (value < min || value > max) drop`, non-finite values dropped; limits are
// floats), in input order; the output may alias the input (fake_particle_risk_voxel.cpp:95-104 filters in place).
// TEST INFRASTRUCTURE (oracle/_ref only).
#ifndef POP_PCL_PASSTHROUGH_STANDIN_H
#define POP_PCL_PASSTHROUGH_STANDIN_H
#include <pcl/point_cloud.h>
#include <cmath>
#include <string>
namespace pcl {
template <class P> class PassThrough {
  typename PointCloud<P>::ConstPtr in_;
  std::string field_ = "x";
  float lo_ = -3.4028235e38f, hi_ = 3.4028235e38f;
 public:
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { in_ = c; }
  void setFilterFieldName(const std::string& f) { field_ = f; }
  void setFilterLimits(const float& lo, const float& hi) { lo_ = lo; hi_ = hi; }
  void filter(PointCloud<P>& out) {
    std::vector<P> keep;
    keep.reserve(in_->points.size());
    for (const P& p : in_->points) {
      if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
      const float v = field_ == "x" ? p.x : (field_ == "y" ? p.y : p.z);
      if (v < lo_ || v > hi_) continue;
      keep.push_back(p);
    }
    out.points.swap(keep);
    out.width = static_cast<uint32_t>(out.points.size());
    out.height = 1;
  }
};
}  // namespace pcl
#endif

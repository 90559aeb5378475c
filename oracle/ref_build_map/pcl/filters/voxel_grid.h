// pcl/filters/voxel_grid.h — a STAND-IN for pcl::VoxelGrid<PointXYZ>: per-leaf centroid, leaves in ascending linear
// index (PCL sorts (leaf index, point index) pairs and averages each run; sums in input order, fp32 Eigen accumulators
// divided by the count).  Leaf coordinates floor(p * inverse_leaf) - floor(min * inverse_leaf) as in
// voxel_grid.hpp.  MapBase::filterPointCloud (map.cpp:107-132) calls it; the repo's A2 oracle restates the same rule.
// TEST INFRASTRUCTURE (oracle/_ref only): a restatement written here, not PCL.
#ifndef POP_PCL_VOXEL_GRID_STANDIN_H
#define POP_PCL_VOXEL_GRID_STANDIN_H
#include <pcl/point_cloud.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
namespace pcl {
template <class P> class VoxelGrid {
  typename PointCloud<P>::ConstPtr in_;
  float leaf_[3] = {0, 0, 0};
 public:
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { in_ = c; }
  void setLeafSize(float x, float y, float z) { leaf_[0] = x; leaf_[1] = y; leaf_[2] = z; }
  void filter(PointCloud<P>& out) {
    out.points.clear();
    const std::vector<P>& pts = in_->points;
    std::vector<int> ok;
    float mn[3] = {3.4028235e38f, 3.4028235e38f, 3.4028235e38f}, mx[3] = {-3.4028235e38f, -3.4028235e38f, -3.4028235e38f};
    for (int i = 0; i < static_cast<int>(pts.size()); i++) {
      const float v[3] = {pts[i].x, pts[i].y, pts[i].z};
      if (!std::isfinite(v[0]) || !std::isfinite(v[1]) || !std::isfinite(v[2])) continue;
      ok.push_back(i);
      for (int k = 0; k < 3; k++) { mn[k] = std::min(mn[k], v[k]); mx[k] = std::max(mx[k], v[k]); }
    }
    if (ok.empty()) return;
    const float inv[3] = {1.0f / leaf_[0], 1.0f / leaf_[1], 1.0f / leaf_[2]};
    int64_t lo[3], div[3];
    for (int k = 0; k < 3; k++) {
      lo[k] = static_cast<int64_t>(std::floor(mn[k] * inv[k]));
      div[k] = static_cast<int64_t>(std::floor(mx[k] * inv[k])) - lo[k] + 1;
    }
    std::vector<std::pair<int64_t, int>> key;
    key.reserve(ok.size());
    for (int i : ok) {
      const int64_t ix = static_cast<int64_t>(std::floor(pts[i].x * inv[0])) - lo[0];
      const int64_t iy = static_cast<int64_t>(std::floor(pts[i].y * inv[1])) - lo[1];
      const int64_t iz = static_cast<int64_t>(std::floor(pts[i].z * inv[2])) - lo[2];
      key.emplace_back(ix + iy * div[0] + iz * div[0] * div[1], i);
    }
    std::stable_sort(key.begin(), key.end(), [](const std::pair<int64_t, int>& a, const std::pair<int64_t, int>& b) { return a.first < b.first; });
    for (size_t s = 0; s < key.size();) {
      size_t e = s;
      float cx = 0, cy = 0, cz = 0;
      while (e < key.size() && key[e].first == key[s].first) { cx += pts[key[e].second].x; cy += pts[key[e].second].y; cz += pts[key[e].second].z; e++; }
      const float n = static_cast<float>(e - s);
      out.points.push_back(P(cx / n, cy / n, cz / n));
      s = e;
    }
    out.width = static_cast<uint32_t>(out.points.size());
    out.height = 1;
  }
};
}  // namespace pcl
#endif

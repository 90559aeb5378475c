// pcl/segmentation/extract_clusters.h — a STAND-IN for pcl::search::KdTree + pcl::EuclideanClusterExtraction (PCL and
// FLANN are absent).  TEST INFRASTRUCTURE (oracle/_ref only); restated from PCL 1.10's extractEuclideanClusters:
// for every unprocessed point in index order, grow a cluster by breadth-first radius searches; keep it when its size
// lies in [min, max]; its indices sorted ascending; finally std::sort(clusters.rbegin(), clusters.rend(), by size) —
// called here exactly like that, so that ties fall as libstdc++'s std::sort leaves them.  The radius search is brute
// force with FLANN's L2_Simple distance (fp32 sum of squared differences in x, y, z order) and its RadiusResultSet
// acceptance `dist < radius * radius` (fp32 product).  The libraries themselves are NOT pinned by this.
#ifndef POP_PCL_EXTRACT_CLUSTERS_STANDIN_H
#define POP_PCL_EXTRACT_CLUSTERS_STANDIN_H
#include <pcl/point_cloud.h>
#include <algorithm>
#include <memory>
#include <vector>
namespace pcl {
struct PointIndices { std::vector<int> indices; };
namespace search {
template <class P> class KdTree {
 public:
  typedef std::shared_ptr<KdTree<P>> Ptr;
  typename PointCloud<P>::ConstPtr cloud_;
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { cloud_ = c; }
  int radiusSearch(int index, double radius, std::vector<int>& idx, std::vector<float>& d2) const {
    idx.clear(); d2.clear();
    const float r2 = static_cast<float>(radius * radius);
    const P& q = cloud_->points[index];
    for (int j = 0; j < static_cast<int>(cloud_->points.size()); j++) {
      const P& p = cloud_->points[j];
      const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
      float d = 0.f;
      d += dx * dx; d += dy * dy; d += dz * dz;
      if (d < r2) { idx.push_back(j); d2.push_back(d); }
    }
    return static_cast<int>(idx.size());
  }
};
}  // namespace search
template <class P> class EuclideanClusterExtraction {
  double tol_ = 0.0;
  int min_ = 1, max_ = 1 << 30;
  typename search::KdTree<P>::Ptr tree_;
  typename PointCloud<P>::ConstPtr cloud_;
 public:
  void setClusterTolerance(double t) { tol_ = t; }
  void setMinClusterSize(int n) { min_ = n; }
  void setMaxClusterSize(int n) { max_ = n; }
  void setSearchMethod(const typename search::KdTree<P>::Ptr& t) { tree_ = t; }
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { cloud_ = c; }
  void extract(std::vector<PointIndices>& clusters) {
    clusters.clear();
    const int n = static_cast<int>(cloud_->points.size());
    std::vector<char> processed(n, 0);
    std::vector<int> nn;
    std::vector<float> nd;
    for (int i = 0; i < n; i++) {
      if (processed[i]) continue;
      std::vector<int> seed;
      size_t sq = 0;
      seed.push_back(i);
      processed[i] = 1;
      while (sq < seed.size()) {
        if (!tree_->radiusSearch(seed[sq], tol_, nn, nd)) { sq++; continue; }
        for (size_t j = 0; j < nn.size(); j++) {
          if (nn[j] == -1 || processed[nn[j]]) continue;
          seed.push_back(nn[j]);
          processed[nn[j]] = 1;
        }
        sq++;
      }
      if (static_cast<int>(seed.size()) >= min_ && static_cast<int>(seed.size()) <= max_) {
        PointIndices r;
        r.indices = seed;
        std::sort(r.indices.begin(), r.indices.end());
        r.indices.erase(std::unique(r.indices.begin(), r.indices.end()), r.indices.end());
        clusters.push_back(r);
      }
    }
    std::sort(clusters.rbegin(), clusters.rend(),
              [](const PointIndices& a, const PointIndices& b) { return a.indices.size() < b.indices.size(); });
  }
};
}  // namespace pcl
#endif

// ref_glue_gridmap.cpp — TEST INFRASTRUCTURE: the reference's log-odds raycast map, plan_env/src/grid_map.cpp, compiled
// UNMODIFIED (together with its header) against the stand-ins of ref_build_gridmap/, ref_build_map/ and ref_build/ into
// oracle/_ref/libref_gridmap.so, with a few C entry points around it.  The ROS wiring of GridMap::initMap is inert;
// a frame is what GridMap::updateOccupancyCallback does after its bookkeeping (grid_map.cpp:606-616): projectDepthImage,
// raycastProcess, clearAndInflateLocalMap — on a depth image, camera rotation and camera position the test supplies
// (what depthPoseCallback / depthOdomCallback would have stored).  plan_env/src/raycast.cpp is linked in as it is.
#define private public
#define protected public
#include <grid_map.cpp>
#undef private
#undef protected
#include <raycast.cpp>

#include <cstring>

extern "C" {

struct refgm_params {
  double resolution, map_size[3], local_update_range[3], obstacles_inflation;
  double p_hit, p_miss, p_min, p_max, p_occ, min_ray_length, max_ray_length, virtual_ceil_height, ground_height;
  double fx, fy, cx, cy, k_depth_scaling_factor, depth_filter_mindist, depth_filter_maxdist, depth_filter_tolerance;
  int local_map_margin, use_depth_filter, depth_filter_margin, skip_pixel;
};

void* refgm_create(const refgm_params* p) {
  ros::NodeHandle nh;
  auto& v = nh.values;
  v["grid_map/resolution"] = p->resolution;
  v["grid_map/map_size_x"] = p->map_size[0]; v["grid_map/map_size_y"] = p->map_size[1]; v["grid_map/map_size_z"] = p->map_size[2];
  v["grid_map/local_update_range_x"] = p->local_update_range[0];
  v["grid_map/local_update_range_y"] = p->local_update_range[1];
  v["grid_map/local_update_range_z"] = p->local_update_range[2];
  v["grid_map/obstacles_inflation"] = p->obstacles_inflation;
  v["grid_map/fx"] = p->fx; v["grid_map/fy"] = p->fy; v["grid_map/cx"] = p->cx; v["grid_map/cy"] = p->cy;
  v["grid_map/use_depth_filter"] = p->use_depth_filter;
  v["grid_map/depth_filter_tolerance"] = p->depth_filter_tolerance;
  v["grid_map/depth_filter_maxdist"] = p->depth_filter_maxdist;
  v["grid_map/depth_filter_mindist"] = p->depth_filter_mindist;
  v["grid_map/depth_filter_margin"] = p->depth_filter_margin;
  v["grid_map/k_depth_scaling_factor"] = p->k_depth_scaling_factor;
  v["grid_map/skip_pixel"] = p->skip_pixel;
  v["grid_map/p_hit"] = p->p_hit; v["grid_map/p_miss"] = p->p_miss; v["grid_map/p_min"] = p->p_min;
  v["grid_map/p_max"] = p->p_max; v["grid_map/p_occ"] = p->p_occ;
  v["grid_map/min_ray_length"] = p->min_ray_length; v["grid_map/max_ray_length"] = p->max_ray_length;
  v["grid_map/visualization_truncate_height"] = 100.0;
  v["grid_map/virtual_ceil_height"] = p->virtual_ceil_height;
  v["grid_map/virtual_ceil_yp"] = 1e9; v["grid_map/virtual_ceil_yn"] = -1e9;
  v["grid_map/local_map_margin"] = p->local_map_margin;
  v["grid_map/ground_height"] = p->ground_height;
  v["grid_map/pose_type"] = 1;
  GridMap* g = new GridMap();
  g->initMap(nh);
  return g;
}
void refgm_destroy(void* h) { delete static_cast<GridMap*>(h); }
void refgm_dims(void* h, int* out3) {
  GridMap* g = static_cast<GridMap*>(h);
  for (int k = 0; k < 3; k++) out3[k] = g->mp_.map_voxel_num_(k);
}
// what projectDepthImage alone leaves in proj_points_ (its first-frame skip included); returns the count
int refgm_project(void* h, const uint16_t* img, int rows, int cols, const double* R, const double* cam, double* out, int cap) {
  GridMap* g = static_cast<GridMap*>(h);
  g->md_.depth_image_.create(rows, cols, CV_16UC1);
  std::memcpy(g->md_.depth_image_.ptr<uint16_t>(0), img, sizeof(uint16_t) * rows * cols);
  for (int i = 0; i < 3; i++) {
    g->md_.camera_pos_(i) = cam[i];
    for (int j = 0; j < 3; j++) g->md_.camera_r_m_(i, j) = R[i * 3 + j];
  }
  g->projectDepthImage();
  const int n = g->md_.proj_points_cnt;
  for (int i = 0; i < n && i < cap; i++)
    for (int k = 0; k < 3; k++) out[i * 3 + k] = g->md_.proj_points_[i](k);
  return n;
}
// raycastProcess + clearAndInflateLocalMap on the points projectDepthImage left (or on points given here)
void refgm_raycast(void* h, const double* pts, int n, const double* cam) {
  GridMap* g = static_cast<GridMap*>(h);
  if (pts) {
    if (static_cast<int>(g->md_.proj_points_.size()) < n) g->md_.proj_points_.resize(n);
    for (int i = 0; i < n; i++) g->md_.proj_points_[i] = Eigen::Vector3d(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]);
    g->md_.proj_points_cnt = n;
    for (int k = 0; k < 3; k++) g->md_.camera_pos_(k) = cam[k];
  }
  g->raycastProcess();
  if (g->md_.local_updated_) g->clearAndInflateLocalMap();
  g->md_.occ_need_update_ = false;
  g->md_.local_updated_ = false;
}
void refgm_download(void* h, double* occ, signed char* inflate, int* bounds6) {
  GridMap* g = static_cast<GridMap*>(h);
  const size_t n = g->md_.occupancy_buffer_.size();
  for (size_t i = 0; i < n; i++) { occ[i] = g->md_.occupancy_buffer_[i]; inflate[i] = g->md_.occupancy_buffer_inflate_[i]; }
  for (int k = 0; k < 3; k++) { bounds6[k] = g->md_.local_bound_min_(k); bounds6[3 + k] = g->md_.local_bound_max_(k); }
}
int refgm_query(void* h, const double* pos, int inflate) {
  GridMap* g = static_cast<GridMap*>(h);
  const Eigen::Vector3d p(pos[0], pos[1], pos[2]);
  return inflate ? g->getInflateOccupancy(p) : g->getOccupancy(p);
}
}  // extern "C"

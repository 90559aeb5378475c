// ref_glue_dsp.cpp — the reference's particle-map node in oracle/_ref/libref_dsp_*.so.  TEST INFRASTRUCTURE; see ref_glue.cpp.
//
// plan_env/src/risk_mapping_node.cpp (filterPointCloud, odomCallback, cloudCallback, publishMap, and main()'s
// configuration of the map) and plan_env/include/plan_env/dsp_dynamic.h (dsp_map::DSPMap: update, mapPrediction,
// mapUpdate, velocityEstimationThread, mapAddNewBornParticlesByObservation, mapOccupancyCalculationAndResample,
// getOccupancyMapWithFutureStatus) are compiled UNMODIFIED: the node's source file is #included below as it lies under
// /root/reference, because the map keeps its state in file-static arrays (voxels_with_particle, pyramids_in_fov, the
// Gaussian tables, ...) that only the translation unit that includes the header can read.  ROS is the inert stand-in
// of oracle/ref_build/ros/ros.h (the node's main() is run for its parameter reads and setter calls; spinning returns
// at once; callbacks are called directly); PCL's VoxelGrid / KdTree / EuclideanClusterExtraction and munkres-cpp are
// the restatements under oracle/ref_build_map (absent third-party libraries: those three are NOT pinned by this build,
// everything the reference itself computes around them is).
//
// Non-determinism of the reference made an input, without touching its source: the noise tables are file-static
// arrays filled from default_random_engine(time(NULL)) — they are overwritten after construction; rand(), srand() and
// time() are defined in this library (linked -Bsymbolic-functions), rand() reading the injected stream.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <sstream>
#include <vector>

#include <plan_env/map_parameters.h>   // the wrapper of oracle/ref_build_map: grid and table size of this build

static std::vector<int32_t> g_rand_stream;
static size_t g_rand_seq = 0;
extern "C" int rand(void) noexcept {
  if (g_rand_stream.empty()) return 0;
  const int r = g_rand_stream[g_rand_seq];
  if (++g_rand_seq >= g_rand_stream.size()) g_rand_seq = 0;
  return r;
}
extern "C" void srand(unsigned) noexcept {}
extern "C" time_t time(time_t* t) noexcept { if (t) *t = 1; return 1; }

#define private public
#define protected public
#include <risk_mapping_node.cpp>   // -I$(REF)/plan_env/src: the reference's file, as is
#undef private
#undef protected

namespace {
struct Quiet {
  std::ostringstream sink;
  std::streambuf* old;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

// out: nx, ny, nz, nt, GAUSSIAN_RANDOMS_NUM, SAFE_PARTICLE_NUM_VOXEL, SAFE_PARTICLE_NUM_PYRAMID, observation_pyramid_num
void refdsp_dims(int32_t* out, double* res) {
  out[0] = MAP_LENGTH_VOXEL_NUM; out[1] = MAP_WIDTH_VOXEL_NUM; out[2] = MAP_HEIGHT_VOXEL_NUM; out[3] = PREDICTION_TIMES;
  out[4] = GAUSSIAN_RANDOMS_NUM; out[5] = SAFE_PARTICLE_NUM_VOXEL; out[6] = SAFE_PARTICLE_NUM_PYRAMID;
  out[7] = observation_pyramid_num;
  *res = VOXEL_RESOLUTION;
}

// The node's own start-up: main() reads its ROS parameters ("map/resolution", "map/risk_threshold_voxel",
// "map/sigma_observation", "map/sigma_localization") and configures dsp_map_ (risk_mapping_node.cpp:214-260).
void refdsp_start(float filter_res, float risk_threshold, float sigma_observation, float sigma_localization) {
  Quiet q;
  std::map<std::string, double>& g = ros::NodeHandle::global_values();
  g["map/resolution"] = filter_res;
  g["map/risk_threshold_voxel"] = risk_threshold;
  g["map/sigma_observation"] = sigma_observation;
  g["map/sigma_localization"] = sigma_localization;
  int argc = 0;
  main(argc, nullptr);
}

// p_gaussian_randoms, v_gaussian_randoms (n = GAUSSIAN_RANDOMS_NUM entries each) and the rand() stream; sequence
// counters back to 0
void refdsp_set_random_tables(const float* pg, const float* vg, const int32_t* ur, int n) {
  std::memcpy(dsp_map::p_gaussian_randoms, pg, sizeof(float) * GAUSSIAN_RANDOMS_NUM);
  std::memcpy(dsp_map::v_gaussian_randoms, vg, sizeof(float) * GAUSSIAN_RANDOMS_NUM);
  g_rand_stream.assign(ur, ur + n);
  g_rand_seq = 0;
  dsp_map_.position_gaussian_random_seq = 0;
  dsp_map_.velocity_gaussian_random_seq = 0;
  dsp_map_.localization_gaussian_random_seq = 0;
}

// "map/odom": odomCallback (risk_mapping_node.cpp:203-211); the message carries doubles
void refdsp_odom(const double* pos, const double* quat_wxyz) {
  nav_msgs::Odometry m;
  m.pose.pose.position.x = pos[0]; m.pose.pose.position.y = pos[1]; m.pose.pose.position.z = pos[2];
  m.pose.pose.orientation.w = quat_wxyz[0]; m.pose.pose.orientation.x = quat_wxyz[1];
  m.pose.pose.orientation.y = quat_wxyz[2]; m.pose.pose.orientation.z = quat_wxyz[3];
  odomCallback(m);
}

// "map/cloud": cloudCallback (:155-188) = filterPointCloud + DSPMap::update + publishMap.  Returns 1 when the map was
// updated (update() accepted the odometry), else 0.  n_valid = points that passed the filter.
int refdsp_cloud(const float* cam, int n, double stamp, int32_t* n_valid) {
  Quiet q;
  const int before = dsp_map_.update_counter;
  std::shared_ptr<sensor_msgs::PointCloud2> m(new sensor_msgs::PointCloud2());
  m->header.stamp = ros::Time(stamp);
  m->xyz.assign(cam, cam + 3 * static_cast<size_t>(n));
  cloudCallback(m);
  // valid_clouds_ holds the filtered body-frame points of this frame; their count is not kept by the node
  if (n_valid) *n_valid = -1;
  return dsp_map_.update_counter != before ? 1 : 0;
}

// filterPointCloud alone (:74-99): out = 5000 x 3
int refdsp_filter(const float* cam, int n, float* out) {
  pcl::PointCloud<pcl::PointXYZ>::Ptr in(new pcl::PointCloud<pcl::PointXYZ>()), outc(new pcl::PointCloud<pcl::PointXYZ>());
  for (int i = 0; i < n; i++) in->points.push_back(pcl::PointXYZ(cam[3 * i], cam[3 * i + 1], cam[3 * i + 2]));
  int nv = 0;
  filterPointCloud(in, outc, out, nv);
  return nv;
}

// DSPMap::update on filtered body-frame points, with the node's current odometry
int refdsp_update_filtered(const float* pts, int n, double stamp) {
  Quiet q;
  std::vector<float> buf(pts, pts + 3 * static_cast<size_t>(n));
  buf.resize(std::max<size_t>(buf.size(), 3));
  return dsp_map_.update(n, 3, buf.data(), pose_.x(), pose_.y(), pose_.z(), stamp, q_.w(), q_.x(), q_.y(), q_.z());
}

// risk_maps_ as publishMap left it ([V][T] future risk), i.e. what "map/future_risk" carries
void refdsp_future_risk(float* out) { std::memcpy(out, &risk_maps_[0][0], sizeof(risk_maps_)); }

// state views in the layouts of tests/oracle_lib.py (OracleDsp)
void refdsp_particles(float* out) {   // [V][S][8]: flag vx vy vz px py pz weight (column 8, the time stamp, dropped)
  for (int i = 0; i < VOXEL_NUM; i++)
    for (int j = 0; j < SAFE_PARTICLE_NUM_VOXEL; j++)
      for (int k = 0; k < 8; k++) out[(static_cast<size_t>(i) * SAFE_PARTICLE_NUM_VOXEL + j) * 8 + k] = dsp_map::voxels_with_particle[i][j][k];
}
void refdsp_voxels(float* out) {   // [V][4]
  for (int i = 0; i < VOXEL_NUM; i++) for (int k = 0; k < 4; k++) out[static_cast<size_t>(i) * 4 + k] = dsp_map::voxels_objects_number[i][k];
}
void refdsp_accumulators(float* out) {   // [V][T]: voxels_objects_number[.][4..] before the export clears them
  for (int i = 0; i < VOXEL_NUM; i++)
    for (int k = 0; k < PREDICTION_TIMES; k++) out[static_cast<size_t>(i) * PREDICTION_TIMES + k] = dsp_map::voxels_objects_number[i][4 + k];
}
void refdsp_observations(float* pc, int32_t* counts, float* maxlen) {
  std::memcpy(pc, &dsp_map_.point_cloud[0][0][0], sizeof(dsp_map_.point_cloud));
  for (int i = 0; i < observation_pyramid_num; i++) { counts[i] = dsp_map_.observation_num_each_pyramid[i]; maxlen[i] = dsp_map_.point_cloud_max_length[i]; }
}
void refdsp_pyramids(int32_t* out) { std::memcpy(out, &dsp_map::pyramids_in_fov[0][0][0], sizeof(dsp_map::pyramids_in_fov)); }
int refdsp_points_with_velocity(float* out, int cap) {
  const std::vector<pcl::PointXYZINormal>& p = dsp_map_.input_cloud_with_velocity->points;
  for (int i = 0; i < static_cast<int>(p.size()) && i < cap; i++) {
    float* o = out + 7 * i;
    o[0] = p[i].x; o[1] = p[i].y; o[2] = p[i].z; o[3] = p[i].intensity; o[4] = p[i].normal_x; o[5] = p[i].normal_y; o[6] = p[i].normal_z;
  }
  return static_cast<int>(p.size());
}
void refdsp_stats(int32_t* out) {   // sequence counters of the three tables, update counter
  out[0] = dsp_map_.position_gaussian_random_seq; out[1] = dsp_map_.velocity_gaussian_random_seq;
  out[2] = static_cast<int32_t>(g_rand_seq); out[3] = dsp_map_.update_counter;
}
// getOccupancyMap (dsp_dynamic.h:402-420) centres of the voxels over thr — NOTE: it clears the accumulators
int refdsp_occupied(float thr, float* out, int cap) {
  int n = 0;
  pcl::PointCloud<pcl::PointXYZ> cloud;
  dsp_map_.getOccupancyMap(n, cloud, thr);
  for (int i = 0; i < n && i < cap; i++) { out[3 * i] = cloud.points[i].x; out[3 * i + 1] = cloud.points[i].y; out[3 * i + 2] = cloud.points[i].z; }
  return n;
}

}  // extern "C"

// ref_glue_map.cpp — the reference's map classes in oracle/_ref/libref_map_*.so.  TEST INFRASTRUCTURE; see ref_glue.cpp.
//
// plan_env/src/map.cpp (MapBase: voxel index math of map.h:153-215, getInflateOccupancy, getObstaclePoints,
// filterPointCloud, pose callbacks) and, per build, plan_env/src/fake_particle_risk_voxel.cpp (REF_MAP_KIND 0: the
// ground-truth SOGM producer FakeParticleRiskVoxel: updateMap, groundTruthStateCallback, addParticlesToRiskMap,
// getClearOcccupancy) or plan_env/src/risk_base.cpp (REF_MAP_KIND 1: RiskBase: futureRiskCallback, addOtherAgents,
// getClearOcccupancy, getObstaclePoints) are compiled UNMODIFIED from /root/reference, together with the swarm
// coordinator (traj_coordinator/src/particles.cpp) and traj_utils/src/bernstein.cpp they call.  ROS, PCL and Eigen are
// the stand-ins of oracle/ref_build and oracle/ref_build_map (inert topics; PassThrough / VoxelGrid / fp32 Eigen vectors,
// quaternions and planes restated there).  The grid size is a compile-time constant of the reference
// (map_parameters.h), so there is one library per size (oracle/Makefile).  This file only marshals flat arrays into
// the reference's message and argument types and calls the reference's callbacks.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <pthread.h>
#include <sstream>

// ParticleATC::getParticlesWithRisk seeds its resampling engine with time(NULL) (particles.cpp:382): the library answers
// that call itself (linked -Bsymbolic-functions), so the seed is an input of the test
#include <ctime>
static time_t g_fake_time = 1;
extern "C" time_t time(time_t* t) noexcept { if (t) *t = g_fake_time; return g_fake_time; }
extern "C" void refmap_set_time_seed(long s) { g_fake_time = static_cast<time_t>(s); }

#define private public
#define protected public
#if REF_MAP_KIND == 0
#include <plan_env/fake_particle_risk_voxel.h>
typedef FakeParticleRiskVoxel RefMap;
#else
#include <plan_env/risk_base.h>
typedef RiskBase RefMap;
#endif
#undef private
#undef protected
#include <plan_env/map_parameters.h>   // after the class header, as fake_particle_risk_voxel.cpp:12-13 includes it

#include "../include/pop.h"

extern "C" void ref_feed_atc(ParticleATC& atc, int n, int ego_id, const double* offs, int max_body, const int32_t* counts,
                             const pop_traj_record* recs, double now);   // ref_glue_particles.cpp

namespace {
struct Quiet {   // the reference prints from its hot loops; the text is discarded
  std::ostringstream sink;
  std::streambuf* old;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};
struct Handle {
  RefMap* map = nullptr;
  ros::NodeHandle nh;
  std::shared_ptr<ParticleATC> atc;
};
// updateMap keeps a float[VOXEL_NUM][PREDICTION_TIMES] on the stack (fake_particle_risk_voxel.cpp:106): run it on a
// thread whose stack holds it
struct Job { void (*fn)(void*); void* arg; };
void* job_main(void* p) { Job* j = static_cast<Job*>(p); j->fn(j->arg); return nullptr; }
void run_on_big_stack(void (*fn)(void*), void* arg) {
  pthread_attr_t at;
  pthread_attr_init(&at);
  pthread_attr_setstacksize(&at, sizeof(float) * static_cast<size_t>(VOXEL_NUM) * PREDICTION_TIMES + (64u << 20));
  pthread_t th;
  Job j{fn, arg};
  pthread_create(&th, &at, job_main, &j);
  pthread_join(th, nullptr);
  pthread_attr_destroy(&at);
}
}  // namespace

extern "C" {

// out: nx, ny, nz, nt, kind (0 FakeParticleRiskVoxel, 1 RiskBase); *res = VOXEL_RESOLUTION
void refmap_dims(int32_t* out, double* res) {
  out[0] = MAP_LENGTH_VOXEL_NUM; out[1] = MAP_WIDTH_VOXEL_NUM; out[2] = MAP_HEIGHT_VOXEL_NUM; out[3] = PREDICTION_TIMES;
  out[4] = REF_MAP_KIND;
  *res = VOXEL_RESOLUTION;
}

// params: time_resolution, risk_threshold_voxel, clearance, ground_height, ceiling_height, risk_threshold_region,
// risk_threshold_region_decay, risk_threshold_voxel_decay — the ROS parameters of map.cpp:15-40 and risk_base.cpp:21-23.
// init() is the reference's own (fake_particle_risk_voxel.cpp:21-73 / risk_base.cpp:16-58): it derives the half ranges
// and the collision footprint.
void* refmap_create(const double* params) {
  Quiet q;
  Handle* h = new Handle();
  h->nh.values["map/time_resolution"] = params[0];
  h->nh.values["map/risk_threshold_voxel"] = params[1];
  h->nh.values["map/clearance"] = params[2];
  h->nh.values["map/ground_height"] = params[3];
  h->nh.values["map/ceiling_height"] = params[4];
  h->nh.values["map/risk_threshold_region"] = params[5];
  h->nh.values["map/risk_threshold_region_decay"] = params[6];
  h->nh.values["map/risk_threshold_voxel_decay"] = params[7];
  h->map = new RefMap();
  // `new T()` leaves risk_maps_ indeterminate; start from zeros as the oracle and libpop do
  std::memset(reinterpret_cast<void*>(&h->map->MapBase::risk_maps_[0][0]), 0, sizeof(h->map->MapBase::risk_maps_));
  std::memset(reinterpret_cast<void*>(&h->map->risk_maps_[0][0]), 0, sizeof(h->map->risk_maps_));
  h->map->init(h->nh);
  return h;
}
void refmap_destroy(void* hv) { Handle* h = static_cast<Handle*>(hv); delete h->map; delete h; }

// the collision footprint init() built: returns the count, offsets to out (cap x 3)
int refmap_kernel(void* hv, int32_t* out, int cap) {
  Handle* h = static_cast<Handle*>(hv);
  const int n = static_cast<int>(h->map->inflate_kernel_.size());
  for (int i = 0; i < n && i < cap; i++) for (int k = 0; k < 3; k++) out[3 * i + k] = h->map->inflate_kernel_[i][k];
  return n;
}

// MapBase::odomCallback (map.cpp:158-170) with the pose as doubles, as nav_msgs/Odometry carries it
void refmap_set_pose(void* hv, const double* pos) {
  Handle* h = static_cast<Handle*>(hv);
  std::shared_ptr<nav_msgs::Odometry> m(new nav_msgs::Odometry());
  m->pose.pose.position.x = pos[0]; m->pose.pose.position.y = pos[1]; m->pose.pose.position.z = pos[2];
  h->map->odomCallback(m);
}

// A coordinator for drone ego_id in a swarm of n that has received every peer's body particles and trajectory
// (see ref_particles_with_risk in ref_glue_particles.cpp), attached with setCoordinator.
void refmap_set_swarm(void* hv, int n, int ego_id, const double* size, float replan_risk_rate, int num_resample,
                      const double* offs, int max_body, const int32_t* counts, const pop_traj_record* recs, double now) {
  Quiet q;
  Handle* h = static_cast<Handle*>(hv);
  ros::Time::set_now(now);
  ros::NodeHandle nh;
  nh.values["drone_id"] = ego_id;
  nh.values["swarm/num_robots"] = n;
  nh.values["swarm/drone_size_x"] = size[0]; nh.values["swarm/drone_size_y"] = size[1]; nh.values["swarm/drone_size_z"] = size[2];
  nh.values["swarm/replan_risk_rate"] = replan_risk_rate;
  nh.values["swarm/num_resample"] = num_resample;
  h->atc.reset(new ParticleATC(nh));
  h->atc->init();
  ref_feed_atc(*h->atc, n, ego_id, offs, max_body, counts, recs, now);
  h->map->setCoordinator(h->atc);
}

#if REF_MAP_KIND == 0
// groundTruthStateCallback (fake_particle_risk_voxel.cpp:244-264) from a MarkerArray built as
// dynamic_forest_seq_sensing.cpp:185-203 builds it: scale.x = width, points[0].z = height, points[1] - points[0] = velocity
void refmap_set_cylinders(void* hv, const pop_cylinder* cyl, int n) {
  Handle* h = static_cast<Handle*>(hv);
  std::shared_ptr<visualization_msgs::MarkerArray> m(new visualization_msgs::MarkerArray());
  m->markers.resize(n);
  for (int i = 0; i < n; i++) {
    visualization_msgs::Marker& mk = m->markers[i];
    mk.id = i; mk.type = cyl[i].type;
    mk.pose.position.x = cyl[i].x; mk.pose.position.y = cyl[i].y; mk.pose.position.z = cyl[i].z;
    mk.pose.orientation.w = cyl[i].qw; mk.pose.orientation.x = cyl[i].qx; mk.pose.orientation.y = cyl[i].qy; mk.pose.orientation.z = cyl[i].qz;
    mk.scale.x = cyl[i].w;
    mk.points.resize(2);
    mk.points[0].x = 0.0; mk.points[0].y = 0.0; mk.points[0].z = cyl[i].h;
    mk.points[1].x = cyl[i].vx; mk.points[1].y = cyl[i].vy; mk.points[1].z = cyl[i].h;
  }
  h->map->groundTruthStateCallback(m);
}

struct UpdateArgs { Handle* h; const float* pts; int n; };
static void do_update(void* p) {
  UpdateArgs* a = static_cast<UpdateArgs*>(p);
  std::shared_ptr<sensor_msgs::PointCloud2> m(new sensor_msgs::PointCloud2());
  m->xyz.assign(a->pts, a->pts + 3 * static_cast<size_t>(a->n));
  a->h->map->cloudCallback(m);   // MapBase::cloudCallback -> FakeParticleRiskVoxel::updateMap (virtual)
}
// "map/cloud" arriving at wall-clock time `now` (which becomes last_update_time_, fake_particle_risk_voxel.cpp:82)
void refmap_update(void* hv, const float* pts, int n, double now) {
  Quiet q;
  ros::Time::set_now(now);
  UpdateArgs a{static_cast<Handle*>(hv), pts, n};
  run_on_big_stack(do_update, &a);
}
#else
// "map/time" then "map/future_risk" (risk_mapping_node.cpp:105-150 publishes [V][T] floats + 4 trailing values with
// stride T): updateTimeCallback + futureRiskCallback (risk_base.cpp:60-80), which projects the peers when a coordinator is set
void refmap_future_risk(void* hv, const float* risk, double stamp, double now) {
  Quiet q;
  Handle* h = static_cast<Handle*>(hv);
  ros::Time::set_now(now);
  std_msgs::Header hd;
  hd.stamp = ros::Time(stamp);
  h->map->updateTimeCallback(hd);
  std::shared_ptr<std_msgs::Float32MultiArray> m(new std_msgs::Float32MultiArray());
  m->layout.dim.resize(1);
  m->layout.dim[0].size = VOXEL_NUM; m->layout.dim[0].stride = PREDICTION_TIMES;
  m->data.assign(risk, risk + static_cast<size_t>(VOXEL_NUM) * PREDICTION_TIMES);
  for (int k = 0; k < 4; k++) m->data.push_back(0.0f);
  h->map->futureRiskCallback(m);
}
#endif

// risk_maps_ of the class under test ([VOXEL_NUM][PREDICTION_TIMES] floats)
void refmap_download(void* hv, float* out) {
  Handle* h = static_cast<Handle*>(hv);
  std::memcpy(out, &h->map->risk_maps_[0][0], sizeof(float) * static_cast<size_t>(VOXEL_NUM) * PREDICTION_TIMES);
}
void refmap_upload(void* hv, const float* in) {
  Handle* h = static_cast<Handle*>(hv);
  std::memcpy(&h->map->risk_maps_[0][0], in, sizeof(float) * static_cast<size_t>(VOXEL_NUM) * PREDICTION_TIMES);
}
void refmap_set_map_time(void* hv, double t) { static_cast<Handle*>(hv)->map->last_update_time_ = ros::Time(t); }

// getClearOcccupancy(pos, int t) / (pos, double dt) of the class under test
void refmap_query(void* hv, const double* pos, const int32_t* t, int n, int8_t* out) {
  Handle* h = static_cast<Handle*>(hv);
  for (int i = 0; i < n; i++)
    out[i] = static_cast<int8_t>(h->map->getClearOcccupancy(Eigen::Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), static_cast<int>(t[i])));
}
void refmap_query_dt(void* hv, const double* pos, const double* dt, int n, int8_t* out) {
  Handle* h = static_cast<Handle*>(hv);
  for (int i = 0; i < n; i++)
    out[i] = static_cast<int8_t>(h->map->getClearOcccupancy(Eigen::Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), static_cast<double>(dt[i])));
}
// MapBase::getInflateOccupancy (map.cpp:345-395): mode 0 (pos), 1 (pos, int t), 2 (pos, double t).  It reads
// MapBase::risk_maps_, which RiskBase shadows with its own array (risk_base.h:46), so it is meaningful for the
// ground-truth map only.
void refmap_inflate_query(void* hv, int mode, const double* pos, const double* t, int n, int8_t* out) {
  Handle* h = static_cast<Handle*>(hv);
  for (int i = 0; i < n; i++) {
    const Eigen::Vector3d p(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    int r;
    if (mode == 0) r = h->map->MapBase::getInflateOccupancy(p);
    else if (mode == 1) r = h->map->MapBase::getInflateOccupancy(p, static_cast<int>(t[i]));
    else r = h->map->MapBase::getInflateOccupancy(p, static_cast<double>(t[i]));
    out[i] = static_cast<int8_t>(r);
  }
}
// getObstaclePoints(points, t0, t1, lc, hc): MapBase's for the ground-truth map (map.cpp:463-531), RiskBase's own
// (risk_base.cpp:295-337).  Returns the count; the first cap points to out.
int refmap_obstacle_points(void* hv, double t0, double t1, const double* lc, const double* hc, double* out, int cap) {
  Quiet q;
  Handle* h = static_cast<Handle*>(hv);
  std::vector<Eigen::Vector3d> pts;
  h->map->getObstaclePoints(pts, t0, t1, Eigen::Vector3d(lc[0], lc[1], lc[2]), Eigen::Vector3d(hc[0], hc[1], hc[2]));
  for (int i = 0; i < static_cast<int>(pts.size()) && i < cap; i++) for (int k = 0; k < 3; k++) out[3 * i + k] = pts[i](k);
  return static_cast<int>(pts.size());
}
// addParticlesToRiskMap(pts (map frame, double), risks, t_index) / addObstaclesToRiskMap(pts, t_index)
void refmap_add_particles(void* hv, const double* pts, const float* risks, int n, int t_index) {
  Handle* h = static_cast<Handle*>(hv);
  std::vector<Eigen::Vector3d> p;
  std::vector<float> r(risks, risks + n);
  for (int i = 0; i < n; i++) p.push_back(Eigen::Vector3d(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
  h->map->addParticlesToRiskMap(p, r, t_index);
}
void refmap_add_obstacles(void* hv, const double* pts, int n, int t_index) {
  Handle* h = static_cast<Handle*>(hv);
  std::vector<Eigen::Vector3d> p;
  for (int i = 0; i < n; i++) p.push_back(Eigen::Vector3d(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
  h->map->addObstaclesToRiskMap(p, t_index);
}
// MapBase::filterPointCloud (map.cpp:107-132): camera-frame cloud -> valid_clouds (cap 5000 x 3); filter_res is
// "map/resolution".  Returns the number of valid points.
int refmap_filter_cloud(void* hv, const float* cam, int n, float filter_res, float* out) {
  Handle* h = static_cast<Handle*>(hv);
  h->map->filter_res_ = filter_res;
  pcl::PointCloud<pcl::PointXYZ>::Ptr in(new pcl::PointCloud<pcl::PointXYZ>()), outc(new pcl::PointCloud<pcl::PointXYZ>());
  for (int i = 0; i < n; i++) in->points.push_back(pcl::PointXYZ(cam[3 * i], cam[3 * i + 1], cam[3 * i + 2]));
  int nv = 0;
  h->map->filterPointCloud(in, outc, out, nv);
  return nv;
}
// voxel index math (map.h:153-215) on local points / indices
void refmap_index_math(void* hv, const float* local, int n, int32_t* out_idx, int8_t* out_in_range, const int32_t* idx, int m,
                       float* out_pos) {
  Handle* h = static_cast<Handle*>(hv);
  for (int i = 0; i < n; i++) {
    const Eigen::Vector3f p(local[3 * i], local[3 * i + 1], local[3 * i + 2]);
    out_in_range[i] = h->map->isInRange(p) ? 1 : 0;
    out_idx[i] = h->map->getVoxelIndex(p);
  }
  for (int i = 0; i < m; i++) {
    const Eigen::Vector3f p = h->map->getVoxelPosition(idx[i]);
    for (int k = 0; k < 3; k++) out_pos[3 * i + k] = p[k];
  }
}

}  // extern "C"

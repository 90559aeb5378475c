// pop.hpp — C++ host side above the C ABI of libpop.so (include/pop.h): header-only classes that keep the
// names, argument meaning and return conventions of the reference's C++ interfaces for the hot path, with
// std::array<double, 3> standing where the reference passes Eigen::Vector3d (this image has neither Eigen nor
// ROS; a plan_manager build swaps the alias below for Eigen types without touching the rest).
//
//   pop::Engine          one handle per GPU: N local agents               plan_manager/src/baseline_fake.cpp:18-51
//   pop::RiskVoxelMap    MapBase / FakeParticleRiskVoxel / RiskBase       plan_env/include/plan_env/map.h:93-146
//   pop::HybridAstar     FakeRiskHybridAstar::search / getPathWithVel     path_searching/include/path_searching/fake_risk_hybrid_a_star.h
//   pop::BezierOpt       traj_opt::BezierOpt                              traj_opt/include/bernstein/bezier_optimizer.hpp:38-87
//   pop::ParticleMapNode risk_mapping_node + dsp_map::DSPMap              plan_env/src/risk_mapping_node.cpp:155-188
//   pop::GridMap         GridMap (raycast map)                            plan_env/include/plan_env/grid_map.h
//
// Errors: the reference reports failures by return value on this path (bool / enum / int); the C ABI returns
// negative codes for misuse and resource failures, which these classes turn into pop::Error exceptions —
// per-agent planning outcomes stay return values (ASTAR_RET, bool optimize(), -1/0/1 occupancy).
#ifndef POP_HPP_
#define POP_HPP_

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "pop.h"

namespace pop {

using Vec3 = std::array<double, 3>;        // Eigen::Vector3d in the reference
using Mat3 = std::array<double, 9>;        // Eigen::Matrix3d, rows pos; vel; acc (bezier_optimizer.cpp:140,160,187)
using HPoly = std::vector<std::array<double, 4>>;   // Eigen::MatrixX4d, row = h0 x + h1 y + h2 z + h3 <= 0

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error("libpop error " + std::to_string(c) + ": " + what), code(c) {}
};

// One pop_handle: all local agents of one GPU.  No CPU fallback: construction throws POP_E_NO_DEVICE without a GPU.
class Engine {
 public:
  explicit Engine(const pop_config& cfg, int device = 0, void* stream = nullptr) : cfg_(cfg) {
    const int rc = pop_create(&cfg_, device, stream, &h_);
    if (rc != POP_OK) { h_ = nullptr; throw Error(rc, "pop_create failed (no CUDA device -> POP_E_NO_DEVICE; there is no CPU fallback)"); }
  }
  ~Engine() { if (h_) pop_destroy(h_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  static pop_config defaultConfig() { pop_config c; pop_default_config(&c); return c; }

  pop_handle* handle() const { return h_; }
  const pop_config& config() const { return cfg_; }
  void check(int rc) const { if (rc != POP_OK) throw Error(rc, pop_last_error(h_)); }

  // "map/cloud", "map/state", odometry, coordinator inputs
  void setWorldCloud(const std::vector<float>& xyz) { check(pop_set_world_cloud(h_, xyz.data(), static_cast<int>(xyz.size() / 3))); }
  void setCylinders(const std::vector<pop_cylinder>& cyl) { check(pop_set_cylinders(h_, cyl.data(), static_cast<int>(cyl.size()))); }
  void setPoses(const std::vector<float>& poses) { check(pop_set_poses(h_, poses.data())); }
  void setBodyParticles(const std::vector<double>& offs, const std::vector<int32_t>& counts) { check(pop_set_body_particles(h_, offs.data(), counts.data())); }
  void setTrajectories(const std::vector<pop_traj_record>& recs) { check(pop_set_trajectories(h_, recs.data())); }
  // FakeParticleRiskVoxel::updateMap for every local agent
  void updateMap(double t_map) { check(pop_map_update_gt(h_, t_map)); }
  // FakeBaselinePlanner::replan x N (map update included); returns per-agent POP_REPLAN_* and fills the records
  std::vector<int32_t> replan(const std::vector<double>& start9, const std::vector<double>& goal3, const std::vector<float>& poses,
                              double t_map, double t_plan, std::vector<pop_traj_record>* traj) {
    std::vector<int32_t> status(cfg_.n_agents_local);
    traj->resize(cfg_.n_agents_local);
    pop_cycle_io io{start9.data(), goal3.data(), poses.data(), t_map, t_plan, status.data(), traj->data()};
    check(pop_replan_cycle(h_, &io));
    return status;
  }
  // BaselinePlanner::replan x N on the SOGM the last map update (e.g. the particle map) left
  std::vector<int32_t> plan(const std::vector<double>& start9, const std::vector<double>& goal3, double t_plan,
                            std::vector<pop_traj_record>* traj) {
    std::vector<int32_t> status(cfg_.n_agents_local);
    traj->resize(cfg_.n_agents_local);
    pop_cycle_io io{start9.data(), goal3.data(), nullptr, 0.0, t_plan, status.data(), traj->data()};
    check(pop_plan_cycle(h_, &io));
    return status;
  }
  // BaselinePlanner::isTrajSafe(T) x N at wall-clock time t_now (1 = safe)
  std::vector<int8_t> isTrajSafe(double t_now, double T) {
    std::vector<int8_t> safe(cfg_.n_agents_local);
    check(pop_traj_safe_batch(h_, t_now, T, safe.data(), nullptr));
    return safe;
  }
  void publishLocal() { check(pop_traj_publish_local(h_)); }
  void advance(double t_next) { check(pop_swarm_advance(h_, t_next)); }
  // "/broadcast_traj" over several GPUs: one handle per GPU (a process or a host thread each), rank 0 makes the id
  static std::vector<uint8_t> commUniqueId() {
    std::vector<uint8_t> id(POP_COMM_ID_BYTES);
    const int rc = pop_comm_unique_id(id.data());
    if (rc != POP_OK) throw Error(rc, "pop_comm_unique_id failed (libnccl.so.2 not loadable?)");
    return id;
  }
  void commInit(const std::vector<uint8_t>& id, int rank, int world) { check(pop_comm_init(h_, world > 1 ? id.data() : nullptr, rank, world)); }
  // after a cycle: the allgather of its records; pipelined = the other ranks' records arrive one call later
  void exchange(bool pipelined = false) { check(pipelined ? pop_exchange_trajs_pipelined(h_) : pop_exchange_trajs(h_)); }
  std::vector<pop_traj_record> trajectories() {
    std::vector<pop_traj_record> t(cfg_.n_agents_total);
    check(pop_get_trajectories(h_, t.data()));
    return t;
  }

 private:
  pop_config cfg_;
  pop_handle* h_ = nullptr;
};

// MapBase's query side for one agent (methods are non-virtual in the reference; callers hold the concrete type)
class RiskVoxelMap {
 public:
  RiskVoxelMap(Engine& e, int agent) : e_(e), agent_(agent) {}
  // 0 free, 1 occupied, -1 out of range
  int getClearOcccupancy(const Vec3& pos, int t) const {
    int8_t out = 0;
    e_.check(pop_map_query(e_.handle(), agent_, pos.data(), &t, 1, &out));
    return out;
  }
  int getClearOcccupancy(const Vec3& pos) const { return getClearOcccupancy(pos, 0); }
  int getClearOcccupancy(const Vec3& pos, double dt) const {
    int8_t out = 0;
    e_.check(pop_map_query_dt(e_.handle(), agent_, pos.data(), &dt, 1, &out));
    return out;
  }
  // MapBase::getInflateOccupancy in its three forms (map.h:120-122)
  int getInflateOccupancy(const Vec3& pos) const {
    int8_t out = 0;
    e_.check(pop_map_query_inflate(e_.handle(), agent_, pos.data(), nullptr, 0, 1, &out));
    return out;
  }
  int getInflateOccupancy(const Vec3& pos, int t) const {
    int8_t out = 0;
    const double td = t;
    e_.check(pop_map_query_inflate(e_.handle(), agent_, pos.data(), &td, 1, 1, &out));
    return out;
  }
  int getInflateOccupancy(const Vec3& pos, double t) const {
    int8_t out = 0;
    e_.check(pop_map_query_inflate(e_.handle(), agent_, pos.data(), &t, 2, 1, &out));
    return out;
  }
  // batched form the planner should prefer: pos = n x 3
  std::vector<int8_t> getClearOcccupancy(const std::vector<double>& pos, const std::vector<int32_t>& t) const {
    std::vector<int8_t> out(t.size());
    e_.check(pop_map_query(e_.handle(), agent_, pos.data(), t.data(), static_cast<int>(t.size()), out.data()));
    return out;
  }
  void getObstaclePoints(std::vector<Vec3>& points, double t_start, double t_end, const Vec3& lower_corner,
                         const Vec3& higher_corner) const {
    points.resize(e_.config().max_obstacle_points);
    int n = 0;
    const int rc = pop_map_obstacle_points(e_.handle(), agent_, t_start, t_end, lower_corner.data(), higher_corner.data(),
                                           points.empty() ? nullptr : points[0].data(), static_cast<int>(points.size()), &n);
    if (rc != POP_OK && rc != POP_E_CAPACITY) e_.check(rc);
    points.resize(n < static_cast<int>(points.size()) ? n : points.size());
  }
  void addParticlesToRiskMap(const std::vector<Vec3>& pts, const std::vector<float>& risks, int t_index) {
    e_.check(pop_map_add_particles(e_.handle(), agent_, pts.empty() ? nullptr : pts[0].data(), risks.data(),
                                   static_cast<int>(pts.size()), t_index));
  }
  // risk_maps_[VOXEL_NUM][PREDICTION_TIMES]
  std::vector<float> riskMaps() const {
    const pop_config& c = e_.config();
    std::vector<float> out(static_cast<size_t>(c.nx) * c.ny * c.nz * c.nt);
    e_.check(pop_map_download(e_.handle(), agent_, out.data()));
    return out;
  }

 private:
  Engine& e_;
  int agent_;
};

// FakeRiskHybridAstar for all local agents at once
class HybridAstar {
 public:
  explicit HybridAstar(Engine& e) : e_(e) {}
  struct Result { std::vector<int32_t> ret, len; std::vector<double> path; };   // ASTAR_RET, samples, n x POP_MAX_PATH x 6
  Result search(const std::vector<double>& start9, const std::vector<double>& goal3, const std::vector<double>& t_start) {
    const int n = e_.config().n_agents_local;
    Result r{std::vector<int32_t>(n), std::vector<int32_t>(n), std::vector<double>(static_cast<size_t>(n) * POP_MAX_PATH * 6)};
    e_.check(pop_astar_search_batch(e_.handle(), start9.data(), goal3.data(), t_start.data(), r.ret.data(), r.path.data(),
                                    r.len.data(), nullptr));
    return r;
  }

 private:
  Engine& e_;
};

// traj_opt::BezierOpt for one problem (the batched entry point is pop_qp_solve_batch)
class BezierOpt {
 public:
  explicit BezierOpt(Engine& e) : e_(e) {}
  void setup(const Mat3& start, const Mat3& end, const std::vector<double>& t, const std::vector<HPoly>& hPolys,
             const double& vmax = 3.0, const double& amax = 3.0) {
    M_ = static_cast<int32_t>(t.size());
    if (M_ < 1 || M_ > POP_MAX_PIECES || hPolys.size() != t.size()) throw Error(POP_E_INVALID, "1..POP_MAX_PIECES pieces, one polytope each");
    init_ = start; goal_ = end; vmax_ = vmax; amax_ = amax;
    dur_.assign(POP_MAX_PIECES, 0.0);
    rows_.assign(POP_MAX_PIECES, 0);
    hpoly_.assign(static_cast<size_t>(POP_MAX_PIECES) * POP_MAX_FACES * 4, 0.0);
    for (int i = 0; i < M_; i++) {
      dur_[i] = t[i];
      if (hPolys[i].size() > POP_MAX_FACES) throw Error(POP_E_CAPACITY, "polytope with more than POP_MAX_FACES faces");
      rows_[i] = static_cast<int32_t>(hPolys[i].size());
      for (size_t r = 0; r < hPolys[i].size(); r++)
        for (int k = 0; k < 4; k++) hpoly_[(static_cast<size_t>(i) * POP_MAX_FACES + r) * 4 + k] = hPolys[i][r][k];
    }
  }
  bool optimize() {   // success iff OSQP status is solved or solved-inaccurate (bezier_optimizer.cpp:279)
    cpts_.assign(static_cast<size_t>(POP_MAX_PIECES) * 15, 0.0);
    e_.check(pop_qp_solve_batch(e_.handle(), 1, &M_, init_.data(), goal_.data(), dur_.data(), hpoly_.data(), rows_.data(), vmax_,
                                amax_, cpts_.data(), &status_, &iters_));
    return status_ == 1 || status_ == 2;
  }
  std::vector<Vec3> getOptCtrlPts() const {   // 5 M control points
    std::vector<Vec3> out(5 * M_);
    for (int i = 0; i < 5 * M_; i++) out[i] = {cpts_[3 * i], cpts_[3 * i + 1], cpts_[3 * i + 2]};
    return out;
  }
  int status() const { return status_; }
  int iterations() const { return iters_; }

 private:
  Engine& e_;
  int32_t M_ = 0, status_ = 0, iters_ = 0;
  Mat3 init_{}, goal_{};
  double vmax_ = 3.0, amax_ = 3.0;
  std::vector<double> dur_, hpoly_, cpts_;
  std::vector<int32_t> rows_;
};

// risk_mapping_node: one DSPMap per local agent behind one call per frame
class ParticleMapNode {
 public:
  explicit ParticleMapNode(Engine& e, const pop_dsp_config* cfg = nullptr) : e_(e) {
    if (cfg) d_ = *cfg; else pop_dsp_default_config(&e.config(), &d_);
    e_.check(pop_dsp_attach(e_.handle(), &d_));
  }
  const pop_dsp_config& config() const { return d_; }
  // cloudCallback: cam = n_local x max_input_points x 3 camera-frame points; returns DSPMap::update's result per agent
  std::vector<int32_t> cloudCallback(const std::vector<float>& cam, const std::vector<int32_t>& counts, const std::vector<float>& poses,
                                     const std::vector<float>& quats, double stamp) {
    std::vector<int32_t> ok(e_.config().n_agents_local);
    e_.check(pop_dsp_update(e_.handle(), cam.data(), counts.data(), poses.data(), quats.data(), stamp, ok.data()));
    return ok;
  }
  // getOccupancyMap: centres of the occupied voxels of one agent
  std::vector<float> getOccupancyMap(int agent) const {
    const pop_config& c = e_.config();
    std::vector<float> out(static_cast<size_t>(c.nx) * c.ny * c.nz * 3);
    int32_t n = 0;
    e_.check(pop_dsp_occupied(e_.handle(), agent, out.data(), c.nx * c.ny * c.nz, &n));
    out.resize(static_cast<size_t>(n) * 3);
    return out;
  }

 private:
  Engine& e_;
  pop_dsp_config d_{};
};

// GridMap: the log-odds raycast map, n independent maps on one GPU
class GridMap {
 public:
  explicit GridMap(const pop_gridmap_config& cfg, int n_maps = 1, int device = 0, void* stream = nullptr) : cfg_(cfg), n_(n_maps) {
    const int rc = pop_gridmap_create(&cfg_, n_maps, device, stream, &g_);
    if (rc != POP_OK) { g_ = nullptr; throw Error(rc, "pop_gridmap_create failed (no CUDA device -> POP_E_NO_DEVICE)"); }
  }
  ~GridMap() { if (g_) pop_gridmap_destroy(g_); }
  GridMap(const GridMap&) = delete;
  GridMap& operator=(const GridMap&) = delete;
  static pop_gridmap_config defaultConfig() { pop_gridmap_config c; pop_gridmap_default_config(&c); return c; }
  void check(int rc) const { if (rc != POP_OK) throw Error(rc, pop_gridmap_last_error(g_)); }
  // updateOccupancyCallback after projectDepthImage: pts = n_maps x max_points x 3
  void updateOccupancy(const std::vector<double>& pts, const std::vector<int32_t>& counts, const std::vector<double>& cam_pos) {
    check(pop_gridmap_update(g_, pts.data(), counts.data(), cam_pos.data()));
  }
  int getOccupancy(int map, const Vec3& pos) const { int8_t o = 0; check(pop_gridmap_query(g_, map, pos.data(), 1, &o, nullptr)); return o; }
  int getInflateOccupancy(int map, const Vec3& pos) const { int8_t i = 0; check(pop_gridmap_query(g_, map, pos.data(), 1, nullptr, &i)); return i; }

 private:
  pop_gridmap_config cfg_;
  int n_;
  pop_gridmap* g_ = nullptr;
};

// traj_utils/BezierTraj on the ROS1 wire
inline std::vector<uint8_t> toWire(const pop_traj_record& rec, double pub_time) {
  const int n = pop_traj_to_wire(&rec, pub_time, nullptr, 0);
  if (n < 0) throw Error(n, "pop_traj_to_wire");
  std::vector<uint8_t> out(n);
  const int m = pop_traj_to_wire(&rec, pub_time, out.data(), n);
  if (m != n) throw Error(m, "pop_traj_to_wire");
  return out;
}
inline pop_traj_record fromWire(const std::vector<uint8_t>& bytes, double* pub_time = nullptr) {
  pop_traj_record rec;
  const int n = pop_traj_from_wire(bytes.data(), static_cast<int>(bytes.size()), &rec, pub_time);
  if (n < 0) throw Error(n, "pop_traj_from_wire");
  return rec;
}

}  // namespace pop
#endif  // POP_HPP_

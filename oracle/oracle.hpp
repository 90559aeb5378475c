// oracle.hpp — CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).
//
// This directory is the parity oracle: a dependency-free C++17 restatement of the
// reference's CPU algorithms for the per-replan path.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build, load or call it; the
// product (pred_occ_planner_b200/) never does.
//
// Parity pinning: Bernstein evaluation, the QP fixture and the separator instance are pinned
// by the reference's own tests (traj_utils/test/test_bernstein.cpp:58-108,
// traj_opt/test/test_bezier_opt.cpp:47-221, utils/separator/src/test_separator.cpp:20-32).
// For SOGM construction, occupancy queries, hybrid A*, FIRI and peer projection the
// reference holds no tests and cannot be compiled here (ROS/Eigen/PCL/OSQP/GLPK absent):
// PARITY UNPINNED for those — they follow the cited source line by line.
//
// Floating-point contract: compiled with -ffp-contract=off so that fp32/fp64 expressions
// round exactly like the reference's x86-64 -O3 build (no FMA).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/pop.h"

namespace orc {

struct V3f { float x, y, z; };
struct V3d { double x, y, z; };
struct V3i { int x, y, z; };

// ---------------------------------------------------------------------------------------
// Grid geometry — plan_env/include/plan_env/map.h:153-215, init at
// plan_env/src/fake_particle_risk_voxel.cpp:27-46 / risk_base.cpp:25-39.
struct Grid {
  int nx, ny, nz, nt;
  float res, dt;
  float rx, ry, rz;  // local_update_range_{x,y,z}_
  float risk_threshold, clearance, ground, ceiling;
  float thr_region, decay_region, decay_voxel;
  int query_variant, kernel_shape;
  int inf_step;
  std::vector<V3i> kernel;  // inflate_kernel_
  V3f pose;                 // pose_
  double t_map;             // last_update_time_
  std::vector<float> risk;  // risk_maps_[VOXEL_NUM][PREDICTION_TIMES]

  void init(const pop_config& c);
  int V() const { return nx * ny * nz; }
  bool in_range(V3f p) const {
    return p.x > -rx && p.x < rx && p.y > -ry && p.y < ry && p.z > -rz && p.z < rz;
  }
  bool in_range(V3i p) const {
    return p.x >= 0 && p.x < nx && p.y >= 0 && p.y < ny && p.z >= 0 && p.z < nz;
  }
  V3i rel_index(V3f p) const {
    return V3i{static_cast<int>((p.x + rx) / res), static_cast<int>((p.y + ry) / res),
               static_cast<int>((p.z + rz) / res)};
  }
  int index(V3f p) const {
    V3i i = rel_index(p);
    return i.z * nx * ny + i.y * nx + i.x;
  }
  int index(V3i p) const { return p.z * nx * ny + p.y * nx + p.x; }
  V3f voxel_position(int idx) const {  // world frame, voxel CORNER (map.h:186-194)
    int x = idx % nx, y = (idx / nx) % ny, z = idx / (nx * ny);
    V3f r{x * res - rx, y * res - ry, z * res - rz};
    return V3f{r.x + pose.x, r.y + pose.y, r.z + pose.z};
  }
  float& at(int idx, int t) { return risk[static_cast<size_t>(idx) * nt + t]; }
  float at(int idx, int t) const { return risk[static_cast<size_t>(idx) * nt + t]; }
};

// FakeParticleRiskVoxel::updateMap without the peer part (fake_particle_risk_voxel.cpp:80-170)
void gt_update(Grid& g, const float* cloud, int n, const pop_cylinder* cyl, int ncyl,
               double t_map);
// addParticlesToRiskMap (fake_particle_risk_voxel.cpp:291-301)
void add_particles(Grid& g, const V3d* pts, const float* risks, int n, int t_index);
// getClearOcccupancy(pos,int t) / (pos,double dt)
int query(const Grid& g, V3d pos, int t);
int query_dt(const Grid& g, V3d pos, double dt);
int inflate_query(const Grid& g, V3d pos, int mode, double t);
// getObstaclePoints(points,t_start,t_end,lc,hc) (map.cpp:463-531 / risk_base.cpp:295-337)
void obstacle_points(const Grid& g, double t_start, double t_end, V3d lc, V3d hc,
                     std::vector<V3d>& out);

// ---------------------------------------------------------------------------------------
// Bernstein / Bezier — traj_utils/include/traj_utils/bernstein.hpp, src/bernstein.cpp
struct Bezier {
  int M = 0;
  std::vector<double> t;      // piece durations
  std::vector<double> cpts;   // (5M) x 3 row-major
  double T = 0;
  Bezier() {}
  Bezier(const std::vector<double>& dur, const std::vector<double>& c);
  int locate(double tt) const;
  V3d pos(double tt) const;
  V3d vel(double tt) const;
  V3d acc(double tt) const;
  double max_vel_rate() const;
  double max_acc_rate() const;
};
// single piece on [t0, tf] (BernsteinPiece)
V3d piece_eval(const double* c5x3, double t0, double tf, double tt, int deriv);
extern const double kBernsteinA[25];  // 5x5 coefficient matrix (bernstein.cpp:80-88)

// ---------------------------------------------------------------------------------------
// Peers — traj_coordinator/src/particles.cpp
struct Swarm {
  int n_agents = 0;
  std::vector<std::vector<V3d>> body;  // particles_buffer_ per global id
  std::vector<pop_traj_record> trajs;  // swarm_trajs_ by id (n_pieces==0 → absent)
  float replan_risk_rate = 0.f;
  int num_resample = 10;
  std::vector<float> noise;            // standard-normal triples (nx, ny, nz) the Gaussian resampling consumes, see o_peers.cpp
};
std::vector<V3d> ego_particles(const double size[3]);  // initEgoParticles :62-75
// getParticlesWithRisk (:346-422) for rate==0; returns validity
bool particles_with_risk(const Swarm& s, int ego_id, int idx_agent, double t0, std::vector<V3d>& pts,
                         std::vector<float>& risks);
// FakeParticleRiskVoxel::updateMap peer part (:178-218) with per-peer aligned (pts,risks)
void add_other_agents(Grid& g, const Swarm& s, int ego_id);

// ---------------------------------------------------------------------------------------
// Hybrid A* — path_searching/src/fake_risk_hybrid_a_star.cpp
struct AstarResult {
  int ret = 0;
  int iters = 0, nodes = 0, retried = 0;
  std::vector<double> path;  // n x 6 from getPathWithVel
};
AstarResult astar_replan(const Grid& g, const pop_config& c, const double start[9],
                         const double goal[3], double t_start, int math_mode);

// ---------------------------------------------------------------------------------------
// Corridor — plan_manager/include/sfc_gen/firi.hpp, traj_utils/sdlp.hpp, baseline_fake.cpp
struct Poly { std::vector<double> h; int rows() const { return static_cast<int>(h.size() / 4); } };
double linprog(int d, const double* c, const double* A, const double* b, int m, double* x,
               uint64_t* rng_state);
bool firi(const double bd[24], const std::vector<V3d>& pc, V3d a, V3d b, Poly& out, double r[3],
          int iterations, uint64_t* rng_state);
struct CorridorResult {
  std::vector<Poly> polys;
  double goal[6];
  bool ok = false;
};
CorridorResult corridor_stage(const Grid& g, const pop_config& c, const std::vector<double>& path,
                              const double start[9], double traj_start_time);

// ---------------------------------------------------------------------------------------
// QP — traj_opt/src/bezier_optimizer.cpp + OSQP 0.6 algorithm
struct QpProblem {
  int n = 0, m = 0;
  std::vector<double> Q;   // dense n x n
  std::vector<double> A;   // dense m x n row-major
  std::vector<double> l, u;
};
void qp_setup(QpProblem& p, const double init[9], const double goal[9],
              const std::vector<double>& dur, const std::vector<Poly>& polys, double vmax,
              double amax);
struct OsqpSettings {
  double rho = 0.1, sigma = 1e-6, alpha = 1.6, eps_abs = 1e-3, eps_rel = 1e-3;
  double eps_prim_inf = 1e-4, eps_dual_inf = 1e-4;
  int max_iter = 4000, scaling = 10, check_termination = 25, adaptive_rho_interval = 25;
  double adaptive_rho_tolerance = 5.0;
};
struct OsqpResult { int status = 0, iters = 0; std::vector<double> x; };
OsqpResult osqp_solve(const QpProblem& p, const OsqpSettings& s);

// Separator — utils/separator/src/separator_glpk.cpp:73-246 (feasibility only)
bool separable(const std::vector<V3d>& A, const std::vector<V3d>& B);

// Cycle — plan_manager/src/baseline_fake.cpp:253-453 + particles.cpp:223-283
bool safe_after_opt(const Swarm& s, int ego_id, const std::vector<double>& cpts, int M, double t0);
extern std::vector<double>* g_sep_trace;   // test hook: when set, safe_after_opt records its point sets instead of solving
bool traj_safe(const Grid& g, const pop_traj_record& r, double tau, double t_now, double T, double* t_unsafe);
int replan(const Grid& g, const pop_config& c, const Swarm* swarm, int ego_id, const double start[9],
           const double goal[3], double t_plan, pop_traj_record* out, int* stats);

}  // namespace orc

// ref_glue_particles.cpp — the reference's swarm coordinator, ParticleATC (traj_coordinator/src/particles.cpp), in
// oracle/_ref/libref.so.  TEST INFRASTRUCTURE; see ref_glue.cpp.  ROS is replaced by the inert stand-ins of
// oracle/ref_build/ros/ros.h (the callbacks are called directly, ros::Time::now() is an input), the two message types
// by plain structs with the same fields, and the GLPK separator (utils/separator/src/separator_glpk.cpp, absent
// library) by a stub that only records the point sets it is handed — isSafeAfterOpt's LP is NOT pinned by this.
#include <cstdint>
#include <iostream>
#include <numeric>
#include <sstream>

#define private public
#define protected public
#include <traj_coordinator/particle.hpp>
#undef private
#undef protected

#include "../include/pop.h"

namespace separator {
struct Separator::PImpl {};
Separator::Separator() {}
Separator::~Separator() {}
// every call of one isSafeAfterOpt: the sizes of both point sets, then set B's points
static std::vector<double> g_calls;
bool Separator::solveModel(Eigen::Vector3d&, double&, const std::vector<Eigen::Vector3d>& a,
                           const std::vector<Eigen::Vector3d>& b) {
  g_calls.push_back(static_cast<double>(a.size()));
  g_calls.push_back(static_cast<double>(b.size()));
  for (const Eigen::Vector3d& p : b) for (int k = 0; k < 3; k++) g_calls.push_back(p(k));
  return true;
}
}  // namespace separator

extern "C" {

// ParticleATC::initEgoParticles (particles.cpp:62-75): returns the count, offsets written to out (cap x 3)
int ref_ego_particles(const double* size, double* out, int cap) {
  ros::NodeHandle nh;
  nh.values["swarm/drone_size_x"] = size[0]; nh.values["swarm/drone_size_y"] = size[1]; nh.values["swarm/drone_size_z"] = size[2];
  nh.values["swarm/num_robots"] = 1;
  ParticleATC atc(nh);
  atc.init();
  const std::vector<Eigen::Vector3d>& p = atc.getEgoParticles();
  for (int i = 0; i < static_cast<int>(p.size()) && i < cap; i++) for (int k = 0; k < 3; k++) out[3 * i + k] = p[i](k);
  return static_cast<int>(p.size());
}

void ref_feed_atc(ParticleATC& atc, int n, int ego_id, const double* offs, int max_body, const int32_t* counts,
                 const pop_traj_record* recs, double now) {
  for (int i = 0; i < n; i++) {
    if (i == ego_id) continue;
    std::shared_ptr<geometry_msgs::PolygonStamped> m(new geometry_msgs::PolygonStamped());
    m->header.frame_id = "uav" + std::to_string(i) + "/base_link";
    m->polygon.points.resize(counts[i]);
    for (int k = 0; k < counts[i]; k++) {
      const double* o = offs + (static_cast<size_t>(i) * max_body + k) * 3;
      m->polygon.points[k].x = static_cast<float>(o[0]);
      m->polygon.points[k].y = static_cast<float>(o[1]);
      m->polygon.points[k].z = static_cast<float>(o[2]);
    }
    atc.particlesCallback(m);
  }
  for (int i = 0; i < n; i++) {
    const pop_traj_record& r = recs[i];
    if (r.n_pieces <= 0) continue;
    std::shared_ptr<traj_utils::BezierTraj> m(new traj_utils::BezierTraj());
    m->drone_id = i; m->traj_id = r.traj_id; m->order = 4;
    m->start_time = ros::Time(r.start_time); m->pub_time = ros::Time(now);
    for (int k = 0; k < r.n_pieces; k++) m->duration.push_back(r.duration[k]);
    m->cpts.resize(static_cast<size_t>(r.n_pieces) * 5);
    for (int k = 0; k < r.n_pieces * 5; k++) { m->cpts[k].x = r.cpts[3 * k]; m->cpts[k].y = r.cpts[3 * k + 1]; m->cpts[k].z = r.cpts[3 * k + 2]; }
    atc.trajectoryCallback(m);
  }
}

// A coordinator of drone `ego_id` in a swarm of n agents (ids 0..9: particlesCallback reads the id from one character
// of the frame name, particles.cpp:90) that has received every peer's body particles (offs n x max_body x 3, counts;
// geometry_msgs/Point32 = float32 on the wire) and trajectory record (recs[i].n_pieces == 0: none yet), then
// getParticlesWithRisk(idx_agent, t0) for nq queries.  out_n[q] = -1 when it returns false, else the particle count;
// points and risks of query q at out_pts + q * cap * 3 and out_risks + q * cap.
void ref_particles_with_risk(int n, int ego_id, const double* size, float replan_risk_rate, const double* offs, int max_body,
                             const int32_t* counts, const pop_traj_record* recs, double now, int nq,
                             const int32_t* q_agent, const double* q_t0, double* out_pts, float* out_risks, int cap,
                             int32_t* out_n) {
  std::ostringstream sink;
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  ros::Time::set_now(now);
  ros::NodeHandle nh;
  nh.values["drone_id"] = ego_id;
  nh.values["swarm/num_robots"] = n;
  nh.values["swarm/drone_size_x"] = size[0]; nh.values["swarm/drone_size_y"] = size[1]; nh.values["swarm/drone_size_z"] = size[2];
  nh.values["swarm/replan_risk_rate"] = replan_risk_rate;
  ParticleATC atc(nh);
  atc.init();
  ref_feed_atc(atc, n, ego_id, offs, max_body, counts, recs, now);
  for (int q = 0; q < nq; q++) {
    std::vector<Eigen::Vector3d> pts;
    std::vector<float> risks;
    const bool ok = atc.getParticlesWithRisk(pts, risks, q_agent[q], q_t0[q]);
    out_n[q] = ok ? static_cast<int32_t>(pts.size()) : -1;
    for (int i = 0; ok && i < static_cast<int>(pts.size()) && i < cap; i++) {
      for (int k = 0; k < 3; k++) out_pts[(static_cast<size_t>(q) * cap + i) * 3 + k] = pts[i](k);
      out_risks[static_cast<size_t>(q) * cap + i] = i < static_cast<int>(risks.size()) ? risks[i] : -1.0f;
    }
  }
  std::cout.rdbuf(old);
}

// isSafeAfterOpt(traj) at wall-clock time `now` (particles.cpp:223-283) up to the separating-plane LP, whose solver
// (GLPK) is absent: returns the number of LPs it would solve and writes, per LP, |A|, |B| and B's points to out.
// cpts = 5M x 3 control points of the ego trajectory, dur[M].
int ref_safe_after_opt_sets(int n, int ego_id, const double* size, const double* offs, int max_body, const int32_t* counts,
                            const pop_traj_record* recs, double now, const double* cpts, const double* dur, int M,
                            double* out, int cap) {
  std::ostringstream sink;
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  ros::Time::set_now(now);
  ros::NodeHandle nh;
  nh.values["drone_id"] = ego_id;
  nh.values["swarm/num_robots"] = n;
  nh.values["swarm/drone_size_x"] = size[0]; nh.values["swarm/drone_size_y"] = size[1]; nh.values["swarm/drone_size_z"] = size[2];
  ParticleATC atc(nh);
  atc.init();
  ref_feed_atc(atc, n, ego_id, offs, max_body, counts, recs, now);
  Eigen::MatrixX3d c(5 * M, 3);
  for (int i = 0; i < 5 * M; i++) for (int k = 0; k < 3; k++) c(i, k) = cpts[3 * i + k];
  Bernstein::Bezier traj(std::vector<double>(dur, dur + M), c);
  separator::g_calls.clear();
  atc.isSafeAfterOpt(traj);
  std::cout.rdbuf(old);
  int lps = 0;
  for (size_t i = 0; i < separator::g_calls.size();) { lps++; i += 2 + 3 * static_cast<size_t>(separator::g_calls[i + 1]); }
  for (size_t i = 0; i < separator::g_calls.size() && static_cast<int>(i) < cap; i++) out[i] = separator::g_calls[i];
  return lps;
}

}  // extern "C"

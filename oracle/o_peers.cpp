// o_peers.cpp — oracle: peer-trajectory particles projected into the SOGM (TEST INFRASTRUCTURE).
// Follows traj_coordinator/src/particles.cpp:62-87,291-422 and
// plan_env/src/fake_particle_risk_voxel.cpp:175-218 of the reference.
// ego_particles and particles_with_risk are PINNED against the reference's own particles.cpp compiled into
// oracle/_ref/libref.so (ROS, the two message types and the GLPK separator replaced by stand-ins): validity, points
// and risks identical (tests/test_ref_pins.py).  add_other_agents (the deposit into the grid,
// fake_particle_risk_voxel.cpp) remains PARITY UNPINNED; property tests in tests/test_oracle_map.py.
#include "oracle.hpp"

namespace orc {

// ParticleATC::initEgoParticles (particles.cpp:62-75)
std::vector<V3d> ego_particles(const double size[3]) {
  std::vector<V3d> out;
  const double STEP = 0.15;
  for (double x = -size[0] / 2; x <= size[0] / 2; x += STEP)
    for (double y = -size[1] / 2; y <= size[1] / 2; y += STEP)
      for (double z = -size[2] / 2; z <= size[2] / 2; z += STEP) out.push_back(V3d{x, y, z});
  return out;
}

static Bezier record_traj(const pop_traj_record& r, double& t_end) {
  // trajectoryCallback (particles.cpp:131-170): durations arrive as float32
  std::vector<double> dur(r.n_pieces);
  t_end = r.start_time;
  for (int i = 0; i < r.n_pieces; i++) {
    t_end += r.duration[i];
    dur[i] = r.duration[i];
  }
  std::vector<double> c(r.cpts, r.cpts + static_cast<size_t>(r.n_pieces) * 15);
  return Bezier(dur, c);
}

// getWaypoints (:316-344) + getParticlesWithRisk (:346-422)
bool particles_with_risk(const Swarm& s, int ego_id, int idx_agent, double t0,
                         std::vector<V3d>& pts, std::vector<float>& risks) {
  const pop_traj_record& r = s.trajs[idx_agent];
  if (r.n_pieces <= 0 || idx_agent == ego_id) return false;  // not in swarm_trajs_
  double t_end;
  Bezier traj = record_traj(r, t_end);
  std::vector<V3d> way;
  bool got;
  if (r.start_time < t0 && t_end > t0) {
    V3d p = traj.pos(t0 - r.start_time);
    for (const V3d& e : s.body[idx_agent]) way.push_back(V3d{p.x + e.x, p.y + e.y, p.z + e.z});
    got = true;
  } else if (r.start_time > t0) {
    got = true;
  } else if (t_end < t0) {
    way.push_back(traj.pos(traj.T));
    got = false;
  } else {
    got = false;
  }
  if (!got || way.empty()) return false;
  float pos_stddev = s.replan_risk_rate * static_cast<float>(t0 - r.start_time);
  pts.clear();
  risks.clear();
  if (pos_stddev < 1e-3) {
    for (const V3d& e : way) { pts.push_back(e); risks.push_back(1.0f); }
  } else {
    // Gaussian resampling (:376-408).  The reference draws nx, ny, nz from three std::normal_distribution<float>(0, stddev)
    // objects over ONE std::default_random_engine seeded with time(NULL) *inside this call*: every call made within the
    // same second replays the same sequence from its start.  Made an input here: `noise` holds the standard-normal
    // triples in consumption order (libstdc++'s distribution returns ret * stddev + mean, so z * stddev + 0.0f is the very
    // value it would have returned), read from the start in every call.
    const int R = s.num_resample;
    const size_t n_tr = s.noise.size() / 3;
    if (n_tr == 0) return false;
    size_t q = 0;
    for (const V3d& e : way) {
      std::vector<float> buf;
      buf.reserve(R);
      for (int i = 0; i < R; i++) {
        const float* z = &s.noise[3 * (q % n_tr)];
        q++;
        const float nx = z[0] * pos_stddev + 0.0f, ny = z[1] * pos_stddev + 0.0f, nz = z[2] * pos_stddev + 0.0f;
        pts.push_back(V3d{e.x + static_cast<double>(nx), e.y + static_cast<double>(ny), e.z + static_cast<double>(nz)});
        buf.push_back(std::exp(-0.5f * (nx * nx + ny * ny + nz * nz) / (pos_stddev * pos_stddev)));
      }
      float sum = 0.0f;
      for (float r : buf) sum += r;                 // std::accumulate(..., 0.0f)
      for (float& r : buf) r = r * R / sum;         // r * num_resample_ / sum, left to right in float
      risks.insert(risks.end(), buf.begin(), buf.end());
    }
  }
  return true;
}

// FakeParticleRiskVoxel::updateMap peer part (:175-218). Deviation: the reference
// shares one `risks` vector that each peer clears, misaligning it with `particles`; here each
// peer's (pts, risks) stay paired.
void add_other_agents(Grid& g, const Swarm& s, int ego_id) {
  std::vector<char> valid(s.n_agents, 1);
  for (int t_idx = 0; t_idx < g.nt; t_idx++) {
    const double t = g.t_map + g.dt * t_idx;  // double + float*int
    for (int i = 0; i < s.n_agents; i++) {
      if (i == ego_id || !valid[i]) continue;
      std::vector<V3d> pts;
      std::vector<float> risks;
      bool ok = particles_with_risk(s, ego_id, i, t, pts, risks);
      valid[i] = ok;
      for (V3d& p : pts) {
        p.x -= static_cast<double>(g.pose.x);
        p.y -= static_cast<double>(g.pose.y);
        p.z -= static_cast<double>(g.pose.z);
      }
      if (!pts.empty()) add_particles(g, pts.data(), risks.data(), static_cast<int>(pts.size()), t_idx);
    }
  }
}

}  // namespace orc

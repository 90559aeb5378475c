// o_capi.cpp — extern "C" surface of the oracle for ctypes (TEST INFRASTRUCTURE).
#include <algorithm>
#include <cstdio>

#include "oracle.hpp"

using namespace orc;

struct OrcMap {
  pop_config cfg;
  Grid grid;
};

static Swarm make_swarm(const pop_config& c, const double* offs, const int32_t* counts,
                        const pop_traj_record* recs) {
  Swarm s;
  s.n_agents = c.n_agents_total;
  s.replan_risk_rate = c.replan_risk_rate;
  s.body.resize(s.n_agents);
  for (int i = 0; i < s.n_agents; i++)
    for (int k = 0; k < counts[i]; k++) {
      const double* p = offs + (static_cast<size_t>(i) * POP_MAX_BODY + k) * 3;
      s.body[i].push_back(V3d{p[0], p[1], p[2]});
    }
  s.trajs.assign(recs, recs + s.n_agents);
  return s;
}

extern "C" {

void* orc_map_create(const pop_config* cfg) {
  OrcMap* m = new OrcMap();
  m->cfg = *cfg;
  m->grid.init(*cfg);
  return m;
}
void orc_map_destroy(void* h) { delete static_cast<OrcMap*>(h); }
void orc_map_set_pose(void* h, const float* p) {
  static_cast<OrcMap*>(h)->grid.pose = V3f{p[0], p[1], p[2]};
}
int orc_map_kernel_size(void* h) { return static_cast<int>(static_cast<OrcMap*>(h)->grid.kernel.size()); }
void orc_map_update_gt(void* h, const float* cloud, int n, const pop_cylinder* cyl, int ncyl,
                       double t_map) {
  gt_update(static_cast<OrcMap*>(h)->grid, cloud, n, cyl, ncyl, t_map);
}
void orc_map_add_other_agents(void* h, const double* offs, const int32_t* counts,
                              const pop_traj_record* recs, int ego_id) {
  OrcMap* m = static_cast<OrcMap*>(h);
  Swarm s = make_swarm(m->cfg, offs, counts, recs);
  add_other_agents(m->grid, s, ego_id);
}
void orc_map_add_particles(void* h, const double* pts, const float* risks, int n, int t) {
  add_particles(static_cast<OrcMap*>(h)->grid, reinterpret_cast<const V3d*>(pts), risks, n, t);
}
void orc_map_query(void* h, const double* pos, const int32_t* t, int n, int8_t* out) {
  const Grid& g = static_cast<OrcMap*>(h)->grid;
  for (int i = 0; i < n; i++)
    out[i] = static_cast<int8_t>(query(g, V3d{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}, t[i]));
}
void orc_map_query_dt(void* h, const double* pos, const double* dt, int n, int8_t* out) {
  const Grid& g = static_cast<OrcMap*>(h)->grid;
  for (int i = 0; i < n; i++)
    out[i] = static_cast<int8_t>(query_dt(g, V3d{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}, dt[i]));
}
int orc_map_obstacle_points(void* h, double t0, double t1, const double* lc, const double* hc,
                            double* out, int cap) {
  std::vector<V3d> v;
  obstacle_points(static_cast<OrcMap*>(h)->grid, t0, t1, V3d{lc[0], lc[1], lc[2]},
                  V3d{hc[0], hc[1], hc[2]}, v);
  int n = static_cast<int>(v.size());
  for (int i = 0; i < std::min(n, cap); i++) { out[3 * i] = v[i].x; out[3 * i + 1] = v[i].y; out[3 * i + 2] = v[i].z; }
  return n;
}
void orc_map_download(void* h, float* out) {
  const Grid& g = static_cast<OrcMap*>(h)->grid;
  std::copy(g.risk.begin(), g.risk.end(), out);
}

int orc_ego_particles(const double* size, double* out, int cap) {
  std::vector<V3d> v = ego_particles(size);
  int n = static_cast<int>(v.size());
  for (int i = 0; i < std::min(n, cap); i++) { out[3 * i] = v[i].x; out[3 * i + 1] = v[i].y; out[3 * i + 2] = v[i].z; }
  return n;
}

// Bezier evaluation: dur[M], cpts[5M x 3]; deriv 0/1/2
void orc_bezier_eval(const double* dur, int M, const double* cpts, const double* t, int n, int deriv,
                     double* out) {
  Bezier b(std::vector<double>(dur, dur + M), std::vector<double>(cpts, cpts + 15 * M));
  for (int i = 0; i < n; i++) {
    V3d p = deriv == 0 ? b.pos(t[i]) : deriv == 1 ? b.vel(t[i]) : b.acc(t[i]);
    out[3 * i] = p.x; out[3 * i + 1] = p.y; out[3 * i + 2] = p.z;
  }
}
void orc_piece_eval(const double* cpts, double t0, double tf, double t, int deriv, double* out) {
  V3d p = piece_eval(cpts, t0, tf, t, deriv);
  out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
double orc_bezier_max_rate(const double* dur, int M, const double* cpts, int order) {
  Bezier b(std::vector<double>(dur, dur + M), std::vector<double>(cpts, cpts + 15 * M));
  return order == 1 ? b.max_vel_rate() : b.max_acc_rate();
}
void orc_bernstein_coeff(double* out25) { std::copy(kBernsteinA, kBernsteinA + 25, out25); }


// Hybrid A*: returns ASTAR_RET; path = n x 6 samples of getPathWithVel(corridor_tau)
int orc_astar(void* h, const double* start9, const double* goal3, double t_start, double* out_path,
              int cap, int* out_len, int* out_stats) {
  OrcMap* m = static_cast<OrcMap*>(h);
  AstarResult r = astar_replan(m->grid, m->cfg, start9, goal3, t_start, 0);
  int n = static_cast<int>(r.path.size() / 6);
  *out_len = n;
  for (int i = 0; i < std::min(n, cap) * 6; i++) out_path[i] = r.path[i];
  if (out_stats) { out_stats[0] = r.iters; out_stats[1] = r.nodes; out_stats[2] = r.retried; out_stats[3] = 0; }
  return r.ret;
}

}  // extern "C"

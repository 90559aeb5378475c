// o_capi.cpp — extern "C" surface of the oracle for ctypes (TEST INFRASTRUCTURE).
#include <algorithm>
#include <cstdio>

#include <cstring>
#include <random>

#include "oracle.hpp"

using namespace orc;

struct OrcMap {
  pop_config cfg;
  Grid grid;
};

static std::vector<float> g_resample_noise;   // orc_set_resample_noise

static Swarm make_swarm(const pop_config& c, const double* offs, const int32_t* counts,
                        const pop_traj_record* recs) {
  Swarm s;
  s.n_agents = c.n_agents_total;
  s.replan_risk_rate = c.replan_risk_rate;
  s.num_resample = c.num_resample;
  s.noise = g_resample_noise;
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

// The reference's simulator defaults (plan_env/include/plan_env/map_parameters.h:5-13, fake_particle_risk_voxel.h:28-29,
// plan_manager/config/sim_fake.yaml with the fall-backs of map.cpp:15-40, fake_risk_hybrid_a_star.cpp:62-82,
// baseline_fake.h:84-113, particles.cpp:28-34; OSQP defaults of iosqp.hpp:87-89) — the oracle's own statement of them, so
// that the CPU arm of bench.py needs nothing from libpop.  tests/test_abi.py checks it against pop_default_config.
void orc_default_config(pop_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->abi_version = POP_ABI_VERSION;
  c->nx = 66; c->ny = 66; c->nz = 20; c->nt = 9;
  c->resolution = 0.15f;
  c->time_resolution = 0.2f;
  c->risk_threshold = 0.2f;
  c->clearance = 0.45f;
  c->ground_height = -0.01f;
  c->ceiling_height = 3.0f;
  c->risk_threshold_region = 1.2f;
  c->risk_region_decay = 0.2f;
  c->risk_voxel_decay = 0.2f;
  c->kernel_shape = POP_KERNEL_DISC_Z0;
  c->query_variant = POP_QUERY_FAKE;
  c->cell_dtype = POP_CELL_F32;
  c->n_agents_local = 1; c->n_agents_total = 1; c->agent_id_base = 0;
  c->drone_size[0] = 0.4; c->drone_size[1] = 0.4; c->drone_size[2] = 0.45;
  c->replan_risk_rate = 0.f;
  c->num_resample = 10;
  c->max_tau = 2.0; c->max_vel = 2.0; c->max_acc = 6.0; c->w_time = 5.0; c->horizon = 5.0;
  c->lambda_heu = 5.0; c->resolution_astar = 0.15; c->time_resolution_astar = 0.3;
  c->allocate_num = 10000; c->check_num = 1; c->tolerance = 1;
  c->corridor_tau = 0.3; c->init_range = 1.2; c->shrink_size = 0.2;
  c->opt_max_vel = 2.0; c->opt_max_acc = 6.0;
  c->max_cloud_points = 1 << 18;
  c->max_cylinders = 256;
  c->max_obstacle_points = 16384;
  c->osqp_max_iter = 4000;
  c->osqp_adaptive_rho_interval = 25;
  c->osqp_check_termination = 25;
  c->osqp_eps_abs = 1e-3; c->osqp_eps_rel = 1e-3;
}

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
// getParticlesWithRisk(idx_agent, t0) of drone ego_id's coordinator for nq queries (tests/test_ref_pins.py compares it
// with the reference's ParticleATC): out_n[q] = -1 when it returns false, else the particle count
void orc_particles_with_risk(const pop_config* cfg, int ego_id, const double* offs, const int32_t* counts,
                             const pop_traj_record* recs, int nq, const int32_t* q_agent, const double* q_t0,
                             double* out_pts, float* out_risks, int cap, int32_t* out_n) {
  Swarm s = make_swarm(*cfg, offs, counts, recs);
  for (int q = 0; q < nq; q++) {
    std::vector<V3d> pts;
    std::vector<float> risks;
    const bool ok = particles_with_risk(s, ego_id, q_agent[q], q_t0[q], pts, risks);
    out_n[q] = ok ? static_cast<int32_t>(pts.size()) : -1;
    for (int i = 0; ok && i < static_cast<int>(pts.size()) && i < cap; i++) {
      double* o = out_pts + (static_cast<size_t>(q) * cap + i) * 3;
      o[0] = pts[i].x; o[1] = pts[i].y; o[2] = pts[i].z;
      out_risks[static_cast<size_t>(q) * cap + i] = risks[i];
    }
  }
}
// the standard-normal triples the peer resampling consumes (replan_risk_rate > 0), for every swarm built afterwards
void orc_set_resample_noise(const float* z, int n_triples) { g_resample_noise.assign(z, z + 3 * static_cast<size_t>(n_triples)); }
// What libstdc++ hands ParticleATC::getParticlesWithRisk (particles.cpp:382-392): std::default_random_engine(seed) feeding
// three std::normal_distribution<float> objects in turn (each keeps its own spare value).  Generated with unit
// deviation: the distribution returns ret * stddev + mean, so these are the `ret`s.
void orc_libstdcxx_normal_triples(unsigned long seed, int n_triples, float* out) {
  std::default_random_engine gen(seed);
  std::normal_distribution<float> nx(0.0, 1.0f), ny(0.0, 1.0f), nz(0.0, 1.0f);
  for (int i = 0; i < n_triples; i++) { out[3 * i] = nx(gen); out[3 * i + 1] = ny(gen); out[3 * i + 2] = nz(gen); }
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
void orc_map_inflate_query(void* h, int mode, const double* pos, const double* t, int n, int8_t* out) {
  const Grid& g = static_cast<OrcMap*>(h)->grid;
  for (int i = 0; i < n; i++)
    out[i] = static_cast<int8_t>(inflate_query(g, V3d{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}, mode, t[i]));
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
// parity helper: load risk_maps_ ([V][T]) as received from "map/future_risk" (risk_base.cpp:60-68)
void orc_map_set_risk(void* h, const float* in) {
  Grid& g = static_cast<OrcMap*>(h)->grid;
  std::copy(in, in + g.risk.size(), g.risk.begin());
}
void orc_map_set_time(void* h, double t_map) { static_cast<OrcMap*>(h)->grid.t_map = t_map; }
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


// sdlp::linprog<d>
double orc_linprog(int d, const double* c, const double* A, const double* b, int m, double* x, uint64_t seed) {
  uint64_t s = seed;
  return linprog(d, c, A, b, m, x, &s);
}

// firi::firi on explicit inputs; returns rows (0 when firi returned false)
int orc_firi(const double* bd24, const double* pc, int n, const double* a, const double* b, int iterations,
             uint64_t seed, double* out_h, int cap_rows) {
  std::vector<V3d> pts(n);
  for (int i = 0; i < n; i++) pts[i] = V3d{pc[3 * i], pc[3 * i + 1], pc[3 * i + 2]};
  Poly poly;
  double r[3] = {1, 1, 1};
  uint64_t s = seed;
  bool ok = firi(bd24, pts, V3d{a[0], a[1], a[2]}, V3d{b[0], b[1], b[2]}, poly, r, iterations, &s);
  if (!ok) return 0;
  int rows = poly.rows();
  for (int i = 0; i < std::min(rows, cap_rows) * 4; i++) out_h[i] = poly.h[i];
  return rows;
}

// the same with the caller's initial radii r (in/out), as firi::firi's own signature has them
int orc_firi_r(const double* bd24, const double* pc, int n, const double* a, const double* b, int iterations,
               uint64_t seed, double* r, double* out_h, int cap_rows) {
  std::vector<V3d> pts(n);
  for (int i = 0; i < n; i++) pts[i] = V3d{pc[3 * i], pc[3 * i + 1], pc[3 * i + 2]};
  Poly poly;
  uint64_t s = seed;
  if (!firi(bd24, pts, V3d{a[0], a[1], a[2]}, V3d{b[0], b[1], b[2]}, poly, r, iterations, &s)) return 0;
  const int rows = poly.rows();
  for (int i = 0; i < std::min(rows, cap_rows) * 4; i++) out_h[i] = poly.h[i];
  return rows;
}

// corridor stage; out_hpoly = POP_MAX_PIECES x POP_MAX_FACES x 4; returns number of corridors
int orc_corridor(void* h, const double* path, int n, const double* start9, double traj_start_time,
                 double* out_hpoly, int32_t* out_rows, double* out_goal) {
  OrcMap* m = static_cast<OrcMap*>(h);
  std::vector<double> pv(path, path + static_cast<size_t>(n) * 6);
  CorridorResult r = corridor_stage(m->grid, m->cfg, pv, start9, traj_start_time);
  if (!r.ok) return 0;
  int np = static_cast<int>(r.polys.size());
  for (int i = 0; i < std::min(np, POP_MAX_PIECES); i++) {
    int rows = r.polys[i].rows();
    out_rows[i] = rows;
    for (int k = 0; k < std::min(rows, POP_MAX_FACES) * 4; k++)
      out_hpoly[(static_cast<size_t>(i) * POP_MAX_FACES) * 4 + k] = r.polys[i].h[k];
  }
  for (int k = 0; k < 6; k++) out_goal[k] = r.goal[k];
  return np;
}


static std::vector<Poly> polys_from(const double* hpoly, const int32_t* rows, int M) {
  std::vector<Poly> polys(M);
  for (int i = 0; i < M; i++)
    polys[i].h.assign(hpoly + static_cast<size_t>(i) * POP_MAX_FACES * 4,
                      hpoly + static_cast<size_t>(i) * POP_MAX_FACES * 4 + static_cast<size_t>(rows[i]) * 4);
  return polys;
}

// BezierOpt::setup: returns m (rows); A (m x n), l, u, Q (n x n) written when non-null
int orc_qp_setup(int M, const double* init9, const double* goal9, const double* dur, const double* hpoly,
                 const int32_t* rows, double vmax, double amax, double* outQ, double* outA, double* outl,
                 double* outu) {
  QpProblem p;
  qp_setup(p, init9, goal9, std::vector<double>(dur, dur + M), polys_from(hpoly, rows, M), vmax, amax);
  if (outQ) std::copy(p.Q.begin(), p.Q.end(), outQ);
  if (outA) std::copy(p.A.begin(), p.A.end(), outA);
  if (outl) std::copy(p.l.begin(), p.l.end(), outl);
  if (outu) std::copy(p.u.begin(), p.u.end(), outu);
  return p.m;
}

// BezierOpt::setup + optimize: returns OSQP status_val; out_cpts = 5M x 3
int orc_qp_solve(const pop_config* cfg, int M, const double* init9, const double* goal9, const double* dur,
                 const double* hpoly, const int32_t* rows, double vmax, double amax, double* out_cpts,
                 int* out_iters) {
  QpProblem p;
  qp_setup(p, init9, goal9, std::vector<double>(dur, dur + M), polys_from(hpoly, rows, M), vmax, amax);
  OsqpSettings S;
  S.eps_abs = cfg->osqp_eps_abs; S.eps_rel = cfg->osqp_eps_rel; S.max_iter = cfg->osqp_max_iter;
  S.check_termination = cfg->osqp_check_termination; S.adaptive_rho_interval = cfg->osqp_adaptive_rho_interval;
  OsqpResult r = osqp_solve(p, S);
  for (int i = 0; i < p.n; i++) out_cpts[i] = r.x[i];
  if (out_iters) *out_iters = r.iters;
  return r.status;
}

// the point sets isSafeAfterOpt would hand to the separator, per active peer: |A|, |B|, B's points; returns the LP count
int orc_safe_after_opt_sets(const pop_config* cfg, const double* offs, const int32_t* counts, const pop_traj_record* recs,
                            int ego_id, const double* cpts, int M, double t0, double* out, int cap) {
  Swarm s = make_swarm(*cfg, offs, counts, recs);
  std::vector<double> trace;
  g_sep_trace = &trace;
  safe_after_opt(s, ego_id, std::vector<double>(cpts, cpts + 15 * M), M, t0);
  g_sep_trace = nullptr;
  int lps = 0;
  for (size_t i = 0; i < trace.size();) { lps++; i += 2 + 3 * static_cast<size_t>(trace[i + 1]); }
  for (size_t i = 0; i < trace.size() && static_cast<int>(i) < cap; i++) out[i] = trace[i];
  return lps;
}
// the ADMM solver alone on a dense problem  min 1/2 x'Qx  s.t. l <= Ax <= u  (what IOSQP::setMats hands to OSQP);
// lets the reference's own planner, compiled into oracle/_ref/libref.so, run end to end (ref_glue_planner.cpp)
int orc_osqp_dense(const pop_config* cfg, int n, int m, const double* Q, const double* A, const double* l, const double* u,
                   double* x, int* iters) {
  QpProblem p;
  p.n = n; p.m = m;
  p.Q.assign(Q, Q + static_cast<size_t>(n) * n);
  p.A.assign(A, A + static_cast<size_t>(m) * n);
  p.l.assign(l, l + m);
  p.u.assign(u, u + m);
  OsqpSettings S;
  S.eps_abs = cfg->osqp_eps_abs; S.eps_rel = cfg->osqp_eps_rel; S.max_iter = cfg->osqp_max_iter;
  S.check_termination = cfg->osqp_check_termination; S.adaptive_rho_interval = cfg->osqp_adaptive_rho_interval;
  OsqpResult r = osqp_solve(p, S);
  for (int i = 0; i < n; i++) x[i] = r.x[i];
  if (iters) *iters = r.iters;
  return r.status;
}
int orc_separable(const double* A, int na, const double* B, int nb) {
  std::vector<V3d> a(na), b(nb);
  for (int i = 0; i < na; i++) a[i] = V3d{A[3 * i], A[3 * i + 1], A[3 * i + 2]};
  for (int i = 0; i < nb; i++) b[i] = V3d{B[3 * i], B[3 * i + 1], B[3 * i + 2]};
  return separable(a, b) ? 1 : 0;
}


// one replan for the agent owning this map (map already updated, peers included);
// offs/counts/recs may be null (single agent)
int orc_replan(void* h, const double* offs, const int32_t* counts, const pop_traj_record* recs, int ego_id,
               const double* start9, const double* goal3, double t_plan, pop_traj_record* out, int* stats) {
  OrcMap* m = static_cast<OrcMap*>(h);
  pop_config c = m->cfg;
  c.agent_id_base = ego_id;  // LP shuffle seed base = global agent id
  if (recs) {
    Swarm s = make_swarm(m->cfg, offs, counts, recs);
    return replan(m->grid, c, &s, ego_id, start9, goal3, t_plan, out, stats);
  }
  return replan(m->grid, c, nullptr, ego_id, start9, goal3, t_plan, out, stats);
}

// isTrajSafe for nq (now, T) pairs on one trajectory record: out[q] 1 safe / 0, t_unsafe[q] relative time or -1
void orc_traj_safe(void* h, const pop_traj_record* rec, double tau, int nq, const double* now, const double* T, int8_t* out,
                   double* t_unsafe) {
  const Grid& g = static_cast<OrcMap*>(h)->grid;
  for (int q = 0; q < nq; q++) out[q] = traj_safe(g, *rec, tau, now[q], T[q], t_unsafe ? &t_unsafe[q] : nullptr) ? 1 : 0;
}
int orc_safe_after_opt(const pop_config* cfg, const double* offs, const int32_t* counts,
                       const pop_traj_record* recs, int ego_id, const double* cpts, int M, double t0) {
  Swarm s = make_swarm(*cfg, offs, counts, recs);
  std::vector<double> x(cpts, cpts + 15 * M);
  return safe_after_opt(s, ego_id, x, M, t0) ? 1 : 0;
}

// multi-threaded batch used by bench.py's CPU-baseline legs: full cycle (map update incl. peers +
// replan) for agents [a0, a1) of a swarm sharing one world; returns seconds of wall time
double orc_cycle_batch(const pop_config* cfg, int a0, int a1, int n_threads, const float* cloud, int n_cloud,
                       const pop_cylinder* cyl, int n_cyl, const float* poses, const double* offs,
                       const int32_t* counts, const pop_traj_record* recs, const double* start, const double* goal,
                       double t_map, double t_plan, int32_t* status_out, pop_traj_record* recs_out);

}  // extern "C"

#include <atomic>
#include <chrono>
#include <thread>

extern "C" double orc_cycle_batch(const pop_config* cfg, int a0, int a1, int n_threads, const float* cloud,
                                  int n_cloud, const pop_cylinder* cyl, int n_cyl, const float* poses,
                                  const double* offs, const int32_t* counts, const pop_traj_record* recs,
                                  const double* start, const double* goal, double t_map, double t_plan,
                                  int32_t* status_out, pop_traj_record* recs_out) {
  const auto t0 = std::chrono::steady_clock::now();
  std::atomic<int> next(a0);
  Swarm swarm;
  if (recs) swarm = make_swarm(*cfg, offs, counts, recs);
  auto worker = [&]() {
    Grid g;
    g.init(*cfg);
    for (;;) {
      const int a = next.fetch_add(1);
      if (a >= a1) break;
      const int local = a - cfg->agent_id_base;
      g.pose = V3f{poses[3 * local], poses[3 * local + 1], poses[3 * local + 2]};
      gt_update(g, cloud, n_cloud, cyl, n_cyl, t_map);
      if (recs) add_other_agents(g, swarm, a);
      pop_config c = *cfg;
      c.agent_id_base = a;
      status_out[local] = replan(g, c, recs ? &swarm : nullptr, a, start + 9 * local, goal + 3 * local, t_plan,
                                 recs_out ? &recs_out[local] : nullptr, nullptr);   // records of failed replans stay untouched
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < std::max(1, n_threads); i++) th.emplace_back(worker);
  for (auto& t : th) t.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

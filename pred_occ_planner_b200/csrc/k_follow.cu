// k_follow.cu — what happens to a trajectory after the optimiser: trajectory-server
// sampling, the closed-loop state hand-over of a perfectly tracking drone, the evaluation metrics, and the
// BezierTraj wire format.
//
// Reference: traj_server/src/bezier_traj_server.cpp:194-275 (queue filled with getPos/getVel/getAcc samples
// of the received Bezier), uav_simulator/fake_drone (position command -> odometry, perfect tracking),
// eval_helper/include/eval_helper/multi_eval.hpp:226-275 (getMinObsDist / getMinUAVDist),
// traj_utils/msg/BezierTraj.msg (ROS1 serialisation of the trajectory message).
#include <cmath>
#include <cstring>

#include "pop_bezier.cuh"
#include "pop_geom.cuh"
#include "pop_internal.h"

namespace {

// traj_.getPos / getVel / getAcc at absolute time t; valid iff a trajectory exists and start <= t <= start + duration
// (fillTrajQueue :213-221 returns outside that window)
__device__ bool sample_record(const pop_traj_record& r, double t, double out[9]) {
  if (r.n_pieces <= 0 || r.n_pieces > POP_MAX_PIECES) return false;
  double T = 0.0;
  for (int k = 0; k < r.n_pieces; k++) T += static_cast<double>(r.duration[k]);
  const double tt = t - r.start_time;
  if (tt < 0.0 || tt > T) return false;
  pop_bezier_eval(r.duration, r.n_pieces, r.cpts, tt, 0, out);
  pop_bezier_eval(r.duration, r.n_pieces, r.cpts, tt, 1, out + 3);
  pop_bezier_eval(r.duration, r.n_pieces, r.cpts, tt, 2, out + 6);
  return true;
}

__global__ void k_traj_sample(const pop_traj_record* __restrict__ recs, int n_agents, const double* __restrict__ t_abs,
                              int n_t, double* __restrict__ out, int8_t* __restrict__ valid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_agents * n_t) return;
  const int a = i / n_t, k = i % n_t;
  double pva[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const bool ok = sample_record(recs[a], t_abs[k], pva);
  for (int q = 0; q < 9; q++) out[static_cast<size_t>(i) * 9 + q] = pva[q];
  valid[i] = ok ? 1 : 0;
}

// BaselinePlanner::isTrajSafe(T) (plan_manager/src/baseline.cpp:45-68 = baseline_fake.cpp:53-76) for every local agent:
// its own trajectory (the last one it planned; pieces of tau = corridor_tau seconds in double, as the optimiser was
// given them — not the float32 durations of the wire record) sampled every 0.1 s from t_now to min(T, duration) by the
// reference's running sum `t += 0.1`; unsafe at the first sample whose getClearOcccupancy(pos, dt) answers 1.
template <typename C>
__global__ void k_traj_safe(PopGeom g, const uint8_t* __restrict__ grid, const float* __restrict__ poses,
                            const pop_traj_record* __restrict__ send, int n_local, double tau, double t_map, double t_now,
                            double T, int8_t* __restrict__ safe, double* __restrict__ t_unsafe) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_local) return;
  const pop_traj_record& r = send[a];
  int8_t ok = 1;
  double tu = -1.0;
  double t0 = __dsub_rn(t_now, r.start_time);
  if (t0 < 0) t0 = 0;
  if (!(t0 > T) && r.n_pieces > 0 && r.n_pieces <= POP_MAX_PIECES) {
    double durs[POP_MAX_PIECES], dur = 0.0;
    for (int k = 0; k < POP_MAX_PIECES; k++) durs[k] = tau;
    for (int k = 0; k < r.n_pieces; k++) dur = __dadd_rn(dur, tau);
    const double Tc = T > dur ? dur : T;
    const uint8_t* ag = grid + a * g.agent_stride;
    const float px = poses[3 * a], py = poses[3 * a + 1], pz = poses[3 * a + 2];
    for (double t = t0; t < Tc; t = __dadd_rn(t, 0.1)) {
      double pos[3];
      pop_bezier_eval(durs, r.n_pieces, r.cpts, t, 0, pos);
      const double dt = __dsub_rn(__dadd_rn(t, r.start_time), t_map);
      int ti = static_cast<int>(floor(__ddiv_rn(dt, static_cast<double>(g.dt))));
      ti = ti > g.nt - 1 ? g.nt - 1 : ti;
      if (geom_query<C>(g, ag, px, py, pz, pos[0], pos[1], pos[2], ti) == 1) { ok = 0; tu = t; break; }
    }
  }
  safe[a] = ok;
  t_unsafe[a] = tu;
}

// perfect tracking: the state handed to the next replan is the commanded state at t_next; past the end of the
// trajectory the drone holds the last position; without a trajectory it stays where it is
__global__ void k_swarm_advance(const pop_traj_record* __restrict__ send, int n_local, double t_next,
                                double* __restrict__ start, float* __restrict__ poses) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_local) return;
  const pop_traj_record& r = send[a];
  double* s = start + 9 * a;
  if (r.n_pieces > 0 && r.n_pieces <= POP_MAX_PIECES && t_next >= r.start_time) {
    double pva[9];
    if (sample_record(r, t_next, pva)) {
      for (int q = 0; q < 9; q++) s[q] = pva[q];
    } else {
      double T = 0.0;
      for (int k = 0; k < r.n_pieces; k++) T += static_cast<double>(r.duration[k]);
      pop_bezier_eval(r.duration, r.n_pieces, r.cpts, T, 0, pva);
      for (int q = 0; q < 3; q++) s[q] = pva[q];
      for (int q = 3; q < 9; q++) s[q] = 0.0;
    }
  } else {
    for (int q = 3; q < 9; q++) s[q] = 0.0;
  }
  for (int q = 0; q < 3; q++) poses[3 * a + q] = static_cast<float>(s[q]);
}

// getMinObsDist (cylinders) and getMinUAVDist for the local agents; peers are sampled from the exchange buffer
__global__ void k_eval_metrics(const double* __restrict__ start, int n_local, int id_base,
                               const pop_traj_record* __restrict__ all, int n_total, double t,
                               const PopCylDev* __restrict__ cyl, int n_cyl, double* __restrict__ min_obs,
                               double* __restrict__ min_uav) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_local) return;
  const double px = start[9 * a], py = start[9 * a + 1], pz = start[9 * a + 2];
  double mo = 1e3;
  for (int c = 0; c < n_cyl; c++) {
    if (cyl[c].type != 3) continue;
    const double dx = static_cast<double>(cyl[c].x) - px, dy = static_cast<double>(cyl[c].y) - py;
    const double d = sqrt(dx * dx + dy * dy + 0.0) - cyl[c].w * 0.5;
    if (d < mo) mo = d;
  }
  double mu = 1e3;
  for (int i = 0; i < n_total; i++) {
    if (i == id_base + a) continue;
    double q[9];
    double ox, oy, oz;
    if (i >= id_base && i < id_base + n_local) {   // a local peer: its current state
      ox = start[9 * (i - id_base)]; oy = start[9 * (i - id_base) + 1]; oz = start[9 * (i - id_base) + 2];
    } else if (sample_record(all[i], t, q)) {
      ox = q[0]; oy = q[1]; oz = q[2];
    } else {
      continue;
    }
    const double dx = ox - px, dy = oy - py, dz = oz - pz;
    const double d = sqrt(dx * dx + dy * dy + dz * dz);
    if (d < mu) mu = d;
  }
  min_obs[a] = mo;
  min_uav[a] = mu;
}

// ---- ROS1 wire helpers -------------------------------------------------------------------------
inline void put_u32(uint8_t*& p, uint32_t v) { memcpy(p, &v, 4); p += 4; }
inline uint32_t get_u32(const uint8_t*& p) { uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
// ros::Time::fromSec: sec = floor(t), nsec = round((t - sec) * 1e9), normalised
inline void put_time(uint8_t*& p, double t) {
  double sec = std::floor(t);
  double nsec = std::round((t - sec) * 1e9);
  if (nsec >= 1e9) { sec += 1.0; nsec -= 1e9; }
  put_u32(p, static_cast<uint32_t>(sec));
  put_u32(p, static_cast<uint32_t>(nsec));
}
inline double get_time(const uint8_t*& p) {
  const uint32_t s = get_u32(p), ns = get_u32(p);
  return static_cast<double>(s) + 1e-9 * static_cast<double>(ns);   // ros::Time::toSec
}

}  // namespace

extern "C" {

int pop_traj_sample_batch(pop_handle* h, int all_agents, const double* t_abs, int n_t, double* out_pva,
                          int8_t* out_valid) {
  if (!h || !t_abs || n_t < 1 || !out_pva || !out_valid) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  const int n = all_agents ? h->cfg.n_agents_total : h->cfg.n_agents_local;
  const pop_traj_record* recs = all_agents ? h->d_traj_all : h->d_traj_send;
  const size_t n_out = static_cast<size_t>(n) * n_t;
  const size_t bytes = sizeof(double) * n_t + sizeof(double) * 9 * n_out + n_out + 64;
  int rc = pop_ensure_scratch(h, bytes, 0);
  if (rc) return rc;
  double* d_t = static_cast<double*>(h->d_scratch);
  double* d_out = d_t + n_t;
  int8_t* d_valid = reinterpret_cast<int8_t*>(d_out + 9 * n_out);
  POP_CUDA(h, cudaMemcpyAsync(d_t, t_abs, sizeof(double) * n_t, cudaMemcpyHostToDevice, h->stream));
  k_traj_sample<<<static_cast<unsigned>((n_out + 127) / 128), 128, 0, h->stream>>>(recs, n, d_t, n_t, d_out, d_valid);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  POP_CUDA(h, cudaMemcpyAsync(out_pva, d_out, sizeof(double) * 9 * n_out, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(out_valid, d_valid, n_out, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

int pop_traj_safe_batch(pop_handle* h, double t_now, double T, int8_t* out_safe, double* out_t_unsafe) {
  if (!h || !out_safe) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->map_ready) return pop_fail(h, POP_E_STATE, "isTrajSafe before the first map update");
  const int nl = h->cfg.n_agents_local;
  int rc = pop_ensure_scratch(h, nl * (sizeof(double) + 1) + 64, 0);
  if (rc) return rc;
  double* d_tu = static_cast<double*>(h->d_scratch);
  int8_t* d_safe = reinterpret_cast<int8_t*>(d_tu + nl);
  if (h->cfg.cell_dtype == POP_CELL_U8)
    k_traj_safe<uint8_t><<<(nl + 63) / 64, 64, 0, h->stream>>>(h->geom, h->d_grid, h->d_poses, h->d_traj_send, nl, h->cfg.corridor_tau,
                                                             h->t_map, t_now, T, d_safe, d_tu);
  else
    k_traj_safe<float><<<(nl + 63) / 64, 64, 0, h->stream>>>(h->geom, h->d_grid, h->d_poses, h->d_traj_send, nl, h->cfg.corridor_tau,
                                                           h->t_map, t_now, T, d_safe, d_tu);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  POP_CUDA(h, cudaMemcpyAsync(out_safe, d_safe, nl, cudaMemcpyDeviceToHost, h->stream));
  if (out_t_unsafe) POP_CUDA(h, cudaMemcpyAsync(out_t_unsafe, d_tu, sizeof(double) * nl, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

int pop_swarm_advance(pop_handle* h, double t_next) {
  if (!h) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->astar) return pop_fail(h, POP_E_STATE, "no resident agent states: run a cycle first");
  const int nl = h->cfg.n_agents_local;
  k_swarm_advance<<<(nl + 127) / 128, 128, 0, h->stream>>>(h->d_traj_send, nl, t_next, popk_astar_start_dev(h), h->d_poses);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  return POP_OK;
}

int pop_swarm_states(pop_handle* h, double* out_start) {
  if (!h || !out_start) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->astar) return pop_fail(h, POP_E_STATE, "no resident agent states: run a cycle first");
  POP_CUDA(h, cudaMemcpyAsync(out_start, popk_astar_start_dev(h), sizeof(double) * 9 * h->cfg.n_agents_local,
                              cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

int pop_eval_metrics(pop_handle* h, double t, double* out_min_obs, double* out_min_uav) {
  if (!h || !out_min_obs || !out_min_uav) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->astar) return pop_fail(h, POP_E_STATE, "no resident agent states: run a cycle first");
  const int nl = h->cfg.n_agents_local;
  int rc = pop_ensure_scratch(h, sizeof(double) * 2 * nl, 0);
  if (rc) return rc;
  double* d_obs = static_cast<double*>(h->d_scratch);
  double* d_uav = d_obs + nl;
  k_eval_metrics<<<(nl + 63) / 64, 64, 0, h->stream>>>(popk_astar_start_dev(h), nl, h->cfg.agent_id_base, h->d_traj_all,
                                                      h->cfg.n_agents_total, t, h->d_cyl, h->n_cyl, d_obs, d_uav);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  POP_CUDA(h, cudaMemcpyAsync(out_min_obs, d_obs, sizeof(double) * nl, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(out_min_uav, d_uav, sizeof(double) * nl, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

// traj_utils/BezierTraj: int32 drone_id, int32 traj_id, time start_time, time pub_time, uint8 order,
// float32[] duration, geometry_msgs/Point[] cpts — ROS1 little-endian serialisation
int pop_traj_to_wire(const pop_traj_record* rec, double pub_time, uint8_t* out, int cap) {
  if (!rec || rec->n_pieces < 0 || rec->n_pieces > POP_MAX_PIECES) return POP_E_INVALID;
  const int M = rec->n_pieces;
  const int need = 4 + 4 + 8 + 8 + 1 + 4 + 4 * M + 4 + 24 * 5 * M;
  if (!out) return need;
  if (cap < need) return POP_E_CAPACITY;
  uint8_t* p = out;
  put_u32(p, static_cast<uint32_t>(rec->drone_id));
  put_u32(p, static_cast<uint32_t>(rec->traj_id));
  put_time(p, rec->start_time);
  put_time(p, pub_time);
  *p++ = 4;   // order (traj_utils/include/traj_utils/bernstein.hpp:22)
  put_u32(p, static_cast<uint32_t>(M));
  memcpy(p, rec->duration, 4 * M); p += 4 * M;
  put_u32(p, static_cast<uint32_t>(5 * M));
  memcpy(p, rec->cpts, 24 * 5 * M); p += 24 * 5 * M;
  return static_cast<int>(p - out);
}

int pop_traj_from_wire(const uint8_t* in, int n, pop_traj_record* rec, double* pub_time) {
  if (!in || !rec || n < 29) return POP_E_INVALID;
  const uint8_t* p = in;
  memset(rec, 0, sizeof(*rec));
  rec->drone_id = static_cast<int32_t>(get_u32(p));
  rec->traj_id = static_cast<int32_t>(get_u32(p));
  rec->start_time = get_time(p);
  const double pub = get_time(p);
  if (pub_time) *pub_time = pub;
  const int order = *p++;
  const uint32_t M = get_u32(p);
  if (order != 4 || M > POP_MAX_PIECES) return POP_E_UNSUPPORTED;
  if (n < static_cast<int>(29 + 4 * M + 4)) return POP_E_INVALID;
  memcpy(rec->duration, p, 4 * M); p += 4 * M;
  const uint32_t R = get_u32(p);
  if (R != 5 * M || n < static_cast<int>(33 + 4 * M + 24 * R)) return POP_E_INVALID;
  memcpy(rec->cpts, p, 24 * R); p += 24 * R;
  rec->n_pieces = static_cast<int32_t>(M);
  rec->valid = M > 0 ? 1 : 0;
  return static_cast<int>(p - in);
}

}  // extern "C"

// pop_dsp.cu — C ABI of the DSP particle map (include/pop.h, "K1d" section): configuration, input staging,
// the per-frame odometry checks of DSPMap::update (dsp_dynamic.h:165-204) and the parity views.
#include <cmath>
#include <cstring>

#include "k_dsp.cuh"

namespace {

// table generator used when the caller supplies no tables: splitmix64 -> Box-Muller.  The reference draws
// from std::default_random_engine seeded with time(NULL) (dsp_dynamic.h:1220-1231); any stream is as valid.
inline uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
inline double u01(uint64_t& s) { return (static_cast<double>(splitmix64(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

int check(pop_handle* h, int agent) {
  if (!h) return POP_E_INVALID;
  if (!h->dsp) return pop_fail(h, POP_E_STATE, "no particle map attached (pop_dsp_attach)");
  if (agent < 0 || agent >= h->cfg.n_agents_local) return pop_fail(h, POP_E_INVALID, "agent %d out of range", agent);
  return POP_OK;
}

int check_device_err(pop_handle* h) {
  int e = 0;
  POP_CUDA(h, cudaMemcpyAsync(&e, h->dsp->err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  if (e == 0) return POP_OK;
  POP_CUDA(h, cudaMemsetAsync(h->dsp->err, 0, sizeof(int), h->stream));
  if (e & 32) return pop_fail(h, POP_E_STATE, "particle re-binning did not reach its fixed point in %d rounds", kDspMaxIter);
  if (e & (2 | 4 | 16)) return pop_fail(h, POP_E_CAPACITY, "particle map capacity exceeded (flags 0x%x: 2 clusters, 4 dynamic clusters, 16 records)", e);
  if (e & 1) return pop_fail(h, POP_E_CAPACITY, "voxel-filter leaf grid too large");
  if (e & 8) return pop_fail(h, POP_E_UNSUPPORTED, "particle with non-zero vz (only LIMIT_MOVEMENT_IN_XY_PLANE builds are supported)");
  return pop_fail(h, POP_E_INVALID, "particle map error flags 0x%x", e);
}

// the odometry checks and static locals of DSPMap::update (:176-210), per agent
void make_frames(pop_handle* h, const int32_t* counts, const float* poses, const float* quats, double stamp,
                 int32_t* out_ok) {
  DspDev* d = h->dsp;
  for (int a = 0; a < d->nl; a++) {
    DspFrame& f = d->h_frame[a];
    memset(&f, 0, sizeof(f));
    const float* p = poses + 3 * a;
    const float* q = quats + 4 * a;
    if (d->first[a]) {
      for (int k = 0; k < 3; k++) d->last_pos[3 * a + k] = p[k];
      d->last_stamp[a] = stamp;
      d->first[a] = 0;
    }
    f.n_in = counts[a];
    for (int k = 0; k < 4; k++) f.q[k] = q[k];
    for (int k = 0; k < 3; k++) f.pos[k] = p[k];
    bool ok = !(std::fabs(q[0]) > 1.001f || std::fabs(q[1]) > 1.001f || std::fabs(q[2]) > 1.001f || std::fabs(q[3]) > 1.001f);
    // non-finite odometry passes the reference's comparisons and poisons the whole map; it is rejected like any
    // other invalid odometry
    for (int k = 0; k < 4; k++) ok = ok && std::isfinite(q[k]);
    for (int k = 0; k < 3; k++) ok = ok && std::isfinite(p[k]);
    ok = ok && std::isfinite(stamp);
    if (ok) {
      const float dx = p[0] - d->last_pos[3 * a], dy = p[1] - d->last_pos[3 * a + 1], dz = p[2] - d->last_pos[3 * a + 2];
      const float dt = static_cast<float>(stamp - d->last_stamp[a]);
      if (std::fabs(dx) > 10.f || std::fabs(dy) > 10.f || std::fabs(dz) > 10.f || dt < 0.f || dt > 10.f) {
        ok = false;
      } else {
        for (int k = 0; k < 3; k++) d->last_pos[3 * a + k] = p[k];
        d->last_stamp[a] = stamp;
        f.d[0] = -dx; f.d[1] = -dy; f.d[2] = -dz;
        f.dt = dt;
      }
    }
    f.ok = ok ? 1 : 0;
    if (out_ok) out_ok[a] = f.ok;
  }
}

int run_update(pop_handle* h, const float* in, int stride_pts, const int32_t* counts, const float* poses,
               const float* quats, double stamp, int32_t* out_ok, int filtered) {
  DspDev* d = h->dsp;
  const int A = d->nl;
  const int cap = filtered ? d->G.MAXPTS : d->G.MAXIN;
  const bool rendered = in == nullptr;   // clouds and their sizes are already on the device (pop_render_depth)
  std::vector<int32_t> zero_counts(A, 0);
  if (rendered) counts = zero_counts.data();
  d->use_cam_counts = rendered;
  for (int a = 0; a < A; a++)
    if (counts[a] < 0 || counts[a] > cap) return pop_fail(h, POP_E_CAPACITY, "agent %d: %d points exceed the capacity %d", a, counts[a], cap);
  if (filtered)   // points that skip the voxel filter (which drops non-finite ones, as PCL does) must be finite
    for (int a = 0; a < A; a++)
      for (int i = 0; i < 3 * counts[a]; i++)
        if (!std::isfinite(in[static_cast<size_t>(a) * stride_pts * 3 + i]))
          return pop_fail(h, POP_E_INVALID, "agent %d: non-finite filtered point", a);
  make_frames(h, counts, poses, quats, stamp, out_ok);
  POP_CUDA(h, cudaMemcpyAsync(d->frame, d->h_frame.data(), sizeof(DspFrame) * A, cudaMemcpyHostToDevice, h->stream));
  float* dst = filtered ? d->pts : d->cam;
  for (int a = 0; a < A; a++)
    if (counts[a] > 0)
      POP_CUDA(h, cudaMemcpyAsync(dst + static_cast<size_t>(a) * cap * 3, in + static_cast<size_t>(a) * stride_pts * 3,
                                  sizeof(float) * 3 * counts[a], cudaMemcpyHostToDevice, h->stream));
  if (filtered) POP_CUDA(h, cudaMemcpyAsync(d->npts, counts, sizeof(int) * A, cudaMemcpyHostToDevice, h->stream));
  // the SOGM this map feeds is centred on the sensor (risk_base.cpp: pose_ from the odometry callbacks)
  POP_CUDA(h, cudaMemcpyAsync(h->d_poses, poses, sizeof(float) * 3 * A, cudaMemcpyHostToDevice, h->stream));
  int rc = popk_dsp_update(h, filtered);
  if (rc) return rc;
  h->t_map = stamp;
  h->map_ready = true;
  // RiskBase::futureRiskCallback -> addOtherAgents (risk_base.cpp:70-72); an agent whose update was rejected
  // published no new future risk, so its grid (peers included) stays as it was
  bool all_ok = true;
  std::vector<int> okv(A);
  for (int a = 0; a < A; a++) { okv[a] = d->h_frame[a].ok; all_ok = all_ok && okv[a]; }
  if (!all_ok) {
    POP_CUDA(h, cudaMemcpyAsync(d->ok_dev, okv.data(), sizeof(int) * A, cudaMemcpyHostToDevice, h->stream));
    POP_CUDA(h, cudaStreamSynchronize(h->stream));   // okv is a stack-owned staging buffer
    h->d_agent_ok = d->ok_dev;
  }
  rc = popk_peer_project(h);
  h->d_agent_ok = nullptr;
  if (rc) return rc;
  return check_device_err(h);
}

}  // namespace

extern "C" {

void pop_dsp_default_config(const pop_config* g, pop_dsp_config* c) {
  memset(c, 0, sizeof(*c));
  // plan_env/include/plan_env/map_parameters.h:9-53
  c->half_fov_h = 43; c->half_fov_v = 29; c->angle_resolution = 1;
  c->max_particle_num_voxel = 7;
  const long long V = static_cast<long long>(g->nx) * g->ny * g->nz;
  c->pyramid_capacity = static_cast<int>(static_cast<long long>(V * 7 + 1e5) / (360 * 180) * 2);
  c->max_points_per_pyramid = 100;
  c->max_points = 5000;           // risk_mapping_node.cpp:47
  c->max_input_points = 32768;
  c->new_born_particle_number_each_point = 20;  // risk_mapping_node.cpp:236
  c->dynamic_cluster_max_point_num = 200;
  c->dynamic_cluster_max_center_height = 1.5f;
  c->new_born_particle_weight = 0.0001f;        // :238
  c->sigma_observation = 0.1f;                  // plan_manager/config/real_cfg.yaml:55
  c->kappa = 0.01f;
  c->p_detection = 0.95f;
  c->voxel_filter_resolution = 0.15f;
  c->obstacle_thickness_for_occlusion = 0.3f;
  c->occupancy_threshold = g->risk_threshold;
  c->position_prediction_stddev = 0.01f;        // setPredictionVariance(0.01, 0.01) :232
  c->velocity_prediction_stddev = 0.01f;
  for (int k = 0; k < POP_DSP_MAX_T; k++) c->prediction_future_time[k] = 0.3f * static_cast<float>(k + 1);
  c->n_randoms = 1 << 22;
  c->seed = 20221023ull;
}

int pop_dsp_attach(pop_handle* h, const pop_dsp_config* c) {
  if (!h || !c) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (h->cfg.cell_dtype != POP_CELL_F32) return pop_fail(h, POP_E_UNSUPPORTED, "the particle map needs f32 SOGM cells");
  if (c->max_particle_num_voxel < 1 || 2 * c->max_particle_num_voxel > kDspSlotStride)
    return pop_fail(h, POP_E_INVALID, "max_particle_num_voxel must be in [1, %d]", kDspSlotStride / 2);
  if (c->max_points < 1 || c->max_points > 8192) return pop_fail(h, POP_E_INVALID, "max_points must be in [1, 8192]");
  if (c->max_input_points < 1 || c->max_input_points > (1 << 20)) return pop_fail(h, POP_E_INVALID, "max_input_points must be in [1, 2^20]");
  if (c->new_born_particle_number_each_point < 1 || c->new_born_particle_number_each_point > 32)
    return pop_fail(h, POP_E_INVALID, "new_born_particle_number_each_point must be in [1, 32]");
  if (c->pyramid_capacity < 1 || c->max_points_per_pyramid < 2 || c->n_randoms < 1024 || h->cfg.nt > POP_DSP_MAX_T ||
      c->angle_resolution < 1 || c->half_fov_h < 1 || c->half_fov_v < 1)
    return pop_fail(h, POP_E_INVALID, "bad particle-map configuration");
  if (static_cast<long long>(h->geom.V) * 2 * c->max_particle_num_voxel >= (1ll << 32))
    return pop_fail(h, POP_E_UNSUPPORTED, "grid too large for 32-bit particle sweep keys");
  int rc = popk_dsp_alloc(h, c);
  if (rc) { popk_dsp_free(h); return rc; }
  // default tables
  const int n = c->n_randoms;
  std::vector<float> pg(n), vg(n);
  std::vector<int32_t> ur(n);
  uint64_t s = c->seed;
  for (int i = 0; i < n; i++) {
    const double r = std::sqrt(-2.0 * std::log(u01(s))), th = 6.283185307179586 * u01(s);
    pg[i] = static_cast<float>(c->position_prediction_stddev * r * std::cos(th));
    vg[i] = static_cast<float>(c->velocity_prediction_stddev * r * std::sin(th));
    ur[i] = static_cast<int32_t>(splitmix64(s) >> 33);
  }
  return pop_dsp_set_random_tables(h, pg.data(), vg.data(), ur.data(), n);
}

int pop_dsp_set_random_tables(pop_handle* h, const float* pg, const float* vg, const int32_t* ur, int n) {
  PopDeviceGuard guard(h);
  int rc = check(h, 0);
  if (rc) return rc;
  DspDev* d = h->dsp;
  if (!pg || !vg || !ur || n != d->G.nrand) return pop_fail(h, POP_E_INVALID, "tables must hold n_randoms=%d entries", d->G.nrand);
  POP_CUDA(h, cudaMemcpyAsync(d->pg, pg, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d->vg, vg, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d->ur, ur, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemsetAsync(d->seq, 0, sizeof(int) * 4 * d->nl, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

int pop_dsp_update(pop_handle* h, const float* cam, const int32_t* counts, const float* poses, const float* quats,
                   double stamp, int32_t* out_ok) {
  PopDeviceGuard guard(h);
  int rc = check(h, 0);
  if (rc) return rc;
  if (!cam || !counts || !poses || !quats) return POP_E_INVALID;
  return run_update(h, cam, h->dsp->G.MAXIN, counts, poses, quats, stamp, out_ok, 0);
}

int pop_dsp_update_rendered(pop_handle* h, const float* poses, const float* quats, double stamp, int32_t* out_ok) {
  PopDeviceGuard guard(h);
  int rc = check(h, 0);
  if (rc) return rc;
  if (!poses || !quats) return POP_E_INVALID;
  if (!h->dsp->have_camera || !h->dsp->use_cam_counts)
    return pop_fail(h, POP_E_STATE, "no rendered clouds: call pop_render_depth first");
  return run_update(h, nullptr, h->dsp->G.MAXIN, nullptr, poses, quats, stamp, out_ok, 0);
}

int pop_dsp_update_filtered(pop_handle* h, const float* pts, const int32_t* counts, const float* poses,
                            const float* quats, double stamp, int32_t* out_ok) {
  PopDeviceGuard guard(h);
  int rc = check(h, 0);
  if (rc) return rc;
  if (!pts || !counts || !poses || !quats) return POP_E_INVALID;
  return run_update(h, pts, h->dsp->G.MAXPTS, counts, poses, quats, stamp, out_ok, 1);
}

int pop_dsp_filter(pop_handle* h, int agent, const float* cam, int n, float* out, int32_t* n_out) {
  PopDeviceGuard guard(h);
  int rc = check(h, agent);
  if (rc) return rc;
  DspDev* d = h->dsp;
  if (n < 0 || n > d->G.MAXIN || !out || !n_out) return POP_E_INVALID;
  for (int a = 0; a < d->nl; a++) { memset(&d->h_frame[a], 0, sizeof(DspFrame)); }
  d->h_frame[agent].n_in = n;
  d->use_cam_counts = false;
  POP_CUDA(h, cudaMemcpyAsync(d->frame, d->h_frame.data(), sizeof(DspFrame) * d->nl, cudaMemcpyHostToDevice, h->stream));
  if (n > 0)
    POP_CUDA(h, cudaMemcpyAsync(d->cam + static_cast<size_t>(agent) * d->G.MAXIN * 3, cam, sizeof(float) * 3 * n,
                                cudaMemcpyHostToDevice, h->stream));
  rc = popk_dsp_filter_only(h);
  if (rc) return rc;
  int cnt = 0;
  POP_CUDA(h, cudaMemcpyAsync(&cnt, d->npts + agent, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  *n_out = cnt;
  if (cnt > 0)
    POP_CUDA(h, cudaMemcpy(out, d->pts + static_cast<size_t>(agent) * d->G.MAXPTS * 3, sizeof(float) * 3 * cnt,
                           cudaMemcpyDeviceToHost));
  return check_device_err(h);
}

int pop_dsp_download_particles(pop_handle* h, int agent, float* out) {
  PopDeviceGuard guard(h);
  int rc = check(h, agent);
  if (rc) return rc;
  const size_t bytes = sizeof(float) * 8 * static_cast<size_t>(h->dsp->G.V) * h->dsp->G.S;
  rc = pop_ensure_scratch(h, bytes, 0);
  if (rc) return rc;
  rc = popk_dsp_download_particles(h, agent, static_cast<float*>(h->d_scratch));
  if (rc) return rc;
  POP_CUDA(h, cudaMemcpyAsync(out, h->d_scratch, bytes, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

int pop_dsp_download_voxels(pop_handle* h, int agent, float* out) {
  PopDeviceGuard guard(h);
  int rc = check(h, agent);
  if (rc) return rc;
  const DspGeo& G = h->dsp->G;
  POP_CUDA(h, cudaMemcpyAsync(out, h->dsp->vobj + static_cast<size_t>(agent) * G.V, sizeof(float4) * G.V,
                              cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

int pop_dsp_download_observations(pop_handle* h, int agent, float* pc, int32_t* counts, float* max_length) {
  PopDeviceGuard guard(h);
  int rc = check(h, agent);
  if (rc) return rc;
  DspDev* d = h->dsp;
  const DspGeo& G = d->G;
  const size_t n = static_cast<size_t>(G.NP) * G.OCAP;
  std::vector<float4> tmp(n);
  POP_CUDA(h, cudaMemcpyAsync(tmp.data(), d->obs + agent * n, sizeof(float4) * n, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(counts, d->obs_num + static_cast<size_t>(agent) * G.NP, sizeof(int) * G.NP, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(max_length, d->obs_maxlen + static_cast<size_t>(agent) * G.NP, sizeof(float) * G.NP, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i < n; i++) {
    float* o = pc + 5 * i;
    o[0] = tmp[i].x; o[1] = tmp[i].y; o[2] = tmp[i].z; o[3] = tmp[i].w;
    o[4] = std::sqrt(tmp[i].x * tmp[i].x + tmp[i].y * tmp[i].y + tmp[i].z * tmp[i].z);
  }
  return POP_OK;
}

int pop_dsp_download_pyramids(pop_handle* h, int agent, int32_t* out) {
  PopDeviceGuard guard(h);
  int rc = check(h, agent);
  if (rc) return rc;
  DspDev* d = h->dsp;
  const DspGeo& G = d->G;
  const size_t n = static_cast<size_t>(G.NP) * G.PCAP;
  std::vector<int2> ref(n);
  std::vector<int> cnt(G.NP);
  POP_CUDA(h, cudaMemcpyAsync(ref.data(), d->pyr_ref + agent * n, sizeof(int2) * n, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(cnt.data(), d->pyr_n + static_cast<size_t>(agent) * G.NP, sizeof(int) * G.NP, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int p = 0; p < G.NP; p++)
    for (int s = 0; s < G.PCAP; s++) {
      int32_t* o = out + (static_cast<size_t>(p) * G.PCAP + s) * 3;
      const bool v = s < cnt[p];
      o[0] = v ? 1 : 0;
      o[1] = v ? ref[static_cast<size_t>(p) * G.PCAP + s].x : 0;
      o[2] = v ? ref[static_cast<size_t>(p) * G.PCAP + s].y : 0;
    }
  return POP_OK;
}

int pop_dsp_download_points_with_velocity(pop_handle* h, int agent, float* out, int cap, int32_t* n_out) {
  PopDeviceGuard guard(h);
  int rc = check(h, agent);
  if (rc) return rc;
  DspDev* d = h->dsp;
  int n = 0;
  POP_CUDA(h, cudaMemcpyAsync(&n, d->ncv + agent, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  *n_out = n;
  const int m = n < cap ? n : cap;
  if (m > 0)
    POP_CUDA(h, cudaMemcpy(out, d->cv + static_cast<size_t>(agent) * d->G.MAXPTS * 7, sizeof(float) * 7 * m, cudaMemcpyDeviceToHost));
  return n > cap ? POP_E_CAPACITY : POP_OK;
}

int pop_dsp_stats(pop_handle* h, int agent, int32_t* out) {
  PopDeviceGuard guard(h);
  int rc = check(h, agent);
  if (rc) return rc;
  DspDev* d = h->dsp;
  int st[8], sq[4];
  POP_CUDA(h, cudaMemcpyAsync(st, d->stats + 8 * agent, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(sq, d->seq + 4 * agent, sizeof(sq), cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  out[0] = st[0]; out[1] = st[1]; out[2] = st[2]; out[3] = st[3];
  out[4] = sq[0]; out[5] = sq[1]; out[6] = sq[2]; out[7] = d->updates;
  return POP_OK;
}

int pop_dsp_occupied(pop_handle* h, int agent, float* out, int cap, int32_t* n_out) {
  PopDeviceGuard guard(h);
  int rc = check(h, agent);
  if (rc) return rc;
  if (cap < 0 || !n_out) return POP_E_INVALID;
  rc = pop_ensure_scratch(h, sizeof(float) * 3 * static_cast<size_t>(cap) + 256, 0);
  if (rc) return rc;
  float* d_out = static_cast<float*>(h->d_scratch);
  int* d_cnt = reinterpret_cast<int*>(static_cast<char*>(h->d_scratch) + ((sizeof(float) * 3 * static_cast<size_t>(cap) + 127) & ~size_t(127)));
  rc = popk_dsp_occupied(h, agent, d_out, cap, d_cnt);
  if (rc) return rc;
  int n = 0;
  POP_CUDA(h, cudaMemcpyAsync(&n, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  *n_out = n;
  const int m = n < cap ? n : cap;
  if (m > 0) POP_CUDA(h, cudaMemcpy(out, d_out, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost));
  return n > cap ? pop_fail(h, POP_E_CAPACITY, "%d occupied voxels exceed the buffer of %d", n, cap) : POP_OK;
}

}  // extern "C"

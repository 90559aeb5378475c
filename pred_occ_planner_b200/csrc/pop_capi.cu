// pop_capi.cu — the C ABI of libpop.so (include/pop.h): handle lifetime, input staging and
// the map-side entry points.  The planning stages live in k_astar.cu / k_corridor.cu / k_qp.cu.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "pop_geom.cuh"
#include "pop_internal.h"

int pop_fail(pop_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

static thread_local std::string g_create_err;

// The planning stages run as per-group chains on up to 32 side streams (k_cycle.cu); ask the driver for
// that many hardware work queues unless the host application already chose a value.  Takes effect
// only if libpop.so is loaded before the CUDA context is created.
namespace {
struct ConnEnv {
  ConnEnv() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
} g_conn_env;
}  // namespace

int pop_ensure_scratch(pop_handle* h, size_t dev_bytes, size_t host_bytes) {
  if (dev_bytes > h->d_scratch_bytes) {
    if (h->d_scratch) cudaFree(h->d_scratch);
    h->d_scratch = nullptr;
    size_t n = dev_bytes + (dev_bytes >> 2) + 4096;
    POP_CUDA(h, cudaMalloc(&h->d_scratch, n));
    h->d_scratch_bytes = n;
  }
  if (host_bytes > h->h_pinned_bytes) {
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    h->h_pinned = nullptr;
    size_t n = host_bytes + (host_bytes >> 2) + 4096;
    POP_CUDA(h, cudaMallocHost(&h->h_pinned, n));
    h->h_pinned_bytes = n;
  }
  return POP_OK;
}

// ---- per-stage CUDA-event profiling ----------------------------------------------------------
static cudaEvent_t prof_event(pop_handle* h) {
  if (!h->prof.pool.empty()) {
    cudaEvent_t e = h->prof.pool.back();
    h->prof.pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
static thread_local cudaEvent_t g_open_begin[POP_STAGE_COUNT];

void pop_prof_begin(pop_handle* h, int stage) {
  if (!h->prof.enabled || (h->prof.map_only && stage > POP_STAGE_PEERS)) return;
  cudaEvent_t e = prof_event(h);
  cudaEventRecord(e, h->stream);
  g_open_begin[stage] = e;
}
void pop_prof_end(pop_handle* h, int stage, int n_launches) {
  h->n_launches += n_launches;
  h->prof.launches[stage] += n_launches;
  if (!h->prof.enabled || (h->prof.map_only && stage > POP_STAGE_PEERS)) return;
  cudaEvent_t e = prof_event(h);
  cudaEventRecord(e, h->stream);
  h->prof.pending.push_back({stage, {g_open_begin[stage], e}});
}

static void derive_geometry(pop_handle* h) {
  const pop_config& c = h->cfg;
  PopGeom& g = h->geom;
  g.nx = c.nx; g.ny = c.ny; g.nz = c.nz; g.nt = c.nt;
  g.V = c.nx * c.ny * c.nz;
  g.res = c.resolution;
  g.dt = c.time_resolution;
  // local_update_range = N / 2 * resolution_ (integer half, fp32 multiply):
  // plan_env/src/fake_particle_risk_voxel.cpp:28-30
  g.rx = static_cast<float>(c.nx / 2) * g.res;
  g.ry = static_cast<float>(c.ny / 2) * g.res;
  g.rz = static_cast<float>(c.nz / 2) * g.res;
  g.risk_threshold = c.risk_threshold;
  g.clearance = c.clearance;
  g.ground = c.ground_height;
  g.ceiling = c.ceiling_height;
  g.thr_region = c.risk_threshold_region;
  g.decay_region = c.risk_region_decay;
  g.decay_voxel = c.risk_voxel_decay;
  g.inf_step = static_cast<int>(c.clearance / c.resolution);  // fp32 divide, :37
  g.kernel_shape = c.kernel_shape;
  g.query_variant = c.query_variant;
  h->cell_bytes = c.cell_dtype == POP_CELL_U8 ? 1 : 4;
  unsigned long long bytes = static_cast<unsigned long long>(g.V) * g.nt * h->cell_bytes;
  g.agent_stride = (bytes + 255ull) & ~255ull;
}

extern "C" {

int pop_abi_version(void) { return POP_ABI_VERSION; }

void pop_default_config(pop_config* c) {
  memset(c, 0, sizeof(*c));
  c->abi_version = POP_ABI_VERSION;
  // plan_env/include/plan_env/map_parameters.h:5-13 and fake_particle_risk_voxel.h:28-29
  c->nx = 66; c->ny = 66; c->nz = 20; c->nt = 9;
  c->resolution = 0.15f;
  // plan_manager/config/sim_fake.yaml (map: section) with map.cpp:15-40 defaults
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
  // search: section of sim_fake.yaml
  c->max_tau = 2.0; c->max_vel = 2.0; c->max_acc = 6.0; c->w_time = 5.0; c->horizon = 5.0;
  c->lambda_heu = 5.0; c->resolution_astar = 0.15; c->time_resolution_astar = 0.3;
  c->allocate_num = 10000; c->check_num = 1; c->tolerance = 1;
  // planner / corridor / optimizer sections
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

int pop_create(const pop_config* cfg, int device, void* stream, pop_handle** out) {
  if (!cfg || !out) return POP_E_INVALID;
  *out = nullptr;
  if (cfg->abi_version != POP_ABI_VERSION) return POP_E_INVALID;
  if (cfg->nx <= 0 || cfg->ny <= 0 || cfg->nz <= 0 || cfg->nt <= 0 || cfg->n_agents_local <= 0 ||
      cfg->n_agents_total < cfg->n_agents_local || cfg->resolution <= 0.f ||
      cfg->time_resolution <= 0.f || cfg->max_cloud_points <= 0 || cfg->max_cylinders <= 0)
    return POP_E_INVALID;
  if (cfg->cell_dtype != POP_CELL_U8 && cfg->cell_dtype != POP_CELL_F32) return POP_E_INVALID;
  if (cfg->replan_risk_rate < 0.f || cfg->num_resample < 1 || cfg->num_resample > 16) return POP_E_INVALID;
  // resampled peer particles carry fractional risks (particles.cpp:395-402): they need float cells
  if (cfg->replan_risk_rate != 0.f && cfg->cell_dtype != POP_CELL_F32) return POP_E_UNSUPPORTED;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return POP_E_NO_DEVICE;  // no CPU fallback by design
  }
  pop_handle* h = new pop_handle();
  h->cfg = *cfg;
  h->device = device;
  struct Restore {   // the caller's current device is the caller's business: put it back on every exit
    int prev = -1;
    Restore() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~Restore() { if (prev >= 0) cudaSetDevice(prev); }
  } restore;
  auto fail = [&](int code) { pop_destroy(h); return code; };
  if (cudaSetDevice(device) != cudaSuccess) return fail(POP_E_NO_DEVICE);
  if (stream) {
    h->stream = static_cast<cudaStream_t>(stream);
  } else {
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(POP_E_CUDA);
    h->own_stream = true;
  }
  derive_geometry(h);
  const int nl = cfg->n_agents_local, nt = cfg->n_agents_total;
  bool ok = true;
  ok &= cudaMalloc(&h->d_grid, h->geom.agent_stride * nl + 256) == cudaSuccess;
  ok &= cudaMalloc(&h->d_cloud, sizeof(float) * 3 * cfg->max_cloud_points) == cudaSuccess;
  ok &= cudaMalloc(&h->d_cyl, sizeof(PopCylDev) * cfg->max_cylinders) == cudaSuccess;
  ok &= cudaMalloc(&h->d_poses, sizeof(float) * 3 * nl) == cudaSuccess;
  ok &= cudaMalloc(&h->d_occ_list, sizeof(int) * static_cast<size_t>(nl) * cfg->max_cloud_points) == cudaSuccess;
  ok &= cudaMalloc(&h->d_occ_count, sizeof(int) * nl) == cudaSuccess;
  ok &= cudaMalloc(&h->d_body, sizeof(double) * 3 * POP_MAX_BODY * nt) == cudaSuccess;
  ok &= cudaMalloc(&h->d_body_count, sizeof(int) * nt) == cudaSuccess;
  ok &= cudaMalloc(&h->d_traj_all, sizeof(pop_traj_record) * nt) == cudaSuccess;
  ok &= cudaMalloc(&h->d_traj_send, sizeof(pop_traj_record) * nl) == cudaSuccess;
  ok &= cudaMalloc(&h->d_peer_pos, sizeof(double) * 3 * nt * cfg->nt) == cudaSuccess;
  ok &= cudaMalloc(&h->d_peer_valid, static_cast<size_t>(nt) * cfg->nt) == cudaSuccess;
  ok &= cudaMalloc(&h->d_peer_sigma, sizeof(float) * nt * cfg->nt) == cudaSuccess;
  if (!ok) { cudaGetLastError(); return fail(POP_E_CUDA); }
  cudaMemsetAsync(h->d_grid, 0, h->geom.agent_stride * nl + 256, h->stream);
  cudaMemsetAsync(h->d_poses, 0, sizeof(float) * 3 * nl, h->stream);
  cudaMemsetAsync(h->d_traj_all, 0, sizeof(pop_traj_record) * nt, h->stream);
  cudaMemsetAsync(h->d_traj_send, 0, sizeof(pop_traj_record) * nl, h->stream);
  cudaMemsetAsync(h->d_body_count, 0, sizeof(int) * nt, h->stream);
  cudaMemsetAsync(h->d_peer_sigma, 0, sizeof(float) * nt * cfg->nt, h->stream);
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return fail(POP_E_CUDA);
  if (cfg->replan_risk_rate != 0.f) {
    // The reference seeds a std::default_random_engine with time(NULL) inside every getParticlesWithRisk call and draws
    // nx, ny, nz from three std::normal_distribution<float> objects over it (particles.cpp:382-392).  Default table: that
    // very sequence for the clock value 1, one triple per (body particle, copy); pop_set_resample_noise replaces it.
    const int n_tr = POP_MAX_BODY * cfg->num_resample;
    std::vector<float> z(3 * static_cast<size_t>(n_tr));
    std::default_random_engine gen(1);
    std::normal_distribution<float> nx(0.0, 1.0f), ny(0.0, 1.0f), nz(0.0, 1.0f);
    for (int i = 0; i < n_tr; i++) { z[3 * i] = nx(gen); z[3 * i + 1] = ny(gen); z[3 * i + 2] = nz(gen); }
    if (pop_set_resample_noise(h, z.data(), n_tr) != POP_OK) return fail(POP_E_CUDA);
  }
  *out = h;
  return POP_OK;
}

int pop_destroy(pop_handle* h) {
  if (!h) return POP_OK;
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  pop_comm_free(h);
  popk_vox_free(h);
  popk_dsp_free(h);
  for (cudaStream_t s : h->side) cudaStreamDestroy(s);
  for (cudaEvent_t e : h->side_done) cudaEventDestroy(e);
  if (h->fork_ev) cudaEventDestroy(h->fork_ev);
  h->side.clear(); h->side_done.clear(); h->fork_ev = nullptr;
  void* ptrs[] = {h->d_grid, h->d_cloud, h->d_cyl, h->d_poses, h->d_occ_list, h->d_occ_count, h->d_body,
                  h->d_body_count, h->d_traj_all, h->d_traj_send, h->d_peer_pos, h->d_peer_valid, h->d_peer_sigma, h->d_noise,
                  h->d_scratch};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  pop_plan_free(h);
  for (auto& p : h->prof.pending) { cudaEventDestroy(p.second.first); cudaEventDestroy(p.second.second); }
  for (cudaEvent_t e : h->prof.pool) cudaEventDestroy(e);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  if (prev >= 0) cudaSetDevice(prev);
  return POP_OK;
}

const char* pop_last_error(const pop_handle* h) { return h ? h->err.c_str() : "null handle"; }

int pop_sync(pop_handle* h) {
  if (!h) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

// ---- inputs ----------------------------------------------------------------------------------
int pop_set_world_cloud(pop_handle* h, const float* pts, int n) {
  if (!h || n < 0 || (n > 0 && !pts)) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (n > h->cfg.max_cloud_points) return pop_fail(h, POP_E_CAPACITY, "cloud of %d points exceeds max_cloud_points=%d", n, h->cfg.max_cloud_points);
  if (n > 0) POP_CUDA(h, cudaMemcpyAsync(h->d_cloud, pts, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n_cloud = n;
  return POP_OK;
}

int pop_set_cylinders(pop_handle* h, const pop_cylinder* cyl, int n) {
  if (!h || n < 0 || (n > 0 && !cyl)) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (n > h->cfg.max_cylinders) return pop_fail(h, POP_E_CAPACITY, "%d cylinders exceed max_cylinders=%d", n, h->cfg.max_cylinders);
  std::vector<PopCylDev> v(n);
  float max_vy = 0.f;
  for (int i = 0; i < n; i++) {
    geom_cyl_prepare(cyl[i], &v[i]);
    if ((v[i].type == 3 || v[i].type == 2) && std::fabs(v[i].vy) > max_vy) max_vy = std::fabs(v[i].vy);
  }
  h->max_abs_vy = max_vy;
  if (n > 0) POP_CUDA(h, cudaMemcpyAsync(h->d_cyl, v.data(), sizeof(PopCylDev) * n, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n_cyl = n;
  return POP_OK;
}

// device-resident world inputs: the same two setters with device pointers, enqueued on the handle's stream (async)
static __global__ void k_cyl_prepare(const pop_cylinder* __restrict__ in, int n, PopCylDev* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) geom_cyl_prepare(in[i], &out[i]);
}
int pop_set_world_cloud_dev(pop_handle* h, const float* pts_dev, int n) {
  if (!h || n < 0 || (n > 0 && !pts_dev)) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (n > h->cfg.max_cloud_points) return pop_fail(h, POP_E_CAPACITY, "cloud of %d points exceeds max_cloud_points=%d", n, h->cfg.max_cloud_points);
  if (n > 0) POP_CUDA(h, cudaMemcpyAsync(h->d_cloud, pts_dev, sizeof(float) * 3 * n, cudaMemcpyDeviceToDevice, h->stream));
  h->n_cloud = n;
  return POP_OK;
}
int pop_set_cylinders_dev(pop_handle* h, const pop_cylinder* cyl_dev, int n, double max_abs_vy) {
  if (!h || n < 0 || (n > 0 && !cyl_dev) || !(max_abs_vy >= 0.0)) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (n > h->cfg.max_cylinders) return pop_fail(h, POP_E_CAPACITY, "%d cylinders exceed max_cylinders=%d", n, h->cfg.max_cylinders);
  if (n > 0) {
    k_cyl_prepare<<<(n + 63) / 64, 64, 0, h->stream>>>(cyl_dev, n, h->d_cyl);
    POP_LAUNCH_CHECK(h);
    h->n_launches++;
  }
  h->max_abs_vy = static_cast<float>(max_abs_vy);
  h->n_cyl = n;
  return POP_OK;
}

int pop_set_poses(pop_handle* h, const float* poses) {
  if (!h || !poses) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  POP_CUDA(h, cudaMemcpyAsync(h->d_poses, poses, sizeof(float) * 3 * h->cfg.n_agents_local, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

int pop_set_body_particles(pop_handle* h, const double* offs, const int32_t* counts) {
  if (!h || !offs || !counts) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  const int nt = h->cfg.n_agents_total;
  for (int i = 0; i < nt; i++)
    if (counts[i] < 0 || counts[i] > POP_MAX_BODY) return pop_fail(h, POP_E_CAPACITY, "agent %d has %d body particles (max %d)", i, counts[i], POP_MAX_BODY);
  POP_CUDA(h, cudaMemcpyAsync(h->d_body, offs, sizeof(double) * 3 * POP_MAX_BODY * nt, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(h->d_body_count, counts, sizeof(int) * nt, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  h->have_body = true;
  return POP_OK;
}

int pop_set_trajectories(pop_handle* h, const pop_traj_record* recs) {
  if (!h || !recs) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  POP_CUDA(h, cudaMemcpyAsync(h->d_traj_all, recs, sizeof(pop_traj_record) * h->cfg.n_agents_total, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  h->have_traj = true;
  return POP_OK;
}

int pop_set_resample_noise(pop_handle* h, const float* z, int n_triples) {
  if (!h || !z || n_triples < 1) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->d_noise) { cudaFree(h->d_noise); h->d_noise = nullptr; h->n_noise = 0; }
  POP_CUDA(h, cudaMalloc(&h->d_noise, sizeof(float) * 3 * n_triples));
  POP_CUDA(h, cudaMemcpyAsync(h->d_noise, z, sizeof(float) * 3 * n_triples, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n_noise = n_triples;
  return POP_OK;
}

int pop_get_trajectories(pop_handle* h, pop_traj_record* recs) {
  if (!h || !recs) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  POP_CUDA(h, cudaMemcpyAsync(recs, h->d_traj_all, sizeof(pop_traj_record) * h->cfg.n_agents_total, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

int pop_traj_buffers(pop_handle* h, void** send_dev, void** recv_dev) {
  if (!h) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (send_dev) *send_dev = h->d_traj_send;
  if (recv_dev) *recv_dev = h->d_traj_all;
  h->have_traj = true;  // the caller fills recv_dev (allgather) before the next map update
  return POP_OK;
}

// ---- map ---------------------------------------------------------------------------------------
int pop_map_update_gt(pop_handle* h, double t_map) {
  if (!h) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  h->t_map = t_map;
  int rc = popk_voxel_update(h);
  if (rc) return rc;
  rc = popk_peer_project(h);
  if (rc) return rc;
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  h->map_ready = true;
  return POP_OK;
}

static int query_common(pop_handle* h, int agent, const double* pos, const int32_t* t_idx,
                        const double* dt, int n, int8_t* out) {
  if (!h || agent < 0 || agent >= h->cfg.n_agents_local || n < 0 || (n > 0 && (!pos || !out))) return POP_E_INVALID;
  if (!h->map_ready) return pop_fail(h, POP_E_STATE, "query before the first map update");
  if (n == 0) return POP_OK;
  const size_t bpos = sizeof(double) * 3 * n, baux = 8 * static_cast<size_t>(n), bout = n;
  int rc = pop_ensure_scratch(h, bpos + baux + bout + 64, 0);
  if (rc) return rc;
  uint8_t* d = static_cast<uint8_t*>(h->d_scratch);
  double* d_pos = reinterpret_cast<double*>(d);
  void* d_aux = d + bpos;
  int8_t* d_out = reinterpret_cast<int8_t*>(d + bpos + baux);
  POP_CUDA(h, cudaMemcpyAsync(d_pos, pos, bpos, cudaMemcpyHostToDevice, h->stream));
  if (dt) POP_CUDA(h, cudaMemcpyAsync(d_aux, dt, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  else POP_CUDA(h, cudaMemcpyAsync(d_aux, t_idx, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
  rc = popk_query(h, agent, d_pos, dt ? nullptr : static_cast<int32_t*>(d_aux), dt ? static_cast<double*>(d_aux) : nullptr, n, d_out);
  if (rc) return rc;
  POP_CUDA(h, cudaMemcpyAsync(out, d_out, bout, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

int pop_map_query(pop_handle* h, int agent, const double* pos, const int32_t* t_idx, int n, int8_t* out) {
  PopDeviceGuard guard(h);
  if (n > 0 && !t_idx) return POP_E_INVALID;
  return query_common(h, agent, pos, t_idx, nullptr, n, out);
}
int pop_map_query_dt(pop_handle* h, int agent, const double* pos, const double* dt, int n, int8_t* out) {
  PopDeviceGuard guard(h);
  if (n > 0 && !dt) return POP_E_INVALID;
  return query_common(h, agent, pos, nullptr, dt, n, out);
}

int pop_map_query_inflate(pop_handle* h, int agent, const double* pos, const double* t, int mode, int n, int8_t* out) {
  if (!h || agent < 0 || agent >= h->cfg.n_agents_local || n < 0 || mode < 0 || mode > 2 ||
      (n > 0 && (!pos || !out || (mode && !t)))) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->map_ready) return pop_fail(h, POP_E_STATE, "query before the first map update");
  if (n == 0) return POP_OK;
  const size_t bpos = sizeof(double) * 3 * n, bt = sizeof(double) * n;
  int rc = pop_ensure_scratch(h, bpos + bt + n + 64, 0);
  if (rc) return rc;
  uint8_t* d = static_cast<uint8_t*>(h->d_scratch);
  double* d_pos = reinterpret_cast<double*>(d);
  double* d_t = reinterpret_cast<double*>(d + bpos);
  int8_t* d_out = reinterpret_cast<int8_t*>(d + bpos + bt);
  POP_CUDA(h, cudaMemcpyAsync(d_pos, pos, bpos, cudaMemcpyHostToDevice, h->stream));
  if (mode) POP_CUDA(h, cudaMemcpyAsync(d_t, t, bt, cudaMemcpyHostToDevice, h->stream));
  rc = popk_query_inflate(h, agent, d_pos, d_t, mode, n, d_out);
  if (rc) return rc;
  POP_CUDA(h, cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

int pop_map_obstacle_points(pop_handle* h, int agent, double t_start, double t_end, const double lc[3],
                            const double hc[3], double* out, int cap, int* n_out) {
  if (!h || agent < 0 || agent >= h->cfg.n_agents_local || !lc || !hc || cap < 0 || !n_out || (cap > 0 && !out)) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->map_ready) return pop_fail(h, POP_E_STATE, "obstacle points before the first map update");
  int rc = pop_ensure_scratch(h, sizeof(double) * 3 * cap + 64, 0);
  if (rc) return rc;
  double* d_out = static_cast<double*>(h->d_scratch);
  int* d_cnt = reinterpret_cast<int*>(static_cast<uint8_t*>(h->d_scratch) + sizeof(double) * 3 * cap);
  rc = popk_obstacle_points(h, agent, t_start, t_end, lc, hc, d_out, cap, d_cnt);
  if (rc) return rc;
  int cnt = 0;
  POP_CUDA(h, cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  const int ncopy = cnt < cap ? cnt : cap;
  if (ncopy > 0) POP_CUDA(h, cudaMemcpy(out, d_out, sizeof(double) * 3 * ncopy, cudaMemcpyDeviceToHost));
  *n_out = cnt;
  if (cnt > cap) return pop_fail(h, POP_E_CAPACITY, "%d obstacle points, buffer holds %d", cnt, cap);
  return POP_OK;
}

int pop_map_add_particles(pop_handle* h, int agent, const double* pts, const float* risks, int n, int t_index) {
  if (!h || agent < 0 || agent >= h->cfg.n_agents_local || n < 0 || t_index < 0 || t_index >= h->cfg.nt || (n > 0 && (!pts || !risks))) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (n == 0) return POP_OK;
  if (h->cfg.cell_dtype == POP_CELL_U8)
    for (int i = 0; i < n; i++)
      if (risks[i] != std::floor(risks[i]) || risks[i] < 0.f || risks[i] > 255.f)
        return pop_fail(h, POP_E_UNSUPPORTED, "u8 cells hold integer risks only (risk[%d]=%g); use POP_CELL_F32", i, risks[i]);
  const size_t bp = sizeof(double) * 3 * n;
  int rc = pop_ensure_scratch(h, bp + sizeof(float) * n, 0);
  if (rc) return rc;
  double* d_pts = static_cast<double*>(h->d_scratch);
  float* d_r = reinterpret_cast<float*>(static_cast<uint8_t*>(h->d_scratch) + bp);
  POP_CUDA(h, cudaMemcpyAsync(d_pts, pts, bp, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d_r, risks, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  rc = popk_add_particles(h, agent, d_pts, d_r, n, t_index);
  if (rc) return rc;
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

int pop_map_future_risk(pop_handle* h, const float* risk, const float* poses, double stamp) {
  if (!h || !risk) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (h->cfg.cell_dtype != POP_CELL_F32) return pop_fail(h, POP_E_UNSUPPORTED, "future-risk grids hold float risks: use POP_CELL_F32");
  const size_t per = sizeof(float) * static_cast<size_t>(h->geom.V) * h->geom.nt;
  int rc = pop_ensure_scratch(h, per, 0);
  if (rc) return rc;
  if (poses) {
    rc = pop_set_poses(h, poses);
    if (rc) return rc;
  }
  for (int a = 0; a < h->cfg.n_agents_local; a++) {
    POP_CUDA(h, cudaMemcpyAsync(h->d_scratch, reinterpret_cast<const uint8_t*>(risk) + per * a, per, cudaMemcpyHostToDevice, h->stream));
    rc = popk_upload(h, a, static_cast<const float*>(h->d_scratch));
    if (rc) return rc;
  }
  h->t_map = stamp;
  rc = popk_peer_project(h);
  if (rc) return rc;
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  h->map_ready = true;
  return POP_OK;
}

int pop_map_download(pop_handle* h, int agent, float* out) {
  if (!h || agent < 0 || agent >= h->cfg.n_agents_local || !out) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  const size_t bytes = sizeof(float) * static_cast<size_t>(h->geom.V) * h->geom.nt;
  int rc = pop_ensure_scratch(h, bytes, 0);
  if (rc) return rc;
  rc = popk_download(h, agent, static_cast<float*>(h->d_scratch));
  if (rc) return rc;
  POP_CUDA(h, cudaMemcpyAsync(out, h->d_scratch, bytes, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

// ---- measurement -------------------------------------------------------------------------------
int pop_profile_enable(pop_handle* h, int on) {
  if (!h) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  h->prof.enabled = on != 0;
  h->prof.map_only = on == 2;
  return POP_OK;
}

int pop_profile_read(pop_handle* h, float* ms, int64_t* launches, int reset) {
  if (!h) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  for (auto& p : h->prof.pending) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, p.second.first, p.second.second) == cudaSuccess) h->prof.ms[p.first] += t;
    h->prof.pool.push_back(p.second.first);
    h->prof.pool.push_back(p.second.second);
  }
  h->prof.pending.clear();
  for (int i = 0; i < POP_STAGE_COUNT; i++) {
    if (ms) ms[i] = h->prof.ms[i];
    if (launches) launches[i] = h->prof.launches[i];
  }
  if (reset)
    for (int i = 0; i < POP_STAGE_COUNT; i++) { h->prof.ms[i] = 0.f; h->prof.launches[i] = 0; }
  return POP_OK;
}

int64_t pop_kernel_launches(const pop_handle* h) { return h ? h->n_launches : 0; }

}  // extern "C"

// pop_internal.h — handle layout and shared device-side geometry of libpop.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pop.h"

// Grid geometry in the reference's fp32 arithmetic
// (plan_env/include/plan_env/map.h:153-215; init at fake_particle_risk_voxel.cpp:27-46).
struct PopGeom {
  int nx, ny, nz, nt;
  int V;               // nx*ny*nz
  float res, dt;       // resolution_, time_resolution_
  float rx, ry, rz;    // local_update_range_{x,y,z}_ = N/2 * res
  float risk_threshold, clearance, ground, ceiling;
  float thr_region, decay_region, decay_voxel;
  int inf_step;        // (int)(clearance_/resolution_) in fp32
  int kernel_shape, query_variant;
  unsigned long long agent_stride;  // bytes between consecutive agents' grids
};

struct PopCylDev {  // pop_cylinder reduced to what the voxel update reads
  float x, y;       // (float)cyl.x, (float)cyl.y
  float vx, vy;     // (float)cyl.vx, (float)cyl.vy
  double w;         // compared in double: dist <= cyl.w + clearance_ (type 3), |w / 2 - dist| < 2 res (type 2)
  int type;         // 3 vertical cylinder, 2 vertical ring (fake_particle_risk_voxel.cpp:131-150)
  float z;          // (float)cyl.z: ring centre
  float n[3], off;  // type 2: the ring's plane, Hyperplane<float, 3>::Through(centre, centre + q*(0,1,0), centre + q*(1,0,0))
};

struct PopProfile {
  bool enabled = false;
  bool map_only = false;   // time the map stages only (the planning stages then run pipelined over agent groups)
  float ms[POP_STAGE_COUNT] = {0};
  int64_t launches[POP_STAGE_COUNT] = {0};
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
  std::vector<cudaEvent_t> pool;
};

struct VoxDev;     // k_map.cu: per-layer entry lists of the tiled voxel update
struct AstarDev;   // k_astar.cu
struct CorridorDev;  // k_corridor.cu
struct QpDev;      // k_qp.cu
struct PlanDev;    // corridor + QP workspaces
struct DspDev;     // k_dsp.cu: DSP particle maps
struct CommDev;    // pop_comm.cu: the NCCL communicator of the trajectory exchange

struct pop_handle {
  pop_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  PopGeom geom;
  int cell_bytes = 1;
  bool map_ready = false;
  double t_map = 0.0;
  int64_t n_launches = 0;

  uint8_t* d_grid = nullptr;      // n_local * agent_stride bytes, layout [agent][t][z][y][x]
  float* d_cloud = nullptr;       // max_cloud_points x 3
  int n_cloud = 0;
  PopCylDev* d_cyl = nullptr;
  int n_cyl = 0;
  float max_abs_vy = 0.f;         // largest |vy| of the obstacles that can own voxels (tile reach of the voxel update)
  float* d_poses = nullptr;       // n_local x 3
  int* d_occ_list = nullptr;      // n_local x max_cloud_points voxel indices of slice 0
  int* d_occ_count = nullptr;     // n_local
  double* d_body = nullptr;       // n_total x POP_MAX_BODY x 3
  int* d_body_count = nullptr;    // n_total
  pop_traj_record* d_traj_all = nullptr;   // n_total (allgather receive buffer)
  pop_traj_record* d_traj_send = nullptr;  // n_local (allgather send buffer)
  double* d_peer_pos = nullptr;   // n_total x nt x 3 world positions of peers per slice
  int8_t* d_peer_valid = nullptr; // n_total x nt
  float* d_peer_sigma = nullptr;  // n_total x nt: pos_stddev of the peer resampling (replan_risk_rate > 0)
  float* d_noise = nullptr;       // n_noise standard-normal triples consumed by the resampling
  int n_noise = 0;
  bool have_body = false, have_traj = false;
  const int* d_agent_ok = nullptr; // optional per-agent mask of the peer projection (set around a particle-map update)

  // scratch for host<->device staging (pinned) and queries
  void* h_pinned = nullptr; size_t h_pinned_bytes = 0;
  void* d_scratch = nullptr; size_t d_scratch_bytes = 0;

  VoxDev* vox = nullptr;
  AstarDev* astar = nullptr;
  CorridorDev* corridor = nullptr;
  QpDev* qp = nullptr;
  PlanDev* plan = nullptr;
  DspDev* dsp = nullptr;
  CommDev* comm = nullptr;
  PopProfile prof;
  // agent-group pipelining of the planning stages (k_cycle.cu)
  std::vector<cudaStream_t> side;
  std::vector<cudaEvent_t> side_done;
  cudaEvent_t fork_ev = nullptr;
};

// error helpers -----------------------------------------------------------------------------
int pop_fail(pop_handle* h, int code, const char* fmt, ...);
#define POP_CUDA(h, expr)                                                                    \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return pop_fail((h), POP_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                      __FILE__, __LINE__);                                                   \
  } while (0)
#define POP_LAUNCH_CHECK(h) POP_CUDA(h, cudaGetLastError())

// Every entry point that touches the device runs with the handle's device current and restores the caller's on exit
// (a process may hold handles on several GPUs, or switch devices between calls).
struct PopDeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit PopDeviceGuard(const pop_handle* h) {
    if (h && cudaGetDevice(&prev) == cudaSuccess && prev != h->device) switched = cudaSetDevice(h->device) == cudaSuccess;
  }
  ~PopDeviceGuard() { if (switched) cudaSetDevice(prev); }
  PopDeviceGuard(const PopDeviceGuard&) = delete;
  PopDeviceGuard& operator=(const PopDeviceGuard&) = delete;
};

void pop_plan_free(pop_handle* h);
void pop_comm_free(pop_handle* h);
void popk_dsp_free(pop_handle* h);
int pop_ensure_scratch(pop_handle* h, size_t dev_bytes, size_t host_bytes);
void pop_prof_begin(pop_handle* h, int stage);
void pop_prof_end(pop_handle* h, int stage, int n_launches);

// map stage launchers (k_map.cu) ---------------------------------------------------------------
int popk_voxel_update(pop_handle* h);        // clear + slice 0 + future slices (K1f)
void popk_vox_free(pop_handle* h);
int popk_peer_project(pop_handle* h);        // peer trajectories → particles → grid (K5)
int popk_query(pop_handle* h, int agent, const double* d_pos, const int32_t* d_t,
               const double* d_dt, int n, int8_t* d_out);
int popk_query_inflate(pop_handle* h, int agent, const double* d_pos, const double* d_t, int mode, int n, int8_t* d_out);
int popk_obstacle_points(pop_handle* h, int agent, double t_start, double t_end, const double* lc,
                         const double* hc, double* d_out, int cap, int* d_count);
int popk_add_particles(pop_handle* h, int agent, const double* d_pts, const float* d_risks, int n,
                       int t_index);
int popk_download(pop_handle* h, int agent, float* d_out);
int popk_upload(pop_handle* h, int agent, const float* d_in);

// planning stage launchers -------------------------------------------------------------------------
int popk_astar_alloc(pop_handle* h);
void popk_astar_free(pop_handle* h);
int popk_astar_run(pop_handle* h, cudaStream_t st, int a0, int na);   // agents [a0, a0+na); inputs resident
double* popk_astar_path_dev(pop_handle* h);
int* popk_astar_len_dev(pop_handle* h);
int* popk_astar_ret_dev(pop_handle* h);
double* popk_astar_start_dev(pop_handle* h);
double* popk_astar_goal_dev(pop_handle* h);
double* popk_astar_tstart_dev(pop_handle* h);

int popk_corridor_alloc(pop_handle* h);
void popk_corridor_free(pop_handle* h);
int popk_corridor_run(pop_handle* h, const double* d_path, const int* d_len, const double* d_start, cudaStream_t st,
                      int a0, int na);
int popk_corridor_check_err(pop_handle* h);
double* popk_corridor_hpoly_dev(pop_handle* h);
int* popk_corridor_rows_dev(pop_handle* h);
int* popk_corridor_npoly_dev(pop_handle* h);
double* popk_corridor_goal_dev(pop_handle* h);
double* popk_corridor_tst_dev(pop_handle* h);

int popk_qp_alloc(pop_handle* h, int n_agents);
void popk_qp_free(pop_handle* h);
int popk_qp_run(pop_handle* h, int n_agents, const double* d_hpoly, const int* d_rows, double vmax, double amax,
                cudaStream_t st = nullptr, int a0 = 0);
int* popk_qp_npieces_dev(pop_handle* h);
double* popk_qp_init_dev(pop_handle* h);
double* popk_qp_goal_dev(pop_handle* h);
double* popk_qp_dur_dev(pop_handle* h);
double* popk_qp_cpts_dev(pop_handle* h);
int* popk_qp_status_dev(pop_handle* h);
int* popk_qp_iters_dev(pop_handle* h);

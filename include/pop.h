/*
 * pop.h — C ABI of libpop.so, the B200-native per-replan hot path of a
 * spatio-temporal-occupancy (SOGM) swarm planner.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no C++/torch
 * types.  Every entry point names the reference interface it replaces
 * (paths relative to the reference tree, file:line).  A handle owns N_local
 * agents' SOGM grids, A* pools and QP workspaces on ONE GPU; it stands where
 * N_local `FakeBaselinePlanner`/`BaselinePlanner` objects stood
 * (plan_manager/src/baseline_fake.cpp:18-51, plan_manager/src/baseline.cpp:15-43).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (POP_E_*);
 *     pop_last_error() gives the text.  Per-agent outcomes mirror the reference
 *     enums (ASTAR_RET, bool replan, int -1/0/1 occupancy).
 *   - all pointer arguments are HOST pointers unless the name ends in _dev.
 *     Calls enqueue on the handle's stream and return after the results are in
 *     the caller's buffers (they synchronise the stream) unless documented async.
 *   - there is NO CPU fallback: if no CUDA device is usable pop_create() fails
 *     with POP_E_NO_DEVICE.
 */
#ifndef POP_H_
#define POP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POP_ABI_VERSION 1

/* error codes */
enum {
  POP_OK            = 0,
  POP_E_INVALID     = -1, /* bad argument / config */
  POP_E_NO_DEVICE   = -2, /* no usable CUDA device (never falls back to CPU) */
  POP_E_CUDA        = -3, /* CUDA runtime error, see pop_last_error */
  POP_E_CAPACITY    = -4, /* a caller-provided buffer or an internal pool is too small */
  POP_E_STATE       = -5, /* call order violated (e.g. query before first map update) */
  POP_E_UNSUPPORTED = -6
};

/* cell storage of the SOGM */
enum { POP_CELL_F32 = 0, POP_CELL_U8 = 1 };
/* collision-kernel shape: FakeParticleRiskVoxel/MapBase build a single z layer
 * (plan_env/src/fake_particle_risk_voxel.cpp:37-46, map.cpp:54-62); RiskBase a cube
 * (plan_env/src/risk_base.cpp:31-39). */
enum { POP_KERNEL_DISC_Z0 = 0, POP_KERNEL_CUBE = 1 };
/* query variant: FAKE returns -1 outside [ground,ceiling] and compares the
 * region sum with risk_threshold (fake_particle_risk_voxel.cpp:309-330); RISKBASE returns 1
 * and compares with risk_threshold_region - t*decay (risk_base.cpp:228-250). */
enum { POP_QUERY_FAKE = 0, POP_QUERY_RISKBASE = 1 };

/* ASTAR_RET (path_searching/include/path_searching/dyn_a_star.h:15) */
enum {
  POP_ASTAR_NO_PATH = 0, POP_ASTAR_INIT_ERR = 1, POP_ASTAR_SEARCH_ERR = 2,
  POP_ASTAR_REACH_HORIZON = 3, POP_ASTAR_REACH_END = 4, POP_ASTAR_NEAR_END = 5
};

/* replan outcome per agent (bool replan() in the reference, plus the stage that failed) */
enum {
  POP_REPLAN_OK = 0,
  POP_REPLAN_FAIL_ASTAR = 1,       /* baseline_fake.cpp:286-289 */
  POP_REPLAN_FAIL_CORRIDOR = 2,    /* :375-391 */
  POP_REPLAN_FAIL_QP = 3,          /* :427-432 */
  POP_REPLAN_FAIL_SEPARATION = 4   /* :441-446 */
};

#define POP_MAX_PIECES 12   /* trajectory_piece_max_size, baseline_fake.h:67 */
#define POP_MAX_BODY   64   /* body particles per agent (27..36 in the reference configs) */
#define POP_MAX_PATH   16   /* samples returned by getPathWithVel */
#define POP_MAX_FACES  48   /* faces kept per corridor polytope */

/* One ground-truth obstacle, plan_env/include/plan_env/fake_particle_risk_voxel.h:31-44.
 * type 3 = vertical cylinder (visualization_msgs::Marker::CYLINDER, dynamic_forest_seq_sensing.cpp:199): x, y, w = diameter;
 * type 2 = vertical ring (Marker::SPHERE, :205-254; MovingCircle): centre x, y, z, w = 2 * inner radius, orientation q;
 * other types own no voxel (the reference prints "unknown type"). */
typedef struct pop_cylinder {
  int32_t type;
  int32_t _pad;
  double  x, y, z;
  double  w, h;
  double  vx, vy;
  double  qw, qx, qy, qz;
} pop_cylinder;

/* Fixed-size trajectory record, the unit of the per-cycle allgather.  Derived from
 * traj_utils/msg/BezierTraj.msg:1-9; `duration` is float32 on the wire (msg:8), so peers
 * evaluate with float-rounded piece durations. 1536 bytes. */
typedef struct pop_traj_record {
  int32_t drone_id;
  int32_t traj_id;
  int32_t n_pieces;      /* 0 = no trajectory published yet */
  int32_t valid;
  double  start_time;
  float   duration[POP_MAX_PIECES];
  double  cpts[POP_MAX_PIECES * 5 * 3];
  double  _pad[3];
} pop_traj_record;

/* All run-time parameters.  The reference spreads them over compile-time #defines
 * (plan_env/include/plan_env/map_parameters.h:5-24) and ROS params (map.cpp:15-40,
 * fake_risk_hybrid_a_star.cpp:62-82, baseline_fake.h:84-113, particles.cpp:28-34). */
typedef struct pop_config {
  int32_t abi_version;            /* = POP_ABI_VERSION */
  /* --- grid --- */
  int32_t nx, ny, nz;             /* MAP_LENGTH/WIDTH/HEIGHT_VOXEL_NUM */
  int32_t nt;                     /* PREDICTION_TIMES */
  float   resolution;             /* VOXEL_RESOLUTION */
  float   time_resolution;        /* map/time_resolution */
  float   risk_threshold;         /* map/risk_threshold_voxel */
  float   clearance;              /* map/clearance */
  float   ground_height;          /* map/ground_height */
  float   ceiling_height;         /* map/ceiling_height */
  float   risk_threshold_region;  /* map/risk_threshold_region       (RISKBASE) */
  float   risk_region_decay;      /* map/risk_threshold_region_decay (RISKBASE) */
  float   risk_voxel_decay;       /* map/risk_threshold_voxel_decay  (RISKBASE) */
  int32_t kernel_shape;           /* POP_KERNEL_* */
  int32_t query_variant;          /* POP_QUERY_* */
  int32_t cell_dtype;             /* POP_CELL_* */
  /* --- swarm --- */
  int32_t n_agents_local;         /* agents owned by this handle (this GPU) */
  int32_t n_agents_total;         /* swarm size over all ranks */
  int32_t agent_id_base;          /* global id of local agent 0 */
  double  drone_size[3];          /* swarm/drone_size_{x,y,z} */
  float   replan_risk_rate;       /* swarm/replan_risk_rate; > 0 resamples peer particles (pop_set_resample_noise), F32 cells */
  int32_t num_resample;           /* swarm/num_resample, 1..16 */
  /* --- hybrid A* (search/...) --- */
  double  max_tau, max_vel, max_acc, w_time, horizon, lambda_heu;
  double  resolution_astar, time_resolution_astar;
  int32_t allocate_num, check_num, tolerance;
  int32_t _pad0;
  /* --- corridor + optimiser (planner/..., corridor/..., optimizer/...) --- */
  double  corridor_tau, init_range, shrink_size;
  double  opt_max_vel, opt_max_acc;
  /* --- capacities --- */
  int32_t max_cloud_points;       /* world cloud capacity */
  int32_t max_cylinders;
  int32_t max_obstacle_points;    /* per corridor segment (reference reserves 2000) */
  /* --- QP (OSQP settings used by traj_opt/include/iosqp.hpp:87-89 + library defaults) --- */
  int32_t osqp_max_iter;          /* 4000 */
  int32_t osqp_adaptive_rho_interval; /* fixed interval replacing OSQP's wall-clock rule */
  int32_t osqp_check_termination; /* 25 */
  double  osqp_eps_abs, osqp_eps_rel;
} pop_config;

typedef struct pop_handle pop_handle;

/* ---- lifetime ------------------------------------------------------------------------ */
/* stream: a cudaStream_t cast to void* on which all work is enqueued (NULL → own stream).
 * Replaces FakeBaselinePlanner::init (baseline_fake.cpp:18-51). */
int         pop_create(const pop_config* cfg, int device, void* stream, pop_handle** out);
int         pop_destroy(pop_handle* h);
const char* pop_last_error(const pop_handle* h);
int         pop_abi_version(void);
int         pop_sync(pop_handle* h);
/* fill *cfg with the reference's simulator defaults (plan_manager/config/sim_fake.yaml) */
void        pop_default_config(pop_config* cfg);

/* ---- map inputs ------------------------------------------------------------------------ */
/* World point cloud shared by all local agents: what arrives on "map/cloud"
 * (fake_particle_risk_voxel.cpp:50, updateMap :80-104). pts = n x (x,y,z) float32. */
int pop_set_world_cloud(pop_handle* h, const float* pts, int n);
/* Ground-truth obstacle states: "map/state" (groundTruthStateCallback :244-264). */
int pop_set_cylinders(pop_handle* h, const pop_cylinder* cyl, int n);
/* The same two inputs from DEVICE memory (a simulator or sensor pipeline that already lives on the GPU): enqueued on
 * the handle's stream, asynchronous, nothing crosses PCIe.  cyl_dev = n pop_cylinder records in device memory;
 * max_abs_vy = the largest |vy| among them (the host-side bound the tiled voxel update sizes its reach with). */
int pop_set_world_cloud_dev(pop_handle* h, const float* pts_dev, int n);
int pop_set_cylinders_dev(pop_handle* h, const pop_cylinder* cyl_dev, int n, double max_abs_vy);
/* Map centre per local agent: MapBase::pose_ set by odom/pose callbacks (map.cpp:139-168).
 * poses = n_agents_local x 3 float32. */
int pop_set_poses(pop_handle* h, const float* poses);
/* Body particle offsets per GLOBAL agent: ParticleATC::particles_buffer_
 * (particles.cpp:62-87, 90-110). offs = n_agents_total x POP_MAX_BODY x 3 doubles,
 * counts = n_agents_total. */
int pop_set_body_particles(pop_handle* h, const double* offs, const int32_t* counts);
/* Peer trajectories as last received: ParticleATC::swarm_trajs_ (particles.cpp:131-191).
 * recs = n_agents_total records indexed by global drone id. */
int pop_set_trajectories(pop_handle* h, const pop_traj_record* recs);

/* Peer-uncertainty resampling (swarm/replan_risk_rate > 0; ParticleATC::getParticlesWithRisk, particles.cpp:365-408):
 * every projected peer particle becomes num_resample Gaussian copies with stddev replan_risk_rate * (t - t_start),
 * weighted by their density and normalised per particle.  The reference draws the offsets from a generator it seeds
 * with time(NULL) inside each call; here they are an input: z = n_triples x 3 standard-normal values (nx, ny, nz before
 * scaling by the stddev), consumed from the start for every (peer, slice), particle j taking triples j * num_resample ...
 * (wrapping).  pop_create installs the sequence libstdc++ yields for the clock value 1.  Needs POP_CELL_F32. */
int pop_set_resample_noise(pop_handle* h, const float* z, int n_triples);
/* The table pop_set_trajectories / the allgather filled, read back (n_agents_total records). */
int pop_get_trajectories(pop_handle* h, pop_traj_record* recs);

/* ---- K1f + K5: SOGM update ------------------------------------------------------------- */
/* FakeParticleRiskVoxel::updateMap for every local agent (fake_particle_risk_voxel.cpp:80-222):
 * clear, slice 0 from the cloud, future slices from the owning cylinder's velocity, then
 * peer-trajectory particles (addParticlesToRiskMap :291-301). t_map = last_update_time_. */
int pop_map_update_gt(pop_handle* h, double t_map);

/* RiskBase::updateTimeCallback + futureRiskCallback (plan_env/src/risk_base.cpp:60-80; risk_base.h:85) for every local
 * agent: the [VOXEL_NUM][PREDICTION_TIMES] float grid a risk_mapping_node published on "map/future_risk"
 * (risk_mapping_node.cpp:105-150) becomes the agent's SOGM, stamp becomes the map time, then the peers are projected
 * (addOtherAgents :136-168).  risk = n_agents_local x V x T floats; poses (n_local x 3, may be NULL = unchanged) are the
 * map centres.  Needs POP_CELL_F32.  (pop_dsp_update produces the same grid on the device without this copy.) */
int pop_map_future_risk(pop_handle* h, const float* risk, const float* poses, double stamp);

/* getClearOcccupancy(pos, int t) (fake_particle_risk_voxel.cpp:309-330 / risk_base.cpp:228-250)
 * pos = n x 3 doubles (world), t_idx = n slice indices, out = n results in {-1,0,1}. */
int pop_map_query(pop_handle* h, int agent, const double* pos, const int32_t* t_idx, int n,
                  int8_t* out);
/* getClearOcccupancy(pos, double dt) (:343-348 / risk_base.cpp:256-260). */
int pop_map_query_dt(pop_handle* h, int agent, const double* pos, const double* dt, int n,
                     int8_t* out);
/* MapBase::getInflateOccupancy (plan_env/include/plan_env/map.h:120-122, map.cpp:345-395): the single-voxel test of
 * the kept map interface.  mode 0 = (pos), 1 = (pos, int t) with t[i] the slice index, 2 = (pos, double t) testing
 * slices ceil(t / time_resolution) and floor(t / time_resolution).  t = n doubles (ignored for mode 0);
 * out = n results in {-1 out of range, 0, 1}.  A slice index equal to PREDICTION_TIMES (which the reference reads one
 * slice past the end) answers -1 in mode 1 and is clamped to nt - 1 in mode 2. */
int pop_map_query_inflate(pop_handle* h, int agent, const double* pos, const double* t, int mode, int n,
                          int8_t* out);
/* MapBase::getObstaclePoints(points, t_start, t_end, lc, hc) (map.cpp:463-531; RiskBase
 * variant risk_base.cpp:295-337). out = cap x 3 doubles in the reference's emission order
 * (z, y, x, slice); *n_out = number the reference would have emitted (may exceed cap →
 * POP_E_CAPACITY with the first cap points written). */
int pop_map_obstacle_points(pop_handle* h, int agent, double t_start, double t_end,
                            const double lc[3], const double hc[3], double* out, int cap,
                            int* n_out);
/* addParticlesToRiskMap(pts, risks, t_index) (fake_particle_risk_voxel.cpp:291-301):
 * pts are already map-frame (pose subtracted) doubles. */
int pop_map_add_particles(pop_handle* h, int agent, const double* pts, const float* risks, int n,
                          int t_index);
/* Parity/debug: the grid in the reference's [VOXEL_NUM][PREDICTION_TIMES] float layout
 * (plan_env/include/plan_env/map.h:52). out = nx*ny*nz*nt floats. */
int pop_map_download(pop_handle* h, int agent, float* out);

/* ---- K2: kinodynamic hybrid A* -------------------------------------------------------- */
/* FakeRiskHybridAstar::reset + search(+retry with init=false) + getPathWithVel(corridor_tau)
 * for n = n_agents_local agents (fake_risk_hybrid_a_star.cpp:85-107,114-421,663-694;
 * call site baseline_fake.cpp:268-296).  With query_variant = POP_QUERY_RISKBASE it is RiskHybridAstar
 * (risk_hybrid_a_star.cpp, call site baseline.cpp:268-296): the same search, except that the one-shot
 * trajectory is checked against slice 0 (:514) instead of its own time (fake_...:521).
 *   start   n x 9  (pos, vel, acc)        goal n x 3
 *   t_start n      (traj_start_time - map time)
 *   out_ret n      ASTAR_RET of the search whose path is returned
 *   out_path n x POP_MAX_PATH x 6 (pos, vel per sample), out_len n samples
 *   out_stats n x 4 (optional, may be NULL): iterations, nodes used, retried, reserved */
int pop_astar_search_batch(pop_handle* h, const double* start, const double* goal,
                           const double* t_start, int32_t* out_ret, double* out_path,
                           int32_t* out_len, int32_t* out_stats);

/* ---- K3: safe flight corridors --------------------------------------------------------- */
/* Corridor stage of replan (baseline_fake.cpp:296-411): per path segment obstacle points →
 * firi::firi (plan_manager/include/sfc_gen/firi.hpp:238-365) → ShrinkCorridor → validity,
 * adjacency and goal-reachability LPs (traj_utils/include/traj_utils/sdlp.hpp:708-785).
 *   path/len as returned by pop_astar_search_batch, start n x 9, traj_start_time n (absolute)
 *   out_hpoly n x POP_MAX_PIECES x POP_MAX_FACES x 4, out_rows n x POP_MAX_PIECES,
 *   out_npoly n (0 = corridor stage failed), out_goal n x 6 (local goal pos, vel) */
int pop_corridor_batch(pop_handle* h, const double* path, const int32_t* len, const double* start,
                       const double* traj_start_time, double* out_hpoly, int32_t* out_rows,
                       int32_t* out_npoly, double* out_goal);

/* ---- K4: corridor-constrained Bézier QP ------------------------------------------------ */
/* BezierOpt::setup + optimize + getOptCtrlPtsMat for a batch (traj_opt/src/bezier_optimizer.cpp:
 * 27-285 → traj_opt/include/iosqp.hpp:40-115 → OSQP).
 *   n_pieces n; init/goal n x 9 (rows pos;vel;acc); dur n x POP_MAX_PIECES;
 *   hpoly n x POP_MAX_PIECES x POP_MAX_FACES x 4 with rows n x POP_MAX_PIECES;
 *   out_cpts n x POP_MAX_PIECES x 5 x 3; out_status n OSQP status_val (1 solved,
 *   2 solved inaccurate, -2 max iter, -3 primal infeasible, -4 dual infeasible);
 *   out_iters n ADMM iterations (may be NULL). */
int pop_qp_solve_batch(pop_handle* h, int n, const int32_t* n_pieces, const double* init,
                       const double* goal, const double* dur, const double* hpoly,
                       const int32_t* rows, double vmax, double amax, double* out_cpts,
                       int32_t* out_status, int32_t* out_iters);

/* ---- K6: post-optimisation separation check ------------------------------------------- */
/* ParticleATC::isSafeAfterOpt (particles.cpp:223-283) → separator::Separator::solveModel
 * (utils/separator/src/separator_glpk.cpp:73-246): for every local agent, its new control
 * points against every active peer's remaining control points; out_safe n (1 safe). */
int pop_separation_check_batch(pop_handle* h, const int32_t* n_pieces, const double* cpts,
                               double t_now, int8_t* out_safe);

/* ---- whole cycle ------------------------------------------------------------------------ */
typedef struct pop_cycle_io {
  /* inputs (host): */
  const double* start;      /* n_local x 9 */
  const double* goal;       /* n_local x 3 */
  const float*  poses;      /* n_local x 3 map centres */
  double        t_map;      /* map stamp of this cycle */
  double        t_plan;     /* traj_start_time of this cycle */
  /* outputs (host): */
  int32_t*         status;  /* n_local POP_REPLAN_* */
  pop_traj_record* traj;    /* n_local records (valid=1 where status==OK) */
} pop_cycle_io;

/* One replan cycle for all local agents = FakeBaselinePlanner::replan x N
 * (baseline_fake.cpp:253-453) preceded by the map update it reads.  Host inputs are copied
 * in, results copied out (this is the call `e2e` in bench.py times). */
int pop_replan_cycle(pop_handle* h, const pop_cycle_io* io);
/* The planning half of the cycle on the SOGM as the last map update left it — BaselinePlanner::replan x N
 * (plan_manager/src/baseline.cpp:253-453) reading the map the particle-map node produced
 * (pop_dsp_update) or a ground-truth update.  io->poses and io->t_map are ignored. */
int pop_plan_cycle(pop_handle* h, const pop_cycle_io* io);
/* BaselinePlanner::isTrajSafe(T) x N (plan_manager/src/baseline.cpp:45-68 = baseline_fake.cpp:53-76), the FSM's trigger
 * for an early replan (fake_plan_manager.cpp:144): every local agent's own trajectory (the last one a cycle published for
 * it; pieces of corridor_tau seconds) is sampled every 0.1 s from wall-clock time t_now up to min(T, its duration) and
 * tested with getClearOcccupancy(pos, dt) on the agent's current SOGM.  out_safe n_local (1 safe, 0 = a sample was
 * occupied); out_t_unsafe (may be NULL) n_local: time of that sample relative to the trajectory start, else -1.
 * T = fsm/colli_check_duration (2.0 in the reference's configs). */
int pop_traj_safe_batch(pop_handle* h, double t_now, double T, int8_t* out_safe, double* out_t_unsafe);
/* Same cycle with inputs already resident on the device (uploaded by the previous
 * pop_replan_cycle/pop_set_* calls); nothing is copied to the host. Async. */
int pop_replan_cycle_resident(pop_handle* h, double t_map, double t_plan);
/* Results of the last enqueued cycle (synchronises): status n_local POP_REPLAN_*, traj n_local
 * records (either may be NULL). */
int pop_cycle_results(pop_handle* h, int32_t* status, pop_traj_record* traj);
/* Single-rank stand-in for the allgather: copies this handle's freshly published records into
 * its own receive buffer at [agent_id_base, agent_id_base + n_agents_local). Async. */
int pop_traj_publish_local(pop_handle* h);
/* ---- multi-GPU: the trajectory exchange, owned by the handle ------------------------------------------------------ */
/* One process (or host thread) per GPU, each with one handle holding an equal block of the swarm in rank order
 * (agent_id_base = rank * n_agents_local).  "/broadcast_traj" (plan_manager/src/fake_plan_manager.cpp:364-399 ->
 * traj_coordinator/src/particles.cpp:131-191) becomes ONE ncclAllGather of n_agents_local 1536-byte records per cycle:
 *   pop_comm_unique_id   rank 0 creates the 128-byte NCCL id (ncclGetUniqueId) and hands it to every rank by the
 *                        host's own means (MPI, a file, torch.distributed, ...);
 *   pop_comm_init        ncclCommInitRank; the handle owns the communicator until pop_destroy.  world = 1 needs no id;
 *   pop_exchange_trajs   after a cycle: allgather of the records the cycle published (send buffer) into the table the
 *                        next map update projects (receive buffer), on the handle's stream, asynchronous.  Without a
 *                        communicator (single GPU) it is the device-to-device publish of pop_traj_publish_local.
 *   pop_exchange_trajs_pipelined   the same allgather, one call behind: this rank's own records enter its table at
 *                        once, the other ranks' records are those of the PREVIOUS call, whose allgather ran on a
 *                        stream of its own while this rank planned the cycle in between.  A rank then waits for the
 *                        slowest rank of the cycle before last instead of the last one — a broadcast in flight for one
 *                        replan period, as "/broadcast_traj" messages are in the reference (the subscriber's callback
 *                        runs whenever the message arrives, fake_plan_manager.cpp:364-399; records carry their own
 *                        start time and are evaluated at absolute time, particles.cpp:131-191).  Every rank must use
 *                        the same variant for every call.  Falls back to pop_exchange_trajs without a communicator.
 * NCCL is loaded at run time (libnccl.so.2); POP_E_UNSUPPORTED if it is not there. */
#define POP_COMM_ID_BYTES 128
int pop_comm_unique_id(void* id_out /* POP_COMM_ID_BYTES */);
int pop_comm_init(pop_handle* h, const void* nccl_unique_id, int rank, int world);
int pop_comm_info(const pop_handle* h, int* rank, int* world, int* nccl_version);
int pop_exchange_trajs(pop_handle* h);
int pop_exchange_trajs_pipelined(pop_handle* h);
/* Device buffers for the per-cycle trajectory allgather ("/broadcast_traj",
 * plan_manager/src/fake_plan_manager.cpp:364-399 → particles.cpp:131-191):
 * *send_dev = n_agents_local records produced by the last cycle,
 * *recv_dev = n_agents_total records consumed by the next map update. */
int pop_traj_buffers(pop_handle* h, void** send_dev, void** recv_dev);

/* ---- K1d: DSP particle map (real-perception SOGM producer) -------------------------------- */
/* Parameters of dsp_map::DSPMap.  The reference fixes most of them at compile time
 * (plan_env/include/plan_env/map_parameters.h:5-53) or in the node that owns the map
 * (plan_env/src/risk_mapping_node.cpp:232-241).  Grid size, resolution and the number of future
 * slices come from the pop_config of the handle the map is attached to (MAP_*_VOXEL_NUM,
 * VOXEL_RESOLUTION, PREDICTION_TIMES). */
#define POP_DSP_MAX_T 40
typedef struct pop_dsp_config {
  int32_t half_fov_h, half_fov_v;       /* 43, 29 degrees (map_parameters.h:22-24) */
  int32_t angle_resolution;             /* ANGLE_RESOLUTION 1 degree */
  int32_t max_particle_num_voxel;       /* MAX_PARTICLE_NUM_VOXEL 7; storage is 2x (SAFE_PARTICLE_NUM_VOXEL), <= 8 */
  int32_t pyramid_capacity;             /* SAFE_PARTICLE_NUM_PYRAMID = (V*7 + 1e5) / (360*180) * 2; 0 = derive */
  int32_t max_points_per_pyramid;       /* observation_max_points_num_one_pyramid 100 */
  int32_t max_points;                   /* MAX_POINT_NUM 5000 valid points after the voxel filter */
  int32_t max_input_points;             /* capacity of one agent's raw camera cloud */
  int32_t new_born_particle_number_each_point; /* 20 (risk_mapping_node.cpp:236) */
  int32_t dynamic_cluster_max_point_num;       /* DYNAMIC_CLUSTER_MAX_POINT_NUM 200 */
  float   dynamic_cluster_max_center_height;   /* DYNAMIC_CLUSTER_MAX_CENTER_HEIGHT 1.5 */
  float   new_born_particle_weight;     /* 0.0001 (risk_mapping_node.cpp:238) */
  float   sigma_observation;            /* map/sigma_observation */
  float   kappa;                        /* 0.01 */
  float   p_detection;                  /* 0.95 */
  float   voxel_filter_resolution;      /* map/resolution of the voxel filter, 0.15 */
  float   obstacle_thickness_for_occlusion; /* 0.3 */
  float   occupancy_threshold;          /* map/risk_threshold_voxel (occupied-voxel export) */
  float   position_prediction_stddev;   /* setPredictionVariance(0.01, 0.01): table generator only */
  float   velocity_prediction_stddev;
  float   prediction_future_time[POP_DSP_MAX_T]; /* nt entries, seconds (map_parameters.h:18) */
  int32_t n_randoms;                    /* GAUSSIAN_RANDOMS_NUM (1e7 in the reference) */
  int32_t _pad;
  uint64_t seed;                        /* table generator seed (the reference uses time(NULL)) */
} pop_dsp_config;

/* reference defaults for the grid of `cfg` (pyramid_capacity derived) */
void pop_dsp_default_config(const pop_config* cfg, pop_dsp_config* out);
/* One DSPMap per local agent, standing where risk_mapping_node's `dsp_map_` stood
 * (risk_mapping_node.cpp:38).  Requires cell_dtype F32.  Tables are generated from `seed`. */
int pop_dsp_attach(pop_handle* h, const pop_dsp_config* cfg);
/* Replace the noise tables: p_gaussian_randoms, v_gaussian_randoms (dsp_dynamic.h:114-116) and the
 * stream rand() would have produced (generateRandomFloat :1687-1689); n entries each, shared by all
 * local agents, every agent's sequence counters reset to 0. */
int pop_dsp_set_random_tables(pop_handle* h, const float* p_gauss, const float* v_gauss,
                              const int32_t* rand_stream, int n);
/* cloudCallback for every local agent (risk_mapping_node.cpp:155-188): filterPointCloud (voxel-grid
 * filter, camera->body axes, range, cap), DSPMap::update (dsp_dynamic.h:165-364), publishMap ->
 * RiskBase::futureRiskCallback (risk_base.cpp:60-80: the [V][T] future risk becomes the agent's SOGM,
 * then peers are projected).  cam = n_local x max_input_points x 3 camera-frame points, counts n_local,
 * poses n_local x 3, quats n_local x 4 (w,x,y,z), one stamp.  out_ok (may be NULL) n_local: 1 updated,
 * 0 rejected (invalid odometry, dsp_dynamic.h:183-204). */
int pop_dsp_update(pop_handle* h, const float* cam, const int32_t* counts, const float* poses,
                   const float* quats, double stamp, int32_t* out_ok);
/* Same, with points already filtered into the body frame (valid_clouds_): pts = n_local x max_points x 3. */
int pop_dsp_update_filtered(pop_handle* h, const float* pts, const int32_t* counts, const float* poses,
                            const float* quats, double stamp, int32_t* out_ok);
/* filterPointCloud alone (risk_mapping_node.cpp:74-99) for one agent: out = max_points x 3. */
int pop_dsp_filter(pop_handle* h, int agent, const float* cam, int n, float* out, int32_t* n_out);
/* Parity/debug views in the reference's layouts:
 *   particles  voxels_with_particle[V][2*max_particle_num_voxel][8] (flag vx vy vz px py pz weight)
 *   voxels     voxels_objects_number[V][0..3] (occupancy weight, mean velocity)
 *   observations point_cloud[NP][max_points_per_pyramid][5], counts[NP], max_length[NP]
 *   pyramids   pyramids_in_fov[NP][pyramid_capacity][3]
 *   points_with_velocity input_cloud_with_velocity: n x 7 (x y z intensity vx vy vz)
 *   stats      8 x int32: particles dropped on a full voxel, on a full pyramid, moved out, newborn
 *              added (cumulative), sequence counters of the three tables, update counter */
int pop_dsp_download_particles(pop_handle* h, int agent, float* out);
int pop_dsp_download_voxels(pop_handle* h, int agent, float* out);
int pop_dsp_download_observations(pop_handle* h, int agent, float* pc, int32_t* counts, float* max_length);
int pop_dsp_download_pyramids(pop_handle* h, int agent, int32_t* out);
int pop_dsp_download_points_with_velocity(pop_handle* h, int agent, float* out, int cap, int32_t* n_out);
int pop_dsp_stats(pop_handle* h, int agent, int32_t* out);
/* getOccupancyMap (dsp_dynamic.h:402-420): centres of the voxels whose occupancy weight exceeds
 * occupancy_threshold, ascending voxel index; out = cap x 3 floats. */
int pop_dsp_occupied(pop_handle* h, int agent, float* out, int cap, int32_t* n_out);

/* ---- K0: depth sensor of the simulator (feeds the particle map without leaving the device) ------ */
/* Pinhole camera of uav_simulator/local_sensing (params/camera.yaml; pcl_render_node.cpp:317-338). */
typedef struct pop_camera {
  float   fx, fy, cx, cy;
  int32_t width, height;
  float   sensing_horizon;       /* metres from the body position (pcl_render_node.cpp:241) */
  int32_t _pad;
  double  cam2body[16];          /* cam02body, row-major 4x4 (:338) */
} pop_camera;
/* reference camera (640x480, fx 387.229, horizon 6 m) scaled to width x height */
void pop_camera_default(pop_camera* cam, int width, int height);
int pop_render_set_camera(pop_handle* h, const pop_camera* cam);
/* DepthRender::render_pose + render_currentpose + render_pcl_world for every local agent
 * (uav_simulator/local_sensing/src/depth_render.cu:2-58, pcl_render_node.cpp:219-285): the world cloud of
 * pop_set_world_cloud is splatted into one depth image per agent (int32 millimetres, 999999 = empty) and
 * back-projected to camera-frame points within the sensing horizon.  pos n_local x 3, quat n_local x 4
 * (w,x,y,z) body poses in double.  The clouds stay on the device as the input of pop_dsp_update_rendered;
 * out_depth_mm (n_local x height x width) and out_counts (n_local) may be NULL. */
int pop_render_depth(pop_handle* h, const double* pos, const double* quat, int32_t* out_depth_mm,
                     int32_t* out_counts);
/* the rendered camera-frame cloud of one agent: out = cap x 3 floats, u-major pixel order */
int pop_render_download_cloud(pop_handle* h, int agent, float* out, int cap, int32_t* n_out);
/* pop_dsp_update on the clouds the last pop_render_depth left on the device (no host cloud at all) */
int pop_dsp_update_rendered(pop_handle* h, const float* poses, const float* quats, double stamp,
                            int32_t* out_ok);

/* ---- after the optimiser: trajectory server, closed loop, evaluation, wire format ------------- */
/* Trajectory-server sampling (traj_server/src/bezier_traj_server.cpp:194-275: traj_.getPos/getVel/getAcc at
 * the command times).  all_agents = 0: the local agents' own last published trajectories; 1: every agent of the
 * swarm as last received.  t_abs n_t absolute times; out_pva (n_agents x n_t x 9: pos, vel, acc), out_valid
 * (n_agents x n_t: 0 outside [start, start + duration] or without a trajectory). */
int pop_traj_sample_batch(pop_handle* h, int all_agents, const double* t_abs, int n_t, double* out_pva,
                          int8_t* out_valid);
/* Closed loop with a perfectly tracking drone (uav_simulator/fake_drone: position command -> odometry): every
 * local agent's resident state (the `start` of the next pop_replan_cycle_resident) and map centre become its
 * commanded state at t_next; past the end of its trajectory it holds the end point, without one it stays. Async. */
int pop_swarm_advance(pop_handle* h, double t_next);
/* The FSM's bookkeeping after a cycle (FiniteStateMachineFake::FSMCallback, plan_manager/src/fake_plan_manager.cpp:85-234):
 * per local agent, from the status of the last cycle — a success resets the failure counter and executes the new
 * trajectory; a failed replan from a trajectory keeps flying the old one and counts; once the count exceeds
 * fsm/replan_max_failures (:187-197), or when a fresh plan fails (:124-131), the agent publishes the reference's "empty"
 * trajectory (publishEmptyTrajectory :389-408: one piece of hover_duration = 1.5 s whose five control points are its
 * current position, traj_id + 1, start = t_now) into the send buffer — it hovers, pop_swarm_advance holds it there and
 * its peers project a stationary body — and plans afresh next time; an executing agent within fsm/goal_tolerance of its
 * goal is done (:149-152) and is left alone afterwards.  out_state / out_failures (n_local, may be NULL). */
enum { POP_FSM_NEW_PLAN = 0, POP_FSM_EXEC_TRAJ = 1, POP_FSM_REPLAN = 2, POP_FSM_GOAL_REACHED = 3 };
typedef struct pop_fsm_params {
  int32_t replan_max_failures;   /* fsm/replan_max_failures, 3 */
  int32_t _pad;
  double  goal_tolerance;        /* fsm/goal_tolerance, 1.0 m */
  double  hover_duration;        /* 1.5 s (fake_plan_manager.cpp:399) */
} pop_fsm_params;
int pop_fsm_update(pop_handle* h, double t_now, const pop_fsm_params* prm, int32_t* out_state, int32_t* out_failures);
/* the resident agent states: out_start n_local x 9 */
int pop_swarm_states(pop_handle* h, double* out_start);
/* MultiAgentLogger::getMinObsDist / getMinUAVDist (eval_helper/include/eval_helper/multi_eval.hpp:240-275) for
 * the local agents at time t: distance to the nearest cylinder surface (as last set) and to the nearest agent. */
int pop_eval_metrics(pop_handle* h, double t, double* out_min_obs, double* out_min_uav);
/* traj_utils/BezierTraj (msg:1-9) in ROS1 serialisation.  to_wire returns the byte count (out = NULL: size
 * only); from_wire returns the bytes consumed.  start/pub times pass through ros::Time (nanoseconds). */
int pop_traj_to_wire(const pop_traj_record* rec, double pub_time, uint8_t* out, int cap);
int pop_traj_from_wire(const uint8_t* in, int n, pop_traj_record* rec, double* pub_time);

/* ---- raycast map: log-odds occupancy from projected depth points ---------------- */
/* Parameters of GridMap (plan_env/src/grid_map.cpp:15-131, "grid_map/..." ROS params). */
typedef struct pop_gridmap_config {
  double  resolution;
  double  map_size[3];            /* map_size_x/y/z; origin = (-x/2, -y/2, ground_height) */
  double  local_update_range[3];
  double  obstacles_inflation;
  double  p_hit, p_miss, p_min, p_max, p_occ;   /* 0.70 0.35 0.12 0.97 0.80 */
  double  max_ray_length;
  double  virtual_ceil_height;
  double  ground_height;
  int32_t local_map_margin;
  int32_t max_points;             /* capacity of one frame's projected points per map */
} pop_gridmap_config;
typedef struct pop_gridmap pop_gridmap;
void pop_gridmap_default_config(pop_gridmap_config* cfg);
/* n_maps independent GridMaps (one per drone) on one GPU, updated by the same launches */
int pop_gridmap_create(const pop_gridmap_config* cfg, int n_maps, int device, void* stream, pop_gridmap** out);
int pop_gridmap_destroy(pop_gridmap* g);
const char* pop_gridmap_last_error(const pop_gridmap* g);
/* updateOccupancyCallback after projectDepthImage (grid_map.cpp:585-640): raycastProcess (:313-456) +
 * clearAndInflateLocalMap (:477-583) for every map.  pts = n_maps x max_points x 3 world-frame doubles
 * (proj_points_), counts n_maps, cam_pos n_maps x 3 (must lie inside the map). */
int pop_gridmap_update(pop_gridmap* g, const double* pts, const int32_t* counts, const double* cam_pos);
/* getOccupancy / getInflateOccupancy (grid_map.h:333-349): -1 outside the map, else 0/1; either output may be NULL */
int pop_gridmap_query(pop_gridmap* g, int map, const double* pos, int n, int8_t* out_occ, int8_t* out_inflate);
/* occupancy_buffer_ (log-odds doubles, x-major), occupancy_buffer_inflate_, local_bound_min/max (6 ints) */
int pop_gridmap_download(pop_gridmap* g, int map, double* occ, int8_t* inflate, int32_t* bounds);
int pop_gridmap_dims(const pop_gridmap* g, int32_t* out3);
/* GridMap::projectDepthImage parameters ("grid_map/..." ROS params, grid_map.cpp:29-42) */
typedef struct pop_depth_params {
  double  fx, fy, cx, cy;
  double  k_depth_scaling_factor;    /* raw depth units per metre (1000 for millimetres) */
  double  depth_filter_mindist, depth_filter_maxdist;
  int32_t use_depth_filter;          /* the reference default is true */
  int32_t depth_filter_margin;
  int32_t skip_pixel;
  int32_t rows, cols;
  int32_t _pad;
} pop_depth_params;
int pop_gridmap_set_depth_params(pop_gridmap* g, const pop_depth_params* p);
/* updateOccupancyCallback from a depth image: projectDepthImage (grid_map.cpp:210-311, including its first-frame
 * skip with the depth filter and its pointer-stepping quirks) + raycastProcess + clearAndInflateLocalMap.
 * depth = n_maps x rows x cols uint16, camera_r = n_maps x 9 row-major (camera_r_m_), cam_pos = n_maps x 3. */
int pop_gridmap_update_depth(pop_gridmap* g, const uint16_t* depth, const double* camera_r, const double* cam_pos);
/* the projected points of the last pop_gridmap_update_depth (proj_points_): out = cap x 3 doubles */
int pop_gridmap_projected_points(pop_gridmap* g, int map, double* out, int cap, int32_t* n_out);

/* ---- measurement ------------------------------------------------------------------------ */
enum { POP_STAGE_VOXEL_UPDATE = 0, POP_STAGE_PEERS = 1, POP_STAGE_ASTAR = 2, POP_STAGE_CORRIDOR = 3,
       POP_STAGE_QP = 4, POP_STAGE_SEPARATION = 5, POP_STAGE_DSP = 6, POP_STAGE_COUNT = 8 };
/* CUDA-event bracketing of the stages on the handle's stream.  on = 1: every stage (the planning
 * stages then run one after the other over the whole batch, so the per-stage times mean something);
 * on = 2: the map stages only (voxel update, peers) — the planning stages stay pipelined over agent
 * groups on side streams, as they are with profiling off; on = 0: off. */
int pop_profile_enable(pop_handle* h, int on);
/* Accumulated milliseconds and launch counts per stage since the last reset. */
int pop_profile_read(pop_handle* h, float* ms /*[POP_STAGE_COUNT]*/,
                     int64_t* launches /*[POP_STAGE_COUNT]*/, int reset);
/* SM cycles the last QP launch spent per problem: out = n x 8 int64 {assembly+scaling+first
 * factorisation, all factorisations, ADMM loop, number of factorisations, and (builds with
 * -DPOP_QP_PHASE_PROF only) gather, KKT solve, row update inside the loop, reserved}. */
int pop_qp_profile(pop_handle* h, int n, int64_t* out);
/* Diagnostic: the device's Seidel LP on its own — min c.x s.t. A x <= b, A row-major rows x d, d = 2..4 — what
 * sdlp::linprog<d> is to the reference (traj_utils/include/traj_utils/sdlp.hpp:708-785), with the shuffle seeded by
 * `seed`.  cooperative = 1 runs the warp-cooperative implementation the planner kernels use, 0 the single-thread one
 * that the host-compiled tests share; both return the same bits.  out_value = the minimum (+inf infeasible, -inf
 * unbounded), out_x = d doubles. */
int pop_lp_solve(pop_handle* h, int d, int rows, const double* c, const double* A, const double* b, uint64_t seed,
                 int cooperative, double* out_x, double* out_value);
/* total kernels launched by this handle since creation */
int64_t pop_kernel_launches(const pop_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* POP_H_ */

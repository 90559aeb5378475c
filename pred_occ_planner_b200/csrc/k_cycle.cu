// k_cycle.cu — K6 separation check and the fused replan cycle (stage glue, trajectory records).
//
// Cycle = FakeBaselinePlanner::replan for every local agent (plan_manager/src/baseline_fake.cpp:
// 253-453) preceded by the map update it reads: SOGM update + peer projection → hybrid A* →
// corridors → QP → ParticleATC::isSafeAfterOpt (traj_coordinator/src/particles.cpp:223-283, LP of
// utils/separator/src/separator_glpk.cpp:73-246) → trajectory record as broadcast by
// FiniteStateMachineFake::publishTrajectory (plan_manager/src/fake_plan_manager.cpp:364-399).
// Everything stays on the device between stages; the host only enqueues kernels.
#include "pop_internal.h"
#include "pop_lp.cuh"

namespace {

constexpr int kSepRows = 2 * POP_MAX_PIECES * 5 + 8;  // ego pts + peer pts + box
constexpr int kSepDoubles = 5 * kSepRows + poplp::work_doubles(kSepRows, 4);
constexpr int kSepInts = poplp::work_ints(kSepRows);
constexpr int kSepWarps = 148 * 8;

// The point sets of one (ego, peer) pair — the ego's new control points and the peer's from the piece that is current at
// t0 on — or false when there is nothing to separate (no trajectory, failed QP, peer not active at t0).
struct SepPair {
  const double* A; const double* B;
  int na, nb, a;
};
__device__ __forceinline__ bool sep_sets(int a0, int n_total, int id_base, double t0, long long pair,
                                         const int* __restrict__ n_pieces, const int* __restrict__ qp_status,
                                         const double* __restrict__ cpts, const pop_traj_record* __restrict__ trajs,
                                         SepPair* out) {
  const int a = a0 + static_cast<int>(pair / n_total), i = static_cast<int>(pair % n_total);
  if (i == id_base + a) return false;
  const int M = n_pieces[a];
  const int st = qp_status[a];
  if (M <= 0 || !(st == 1 || st == 2)) return false;
  const pop_traj_record& r = trajs[i];
  if (r.n_pieces <= 0 || r.n_pieces > POP_MAX_PIECES) return false;
  double t_end = r.start_time;
  for (int k = 0; k < r.n_pieces; k++) t_end += r.duration[k];
  if (!(r.start_time < t0 && t0 < t_end)) return false;
  double rem = t0 - r.start_time;  // Bezier::locatePiece
  int piece = r.n_pieces - 1;
  for (int k = 0; k < r.n_pieces; k++) {
    rem -= static_cast<double>(r.duration[k]);
    if (rem < 0) { piece = k; break; }
  }
  const double* A = cpts + static_cast<size_t>(a) * POP_MAX_PIECES * 15;
  const int na = 5 * M, nb = (r.n_pieces - piece) * 5;
  const double* B = r.cpts + piece * 15;
  out->A = A; out->B = B; out->na = na; out->nb = nb; out->a = a;
  return true;
}
// Screening of a pair's point sets: 0 = safe without the solver (an axis-aligned gap, or the centroid-direction
// certificate holds), 1 = the LP has to decide.
__device__ __forceinline__ int sep_screen(const SepPair& sp) {
  const double* A = sp.A; const double* B = sp.B;
  const int na = sp.na, nb = sp.nb;
  // axis-aligned gap: an LP with margin 1 inside |n|,|d| <= 1e4 is feasible for any gap > 2e-4
  double amin[3] = {1e300, 1e300, 1e300}, amax[3] = {-1e300, -1e300, -1e300};
  double bmin[3] = {1e300, 1e300, 1e300}, bmax[3] = {-1e300, -1e300, -1e300};
  for (int k = 0; k < na; k++) for (int q = 0; q < 3; q++) { amin[q] = fmin(amin[q], A[k * 3 + q]); amax[q] = fmax(amax[q], A[k * 3 + q]); }
  for (int k = 0; k < nb; k++) for (int q = 0; q < 3; q++) { bmin[q] = fmin(bmin[q], B[k * 3 + q]); bmax[q] = fmax(bmax[q], B[k * 3 + q]); }
  for (int q = 0; q < 3; q++) if (bmin[q] - amax[q] > 1e-2 || amin[q] - bmax[q] > 1e-2) return 0;
  // GLPK's variables are free, so the reference's answer does not depend on where the pair sits; inside the box
  // of the LP that only holds with the points taken relative to one of them (A's first), which keeps d small.
  const double o[3] = {A[0], A[1], A[2]};
  // Certificate before the LP: along u = centroid(A) - centroid(B), n = 2 (1 + 1e-9) u / gap and d halfway between the
  // two sets.  If this (n, d) satisfies every row of the LP — checked row by row, in double, with the box — the LP is
  // feasible and the pair is safe; the solver is only needed for pairs that nearly touch or interpenetrate.
  double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
  for (int k = 0; k < na; k++) for (int q = 0; q < 3; q++) ca[q] += A[k * 3 + q] - o[q];
  for (int k = 0; k < nb; k++) for (int q = 0; q < 3; q++) cb[q] += B[k * 3 + q] - o[q];
  double u[3];
  for (int q = 0; q < 3; q++) u[q] = ca[q] / na - cb[q] / nb;
  double lo_a = 1e300, hi_b = -1e300;
  for (int k = 0; k < na; k++) lo_a = fmin(lo_a, u[0] * (A[k * 3] - o[0]) + u[1] * (A[k * 3 + 1] - o[1]) + u[2] * (A[k * 3 + 2] - o[2]));
  for (int k = 0; k < nb; k++) hi_b = fmax(hi_b, u[0] * (B[k * 3] - o[0]) + u[1] * (B[k * 3 + 1] - o[1]) + u[2] * (B[k * 3 + 2] - o[2]));
  const double gap = lo_a - hi_b;
  if (gap > 0.0) {
    const double sc = 2.0 * (1.0 + 1e-9) / gap;
    const double nn[3] = {sc * u[0], sc * u[1], sc * u[2]};
    const double dd = -0.5 * sc * (lo_a + hi_b);
    bool ok = fabs(nn[0]) <= 9.99e3 && fabs(nn[1]) <= 9.99e3 && fabs(nn[2]) <= 9.99e3 && fabs(dd) <= 9.99e3;
    for (int k = 0; ok && k < na; k++) ok = nn[0] * (A[k * 3] - o[0]) + nn[1] * (A[k * 3 + 1] - o[1]) + nn[2] * (A[k * 3 + 2] - o[2]) + dd >= 1.0;
    for (int k = 0; ok && k < nb; k++) ok = nn[0] * (B[k * 3] - o[0]) + nn[1] * (B[k * 3 + 1] - o[1]) + nn[2] * (B[k * 3 + 2] - o[2]) + dd <= -1.0;
    if (ok) return 0;
  }
  return 1;
}

// Separation in two launches.  k_sep_screen: one thread per (ego, peer) pair — almost every pair of a swarm is settled
// there; the pairs that nearly touch are appended to the launch's list.  k_sep_lp: one warp (lane 0: Seidel's LP is a
// sequential algorithm) per listed pair, so the few LPs of a cycle run side by side instead of behind one another.
// Agents [a0, a0 + n_local) against every peer; list and counter of the launch are those of agent a0.
__global__ void __launch_bounds__(128) k_sep_screen(int a0, int n_local, int n_total, int id_base, double t0,
                                                    const int* __restrict__ n_pieces, const int* __restrict__ qp_status,
                                                    const double* __restrict__ cpts,
                                                    const pop_traj_record* __restrict__ trajs, int* __restrict__ list,
                                                    int* __restrict__ count) {
  const long long pair = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pair >= static_cast<long long>(n_local) * n_total) return;
  SepPair sp;
  if (!sep_sets(a0, n_total, id_base, t0, pair, n_pieces, qp_status, cpts, trajs, &sp)) return;
  if (!sep_screen(sp)) return;
  list[static_cast<size_t>(a0) * n_total + atomicAdd(&count[a0], 1)] = static_cast<int>(pair);
}

// warp w of this launch works in workspace slot w0 + w
__global__ void __launch_bounds__(128) k_sep_lp(int a0, int w0, int n_total, int id_base, double t0,
                                                const int* __restrict__ n_pieces, const int* __restrict__ qp_status,
                                                const double* __restrict__ cpts,
                                                const pop_traj_record* __restrict__ trajs, const int* __restrict__ list,
                                                const int* __restrict__ count, double* __restrict__ wd,
                                                int* __restrict__ wi, int* __restrict__ safe) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total_warps = (gridDim.x * blockDim.x) >> 5;
  double* rowsM = wd + static_cast<size_t>(w0 + warp) * kSepDoubles;
  double* rhs = rowsM + 4 * kSepRows;
  double* lpw = rhs + kSepRows;
  int* lpi = wi + static_cast<size_t>(w0 + warp) * kSepInts;
  const int n = count[a0];
  for (int e = warp; e < n; e += total_warps) {
    SepPair sp;
    if (!sep_sets(a0, n_total, id_base, t0, list[static_cast<size_t>(a0) * n_total + e], n_pieces, qp_status, cpts, trajs, &sp)) continue;
    const double* A = sp.A; const double* B = sp.B;
    const int na = sp.na, nb = sp.nb;
    const double o[3] = {A[0], A[1], A[2]};
    __syncwarp();
    for (int k = lane; k < na + nb + 8; k += 32) {
      if (k < na) {
        for (int q = 0; q < 3; q++) rowsM[k * 4 + q] = -(A[k * 3 + q] - o[q]);
        rowsM[k * 4 + 3] = -1.0; rhs[k] = -1.0;
      } else if (k < na + nb) {
        const int kb = k - na;
        for (int q = 0; q < 3; q++) rowsM[k * 4 + q] = B[kb * 3 + q] - o[q];
        rowsM[k * 4 + 3] = 1.0; rhs[k] = -1.0;
      } else {   // the box |n_i|, |d| <= 1e4, two rows per variable
        const int kk = k - na - nb, var = kk >> 1;
        for (int q = 0; q < 4; q++) rowsM[k * 4 + q] = q == var ? ((kk & 1) ? -1.0 : 1.0) : 0.0;
        rhs[k] = 1e4;
      }
    }
    __syncwarp();
    const int rr = na + nb + 8;
    const double c[4] = {0, 0, 0, 0};
    double x[4];
    unsigned long long rng = 0x5eed5eedull + static_cast<unsigned long long>(na) * 131 + nb;
    const double v = poplp::warp::linprog<4>(c, rowsM, 4, rhs, 1, rr, x, &rng, lpw, lpi);
    if (lane == 0 && isinf(v)) atomicExch(&safe[sp.a], 0);
  }
}

__global__ void k_fill_times(int n, double t_rel, double t_abs, double* tstart, double* tst, int* safe, int* sep_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  tstart[i] = t_rel;
  tst[i] = t_abs;
  safe[i] = 1;
  sep_count[i] = 0;
}

// QP inputs from the A* / corridor outputs (baseline_fake.cpp:393-427)
__global__ void k_prepare_qp(int a0, int n, double tau, const double* __restrict__ start, const int* __restrict__ ret,
                             const int* __restrict__ npoly, const double* __restrict__ cgoal,
                             int* __restrict__ n_pieces, double* __restrict__ init, double* __restrict__ goal,
                             double* __restrict__ dur) {
  const int a = a0 + blockIdx.x * blockDim.x + threadIdx.x;   // n = end of the range
  if (a >= n) return;
  const int M = ret[a] == POP_ASTAR_NO_PATH ? 0 : npoly[a];
  n_pieces[a] = M;
  for (int k = 0; k < 9; k++) init[a * 9 + k] = start[a * 9 + k];
  for (int k = 0; k < 6; k++) goal[a * 9 + k] = cgoal[a * 6 + k];
  for (int k = 6; k < 9; k++) goal[a * 9 + k] = 0.0;
  for (int k = 0; k < POP_MAX_PIECES; k++) dur[a * POP_MAX_PIECES + k] = k < M ? tau : 0.0;
}

__global__ void k_finalize(int a0, int n, int id_base, double t_plan, double tau, const int* __restrict__ ret,
                           const int* __restrict__ npoly, const int* __restrict__ qp_status,
                           const int* __restrict__ safe, const double* __restrict__ cpts,
                           int* __restrict__ status, pop_traj_record* __restrict__ send, int* __restrict__ traj_ids) {
  const int a = a0 + blockIdx.x * blockDim.x + threadIdx.x;   // n = end of the range
  if (a >= n) return;
  int st = POP_REPLAN_OK;
  if (ret[a] == POP_ASTAR_NO_PATH || ret[a] == POP_ASTAR_SEARCH_ERR) st = POP_REPLAN_FAIL_ASTAR;
  else if (npoly[a] <= 0) st = POP_REPLAN_FAIL_CORRIDOR;
  else if (!(qp_status[a] == 1 || qp_status[a] == 2)) st = POP_REPLAN_FAIL_QP;
  else if (!safe[a]) st = POP_REPLAN_FAIL_SEPARATION;
  status[a] = st;
  if (st != POP_REPLAN_OK) return;  // nothing is published; peers keep the previous trajectory
  pop_traj_record& r = send[a];
  const int M = npoly[a];
  r.drone_id = id_base + a;
  r.traj_id = ++traj_ids[a];
  r.n_pieces = M;
  r.valid = 1;
  r.start_time = t_plan;
  for (int k = 0; k < POP_MAX_PIECES; k++) r.duration[k] = k < M ? static_cast<float>(tau) : 0.f;
  const double* c = cpts + static_cast<size_t>(a) * POP_MAX_PIECES * 15;
  for (int k = 0; k < POP_MAX_PIECES * 15; k++) r.cpts[k] = k < 15 * M ? c[k] : 0.0;
}

// FiniteStateMachineFake::FSMCallback, the branches that follow a replan (plan_manager/src/fake_plan_manager.cpp):
//   REPLAN (:157-197): success -> num_replan_failures_ = 0, EXEC_TRAJ; failure -> counter + 1, and once it exceeds
//     fsm/replan_max_failures -> NEW_PLAN + publishEmptyTrajectory (the counter is only ever reset by a success);
//   NEW_PLAN (:117-138): success -> EXEC_TRAJ; failure -> publishEmptyTrajectory;
//   EXEC_TRAJ (:149-152): |odom - goal| < fsm/goal_tolerance -> GOAL_REACHED.
// publishEmptyTrajectory (:389-408): traj_id + 1, start = now, one piece of 1.5 s whose five control points are the
// current position — the drone hovers, and its peers project a stationary body.
__global__ void k_fsm_update(int n, int id_base, double t_now, int max_failures, double goal_tol, double hover_dur,
                             const int* __restrict__ status, const double* __restrict__ start,
                             const double* __restrict__ goal, pop_traj_record* __restrict__ send,
                             int* __restrict__ traj_ids, int* __restrict__ fsm_state, int* __restrict__ fsm_fail) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  int st = fsm_state[a], fails = fsm_fail[a];
  if (st == POP_FSM_GOAL_REACHED) return;
  const bool ok = status[a] == POP_REPLAN_OK;
  bool hover = false;
  if (st == POP_FSM_NEW_PLAN) {
    if (ok) st = POP_FSM_EXEC_TRAJ; else hover = true;
  } else {   // EXEC_TRAJ / REPLAN: a replan from the current trajectory
    if (ok) { fails = 0; st = POP_FSM_EXEC_TRAJ; }
    else {
      fails++;
      st = POP_FSM_REPLAN;
      if (fails > max_failures) { st = POP_FSM_NEW_PLAN; hover = true; }
    }
  }
  const double* p = start + 9 * a;
  if (hover) {
    pop_traj_record& r = send[a];
    r.drone_id = id_base + a;
    r.traj_id = ++traj_ids[a];
    r.n_pieces = 1;
    r.valid = 1;
    r.start_time = t_now;
    for (int k = 0; k < POP_MAX_PIECES; k++) r.duration[k] = k == 0 ? static_cast<float>(hover_dur) : 0.f;
    for (int k = 0; k < POP_MAX_PIECES * 15; k++) r.cpts[k] = k < 15 ? p[k % 3] : 0.0;
  }
  if (st == POP_FSM_EXEC_TRAJ) {
    const double dx = p[0] - goal[3 * a], dy = p[1] - goal[3 * a + 1], dz = p[2] - goal[3 * a + 2];
    if (sqrt(dx * dx + dy * dy + dz * dz) < goal_tol) st = POP_FSM_GOAL_REACHED;
  }
  fsm_state[a] = st;
  fsm_fail[a] = fails;
}

}  // namespace

struct PlanDev {
  double* sep_wd = nullptr;
  int* sep_wi = nullptr;
  int* safe = nullptr;      // [n_local]
  int* status = nullptr;    // [n_local]
  int* traj_ids = nullptr;  // [n_local]
  int* fsm_state = nullptr; // [n_local] POP_FSM_*
  int* fsm_fail = nullptr;  // [n_local] num_replan_failures_
  int* sep_np = nullptr;    // staging for pop_separation_check_batch
  int* sep_st = nullptr;
  double* sep_cpts = nullptr;
  int* sep_list = nullptr;  // [n_local][n_total] pairs the screening left to the LP, per launch from its first agent's row
  int* sep_count = nullptr; // [n_local]
};

static int plan_alloc(pop_handle* h) {
  if (h->plan) return POP_OK;
  PlanDev* d = new PlanDev();
  h->plan = d;
  const int nl = h->cfg.n_agents_local;
  bool ok = true;
  ok &= cudaMalloc(&d->sep_wd, sizeof(double) * kSepDoubles * kSepWarps) == cudaSuccess;
  ok &= cudaMalloc(&d->sep_wi, sizeof(int) * kSepInts * kSepWarps) == cudaSuccess;
  ok &= cudaMalloc(&d->safe, sizeof(int) * nl) == cudaSuccess;
  ok &= cudaMalloc(&d->status, sizeof(int) * nl) == cudaSuccess;
  ok &= cudaMalloc(&d->traj_ids, sizeof(int) * nl) == cudaSuccess;
  ok &= cudaMalloc(&d->fsm_state, sizeof(int) * nl) == cudaSuccess;
  ok &= cudaMalloc(&d->fsm_fail, sizeof(int) * nl) == cudaSuccess;
  ok &= cudaMalloc(&d->sep_np, sizeof(int) * nl) == cudaSuccess;
  ok &= cudaMalloc(&d->sep_st, sizeof(int) * nl) == cudaSuccess;
  ok &= cudaMalloc(&d->sep_cpts, sizeof(double) * POP_MAX_PIECES * 15 * nl) == cudaSuccess;
  ok &= cudaMalloc(&d->sep_list, sizeof(int) * static_cast<size_t>(nl) * h->cfg.n_agents_total) == cudaSuccess;
  ok &= cudaMalloc(&d->sep_count, sizeof(int) * nl) == cudaSuccess;
  auto finish = [&]() -> int {
    if (!ok) { cudaGetLastError(); return pop_fail(h, POP_E_CUDA, "cycle workspace allocation failed"); }
    POP_CUDA(h, cudaMemsetAsync(d->traj_ids, 0, sizeof(int) * nl, h->stream));
    POP_CUDA(h, cudaMemsetAsync(d->fsm_state, 0, sizeof(int) * nl, h->stream));   // POP_FSM_NEW_PLAN
    POP_CUDA(h, cudaMemsetAsync(d->fsm_fail, 0, sizeof(int) * nl, h->stream));
    POP_CUDA(h, cudaMemsetAsync(d->status, 0, sizeof(int) * nl, h->stream));
    return POP_OK;
  };
  const int rc = finish();
  if (rc != POP_OK) {   // leave no half-built workspace behind: the next call allocates again
    void* ptrs[] = {d->sep_wd, d->sep_wi, d->safe, d->status, d->traj_ids, d->fsm_state, d->fsm_fail, d->sep_np, d->sep_st, d->sep_cpts, d->sep_list, d->sep_count};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete d;
    h->plan = nullptr;
  }
  return rc;
}

void pop_plan_free(pop_handle* h) {
  popk_astar_free(h);
  popk_corridor_free(h);
  popk_qp_free(h);
  PlanDev* d = h->plan;
  if (!d) return;
  void* ptrs[] = {d->sep_wd, d->sep_wi, d->safe, d->status, d->traj_ids, d->fsm_state, d->fsm_fail, d->sep_np, d->sep_st, d->sep_cpts, d->sep_list, d->sep_count};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete d;
  h->plan = nullptr;
}

// separation check of agents [a0, a0+na) on stream st with the workspace slots [w0, w0+nw)
static int run_separation(pop_handle* h, double t_now, const int* d_np, const int* d_st, const double* d_cpts,
                          cudaStream_t st, int a0, int na, int w0, int nw, bool timed) {
  PlanDev* d = h->plan;
  if (timed) pop_prof_begin(h, POP_STAGE_SEPARATION);
  int launches = 0;
  if (h->have_traj && h->cfg.n_agents_total > 1) {
    const int nt = h->cfg.n_agents_total;
    const long long pairs = static_cast<long long>(na) * nt;
    k_sep_screen<<<static_cast<unsigned>((pairs + 127) / 128), 128, 0, st>>>(a0, na, nt, h->cfg.agent_id_base, t_now, d_np, d_st,
                                                                             d_cpts, h->d_traj_all, d->sep_list, d->sep_count);
    POP_LAUNCH_CHECK(h);
    const int blocks = nw / 4 > 0 ? nw / 4 : 1;
    k_sep_lp<<<blocks, 128, 0, st>>>(a0, w0, nt, h->cfg.agent_id_base, t_now, d_np, d_st, d_cpts, h->d_traj_all, d->sep_list,
                                     d->sep_count, d->sep_wd, d->sep_wi, d->safe);
    POP_LAUNCH_CHECK(h);
    launches = 2;
  }
  if (timed) pop_prof_end(h, POP_STAGE_SEPARATION, launches);
  else h->n_launches += launches;
  return POP_OK;
}

// number of agent groups the planning stages are pipelined over (POP_PLAN_GROUPS; default = the
// number of hardware work queues, at most 32);
// creates the side streams on first use
static int plan_groups(pop_handle* h, int nl) {
  static int want = -1;
  if (want < 0) {
    // more groups than hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS, 8 unless raised before the
    // context is created) serialise falsely; measured on B200 at 64 agents: 8 groups 20.8 ms, 32 groups
    // with 32 connections 19.5 ms, 16 groups with 8 connections 33.9 ms per cycle
    const char* c = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
    const int conn = c ? atoi(c) : 8;
    const char* e = getenv("POP_PLAN_GROUPS");
    want = e ? atoi(e) : (conn >= 32 ? 32 : (conn >= 8 ? 8 : conn));
    if (want < 1) want = 1;
    if (want > 64) want = 64;
  }
  const int g = want < nl ? want : nl;
  if (g <= 1) return 1;
  while (static_cast<int>(h->side.size()) < g) {
    cudaStream_t s;
    cudaEvent_t e;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return 1;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaStreamDestroy(s); return 1; }
    h->side.push_back(s);
    h->side_done.push_back(e);
  }
  if (!h->fork_ev && cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming) != cudaSuccess) return 1;
  return g;
}

// the whole cycle on the stream; inputs (cloud, cylinders, poses, peers, start, goal) resident
static int cycle_enqueue(pop_handle* h, double t_map, double t_plan, bool update_map = true) {
  int rc;
  if ((rc = popk_astar_alloc(h))) return rc;
  if ((rc = popk_corridor_alloc(h))) return rc;
  if ((rc = popk_qp_alloc(h, h->cfg.n_agents_local))) return rc;
  if ((rc = plan_alloc(h))) return rc;
  PlanDev* d = h->plan;
  const pop_config& c = h->cfg;
  const int nl = c.n_agents_local;
  if (update_map) {
    h->t_map = t_map;
    if ((rc = popk_voxel_update(h))) return rc;
    if ((rc = popk_peer_project(h))) return rc;
    h->map_ready = true;
  } else {
    t_map = h->t_map;   // the SOGM as the last map update (ground truth or particle map) left it
  }
  k_fill_times<<<(nl + 127) / 128, 128, 0, h->stream>>>(nl, t_plan - t_map, t_plan, popk_astar_tstart_dev(h),
                                                       popk_corridor_tst_dev(h), d->safe, d->sep_count);
  POP_LAUNCH_CHECK(h);
  // A* → corridors → QP are independent per agent: unless every stage is being timed, run them as
  // chains over agent groups on side streams, so one agent's 4 000-iteration QP overlaps the other
  // groups' searches and corridors instead of waiting for the slowest agent of every stage.
  const int n_groups = (h->prof.enabled && !h->prof.map_only) ? 1 : plan_groups(h, nl);
  if (n_groups > 1) {
    POP_CUDA(h, cudaEventRecord(h->fork_ev, h->stream));
    for (int gi = 0; gi < n_groups; gi++) POP_CUDA(h, cudaStreamWaitEvent(h->side[gi], h->fork_ev, 0));
  }
  const int per = (nl + n_groups - 1) / n_groups;
  // one group's chain on its stream; a failure part-way leaves work queued on the side streams, so the caller below
  // joins every stream that was forked before reporting it (the handle's stream then still orders after all of it)
  auto enqueue_group = [&](int gi, int a0, int na, cudaStream_t st) -> int {
    int rc;
    if ((rc = popk_astar_run(h, st, a0, na))) return rc;
    if ((rc = popk_corridor_run(h, popk_astar_path_dev(h), popk_astar_len_dev(h), popk_astar_start_dev(h), st, a0, na))) return rc;
    k_prepare_qp<<<(na + 127) / 128, 128, 0, st>>>(a0, a0 + na, c.corridor_tau, popk_astar_start_dev(h),
                                                  popk_astar_ret_dev(h), popk_corridor_npoly_dev(h),
                                                  popk_corridor_goal_dev(h), popk_qp_npieces_dev(h),
                                                  popk_qp_init_dev(h), popk_qp_goal_dev(h), popk_qp_dur_dev(h));
    POP_LAUNCH_CHECK(h);
    if ((rc = popk_qp_run(h, na, popk_corridor_hpoly_dev(h), popk_corridor_rows_dev(h), c.opt_max_vel, c.opt_max_acc, st, a0))) return rc;
    const int wper = (kSepWarps / n_groups) & ~3;   // workspace slots of this group (multiple of 4 warps)
    if ((rc = run_separation(h, t_plan, popk_qp_npieces_dev(h), popk_qp_status_dev(h), popk_qp_cpts_dev(h), st, a0, na,
                             gi * wper, wper, n_groups == 1))) return rc;
    k_finalize<<<(na + 127) / 128, 128, 0, st>>>(a0, a0 + na, c.agent_id_base, t_plan, c.corridor_tau, popk_astar_ret_dev(h),
                                                popk_qp_npieces_dev(h), popk_qp_status_dev(h), d->safe,
                                                popk_qp_cpts_dev(h), d->status, h->d_traj_send, d->traj_ids);
    POP_LAUNCH_CHECK(h);
    h->n_launches += 2;   // k_prepare_qp + k_finalize
    return POP_OK;
  };
  int first_error = POP_OK;
  for (int gi = 0; gi < n_groups; gi++) {
    const int a0 = gi * per, na = min(per, nl - a0);
    if (na <= 0) break;
    cudaStream_t st = n_groups > 1 ? h->side[gi] : h->stream;
    if (first_error == POP_OK) first_error = enqueue_group(gi, a0, na, st);
    if (n_groups > 1) {   // also after an error: every forked stream is joined
      cudaEventRecord(h->side_done[gi], st);
      cudaStreamWaitEvent(h->stream, h->side_done[gi], 0);
    }
  }
  if (first_error != POP_OK) return first_error;
  h->n_launches += 1;   // k_fill_times
  return POP_OK;
}

extern "C" {

int pop_separation_check_batch(pop_handle* h, const int32_t* n_pieces, const double* cpts, double t_now,
                               int8_t* out_safe) {
  if (!h || !n_pieces || !cpts || !out_safe) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  int rc = plan_alloc(h);
  if (rc) return rc;
  PlanDev* d = h->plan;
  const int nl = h->cfg.n_agents_local;
  std::vector<int> ones(nl, 1);
  for (int i = 0; i < nl; i++) if (n_pieces[i] < 0 || n_pieces[i] > POP_MAX_PIECES) return POP_E_INVALID;
  POP_CUDA(h, cudaMemcpyAsync(d->sep_np, n_pieces, sizeof(int) * nl, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d->sep_st, ones.data(), sizeof(int) * nl, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d->safe, ones.data(), sizeof(int) * nl, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemsetAsync(d->sep_count, 0, sizeof(int) * nl, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d->sep_cpts, cpts, sizeof(double) * POP_MAX_PIECES * 15 * nl, cudaMemcpyHostToDevice, h->stream));
  rc = run_separation(h, t_now, d->sep_np, d->sep_st, d->sep_cpts, h->stream, 0, nl, 0, kSepWarps, true);
  if (rc) return rc;
  std::vector<int> safe(nl);
  POP_CUDA(h, cudaMemcpyAsync(safe.data(), d->safe, sizeof(int) * nl, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < nl; i++) out_safe[i] = static_cast<int8_t>(safe[i]);
  return POP_OK;
}

int pop_fsm_update(pop_handle* h, double t_now, const pop_fsm_params* prm, int32_t* out_state, int32_t* out_failures) {
  if (!h || !prm || prm->replan_max_failures < 0 || !(prm->goal_tolerance >= 0.0) || !(prm->hover_duration > 0.0)) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->plan || !h->astar) return pop_fail(h, POP_E_STATE, "no cycle has run yet");
  PlanDev* d = h->plan;
  const int nl = h->cfg.n_agents_local;
  k_fsm_update<<<(nl + 127) / 128, 128, 0, h->stream>>>(nl, h->cfg.agent_id_base, t_now, prm->replan_max_failures, prm->goal_tolerance,
                                                       prm->hover_duration, d->status, popk_astar_start_dev(h), popk_astar_goal_dev(h),
                                                       h->d_traj_send, d->traj_ids, d->fsm_state, d->fsm_fail);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  if (out_state) POP_CUDA(h, cudaMemcpyAsync(out_state, d->fsm_state, sizeof(int) * nl, cudaMemcpyDeviceToHost, h->stream));
  if (out_failures) POP_CUDA(h, cudaMemcpyAsync(out_failures, d->fsm_fail, sizeof(int) * nl, cudaMemcpyDeviceToHost, h->stream));
  if (out_state || out_failures) POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}

int pop_replan_cycle_resident(pop_handle* h, double t_map, double t_plan) {
  if (!h) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->astar) return pop_fail(h, POP_E_STATE, "no resident inputs: call pop_replan_cycle (or pop_astar_search_batch) once first");
  return cycle_enqueue(h, t_map, t_plan);
}

static int replan_cycle_io(pop_handle* h, const pop_cycle_io* io, bool update_map);

int pop_replan_cycle(pop_handle* h, const pop_cycle_io* io) {
  if (!h || !io || !io->start || !io->goal || !io->poses || !io->status || !io->traj) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  return replan_cycle_io(h, io, true);
}

int pop_plan_cycle(pop_handle* h, const pop_cycle_io* io) {
  if (!h || !io || !io->start || !io->goal || !io->status || !io->traj) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->map_ready) return pop_fail(h, POP_E_STATE, "no map yet: call pop_dsp_update or pop_map_update_gt first");
  return replan_cycle_io(h, io, false);
}

static int replan_cycle_io(pop_handle* h, const pop_cycle_io* io, bool update_map) {
  int rc;
  if ((rc = popk_astar_alloc(h))) return rc;
  if ((rc = plan_alloc(h))) return rc;
  const int nl = h->cfg.n_agents_local;
  if (update_map)
    POP_CUDA(h, cudaMemcpyAsync(h->d_poses, io->poses, sizeof(float) * 3 * nl, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(popk_astar_start_dev(h), io->start, sizeof(double) * 9 * nl, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(popk_astar_goal_dev(h), io->goal, sizeof(double) * 3 * nl, cudaMemcpyHostToDevice, h->stream));
  if ((rc = cycle_enqueue(h, io->t_map, io->t_plan, update_map))) return rc;
  POP_CUDA(h, cudaMemcpyAsync(io->status, h->plan->status, sizeof(int) * nl, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(io->traj, h->d_traj_send, sizeof(pop_traj_record) * nl, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < nl; i++) if (io->status[i] != POP_REPLAN_OK) io->traj[i].valid = 0;
  return popk_corridor_check_err(h);
}

}  // extern "C"

extern "C" int pop_cycle_results(pop_handle* h, int32_t* status, pop_traj_record* traj) {
  if (!h) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->plan) return pop_fail(h, POP_E_STATE, "no cycle has run yet");
  const int nl = h->cfg.n_agents_local;
  if (status) POP_CUDA(h, cudaMemcpyAsync(status, h->plan->status, sizeof(int) * nl, cudaMemcpyDeviceToHost, h->stream));
  if (traj) POP_CUDA(h, cudaMemcpyAsync(traj, h->d_traj_send, sizeof(pop_traj_record) * nl, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return popk_corridor_check_err(h);
}

extern "C" int pop_traj_publish_local(pop_handle* h) {
  if (!h) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  POP_CUDA(h, cudaMemcpyAsync(h->d_traj_all + h->cfg.agent_id_base, h->d_traj_send,
                              sizeof(pop_traj_record) * h->cfg.n_agents_local, cudaMemcpyDeviceToDevice, h->stream));
  h->have_traj = true;
  return POP_OK;
}

// ---- diagnostic: the device LP on its own (both implementations), for tests ---------------------------------------
namespace {
template <int d>
__global__ void k_lp_solve(int rows, const double* __restrict__ c, const double* __restrict__ A, const double* __restrict__ b,
                           unsigned long long seed, int cooperative, double* __restrict__ out, double* wd, int* wi) {
  double x[d];
  unsigned long long rng = seed;
  double v;
  if (cooperative) {
    v = poplp::warp::linprog<d>(c, A, d, b, 1, rows, x, &rng, wd, wi);
  } else {
    if (threadIdx.x != 0) return;
    v = poplp::linprog<d>(c, A, d, b, 1, rows, x, &rng, wd, wi);
  }
  if (threadIdx.x == 0) {
    out[0] = v;
    for (int k = 0; k < d; k++) out[1 + k] = x[k];
  }
}
}  // namespace

extern "C" int pop_lp_solve(pop_handle* h, int d, int rows, const double* c, const double* A, const double* b,
                            uint64_t seed, int cooperative, double* out_x, double* out_value) {
  if (!h || d < 2 || d > 4 || rows < 0 || rows > 4096 || !c || (rows > 0 && (!A || !b)) || !out_x || !out_value) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  const size_t nwd = poplp::work_doubles(rows, 4) + 8, nwi = poplp::work_ints(rows) + 8;
  double *dc = nullptr, *dA = nullptr, *db = nullptr, *dout = nullptr, *wd = nullptr;
  int* wi = nullptr;
  bool ok = cudaMalloc(&dc, sizeof(double) * 4) == cudaSuccess && cudaMalloc(&dA, sizeof(double) * (rows * d + 1)) == cudaSuccess &&
            cudaMalloc(&db, sizeof(double) * (rows + 1)) == cudaSuccess && cudaMalloc(&dout, sizeof(double) * 5) == cudaSuccess &&
            cudaMalloc(&wd, sizeof(double) * nwd) == cudaSuccess && cudaMalloc(&wi, sizeof(int) * nwi) == cudaSuccess;
  double out[5] = {0, 0, 0, 0, 0};
  if (ok) {
    cudaMemcpyAsync(dc, c, sizeof(double) * d, cudaMemcpyHostToDevice, h->stream);
    if (rows > 0) {
      cudaMemcpyAsync(dA, A, sizeof(double) * rows * d, cudaMemcpyHostToDevice, h->stream);
      cudaMemcpyAsync(db, b, sizeof(double) * rows, cudaMemcpyHostToDevice, h->stream);
    }
    if (d == 2) k_lp_solve<2><<<1, 32, 0, h->stream>>>(rows, dc, dA, db, seed, cooperative, dout, wd, wi);
    else if (d == 3) k_lp_solve<3><<<1, 32, 0, h->stream>>>(rows, dc, dA, db, seed, cooperative, dout, wd, wi);
    else k_lp_solve<4><<<1, 32, 0, h->stream>>>(rows, dc, dA, db, seed, cooperative, dout, wd, wi);
    cudaMemcpyAsync(out, dout, sizeof(out), cudaMemcpyDeviceToHost, h->stream);
    ok = cudaStreamSynchronize(h->stream) == cudaSuccess;
  }
  void* ptrs[] = {dc, dA, db, dout, wd, wi};
  for (void* q : ptrs) if (q) cudaFree(q);
  if (!ok) { cudaGetLastError(); return pop_fail(h, POP_E_CUDA, "pop_lp_solve: allocation or launch failed"); }
  *out_value = out[0];
  for (int k = 0; k < d; k++) out_x[k] = out[1 + k];
  return POP_OK;
}

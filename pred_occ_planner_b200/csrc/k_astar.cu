// k_astar.cu — K2: batched kinodynamic hybrid A* in (x, y, z, t) over the SOGM.
//
// One CTA per agent.  An expansion evaluates its 75 motion primitives in parallel (state
// transition, closed-set lookup, velocity/same-cell pruning, SOGM region-risk gathers,
// quartic heuristic), then one thread commits them in the reference's order (same-parent
// pruning, node allocation, heap push) and the new nodes are published to the hash in parallel.
// Open set and f-scores live in shared memory; node payloads and the 4-D hash in global memory.
//
// Restates path_searching/src/fake_risk_hybrid_a_star.cpp:114-421 (search), :423-457
// (estimateHeuristic), :459-519 (computeShotTraj), :521-585 (cubic/quartic), :663-694
// (getPathWithVel), :795-836, with the call sequence of plan_manager/src/baseline_fake.cpp:264-296
// (search with init=true, one retry with init=false when NO_PATH).  The open set reproduces
// std::priority_queue (libstdc++ __push_heap/__adjust_heap) including the reference's in-place
// re-scoring without re-heapify; hash inserts use the (int)time key quirk of :387.
#include <cmath>

#include "pop_bezier.cuh"
#include "pop_geom.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kMaxPrim = 75;
constexpr int kSmemNodes = 16384;  // f (double) + heap (int) = 12 B per node in shared memory

enum { IN_CLOSE_SET = 1, IN_OPEN_SET = 2, NOT_EXPAND = 3 };

struct AstarParams {
  double max_tau, max_vel, max_acc, w_time, horizon, lambda_heu;
  double inv_res, inv_tres, tau, tie_breaker, corridor_tau;
  int cap, check_num, tolerance, hash_mask;
};

}  // namespace

struct AstarDev {
  int cap = 0, hash_cap = 0, n_agents = 0;
  double* state = nullptr;   // [agent][cap][6]
  double* input = nullptr;   // [agent][cap][3]
  double* time = nullptr;    // [agent][cap]
  double* g = nullptr;       // [agent][cap]
  int* parent = nullptr;     // [agent][cap]
  int* time_idx = nullptr;   // [agent][cap]
  int* index = nullptr;      // [agent][cap][3]
  int* nstate = nullptr;     // [agent][cap]
  unsigned long long* hkey = nullptr;  // [agent][hash_cap]  gen:16 | x:14 | y:14 | z:14 | t:6
  int* hval = nullptr;       // [agent][hash_cap]
  unsigned* gen = nullptr;   // [agent] search generation (hash entries of older generations are free)
  double* d_start = nullptr; // [agent][9]
  double* d_goal = nullptr;  // [agent][3]
  double* d_tstart = nullptr;
  int* d_ret = nullptr;
  double* d_path = nullptr;  // [agent][POP_MAX_PATH][6]
  int* d_len = nullptr;
  int* d_stats = nullptr;    // [agent][4]
};

namespace {

__device__ __forceinline__ unsigned long long pack_key(unsigned gen, int x, int y, int z, int t) {
  return (static_cast<unsigned long long>(gen & 0xffffu) << 48) |
         (static_cast<unsigned long long>(x & 0x3fff) << 34) |
         (static_cast<unsigned long long>(y & 0x3fff) << 20) |
         (static_cast<unsigned long long>(z & 0x3fff) << 6) | static_cast<unsigned long long>(t & 0x3f);
}
__device__ __forceinline__ unsigned hash_slot(unsigned long long k, unsigned mask) {
  k &= 0xffffffffffffull;
  k ^= k >> 23; k *= 0x2127599bf4325c37ull; k ^= k >> 29;
  return static_cast<unsigned>(k) & mask;
}
__device__ int hash_find(const unsigned long long* hkey, const int* hval, unsigned mask, unsigned gen,
                         unsigned long long key) {
  unsigned pos = hash_slot(key, mask);
  for (unsigned probe = 0; probe <= mask; probe++) {
    const unsigned long long k = hkey[pos];
    if ((k >> 48) != (gen & 0xffffu)) return -1;
    if (k == key) return hval[pos];
    pos = (pos + 1) & mask;
  }
  return -1;
}
// unordered_map::insert semantics: keeps the existing entry when the key is present
__device__ void hash_insert(unsigned long long* hkey, int* hval, unsigned mask, unsigned gen,
                            unsigned long long key, int val) {
  unsigned pos = hash_slot(key, mask);
  for (unsigned probe = 0; probe <= mask; probe++) {
    unsigned long long k = hkey[pos];
    while ((k >> 48) != (gen & 0xffffu)) {
      const unsigned long long old = atomicCAS(&hkey[pos], k, key);
      if (old == k) { hval[pos] = val; return; }
      k = old;
    }
    if (k == key) return;
    pos = (pos + 1) & mask;
  }
}

// stateTransit (:812-824): state1 = phi*state0 + [0.5*tau^2*um ; tau*um]
__device__ __forceinline__ void transit(const double* s0, double* s1, const double* um, double tau) {
  const double h = __dmul_rn(0.5, __dmul_rn(tau, tau));
#pragma unroll
  for (int i = 0; i < 3; i++) {
    s1[i] = __dadd_rn(__dadd_rn(s0[i], __dmul_rn(tau, s0[i + 3])), __dmul_rn(h, um[i]));
    s1[i + 3] = __dadd_rn(s0[i + 3], __dmul_rn(tau, um[i]));
  }
}

// cubic / quartic (:521-585); arithmetic kept in source order, no contraction (-fmad=false)
__device__ int cubic(double a, double b, double c, double d, double* dts) {
  const double a2 = b / a, a1 = c / a, a0 = d / a;
  const double Q = (3 * a1 - a2 * a2) / 9;
  const double R = (9 * a1 * a2 - 27 * a0 - 2 * a2 * a2 * a2) / 54;
  const double D = Q * Q * Q + R * R;
  if (D > 0) {
    const double S = cbrt(R + sqrt(D));
    const double T = cbrt(R - sqrt(D));
    dts[0] = -a2 / 3 + (S + T);
    return 1;
  } else if (D == 0) {
    const double S = cbrt(R);
    dts[0] = -a2 / 3 + S + S;
    dts[1] = -a2 / 3 - S;
    return 2;
  }
  const double theta = acos(R / sqrt(-Q * Q * Q));
  dts[0] = 2 * sqrt(-Q) * cos(theta / 3) - a2 / 3;
  dts[1] = 2 * sqrt(-Q) * cos((theta + 2 * M_PI) / 3) - a2 / 3;
  dts[2] = 2 * sqrt(-Q) * cos((theta + 4 * M_PI) / 3) - a2 / 3;
  return 3;
}
__device__ int quartic(double a, double b, double c, double d, double e, double* dts) {
  const double a3 = b / a, a2 = c / a, a1 = d / a, a0 = e / a;
  double ys[3];
  cubic(1, -a2, a1 * a3 - 4 * a0, 4 * a2 * a0 - a1 * a1 - a3 * a3 * a0, ys);
  const double y1 = ys[0];
  const double r = a3 * a3 / 4 - a2 + y1;
  if (r < 0) return 0;
  const double R = sqrt(r);
  double D, E;
  if (R != 0) {
    D = sqrt(0.75 * a3 * a3 - R * R - 2 * a2 + 0.25 * (4 * a3 * a2 - 8 * a1 - a3 * a3 * a3) / R);
    E = sqrt(0.75 * a3 * a3 - R * R - 2 * a2 - 0.25 * (4 * a3 * a2 - 8 * a1 - a3 * a3 * a3) / R);
  } else {
    D = sqrt(0.75 * a3 * a3 - 2 * a2 + 2 * sqrt(y1 * y1 - 4 * a0));
    E = sqrt(0.75 * a3 * a3 - 2 * a2 - 2 * sqrt(y1 * y1 - 4 * a0));
  }
  int n = 0;
  if (!isnan(D)) { dts[n++] = -a3 / 4 + R / 2 + D / 2; dts[n++] = -a3 / 4 + R / 2 - D / 2; }
  if (!isnan(E)) { dts[n++] = -a3 / 4 - R / 2 + E / 2; dts[n++] = -a3 / 4 - R / 2 - E / 2; }
  return n;
}
// estimateHeuristic (:423-457)
__device__ double heuristic(const AstarParams& P, const double* x1, const double* x2, double* opt_t) {
  double dp[3], v0[3], v1[3], vs[3];
#pragma unroll
  for (int i = 0; i < 3; i++) { dp[i] = x2[i] - x1[i]; v0[i] = x1[i + 3]; v1[i] = x2[i + 3]; vs[i] = v0[i] + v1[i]; }
  const double c1 = -36 * (dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]);
  const double c2 = 24 * (vs[0] * dp[0] + vs[1] * dp[1] + vs[2] * dp[2]);
  const double c3 = -4 * ((v0[0] * v0[0] + v0[1] * v0[1] + v0[2] * v0[2]) +
                          (v0[0] * v1[0] + v0[1] * v1[1] + v0[2] * v1[2]) +
                          (v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]));
  double ts[5];
  int n = quartic(P.w_time, 0.0, c3, c2, c1, ts);
  const double v_max = P.max_vel * 0.5;
  const double linf = fmax(fmax(fabs(x1[0] - x2[0]), fabs(x1[1] - x2[1])), fabs(x1[2] - x2[2]));
  const double t_bar = linf / v_max;
  ts[n++] = t_bar;
  double cost = 100000000, t_d = t_bar;
  for (int i = 0; i < n; i++) {
    const double t = ts[i];
    if (t < t_bar) continue;
    const double c = -c1 / (3 * t * t * t) - c2 / (2 * t * t) - c3 / t + P.w_time * t;
    if (c < cost) { cost = c; t_d = t; }
  }
  if (opt_t) *opt_t = t_d;
  return 1.0 * (1 + P.tie_breaker) * cost;
}

template <typename C>
__device__ __forceinline__ int query_dt(const PopGeom& g, const uint8_t* ag, const float* pose,
                                        const double* p, double dt) {
  int t = static_cast<int>(floor(__ddiv_rn(dt, static_cast<double>(g.dt))));
  t = t > g.nt - 1 ? g.nt - 1 : t;
  return geom_query<C>(g, ag, pose[0], pose[1], pose[2], p[0], p[1], p[2], t);
}

// ---- open set: std::priority_queue<Node*, vector, NodeComparator{f1 > f2}> (libstdc++) --------
__device__ __forceinline__ void heap_push_up(int* heap, const double* f, int hole, int top, int value) {
  const double fv = f[value];
  int parent = (hole - 1) / 2;
  while (hole > top && f[heap[parent]] > fv) {
    heap[hole] = heap[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  heap[hole] = value;
}
__device__ void heap_push(int* heap, const double* f, int& size, int value) {
  heap[size] = value;
  size++;
  heap_push_up(heap, f, size - 1, 0, value);
}
__device__ void heap_pop(int* heap, const double* f, int& size) {
  if (size > 1) {
    const int last = size - 1;
    const int value = heap[last];
    heap[last] = heap[0];
    const int len = last;
    int hole = 0, child = 0;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (f[heap[child]] > f[heap[child - 1]]) child--;
      heap[hole] = heap[child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      heap[hole] = heap[child - 1];
      hole = child - 1;
    }
    heap_push_up(heap, f, hole, 0, value);
  }
  size--;
}

struct Prim {  // per-primitive scratch in shared memory
  double state[6];
  double g, f;
  int id[3];
  int node;      // existing node found in the hash (-1 none)
  int nstate;    // its state
  int valid;
};

template <typename C>
__global__ void __launch_bounds__(kThreads) k_astar(PopGeom g, AstarParams P, AstarDev D,
                                                    const uint8_t* __restrict__ grid,
                                                    const float* __restrict__ poses, int a0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* f = reinterpret_cast<double*>(smem_raw);            // [cap]
  int* heap = reinterpret_cast<int*>(f + P.cap);              // [cap]
  __shared__ Prim prim[kMaxPrim];
  __shared__ int s_tmp_node[kMaxPrim], s_tmp_owner[kMaxPrim], s_tmp_src[kMaxPrim];
  __shared__ int s_leader[kMaxPrim], s_slot[kMaxPrim];
  __shared__ int s_ntmp, s_heap_size, s_use, s_iter, s_ret, s_done, s_cur, s_init, s_shot, s_term;
  __shared__ double s_cur_state[6], s_cur_time, s_cur_g, s_end[6], s_start[9], s_time_origin, s_center[3];
  __shared__ int s_cur_index[3], s_cur_tidx, s_end_index[3];
  __shared__ unsigned s_gen;
  __shared__ int s_wrap;

  const int a = blockIdx.x + a0;
  const int tid = threadIdx.x;
  const uint8_t* ag = grid + a * g.agent_stride;
  const float* pose = poses + 3 * a;
  const size_t nb = static_cast<size_t>(a) * P.cap;
  double* n_state = D.state + nb * 6;
  double* n_input = D.input + nb * 3;
  double* n_time = D.time + nb;
  double* n_g = D.g + nb;
  int* n_parent = D.parent + nb;
  int* n_tidx = D.time_idx + nb;
  int* n_index = D.index + nb * 3;
  int* n_nstate = D.nstate + nb;
  unsigned long long* hkey = D.hkey + static_cast<size_t>(a) * D.hash_cap;
  int* hval = D.hval + static_cast<size_t>(a) * D.hash_cap;
  const unsigned mask = static_cast<unsigned>(P.hash_mask);

  auto pos_to_index = [&](const double* p, int* idx) {
#pragma unroll
    for (int i = 0; i < 3; i++) idx[i] = static_cast<int>(floor(__dmul_rn(__dsub_rn(p[i], s_center[i]), P.inv_res)));
  };
  auto time_to_index = [&](double t) {
    return static_cast<int>(floor(__dmul_rn(__dsub_rn(t, s_time_origin), P.inv_tres)));
  };

  int total_iter = 0, retried = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    // ---- reset() + search() prologue (:85-107, :122-165) ----
    if (tid == 0) {
      unsigned gen = D.gen[a] + 1;
      s_wrap = (gen & 0xffffu) == 0;    // 16-bit generation wrapped: stale entries could alias
      if (s_wrap) gen++;                // generation 0 marks never-used slots
      D.gen[a] = gen;
      s_gen = gen;
    }
    __syncthreads();
    if (s_wrap) {
      for (int i = tid; i < D.hash_cap; i += kThreads) hkey[i] = 0ull;
      __syncthreads();
    }
    if (tid == 0) {
      for (int i = 0; i < 9; i++) s_start[i] = D.d_start[a * 9 + i];
      for (int i = 0; i < 3; i++) { s_end[i] = D.d_goal[a * 3 + i]; s_end[i + 3] = 0.0; s_center[i] = static_cast<double>(pose[i]); }
      s_time_origin = D.d_tstart[a];
      s_heap_size = 0; s_use = 0; s_iter = 0; s_done = 0; s_shot = 0; s_term = -1;
      s_ret = POP_ASTAR_NO_PATH;
      s_init = attempt == 0 ? 1 : 0;
      int idx0[3];
      pos_to_index(s_start, idx0);
      pos_to_index(s_end, s_end_index);
      for (int i = 0; i < 6; i++) n_state[i] = s_start[i];
      for (int i = 0; i < 3; i++) { n_index[i] = idx0[i]; n_input[i] = 0.0; }
      n_parent[0] = -1;
      n_g[0] = 0.0;
      f[0] = __dmul_rn(P.lambda_heu, heuristic(P, s_start, s_end, nullptr));
      n_nstate[0] = IN_OPEN_SET;
      n_time[0] = s_time_origin;
      n_tidx[0] = time_to_index(s_time_origin);
      heap_push(heap, f, s_heap_size, 0);
      s_use = 1;
      hash_insert(hkey, hval, mask, s_gen, pack_key(s_gen, idx0[0], idx0[1], idx0[2], n_tidx[0]), 0);
    }
    __syncthreads();

    while (true) {
      // ---- pop / termination (:167-225) ----
      if (tid == 0) {
        if (s_heap_size == 0) {
          s_done = 1; s_ret = POP_ASTAR_NO_PATH; s_term = -1;
        } else {
          const int cur = heap[0];
          double cs[6];
          for (int i = 0; i < 6; i++) cs[i] = n_state[cur * 6 + i];
          const double ctime = n_time[cur];
          int cidx[3] = {n_index[cur * 3], n_index[cur * 3 + 1], n_index[cur * 3 + 2]};
          const double dx = cs[0] - s_start[0], dy = cs[1] - s_start[1], dz = cs[2] - s_start[2];
          const bool reach_horizon = sqrt(dx * dx + dy * dy + dz * dz) >= P.horizon;
          const bool near_end = abs(cidx[0] - s_end_index[0]) <= P.tolerance &&
                                abs(cidx[1] - s_end_index[1]) <= P.tolerance &&
                                abs(cidx[2] - s_end_index[2]) <= P.tolerance;
          const bool exceed_time = ctime >= P.max_tau;
          if (reach_horizon || near_end || exceed_time) {
            s_term = cur;
            if (near_end) {
              // computeShotTraj (:459-519); FakeRiskHybridAstar queries with the time ALONG the shot
              // (fake_risk_hybrid_a_star.cpp:521), RiskHybridAstar with the current slice (risk_hybrid_a_star.cpp:514)
              double t_d;
              heuristic(P, cs, s_end, &t_d);
              double coef[3][4];
              for (int i = 0; i < 3; i++) {
                const double p0 = cs[i], dpp = s_end[i] - p0, v0 = cs[i + 3], v1 = s_end[i + 3], dv = v1 - v0;
                coef[i][3] = 1.0 / 6.0 * (-12.0 / (t_d * t_d * t_d) * (dpp - v0 * t_d) + 6 / (t_d * t_d) * dv);
                coef[i][2] = 0.5 * (6.0 / (t_d * t_d) * (dpp - v0 * t_d) - 2 / t_d * dv);
                coef[i][1] = v0;
                coef[i][0] = p0;
              }
              const double t_delta = t_d / 10;
              // t_d == 0 (goal reached before the first expansion) makes the reference's loop spin forever and a
              // NaN t_d passes as a successful shot: both count as "no shot"
              bool ok = t_delta > 0.0 && isfinite(t_d);
              for (double time = t_delta; ok && time <= t_d; time += t_delta) {
                const double tp[4] = {1.0, time, time * time, pop_pow_int(time, 3)};
                double coord[3];
                for (int d = 0; d < 3; d++)
                  coord[d] = (coef[d][0] * tp[0] + coef[d][2] * tp[2]) + (coef[d][1] * tp[1] + coef[d][3] * tp[3]);
                if (query_dt<C>(g, ag, pose, coord, g.query_variant == POP_QUERY_FAKE ? time : 0.0) != 0) { ok = false; break; }
              }
              if (ok) s_shot = 1;
            }
            s_done = 1;
            if (reach_horizon) s_ret = s_shot ? POP_ASTAR_REACH_END : POP_ASTAR_REACH_HORIZON;
            else if (near_end) s_ret = s_shot ? POP_ASTAR_REACH_END : (n_parent[cur] >= 0 ? POP_ASTAR_NEAR_END : POP_ASTAR_NO_PATH);
            else s_ret = POP_ASTAR_REACH_HORIZON;
          } else {
            heap_pop(heap, f, s_heap_size);
            n_nstate[cur] = IN_CLOSE_SET;
            s_iter++;
            s_cur = cur;
            for (int i = 0; i < 6; i++) s_cur_state[i] = cs[i];
            s_cur_time = ctime;
            s_cur_g = n_g[cur];
            s_cur_tidx = n_tidx[cur];
            for (int i = 0; i < 3; i++) s_cur_index[i] = cidx[i];
            s_ntmp = 0;
          }
        }
      }
      __syncthreads();
      if (s_done) break;

      // ---- phase 1: evaluate primitives in parallel (:229-343) ----
      const int n_inputs = s_init ? 1 : kMaxPrim;
      if (tid < n_inputs) {
        Prim& pr = prim[tid];
        double um[3];
        if (s_init) {
          um[0] = s_start[6]; um[1] = s_start[7]; um[2] = s_start[8];
        } else {
          // ax, ay in {-A, -A/2, 0, A/2, A}, az in {-A/2, 0, A/2}; loop order ax -> ay -> az,
          // values accumulated like `ax += max_acc_ * res`
          const double step = __dmul_rn(P.max_acc, 0.5);
          const int ia = tid / 15, ib = (tid / 3) % 5, ic = tid % 3;
          double ax = -P.max_acc, ay = -P.max_acc, az = __dmul_rn(-0.5, P.max_acc);
          for (int q = 0; q < ia; q++) ax = __dadd_rn(ax, step);
          for (int q = 0; q < ib; q++) ay = __dadd_rn(ay, step);
          for (int q = 0; q < ic; q++) az = __dadd_rn(az, step);
          um[0] = ax; um[1] = ay; um[2] = az;
        }
        const double tau = P.tau;
        double ps[6];
        transit(s_cur_state, ps, um, tau);
        const double pro_t = __dadd_rn(s_cur_time, tau);
        int pid[3];
        pos_to_index(ps, pid);
        const int pro_t_id = time_to_index(pro_t);
        int valid = 1;
        const int node = hash_find(hkey, hval, mask, s_gen, pack_key(s_gen, pid[0], pid[1], pid[2], pro_t_id));
        int nst = NOT_EXPAND;
        if (node >= 0) {
          nst = n_nstate[node];
          if (nst == IN_CLOSE_SET) valid = 0;
        }
        if (valid && (fabs(ps[3]) > P.max_vel || fabs(ps[4]) > P.max_vel || fabs(ps[5]) > P.max_vel)) valid = 0;
        if (valid && pid[0] == s_cur_index[0] && pid[1] == s_cur_index[1] && pid[2] == s_cur_index[2] &&
            pro_t_id - s_cur_tidx == 0)
          valid = 0;
        if (valid) {
          for (int k = 1; k <= P.check_num; ++k) {
            const double dt = __ddiv_rn(__dmul_rn(tau, static_cast<double>(k)), static_cast<double>(P.check_num));
            double xt[6];
            transit(s_cur_state, xt, um, dt);
            const double t = __dadd_rn(s_cur_time, dt);
            if (query_dt<C>(g, ag, pose, xt, t) != 0) { valid = 0; break; }
          }
        }
        if (valid) {
          const double usq = um[0] * um[0] + um[1] * um[1] + um[2] * um[2];
          pr.g = (usq + P.w_time) * tau + s_cur_g;
          pr.f = pr.g + P.lambda_heu * heuristic(P, ps, s_end, nullptr);
        }
        for (int i = 0; i < 6; i++) pr.state[i] = ps[i];
        for (int i = 0; i < 3; i++) pr.id[i] = pid[i];
        pr.node = node;
        pr.nstate = nst;
        pr.valid = valid;
      }
      __syncthreads();
      // same-voxel duplicates among this expansion's new nodes (:345-360): the reference scans the list of nodes
      // created so far for every primitive; here every primitive finds, in parallel, the first earlier valid
      // primitive of its voxel (the one that will have created the node), so the ordered commit below is O(n)
      if (tid < n_inputs) {
        int leader = tid;
        const Prim& pr = prim[tid];
        if (pr.valid && pr.node < 0) {
          for (int j = 0; j < tid; j++) {
            const Prim& q = prim[j];
            if (q.valid && q.id[0] == pr.id[0] && q.id[1] == pr.id[1] && q.id[2] == pr.id[2]) { leader = j; break; }
          }
        }
        s_leader[tid] = leader;
      }
      __syncthreads();

      // ---- phase 2: ordered commit (:345-419) ----
      if (tid == 0) {
        const int pro_t_id = time_to_index(__dadd_rn(s_cur_time, P.tau));
        for (int i = 0; i < n_inputs; i++) {
          Prim& pr = prim[i];
          if (!pr.valid) continue;
          if (pr.node < 0 && s_leader[i] != i) {   // an earlier primitive of this expansion created the node
            const int j = s_slot[s_leader[i]];
            const int en = s_tmp_node[j];
            if (pr.f < f[en]) { f[en] = pr.f; s_tmp_owner[j] = i; }
            continue;
          }
          if (pr.node < 0) {
            const int id = s_use;
            f[id] = pr.f;
            heap_push(heap, f, s_heap_size, id);
            s_tmp_node[s_ntmp] = id; s_tmp_owner[s_ntmp] = i; s_tmp_src[s_ntmp] = i;
            s_slot[i] = s_ntmp;
            s_ntmp++;
            s_use++;
            if (s_use == P.cap) { s_done = 1; s_ret = POP_ASTAR_NO_PATH; s_term = -1; break; }
          } else if (n_nstate[pr.node] == IN_OPEN_SET) {
            const int nd = pr.node;
            if (pr.g < n_g[nd]) {
              for (int q = 0; q < 6; q++) n_state[nd * 6 + q] = pr.state[q];
              f[nd] = pr.f;
              n_g[nd] = pr.g;
              const double step = __dmul_rn(P.max_acc, 0.5);
              double ax = -P.max_acc, ay = -P.max_acc, az = __dmul_rn(-0.5, P.max_acc);
              if (s_init) { ax = s_start[6]; ay = s_start[7]; az = s_start[8]; }
              else {
                for (int q = 0; q < i / 15; q++) ax = __dadd_rn(ax, step);
                for (int q = 0; q < (i / 3) % 5; q++) ay = __dadd_rn(ay, step);
                for (int q = 0; q < i % 3; q++) az = __dadd_rn(az, step);
              }
              n_input[nd * 3] = ax; n_input[nd * 3 + 1] = ay; n_input[nd * 3 + 2] = az;
              n_parent[nd] = s_cur;
              n_time[nd] = __dadd_rn(s_cur_time, P.tau);
            }
          } else {
            s_done = 1; s_ret = POP_ASTAR_SEARCH_ERR; s_term = -1; break;
          }
        }
        (void)pro_t_id;
        if (s_init) s_init = 0;
      }
      __syncthreads();

      // ---- phase 3: publish new nodes (payload + hash) in parallel ----
      if (tid < s_ntmp) {
        const int id = s_tmp_node[tid], owner = s_tmp_owner[tid], src = s_tmp_src[tid];
        const Prim& pr = prim[owner];
        for (int q = 0; q < 6; q++) n_state[id * 6 + q] = pr.state[q];
        double um[3];
        if (n_inputs == 1) { um[0] = s_start[6]; um[1] = s_start[7]; um[2] = s_start[8]; }
        else {
          const double step = __dmul_rn(P.max_acc, 0.5);
          um[0] = -P.max_acc; um[1] = -P.max_acc; um[2] = __dmul_rn(-0.5, P.max_acc);
          for (int q = 0; q < owner / 15; q++) um[0] = __dadd_rn(um[0], step);
          for (int q = 0; q < (owner / 3) % 5; q++) um[1] = __dadd_rn(um[1], step);
          for (int q = 0; q < owner % 3; q++) um[2] = __dadd_rn(um[2], step);
        }
        for (int q = 0; q < 3; q++) { n_input[id * 3 + q] = um[q]; n_index[id * 3 + q] = prim[src].id[q]; }
        const double t = __dadd_rn(s_cur_time, P.tau);
        n_time[id] = t;
        n_tidx[id] = time_to_index(t);
        n_g[id] = pr.g;
        n_parent[id] = s_cur;
        n_nstate[id] = IN_OPEN_SET;
        // expanded_nodes_.insert(pro_id, pro_node->time, ...): the double time is truncated to int
        hash_insert(hkey, hval, mask, s_gen,
                    pack_key(s_gen, prim[src].id[0], prim[src].id[1], prim[src].id[2], static_cast<int>(t)), id);
      }
      __threadfence_block();
      __syncthreads();
      if (s_done) break;
    }
    total_iter += s_iter;
    __syncthreads();
    if (s_ret != POP_ASTAR_NO_PATH || attempt == 1) break;
    retried = 1;
    __syncthreads();
  }

  // ---- retrievePath + getPathWithVel(corridor_tau) (:826-836, :663-694) ----
  if (tid == 0) {
    int len = 0;
    double* out = D.d_path + static_cast<size_t>(a) * POP_MAX_PATH * 6;
    if (s_ret != POP_ASTAR_NO_PATH && s_ret != POP_ASTAR_SEARCH_ERR && s_term >= 0) {
      double tmp[POP_MAX_PATH * 2][6];
      int n = 0;
      int node = s_term;
      const double delta_t = P.corridor_tau;
      double t_sample = delta_t;
      bool overflow = false;
      for (int q = 0; q < 6; q++) tmp[n][q] = n_state[node * 6 + q];
      n++;
      while (n_parent[node] >= 0 && !overflow) {
        const int par = n_parent[node];
        double ut[3] = {n_input[node * 3], n_input[node * 3 + 1], n_input[node * 3 + 2]};
        double x0[6];
        for (int q = 0; q < 6; q++) x0[q] = n_state[par * 6 + q];
        double t_node = P.tau;  // node->duration
        while (true) {
          if (t_sample > t_node) {
            node = par;
            t_sample = __dsub_rn(t_sample, t_node);
            break;
          }
          t_node = __dsub_rn(t_node, t_sample);
          if (n >= POP_MAX_PATH * 2) { overflow = true; break; }
          transit(x0, tmp[n], ut, t_node);
          n++;
          t_sample = delta_t;
        }
      }
      if (overflow || n > POP_MAX_PATH) {
        len = -1;
      } else {
        len = n;
        for (int i = 0; i < n; i++)
          for (int q = 0; q < 6; q++) out[i * 6 + q] = tmp[n - 1 - i][q];
      }
    }
    D.d_len[a] = len;
    D.d_ret[a] = s_ret;
    D.d_stats[a * 4 + 0] = s_iter;
    D.d_stats[a * 4 + 1] = s_use;
    D.d_stats[a * 4 + 2] = retried;
    D.d_stats[a * 4 + 3] = total_iter;
  }
}

}  // namespace

// ---- host side -----------------------------------------------------------------------------------
int popk_astar_alloc(pop_handle* h) {
  if (h->astar) return POP_OK;
  const pop_config& c = h->cfg;
  if (c.allocate_num <= 1 || c.allocate_num > kSmemNodes)
    return pop_fail(h, POP_E_UNSUPPORTED, "search/allocate_num=%d outside (1, %d] supported by the shared-memory open set", c.allocate_num, kSmemNodes);
  AstarDev* d = new AstarDev();
  d->cap = c.allocate_num;
  int hc = 1;
  while (hc < 2 * d->cap) hc <<= 1;
  d->hash_cap = hc;
  d->n_agents = c.n_agents_local;
  const size_t n = static_cast<size_t>(d->n_agents) * d->cap;
  const size_t nh = static_cast<size_t>(d->n_agents) * hc;
  bool ok = true;
  ok &= cudaMalloc(&d->state, n * 6 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->input, n * 3 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->time, n * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->g, n * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->parent, n * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->time_idx, n * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->index, n * 3 * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->nstate, n * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->hkey, nh * sizeof(unsigned long long)) == cudaSuccess;
  ok &= cudaMalloc(&d->hval, nh * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->gen, d->n_agents * sizeof(unsigned)) == cudaSuccess;
  ok &= cudaMalloc(&d->d_start, d->n_agents * 9 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->d_goal, d->n_agents * 3 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->d_tstart, d->n_agents * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->d_ret, d->n_agents * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->d_path, static_cast<size_t>(d->n_agents) * POP_MAX_PATH * 6 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->d_len, d->n_agents * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->d_stats, d->n_agents * 4 * sizeof(int)) == cudaSuccess;
  // the handle keeps the workspace only if every allocation and attribute call succeeded: a failed attempt frees
  // what it got and leaves h->astar null, so the next call tries again instead of launching on null pointers
  h->astar = d;
  auto finish = [&]() -> int {
    if (!ok) { cudaGetLastError(); return pop_fail(h, POP_E_CUDA, "A* pool allocation failed"); }
    POP_CUDA(h, cudaMemsetAsync(d->hkey, 0, nh * sizeof(unsigned long long), h->stream));
    POP_CUDA(h, cudaMemsetAsync(d->gen, 0, d->n_agents * sizeof(unsigned), h->stream));
    POP_CUDA(h, cudaMemsetAsync(d->d_path, 0, static_cast<size_t>(d->n_agents) * POP_MAX_PATH * 6 * sizeof(double), h->stream));
    const size_t smem = static_cast<size_t>(d->cap) * 12 + 16;
    POP_CUDA(h, cudaFuncSetAttribute(k_astar<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    POP_CUDA(h, cudaFuncSetAttribute(k_astar<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    return POP_OK;
  };
  const int rc = finish();
  if (rc != POP_OK) { popk_astar_free(h); return rc; }
  return POP_OK;
}

void popk_astar_free(pop_handle* h) {
  AstarDev* d = h->astar;
  if (!d) return;
  void* ptrs[] = {d->state, d->input, d->time, d->g, d->parent, d->time_idx, d->index, d->nstate, d->hkey,
                  d->hval, d->gen, d->d_start, d->d_goal, d->d_tstart, d->d_ret, d->d_path, d->d_len, d->d_stats};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete d;
  h->astar = nullptr;
}

// device-resident search: inputs already in astar->d_start/d_goal/d_tstart
int popk_astar_run(pop_handle* h, cudaStream_t st, int a0, int na) {
  AstarDev* d = h->astar;
  const pop_config& c = h->cfg;
  AstarParams P;
  P.max_tau = c.max_tau; P.max_vel = c.max_vel; P.max_acc = c.max_acc; P.w_time = c.w_time;
  P.horizon = c.horizon; P.lambda_heu = c.lambda_heu;
  P.inv_res = 1.0 / c.resolution_astar;
  P.inv_tres = 1.0 / c.time_resolution_astar;
  P.tau = c.time_resolution_astar;
  P.tie_breaker = 1.0 + 1.0 / 10000;
  P.corridor_tau = c.corridor_tau;
  P.cap = d->cap; P.check_num = c.check_num; P.tolerance = c.tolerance; P.hash_mask = d->hash_cap - 1;
  const size_t smem = static_cast<size_t>(d->cap) * 12 + 16;
  pop_prof_begin(h, POP_STAGE_ASTAR);
  if (c.cell_dtype == POP_CELL_U8)
    k_astar<uint8_t><<<na, kThreads, smem, st>>>(h->geom, P, *d, h->d_grid, h->d_poses, a0);
  else
    k_astar<float><<<na, kThreads, smem, st>>>(h->geom, P, *d, h->d_grid, h->d_poses, a0);
  POP_LAUNCH_CHECK(h);
  pop_prof_end(h, POP_STAGE_ASTAR, 1);
  return POP_OK;
}

double* popk_astar_path_dev(pop_handle* h) { return h->astar->d_path; }
int* popk_astar_len_dev(pop_handle* h) { return h->astar->d_len; }
int* popk_astar_ret_dev(pop_handle* h) { return h->astar->d_ret; }
double* popk_astar_start_dev(pop_handle* h) { return h->astar->d_start; }
double* popk_astar_goal_dev(pop_handle* h) { return h->astar->d_goal; }
double* popk_astar_tstart_dev(pop_handle* h) { return h->astar->d_tstart; }

extern "C" int pop_astar_search_batch(pop_handle* h, const double* start, const double* goal,
                                      const double* t_start, int32_t* out_ret, double* out_path,
                                      int32_t* out_len, int32_t* out_stats) {
  if (!h || !start || !goal || !t_start || !out_ret || !out_path || !out_len) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->map_ready) return pop_fail(h, POP_E_STATE, "A* before the first map update");
  int rc = popk_astar_alloc(h);
  if (rc) return rc;
  AstarDev* d = h->astar;
  const int n = d->n_agents;
  POP_CUDA(h, cudaMemcpyAsync(d->d_start, start, sizeof(double) * 9 * n, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d->d_goal, goal, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d->d_tstart, t_start, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  rc = popk_astar_run(h, h->stream, 0, h->astar->n_agents);
  if (rc) return rc;
  POP_CUDA(h, cudaMemcpyAsync(out_ret, d->d_ret, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(out_len, d->d_len, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(out_path, d->d_path, sizeof(double) * 6 * POP_MAX_PATH * n, cudaMemcpyDeviceToHost, h->stream));
  if (out_stats) POP_CUDA(h, cudaMemcpyAsync(out_stats, d->d_stats, sizeof(int) * 4 * n, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < n; i++)
    if (out_len[i] < 0) return pop_fail(h, POP_E_CAPACITY, "agent %d: A* path longer than POP_MAX_PATH=%d samples", i, POP_MAX_PATH);
  return POP_OK;
}

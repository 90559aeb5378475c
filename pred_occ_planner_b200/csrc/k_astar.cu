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

constexpr int kRoleThreads = 96;            // three warps cover the 75 primitives
constexpr int kThreads = 3 * kRoleThreads;  // lookup, gather and heuristic roles
constexpr int kMaxPrim = 75;
constexpr int kSmemNodes = 16384;  // f (double) + heap slot (16-bit node id) = 10 B per node in shared memory: two
                                   // searches of the default 10 000-node pool share an SM
using HeapIdx = unsigned short;
constexpr int kSmemBytesPerNode = 8 + sizeof(HeapIdx);
static_assert(kSmemNodes <= 65536, "node ids must fit the heap slot");

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
// Linear probing, two slots per round: keys and values of both are loaded before any is looked at, so the usual
// lookup (hit or miss within two slots) costs one memory round trip instead of two or three dependent ones.
__device__ int hash_find(const unsigned long long* hkey, const int* hval, unsigned mask, unsigned gen,
                         unsigned long long key) {
  unsigned pos = hash_slot(key, mask);
  for (unsigned probe = 0; probe <= mask; probe += 2) {
    const unsigned pos1 = (pos + 1) & mask;
    const unsigned long long k0 = hkey[pos], k1 = hkey[pos1];
    const int v0 = hval[pos], v1 = hval[pos1];
    if ((k0 >> 48) != (gen & 0xffffu)) return -1;
    if (k0 == key) return v0;
    if ((k1 >> 48) != (gen & 0xffffu)) return -1;
    if (k1 == key) return v1;
    pos = (pos + 2) & mask;
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
  // at most 4 roots + t_bar; the candidates' costs do not depend on one another, so they are evaluated side by side
  // (three divisions each, the long pole of this function) and only the selection walks them in order
  double cs[5];
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const double t = i < n ? ts[i] : 1.0;
    cs[i] = -c1 / (3 * t * t * t) - c2 / (2 * t * t) - c3 / t + P.w_time * t;
  }
#pragma unroll
  for (int i = 0; i < 5; i++) {
    if (i >= n || ts[i] < t_bar) continue;
    if (cs[i] < cost) { cost = cs[i]; t_d = ts[i]; }
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
__device__ __forceinline__ void heap_push_up(HeapIdx* heap, const double* f, int hole, int top, int value) {
  const double fv = f[value];
  int parent = (hole - 1) / 2;
  while (hole > top && f[heap[parent]] > fv) {
    heap[hole] = heap[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  heap[hole] = static_cast<HeapIdx>(value);
}
__device__ void heap_push(HeapIdx* heap, const double* f, int& size, int value) {
  heap[size] = static_cast<HeapIdx>(value);
  size++;
  heap_push_up(heap, f, size - 1, 0, value);
}
__device__ void heap_pop(HeapIdx* heap, const double* f, int& size) {
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
  int id[3];
  double node_g; // g of the existing node
  int node;      // existing node found in the hash (-1 none)
  int nstate;    // its state
  int valid;
};
enum { EV_NONE = 0, EV_NEW, EV_LOWER, EV_UPD, EV_ERR };

template <typename C>
__global__ void __launch_bounds__(kThreads, 1) k_astar(PopGeom g, AstarParams P, AstarDev D,
                                                    const uint8_t* __restrict__ grid,
                                                    const float* __restrict__ poses, int a0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* f = reinterpret_cast<double*>(smem_raw);            // [cap]
  HeapIdx* heap = reinterpret_cast<HeapIdx*>(f + P.cap);      // [cap]
  __shared__ Prim prim[kMaxPrim];
  __shared__ unsigned long long s_key[kMaxPrim];
  __shared__ double s_score[kMaxPrim], s_pg[kMaxPrim], s_pf[kMaxPrim];
  __shared__ int s_leader[kMaxPrim], s_newid[kMaxPrim], s_ev[kMaxPrim], s_owns[kMaxPrim], s_bad[kMaxPrim];
  __shared__ unsigned s_evmask[3];
  __shared__ int s_heap_size, s_use, s_iter, s_ret, s_done, s_cur, s_init, s_shot, s_term;
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

#ifdef POP_ASTAR_PROF   // cycles per section of an expansion, thread 0's clock (scripts/gpu_astar_phase.sh)
  long long prof_acc[6] = {0, 0, 0, 0, 0, 0}, prof_t = clock64();
  __shared__ long long s_prof_role[3];
  if (tid < 3) s_prof_role[tid] = 0;
#define POP_ASTAR_TICK(k) do { if (tid == 0) { const long long c_ = clock64(); prof_acc[k] += c_ - prof_t; prof_t = c_; } } while (0)
#else
#define POP_ASTAR_TICK(k) do { } while (0)
#endif
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
      // The sift-down of pop() needs only the heap; the termination tests need only the top node's payload.  Two
      // threads of different warps do one each: if the node turns out to be terminal the search ends and the popped
      // heap is never looked at again.
      POP_ASTAR_TICK(0);
      const int top_node = s_heap_size > 0 ? heap[0] : -1;
      __syncthreads();
      if (tid == 0 && top_node >= 0) heap_pop(heap, f, s_heap_size);
      if (tid == 32) {
        if (top_node < 0) {
          s_done = 1; s_ret = POP_ASTAR_NO_PATH; s_term = -1;
        } else {
          const int cur = top_node;
          double cs[6];
          for (int i = 0; i < 6; i++) cs[i] = n_state[cur * 6 + i];
          const double ctime = n_time[cur];
          int cidx[3] = {n_index[cur * 3], n_index[cur * 3 + 1], n_index[cur * 3 + 2]};
          const double cg = n_g[cur];
          const int ctidx = n_tidx[cur];
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
            n_nstate[cur] = IN_CLOSE_SET;
            s_iter++;
            s_cur = cur;
            for (int i = 0; i < 6; i++) s_cur_state[i] = cs[i];
            s_cur_time = ctime;
            s_cur_g = cg;
            s_cur_tidx = ctidx;
            for (int i = 0; i < 3; i++) s_cur_index[i] = cidx[i];
          }
        }
      }
      __threadfence_block();
      __syncthreads();
      if (s_done) break;
      POP_ASTAR_TICK(1);

      // ---- evaluate the primitives in parallel (:229-343) ----
      // Three roles, one warp group each, all starting from the popped node's state: the lookups (closed set /
      // velocity / same voxel pruning), the SOGM region-risk gathers and the quartic heuristics do not need one
      // another's results, only the commit needs all three.  Gathers and heuristics of primitives the lookup prunes
      // are wasted work on otherwise idle warps; the expansion's latency is that of the slowest role, not their sum.
      const int n_inputs = s_init ? 1 : kMaxPrim;
      auto prim_input = [&](int i, double* um) {
        if (n_inputs == 1) {
          um[0] = s_start[6]; um[1] = s_start[7]; um[2] = s_start[8];
        } else {
          // ax, ay in {-A, -A/2, 0, A/2, A}, az in {-A/2, 0, A/2}; loop order ax -> ay -> az,
          // values accumulated like `ax += max_acc_ * res`
          const double step = __dmul_rn(P.max_acc, 0.5);
          const int ia = i / 15, ib = (i / 3) % 5, ic = i % 3;
          double ax = -P.max_acc, ay = -P.max_acc, az = __dmul_rn(-0.5, P.max_acc);
          for (int q = 0; q < ia; q++) ax = __dadd_rn(ax, step);
          for (int q = 0; q < ib; q++) ay = __dadd_rn(ay, step);
          for (int q = 0; q < ic; q++) az = __dadd_rn(az, step);
          um[0] = ax; um[1] = ay; um[2] = az;
        }
      };
      const int role = tid / kRoleThreads, ri = tid % kRoleThreads;
#ifdef POP_ASTAR_PROF
      const long long role_t0 = clock64();
#endif
      if (ri < n_inputs) {
        double um[3], ps[6];
        prim_input(ri, um);
        transit(s_cur_state, ps, um, P.tau);
        const bool vel_ok = !(fabs(ps[3]) > P.max_vel || fabs(ps[4]) > P.max_vel || fabs(ps[5]) > P.max_vel);
        if (role == 0) {
          Prim& pr = prim[ri];
          const double pro_t = __dadd_rn(s_cur_time, P.tau);
          int pid[3];
          pos_to_index(ps, pid);
          const int pro_t_id = time_to_index(pro_t);
          int valid = 1;
          const int node = hash_find(hkey, hval, mask, s_gen, pack_key(s_gen, pid[0], pid[1], pid[2], pro_t_id));
          int nst = NOT_EXPAND;
          double node_g = 0.0;
          if (node >= 0) {
            nst = n_nstate[node];
            node_g = n_g[node];
            if (nst == IN_CLOSE_SET) valid = 0;
          }
          if (valid && !vel_ok) valid = 0;
          if (valid && pid[0] == s_cur_index[0] && pid[1] == s_cur_index[1] && pid[2] == s_cur_index[2] &&
              pro_t_id - s_cur_tidx == 0)
            valid = 0;
          for (int i = 0; i < 6; i++) pr.state[i] = ps[i];
          for (int i = 0; i < 3; i++) pr.id[i] = pid[i];
          pr.node = node;
          pr.nstate = nst;
          pr.node_g = node_g;
          pr.valid = valid;
          s_newid[ri] = -1;
        } else if (role == 1) {
          int bad = 0;
          if (vel_ok) {
            for (int k = 1; k <= P.check_num && !bad; ++k) {
              const double dt = __ddiv_rn(__dmul_rn(P.tau, static_cast<double>(k)), static_cast<double>(P.check_num));
              double xt[6];
              transit(s_cur_state, xt, um, dt);
              bad = query_dt<C>(g, ag, pose, xt, __dadd_rn(s_cur_time, dt)) != 0;
            }
          }
          s_bad[ri] = bad;
        } else if (vel_ok) {
          const double usq = um[0] * um[0] + um[1] * um[1] + um[2] * um[2];
          const double gg = (usq + P.w_time) * P.tau + s_cur_g;
          s_pg[ri] = gg;
          s_pf[ri] = gg + P.lambda_heu * heuristic(P, ps, s_end, nullptr);
        }
      }
#ifdef POP_ASTAR_PROF
      if (ri == 37) s_prof_role[role] += clock64() - role_t0;   // a middle primitive of each role
#endif
      __syncthreads();
      POP_ASTAR_TICK(2);
      // ---- ordered commit (:345-419), resolved in parallel ----
      // All valid primitives that land in one voxel address one node (the time index is the same for all of them):
      // a new one, created by the first of them ("leader", the same-voxel scan of :345-360), or one already open.  The
      // reference walks the primitives in order; what that walk does to a node depends only on the primitives of
      // its voxel, so every primitive finds out by itself, from the others' voxel and score, whether the walk
      //   NEW    creates the node with its f (it is the leader),
      //   LOWER  lowers the new node's f in place (its f is below that of every earlier primitive of the voxel),
      //   UPD    re-parents the open node (its g is below the node's and every earlier primitive's of the voxel),
      // and whether it ends up owning the node's payload (lowest score of the voxel, first among equals).  Only the
      // heap order still needs the walk: thread 0 replays those events, in primitive order, on f[] and the heap.
      if (tid < kMaxPrim) {   // slots past n_inputs (the first expansion has one input) get a key nothing matches
        Prim& pr = prim[tid];
        const int valid = tid < n_inputs && pr.valid && !s_bad[tid];
        pr.valid = valid;
        s_key[tid] = valid ? pack_key(0, pr.id[0], pr.id[1], pr.id[2], 0) : (~0ull - static_cast<unsigned long long>(tid));
        s_score[tid] = pr.node < 0 ? s_pf[tid] : s_pg[tid];
      }
      __syncthreads();
      int ev = EV_NONE;
      if (tid < n_inputs) {
        const Prim& pr = prim[tid];
        int owns = 0;
        if (pr.valid) {
          const unsigned long long key = s_key[tid];
          const double mine = s_score[tid];
          int leader = tid, beaten = 0;
          double prev_min = pr.node < 0 ? INFINITY : pr.node_g;
          // the key compares are branch-free and unrolled (their loads do not wait for one another); only the few
          // primitives that share the voxel are then looked at
#pragma unroll
          for (int c = 0; c < 3; c++) {
            unsigned m = 0u;
#pragma unroll
            for (int k = 0; k < 25; k++) m |= (s_key[c * 25 + k] == key ? 1u : 0u) << k;
            if (tid / 25 == c) m &= ~(1u << (tid % 25));
            while (m) {
              const int j = c * 25 + __ffs(m) - 1;
              m &= m - 1;
              const double sj = s_score[j];
              if (j < tid) { leader = min(leader, j); prev_min = fmin(prev_min, sj); }
              if (sj < mine || (sj == mine && j < tid)) beaten = 1;
            }
          }
          s_leader[tid] = leader;
          if (pr.node < 0) {
            ev = leader == tid ? EV_NEW : (mine < prev_min ? EV_LOWER : EV_NONE);
            owns = !beaten;
          } else if (pr.nstate == IN_OPEN_SET) {
            ev = mine < prev_min ? EV_UPD : EV_NONE;
            owns = !beaten && mine < pr.node_g;
          } else {
            ev = EV_ERR;
          }
        }
        s_owns[tid] = owns;
        s_ev[tid] = ev;
      }
      if (tid < 96) {   // warps 0..2 cover the 75 primitives
        const unsigned m = __ballot_sync(0xffffffffu, ev != EV_NONE);
        if ((tid & 31) == 0) s_evmask[tid >> 5] = m;
      }
      __syncthreads();
      POP_ASTAR_TICK(3);
      if (tid == 0) {
        bool stop = false;
        for (int w = 0; w < 3 && !stop; w++) {
          unsigned m = s_evmask[w];
          while (m) {
            const int i = w * 32 + __ffs(m) - 1;
            m &= m - 1;
            const int e = s_ev[i];
            if (e == EV_NEW) {
              const int id = s_use;
              f[id] = s_pf[i];
              heap_push(heap, f, s_heap_size, id);
              s_newid[i] = id;
              s_use++;
              if (s_use == P.cap) { s_done = 1; s_ret = POP_ASTAR_NO_PATH; s_term = -1; stop = true; break; }
            } else if (e == EV_LOWER) {
              f[s_newid[s_leader[i]]] = s_pf[i];
            } else if (e == EV_UPD) {
              f[prim[i].node] = s_pf[i];
            } else {
              s_done = 1; s_ret = POP_ASTAR_SEARCH_ERR; s_term = -1; stop = true; break;
            }
          }
        }
        if (s_init) s_init = 0;
      }
      __syncthreads();
      POP_ASTAR_TICK(4);

      // ---- publish the payloads (new nodes and re-parented open nodes) and the hash entries in parallel ----
      if (tid < n_inputs && s_owns[tid]) {
        const Prim& pr = prim[tid];
        const int id = pr.node < 0 ? s_newid[s_leader[tid]] : pr.node;
        if (id >= 0) {   // -1: the walk stopped (pool exhausted) before the leader's turn
          for (int q = 0; q < 6; q++) n_state[id * 6 + q] = pr.state[q];
          double um[3];
          prim_input(tid, um);
          for (int q = 0; q < 3; q++) n_input[id * 3 + q] = um[q];
          const double t = __dadd_rn(s_cur_time, P.tau);
          n_time[id] = t;
          n_g[id] = s_pg[tid];
          n_parent[id] = s_cur;
          if (pr.node < 0) {
            for (int q = 0; q < 3; q++) n_index[id * 3 + q] = pr.id[q];
            n_tidx[id] = time_to_index(t);
            n_nstate[id] = IN_OPEN_SET;
            // expanded_nodes_.insert(pro_id, pro_node->time, ...): the double time is truncated to int
            hash_insert(hkey, hval, mask, s_gen, pack_key(s_gen, pr.id[0], pr.id[1], pr.id[2], static_cast<int>(t)), id);
          }
        }
      }
      __threadfence_block();
      __syncthreads();
      POP_ASTAR_TICK(5);
      if (s_done) break;
    }
    total_iter += s_iter;
    __syncthreads();
    if (s_ret != POP_ASTAR_NO_PATH || attempt == 1) break;
    retried = 1;
    __syncthreads();
  }

#ifdef POP_ASTAR_PROF
  if (tid == 0 && total_iter >= 100)
    printf("astar_prof agent %d expansions %d nodes %d cycles/expansion: pop %lld evaluate %lld (lookup %lld gather %lld heuristic %lld) resolve %lld commit %lld publish %lld\n", a,
           total_iter, s_use, prof_acc[1] / total_iter, prof_acc[2] / total_iter, s_prof_role[0] / total_iter, s_prof_role[1] / total_iter, s_prof_role[2] / total_iter, prof_acc[3] / total_iter,
           prof_acc[4] / total_iter, prof_acc[5] / total_iter);
#endif
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
#ifndef POP_ASTAR_DEVICE_ONLY   // scripts/ubench/heur_lat.cu includes the device code alone
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
    const size_t smem = static_cast<size_t>(d->cap) * kSmemBytesPerNode + 16;
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
  const size_t smem = static_cast<size_t>(d->cap) * kSmemBytesPerNode + 16;
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
#endif  // POP_ASTAR_DEVICE_ONLY

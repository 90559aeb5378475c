// pop_corridor_finish.cuh — the sequential tail of the corridor stage for ONE agent: cut at the first invalid
// corridor, adjacent-intersection LPs, goal reachability (FakeBaselinePlanner::replan, baseline_fake.cpp:356-411;
// BaselinePlanner::replan, baseline.cpp:353-403).  Plain C++ over flat arrays (no thread indices, no intrinsics), so
// the same source also compiles for the host: tests/csrc/corridor_finish_host.cpp builds it with g++ and
// tests/test_device_logic_on_host.py checks it on the CPU (the only way to exercise device logic without a GPU).
//
// The two planners of the reference differ here, selected by `variant` (= pop_config.query_variant, the kind of map
// the planner sits on):
//   POP_QUERY_FAKE      (baseline_fake.cpp)  non-intersecting corridors i, i+1 (i >= 2): keep 0..i      (:374)
//                                            fail when no corridor is left                               (:384)
//                                            goal search: from the last corridor backwards               (:394-407)
//   POP_QUERY_RISKBASE  (baseline.cpp)       non-intersecting corridors i, i+1 (i >= 2): keep 0..i-1    (:371)
//                                            fail when at most one corridor is left                      (:381)
//                                            the last corridor is tested once first (:389) — a test that moves the
//                                            goal into the corridor when it fails — and only then searched backwards
#pragma once
#include <math.h>

#include "../../include/pop.h"
#include "pop_lp.cuh"

__device__ __forceinline__ unsigned long long lp_seed(unsigned long long base, int segment, int kind) {
  unsigned long long z = base * 0x9e3779b97f4a7c15ull + static_cast<unsigned long long>(segment) * 16 + kind + 1;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// checkGoalReachability (baseline_fake.cpp:143-182 = baseline.cpp:143-181): true when the goal satisfies every face of
// corridor h (m rows); otherwise the goal is replaced by the midpoint of the corridor's extreme points along the
// start -> goal direction and false is returned.  kind0 / kind0 + 1 pick the LP shuffle streams.
__device__ inline bool corridor_goal_reachable(const double* h, int m, const double* sp, double* goal_pos,
                                               unsigned long long seed_base, int it, int kind0, double* ab, double* lpw,
                                               int* scri) {
  if (m <= 0) return true;
  double mx = -INFINITY;
  for (int k = 0; k < m; k++) {
    const double v = ((h[k * 4] * goal_pos[0] + h[k * 4 + 1] * goal_pos[1]) + h[k * 4 + 2] * goal_pos[2]) + h[k * 4 + 3] * 1.0;
    mx = fmax(mx, v);
  }
  if (mx <= 0) return true;
  for (int k = 0; k < m; k++) { for (int q = 0; q < 3; q++) ab[k * 4 + q] = h[k * 4 + q]; ab[k * 4 + 3] = -h[k * 4 + 3]; }
  double c[3] = {-goal_pos[0] + sp[0], -goal_pos[1] + sp[1], -goal_pos[2] + sp[2]};
  double gmax[3], gmin[3];
  unsigned long long r1 = lp_seed(seed_base, it, kind0), r2 = lp_seed(seed_base, it, kind0 + 1);
  poplp::linprog<3>(c, ab, 4, ab + 3, 4, m, gmax, &r1, lpw, scri);
  for (int q = 0; q < 3; q++) c[q] = goal_pos[q] - sp[q];
  poplp::linprog<3>(c, ab, 4, ab + 3, 4, m, gmin, &r2, lpw, scri);
  for (int q = 0; q < 3; q++) goal_pos[q] = 0.5 * (gmax[q] + gmin[q]);
  return false;
}

// Do corridors i and i + 1 intersect?  (the feasibility LP of baseline_fake.cpp:366-378 = baseline.cpp:363-375)
__device__ inline bool corridor_pair_intersects(const double* H, const int* rows, int i, unsigned long long seed_base,
                                                double* ab, double* lpw, int* scri) {
  const int m0 = rows[i], m1 = rows[i + 1];
  for (int k = 0; k < m0; k++) {
    const double* h = H + (static_cast<size_t>(i) * POP_MAX_FACES + k) * 4;
    for (int q = 0; q < 3; q++) ab[k * 4 + q] = h[q];
    ab[k * 4 + 3] = -h[3];
  }
  for (int k = 0; k < m1; k++) {
    const double* h = H + (static_cast<size_t>(i + 1) * POP_MAX_FACES + k) * 4;
    for (int q = 0; q < 3; q++) ab[(m0 + k) * 4 + q] = h[q];
    ab[(m0 + k) * 4 + 3] = -h[3];
  }
  const double c[3] = {0, 0, 0};
  double x[3];
  unsigned long long rng = lp_seed(seed_base, i, 2);
  const double rst = poplp::linprog<3>(c, ab, 4, ab + 3, 4, m0 + m1, x, &rng, lpw, scri);
  return !isinf(rst);
}

// number of leading corridors whose validity check passed
__device__ inline int corridor_valid_prefix(int n, const int* valid) {
  int np = 0;
  for (int i = 0; i < n - 1 && i < POP_MAX_PIECES; i++) {
    if (!valid[i]) break;
    np++;
  }
  return np;
}

// Returns the number of corridors kept (0 = the replan fails here) and writes the local goal (pos, vel).
//   n        samples of the A* path P (n x 6)
//   H, rows  the agent's corridors, POP_MAX_PIECES x POP_MAX_FACES x 4 and their row counts
//   valid    per-segment result of checkCorridorValidity
//   sp       start state (position first)
//   ab       4 * (2 * POP_MAX_FACES + 8) doubles, lpw / scri the LP work arrays
//   inter    optional: inter[i] = corridor_pair_intersects(i) for every i + 1 < valid prefix, solved beforehand (the
//            pairs do not depend on one another; the kernel gives each its own warp); nullptr = solve them here
__device__ inline int corridor_finish_agent(int variant, unsigned long long seed_base, int n, const double* P,
                                            const double* H, const int* rows, const int* valid, const double* sp,
                                            double* ab, double* lpw, int* scri, double* out_goal,
                                            const int* inter = nullptr) {
  const bool real = variant != POP_QUERY_FAKE;
  int np = corridor_valid_prefix(n, valid);
  bool fail = false;
  // adjacent corridors must intersect (baseline_fake.cpp:366-378, baseline.cpp:363-375)
  for (int i = 0; i + 1 < np; i++) {
    const bool meets = inter ? inter[i] != 0 : corridor_pair_intersects(H, rows, i, seed_base, ab, lpw, scri);
    if (!meets) {
      if (i < 2) { fail = true; break; }
      np = real ? i : i + 1;
      break;
    }
  }
  if (fail || np <= (real ? 1 : 0)) return 0;
  int gi = np - 1;
  double goal_pos[3] = {P[gi * 6], P[gi * 6 + 1], P[gi * 6 + 2]};
  double goal_vel[3] = {P[gi * 6 + 3], P[gi * 6 + 4], P[gi * 6 + 5]};
  bool search = true;
  if (real)   // baseline.cpp:389: only an unreachable goal starts the backward search (and has been moved by now)
    search = !corridor_goal_reachable(H + static_cast<size_t>(np - 1) * POP_MAX_FACES * 4, rows[np - 1], sp, goal_pos,
                                      seed_base, np - 1, 5, ab, lpw, scri);
  for (int it = np - 1; search && it != 0; it--) {
    if (corridor_goal_reachable(H + static_cast<size_t>(it) * POP_MAX_FACES * 4, rows[it], sp, goal_pos, seed_base, it, 3,
                                ab, lpw, scri)) {
      np = it + 1;
      const int idx = np - 1;
      for (int q = 0; q < 3; q++) { goal_pos[q] = P[idx * 6 + q]; goal_vel[q] = P[idx * 6 + 3 + q]; }
      break;
    }
  }
  for (int q = 0; q < 3; q++) { out_goal[q] = goal_pos[q]; out_goal[3 + q] = goal_vel[q]; }
  return np;
}

// pop_lp.cuh — device Seidel LP (d = 3, 4): min c.x s.t. A x <= b.
// Incremental randomized LP on homogeneous half-spaces with recursion on the boundary of a
// violated constraint, the algorithm behind sdlp::linprog<d>
// (traj_utils/include/traj_utils/sdlp.hpp:25-785 of the reference).  Constraint order comes
// from a Fisher-Yates shuffle driven by a caller-provided 64-bit LCG state (the reference
// shuffles with a function-static mt19937_64 whose state depends on global call history).
//
// The algorithm and its structure are those of sdlp.hpp, whose notices are preserved here as they ask:
//
//   Copyright (c) 1990 Michael E. Hohmeyer,
//         hohmeyer@icemcfd.com
//   Permission is granted to modify and re-distribute this code in any manner
//   as long as this notice is preserved.  All standard disclaimers apply.
//
//   R. Seidel's algorithm for solving LPs (linear programs.)
//
//   Copyright (c) 2021 Zhepei Wang,
//         wangzhepei@live.com
//   1. Bug fix in "move_to_front" function that "prev[m]" is illegally accessed
//      while "prev" originally has only m ints. It is fixed by allocating a
//      "prev" with m + 1 ints.
//   2. Add Eigen interface.
//   3. Resursive template.
//   Permission is granted to modify and re-distribute this code in any manner
//   as long as this notice is preserved.  All standard disclaimers apply.
//
//   Ref: Seidel, R. (1991), "Small-dimensional linear programming and convex
//        hulls made easy", Discrete & Computational Geometry 6 (1): 423-434,
//        doi:10.1007/BF02574699
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace poplp {

constexpr double kEps = 1.0e-12;
enum { LP_MINIMUM = 0, LP_INFEASIBLE = 1, LP_UNBOUNDED = 2, LP_AMBIGUOUS = 3 };

__device__ __forceinline__ double d2(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1]; }
__device__ __forceinline__ double x2(const double* a, const double* b) { return a[0] * b[1] - a[1] * b[0]; }
__device__ __forceinline__ bool norm2(const double* a, double* b) {
  const double mag = sqrt(a[0] * a[0] + a[1] * a[1]);
  if (mag < 2.0 * kEps) return true;
  b[0] = a[0] / mag;
  b[1] = a[1] / mag;
  return false;
}

template <int d>
__device__ bool normalize_h(double* a) {
  double mag = 0.0;
  for (int i = 0; i <= d; i++) mag += a[i] * a[i];
  if (mag < (d + 1) * kEps * kEps) return true;
  mag = 1.0 / sqrt(mag);
  for (int i = 0; i <= d; i++) a[i] *= mag;
  return false;
}

template <int d>
__device__ int free_optimum(const double* nv, const double* dv, double* opt) {
  double nd = 0.0, dd = 0.0;
  for (int i = 0; i <= d; i++) { nd += nv[i] * dv[i]; dd += dv[i] * dv[i]; }
  if (dd < kEps * kEps) { nd = 0.0; dd = 1.0; }
  for (int i = 0; i <= d; i++) opt[i] = -nv[i] + dv[i] * nd / dd;
  if (normalize_h<d>(opt)) { opt[d] = 1.0; return LP_AMBIGUOUS; }
  return LP_MINIMUM;
}

__device__ __forceinline__ int to_front(int i, int* next, int* prev) {
  if (i == 0 || i == next[0]) return i;
  const int before = prev[i];
  next[prev[i]] = next[i];
  prev[next[i]] = prev[i];
  next[i] = next[0];
  prev[i] = 0;
  prev[next[i]] = i;
  next[0] = i;
  return before;
}

__device__ inline void rational_min(bool degen, const double* cw, const double* ccw, const double* nv,
                                    const double* dv, double* opt) {
  const double d_cw = d2(cw, dv), d_ccw = d2(ccw, dv), n_cw = d2(cw, nv), n_ccw = d2(ccw, nv);
  const double* pick;
  if (degen) {
    pick = n_cw / d_cw < n_ccw / d_ccw ? cw : ccw;
  } else if (fabs(d_cw) > 2.0 * kEps && fabs(d_ccw) > 2.0 * kEps) {
    if (d_cw * d_ccw > 0.0) {
      pick = n_cw / d_cw < n_ccw / d_ccw ? cw : ccw;
    } else if (d_cw > 0.0) {
      opt[0] = -dv[1]; opt[1] = dv[0];
      return;
    } else {
      opt[0] = dv[1]; opt[1] = -dv[0];
      return;
    }
  } else if (fabs(d_cw) > 2.0 * kEps) {
    pick = n_ccw * d_cw > 0.0 ? cw : ccw;
  } else if (fabs(d_ccw) > 2.0 * kEps) {
    pick = n_cw * d_ccw > 2.0 * kEps ? ccw : cw;
  } else {
    pick = x2(dv, nv) > 0.0 ? cw : ccw;
  }
  opt[0] = pick[0];
  opt[1] = pick[1];
}

__device__ inline int feasible_arc(const double* hv /*[m][2]*/, int m, int* next, int* prev, double* cw,
                                   double* ccw, bool* degen) {
  int i;
  *degen = false;
  for (i = 0; i != m; i = next[i]) {
    if (!norm2(hv + 2 * i, ccw)) {
      cw[0] = ccw[1]; cw[1] = -ccw[0];
      ccw[0] = -cw[0]; ccw[1] = -cw[1];
      break;
    }
  }
  if (i == m) return LP_UNBOUNDED;
  i = 0;
  while (i != m) {
    bool offensive = false;
    const double* h = hv + 2 * i;
    const double d_cw = d2(cw, h), d_ccw = d2(ccw, h);
    if (d_ccw >= 2.0 * kEps) {
      if (d_cw <= -2.0 * kEps) {
        cw[0] = h[1]; cw[1] = -h[0];
        norm2(cw, cw);
        offensive = true;
      }
    } else if (d_cw >= 2.0 * kEps) {
      if (d_ccw <= -2.0 * kEps) {
        ccw[0] = -h[1]; ccw[1] = h[0];
        norm2(ccw, ccw);
        offensive = true;
      }
    } else if (d_ccw <= -2.0 * kEps && d_cw <= -2.0 * kEps) {
      return LP_INFEASIBLE;
    } else if (d_cw <= -2.0 * kEps || d_ccw <= -2.0 * kEps || x2(cw, h) < 0.0) {
      if (d_cw <= -2.0 * kEps) norm2(ccw, cw);
      else if (d_ccw <= -2.0 * kEps) norm2(cw, ccw);
      *degen = true;
      offensive = true;
    }
    if (offensive) i = to_front(i, next, prev);
    i = next[i];
    if (*degen) break;
  }
  if (*degen) {
    while (i != m) {
      const double* h = hv + 2 * i;
      const double d_cw = d2(cw, h), d_ccw = d2(ccw, h);
      if (d_cw < -2.0 * kEps) {
        if (d_ccw < -2.0 * kEps) return LP_INFEASIBLE;
        cw[0] = ccw[0]; cw[1] = ccw[1];
      } else if (d_ccw < -2.0 * kEps) {
        ccw[0] = cw[0]; ccw[1] = cw[1];
      }
      i = next[i];
    }
  }
  return LP_MINIMUM;
}

__device__ inline int solve_1d(const double* hv, int m, const double* nv, const double* dv, double* opt,
                               int* next, int* prev) {
  double cw[2], ccw[2];
  bool degen;
  int status = feasible_arc(hv, m, next, prev, cw, ccw, &degen);
  if (status == LP_INFEASIBLE) return status;
  if (status == LP_UNBOUNDED) return free_optimum<1>(nv, dv, opt);
  if (fabs(x2(nv, dv)) < 2.0 * kEps * kEps) {
    if (d2(nv, nv) < 2.0 * kEps * kEps || d2(dv, dv) > 2.0 * kEps * kEps) {
      opt[0] = cw[0]; opt[1] = cw[1];
      status = LP_AMBIGUOUS;
    } else {
      if (!degen && x2(cw, nv) <= 0.0 && x2(nv, ccw) <= 0.0) {
        opt[0] = -nv[0]; opt[1] = -nv[1];
      } else if (d2(nv, cw) > d2(nv, ccw)) {
        opt[0] = ccw[0]; opt[1] = ccw[1];
      } else {
        opt[0] = cw[0]; opt[1] = cw[1];
      }
      status = LP_MINIMUM;
    }
  } else {
    rational_min(degen, cw, ccw, nv, dv, opt);
    status = LP_MINIMUM;
  }
  return status;
}

template <int d>
__device__ int solve_nd(const double* hv, int max_size, int m, const double* nv, const double* dv,
                        double* opt, double* work, int* next, int* prev) {
  if constexpr (d == 1) {
    if (m > 0) return solve_1d(hv, m, nv, dv, opt, next, prev);
    return free_optimum<1>(nv, dv, opt);
  } else {
    double val = 0.0;
    for (int j = 0; j <= d; j++) val += dv[j] * dv[j];
    const bool dv_zero = val < (d + 1) * kEps * kEps;
    int status = free_optimum<d>(nv, dv, opt);
    if (m <= 0) return status;
    double* sub_opt = work;
    double* sub_n = sub_opt + d;
    double* sub_d = sub_n + d;
    double* sub_h = sub_d + d;
    double* sub_work = sub_h + max_size * d;
    for (int i = 0; i != m; i = next[i]) {
      const double* pl = hv + i * (d + 1);
      val = 0.0;
      for (int j = 0; j <= d; j++) val += opt[j] * pl[j];
      if (val < -(d + 1) * kEps) {
        int imax = 0;
        double rmax = fabs(pl[0]);
        for (int j = 1; j <= d; j++) {
          const double ab = fabs(pl[j]);
          if (ab > rmax) { imax = j; rmax = ab; }
        }
        if (i != 0) {
          const double fac = 1.0 / pl[imax];
          for (int j = 0; j != i; j = next[j]) {
            const double* op = hv + j * (d + 1);
            const double crit = op[imax] * fac;
            double* np = sub_h + j * d;
            for (int k = 0; k <= d; k++) {
              const int l = k < imax ? k : k - 1;
              np[l] = k != imax ? op[k] - pl[k] * crit : np[l];
            }
          }
        }
        if (dv_zero) {
          double ve = 0.0, ee = 0.0;
          for (int j = 0; j <= d; j++) { ve += nv[j] * pl[j]; ee += pl[j] * pl[j]; }
          const double fac = ve / ee;
          for (int j = 0; j <= d; j++)
            if (j != imax) sub_n[j < imax ? j : j - 1] = nv[j] - pl[j] * fac;
          for (int j = 0; j < d; j++) sub_d[j] = 0.0;
        } else {
          const double cn = nv[imax] / pl[imax], cd = dv[imax] / pl[imax];
          for (int j = 0; j <= d; j++)
            if (j != imax) {
              sub_n[j < imax ? j : j - 1] = nv[j] - pl[j] * cn;
              sub_d[j < imax ? j : j - 1] = dv[j] - pl[j] * cd;
            }
        }
        status = solve_nd<d - 1>(sub_h, max_size, i, sub_n, sub_d, sub_opt, sub_work, next, prev);
        if (status == LP_INFEASIBLE) return status;
        opt[imax] = 0.0;
        for (int j = 0; j <= d; j++)
          if (j != imax) {
            const int l = j < imax ? j : j - 1;
            opt[j] = sub_opt[l];
            opt[imax] -= pl[j] * sub_opt[l];
          }
        opt[imax] /= pl[imax];
        double mag = 0.0;
        for (int j = 0; j <= d; j++) mag += opt[j] * opt[j];
        mag = 1.0 / sqrt(mag);
        for (int j = 0; j <= d; j++) opt[j] *= mag;
        i = to_front(i, next, prev);
      }
    }
    return status;
  }
}

__device__ __forceinline__ unsigned lcg_next(unsigned long long* s) {
  *s = *s * 6364136223846793005ull + 1442695040888963407ull;
  return static_cast<unsigned>(*s >> 33);
}

// workspace sizes for `rows` constraints in dimension d
__host__ __device__ constexpr int work_doubles(int rows, int d) {
  return (d + 1) * (rows + 1) + (rows + 4) * (d + 2) * (d - 1) / 2 + 8;
}
__host__ __device__ constexpr int work_ints(int rows) { return 3 * (rows + 2); }

// A is rows x d (row stride lda), b has rows entries (stride ldb); single-thread routine.
template <int d>
__device__ double linprog(const double* c, const double* A, int lda, const double* b, int ldb, int rows,
                          double* x, unsigned long long* rng, double* wd, int* wi) {
  const int m = rows + 1;
  for (int i = 0; i < d; i++) x[i] = 0.0;
  if (m <= 1) {
    double mx = 0.0;
    for (int i = 0; i < d; i++) mx = fmax(mx, fabs(c[i]));
    return mx > 0.0 ? -INFINITY : 0.0;
  }
  double* hv = wd;
  double* work = hv + (d + 1) * m;
  int* perm = wi;
  int* next = perm + m;
  int* prev = next + m + 1;
  double nv[d + 1], dv[d + 1], opt[d + 1];
  for (int j = 0; j < d; j++) hv[j] = 0.0;
  hv[d] = 1.0;
  for (int i = 0; i < rows; i++) {
    double* h = hv + (i + 1) * (d + 1);
    double sq = 0.0;
    for (int j = 0; j < d; j++) { h[j] = -A[i * lda + j]; sq += h[j] * h[j]; }
    h[d] = b[i * ldb];
    sq += h[d] * h[d];
    const double nrm = sqrt(sq);
    if (nrm > 0.0)
      for (int j = 0; j <= d; j++) h[j] /= nrm;
  }
  for (int j = 0; j < d; j++) { nv[j] = c[j]; dv[j] = 0.0; }
  nv[d] = 0.0;
  dv[d] = 1.0;
  for (int i = 0; i < m - 1; i++) perm[i] = i;
  for (int i = 0; i < m - 1; i++) {
    const int j = static_cast<int>(lcg_next(rng) % static_cast<unsigned>(m - 1 - i)) + i;
    const int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
  }
  prev[0] = 0;
  next[0] = perm[0] + 1;
  prev[perm[0] + 1] = 0;
  for (int i = 0; i < m - 2; i++) {
    next[perm[i] + 1] = perm[i + 1] + 1;
    prev[perm[i + 1] + 1] = perm[i] + 1;
  }
  next[perm[m - 2] + 1] = m;
  const int status = solve_nd<d>(hv, m, m, nv, dv, opt, work, next, prev);
  double minimum = INFINITY;
  if (status != LP_INFEASIBLE) {
    if (opt[d] != 0.0 && status != LP_UNBOUNDED) {
      minimum = 0.0;
      for (int j = 0; j < d; j++) x[j] = opt[j] / opt[d];
      for (int j = 0; j < d; j++) minimum += c[j] * x[j];
    }
    if (opt[d] == 0.0 || status == LP_UNBOUNDED) {
      for (int j = 0; j < d; j++) x[j] = opt[j];
      minimum = -INFINITY;
    }
  }
  return minimum;
}


// =================================================================================================================
// The same algorithm for a whole warp (device only).  Seidel's method walks the constraints in a fixed (shuffled,
// move-to-front) order and almost all of them are no-ops: a constraint the current optimum already satisfies changes
// nothing.  So the 32 lanes test the next 32 constraints of the order at once, the first one that is violated is
// handled exactly as the sequential code handles it (all lanes carry the same optimum in registers and do the same
// scalar arithmetic), and the walk resumes behind it.  The projection of the earlier constraints onto a violated
// one and the move-to-front are done lane-parallel.  The order lives in an array (position -> constraint) instead
// of a linked list, so "the constraints before position p" is a range.  Same arithmetic, same order of events, same
// result bit for bit as linprog<d> (tests/test_gpu_lp.py compares them on the device).
// Every lane of a converged warp must call it with the same arguments; the result is returned to every lane.
// =================================================================================================================
#ifdef __CUDACC__
namespace warp {

constexpr unsigned kFull = 0xffffffffu;
__device__ __forceinline__ int lane_id() { return static_cast<int>(threadIdx.x & 31u); }

// ord[1..p] rotated right by one: the constraint at position p goes to position 1 (to_front); p <= 1: nothing moves
__device__ __forceinline__ void rotate_to_front(int* ord, int p) {
  if (p <= 1) return;
  const int lane = lane_id();
  const int moved = ord[p];
  __syncwarp();
  for (int base = p; base > 1; base -= 32) {
    const int k = base - lane;
    const int v = k > 1 ? ord[k - 1] : 0;
    __syncwarp();
    if (k > 1) ord[k] = v;
    __syncwarp();
  }
  if (lane == 0) ord[1] = moved;
  __syncwarp();
}

// first position q in [from, to) for which pred(q) holds, or `to`
template <typename Pred>
__device__ __forceinline__ int first_where(int from, int to, Pred pred) {
  const int lane = lane_id();
  for (int base = from; base < to; base += 32) {
    const int q = base + lane;
    const unsigned hit = __ballot_sync(kFull, q < to && pred(q));
    if (hit) return base + __ffs(hit) - 1;
  }
  return to;
}

// what feasible_arc's first loop does with constraint h given the current arc: 0 nothing, 1 cw update, 2 ccw update,
// 3 infeasible, 4 degenerate
__device__ __forceinline__ int arc_event(const double* cw, const double* ccw, const double* h) {
  const double d_cw = d2(cw, h), d_ccw = d2(ccw, h);
  if (d_ccw >= 2.0 * kEps) return d_cw <= -2.0 * kEps ? 1 : 0;
  if (d_cw >= 2.0 * kEps) return d_ccw <= -2.0 * kEps ? 2 : 0;
  if (d_ccw <= -2.0 * kEps && d_cw <= -2.0 * kEps) return 3;
  if (d_cw <= -2.0 * kEps || d_ccw <= -2.0 * kEps || x2(cw, h) < 0.0) return 4;
  return 0;
}

__device__ inline int feasible_arc(const double* hv /*[id][2]*/, int m, int* ord, double* cw, double* ccw, bool* degen) {
  *degen = false;
  int q = first_where(0, m, [&](int k) { double t[2]; return !norm2(hv + 2 * ord[k], t); });
  if (q == m) return LP_UNBOUNDED;
  norm2(hv + 2 * ord[q], ccw);
  cw[0] = ccw[1]; cw[1] = -ccw[0];
  ccw[0] = -cw[0]; ccw[1] = -cw[1];
  q = 0;
  while (true) {
    q = first_where(q, m, [&](int k) { return arc_event(cw, ccw, hv + 2 * ord[k]) != 0; });
    if (q == m) break;
    const double* h = hv + 2 * ord[q];
    const int ev = arc_event(cw, ccw, h);
    if (ev == 3) return LP_INFEASIBLE;
    if (ev == 1) {
      cw[0] = h[1]; cw[1] = -h[0];
      norm2(cw, cw);
    } else if (ev == 2) {
      ccw[0] = -h[1]; ccw[1] = h[0];
      norm2(ccw, ccw);
    } else {
      const double d_cw = d2(cw, h), d_ccw = d2(ccw, h);
      if (d_cw <= -2.0 * kEps) norm2(ccw, cw);
      else if (d_ccw <= -2.0 * kEps) norm2(cw, ccw);
      *degen = true;
    }
    rotate_to_front(ord, q);   // every event of this loop is "offensive"
    q++;
    if (*degen) break;
  }
  if (*degen) {
    while (true) {
      q = first_where(q, m, [&](int k) {
        const double* g = hv + 2 * ord[k];
        return d2(cw, g) < -2.0 * kEps || d2(ccw, g) < -2.0 * kEps;
      });
      if (q == m) break;
      const double* h = hv + 2 * ord[q];
      const double d_cw = d2(cw, h), d_ccw = d2(ccw, h);
      if (d_cw < -2.0 * kEps) {
        if (d_ccw < -2.0 * kEps) return LP_INFEASIBLE;
        cw[0] = ccw[0]; cw[1] = ccw[1];
      } else {
        ccw[0] = cw[0]; ccw[1] = cw[1];
      }
      q++;
    }
  }
  return LP_MINIMUM;
}

__device__ inline int solve_1d(const double* hv, int m, const double* nv, const double* dv, double* opt, int* ord) {
  double cw[2], ccw[2];
  bool degen;
  int status = feasible_arc(hv, m, ord, cw, ccw, &degen);
  if (status == LP_INFEASIBLE) return status;
  if (status == LP_UNBOUNDED) return free_optimum<1>(nv, dv, opt);
  if (fabs(x2(nv, dv)) < 2.0 * kEps * kEps) {
    if (d2(nv, nv) < 2.0 * kEps * kEps || d2(dv, dv) > 2.0 * kEps * kEps) {
      opt[0] = cw[0]; opt[1] = cw[1];
      status = LP_AMBIGUOUS;
    } else {
      if (!degen && x2(cw, nv) <= 0.0 && x2(nv, ccw) <= 0.0) {
        opt[0] = -nv[0]; opt[1] = -nv[1];
      } else if (d2(nv, cw) > d2(nv, ccw)) {
        opt[0] = ccw[0]; opt[1] = ccw[1];
      } else {
        opt[0] = cw[0]; opt[1] = cw[1];
      }
      status = LP_MINIMUM;
    }
  } else {
    rational_min(degen, cw, ccw, nv, dv, opt);
    status = LP_MINIMUM;
  }
  return status;
}

// m = number of order positions in play (the constraints before the one the caller is handling)
template <int d>
__device__ int solve_nd(const double* hv, int max_size, int m, const double* nv, const double* dv, double* opt,
                        double* work, int* ord) {
  if constexpr (d == 1) {
    if (m > 0) return solve_1d(hv, m, nv, dv, opt, ord);
    return free_optimum<1>(nv, dv, opt);
  } else {
    const int lane = lane_id();
    double val = 0.0;
    for (int j = 0; j <= d; j++) val += dv[j] * dv[j];
    const bool dv_zero = val < (d + 1) * kEps * kEps;
    int status = free_optimum<d>(nv, dv, opt);
    if (m <= 0) return status;
    double* sub_h = work;                       // [constraint id][d]
    double* sub_work = sub_h + max_size * d;
    int p = 0;
    while (true) {
      p = first_where(p, m, [&](int k) {
        const double* pl = hv + ord[k] * (d + 1);
        double v = 0.0;
        for (int j = 0; j <= d; j++) v += opt[j] * pl[j];
        return v < -(d + 1) * kEps;
      });
      if (p == m) break;
      const int i = ord[p];
      const double* pl = hv + i * (d + 1);
      int imax = 0;
      double rmax = fabs(pl[0]);
      for (int j = 1; j <= d; j++) {
        const double ab = fabs(pl[j]);
        if (ab > rmax) { imax = j; rmax = ab; }
      }
      if (p != 0) {
        const double fac = 1.0 / pl[imax];
        for (int q = lane; q < p; q += 32) {
          const int j = ord[q];
          const double* op = hv + j * (d + 1);
          const double crit = op[imax] * fac;
          double* np = sub_h + j * d;
          for (int k = 0; k <= d; k++) {
            const int l = k < imax ? k : k - 1;
            if (k != imax) np[l] = op[k] - pl[k] * crit;
          }
        }
      }
      double sub_n[d], sub_d[d], sub_opt[d];
      if (dv_zero) {
        double ve = 0.0, ee = 0.0;
        for (int j = 0; j <= d; j++) { ve += nv[j] * pl[j]; ee += pl[j] * pl[j]; }
        const double fac = ve / ee;
        for (int j = 0; j <= d; j++)
          if (j != imax) sub_n[j < imax ? j : j - 1] = nv[j] - pl[j] * fac;
        for (int j = 0; j < d; j++) sub_d[j] = 0.0;
      } else {
        const double cn = nv[imax] / pl[imax], cd = dv[imax] / pl[imax];
        for (int j = 0; j <= d; j++)
          if (j != imax) {
            sub_n[j < imax ? j : j - 1] = nv[j] - pl[j] * cn;
            sub_d[j < imax ? j : j - 1] = dv[j] - pl[j] * cd;
          }
      }
      __syncwarp();   // the projected constraints are read by every lane below
      status = solve_nd<d - 1>(sub_h, max_size, p, sub_n, sub_d, sub_opt, sub_work, ord);
      if (status == LP_INFEASIBLE) return status;
      opt[imax] = 0.0;
      for (int j = 0; j <= d; j++)
        if (j != imax) {
          const int l = j < imax ? j : j - 1;
          opt[j] = sub_opt[l];
          opt[imax] -= pl[j] * sub_opt[l];
        }
      opt[imax] /= pl[imax];
      double mag = 0.0;
      for (int j = 0; j <= d; j++) mag += opt[j] * opt[j];
      mag = 1.0 / sqrt(mag);
      for (int j = 0; j <= d; j++) opt[j] *= mag;
      rotate_to_front(ord, p);
      p++;
    }
    return status;
  }
}

// workspace for `rows` constraints in dimension d (the linked-list version's sizes cover it)
template <int d>
__device__ double linprog(const double* c, const double* A, int lda, const double* b, int ldb, int rows, double* x,
                          unsigned long long* rng, double* wd, int* wi) {
  const int lane = lane_id();
  const int m = rows + 1;
  for (int i = 0; i < d; i++) x[i] = 0.0;
  if (m <= 1) {
    double mx = 0.0;
    for (int i = 0; i < d; i++) mx = fmax(mx, fabs(c[i]));
    return mx > 0.0 ? -INFINITY : 0.0;
  }
  double* hv = wd;
  double* work = hv + (d + 1) * m;
  int* perm = wi;
  int* ord = perm + m;
  double nv[d + 1], dv[d + 1], opt[d + 1];
  __syncwarp();
  if (lane == 0) {
    for (int j = 0; j < d; j++) hv[j] = 0.0;
    hv[d] = 1.0;
  }
  for (int i = lane; i < rows; i += 32) {
    double* h = hv + (i + 1) * (d + 1);
    double sq = 0.0;
    for (int j = 0; j < d; j++) { h[j] = -A[i * lda + j]; sq += h[j] * h[j]; }
    h[d] = b[i * ldb];
    sq += h[d] * h[d];
    const double nrm = sqrt(sq);
    if (nrm > 0.0)
      for (int j = 0; j <= d; j++) h[j] /= nrm;
  }
  for (int j = 0; j < d; j++) { nv[j] = c[j]; dv[j] = 0.0; }
  nv[d] = 0.0;
  dv[d] = 1.0;
  unsigned long long state = *rng;
  if (lane == 0) {   // the shuffle is a sequential recurrence; it is a few hundred integer operations
    for (int i = 0; i < m - 1; i++) perm[i] = i;
    for (int i = 0; i < m - 1; i++) {
      const int j = static_cast<int>(lcg_next(&state) % static_cast<unsigned>(m - 1 - i)) + i;
      const int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
  }
  state = __shfl_sync(kFull, state, 0);
  *rng = state;
  __syncwarp();
  for (int i = lane; i < m; i += 32) ord[i] = i == 0 ? 0 : perm[i - 1] + 1;
  __syncwarp();
  const int status = solve_nd<d>(hv, m, m, nv, dv, opt, work, ord);
  double minimum = INFINITY;
  if (status != LP_INFEASIBLE) {
    if (opt[d] != 0.0 && status != LP_UNBOUNDED) {
      minimum = 0.0;
      for (int j = 0; j < d; j++) x[j] = opt[j] / opt[d];
      for (int j = 0; j < d; j++) minimum += c[j] * x[j];
    }
    if (opt[d] == 0.0 || status == LP_UNBOUNDED) {
      for (int j = 0; j < d; j++) x[j] = opt[j];
      minimum = -INFINITY;
    }
  }
  __syncwarp();
  return minimum;
}

}  // namespace warp
#endif  // __CUDACC__

}  // namespace poplp

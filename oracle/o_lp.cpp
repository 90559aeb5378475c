// o_lp.cpp — oracle: Seidel's randomized incremental LP in d = 1..4 (TEST INFRASTRUCTURE).
// Algorithm of traj_utils/include/traj_utils/sdlp.hpp (linfracprog / lp_base_case / wedge /
// linprog, :25-785) restated on flat arrays: the problem  min c.x  s.t. A x <= b  is lifted
// to homogeneous half-spaces  h.(x,1) >= 0, the optimum of a linear-fractional objective is
// maintained while constraints are added one at a time, and a violated constraint triggers
// a recursive solve on its boundary hyperplane (dimension d-1).
// Deviation: the reference shuffles constraints with a function-static
// mt19937_64 (result depends on global call history); here the shuffle is a Fisher-Yates
// driven by a caller-owned 64-bit LCG so CPU and GPU agree call by call.
// PINNED against the reference's own sdlp.hpp compiled into oracle/_ref/libref.so: optimum and minimiser within
// 1e-9 / 1e-7 on bounded problems, same infeasible / unbounded classification (tests/test_ref_pins.py).
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
#include <cmath>
#include <limits>
#include <vector>

#include "oracle.hpp"

namespace orc {
namespace {

constexpr double kEps = 1.0e-12;
enum { LP_MINIMUM = 0, LP_INFEASIBLE = 1, LP_UNBOUNDED = 2, LP_AMBIGUOUS = 3 };

inline double d2(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1]; }
inline double x2(const double* a, const double* b) { return a[0] * b[1] - a[1] * b[0]; }
inline bool norm2(const double* a, double* b) {  // true when a is (numerically) zero
  const double mag = std::sqrt(a[0] * a[0] + a[1] * a[1]);
  if (mag < 2.0 * kEps) return true;
  b[0] = a[0] / mag;
  b[1] = a[1] / mag;
  return false;
}

template <int d>
bool normalize_h(double* a) {
  double mag = 0.0;
  for (int i = 0; i <= d; i++) mag += a[i] * a[i];
  if (mag < (d + 1) * kEps * kEps) return true;
  mag = 1.0 / std::sqrt(mag);
  for (int i = 0; i <= d; i++) a[i] *= mag;
  return false;
}

// unconstrained optimum of n.x / d.x on the unit sphere
template <int d>
int free_optimum(const double* nv, const double* dv, double* opt) {
  double nd = 0.0, dd = 0.0;
  for (int i = 0; i <= d; i++) { nd += nv[i] * dv[i]; dd += dv[i] * dv[i]; }
  if (dd < kEps * kEps) { nd = 0.0; dd = 1.0; }
  for (int i = 0; i <= d; i++) opt[i] = -nv[i] + dv[i] * nd / dd;
  if (normalize_h<d>(opt)) { opt[d] = 1.0; return LP_AMBIGUOUS; }
  return LP_MINIMUM;
}

// linked-list helper: move constraint i right behind the head; returns the index now in i's place
inline int to_front(int i, int* next, int* prev) {
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

void rational_min(bool degen, const double* cw, const double* ccw, const double* nv, const double* dv,
                  double* opt) {
  const double d_cw = d2(cw, dv), d_ccw = d2(ccw, dv), n_cw = d2(cw, nv), n_ccw = d2(ccw, nv);
  auto take = [&](const double* v) { opt[0] = v[0]; opt[1] = v[1]; };
  if (degen) {
    take(n_cw / d_cw < n_ccw / d_ccw ? cw : ccw);
  } else if (std::fabs(d_cw) > 2.0 * kEps && std::fabs(d_ccw) > 2.0 * kEps) {
    if (d_cw * d_ccw > 0.0) {
      take(n_cw / d_cw < n_ccw / d_ccw ? cw : ccw);
    } else if (d_cw > 0.0) {
      opt[0] = -dv[1]; opt[1] = dv[0];
    } else {
      opt[0] = dv[1]; opt[1] = -dv[0];
    }
  } else if (std::fabs(d_cw) > 2.0 * kEps) {
    take(n_ccw * d_cw > 0.0 ? cw : ccw);
  } else if (std::fabs(d_ccw) > 2.0 * kEps) {
    take(n_cw * d_ccw > 2.0 * kEps ? ccw : cw);
  } else {
    take(x2(dv, nv) > 0.0 ? cw : ccw);
  }
}

// feasible arc [cw, ccw] of the projective line under the listed half-lines
int feasible_arc(const double (*hv)[2], int m, int* next, int* prev, double* cw, double* ccw, bool* degen) {
  int i;
  *degen = false;
  for (i = 0; i != m; i = next[i]) {
    if (!norm2(hv[i], ccw)) {
      cw[0] = ccw[1]; cw[1] = -ccw[0];
      ccw[0] = -cw[0]; ccw[1] = -cw[1];
      break;
    }
  }
  if (i == m) return LP_UNBOUNDED;
  i = 0;
  while (i != m) {
    bool offensive = false;
    double d_cw = d2(cw, hv[i]), d_ccw = d2(ccw, hv[i]);
    if (d_ccw >= 2.0 * kEps) {
      if (d_cw <= -2.0 * kEps) {
        cw[0] = hv[i][1]; cw[1] = -hv[i][0];
        norm2(cw, cw);
        offensive = true;
      }
    } else if (d_cw >= 2.0 * kEps) {
      if (d_ccw <= -2.0 * kEps) {
        ccw[0] = -hv[i][1]; ccw[1] = hv[i][0];
        norm2(ccw, ccw);
        offensive = true;
      }
    } else if (d_ccw <= -2.0 * kEps && d_cw <= -2.0 * kEps) {
      return LP_INFEASIBLE;
    } else if (d_cw <= -2.0 * kEps || d_ccw <= -2.0 * kEps || x2(cw, hv[i]) < 0.0) {
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
      double d_cw = d2(cw, hv[i]), d_ccw = d2(ccw, hv[i]);
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

int solve_1d(const double (*hv)[2], int m, const double* nv, const double* dv, double* opt, int* next,
             int* prev) {
  double cw[2], ccw[2];
  bool degen;
  int status = feasible_arc(hv, m, next, prev, cw, ccw, &degen);
  if (status == LP_INFEASIBLE) return status;
  if (status == LP_UNBOUNDED) return free_optimum<1>(nv, dv, opt);
  if (std::fabs(x2(nv, dv)) < 2.0 * kEps * kEps) {
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
int solve_nd(const double* hv, int max_size, int m, const double* nv, const double* dv, double* opt,
             double* work, int* next, int* prev) {
  if constexpr (d == 1) {
    if (m > 0) return solve_1d(reinterpret_cast<const double(*)[2]>(hv), m, nv, dv, opt, next, prev);
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
        // eliminate the largest coefficient of the violated plane
        int imax = 0;
        double rmax = std::fabs(pl[0]);
        for (int j = 1; j <= d; j++) {
          const double ab = std::fabs(pl[j]);
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
        // back-substitute the eliminated variable, renormalise
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
        mag = 1.0 / std::sqrt(mag);
        for (int j = 0; j <= d; j++) opt[j] *= mag;
        i = to_front(i, next, prev);
      }
    }
    return status;
  }
}

inline uint32_t lcg_next(uint64_t* s) {
  *s = *s * 6364136223846793005ull + 1442695040888963407ull;
  return static_cast<uint32_t>(*s >> 33);
}

template <int d>
double linprog_d(const double* c, const double* A, const double* b, int rows, double* x, uint64_t* rng) {
  const int m = rows + 1;
  for (int i = 0; i < d; i++) x[i] = 0.0;
  if (m <= 1) {
    double mx = 0.0;
    for (int i = 0; i < d; i++) mx = std::max(mx, std::fabs(c[i]));
    return mx > 0.0 ? -std::numeric_limits<double>::infinity() : 0.0;
  }
  std::vector<int> perm(m - 1), next(m), prev(m + 1);
  std::vector<double> hv(static_cast<size_t>(d + 1) * m), work(static_cast<size_t>(m + 3) * (d + 2) * (d - 1) / 2 + 8);
  double nv[d + 1], dv[d + 1], opt[d + 1];
  for (int j = 0; j < d; j++) hv[j] = 0.0;
  hv[d] = 1.0;
  for (int i = 0; i < rows; i++) {
    double* h = &hv[static_cast<size_t>(i + 1) * (d + 1)];
    double sq = 0.0;
    for (int j = 0; j < d; j++) { h[j] = -A[i * d + j]; sq += h[j] * h[j]; }
    h[d] = b[i];
    sq += h[d] * h[d];
    // Eigen colwise().normalize(): v /= sqrt(squaredNorm) when the norm is positive
    const double nrm = std::sqrt(sq);
    if (nrm > 0.0)
      for (int j = 0; j <= d; j++) h[j] /= nrm;
  }
  for (int j = 0; j < d; j++) { nv[j] = c[j]; dv[j] = 0.0; }
  nv[d] = 0.0;
  dv[d] = 1.0;
  for (int i = 0; i < m - 1; i++) perm[i] = i;
  for (int i = 0; i < m - 1; i++) {
    const int j = static_cast<int>(lcg_next(rng) % static_cast<uint32_t>(m - 1 - i)) + i;
    std::swap(perm[i], perm[j]);
  }
  prev[0] = 0;
  next[0] = perm[0] + 1;
  prev[perm[0] + 1] = 0;
  for (int i = 0; i < m - 2; i++) {
    next[perm[i] + 1] = perm[i + 1] + 1;
    prev[perm[i + 1] + 1] = perm[i] + 1;
  }
  next[perm[m - 2] + 1] = m;
  const int status = solve_nd<d>(hv.data(), m, m, nv, dv, opt, work.data(), next.data(), prev.data());
  double minimum = std::numeric_limits<double>::infinity();
  if (status != LP_INFEASIBLE) {
    if (opt[d] != 0.0 && status != LP_UNBOUNDED) {
      minimum = 0.0;
      for (int j = 0; j < d; j++) { x[j] = opt[j] / opt[d]; }
      for (int j = 0; j < d; j++) minimum += c[j] * x[j];
    }
    if (opt[d] == 0.0 || status == LP_UNBOUNDED) {
      for (int j = 0; j < d; j++) x[j] = opt[j];
      minimum = -std::numeric_limits<double>::infinity();
    }
  }
  return minimum;
}

}  // namespace

// sdlp::linprog<d>: min c.x s.t. A x <= b; returns the minimum (inf infeasible, -inf unbounded)
double linprog(int d, const double* c, const double* A, const double* b, int m, double* x,
               uint64_t* rng_state) {
  if (d == 3) return linprog_d<3>(c, A, b, m, x, rng_state);
  if (d == 4) return linprog_d<4>(c, A, b, m, x, rng_state);
  if (d == 2) return linprog_d<2>(c, A, b, m, x, rng_state);
  return std::numeric_limits<double>::quiet_NaN();
}

}  // namespace orc

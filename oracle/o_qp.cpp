// o_qp.cpp — oracle: corridor-constrained Bezier QP and its ADMM solver (TEST INFRASTRUCTURE).
//
// qp_setup restates traj_opt/src/bezier_optimizer.cpp:27-260 (calcCtrlPtsCvtMat, calcMinJerkCost,
// addContinuityConstraints, addDynamicalConstraints, addSafetyConstraints) on dense row-major
// arrays.  osqp_solve restates the algorithm of OSQP 0.6.x, the un-vendored, un-pinned solver
// the reference links (README.md:13-22; traj_opt/include/iosqp.hpp:40-115 uses the 0.6 C API
// with defaults except eps_abs = eps_rel = 1e-3): Ruiz equilibration (scaling = 10), per-row
// rho (equalities 1e3 * rho), over-relaxed ADMM (alpha 1.6, sigma 1e-6), unscaled termination
// every 25 iterations, residual-balancing rho adaptation.  Two documented deviations:
//   (1) OSQP picks the rho-adaptation interval from wall-clock time (non-deterministic); here it
//       is the fixed pop_config.osqp_adaptive_rho_interval (25 = what the clock rule yields for
//       problems this small);
//   (2) the quasi-definite KKT solve (QDLDL) is replaced by the algebraically identical
//       reduced system (P + sigma I + A' diag(rho) A) x = rhs with a dense Cholesky.
// PARITY UNPINNED against OSQP itself (not available here); pinned properties: the reference's
// own QP test fixture (traj_opt/test/test_bezier_opt.cpp:47-221).
// qp_setup is PINNED against the reference's own bezier_optimizer.cpp compiled into oracle/_ref/libref.so: Q, A, lb,
// ub identical (tests/test_ref_pins.py).  osqp_solve has no reference build to compare with (OSQP is absent): pinned
// only by the reference's QP fixture (tests/test_golden_reference.py) and a tight-tolerance cross-solve.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "oracle.hpp"

namespace orc {

namespace {
constexpr int N_ = 4, DIM = 3;
constexpr double OSQP_INFTY = 1e30;
constexpr double MIN_SCALING = 1e-04, MAX_SCALING = 1e+04;
constexpr double RHO_MIN = 1e-06, RHO_MAX = 1e+06, RHO_EQ_OVER_RHO_INEQ = 1e+03, RHO_TOL = 1e-04;
}  // namespace

// BezierOpt::setup (bezier_optimizer.cpp:27-54). init/goal rows = pos; vel; acc.
void qp_setup(QpProblem& p, const double init[9], const double goal[9], const std::vector<double>& t,
              const std::vector<Poly>& polys, double vmax, double amax) {
  const int M = static_cast<int>(t.size());
  const int n = DIM * M * (N_ + 1);
  // calcCtrlPtsCvtMat (:56-83): p2v 12x15, v2a 9x12, a2j 6x9
  double p2v[12][15] = {{0}}, v2a[9][12] = {{0}}, a2j[6][9] = {{0}};
  for (int i = 0; i < N_; i++)
    for (int d = 0; d < DIM; d++) { p2v[i * 3 + d][i * 3 + d] = -N_; p2v[i * 3 + d][(i + 1) * 3 + d] = N_; }
  for (int i = 0; i < N_ - 1; i++)
    for (int d = 0; d < DIM; d++) { v2a[i * 3 + d][i * 3 + d] = -(N_ - 1); v2a[i * 3 + d][(i + 1) * 3 + d] = N_ - 1; }
  for (int i = 0; i < N_ - 2; i++)
    for (int d = 0; d < DIM; d++) { a2j[i * 3 + d][i * 3 + d] = -(N_ - 2); a2j[i * 3 + d][(i + 1) * 3 + d] = N_ - 2; }
  double p2a[9][15] = {{0}}, p2j[6][15] = {{0}};
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 15; j++) { double s = 0; for (int k = 0; k < 12; k++) s += v2a[i][k] * p2v[k][j]; p2a[i][j] = s; }
  // p2j = a2j * v2a * p2v evaluated left to right: (a2j*v2a)*p2v
  double a2j_v2a[6][12] = {{0}};
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 12; j++) { double s = 0; for (int k = 0; k < 9; k++) s += a2j[i][k] * v2a[k][j]; a2j_v2a[i][j] = s; }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 15; j++) { double s = 0; for (int k = 0; k < 12; k++) s += a2j_v2a[i][k] * p2v[k][j]; p2j[i][j] = s; }
  // calcMinJerkCost (:91-105): QM = p2j' * [[I/3, I/6],[I/6, I/3]] * p2j on every diagonal block
  double Pm[6][6] = {{0}};
  for (int d = 0; d < 3; d++) { Pm[d][d] = 1.0 / 3; Pm[d][3 + d] = 1.0 / 6; Pm[3 + d][d] = 1.0 / 6; Pm[3 + d][3 + d] = 1.0 / 3; }
  double tP[15][6] = {{0}}, QM[15][15] = {{0}};
  for (int i = 0; i < 15; i++)
    for (int j = 0; j < 6; j++) { double s = 0; for (int k = 0; k < 6; k++) s += p2j[k][i] * Pm[k][j]; tP[i][j] = s; }
  for (int i = 0; i < 15; i++)
    for (int j = 0; j < 15; j++) { double s = 0; for (int k = 0; k < 6; k++) s += tP[i][k] * p2j[k][j]; QM[i][j] = s; }
  p.n = n;
  p.Q.assign(static_cast<size_t>(n) * n, 0.0);
  for (int i = 0; i < M; i++)
    for (int a = 0; a < 15; a++)
      for (int b = 0; b < 15; b++) p.Q[static_cast<size_t>(i * 15 + a) * n + i * 15 + b] = QM[a][b];

  // addConstraints (:107-131)
  int num_const = 0;
  for (const Poly& c : polys) num_const += c.rows();
  num_const *= N_ + 1;
  const int num_cont = (1 + M) * DIM * 3, num_dyn = M * (DIM * N_ + DIM * (N_ - 1));
  const int m = num_const + num_cont + num_dyn;
  p.m = m;
  p.A.assign(static_cast<size_t>(m) * n, 0.0);
  p.l.assign(m, 0.0);
  p.u.assign(m, 0.0);
  auto A = [&](int r, int c) -> double& { return p.A[static_cast<size_t>(r) * n + c]; };
  int idx = 0;
  const double t0 = t[0], tM = t[M - 1];
  // ---- addContinuityConstraints (:137-207) ----
  for (int d = 0; d < 3; d++) { A(idx + d, d) = 1; p.u[idx + d] = init[d]; p.l[idx + d] = init[d]; }
  idx += 3;
  for (int i = 1; i < M; i++) {
    for (int d = 0; d < 3; d++) { A(idx + d, i * 15 + d) = 1; A(idx + d, i * 15 - 3 + d) = -1; }
    idx += 3;
  }
  for (int d = 0; d < 3; d++) { A(idx + d, M * 15 - 3 + d) = 1; p.u[idx + d] = goal[d]; p.l[idx + d] = goal[d]; }
  idx += 3;
  for (int d = 0; d < 3; d++) { A(idx + d, d) = -N_; A(idx + d, 3 + d) = N_; p.u[idx + d] = init[3 + d] * t0; p.l[idx + d] = init[3 + d] * t0; }
  idx += 3;
  for (int i = 1; i < M; i++) {
    const double t1 = t[i], t1_ = t[i - 1];
    for (int d = 0; d < 3; d++) {
      A(idx + d, i * 15 + d) = -N_ * 1.0 / t1;
      A(idx + d, i * 15 + 3 + d) = N_ * 1.0 / t1;
      A(idx + d, i * 15 - 3 + d) = -N_ * 1.0 / t1_;
      A(idx + d, i * 15 - 6 + d) = N_ * 1.0 / t1_;
    }
    idx += 3;
  }
  for (int d = 0; d < 3; d++) { A(idx + d, M * 15 - 6 + d) = -N_; A(idx + d, M * 15 - 3 + d) = N_; p.u[idx + d] = goal[3 + d] * tM; p.l[idx + d] = goal[3 + d] * tM; }
  idx += 3;
  for (int d = 0; d < 3; d++) {
    for (int c = 0; c < 9; c++) A(idx + d, c) = p2a[d][c];
    p.u[idx + d] = init[6 + d] * t0 * t0; p.l[idx + d] = init[6 + d] * t0 * t0;
  }
  idx += 3;
  for (int i = 1; i < M; i++) {
    const double t2 = std::pow(t[i], 2), t2_ = std::pow(t[i - 1], 2);
    for (int d = 0; d < 3; d++)
      for (int c = 0; c < 9; c++) { A(idx + d, i * 15 + c) = p2a[d][c] / t2; A(idx + d, i * 15 - 9 + c) = -p2a[d][c] / t2_; }
    idx += 3;
  }
  for (int d = 0; d < 3; d++) {
    for (int c = 0; c < 9; c++) A(idx + d, M * 15 - 9 + c) = p2a[d][c];
    p.u[idx + d] = goal[6 + d] * tM * tM; p.l[idx + d] = goal[6 + d] * tM * tM;
  }
  idx += 3;
  // ---- addDynamicalConstraints (:209-240) ----
  for (int i = 0; i < M; i++)
    for (int j = 0; j < N_; j++) {
      for (int d = 0; d < 3; d++) {
        for (int c = 0; c < 6; c++) A(idx + d, i * 15 + j * 3 + c) = p2v[d][c];
        p.u[idx + d] = vmax * 1.0 * t[i]; p.l[idx + d] = -vmax * 1.0 * t[i];
      }
      idx += 3;
    }
  for (int i = 0; i < M; i++)
    for (int j = 0; j < N_ - 1; j++) {
      for (int d = 0; d < 3; d++) {
        for (int c = 0; c < 9; c++) A(idx + d, i * 15 + j * 3 + c) = p2a[d][c];
        p.u[idx + d] = amax * 1.0 * t[i] * t[i]; p.l[idx + d] = -amax * 1.0 * t[i] * t[i];
      }
      idx += 3;
    }
  // ---- addSafetyConstraints (:242-260) ----
  for (int i = 0; i < M; i++)
    for (int j = 0; j < polys[i].rows(); j++) {
      const double* h = &polys[i].h[static_cast<size_t>(j) * 4];
      for (int k = 0; k < N_ + 1; k++) {
        for (int d = 0; d < 3; d++) A(idx, i * 15 + k * 3 + d) = h[d];
        p.u[idx] = -h[3];
        p.l[idx] = -OSQP_INFTY;
        idx++;
      }
    }
}

// ---- OSQP 0.6 ADMM ---------------------------------------------------------------------------------
namespace {

inline double limit_scaling(double v) {
  v = v < MIN_SCALING ? 1.0 : v;
  v = v > MAX_SCALING ? MAX_SCALING : v;
  return v;
}
inline double norm_inf(const std::vector<double>& v) {
  double m = 0;
  for (double x : v) m = std::max(m, std::fabs(x));
  return m;
}

struct Work {
  int n, m;
  std::vector<double> P, A, q, l, u;          // scaled data (P full symmetric dense)
  std::vector<double> D, E, Dinv, Einv;
  double c = 1.0, cinv = 1.0;
  std::vector<double> rho_vec, rho_inv_vec;
  std::vector<int> constr_type;
  double rho;
  std::vector<double> K;                       // Cholesky factor of P + sigma I + A' R A
};

void set_rho_vec(Work& w) {
  w.rho = std::min(std::max(w.rho, RHO_MIN), RHO_MAX);
  for (int i = 0; i < w.m; i++) {
    if (w.l[i] < -OSQP_INFTY * MIN_SCALING && w.u[i] > OSQP_INFTY * MIN_SCALING) {
      w.constr_type[i] = -1;
      w.rho_vec[i] = RHO_MIN;
    } else if (w.u[i] - w.l[i] < RHO_TOL) {
      w.constr_type[i] = 1;
      w.rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w.rho;
    } else {
      w.constr_type[i] = 0;
      w.rho_vec[i] = w.rho;
    }
    w.rho_inv_vec[i] = 1. / w.rho_vec[i];
  }
}

bool factorize(Work& w, double sigma) {
  const int n = w.n, m = w.m;
  w.K.assign(static_cast<size_t>(n) * n, 0.0);
  for (int a = 0; a < n; a++)
    for (int b = 0; b <= a; b++) w.K[static_cast<size_t>(a) * n + b] = w.P[static_cast<size_t>(a) * n + b] + (a == b ? sigma : 0.0);
  for (int i = 0; i < m; i++) {
    const double* row = &w.A[static_cast<size_t>(i) * n];
    const double r = w.rho_vec[i];
    for (int a = 0; a < n; a++) {
      if (row[a] == 0.0) continue;
      const double ra = r * row[a];
      for (int b = 0; b <= a; b++)
        if (row[b] != 0.0) w.K[static_cast<size_t>(a) * n + b] += ra * row[b];
    }
  }
  for (int j = 0; j < n; j++) {
    double d = w.K[static_cast<size_t>(j) * n + j];
    for (int k = 0; k < j; k++) d -= w.K[static_cast<size_t>(j) * n + k] * w.K[static_cast<size_t>(j) * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    w.K[static_cast<size_t>(j) * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = w.K[static_cast<size_t>(i) * n + j];
      for (int k = 0; k < j; k++) s -= w.K[static_cast<size_t>(i) * n + k] * w.K[static_cast<size_t>(j) * n + k];
      w.K[static_cast<size_t>(i) * n + j] = s / d;
    }
  }
  return true;
}

void chol_solve(const Work& w, std::vector<double>& b) {
  const int n = w.n;
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= w.K[static_cast<size_t>(i) * n + k] * b[k];
    b[i] = s / w.K[static_cast<size_t>(i) * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < n; k++) s -= w.K[static_cast<size_t>(k) * n + i] * b[k];
    b[i] = s / w.K[static_cast<size_t>(i) * n + i];
  }
}

void mat_vec(const std::vector<double>& Mx, int rows, int cols, const std::vector<double>& x, std::vector<double>& y) {
  for (int i = 0; i < rows; i++) {
    double s = 0;
    const double* r = &Mx[static_cast<size_t>(i) * cols];
    for (int j = 0; j < cols; j++) s += r[j] * x[j];
    y[i] = s;
  }
}
void mat_tvec(const std::vector<double>& Mx, int rows, int cols, const std::vector<double>& x, std::vector<double>& y) {
  std::fill(y.begin(), y.end(), 0.0);
  for (int i = 0; i < rows; i++) {
    const double* r = &Mx[static_cast<size_t>(i) * cols];
    const double xi = x[i];
    if (xi == 0.0) continue;
    for (int j = 0; j < cols; j++) y[j] += r[j] * xi;
  }
}

// scale_data (osqp/src/scaling.c)
void scale_data(Work& w, int iters) {
  const int n = w.n, m = w.m;
  w.D.assign(n, 1.0); w.E.assign(m, 1.0); w.c = 1.0;
  std::vector<double> Dt(n), Et(m);
  for (int it = 0; it < iters; it++) {
    for (int j = 0; j < n; j++) {
      double cp = 0, ca = 0;
      for (int i = 0; i < n; i++) cp = std::max(cp, std::fabs(w.P[static_cast<size_t>(i) * n + j]));
      for (int i = 0; i < m; i++) ca = std::max(ca, std::fabs(w.A[static_cast<size_t>(i) * n + j]));
      Dt[j] = std::max(cp, ca);
    }
    for (int i = 0; i < m; i++) {
      double r = 0;
      for (int j = 0; j < n; j++) r = std::max(r, std::fabs(w.A[static_cast<size_t>(i) * n + j]));
      Et[i] = r;
    }
    for (int j = 0; j < n; j++) Dt[j] = 1.0 / std::sqrt(limit_scaling(Dt[j]));
    for (int i = 0; i < m; i++) Et[i] = 1.0 / std::sqrt(limit_scaling(Et[i]));
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) w.P[static_cast<size_t>(i) * n + j] = (Dt[i] * w.P[static_cast<size_t>(i) * n + j]) * Dt[j];
    for (int i = 0; i < m; i++)
      for (int j = 0; j < n; j++) w.A[static_cast<size_t>(i) * n + j] = (Et[i] * w.A[static_cast<size_t>(i) * n + j]) * Dt[j];
    for (int j = 0; j < n; j++) { w.q[j] *= Dt[j]; w.D[j] *= Dt[j]; }
    for (int i = 0; i < m; i++) w.E[i] *= Et[i];
    // cost normalisation
    double mean = 0;
    for (int j = 0; j < n; j++) {
      double cp = 0;
      for (int i = 0; i < n; i++) cp = std::max(cp, std::fabs(w.P[static_cast<size_t>(i) * n + j]));
      mean += cp;
    }
    mean /= n;
    double qn = limit_scaling(norm_inf(w.q));
    double ct = limit_scaling(std::max(mean, qn));
    ct = 1. / ct;
    for (double& v : w.P) v *= ct;
    for (double& v : w.q) v *= ct;
    w.c *= ct;
  }
  w.Dinv.resize(n); w.Einv.resize(m);
  for (int j = 0; j < n; j++) w.Dinv[j] = 1.0 / w.D[j];
  for (int i = 0; i < m; i++) w.Einv[i] = 1.0 / w.E[i];
  w.cinv = 1.0 / w.c;
  for (int i = 0; i < m; i++) { w.l[i] *= w.E[i]; w.u[i] *= w.E[i]; }
}

}  // namespace

OsqpResult osqp_solve(const QpProblem& p, const OsqpSettings& S) {
  OsqpResult res;
  Work w;
  const int n = w.n = p.n, m = w.m = p.m;
  w.P = p.Q; w.A = p.A; w.q.assign(n, 0.0); w.l = p.l; w.u = p.u;
  // osqp_setup clips infinite bounds to +-OSQP_INFTY
  for (int i = 0; i < m; i++) { w.l[i] = std::max(w.l[i], -OSQP_INFTY); w.u[i] = std::min(w.u[i], OSQP_INFTY); }
  if (S.scaling) scale_data(w, S.scaling);
  else { w.D.assign(n, 1.0); w.E.assign(m, 1.0); w.Dinv = w.D; w.Einv = w.E; }
  w.rho = S.rho;
  w.rho_vec.resize(m); w.rho_inv_vec.resize(m); w.constr_type.resize(m);
  set_rho_vec(w);
  if (!factorize(w, S.sigma)) { res.status = -7; res.x.assign(n, 0.0); return res; }  // OSQP_NON_CVX

  std::vector<double> x(n, 0.0), z(m, 0.0), y(m, 0.0), x_prev(n), z_prev(m), xt(n), zt(m), delta_x(n), delta_y(m);
  std::vector<double> Ax(m), Px(n), Aty(n), tmpn(n), tmpm(m), rhs(n), resid_p(m), resid_d(n);
  double pri_res = 0, dua_res = 0;
  int status = 0;  // OSQP_UNSOLVED

  auto update_info = [&]() {
    mat_vec(w.A, m, n, x, Ax);
    for (int i = 0; i < m; i++) resid_p[i] = Ax[i] - z[i];
    double r = 0;
    for (int i = 0; i < m; i++) r = std::max(r, std::fabs(w.Einv[i] * resid_p[i]));
    pri_res = r;
    mat_vec(w.P, n, n, x, Px);
    mat_tvec(w.A, m, n, y, Aty);
    for (int j = 0; j < n; j++) resid_d[j] = (w.q[j] + Px[j]) + Aty[j];
    r = 0;
    for (int j = 0; j < n; j++) r = std::max(r, std::fabs(w.Dinv[j] * resid_d[j]));
    dua_res = w.cinv * r;
  };
  auto primal_infeasible = [&](double eps) {
    for (int i = 0; i < m; i++) {
      if (w.u[i] > OSQP_INFTY * MIN_SCALING) {
        if (w.l[i] < -OSQP_INFTY * MIN_SCALING) delta_y[i] = 0.0;
        else delta_y[i] = std::min(delta_y[i], 0.0);
      } else if (w.l[i] < -OSQP_INFTY * MIN_SCALING) {
        delta_y[i] = std::max(delta_y[i], 0.0);
      }
    }
    double nd = 0;
    for (int i = 0; i < m; i++) nd = std::max(nd, std::fabs(w.E[i] * delta_y[i]));
    if (nd > eps) {
      double lhs = 0;
      for (int i = 0; i < m; i++) lhs += w.u[i] * std::max(delta_y[i], 0.0) + w.l[i] * std::min(delta_y[i], 0.0);
      if (lhs < -eps * nd) {
        mat_tvec(w.A, m, n, delta_y, tmpn);
        double r = 0;
        for (int j = 0; j < n; j++) r = std::max(r, std::fabs(w.Dinv[j] * tmpn[j]));
        return r < eps * nd;
      }
    }
    return false;
  };
  auto check_termination = [&](bool approximate) {
    double eps_abs = S.eps_abs, eps_rel = S.eps_rel, eps_pi = S.eps_prim_inf;
    if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pi *= 10; }
    if (pri_res > OSQP_INFTY || dua_res > OSQP_INFTY) { status = -7; return true; }
    bool prim_ok = false, dual_ok = false, prim_inf = false;
    if (m == 0) prim_ok = true;
    else {
      double a = 0, b = 0;
      for (int i = 0; i < m; i++) { a = std::max(a, std::fabs(w.Einv[i] * z[i])); b = std::max(b, std::fabs(w.Einv[i] * Ax[i])); }
      const double eps_prim = eps_abs + eps_rel * std::max(a, b);
      if (pri_res < eps_prim) prim_ok = true;
      else prim_inf = primal_infeasible(eps_pi);
    }
    {
      double a = 0, b = 0, cq = 0;
      for (int j = 0; j < n; j++) {
        cq = std::max(cq, std::fabs(w.Dinv[j] * w.q[j]));
        a = std::max(a, std::fabs(w.Dinv[j] * Aty[j]));
        b = std::max(b, std::fabs(w.Dinv[j] * Px[j]));
      }
      const double eps_dual = eps_abs + eps_rel * (w.cinv * std::max(std::max(cq, a), b));
      if (dua_res < eps_dual) dual_ok = true;
      // dual infeasibility needs q' dx < 0; q = 0 for this problem family → never triggers
    }
    if (prim_ok && dual_ok) { status = approximate ? 2 : 1; return true; }
    if (prim_inf) { status = approximate ? 3 : -3; return true; }
    return false;
  };

  int iter;
  bool can_check = false;
  for (iter = 1; iter <= S.max_iter; iter++) {
    x_prev.swap(x);
    z_prev.swap(z);
    // update_xz_tilde: solve [P + sigma I, A'; A, -1/rho] [xt; nu] = [sigma x_prev - q; z_prev - y/rho]
    for (int i = 0; i < m; i++) tmpm[i] = w.rho_vec[i] * z_prev[i] - y[i];
    mat_tvec(w.A, m, n, tmpm, rhs);
    for (int j = 0; j < n; j++) rhs[j] = (S.sigma * x_prev[j] - w.q[j]) + rhs[j];
    chol_solve(w, rhs);
    xt = rhs;
    mat_vec(w.A, m, n, xt, zt);  // z_tilde = A x_tilde
    for (int j = 0; j < n; j++) { x[j] = S.alpha * xt[j] + (1.0 - S.alpha) * x_prev[j]; delta_x[j] = x[j] - x_prev[j]; }
    for (int i = 0; i < m; i++) {
      double v = (S.alpha * zt[i] + (1.0 - S.alpha) * z_prev[i]) + w.rho_inv_vec[i] * y[i];
      z[i] = std::min(std::max(v, w.l[i]), w.u[i]);
    }
    for (int i = 0; i < m; i++) {
      delta_y[i] = w.rho_vec[i] * ((S.alpha * zt[i] + (1.0 - S.alpha) * z_prev[i]) - z[i]);
      y[i] += delta_y[i];
    }
    can_check = S.check_termination && (iter % S.check_termination == 0);
    if (can_check) {
      update_info();
      if (check_termination(false)) break;
    }
    if (S.adaptive_rho_interval && (iter % S.adaptive_rho_interval == 0)) {
      if (!can_check) update_info();
      // compute_rho_estimate: scaled residuals, normalised
      double pr = norm_inf(resid_p), dr = norm_inf(resid_d);
      double pn = std::max(norm_inf(z), norm_inf(Ax));
      pr /= (pn + 1e-10);
      double dn = std::max(std::max(norm_inf(w.q), norm_inf(Aty)), norm_inf(Px));
      dr /= (dn + 1e-10);
      double rho_new = w.rho * std::sqrt(pr / (dr + 1e-10));
      rho_new = std::min(std::max(rho_new, RHO_MIN), RHO_MAX);
      if (getenv("ORC_QP_DEBUG")) fprintf(stderr, "iter %d pri %.3e dua %.3e rho %.3e rho_new %.3e\n", iter, pri_res, dua_res, w.rho, rho_new);
      if (rho_new > w.rho * S.adaptive_rho_tolerance || rho_new < w.rho / S.adaptive_rho_tolerance) {
        w.rho = rho_new;
        for (int i = 0; i < m; i++) {
          if (w.constr_type[i] == 0) w.rho_vec[i] = w.rho;
          else if (w.constr_type[i] == 1) w.rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w.rho;
          w.rho_inv_vec[i] = 1. / w.rho_vec[i];
        }
        factorize(w, S.sigma);
      }
    }
  }
  if (status == 0) {
    if (!can_check) { update_info(); check_termination(false); }
  }
  if (status == 0) {
    if (!check_termination(true)) status = -2;  // OSQP_MAX_ITER_REACHED
  }
  res.iters = std::min(iter, S.max_iter);
  res.status = status;
  res.x.resize(n);
  for (int j = 0; j < n; j++) res.x[j] = w.D[j] * x[j];  // unscale_solution
  return res;
}

}  // namespace orc

// o_firi.cpp — oracle: safe-flight-corridor stage of a replan (TEST INFRASTRUCTURE).
// Follows plan_manager/include/sfc_gen/firi.hpp:44-365 (chol3d, smoothedL1, costMVIE,
// maxVolInsEllipsoid, firi), plan_manager/include/sfc_gen/lbfgs.hpp (Lewis-Overton line search
// :96-190, lbfgs_optimize :192-430) and the corridor part of
// plan_manager/src/baseline_fake.cpp:127-228 (helpers) and :296-411 (stage).
// Deviations: Eigen::JacobiSVD of the 3x3 factor is replaced by a cyclic Jacobi eigen-solve of
// L L^T (same ellipsoid, axis signs may differ); LP shuffles are seeded per call (see o_lp.cpp).
// PINNED against the reference's own firi.hpp / lbfgs.hpp / sdlp.hpp compiled unmodified into oracle/_ref/libref.so
// (Eigen replaced by a stand-in, oracle/ref_build/Eigen/Eigen): costMVIE and whole L-BFGS runs bit-identical, the
// first FIRI polytope identical, maxVolInsEllipsoid the same optimum to the optimiser's own stopping tolerance,
// iterated corridors compared as sets — tests/test_ref_pins.py.  The corridor stage around firi is pinned through the
// whole replan of both reference planners (see o_replan.cpp).
#include <algorithm>
#include <cfloat>
#include <limits>

#include "oracle.hpp"

namespace orc {
namespace {

inline uint64_t lp_seed(uint64_t base, int segment, int kind) {
  uint64_t z = base * 0x9e3779b97f4a7c15ull + static_cast<uint64_t>(segment) * 16 + kind + 1;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// ---- MVIE cost (firi.hpp:73-142) ---------------------------------------------------------------
struct MvieData {
  int M;
  double smooth_eps, penalty;
  std::vector<double> A;  // M x 3 row-major
};

inline bool smoothed_l1(double mu, double x, double& f, double& df) {
  if (x < 0.0) return false;
  if (x > mu) { f = x - 0.5 * mu; df = 1.0; return true; }
  const double xdmu = x / mu, sqr = xdmu * xdmu, mumxd2 = mu - 0.5 * x;
  f = mumxd2 * sqr * xdmu;
  df = sqr * ((-0.5) * xdmu + 3.0 * mumxd2 / mu);
  return true;
}

double cost_mvie(const MvieData& D, const double* x, double* grad) {
  const double* p = x; const double* rtd = x + 3; const double* cde = x + 6;
  double gdp[3] = {0, 0, 0}, gdrtd[3] = {0, 0, 0}, gdcde[3] = {0, 0, 0};
  double cost = 0;
  const double L00 = rtd[0] * rtd[0] + DBL_EPSILON, L10 = cde[0], L11 = rtd[1] * rtd[1] + DBL_EPSILON,
               L20 = cde[2], L21 = cde[1], L22 = rtd[2] * rtd[2] + DBL_EPSILON;
  for (int i = 0; i < D.M; ++i) {
    const double* a = &D.A[i * 3];
    const double al0 = (a[0] * L00 + a[1] * L10) + a[2] * L20;
    const double al1 = (a[0] * 0.0 + a[1] * L11) + a[2] * L21;
    const double al2 = (a[0] * 0.0 + a[1] * 0.0) + a[2] * L22;
    const double nrm = std::sqrt((al0 * al0 + al1 * al1) + al2 * al2);
    const double adj0 = al0 / nrm, adj1 = al1 / nrm, adj2 = al2 / nrm;
    const double viol = (nrm + ((a[0] * p[0] + a[1] * p[1]) + a[2] * p[2])) - 1.0;
    double c, dc;
    if (smoothed_l1(D.smooth_eps, viol, c, dc)) {
      cost += c;
      const double v0 = dc * a[0], v1 = dc * a[1], v2 = dc * a[2];
      gdp[0] += v0; gdp[1] += v1; gdp[2] += v2;
      gdrtd[0] += adj0 * v0; gdrtd[1] += adj1 * v1; gdrtd[2] += adj2 * v2;
      gdcde[0] += adj0 * v1;
      gdcde[1] += adj1 * v2;
      gdcde[2] += adj0 * v2;
    }
  }
  cost *= D.penalty;
  for (int k = 0; k < 3; k++) { gdp[k] *= D.penalty; gdrtd[k] *= D.penalty; gdcde[k] *= D.penalty; }
  cost -= std::log(L00) + std::log(L11) + std::log(L22);
  gdrtd[0] -= 1.0 / L00; gdrtd[1] -= 1.0 / L11; gdrtd[2] -= 1.0 / L22;
  gdrtd[0] *= 2.0 * rtd[0]; gdrtd[1] *= 2.0 * rtd[1]; gdrtd[2] *= 2.0 * rtd[2];
  for (int k = 0; k < 3; k++) { grad[k] = gdp[k]; grad[3 + k] = gdrtd[k]; grad[6 + k] = gdcde[k]; }
  return cost;
}

// ---- L-BFGS with Lewis-Overton line search, n = 9 (lbfgs.hpp) -----------------------------------
constexpr int N9 = 9;
struct LbfgsParams {
  int mem_size = 18, past = 3, max_iterations = 0, max_linesearch = 64;
  double g_epsilon = 0.0, delta = 1.0e-7, min_step = 1.0e-32, max_step = 1.0e+20;
  double f_dec_coeff = 1.0e-4, s_curv_coeff = 0.9, cautious_factor = 1.0e-6, machine_prec = 1.0e-16;
};
inline double dot9(const double* a, const double* b) {
  double s = 0;
  for (int i = 0; i < N9; i++) s += a[i] * b[i];
  return s;
}
inline double maxabs9(const double* a) {
  double m = 0;
  for (int i = 0; i < N9; i++) m = std::max(m, std::fabs(a[i]));
  return m;
}

int line_search(const MvieData& D, double* x, double& f, double* g, double& stp, const double* s,
                const double* xp, const double* gp, double stpmin, double stpmax, const LbfgsParams& P) {
  int count = 0;
  bool brackt = false, touched = false;
  double mu = 0.0, nu = stpmax;
  if (!(stp > 0.0)) return -1;
  const double dginit = dot9(gp, s);
  if (0.0 < dginit) return -2;
  const double finit = f, dgtest = P.f_dec_coeff * dginit, dstest = P.s_curv_coeff * dginit;
  while (true) {
    for (int i = 0; i < N9; i++) x[i] = xp[i] + stp * s[i];
    f = cost_mvie(D, x, g);
    ++count;
    if (std::isinf(f) || std::isnan(f)) return -3;
    if (f > finit + stp * dgtest) {
      nu = stp;
      brackt = true;
    } else {
      if (dot9(g, s) < dstest) mu = stp;
      else return count;
    }
    if (P.max_linesearch <= count) return -4;
    if (brackt && (nu - mu) < P.machine_prec * nu) return -5;
    if (brackt) stp = 0.5 * (mu + nu);
    else stp *= 2.0;
    if (stp < stpmin) return -6;
    if (stp > stpmax) {
      if (touched) return -7;
      touched = true;
      stp = stpmax;
    }
  }
}

int lbfgs9(const MvieData& D, double* x, double& fout, const LbfgsParams& P, int* iters_out) {
  const int m = P.mem_size;
  double xp[N9], g[N9], gp[N9], d[N9], pf[8];
  std::vector<double> lm_alpha(m, 0.0), lm_s(static_cast<size_t>(N9) * m, 0.0), lm_y(static_cast<size_t>(N9) * m, 0.0),
      lm_ys(m, 0.0);
  int ret = 0, k = 1, end = 0, bound = 0;
  double fx = cost_mvie(D, x, g);
  pf[0] = fx;
  for (int i = 0; i < N9; i++) d[i] = -g[i];
  double gnorm = maxabs9(g), xnorm = maxabs9(x);
  if (gnorm / std::max(1.0, xnorm) < P.g_epsilon) {
    ret = 0;
  } else {
    double step = 1.0 / std::sqrt(dot9(d, d));
    while (true) {
      for (int i = 0; i < N9; i++) { xp[i] = x[i]; gp[i] = g[i]; }
      int ls = line_search(D, x, fx, g, step, d, xp, gp, P.min_step, P.max_step, P);
      if (ls < 0) {
        for (int i = 0; i < N9; i++) { x[i] = xp[i]; g[i] = gp[i]; }
        ret = ls;
        break;
      }
      gnorm = maxabs9(g); xnorm = maxabs9(x);
      if (gnorm / std::max(1.0, xnorm) < P.g_epsilon) { ret = 0; break; }
      if (0 < P.past) {
        if (P.past <= k) {
          const double rate = std::fabs(pf[k % P.past] - fx) / std::max(1.0, std::fabs(fx));
          if (rate < P.delta) { ret = 1; break; }
        }
        pf[k % P.past] = fx;
      }
      if (P.max_iterations != 0 && P.max_iterations <= k) { ret = -8; break; }
      ++k;
      double* sc = &lm_s[static_cast<size_t>(end) * N9];
      double* yc = &lm_y[static_cast<size_t>(end) * N9];
      for (int i = 0; i < N9; i++) { sc[i] = x[i] - xp[i]; yc[i] = g[i] - gp[i]; }
      const double ys = dot9(yc, sc), yy = dot9(yc, yc);
      lm_ys[end] = ys;
      for (int i = 0; i < N9; i++) d[i] = -g[i];
      const double cau = dot9(sc, sc) * std::sqrt(dot9(gp, gp)) * P.cautious_factor;
      if (ys > cau) {
        ++bound;
        bound = m < bound ? m : bound;
        end = (end + 1) % m;
        int j = end;
        for (int i = 0; i < bound; ++i) {
          j = (j + m - 1) % m;
          lm_alpha[j] = dot9(&lm_s[static_cast<size_t>(j) * N9], d) / lm_ys[j];
          for (int q = 0; q < N9; q++) d[q] += (-lm_alpha[j]) * lm_y[static_cast<size_t>(j) * N9 + q];
        }
        for (int q = 0; q < N9; q++) d[q] *= ys / yy;
        for (int i = 0; i < bound; ++i) {
          const double beta = dot9(&lm_y[static_cast<size_t>(j) * N9], d) / lm_ys[j];
          for (int q = 0; q < N9; q++) d[q] += (lm_alpha[j] - beta) * lm_s[static_cast<size_t>(j) * N9 + q];
          j = (j + 1) % m;
        }
      }
      step = 1.0;
    }
  }
  fout = fx;
  if (iters_out) *iters_out = k;
  return ret;
}

// ---- small dense helpers --------------------------------------------------------------------------
void chol3(const double A[9], double L[9]) {  // firi.hpp:44-55, row-major 3x3
  L[0] = std::sqrt(A[0]); L[1] = 0; L[2] = 0;
  L[3] = 0.5 * (A[1] + A[3]) / L[0];
  L[4] = std::sqrt(A[4] - L[3] * L[3]); L[5] = 0;
  L[6] = 0.5 * (A[2] + A[6]) / L[0];
  L[7] = (0.5 * (A[5] + A[7]) - L[6] * L[3]) / L[4];
  L[8] = std::sqrt(A[8] - L[6] * L[6] - L[7] * L[7]);
}

// U, S of a 3x3 matrix via cyclic Jacobi on B = L L^T; singular values sorted descending
void svd3_us(const double Lm[9], double U[9], double S[3]) {
  double B[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += Lm[i * 3 + k] * Lm[j * 3 + k];
      B[i * 3 + j] = s;
    }
  double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int sweep = 0; sweep < 60; sweep++) {
    const double off = std::fabs(B[1]) + std::fabs(B[2]) + std::fabs(B[5]);
    const double diag = std::fabs(B[0]) + std::fabs(B[4]) + std::fabs(B[8]);
    if (off <= 1e-300 || off <= 1e-17 * diag) break;
    const int pq[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    for (int e = 0; e < 3; e++) {
      const int p = pq[e][0], q = pq[e][1];
      const double apq = B[p * 3 + q];
      if (apq == 0.0) continue;
      const double theta = (B[q * 3 + q] - B[p * 3 + p]) / (2.0 * apq);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
      const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; k++) {  // B <- B J
        const double bkp = B[k * 3 + p], bkq = B[k * 3 + q];
        B[k * 3 + p] = c * bkp - s * bkq;
        B[k * 3 + q] = s * bkp + c * bkq;
      }
      for (int k = 0; k < 3; k++) {  // B <- J^T B
        const double bpk = B[p * 3 + k], bqk = B[q * 3 + k];
        B[p * 3 + k] = c * bpk - s * bqk;
        B[q * 3 + k] = s * bpk + c * bqk;
      }
      for (int k = 0; k < 3; k++) {  // V <- V J
        const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
        V[k * 3 + p] = c * vkp - s * vkq;
        V[k * 3 + q] = s * vkp + c * vkq;
      }
    }
  }
  int order[3] = {0, 1, 2};
  double ev[3] = {B[0], B[4], B[8]};
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2 - i; j++)
      if (ev[order[j]] < ev[order[j + 1]]) std::swap(order[j], order[j + 1]);
  for (int j = 0; j < 3; j++) {
    S[j] = std::sqrt(std::max(ev[order[j]], 0.0));
    for (int k = 0; k < 3; k++) U[k * 3 + j] = V[k * 3 + order[j]];
  }
}

double det3(const double M[9]) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// maxVolInsEllipsoid (firi.hpp:146-236); R row-major 3x3
bool max_vol_ins_ellipsoid(const std::vector<double>& hPoly, double R[9], double p[3], double r[3],
                           uint64_t* rng) {
  const int M = static_cast<int>(hPoly.size() / 4);
  std::vector<double> Alp(static_cast<size_t>(M) * 4), blp(M);
  for (int i = 0; i < M; i++) {
    const double* h = &hPoly[i * 4];
    const double hn = std::sqrt((h[0] * h[0] + h[1] * h[1]) + h[2] * h[2]);
    for (int k = 0; k < 3; k++) Alp[i * 4 + k] = h[k] / hn;
    Alp[i * 4 + 3] = 1.0;
    blp[i] = -h[3] / hn;
  }
  const double clp[4] = {0, 0, 0, -1.0};
  double xlp[4];
  const double maxdepth = -linprog(4, clp, Alp.data(), blp.data(), M, xlp, rng);
  if (!(maxdepth > 0.0) || std::isinf(maxdepth)) return false;
  const double interior[3] = {xlp[0], xlp[1], xlp[2]};
  MvieData D;
  D.M = M;
  D.smooth_eps = 1.0e-2;
  D.penalty = 1.0e+3;
  D.A.resize(static_cast<size_t>(M) * 3);
  for (int i = 0; i < M; i++) {
    const double* a = &Alp[i * 4];
    const double den = blp[i] - ((a[0] * interior[0] + a[1] * interior[1]) + a[2] * interior[2]);
    for (int k = 0; k < 3; k++) D.A[i * 3 + k] = a[k] / den;
  }
  double RD[9], Q[9], L[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) RD[i * 3 + j] = R[i * 3 + j] * (r[j] * r[j]);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += RD[i * 3 + k] * R[j * 3 + k];
      Q[i * 3 + j] = s;
    }
  chol3(Q, L);
  double x[9];
  for (int k = 0; k < 3; k++) x[k] = p[k] - interior[k];
  x[3] = std::sqrt(L[0]); x[4] = std::sqrt(L[4]); x[5] = std::sqrt(L[8]);
  x[6] = L[3]; x[7] = L[7]; x[8] = L[6];
  double min_cost;
  LbfgsParams P;
  const int ret = lbfgs9(D, x, min_cost, P, nullptr);
  for (int k = 0; k < 3; k++) p[k] = x[k] + interior[k];
  const double Lf[9] = {x[3] * x[3], 0, 0, x[6], x[4] * x[4], 0, x[8], x[7], x[5] * x[5]};
  double U[9], S[3];
  svd3_us(Lf, U, S);
  if (det3(U) < 0.0) {
    for (int k = 0; k < 3; k++) { R[k * 3 + 0] = U[k * 3 + 1]; R[k * 3 + 1] = U[k * 3 + 0]; R[k * 3 + 2] = U[k * 3 + 2]; }
    r[0] = S[1]; r[1] = S[0]; r[2] = S[2];
  } else {
    for (int k = 0; k < 9; k++) R[k] = U[k];
    for (int k = 0; k < 3; k++) r[k] = S[k];
  }
  return ret >= 0;
}

inline double dot3(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline double norm3(const double* a) { return std::sqrt(dot3(a, a)); }

}  // namespace

// firi::firi (firi.hpp:238-365); bd = 6 x 4 row-major
bool firi(const double bd[24], const std::vector<V3d>& pc, V3d a_, V3d b_, Poly& out, double r[3],
          int iterations, uint64_t* rng) {
  const double epsilon = 1.0e-6;
  const double a[3] = {a_.x, a_.y, a_.z}, b[3] = {b_.x, b_.y, b_.z};
  const int M = 6, N = static_cast<int>(pc.size());
  for (int i = 0; i < M; i++) {
    if (dot3(&bd[i * 4], a) + bd[i * 4 + 3] > 0.0) return false;   // (bd * ah).maxCoeff() > 0
  }
  for (int i = 0; i < M; i++) {
    if (dot3(&bd[i * 4], b) + bd[i * 4 + 3] > 0.0) return false;
  }
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double p[3] = {0.5 * (a[0] + b[0]), 0.5 * (a[1] + b[1]), 0.5 * (a[2] + b[2])};
  std::vector<double> forwardH(static_cast<size_t>(M + N) * 4);
  std::vector<double> fpc(static_cast<size_t>(N) * 3), tang(static_cast<size_t>(N) * 4), distRs(N);
  std::vector<uint8_t> pcFlags(N);
  int nH = 0;
  for (int loop = 0; loop < iterations; ++loop) {
    double fwd[9], bwd[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        fwd[i * 3 + j] = (1.0 / r[i]) * R[j * 3 + i];
        bwd[i * 3 + j] = R[i * 3 + j] * r[j];
      }
    double fB[18], fD[6], distDs[6];
    for (int i = 0; i < M; i++) {
      for (int j = 0; j < 3; j++)
        fB[i * 3 + j] = (bd[i * 4] * bwd[j] + bd[i * 4 + 1] * bwd[3 + j]) + bd[i * 4 + 2] * bwd[6 + j];
      fD[i] = bd[i * 4 + 3] + dot3(&bd[i * 4], p);
      distDs[i] = std::fabs(fD[i]) / norm3(&fB[i * 3]);
    }
    auto fwd_mul = [&](const double* v, double* o) {
      for (int i = 0; i < 3; i++) o[i] = (fwd[i * 3] * v[0] + fwd[i * 3 + 1] * v[1]) + fwd[i * 3 + 2] * v[2];
    };
    double fa[3], fb[3], tmp[3];
    for (int k = 0; k < 3; k++) tmp[k] = a[k] - p[k];
    fwd_mul(tmp, fa);
    for (int k = 0; k < 3; k++) tmp[k] = b[k] - p[k];
    fwd_mul(tmp, fb);
    for (int i = 0; i < N; i++) {
      const double v[3] = {pc[i].x - p[0], pc[i].y - p[1], pc[i].z - p[2]};
      double* q = &fpc[i * 3];
      fwd_mul(v, q);
      double* tg = &tang[i * 4];
      distRs[i] = norm3(q);
      tg[3] = -distRs[i];
      for (int k = 0; k < 3; k++) tg[k] = q[k] / distRs[i];
      if (dot3(tg, fa) + tg[3] > epsilon) {
        const double delta[3] = {q[0] - fa[0], q[1] - fa[1], q[2] - fa[2]};
        const double coef = dot3(delta, fa) / dot3(delta, delta);
        for (int k = 0; k < 3; k++) tg[k] = fa[k] - coef * delta[k];
        distRs[i] = norm3(tg);
        tg[3] = -distRs[i];
        for (int k = 0; k < 3; k++) tg[k] /= distRs[i];
      }
      if (dot3(tg, fb) + tg[3] > epsilon) {
        const double delta[3] = {q[0] - fb[0], q[1] - fb[1], q[2] - fb[2]};
        const double coef = dot3(delta, fb) / dot3(delta, delta);
        for (int k = 0; k < 3; k++) tg[k] = fb[k] - coef * delta[k];
        distRs[i] = norm3(tg);
        tg[3] = -distRs[i];
        for (int k = 0; k < 3; k++) tg[k] /= distRs[i];
      }
      if (dot3(tg, fa) + tg[3] > epsilon) {
        const double u[3] = {fa[0] - q[0], fa[1] - q[1], fa[2] - q[2]};
        const double w[3] = {fb[0] - q[0], fb[1] - q[1], fb[2] - q[2]};
        double c[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
        const double n2 = dot3(c, c);
        if (n2 > 0.0) { const double n = std::sqrt(n2); for (int k = 0; k < 3; k++) c[k] /= n; }
        for (int k = 0; k < 3; k++) tg[k] = c[k];
        tg[3] = -dot3(tg, fa);
        const double sgn = tg[3] > 0.0 ? -1.0 : 1.0;
        for (int k = 0; k < 4; k++) tg[k] *= sgn;
      }
    }
    uint8_t bdFlags[6] = {1, 1, 1, 1, 1, 1};
    std::fill(pcFlags.begin(), pcFlags.end(), 1);
    nH = 0;
    bool completed = false;
    int bdMinId = 0, pcMinId = 0;
    double minSqrD = distDs[0];
    for (int j = 1; j < M; j++) if (distDs[j] < minSqrD) { minSqrD = distDs[j]; bdMinId = j; }
    double minSqrR = INFINITY;
    if (N != 0) {
      minSqrR = distRs[0];
      for (int j = 1; j < N; j++) if (distRs[j] < minSqrR) { minSqrR = distRs[j]; pcMinId = j; }
    }
    for (int i = 0; !completed && i < (M + N); ++i) {
      double* fh = &forwardH[static_cast<size_t>(nH) * 4];
      if (minSqrD < minSqrR) {
        for (int k = 0; k < 3; k++) fh[k] = fB[bdMinId * 3 + k];
        fh[3] = fD[bdMinId];
        bdFlags[bdMinId] = 0;
      } else {
        for (int k = 0; k < 4; k++) fh[k] = tang[pcMinId * 4 + k];
        pcFlags[pcMinId] = 0;
      }
      completed = true;
      minSqrD = INFINITY;
      for (int j = 0; j < M; ++j)
        if (bdFlags[j]) {
          completed = false;
          if (minSqrD > distDs[j]) { bdMinId = j; minSqrD = distDs[j]; }
        }
      minSqrR = INFINITY;
      for (int j = 0; j < N; ++j)
        if (pcFlags[j]) {
          if (dot3(fh, &fpc[j * 3]) + fh[3] > -epsilon) {
            pcFlags[j] = 0;
          } else {
            completed = false;
            if (minSqrR > distRs[j]) { pcMinId = j; minSqrR = distRs[j]; }
          }
        }
      ++nH;
    }
    out.h.assign(static_cast<size_t>(nH) * 4, 0.0);
    for (int i = 0; i < nH; ++i) {
      const double* fh = &forwardH[static_cast<size_t>(i) * 4];
      double* h = &out.h[static_cast<size_t>(i) * 4];
      for (int j = 0; j < 3; j++) h[j] = (fh[0] * fwd[j] + fh[1] * fwd[3 + j]) + fh[2] * fwd[6 + j];
      h[3] = fh[3] - dot3(h, p);
    }
    if (loop == iterations - 1) break;
    max_vol_ins_ellipsoid(out.h, R, p, r, rng);
  }
  return true;
}

namespace {

// checkCorridorValidity (baseline_fake.cpp:191-204): feasibility LP with zero objective
bool corridor_valid(const std::vector<double>& h, uint64_t* rng) {
  const int m = static_cast<int>(h.size() / 4);
  std::vector<double> A(static_cast<size_t>(m) * 3), b(m);
  for (int i = 0; i < m; i++) { for (int k = 0; k < 3; k++) A[i * 3 + k] = h[i * 4 + k]; b[i] = -h[i * 4 + 3]; }
  const double c[3] = {0, 0, 0};
  double x[3];
  const double rst = linprog(3, c, A.data(), b.data(), m, x, rng);
  return !std::isinf(rst);
}

// ShrinkCorridor(corridor, path): the simulator planner leaves front / back and top / bottom faces alone
// (baseline_fake.cpp:215-228); the real-perception planner shrinks every face (baseline.cpp:213-226)
void shrink(std::vector<double>& h, const double path[3], double shrink_size, bool keep_end_faces) {
  const int m = static_cast<int>(h.size() / 4);
  for (int i = 0; i < m; i++) {
    const double A = h[i * 4], B = h[i * 4 + 1], C = h[i * 4 + 2];
    const double n[3] = {A, B, C}, z[3] = {0, 0, 1};
    if (keep_end_faces && dot3(n, path) / norm3(n) / norm3(path) > 0.8) continue;
    if (keep_end_faces && std::abs(dot3(n, z)) / norm3(n) > 0.8) continue;
    h[i * 4 + 3] += std::sqrt(A * A + B * B + C * C) * shrink_size;
  }
}

// checkGoalReachability (baseline_fake.cpp:143-182)
bool goal_reachable(const std::vector<double>& h, const double start[3], double goal[3], uint64_t seed_base,
                    int corridor_idx, int kind0 = 3) {
  const int m = static_cast<int>(h.size() / 4);
  if (m <= 0) return true;
  double mx = -INFINITY;
  for (int i = 0; i < m; i++) {
    const double v = ((h[i * 4] * goal[0] + h[i * 4 + 1] * goal[1]) + h[i * 4 + 2] * goal[2]) + h[i * 4 + 3] * 1.0;
    mx = std::max(mx, v);
  }
  if (mx <= 0) return true;
  std::vector<double> A(static_cast<size_t>(m) * 3), b(m);
  for (int i = 0; i < m; i++) { for (int k = 0; k < 3; k++) A[i * 3 + k] = h[i * 4 + k]; b[i] = -h[i * 4 + 3]; }
  double c[3] = {-goal[0] + start[0], -goal[1] + start[1], -goal[2] + start[2]};
  double gmax[3], gmin[3];
  uint64_t r1 = lp_seed(seed_base, corridor_idx, kind0), r2 = lp_seed(seed_base, corridor_idx, kind0 + 1);
  linprog(3, c, A.data(), b.data(), m, gmax, &r1);
  for (int k = 0; k < 3; k++) c[k] = goal[k] - start[k];
  linprog(3, c, A.data(), b.data(), m, gmin, &r2);
  for (int k = 0; k < 3; k++) goal[k] = 0.5 * (gmax[k] + gmin[k]);
  return false;
}

// The tail of the corridor stage: adjacent corridors must intersect, enough corridors must be left, the local goal is
// pulled back to the last corridor that contains it.  The reference's two planners differ here; `real` selects
// BaselinePlanner::replan (baseline.cpp:363-403) against FakeBaselinePlanner::replan (baseline_fake.cpp:366-411):
//   non-intersecting corridors i, i+1 (i >= 2): the real planner also drops corridor i (:371 against :374);
//   it fails with at most one corridor left (:381 against "none", :384);
//   it tests the last corridor once before the backward search (:389) — and that test moves the goal when it fails.
// polys holds the valid prefix of corridors; returns false when the replan fails here.
bool corridor_tail(bool real, uint64_t seed_base, std::vector<Poly>& polys, const std::vector<double>& path,
                   const double start[9], double goal[6]) {
  // the reference underflows `size() - 1` on an empty list
  for (int i = 0; i + 1 < static_cast<int>(polys.size()); i++) {
    std::vector<double> both = polys[i].h;
    both.insert(both.end(), polys[i + 1].h.begin(), polys[i + 1].h.end());
    uint64_t rng_i = lp_seed(seed_base, i, 2);
    if (!corridor_valid(both, &rng_i)) {
      if (i < 2) { polys.clear(); return false; }
      polys.erase(polys.begin() + (real ? i : i + 1), polys.end());
      break;
    }
  }
  if (polys.size() <= (real ? 1u : 0u)) { polys.clear(); return false; }
  int gi = static_cast<int>(polys.size()) - 1;
  double goal_pos[3] = {path[gi * 6], path[gi * 6 + 1], path[gi * 6 + 2]};
  double goal_vel[3] = {path[gi * 6 + 3], path[gi * 6 + 4], path[gi * 6 + 5]};
  bool search = true;
  if (real) search = !goal_reachable(polys.back().h, start, goal_pos, seed_base, gi, 5);
  for (int it = static_cast<int>(polys.size()) - 1; search && it != 0; it--) {
    if (goal_reachable(polys[it].h, start, goal_pos, seed_base, it)) {
      polys.erase(polys.begin() + it + 1, polys.end());
      const int idx = static_cast<int>(polys.size()) - 1;
      for (int k = 0; k < 3; k++) { goal_pos[k] = path[idx * 6 + k]; goal_vel[k] = path[idx * 6 + 3 + k]; }
      break;
    }
  }
  for (int k = 0; k < 3; k++) { goal[k] = goal_pos[k]; goal[3 + k] = goal_vel[k]; }
  return true;
}

}  // namespace

// Corridor stage of FakeBaselinePlanner::replan (baseline_fake.cpp:296-411) or, on a RiskBase-style map
// (pop_config.query_variant), of BaselinePlanner::replan (baseline.cpp:296-403): the same code up to the shrink rule
// and the tail rules above.
// path = n x 6 from getPathWithVel; the LP seed base is pop_config.agent_id_base (callers set it
// to the agent's global id).
CorridorResult corridor_stage(const Grid& g, const pop_config& c, const std::vector<double>& path,
                              const double start[9], double traj_start_time) {
  CorridorResult res;
  const bool real = g.query_variant != POP_QUERY_FAKE;
  const uint64_t seed_base = static_cast<uint64_t>(c.agent_id_base);
  const int n = static_cast<int>(path.size() / 6);
  std::vector<V3d> wpts(n);
  for (int i = 0; i < n; i++) {
    wpts[i] = V3d{path[i * 6], path[i * 6 + 1], path[i * 6 + 2]};
    if (wpts[i].z < 0) wpts[i].z = 0.1;
  }
  V3d lower{-4 + start[0], -4 + start[1], -1 + start[2]}, higher{4 + start[0], 4 + start[1], 1 + start[2]};
  if (lower.z < 0) lower.z = 0;
  if (higher.z > 4) higher.z = 4;
  std::vector<V3d> pc;
  for (int i = 0; i < n - 1; i++) {
    V3d lhc, llc;
    lhc.x = std::min(std::max(wpts[i].x, wpts[i + 1].x) + c.init_range, higher.x);
    lhc.y = std::min(std::max(wpts[i].y, wpts[i + 1].y) + c.init_range, higher.y);
    lhc.z = std::min(std::max(wpts[i].z, wpts[i + 1].z) + c.init_range, higher.z);
    llc.x = std::max(std::min(wpts[i].x, wpts[i + 1].x) - c.init_range, lower.x);
    llc.y = std::max(std::min(wpts[i].y, wpts[i + 1].y) - c.init_range, lower.y);
    llc.z = std::max(std::min(wpts[i].z, wpts[i + 1].z) - c.init_range, lower.z);
    const double bd[24] = {1, 0, 0, -lhc.x, 0, 1, 0, -lhc.y, 0, 0, 1, -lhc.z,
                           -1, 0, 0, llc.x, 0, -1, 0, llc.y, 0, 0, -1, llc.z};
    pc.clear();
    const double t1 = traj_start_time + i * c.corridor_tau;
    const double t2 = traj_start_time + (i + 1) * c.corridor_tau;
    obstacle_points(g, t1, t2, llc, lhc, pc);
    Poly poly;
    double r[3] = {1, 1, 1};
    uint64_t rng = lp_seed(seed_base, i, 0);
    firi(bd, pc, wpts[i], wpts[i + 1], poly, r, 2, &rng);
    const double dir[3] = {wpts[i + 1].x - wpts[i].x, wpts[i + 1].y - wpts[i].y, wpts[i + 1].z - wpts[i].z};
    shrink(poly.h, dir, c.shrink_size, !real);
    uint64_t rng_v = lp_seed(seed_base, i, 1);
    if (!corridor_valid(poly.h, &rng_v)) break;
    res.polys.push_back(poly);
  }
  if (!corridor_tail(real, seed_base, res.polys, path, start, res.goal)) return res;
  res.ok = true;
  return res;
}

}  // namespace orc

// test hooks: the MVIE cost at one point and maxVolInsEllipsoid alone (tests/test_ref_pins.py compares them with
// the reference's own firi.hpp compiled into oracle/_ref/libref.so)
extern "C" double orc_cost_mvie(const double* A, int M, double smooth_eps, double penalty_wt, const double* x, double* grad) {
  orc::MvieData D;
  D.M = M; D.smooth_eps = smooth_eps; D.penalty = penalty_wt;
  D.A.assign(A, A + static_cast<size_t>(M) * 3);
  return orc::cost_mvie(D, x, grad);
}
extern "C" int orc_mvie(const double* hpoly, int M, double* R, double* p, double* r, uint64_t seed) {
  std::vector<double> h(hpoly, hpoly + static_cast<size_t>(M) * 4);
  uint64_t s = seed;
  return orc::max_vol_ins_ellipsoid(h, R, p, r, &s) ? 1 : 0;
}
extern "C" int orc_lbfgs_mvie(const double* A, int M, double smooth_eps, double penalty_wt, double* x, double* fx) {
  orc::MvieData D;
  D.M = M; D.smooth_eps = smooth_eps; D.penalty = penalty_wt;
  D.A.assign(A, A + static_cast<size_t>(M) * 3);
  orc::LbfgsParams P;
  return orc::lbfgs9(D, x, *fx, P, nullptr);
}
// the tail alone on given corridors (hpoly POP_MAX_PIECES x POP_MAX_FACES x 4, rows, valid per segment): returns the
// number kept (0 = fail); tests/test_device_logic_on_host.py compares it with the device source compiled for the host
extern "C" int orc_corridor_tail(int variant, uint64_t seed_base, int n, const double* path, const double* hpoly,
                                 const int32_t* rows, const int32_t* valid, const double* start9, double* out_goal) {
  std::vector<orc::Poly> polys;
  for (int i = 0; i < n - 1 && i < POP_MAX_PIECES; i++) {
    if (!valid[i]) break;
    orc::Poly p;
    p.h.assign(hpoly + static_cast<size_t>(i) * POP_MAX_FACES * 4, hpoly + (static_cast<size_t>(i) * POP_MAX_FACES + rows[i]) * 4);
    polys.push_back(p);
  }
  std::vector<double> pv(path, path + static_cast<size_t>(n) * 6);
  if (!orc::corridor_tail(variant != POP_QUERY_FAKE, seed_base, polys, pv, start9, out_goal)) return 0;
  return static_cast<int>(polys.size());
}
// ShrinkCorridor(corridor, path) alone, in place
extern "C" void orc_shrink(int variant, double* h, int m, const double* dir, double shrink_size) {
  std::vector<double> v(h, h + static_cast<size_t>(m) * 4);
  orc::shrink(v, dir, shrink_size, variant == POP_QUERY_FAKE);
  std::copy(v.begin(), v.end(), h);
}

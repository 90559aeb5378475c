// ref_glue.cpp — C entry points into the REFERENCE's own code (TEST INFRASTRUCTURE; builds oracle/_ref/libref.so).
//
// Compiled together with the reference files listed in oracle/Makefile (raycast.cpp, bernstein.cpp, bezier_optimizer.cpp,
// both hybrid A* sources, particles.cpp, and the headers firi.hpp / lbfgs.hpp / sdlp.hpp) — unmodified, from where
// they lie — against oracle/ref_build/Eigen/Eigen (a stand-in for the absent Eigen; see that file for what it does
// and does not reproduce).  Nothing here restates reference arithmetic: each function only marshals flat arrays
// into the reference's argument types and back.  Used by tests/test_ref_pins.py to pin oracle/o_gridmap.cpp's
// DDA, oracle/o_lp.cpp and oracle/o_firi.cpp, and by scripts/make_golden_ref.py to write tests/golden/ref_pins.npz.
// Never loaded by the product (pred_occ_planner_b200/) — tests/test_abi.py checks that.
#include <cstdint>
#include <cstring>

#include <plan_env/raycast.h>
#include "sfc_gen/firi.hpp"
#include "traj_utils/bernstein.hpp"
#include "bernstein/bezier_optimizer.hpp"   // finds oracle/ref_build/iosqp.hpp (a recorder, not a solver) before the reference's

#include <iostream>
#include <sstream>

// the hybrid A*: its private counters (use_node_num_, iter_num_) are read back below
#define private public
#define protected public
#include <path_searching/fake_risk_hybrid_a_star.h>
#undef private
#undef protected
#include "ref_astar_driver.hpp"

// path_searching/src/dyn_a_star.cpp (the base class's own grid A*, not on the path, needs the real GridMap) is not
// compiled; its out-of-line members are stubbed so that the vtable links
AStar::~AStar() {}
void AStar::reset() {}
void AStar::init() {}
void AStar::init(const Eigen::Vector3d&, const Eigen::Vector3i&) {}
void AStar::setParam(ros::NodeHandle&) {}
void AStar::setEnvironment(const GridMap::Ptr&) {}
void AStar::initEnvironment(GridMap::Ptr, const Eigen::Vector3i) {}
double AStar::getDiagHeu(GridNodePtr, GridNodePtr) { return 0.0; }
double AStar::getManhHeu(GridNodePtr, GridNodePtr) { return 0.0; }
double AStar::getEuclHeu(GridNodePtr, GridNodePtr) { return 0.0; }
ASTAR_RET AStar::search(Eigen::Vector3d, Eigen::Vector3d) { return NO_PATH; }
std::vector<Eigen::Vector3d> AStar::getPath() { return {}; }

// sdlp::linprog<d>: min c.x s.t. A x <= b, A row-major m x d (sdlp.hpp:708-785)
template <int d>
static double lp(const double* c, const double* A, const double* b, int m, double* x) {
  Eigen::Matrix<double, d, 1> cv, xv;
  Eigen::Matrix<double, -1, d> Am(m, d);
  Eigen::Matrix<double, -1, 1> bv(m);
  for (int k = 0; k < d; k++) cv(k) = c[k];
  for (int i = 0; i < m; i++) {
    bv(i) = b[i];
    for (int k = 0; k < d; k++) Am(i, k) = A[i * d + k];
  }
  const double v = sdlp::linprog<d>(cv, Am, bv, xv);
  for (int k = 0; k < d; k++) x[k] = xv(k);
  return v;
}

extern "C" {

// RayCaster as GridMap::raycastProcess drives it (grid_map.cpp:379-381): setInput(start, end), then every ray_pt for
// which step() returned true.  Returns the number of such points (written up to cap), *set_ok = setInput's result.
int ref_raycaster(const double* start, const double* end, double* out, int cap, int* set_ok) {
  RayCaster rc;
  Eigen::Vector3d s(start[0], start[1], start[2]), e(end[0], end[1], end[2]), pt;
  const bool ok = rc.setInput(s, e);
  if (set_ok) *set_ok = ok ? 1 : 0;
  int n = 0;
  while (rc.step(pt)) {
    if (n < cap) { out[3 * n] = pt(0); out[3 * n + 1] = pt(1); out[3 * n + 2] = pt(2); }
    if (++n > (1 << 22)) break;   // the reference loop has no bound; inputs of the tests never get here
  }
  return n;
}

// sdlp::linprog<d> (sdlp.hpp:708-785)
double ref_linprog(int d, const double* c, const double* A, const double* b, int m, double* x) {
  switch (d) {
    case 2: return lp<2>(c, A, b, m, x);
    case 3: return lp<3>(c, A, b, m, x);
    case 4: return lp<4>(c, A, b, m, x);
    default: return NAN;
  }
}

// firi::firi (firi.hpp:216-365): bd M x 4 and pc N x 3 row-major; r in/out (the caller's initial radii);
// hPoly rows written to out (cap rows); returns the row count, or 0 when firi returns false
int ref_firi(const double* bd, int M, const double* pc, int N, const double* a, const double* b, int iterations,
             double* r, double* out, int cap) {
  Eigen::MatrixX4d bdm(M, 4), h;
  Eigen::Matrix3Xd pcm(3, N);
  for (int i = 0; i < M; i++) for (int k = 0; k < 4; k++) bdm(i, k) = bd[4 * i + k];
  for (int i = 0; i < N; i++) for (int k = 0; k < 3; k++) pcm(k, i) = pc[3 * i + k];
  Eigen::Vector3d av(a[0], a[1], a[2]), bv(b[0], b[1], b[2]), rv(r[0], r[1], r[2]);
  if (!firi::firi(bdm, pcm, av, bv, h, rv, iterations)) return 0;
  for (int k = 0; k < 3; k++) r[k] = rv(k);
  for (int i = 0; i < h.rows() && i < cap; i++) for (int k = 0; k < 4; k++) out[4 * i + k] = h(i, k);
  return h.rows();
}

// firi::maxVolInsEllipsoid (firi.hpp:143-214): hPoly M x 4 row-major; R (row-major 3 x 3), p, r in/out
int ref_mvie(const double* hpoly, int M, double* R, double* p, double* r) {
  Eigen::MatrixX4d h(M, 4);
  for (int i = 0; i < M; i++) for (int k = 0; k < 4; k++) h(i, k) = hpoly[4 * i + k];
  Eigen::Matrix3d Rm;
  for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) Rm(i, k) = R[3 * i + k];
  Eigen::Vector3d pv(p[0], p[1], p[2]), rv(r[0], r[1], r[2]);
  const bool ok = firi::maxVolInsEllipsoid(h, Rm, pv, rv);
  for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) R[3 * i + k] = Rm(i, k);
  for (int k = 0; k < 3; k++) { p[k] = pv(k); r[k] = rv(k); }
  return ok ? 1 : 0;
}

// firi::costMVIE (firi.hpp:74-137) at one point: A M x 3 row-major, x[9] -> cost, grad[9]
double ref_cost_mvie(const double* A, int M, double smooth_eps, double penalty_wt, const double* x, double* grad) {
  std::vector<uint8_t> buf(sizeof(int) + (2 + 3 * static_cast<size_t>(M)) * sizeof(double));
  int* pM = reinterpret_cast<int*>(buf.data());
  double* pS = reinterpret_cast<double*>(pM + 1);
  *pM = M; pS[0] = smooth_eps; pS[1] = penalty_wt;
  for (int i = 0; i < M; i++) for (int k = 0; k < 3; k++) pS[2 + k * M + i] = A[3 * i + k];   // column-major M x 3
  Eigen::VectorXd xv(9), g(9);
  for (int k = 0; k < 9; k++) xv(k) = x[k];
  const double c = firi::costMVIE(buf.data(), xv, g);
  for (int k = 0; k < 9; k++) grad[k] = g(k);
  return c;
}

// lbfgs::lbfgs_optimize (lbfgs.hpp:434-) on firi::costMVIE with maxVolInsEllipsoid's parameter set (firi.hpp:186-196):
// A M x 3 row-major, x[9] in/out, *fx the final cost; returns lbfgs_optimize's status
int ref_lbfgs_mvie(const double* A, int M, double smooth_eps, double penalty_wt, double* x, double* fx) {
  std::vector<uint8_t> buf(sizeof(int) + (2 + 3 * static_cast<size_t>(M)) * sizeof(double));
  int* pM = reinterpret_cast<int*>(buf.data());
  double* pS = reinterpret_cast<double*>(pM + 1);
  *pM = M; pS[0] = smooth_eps; pS[1] = penalty_wt;
  for (int i = 0; i < M; i++) for (int k = 0; k < 3; k++) pS[2 + k * M + i] = A[3 * i + k];
  Eigen::VectorXd xv(9);
  for (int k = 0; k < 9; k++) xv(k) = x[k];
  lbfgs::lbfgs_parameter_t prm;
  prm.mem_size = 18; prm.g_epsilon = 0.0; prm.min_step = 1.0e-32; prm.past = 3; prm.delta = 1.0e-7;
  double f = 0.0;
  const int ret = lbfgs::lbfgs_optimize(xv, f, &firi::costMVIE, nullptr, nullptr, buf.data(), prm);
  for (int k = 0; k < 9; k++) x[k] = xv(k);
  *fx = f;
  return ret;
}

// Bernstein::Bezier (traj_utils/src/bernstein.cpp): dur[M], cpts[5M x 3] row-major; deriv 0/1/2 = getPos/getVel/getAcc
void ref_bezier_eval(const double* dur, int M, const double* cpts, const double* t, int n, int deriv, double* out) {
  Eigen::MatrixX3d c(5 * M, 3);
  for (int i = 0; i < 5 * M; i++) for (int k = 0; k < 3; k++) c(i, k) = cpts[3 * i + k];
  Bernstein::Bezier b(std::vector<double>(dur, dur + M), c);
  for (int i = 0; i < n; i++) {
    const Eigen::Vector3d p = deriv == 0 ? b.getPos(t[i]) : deriv == 1 ? b.getVel(t[i]) : b.getAcc(t[i]);
    for (int k = 0; k < 3; k++) out[3 * i + k] = p(k);
  }
}
// getMaxVelRate (order 1) / getMaxAccRate (order 2)
double ref_bezier_max_rate(const double* dur, int M, const double* cpts, int order) {
  Eigen::MatrixX3d c(5 * M, 3);
  for (int i = 0; i < 5 * M; i++) for (int k = 0; k < 3; k++) c(i, k) = cpts[3 * i + k];
  Bernstein::Bezier b(std::vector<double>(dur, dur + M), c);
  return order == 1 ? b.getMaxVelRate() : b.getMaxAccRate();
}

// traj_opt::BezierOpt::setup (bezier_optimizer.cpp:26-285): init9 / goal9 = [pos; vel; acc] row-major 3 x 3, dur[M],
// hpoly = M blocks of max_rows x 4 with rows[i] valid rows each.  Writes Q (n x n), A (m x n), lb, ub row-major when
// non-null (n = 15 M) and returns m.  The reference prints its row counts to stdout; that is swallowed here.
int ref_qp_setup(int M, const double* init9, const double* goal9, const double* dur, const double* hpoly, int max_rows,
                 const int32_t* rows, double vmax, double amax, double* outQ, double* outA, double* outl, double* outu) {
  Eigen::Matrix3d s, e;
  for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) { s(i, k) = init9[3 * i + k]; e(i, k) = goal9[3 * i + k]; }
  std::vector<traj_opt::PolyhedronH> cs;
  for (int i = 0; i < M; i++) {
    traj_opt::PolyhedronH h(rows[i], 4);
    for (int j = 0; j < rows[i]; j++) for (int k = 0; k < 4; k++) h(j, k) = hpoly[(static_cast<size_t>(i) * max_rows + j) * 4 + k];
    cs.push_back(h);
  }
  std::ostringstream sink;
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  traj_opt::BezierOpt opt;
  opt.setup(s, e, std::vector<double>(dur, dur + M), cs, vmax, amax);
  std::cout.rdbuf(old);
  const Eigen::MatrixXd Q = opt.getQ(), A = opt.getA(), lb = opt.getlb();
  const Eigen::VectorXd ub = opt.getb();
  const int n = Q.rows(), m = A.rows();
  if (outQ) for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) outQ[static_cast<size_t>(i) * n + j] = Q(i, j);
  if (outA) for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) outA[static_cast<size_t>(i) * n + j] = A(i, j);
  if (outl) for (int i = 0; i < m; i++) outl[i] = lb(i);
  if (outu) for (int i = 0; i < m; i++) outu[i] = ub(i);
  return m;
}

// FakeRiskHybridAstar on FakeParticleRiskVoxel (see ref_astar_driver.hpp); RiskHybridAstar is ref_astar_risk in ref_glue_risk.cpp
int ref_astar(const double* prm, FakeParticleRiskVoxel::query_dt_fn query, void* map, const float* pose3,
              const double* start9, const double* goal3, double t_start, double* out_path, int cap, int* out_n, int* stats) {
  return run_astar<FakeRiskHybridAstar, FakeParticleRiskVoxel>(prm, query, map, pose3, start9, goal3, t_start, out_path,
                                                               cap, out_n, stats);
}

}  // extern "C"

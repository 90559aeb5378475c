// iosqp.hpp — a STAND-IN for traj_opt/include/iosqp.hpp, the reference's thin wrapper over OSQP (third party, absent
// here).  TEST INFRASTRUCTURE (oracle/_ref only).  It lets traj_opt/src/bezier_optimizer.cpp compile unmodified so
// that the QP it ASSEMBLES (BezierOpt::setup -> Q, A, lb, ub) can be read back and compared with the oracle's
// assembly; it does not solve anything: setMats records the problem, solve() is a no-op, the status is "unsolved".
// The solver itself (OSQP 0.6) is restated in oracle/o_qp.cpp from its published algorithm and is NOT pinned by this.
#ifndef POP_IOSQP_STANDIN_H
#define POP_IOSQP_STANDIN_H
#include <Eigen/Eigen>
#include <vector>

typedef long long c_int;
typedef double c_float;
#ifndef OSQP_INFTY
#define OSQP_INFTY ((c_float)1e30)   // osqp/include/constants.h
#endif

// Optional solver behind the recorder: when set, solve() hands the recorded problem (dense, row-major) to it.
// ref_glue_planner.cpp points it at the oracle's OSQP restatement so that the reference's planner can run end to end.
typedef int (*pop_iosqp_solver)(int n, int m, const double* P, const double* q, const double* A, const double* l,
                                const double* u, double eps_abs, double eps_rel, double* x);
inline pop_iosqp_solver& pop_iosqp_solver_slot() { static pop_iosqp_solver f = nullptr; return f; }

class IOSQP {
 public:
  Eigen::MatrixXd P_, A_;
  Eigen::VectorXd q_, l_, u_;
  mutable Eigen::VectorXd x_;
  mutable c_int status_ = -10;   // OSQP_UNSOLVED
  double eps_abs_ = 0, eps_rel_ = 0;
  c_int setMats(Eigen::SparseMatrix<double>& P, Eigen::VectorXd& q, Eigen::SparseMatrix<double>& A, Eigen::VectorXd& l,
                Eigen::VectorXd& u, const double& eps_abs, const double& eps_rel) {
    P_ = P; q_ = q; A_ = A; l_ = l; u_ = u; eps_abs_ = eps_abs; eps_rel_ = eps_rel;
    return 0;
  }
  c_int solve() const {
    const int n = P_.rows(), m = A_.rows();
    x_ = Eigen::VectorXd::Zero(n);
    if (!pop_iosqp_solver_slot()) return 0;
    std::vector<double> P(static_cast<size_t>(n) * n), A(static_cast<size_t>(m) * n);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) P[static_cast<size_t>(i) * n + j] = P_(i, j);
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) A[static_cast<size_t>(i) * n + j] = A_(i, j);
    status_ = pop_iosqp_solver_slot()(n, m, P.data(), q_.data(), A.data(), l_.data(), u_.data(), eps_abs_, eps_rel_, x_.data());
    return 0;
  }
  c_int getStatus() const { return status_; }
  Eigen::VectorXd getPrimalSol() const { return x_; }
};
#endif

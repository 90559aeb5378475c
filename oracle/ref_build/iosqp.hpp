// iosqp.hpp — a STAND-IN for traj_opt/include/iosqp.hpp, the reference's thin wrapper over OSQP (third party, absent
// here).  TEST INFRASTRUCTURE (oracle/_ref only).  It lets traj_opt/src/bezier_optimizer.cpp compile unmodified so
// that the QP it ASSEMBLES (BezierOpt::setup -> Q, A, lb, ub) can be read back and compared with the oracle's
// assembly; it does not solve anything: setMats records the problem, solve() is a no-op, the status is "unsolved".
// The solver itself (OSQP 0.6) is restated in oracle/o_qp.cpp from its published algorithm and is NOT pinned by this.
#ifndef POP_IOSQP_STANDIN_H
#define POP_IOSQP_STANDIN_H
#include <Eigen/Eigen>

typedef long long c_int;
typedef double c_float;
#ifndef OSQP_INFTY
#define OSQP_INFTY ((c_float)1e30)   // osqp/include/constants.h
#endif

class IOSQP {
 public:
  Eigen::MatrixXd P_, A_;
  Eigen::VectorXd q_, l_, u_;
  double eps_abs_ = 0, eps_rel_ = 0;
  c_int setMats(Eigen::SparseMatrix<double>& P, Eigen::VectorXd& q, Eigen::SparseMatrix<double>& A, Eigen::VectorXd& l,
                Eigen::VectorXd& u, const double& eps_abs, const double& eps_rel) {
    P_ = P; q_ = q; A_ = A; l_ = l; u_ = u; eps_abs_ = eps_abs; eps_rel_ = eps_rel;
    return 0;
  }
  c_int solve() const { return 0; }
  c_int getStatus() const { return -10; }   // OSQP_UNSOLVED
  Eigen::VectorXd getPrimalSol() const { return Eigen::VectorXd::Zero(static_cast<int>(q_.size())); }
};
#endif

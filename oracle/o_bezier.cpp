// o_bezier.cpp — oracle: Bernstein/Bezier evaluation (TEST INFRASTRUCTURE).
// Follows traj_utils/src/bernstein.cpp:25-59,80-88,117-196 and
// traj_utils/include/traj_utils/bernstein.hpp:98-177 of the reference.
// PINNED twice: the reference's own known-answer tests (tests/test_golden_reference.py) and its bernstein.cpp
// compiled into oracle/_ref/libref.so (1e-14 relative, 99 % of samples identical; tests/test_ref_pins.py).
#include "oracle.hpp"

namespace orc {

// bernstein.cpp:80-88 (case 4)
const double kBernsteinA[25] = {1, -4, 6,  -4,  1,  0, 4, -12, 12, -4, 0, 0, 6,
                                -12, 6, 0, 0,   0,  4, -4, 0,  0,  0,  0, 1};

// BernsteinPiece::getPos/getVel/getAcc: p = C^T * A * S (left to right), S from pow(s, i)
V3d piece_eval(const double* c, double t0, double tf, double tt, int deriv) {
  const double dur = tf - t0;
  const double s = (tt - t0) / dur;
  double S[5];
  if (deriv == 0) {
    S[0] = 1;
    for (int i = 1; i <= 4; i++) S[i] = std::pow(s, i);
  } else if (deriv == 1) {
    S[0] = 0; S[1] = 1;
    for (int i = 2; i <= 4; i++) S[i] = i * std::pow(s, i - 1);
  } else {
    S[0] = 0; S[1] = 0; S[2] = 2;
    for (int i = 3; i <= 4; i++) S[i] = i * (i - 1) * std::pow(s, i - 2);
  }
  double out[3];
  for (int d = 0; d < 3; d++) {
    double CA[5];
    for (int j = 0; j < 5; j++) {
      double acc = 0;
      for (int i = 0; i < 5; i++) acc += c[i * 3 + d] * kBernsteinA[i * 5 + j];
      CA[j] = acc;
    }
    double acc = 0;
    for (int j = 0; j < 5; j++) acc += CA[j] * S[j];
    out[d] = acc;
  }
  if (deriv == 1) { for (double& o : out) o = o / dur; }
  if (deriv == 2) { const double d2 = std::pow(dur, 2); for (double& o : out) o = o / d2; }
  return V3d{out[0], out[1], out[2]};
}

Bezier::Bezier(const std::vector<double>& dur, const std::vector<double>& c) : t(dur), cpts(c) {
  M = static_cast<int>(t.size());
  T = 0;
  for (int i = 0; i < M; i++) T += t[i];
}

// bernstein.hpp:164-172
int Bezier::locate(double tt) const {
  for (int i = 0; i < M; i++) {
    tt -= t[i];
    if (tt < 0) return i;
  }
  return M - 1;
}

static void piece_span(const Bezier& b, int i, double& t0, double& tf) {
  double acc = 0;  // Bezier::calcPieces: t0 = running sum, tf = t0 + t_[i]
  for (int k = 0; k < i; k++) acc += b.t[k];
  t0 = acc;
  tf = acc + b.t[i];
}

V3d Bezier::pos(double tt) const {
  int i = locate(tt); double t0, tf; piece_span(*this, i, t0, tf);
  return piece_eval(&cpts[static_cast<size_t>(i) * 15], t0, tf, tt, 0);
}
V3d Bezier::vel(double tt) const {
  int i = locate(tt); double t0, tf; piece_span(*this, i, t0, tf);
  return piece_eval(&cpts[static_cast<size_t>(i) * 15], t0, tf, tt, 1);
}
V3d Bezier::acc(double tt) const {
  int i = locate(tt); double t0, tf; piece_span(*this, i, t0, tf);
  return piece_eval(&cpts[static_cast<size_t>(i) * 15], t0, tf, tt, 2);
}

// bernstein.cpp:131-157,198-219
static double max_rate(const Bezier& b, int order) {
  double best = 0;
  for (int p = 0; p < b.M; p++) {
    double t0, tf; piece_span(b, p, t0, tf);
    const double dur = tf - t0;
    double d[5][3];
    int n = 4;
    for (int i = 0; i < 5; i++) for (int k = 0; k < 3; k++) d[i][k] = b.cpts[(p * 5 + i) * 3 + k];
    for (int o = 0; o < order; o++) {
      for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) d[i][k] = n * (d[i + 1][k] - d[i][k]);
      n--;
    }
    double m = 0;
    for (int i = 0; i <= n; i++) {
      double r = std::sqrt(d[i][0] * d[i][0] + d[i][1] * d[i][1] + d[i][2] * d[i][2]);
      if (r > m) m = r;
    }
    double rate = order == 1 ? m / dur : m / std::pow(dur, 2);
    if (rate > best) best = rate;
  }
  return best;
}
double Bezier::max_vel_rate() const { return max_rate(*this, 1); }
double Bezier::max_acc_rate() const { return max_rate(*this, 2); }

}  // namespace orc

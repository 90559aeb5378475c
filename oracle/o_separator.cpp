// o_separator.cpp — oracle: linear separability of two point sets (TEST INFRASTRUCTURE).
// utils/separator/src/separator_glpk.cpp:73-246 poses the LP  n.a + d >= 1 (a in A),
// n.b + d <= -1 (b in B) with a zero objective and reports GLP_OPT/GLP_FEAS, i.e. pure
// feasibility; GLPK 4.65 (fetched at build time by utils/separator/cmake/glpk.cmake.in:6) is
// not available here, so feasibility is decided by the same Seidel LP used for corridors, inside
// the box |n_i|, |d| <= 1e4 (Seidel needs a bounded region).  GLPK's variables are free, so its answer does
// not depend on where the two sets sit; to keep that property inside a box the points are taken relative to
// A's first point (d then stays of the order of |n| x the sets' extent, and only |n| <= 1e4, i.e. a gap
// above 2e-4 m, can bind).  PARITY UNPINNED beyond the reference's own instance
// (utils/separator/src/test_separator.cpp:20-32, separable).
#include <cmath>

#include "oracle.hpp"

namespace orc {

bool separable(const std::vector<V3d>& A, const std::vector<V3d>& B) {
  const int m = static_cast<int>(A.size() + B.size()) + 8;
  std::vector<double> M(static_cast<size_t>(m) * 4), rhs(m);
  int r = 0;
  const V3d o = A.empty() ? V3d{0, 0, 0} : A[0];
  for (const V3d& a : A) { M[r * 4] = -(a.x - o.x); M[r * 4 + 1] = -(a.y - o.y); M[r * 4 + 2] = -(a.z - o.z); M[r * 4 + 3] = -1.0; rhs[r] = -1.0; r++; }
  for (const V3d& b : B) { M[r * 4] = b.x - o.x; M[r * 4 + 1] = b.y - o.y; M[r * 4 + 2] = b.z - o.z; M[r * 4 + 3] = 1.0; rhs[r] = -1.0; r++; }
  for (int k = 0; k < 4; k++) {
    for (int j = 0; j < 4; j++) { M[r * 4 + j] = j == k ? 1.0 : 0.0; M[(r + 1) * 4 + j] = j == k ? -1.0 : 0.0; }
    rhs[r] = 1e4; rhs[r + 1] = 1e4;
    r += 2;
  }
  const double c[4] = {0, 0, 0, 0};
  double x[4];
  uint64_t rng = 0x5eed5eedull + A.size() * 131 + B.size();
  const double v = linprog(4, c, M.data(), rhs.data(), m, x, &rng);
  return !std::isinf(v);
}

}  // namespace orc

// munkres.h — a STAND-IN for munkres-cpp (github.com/saebyn/munkres-cpp, installed as /usr/local/lib/libmunkres.a by the
// reference's plan_env/CMakeLists.txt:26-32; absent here): the two classes dsp_dynamic.h:1566-1589 uses.  TEST
// INFRASTRUCTURE (oracle/_ref only).  Munkres<T>::solve replaces a cost matrix by 0 at the assigned (row, column)
// pairs of a minimum-cost assignment and -1 elsewhere; a rectangular matrix is solved as if padded to square.  The
// assignment here is found by shortest augmenting paths (ties: lowest column), written for this repo — the optimum is
// unique up to ties, which is all the caller depends on.  The library itself is NOT pinned by this.
#ifndef POP_MUNKRES_STANDIN_H
#define POP_MUNKRES_STANDIN_H
#include <limits>
#include <vector>
template <class T>
class Matrix {
  std::vector<T> v_;
  size_t r_ = 0, c_ = 0;
 public:
  Matrix() {}
  Matrix(size_t rows, size_t columns) : v_(rows * columns, T(0)), r_(rows), c_(columns) {}
  T& operator()(size_t r, size_t c) { return v_[r * c_ + c]; }
  const T& operator()(size_t r, size_t c) const { return v_[r * c_ + c]; }
  size_t rows() const { return r_; }
  size_t columns() const { return c_; }
};
template <class T>
class Munkres {
 public:
  void solve(Matrix<T>& m) {
    const int nr = static_cast<int>(m.rows()), nc = static_cast<int>(m.columns());
    const bool tr = nr > nc;
    const int n = tr ? nc : nr, k = tr ? nr : nc;   // n <= k
    auto a = [&](int i, int j) { return static_cast<double>(tr ? m(j, i) : m(i, j)); };
    const double INF = std::numeric_limits<double>::infinity();
    std::vector<double> u(n + 1, 0.0), v(k + 1, 0.0);
    std::vector<int> p(k + 1, 0), way(k + 1, 0);
    for (int i = 1; i <= n; ++i) {
      p[0] = i;
      int j0 = 0;
      std::vector<double> minv(k + 1, INF);
      std::vector<char> used(k + 1, 0);
      do {
        used[j0] = 1;
        const int i0 = p[j0];
        double delta = INF;
        int j1 = 0;
        for (int j = 1; j <= k; ++j)
          if (!used[j]) {
            const double cur = a(i0 - 1, j - 1) - u[i0] - v[j];
            if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
            if (minv[j] < delta) { delta = minv[j]; j1 = j; }
          }
        for (int j = 0; j <= k; ++j)
          if (used[j]) { u[p[j]] += delta; v[j] -= delta; } else { minv[j] -= delta; }
        j0 = j1;
      } while (p[j0] != 0);
      do { const int j1 = way[j0]; p[j0] = p[j1]; j0 = j1; } while (j0);
    }
    for (int r = 0; r < nr; r++) for (int c = 0; c < nc; c++) m(r, c) = T(-1);
    for (int j = 1; j <= k; ++j)
      if (p[j] > 0) { if (tr) m(j - 1, p[j] - 1) = T(0); else m(p[j] - 1, j - 1) = T(0); }
  }
};
#endif

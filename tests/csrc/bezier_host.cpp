// bezier_host.cpp — TEST INFRASTRUCTURE: the product's device Bezier evaluation (pred_occ_planner_b200/csrc/
// pop_bezier.cuh) compiled for the host.  The CUDA rounding intrinsics it uses are IEEE round-to-nearest operations,
// which is what the plain operators are on the host with contraction off (-ffp-contract=off); __fma_rn is std::fma.
#include <cmath>
#include <cuda_runtime.h>

static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }

#include "../../pred_occ_planner_b200/csrc/pop_bezier.cuh"

extern "C" {
// dur[M], cpts 5M x 3, t[n] -> out n x 3 (deriv 0/1/2)
void dev_bezier_eval(const double* dur, int M, const double* cpts, const double* t, int n, int deriv, double* out) {
  for (int i = 0; i < n; i++) pop_bezier_eval(dur, M, cpts, t[i], deriv, out + 3 * i);
}
double dev_pow_int(double s, int n) { return pop_pow_int(s, n); }
}

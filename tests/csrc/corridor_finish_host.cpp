// corridor_finish_host.cpp — TEST INFRASTRUCTURE: compiles the PRODUCT's device source for the host.
// pred_occ_planner_b200/csrc/pop_lp.cuh (Seidel LP) and pop_corridor_finish.cuh (the per-agent tail of the corridor
// stage) are plain C++ behind __device__; with g++ (cuda_runtime.h makes the qualifiers empty outside nvcc) the very
// same source runs on the CPU, so tests/test_device_logic_on_host.py can check it where no GPU exists.
#include <cstdint>
#include <vector>

#include "../../pred_occ_planner_b200/csrc/pop_corridor_finish.cuh"

extern "C" {

// corridor_finish_agent with the work arrays k_corridor.cu gives it
int dev_corridor_finish(int variant, uint64_t seed_base, int n, const double* path, const double* hpoly, const int32_t* rows,
                        const int32_t* valid, const double* start9, double* out_goal) {
  constexpr int kLpRows = 2 * POP_MAX_FACES + 8;
  std::vector<double> ab(4 * kLpRows), lpw(poplp::work_doubles(kLpRows, 4));
  std::vector<int> scri(poplp::work_ints(kLpRows));
  return corridor_finish_agent(variant, seed_base, n, path, hpoly, rows, valid, start9, ab.data(), lpw.data(), scri.data(),
                               out_goal);
}

// poplp::linprog<d> alone: min c.x s.t. A x <= b (A row-major m x d)
double dev_linprog(int d, const double* c, const double* A, const double* b, int m, double* x, uint64_t seed) {
  std::vector<double> lpw(poplp::work_doubles(m + 8, 4));
  std::vector<int> scri(poplp::work_ints(m + 8));
  unsigned long long rng = seed;
  switch (d) {
    case 2: return poplp::linprog<2>(c, A, 2, b, 1, m, x, &rng, lpw.data(), scri.data());
    case 3: return poplp::linprog<3>(c, A, 3, b, 1, m, x, &rng, lpw.data(), scri.data());
    case 4: return poplp::linprog<4>(c, A, 4, b, 1, m, x, &rng, lpw.data(), scri.data());
    default: return NAN;
  }
}

}  // extern "C"

// heur_lat.cu — cycles of one estimateHeuristic (quartic + cubic + candidate costs) on its own, one warp, cold and warm
// instruction cache, against the ~7 000 cycles it takes inside an A* expansion.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false --expt-relaxed-constexpr -I../../include \
//        -I../../pred_occ_planner_b200/csrc heur_lat.cu -o heur_lat && ./heur_lat
#include <cstdio>
#define POP_ASTAR_DEVICE_ONLY
#include "../../pred_occ_planner_b200/csrc/k_astar.cu"

__global__ void k_heur(AstarParams P, double* out, long long* cyc, int reps) {
  double x1[6] = {0.3 + 0.01 * threadIdx.x, -0.2 + 0.02 * threadIdx.x, 1.0, 0.5, -0.1 * (threadIdx.x % 3), 0.0};
  double x2[6] = {8.0, 3.0 - 0.1 * threadIdx.x, 1.2, 0, 0, 0};
  double acc = 0.0;
  for (int r = 0; r < reps; r++) {
    const long long t0 = clock64();
    const double h = heuristic(P, x1, x2, nullptr);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[r] = t1 - t0;
    acc += h;
    x1[0] += 1e-3 * (h > 0 ? 1 : -1);   // the next call depends on this one
    x1[3] += 1e-4;
  }
  out[threadIdx.x] = acc;
}

int main() {
  AstarParams P{};
  P.max_tau = 0.6; P.max_vel = 3.0; P.max_acc = 3.0; P.w_time = 10.0; P.horizon = 7.0; P.lambda_heu = 5.0;
  P.tie_breaker = 1.0 + 1.0 / 10000;
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * sizeof(double)); cudaMalloc(&cyc, 64 * sizeof(long long));
  k_heur<<<1, 32>>>(P, out, cyc, 16);
  long long h[16];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("heuristic cycles per call (one warp, dependent calls):");
  for (int i = 0; i < 16; i++) printf(" %lld", h[i]);
  printf("\n");
  return 0;
}

// Micro-benchmark of the QP kernel's K^-1 application with 6x6 register tiles (384 threads, n = 150):
// cycles per call of the full step and of variants with parts removed, to see what bounds it.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kT = 384;
struct TileId { int rb, cb, NB, NBp; bool act; };

template <int MODE>   // 0 full, 1 no reduction loads, 2 no FMAs, 3 barriers only
__device__ __forceinline__ void kkt_tiles(double* rhs, double* xt_out, const double (&T)[36], const TileId& ti,
                                          double* __restrict__ partA, double* __restrict__ partB, int n) {
  const int tid = threadIdx.x;
  const int NB6 = 6 * ti.NB;
  if (MODE != 3 && ti.act) {
    const double2* xc = reinterpret_cast<const double2*>(rhs + 6 * ti.cb);
    const double2* xr = reinterpret_cast<const double2*>(rhs + 6 * ti.rb);
    const double2 c01 = xc[0], c23 = xc[1], c45 = xc[2];
    double* outA = partA + (6 * ti.rb) * ti.NBp + ti.cb;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      double s = T[i * 6] * c01.x;
      if (MODE != 2) {
        s = fma(T[i * 6 + 1], c01.y, s); s = fma(T[i * 6 + 2], c23.x, s); s = fma(T[i * 6 + 3], c23.y, s);
        s = fma(T[i * 6 + 4], c45.x, s); s = fma(T[i * 6 + 5], c45.y, s);
      }
      outA[i * ti.NBp] = s;
    }
    if (ti.rb != ti.cb) {
      const double2 r01 = xr[0], r23 = xr[1], r45 = xr[2];
      double o[6];
#pragma unroll
      for (int j = 0; j < 6; j++) {
        double s = T[j] * r01.x;
        if (MODE != 2) {
          s = fma(T[6 + j], r01.y, s); s = fma(T[12 + j], r23.x, s); s = fma(T[18 + j], r23.y, s);
          s = fma(T[24 + j], r45.x, s); s = fma(T[30 + j], r45.y, s);
        }
        o[j] = s;
      }
      double2* outB = reinterpret_cast<double2*>(partB + ti.rb * NB6 + 6 * ti.cb);
      outB[0] = make_double2(o[0], o[1]); outB[1] = make_double2(o[2], o[3]); outB[2] = make_double2(o[4], o[5]);
    }
  }
  __syncthreads();
  if (MODE != 3) {
    const int i = tid >> 1, side = tid & 1;
    double a = 0, b = 0;
    if (i < n && MODE != 1) {
      if (side == 0) {
        const double2* pr = reinterpret_cast<const double2*>(partA + i * ti.NBp);
        const int cnt = ti.NBp >> 1;
#pragma unroll 4
        for (int k = 0; k < cnt; k++) { const double2 v = pr[k]; a += v.x; b += v.y; }
      } else {
        const double* pc = partB + i;
        const int cnt = (ti.NB + 1) & ~1;
#pragma unroll 4
        for (int k = 0; k < cnt; k += 2) { a += pc[k * NB6]; b += pc[(k + 1) * NB6]; }
      }
    }
    double xt = a + b;
    xt += __shfl_xor_sync(0xffffffffu, xt, 1);
    if (i < n && side == 0) { xt_out[i] = xt; rhs[i] = 1e-3 * xt + 0.5; }
  }
  __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(kT, 1) k_bench(double* out, long long* cyc, int n, int iters) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* rhs = reinterpret_cast<double*>(smem);
  double* xt = rhs + 192;
  double* partA = xt + 192;
  TileId ti;
  ti.NB = (n + 5) / 6;
  ti.NBp = ((ti.NB + 1) & ~1) | 2;
  if (ti.NBp < ti.NB) ti.NBp += 4;
  double* partB = partA + 6 * ti.NB * ti.NBp;
  const int tid = threadIdx.x;
  int rb = static_cast<int>((sqrtf(8.0f * tid + 1.0f) - 1.0f) * 0.5f);
  while ((rb + 1) * (rb + 2) / 2 <= tid) rb++;
  while (rb * (rb + 1) / 2 > tid) rb--;
  ti.rb = rb; ti.cb = tid - rb * (rb + 1) / 2;
  ti.act = ti.rb < ti.NB;
  double T[36];
  for (int k = 0; k < 36; k++) T[k] = 1e-3 * (k + 1) + 1e-5 * tid;
  for (int e = tid; e < 6 * ti.NB * ti.NBp + ((ti.NB + 1) & ~1) * 6 * ti.NB; e += kT) partA[e] = 0.0;
  if (tid < 192) { rhs[tid] = 1.0 + tid * 1e-3; xt[tid] = 0; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) kkt_tiles<MODE>(rhs, xt, T, ti, partA, partB, n);
  const long long t1 = clock64();
  if (tid < n) out[tid] = xt[tid];
  if (tid == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
  const int smem = 150 * 1024, iters = 2000;
  cudaFuncSetAttribute(k_bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_bench<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int n : {150, 105}) {
    k_bench<0><<<1, kT, smem>>>(out, cyc, n, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("n=%d full            %.0f cycles/call\n", n, (double)h / iters);
    k_bench<1><<<1, kT, smem>>>(out, cyc, n, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("n=%d no reduce loads %.0f cycles/call\n", n, (double)h / iters);
    k_bench<2><<<1, kT, smem>>>(out, cyc, n, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("n=%d no FMAs         %.0f cycles/call\n", n, (double)h / iters);
    k_bench<3><<<1, kT, smem>>>(out, cyc, n, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("n=%d barriers only   %.0f cycles/call\n", n, (double)h / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Micro-benchmarks behind the QP kernel's design numbers: dependent-chain latency and per-SM
// throughput of fp64 FMA/ADD, shared-memory load latency, and __syncthreads cost at 384 threads.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dep(double* out, long long* cyc, int n) {
  double a = out[0], b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
  long long t1 = clock64();
  double d = out[1];
  long long t2 = clock64();
  for (int i = 0; i < n; i++) { d = d + c; d = d + b; d = d + c; d = d + b; }
  long long t3 = clock64();
  out[threadIdx.x] = a + d;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; }
}
__global__ void k_thr(double* out, long long* cyc, int n) {
  double a[8];
  for (int k = 0; k < 8; k++) a[k] = out[k] + threadIdx.x;
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++)
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = fma(a[k], b, c);
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < 8; k++) s += a[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc, int n) {
  __shared__ int idx[1024];
  __shared__ double val[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) { idx[i] = (i * 7 + 1) & 1023; val[i] = i; }
  __syncthreads();
  int j = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) j = idx[j];
  long long t1 = clock64();
  double s = 0;
  int q = threadIdx.x;
  long long t2 = clock64();
  for (int i = 0; i < n; i++) { s += val[q]; q = (q + 33) & 1023; }
  long long t3 = clock64();
  out[threadIdx.x] = j + s;
  long long t4 = clock64();
  for (int i = 0; i < n; i++) __syncthreads();
  long long t5 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; cyc[2] = t5 - t4; }
}
int main() {
  double* out; long long* cyc; long long h[4];
  cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 8192);
  const int n = 4096;
  k_dep<<<1, 32>>>(out, cyc, n); cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
  printf("dependent DFMA latency %.2f cycles, dependent DADD latency %.2f cycles\n", h[0] / (4.0 * n), h[1] / (4.0 * n));
  for (int threads : {32, 128, 384, 512, 1024}) {
    k_thr<<<1, threads>>>(out, cyc, n); cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    printf("%4d threads: %.2f DFMA lanes/clk/SM\n", threads, 8.0 * n * threads / h[0]);
  }
  k_lds<<<1, 384>>>(out, cyc, n); cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
  printf("dependent LDS latency %.2f cycles; LDS+DADD chain step %.2f cycles; __syncthreads (384 thr) %.2f cycles\n",
         h[0] / (double)n, h[1] / (double)n, h[2] / (double)n);
  k_lds<<<1, 32>>>(out, cyc, n); cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
  printf("1 warp: dependent LDS latency %.2f cycles; LDS+DADD chain step %.2f; __syncthreads %.2f cycles\n",
         h[0] / (double)n, h[1] / (double)n, h[2] / (double)n);
  return 0;
}

// Micro-probe: latency of a dependent DFMA / double-shuffle chain and DFMA throughput per SM (is fp64 a slow path on this part?).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a fp64_probe.cu -o fp64_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double *out, long long *clk, double a, double b, int n) {
  double x = a;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  double y = a;
  for (int i = 0; i < n; ++i) y = __shfl_sync(0xffffffffu, y, (threadIdx.x + 1) & 31);
  long long t2 = clock64();
  float z = (float)a;
  for (int i = 0; i < n; ++i) z = fmaf(z, (float)b, (float)a);
  long long t3 = clock64();
  if (threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; }
  out[threadIdx.x] = x + y + z;
}
__global__ void thr(double *out, long long *clk, double a, double b, int n) {
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
    x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[0] = t1 - t0;
  out[threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
  double *o; long long *c; cudaMalloc(&o, 8 * 1024); cudaMalloc(&c, 64);
  long long h[4];
  const int n = 4096;
  lat<<<1, 32>>>(o, c, 1.0, 0.999, n); lat<<<1, 32>>>(o, c, 1.0, 0.999, n);
  cudaDeviceSynchronize(); cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
  printf("dependent chain: DFMA %.1f clk, double shuffle %.1f clk, FFMA %.1f clk\n", (double)h[0] / n, (double)h[1] / n, (double)h[2] / n);
  for (int threads : {32, 128, 256, 512, 1024}) {
    thr<<<1, threads>>>(o, c, 1.0, 0.999, n); thr<<<1, threads>>>(o, c, 1.0, 0.999, n);
    cudaDeviceSynchronize(); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
    printf("%4d threads x 8 independent DFMA chains: %.2f clk per warp-DFMA per SM (%.1f DFMA lanes / clk / SM)\n", threads,
           (double)h[0] / ((double)n * 8 * (threads / 32)), (double)n * 8 * threads / (double)h[0]);
  }
  return 0;
}

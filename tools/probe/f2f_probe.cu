// Micro-probe: throughput of the fp32 -> fp64 conversion (F2F.F64.F32) per SM, alone and next to DFMA, and of a conversion built from
// integer operations (normal numbers and zeros only). The history pass of the direction converts every element it streams.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a f2f_probe.cu -o f2f_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double cvt_int(float f) { // exact for normal floats and +-0; denormals -> 0 (Inf / NaN not handled)
  const unsigned b = __float_as_uint(f);
  const unsigned mag = b & 0x7fffffffu;
  unsigned hi = (mag >> 3) + 0x38000000u;
  if (mag < 0x00800000u) hi = 0u;
  hi |= b & 0x80000000u;
  const unsigned lo = mag < 0x00800000u ? 0u : b << 29;
  return __hiloint2double((int)hi, (int)lo);
}
template <int MODE> // 0: FMUL + F2F + DFMA, 1: FMUL + DFMA (no conversion: the double is reused), 2: FMUL + integer conversion + DFMA, 3: FMUL + F2F + DADD-free (F2F only, xor-combined)
__global__ void thr(double *out, long long *clk, float a, double b, int n) {
  float f[8]; double x[8]; double keep[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { f[j] = a + j + threadIdx.x; x[j] = j; keep[j] = a + j; }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = f[j] * 1.0000001f;
      if (MODE == 0) x[j] = fma((double)f[j], b, x[j]);
      if (MODE == 1) { keep[j] = keep[j]; x[j] = fma(keep[j], b, x[j]); }
      if (MODE == 2) x[j] = fma(cvt_int(f[j]), b, x[j]);
      if (MODE == 3) { const double d = (double)f[j]; x[j] = __hiloint2double(__double2hiint(x[j]) ^ __double2hiint(d), __double2loint(x[j]) ^ __double2loint(d)); }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[0] = t1 - t0;
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j] + f[j];
  out[threadIdx.x] = s;
}
int main() {
  double *o; long long *c; cudaMalloc(&o, 8 * 1024); cudaMalloc(&c, 64);
  long long h[1];
  const int n = 4096;
  const char *names[4] = {"FMUL + F2F.F64.F32 + DFMA", "FMUL + DFMA", "FMUL + integer conversion + DFMA", "FMUL + F2F.F64.F32 + 2 LOP3"};
  for (int mode = 0; mode < 4; ++mode)
    for (int threads : {256, 512, 1024}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) thr<0><<<1, threads>>>(o, c, 1.0f, 0.999, n);
        if (mode == 1) thr<1><<<1, threads>>>(o, c, 1.0f, 0.999, n);
        if (mode == 2) thr<2><<<1, threads>>>(o, c, 1.0f, 0.999, n);
        if (mode == 3) thr<3><<<1, threads>>>(o, c, 1.0f, 0.999, n);
      }
      cudaDeviceSynchronize(); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
      printf("%-36s %4d threads: %.2f clk per warp-step per SM (%.1f lanes / clk / SM)\n", names[mode], threads,
             (double)h[0] / ((double)n * 8 * (threads / 32)), (double)n * 8 * threads / (double)h[0]);
    }
  return 0;
}

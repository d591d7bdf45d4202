// Micro-probe: issue / completion rate of tcgen05.mma.cta_group::2 (M = 256, both CTAs of a cluster) against cta_group::1.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../lbfgs_ffnn_b200/csrc pair_probe.cu -o pair_probe -lcuda
#include "tc_ptx.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace b200::tcx;

__device__ __forceinline__ void umma2_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_f16(int n, int m) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

__global__ void probe(int n, int N, long long *out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar2;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(raw)[i] = 0;
  if (threadIdx.x < 32) tmem_alloc_pair(smem_u32(&slot), 512u);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar2), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t id = idesc_f16(N, 256);
    const uint64_t dA = desc_k_major(base), dB = desc_k_major(base + 64 * 1024);
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) umma2_f16(tm, dA + (uint64_t)((i & 3) * 2 + ((i >> 2) & 3) * 1024), dB + (uint64_t)((i & 3) * 2 + ((i >> 2) & 3) * 1024), id, i > 0);
    const long long t1 = clock64();
    umma_commit_pair(smem_u32(&bar2));
    mbar_wait(smem_u32(&bar2), 0);
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  if (threadIdx.x == 0 && rank == 1) mbar_wait(smem_u32(&bar2), 0);
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc_pair(tm, 512u); }
}

int main() {
  long long *d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n = 256;
  for (int N : {64, 128, 256}) {
    long long h[2] = {0, 0};
    for (int rep = 0; rep < 2; ++rep) {
      cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, probe, n, N, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    }
    printf("cta_group::2 M 256 N %3d : issue %.1f clk/mma, complete %.1f clk/mma\n", N, (double)h[0] / n, (double)h[1] / n);
  }
  return 0;
}

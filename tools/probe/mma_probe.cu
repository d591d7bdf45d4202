// Micro-probe: issue rate of tcgen05.mma kind::f16 / kind::tf32 (M=128) from one thread, with and without commits.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../lbfgs_ffnn_b200/csrc mma_probe.cu -o mma_probe -lcuda
#include "tc_ptx.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace b200::tcx;

__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }

__device__ __forceinline__ bool elect_one() { // one lane of a converged warp; the compiler knows the branch it guards is single-threaded
  uint32_t pred = 0, laneid = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %2;\n\t@px mov.s32 %1, 1;\n\tmov.s32 %0, rx;\n\t}" : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}
// mode: 0 = n MMAs into one accumulator; 1 = alternate two accumulators; 2 = one accumulator + commit every 4; kind: 0 f16, 1 tf32
__global__ void probe(int n, int N, int mode, int kind, int same_a, long long *out, int elect) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t slot;
  __shared__ uint64_t bar, bar2;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(raw)[i] = 0;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (elect && threadIdx.x < 32) { // the whole warp runs the loop; one elected lane issues
    const uint32_t id = kind == 0 ? idesc_f16(N) : make_idesc(0, 0, N);
    const uint64_t dA = desc_k_major(base), dB = desc_k_major(base + 64 * 1024);
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const uint64_t da = dA + (same_a ? 0 : (uint64_t)(((i & 3) * 2) + ((i >> 2) & 3) * 1024)), db = dB + (uint64_t)((i & 3) * 2);
      const uint32_t d = tm + ((mode == 1 && (i & 1)) ? 256u : 0u);
      if (elect_one()) {
        if (kind == 0) umma_f16(d, da, db, id, i > 0);
        else umma_tf32(d, da, db, id, i > 0);
        if (mode == 2 && (i & 3) == 3) umma_commit(smem_u32(&bar));
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit(smem_u32(&bar2));
    __syncwarp();
    mbar_wait(smem_u32(&bar2), 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (!elect && threadIdx.x == 0) {
    const uint32_t id = kind == 0 ? idesc_f16(N) : make_idesc(0, 0, N);
    const uint64_t dA = desc_k_major(base), dB = desc_k_major(base + 64 * 1024);
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const uint64_t da = dA + (same_a ? 0 : (uint64_t)(((i & 3) * 2) + ((i >> 2) & 3) * 1024)), db = dB + (uint64_t)((i & 3) * 2);
      const uint32_t d = tm + ((mode == 1 && (i & 1)) ? 256u : 0u);
      if (kind == 0) umma_f16(d, da, db, id, i > 0);
      else umma_tf32(d, da, db, id, i > 0);
      if (mode == 2 && (i & 3) == 3) umma_commit(smem_u32(&bar));
    }
    const long long t1 = clock64();
    umma_commit(smem_u32(&bar2)); // tracks every MMA issued above
    mbar_wait(smem_u32(&bar2), 0);
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory"); }
}

int main() {
  long long *d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n = 256;
  for (int kind = 0; kind < 2; ++kind)
    for (int N : {64, 128, 256})
      for (int mode = 0; mode < 3; ++mode)
        for (int elect = 0; elect < 2; ++elect) {
          const int same_a = 0;
          long long h[2] = {0, 0};
          for (int rep = 0; rep < 2; ++rep) {
            probe<<<1, 128, 200 * 1024>>>(n, N, mode, kind, same_a, d, elect);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          }
          fflush(stdout); printf("kind %s N %3d mode %d elect %d : issue %.1f clk/mma, complete %.1f clk/mma\n", kind ? "tf32" : "f16 ", N, mode, elect,
                 (double)h[0] / n, (double)h[1] / n);
        }
  return 0;
}

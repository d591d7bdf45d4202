// Micro-probe: steady-state cost of one K block of the issuer / producer ring of the fp16 kernels, without any memory traffic:
//   producer warp: wait empty(s) -> arrive full(s)            (x2 producers: "X" and "W")
//   issuer:        wait fullX(s), wait fullW(s) -> fence -> [4 MMAs N = 256] -> commit empty(s)
// variants: lane-0 branch vs elect.sync issue, MMAs on / off, one or two full barriers, NS stages.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../lbfgs_ffnn_b200/csrc loop_probe.cu -o loop_probe -lcuda
#include "tc_ptx.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace b200::tcx;

__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0, laneid = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %2;\n\t@px mov.s32 %1, 1;\n\tmov.s32 %0, rx;\n\t}" : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}

template <int NS, bool ELECT, bool TWO>
__global__ void probe(int n, int mma, int fence, long long *out, int spin_mode) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bars[3 * 8 + 2];
  auto fullx = [&](int s) { return smem_u32(&bars[s]); };
  auto fullw = [&](int s) { return smem_u32(&bars[8 + s]); };
  auto empty = [&](int s) { return smem_u32(&bars[16 + s]); };
  const uint32_t done = smem_u32(&bars[24]);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(raw)[i] = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(fullx(s), 1); mbar_init(fullw(s), 1); mbar_init(empty(s), TWO ? 1 : 1); }
    mbar_init(done, 1); mbar_init(smem_u32(&bars[25]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0 || (warp == 3 && TWO)) { // producers
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int i = 0; i < n; ++i) {
        mbar_wait(empty(s), ph ^ 1);
        mbar_arrive(warp == 0 ? fullx(s) : fullw(s));
        if (++s == NS) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t id = idesc_f16(256);
    const uint64_t dA = desc_k_major(base), dB = desc_k_major(base + 64 * 1024);
    if (ELECT) {
      int s = 0; uint32_t ph = 0;
      const long long t0 = clock64();
      for (int i = 0; i < n; ++i) {
        mbar_wait(fullx(s), ph);
        if (TWO) mbar_wait(fullw(s), ph);
        if (fence) tc_fence_after();
        if (elect_one()) {
          if (mma) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_f16(tm, dA + (uint64_t)(s * 1024 + 2 * ks), dB + (uint64_t)(s * 2048 + 2 * ks), id, (i > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(empty(s));
        }
        __syncwarp();
        if (++s == NS) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(done);
      __syncwarp();
      mbar_wait(done, 0);
      if (lane == 0) out[0] = clock64() - t0;
    } else if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      const long long t0 = clock64();
      for (int i = 0; i < n; ++i) {
        mbar_wait(fullx(s), ph);
        if (TWO) mbar_wait(fullw(s), ph);
        if (fence) tc_fence_after();
        if (mma) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_f16(tm, dA + (uint64_t)(s * 1024 + 2 * ks), dB + (uint64_t)(s * 2048 + 2 * ks), id, (i > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(empty(s));
        if (++s == NS) { s = 0; ph ^= 1; }
      }
      umma_commit(done);
      mbar_wait(done, 0);
      out[0] = clock64() - t0;
    }
  }
  if (warp >= 4) { // "epilogue" warps waiting for something that only happens at the end
    if (spin_mode == 1) mbar_wait(smem_u32(&bars[25]), 0);                       // every lane polls
    else if (spin_mode == 2) { if (lane == 0) mbar_wait(smem_u32(&bars[25]), 0); __syncwarp(); } // one lane polls
  }
  if (warp == 1 && lane == 0) mbar_arrive(smem_u32(&bars[25]));
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory"); }
}

template <int NS, bool ELECT, bool TWO> void run(long long *d, const char *name) {
  cudaFuncSetAttribute(probe<NS, ELECT, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n = 512;
  for (int mma = 0; mma < 2; ++mma)
    for (int spin = 0; spin < 3; ++spin) {
      const int fence = 1;
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        probe<NS, ELECT, TWO><<<1, 384, 200 * 1024>>>(n, mma, fence, d, spin);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      }
      printf("%s NS %d mma %d spinning warps mode %d : %.1f clk per K block\n", name, NS, mma, spin, (double)h / n);
      fflush(stdout);
    }
}
int main() {
  long long *d; cudaMalloc(&d, 16);
  run<4, false, true>(d, "lane0 two-full");
  run<4, true, true>(d, "elect two-full");
  run<4, false, false>(d, "lane0 one-full");
  run<4, true, false>(d, "elect one-full");
  return 0;
}

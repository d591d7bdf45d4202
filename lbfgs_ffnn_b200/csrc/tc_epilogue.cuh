// Epilogue helpers shared by the tcgen05 kernels: compile-time activation tags and the coalesced 32x32 block store.
#pragma once

#include "common.cuh"

namespace b200 {
namespace tcx {

// activation with a compile-time tag when ACT >= 0 (the ReLU fast path), else the runtime switch. A per-element
// runtime switch puts a uniform branch between every pair of independent FMAs and serialises the epilogue.
template <int ACT> __device__ __forceinline__ float act_apply_c(int act, float v) {
  if constexpr (ACT == B200_ACT_RELU) return fmaxf(v, 0.0f);
  else if constexpr (ACT == B200_ACT_LINEAR) return v;
  else return act_apply(act, v);
}
template <int ACT> __device__ __forceinline__ float act_deriv_c(int act, float a) {
  if constexpr (ACT == B200_ACT_RELU) return a > 0.0f ? 1.0f : 0.0f;
  else if constexpr (ACT == B200_ACT_LINEAR) return 1.0f;
  else return act_deriv_from_output(act, a);
}
template <int V> struct IntTag { static constexpr int value = V; };

// Store a 32(row) x 32(col) fp32 block held one row per lane as 32 fully coalesced 128-byte row segments
// (a per-thread row store would touch 32 different lines per instruction, 8x the L2 write transactions).
// scratch: this warp's private 32 x 33 floats of shared memory.
__device__ __forceinline__ void store_block_coalesced(const float (&r)[32], float *scratch, float *gbase, long ld,
                                                      int rows_ok, int lane) {
#pragma unroll
  for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = r[j];
  __syncwarp();
#pragma unroll 8
  for (int rr = 0; rr < 32; ++rr)
    if (rr < rows_ok) gbase[(long)rr * ld + lane] = scratch[rr * 33 + lane];
  __syncwarp();
}

} // namespace tcx
} // namespace b200

// The skinny last layer (out <= 12, in = 32 / 64 / 128) in ONE pass over the penultimate activations:
//   forward  A_L = act(A_{L-1} W_L + b_L)                      (src/cuda/layer.cuh:48-58, kernels.cuh:74-106)
//   loss     0.5 ||A_L - T||^2 / B, delta_L = (A_L - T)/B .* act'   (src/cuda/network.cuh:100-107, layer.cuh:69-79)
//   dX       delta_{L-1} = (delta_L W_L^T) .* act'(A_{L-1})     (layer.cuh:89-103, kernels.cuh:109-133)
//   dW, db   [A_{L-1} | 1]^T delta_L split-K partials           (layer.cuh:81-86, kernels.cuh:144-153)
// i.e. 3 SGEMMs + 7 element-wise kernels + a blocking dot of the reference. A_{L-1} is read once (coalesced rows) and
// delta_{L-1} written once; everything else stays in registers: a lane owns in/32 consecutive features and the matching
// rows of W_L, a warp takes two samples at a time and reduces their 2 x 16 partial pre-activations with a 27-shuffle
// transpose-reduce (lane v ends up with value v), so no shared memory is touched until the per-CTA dW combine.
// HBM-bound by design (bytes: A_{L-1} + delta_{L-1}); fp32 FFMA throughout, so it serves every precision mode.
#include "network.cuh"
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

namespace b200 {

namespace {

struct TailParams {
  const float *A;      // [B][in] penultimate activations
  const float *W;      // [in][out] then out biases
  const float *T;      // [B][out]
  float *out_last;     // [B][out]
  float *delta_last;   // [B][ldd]
  float *delta_prev;   // [B][in] fp32 (nullptr: not needed)
  __half *d16;         // [B][2 in] fp16 {hi | lo} of S * delta_prev (nullptr: not needed)
  float *scale16_inv;  // device scalar 1 / S of the chained delta_0 pair (deeper nets, block 0 of the backward kernel writes it)
  float *scale_d16_inv; // device scalar 1 / S of the d16 pair this kernel writes
  float *db_prev_part; // [grid][in] per-CTA column sums of delta_{L-1} (= the db_{L-1} partials), or nullptr
  // deeper nets: block 0 publishes S (and 1 / S) of delta_0 from max |delta_L| and the per-layer row-norm maxima (ChainW)
  const float *chain_cw;
  int chain_nl, chain_nctas;
  float *scale16;
  float *amax_part;    // [grid] per-CTA max |delta_L|
  int n_amax;
  float *partial;      // [grid][(in+1)*out]
  double *loss_part;   // [grid]
  long batch;
  int chunk;           // samples per CTA (even)
  int out, ldd, act_last, act_prev;
  float inv_batch;
  const SpecState *spec_st; // speculative launch on a wrong guess: return at once (common.cuh)
  int spec;
};

template <int FPL> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };

template <int FPL> __device__ __forceinline__ void load_row(const float *p, float (&a)[FPL]) {
  using V = typename VecT<FPL>::type;
  const V v = __ldg(reinterpret_cast<const V *>(p));
  const float *f = reinterpret_cast<const float *>(&v);
#pragma unroll
  for (int c = 0; c < FPL; ++c) a[c] = f[c];
}
template <int FPL> __device__ __forceinline__ void load_row_shared(const float *p, float (&a)[FPL]) {
  using V = typename VecT<FPL>::type;
  const V v = *reinterpret_cast<const V *>(p);
  const float *f = reinterpret_cast<const float *>(&v);
#pragma unroll
  for (int c = 0; c < FPL; ++c) a[c] = f[c];
}
constexpr int kFwdStages = 5;  // tail_fwd: steps (2 rows per warp) in flight per warp
constexpr int kBwdStages = 5;  // tail_bwd: steps (2 rows + their delta_L per warp) in flight per warp
template <int FPL> __device__ __forceinline__ void store_row(float *p, const float (&a)[FPL]) {
  using V = typename VecT<FPL>::type;
  V v;
  float *f = reinterpret_cast<float *>(&v);
#pragma unroll
  for (int c = 0; c < FPL; ++c) f[c] = a[c];
  *reinterpret_cast<V *>(p) = v;
}

// transpose-reduce of 2 x 16 per-lane partials over the warp: lane v ends up with the warp total of value v
// (v = 16 * sample + output); 27 shuffles for OLP <= 12 instead of 5 per value
template <int OLP>
__device__ __forceinline__ float transpose_reduce(const float (&p0)[OLP], const float (&p1)[OLP], int lane) {
  float q[16];
  const int half = lane >> 4;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j < OLP) {
      const float send = half ? p0[j] : p1[j], keep = half ? p1[j] : p0[j];
      q[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    } else {
      q[j] = 0.0f;
    }
  }
  float r8[8], r4[4], r2[2];
  {
    const bool b = (lane >> 3) & 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float send = b ? q[j] : q[8 + j], keep = b ? q[8 + j] : q[j];
      r8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool b = (lane >> 2) & 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float send = b ? r8[j] : r8[4 + j], keep = b ? r8[4 + j] : r8[j];
      r4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool b = (lane >> 1) & 1;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float send = b ? r4[j] : r4[2 + j], keep = b ? r4[2 + j] : r4[j];
      r2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
  }
  const bool b = lane & 1;
  const float send = b ? r2[0] : r2[1], keep = b ? r2[1] : r2[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

// ---- pass 1: last-layer forward, loss, delta_L; per-CTA max |delta_L| (the scale of the fp16 delta_{L-1} needs it) ----
// OLP: compile-time padded output count (10 = the MNIST fast path, 12 = anything up to 12 with zero-padded weights)
template <int FPL, int OLP>
__global__ void __launch_bounds__(256, 2) tail_fwd_kernel(const TailParams p) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  if (spec_skip(p.spec_st, p.spec)) return;
  constexpr int IN = 32 * FPL;
  __shared__ double lred[32];
  __shared__ float mred[8];
  __shared__ float dbred[8][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int OL = (OLP == 10) ? 10 : p.out; // launch_tail: the 10-wide instantiation serves exactly out == 10
  float w[FPL][OLP];
#pragma unroll
  for (int c = 0; c < FPL; ++c)
#pragma unroll
    for (int j = 0; j < OLP; ++j) w[c][j] = (j < OL) ? __ldg(p.W + (size_t)(lane * FPL + c) * OL + j) : 0.0f;
  const int myj = lane & 15, half = lane >> 4;
  const float bj = (myj < OL) ? __ldg(p.W + (size_t)IN * OL + myj) : 0.0f;
  double lsum = 0.0;
  float amax = 0.0f, accd = 0.0f; // accd: this lane's (sample parity, output) share of db_L = sum_s delta_L[s][:]
  const long b0 = (long)blockIdx.x * p.chunk, b1 = min(p.batch, b0 + (long)p.chunk);
  // Activations arrive through a per-warp ring of bulk async copies (two rows = one step of this warp per stage, kFwdStages
  // steps ahead): with register prefetch one step ahead a warp kept 1 KB in flight, 16 KB per SM, and the kernel ran at the
  // 1.6 TB/s that Little's law gives for that; the ring keeps kFwdStages KB per warp in flight. Each warp is producer and consumer
  // of its own slots, so there is no cross-warp synchronisation.
  __shared__ __align__(128) float ring[8][kFwdStages][2 * IN];
  __shared__ __align__(8) unsigned long long bars[8][kFwdStages];
  const uint32_t ring_w = tcx::smem_u32(&ring[warp][0][0]), bar_w = tcx::smem_u32(&bars[warp][0]);
  if (lane == 0) {
    for (int st = 0; st < kFwdStages; ++st) tcx::mbar_init(bar_w + 8 * st, 1);
    tcx::fence_mbar_init();
  }
  __syncwarp();
  auto issue = [&](long s, int st) { // lane 0: rows s, s + 1 of this warp's step (clipped at the end of the CTA's slice)
    if (s < b1) {
      const uint32_t bytes = (uint32_t)min(2L, b1 - s) * (IN * 4);
      tcx::mbar_expect_tx(bar_w + 8 * st, bytes);
      tcx::bulk_load_1d(ring_w + st * (2 * IN * 4), p.A + s * IN, bytes, bar_w + 8 * st);
    }
  };
  float a0[FPL], a1[FPL];
  long s = b0 + 2 * warp;
  if (lane == 0)
    for (int st = 0; st < kFwdStages; ++st) issue(s + 16L * st, st);
  int st = 0;
  uint32_t ph = 0;
  for (; s < b1; s += 16) {
    tcx::mbar_wait(bar_w + 8 * st, ph);
    load_row_shared<FPL>(&ring[warp][st][lane * FPL], a0);
    load_row_shared<FPL>(&ring[warp][st][IN + lane * FPL], a1);
    if (s + 1 >= b1) {
#pragma unroll
      for (int c = 0; c < FPL; ++c) a1[c] = 0.0f;
    }
    const long smy = s + half;
    float tj = 0.0f;
    if (myj < OL && smy < b1) tj = __ldg(p.T + smy * OL + myj);
    float p0[OLP], p1[OLP];
#pragma unroll
    for (int j = 0; j < OLP; ++j) { p0[j] = 0.0f; p1[j] = 0.0f; }
#pragma unroll
    for (int c = 0; c < FPL; ++c)
#pragma unroll
      for (int j = 0; j < OLP; j += 2) { // packed fp32 FMA: two outputs per issue slot
        tcx::ffma2(p0[j], p0[j + 1], a0[c], a0[c], w[c][j], w[c][j + 1]);
        tcx::ffma2(p1[j], p1[j + 1], a1[c], a1[c], w[c][j], w[c][j + 1]);
      }
    const float z = transpose_reduce<OLP>(p0, p1, lane); // pre-activation of (sample smy, output myj)
    if (myj < OL && smy < b1) {
      const float o = act_apply(p.act_last, z + bj);
      const float d = o - tj;
      const float dl = d * p.inv_batch * act_deriv_from_output(p.act_last, o);
      p.out_last[smy * OL + myj] = o;
      p.delta_last[smy * p.ldd + myj] = dl;
      lsum += (double)d * (double)d;
      amax = fmaxf(amax, fabsf(dl));
      accd += dl;
    }
    __syncwarp(); // every lane has consumed its part of the slot: refill it for the step kFwdStages ahead
    if (lane == 0) issue(s + 16L * kFwdStages, st);
    if (++st == kFwdStages) { st = 0; ph ^= 1; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) mred[warp] = amax;
  const float dbv = accd + __shfl_xor_sync(0xffffffffu, accd, 16); // lanes j and 16 + j hold the two sample parities
  if (lane < 16) dbred[warp][lane] = dbv;
  const double tot = block_sum(lsum, lred); // (contains the __syncthreads that publishes mred and dbred)
  if (threadIdx.x < OL) { // bias row of this CTA's [dW_L; db_L] partial (tail_bwd_kernel writes the dW rows), fixed warp order
    float v = dbred[0][threadIdx.x];
    for (int i = 1; i < 8; ++i) v += dbred[i][threadIdx.x];
    p.partial[(size_t)blockIdx.x * (size_t)(IN + 1) * OL + (size_t)IN * OL + threadIdx.x] = v;
  }
  if (threadIdx.x == 0) {
    p.loss_part[blockIdx.x] = tot;
    float m = mred[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, mred[i]);
    p.amax_part[blockIdx.x] = m;
  }
}

// ---- pass 1, sample-per-lane form ------------------------------------------------------------------------------------------
// A lane owns one SAMPLE of a 32-sample tile and all OLP pre-activations of it, so no cross-lane reduction is needed (the
// feature-per-lane form above spends 40 % of its instructions in the transpose-reduce). The tile's activations arrive in stages
// of 32 features x 32 samples = one TMA box {32 floats, 32 rows} with SWIZZLE_128B: row r's 16-byte chunk c lands at chunk
// c ^ (r & 7) of its 128-byte row, so the lanes' reads down a column of rows are conflict-free (a linear copy would put all 32 on
// one bank, and per-row bulk copies into padded rows serialise on the uniform datapath: 32 issues per stage). Rows past the
// batch are zero-filled. W_L sits in shared memory and is read as broadcasts. 16 warps per CTA, one CTA per SM, kF2Stages stages
// in flight per warp; a warp's tiles are g = cta + grid * (warp + 16 i).
// Per-CTA results as in tail_fwd_kernel (loss, max |delta_L|, bias row of the [dW_L; db_L] partial); tail_bwd_kernel runs on a
// larger grid (n_bwd partials): the bias rows of the partials this grid does not own are zeroed here.
constexpr int kF2Warps = 16, kF2Stages = 3, kF2StageBytes = 32 * 32 * 4;
template <int IN> struct Fwd2Plan {
  static constexpr int NH = IN / 32;
  static constexpr int kWFloats = (IN + 1) * 12;
  static constexpr int kSmemBytes = kF2Warps * kF2Stages * kF2StageBytes + kWFloats * 4 + 1024;
};
template <int FPL, int OLP>
__global__ void __launch_bounds__(kF2Warps * 32, 1) tail_fwd2_kernel(const __grid_constant__ CUtensorMap tmA, const TailParams p, int n_bwd) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  if (spec_skip(p.spec_st, p.spec)) return;
  constexpr int IN = 32 * FPL;
  constexpr int NH = Fwd2Plan<IN>::NH;
  extern __shared__ uint8_t fsm_raw[];
  const uint32_t base = (tcx::smem_u32(fsm_raw) + 1023u) & ~1023u; // the swizzle pattern is a function of the shared-memory address
  uint8_t *bp = fsm_raw + (base - tcx::smem_u32(fsm_raw));
  float *Ws = reinterpret_cast<float *>(bp + kF2Warps * kF2Stages * kF2StageBytes); // [IN + 1][12]: W_L rows zero padded; row IN = bias
  __shared__ __align__(8) unsigned long long bars[kF2Warps][kF2Stages];
  __shared__ double lred[32];
  __shared__ float mred[kF2Warps];
  __shared__ float dbred[kF2Warps][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int OL = (OLP == 10) ? 10 : p.out;
  for (int e = threadIdx.x; e < (IN + 1) * 12; e += blockDim.x) {
    const int f = e / 12, j = e - f * 12;
    Ws[e] = (j < OL) ? __ldg(p.W + (size_t)f * OL + j) : 0.0f;
  }
  const uint32_t ring_w = base + warp * (kF2Stages * kF2StageBytes), bar_w = tcx::smem_u32(&bars[warp][0]);
  const uint8_t *ring_mine = bp + warp * (kF2Stages * kF2StageBytes);
  if (lane == 0) {
    for (int st = 0; st < kF2Stages; ++st) tcx::mbar_init(bar_w + 8 * st, 1);
    tcx::fence_mbar_init();
  }
  if (threadIdx.x == 0) tcx::tma_prefetch_desc(&tmA);
  __syncthreads();
  const long ntiles = (p.batch + 31) / 32;
  const long g0 = blockIdx.x + (long)gridDim.x * warp, gstep = (long)gridDim.x * kF2Warps;
  const long my_tiles = g0 < ntiles ? (ntiles - g0 + gstep - 1) / gstep : 0;
  const long nq = my_tiles * NH; // stages this warp consumes: (tile round i, feature block h) = (q / NH, q % NH)
  auto issue = [&](long q) {     // lane 0
    if (q < nq) {
      const int st = (int)(q % kF2Stages);
      tcx::mbar_expect_tx(bar_w + 8 * st, kF2StageBytes);
      tcx::tma_load_2d(ring_w + st * kF2StageBytes, &tmA, bar_w + 8 * st, (int)(q % NH) * 32, (int)((g0 + (q / NH) * gstep) * 32));
    }
  };
  if (lane == 0)
    for (int q = 0; q < kF2Stages; ++q) issue(q);

  double lsum = 0.0;
  float amax = 0.0f;
  float accd[OLP];
#pragma unroll
  for (int j = 0; j < OLP; ++j) accd[j] = 0.0f;
  long q = 0;
  for (long i = 0; i < my_tiles; ++i) {
    const long s = (g0 + i * gstep) * 32 + lane; // this lane's sample
    const bool ok = s < p.batch;
    float tg[OLP];
#pragma unroll
    for (int j = 0; j < OLP; ++j) tg[j] = (ok && j < OL) ? __ldg(p.T + s * OL + j) : 0.0f; // in flight during the dot products
    float z[OLP];
#pragma unroll
    for (int j = 0; j < OLP; ++j) z[j] = 0.0f;
#pragma unroll 1
    for (int h = 0; h < NH; ++h, ++q) {
      const int st = (int)(q % kF2Stages);
      tcx::mbar_wait(bar_w + 8 * st, (uint32_t)((q / kF2Stages) & 1));
      const float4 *ar = reinterpret_cast<const float4 *>(ring_mine + st * kF2StageBytes + lane * 128);
      const float4 *wr = reinterpret_cast<const float4 *>(Ws + (size_t)h * 32 * 12);
#pragma unroll 2
      for (int c = 0; c < 8; ++c) {
        const float4 av = ar[c ^ (lane & 7)];
        const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 w0 = wr[(4 * c + e) * 3 + 0], w1 = wr[(4 * c + e) * 3 + 1];
          tcx::ffma2(z[0], z[1], a4[e], a4[e], w0.x, w0.y);
          tcx::ffma2(z[2], z[3], a4[e], a4[e], w0.z, w0.w);
          tcx::ffma2(z[4], z[5], a4[e], a4[e], w1.x, w1.y);
          tcx::ffma2(z[6], z[7], a4[e], a4[e], w1.z, w1.w);
          if constexpr (OLP == 10) {
            const float2 w2 = *reinterpret_cast<const float2 *>(&wr[(4 * c + e) * 3 + 2]);
            tcx::ffma2(z[8], z[9], a4[e], a4[e], w2.x, w2.y);
          } else {
            const float4 w2 = wr[(4 * c + e) * 3 + 2];
            tcx::ffma2(z[8], z[9], a4[e], a4[e], w2.x, w2.y);
            tcx::ffma2(z[10], z[11], a4[e], a4[e], w2.z, w2.w);
          }
        }
      }
      __syncwarp(); // the stage has been read by every lane
      if (lane == 0) issue(q + kF2Stages);
    }
    if (ok) {
      const float *brow = Ws + (size_t)IN * 12;
      float o[12], dl[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) { o[j] = 0.0f; dl[j] = 0.0f; }
#pragma unroll
      for (int j = 0; j < OLP; ++j) {
        if (j < OL) {
          o[j] = act_apply(p.act_last, z[j] + brow[j]);
          const float d = o[j] - tg[j];
          dl[j] = d * p.inv_batch * act_deriv_from_output(p.act_last, o[j]);
          lsum += (double)d * (double)d;
          amax = fmaxf(amax, fabsf(dl[j]));
          accd[j] += dl[j];
        }
      }
      float *orow = p.out_last + s * OL;
      if (OL == 10) { // 40-byte rows: 8-byte aligned
#pragma unroll
        for (int j = 0; j < 10; j += 2) *reinterpret_cast<float2 *>(orow + j) = make_float2(o[j], o[j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < OLP; ++j) if (j < OL) orow[j] = o[j];
      }
      float *drow = p.delta_last + s * p.ldd; // ldd = out rounded up to 4: the padding is written as zeros
#pragma unroll
      for (int j = 0; j < 12; j += 4)
        if (j < p.ldd) *reinterpret_cast<float4 *>(drow + j) = make_float4(dl[j], dl[j + 1], dl[j + 2], dl[j + 3]);
    }
  }
  // ---- per-CTA results (fixed order: deterministic) ----
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
#pragma unroll
  for (int j = 0; j < OLP; ++j) {
    float v = accd[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) dbred[warp][j] = v;
  }
  if (lane == 0) mred[warp] = amax;
  const double tot = block_sum(lsum, lred); // (contains the __syncthreads that publishes mred and dbred)
  const size_t pstride = (size_t)(IN + 1) * OL;
  if (threadIdx.x < OL) {
    float v = dbred[0][threadIdx.x];
    for (int i = 1; i < kF2Warps; ++i) v += dbred[i][threadIdx.x];
    p.partial[(size_t)blockIdx.x * pstride + (size_t)IN * OL + threadIdx.x] = v;
    for (int c = blockIdx.x + gridDim.x; c < n_bwd; c += gridDim.x) p.partial[(size_t)c * pstride + (size_t)IN * OL + threadIdx.x] = 0.0f;
  }
  if (threadIdx.x == 0) {
    p.loss_part[blockIdx.x] = tot;
    float m = mred[0];
    for (int i = 1; i < kF2Warps; ++i) m = fmaxf(m, mred[i]);
    p.amax_part[blockIdx.x] = m;
  }
}

// ---- pass 2: delta_{L-1} (fp32 for a dX GEMM and / or scaled fp16 hi|lo for the fp16 dW GEMM) and the [dW_L; db_L] partials ----
template <int FPL, int OLP, bool RELU>
__global__ void __launch_bounds__(256, 2) tail_bwd_kernel(const TailParams p) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  static_assert(OLP % 2 == 0, "packed FMAs take the outputs in pairs");
  if (spec_skip(p.spec_st, p.spec)) return;
  constexpr int IN = 32 * FPL;
  // the activation ring of the main loop and, after it, four copies of the dW accumulators (warps w and w + 4 share one)
  constexpr int kRingFloats = 8 * kBwdStages * 2 * IN, kRedFloats = 4 * IN * OLP;
  __shared__ __align__(128) float ring_red[kRingFloats > kRedFloats ? kRingFloats : kRedFloats];
  __shared__ __align__(16) float ring_d[8][kBwdStages][2 * 12];
  __shared__ __align__(8) unsigned long long bars[8][kBwdStages];
  float *red = ring_red;
  __shared__ float mred[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long b0 = (long)blockIdx.x * p.chunk, b1 = min(p.batch, b0 + (long)p.chunk);
  // two consecutive samples per warp and step; their activations (2 IN floats) and delta_L (2 ldd floats, broadcast to every
  // lane) arrive through a per-warp ring of bulk async copies kBwdStages steps ahead (see tail_fwd_kernel)
  float *ring_mine = ring_red + warp * (kBwdStages * 2 * IN);
  const uint32_t ring_w = tcx::smem_u32(ring_mine), ringd_w = tcx::smem_u32(&ring_d[warp][0][0]);
  const uint32_t bar_w = tcx::smem_u32(&bars[warp][0]);
  if (lane == 0) {
    for (int st = 0; st < kBwdStages; ++st) tcx::mbar_init(bar_w + 8 * st, 1);
    tcx::fence_mbar_init();
  }
  __syncwarp();
  const int ldd = p.ldd;
  auto issue = [&](long s, int st) { // lane 0
    if (s < b1) {
      const uint32_t rows = (uint32_t)min(2L, b1 - s);
      tcx::mbar_expect_tx(bar_w + 8 * st, rows * (IN * 4 + ldd * 4));
      tcx::bulk_load_1d(ring_w + st * (2 * IN * 4), p.A + s * IN, rows * (IN * 4), bar_w + 8 * st);
      tcx::bulk_load_1d(ringd_w + st * 96, p.delta_last + s * ldd, rows * (ldd * 4), bar_w + 8 * st);
    }
  };
  long s = b0 + 2 * warp;
  if (lane == 0) // first loads out before the weight / scale prologue below
    for (int st0 = 0; st0 < kBwdStages; ++st0) issue(s + 16L * st0, st0);
  const int OL = (OLP == 10) ? 10 : p.out; // launch_tail: the 10-wide instantiation serves exactly out == 10
  float w[FPL][OLP];
#pragma unroll
  for (int c = 0; c < FPL; ++c)
#pragma unroll
    for (int j = 0; j < OLP; ++j) w[c][j] = (j < OL) ? __ldg(p.W + (size_t)(lane * FPL + c) * OL + j) : 0.0f;
  // power-of-two scale S of the fp16 copy: |delta_{L-1}| <= max_f ||W_L[f,:]||_1 * max |delta_L| =: bound; S * bound in [2^13, 2^14)
  float S = 1.0f;
  if (p.d16) {
    float m = 0.0f;
    for (int i = threadIdx.x; i < p.n_amax; i += blockDim.x) m = fmaxf(m, __ldg(p.amax_part + i));
    float cw = 0.0f;
#pragma unroll
    for (int c = 0; c < FPL; ++c) {
      float r = 0.0f;
#pragma unroll
      for (int j = 0; j < OLP; ++j) r += fabsf(w[c][j]);
      cw = fmaxf(cw, r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      cw = fmaxf(cw, __shfl_xor_sync(0xffffffffu, cw, o));
    }
    if (lane == 0) mred[warp] = m;
    __syncthreads();
    m = mred[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, mred[i]);
    const float bound = m * cw;
    int e = 0;
    if (bound > 0.0f && bound < 3.0e38f) frexpf(bound, &e);
    e = max(-100, min(100, e));
    S = ldexpf(1.0f, 14 - e);
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.scale_d16_inv = ldexpf(1.0f, e - 14);
  }
  if (p.chain_cw && blockIdx.x == 0) { // block-uniform
    float m = 0.0f;
    for (int i = threadIdx.x; i < p.n_amax; i += blockDim.x) m = fmaxf(m, __ldg(p.amax_part + i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) mred[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < 8; ++i) m = fmaxf(m, mred[i]);
      float bound = m;
      for (int l = 0; l < p.chain_nl; ++l) {
        float cwl = 0.0f;
        for (int i = 0; i < p.chain_nctas; ++i) cwl = fmaxf(cwl, __ldg(p.chain_cw + i * kMaxChain + l));
        bound *= cwl;
      }
      int e = 0;
      if (bound > 0.0f && bound < 3.0e38f) frexpf(bound, &e);
      e = max(-100, min(100, e));
      *p.scale16 = ldexpf(1.0f, 14 - e);
      *p.scale16_inv = ldexpf(1.0f, e - 14);
    }
  }
  float acc[FPL][OLP];
#pragma unroll
  for (int c = 0; c < FPL; ++c)
#pragma unroll
    for (int j = 0; j < OLP; ++j) acc[c][j] = 0.0f;
  float gsum[FPL]; // column sums of delta_{L-1} over this warp's samples (db_{L-1})
#pragma unroll
  for (int c = 0; c < FPL; ++c) gsum[c] = 0.0f;

  auto process = [&](long s, const float *arow, const float *drow) { // one sample: delta_{L-1}[s][:] and the dW update
    float a[FPL];
    load_row_shared<FPL>(arow + lane * FPL, a);
    float d[OLP];
    {
      const float4 *dp = reinterpret_cast<const float4 *>(drow);
      const float4 d0 = dp[0], d1 = ldd > 4 ? dp[1] : make_float4(0.f, 0.f, 0.f, 0.f), d2 = ldd > 8 ? dp[2] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float t[12] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y, d2.z, d2.w};
#pragma unroll
      for (int j = 0; j < OLP; ++j) d[j] = (j < OL) ? t[j] : 0.0f;
    }
    float g[FPL];
#pragma unroll
    for (int c = 0; c < FPL; ++c) {
      float xe = 0.0f, xo = 0.0f;
#pragma unroll
      for (int j = 0; j < OLP; j += 2) {
        tcx::ffma2(xe, xo, w[c][j], w[c][j + 1], d[j], d[j + 1]);
        tcx::ffma2(acc[c][j], acc[c][j + 1], a[c], a[c], d[j], d[j + 1]);
      }
      const float x = xe + xo;
      g[c] = RELU ? (a[c] > 0.0f ? x : 0.0f) : x * act_deriv_from_output(p.act_prev, a[c]);
      gsum[c] += g[c];
    }
    if (p.delta_prev) store_row<FPL>(p.delta_prev + s * IN + lane * FPL, g);
    if (p.d16) { // [s][0..IN) = hi, [s][IN..2 IN) = lo of S * delta
      __align__(8) __half hi[FPL], lo[FPL];
      if constexpr (FPL % 2 == 0) {
#pragma unroll
        for (int c = 0; c < FPL; c += 2) { // packed conversions: 2 values per instruction
          const float x0 = g[c] * S, x1 = g[c + 1] * S;
          const __half2 h = __floats2half2_rn(x0, x1);
          const float2 hf = __half22float2(h);
          *reinterpret_cast<__half2 *>(&hi[c]) = h;
          *reinterpret_cast<__half2 *>(&lo[c]) = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
        }
      } else {
#pragma unroll
        for (int c = 0; c < FPL; ++c) {
          const float x = g[c] * S;
          hi[c] = __float2half_rn(x);
          lo[c] = __float2half_rn(x - __half2float(hi[c]));
        }
      }
      __half *row = p.d16 + s * (2 * IN) + lane * FPL;
      if constexpr (FPL == 4) {
        *reinterpret_cast<uint2 *>(row) = *reinterpret_cast<const uint2 *>(hi);
        *reinterpret_cast<uint2 *>(row + IN) = *reinterpret_cast<const uint2 *>(lo);
      } else if constexpr (FPL == 2) {
        *reinterpret_cast<uint32_t *>(row) = *reinterpret_cast<const uint32_t *>(hi);
        *reinterpret_cast<uint32_t *>(row + IN) = *reinterpret_cast<const uint32_t *>(lo);
      } else {
        row[0] = hi[0];
        row[IN] = lo[0];
      }
    }
  };
  int st = 0;
  uint32_t ph = 0;
  for (; s < b1; s += 16) {
    tcx::mbar_wait(bar_w + 8 * st, ph);
    const float *ar = ring_mine + st * (2 * IN), *dr = &ring_d[warp][st][0];
    process(s, ar, dr);
    if (s + 1 < b1) process(s + 1, ar + IN, dr + ldd);
    __syncwarp(); // the slot has been consumed by every lane: refill it for the step kBwdStages ahead
    if (lane == 0) issue(s + 16L * kBwdStages, st);
    if (++st == kBwdStages) { st = 0; ph ^= 1; }
  }
  __syncthreads(); // every warp is done with its ring (nothing is in flight: each copy issued was waited for) before `red` reuses it

  // ---- per-CTA combine: warps 0-3 park their accumulators, warps 4-7 add theirs on top, then each thread adds the four copies
  // of its elements in a fixed order (deterministic; three barriers instead of eight serialised rounds) ---------------------
  if (warp < 4) {
#pragma unroll
    for (int c = 0; c < FPL; ++c)
#pragma unroll
      for (int j = 0; j < OLP; ++j) red[warp * (IN * OLP) + (lane * FPL + c) * OLP + j] = acc[c][j];
  }
  __syncthreads();
  if (warp >= 4) {
#pragma unroll
    for (int c = 0; c < FPL; ++c)
#pragma unroll
      for (int j = 0; j < OLP; ++j) red[(warp - 4) * (IN * OLP) + (lane * FPL + c) * OLP + j] += acc[c][j];
  }
  __syncthreads();
  float *dst = p.partial + (size_t)blockIdx.x * (size_t)(IN + 1) * OL; // the bias row was written by tail_fwd_kernel
  for (int e = threadIdx.x; e < IN * OL; e += blockDim.x) {
    const int i = e / OL, j = e - i * OL;
    float v = red[i * OLP + j];
#pragma unroll
    for (int wv = 1; wv < 4; ++wv) v += red[wv * (IN * OLP) + i * OLP + j];
    dst[e] = v;
  }
  if (p.db_prev_part) { // per-CTA column sums of delta_{L-1}, warps combined in a fixed order
    __syncthreads();
#pragma unroll
    for (int c = 0; c < FPL; ++c) red[warp * IN + lane * FPL + c] = gsum[c];
    __syncthreads();
    for (int i = threadIdx.x; i < IN; i += blockDim.x) {
      float v = red[i];
#pragma unroll
      for (int wv = 1; wv < 8; ++wv) v += red[wv * IN + i];
      p.db_prev_part[(size_t)blockIdx.x * IN + i] = v;
    }
  }
}

template <int FPL, int OLP> int launch_tail_fwd2(const TailParams &p, int grid_fwd, int grid_bwd, cudaStream_t st) {
  auto kern = tail_fwd2_kernel<FPL, OLP>;
  constexpr int smem = Fwd2Plan<32 * FPL>::kSmemBytes;
  CUtensorMap tmA; // A_{L-1} [batch][IN] fp32: box {32 features, 32 samples}
  B200_TRY(tc_make_map_2d_f32(&tmA, p.A, 32 * FPL, (unsigned long long)p.batch, 32, 32));
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  B200_CUDA(launch_ex(kern, dim3(grid_fwd), dim3(kF2Warps * 32), (size_t)smem, st, 1, tmA, p, grid_bwd));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

// grid_fwd > 0: sample-per-lane forward on its own grid (p.n_amax = grid_fwd); else the feature-per-lane forward on `grid`
template <int FPL> int launch_tail(b200_ctx *ctx, const TailParams &p, int grid, int grid_fwd, cudaStream_t st) {
  if (p.out == 10) {
    { ProfScope ps(ctx, "tail_fwd");
      if (grid_fwd > 0) B200_TRY((launch_tail_fwd2<FPL, 10>(p, grid_fwd, grid, st)));
      else B200_LAUNCH((tail_fwd_kernel<FPL, 10>), grid, 256, 0, st, p); }
    { ProfScope ps(ctx, "tail_bwd"); if (p.act_prev == B200_ACT_RELU) B200_LAUNCH((tail_bwd_kernel<FPL, 10, true>), grid, 256, 0, st, p);
      else B200_LAUNCH((tail_bwd_kernel<FPL, 10, false>), grid, 256, 0, st, p); }
  } else {
    { ProfScope ps(ctx, "tail_fwd");
      if (grid_fwd > 0) B200_TRY((launch_tail_fwd2<FPL, 12>(p, grid_fwd, grid, st)));
      else B200_LAUNCH((tail_fwd_kernel<FPL, 12>), grid, 256, 0, st, p); }
    { ProfScope ps(ctx, "tail_bwd"); if (p.act_prev == B200_ACT_RELU) B200_LAUNCH((tail_bwd_kernel<FPL, 12, true>), grid, 256, 0, st, p);
      else B200_LAUNCH((tail_bwd_kernel<FPL, 12, false>), grid, 256, 0, st, p); }
  }
  return B200_OK;
}

} // namespace

bool tail_applicable(const b200_net *net) {
  if (env().tail == 0) return false; // debugging aid: 0 = per-layer kernels
  const int L = net->nlayers();
  if (L < 2) return false;
  const int in = net->dims[L - 1], out = net->dims[L];
  return out <= 12 && (in == 32 || in == 64 || in == 128) && net->ldd[L - 2] == in;
}

// Runs the last layer forward, loss, delta_L, delta_{L-1} and the [dW_L; db_L] partials. Requires net_ensure(batch) and the
// penultimate activations in net->act[L-2]. want16: also (or, with !want32, only) write delta_{L-1} as scaled fp16 {hi | lo}
// into net->delta16 for the fp16 dW GEMM of layer L-2 (gemm_dw16.cu), with 1 / scale in net->scale16_inv.
int tail_layer(b200_net *net, const float *params, const float *t, long batch, float inv_batch, bool want32, bool want16,
               bool chain16) {
  const bool mid = net->m16.on; // delta_{L-1} = delta_1 leaves as its own fp16 pair (and its column sums), delta_0's scale is chained
  const int L = net->nlayers();
  const int in = net->dims[L - 1], out = net->dims[L];
  B200_TRY(tail_ensure_scalars(net));
  if (((want16 && !mid) || chain16) && net->delta16_cap < batch) {
    if (net->delta16) cudaFree(net->delta16);
    net->delta16 = nullptr;
    B200_CUDA(cudaMalloc(&net->delta16, sizeof(__half) * 2 * (size_t)net->dims[1] * net->cap)); // rows of delta_0 (L == 2: in == dims[1])
    ++net->config_gen;
    net->delta16_cap = net->cap;
  }
  TailParams p{};
  p.A = net->act[L - 2];
  p.W = params + net->offs[L - 1];
  p.T = t;
  p.out_last = net->act[L - 1];
  p.delta_last = net->delta[L - 1];
  p.delta_prev = want32 ? net->delta[L - 2] : nullptr;
  p.d16 = want16 ? (__half *)(mid ? net->m16.d16 : net->delta16) : nullptr;
  p.scale16_inv = net->scale16_inv;
  p.scale_d16_inv = mid ? net->m16.scale1_inv : net->scale16_inv;
  p.db_prev_part = mid ? net->m16.db_part : nullptr;
  p.amax_part = net->amax_part;
  if (chain16) {
    ChainW c{};
    tail_chain_fill(net, params, &c);
    p.chain_cw = net->chain_cw; p.chain_nl = c.nl; p.chain_nctas = c.nctas; p.scale16 = net->scale16;
  }
  p.partial = net->partials + net->part_off[L - 1];
  p.loss_part = net->loss_part;
  p.batch = batch;
  p.out = out; p.ldd = net->ldd[L - 1];
  p.act_last = net->acts[L - 1]; p.act_prev = net->acts[L - 2];
  p.inv_batch = inv_batch;
  p.spec_st = net->spec_st; p.spec = net->spec_flag;
  const int max_grid = std::min(std::min(net->skinny_splits[L - 1], net->loss_part_cap), 2 * net->ctx->num_sms);
  int grid = std::max(1, std::min(max_grid, ceil_div(batch, 32)));
  p.chunk = ceil_div(ceil_div(batch, grid), 16) * 16;
  grid = ceil_div(batch, p.chunk);
  // forward: sample-per-lane kernel on one CTA per SM (B200_TAIL_FWD=1: the feature-per-lane kernel on the backward's grid)
  const bool fwd2 = env().tail_fwd != 1 && (reinterpret_cast<uintptr_t>(p.A) & 15u) == 0 && (p.ldd % 4) == 0 &&
                    (reinterpret_cast<uintptr_t>(p.delta_last) & 15u) == 0 && (out != 10 || (reinterpret_cast<uintptr_t>(p.out_last) & 7u) == 0);
  // (never more forward CTAs than backward partials: each forward CTA leaves its bias-gradient row in the partial of its index)
  const int grid_fwd = fwd2 ? std::max(1, std::min(std::min(std::min(net->ctx->num_sms, net->loss_part_cap), grid), (int)ceil_div(batch, 32))) : 0;
  p.n_amax = fwd2 ? grid_fwd : grid;
  if (in == 128) B200_TRY(launch_tail<4>(net->ctx, p, grid, grid_fwd, net->ctx->stream));
  else if (in == 64) B200_TRY(launch_tail<2>(net->ctx, p, grid, grid_fwd, net->ctx->stream));
  else B200_TRY(launch_tail<1>(net->ctx, p, grid, grid_fwd, net->ctx->stream));
  net->splits_used[L - 1] = grid;
  if (mid) net->m16.db_splits = grid;
  net->loss_part_n = fwd2 ? grid_fwd : grid;
  return B200_OK;
}

// deeper nets: a few CTAs scan the weight matrices of layers 1..L-1, one warp per row (ChainW), so they must be small
bool tail_chain16_applicable(const b200_net *net) {
  const int L = net->nlayers();
  if (L < 3 || L - 1 > kMaxChain) return false;
  size_t total = 0;
  for (int l = 1; l < L; ++l) total += (size_t)net->dims[l] * net->dims[l + 1];
  int rows = 0;
  for (int l = 1; l < L; ++l) rows += net->dims[l];
  return total <= (size_t)1 << 20 && rows <= kChainRowsPerCta * kMaxChainCtas;
}

int tail_ensure_scalars(b200_net *net) {
  if (!net->amax_part) {
    B200_CUDA(cudaMalloc(&net->amax_part, sizeof(float) * 2 * net->ctx->num_sms));
    B200_CUDA(cudaMalloc(&net->scale16_inv, (2 + kMaxChain * kMaxChainCtas) * sizeof(float)));
    net->scale16 = net->scale16_inv + 1;
    net->chain_cw = net->scale16_inv + 2;
    ++net->config_gen;
  }
  return B200_OK;
}

void tail_chain_fill(const b200_net *net, const float *params, ChainW *c) {
  const int L = net->nlayers();
  c->nl = L - 1;
  for (int l = 1; l < L; ++l) { c->W[l - 1] = params + net->offs[l]; c->in[l - 1] = net->dims[l]; c->out[l - 1] = net->dims[l + 1]; }
  int rows = 0;
  for (int l = 1; l < L; ++l) rows += net->dims[l];
  c->nctas = ceil_div(rows, kChainRowsPerCta);
  c->cw_part = net->chain_cw;
}

void tail_release(b200_net *net) {
  if (net->amax_part) cudaFree(net->amax_part);
  if (net->scale16_inv) cudaFree(net->scale16_inv);
  if (net->delta16) cudaFree(net->delta16);
  net->amax_part = net->scale16_inv = net->scale16 = net->chain_cw = nullptr;
  net->delta16 = nullptr;
  net->delta16_cap = 0;
}

} // namespace b200

// Persistent forward kernel of layer 0 for an input that is exactly u/255 (8-bit pixels), fp16 tensor-core operands.
//
// Replaces, for the first dense layer, CudaDenseLayer::forward (src/cuda/layer.cuh:48-58: SGEMM + add_bias_kernel +
// activation_kernel).
//
// Why fp16: u in [0, 255] is exact in fp16, so X needs no hi/lo split: net_quantize_input keeps an fp16 copy of X (value u,
// half the bytes of the fp32 array) that TMA drops straight into the UMMA K-major SWIZZLE_128B layout. Each weight is split ONCE
// per evaluation into hi = fp16(s_o * w), lo = fp16(s_o * w - hi) with a per-output-neuron power-of-two scale s_o
// (split_w16_kernel), i.e. 22 mantissa bits, and the accumulator is rescaled by 1/(255 s_o) in the epilogue. [W_hi; W_lo] is one
// 2*BN-row B operand, so ONE kind::f16 MMA of N = 256 per 16 features yields D = [hi | lo] in adjacent TMEM columns:
// fp32-level products, 4x fewer tensor-core instructions than the 3xTF32 N = 128, K = 8 kernel (tcgen05.mma issues at one
// instruction per ~140-160 clk whatever its N: tools/probe/mma_probe.cu).
//
// Structure (one CTA per SM, 384 threads, tiles of 128 samples taken round-robin):
//   warp 0      TMA producer of the X tiles          (kNC stages x [128 rows][64 fp16], from HBM)
//   warp 3      TMA producer of the weight tiles     (kNW stages x {hi, lo} [BN rows][64 fp16], L2-resident)
//   warp 1      tcgen05.mma.kind::f16 issuer; accumulators [hi | lo] DOUBLE-BUFFERED in TMEM (2 x 2 x BN columns)
//   warps 4-11  epilogue of tile i while the main loop of tile i+1 runs: TMEM -> registers -> scale, bias, activation ->
//               swizzled shared-memory staging -> TMA tile store
// History of the measurements that shaped it (profiles/r01_*.md): a uint8 -> fp16 converter stage in shared memory was bound by
// shared-memory bandwidth (MMA operand reads + TMA fills + converter traffic); per-thread row stores from registers throttled the
// whole SM (24 k of 50 k clk per CTA) until the epilogue moved to TMA stores; per-MMA descriptor arithmetic bound the issuer.
#include "gemm_tc.cuh"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace b200 {

namespace {

using namespace tcx;

constexpr int kFM = 128;                   // samples per tile = UMMA M
constexpr int kFK = 64;                    // K elements per stage (64 fp16 = one 128-byte swizzle row)
constexpr int kConvBytes = kFM * kFK * 2;  // X tile: 16 KB
constexpr int kNC = 4, kNW = 4;            // X and weight rings share one stage index, so ONE tcgen05.commit frees both
constexpr int kFThreads = 384;            // warps 0-3: X TMA, MMA issue, TMEM alloc, weight TMA; warps 4-11: epilogue
constexpr int kEpiWarp0 = 4, kEpiThreads = 256;
constexpr int kStageOutBytes = 32 * 128;      // per epilogue warp: one [32 rows][32 floats] TMA-store box

struct F16Params {
  int rows_valid, cols_valid, k_total, k_blocks, tiles;
  int row0;              // first sample of this evaluation inside the fp16 copy of the input
  int act;
  const float *bias;     // [N]
  const float *colscale; // [N]: 1 / (255 s_o)
  float *out;            // activations [rows][ld_out]
  long ld_out;
  const SpecState *spec_st; // speculative launch on a wrong guess: return at once (common.cuh)
  int spec;
  long long *dbg;        // B200_TC_TIMING: per CTA {total, epilogue busy, epilogue waiting for the accumulator, issuer waiting}
};

static_assert(kNC == kNW, "the X and weight rings share their empty barriers");
template <int BN, bool X2> struct FPlan {
  static constexpr int kWStage = BN * 128 * (X2 ? 2 : 1);
  static constexpr int kOffConv = 0;
  static constexpr int kOffW = kNC * kConvBytes;
  static constexpr int kOffOut = kOffW + kNW * kWStage; // epilogue staging tiles (TMA store), 1024-byte aligned
  static constexpr int kOffCol = kOffOut + 8 * kStageOutBytes; // colscale[128], bias[128]
  static constexpr int kOffBar = kOffCol + 1024;
  static constexpr int kTotal = kOffBar + 256 + 1024;
  static constexpr int kTmemCols = (4 * BN <= 256) ? 256 : 512;
};

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// fp32 accumulate, fp16 x fp16, both operands K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kFM >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int BN, bool X2>
__global__ void __launch_bounds__(kFThreads, 1)
fwd16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWh,
             const __grid_constant__ CUtensorMap tmWl, const __grid_constant__ CUtensorMap tmOut, const F16Params p) {
  if (spec_skip(p.spec_st, p.spec)) return; // before any barrier / TMEM allocation
  using Plan = FPlan<BN, X2>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *bp = smem_raw + (base - smem_u32(smem_raw));
  auto conv_a = [&](int s) { return base + Plan::kOffConv + s * kConvBytes; };
  auto w_a = [&](int s, int lo) { return base + Plan::kOffW + s * Plan::kWStage + lo * (BN * 128); };
  const uint32_t bars = base + Plan::kOffBar;
  auto conv_full = [&](int s) { return bars + 8 * (s); };
  auto conv_empty = [&](int s) { return bars + 8 * (kNC + s); };
  auto w_full = [&](int s) { return bars + 8 * (2 * kNC + s); };
  auto w_empty = [&](int s) { return conv_empty(s); }; // shared: the MMA warp's single commit per K block releases both tiles
  auto tm_full = [&](int b) { return bars + 8 * (2 * kNC + 2 * kNW + b); };
  auto tm_empty = [&](int b) { return bars + 8 * (2 * kNC + 2 * kNW + 2 + b); };
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(bp + Plan::kOffBar + 8 * (2 * kNC + 2 * kNW + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_start = p.dbg ? clock64() : 0;

  // ---- one-time setup -----------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWh);
    if (X2) tma_prefetch_desc(&tmWl);
    for (int s = 0; s < kNC; ++s) { mbar_init(conv_full(s), 1); mbar_init(conv_empty(s), 1); }
    for (int s = 0; s < kNW; ++s) mbar_init(w_full(s), 1);
    for (int b = 0; b < 2; ++b) { mbar_init(tm_full(b), 1); mbar_init(tm_empty(b), kEpiThreads / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)),
                 "r"((uint32_t)Plan::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  float *colsc = reinterpret_cast<float *>(bp + Plan::kOffCol), *biass = colsc + 128;
  for (int i = threadIdx.x; i < 128; i += kFThreads) {
    colsc[i] = (i < p.cols_valid) ? __ldg(p.colscale + i) : 0.0f;
    biass[i] = (i < p.cols_valid) ? __ldg(p.bias + i) : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int umma_n = min(BN, (p.cols_valid + 31) & ~31);

  if (warp == 0) {
    if (lane == 0) { // ===== X producer: fp16 rows straight from HBM into the K-major SWIZZLE_128B operand tile ===========
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(conv_empty(s), ph ^ 1);
          mbar_expect_tx(conv_full(s), kConvBytes);
          tma_load_3d(conv_a(s), &tmX, conv_full(s), 0, p.row0 + tile * kFM, kb);
          if (++s == kNC) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    if (lane == 0) { // ===== weight producer ===========================================================
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(w_empty(s), ph ^ 1);
          mbar_expect_tx(w_full(s), Plan::kWStage);
          tma_load_2d(w_a(s, 0), &tmWh, w_full(s), kb * kFK, 0);
          if (X2) tma_load_2d(w_a(s, 1), &tmWl, w_full(s), kb * kFK, 0);
          if (++s == kNW) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) { // ===== MMA issuer ================================================================
      // One MMA per K step: with the lo tile stored right behind the hi tile, B = [W_hi; W_lo] is a single 2*BN-row K-major
      // operand and D = [hi | lo] lands in adjacent TMEM columns. Descriptors are built once; per stage / K step only the
      // 14-bit start-address field moves (the issuing thread is otherwise bound by descriptor arithmetic, not by the tensor pipe).
      const uint32_t idesc = make_idesc_f16(X2 ? 2 * BN : umma_n);
      const uint64_t dA0 = desc_k_major(conv_a(0)), dB0 = desc_k_major(w_a(0, 0));
      int cs = 0, ws = 0, it = 0;
      uint32_t cph = 0, wph = 0;
      long long waited = 0, waited_w = 0, waited_tm = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t d_acc = tmem_base + (uint32_t)(buf * 2 * BN);
        const long long tt0 = p.dbg ? clock64() : 0;
        mbar_wait(tm_empty(buf), ((it >> 1) & 1) ^ 1);
        if (p.dbg) waited_tm += clock64() - tt0;
        tc_fence_after();
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          const long long t0 = p.dbg ? clock64() : 0;
          mbar_wait(conv_full(cs), cph);
          const long long t0b = p.dbg ? clock64() : 0;
          mbar_wait(w_full(ws), wph);
          if (p.dbg) { waited += t0b - t0; waited_w += clock64() - t0b; }
          tc_fence_after();
          const int nks = min(kFK / 16, (p.k_total - kb * kFK + 15) / 16);
          const uint64_t da = dA0 + (uint64_t)(cs * (kConvBytes >> 4)), db = dB0 + (uint64_t)(ws * (Plan::kWStage >> 4));
#pragma unroll
          for (int ks = 0; ks < kFK / 16; ++ks)
            if (ks < nks) umma_f16(d_acc, da + 2 * ks, db + 2 * ks, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
          umma_commit(conv_empty(cs)); // == w_empty(ws): kNC == kNW and both rings advance together
          if (++cs == kNC) { cs = 0; cph ^= 1; }
          if (++ws == kNW) { ws = 0; wph ^= 1; }
        }
        umma_commit(tm_full(buf));
      }
      if (p.dbg) { p.dbg[8 * blockIdx.x + 3] = waited; p.dbg[8 * blockIdx.x + 4] = waited_w; p.dbg[8 * blockIdx.x + 5] = waited_tm; }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue warps 4-11: warp w owns TMEM lanes (= samples) 32*(w%4).., the two warps of a lane quarter split the
    // 32-column chunks. Each warp stages its [32 rows][32 floats] block in shared memory (SWIZZLE_128B, conflict-free 16-byte
    // stores) and ONE lane issues a TMA tile store: full 128-byte lines, asynchronous, rows past the batch clipped by the
    // tensor map. (Per-thread row stores from registers hit 32 different lines per instruction and throttled the whole SM:
    // 24 k of 50 k clk per CTA.) ===================================================================================
    auto epilogue = [&](auto act_tag) {
      constexpr int ACT = decltype(act_tag)::value;
      const int q = warp & 3, half = (warp - kEpiWarp0) >> 2;
      const float4 *colsc4 = reinterpret_cast<const float4 *>(colsc), *bias4 = reinterpret_cast<const float4 *>(biass);
      const uint32_t stage_a = base + Plan::kOffOut + (warp - kEpiWarp0) * kStageOutBytes;
      uint8_t *stage_p = bp + Plan::kOffOut + (warp - kEpiWarp0) * kStageOutBytes;
      long long busy = 0, waiting = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const long long t0 = p.dbg ? clock64() : 0;
        mbar_wait(tm_full(buf), (it >> 1) & 1);
        tc_fence_after();
        const long long t1 = p.dbg ? clock64() : 0;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * BN);
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) {
          const int c0 = half * 32 + 64 * i;
          if (c0 < p.cols_valid) {
            uint32_t v[32];
            tmem_ld32_nowait(lane_addr + c0, v);
            if (X2) { // hi + lo accumulators, added in RN fp32
              uint32_t w[32];
              tmem_ld32_nowait(lane_addr + BN + c0, w);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
            } else {
              tmem_ld_wait();
            }
            if (lane == 0) tma_store_wait_read(); // the previous block of this warp has left the staging tile
            __syncwarp();
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const float4 s4 = colsc4[c0 / 4 + qq], b4 = bias4[c0 / 4 + qq];
              float4 r;
              r.x = act_apply_c<ACT>(p.act, fmaf(__uint_as_float(v[4 * qq + 0]), s4.x, b4.x));
              r.y = act_apply_c<ACT>(p.act, fmaf(__uint_as_float(v[4 * qq + 1]), s4.y, b4.y));
              r.z = act_apply_c<ACT>(p.act, fmaf(__uint_as_float(v[4 * qq + 2]), s4.z, b4.z));
              r.w = act_apply_c<ACT>(p.act, fmaf(__uint_as_float(v[4 * qq + 3]), s4.w, b4.w));
              *reinterpret_cast<float4 *>(stage_p + lane * 128 + ((qq ^ (lane & 7)) << 4)) = r;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmOut, stage_a, c0, tile * kFM + q * 32);
              tma_store_commit();
            }
          }
        }
        tc_fence_before(); // accumulator consumed: the issuer may start tile it+2 in this buffer
        __syncwarp();
        if (lane == 0) mbar_arrive(tm_empty(buf));
        if (p.dbg) { const long long t2 = clock64(); waiting += t1 - t0; busy += t2 - t1; }
      }
      if (lane == 0) tma_store_wait_all();
      if (p.dbg && threadIdx.x == kEpiWarp0 * 32) { p.dbg[8 * blockIdx.x + 1] = busy; p.dbg[8 * blockIdx.x + 2] = waiting; }
    };
    if (p.act == B200_ACT_RELU) epilogue(IntTag<B200_ACT_RELU>{});
    else if (p.act == B200_ACT_LINEAR) epilogue(IntTag<B200_ACT_LINEAR>{});
    else epilogue(IntTag<-1>{});
  }
  tc_fence_before();
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) p.dbg[8 * blockIdx.x] = clock64() - t_start;
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Plan::kTmemCols) : "memory");
  }
}

// W_l [K][N] fp32 (N contiguous) -> wh, wl [N][ldk] fp16 (K contiguous) with a per-neuron power-of-two scale:
// s_o * max_k |w| in [2^13, 2^14) (fp16 overflows at 65504); colscale[o] = pre / s_o. One CTA of 1024 threads per 8 neurons
// (one 32-byte sector per weight row), 128 K-slices per neuron: every weight is loaded ONCE (<= 8 per thread, all in flight
// together), kept in registers across the max reduction, split, and written out through a shared-memory transpose.
constexpr int kSplitNeurons = 8, kSplitMaxK = 1024;
__global__ void __launch_bounds__(1024) split_w16_kernel(const float *__restrict__ W, int K, int N, int ldk, float pre,
                                                       __half *__restrict__ wh, __half *__restrict__ wlo,
                                                       float *__restrict__ colscale, const SpecState *spec_st, int spec,
                                                       const ChainW chain) {
  if (spec_skip(spec_st, spec)) return;
  __shared__ float red[128][kSplitNeurons + 1];
  if (chain.nl > 0 && blockIdx.x >= gridDim.x - chain.nctas) { // the extra CTAs (see ChainW, network.cuh)
    chain_cw_block(chain, (int)(blockIdx.x - (gridDim.x - chain.nctas)), &red[0][0]);
    return;
  }
  __shared__ float sc[kSplitNeurons];
  __shared__ __align__(16) __half th[kSplitNeurons][kSplitMaxK + 8], tl[kSplitNeurons][kSplitMaxK + 8];
  const int o = threadIdx.x & (kSplitNeurons - 1), kq = threadIdx.x / kSplitNeurons, o0 = blockIdx.x * kSplitNeurons;
  const bool ok = o0 + o < N;
  float v[kSplitMaxK / 128];
  float amax = 0.0f;
#pragma unroll
  for (int i = 0; i < kSplitMaxK / 128; ++i) {
    const int k = kq + 128 * i;
    v[i] = (ok && k < K) ? __ldg(W + (size_t)k * N + o0 + o) : 0.0f;
  }
#pragma unroll
  for (int i = 0; i < kSplitMaxK / 128; ++i) amax = fmaxf(amax, fabsf(v[i]));
  red[kq][o] = amax;
  __syncthreads();
  if (threadIdx.x < kSplitNeurons) {
    float m = 0.0f;
    for (int i = 0; i < 128; ++i) m = fmaxf(m, red[i][threadIdx.x]);
    int e = 0;
    if (m > 0.0f && m < 3.0e38f) frexpf(m, &e); // m = f * 2^e, f in [0.5, 1)
    e = max(-100, min(100, e));
    sc[threadIdx.x] = ldexpf(1.0f, 14 - e);      // s * m in [2^13, 2^14)
    if (o0 + threadIdx.x < N) colscale[o0 + threadIdx.x] = pre * ldexpf(1.0f, e - 14);
  }
  __syncthreads();
  const float s = sc[o];
#pragma unroll
  for (int i = 0; i < kSplitMaxK / 128; ++i) {
    const int k = kq + 128 * i;
    const float x = v[i] * s;
    const __half h = __float2half_rn(x);
    th[o][k] = h;
    tl[o][k] = __float2half_rn(x - __half2float(h));
  }
  __syncthreads();
  const int vec_per_row = ldk / 8; // ldk is a multiple of 8: 16-byte stores
  for (int idx = threadIdx.x; idx < kSplitNeurons * vec_per_row; idx += 1024) {
    const int oo = idx / vec_per_row, kv = idx - oo * vec_per_row;
    if (o0 + oo < N) {
      *reinterpret_cast<uint4 *>(wh + (size_t)(o0 + oo) * ldk + 8 * kv) = *reinterpret_cast<const uint4 *>(&th[oo][8 * kv]);
      *reinterpret_cast<uint4 *>(wlo + (size_t)(o0 + oo) * ldk + 8 * kv) = *reinterpret_cast<const uint4 *>(&tl[oo][8 * kv]);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map_2d(CUtensorMap *tm, CUtensorMapDataType dt, const void *ptr, unsigned long long dim0, unsigned long long dim1,
                unsigned long long stride_bytes, unsigned box0, unsigned box1, CUtensorMapSwizzle sw) {
  static EncodeTiledFn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return (EncodeTiledFn) nullptr;
    return (EncodeTiledFn)f;
  }();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B200_ERR_CUDA;
  }
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {stride_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, dt, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): dims %llu x %llu stride %llu box %u x %u ptr %p", (int)r, dim0, dim1,
              stride_bytes, box0, box1, ptr);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

// fp16 {dim0 contiguous, dim1, dim2} with dense strides, box {box0, box1, 1}, SWIZZLE_128B (the block-major fp16 input copy)
int make_map_3d(CUtensorMap *tm, const void *ptr, unsigned long long dim0, unsigned long long dim1, unsigned long long dim2,
                unsigned box0, unsigned box1) {
  void *fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B200_ERR_CUDA;
  }
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {dim0 * 2, dim0 * dim1 * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = ((EncodeTiledFn)fp)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(ptr), dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d fp16) failed (%d): dims %llu x %llu x %llu ptr %p", (int)r, dim0, dim1, dim2, ptr);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

template <int BN, bool X2>
int launch_fwd16(const CUtensorMap &tx, const CUtensorMap &twh, const CUtensorMap &twl, const CUtensorMap &tout, const F16Params &p, int grid,
                 cudaStream_t st) {
  auto kern = fwd16_kernel<BN, X2>;
  constexpr int smem = FPlan<BN, X2>::kTotal;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  static long long *dbg = nullptr;
  const bool timing = env().tc_timing;
  F16Params pp = p;
  if (timing) {
    if (!dbg) B200_CUDA(cudaMalloc(&dbg, sizeof(long long) * 8 * 1024));
    B200_CUDA(cudaMemsetAsync(dbg, 0, sizeof(long long) * 8 * 1024, st));
    pp.dbg = dbg;
  }
  kern<<<grid, kFThreads, smem, st>>>(tx, twh, twl, tout, pp);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  B200_CUDA(cudaGetLastError());
  if (timing) {
    std::vector<long long> h(8 * 1024);
    B200_CUDA(cudaMemcpyAsync(h.data(), dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < grid; ++i)
      for (int j = 0; j < 8; ++j) a[j] += (double)h[8 * i + j] / grid;
    fprintf(stderr, "[fwd16 timing] BN %d x2 %d grid %d tiles %d: per CTA total %.0f clk | epilogue busy %.0f, waiting %.0f | issuer waits: X %.0f, "
            "weights %.0f, tmem %.0f\n", BN, (int)X2, grid, p.tiles, a[0], a[1], a[2], a[3], a[4], a[5]);
  }
  return B200_OK;
}

} // namespace

int tc_make_map_2d_f32(CUtensorMap *tm, const float *ptr, unsigned long long dim0, unsigned long long dim1, unsigned box0, unsigned box1) {
  return make_map_2d(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ptr, dim0, dim1, dim0 * 4, box0, box1, CU_TENSOR_MAP_SWIZZLE_128B);
}

static bool fwd16_shape_ok(const b200_net *net) {
  if (env().fwd16 == 0) return false; // debugging aid: 0 = use the generic tcgen05 kernel
  const int K = net->dims[0], N = net->dims[1];
  return net->nlayers() >= 2 && net->prec != B200_PREC_FP32 && K % 16 == 0 && K <= kSplitMaxK && N % 32 == 0 && N <= 128;
}

// Per evaluation, before the forward sweep: the scaled fp16 {hi, lo} split of W_0 (its own launch, so that it is timed and
// profiled apart from the GEMM). No-op when the fp16 forward does not apply.
int fwd16_prepare(b200_net *net, const float *params) {
  net->w16_params = nullptr;
  net->chain_ready = false;
  if (!fwd16_shape_ok(net)) return B200_OK;
  const int K = net->dims[0], N = net->dims[1];
  const int ldk = (K + 7) & ~7;
  if (!net->w16h) {
    B200_CUDA(cudaMalloc(&net->w16h, sizeof(__half) * (size_t)N * ldk));
    B200_CUDA(cudaMalloc(&net->w16l, sizeof(__half) * (size_t)N * ldk));
    B200_CUDA(cudaMalloc(&net->colscale, sizeof(float) * N));
    ++net->config_gen;
  }
  ChainW chain{};
  if (net->nlayers() > 2 && dw16_applicable(net)) {
    B200_TRY(tail_ensure_scalars(net));
    tail_chain_fill(net, params, &chain);
    net->chain_ready = true;
  }
  ProfScope ps(net->ctx, "split16");
  B200_LAUNCH(split_w16_kernel, ceil_div(N, kSplitNeurons) + (chain.nl > 0 ? chain.nctas : 0), 1024, 0, net->ctx->stream, params + net->offs[0], K, N,
              ldk, 1.0f / 255.0f, (__half *)net->w16h, (__half *)net->w16l, net->colscale, net->spec_st, net->spec_flag, chain);
  net->w16_params = params;
  return B200_OK;
}

// Layer 0 forward on the fp16 copy of an 8-bit-pixel input (x16: feature-block-major [blocks of 64 features][rows][64] halves, value u = 255 x). Sets *done when it ran.
int fwd16_forward_layer(b200_net *net, int l, const float *params, const X16View &x16, long batch, bool *done) {
  *done = false;
  if (l != 0 || !x16.base || !fwd16_shape_ok(net) || (reinterpret_cast<uintptr_t>(net->act[0]) & 15u)) return B200_OK;
  if (net->w16_params != params) B200_TRY(fwd16_prepare(net, params)); // callers normally prepare before the sweep
  const int K = net->dims[0], N = net->dims[1];
  const bool x2 = net->prec == B200_PREC_TF32X3;
  const int ldk = (K + 7) & ~7;
  cudaStream_t st = net->ctx->stream;
  const float *W = params + net->offs[0];
  CUtensorMap tx, twh, twl, tout;
  // X tiles: box {64 features, 128 samples, 1 block} of the block-major fp16 copy = 16 contiguous KB of DRAM each
  B200_TRY(make_map_3d(&tx, x16.base, 64, (unsigned long long)x16.rows_total, (unsigned long long)x16.nblocks, 64, kFM));
  const unsigned bn = N > 64 ? 128 : 64;
  B200_TRY(make_map_2d(&twh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, net->w16h, ldk, N, (unsigned long long)ldk * 2, kFK, bn,
                       CU_TENSOR_MAP_SWIZZLE_128B));
  B200_TRY(make_map_2d(&twl, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, net->w16l, ldk, N, (unsigned long long)ldk * 2, kFK, bn,
                       CU_TENSOR_MAP_SWIZZLE_128B));
  B200_TRY(make_map_2d(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, net->act[0], N, batch, (unsigned long long)N * 4, 32, 32,
                       CU_TENSOR_MAP_SWIZZLE_128B)); // activations [batch][N], stored as [32 rows][32 floats] boxes
  F16Params p{};
  p.rows_valid = (int)batch; p.cols_valid = N; p.k_total = K;
  p.row0 = (int)x16.row0;
  p.k_blocks = ceil_div(K, kFK);
  p.tiles = ceil_div(batch, kFM);
  p.act = net->acts[0];
  p.bias = W + (size_t)K * N;
  p.colscale = net->colscale;
  p.out = net->act[0]; p.ld_out = N;
  p.spec_st = net->spec_st; p.spec = net->spec_flag;
  const int grid = std::min(net->ctx->num_sms, p.tiles);
  if (bn == 128) {
    if (x2) B200_TRY((launch_fwd16<128, true>(tx, twh, twl, tout, p, grid, st)));
    else B200_TRY((launch_fwd16<128, false>(tx, twh, twl, tout, p, grid, st)));
  } else {
    if (x2) B200_TRY((launch_fwd16<64, true>(tx, twh, twl, tout, p, grid, st)));
    else B200_TRY((launch_fwd16<64, false>(tx, twh, twl, tout, p, grid, st)));
  }
  *done = true;
  return B200_OK;
}

void fwd16_release(b200_net *net) {
  if (net->w16h) cudaFree(net->w16h);
  if (net->w16l) cudaFree(net->w16l);
  if (net->colscale) cudaFree(net->colscale);
  net->w16h = net->w16l = nullptr;
  net->colscale = nullptr;
}

} // namespace b200

// tcgen05 / TMEM / TMA GEMMs of the MLP objective (precision modes B200_PREC_TF32 and B200_PREC_TF32X3).
//
// One warp-specialised kernel template covers the three GEMM roles of a dense layer
// (src/cuda/layer.cuh:48-58 forward, :81-86 dW + db, :89-103 dX) with the reference's separate
// element-wise kernels (kernels.cuh:74-153) folded into the epilogue:
//
//   role  D rows (M=128/CTA)  D cols (N)   K        A operand (smem)            B operand (smem)
//   FWD   samples             out          in       act_{l-1} [B][in]  K-major   W_l [in][out]      MN-major
//   DX    samples             in           out      delta_l   [B][out] K-major   W_l (n=in, k=out)  K-major
//   DW    out                 in           samples  delta_l   [B][out] MN-major  act_{l-1} [B][in]  MN-major
//
// Data path: TMA (cp.async.bulk.tensor, SWIZZLE_128B, OOB zero fill handles every ragged edge) -> shared
// memory ring (3 stages) -> tcgen05.mma.kind::tf32 issued by one thread, fp32 accumulator in TMEM ->
// tcgen05.ld -> fused epilogue. All tiles are 32 floats (one 128-byte swizzle row) deep in K.
// 3xTF32 (fp32-accurate mode): four otherwise idle warps split every landed tile in place into
// hi = rna_tf32(a) and lo = a - hi (exact), and the issuer runs hi*hi + hi*lo + lo*hi into the same
// accumulator, i.e. fp32-level products at one third of the TF32 rate with no extra HBM traffic.
// DW is split-K over the batch with per-split partial tiles (deterministic, combined in fp64 by
// finalize_grad_kernel); the bias gradient is one extra N=16 MMA per K step against a tile of ones.
#include <cuda_fp16.h>
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace b200 {

namespace {

constexpr int BM = 128;           // UMMA M (cta_group::1)
constexpr int BK = 32;            // floats per K block = one 128-byte swizzle row
constexpr int kATileBytes = BM * BK * 4;
constexpr int kSplitThreads = 192; // warps 2-7 prepare operands (3xTF32 split / uint8 conversion) during the main loop
constexpr int kTcThreads = 256;   // warp 0 TMA, warp 1 MMA issue, warp 2 TMEM alloc, warps 4-7 epilogue / splitter
constexpr int kOnesBytes = 2048;  // 16 rows x 128 B of 1.0f (bias-gradient B operand; layout-agnostic)
constexpr int kLastCols = 12;                          // fused last layer: out <= 12, rows padded to 12 floats
constexpr int kLastBytes = (128 + 1) * kLastCols * 4;  // [W_last | b_last] staged in shared memory

enum { MAJOR_K = 0, MAJOR_MN = 1 };
enum { TC_FWD = 0, TC_DX = 1, TC_DW = 2 };
constexpr int kShallowKBlocks = 8; // FWD / DX in 3xTF32 with at most this many K blocks of 32 run the two-stage variant

struct TcParams {
  int rows_valid;   // FWD/DX: batch, DW: out
  int cols_valid;   // FWD: out, DX: in, DW: in
  int k_blocks;     // total K blocks of 32
  int kb_per_split; // DW: K blocks per split (others: k_blocks)
  int act;          // FWD: this layer's activation, DX: previous layer's
  const float *bias;
  float *out;       // FWD: activations, DX: delta_prev ; row-major [rows][ld_out]
  long ld_out;
  const float *aux; // DX: act_{l-1}
  float *partial;   // DW: [split][(in+1)*out]
  unsigned long long partial_stride;
  int out_dim, in_dim; // DW
  float acc_scale;     // uint8 operand variants: 1/255 (accumulator holds 255 * result)
  // FWD of the penultimate layer with the (skinny, out <= 16) last layer fused into the epilogue
  int fuse_last, last_out, last_act;
  const float *w_last;   // [cols_valid][last_out] followed by the last_out biases
  const float *targets;  // [rows][last_out]
  float *out_last, *delta_last, *delta_prev;
  int ld_delta_last;     // row stride of delta_last (out rounded up to 4)
  float inv_batch;
  double *loss_part;     // [gridDim.x * 4]
  // DX of layer 1 feeding the fp16 dW GEMM of layer 0 (gemm_dw16.cu): delta_0 leaves as S * delta in fp16 {hi | lo},
  // row s = [cols_valid hi | cols_valid lo], INSTEAD of the fp32 rows in `out`; S is the device scalar *scale16
  __half *d16;
  const float *scale16;
  long long *dbg;        // B200_TC_TIMING=1: per-CTA {main loop, epilogue} clock64 durations
  const SpecState *spec_st; // speculative launch on a wrong guess: return at once (common.cuh)
  int spec;
};

using namespace tcx;

// U8: 0 = both operands fp32 in HBM; 1 = the A operand (FWD: the input X) is stored as uint8 = 255*x;
//     2 = the B operand (DW: X) is. A uint8 operand is converted to fp32 in shared memory by the splitter warps;
//     it is exact in TF32, so it needs no lo tile and contributes no lo*hi product.
// SHALLOW: two stages, for GEMMs with a handful of K blocks (the middle layers of a deep net): a 3xTF32 stage is 48 KB at BN = 64,
// so two stages let TWO CTAs share an SM (256 TMEM columns each) where the deep pipeline ran one CTA per SM in 3.2 waves
template <int BN, bool X3, int U8, int SHALLOW = 0>
struct SmemPlan {
  static constexpr int kBTileBytes = BN * BK * 4;
  static constexpr bool kSplitA = X3 && U8 != 1, kSplitB = X3 && U8 != 2;
  static constexpr int kOffBHi = kATileBytes;
  static constexpr int kOffALo = kATileBytes + kBTileBytes;
  static constexpr int kOffBLo = kOffALo + (kSplitA ? kATileBytes : 0);
  static constexpr int kOffU8 = kOffBLo + (kSplitB ? kBTileBytes : 0);
  static constexpr int kU8Bytes = U8 == 1 ? BM * BK : (U8 == 2 ? BN * BK : 0);
  static constexpr int kStageBytes = ((kOffU8 + kU8Bytes + 1023) / 1024) * 1024;
  // 3 stages when that leaves room for two CTAs per SM (epilogue of one overlaps the main loop of the other),
  // otherwise as deep as one CTA's 200 KB allows: HBM latency under load is ~4 us, so bytes in flight are what
  // bound the streaming GEMMs (Little's law), not the tensor pipe
  static constexpr int kFit = (200 * 1024) / kStageBytes;
  // uint8-A forward kernel: HBM-light (4 KB per K block), epilogue-heavy -> 2 stages so that TWO CTAs fit an SM and one's
  // epilogue overlaps the other's main loop
  static constexpr int kStages = SHALLOW ? 2 : (U8 == 1 && kStageBytes * 2 <= 104 * 1024) ? 2
                               : (kStageBytes * 3 <= 100 * 1024) ? 3 : (kFit > 6 ? 6 : (kFit < 2 ? 2 : kFit));
  static constexpr int kTxBytes = (U8 == 1 ? BM * BK : kATileBytes) + (U8 == 2 ? BN * BK : kBTileBytes);
  static constexpr int kOnes = (U8 == 1) ? 0 : kOnesBytes; // the ones tile is a dW-only operand; the uint8-A kernel is forward-only
  static constexpr int kBarOffset = kStages * kStageBytes + kOnes + ((kLastBytes + 127) / 128) * 128;
  static constexpr int kTotal = kBarOffset + 128 + 1024; // barriers + tmem slot + alignment slack
};

template <int A_MAJOR, int B_MAJOR, int ROLE, int BN, bool X3, int U8, int SHALLOW = 0>
__global__ void __launch_bounds__(kTcThreads, ((SmemPlan<BN, X3, U8, SHALLOW>::kTotal + 1024) * 2 <= 227 * 1024) ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmBlo, const TcParams p) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  if (spec_skip(p.spec_st, p.spec)) return; // before any barrier / TMEM allocation
  using Plan = SmemPlan<BN, X3, U8, SHALLOW>;
  constexpr bool kSplitA = Plan::kSplitA, kSplitB = Plan::kSplitB;
  // FWD / DX: the B operand is the weight matrix, the same for every CTA: its hi / lo parts are split ONCE per
  // evaluation into two global arrays (split_params_kernel) and arrive as two TMA loads; only A is split in-kernel
  constexpr bool kPresplitB = kSplitB && ROLE != TC_DW;
  constexpr bool kUseSplit = X3 || U8 != 0; // the issuer waits for the splitter / converter warps
  constexpr int kBTileBytes = Plan::kBTileBytes;
  constexpr int kStages = Plan::kStages;
  constexpr bool kBias = (ROLE == TC_DW);
  // 3xTF32 accumulator plan. The tensor core's fp32 accumulate truncates (measured: a ~3e-8 relative bias per
  // accumulating MMA on same-signed sums), so the fp32-accurate mode (a) rotates the hi*hi products of successive
  // K blocks over kMain accumulators and (b) keeps the 2^-12-sized hi*lo + lo*hi corrections in their own
  // accumulator; the epilogue adds them in round-to-nearest fp32. DW keeps one (its sums are sign-mixed).
  constexpr bool kSplitAcc = X3 && ROLE != TC_DW;
  constexpr int kMain = kSplitAcc ? ((BN <= 128 && U8 != 1) ? 3 : 1) : 1; // uint8-A: 256 TMEM columns per CTA, two CTAs per SM
  constexpr int kColsNeeded = kMain * BN + (kSplitAcc ? BN : 0) + (kBias ? 32 : 0);
  constexpr int kTmemCols = kColsNeeded <= 32 ? 32 : kColsNeeded <= 64 ? 64 : kColsNeeded <= 128 ? 128
                          : kColsNeeded <= 256 ? 256 : 512;
  static_assert(kColsNeeded <= 512, "TMEM has 512 columns");
  constexpr int kSmallCol = kMain * BN;                          // hi*lo + lo*hi accumulator (kSplitAcc)
  constexpr int kBiasCol = kMain * BN + (kSplitAcc ? BN : 0);    // bias-gradient accumulator (DW)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *base_ptr = smem_raw + (base - smem_u32(smem_raw));
  auto a_tile = [&](int s, int lo) { return base + s * Plan::kStageBytes + (lo ? Plan::kOffALo : 0); };
  auto b_tile = [&](int s, int lo) { return base + s * Plan::kStageBytes + (lo ? Plan::kOffBLo : Plan::kOffBHi); };
  auto u8_tile = [&](int s) { return base + s * Plan::kStageBytes + Plan::kOffU8; };
  const uint32_t ones = base + kStages * Plan::kStageBytes;
  const uint32_t bars = base + Plan::kBarOffset;
  auto bar_full = [&](int s) { return bars + 8 * s; };
  auto bar_empty = [&](int s) { return bars + 8 * (kStages + s); };
  auto bar_split = [&](int s) { return bars + 8 * (2 * kStages + s); };
  const uint32_t bar_accum = bars + 8 * (3 * kStages);
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(base_ptr + Plan::kBarOffset + 8 * (3 * kStages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_start = p.dbg ? clock64() : 0;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  int kb_begin = 0, kb_end = p.k_blocks;
  if (ROLE == TC_DW) {
    kb_begin = blockIdx.z * p.kb_per_split;
    kb_end = min(p.k_blocks, kb_begin + p.kb_per_split);
  }
  const int nkb = max(0, kb_end - kb_begin);
  // UMMA N of this CTA: the valid columns rounded up to a whole 32-float swizzle atom
  const int n_valid = min(BN, p.cols_valid - n0);
  const int umma_n = (n_valid + 31) & ~31;
  const bool do_bias = kBias && (blockIdx.y == gridDim.y - 1);

  // ---- one-time setup ---------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
      mbar_init(bar_split(s), kSplitThreads);
    }
    mbar_init(bar_accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (kBias && warp == 3) { // tile of ones for the bias-gradient MMA
    float4 *o4 = reinterpret_cast<float4 *>(base_ptr + kStages * Plan::kStageBytes);
    for (int i = lane; i < kOnesBytes / 16; i += 32) o4[i] = make_float4(1.f, 1.f, 1.f, 1.f);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  float *wl = reinterpret_cast<float *>(base_ptr + kStages * Plan::kStageBytes + Plan::kOnes); // fused last layer
  const bool fuse = (ROLE == TC_FWD) && p.fuse_last;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer ========================================================================
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(bar_empty(s), ph ^ 1);
        mbar_expect_tx(bar_full(s), Plan::kTxBytes + (kPresplitB ? kBTileBytes : 0));
        const int k0 = kb * BK;
        if (U8 == 1) {
          tma_load_2d(u8_tile(s), &tmA, bar_full(s), k0, m0);   // uint8 box {32 K bytes, 128 rows}, no swizzle
        } else if (A_MAJOR == MAJOR_K) {
          tma_load_2d(a_tile(s, 0), &tmA, bar_full(s), k0, m0); // box {32 K, 128 rows}
        } else {
#pragma unroll
          for (int g = 0; g < BM / 32; ++g) tma_load_2d(a_tile(s, 0) + g * 4096, &tmA, bar_full(s), m0 + 32 * g, k0);
        }
        if (U8 == 2) {
#pragma unroll
          for (int g = 0; g < BN / 32; ++g) tma_load_2d(u8_tile(s) + g * 1024, &tmB, bar_full(s), n0 + 32 * g, k0); // uint8 {32 N, 32 K}
        } else if (B_MAJOR == MAJOR_K) {
          tma_load_2d(b_tile(s, 0), &tmB, bar_full(s), k0, n0); // box {32 K, BN rows}
          if (kPresplitB) tma_load_2d(b_tile(s, 1), &tmBlo, bar_full(s), k0, n0);
        } else {
#pragma unroll
          for (int g = 0; g < BN / 32; ++g) tma_load_2d(b_tile(s, 0) + g * 4096, &tmB, bar_full(s), n0 + 32 * g, k0);
          if (kPresplitB) {
#pragma unroll
            for (int g = 0; g < BN / 32; ++g) tma_load_2d(b_tile(s, 1) + g * 4096, &tmBlo, bar_full(s), n0 + 32 * g, k0);
          }
        }
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer ==========================================================================
      const uint32_t idesc = make_idesc(A_MAJOR, B_MAJOR, umma_n);
      const uint32_t idesc_bias = make_idesc(A_MAJOR, MAJOR_K, 16);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(kUseSplit ? bar_split(s) : bar_full(s), ph);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint32_t acc = (kb > kb_begin || ks > 0) ? 1u : 0u;
          uint64_t da[2], db[2];
#pragma unroll
          for (int lo = 0; lo < (X3 ? 2 : 1); ++lo) {
            da[lo] = (A_MAJOR == MAJOR_K) ? desc_k_major(a_tile(s, lo) + ks * 32) : desc_mn_major(a_tile(s, lo) + ks * 1024);
            db[lo] = (B_MAJOR == MAJOR_K) ? desc_k_major(b_tile(s, lo) + ks * 32) : desc_mn_major(b_tile(s, lo) + ks * 1024);
          }
          const int rel = kb - kb_begin;
          const uint32_t d_main = tmem_base + (uint32_t)((rel % kMain) * BN);
          const uint32_t acc_main = (rel >= kMain || ks > 0) ? 1u : 0u; // first touch of this accumulator overwrites
          umma_tf32(d_main, da[0], db[0], idesc, acc_main);
          if (X3) {
            const uint32_t d_small = kSplitAcc ? tmem_base + kSmallCol : d_main;
            uint32_t acc_small = kSplitAcc ? acc : 1u;
            if (kSplitB) { umma_tf32(d_small, da[0], db[1], idesc, acc_small); acc_small = 1u; } // hi * lo
            if (kSplitA) { umma_tf32(d_small, da[1], db[0], idesc, acc_small); }                 // lo * hi
          }
          if (do_bias) {
            const uint64_t dones = desc_k_major(ones);
            umma_tf32(tmem_base + kBiasCol, da[0], dones, idesc_bias, acc);
            if (kSplitA) umma_tf32(tmem_base + kBiasCol, da[1], dones, idesc_bias, 1u);
          }
        }
        umma_commit(bar_empty(s)); // smem slot reusable once these MMAs have read it
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
      umma_commit(bar_accum); // accumulator complete
    }
    __syncwarp();
  }
  if (warp >= 2 && warp < 4 && fuse) { // stage [W_last | b_last], rows zero-padded to kLastCols floats (row cols_valid = the bias)
    const int OL = p.last_out;
    for (int i = threadIdx.x - 64; i < (p.cols_valid + 1) * kLastCols; i += 64) {
      const int k = i / kLastCols, j = i - k * kLastCols;
      wl[i] = (j < OL) ? __ldg(p.w_last + (size_t)k * OL + j) : 0.0f;
    }
  }
  if (warp < 2) {
  } else if (kUseSplit) {
    // ===== operand preparation warps (4-7) =======================================================
    // 3xTF32: hi = rna_tf32(a) in place, lo = a - hi (exact), element-wise hence layout-agnostic.
    // uint8 operand: u -> float(u) written straight into the UMMA canonical layout (exact in TF32, no lo tile).
    const int et = threadIdx.x - 64; // 0..191
    auto split_tile = [&](uint8_t *hi_p, uint8_t *lo_p, int bytes) {
      float4 *hi = reinterpret_cast<float4 *>(hi_p);
      float4 *lo = reinterpret_cast<float4 *>(lo_p);
#pragma unroll 4
      for (int i = et; i < bytes / 16; i += kSplitThreads) {
        // hi = a rounded to the nearest TF32 (ties away): add half an ulp of the 10-bit mantissa, clear the low 13
        // bits. cvt.rna.tf32.f32 is emulated with ~12 integer instructions on sm_100; this is 2 (finite inputs).
        const float4 a = hi[i];
        float4 h, l;
        h.x = __uint_as_float((__float_as_uint(a.x) + 0x1000u) & 0xFFFFE000u); l.x = a.x - h.x;
        h.y = __uint_as_float((__float_as_uint(a.y) + 0x1000u) & 0xFFFFE000u); l.y = a.y - h.y;
        h.z = __uint_as_float((__float_as_uint(a.z) + 0x1000u) & 0xFFFFE000u); l.z = a.z - h.z;
        h.w = __uint_as_float((__float_as_uint(a.w) + 0x1000u) & 0xFFFFE000u); l.w = a.w - h.w;
        hi[i] = h;
        lo[i] = l;
      }
    };
    auto u8x4_to_f4 = [](uint32_t w) { // float(u) = as_float(0x4B000000 | u) - 2^23, exact
      return make_float4(__uint_as_float(0x4B000000u | (w & 0xFFu)) - 8388608.0f,
                         __uint_as_float(0x4B000000u | ((w >> 8) & 0xFFu)) - 8388608.0f,
                         __uint_as_float(0x4B000000u | ((w >> 16) & 0xFFu)) - 8388608.0f,
                         __uint_as_float(0x4B000000u | (w >> 24)) - 8388608.0f);
    };
    int s = 0;
    uint32_t ph = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait(bar_full(s), ph);
      uint8_t *st = base_ptr + s * Plan::kStageBytes;
      if (U8 == 1) {
        // raw [128 rows][32 B] -> K-major SWIZZLE_128B tile: row r at r*128, 16-byte chunk c at (c ^ (r & 7))*16
        const uint32_t *raw = reinterpret_cast<const uint32_t *>(st + Plan::kOffU8);
#pragma unroll 2
        for (int idx = et; idx < BM * BK / 4; idx += kSplitThreads) {
          const int r = idx >> 3, c = idx & 7;
          *reinterpret_cast<float4 *>(st + r * 128 + ((c ^ (r & 7)) << 4)) = u8x4_to_f4(raw[idx]);
        }
      } else if (kSplitA) {
        split_tile(st, st + Plan::kOffALo, kATileBytes);
      }
      if (U8 == 2) {
        // raw [BN/32 groups][32 K rows][32 B] -> MN-major SWIZZLE_128B_BASE32B: group g at g*4096, row r at r*128,
        // 32-byte chunk q/2 at ((q/2) ^ (r & 3))*32, half (q & 1)*16
        const uint32_t *raw = reinterpret_cast<const uint32_t *>(st + Plan::kOffU8);
#pragma unroll 4
        for (int idx = et; idx < BN * BK / 4; idx += kSplitThreads) {
          const int g = idx >> 8, r = (idx >> 3) & 31, q = idx & 7;
          *reinterpret_cast<float4 *>(st + Plan::kOffBHi + g * 4096 + r * 128 + ((((q >> 1) ^ (r & 3)) << 5) | ((q & 1) << 4))) =
              u8x4_to_f4(raw[idx]);
        }
      } else if (kSplitB && !kPresplitB) {
        split_tile(st + Plan::kOffBHi, st + Plan::kOffBLo, kBTileBytes);
      }
      fence_async_smem(); // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(bar_split(s));
      if (++s == kStages) { s = 0; ph ^= 1; }
    }
  }

  // ===== epilogue: all 8 warps. Warps w and w+4 own TMEM lanes (= D rows) 32*(w%4).., and split the 32-column
  // chunks of the accumulator between them (even chunks: warps 0-3, odd chunks: warps 4-7) ===================
  __syncthreads(); // roles done issuing; wl staged
  // fused last layer: fetch this row's targets now; the ~4 us loaded HBM latency hides behind the rest of the main loop
  float tgt[kLastCols];
#pragma unroll
  for (int j = 0; j < kLastCols; ++j) tgt[j] = 0.0f;
  if (fuse) {
    const long trow = (long)m0 + (warp & 3) * 32 + lane;
    if (trow < p.rows_valid) {
#pragma unroll
      for (int j = 0; j < kLastCols; ++j)
        if (j < p.last_out) tgt[j] = __ldg(p.targets + trow * p.last_out + j);
    }
  }
  // DX: act'(A_{l-1}) of this warp's first 32 x 32 block, one column per lane (full 128-byte segments), fetched while the MMAs
  // are still running; later blocks are fetched ahead of their TMEM load
  float auxv[32];
  auto fetch_aux = [&](int c0) {
    const long wr0 = (long)m0 + (warp & 3) * 32;
    const int rok = (int)max(0L, min(32L, (long)p.rows_valid - wr0));
    const float *ab = p.aux + wr0 * p.ld_out + n0 + c0 + lane;
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) auxv[rr] = (rr < rok && n0 + c0 + 32 <= p.cols_valid) ? __ldg(ab + (long)rr * p.ld_out) : 0.0f;
  };
  if (ROLE == TC_DX) fetch_aux((warp >> 2) * 32);
  if (nkb > 0) {
    mbar_wait(bar_accum, 0);
    tc_fence_after();
  }
  const long long t_acc = p.dbg ? clock64() : 0;
  auto epilogue = [&](auto act_tag) {
    constexpr int ACT = decltype(act_tag)::value;
    const int half = warp >> 2;
    const int row = (warp & 3) * 32 + lane; // TMEM lane == D row
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const long grow = (long)m0 + row;
    const bool row_ok = grow < p.rows_valid;
    // the pipeline stages are idle once the accumulator is complete: they double as store-transpose scratch
    // (8 x 32x33 floats) and as the exchange buffer of the fused last layer
    float *scratch = reinterpret_cast<float *>(base_ptr) + warp * (32 * 33);
    float *zbuf = reinterpret_cast<float *>(base_ptr) + 8 * (32 * 33); // [2][128][kLastCols]
    const long wrow0 = (long)m0 + (warp & 3) * 32;                          // first row of this warp
    const int rows_ok = (int)max(0L, min(32L, (long)p.rows_valid - wrow0)); // valid rows of this warp
    // accumulator chunk c0..c0+31 of this thread's row: (main_0 + main_1 + main_2) + small in RN fp32
    auto load_chunk = [&](int c0, uint32_t (&v)[32]) {
      if (nkb > 0) {
        tmem_ld32(lane_addr + c0, v);
        if (kSplitAcc) {
          uint32_t w[32];
#pragma unroll
          for (int a = 1; a < kMain; ++a) {
            if (a < nkb) {
              tmem_ld32(lane_addr + a * BN + c0, w);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
            }
          }
          tmem_ld32(lane_addr + kSmallCol + c0, w);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
    };
    float z[kLastCols]; // fused last layer: this half's share of the pre-activations of this thread's sample
#pragma unroll
    for (int j = 0; j < kLastCols; ++j) z[j] = 0.0f;
    uint32_t relu_mask[(BN + 63) / 64]; // ReLU fast path: act'(a) of this thread's elements, one bit each
#pragma unroll
    for (int j = 0; j < (BN + 63) / 64; ++j) relu_mask[j] = 0u;
    for (int c0 = half * 32; c0 < umma_n; c0 += 64) {
      uint32_t v[32];
      if (ROLE == TC_DX && c0 != half * 32) fetch_aux(c0);
      load_chunk(c0, v);
      if (U8 != 0) { // uint8 operand = 255 * x: undo the scale once per accumulator element
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * p.acc_scale);
      }
      const int gcol0 = n0 + c0;
      if (ROLE == TC_FWD || ROLE == TC_DX) {
        if (gcol0 + 32 <= p.cols_valid) {
          float r[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (ROLE == TC_FWD) {
              const float4 b4 = __ldg(reinterpret_cast<const float4 *>(p.bias + gcol0) + q);
              r[4 * q + 0] = act_apply_c<ACT>(p.act, __uint_as_float(v[4 * q + 0]) + b4.x);
              r[4 * q + 1] = act_apply_c<ACT>(p.act, __uint_as_float(v[4 * q + 1]) + b4.y);
              r[4 * q + 2] = act_apply_c<ACT>(p.act, __uint_as_float(v[4 * q + 2]) + b4.z);
              r[4 * q + 3] = act_apply_c<ACT>(p.act, __uint_as_float(v[4 * q + 3]) + b4.w);
              if (fuse) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float a = r[4 * q + e];
                  if (ACT == B200_ACT_RELU && a > 0.0f) relu_mask[c0 / 64] |= 1u << (4 * q + e);
                  const float4 *wr = reinterpret_cast<const float4 *>(wl + (size_t)(gcol0 + 4 * q + e) * kLastCols);
                  const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2];
                  ffma2(z[0], z[1], a, a, w0.x, w0.y); ffma2(z[2], z[3], a, a, w0.z, w0.w);
                  ffma2(z[4], z[5], a, a, w1.x, w1.y); ffma2(z[6], z[7], a, a, w1.z, w1.w);
                  ffma2(z[8], z[9], a, a, w2.x, w2.y); ffma2(z[10], z[11], a, a, w2.z, w2.w);
                }
              }
            }
          }
          if (ROLE == TC_DX) {
            // accumulators through the transpose scratch; act'(A_{l-1}) is applied on the far side, where lanes run along a
            // row: the loads of A_{l-1} and the stores are then full 128-byte segments (a per-thread row read touches 32 lines
            // per instruction)
#pragma unroll
            for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = __uint_as_float(v[j]);
            __syncwarp();
            if (p.d16) {
              // fp16 {hi | lo} of S * delta: a lane packs its element as {hi, lo}, swaps with its neighbour, and the even lane
              // stores the hi pair of columns (2i, 2i + 1), the odd lane their lo pair: two 64-byte runs per instruction
              const float S = __ldg(p.scale16);
              __half *gb = p.d16 + wrow0 * (2L * p.cols_valid) + gcol0 + ((lane & 1) ? p.cols_valid : 0);
#pragma unroll
              for (int rr = 0; rr < 32; rr += 2) { // two rows per step: packed fp32 <-> fp16 conversions (the scalar ones issue at 1/4 rate)
                const float xa = scratch[rr * 33 + lane] * act_deriv_c<ACT>(p.act, auxv[rr]) * S;
                const float xb = scratch[(rr + 1) * 33 + lane] * act_deriv_c<ACT>(p.act, auxv[rr + 1]) * S;
                const __half2 h2 = __floats2half2_rn(xa, xb);
                const float2 hf = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(xa - hf.x, xb - hf.y);
                const uint32_t hb = *reinterpret_cast<const uint32_t *>(&h2), lb = *reinterpret_cast<const uint32_t *>(&l2);
                const uint32_t mine_a = __byte_perm(hb, lb, 0x5410), mine_b = __byte_perm(hb, lb, 0x7632); // {hi, lo} of each row
                const uint32_t other_a = __shfl_xor_sync(0xffffffffu, mine_a, 1), other_b = __shfl_xor_sync(0xffffffffu, mine_b, 1);
                // even lane: hi pair {own, neighbour}; odd lane: lo pair {neighbour, own}
                const uint32_t wa = (lane & 1) ? __byte_perm(other_a, mine_a, 0x7632) : __byte_perm(mine_a, other_a, 0x5410);
                const uint32_t wb = (lane & 1) ? __byte_perm(other_b, mine_b, 0x7632) : __byte_perm(mine_b, other_b, 0x5410);
                if (rr < rows_ok) reinterpret_cast<uint32_t *>(gb + (long)rr * (2L * p.cols_valid))[lane >> 1] = wa;
                if (rr + 1 < rows_ok) reinterpret_cast<uint32_t *>(gb + (long)(rr + 1) * (2L * p.cols_valid))[lane >> 1] = wb;
              }
            } else {
              float *gbase = p.out + wrow0 * p.ld_out + gcol0;
#pragma unroll
              for (int rr = 0; rr < 32; ++rr)
                if (rr < rows_ok) gbase[(long)rr * p.ld_out + lane] = scratch[rr * 33 + lane] * act_deriv_c<ACT>(p.act, auxv[rr]);
            }
            __syncwarp();
          } else {
            store_block_coalesced(r, scratch, p.out + wrow0 * p.ld_out + gcol0, p.ld_out, rows_ok, lane);
          }
        } else if (row_ok) {
          float *dst = p.out + grow * p.ld_out + gcol0;
          const float *aux = (ROLE == TC_DX) ? p.aux + grow * p.ld_out + gcol0 : nullptr;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (gcol0 + j < p.cols_valid) {
              const float acc = __uint_as_float(v[j]);
              dst[j] = (ROLE == TC_FWD) ? act_apply_c<ACT>(p.act, acc + __ldg(p.bias + gcol0 + j))
                                        : acc * act_deriv_c<ACT>(p.act, __ldg(aux + j));
            }
          }
        }
      } else { // TC_DW: D[o][i] -> partial[split][i*out + o]; lanes are consecutive o => coalesced per column
        if (row_ok) {
          float *dst = p.partial + (unsigned long long)blockIdx.z * p.partial_stride + grow;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (gcol0 + j < p.cols_valid) dst[(long)(gcol0 + j) * p.out_dim] = __uint_as_float(v[j]);
        }
      }
    }
    const long long t_p1 = p.dbg ? clock64() : 0;
    if (p.dbg && threadIdx.x == 0 && blockIdx.x < 1024) p.dbg[8192 + 4 * blockIdx.x + 0] = t_p1 - t_acc;
    if (fuse) {
      // last layer forward, loss, delta_L (src/cuda/network.cuh:100-107) and delta_{L-1} (layer.cuh:89-103 +
      // kernels.cuh:109-133) for this thread's sample. The two warps of a row exchange their half sums of z.
#pragma unroll
      for (int j = 0; j < kLastCols; ++j) zbuf[(half * 128 + row) * kLastCols + j] = z[j];
      __syncthreads();
      const int OL = p.last_out;
      float dl[kLastCols];
      double lsum = 0.0;
      const float *brow = wl + (size_t)p.cols_valid * kLastCols;
      const float *zo = zbuf + ((half ^ 1) * 128 + row) * kLastCols;
#pragma unroll
      for (int j = 0; j < kLastCols; ++j) {
        dl[j] = 0.0f;
        if (j < OL && row_ok) {
          const float zz = (half == 0) ? (z[j] + zo[j]) : (zo[j] + z[j]); // same operand order in both warps
          const float o = act_apply(p.last_act, zz + brow[j]);
          const float d = o - tgt[j];
          dl[j] = d * p.inv_batch * act_deriv_from_output(p.last_act, o);
          if (half == 0) {
            p.out_last[grow * OL + j] = o;
            p.delta_last[grow * p.ld_delta_last + j] = dl[j];
            lsum += (double)d * (double)d;
          }
        }
      }
      if (half == 0) {
        lsum = warp_sum(lsum);
        if (lane == 0) p.loss_part[blockIdx.x * 4 + (warp & 3)] = lsum;
      }
      if (p.dbg && threadIdx.x == 0 && blockIdx.x < 1024) p.dbg[8192 + 4 * blockIdx.x + 1] = clock64() - t_p1;
      for (int c0 = half * 32; c0 < umma_n; c0 += 64) {
        uint32_t v[32];
        if (ACT != B200_ACT_RELU) { // generic activations need the activation value again for act'(a)
          load_chunk(c0, v);
          if (U8 != 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * p.acc_scale);
          }
        }
        float r[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float bb[4] = {0.f, 0.f, 0.f, 0.f};
          if (ACT != B200_ACT_RELU) {
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(p.bias + c0) + q);
            bb[0] = b4.x; bb[1] = b4.y; bb[2] = b4.z; bb[3] = b4.w;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 *wr = reinterpret_cast<const float4 *>(wl + (size_t)(c0 + 4 * q + e) * kLastCols);
            const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2];
            float g0 = 0.0f, g1 = 0.0f;
            ffma2(g0, g1, w0.x, w0.y, dl[0], dl[1]); ffma2(g0, g1, w0.z, w0.w, dl[2], dl[3]);
            ffma2(g0, g1, w1.x, w1.y, dl[4], dl[5]); ffma2(g0, g1, w1.z, w1.w, dl[6], dl[7]);
            ffma2(g0, g1, w2.x, w2.y, dl[8], dl[9]); ffma2(g0, g1, w2.z, w2.w, dl[10], dl[11]);
            if (ACT == B200_ACT_RELU) {
              r[4 * q + e] = ((relu_mask[c0 / 64] >> (4 * q + e)) & 1u) ? (g0 + g1) : 0.0f;
            } else {
              const float a = act_apply_c<ACT>(p.act, __uint_as_float(v[4 * q + e]) + bb[e]);
              r[4 * q + e] = (g0 + g1) * act_deriv_c<ACT>(p.act, a);
            }
          }
        }
        store_block_coalesced(r, scratch, p.delta_prev + wrow0 * p.ld_out + c0, p.ld_out, rows_ok, lane);
      }
    }
    if (do_bias && half == 0) { // db[o] = sum_b delta[b][o]: first column of the bias accumulator
      uint32_t v[32];
      if (nkb > 0) {
        tmem_ld32(lane_addr + kBiasCol, v);
      } else {
        v[0] = 0u;
      }
      if (row_ok)
        p.partial[(unsigned long long)blockIdx.z * p.partial_stride + (unsigned long long)p.in_dim * p.out_dim + grow] =
            __uint_as_float(v[0]);
    }
    tc_fence_before();
  };
  if (p.act == B200_ACT_RELU) epilogue(IntTag<B200_ACT_RELU>{});
  else if (p.act == B200_ACT_LINEAR) epilogue(IntTag<B200_ACT_LINEAR>{});
  else epilogue(IntTag<-1>{});
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) {
    const long long t_end = clock64();
    const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (cta < 4096) { p.dbg[2 * cta] = t_acc - t_start; p.dbg[2 * cta + 1] = t_end - t_acc; }
  }
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return (EncodeTiledFn) nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// 2-D fp32 tensor map: dim0 (contiguous) x dim1 with row stride ld floats; box {32, box1}; SWIZZLE_128B; OOB -> 0
int make_map(CUtensorMap *tm, const float *ptr, unsigned long long dim0, unsigned long long dim1, unsigned long long ld,
             unsigned box1, int major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B200_ERR_CUDA;
  }
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {32, box1};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        major == MAJOR_K ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): dims %llu x %llu ld %llu box1 %u ptr %p", (int)r, dim0, dim1, ld, box1, ptr);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

// 2-D uint8 tensor map (the quantised input): dim0 bytes contiguous, box {32, box1}, no swizzle, OOB -> 0
int make_map_u8(CUtensorMap *tm, const uint8_t *ptr, unsigned long long dim0, unsigned long long dim1, unsigned box1) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B200_ERR_CUDA;
  }
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {dim0};
  cuuint32_t box[2] = {32, box1};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(uint8) failed (%d): dims %llu x %llu box1 %u ptr %p", (int)r, dim0, dim1, box1, ptr);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

// speculation gate of the evaluation being launched (host-side, single-threaded per context like the rest of the library)
const SpecState *g_tc_spec_st = nullptr;
int g_tc_spec = 0;
struct TcSpecScope {
  explicit TcSpecScope(const b200_net *net) { g_tc_spec_st = net->spec_st; g_tc_spec = net->spec_flag; }
  ~TcSpecScope() { g_tc_spec_st = nullptr; g_tc_spec = 0; }
};

bool tma_ok(const float *ptr, long ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0 && (ld % 4) == 0; }

template <int A_MAJOR, int B_MAJOR, int ROLE, int BN, bool X3, int U8, int SHALLOW = 0>
int launch_tc(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &tblo, const TcParams &p, dim3 grid, cudaStream_t st) {
  auto kern = gemm_tc_kernel<A_MAJOR, B_MAJOR, ROLE, BN, X3, U8, SHALLOW>;
  constexpr int smem = SmemPlan<BN, X3, U8, SHALLOW>::kTotal;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  static long long *dbg = nullptr;
  const bool timing = env().tc_timing;
  TcParams pp = p;
  pp.spec_st = g_tc_spec_st; pp.spec = g_tc_spec; // set by the tc_*_layer entry points from the network being evaluated
  if (timing) {
    if (!dbg) B200_CUDA(cudaMalloc(&dbg, sizeof(long long) * 4 * 4096));
    B200_CUDA(cudaMemsetAsync(dbg, 0, sizeof(long long) * 4 * 4096, st));
    pp.dbg = dbg;
  }
  B200_CUDA(launch_ex(kern, grid, dim3(kTcThreads), (size_t)smem, st, 1, ta, tb, tblo, pp));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  B200_CUDA(cudaGetLastError());
  if (timing) { // debugging aid: average main-loop and epilogue duration per CTA, in SM clocks
    std::vector<long long> h(4 * 4096);
    B200_CUDA(cudaMemcpyAsync(h.data(), dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    const size_t n = std::min<size_t>(4096, (size_t)grid.x * grid.y * grid.z);
    double m = 0, e = 0;
    for (size_t i = 0; i < n; ++i) { m += h[2 * i]; e += h[2 * i + 1]; }
    double p1 = 0, mid = 0;
    const size_t n2 = std::min<size_t>(1024, n);
    for (size_t i = 0; i < n2; ++i) { p1 += h[8192 + 4 * i]; mid += h[8192 + 4 * i + 1]; }
    fprintf(stderr, "[tc timing] role %d BN %d x3 %d u8 %d grid %ux%ux%u: main loop %.0f clk, epilogue %.0f clk per CTA (pass1 %.0f, last layer %.0f)\n",
            ROLE, BN, (int)X3, U8, grid.x, grid.y, grid.z, m / n, e / n, p1 / n2, mid / n2);
  }
  return B200_OK;
}

template <int A_MAJOR, int B_MAJOR, int ROLE, int BN, int U8 = 0>
int launch_tc_prec(bool x3, const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &tblo, const TcParams &p, dim3 grid,
                   cudaStream_t st) {
  return x3 ? launch_tc<A_MAJOR, B_MAJOR, ROLE, BN, true, U8>(ta, tb, tblo, p, grid, st)
            : launch_tc<A_MAJOR, B_MAJOR, ROLE, BN, false, U8>(ta, tb, tblo, p, grid, st);
}

__global__ void __launch_bounds__(256) split_params_kernel(const float *__restrict__ w, unsigned long long n, float *__restrict__ hi,
                                                         float *__restrict__ lo, const SpecState *spec_st, int spec) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  if (spec_skip(spec_st, spec)) return;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const float a = w[i];
    const float h = __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xFFFFE000u);
    hi[i] = h;
    lo[i] = a - h;
  }
}

int tc_mask() { return env().tc_mask; } // debugging aid: B200_TC_MASK bit0 = FWD, bit1 = DX, bit2 = DW (default all)

} // namespace

static int tc_ensure_split(b200_net *net);

// A_l = act(A_{l-1} W_l + b_l) for hidden layers. With `fuse` (l is the penultimate layer and the last layer has
// out <= 16) the epilogue also runs the last layer, the loss partials, delta_L and delta_{L-1}.
int tc_forward_layer(b200_net *net, int l, const float *params, const float *in, long batch, const TcFuseLast *fuse,
                     bool *done, bool *fused) {
  *done = false;
  if (fused) *fused = false;
  if (!(tc_mask() & 1)) return B200_OK;
  if (wide16_applicable(net, l, 0, batch)) { // wide layer: both operands as fp16 pairs on the CTA-pair kernel (gemm_fwd16.cu)
    B200_TRY(wide16_forward_layer(net, l, params, in, batch));
    *done = true;
    return B200_OK;
  }
  TcSpecScope spec_scope(net);
  const int K = net->dims[l], N = net->dims[l + 1];
  const float *W = params + net->offs[l];
  if (!tma_ok(in, K) || !tma_ok(W, N) || !tma_ok(net->act[l], N) || N % 32 != 0) return B200_OK;
  const bool x3 = net->prec == B200_PREC_TF32X3;
  // layer 0 with an input that is exactly u/255: read the 4x smaller uint8 copy (net_quantize_input)
  const uint8_t *xq = (l == 0 && (tc_mask() & 16) == 0) ? net_xq_lookup(net, in, batch) : nullptr;
  if (l == 0 && (tc_mask() & 16) == 0 && !fuse) { // persistent fp16 kernel (gemm_fwd16.cu); the last layer then runs in tail_layer.cu
    X16View xv;
    if (net_x16_view(net, in, batch, &xv)) B200_TRY(fwd16_forward_layer(net, l, params, xv, batch, done));
    if (*done) return B200_OK;
  }
  B200_TRY(tc_ensure_split(net));
  CUtensorMap ta, tb;
  if (xq) B200_TRY(make_map_u8(&ta, xq, K, batch, BM));        // A: uint8 {K bytes, rows}, box {32, 128}
  else B200_TRY(make_map(&ta, in, K, batch, K, BM, MAJOR_K)); // A: {K, rows}, box {32, 128}
  // 3xTF32: B = the pre-split weights (net_split_params); otherwise the raw weights (tblo unused)
  const float *Whi = x3 ? net->w_hi + net->offs[l] : W, *Wlo = x3 ? net->w_lo + net->offs[l] : W;
  CUtensorMap tblo;
  B200_TRY(make_map(&tb, Whi, N, K, N, 32, MAJOR_MN));        // B: {N, K}, box {32, 32}
  B200_TRY(make_map(&tblo, Wlo, N, K, N, 32, MAJOR_MN));
  TcParams p{};
  p.acc_scale = xq ? 1.0f / 255.0f : 1.0f;
  p.rows_valid = (int)batch; p.cols_valid = N;
  p.k_blocks = ceil_div(K, BK); p.kb_per_split = p.k_blocks;
  p.act = net->acts[l]; p.bias = W + (size_t)K * N;
  p.out = net->act[l]; p.ld_out = N;
  const int L = net->nlayers();
  if (fuse && (tc_mask() & 8) == 0 && l == L - 2 && N <= 128 && net->dims[L] <= kLastCols && tma_ok(net->delta[l], net->ldd[l]) && net->ldd[l] == N) {
    p.fuse_last = 1;
    p.last_out = net->dims[L];
    p.last_act = net->acts[L - 1];
    p.w_last = params + net->offs[L - 1];
    p.targets = fuse->targets;
    p.out_last = net->act[L - 1];
    p.delta_last = net->delta[L - 1];
    p.ld_delta_last = net->ldd[L - 1];
    p.delta_prev = net->delta[l];
    p.inv_batch = fuse->inv_batch;
    p.loss_part = net->loss_part;
    net->loss_part_n = ceil_div(batch, BM) * 4;
    if (fused) *fused = true;
  }
  cudaStream_t st = net->ctx->stream;
  const dim3 g64(ceil_div(batch, BM), ceil_div(N, 64)), g128(ceil_div(batch, BM), ceil_div(N, 128));
  if (xq) {
    if (N <= 64) B200_TRY((launch_tc_prec<MAJOR_K, MAJOR_MN, TC_FWD, 64, 1>(x3, ta, tb, tblo, p, g64, st)));
    else B200_TRY((launch_tc_prec<MAJOR_K, MAJOR_MN, TC_FWD, 128, 1>(x3, ta, tb, tblo, p, g128, st)));
  } else {
    if (x3 && p.k_blocks <= kShallowKBlocks && !p.fuse_last)
      B200_TRY((launch_tc<MAJOR_K, MAJOR_MN, TC_FWD, 64, true, 0, 1>(ta, tb, tblo, p, dim3(g64.x, ceil_div(N, 64)), st)));
    else if (N <= 64) B200_TRY((launch_tc_prec<MAJOR_K, MAJOR_MN, TC_FWD, 64>(x3, ta, tb, tblo, p, g64, st)));
    else B200_TRY((launch_tc_prec<MAJOR_K, MAJOR_MN, TC_FWD, 128>(x3, ta, tb, tblo, p, g128, st)));
  }
  *done = true;
  return B200_OK;
}

// delta_{l-1} = (delta_l W_l^T) .* act'_{l-1}(A_{l-1})
int tc_dx_layer(b200_net *net, int l, const float *params, long batch, bool *done, bool *emit16) {
  *done = false;
  const bool want16 = emit16 && *emit16 && net->delta16 && net->scale16;
  if (emit16) *emit16 = false;
  if (!(tc_mask() & 2)) return B200_OK;
  if (wide16_applicable(net, l, 1, batch)) {
    B200_TRY(wide16_dx_layer(net, l, params, batch));
    *done = true;
    return B200_OK;
  }
  TcSpecScope spec_scope(net);
  const int Kin = net->dims[l], Nout = net->dims[l + 1]; // contraction over out, result width in
  const float *W = params + net->offs[l];
  if (!tma_ok(net->delta[l], net->ldd[l]) || !tma_ok(W, Nout) || !tma_ok(net->delta[l - 1], Kin) || net->ldd[l - 1] != Kin ||
      !tma_ok(net->act[l - 1], Kin) ||
      Kin % 32 != 0)
    return B200_OK;
  const bool x3 = net->prec == B200_PREC_TF32X3;
  B200_TRY(tc_ensure_split(net));
  CUtensorMap ta, tb;
  B200_TRY(make_map(&ta, net->delta[l], Nout, batch, net->ldd[l], BM, MAJOR_K)); // A: {K = out, rows}, box {32, 128}
  TcParams p{};
  p.rows_valid = (int)batch; p.cols_valid = Kin;
  p.k_blocks = ceil_div(Nout, BK); p.kb_per_split = p.k_blocks;
  p.act = net->acts[l - 1];
  p.acc_scale = 1.0f;
  p.out = net->delta[l - 1]; p.ld_out = Kin; p.aux = net->act[l - 1];
  if (want16) { p.d16 = (__half *)net->delta16; p.scale16 = net->scale16; }
  cudaStream_t st = net->ctx->stream;
  const float *Whi = x3 ? net->w_hi + net->offs[l] : W, *Wlo = x3 ? net->w_lo + net->offs[l] : W;
  CUtensorMap tblo;
  if (x3 && p.k_blocks <= kShallowKBlocks) {
    B200_TRY(make_map(&tb, Whi, Nout, Kin, Nout, 64, MAJOR_K));
    B200_TRY(make_map(&tblo, Wlo, Nout, Kin, Nout, 64, MAJOR_K));
    B200_TRY((launch_tc<MAJOR_K, MAJOR_K, TC_DX, 64, true, 0, 1>(ta, tb, tblo, p, dim3(ceil_div(batch, BM), ceil_div(Kin, 64)), st)));
  } else if (Kin <= 64) {
    B200_TRY(make_map(&tb, Whi, Nout, Kin, Nout, 64, MAJOR_K)); // B: {K = out, N = in}, box {32, BN}
    B200_TRY(make_map(&tblo, Wlo, Nout, Kin, Nout, 64, MAJOR_K));
    B200_TRY((launch_tc_prec<MAJOR_K, MAJOR_K, TC_DX, 64>(x3, ta, tb, tblo, p, dim3(ceil_div(batch, BM), ceil_div(Kin, 64)), st)));
  } else {
    B200_TRY(make_map(&tb, Whi, Nout, Kin, Nout, 128, MAJOR_K));
    B200_TRY(make_map(&tblo, Wlo, Nout, Kin, Nout, 128, MAJOR_K));
    B200_TRY((launch_tc_prec<MAJOR_K, MAJOR_K, TC_DX, 128>(x3, ta, tb, tblo, p, dim3(ceil_div(batch, BM), ceil_div(Kin, 128)), st)));
  }
  if (emit16) *emit16 = want16;
  *done = true;
  return B200_OK;
}

// [dW_l ; db_l] split-K partials = [A_{l-1} | 1]^T delta_l
// N tile of the dW kernel: 256 columns (fewer re-reads of delta) except in 3xTF32, where a 256-wide stage is 72-96 KB and
// only two fit — the TMA -> split/convert -> MMA chain then runs latency-bound; 128-wide stages pipeline three deep
static int tc_dw_bn(const b200_net *net) {
  if (env().dw_bn) return env().dw_bn == 128 ? 128 : 256;
  return net->prec == B200_PREC_TF32X3 ? 128 : 256;
}

int tc_dw_plan(b200_net *net, int l, long batch, int *splits) {
  const int Kin = net->dims[l], Nout = net->dims[l + 1];
  const int tiles = ceil_div(Nout, BM) * ceil_div(Kin, tc_dw_bn(net));
  const int kblocks = ceil_div(batch, BK);
  // one CTA per SM (the stages take most of the shared memory): never more CTAs than SMs, or a second wave doubles the time
  int s = std::max(1, std::min(net->ctx->num_sms / tiles, kblocks));
  const int per = ceil_div(kblocks, s);
  *splits = ceil_div(kblocks, per);
  return per;
}

int tc_dw_layer(b200_net *net, int l, const float *in, long batch, bool *done) {
  *done = false;
  if (!(tc_mask() & 4)) return B200_OK;
  if (wide16_applicable(net, l, 2, batch) && (reinterpret_cast<uintptr_t>(net->partials + net->part_off[l]) & 15u) == 0) {
    B200_TRY(wide16_dw_layer(net, l, in, batch));
    *done = true;
    return B200_OK;
  }
  TcSpecScope spec_scope(net);
  const int Kin = net->dims[l], Nout = net->dims[l + 1];
  if (!tma_ok(net->delta[l], net->ldd[l]) || !tma_ok(in, Kin)) return B200_OK;
  const bool x3 = net->prec == B200_PREC_TF32X3;
  const uint8_t *xq = (l == 0 && (tc_mask() & 16) == 0) ? net_xq_lookup(net, in, batch) : nullptr;
  CUtensorMap ta, tb;
  B200_TRY(make_map(&ta, net->delta[l], Nout, batch, net->ldd[l], 32, MAJOR_MN)); // A: {M = out, K = batch}, box {32, 32}
  if (xq) B200_TRY(make_map_u8(&tb, xq, Kin, batch, 32));                  // B: uint8 {N = in bytes, K = batch}, box {32, 32}
  else B200_TRY(make_map(&tb, in, Kin, batch, Kin, 32, MAJOR_MN));         // B: {N = in, K = batch}, box {32, 32}
  int splits = 1;
  const int per = tc_dw_plan(net, l, batch, &splits);
  TcParams p{};
  p.acc_scale = xq ? 1.0f / 255.0f : 1.0f;
  p.rows_valid = Nout; p.cols_valid = Kin;
  p.k_blocks = ceil_div(batch, BK); p.kb_per_split = per;
  p.partial = net->partials + net->part_off[l];
  p.partial_stride = (unsigned long long)(Kin + 1) * Nout;
  p.out_dim = Nout; p.in_dim = Kin;
  const int bn = tc_dw_bn(net);
  const dim3 grid(ceil_div(Nout, BM), ceil_div(Kin, bn), splits);
  if (bn == 128) {
    if (xq) B200_TRY((launch_tc_prec<MAJOR_MN, MAJOR_MN, TC_DW, 128, 2>(x3, ta, tb, tb, p, grid, net->ctx->stream)));
    else B200_TRY((launch_tc_prec<MAJOR_MN, MAJOR_MN, TC_DW, 128>(x3, ta, tb, tb, p, grid, net->ctx->stream)));
  } else {
    if (xq) B200_TRY((launch_tc_prec<MAJOR_MN, MAJOR_MN, TC_DW, 256, 2>(x3, ta, tb, tb, p, grid, net->ctx->stream)));
    else B200_TRY((launch_tc_prec<MAJOR_MN, MAJOR_MN, TC_DW, 256>(x3, ta, tb, tb, p, grid, net->ctx->stream)));
  }
  net->splits_used[l] = splits; // finalize_grad_kernel combines exactly the splits this launch wrote
  *done = true;
  return B200_OK;
}

// 3xTF32: hi / lo parts of the whole parameter vector, once per evaluation (read by the FWD and DX kernels)
int tc_split_params(b200_net *net, const float *params) {
  net->split_src = (net->prec == B200_PREC_TF32X3) ? params : nullptr; // launched by the first kernel that reads the split
  return B200_OK;
}

static int tc_ensure_split(b200_net *net) {
  const float *params = net->split_src;
  if (!params) return B200_OK;
  net->split_src = nullptr;
  if (!net->w_hi) {
    B200_CUDA(cudaMalloc(&net->w_hi, sizeof(float) * net->n));
    B200_CUDA(cudaMalloc(&net->w_lo, sizeof(float) * net->n));
    ++net->config_gen;
  }
  const int blocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)4 * net->ctx->num_sms, (net->n + 255) / 256));
  B200_LAUNCH(split_params_kernel, blocks, 256, 0, net->ctx->stream, params, (unsigned long long)net->n, net->w_hi, net->w_lo,
              net->spec_st, net->spec_flag);
  return B200_OK;
}

void tc_release(b200_net *net) {
  if (net->w_hi) cudaFree(net->w_hi);
  if (net->w_lo) cudaFree(net->w_lo);
  net->w_hi = net->w_lo = nullptr;
  fwd16_release(net);
}

} // namespace b200

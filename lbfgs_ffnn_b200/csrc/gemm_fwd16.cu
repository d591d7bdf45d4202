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
// X ring: NS stages; weight ring: NW stages (default NS: one stage index, ONE tcgen05.commit frees both tiles; NW != NS: the X
// ring, which comes from HBM, is made deeper than the ring of the weights, which come from L2, and each gets its own commit)
enum { EPI_F32 = 0,     // activations as fp32 rows (layer 0 of a two-layer net, layer 1 of a three-layer net)
       EPI_EMIT16 = 1,  // activations ONLY as the per-feature-scaled fp16 pair, block-major (b200_net::Mid16::a16)
       EPI_DX = 2,      // dX role: A = delta pair, B = W^T; epilogue: * act'(A_prev) from the pair, emits delta_prev's fp16 pair
       EPI_WIDE = 3 };  // wide layers (WIDE form below): fp32 rows out; forward (bias, activation), dX (* act'(A_prev) read as fp32)
                        // or plain (dW) by the run-time fields of F16Params
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
  int diag;              // B200_DIAG (timing experiments only, results are wrong): bit0 no X loads, bit1 no weight loads, bit2 no MMAs, bit3 no stores
  int x_block_first;     // coordinates of the A-operand map are {0, block, row} instead of {0, row, block} (row-major pair rows)
  // EPI_EMIT16
  const float *tscale;   // [N]: t_f
  int nb_out;            // blocks of 64 features of the output pair (the lo blocks start at nb_out)
  // EPI_DX
  const float *scale_in_inv; // device scalar: 1 / scale of the A operand (delta pair)
  const float *scale_out;    // device scalar: scale of the emitted pair
  // WIDE
  int n_tiles;               // tiles of BN output columns (units = sample-tile pairs x n_tiles, n fastest)
  int k_chunk;               // K blocks accumulated in TMEM before the epilogue warps add them to their fp32 registers
  int k_slices, kb_per_slice; // split-K: a unit is (tile pair, column tile, slice of the contraction); slice s is stored at index s of
                             // the third dimension of the output map (the dW of a layer with few output tiles: one partial per slice)
  const float *sa_inv, *sb_inv; // device scalars: 1 / scale of the A pair, 1 / scale of the B pair
  const float *aux32;        // dX: A_prev as fp32 rows (act' is taken from it), else nullptr
  long ld_aux;
  int ones_row;              // row of the A operand that holds unscaled ones (the bias-gradient row of a dW), or -1
};

// PAIR: two CTAs of a cluster work as ONE tcgen05 cta_group::2 unit on two neighbouring sample tiles (M = 256). The B operand
// [W_hi; W_lo] is split between them — the even CTA (the leader) stages W_hi, the odd one W_lo — so a K block costs each SM 16 KB
// of X + 16 KB of weights instead of 16 + 32: the weight re-stream out of L2 (2/3 of this kernel's L2 -> SM bytes, the path that
// bounds it: profiles/r02_*) is halved. Same MMA shapes per SM and the same accumulation order, so the results are bit-identical
// to the one-CTA form.
// WIDE (layers with hundreds of output columns, gemm "wide16" entry points at the end of this file): units are (pair of sample
// tiles, tile of BN output columns); the A operand is a pair too, and BOTH its tiles of a K block (hi, lo) sit in one stage and
// meet the same weight stage (two MMAs per K step), so a K block of 64 costs an SM 32 KB of A + 16 KB of B out of L2.
template <int BN, bool X2, int EPI, int NS, bool PAIR = false, int NW = NS, bool WIDE = false> struct FPlan {
  static_assert(!PAIR || X2, "the pair form splits [W_hi; W_lo] between the two CTAs");
  static_assert(!WIDE || (PAIR && X2 && BN == 128 && NW == NS && EPI == EPI_WIDE), "the wide form is a CTA-pair kernel");
  static constexpr int kXStage = kConvBytes * (WIDE ? 2 : 1);
  static constexpr int kWStage = BN * 128 * ((X2 && !PAIR) ? 2 : 1);
  static constexpr int kOffConv = 0;
  static constexpr int kOffW = NS * kXStage;
  static constexpr int kOffOut = kOffW + NW * kWStage; // epilogue staging tiles (TMA store), 1024-byte aligned
  static constexpr int kAuxTile = (BN / 64) * kConvBytes; // EPI_DX: hi blocks of the previous layer's activation pair, [128 rows][64] each
  static constexpr int kOffAux = kOffOut + 8 * kStageOutBytes;
  static constexpr int kOffCol = kOffAux + (EPI == EPI_DX ? 2 * kAuxTile : 0); // colscale[128], bias[128]
  static constexpr int kOffBar = kOffCol + 1024;
  static constexpr int kTotal = kOffBar + 512 + 1024; // 64 barrier slots, then the slack of the 1024-byte alignment
  static constexpr int kTmemCols = (4 * BN <= 256) ? 256 : 512;
  static_assert(kTotal <= 227 * 1024, "shared memory plan exceeds the SM");
};

template <bool PAIR = false>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (PAIR)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// fp32 accumulate, fp16 x fp16, both operands K-major, M = 128 (one CTA) or 256 (a CTA pair)
__host__ __device__ constexpr uint32_t make_idesc_f16(int n, int m = kFM) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int BN, bool X2, int EPI, int NS, bool PAIR, int NW, bool WIDE>
__global__ void __launch_bounds__(kFThreads, 1)
fwd16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWh,
             const __grid_constant__ CUtensorMap tmWl, const __grid_constant__ CUtensorMap tmOut,
             const __grid_constant__ CUtensorMap tmAux, const F16Params p) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  if (spec_skip(p.spec_st, p.spec)) return; // before any barrier / TMEM allocation (both CTAs of a pair read the same flag)
  using Plan = FPlan<BN, X2, EPI, NS, PAIR, NW, WIDE>;
  // work units: sample tiles (one CTA) or pairs of neighbouring sample tiles (a CTA pair: rank r takes tile 2 * unit + r);
  // WIDE: (pair of sample tiles, tile of BN output columns), the column tile running fastest
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_tiles = WIDE ? p.n_tiles : 1, k_slices = WIDE ? p.k_slices : 1;
  const int units = (PAIR ? (p.tiles + 1) >> 1 : p.tiles) * n_tiles * k_slices;
  auto tile_of = [&](int unit) { return WIDE ? 2 * (unit / (n_tiles * k_slices)) + (int)rank : (PAIR ? 2 * unit + (int)rank : unit); };
  auto nt_of = [&](int unit) { return WIDE ? (unit / k_slices) % n_tiles : 0; };
  auto ks_of = [&](int unit) { return WIDE ? unit % k_slices : 0; };
  auto kb_lo_of = [&](int unit) { return WIDE ? ks_of(unit) * p.kb_per_slice : 0; };
  auto kb_hi_of = [&](int unit) { return WIDE ? min(p.k_blocks, (ks_of(unit) + 1) * p.kb_per_slice) : p.k_blocks; };
  constexpr int kNC = NS, kNW = NW;
  static_assert(2 * kNC + 2 * kNW + 9 <= 64, "barrier table");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *bp = smem_raw + (base - smem_u32(smem_raw));
  auto conv_a = [&](int s) { return base + Plan::kOffConv + s * Plan::kXStage; };
  auto w_a = [&](int s, int lo) { return base + Plan::kOffW + s * Plan::kWStage + lo * (BN * 128); };
  const uint32_t bars = base + Plan::kOffBar;
  auto conv_full = [&](int s) { return bars + 8 * (s); };
  auto conv_empty = [&](int s) { return bars + 8 * (kNC + s); };
  auto w_full = [&](int s) { return bars + 8 * (2 * kNC + s); };
  // (equal ring depths: shared with conv_empty, the MMA warp's single commit per K block releases both tiles)
  auto w_empty = [&](int s) { return kNW == kNC ? conv_empty(s) : bars + 8 * (2 * kNC + kNW + s); };
  auto tm_full = [&](int b) { return bars + 8 * (2 * kNC + 2 * kNW + b); };
  auto tm_empty = [&](int b) { return bars + 8 * (2 * kNC + 2 * kNW + 2 + b); };
  auto aux_full = [&](int b) { return bars + 8 * (2 * kNC + 2 * kNW + 4 + b); };
  auto aux_empty = [&](int b) { return bars + 8 * (2 * kNC + 2 * kNW + 6 + b); };
  auto aux_a = [&](int b) { return base + Plan::kOffAux + b * Plan::kAuxTile; };
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(bp + Plan::kOffBar + 8 * (2 * kNC + 2 * kNW + 8));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_start = p.dbg ? clock64() : 0;
  if (p.dbg && threadIdx.x == 0) p.dbg[8 * blockIdx.x + 6] = (long long)globaltimer_ns();
  // PAIR: the "full" barriers and tm_empty that count are the LEADER's (same offsets in its shared memory)
  const uint32_t lead_off = PAIR ? mapa_shared(bars, 0) - bars : 0u;

  // ---- one-time setup -----------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWh);
    if (X2) tma_prefetch_desc(&tmWl);
    for (int s = 0; s < kNC; ++s) { mbar_init(conv_full(s), 1); mbar_init(conv_empty(s), 1); }
    for (int s = 0; s < kNW; ++s) { mbar_init(w_full(s), 1); if (kNW != kNC) mbar_init(w_empty(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tm_full(b), 1); mbar_init(tm_empty(b), (PAIR ? 2 : 1) * (kEpiThreads / 32)); // (PAIR: both CTAs' epilogue warps)
      mbar_init(aux_full(b), 1); mbar_init(aux_empty(b), kEpiThreads / 32);
    }
    if (EPI == EPI_DX) tma_prefetch_desc(&tmAux);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc_pair(smem_u32((const void *)tmem_slot), (uint32_t)Plan::kTmemCols);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)),
                   "r"((uint32_t)Plan::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  float *colsc = reinterpret_cast<float *>(bp + Plan::kOffCol), *biass = colsc + 128;
  if constexpr (!WIDE)
  for (int i = threadIdx.x; i < 128; i += kFThreads) {
    // EMIT16 (ReLU / Linear only): t_f act(z) = act(t_f z) for t_f > 0, so the pair's scale is folded into colscale and bias
    const float t = (EPI == EPI_EMIT16 && i < p.cols_valid) ? __ldg(p.tscale + i) : 1.0f;
    colsc[i] = (i < p.cols_valid) ? __ldg(p.colscale + i) * t : 0.0f;
    biass[i] = (i < p.cols_valid && EPI != EPI_DX) ? __ldg(p.bias + i) * t : 0.0f;
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int umma_n = WIDE ? BN : min(BN, (p.cols_valid + 31) & ~31);

  if (warp == 0) {
    { // ===== X producer: fp16 rows straight from HBM into the K-major SWIZZLE_128B operand tile (whole warp in the loop, one
      // elected lane issues: the TMA instructions then need no loop over the active lanes) ===========
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int unit = unit0; unit < units; unit += unit_step, ++it) {
        const int tile = tile_of(unit);
        if (EPI == EPI_DX) { // the hi blocks of A_prev's pair for this tile's act' (double-buffered like the accumulators)
          const int buf = it & 1;
          mbar_wait(aux_empty(buf), ((it >> 1) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(aux_full(buf), Plan::kAuxTile);
#pragma unroll
            for (int b = 0; b < BN / 64; ++b) tma_load_3d(aux_a(buf) + b * kConvBytes, &tmAux, aux_full(buf), 0, p.row0 + tile * kFM, b);
          }
          __syncwarp();
        }
        for (int kb = kb_lo_of(unit), kb_hi = kb_hi_of(unit); kb < kb_hi; ++kb) {
          mbar_wait(conv_empty(s), ph ^ 1);
          if (elect_one()) {
            if (p.diag & 1) { if (leader) mbar_arrive(conv_full(s)); }
            else if constexpr (WIDE) { // the hi and the lo tile of this K block, of both CTAs, counted on the leader's barrier
              if (leader) mbar_expect_tx(conv_full(s), 4 * kConvBytes);
              tma_load_3d_pair(conv_a(s), &tmX, conv_full(s) + lead_off, 0, kb, p.row0 + tile * kFM);
              tma_load_3d_pair(conv_a(s) + kConvBytes, &tmX, conv_full(s) + lead_off, 0, p.k_blocks + kb, p.row0 + tile * kFM);
            } else if constexpr (PAIR) { // both CTAs' tiles are counted on the leader's barrier
              if (leader) mbar_expect_tx(conv_full(s), 2 * kConvBytes);
              if (p.x_block_first) tma_load_3d_pair(conv_a(s), &tmX, conv_full(s) + lead_off, 0, kb, p.row0 + tile * kFM);
              else tma_load_3d_pair(conv_a(s), &tmX, conv_full(s) + lead_off, 0, p.row0 + tile * kFM, kb);
            } else {
              mbar_expect_tx(conv_full(s), kConvBytes);
              if (p.x_block_first) tma_load_3d(conv_a(s), &tmX, conv_full(s), 0, kb, p.row0 + tile * kFM);
              else tma_load_3d(conv_a(s), &tmX, conv_full(s), 0, p.row0 + tile * kFM, kb);
            }
          }
          __syncwarp();
          if (++s == kNC) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    { // ===== weight producer ===========================================================
      int s = 0;
      uint32_t ph = 0;
      for (int unit = unit0; unit < units; unit += unit_step) {
        const int wrow0 = nt_of(unit) * BN; // (WIDE: the rows of the B operand of this unit's column tile)
        for (int kb = kb_lo_of(unit), kb_hi = kb_hi_of(unit); kb < kb_hi; ++kb) {
          mbar_wait(w_empty(s), ph ^ 1);
          if (elect_one()) {
            if (p.diag & 2) { if (leader) mbar_arrive(w_full(s)); }
            else if constexpr (WIDE) { // the leader stages B_hi, the peer B_lo: [B_hi; B_lo] is the N = 256 operand of the pair
              if (leader) mbar_expect_tx(w_full(s), 2 * Plan::kWStage);
              tma_load_2d_pair(w_a(s, 0), leader ? &tmWh : &tmWl, w_full(s) + lead_off, kb * kFK, wrow0);
            } else if constexpr (PAIR) { // this CTA's half of B = [W_hi; W_lo]: rows 0 .. BN-1 (hi) in the leader, BN .. 2 BN-1 (lo) in the peer
              if (leader) mbar_expect_tx(w_full(s), 2 * Plan::kWStage);
              tma_load_2d_pair(w_a(s, 0), leader ? &tmWh : &tmWl, w_full(s) + lead_off, kb * kFK, 0);
            } else {
              mbar_expect_tx(w_full(s), Plan::kWStage);
              tma_load_2d(w_a(s, 0), &tmWh, w_full(s), kb * kFK, 0);
              if (X2) tma_load_2d(w_a(s, 1), &tmWl, w_full(s), kb * kFK, 0);
            }
          }
          __syncwarp();
          if (++s == kNW) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader) { // ===== MMA issuer (PAIR: the leader issues for both CTAs). The WHOLE warp runs the loop (uniform control flow,
      // operands in uniform registers) and one elected lane issues the tcgen05 instructions. =====================================
      // One MMA per K step: with the lo tile stored right behind the hi tile, B = [W_hi; W_lo] is a single 2*BN-row K-major
      // operand and D = [hi | lo] lands in adjacent TMEM columns. Descriptors are built once; per stage / K step only the
      // 14-bit start-address field moves.
      const uint32_t idesc = make_idesc_f16(X2 ? 2 * BN : umma_n, PAIR ? 2 * kFM : kFM);
      const uint64_t dA0 = desc_k_major(conv_a(0)), dB0 = desc_k_major(w_a(0, 0));
      int cs = 0, ws = 0, it = 0;
      uint32_t cph = 0, wph = 0;
      long long waited = 0, waited_w = 0, waited_tm = 0, t_mma = 0, t_commit = 0;
      if constexpr (WIDE) {
        // The tensor core's fp32 accumulate truncates (a bias of ~3e-8 of the accumulator per accumulating MMA): the contraction
        // is cut into chunks of p.k_chunk K blocks, each chunk accumulated from zero in one of the two TMEM buffers and added to
        // fp32 registers (round to nearest) by the epilogue warps while the next chunk runs in the other buffer.
        int cnt = 0;
        for (int unit = unit0; unit < units; unit += unit_step) {
          const int kb_hi = kb_hi_of(unit);
          for (int kb0 = kb_lo_of(unit); kb0 < kb_hi; kb0 += p.k_chunk, ++cnt) {
            const int buf = cnt & 1;
            const uint32_t d_acc = tmem_base + (uint32_t)(buf * 2 * BN);
            const long long tt0 = p.dbg ? clock64() : 0;
            mbar_wait(tm_empty(buf), ((cnt >> 1) & 1) ^ 1);
            if (p.dbg) waited_tm += clock64() - tt0;
            tc_fence_after();
            const int kb1 = min(kb_hi, kb0 + p.k_chunk);
            for (int kb = kb0; kb < kb1; ++kb) {
              const long long t0 = p.dbg ? clock64() : 0;
              mbar_wait(conv_full(cs), cph);
              const long long t0b = p.dbg ? clock64() : 0;
              mbar_wait(w_full(ws), wph);
              if (p.dbg) { const long long t1 = clock64(); waited += t0b - t0; waited_w += t1 - t0b; }
              tc_fence_after();
              const uint64_t da = dA0 + (uint64_t)(cs * (Plan::kXStage >> 4)), db = dB0 + (uint64_t)(ws * (Plan::kWStage >> 4));
              if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < kFK / 16; ++ks) {
                  umma_f16<PAIR>(d_acc, da + 2 * ks, db + 2 * ks, idesc, (kb > kb0 || ks > 0) ? 1u : 0u); // A_hi x [B_hi; B_lo]
                  umma_f16<PAIR>(d_acc, da + (kConvBytes >> 4) + 2 * ks, db + 2 * ks, idesc, 1u);          // A_lo x [B_hi; B_lo]
                }
                umma_commit_pair(conv_empty(cs)); // frees the stage (A and B tiles) in both CTAs
              }
              __syncwarp();
              if (++cs == kNC) { cs = 0; cph ^= 1; }
              if (++ws == kNW) { ws = 0; wph ^= 1; }
            }
            if (elect_one()) umma_commit_pair(tm_full(buf));
            __syncwarp();
          }
        }
        if (p.dbg && lane == 0) { p.dbg[8 * blockIdx.x + 3] = waited; p.dbg[8 * blockIdx.x + 4] = waited_w; p.dbg[8 * blockIdx.x + 5] = waited_tm; }
      }
      if constexpr (!WIDE)
      for (int unit = unit0; unit < units; unit += unit_step, ++it) {
        const int buf = it & 1;
        const uint32_t d_acc = tmem_base + (uint32_t)(buf * 2 * BN);
        const long long tt0 = p.dbg ? clock64() : 0;
        if (p.dbg && blockIdx.x == 0 && it < 8 && lane == 0) { p.dbg[4096 + 2 * it] = tt0; p.dbg[4096 + 2 * it + 1] = (long long)globaltimer_ns(); }
        mbar_wait(tm_empty(buf), ((it >> 1) & 1) ^ 1);
        if (p.dbg) waited_tm += clock64() - tt0;
        tc_fence_after();
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          const long long t0 = p.dbg ? clock64() : 0;
          mbar_wait(conv_full(cs), cph);
          const long long t0b = p.dbg ? clock64() : 0;
          mbar_wait(w_full(ws), wph);
          const long long t1 = p.dbg ? clock64() : 0;
          if (p.dbg) { waited += t0b - t0; waited_w += t1 - t0b; }
          tc_fence_after();
          const int nks = (p.diag & 4) ? 0 : min(kFK / 16, (p.k_total - kb * kFK + 15) / 16);
          const uint64_t da = dA0 + (uint64_t)(cs * (Plan::kXStage >> 4)), db = dB0 + (uint64_t)(ws * (Plan::kWStage >> 4));
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < kFK / 16; ++ks)
              if (ks < nks) umma_f16<PAIR>(d_acc, da + 2 * ks, db + 2 * ks, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
            if constexpr (PAIR) umma_commit_pair(conv_empty(cs)); // frees the stage in both CTAs
            else umma_commit(conv_empty(cs)); // (== w_empty(ws) when both rings have the same depth)
            if constexpr (kNW != kNC) { if constexpr (PAIR) umma_commit_pair(w_empty(ws)); else umma_commit(w_empty(ws)); }
          }
          __syncwarp();
          if (p.dbg) t_mma += clock64() - t1;
          if (++cs == kNC) { cs = 0; cph ^= 1; }
          if (++ws == kNW) { ws = 0; wph ^= 1; }
        }
        if (elect_one()) {
          if constexpr (PAIR) umma_commit_pair(tm_full(buf));
          else umma_commit(tm_full(buf));
        }
        __syncwarp();
      }
      if (p.dbg && blockIdx.x == 0 && it < 8 && lane == 0) {
        p.dbg[4096 + 2 * it] = clock64(); p.dbg[4096 + 2 * it + 1] = (long long)globaltimer_ns(); p.dbg[4096 + 16] = it;
        p.dbg[4096 + 20] = t_mma; p.dbg[4096 + 21] = t_commit;
      }
      if (p.dbg && lane == 0) { p.dbg[8 * blockIdx.x + 3] = waited; p.dbg[8 * blockIdx.x + 4] = waited_w; p.dbg[8 * blockIdx.x + 5] = waited_tm; }
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
      if constexpr (WIDE) {
        int cnt = 0;
        for (int unit = unit0; unit < units; unit += unit_step) {
          const int tile = tile_of(unit);
          float accr[BN / 64][32]; // this thread's row of the tile, its BN / 64 chunks of 32 columns, summed over the K chunks in RN fp32
#pragma unroll
          for (int i = 0; i < BN / 64; ++i)
#pragma unroll
            for (int j = 0; j < 32; ++j) accr[i][j] = 0.0f;
          for (int kb0 = kb_lo_of(unit), kb_hi = kb_hi_of(unit); kb0 < kb_hi; kb0 += p.k_chunk, ++cnt) {
            const int buf = cnt & 1;
            const long long t0 = p.dbg ? clock64() : 0;
            mbar_wait(tm_full(buf), (cnt >> 1) & 1);
            if (p.dbg) waiting += clock64() - t0;
            tc_fence_after();
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * BN);
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) {
              const int c0 = half * 32 + 64 * i;
              uint32_t v[32], w[32];
              tmem_ld32_nowait(lane_addr + c0, v);
              tmem_ld32_nowait(lane_addr + BN + c0, w);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) accr[i][j] += __uint_as_float(v[j]) + __uint_as_float(w[j]);
            }
            tc_fence_before(); // chunk consumed: the issuer may start chunk cnt + 2 in this buffer
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tm_empty(buf) + lead_off);
          }
          // out = acc / (S_A S_B) [+ bias -> activation | * act'(A_prev)]; the row of ones of a dW's A operand is unscaled
          const long grow = (long)tile * kFM + q * 32 + lane;
          const float sb = __ldg(p.sb_inv), sc = (grow == (long)p.ones_row) ? sb : __ldg(p.sa_inv) * sb;
          const bool rok = grow < (long)p.rows_valid;
#pragma unroll
          for (int i = 0; i < BN / 64; ++i) {
            const int col0 = nt_of(unit) * BN + half * 32 + 64 * i;
            const float4 *ap = reinterpret_cast<const float4 *>(p.aux32 + (rok ? grow : 0) * p.ld_aux + col0);
            if (lane == 0) tma_store_wait_read(); // the previous block of this warp has left the staging tile
            __syncwarp();
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              float4 r;
              r.x = accr[i][4 * qq + 0] * sc; r.y = accr[i][4 * qq + 1] * sc; r.z = accr[i][4 * qq + 2] * sc; r.w = accr[i][4 * qq + 3] * sc;
              if (p.aux32) {
                const float4 a = rok ? __ldg(ap + qq) : make_float4(0.f, 0.f, 0.f, 0.f);
                r.x *= act_deriv_c<ACT>(p.act, a.x); r.y *= act_deriv_c<ACT>(p.act, a.y);
                r.z *= act_deriv_c<ACT>(p.act, a.z); r.w *= act_deriv_c<ACT>(p.act, a.w);
              } else if (p.bias) {
                const float *bp4 = p.bias + col0 + 4 * qq; // (the same address in every lane: one broadcast load each)
                r.x = act_apply_c<ACT>(p.act, r.x + __ldg(bp4)); r.y = act_apply_c<ACT>(p.act, r.y + __ldg(bp4 + 1));
                r.z = act_apply_c<ACT>(p.act, r.z + __ldg(bp4 + 2)); r.w = act_apply_c<ACT>(p.act, r.w + __ldg(bp4 + 3));
              }
              *reinterpret_cast<float4 *>(stage_p + lane * 128 + ((qq ^ (lane & 7)) << 4)) = r;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tmOut, stage_a, col0, tile * kFM + q * 32, ks_of(unit));
              tma_store_commit();
            }
          }
        }
        if (lane == 0) tma_store_wait_all();
        if (p.dbg && threadIdx.x == kEpiWarp0 * 32) p.dbg[8 * blockIdx.x + 2] = waiting;
      }
      if constexpr (!WIDE)
      for (int unit = unit0; unit < units; unit += unit_step, ++it) {
        const int tile = tile_of(unit); // (PAIR, odd tile count: the peer's last tile lies past the batch; its stores are clipped)
        const int buf = it & 1;
        const long long t0 = p.dbg ? clock64() : 0;
        mbar_wait(tm_full(buf), (it >> 1) & 1);
        tc_fence_after();
        if (EPI == EPI_DX) mbar_wait(aux_full(buf), (it >> 1) & 1);
        const long long t1 = p.dbg ? clock64() : 0;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * BN);
        float s_in = 1.0f, s_out = 1.0f;
        if (EPI == EPI_DX) { s_in = __ldg(p.scale_in_inv); s_out = __ldg(p.scale_out); }
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) {
          const int c0 = half * 32 + 64 * i;
          if (c0 < p.cols_valid && !(p.diag & 16)) {
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
            if constexpr (EPI == EPI_F32) {
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
              if (lane == 0 && !(p.diag & 8)) {
                tma_store_2d(&tmOut, stage_a, c0, tile * kFM + q * 32);
                tma_store_commit();
              }
            } else {
              // the 32 values of this thread's row as a scaled fp16 pair: hi = fp16(x), lo = fp16(x - hi). Staged as two plain
              // [32 rows][32 halves] blocks (64-byte rows; the few bank conflicts of these 8 stores per block do not show) and
              // stored by TMA: EMIT16 into blocks c0/64 (hi) and nb_out + c0/64 (lo) of the block-major pair, DX into columns
              // c0 (hi) and cols + c0 (lo) of the row-major pair
              float x[32];
              if constexpr (EPI == EPI_EMIT16) {
#pragma unroll
                for (int qq = 0; qq < 8; ++qq) {
                  const float4 s4 = colsc4[c0 / 4 + qq], b4 = bias4[c0 / 4 + qq]; // (both carry t_f)
                  x[4 * qq + 0] = act_apply_c<ACT>(p.act, fmaf(__uint_as_float(v[4 * qq + 0]), s4.x, b4.x));
                  x[4 * qq + 1] = act_apply_c<ACT>(p.act, fmaf(__uint_as_float(v[4 * qq + 1]), s4.y, b4.y));
                  x[4 * qq + 2] = act_apply_c<ACT>(p.act, fmaf(__uint_as_float(v[4 * qq + 2]), s4.z, b4.z));
                  x[4 * qq + 3] = act_apply_c<ACT>(p.act, fmaf(__uint_as_float(v[4 * qq + 3]), s4.w, b4.w));
                }
              } else { // EPI_DX: (delta W^T) * colscale / S_in, * act'(A_prev) read from the hi block of its pair, * S_out
                const int row = q * 32 + lane;
                const uint8_t *ap = bp + Plan::kOffAux + buf * Plan::kAuxTile + (c0 >> 6) * kConvBytes + row * 128;
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8) {
                  const uint4 hv = *reinterpret_cast<const uint4 *>(ap + (((((c0 & 63) >> 3) + k8) ^ (row & 7)) << 4));
                  const __half2 *h2 = reinterpret_cast<const __half2 *>(&hv);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 af = __half22float2(h2[e]);
                    const int j = 8 * k8 + 2 * e;
                    const float g0 = __uint_as_float(v[j]) * colsc[c0 + j] * s_in, g1 = __uint_as_float(v[j + 1]) * colsc[c0 + j + 1] * s_in;
                    x[j] = (ACT == B200_ACT_RELU ? (af.x > 0.0f ? g0 : 0.0f) : g0) * s_out;
                    x[j + 1] = (ACT == B200_ACT_RELU ? (af.y > 0.0f ? g1 : 0.0f) : g1) * s_out;
                  }
                }
              }
#pragma unroll
              for (int k8 = 0; k8 < 4; ++k8) {
                __align__(16) __half2 hh[4], ll[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float x0 = x[8 * k8 + 2 * e], x1 = x[8 * k8 + 2 * e + 1];
                  const __half2 h = __floats2half2_rn(x0, x1);
                  const float2 hf = __half22float2(h);
                  hh[e] = h;
                  ll[e] = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
                }
                *reinterpret_cast<uint4 *>(stage_p + lane * 64 + k8 * 16) = *reinterpret_cast<const uint4 *>(hh);
                *reinterpret_cast<uint4 *>(stage_p + 2048 + lane * 64 + k8 * 16) = *reinterpret_cast<const uint4 *>(ll);
              }
              fence_async_smem();
              __syncwarp();
              if (lane == 0 && !(p.diag & 8)) {
                if constexpr (EPI == EPI_EMIT16) {
                  tma_store_3d(&tmOut, stage_a, c0 & 63, tile * kFM + q * 32, c0 >> 6);
                  tma_store_3d(&tmOut, stage_a + 2048, c0 & 63, tile * kFM + q * 32, p.nb_out + (c0 >> 6));
                } else {
                  tma_store_2d(&tmOut, stage_a, c0, tile * kFM + q * 32);
                  tma_store_2d(&tmOut, stage_a + 2048, p.cols_valid + c0, tile * kFM + q * 32);
                }
                tma_store_commit();
              }
            }
          }
        }
        tc_fence_before(); // accumulator consumed: the issuer may start tile it+2 in this buffer
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster(tm_empty(buf) + lead_off); // the leader's issuer waits for both CTAs' epilogues
          else mbar_arrive(tm_empty(buf));
          if (EPI == EPI_DX) mbar_arrive(aux_empty(buf));
        }
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
  if constexpr (PAIR) cluster_sync_all(); // neither CTA leaves while the other may still signal its barriers or feed its MMAs
  else __syncthreads();
  if (p.dbg && threadIdx.x == 0) { p.dbg[8 * blockIdx.x] = clock64() - t_start; p.dbg[8 * blockIdx.x + 7] = (long long)globaltimer_ns(); }
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, (uint32_t)Plan::kTmemCols);
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Plan::kTmemCols) : "memory");
  }
}

// Operand preparation, once per evaluation. A job turns one weight matrix (elements (r, n) at W[r * sr + n * sn]: r the
// contraction index, n the output row) into the scaled fp16 pair wh, wl [rows n][ldk] (contraction index contiguous):
//     v = W(r, n) * rscale[r]          (rscale: 1 / t_r of the layer that produced the A operand, see rs_mode; or none)
//     s_n = 2^(14 - e), max_r |v| in [2^(e-1), 2^e)  =>  s_n * max |v| in [2^13, 2^14)   (fp16 overflows at 65504)
//     hi = fp16(s_n v), lo = fp16(s_n v - hi), colscale[n] = pre / s_n; with dup = 2 the row is written twice, [v | v]: the hi
//     and the lo blocks of a K-stacked A operand meet the same weights.
// One CTA of 1024 threads per 8 output rows (one 32-byte sector per weight row when sn == 1), 128 slices of the contraction
// index per row: every weight is loaded ONCE (<= 8 per thread, all in flight together), kept in registers across the
// reductions, split, and written out through a shared-memory transpose.
constexpr int kSplitNeurons = 8, kSplitMaxK = 1024;
struct PrepJob {
  const float *W;
  int R, Nn;            // contraction length, output rows
  long sr, sn;
  int ldk, dup;
  float pre;
  __half *wh, *wl;
  float *colscale;
  // rscale computed here (so that this job does not wait for another launch): the A operand is the pair of the PREVIOUS layer's
  // activations, whose per-feature scale t_k comes from the bound sum_r |Wp(r, k)| + |bp_k| over that layer's weights Wp
  // ([rs_R][128], k contiguous; inputs in [0, 1]) — mode 1 — or is 1 for a bounded activation (mode 2). Every CTA of the job
  // computes all 128 scales with the same arithmetic (identical results); the first one publishes t_k and 1 / t_k.
  // Mode 3 (default when the launch also prepares that previous layer, job 0): the job-0 CTAs, which hold the previous layer's
  // weights in registers anyway, publish t_k and 1 / t_k of their 8 neurons (pub_*) and count themselves in sync[0]; the CTAs of
  // this job fetch their own weights, wait for the count (all CTAs of the launch are co-resident: at most ~45 of 1024 threads)
  // and read the 128 scales — instead of every CTA re-reading all of W_0 (401 KB through one SM's L2 port: ~8 us of a 17 us kernel).
  int rs_mode;
  const float *rs_W, *rs_bias;
  int rs_R;
  float *rs_tscale, *rs_tinv;
  float *pub_tscale, *pub_tinv; // job 0 as the publisher (nullptr: not)
  const float *pub_bias;
  int nblocks;          // CTAs of this job
};
// sync: {publishers that have finished, CTAs of the jobs that have finished}; the last CTA to finish zeroes both
__global__ void __launch_bounds__(1024) prep_w16_kernel(const __grid_constant__ PrepJob j0, const __grid_constant__ PrepJob j1,
                                                      const __grid_constant__ PrepJob j2, const SpecState *spec_st, int spec,
                                                      const __grid_constant__ ChainW chain, unsigned *sync, double *host_err) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  if (spec_skip(spec_st, spec)) return;
  __shared__ float red[128][kSplitNeurons + 1];
  if (chain.nl > 0 && blockIdx.x >= gridDim.x - chain.nctas) { // the extra CTAs (see ChainW, network.cuh)
    chain_cw_block(chain, (int)(blockIdx.x - (gridDim.x - chain.nctas)), &red[0][0]);
    return;
  }
  const int which = (int)blockIdx.x < j0.nblocks ? 0 : ((int)blockIdx.x < j0.nblocks + j1.nblocks ? 1 : 2);
  const PrepJob &j = *(which == 0 ? &j0 : (which == 1 ? &j1 : &j2)); // (a pointer into the parameter space: no local copy)
  __shared__ float sc[kSplitNeurons], rs_sh[128];
  __shared__ __align__(16) __half th[kSplitNeurons][kSplitMaxK + 8], tl[kSplitNeurons][kSplitMaxK + 8];
  const int o = threadIdx.x & (kSplitNeurons - 1), kq = threadIdx.x / kSplitNeurons;
  const int o0 = ((int)blockIdx.x - (which == 0 ? 0 : (which == 1 ? j0.nblocks : j0.nblocks + j1.nblocks))) * kSplitNeurons;
  const bool ok = o0 + o < j.Nn;
  float v[kSplitMaxK / 128]; // this thread's weights: in flight while the scales below are computed / waited for
#pragma unroll
  for (int i = 0; i < kSplitMaxK / 128; ++i) {
    const int k = kq + 128 * i;
    v[i] = (ok && k < j.R) ? __ldg(j.W + (size_t)k * j.sr + (size_t)(o0 + o) * j.sn) : 0.0f;
  }
  if (j.rs_mode == 3) { // block-uniform: the scales come from the job-0 CTAs of this launch
    if (threadIdx.x == 0) {
      const volatile unsigned *c = sync;
      unsigned long long polls = 0;
      while (*c < (unsigned)j0.nblocks && ++polls < (1ull << 22)) __nanosleep(64); // (bounded: never hang the device)
      if (*c < (unsigned)j0.nblocks && host_err) { *host_err = 2.0; __threadfence_system(); } // reported at the caller's next synchronisation
      __threadfence();
    }
    __syncthreads();
    if (threadIdx.x < 128) rs_sh[threadIdx.x] = __ldcg(j.rs_tinv + threadIdx.x);
    __syncthreads();
  } else if (j.rs_mode) { // block-uniform
    // 32 slices of the contraction index x 32 groups of four features: every thread's (<= 32) 16-byte loads are in flight in
    // two batches, the partial sums meet in shared memory (the split tiles' space, not yet in use)
    float *part = reinterpret_cast<float *>(&th[0][0]); // [32 slices][128 features]
    static_assert(sizeof(th) >= 32 * 128 * sizeof(float), "partial sums of the feature bounds");
    const int k4 = (threadIdx.x & 31) * 4, slice = threadIdx.x >> 5;
    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool w_vec = (reinterpret_cast<uintptr_t>(j.rs_W) & 15u) == 0; // (the caller's parameter buffer)
    if (j.rs_mode == 1) {
      for (int r0 = slice; r0 < j.rs_R; r0 += 32 * 16) {
        float4 t[16];
        if (w_vec) { // sixteen 16-byte loads in flight (pinned: the compiler otherwise pairs each load with its add, one L2 round
                     // trip per row; rows past the end re-read the last row and are masked out)
#pragma unroll
          for (int u = 0; u < 16; ++u) t[u] = ldg4_pinned(reinterpret_cast<const float4 *>(j.rs_W + (size_t)min(r0 + 32 * u, j.rs_R - 1) * 128 + k4));
          B200_PIN16_F4(t);
#pragma unroll
          for (int u = 0; u < 16; ++u) if (r0 + 32 * u >= j.rs_R) t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float *q = j.rs_W + (size_t)min(r0 + 32 * u, j.rs_R - 1) * 128 + k4;
            t[u] = (r0 + 32 * u >= j.rs_R) ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
          }
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) { a4.x += fabsf(t[u].x); a4.y += fabsf(t[u].y); a4.z += fabsf(t[u].z); a4.w += fabsf(t[u].w); }
      }
    }
    *reinterpret_cast<float4 *>(part + slice * 128 + k4) = a4;
    __syncthreads();
    const int k = threadIdx.x & 127;
    if (threadIdx.x < 128) {
      float bound = 1.0f;
      if (j.rs_mode == 1) {
        float b1 = 0.0f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) b1 += (part[i * 128 + k] + part[(i + 1) * 128 + k]) + (part[(i + 2) * 128 + k] + part[(i + 3) * 128 + k]);
        bound = b1 * 1.0001f + fabsf(__ldg(j.rs_bias + k)) + 1e-30f; // (fp32 summation slack)
      }
      int eb = 0;
      frexpf(bound, &eb);
      eb = max(-100, min(100, eb));
      rs_sh[k] = ldexpf(1.0f, eb - 14);
      if (o0 == 0) { j.rs_tscale[k] = ldexpf(1.0f, 14 - eb); j.rs_tinv[k] = ldexpf(1.0f, eb - 14); }
    }
    __syncthreads();
  }
  float amax = 0.0f;
#pragma unroll
  for (int i = 0; i < kSplitMaxK / 128; ++i) {
    if (j.rs_mode) { const int k = kq + 128 * i; v[i] *= (k < 128) ? rs_sh[k] : 0.0f; } // (R <= 128 on this path)
    amax = fmaxf(amax, fabsf(v[i]));
  }
  red[kq][o] = amax;
  __shared__ float reds[128][kSplitNeurons + 1], red2s[8][kSplitNeurons];
  const bool publish = which == 0 && j.pub_tscale != nullptr; // block-uniform
  if (publish) {
    float asum = 0.0f;
#pragma unroll
    for (int i = 0; i < kSplitMaxK / 128; ++i) asum += fabsf(v[i]);
    reds[kq][o] = asum;
  }
  __syncthreads();
  // max over the 128 slices of each of the 8 rows: 256 threads take four slices each, shuffles combine the four parts a warp
  // holds per row, the eight warps meet in shared memory (a 128-step serial scan per row cost ~2 us of a ~8 us kernel)
  __shared__ float red2[8][kSplitNeurons];
  if (threadIdx.x < 256) {
    const int oo = threadIdx.x & (kSplitNeurons - 1), part = threadIdx.x >> 3; // 32 parts of 4 slices
    float m = fmaxf(fmaxf(red[4 * part][oo], red[4 * part + 1][oo]), fmaxf(red[4 * part + 2][oo], red[4 * part + 3][oo]));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
    if ((threadIdx.x & 31) < kSplitNeurons) red2[threadIdx.x >> 5][oo] = m;
    if (publish) {
      float a = (reds[4 * part][oo] + reds[4 * part + 1][oo]) + (reds[4 * part + 2][oo] + reds[4 * part + 3][oo]);
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      if ((threadIdx.x & 31) < kSplitNeurons) red2s[threadIdx.x >> 5][oo] = a;
    }
  }
  __syncthreads();
  if (threadIdx.x < kSplitNeurons) {
    float m = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) m = fmaxf(m, red2[i][threadIdx.x]);
    int e = 0;
    if (m > 0.0f && m < 3.0e38f) frexpf(m, &e); // m = f * 2^e, f in [0.5, 1)
    e = max(-100, min(100, e));
    sc[threadIdx.x] = ldexpf(1.0f, 14 - e);      // s * m in [2^13, 2^14)
    const int n = o0 + threadIdx.x;
    if (n < j.Nn) j.colscale[n] = j.pre * ldexpf(1.0f, e - 14);
    if (publish && n < j.Nn) { // t_n of this layer's output n (the next layer's input feature): bound sum_k |W(k, n)| + |b_n|
      float b1 = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) b1 += red2s[i][threadIdx.x];
      const float bound = b1 * 1.0001f + fabsf(__ldg(j.pub_bias + n)) + 1e-30f; // (fp32 summation slack)
      int eb = 0;
      frexpf(bound, &eb);
      eb = max(-100, min(100, eb));
      j.pub_tscale[n] = ldexpf(1.0f, 14 - eb);
      j.pub_tinv[n] = ldexpf(1.0f, eb - 14);
    }
  }
  __syncthreads();
  if (publish && threadIdx.x == 0) { // (the barrier above ordered the eight writers before this thread)
    __threadfence();
    atomicAdd(sync, 1u);
  }
  const float s = sc[o];
#pragma unroll
  for (int i = 0; i < kSplitMaxK / 128; ++i) {
    const int k = kq + 128 * i;
    if (k < j.R) {
      const float x = v[i] * s;
      const __half h = __float2half_rn(x);
      const __half l = __float2half_rn(x - __half2float(h));
      th[o][k] = h;
      tl[o][k] = l;
      if (j.dup == 2) { th[o][j.R + k] = h; tl[o][j.R + k] = l; }
    }
  }
  for (int k = j.dup * j.R + kq; k < j.ldk; k += 128) { th[o][k] = __float2half_rn(0.0f); tl[o][k] = __float2half_rn(0.0f); } // row padding
  __syncthreads();
  const int vec_per_row = j.ldk / 8; // ldk is a multiple of 8: 16-byte stores
  for (int idx = threadIdx.x; idx < kSplitNeurons * vec_per_row; idx += 1024) {
    const int oo = idx / vec_per_row, kv = idx - oo * vec_per_row;
    if (o0 + oo < j.Nn) {
      *reinterpret_cast<uint4 *>(j.wh + (size_t)(o0 + oo) * j.ldk + 8 * kv) = *reinterpret_cast<const uint4 *>(&th[oo][8 * kv]);
      *reinterpret_cast<uint4 *>(j.wl + (size_t)(o0 + oo) * j.ldk + 8 * kv) = *reinterpret_cast<const uint4 *>(&tl[oo][8 * kv]);
    }
  }
  if (sync) { // the last CTA of the jobs to finish re-arms the two counters for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(sync + 1, 1u) == (unsigned)(j0.nblocks + j1.nblocks + j2.nblocks) - 1u) {
        sync[0] = 0u;
        sync[1] = 0u;
        __threadfence();
      }
    }
  }
}

// fp32 A_1 from its fp16 pair (debug read-back of a hidden activation that only exists as the pair)
__global__ void __launch_bounds__(256) pair_to_f32_kernel(const __half *__restrict__ a16, long rows, int width, const float *__restrict__ tinv,
                                                          float *__restrict__ out) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  const int nb = width / 64;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * width; i += (long)gridDim.x * blockDim.x) {
    const long r = i / width;
    const int f = (int)(i - r * width);
    const float hi = __half2float(a16[((long)(f >> 6) * rows + r) * 64 + (f & 63)]);
    const float lo = __half2float(a16[((long)(nb + (f >> 6)) * rows + r) * 64 + (f & 63)]);
    out[i] = (hi + lo) * tinv[f];
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return (EncodeTiledFn) nullptr;
    return (EncodeTiledFn)f;
  }();
  if (!fn) set_error("cuTensorMapEncodeTiled is not available from the driver");
  return fn;
}

int make_map_2d(CUtensorMap *tm, CUtensorMapDataType dt, const void *ptr, unsigned long long dim0, unsigned long long dim1,
                unsigned long long stride_bytes, unsigned box0, unsigned box1, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return B200_ERR_CUDA;
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

// fp16 {dim0 contiguous, dim1, dim2}, strides in bytes for dim1 / dim2, box {box0, box1, box2}
int make_map_3d_ex(CUtensorMap *tm, const void *ptr, unsigned long long dim0, unsigned long long dim1, unsigned long long dim2,
                   unsigned long long stride1, unsigned long long stride2, unsigned box0, unsigned box1, unsigned box2,
                   CUtensorMapSwizzle sw, CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return B200_ERR_CUDA;
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {stride1, stride2};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(tm, dt, 3, const_cast<void *>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d fp16) failed (%d): dims %llu x %llu x %llu strides %llu, %llu ptr %p", (int)r, dim0, dim1, dim2,
              stride1, stride2, ptr);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}
// the block-major fp16 copies [blocks][rows][64]: dense strides, box {box0, box1, 1}
int make_map_3d(CUtensorMap *tm, const void *ptr, unsigned long long dim0, unsigned long long dim1, unsigned long long dim2,
                unsigned box0, unsigned box1, CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B) {
  return make_map_3d_ex(tm, ptr, dim0, dim1, dim2, dim0 * 2, dim0 * dim1 * 2, box0, box1, 1, sw);
}

template <int BN, bool X2, int EPI, int NS, bool PAIR = false, int NW = NS, bool WIDE = false>
int launch_fwd16(const CUtensorMap &tx, const CUtensorMap &twh, const CUtensorMap &twl, const CUtensorMap &tout, const CUtensorMap &taux,
                 const F16Params &p, int grid, cudaStream_t st) {
  auto kern = fwd16_kernel<BN, X2, EPI, NS, PAIR, NW, WIDE>;
  constexpr int smem = FPlan<BN, X2, EPI, NS, PAIR, NW, WIDE>::kTotal;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  static long long *dbg = nullptr;
  const bool timing = env().tc_timing;
  F16Params pp = p;
  pp.diag = env().diag;
  if (timing) {
    if (!dbg) B200_CUDA(cudaMalloc(&dbg, sizeof(long long) * 8 * 1024));
    B200_CUDA(cudaMemsetAsync(dbg, 0, sizeof(long long) * 8 * 1024, st));
    pp.dbg = dbg;
  }
  B200_CUDA(launch_ex(kern, dim3((unsigned)grid), dim3(kFThreads), (size_t)smem, st, PAIR ? 2 : 1, tx, twh, twl, tout, taux, pp)); // (PAIR: clusters of two CTAs = the two SMs of a TPC)
  g_launches.fetch_add(1, std::memory_order_relaxed);
  B200_CUDA(cudaGetLastError());
  if (timing) {
    std::vector<long long> h(8 * 1024);
    B200_CUDA(cudaMemcpyAsync(h.data(), dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_first = h[6], t_last = h[7], t_first_end = h[7], t_last_start = h[6];
    for (int i = 0; i < grid; ++i) {
      for (int j = 0; j < 6; ++j) a[j] += (double)h[8 * i + j] / grid;
      a[6] = 0; a[7] += (double)(h[8 * i + 7] - h[8 * i + 6]) / grid; // (ns of this CTA)
      t_first = std::min(t_first, h[8 * i + 6]); t_last = std::max(t_last, h[8 * i + 7]);
      t_first_end = std::min(t_first_end, h[8 * i + 7]); t_last_start = std::max(t_last_start, h[8 * i + 6]);
    }
    fprintf(stderr, "[fwd16 timing] BN %d x2 %d epi %d pair %d grid %d tiles %d: per CTA total %.0f clk = %.0f ns (%.2f clk/ns) | epilogue busy %.0f, "
            "waiting %.0f | issuer waits: X %.0f, weights %.0f, tmem %.0f | first CTA start -> last CTA end %.1f us, starts spread %.1f us, "
            "ends spread %.1f us\n", BN, (int)X2, EPI, (int)PAIR, grid, p.tiles, a[0], a[7] - a[6], a[0] / (a[7] - a[6]), a[1], a[2], a[3], a[4], a[5],
            (t_last - t_first) * 1e-3, (t_last_start - t_first) * 1e-3, (t_last - t_first_end) * 1e-3);
    fprintf(stderr, "   CTA 0 issuer, per tile: ");
    for (int i = 0; i < (int)h[4096 + 16] && i < 7; ++i)
      fprintf(stderr, "%lld clk / %lld ns; ", h[4096 + 2 * (i + 1)] - h[4096 + 2 * i], h[4096 + 2 * (i + 1) + 1] - h[4096 + 2 * i + 1]);
    fprintf(stderr, " | CTA 0 issuer: fence + MMA issue %lld clk, commit %lld clk\n", h[4096 + 20], h[4096 + 21]);
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

// ---- hidden layer 1 of a three-layer net on the fp16 kernels (b200_net::Mid16) ----------------------------------------------
bool mid16_applicable(const b200_net *net) {
  if (env().mid16 == 0 || env().fwd16 == 0 || env().dw16 == 0 || env().tail == 0) return false;
  if (net->nlayers() != 3 || net->prec != B200_PREC_TF32X3) return false;
  const int N0 = net->dims[1], N1 = net->dims[2];
  // A_1's act' is read back from the hi half of its pair: exact for ReLU / Linear only (the sign of a, or nothing)
  const bool act0_ok = net->acts[0] == B200_ACT_RELU || net->acts[0] == B200_ACT_LINEAR;
  // (N0 == 128: the hi and the lo halves of A_1's pair are exactly the two M tiles of a dW CTA)
  return fwd16_shape_ok(net) && dw16_applicable(net) && tail_applicable(net) && act0_ok && N0 == 128 && (N1 == 64 || N1 == 128);
}

int mid16_ensure(b200_net *net, long batch) {
  b200_net::Mid16 &m = net->m16;
  const int N0 = net->dims[1], N1 = net->dims[2];
  if (!m.tscale) {
    B200_CUDA(cudaMalloc(&m.tscale, sizeof(float) * (2 * N0 + N1 + N0 + 4)));
    m.tinv = m.tscale + N0;
    m.colscale_f = m.tinv + N0;
    m.colscale_d = m.colscale_f + N1;
    m.scale1_inv = m.colscale_d + N0;
    B200_CUDA(cudaMalloc(&m.wfh, sizeof(__half) * (size_t)N1 * 2 * N0));
    B200_CUDA(cudaMalloc(&m.wfl, sizeof(__half) * (size_t)N1 * 2 * N0));
    B200_CUDA(cudaMalloc(&m.wdh, sizeof(__half) * (size_t)N0 * 2 * N1));
    B200_CUDA(cudaMalloc(&m.wdl, sizeof(__half) * (size_t)N0 * 2 * N1));
    B200_CUDA(cudaMalloc(&m.db_part, sizeof(float) * (size_t)2 * net->ctx->num_sms * N1));
    B200_CUDA(cudaMalloc(&m.prep_sync, sizeof(unsigned) * 2));
    B200_CUDA(cudaMemsetAsync(m.prep_sync, 0, sizeof(unsigned) * 2, net->ctx->stream)); // (ordered before the first launch that counts in it)
    ++net->config_gen;
  }
  if (m.a16_rows < net->cap || m.d16_rows < net->cap) {
    if (m.a16) cudaFree(m.a16);
    if (m.d16) cudaFree(m.d16);
    m.a16 = m.d16 = nullptr;
    B200_CUDA(cudaMalloc(&m.a16, sizeof(__half) * 2 * (size_t)N0 * net->cap));
    B200_CUDA(cudaMalloc(&m.d16, sizeof(__half) * 2 * (size_t)N1 * net->cap));
    m.a16_rows = m.d16_rows = net->cap;
    ++net->config_gen;
  }
  (void)batch;
  return B200_OK;
}

void mid16_release(b200_net *net) {
  b200_net::Mid16 &m = net->m16;
  for (void *p : {(void *)m.tscale, m.wfh, m.wfl, m.wdh, m.wdl, (void *)m.db_part, m.a16, m.d16, (void *)m.prep_sync})
    if (p) cudaFree(p);
  m = b200_net::Mid16{};
}

int mid16_reconstruct_act0(b200_net *net) {
  b200_net::Mid16 &m = net->m16;
  if (!m.act0_stale || !m.a16 || net->last_batch <= 0) return B200_OK;
  const long total = net->last_batch * net->dims[1];
  B200_LAUNCH(pair_to_f32_kernel, (int)std::min<long>(1184, (total + 255) / 256), 256, 0, net->ctx->stream, (const __half *)m.a16,
              net->last_batch, net->dims[1], m.tinv, net->act[0]);
  m.act0_stale = false;
  return B200_OK;
}

// Per evaluation, before the forward sweep: the scaled fp16 {hi, lo} split of W_0 (its own launch, so that it is timed and
// profiled apart from the GEMM) and, on the mid16 path, the operands of layer 1. No-op when the fp16 forward does not apply.
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
  const float *W0 = params + net->offs[0];
  PrepJob j0{};
  j0.W = W0; j0.R = K; j0.Nn = N; j0.sr = N; j0.sn = 1; j0.ldk = ldk; j0.dup = 1; j0.pre = 1.0f / 255.0f;
  j0.wh = (__half *)net->w16h; j0.wl = (__half *)net->w16l; j0.colscale = net->colscale;
  j0.nblocks = ceil_div(N, kSplitNeurons);
  PrepJob j1{}, j2{}; // (no CTAs)
  unsigned *sync = nullptr;
  if (net->m16.on) {
    b200_net::Mid16 &m = net->m16;
    const int N1 = net->dims[2];
    const float *W1 = params + net->offs[1];
    // dX operand of layer 1: rows = layer-1 inputs f, contraction over its outputs o (W_1 is [f][o], o contiguous)
    j1.W = W1; j1.R = N1; j1.Nn = N; j1.sr = 1; j1.sn = N1; j1.ldk = 2 * N1; j1.dup = 2; j1.pre = 1.0f;
    j1.wh = (__half *)m.wdh; j1.wl = (__half *)m.wdl; j1.colscale = m.colscale_d;
    j1.nblocks = ceil_div(N, kSplitNeurons);
    // forward operand of layer 1, rows scaled by 1 / t_k of layer 0's outputs (computed inside the job from W_0: N == 128)
    j2.W = W1; j2.R = N; j2.Nn = N1; j2.sr = N1; j2.sn = 1; j2.ldk = 2 * N; j2.dup = 2; j2.pre = 1.0f;
    j2.wh = (__half *)m.wfh; j2.wl = (__half *)m.wfl; j2.colscale = m.colscale_f;
    j2.rs_mode = (net->acts[0] == B200_ACT_TANH || net->acts[0] == B200_ACT_SIGMOID) ? 2 : 1;
    j2.rs_W = W0; j2.rs_bias = W0 + (size_t)K * N; j2.rs_R = K; j2.rs_tscale = m.tscale; j2.rs_tinv = m.tinv;
    j2.nblocks = ceil_div(N1, kSplitNeurons);
    if (j2.rs_mode == 1 && env().prep_pub) { // job 0 (the CTAs that hold W_0 anyway) publishes the scales, job 2 waits for them
      j0.pub_tscale = m.tscale; j0.pub_tinv = m.tinv; j0.pub_bias = j2.rs_bias;
      j2.rs_mode = 3;
      sync = m.prep_sync;
    }
  }
  if (env().diag & 32) { j1.nblocks = 0; j2.nblocks = 0; }  // (timing experiments: results are wrong)
  if (env().diag & 64) { j2.rs_mode = 0; j0.pub_tscale = nullptr; sync = nullptr; }
  if (env().diag & 128) chain.nl = 0;
  {
    ProfScope ps(net->ctx, "split16");
    B200_LAUNCH(prep_w16_kernel, j0.nblocks + j1.nblocks + j2.nblocks + (chain.nl > 0 ? chain.nctas : 0), 1024, 0, net->ctx->stream, j0,
                j1, j2, net->spec_st, net->spec_flag, chain, sync, net->ctx->h_scalars + kHostErrSlot);
  }
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
  const bool emit16 = net->m16.on; // A_1 leaves only as the scaled fp16 pair (layer 1 runs on the fp16 kernels)
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
  if (emit16) // the pair [2 N / 64 blocks][batch][64]: stored as plain {32 halves, 32 rows} boxes
    B200_TRY(make_map_3d(&tout, net->m16.a16, 64, (unsigned long long)batch, (unsigned long long)(2 * N / 64), 32, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  else
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
  p.tscale = net->m16.tscale; p.nb_out = N / 64;
  const int grid = std::min(net->ctx->num_sms, p.tiles);
  // CTA pairs (cta_group::2) when there are at least two sample tiles per pair of SMs to share a weight stream over
  const bool pair = x2 && bn == 128 && (env().pair & 1) && p.tiles >= 4;
  const int grid_pair = 2 * std::min(net->ctx->num_sms / 2, (p.tiles + 1) / 2);
  const bool deep_x = (env().ring & 1) != 0; // X ring (HBM) 6 stages deep, weight ring (L2) 3
  if (emit16) {
    if (pair) B200_TRY((launch_fwd16<128, true, EPI_EMIT16, 6, true>(tx, twh, twl, tout, tout, p, grid_pair, st)));
    else if (bn == 128 && deep_x) B200_TRY((launch_fwd16<128, true, EPI_EMIT16, 6, false, 3>(tx, twh, twl, tout, tout, p, grid, st)));
    else if (bn == 128) B200_TRY((launch_fwd16<128, true, EPI_EMIT16, 4>(tx, twh, twl, tout, tout, p, grid, st)));
    else B200_TRY((launch_fwd16<64, true, EPI_EMIT16, 4>(tx, twh, twl, tout, tout, p, grid, st)));
    net->m16.act0_stale = true;
  } else if (bn == 128) {
    if (pair) B200_TRY((launch_fwd16<128, true, EPI_F32, 6, true>(tx, twh, twl, tout, tout, p, grid_pair, st)));
    else if (x2 && deep_x) B200_TRY((launch_fwd16<128, true, EPI_F32, 6, false, 3>(tx, twh, twl, tout, tout, p, grid, st)));
    else if (x2) B200_TRY((launch_fwd16<128, true, EPI_F32, 4>(tx, twh, twl, tout, tout, p, grid, st)));
    else B200_TRY((launch_fwd16<128, false, EPI_F32, 4>(tx, twh, twl, tout, tout, p, grid, st)));
  } else {
    if (x2) B200_TRY((launch_fwd16<64, true, EPI_F32, 4>(tx, twh, twl, tout, tout, p, grid, st)));
    else B200_TRY((launch_fwd16<64, false, EPI_F32, 4>(tx, twh, twl, tout, tout, p, grid, st)));
  }
  *done = true;
  return B200_OK;
}

// mid16: A_2 = act(A_1 W_1 + b_1) with A_1 as its fp16 pair: the layer-0 kernel on that copy, K = 2 * width (hi blocks, then lo
// blocks, against weight rows written twice). Output fp32 rows (the last-layer kernels read them).
int mid16_forward_layer1(b200_net *net, const float *params, long batch) {
  b200_net::Mid16 &m = net->m16;
  const int K = net->dims[1], N = net->dims[2];
  cudaStream_t st = net->ctx->stream;
  const float *W = params + net->offs[1];
  CUtensorMap tx, twh, twl, tout;
  B200_TRY(make_map_3d(&tx, m.a16, 64, (unsigned long long)batch, (unsigned long long)(2 * K / 64), 64, kFM));
  const unsigned bn = N > 64 ? 128 : 64;
  B200_TRY(make_map_2d(&twh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, m.wfh, 2 * K, N, (unsigned long long)2 * K * 2, kFK, bn, CU_TENSOR_MAP_SWIZZLE_128B));
  B200_TRY(make_map_2d(&twl, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, m.wfl, 2 * K, N, (unsigned long long)2 * K * 2, kFK, bn, CU_TENSOR_MAP_SWIZZLE_128B));
  B200_TRY(make_map_2d(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, net->act[1], N, batch, (unsigned long long)N * 4, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
  F16Params p{};
  p.rows_valid = (int)batch; p.cols_valid = N; p.k_total = 2 * K;
  p.row0 = 0;
  p.k_blocks = 2 * K / kFK;
  p.tiles = ceil_div(batch, kFM);
  p.act = net->acts[1];
  p.bias = W + (size_t)K * N;
  p.colscale = m.colscale_f;
  p.out = net->act[1]; p.ld_out = N;
  p.spec_st = net->spec_st; p.spec = net->spec_flag;
  const int grid = std::min(net->ctx->num_sms, p.tiles);
  if (bn == 128) B200_TRY((launch_fwd16<128, true, EPI_F32, 4>(tx, twh, twl, tout, tout, p, grid, st)));
  else B200_TRY((launch_fwd16<64, true, EPI_F32, 4>(tx, twh, twl, tout, tout, p, grid, st)));
  return B200_OK;
}

// mid16: delta_0 = (delta_1 W_1^T) .* act_0'(A_1), both deltas as scaled fp16 pairs: A = delta_1's pair [rows][hi N1 | lo N1]
// (two "blocks" of a row), B = W_1^T rows written twice, act' from the hi blocks of A_1's pair, output = delta_0's pair
// [rows][hi N0 | lo N0] in net->delta16 with the chained scale (tail_layer.cu) for layer 0's dW kernel.
int mid16_dx_layer1(b200_net *net, long batch) {
  b200_net::Mid16 &m = net->m16;
  const int N0 = net->dims[1], N1 = net->dims[2];
  cudaStream_t st = net->ctx->stream;
  CUtensorMap tx, twh, twl, tout, taux;
  // dims ordered so that the strides ascend: {64 halves, N1 / 64 * 2 blocks (128 B apart), rows (2 N1 halves apart)}
  B200_TRY(make_map_3d_ex(&tx, m.d16, 64, (unsigned long long)(2 * N1 / 64), (unsigned long long)batch, 128, (unsigned long long)2 * N1 * 2,
                          64, 1, kFM, CU_TENSOR_MAP_SWIZZLE_128B));
  const unsigned bn = N0 > 64 ? 128 : 64;
  B200_TRY(make_map_2d(&twh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, m.wdh, 2 * N1, N0, (unsigned long long)2 * N1 * 2, kFK, bn, CU_TENSOR_MAP_SWIZZLE_128B));
  B200_TRY(make_map_2d(&twl, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, m.wdl, 2 * N1, N0, (unsigned long long)2 * N1 * 2, kFK, bn, CU_TENSOR_MAP_SWIZZLE_128B));
  B200_TRY(make_map_2d(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, net->delta16, 2 * N0, batch, (unsigned long long)2 * N0 * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  B200_TRY(make_map_3d(&taux, m.a16, 64, (unsigned long long)batch, (unsigned long long)(2 * N0 / 64), 64, kFM));
  F16Params p{};
  p.rows_valid = (int)batch; p.cols_valid = N0; p.k_total = 2 * N1;
  p.row0 = 0;
  p.k_blocks = 2 * N1 / kFK;
  p.tiles = ceil_div(batch, kFM);
  p.act = net->acts[0];
  p.colscale = m.colscale_d;
  p.bias = m.colscale_d; // (unused by this epilogue)
  p.spec_st = net->spec_st; p.spec = net->spec_flag;
  p.x_block_first = 1;
  p.scale_in_inv = m.scale1_inv;
  p.scale_out = net->scale16;
  const int grid = std::min(net->ctx->num_sms, p.tiles);
  if (bn == 128) B200_TRY((launch_fwd16<128, true, EPI_DX, 2>(tx, twh, twl, tout, taux, p, grid, st)));
  else B200_TRY((launch_fwd16<64, true, EPI_DX, 2>(tx, twh, twl, tout, taux, p, grid, st)));
  return B200_OK;
}


// ---- wide hidden layers on the fp16 pair kernels (b200_net::Wide16) ---------------------------------------------------------
// Replaces, for layers whose GEMMs are compute-bound (BASELINE configs[4]: 784-4096-4096-10), the three cublasSgemm calls of
// CudaDenseLayer (src/cuda/layer.cuh:51-54 forward, :81-84 dW, :89-92 dX) and their element-wise kernels. The generic kernel
// (gemm_tc.cu) runs them as 3xTF32 with N = 128, K = 8 MMAs at the 133-clk issue floor of cta_group::1; here both operands are
// fp16 pairs and a CTA pair issues M = 256, N = 256, K = 16 MMAs (130 clk): four products per K step (hi hi, hi lo, lo hi, lo lo)
// in two instructions.
namespace {

// per-CTA max |x| over a linear array (padding columns of a delta buffer are zero)
__global__ void __launch_bounds__(256) wide_amax_kernel(const float *__restrict__ x, unsigned long long n, float *__restrict__ part,
                                                      const SpecState *spec_st, int spec) {
  pdl_enter();
  if (spec_skip(spec_st, spec)) return;
  float m = 0.0f;
  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
  const unsigned long long n4 = vec ? n / 4 : 0;
  const float4 *x4 = reinterpret_cast<const float4 *>(x);
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (unsigned long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x4 + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  for (unsigned long long i = 4 * n4 + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(x + i)));
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    part[blockIdx.x] = m;
  }
}

// x[rows][ld] (cols valid) -> P[rows][hi Cp | lo Cp] and (PT != nullptr) PT[cols (+ ones)][hi Rp | lo Rp], zero padded; the scale
// S = 2^(14 - e) with max |x| in [2^(e-1), 2^e) from the per-CTA maxima; block 0 publishes {S, 1 / S}
// GEN: x is not read but generated, element (r, c) = act'(a[r][c]) * sum_j dl[r][j] Wl[c][j]: the dX of a skinny last layer
// (src/cuda/layer.cuh:89-103 with out <= 12) fused with the split of its result — delta_{L-1} then exists only as the pair. The
// per-CTA maxima are those of delta_L, and the scale comes from the bound max |delta_L| * max_c sum_j |Wl[c][j]| (|act'| <= 1).
constexpr int kGenMaxOut = 12;
struct WideGen {
  const float *dl; long ldl; int out;
  const float *Wl;  // [cols][out]
  const float *a;   // [rows][cols] activations whose act' multiplies
  int act;
};
template <bool GEN>
__global__ void __launch_bounds__(256, 3) wide_split_kernel(const float *__restrict__ x, long rows, int cols, long ld, const float *__restrict__ part,
                                                       int npart, __half *__restrict__ P, int Cp, __half *__restrict__ PT, long Rp, int ones,
                                                       float *__restrict__ scal, const SpecState *spec_st, int spec, const WideGen gen) {
  pdl_enter();
  if (spec_skip(spec_st, spec)) return;
  __shared__ float red[8];
  __shared__ __half th[64][72], tl[64][72]; // [column of the tile][row of the tile] (rows padded: 16-byte reads at 144-byte pitch)
  float m = 0.0f;
  for (int i = threadIdx.x; i < npart; i += 256) m = fmaxf(m, __ldcg(part + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  for (int i = 0; i < 8; ++i) m = fmaxf(m, red[i]);
  if constexpr (GEN) { // * max_c sum_j |Wl[c][j]|
    float wn = 0.0f;
    for (int c = threadIdx.x; c < cols; c += 256) {
      float sum = 0.0f;
      for (int j = 0; j < gen.out; ++j) sum += fabsf(__ldg(gen.Wl + (long)c * gen.out + j));
      wn = fmaxf(wn, sum);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wn = fmaxf(wn, __shfl_xor_sync(0xffffffffu, wn, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = wn;
    __syncthreads();
    for (int i = 0; i < 8; ++i) wn = fmaxf(wn, red[i]);
    m *= wn * 1.0001f; // (fp32 summation slack)
  }
  int e = 0;
  if (m > 0.0f && m < 3.0e38f) frexpf(m, &e);
  e = max(-100, min(100, e));
  const float S = ldexpf(1.0f, 14 - e);
  if (blockIdx.x == 0 && threadIdx.x == 0) { scal[0] = S; scal[1] = ldexpf(1.0f, e - 14); }
  const long ldp = 2L * Cp, ldt = 2L * Rp;
  const int tiles_c = Cp / 64;
  const long tiles_r = Rp / 64;
  // a CTA keeps one tile column (the grid is a multiple of tiles_c) and walks down the rows: GEN loads its four columns of the last
  // layer's weights once. A thread owns 4 consecutive rows x 4 consecutive columns of the 64 x 64 tile.
  const int G = gridDim.x / tiles_c;
  const int c0 = (int)(blockIdx.x % tiles_c) * 64;
  const int tr_ = threadIdx.x >> 4, tc4 = (threadIdx.x & 15) * 4;
  const bool vec = (cols % 4 == 0) && (GEN ? ((reinterpret_cast<uintptr_t>(gen.a) & 15u) == 0)
                                           : (ld % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0));
  // GEN: the last layer's weights of this CTA's 64 columns (once) and delta_L of the tile's 64 rows (per tile) in shared memory,
  // [.][out] padded to 13 floats; a thread's 4 x 4 values are then 16 accumulators fed by 8 shared-memory loads per output j
  // (output-major: a thread's four columns / four rows of one output j are ONE conflict-free 16-byte load)
  __shared__ __align__(16) float wls[GEN ? kGenMaxOut : 1][GEN ? 64 : 1], dls[GEN ? kGenMaxOut : 1][GEN ? 64 : 1];
  if constexpr (GEN) {
    for (int e = threadIdx.x; e < 64 * kGenMaxOut; e += 256) {
      const int c = e / kGenMaxOut, j = e - c * kGenMaxOut;
      wls[j][c] = (j < gen.out && c0 + c < cols) ? __ldg(gen.Wl + (long)(c0 + c) * gen.out + j) : 0.0f;
    }
  }
  // this thread's 4 x 4 values of tile row rb (zero past the edges); the loads of the NEXT tile are issued before the current one is
  // processed, so a thread always has a tile's worth of loads in flight
  auto load_tile = [&](long rb, float4 (&dst)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long r = rb * 64 + 4 * tr_ + i;
      const float *src = GEN ? gen.a + r * (long)cols + c0 + tc4 : x + r * ld + c0 + tc4;
      if (r < rows && vec && c0 + tc4 + 3 < cols) dst[i] = __ldg(reinterpret_cast<const float4 *>(src));
      else {
        dst[i].x = (r < rows && c0 + tc4 + 0 < cols) ? __ldg(src + 0) : 0.0f;
        dst[i].y = (r < rows && c0 + tc4 + 1 < cols) ? __ldg(src + 1) : 0.0f;
        dst[i].z = (r < rows && c0 + tc4 + 2 < cols) ? __ldg(src + 2) : 0.0f;
        dst[i].w = (r < rows && c0 + tc4 + 3 < cols) ? __ldg(src + 3) : 0.0f;
      }
    }
  };
  float4 nxt[4];
  if ((long)(blockIdx.x / tiles_c) < tiles_r) load_tile(blockIdx.x / tiles_c, nxt);
  for (long rb = blockIdx.x / tiles_c; rb < tiles_r; rb += G) {
    const long r0 = rb * 64;
    float v[4][4];
    {
      float4 raw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) raw[i] = nxt[i];
      if (rb + G < tiles_r) load_tile(rb + G, nxt);
      if constexpr (GEN) {
        for (int e = threadIdx.x; e < 64 * kGenMaxOut; e += 256) { // delta_L of the tile's rows (zero past the batch / past `out`)
          const int rr = e / kGenMaxOut, j = e - rr * kGenMaxOut;
          dls[j][rr] = (j < gen.out && r0 + rr < rows) ? __ldg(gen.dl + (r0 + rr) * gen.ldl + j) : 0.0f;
        }
        __syncthreads(); // (also orders the first tile after the staging of wls)
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[i][k] = 0.0f;
#pragma unroll
        for (int j = 0; j < kGenMaxOut; ++j) {
          const float4 d4 = *reinterpret_cast<const float4 *>(&dls[j][4 * tr_]), w4 = *reinterpret_cast<const float4 *>(&wls[j][tc4]);
          const float d[4] = {d4.x, d4.y, d4.z, d4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(d[i], wv[k], acc[i][k]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float av[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w}; // (zero past the edges: act' of it times a zero sum)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool ok = r0 + 4 * tr_ + i < rows && c0 + tc4 + k < cols;
            v[i][k] = ok ? acc[i][k] * act_deriv_from_output(gen.act, av[k]) * S : 0.0f;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i][0] = raw[i].x * S; v[i][1] = raw[i].y * S; v[i][2] = raw[i].z * S; v[i][3] = raw[i].w * S; }
      }
    }
    __align__(8) __half h[4][4], l[4][4]; // [row i][col k]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        h[i][k] = __float2half_rn(v[i][k]);
        l[i][k] = __float2half_rn(v[i][k] - __half2float(h[i][k]));
      }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long r = r0 + 4 * tr_ + i;
      if (r < rows) {
        *reinterpret_cast<uint2 *>(P + r * ldp + c0 + tc4) = *reinterpret_cast<const uint2 *>(h[i]);
        *reinterpret_cast<uint2 *>(P + r * ldp + Cp + c0 + tc4) = *reinterpret_cast<const uint2 *>(l[i]);
      }
    }
    if (PT) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { // four consecutive rows of column tc4 + k: one 8-byte store each for hi and lo
        __align__(8) __half hc[4] = {h[0][k], h[1][k], h[2][k], h[3][k]}, lc[4] = {l[0][k], l[1][k], l[2][k], l[3][k]};
        *reinterpret_cast<uint2 *>(&th[tc4 + k][4 * tr_]) = *reinterpret_cast<const uint2 *>(hc);
        *reinterpret_cast<uint2 *>(&tl[tc4 + k][4 * tr_]) = *reinterpret_cast<const uint2 *>(lc);
      }
      __syncthreads();
      const int col = threadIdx.x >> 2, seg = (threadIdx.x & 3) * 16;
      if (c0 + col < cols) {
        __half *q = PT + (long)(c0 + col) * ldt + r0 + seg;
        *reinterpret_cast<uint4 *>(q) = *reinterpret_cast<const uint4 *>(&th[col][seg]);
        *reinterpret_cast<uint4 *>(q + 8) = *reinterpret_cast<const uint4 *>(&th[col][seg + 8]);
        *reinterpret_cast<uint4 *>(q + Rp) = *reinterpret_cast<const uint4 *>(&tl[col][seg]);
        *reinterpret_cast<uint4 *>(q + Rp + 8) = *reinterpret_cast<const uint4 *>(&tl[col][seg + 8]);
      }
      __syncthreads();
    } else if constexpr (GEN) {
      __syncthreads(); // dls is refilled by the next tile
    }
  }
  if (PT && ones) { // row `cols` of the transpose: unscaled ones over the valid rows (the bias-gradient row of the dW GEMM)
    __half *q = PT + (long)cols * ldt;
    for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < Rp; r += (long)gridDim.x * blockDim.x) {
      q[r] = __float2half_rn(r < rows ? 1.0f : 0.0f);
      q[Rp + r] = __float2half_rn(0.0f);
    }
  }
}

inline long up64(long v) { return (v + 63) / 64 * 64; }

int wide_buf(b200_net *net, b200_net::Wide16::Buf &b, size_t halves) {
  if (b.halves >= halves) return B200_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr; b.halves = 0;
  B200_CUDA(cudaMalloc(&b.p, halves * sizeof(__half)));
  b.halves = halves;
  ++net->config_gen;
  return B200_OK;
}

int wide_ensure(b200_net *net) {
  b200_net::Wide16 &w = net->w16x;
  const int L = net->nlayers();
  if ((int)w.a.size() != L) {
    w.a.resize(L); w.aT.resize(L); w.wf.resize(L); w.wd.resize(L);
    w.a_ready.assign(L, 0); w.w_ready.assign(L, 0); w.d_ready.assign(L, 0);
  }
  if (!w.scal || w.scal_layers < L) {
    if (w.scal) cudaFree(w.scal);
    if (w.amax_part) cudaFree(w.amax_part);
    w.scal = w.amax_part = nullptr;
    w.scal_layers = L;
    B200_CUDA(cudaMalloc(&w.scal, sizeof(float) * 8 * L));
    w.amax_n = 4 * net->ctx->num_sms;
    B200_CUDA(cudaMalloc(&w.amax_part, sizeof(float) * w.amax_n));
    ++net->config_gen;
  }
  return B200_OK;
}

// grid of the split kernel: a multiple of the tile columns (a CTA keeps its tile column), ~8 CTAs per SM
int wide_split_grid(const b200_net *net, long Rp, long Cp) {
  const long tiles_c = Cp / 64, tiles_r = Rp / 64;
  const long G = std::max<long>(1, std::min<long>(tiles_r, (8L * net->ctx->num_sms) / tiles_c));
  return (int)(tiles_c * G);
}

// split x[rows][ld] into P (and PT); scal -> {S, 1 / S}
int wide_split(b200_net *net, const float *x, long rows, int cols, long ld, void *P, int Cp, void *PT, long Rp, int ones, float *scal) {
  b200_net::Wide16 &w = net->w16x;
  cudaStream_t st = net->ctx->stream;
  const unsigned long long n = (unsigned long long)rows * ld - (unsigned long long)(ld - cols); // (the last row ends at its last valid column)
  const int ga = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((unsigned long long)w.amax_n, (n + 1023) / 1024));
  B200_LAUNCH(wide_amax_kernel, ga, 256, 0, st, x, n, w.amax_part, net->spec_st, net->spec_flag);
  const int gs = wide_split_grid(net, Rp, Cp);
  B200_LAUNCH(wide_split_kernel<false>, gs, 256, 0, st, x, rows, cols, ld, (const float *)w.amax_part, ga, (__half *)P, Cp, (__half *)PT, Rp, ones, scal,
              net->spec_st, net->spec_flag, WideGen{});
  return B200_OK;
}

int wide_weights(b200_net *net, int l, const float *params) {
  b200_net::Wide16 &w = net->w16x;
  if (w.w_ready[l]) return B200_OK;
  const int K = net->dims[l], N = net->dims[l + 1];
  const long Kp = up64(K), Np = up64(N);
  B200_TRY(wide_buf(net, w.wd[l], (size_t)K * 2 * Np));
  B200_TRY(wide_buf(net, w.wf[l], (size_t)N * 2 * Kp));
  // W_l [K][N]: P = [K][hi Np | lo Np] (dX: rows = the columns of delta_{l-1}), PT = [N][hi Kp | lo Kp] (forward: rows = outputs)
  B200_TRY(wide_split(net, params + net->offs[l], K, N, N, w.wd[l].p, (int)Np, w.wf[l].p, Kp, 0, w.scal + 8 * l + 2));
  w.w_ready[l] = 1;
  return B200_OK;
}

int wide_input(b200_net *net, int l, const float *in, long batch) {
  b200_net::Wide16 &w = net->w16x;
  if (w.a_ready[l]) return B200_OK;
  const int K = net->dims[l];
  const long Kp = up64(K), Bp = up64(net->cap);
  // layer 0 on an input the caller holds constant (Wide16::x_src): the split of an earlier evaluation is still there
  const bool held = l == 0 && w.x_src == in && w.x_rows == batch;
  if (l == 0 && (!held || w.a[0].halves < (size_t)net->cap * 2 * Kp || w.aT[0].halves < (size_t)(K + 1) * 2 * Bp)) w.x_done = false;
  B200_TRY(wide_buf(net, w.a[l], (size_t)net->cap * 2 * Kp));
  B200_TRY(wide_buf(net, w.aT[l], (size_t)(K + 1) * 2 * Bp));
  if (!(held && w.x_done)) B200_TRY(wide_split(net, in, batch, K, K, w.a[l].p, (int)Kp, w.aT[l].p, up64(batch), 1, w.scal + 8 * l));
  if (held) w.x_done = true;
  w.a_ready[l] = 1;
  return B200_OK;
}

int wide_delta(b200_net *net, int l, long batch) {
  b200_net::Wide16 &w = net->w16x;
  if (w.d_ready[l]) return B200_OK;
  const int N = net->dims[l + 1];
  const long Np = up64(N), Bp = up64(net->cap);
  int maxN = 0;
  for (int j = 0; j < net->nlayers(); ++j) maxN = std::max(maxN, net->dims[j + 1]);
  B200_TRY(wide_buf(net, w.d, (size_t)net->cap * 2 * up64(maxN)));
  B200_TRY(wide_buf(net, w.dT, (size_t)maxN * 2 * Bp));
  B200_TRY(wide_split(net, net->delta[l], batch, N, net->ldd[l], w.d.p, (int)Np, w.dT.p, up64(batch), 0, w.scal + 8 * l + 4));
  std::fill(w.d_ready.begin(), w.d_ready.end(), 0); // (one buffer for every layer: it now holds layer l's)
  w.d_ready[l] = 1;
  return B200_OK;
}

// C[rows][ncols] = A B^T / (S_A S_B): A = pair rows [rows][hi Kc | lo Kc] (Kc the padded contraction length), B = pair rows
// [ncols][hi Kc | lo Kc]
struct WideGemm {
  const void *A; long rows; const void *B; int ncols; long Kc;
  float *out; long ld_out;
  int slices = 1; unsigned long long slice_stride = 0; // split-K: slice s of the contraction is stored at out + s * slice_stride
  const float *sa_inv, *sb_inv, *bias, *aux32; long ld_aux; int act, ones_row;
};
int wide_gemm(b200_net *net, const WideGemm &g, int *slices_used = nullptr) {
  CUtensorMap tx, twh, twl, tout;
  // A: dims ordered so that the strides ascend {64 halves, 2 Kc / 64 blocks (128 B apart), rows}
  B200_TRY(make_map_3d_ex(&tx, g.A, 64, (unsigned long long)(2 * g.Kc / 64), (unsigned long long)g.rows, 128, (unsigned long long)2 * g.Kc * 2, 64, 1,
                          kFM, CU_TENSOR_MAP_SWIZZLE_128B));
  B200_TRY(make_map_2d(&twh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, g.B, (unsigned long long)g.Kc, (unsigned long long)g.ncols,
                       (unsigned long long)2 * g.Kc * 2, kFK, 128, CU_TENSOR_MAP_SWIZZLE_128B));
  B200_TRY(make_map_2d(&twl, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (const __half *)g.B + g.Kc, (unsigned long long)g.Kc, (unsigned long long)g.ncols,
                       (unsigned long long)2 * g.Kc * 2, kFK, 128, CU_TENSOR_MAP_SWIZZLE_128B));
  const int k_blocks = (int)(g.Kc / kFK);
  const int kb_per_slice = ceil_div(k_blocks, std::max(1, g.slices)), slices = ceil_div(k_blocks, kb_per_slice); // (no empty slice)
  B200_TRY(make_map_3d_ex(&tout, g.out, (unsigned long long)g.ncols, (unsigned long long)g.rows, (unsigned long long)slices,
                          (unsigned long long)g.ld_out * 4, (slices > 1 ? g.slice_stride : (unsigned long long)g.rows * g.ld_out) * 4, 32, 32, 1,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT32));
  F16Params p{};
  p.k_slices = slices; p.kb_per_slice = kb_per_slice;
  p.rows_valid = (int)g.rows; p.cols_valid = g.ncols; p.k_total = (int)g.Kc;
  p.row0 = 0;
  p.k_blocks = (int)(g.Kc / kFK);
  p.tiles = ceil_div(g.rows, kFM);
  p.n_tiles = g.ncols / 128;
  p.k_chunk = env().wide_chunk > 0 ? env().wide_chunk : 4;
  p.act = g.act;
  p.bias = g.bias;
  p.out = g.out; p.ld_out = g.ld_out;
  p.spec_st = net->spec_st; p.spec = net->spec_flag;
  p.x_block_first = 1;
  p.sa_inv = g.sa_inv; p.sb_inv = g.sb_inv;
  p.aux32 = g.aux32; p.ld_aux = g.ld_aux;
  p.ones_row = g.ones_row;
  const long units = (long)((p.tiles + 1) / 2) * p.n_tiles * slices;
  const int grid = 2 * (int)std::min<long>(net->ctx->num_sms / 2, units);
  if (slices_used) *slices_used = slices;
  return launch_fwd16<128, true, EPI_WIDE, 4, true, 4, true>(tx, twh, twl, tout, tout, p, grid, net->ctx->stream);
}

} // namespace

bool wide16_applicable(const b200_net *net, int l, int role, long batch) {
  if (!env().wide16 || net->prec != B200_PREC_TF32X3 || net->m16.on) return false;
  const int K = net->dims[l], N = net->dims[l + 1];
  if (K < 256 || N < 128 || N % 128 != 0) return false;
  // below ~1e9 multiply-adds a GEMM is launch- and latency-bound (S-LBFGS mini-batches on the 256-wide pair network): the two
  // extra launches of an operand split cost more than the tensor time saved (B200_WIDE16_MIN, in multiply-adds)
  if ((double)batch * K * N < (double)env().wide16_min) return false;
  if (l == 0 && fwd16_shape_ok(net) && net->xq.valid) return false; // an 8-bit-pixel input: the exact-fp16 kernels of §3.1 take layer 0
  if (role == 1 && (l == 0 || K % 128 != 0 || net->ldd[l - 1] % 4 != 0)) return false; // (dX: the output tile is 128 columns of delta_{l-1})
  return true;
}

void wide16_begin(b200_net *net) {
  b200_net::Wide16 &w = net->w16x;
  std::fill(w.a_ready.begin(), w.a_ready.end(), 0);
  std::fill(w.w_ready.begin(), w.w_ready.end(), 0);
  std::fill(w.d_ready.begin(), w.d_ready.end(), 0);
}

// slices the partial buffer of layer l has room for (net_ensure sizes it for the FFMA / generic split-K plans)
int wide16_dw_slices_cap(const b200_net *net, int l) {
  const int L = net->nlayers();
  const size_t per = (size_t)(net->dims[l] + 1) * net->dims[l + 1];
  const size_t end = (l + 1 < L) ? net->part_off[l + 1] : net->partials_cap;
  return (int)std::min<size_t>(64, (end - net->part_off[l]) / per);
}

// A_l = act(A_{l-1} W_l + b_l)
int wide16_forward_layer(b200_net *net, int l, const float *params, const float *in, long batch) {
  B200_TRY(wide_ensure(net));
  b200_net::Wide16 &w = net->w16x;
  const int K = net->dims[l], N = net->dims[l + 1];
  B200_TRY(wide_weights(net, l, params));
  B200_TRY(wide_input(net, l, in, batch));
  WideGemm g{};
  g.A = w.a[l].p; g.rows = batch; g.B = w.wf[l].p; g.ncols = N; g.Kc = up64(K);
  g.out = net->act[l]; g.ld_out = N;
  g.sa_inv = w.scal + 8 * l + 1; g.sb_inv = w.scal + 8 * l + 3;
  g.bias = params + net->offs[l] + (size_t)K * N; g.act = net->acts[l]; g.ones_row = -1;
  return wide_gemm(net, g);
}

// delta_{l-1} = (delta_l W_l^T) .* act'_{l-1}(A_{l-1})
int wide16_dx_layer(b200_net *net, int l, const float *params, long batch) {
  B200_TRY(wide_ensure(net));
  b200_net::Wide16 &w = net->w16x;
  const int K = net->dims[l], N = net->dims[l + 1];
  B200_TRY(wide_weights(net, l, params));
  B200_TRY(wide_delta(net, l, batch));
  WideGemm g{};
  g.A = w.d.p; g.rows = batch; g.B = w.wd[l].p; g.ncols = K; g.Kc = up64(N);
  g.out = net->delta[l - 1]; g.ld_out = net->ldd[l - 1];
  g.sa_inv = w.scal + 8 * l + 5; g.sb_inv = w.scal + 8 * l + 3;
  g.aux32 = net->act[l - 1]; g.ld_aux = K; g.act = net->acts[l - 1]; g.ones_row = -1;
  return wide_gemm(net, g);
}

// The layer below a skinny last layer (out <= 12): its delta pair straight from delta_L, W_L and act'(A_{L-1}) — replaces the last
// layer's dX GEMM + activation_deriv (src/cuda/layer.cuh:89-103), whose fp32 result (2 GB at configs[4]'s per-GPU share) would be
// written only to be read back twice (maximum, split).
bool wide16_last_dx_applicable(const b200_net *net, long batch) {
  const int L = net->nlayers();
  if (L < 2 || net->dims[L] > kGenMaxOut || !env().wide16 || (env().wide16 & 2)) return false; // (B200_WIDE16=3: separate dX + split)
  const int l = L - 2;
  return wide16_applicable(net, l, 2, batch) && (l == 0 || wide16_applicable(net, l, 1, batch));
}

int wide16_last_dx(b200_net *net, const float *params, long batch) {
  B200_TRY(wide_ensure(net));
  b200_net::Wide16 &w = net->w16x;
  const int L = net->nlayers(), l = L - 2;
  const int N = net->dims[l + 1], out = net->dims[L];
  const long Np = up64(N), Bp = up64(net->cap);
  int maxN = 0;
  for (int j = 0; j < L; ++j) maxN = std::max(maxN, net->dims[j + 1]);
  B200_TRY(wide_buf(net, w.d, (size_t)net->cap * 2 * up64(maxN)));
  B200_TRY(wide_buf(net, w.dT, (size_t)maxN * 2 * Bp));
  cudaStream_t st = net->ctx->stream;
  const float *dl = net->delta[L - 1];
  const long ldl = net->ldd[L - 1];
  const unsigned long long n = (unsigned long long)batch * ldl - (unsigned long long)(ldl - out);
  const int ga = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((unsigned long long)w.amax_n, (n + 1023) / 1024));
  B200_LAUNCH(wide_amax_kernel, ga, 256, 0, st, dl, n, w.amax_part, net->spec_st, net->spec_flag);
  const long Rp = up64(batch);
  const int gs = wide_split_grid(net, Rp, Np);
  WideGen gen{dl, ldl, out, params + net->offs[L - 1], net->act[l], net->acts[l]};
  B200_LAUNCH(wide_split_kernel<true>, gs, 256, 0, st, (const float *)nullptr, batch, N, (long)N, (const float *)w.amax_part, ga, (__half *)w.d.p, (int)Np,
              (__half *)w.dT.p, Rp, 0, w.scal + 8 * l + 4, net->spec_st, net->spec_flag, gen);
  std::fill(w.d_ready.begin(), w.d_ready.end(), 0);
  w.d_ready[l] = 1;
  return B200_OK;
}

// [dW_l; db_l] = [A_{l-1} | 1]^T delta_l, one slice (the whole batch) straight into the layer's partial
int wide16_dw_layer(b200_net *net, int l, const float *in, long batch) {
  B200_TRY(wide_ensure(net));
  b200_net::Wide16 &w = net->w16x;
  const int K = net->dims[l], N = net->dims[l + 1];
  B200_TRY(wide_input(net, l, in, batch)); // (made by the forward pass of this evaluation)
  B200_TRY(wide_delta(net, l, batch));
  WideGemm g{};
  g.A = w.aT[l].p; g.rows = K + 1; g.B = w.dT.p; g.ncols = N; g.Kc = up64(batch);
  g.out = net->partials + net->part_off[l]; g.ld_out = N;
  g.sa_inv = w.scal + 8 * l + 1; g.sb_inv = w.scal + 8 * l + 5;
  g.act = B200_ACT_LINEAR; g.ones_row = K;
  // few output tiles (a 128-wide layer 0: four units): slices of the samples fill the machine, one partial per slice, combined by
  // finalize_grad_kernel like every other split-K; never more slices than the layer's partials were sized for (net_ensure)
  const long out_units = (long)((ceil_div(K + 1, kFM) + 1) / 2) * (N / 128);
  const int want = (int)std::min<long>(std::max<long>(1, (2L * (net->ctx->num_sms / 2)) / out_units), std::max<long>(1, g.Kc / kFK / 8));
  g.slices = std::max(1, std::min(want, wide16_dw_slices_cap(net, l)));
  g.slice_stride = (unsigned long long)(K + 1) * N;
  int used = 1;
  B200_TRY(wide_gemm(net, g, &used));
  net->splits_used[l] = used;
  return B200_OK;
}

void wide16_release(b200_net *net) {
  b200_net::Wide16 &w = net->w16x;
  for (auto *v : {&w.a, &w.aT, &w.wf, &w.wd})
    for (auto &b : *v) if (b.p) cudaFree(b.p);
  if (w.d.p) cudaFree(w.d.p);
  if (w.dT.p) cudaFree(w.dT.p);
  if (w.scal) cudaFree(w.scal);
  if (w.amax_part) cudaFree(w.amax_part);
  w = b200_net::Wide16{};
}

void fwd16_release(b200_net *net) {
  if (net->w16h) cudaFree(net->w16h);
  if (net->w16l) cudaFree(net->w16l);
  if (net->colscale) cudaFree(net->colscale);
  net->w16h = net->w16l = nullptr;
  net->colscale = nullptr;
  mid16_release(net);
  wide16_release(net);
}

} // namespace b200

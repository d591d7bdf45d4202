// [dW_0 ; db_0] split-K partials of layer 0 on the fp16 tensor cores, for an input that is exactly u/255 (8-bit pixels).
//
// Replaces CudaDenseLayer::backward's dW SGEMM (K = batch) and the serial sum_rows_kernel (src/cuda/layer.cuh:81-86,
// src/cuda/kernels.cuh:144-153) for the first layer:
//     D[f][o] = sum_s X[s][f] * delta[s][o]          f = input feature (M, tiles of 128), o = output neuron, s = sample (K)
// Operands: A = X^T from the fp16 copy of X (value u, exact), loaded by TMA as MN-major SWIZZLE_128B atoms; B = [delta_hi |
// delta_lo], the scaled fp16 split of delta written by tail_bwd_kernel (22 mantissa bits), also MN-major by TMA. One kind::f16
// MMA of N = 2 * out per 16 samples yields D = [hi | lo] in adjacent TMEM columns; the epilogue adds them, undoes the scales
// (1/255 of X, 1/S of delta) and leaves through TMA tile stores. The bias gradient is the row f = in of the same product: the fp16
// copy of X carries a column of ones there (sum_s 1 * delta[s][o]), exactly the [W | b] layout of the flat gradient.
// A CTA owns TWO feature tiles (512 TMEM columns) so that every delta tile fetched from L2 feeds two MMAs, and a slice of the
// samples (split-K, combined deterministically by finalize_grad_kernel).
// tcgen05.mma issues at one instruction per ~140-160 clk regardless of N (tools/probe/mma_probe.cu), so the N = 256, K = 16
// shape is what makes this 4x cheaper in tensor-core instructions than the TF32 N = 128, K = 8 kernel it replaces.
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace b200 {

namespace {

using namespace tcx;

constexpr int kDM = 128;                  // input features per M tile = UMMA M
constexpr int kDMT = 2;                   // M tiles per CTA: every delta tile read from L2 feeds two MMAs (the L2 -> SM path,
                                          // ~43 B/clk per SM, is what bounds a one-tile CTA), accumulators fill the 512 TMEM columns
constexpr int kDK = 32;                   // samples per stage
constexpr int kDConvTile = kDK * kDM * 2; // fp16: two feature atoms of [32 K-rows][128 B]
constexpr int kDConvBytes = kDMT * kDConvTile;
constexpr int kDAtom = kDK * 128;         // one MN atom column of a stage: [32 K-rows][64 elements]
constexpr int kDNS = 6;                   // default depth: A and B rings share one stage index (one commit frees both)
constexpr int kDThreads = 256;            // warp 0 X TMA, 1 MMA issue, 2 TMEM alloc, 3 delta TMA, 4-7 epilogue
constexpr int kMinKbPerSplit = 12;        // split-K plan: K blocks (of 32 samples) a slice holds at least

struct Dw16Params {
  int in_dim, out_dim;    // layer 0: in (784), out (<= 128)
  int k_blocks, kb_per_split;
  // grid: full_groups * splits CTAs that own two feature tiles (split-major, so the CTAs of one slice of the samples run side
  // by side and share its delta tiles in L2), then tail_splits CTAs for a last group that holds ONE tile: it has half the MMAs
  // per K block, so it gets fewer, longer slices and the CTAs finish together (784 + 1 features = 3 full groups + 17 rows)
  int full_groups, splits, tail_splits, kb_per_tail;
  float *partial;         // [split][(in+1)*out]
  unsigned long long partial_stride;
  const float *scale_inv; // device scalar 1 / S of the fp16 delta
  int row0;               // first sample of this evaluation inside the fp16 copy of the input
  const SpecState *spec_st; // speculative launch on a wrong guess: return at once (common.cuh)
  int spec;
  long long *dbg;
  int diag;               // B200_DIAG (timing experiments only): bit0 no X loads, bit1 no delta loads, bit2 no MMAs, bit3 no stores
  const float *rowscale;  // FOLD: [128] 1 / t_f of the activation pair
};

// NX / ND: depth of the X ring (HBM) and of the delta ring (L2-resident after the first feature group has read it)
template <int NB, int NX = kDNS, int ND = kDNS> struct DPlan { // NB = 2 * out rounded up to 128 / 256: width of [delta_hi | delta_lo]
  static constexpr int kBStage = NB * kDK * 2;
  static constexpr int kOffConv = 0;
  static constexpr int kOffB = NX * kDConvBytes;
  static constexpr int kOffBar = kOffB + ND * kBStage;
  static constexpr int kTotal = kOffBar + 512 + 1024;
  static_assert(kTotal <= 227 * 1024, "shared memory plan exceeds the SM");
};

__device__ __forceinline__ void umma_f16_d(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// fp32 accumulate, fp16 x fp16, both operands MN-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_f16_mn(int n) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kDM >> 4) << 24);
}
// MN-major SWIZZLE_128B, 16-bit: atoms of [8 K-rows][64 MN elements] (1024 B); next 8-row K group at SBO = 1024 B, next MN atom
// (64 elements further along M / N) at LBO = one [32 K-rows][128 B] box = 4096 B
__device__ __forceinline__ uint64_t desc_mn16(uint32_t saddr) { return make_desc(saddr, kDAtom, 1024, 2); }

// FOLD (hidden layer 1, b200_net::Mid16): the A operand is an activation PAIR — M tile 0 holds the hi halves of the 128
// features, tile 1 their lo halves — so the two accumulators of a CTA are two terms of the same product and the epilogue adds
// them (with the hi | lo column halves of the delta pair: four terms), scales row f by 1 / (t_f S) and writes ONE 128-row tile.
template <int NB, bool FOLD, int NX, int ND>
__global__ void __launch_bounds__(kDThreads, 1)
dw16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmOut,
            const Dw16Params p) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  if (spec_skip(p.spec_st, p.spec)) return; // before any barrier / TMEM allocation
  using Plan = DPlan<NB, NX, ND>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *bp = smem_raw + (base - smem_u32(smem_raw));
  auto conv_a = [&](int s) { return base + Plan::kOffConv + s * kDConvBytes; };
  auto b_a = [&](int s) { return base + Plan::kOffB + s * Plan::kBStage; };
  const uint32_t bars = base + Plan::kOffBar;
  auto conv_full = [&](int s) { return bars + 8 * (s); };
  auto b_full = [&](int s) { return bars + 8 * (NX + s); };
  auto st_empty = [&](int s) { return bars + 8 * (NX + ND + s); }; // MMAs of the stage done: the A tile (NX == ND: and the B tile) reusable
  auto b_empty = [&](int s) { return NX == ND ? st_empty(s) : bars + 8 * (2 * NX + ND + s); };
  const uint32_t acc_full = bars + 8 * (2 * NX + 2 * ND);
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(bp + Plan::kOffBar + 8 * (2 * NX + 2 * ND + 1));
  static_assert(2 * NX + 2 * ND + 2 <= 64, "barrier table");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_start = p.dbg ? clock64() : 0;
  const bool tail_cta = (int)blockIdx.x >= p.full_groups * p.splits;
  const int split = tail_cta ? (int)blockIdx.x - p.full_groups * p.splits : (int)blockIdx.x / p.full_groups;
  const int group = tail_cta ? p.full_groups : (int)blockIdx.x - split * p.full_groups;
  const int per = tail_cta ? p.kb_per_tail : p.kb_per_split;
  const int m0 = group * (kDMT * kDM);
  const int nmt = min(kDMT, (p.in_dim + 1 - m0 + kDM - 1) / kDM); // M tiles of this CTA that hold features (the last group may hold one)
  const int kb_begin = split * per, kb_end = min(p.k_blocks, kb_begin + per);
  const int nkb = max(0, kb_end - kb_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < NX; ++s) { mbar_init(conv_full(s), 1); mbar_init(st_empty(s), 1); }
    for (int s = 0; s < ND; ++s) { mbar_init(b_full(s), 1); if (NX != ND) mbar_init(b_empty(s), 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)),
                 "r"((uint32_t)(kDMT * NB))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    { // ===== X producer: fp16 rows -> MN-major SWIZZLE_128B atoms, two boxes {64 features, 32 samples} per M tile (whole warp in
      // the loop, one elected lane issues) ==
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(st_empty(s), ph ^ 1);
        if (elect_one()) {
          if (p.diag & 1) mbar_arrive(conv_full(s));
          else {
            mbar_expect_tx(conv_full(s), nmt * kDConvTile);
            for (int t = 0; t < nmt; ++t)
#pragma unroll
              for (int j = 0; j < 2; ++j)
                tma_load_3d(conv_a(s) + t * kDConvTile + j * kDAtom, &tmX, conv_full(s), 0, p.row0 + kb * kDK, (m0 + t * kDM) / 64 + j);
          }
        }
        __syncwarp();
        if (++s == NX) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    { // ===== delta producer: NB / 64 boxes {64 halves, 64 samples}, MN-major SWIZZLE_128B ==========
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(b_empty(s), ph ^ 1);
        if (elect_one()) {
          if (p.diag & 2) mbar_arrive(b_full(s));
          else {
            mbar_expect_tx(b_full(s), Plan::kBStage);
#pragma unroll
            for (int j = 0; j < NB / 64; ++j) tma_load_2d(b_a(s) + j * kDAtom, &tmD, b_full(s), 64 * j, kb * kDK);
          }
        }
        __syncwarp();
        if (++s == ND) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    { // ===== MMA issuer: the whole warp runs the loop (uniform control flow), one elected lane issues the tcgen05 instructions ====
      const uint32_t idesc = make_idesc_f16_mn(NB);
      const uint64_t dA0 = desc_mn16(conv_a(0)), dB0 = desc_mn16(b_a(0));
      int s = 0, sb = 0;
      uint32_t ph = 0, phb = 0;
      long long waited = 0, waited_b = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const long long t0 = p.dbg ? clock64() : 0;
        mbar_wait(conv_full(s), ph);
        const long long t1 = p.dbg ? clock64() : 0;
        mbar_wait(b_full(sb), phb);
        if (p.dbg) { waited += t1 - t0; waited_b += clock64() - t1; }
        tc_fence_after();
        const uint64_t da = dA0 + (uint64_t)(s * (kDConvBytes >> 4)), db = dB0 + (uint64_t)(sb * (Plan::kBStage >> 4));
        const int nt = (p.diag & 4) ? 0 : nmt;
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < kDMT; ++t) {
            if (t < nt) {
#pragma unroll
              for (int ks = 0; ks < kDK / 16; ++ks) // 16 samples = two 8-row K groups = 2048 B further into every atom
                umma_f16_d(tmem_base + t * NB, da + t * (kDConvTile >> 4) + 128 * ks, db + 128 * ks, idesc,
                           (kb > kb_begin || ks > 0) ? 1u : 0u);
            }
          }
          umma_commit(st_empty(s));
          if constexpr (NX != ND) umma_commit(b_empty(sb));
        }
        __syncwarp();
        if (++s == NX) { s = 0; ph ^= 1; }
        if (++sb == ND) { sb = 0; phb ^= 1; }
      }
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
      if (p.dbg && lane == 0) { p.dbg[4 * blockIdx.x + 1] = waited; p.dbg[4 * blockIdx.x + 2] = waited_b; }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===== epilogue (same warps: warp w owns TMEM lanes 32 * (w % 4) ..): partial[split][f * out + o] = (hi + lo) * scale ====
    const int OUT = p.out_dim;
    {
    if (nkb > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    const float sinv = __ldg(p.scale_inv);
    // Each warp stages [32 feature rows][32 outputs] blocks in the (now idle) pipeline memory and one lane issues a 3-D TMA
    // tile store {output, feature, split}: full lines, and rows past feature `in` are clipped by the tensor map.
    const uint32_t stage_a = base + (warp - 4) * 4096;
    uint8_t *stage_p = bp + (warp - 4) * 4096;
    for (int t2 = 0; t2 < (FOLD ? 1 : nmt); ++t2) {
      const int row = (warp & 3) * 32 + lane, f = m0 + t2 * kDM + row;
      const float scale = FOLD ? sinv * __ldg(p.rowscale + row) : ((f < p.in_dim) ? sinv * (1.0f / 255.0f) : sinv);
      const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t2 * NB);
      for (int c0 = 0; c0 < OUT; c0 += 32) {
        uint32_t v[32], w[32];
        if (nkb > 0) {
          tmem_ld32(lane_addr + c0, v);
          tmem_ld32(lane_addr + NB / 2 + c0, w);
          if (FOLD) { // + the lo-feature tile's two column halves
            uint32_t v2[32], w2[32];
            tmem_ld32(lane_addr + NB + c0, v2);
            tmem_ld32(lane_addr + NB + NB / 2 + c0, w2);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
              w[j] = __float_as_uint(__uint_as_float(w[j]) + __uint_as_float(w2[j]));
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = 0u; w[j] = 0u; }
        }
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
          float4 r;
          r.x = (__uint_as_float(v[4 * qq + 0]) + __uint_as_float(w[4 * qq + 0])) * scale;
          r.y = (__uint_as_float(v[4 * qq + 1]) + __uint_as_float(w[4 * qq + 1])) * scale;
          r.z = (__uint_as_float(v[4 * qq + 2]) + __uint_as_float(w[4 * qq + 2])) * scale;
          r.w = (__uint_as_float(v[4 * qq + 3]) + __uint_as_float(w[4 * qq + 3])) * scale;
          *reinterpret_cast<float4 *>(stage_p + lane * 128 + ((qq ^ (lane & 7)) << 4)) = r;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0 && !(p.diag & 8)) {
          tma_store_3d(&tmOut, stage_a, c0, m0 + t2 * kDM + (warp & 3) * 32, split);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) p.dbg[4 * blockIdx.x] = clock64() - t_start;
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(kDMT * NB)) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map_3d_d(CUtensorMap *tm, const float *ptr, unsigned long long dim0, unsigned long long dim1, unsigned long long dim2,
                  unsigned box0, unsigned box1); // defined below

int make_map_2d_d(CUtensorMap *tm, CUtensorMapDataType dt, const void *ptr, unsigned long long dim0, unsigned long long dim1,
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
int make_map_3d_h(CUtensorMap *tm, const void *ptr, unsigned long long dim0, unsigned long long dim1, unsigned long long dim2,
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

// fp32 {dim0 contiguous, dim1, dim2} with dense strides, box {box0, box1, 1}, SWIZZLE_128B (the split-K partial tensor)
int make_map_3d_d(CUtensorMap *tm, const float *ptr, unsigned long long dim0, unsigned long long dim1, unsigned long long dim2,
                  unsigned box0, unsigned box1) {
  void *f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B200_ERR_CUDA;
  }
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {dim0 * 4, dim0 * dim1 * 4};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = ((EncodeTiledFn)f)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(ptr), dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed (%d): dims %llu x %llu x %llu ptr %p", (int)r, dim0, dim1, dim2, ptr);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

template <int NB, bool FOLD, int NX = kDNS, int ND = kDNS>
int launch_dw16(const CUtensorMap &tx, const CUtensorMap &td, const CUtensorMap &tout, const Dw16Params &p, dim3 grid, cudaStream_t st) {
  auto kern = dw16_kernel<NB, FOLD, NX, ND>;
  constexpr int smem = DPlan<NB, NX, ND>::kTotal;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  static long long *dbg = nullptr;
  const bool timing = env().tc_timing;
  Dw16Params pp = p;
  pp.diag = env().diag;
  if (timing) {
    if (!dbg) B200_CUDA(cudaMalloc(&dbg, sizeof(long long) * 4 * 1024));
    B200_CUDA(cudaMemsetAsync(dbg, 0, sizeof(long long) * 4 * 1024, st));
    pp.dbg = dbg;
  }
  B200_CUDA(launch_ex(kern, grid, dim3(kDThreads), (size_t)smem, st, 1, tx, td, tout, pp));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (timing) {
    std::vector<long long> h(4 * 1024);
    B200_CUDA(cudaMemcpyAsync(h.data(), dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    const int n = std::min(1024, (int)grid.x), nf = std::min(n, p.full_groups * p.splits);
    double tot = 0, iw = 0, ib = 0, tot_t = 0;
    for (int i = 0; i < nf; ++i) { tot += h[4 * i]; iw += h[4 * i + 1]; ib += h[4 * i + 2]; }
    for (int i = nf; i < n; ++i) tot_t += h[4 * i];
    fprintf(stderr, "[dw16 timing] NB %d rings %d/%d grid %d x %d + %d, K blocks/CTA %d (tail %d): per CTA total %.0f clk (tail CTAs %.0f), issuer waiting: "
            "X %.0f, delta %.0f\n", NB, NX, ND, p.full_groups, p.splits, p.tail_splits, p.kb_per_split, p.kb_per_tail, tot / std::max(1, nf),
            tot_t / std::max(1, n - nf), iw / std::max(1, nf), ib / std::max(1, nf));
  }
  return B200_OK;
}

} // namespace

bool dw16_applicable(const b200_net *net) {
  if (env().dw16 == 0) return false; // debugging aid: 0 = generic tcgen05 kernel
  const int K0 = net->dims[0], N0 = net->dims[1];
  if (!(net->prec != B200_PREC_FP32 && K0 % 16 == 0 && (N0 == 64 || N0 == 128) && tail_applicable(net))) return false;
  // two layers: the tail writes the fp16 delta_0; deeper: the DX kernel of layer 1 does, with a chained scale bound
  return net->nlayers() == 2 || tail_chain16_applicable(net);
}

// Split plan: one CTA per (group of two feature tiles, slice of the samples), never more CTAs than SMs. A last group that holds
// one tile only (tail) runs half the MMAs per K block: it is given ~0.6x as many, longer slices (its CTAs still fetch whole delta
// tiles), so that all CTAs finish together. Returns the K blocks per slice of the full groups.
struct Dw16Plan { int full_groups, splits, per, tail_splits, per_tail, tail_row0; };
static Dw16Plan dw16_plan_full(const b200_net *net, long batch) {
  Dw16Plan pl{};
  const int rows = net->dims[0] + 1, group_rows = kDMT * kDM;
  const int kblocks = ceil_div(batch, kDK);
  const int rem = rows % group_rows;
  const bool tail = rem > 0 && rem <= kDM && env().dw_tail != 0;
  pl.full_groups = tail ? rows / group_rows : ceil_div(rows, group_rows);
  // at least kMinKbPerSplit K blocks per slice: a short shard (several GPUs, mini-batches) then writes fewer slices, and the
  // combine pass (finalize_grad_kernel), whose cost is the number of slices, shrinks with it
  const int by_k = ceil_div(kblocks, kMinKbPerSplit);
  const int sms = net->ctx->num_sms;
  int s = tail ? (int)((double)sms / (pl.full_groups + 0.6)) : sms / std::max(1, pl.full_groups);
  s = std::max(1, std::min(s, by_k));
  pl.per = ceil_div(kblocks, s);
  pl.splits = ceil_div(kblocks, pl.per);
  if (tail) {
    int st = std::min(sms - pl.full_groups * pl.splits, (int)(0.6 * pl.splits + 0.5));
    st = std::max(1, std::min(st, pl.splits));
    if (pl.full_groups == 0) { st = std::max(1, std::min(sms, by_k)); }
    pl.per_tail = ceil_div(kblocks, st);
    pl.tail_splits = ceil_div(kblocks, pl.per_tail);
    pl.tail_row0 = pl.full_groups * group_rows;
  }
  return pl;
}
int dw16_plan(const b200_net *net, long batch, int *splits) {
  const Dw16Plan pl = dw16_plan_full(net, batch);
  *splits = std::max(pl.splits, pl.tail_splits); // slices the partial buffer must hold
  return pl.per;
}

// layer 0 [dW; db] partials from the uint8 input copy and the fp16 {hi | lo} delta written by tail_layer(want16)
int dw16_layer(b200_net *net, const X16View &x16, long batch, bool *done) {
  *done = false;
  if (!x16.base || !net->delta16) return B200_OK;
  const int K0 = net->dims[0], N0 = net->dims[1];
  CUtensorMap tx, td, tout;
  // A atoms: box {64 features, 32 samples, 1 block} of the block-major fp16 copy = 4 contiguous KB each; feature K0 reads 1
  // (its row of D is the bias gradient), blocks past the last one are out of bounds = zero
  B200_TRY(make_map_3d_h(&tx, x16.base, 64, (unsigned long long)x16.rows_total, (unsigned long long)x16.nblocks, 64, kDK));
  B200_TRY(make_map_2d_d(&td, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, net->delta16, 2 * N0, batch, (unsigned long long)2 * N0 * 2, 64, kDK,
                         CU_TENSOR_MAP_SWIZZLE_128B));
  const Dw16Plan pl = dw16_plan_full(net, batch);
  const int splits = std::max(pl.splits, pl.tail_splits);
  Dw16Params p{};
  p.in_dim = K0; p.out_dim = N0;
  p.k_blocks = ceil_div(batch, kDK); p.kb_per_split = pl.per;
  p.full_groups = pl.full_groups; p.splits = pl.full_groups > 0 ? pl.splits : 0; p.tail_splits = pl.tail_splits; p.kb_per_tail = pl.per_tail;
  p.partial = net->partials + net->part_off[0];
  p.partial_stride = (unsigned long long)(K0 + 1) * N0;
  p.scale_inv = net->scale16_inv;
  p.row0 = (int)x16.row0;
  p.spec_st = net->spec_st; p.spec = net->spec_flag;
  const dim3 grid(p.full_groups * p.splits + p.tail_splits);
  B200_TRY(make_map_3d_d(&tout, p.partial, N0, K0 + 1, splits, 32, 32));
  if (N0 == 128 && (env().ring & 2)) B200_TRY((launch_dw16<256, false, 8, 4>(tx, td, tout, p, grid, net->ctx->stream)));
  else if (N0 == 128) B200_TRY((launch_dw16<256, false>(tx, td, tout, p, grid, net->ctx->stream)));
  else B200_TRY((launch_dw16<128, false>(tx, td, tout, p, grid, net->ctx->stream)));
  // finalize_grad_kernel combines exactly the slices each row range has: rows [0, tail_row0) splits, the rest tail_splits
  net->splits_used[0] = pl.full_groups > 0 ? pl.splits : pl.tail_splits;
  net->dw0_tail_row0 = (pl.tail_splits > 0 && pl.full_groups > 0) ? pl.tail_row0 : -1;
  net->dw0_tail_splits = pl.tail_splits;
  *done = true;
  return B200_OK;
}

// split plan of the mid16 dW: ONE group of two M tiles (hi / lo of the 128 features), so up to one split per SM
int mid16_dw_plan(const b200_net *net, long batch, int *splits) {
  const int kblocks = ceil_div(batch, kDK);
  const int s = std::max(1, std::min(net->ctx->num_sms, ceil_div(kblocks, kMinKbPerSplit)));
  const int per = ceil_div(kblocks, s);
  *splits = ceil_div(kblocks, per);
  return per;
}

// mid16: dW_1 partials = A_1^T delta_1 from the two fp16 pairs (b200_net::Mid16); the bias gradient comes from the last-layer
// backward kernel's column sums (db_part)
int mid16_dw_layer1(b200_net *net, long batch) {
  b200_net::Mid16 &m = net->m16;
  const int K1 = net->dims[1], N1 = net->dims[2];
  CUtensorMap tx, td, tout;
  B200_TRY(make_map_3d_h(&tx, m.a16, 64, (unsigned long long)batch, (unsigned long long)(2 * K1 / 64), 64, kDK));
  B200_TRY(make_map_2d_d(&td, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, m.d16, 2 * N1, batch, (unsigned long long)2 * N1 * 2, 64, kDK,
                         CU_TENSOR_MAP_SWIZZLE_128B));
  int splits = 1;
  const int per = mid16_dw_plan(net, batch, &splits);
  Dw16Params p{};
  p.in_dim = 2 * K1 - 1; p.out_dim = N1; // (in_dim + 1 = the 256 rows of the pair: both M tiles hold features)
  p.k_blocks = ceil_div(batch, kDK); p.kb_per_split = per;
  p.full_groups = 1; p.splits = splits; p.tail_splits = 0; p.kb_per_tail = per;
  p.partial = net->partials + net->part_off[1];
  p.partial_stride = (unsigned long long)(K1 + 1) * N1;
  p.scale_inv = m.scale1_inv;
  p.rowscale = m.tinv;
  p.row0 = 0;
  p.spec_st = net->spec_st; p.spec = net->spec_flag;
  const dim3 grid(splits);
  B200_TRY(make_map_3d_d(&tout, p.partial, N1, K1 + 1, splits, 32, 32));
  if (N1 == 128) B200_TRY((launch_dw16<256, true>(tx, td, tout, p, grid, net->ctx->stream)));
  else B200_TRY((launch_dw16<128, true>(tx, td, tout, p, grid, net->ctx->stream)));
  net->splits_used[1] = splits;
  return B200_OK;
}

} // namespace b200

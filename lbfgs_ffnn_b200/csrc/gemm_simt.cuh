// FP32 (FFMA) tiled GEMM with fused epilogues — the B200_PREC_FP32 arithmetic mode and the
// path for the skinny layers (out = 10, K = 10) that cannot fill a tensor-core tile.
//
// One kernel template covers the three GEMM roles of a dense layer; everything the reference
// does in separate element-wise kernels is folded into the epilogue:
//   FWD      A = act(X W + b)                      (src/cuda/layer.cuh:48-58, kernels.cuh:74-106)
//   FWD_LAST same + diff, 0.5||diff||^2 partials, delta = diff/B * act'(A)
//                                                   (src/cuda/network.cuh:100-107, kernels.cuh:136-141)
//   DX       delta_prev = (delta W^T) .* act'(A_prev)  (layer.cuh:89-103 + kernels.cuh:109-133 of the
//                                                   next backward() call, fused into the producer)
//   DW       [dW; db] partial = [A_prev | 1]^T delta over one batch slice (split-K over the batch)
//                                                   (layer.cuh:81-86; replaces the serial sum_rows_kernel)
//
// Storage (reference layout, column-major == "sample-major"): activations/deltas are [B][features],
// W is [in][out] (out contiguous), so with C[m][n] = sum_k A(m,k) B(k,n):
//   FWD: m = sample, n = out, k = in : A k-contiguous,  B n-contiguous
//   DX : m = sample, n = in,  k = out: A k-contiguous,  B k-contiguous
//   DW : m = in(+1), n = out, k = sample: A m-contiguous, B n-contiguous
//
// Tile: 128 x (16*TN) x 16, 256 threads, 8 x TN register micro-tile, register-prefetch double buffer.
#pragma once

#include "common.cuh"

namespace b200 {

enum { EPI_FWD = 0, EPI_FWD_LAST = 1, EPI_DX = 2, EPI_DW = 3 };

struct GemmParams {
  const float *A;
  const float *B;
  long lda, ldb;
  int M, N, K;
  int vecA, vecB; // 16-byte vector loads legal
  int a_ones_row; // DW: index m that reads as 1.0 (the bias row), else -1
  int k_chunk;    // DW: samples per split (multiple of 16); others: K
  // epilogue
  const float *bias;  // FWD
  int act;            // FWD: this layer's activation; DX: previous layer's activation
  float *out;         // FWD: activations [M][N]; DX: delta_prev [M][N]; DW: partials [split][M*N]
  long ldo;
  const float *aux;   // FWD_LAST: targets [M][N]; DX: A_prev [M][N]
  long ld_aux;        // row stride of aux (DX: act_{l-1} is [M][N] while delta_{l-1} = out is [M][ldo])
  float *delta;       // FWD_LAST: delta out [M][ldd]
  long ldd;
  float inv_batch;    // FWD_LAST: 1/B_global
  double *loss_part;  // FWD_LAST: one partial per CTA
  const SpecState *spec_st; // speculative launch on a wrong guess: return at once (common.cuh)
  int spec;
};

constexpr int kBM = 128, kBK = 16, kThreads = 256, kPad = 4;

template <int MN, bool KC>
struct TileRegs {
  static constexpr int kVecs = (MN * kBK / 4 + kThreads - 1) / kThreads;
  float v[kVecs][4];
};

// Load a MN x kBK tile into registers. Element (mn, k) lives at P[mn*ld + k] (KC) or P[k*ld + mn].
template <int MN, bool KC>
__device__ __forceinline__ void tile_load(TileRegs<MN, KC> &r, const float *__restrict__ P, long ld, int mn0, int k0,
                                          int mn_end, int k_end, bool vec, int ones_mn) {
  constexpr int kTotal = MN * kBK / 4;
#pragma unroll
  for (int i = 0; i < TileRegs<MN, KC>::kVecs; ++i) {
    const int j = threadIdx.x + i * kThreads;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < kTotal) {
      if constexpr (KC) {
        const int mn = mn0 + (j >> 2), k = k0 + ((j & 3) << 2);
        if (mn < mn_end) {
          const float *src = P + (long)mn * ld + k;
          if (vec && k + 3 < k_end) {
            val = __ldg(reinterpret_cast<const float4 *>(src));
          } else {
            if (k + 0 < k_end) val.x = __ldg(src + 0);
            if (k + 1 < k_end) val.y = __ldg(src + 1);
            if (k + 2 < k_end) val.z = __ldg(src + 2);
            if (k + 3 < k_end) val.w = __ldg(src + 3);
          }
        }
      } else {
        constexpr int kPerRow = MN / 4;
        const int k = k0 + j / kPerRow, mn = mn0 + (j % kPerRow) * 4;
        if (k < k_end) {
          const float *src = P + (long)k * ld + mn;
          if (vec && mn + 3 < mn_end && (ones_mn < mn || ones_mn > mn + 3)) {
            val = __ldg(reinterpret_cast<const float4 *>(src));
          } else {
            float t[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const int q = mn + s;
              t[s] = (q == ones_mn) ? 1.0f : ((q < mn_end) ? __ldg(src + s) : 0.0f);
            }
            val = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
      }
    }
    r.v[i][0] = val.x; r.v[i][1] = val.y; r.v[i][2] = val.z; r.v[i][3] = val.w;
  }
}

// Store the registers into smem laid out as S[k][mn] with row stride MN + kPad.
template <int MN, bool KC>
__device__ __forceinline__ void tile_store(const TileRegs<MN, KC> &r, float *__restrict__ S) {
  constexpr int kTotal = MN * kBK / 4;
  constexpr int kLd = MN + kPad;
#pragma unroll
  for (int i = 0; i < TileRegs<MN, KC>::kVecs; ++i) {
    const int j = threadIdx.x + i * kThreads;
    if (j < kTotal) {
      if constexpr (KC) {
        const int mn = j >> 2, k = (j & 3) << 2;
#pragma unroll
        for (int s = 0; s < 4; ++s) S[(k + s) * kLd + mn] = r.v[i][s];
      } else {
        constexpr int kPerRow = MN / 4;
        const int k = j / kPerRow, mn = (j % kPerRow) * 4;
        *reinterpret_cast<float4 *>(&S[k * kLd + mn]) = make_float4(r.v[i][0], r.v[i][1], r.v[i][2], r.v[i][3]);
      }
    }
  }
}

template <int TN, bool A_KC, bool B_KC, int EPI>
__global__ void __launch_bounds__(kThreads) gemm_simt_kernel(const GemmParams p) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  if (spec_skip(p.spec_st, p.spec)) return;
  constexpr int BN = 16 * TN;
  constexpr int kLdA = kBM + kPad, kLdB = BN + kPad;
  __shared__ __align__(16) float sA[2][kBK * kLdA];
  __shared__ __align__(16) float sB[2][kBK * kLdB];
  __shared__ double s_red[32];

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
  int k_begin = 0, k_end = p.K;
  if (EPI == EPI_DW) {
    k_begin = blockIdx.z * p.k_chunk;
    k_end = min(p.K, k_begin + p.k_chunk);
  }
  // DW: the ones row is part of M but not of the stored matrix
  const int a_mn_end = (EPI == EPI_DW) ? ((p.a_ones_row >= 0) ? p.M - 1 : p.M) : p.M;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  TileRegs<kBM, A_KC> ra;
  TileRegs<BN, B_KC> rb;
  const int nk = (k_end - k_begin + kBK - 1) / kBK;
  if (nk > 0) {
    tile_load<kBM, A_KC>(ra, p.A, p.lda, m0, k_begin, a_mn_end, k_end, p.vecA, p.a_ones_row);
    tile_load<BN, B_KC>(rb, p.B, p.ldb, n0, k_begin, p.N, k_end, p.vecB, -1);
    tile_store<kBM, A_KC>(ra, sA[0]);
    tile_store<BN, B_KC>(rb, sB[0]);
  }
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      const int k0 = k_begin + (kt + 1) * kBK;
      tile_load<kBM, A_KC>(ra, p.A, p.lda, m0, k0, a_mn_end, k_end, p.vecA, p.a_ones_row);
      tile_load<BN, B_KC>(rb, p.B, p.ldb, n0, k0, p.N, k_end, p.vecB, -1);
    }
    const float *a_s = sA[cur], *b_s = sB[cur];
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4 *>(&a_s[k * kLdA + ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&a_s[k * kLdA + 64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      if constexpr (TN == 8) {
        const float4 b0 = *reinterpret_cast<const float4 *>(&b_s[k * kLdB + tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4 *>(&b_s[k * kLdB + 64 + tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
      } else if constexpr (TN == 4) {
        const float4 b0 = *reinterpret_cast<const float4 *>(&b_s[k * kLdB + tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      } else if constexpr (TN == 2) {
        const float2 b0 = *reinterpret_cast<const float2 *>(&b_s[k * kLdB + tx * 2]);
        b[0] = b0.x; b[1] = b0.y;
      } else {
        b[0] = b_s[k * kLdB + tx];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      tile_store<kBM, A_KC>(ra, sA[cur ^ 1]);
      tile_store<BN, B_KC>(rb, sB[cur ^ 1]);
    }
    __syncthreads();
  }

  // ---- epilogue --------------------------------------------------------------------------------
  double loss_local = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)));
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n;
      if constexpr (TN == 8) n = n0 + ((j < 4) ? (tx * 4 + j) : (64 + tx * 4 + (j - 4)));
      else n = n0 + tx * TN + j;
      if (n >= p.N) continue;
      const float v = acc[i][j];
      if constexpr (EPI == EPI_FWD || EPI == EPI_FWD_LAST) {
        const float a = act_apply(p.act, v + __ldg(p.bias + n));
        p.out[(long)m * p.ldo + n] = a;
        if constexpr (EPI == EPI_FWD_LAST) {
          const float d = a - __ldg(p.aux + (long)m * p.ld_aux + n);
          loss_local += (double)d * (double)d;
          p.delta[(long)m * p.ldd + n] = d * p.inv_batch * act_deriv_from_output(p.act, a);
        }
      } else if constexpr (EPI == EPI_DX) {
        const float ap = __ldg(p.aux + (long)m * p.ld_aux + n);
        p.out[(long)m * p.ldo + n] = v * act_deriv_from_output(p.act, ap);
      } else { // EPI_DW: partial of the flat [ (in+1) x out ] gradient block of this layer
        p.out[(long)blockIdx.z * ((long)p.M * p.N) + (long)m * p.N + n] = v;
      }
    }
  }
  if constexpr (EPI == EPI_FWD_LAST) {
    const double s = block_sum(loss_local, s_red);
    if (threadIdx.x == 0) p.loss_part[blockIdx.y * gridDim.x + blockIdx.x] = s;
  }
}

// host-side dispatch over TN
template <bool A_KC, bool B_KC, int EPI>
inline int launch_gemm_simt(const GemmParams &p, int splits, cudaStream_t stream) {
  const int tn = (p.N > 64) ? 8 : (p.N > 32 ? 4 : (p.N > 16 ? 2 : 1));
  const int bn = 16 * tn;
  dim3 grid(ceil_div(p.M, kBM), ceil_div(p.N, bn), splits);
  switch (tn) {
  case 8: B200_LAUNCH((gemm_simt_kernel<8, A_KC, B_KC, EPI>), grid, kThreads, 0, stream, p); break;
  case 4: B200_LAUNCH((gemm_simt_kernel<4, A_KC, B_KC, EPI>), grid, kThreads, 0, stream, p); break;
  case 2: B200_LAUNCH((gemm_simt_kernel<2, A_KC, B_KC, EPI>), grid, kThreads, 0, stream, p); break;
  default: B200_LAUNCH((gemm_simt_kernel<1, A_KC, B_KC, EPI>), grid, kThreads, 0, stream, p); break;
  }
  return B200_OK;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

} // namespace b200

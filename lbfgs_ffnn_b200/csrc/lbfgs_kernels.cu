// Kernels of the compact-form L-BFGS direction and the fused line-search vector operations.
// See lbfgs_kernels.cuh for the algebra and the reference lines each kernel replaces.
#include "lbfgs_kernels.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <vector>

namespace b200 {

namespace {

// a / b in double without the ~40-instruction IEEE division routine on the kernel's serial path: fp32 reciprocal refined by
// two Newton steps (relative error ~1e-16, not correctly rounded; b normal and finite)
__device__ __forceinline__ double div_fast(double a, double b) {
  double r = (double)(1.0f / (float)b);
  r = fma(fma(-b, r, 1.0), r, r);
  r = fma(fma(-b, r, 1.0), r, r);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q); // one correction of the quotient
}
__device__ __forceinline__ int ring_phys(int head, int count, int mod, int logical) {
  int start = (head - count) % mod; // src/cuda/lbfgs.cuh:225-230
  if (start < 0) start += mod;
  return (start + logical) % mod;
}

__device__ __forceinline__ float4 ld4(const float *p, size_t i, size_t n, bool vec) {
  if (vec && i + 3 < n) return __ldg(reinterpret_cast<const float4 *>(p + i));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i + 0 < n) v.x = __ldg(p + i + 0);
  if (i + 1 < n) v.y = __ldg(p + i + 1);
  if (i + 2 < n) v.z = __ldg(p + i + 2);
  if (i + 3 < n) v.w = __ldg(p + i + 3);
  return v;
}
__device__ __forceinline__ void st4(float *p, size_t i, size_t n, bool vec, float4 v) {
  if (vec && i + 3 < n) {
    *reinterpret_cast<float4 *>(p + i) = v;
    return;
  }
  if (i + 0 < n) p[i + 0] = v.x;
  if (i + 1 < n) p[i + 1] = v.y;
  if (i + 2 < n) p[i + 2] = v.z;
  if (i + 3 < n) p[i + 3] = v.w;
}

// ------------------------------------------------------------------------------------------------
// (1) batched GEMV over the ring. Each CTA streams tiles of kDotsTile elements: the tile of
// g / s_new / y_new is staged once in shared memory (as doubles), then every warp owns the ring rows
// r = warp, warp + 8, ... and accumulates its 5 dot products per row in fp64 registers across ALL tiles
// of the CTA, so the shuffle reduction happens once per kernel, not once per tile.
// RPW = rows per warp (compile-time so the accumulators stay in registers).
// ------------------------------------------------------------------------------------------------
// HS = 2 (fused direction kernel, 16 warps): warps 8-15 take the second half of every tile row of the SAME ring rows, so a warp
// waits for one round of loads per row instead of two; the halves are combined through shared memory in a fixed order.
template <int RPW, bool PAIR, int HS = 1>
__device__ __forceinline__ void dots_body(const DotsArgs &a) {
  __shared__ __align__(16) double sh_g[kDotsTile];
  __shared__ __align__(16) double sh_s[PAIR ? kDotsTile : 2];
  __shared__ __align__(16) double sh_y[PAIR ? kDotsTile : 2];
  __shared__ int sh_rows[kMaxSlots];
  __shared__ int sh_nrows, sh_w;
  __shared__ double sh_red[32];
  __shared__ double sh_half[HS == 2 ? 2 * RPW * kDotsWarps * kDotsCols : 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = (tid >> 5) % kDotsWarps, whalf = (tid >> 5) / kDotsWarps;
  const bool stager = tid < kDotsThreads; // (HS == 2: the first 256 threads stage the tile)
  const int mp = a.st.h->mp, mod = a.st.h->mod;
  if (tid == 0) {
    int head = a.st.h->head, count = a.st.h->count;
    if (a.reset_first) head = count = 0;
    const int w = head;
    int nr = 0;
    for (int p = 0; p < mod; ++p) {
      int rel = (p - (head - count)) % mod;
      if (rel < 0) rel += mod;
      const bool valid = rel < count;
      if (valid || (PAIR && p == w)) sh_rows[nr++] = p;
    }
    sh_nrows = nr;
    sh_w = w;
  }
  __syncthreads();
  const int nrows = sh_nrows, w = sh_w;

  const bool vec = ((reinterpret_cast<uintptr_t>(a.g) | reinterpret_cast<uintptr_t>(a.S) |
                     reinterpret_cast<uintptr_t>(a.Y) | reinterpret_cast<uintptr_t>(a.x) |
                     reinterpret_cast<uintptr_t>(a.x_prev) | reinterpret_cast<uintptr_t>(a.g_prev)) & 15u) == 0 &&
                   (a.ld % 4 == 0);

  double acc[RPW][kDotsCols];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int c = 0; c < kDotsCols; ++c) acc[r][c] = 0.0;
  double gg = 0.0;

  const size_t ntiles = (a.n + kDotsTile - 1) / kDotsTile;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t base = tile * kDotsTile;
    __syncthreads(); // previous tile fully consumed
    if (stager) { // stage: one float4 per thread
      const size_t i = base + (size_t)tid * 4;
      const float4 g4 = ld4(a.g, i, a.n, vec);
      if (a.row_begin == 0) gg += (double)g4.x * g4.x + (double)g4.y * g4.y + (double)g4.z * g4.z + (double)g4.w * g4.w;
      sh_g[tid * 4 + 0] = g4.x; sh_g[tid * 4 + 1] = g4.y; sh_g[tid * 4 + 2] = g4.z; sh_g[tid * 4 + 3] = g4.w;
      if constexpr (PAIR) {
        float4 s4, y4;
        if (a.mode == DOTS_FORM_PAIR && a.row_begin > 0) { // already formed and stored by the first launch
          s4 = ld4(a.S + (size_t)w * a.ld, i, a.n, vec);
          y4 = ld4(a.Y + (size_t)w * a.ld, i, a.n, vec);
        } else if (a.mode == DOTS_FORM_PAIR) {
          const float4 x4 = ld4(a.x, i, a.n, vec), xp4 = ld4(a.x_prev, i, a.n, vec), gp4 = ld4(a.g_prev, i, a.n, vec);
          s4 = make_float4(x4.x - xp4.x, x4.y - xp4.y, x4.z - xp4.z, x4.w - xp4.w);
          y4 = make_float4(g4.x - gp4.x, g4.y - gp4.y, g4.z - gp4.z, g4.w - gp4.w);
          st4(a.S + (size_t)w * a.ld, i, a.n, vec, s4);
          st4(a.Y + (size_t)w * a.ld, i, a.n, vec, y4);
        } else {
          s4 = ld4(a.S + (size_t)w * a.ld, i, a.n, vec);
          y4 = ld4(a.Y + (size_t)w * a.ld, i, a.n, vec);
        }
        sh_s[tid * 4 + 0] = s4.x; sh_s[tid * 4 + 1] = s4.y; sh_s[tid * 4 + 2] = s4.z; sh_s[tid * 4 + 3] = s4.w;
        sh_y[tid * 4 + 0] = y4.x; sh_y[tid * 4 + 1] = y4.y; sh_y[tid * 4 + 2] = y4.z; sh_y[tid * 4 + 3] = y4.w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int ri = a.row_begin + warp + r * kDotsWarps;
      if (ri >= nrows) break;
      const int p = sh_rows[ri];
      const float *Sp = a.S + (size_t)p * a.ld, *Yp = a.Y + (size_t)p * a.ld;
      const bool self = PAIR && (p == w) && (a.mode == DOTS_FORM_PAIR); // row being written this pass: use smem copy
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (HS == 2 && half != whalf) continue;
        float4 sv[4], yv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = (half * 4 + j) * 128 + lane * 4;
          if (self) {
            sv[j] = make_float4((float)sh_s[e], (float)sh_s[e + 1], (float)sh_s[e + 2], (float)sh_s[e + 3]);
            yv[j] = make_float4((float)sh_y[e], (float)sh_y[e + 1], (float)sh_y[e + 2], (float)sh_y[e + 3]);
          } else {
            sv[j] = ld4(Sp, base + e, a.n, vec);
            yv[j] = ld4(Yp, base + e, a.n, vec);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = (half * 4 + j) * 128 + lane * 4;
          const double s_[4] = {sv[j].x, sv[j].y, sv[j].z, sv[j].w};
          const double y_[4] = {yv[j].x, yv[j].y, yv[j].z, yv[j].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double gq = sh_g[e + q];
            acc[r][0] = fma(s_[q], gq, acc[r][0]);
            acc[r][2] = fma(y_[q], gq, acc[r][2]);
            if constexpr (PAIR) {
              const double sn = sh_s[e + q], yn = sh_y[e + q];
              acc[r][1] = fma(s_[q], yn, acc[r][1]);
              acc[r][3] = fma(y_[q], sn, acc[r][3]);
              acc[r][4] = fma(y_[q], yn, acc[r][4]);
            }
          }
        }
      }
    }
  }

  // one reduction per kernel
  const int ncols = kDotsCols * mp + 1;
  double *out = a.partials + (size_t)blockIdx.x * ncols;
  if (a.row_begin == 0)
    for (int c = tid; c < ncols; c += blockDim.x) out[c] = 0.0;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int ri = a.row_begin + warp + r * kDotsWarps;
#pragma unroll
    for (int c = 0; c < kDotsCols; ++c) {
      const double v = warp_sum(acc[r][c]);
      if (HS == 2) { if (lane == 0) sh_half[((whalf * RPW + r) * kDotsWarps + warp) * kDotsCols + c] = v; }
      else if (lane == 0 && ri < nrows) out[sh_rows[ri] * kDotsCols + c] = v;
    }
  }
  if (HS == 2) {
    __syncthreads();
    if (tid < RPW * kDotsWarps * kDotsCols) {
      const int c = tid % kDotsCols, wr = tid / kDotsCols, w_ = wr % kDotsWarps, r = wr / kDotsWarps;
      const int ri = a.row_begin + w_ + r * kDotsWarps;
      if (ri < nrows) out[sh_rows[ri] * kDotsCols + c] = sh_half[tid] + sh_half[RPW * kDotsWarps * kDotsCols + tid];
    }
  }
  gg = block_sum(gg, sh_red);
  if (tid == 0 && a.row_begin == 0) out[kDotsCols * mp] = gg;
}
template <int RPW, bool PAIR>
__global__ void __launch_bounds__(kDotsThreads) lbfgs_dots_kernel(const DotsArgs a) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  dots_body<RPW, PAIR>(a);
}

// ------------------------------------------------------------------------------------------------
// (1b) the same batched GEMV for LONG vectors (BASELINE configs[4]: n = 2·10⁷, m = 20 — the history no longer fits L2 and
// the pass is HBM-bound). dots_body keeps its loads in registers: at three rows per warp the compiler holds 48 16-byte loads per
// lane (255 registers, one CTA of 8 warps per SM), and the warps of a CTA load and compute in lockstep between the per-tile
// barriers — 2.4 TB/s (measured with the ring full: 1 550 µs for 3.7 GB). Here the bytes in flight do not depend on registers:
// four producer warps (one per scheduler) stream every ring row's 1 KB chunk of a 256-element tile (plus g and, when the pair is
// formed, x / x_prev / g_prev) into a shared-memory ring with plain bulk async copies (cp.async.bulk, completion counted on an
// mbarrier); the consumer warps own the same ring rows as in dots_body (row = warp, warp + 8, ...), HS warps per row group
// with 256 / HS elements of the tile each, and keep g / s_new / y_new of their positions in registers as doubles across their
// rows. Per-CTA partials in the layout of dots_body, so the solve kernel does not know which of the two ran. Needs 16-byte
// aligned rows (the `vec` condition of dots_body); the up to three elements past the last multiple of four are added by lane 0
// of the first warp of every row group from global memory.
// Measured at n = 2·10⁷, m = 20, ring full (B200_DIAG=1024 prints where the warps of CTA 0 wait): one producer warp next to two
// busy consumer warps issued a copy per 75 clk — 3 500 clk per 46 KB stage, twice what HBM needs; four producers: the copies
// alone (B200_DIAG=256) run at 8.0 TB/s.
// ------------------------------------------------------------------------------------------------
constexpr int kBulkTile = 256;                  // floats per row per stage
constexpr int kBulkSlot = kBulkTile * 4;        // bytes of one row chunk
constexpr int kBulkProducers = 4;               // producer warps (one per scheduler of the SM)
constexpr int kBulkMaxStages = 8;
constexpr int kBulkVecSlots = 4;                // g, x, x_prev, g_prev
__host__ __device__ constexpr int bulk_threads(int hs) { return kDotsThreads * hs + 32 * kBulkProducers; }

template <int RPW, bool PAIR, int HS, bool SHARE = false>
__global__ void __launch_bounds__(bulk_threads(HS), 1) lbfgs_dots_bulk_kernel(const DotsArgs a, int stages, int stage_slots, int nparts, int diag) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  constexpr int kCW = kDotsWarps * HS;  // consumer warps
  constexpr int kPos = 2 / HS;          // 16-byte positions of a tile row per lane
  constexpr int kE = 4 * kPos;          // elements per lane and tile row
  extern __shared__ __align__(128) uint8_t bulk_smem[];
  __shared__ __align__(8) uint64_t bars[2 * kBulkMaxStages];
  __shared__ int sh_rows[kMaxSlots];
  __shared__ int sh_nrows, sh_w, sh_wri;
  __shared__ double sh_half[HS == 2 ? 2 * RPW * kDotsWarps * kDotsCols : 1];
  __shared__ double sh_gg[kDotsWarps * HS];
  // SHARE: g / s_new / y_new of the tile being consumed as doubles, converted once per CTA (two buffers, see the consumers)
  __shared__ __align__(16) double vec_d[SHARE ? 2 * 3 * kBulkTile : 2];
  static_assert(!SHARE || HS == 2, "the shared conversion is written for sixteen consumer warps");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = warp % kDotsWarps, wh = warp / kDotsWarps; // consumers: row group, part of the tile
  const int mp = a.st.h->mp, mod = a.st.h->mod;
  const bool form = PAIR && a.mode == DOTS_FORM_PAIR;
  if (tid == 0) {
    int head = a.st.h->head, count = a.st.h->count;
    if (a.reset_first) head = count = 0;
    const int w = head;
    int nr = 0, wri = -1;
    for (int p = 0; p < mod; ++p) {
      int rel = (p - (head - count)) % mod;
      if (rel < 0) rel += mod;
      const bool valid = rel < count;
      if (valid || (PAIR && p == w)) {
        if (p == w) wri = nr;
        sh_rows[nr++] = p;
      }
    }
    sh_nrows = nr;
    sh_w = w;
    sh_wri = wri;
    for (int s = 0; s < stages; ++s) {
      tcx::mbar_init(tcx::smem_u32(&bars[s]), kBulkProducers);            // full: the producers' arrives + the bytes of the copies
      tcx::mbar_init(tcx::smem_u32(&bars[kBulkMaxStages + s]), kCW);      // empty: one arrive per consumer warp
    }
    tcx::fence_mbar_init();
  }
  // a short last tile leaves the end of its chunks as they were: zeros or earlier (finite) history values, never uninitialised
  // shared memory, so masking g / s_new / y_new is enough
  const int stage_bytes = stage_slots * kBulkSlot;
  for (int i = tid; i < stages * stage_bytes / 16; i += bulk_threads(HS)) reinterpret_cast<float4 *>(bulk_smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  tcx::fence_async_smem();
  __syncthreads();
  const int nrows = sh_nrows, w = sh_w, wri = sh_wri;
  const size_t n4 = a.n & ~(size_t)3;
  const size_t ntiles = (n4 + kBulkTile - 1) / kBulkTile;
  const uint32_t smem0 = tcx::smem_u32(bulk_smem);

  double acc[RPW][kDotsCols];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int c = 0; c < kDotsCols; ++c) acc[r][c] = 0.0;
  double gg = 0.0;
  long long dbg_wait = 0, dbg_work = 0, dbg_n = 0; // (B200_DIAG 1024: where the warps of CTA 0 wait)

  if (warp >= kCW) { // ---- producers: one copy per lane, the same slot for the whole kernel ----
    const int nvec = form ? kBulkVecSlots : 1;
    const int skip = form ? wri : -1; // the row being formed holds last round's pair: not loaded
    const int c = (warp - kCW) + kBulkProducers * lane;
    bool active = c < kBulkVecSlots + 2 * nrows;
    const float *src = nullptr;
    if (active) {
      if (c < kBulkVecSlots) {
        active = c < nvec;
        src = c == 0 ? a.g : c == 1 ? a.x : c == 2 ? a.x_prev : a.g_prev;
      } else {
        const int ri = (c - kBulkVecSlots) >> 1;
        active = ri != skip && !(diag & 512); // (B200_DIAG 512: timing without the row copies)
        src = (((c - kBulkVecSlots) & 1) ? a.Y : a.S) + (size_t)sh_rows[ri] * a.ld;
      }
    }
    const uint32_t nact = (uint32_t)__popc(__ballot_sync(0xffffffffu, active));
    int s = 0;        // stage and its phase as running counters (no division per stage)
    uint32_t ph = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long tw0 = (diag & 1024) ? clock64() : 0;
      tcx::mbar_wait(tcx::smem_u32(&bars[kBulkMaxStages + s]), ph ^ 1u);
      if (diag & 1024) { dbg_wait += clock64() - tw0; ++dbg_n; }
      const size_t base = tile * kBulkTile;
      const uint32_t bytes = (n4 - base < (size_t)kBulkTile ? (uint32_t)(n4 - base) : (uint32_t)kBulkTile) * 4u;
      const uint32_t full = tcx::smem_u32(&bars[s]);
      if (lane == 0) tcx::mbar_expect_tx(full, bytes * nact);
      __syncwarp();
      if (active) tcx::bulk_load_1d(smem0 + (uint32_t)(s * stage_bytes + c * kBulkSlot), src + base, bytes, full);
      if (diag & 1024) { __syncwarp(); dbg_work += clock64() - tw0; }
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
    if ((diag & 1024) && blockIdx.x == 0 && lane == 0)
      printf("[dots_bulk] producer warp %d: %lld stages, %lld clk waiting for a free stage, %lld clk in all per stage\n", warp, dbg_n, dbg_wait / max(dbg_n, 1LL), dbg_work / max(dbg_n, 1LL));
  } else { // ---- consumers ----
    // this lane's element offset inside a row chunk, and whether this warp holds the row being formed (its pair goes through the
    // row's own — never loaded — slots of the stage, so that the row loop below has ONE form)
    const int e0 = (HS == 2 ? wh : 0) * 128 + lane * 4;
    const bool owner = form && wri >= 0 && (wri % kDotsWarps) == wr;
    int s = 0, turn = 0, par = 0; // stage, its phase, the row group whose turn it is to store the new pair, (SHARE) the buffer: running counters
    uint32_t ph = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long tw0 = (diag & 1024) ? clock64() : 0;
      tcx::mbar_wait(tcx::smem_u32(&bars[s]), ph);
      if (diag & 1024) { dbg_wait += clock64() - tw0; ++dbg_n; }
      float *stg = reinterpret_cast<float *>(bulk_smem + (size_t)s * stage_bytes) + e0;
      const size_t base = tile * kBulkTile;
      if (diag & 256) { // (B200_DIAG 256: timing without the consumers' work)
        __syncwarp();
        if (lane == 0) tcx::mbar_arrive(tcx::smem_u32(&bars[kBulkMaxStages + s]));
        if (++s == stages) { s = 0; ph ^= 1u; }
        continue;
      }
      const int valid = n4 - base < (size_t)kBulkTile ? (int)(n4 - base) : kBulkTile;
      double gd[kE], snd[PAIR ? kE : 1], ynd[PAIR ? kE : 1];
      if constexpr (SHARE) {
        // One conversion per element and CTA instead of one per element and row group: F2F.F64.F32 runs at 16 lanes per clock and SM
        // and bounds this kernel at the power-limited clock (576 -> 408 per stage). Thread t < 256 takes g and s_new of element t,
        // thread 256 + t y_new; the doubles go to a buffer laid out so that a lane reads its four elements with two conflict-free
        // 16-byte loads per vector. Two buffers: a warp that runs ahead writes the other one, and cannot come back to this one
        // before every warp has passed the next stage's barrier, i.e. has finished reading it.
        float *st0 = stg - e0;
        const int e = tid & (kBulkTile - 1);
        const bool in = e < valid;
        const int idx = ((((e >> 7) * 2 + ((e & 3) >> 1)) * 32 + ((e & 127) >> 2)) << 1) + (e & 1);
        double *vd = vec_d + par * (3 * kBulkTile);
        const float gf = in ? st0[e] : 0.f;
        if (tid < kBulkTile) {
          vd[idx] = gf;
          gg = fma((double)gf, (double)gf, gg);
          if constexpr (PAIR) {
            float sf = form ? st0[kBulkTile + e] - st0[2 * kBulkTile + e] : st0[(kBulkVecSlots + 2 * wri) * kBulkTile + e];
            if (!in) sf = 0.f;
            if (form) {
              if (in) a.S[(size_t)w * a.ld + base + e] = sf;
              st0[(kBulkVecSlots + 2 * wri) * kBulkTile + e] = sf; // (the formed row's own, never loaded, slot: one form of the row loop)
            }
            vd[kBulkTile + idx] = sf;
          }
        } else if constexpr (PAIR) {
          float yf = form ? gf - st0[3 * kBulkTile + e] : st0[(kBulkVecSlots + 2 * wri + 1) * kBulkTile + e];
          if (!in) yf = 0.f;
          if (form) {
            if (in) a.Y[(size_t)w * a.ld + base + e] = yf;
            st0[(kBulkVecSlots + 2 * wri + 1) * kBulkTile + e] = yf;
          }
          vd[2 * kBulkTile + idx] = yf;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kDotsThreads * HS) : "memory");
        const double2 *v2 = reinterpret_cast<const double2 *>(vd) + (wh * 2) * 32 + lane;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const double2 g2 = v2[j * 32];
          gd[2 * j] = g2.x; gd[2 * j + 1] = g2.y;
          if constexpr (PAIR) {
            const double2 s2 = v2[kBulkTile / 2 + j * 32], y2 = v2[kBulkTile + j * 32];
            snd[2 * j] = s2.x; snd[2 * j + 1] = s2.y;
            ynd[2 * j] = y2.x; ynd[2 * j + 1] = y2.y;
          }
        }
        par ^= 1;
      } else {
#pragma unroll
      for (int h = 0; h < kPos; ++h) {
        float *sp = stg + h * 128;
        float4 g4 = *reinterpret_cast<const float4 *>(sp);
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), y4 = s4;
        if constexpr (PAIR) {
          if (form) {
            const float4 x4 = *reinterpret_cast<const float4 *>(sp + kBulkTile);
            const float4 xp4 = *reinterpret_cast<const float4 *>(sp + 2 * kBulkTile);
            const float4 gp4 = *reinterpret_cast<const float4 *>(sp + 3 * kBulkTile);
            s4 = make_float4(x4.x - xp4.x, x4.y - xp4.y, x4.z - xp4.z, x4.w - xp4.w);
            y4 = make_float4(g4.x - gp4.x, g4.y - gp4.y, g4.z - gp4.z, g4.w - gp4.w);
          } else {
            s4 = *reinterpret_cast<const float4 *>(sp + (kBulkVecSlots + 2 * wri) * kBulkTile);
            y4 = *reinterpret_cast<const float4 *>(sp + (kBulkVecSlots + 2 * wri + 1) * kBulkTile);
          }
        }
        if (valid < kBulkTile && e0 + h * 128 >= valid) g4 = s4 = y4 = make_float4(0.f, 0.f, 0.f, 0.f); // (the short last tile only)
        if constexpr (PAIR) {
          if (form) {
            if (wr == turn && e0 + h * 128 < valid) { // the row groups take turns storing the new pair
              *reinterpret_cast<float4 *>(a.S + (size_t)w * a.ld + base + e0 + h * 128) = s4;
              *reinterpret_cast<float4 *>(a.Y + (size_t)w * a.ld + base + e0 + h * 128) = y4;
            }
            if (owner) {
              *reinterpret_cast<float4 *>(sp + (kBulkVecSlots + 2 * wri) * kBulkTile) = s4;
              *reinterpret_cast<float4 *>(sp + (kBulkVecSlots + 2 * wri + 1) * kBulkTile) = y4;
            }
          }
          snd[4 * h + 0] = s4.x; snd[4 * h + 1] = s4.y; snd[4 * h + 2] = s4.z; snd[4 * h + 3] = s4.w;
          ynd[4 * h + 0] = y4.x; ynd[4 * h + 1] = y4.y; ynd[4 * h + 2] = y4.z; ynd[4 * h + 3] = y4.w;
        }
        gd[4 * h + 0] = g4.x; gd[4 * h + 1] = g4.y; gd[4 * h + 2] = g4.z; gd[4 * h + 3] = g4.w;
      }
      if (owner) __syncwarp(); // (each lane reads back exactly what it wrote; this orders the two for the compiler as well)
      if (wr == 0) {
#pragma unroll
        for (int q = 0; q < kE; ++q) gg = fma(gd[q], gd[q], gg);
      }
      }
      const float *rowp = stg + (kBulkVecSlots + 2 * wr) * kBulkTile;
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        if (wr + r * kDotsWarps >= nrows) break;
        double sd[kE], yd[kE];
#pragma unroll
        for (int h = 0; h < kPos; ++h) {
          const float4 s4 = *reinterpret_cast<const float4 *>(rowp + r * (2 * kDotsWarps * kBulkTile) + h * 128);
          const float4 y4 = *reinterpret_cast<const float4 *>(rowp + r * (2 * kDotsWarps * kBulkTile) + kBulkTile + h * 128);
          sd[4 * h + 0] = s4.x; sd[4 * h + 1] = s4.y; sd[4 * h + 2] = s4.z; sd[4 * h + 3] = s4.w;
          yd[4 * h + 0] = y4.x; yd[4 * h + 1] = y4.y; yd[4 * h + 2] = y4.z; yd[4 * h + 3] = y4.w;
        }
#pragma unroll
        for (int q = 0; q < kE; ++q) {
          acc[r][0] = fma(sd[q], gd[q], acc[r][0]);
          acc[r][2] = fma(yd[q], gd[q], acc[r][2]);
          if constexpr (PAIR) {
            acc[r][1] = fma(sd[q], ynd[q], acc[r][1]);
            acc[r][3] = fma(yd[q], snd[q], acc[r][3]);
            acc[r][4] = fma(yd[q], ynd[q], acc[r][4]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) tcx::mbar_arrive(tcx::smem_u32(&bars[kBulkMaxStages + s]));
      if (diag & 1024) dbg_work += clock64() - tw0;
      if (++s == stages) { s = 0; ph ^= 1u; }
      turn = (turn + 1) & (kDotsWarps - 1);
    }
    if ((diag & 1024) && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kCW - 1))
      printf("[dots_bulk] consumer warp %d: %lld stages, %lld clk waiting for a full stage, %lld clk in all per stage\n", warp, dbg_n, dbg_wait / max(dbg_n, 1LL), dbg_work / max(dbg_n, 1LL));
    // the up to three elements past the last multiple of four: lane 0 of the first warp of every row group, from global memory,
    // in the CTA that would own the next tile (any fixed choice keeps the partials deterministic)
    if (n4 < a.n && lane == 0 && wh == 0 && blockIdx.x == (unsigned)(ntiles % gridDim.x)) {
      for (size_t e = n4; e < a.n; ++e) {
        const float gf = __ldg(a.g + e);
        float sf = 0.f, yf = 0.f;
        if constexpr (PAIR) {
          if (form) {
            sf = __ldg(a.x + e) - __ldg(a.x_prev + e);
            yf = gf - __ldg(a.g_prev + e);
            if (wr == 0) { a.S[(size_t)w * a.ld + e] = sf; a.Y[(size_t)w * a.ld + e] = yf; }
          } else {
            sf = __ldg(a.S + (size_t)w * a.ld + e);
            yf = __ldg(a.Y + (size_t)w * a.ld + e);
          }
        }
        const double g1 = gf, sn1 = sf, yn1 = yf;
        if (wr == 0) gg = fma(g1, g1, gg);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          const int ri = wr + r * kDotsWarps;
          if (ri >= nrows) break;
          const bool self = form && ri == wri;
          const double s1 = self ? sn1 : (double)__ldg(a.S + (size_t)sh_rows[ri] * a.ld + e);
          const double y1 = self ? yn1 : (double)__ldg(a.Y + (size_t)sh_rows[ri] * a.ld + e);
          acc[r][0] = fma(s1, g1, acc[r][0]);
          acc[r][2] = fma(y1, g1, acc[r][2]);
          if constexpr (PAIR) {
            acc[r][1] = fma(s1, yn1, acc[r][1]);
            acc[r][3] = fma(y1, sn1, acc[r][3]);
            acc[r][4] = fma(y1, yn1, acc[r][4]);
          }
        }
      }
    }
  }

  // one reduction per kernel, partials as dots_body writes them (HS == 2: the two parts of the tile combined in a fixed order)
  const int ncols = kDotsCols * mp + 1;
  double *out = a.partials + (size_t)blockIdx.x * ncols;
  for (int c = tid; c < ncols; c += bulk_threads(HS)) out[c] = 0.0;
  // one CTA per SM is resident (the ring takes the shared memory), so the grid is one wave; the solve kernel sums `nparts` rows
  // of partials (sized for the register-staged kernel's two CTAs per SM): the rows without a CTA are zeroed here
  for (int b = blockIdx.x + gridDim.x; b < nparts; b += gridDim.x)
    for (int c = tid; c < ncols; c += bulk_threads(HS)) a.partials[(size_t)b * ncols + c] = 0.0;
  if (warp < kCW) {
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
#pragma unroll
      for (int c = 0; c < kDotsCols; ++c) {
        const double v = warp_sum(acc[r][c]);
        if (HS == 2) { if (lane == 0) sh_half[((wh * RPW + r) * kDotsWarps + wr) * kDotsCols + c] = v; }
        else acc[r][c] = v;
      }
    }
    gg = warp_sum(gg); // (zero in the warps that do not accumulate it)
    if (lane == 0) sh_gg[warp] = gg;
  }
  __syncthreads();
  if (HS == 2) {
    if (tid < RPW * kDotsWarps * kDotsCols) {
      const int c = tid % kDotsCols, wq = tid / kDotsCols, w_ = wq % kDotsWarps, r = wq / kDotsWarps;
      const int ri = w_ + r * kDotsWarps;
      if (ri < nrows) out[sh_rows[ri] * kDotsCols + c] = sh_half[tid] + sh_half[RPW * kDotsWarps * kDotsCols + tid];
    }
    if (tid == 0) {
      double sgg = 0.0;
      for (int i = 0; i < kCW; ++i) sgg += sh_gg[i]; // fixed order
      out[kDotsCols * mp] = sgg;
    }
  } else if (warp < kCW) {
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const int ri = wr + r * kDotsWarps;
#pragma unroll
        for (int c = 0; c < kDotsCols; ++c)
          if (ri < nrows) out[sh_rows[ri] * kDotsCols + c] = acc[r][c];
      }
      if (wr == 0) out[kDotsCols * mp] = gg;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// (2) one CTA: reduce the per-CTA partials (fixed order => deterministic), update the Gram blocks for
// the slot written by (1), curvature test + ring advance, then the two-loop recurrences in fp64.
// ------------------------------------------------------------------------------------------------
// Per-CTA results of the solve, kept in shared memory for a fused apply phase.
struct SolveLocal {
  int k;
  double cg, alpha0;
  double *cs, *cy; // [mp] (shared memory)
  int *phys;       // [mp]
};

// `leader` CTAs write the device state; in the fused direction kernel every CTA runs this redundantly on the same inputs
// (identical results) so that no second grid-wide barrier is needed, and only CTA 0 is the leader. head_in / count_in: the
// ring header as it was BEFORE this direction (the leader overwrites it while other CTAs may still be reading).
// STAGED (compile time): the Gram blocks are known to sit in shared memory, so the recurrences read them with shared-memory
// loads; otherwise a.stage_gram decides at run time and the accesses are generic
template <bool STAGED = false>
__device__ __forceinline__ void solve_body(const SolveArgs &a, double *sh, bool leader, int head_in, int count_in, SolveLocal *res,
                                           long long *dbg = nullptr) {
  const long long td0 = dbg ? clock64() : 0;
  // sh: tot[ncols] | alpha[mp] | dlt[mp] | cs[mp] | cy[mp] | (staged) SY[mp*mp] | YY[mp*mp] | rho, sg, yg [mp] | phys[mp] (int)
  LbfgsHeader *h = a.st.h;
  const int mp = h->mp, mod = h->mod;
  const int ncols = kDotsCols * mp + 1;
  double *tot = sh, *alpha = sh + ncols, *dlt = alpha + mp, *lcs = dlt + mp, *lcy = lcs + mp;
  __shared__ int s_head, s_count, s_w, s_k;
  __shared__ double s_cg, s_alpha0;
  const bool stage = STAGED || a.stage_gram != 0;
  double *SYp = a.st.SY, *YYp = a.st.YY, *rhop = a.st.rho, *sgp = a.st.sg, *ygp = a.st.yg;
  int *physp = a.st.phys;
  if (stage) {
    SYp = lcy + mp; YYp = SYp + mp * mp; rhop = YYp + mp * mp; sgp = rhop + mp; ygp = sgp + mp;
    physp = reinterpret_cast<int *>(ygp + mp);
  }

  // fixed-order (deterministic) reduction of the per-CTA partials. Short histories (all columns fit the CTA with several
  // threads each): ONE pass, every thread sums a strided slice of the blocks with 16 loads in flight per step and the threads
  // of a column are combined by shuffles, i.e. ~nblocks / (16 tpc) dependent L2 round trips. Otherwise 32 columns at a time,
  // 8 threads per column.
  int tpc = 1; // (from 256 threads whatever the CTA size: the fused and the stand-alone solve then sum in the same order)
  while (tpc < 8 && ncols * (tpc * 2) <= 256) tpc *= 2;
  if (tpc >= 2) {
    const int c = threadIdx.x / tpc, sub = threadIdx.x % tpc;
    double s = 0.0;
    if (c < ncols) {
      for (int b0 = sub; b0 < a.nblocks; b0 += 16 * tpc) {
        double t[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) t[u] = (b0 + tpc * u < a.nblocks) ? __ldcg(a.partials + (size_t)(b0 + tpc * u) * ncols + c) : 0.0;
        s += (((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]))) +
             (((t[8] + t[9]) + (t[10] + t[11])) + ((t[12] + t[13]) + (t[14] + t[15])));
      }
    }
    for (int o = 1; o < tpc; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (c < ncols && sub == 0) tot[c] = s;
  } else {
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      const int c = c0 + (threadIdx.x >> 3), sub = threadIdx.x & 7;
      double s = 0.0;
      if (c < ncols) {
        for (int b0 = sub; b0 < a.nblocks; b0 += 64) { // eight independent loads in flight per step, same summation order
          double t[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] = (b0 + 8 * u < a.nblocks) ? __ldcg(a.partials + (size_t)(b0 + 8 * u) * ncols + c) : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) s += t[u];
        }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (c < ncols && sub == 0) tot[c] = s;
    }
  }
  const long long td1 = dbg ? clock64() : 0;
  if (stage) { // everything the 2k dependent recurrence steps touch lives in shared memory (an L2 round trip per step otherwise)
    for (int i = threadIdx.x; i < mp * mp; i += blockDim.x) { SYp[i] = a.st.SY[i]; YYp[i] = a.st.YY[i]; }
    for (int i = threadIdx.x; i < mp; i += blockDim.x) { rhop[i] = a.st.rho[i]; sgp[i] = a.st.sg[i]; ygp[i] = a.st.yg[i]; }
  }
  if (threadIdx.x == 0) {
    int head = head_in, count = count_in;
    if (a.reset_first) head = count = 0;
    s_head = head; s_count = count; s_w = head;
  }
  __syncthreads();
  const int w = s_w;
  const bool pair = a.mode != DOTS_NONE;
  const bool wr_global = leader && stage; // staged: mirror the updates into the device state; unstaged: the pointers ARE the state

  // Gram update for every row the dots kernel visited (valid rows + w)
  for (int p = threadIdx.x; p < mod; p += blockDim.x) {
    int rel = (p - (s_head - s_count)) % mod;
    if (rel < 0) rel += mod;
    const bool valid = rel < s_count;
    if (!(valid || (pair && p == w))) continue;
    const double v0 = tot[p * kDotsCols + 0], v1 = tot[p * kDotsCols + 1], v2 = tot[p * kDotsCols + 2],
                 v3 = tot[p * kDotsCols + 3], v4 = tot[p * kDotsCols + 4];
    if (stage || leader) {
      sgp[p] = v0;
      ygp[p] = v2;
      if (pair) {
        SYp[p * mp + w] = v1; // s_p . y_w
        SYp[w * mp + p] = v3; // s_w . y_p
        YYp[p * mp + w] = v4;
        YYp[w * mp + p] = v4;
      }
    }
    if (wr_global) {
      a.st.sg[p] = v0;
      a.st.yg[p] = v2;
      if (pair) {
        a.st.SY[p * mp + w] = v1;
        a.st.SY[w * mp + p] = v3;
        a.st.YY[p * mp + w] = v4;
        a.st.YY[w * mp + p] = v4;
      }
    }
  }
  __syncthreads();

  if (threadIdx.x == 0) {
    int head = s_head, count = s_count, flags = 0;
    double ys = 0.0, yy = 0.0;
    if (pair) {
      ys = SYp[w * mp + w];
      yy = YYp[w * mp + w];
      // curvature filter y^T s > 1e-10 (src/cuda/lbfgs.cuh:161-169, src/minimizer/lbfgs.hpp:77-84)
      // S-LBFGS: |y^T s| > 1e-10 (src/minimizer/s_lbfgs.hpp:253-258)
      const bool accept = a.force_accept ? true : (a.policy == POLICY_SLBFGS ? fabs(ys) > 1e-10 : ys > 1e-10);
      if (accept) {
        const double r = a.force_accept ? a.ext_rho : ((fabs(ys) > 1e-30 && fabs(ys) < 1e30) ? div_fast(1.0, ys) : 1.0 / ys);
        if (stage || leader) rhop[w] = r;
        if (wr_global) a.st.rho[w] = r;
        head = (head + 1) % mod;
        count = min(count + 1, h->m);
        flags |= FLAG_PAIR_ACCEPTED;
      }
    }
    if (leader) {
      if (pair) { h->ys_new = ys; h->yy_new = yy; }
      h->head = head; h->count = count; h->flags = flags;
      h->gnorm2 = tot[kDotsCols * mp];
    }
    s_head = head; s_count = count;
  }
  __syncthreads();

  const long long td2 = dbg ? clock64() : 0;
  // two-loop recurrences on the Gram blocks: warp 0, lanes parallel over the inner sums
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int head = s_head, k = s_count;
    const double gg = tot[kDotsCols * mp];
    for (int i = lane; i < k; i += 32) {
      const int ph = ring_phys(head, k, mod, i);
      physp[i] = ph;
      if (wr_global) a.st.phys[i] = ph;
    }
    __syncwarp();
    double gamma = 1.0, cg = -1.0, gdotp = -gg;
    if (k > 0 && k <= 32) {
      // Histories of up to 32 pairs: lane i owns logical slot i and both recurrences run in their column (axpy) form — at step j
      // the lane of slot j finishes its value, broadcasts it with one shuffle and every lane still waiting folds it into its own
      // running sum. The dependent chain is one shuffle + one FMA per slot (~50 clk) instead of the j-term dot product a single
      // lane used to walk through shared memory (the old form took ~15 k clk at k = 10: 40 % of the whole direction kernel).
      if (dbg && lane == 0) dbg[4] = clock64() - td2;
      const bool on = lane < k;
      const int pi = on ? physp[lane] : 0;
      const double rho_i = on ? rhop[pi] : 0.0, sg_i = on ? sgp[pi] : 0.0, yg_i = on ? ygp[pi] : 0.0;
      // the matrix entries this lane needs come out of shared memory BEFORE the chains start (they do not depend on them):
      // row i of S^T Y and Y^T Y and column i of S^T Y in logical order, pre-multiplied by rho_i where the recurrences use them so
      const double *sy_row = SYp + pi * mp, *yy_row = YYp + pi * mp;
      // alpha_j = rho_j (s_j.g - sum_{l > j} alpha_l s_j.y_l): lane i keeps a_i = rho_i (s_i.g - partial sum); the chain per slot is
      // one shuffle + one FMA
      double av = rho_i * sg_i, alpha_i = 0.0;
      {
        double m_next = -rho_i * sy_row[physp[k - 1]];
        for (int j = k - 1; j >= 0; --j) {
          const double mj = m_next;
          if (j > 0) m_next = -rho_i * sy_row[physp[j - 1]];
          const double aj = __shfl_sync(0xffffffffu, av, j);
          if (lane == j) alpha_i = aj;
          if (lane < j) av = fma(aj, mj, av);
        }
      }
      if (dbg && lane == 0) dbg[5] = clock64() - td2;
      const int pl = physp[k - 1];
      const double ys = SYp[pl * mp + pl], yy = YYp[pl * mp + pl];
      const bool yy_ok = fabs(yy) > 1e-30 && fabs(yy) < 1e30; // (float range of the seed reciprocal; else the IEEE routine)
      const double ratio = yy_ok ? div_fast(ys, yy) : ys / yy;
      if (a.policy == POLICY_ARMIJO) gamma = (yy > 0.0) ? ratio : 1.0; // src/cuda/lbfgs.cuh:244-247
      else if (a.policy == POLICY_WOLFE) gamma = ratio;                 // src/minimizer/lbfgs.hpp:124-125
      else {                                                            // src/minimizer/s_lbfgs.hpp:116-124
        gamma = (fabs(yy) < 1e-12) ? 1.0 : ratio;
        gamma = fmin(fmax(gamma, 1e-6), 1e6);
      }
      // sum_j alpha_j y_i.y_j: no dependence on the second recurrence; four partial sums keep the FMA chain short
      double s1p[4] = {0.0, 0.0, 0.0, 0.0};
      for (int j0 = 0; j0 < k; j0 += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = min(j0 + u, k - 1);
          const double aj = __shfl_sync(0xffffffffu, alpha_i, j);
          s1p[u] = fma(j0 + u < k ? aj : 0.0, yy_row[physp[j]], s1p[u]);
        }
      }
      const double s1 = (s1p[0] + s1p[1]) + (s1p[2] + s1p[3]);
      if (dbg && lane == 0) dbg[6] = clock64() - td2;
      // delta_j = alpha_j - rho_j (gamma (y_j.g - s1_j) + sum_{l < j} delta_l s_l.y_j): lane i keeps d_i = alpha_i - rho_i (...)
      double dv = alpha_i - rho_i * (gamma * (yg_i - s1)), dlt_i = 0.0;
      {
        double m_next = -rho_i * SYp[physp[0] * mp + pi];
        for (int j = 0; j < k; ++j) {
          const double mj = m_next;
          if (j + 1 < k) m_next = -rho_i * SYp[physp[j + 1] * mp + pi];
          const double dj = __shfl_sync(0xffffffffu, dv, j);
          if (lane == j) dlt_i = dj;
          if (lane > j) dv = fma(dj, mj, dv);
        }
      }
      if (dbg && lane == 0) dbg[7] = clock64() - td2;
      cg = -gamma;
      const double csj = -dlt_i, cyj = gamma * alpha_i;
      if (on) {
        lcs[lane] = csj;
        lcy[lane] = cyj;
        if (leader) { a.st.cs[lane] = csj; a.st.cy[lane] = cyj; }
      }
      const double gp = warp_sum(on ? csj * sg_i + cyj * yg_i : 0.0);
      gdotp = cg * gg + gp;
    } else if (k > 0) {
      for (int i = k - 1; i >= 0; --i) {
        const int pi = physp[i];
        double s = 0.0;
        for (int j = i + 1 + lane; j < k; j += 32) s += alpha[j] * SYp[pi * mp + physp[j]];
        s = warp_sum(s);
        if (lane == 0) alpha[i] = rhop[pi] * (sgp[pi] - s);
        __syncwarp();
      }
      const int pl = physp[k - 1];
      const double ys = SYp[pl * mp + pl], yy = YYp[pl * mp + pl];
      if (a.policy == POLICY_ARMIJO) gamma = (yy > 0.0) ? ys / yy : 1.0; // src/cuda/lbfgs.cuh:244-247
      else if (a.policy == POLICY_WOLFE) gamma = ys / yy;                 // src/minimizer/lbfgs.hpp:124-125
      else {                                                              // src/minimizer/s_lbfgs.hpp:116-124
        gamma = (fabs(yy) < 1e-12) ? 1.0 : ys / yy;
        gamma = fmin(fmax(gamma, 1e-6), 1e6);
      }
      for (int i = 0; i < k; ++i) {
        const int pi = physp[i];
        double s1 = 0.0, s2 = 0.0;
        for (int j = lane; j < k; j += 32) s1 += alpha[j] * YYp[pi * mp + physp[j]];
        for (int j = lane; j < i; j += 32) s2 += dlt[j] * SYp[physp[j] * mp + pi];
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) {
          const double beta = rhop[pi] * (gamma * (ygp[pi] - s1) + s2);
          dlt[i] = alpha[i] - beta;
        }
        __syncwarp();
      }
      cg = -gamma;
      double gp = 0.0;
      for (int j = lane; j < k; j += 32) {
        const double csj = -dlt[j], cyj = gamma * alpha[j];
        lcs[j] = csj;
        lcy[j] = cyj;
        if (leader) { a.st.cs[j] = csj; a.st.cy[j] = cyj; }
        gp += csj * sgp[physp[j]] + cyj * ygp[physp[j]];
      }
      gp = warp_sum(gp);
      gdotp = cg * gg + gp;
    }
    __syncwarp();
    if (dbg && lane == 0) dbg[2] = clock64() - td2;
    if (lane == 0) {
      int kk = k;
      // non-descent direction: steepest descent + history reset (src/cuda/lbfgs.cuh:97-104).
      // The CPU backend has no such check (src/minimizer/lbfgs.hpp:56-65).
      const bool sd = a.policy == POLICY_ARMIJO && kk > 0 && !(gdotp < 0.0);
      if (sd) { kk = 0; cg = -1.0; gdotp = -gg; gamma = 1.0; }
      double alpha0 = 1.0; // lbfgs.cuh:108, lbfgs.hpp:60-61
      if (a.first_iter) alpha0 = fmin(1.0, 1.0 / sqrt(gg)); // (the square root and the division only where they are used)
      if (leader) {
        if (sd) { h->head = 0; h->count = 0; h->flags |= FLAG_SD_FALLBACK; }
        h->k = kk; h->cg = cg; h->gdotp = gdotp; h->gamma = gamma;
        h->alpha0 = alpha0;
        if (a.host_hdr) { // zero-copy mailbox: the host reads it after the graph's end event
          *a.host_hdr = *h;
          __threadfence_system();
        }
      }
      s_k = kk; s_cg = cg; s_alpha0 = alpha0;
    }
  }
  __syncthreads();
  if (dbg && threadIdx.x == 0) { dbg[0] = td1 - td0; dbg[1] = td2 - td1; dbg[3] = clock64() - td2; }
  if (res) { res->k = s_k; res->cg = s_cg; res->alpha0 = s_alpha0; res->cs = lcs; res->cy = lcy; res->phys = physp; }
}

__global__ void __launch_bounds__(256) lbfgs_solve_kernel(const SolveArgs a) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  extern __shared__ double sh[];
  solve_body(a, sh, true, a.st.h->head, a.st.h->count, nullptr);
}

// ------------------------------------------------------------------------------------------------
// (3) p = cg g + sum_j cs_j s_j + cy_j y_j  (fp64 accumulate, one rounding); x_prev = x; x += alpha0 p
// ------------------------------------------------------------------------------------------------
// COHERENT: the history slot written earlier in the SAME kernel (fused direction) must not come through the read-only path
template <bool COHERENT> __device__ __forceinline__ float4 ld4c(const float *p, size_t i, size_t n, bool vec) {
  if constexpr (!COHERENT) return ld4(p, i, n, vec);
  if (vec && i + 3 < n) return __ldcg(reinterpret_cast<const float4 *>(p + i));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i + 0 < n) v.x = __ldcg(p + i + 0);
  if (i + 1 < n) v.y = __ldcg(p + i + 1);
  if (i + 2 < n) v.z = __ldcg(p + i + 2);
  if (i + 3 < n) v.w = __ldcg(p + i + 3);
  return v;
}

// U: history pairs whose loads are in flight together (2 U 16-byte loads per step)
template <bool COHERENT, int U = 4>
__device__ __forceinline__ void apply_body(const ApplyArgs &a, int k, double cg_in, double alpha0_in, const double *s_cs,
                                           const double *s_cy, const int *s_ph) {
  const double cg = cg_in * a.sign;
  const float alpha0 = (a.step != 0.0f) ? a.step : (float)alpha0_in;
  const bool vec = ((reinterpret_cast<uintptr_t>(a.g) | reinterpret_cast<uintptr_t>(a.S) |
                     reinterpret_cast<uintptr_t>(a.Y) | reinterpret_cast<uintptr_t>(a.p) |
                     reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.x_prev) |
                     reinterpret_cast<uintptr_t>(a.x_copy)) & 15u) == 0 &&
                   (a.ld % 4 == 0);
  const size_t nv = (a.n + 3) / 4;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (size_t)gridDim.x * blockDim.x) {
    const size_t i = v * 4;
    const float4 g4 = ld4c<COHERENT>(a.g, i, a.n, vec);
    double r0 = cg * g4.x, r1 = cg * g4.y, r2 = cg * g4.z, r3 = cg * g4.w;
    for (int j0 = 0; j0 < k; j0 += U) { // U pairs (2 U 16-byte loads) in flight per step; same order of the fp64 sums
      float4 s4[U], y4[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = min(j0 + u, k - 1);
        s4[u] = ld4c<COHERENT>(a.S + (size_t)s_ph[j] * a.ld, i, a.n, vec);
        y4[u] = ld4c<COHERENT>(a.Y + (size_t)s_ph[j] * a.ld, i, a.n, vec);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j0 + u < k) {
          const double cs = s_cs[j0 + u] * a.sign, cy = s_cy[j0 + u] * a.sign;
          r0 = fma(cs, (double)s4[u].x, r0); r1 = fma(cs, (double)s4[u].y, r1);
          r2 = fma(cs, (double)s4[u].z, r2); r3 = fma(cs, (double)s4[u].w, r3);
          r0 = fma(cy, (double)y4[u].x, r0); r1 = fma(cy, (double)y4[u].y, r1);
          r2 = fma(cy, (double)y4[u].z, r2); r3 = fma(cy, (double)y4[u].w, r3);
        }
      }
    }
    const float4 p4 = make_float4((float)r0, (float)r1, (float)r2, (float)r3);
    st4(a.p, i, a.n, vec, p4);
    if (a.x) {
      const float4 x4 = ld4c<COHERENT>(a.x, i, a.n, vec);
      if (a.x_prev) st4(a.x_prev, i, a.n, vec, x4);
      // one rounding per element like the reference's copy + axpy (lbfgs.cuh:116-117, cuBLAS axpy is an FMA)
      const float4 xn = make_float4(fmaf(alpha0, p4.x, x4.x), fmaf(alpha0, p4.y, x4.y), fmaf(alpha0, p4.z, x4.z),
                                    fmaf(alpha0, p4.w, x4.w));
      st4(a.x, i, a.n, vec, xn);
      if (a.x_copy) st4(a.x_copy, i, a.n, vec, xn);
    }
  }
}

__global__ void __launch_bounds__(256) lbfgs_apply_kernel(const ApplyArgs a) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  __shared__ double s_cs[kMaxSlots], s_cy[kMaxSlots];
  __shared__ int s_ph[kMaxSlots];
  const LbfgsHeader *h = a.st.h;
  const int k = h->k;
  if (threadIdx.x < k) {
    s_cs[threadIdx.x] = a.st.cs[threadIdx.x];
    s_cy[threadIdx.x] = a.st.cy[threadIdx.x];
    s_ph[threadIdx.x] = a.st.phys[threadIdx.x];
  }
  if (k > (int)blockDim.x)
    for (int j = blockDim.x + threadIdx.x; j < k; j += blockDim.x) { s_cs[j] = a.st.cs[j]; s_cy[j] = a.st.cy[j]; s_ph[j] = a.st.phys[j]; }
  __syncthreads();
  apply_body<false>(a, k, h->cg, h->alpha0, s_cs, s_cy, s_ph);
}

// ------------------------------------------------------------------------------------------------
// (1)+(2)+(3) in ONE launch (histories of <= 32 slots, one GPU): dots pass -> grid-wide barrier -> every CTA reduces the
// partials and runs the O(k^2) solve redundantly (identical inputs, identical results; only CTA 0 writes the device state)
// -> apply pass on the same tiles the CTA streamed in phase 1. Saves two launches and their dependency latency per direction;
// needs all CTAs co-resident (checked on the host with the occupancy API).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned nblocks) { // bar[0] arrivals, bar[1] generation
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned *gen = bar + 1;
    const unsigned g = *gen; // read before arriving: it can only change after EVERY CTA has arrived
    __threadfence();
    if (atomicAdd(bar, 1u) == nblocks - 1) {
      bar[0] = 0u; // ready for the next launch
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*gen == g) __nanosleep(40);
    }
    __threadfence();
  }
  __syncthreads();
}

template <int RPW, bool PAIR, int HS>
__global__ void __launch_bounds__(kDotsThreads * HS) lbfgs_direction_kernel(const DotsArgs da, const SolveArgs sa, const ApplyArgs aa,
                                                                      unsigned *bar, const SpecState *spec_st, int spec, long long *dbg) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  if (spec_skip(spec_st, spec)) return; // every CTA takes the same branch: nobody reaches the grid barrier
  extern __shared__ double sh[];
  __shared__ int s_head0, s_count0;
  const long long t0 = dbg ? clock64() : 0;
  if (threadIdx.x == 0) { s_head0 = da.st.h->head; s_count0 = da.st.h->count; } // before anyone can overwrite the header
  dots_body<RPW, PAIR, HS>(da); // (its barriers publish s_head0 / s_count0)
  const long long t1 = dbg ? clock64() : 0;
  grid_barrier(bar, gridDim.x);
  const long long t2 = dbg ? clock64() : 0;
  __shared__ SolveLocal res;
  SolveLocal r;
  solve_body<true>(sa, sh, blockIdx.x == 0, s_head0, s_count0, &r, (dbg && blockIdx.x <= 1) ? dbg + 4096 + 8 * blockIdx.x : nullptr);
  if (threadIdx.x == 0) res = r;
  __syncthreads();
  const long long t3 = dbg ? clock64() : 0;
  apply_body<true, 6>(aa, res.k, res.cg, res.alpha0, res.cs, res.cy, res.phys);
  if (dbg && threadIdx.x == 0) { // B200_TC_TIMING: phases of this CTA in SM clocks
    dbg[4 * blockIdx.x + 0] = t1 - t0; dbg[4 * blockIdx.x + 1] = t2 - t1; dbg[4 * blockIdx.x + 2] = t3 - t2; dbg[4 * blockIdx.x + 3] = clock64() - t3;
  }
}

// totals[c] = sum_b partials[b][c] in a fixed order (sharded history: the totals are then all-reduced over ranks)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double *__restrict__ partials, int nblocks, int ncols,
                                                             double *__restrict__ totals) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  for (int c0 = blockIdx.x * 32; c0 < ncols; c0 += gridDim.x * 32) {
    const int c = c0 + (threadIdx.x >> 3), sub = threadIdx.x & 7;
    double s = 0.0;
    if (c < ncols)
      for (int b = sub; b < nblocks; b += 8) s += partials[(size_t)b * ncols + c];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (c < ncols && sub == 0) totals[c] = s;
  }
}

__global__ void lbfgs_init_kernel(LbfgsView v, int m, int mod) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  const int mp = m + 1;
  for (int i = threadIdx.x; i < mp * mp; i += blockDim.x) { v.SY[i] = 0.0; v.YY[i] = 0.0; }
  for (int i = threadIdx.x; i < mp; i += blockDim.x) {
    v.rho[i] = 0.0; v.sg[i] = 0.0; v.yg[i] = 0.0; v.cs[i] = 0.0; v.cy[i] = 0.0; v.phys[i] = 0;
  }
  if (threadIdx.x == 0) {
    LbfgsHeader h{};
    h.m = m; h.mp = mp; h.mod = mod; h.cg = -1.0; h.alpha0 = 1.0; h.gamma = 1.0;
    *v.h = h;
  }
}

// copy an explicit (s, y) pair into slot `head` (the Gram update happens in the following dots/solve)
__global__ void __launch_bounds__(256) store_pair_kernel(float *S, float *Y, size_t n, size_t ld, LbfgsView st,
                                                         const float *s, const float *y) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  const int w = st.h->head;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    S[(size_t)w * ld + i] = s[i];
    Y[(size_t)w * ld + i] = y[i];
  }
}

// ---- BLAS-1 replacements -----------------------------------------------------------------------
__global__ void __launch_bounds__(256) trial_point_kernel(size_t n, const float *__restrict__ x0, float alpha,
                                                          const float *__restrict__ p, float *__restrict__ y) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = fmaf(alpha, p[i], x0[i]);
}
__global__ void __launch_bounds__(256) axpy_kernel(size_t n, float alpha, const float *__restrict__ x, float *y) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = fmaf(alpha, x[i], y[i]);
}
__global__ void __launch_bounds__(256) scal_kernel(size_t n, float alpha, float *x) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] *= alpha;
}
__global__ void __launch_bounds__(256) momentum_step_kernel(size_t n, float mu, float lr, const float *__restrict__ g,
                                                            float *v, float *x) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    // v = mu*v; v += -lr*g; x += v  (src/cuda/gd.cuh:77-81: scal, axpy, axpy)
    const float vv = fmaf(-lr, g[i], mu * v[i]);
    v[i] = vv;
    x[i] += vv;
  }
}
__global__ void __launch_bounds__(256) f64_to_f32_kernel(size_t n, const double *__restrict__ s, float *__restrict__ d) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    d[i] = (float)s[i];
}
__global__ void __launch_bounds__(256) dot_part_kernel(const float *__restrict__ x, const float *__restrict__ y, size_t n,
                                                       double *part) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  __shared__ double red[32];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    s = fma((double)x[i], (double)y[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) dot_final_kernel(const double *part, int nparts, double *out) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += part[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out = s;
}

// profiling aid: keeps the GPU busy for ~`clocks` SM clocks so that the host can enqueue the next events and launches behind it;
// without it the first profiled kernel after a host synchronisation is charged the host's launch latency
__global__ void prof_spacer_kernel(long long clocks) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  const long long t0 = clock64();
  while (clock64() - t0 < clocks) {}
}

inline int vec_blocks(size_t n, int cap) { return (int)std::max<size_t>(1, std::min<size_t>((size_t)cap, (n + 255) / 256)); }

} // namespace

int lbfgs_dots_blocks(b200_ctx *ctx, size_t n) {
  const size_t tiles = (n + kDotsTile - 1) / kDotsTile;
  return (int)std::max<size_t>(1, std::min<size_t>((size_t)2 * ctx->num_sms, tiles));
}
int dot_blocks(b200_ctx *ctx, size_t n) { return vec_blocks(n, 2 * ctx->num_sms); }

int launch_lbfgs_dots(const DotsArgs &a0, int mp, int nblocks, cudaStream_t st) {
  B200_REQUIRE(mp <= kMaxSlots, "history size above 256 is not supported");
  const bool pair = a0.mode != DOTS_NONE;
  // long vectors (the history does not fit L2): the bulk-copy form, rows streamed through a shared-memory ring (1b)
  const bool aligned = ((reinterpret_cast<uintptr_t>(a0.g) | reinterpret_cast<uintptr_t>(a0.S) | reinterpret_cast<uintptr_t>(a0.Y) |
                         reinterpret_cast<uintptr_t>(a0.x) | reinterpret_cast<uintptr_t>(a0.x_prev) |
                         reinterpret_cast<uintptr_t>(a0.g_prev)) & 15u) == 0 && a0.ld % 4 == 0;
  if (env().dots_bulk && aligned && mp <= kRowsPerLaunch && (long)a0.n >= env().dots_bulk_min &&
      (a0.mode != DOTS_FORM_PAIR || (a0.x && a0.x_prev && a0.g_prev))) {
    DotsArgs a = a0;
    a.row_begin = 0;
    const int stage_slots = kBulkVecSlots + 2 * mp;
    const int stages = std::min(kBulkMaxStages, (204 * 1024) / (stage_slots * kBulkSlot)); // (+ 13.5 KB static with the shared conversion)
    const size_t smem = (size_t)stages * stage_slots * kBulkSlot;
    const int rpw = ceil_div(mp, kDotsWarps);
    const int hs = env().dots_bulk == 1 ? 1 : 2; // (B200_DOTS_BULK=1: eight consumer warps with two positions per lane; default sixteen)
    const bool share = env().dots_bulk >= 3;     // (default; B200_DOTS_BULK=2: every row group converts g / s_new / y_new itself)
    if (stages >= 2 && rpw >= 1 && rpw <= 4) {
      auto run = [&](auto kern) -> int {
        static bool attr_set[5][2][4] = {}; // (the instantiations share one pointer type, hence one copy of this lambda)
        if (!attr_set[rpw][pair][hs + share]) {
          B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024));
          attr_set[rpw][pair][hs + share] = true;
        }
        static int sms = 0;
        if (!sms) {
          int dev = 0;
          B200_CUDA(cudaGetDevice(&dev));
          B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        }
        B200_CUDA(launch_ex(kern, dim3(std::min(nblocks, sms)), dim3(bulk_threads(hs)), smem, st, 1, a, stages, stage_slots, nblocks, env().diag));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        B200_CUDA(cudaGetLastError());
        return B200_OK;
      };
#define B200_BULK_CASE(R)                                                      \
  case R:                                                                      \
    if (share) {                                                               \
      if (pair) return run(lbfgs_dots_bulk_kernel<R, true, 2, true>);          \
      return run(lbfgs_dots_bulk_kernel<R, false, 2, true>);                   \
    }                                                                          \
    if (hs == 2) {                                                             \
      if (pair) return run(lbfgs_dots_bulk_kernel<R, true, 2>);                \
      return run(lbfgs_dots_bulk_kernel<R, false, 2>);                         \
    }                                                                          \
    if (pair) return run(lbfgs_dots_bulk_kernel<R, true, 1>);                  \
    return run(lbfgs_dots_bulk_kernel<R, false, 1>);
      switch (rpw) {
        B200_BULK_CASE(1)
        B200_BULK_CASE(2)
        B200_BULK_CASE(3)
        B200_BULK_CASE(4)
      default:
        break;
      }
#undef B200_BULK_CASE
    }
  }
  // the row count is device state (<= mp); launches beyond the live rows exit after the staging loop
  for (int row_begin = 0; row_begin < mp; row_begin += kRowsPerLaunch) {
    DotsArgs a = a0;
    a.row_begin = row_begin;
    const int rpw = ceil_div(std::min(mp - row_begin, kRowsPerLaunch), kDotsWarps);
#define B200_DOTS_CASE(R)                                                                            \
  case R:                                                                                            \
    if (pair) B200_LAUNCH((lbfgs_dots_kernel<R, true>), nblocks, kDotsThreads, 0, st, a);            \
    else B200_LAUNCH((lbfgs_dots_kernel<R, false>), nblocks, kDotsThreads, 0, st, a);                \
    break;
    switch (rpw) {
      B200_DOTS_CASE(1)
      B200_DOTS_CASE(2)
      B200_DOTS_CASE(3)
      B200_DOTS_CASE(4)
    default:
      set_error("unsupported rows per warp %d", rpw);
      return B200_ERR_INVALID;
    }
#undef B200_DOTS_CASE
  }
  return B200_OK;
}

static size_t solve_smem_bytes(int mp, bool stage) {
  size_t b = sizeof(double) * (kDotsCols * mp + 1 + 4 * (size_t)mp);
  if (stage) b += sizeof(double) * (2 * (size_t)mp * mp + 3 * mp) + sizeof(int) * mp;
  return b;
}

int launch_lbfgs_solve(const SolveArgs &a0, int mp, cudaStream_t st) {
  SolveArgs a = a0;
  a.stage_gram = solve_smem_bytes(mp, true) <= 200 * 1024;
  const size_t smem = solve_smem_bytes(mp, a.stage_gram != 0);
  if (smem > 48 * 1024) {
    static bool attr_set = false;
    if (!attr_set) {
      B200_CUDA(cudaFuncSetAttribute(lbfgs_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
  }
  B200_LAUNCH(lbfgs_solve_kernel, 1, 256, smem, st, a);
  return B200_OK;
}

// Fused direction: returns false in *done when the shape does not qualify (the caller then issues the three kernels).
int launch_lbfgs_direction(b200_ctx *ctx, const DotsArgs &da0, const SolveArgs &sa0, const ApplyArgs &aa, int mp, int nblocks,
                           unsigned *bar, cudaStream_t st, bool *done, const SpecState *spec_st, int spec) {
  *done = false;
  const size_t smem = solve_smem_bytes(mp, true);
  if (mp > kRowsPerLaunch || smem > 16 * 1024 || !bar) return B200_OK;
  const bool pair = da0.mode != DOTS_NONE;
  const int rpw = ceil_div(mp, kDotsWarps);
  DotsArgs da = da0;
  da.row_begin = 0;
  SolveArgs sa = sa0;
  sa.stage_gram = 1;
  auto run = [&](auto kern, int threads) -> int {
    int per_sm = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if ((long)per_sm * ctx->num_sms < nblocks) return B200_OK; // the grid barrier needs every CTA resident
    static long long *dbg = nullptr;
    const bool timing = env().tc_timing;
    if (timing && !dbg) B200_CUDA(cudaMalloc(&dbg, sizeof(long long) * (4 * 1024 + 64)));
    B200_CUDA(launch_ex(kern, dim3(nblocks), dim3(threads), (size_t)smem, st, 1, da, sa, aa, bar, spec_st, spec, timing ? dbg : (long long *)nullptr));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    B200_CUDA(cudaGetLastError());
    if (timing) {
      std::vector<long long> h(4 * 1024 + 64);
      B200_CUDA(cudaMemcpyAsync(h.data(), dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, st));
      B200_CUDA(cudaStreamSynchronize(st));
      fprintf(stderr, "[solve timing] CTA 0 / CTA 1: partial reduction %lld / %lld clk, Gram staging + update + header %lld / %lld, recurrences %lld / %lld, "
              "recurrences + tail %lld / %lld | CTA 1 cumulative: ring order %lld, first recurrence %lld, s1 %lld, second recurrence %lld\n", h[4096], h[4104], h[4097], h[4105], h[4098], h[4106], h[4099], h[4107], h[4108], h[4109], h[4110], h[4111]);
      double a[4] = {0, 0, 0, 0};
      for (int i = 0; i < nblocks && i < 1024; ++i) for (int j = 0; j < 4; ++j) a[j] += (double)h[4 * i + j] / std::min(nblocks, 1024);
      fprintf(stderr, "[direction timing] %d CTAs, mp %d: dots %.0f clk, grid barrier %.0f, reduce + solve %.0f (CTA 0: %lld), apply %.0f\n", nblocks, mp,
              a[0], a[1], a[2], h[2], a[3]);
    }
    *done = true;
    return B200_OK;
  };
  // up to two rows per warp: 16 warps, the two halves of a tile row on different warps (one round of loads per row)
#define B200_DIR_CASE(R, HS)                                                              \
  case R:                                                                                 \
    if (pair) return run(lbfgs_direction_kernel<R, true, HS>, kDotsThreads * HS);         \
    return run(lbfgs_direction_kernel<R, false, HS>, kDotsThreads * HS);
  switch (rpw) {
    B200_DIR_CASE(1, 2)
    B200_DIR_CASE(2, 2)
    B200_DIR_CASE(3, 1)
    B200_DIR_CASE(4, 1)
  default:
    return B200_OK;
  }
#undef B200_DIR_CASE
}

int launch_lbfgs_apply(const ApplyArgs &a, int nblocks, cudaStream_t st) {
  // a grid-stride kernel: never more CTAs than are resident at once (4 per SM were asked for, 3 fit at 77 registers — the 148 CTAs
  // of a second wave ran at a third of the occupancy)
  static int resident = 0;
  if (!resident) {
    int dev = 0, sms = 0, per_sm = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbfgs_apply_kernel, 256, 0));
    resident = std::max(1, per_sm * sms);
  }
  nblocks = std::min(nblocks, resident);
  B200_LAUNCH(lbfgs_apply_kernel, nblocks, 256, 0, st, a);
  return B200_OK;
}

int launch_reduce_partials(const double *partials, int nblocks, int ncols, double *totals, cudaStream_t st) {
  B200_LAUNCH(reduce_partials_kernel, std::max(1, std::min(8, (ncols + 31) / 32)), 256, 0, st, partials, nblocks, ncols, totals);
  return B200_OK;
}

int lbfgs_init_state(LbfgsView v, int m, int mod, cudaStream_t st) {
  B200_LAUNCH(lbfgs_init_kernel, 1, 256, 0, st, v, m, mod);
  return B200_OK;
}

int launch_lbfgs_store_pair(float *S, float *Y, size_t n, size_t ld, LbfgsView stv, const float *s, const float *y,
                            cudaStream_t stream) {
  B200_LAUNCH(store_pair_kernel, vec_blocks(n, 1184), 256, 0, stream, S, Y, n, ld, stv, s, y);
  return B200_OK;
}

int launch_prof_spacer(cudaStream_t st) {
  prof_spacer_kernel<<<1, 1, 0, st>>>(80000); // ~40 us; not counted as a product launch
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int launch_trial_point(size_t n, const float *x0, float alpha, const float *p, float *y, cudaStream_t st) {
  B200_LAUNCH(trial_point_kernel, vec_blocks(n, 1184), 256, 0, st, n, x0, alpha, p, y);
  return B200_OK;
}
int launch_axpy(size_t n, float alpha, const float *x, float *y, cudaStream_t st) {
  B200_LAUNCH(axpy_kernel, vec_blocks(n, 1184), 256, 0, st, n, alpha, x, y);
  return B200_OK;
}
int launch_scal(size_t n, float alpha, float *x, cudaStream_t st) {
  B200_LAUNCH(scal_kernel, vec_blocks(n, 1184), 256, 0, st, n, alpha, x);
  return B200_OK;
}
int launch_momentum_step(size_t n, float mu, float lr, const float *g, float *v, float *x, cudaStream_t st) {
  B200_LAUNCH(momentum_step_kernel, vec_blocks(n, 1184), 256, 0, st, n, mu, lr, g, v, x);
  return B200_OK;
}
int launch_f64_to_f32(size_t n, const double *src, float *dst, cudaStream_t st) {
  B200_LAUNCH(f64_to_f32_kernel, vec_blocks(n, 1184), 256, 0, st, n, src, dst);
  return B200_OK;
}
int launch_dot(b200_ctx *ctx, const float *x, const float *y, size_t n, double *part, double *out) {
  const int nb = dot_blocks(ctx, n);
  B200_LAUNCH(dot_part_kernel, nb, 256, 0, ctx->stream, x, y, n, part);
  B200_LAUNCH(dot_final_kernel, 1, 256, 0, ctx->stream, part, nb, out);
  return B200_OK;
}

} // namespace b200

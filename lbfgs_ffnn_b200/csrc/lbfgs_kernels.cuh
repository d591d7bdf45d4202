// L-BFGS direction in compact (Gram) form + fused line-search vector kernels.
//
// The reference computes p = -H g with the two-loop recursion as 2k+2 blocking cublasSdot calls,
// 2k cublasSaxpy, 3 D2D memcpys and 2 cublasSscal per direction (src/cuda/lbfgs.cuh:206-261) —
// latency bound (~48 us per stored pair). Here the same recursion is unrolled
// algebraically: with the Gram blocks SY[i][j] = s_i.y_j, YY[i][j] = y_i.y_j and the projections
// S^T g, Y^T g every alpha_i / beta_i is a scalar recurrence,
//     alpha_i = rho_i (s_i.g - sum_{j>i} alpha_j SY[i][j])
//     y_i.q   = y_i.g - sum_j alpha_j YY[i][j]
//     beta_i  = rho_i (gamma y_i.q + sum_{j<i} (alpha_j - beta_j) SY[j][i])
//     p       = -(gamma g - gamma sum_j alpha_j y_j + sum_j (alpha_j - beta_j) s_j)
// so one direction is
//   (1) lbfgs_dots_kernel : ONE coalesced pass over the ring — a batched GEMV [S Y]^T [g s_new y_new]
//       that also forms the newest pair s = x - x_prev, y = g - g_prev in flight and stores it;
//   (2) lbfgs_solve_kernel: one CTA, fp64: Gram row/column update for the new slot, curvature test,
//       ring advance, the recurrences above, gamma, g.p, steepest-descent fallback, first step length;
//   (3) lbfgs_apply_kernel: ONE fused pass p = cg g + [S Y] c, x_prev = x, x = x + alpha0 p.
// All scalars stay on the device; rho_i is kept separately from SY[i][i] so the reference's
// "slot written before the curvature test" behaviour (lbfgs.cuh:149-169) is reproduced exactly.
//
// Algorithmic HBM bytes: (1) reads (2k+4) n floats, writes 2n; (3) reads (2k+2) n, writes 3n.
#pragma once

#include "common.cuh"

namespace b200 {

enum { DOTS_NONE = 0, DOTS_FORM_PAIR = 1, DOTS_PAIR_IN_SLOT = 2 };
enum { POLICY_ARMIJO = 0, POLICY_WOLFE = 1, POLICY_SLBFGS = 2 };
enum { FLAG_SD_FALLBACK = 1, FLAG_PAIR_ACCEPTED = 2 };

constexpr int kMaxSlots = 257; // physical ring slots (m <= 256)
constexpr int kRowsPerLaunch = 32; // ring rows handled by one dots launch (4 per warp)
constexpr int kDotsWarps = 8;
constexpr int kDotsThreads = kDotsWarps * 32;
constexpr int kDotsTile = 1024; // floats staged per CTA step
constexpr int kDotsCols = 5;    // per row: S.g, S.y_new, Y.g, Y.s_new, Y.y_new

// Device-resident optimizer state. One allocation; arrays follow the header.
struct LbfgsHeader {
  int m;      // history size
  int mp;     // physical slots allocated = m + 1
  int mod;    // ring modulus: m (ARMIJO: reference CUDA ring) or m + 1 (WOLFE / S-LBFGS: RingBuffer semantics)
  int head;   // next slot to write
  int count;  // valid pairs
  int k;      // pairs used by the last direction
  int flags;  // FLAG_*
  int pad;
  double gnorm2, gdotp, alpha0, gamma, ys_new, yy_new;
  double cg;  // coefficient of g in p
};

struct LbfgsView { // pointers into the state allocation (computed on host, passed by value)
  LbfgsHeader *h;
  double *rho;  // [mp]
  double *SY;   // [mp*mp]  SY[p*mp+q] = s_p . y_q
  double *YY;   // [mp*mp]
  double *sg;   // [mp] s_p . g  (physical index)
  double *yg;   // [mp]
  double *cs;   // [mp] coefficient of s (logical order of the last direction)
  double *cy;   // [mp]
  int *phys;    // [mp] logical -> physical of the last direction
};

inline size_t lbfgs_state_bytes(int m) {
  const size_t mp = m + 1;
  size_t b = (sizeof(LbfgsHeader) + 15) & ~size_t(15);
  b += sizeof(double) * (mp + 2 * mp * mp + 4 * mp);
  b += sizeof(int) * mp;
  return (b + 255) & ~size_t(255);
}
inline LbfgsView lbfgs_view(void *base, int m) {
  const size_t mp = m + 1;
  LbfgsView v;
  char *c = (char *)base;
  v.h = (LbfgsHeader *)c;
  size_t off = (sizeof(LbfgsHeader) + 15) & ~size_t(15);
  v.rho = (double *)(c + off); off += sizeof(double) * mp;
  v.SY = (double *)(c + off);  off += sizeof(double) * mp * mp;
  v.YY = (double *)(c + off);  off += sizeof(double) * mp * mp;
  v.sg = (double *)(c + off);  off += sizeof(double) * mp;
  v.yg = (double *)(c + off);  off += sizeof(double) * mp;
  v.cs = (double *)(c + off);  off += sizeof(double) * mp;
  v.cy = (double *)(c + off);  off += sizeof(double) * mp;
  v.phys = (int *)(c + off);
  return v;
}

struct DotsArgs {
  float *S, *Y;            // ring storage, slot p at S + p*ld
  size_t n, ld;
  LbfgsView st;
  const float *g;          // current gradient
  const float *x, *x_prev; // DOTS_FORM_PAIR: s = x - x_prev
  const float *g_prev;     // DOTS_FORM_PAIR: y = g - g_prev
  int mode;
  int reset_first;         // apply a history reset (head = count = 0) before anything else
  int row_begin;           // first ring row (index into the row list) handled by this launch
  double *partials;        // [gridDim.x][kDotsCols*mp + 1]
};

struct SolveArgs {
  LbfgsView st;
  const double *partials;
  int nblocks;
  int mode, reset_first, policy;
  int first_iter;        // alpha0 = min(1, 1/||g||), else 1
  int force_accept;      // direct API: accept the pair with the caller's rho
  double ext_rho;
  int allreduce_done;    // unused on one GPU (partials are already global)
  int stage_gram;        // set by launch_lbfgs_solve: the Gram blocks fit the shared-memory staging area
  LbfgsHeader *host_hdr; // optional pinned-host copy of the header, written by the leader (replaces a D2H copy node)
};

struct ApplyArgs {
  const float *S, *Y;
  size_t n, ld;
  LbfgsView st;
  const float *g;
  float *p;
  float *x, *x_prev; // if x != nullptr: x_prev = x (when x_prev != nullptr); x += step * p
  double sign;       // +1: p as computed (-H g); -1: +H g (S-LBFGS two-loop returns H v)
  float step;        // 0: use the header's alpha0; otherwise this step length
  float *x_copy;     // optional second copy of the NEW x (S-LBFGS iterate history)
};

int launch_lbfgs_dots(const DotsArgs &a, int mp, int nblocks, cudaStream_t st);
int launch_lbfgs_solve(const SolveArgs &a, int mp, cudaStream_t st);
int launch_lbfgs_apply(const ApplyArgs &a, int nblocks, cudaStream_t st);
// dots + solve + apply in one launch (grid-wide barrier); *done = false when the shape does not qualify
int launch_lbfgs_direction(b200_ctx *ctx, const DotsArgs &da, const SolveArgs &sa, const ApplyArgs &aa, int mp, int nblocks,
                           unsigned *bar, cudaStream_t st, bool *done, const SpecState *spec_st = nullptr, int spec = 0);
int lbfgs_dots_blocks(b200_ctx *ctx, size_t n);
int lbfgs_init_state(LbfgsView v, int m, int mod, cudaStream_t st);
int launch_reduce_partials(const double *partials, int nblocks, int ncols, double *totals, cudaStream_t st);
// ring push of an explicit pair (S-LBFGS curvature pairs): s, y device vectors copied into slot head
int launch_lbfgs_store_pair(float *S, float *Y, size_t n, size_t ld, LbfgsView st, const float *s, const float *y,
                            cudaStream_t stream);

int launch_prof_spacer(cudaStream_t st); // per-launch profiling only (see lbfgs_kernels.cu)
// y = x0 + alpha * p
int launch_trial_point(size_t n, const float *x0, float alpha, const float *p, float *y, cudaStream_t st);
// two-stage deterministic dot: *out = x.y (double, device). part must hold >= dot_blocks() doubles.
int launch_dot(b200_ctx *ctx, const float *x, const float *y, size_t n, double *part, double *out);
int launch_axpy(size_t n, float alpha, const float *x, float *y, cudaStream_t st);
int launch_scal(size_t n, float alpha, float *x, cudaStream_t st);
// v = mu*v - lr*g ; x += v      (CudaGD / CudaSGD momentum step, src/cuda/gd.cuh:77-84)
int launch_momentum_step(size_t n, float mu, float lr, const float *g, float *v, float *x, cudaStream_t st);
int launch_f64_to_f32(size_t n, const double *src, float *dst, cudaStream_t st);
int dot_blocks(b200_ctx *ctx, size_t n);

} // namespace b200

// Stochastic L-BFGS (Moritz et al. 2016) on the GPU — SLBFGS::stochastic_solve
// (src/minimizer/s_lbfgs.hpp:165-290) with the objective closures of UnifiedSLBFGS_CPU::optimize
// (src/unified_optimization.hpp:314-407: mean-squared loss / batch + 0.5*lambda*||w||^2).
//
// New functionality: the reference static_asserts on UnifiedSLBFGS<CudaBackend>
// (src/unified_optimization.hpp:639-641). The control flow and the RNG consumption order
// (mini-batch draws, H-batch draws, anchor pick from ONE std::mt19937(seed)) are the reference's;
// all index streams of an epoch are drawn on the host up front (they do not depend on device results)
// and uploaded once, so an epoch runs without a single host synchronisation:
//   per inner step: gather -> eval(w_t) -> eval(w~) -> v = g_t - g_k + mu -> dots -> solve -> apply
//   every L steps : iterate mean u, s = u - u_prev, w+- = u +- eps*s, 2 evaluations on the H-batch,
//                   y = (g+ - g-)/(2 eps), ring push with the |y.s| > 1e-10 filter decided on the device.
#include "lbfgs_kernels.cuh"
#include "network.cuh"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <random>
#include <vector>

namespace b200 {

namespace {

// Partial Fisher-Yates over the identity permutation, one uniform_int_distribution<size_t>(i, N-1) draw
// per element (s_lbfgs.hpp:141-161). The reference rebuilds iota(N) for every batch; undoing the swaps
// gives the same result in O(b).
struct Sampler {
  std::mt19937 rng;
  std::vector<size_t> idx;
  std::vector<std::pair<size_t, size_t>> undo;
  explicit Sampler(unsigned seed, size_t N) : rng(seed), idx(N) { std::iota(idx.begin(), idx.end(), 0); }
  void draw(size_t N, size_t b, uint32_t *out) {
    if (b >= N) {
      for (size_t i = 0; i < N; ++i) out[i] = (uint32_t)i;
      return;
    }
    undo.clear();
    for (size_t i = 0; i < b; ++i) {
      std::uniform_int_distribution<size_t> dist(i, N - 1);
      const size_t j = dist(rng);
      std::swap(idx[i], idx[j]);
      undo.emplace_back(i, j);
    }
    for (size_t i = 0; i < b; ++i) out[i] = (uint32_t)idx[i];
    for (size_t q = undo.size(); q-- > 0;) std::swap(idx[undo[q].first], idx[undo[q].second]);
  }
  size_t pick(size_t hi_inclusive) {
    std::uniform_int_distribution<size_t> d(0, hi_inclusive);
    return d(rng);
  }
};

// rows of X / T selected by idx into contiguous batch buffers (one warp per sample row)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ X, const float *__restrict__ T,
                                                          const uint32_t *__restrict__ idx, int count, int in_dim,
                                                          int out_dim, float *__restrict__ Xb, float *__restrict__ Tb,
                                                          float *__restrict__ Tb2 = nullptr) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh) // Tb2: rows [T | T] for the pair network
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < count; i += warps) {
    const size_t src = idx[i];
    const float *xs = X + src * in_dim;
    float *xd = Xb + (size_t)i * in_dim;
    if ((in_dim & 3) == 0 && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Xb)) & 15u) == 0) {
      for (int c = lane; c < in_dim / 4; c += 32)
        reinterpret_cast<float4 *>(xd)[c] = __ldg(reinterpret_cast<const float4 *>(xs) + c);
    } else {
      for (int c = lane; c < in_dim; c += 32) xd[c] = __ldg(xs + c);
    }
    for (int c = lane; c < out_dim; c += 32) {
      const float tv = __ldg(T + src * out_dim + c);
      Tb[(size_t)i * out_dim + c] = tv;
      if (Tb2) { Tb2[(size_t)i * 2 * out_dim + c] = tv; Tb2[(size_t)i * 2 * out_dim + out_dim + c] = tv; }
    }
  }
}

// the same rows of the quantised input's fp16 copy (block-major [nblocks][rows_src][64] halves) into a block-major copy of the
// mini-batch [nblocks][count][64]: one warp per (sample, block) pair moves one 128-byte run
__global__ void __launch_bounds__(256) gather16_rows_kernel(const uint32_t *__restrict__ src16, long rows_src, int nblocks,
                                                            const uint32_t *__restrict__ idx, int count, uint32_t *__restrict__ dst16) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  const int lane = threadIdx.x & 31;
  const long warps = ((long)gridDim.x * blockDim.x) >> 5;
  const long total = (long)count * nblocks;
  for (long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += warps) {
    const int b = (int)(w / count), i = (int)(w - (long)b * count);
    dst16[((long)b * count + i) * 32 + lane] = __ldg(src16 + ((long)b * rows_src + idx[i]) * 32 + lane);
  }
}
// gathers the fp16 rows next to the fp32 ones and registers them as the batch's fp16 view (no-op unless the input is quantised)
static int gather16(b200_net *net, const float *Xb, const uint32_t *d_idx, int count, cudaStream_t st) {
  long rows_src = 0;
  int nb = 0;
  const void *src = net_x16_source(net, &rows_src, &nb);
  void *dst = src ? net_gather16_buffer(net, count) : nullptr;
  if (!dst) return B200_OK;
  B200_LAUNCH(gather16_rows_kernel, std::min(1184, ceil_div((long)count * nb, 8)), 256, 0, st, (const uint32_t *)src, rows_src, nb, d_idx,
              count, (uint32_t *)dst);
  net_gather16_register(net, Xb, count);
  return B200_OK;
}

// v = g_t - g_k + mu   (s_lbfgs.hpp:225-228)
__global__ void __launch_bounds__(256) vr_combine_kernel(size_t n, const float *__restrict__ gt, const float *__restrict__ gk,
                                                         const float *__restrict__ mu, float *__restrict__ v) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    v[i] = (gt[i] - gk[i]) + mu[i];
}

// u = mean of `count` ring slots (s_lbfgs.hpp:238-242); s = u - u_prev; w+- = u +- eps*s (:92-93)
__global__ void __launch_bounds__(256) hvp_points_kernel(size_t n, size_t ld, const float *__restrict__ W, int count,
                                                         float *__restrict__ u, const float *__restrict__ u_prev,
                                                         int have_prev, float eps, float *__restrict__ s,
                                                         float *__restrict__ wp, float *__restrict__ wm) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.0f;
    for (int r = 0; r < count; ++r) acc += W[(size_t)r * ld + i]; // slot order == logical order only up to rotation: a sum
    const float uu = acc / (float)count;
    u[i] = uu;
    if (have_prev) {
      const float ss = uu - u_prev[i];
      s[i] = ss;
      wp[i] = fmaf(eps, ss, uu);
      wm[i] = fmaf(-eps, ss, uu);
    }
  }
}

// y = (g+ - g-) / (2 eps)   (s_lbfgs.hpp:100)
__global__ void __launch_bounds__(256) hvp_diff_kernel(size_t n, const float *__restrict__ gp, const float *__restrict__ gm,
                                                       float inv_2eps, float *__restrict__ y) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = (gp[i] - gm[i]) * inv_2eps;
}

// ---- the two evaluations of a step as ONE forward/backward ---------------------------------------------------------------
// Both evaluations of an inner step (w_t and the anchor) and both of a curvature pair (u + eps s, u - eps s) see the SAME
// mini-batch. They run as a single evaluation of the "pair network": the two parameter sets stacked along the output dimension,
//   layer 0: W = [W_a | W_b] (in x 2 out);   layer l >= 1: W = blockdiag(W_a, W_b) (2 in x 2 out);   b = [b_a | b_b];   T = [T | T]
// so X is read once, every kernel is launched once, and the loss is L_a + L_b: its gradient holds g_a and g_b in the diagonal
// blocks (the off-diagonal blocks belong to weights that are identically zero and are ignored). Exact: the zero blocks add
// exact zeros to every sum. pack_pair_kernel builds the stacked parameters, unpack_pair_kernel reads the two gradients back
// and forms the quantity the algorithm wants in the same pass: v = g_a - g_b + mu (s_lbfgs.hpp:225-228) or
// y = (g_a - g_b) / (2 eps) (:100).
struct PairLayout {
  int nl;
  int in[16], out[16];               // of ONE network
  unsigned long long off[16], poff[16]; // parameter offsets of layer l in one network / in the pair network
};
__global__ void __launch_bounds__(256) pack_pair_kernel(const PairLayout L, unsigned long long n_pair, const float *__restrict__ wa,
                                                        const float *__restrict__ wb, float *__restrict__ wp) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pair;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    int l = 0;
    while (l + 1 < L.nl && i >= L.poff[l + 1]) ++l;
    const unsigned long long e = i - L.poff[l];
    const int K = L.in[l], N = L.out[l], Kp = (l == 0) ? K : 2 * K, Np = 2 * N;
    float v = 0.0f;
    if (e < (unsigned long long)Kp * Np) {
      const int kp = (int)(e / Np), op = (int)(e - (unsigned long long)kp * Np);
      const bool second = op >= N;
      const int o = second ? op - N : op;
      if (l == 0) v = (second ? wb : wa)[L.off[l] + (unsigned long long)kp * N + o];
      else if ((kp >= K) == second) v = (second ? wb : wa)[L.off[l] + (unsigned long long)(second ? kp - K : kp) * N + o];
    } else {
      const int op = (int)(e - (unsigned long long)Kp * Np);
      v = (op >= N ? wb : wa)[L.off[l] + (unsigned long long)K * N + (op >= N ? op - N : op)];
    }
    wp[i] = v;
  }
}
// out[i] = (g_a[i] - g_b[i]) * scale + (add ? add[i] : 0)
__global__ void __launch_bounds__(256) unpack_pair_kernel(const PairLayout L, unsigned long long n, const float *__restrict__ gp,
                                                          float scale, const float *__restrict__ add, float *__restrict__ out) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    int l = 0;
    while (l + 1 < L.nl && i >= L.off[l + 1]) ++l;
    const unsigned long long e = i - L.off[l];
    const int K = L.in[l], N = L.out[l], Kp = (l == 0) ? K : 2 * K, Np = 2 * N;
    unsigned long long ia, ib;
    if (e < (unsigned long long)K * N) {
      const int k = (int)(e / N), o = (int)(e - (unsigned long long)k * N);
      ia = L.poff[l] + (unsigned long long)k * Np + o;
      ib = L.poff[l] + (unsigned long long)(l == 0 ? k : K + k) * Np + N + o;
    } else {
      const int o = (int)(e - (unsigned long long)K * N);
      ia = L.poff[l] + (unsigned long long)Kp * Np + o;
      ib = ia + N;
    }
    const float d = (gp[ia] - gp[ib]) * scale;
    out[i] = add ? d + add[i] : d;
  }
}

inline int vblocks(size_t n) { return (int)std::max<size_t>(1, std::min<size_t>(1184, (n + 255) / 256)); }

} // namespace

// StochasticGradientDescent::stochastic_solve (src/minimizer/s_gd.hpp:63-145) with the objective closures of UnifiedSGD_CPU
// (src/unified_optimization.hpp:219-300): per epoch m = N / b mini-batches of b random indices (the sampler above: the same
// draws as s_gd.hpp:148-170) from ONE mt19937(seed); w -= step * (mean gradient of the batch). No L2 term, no decay, no stop
// test. All index lists of an epoch are drawn up front and uploaded once: an epoch runs without a host synchronisation.
// Recorder (:122-139): full-batch loss and full-gradient norm after every epoch.
int sgd_random_solve(b200_ctx *ctx, b200_net *net, int n, float *params, const float *input, const float *target, int total_samples,
                     const b200_sgd_opts &o, b200_history *hist) {
  B200_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const long launches0 = b200_launch_count();
  const int N = total_samples, b = std::min(o.batch_size, N);
  int m = N / o.batch_size; // unified_optimization.hpp:232-233
  if (m == 0) m = 1;
  const int in_dim = net->dims.front(), out_dim = net->dims.back();
  const size_t Nn = (size_t)n;
  const size_t vec = (sizeof(float) * Nn + 255) & ~size_t(255);
  const size_t xb_bytes = (sizeof(float) * (size_t)b * in_dim + 255) & ~size_t(255);
  const size_t tb_bytes = (sizeof(float) * (size_t)b * out_dim + 255) & ~size_t(255);
  const size_t idx_bytes = (sizeof(uint32_t) * (size_t)m * b + 255) & ~size_t(255);
  char *ws = nullptr;
  B200_CUDA(cudaMalloc(&ws, 2 * vec + xb_bytes + tb_bytes + idx_bytes + 256));
  struct Free { char *p; ~Free() { cudaFree(p); } } free_ws{ws};
  float *grad = (float *)ws, *vel = (float *)(ws + vec);
  float *Xb = (float *)(ws + 2 * vec), *Tb = (float *)(ws + 2 * vec + xb_bytes);
  uint32_t *d_idx = (uint32_t *)(ws + 2 * vec + xb_bytes + tb_bytes);
  EvalOut *scratch = (EvalOut *)(ws + 2 * vec + xb_bytes + tb_bytes + idx_bytes);
  if (o.momentum > 0.0f) B200_CUDA(cudaMemsetAsync(vel, 0, sizeof(float) * Nn, st));
  uint32_t *h_idx = nullptr;
  B200_CUDA(cudaMallocHost(&h_idx, sizeof(uint32_t) * (size_t)m * b));
  struct FreeHost { uint32_t *p; ~FreeHost() { cudaFreeHost(p); } } free_h{h_idx};
  Sampler sampler(o.seed, (size_t)N);
  double *h_mail = ctx->h_scalars;
  long evals = 0;
  int iters = 0;
  float elapsed = 0.f;
  cudaEvent_t ev0 = ctx->ev_a, ev1 = ctx->ev_b;
  // 8-bit-pixel inputs: the fp16 copy of the data set (and of every gathered mini-batch) feeds the fp16 layer-0 kernels
  if (net->prec != B200_PREC_FP32) B200_TRY(net_quantize_input(net, input, N, true));
  struct ClearQ { b200_net *n; ~ClearQ() { net_xq_release_solver(n); } } clear_q{net};
  while (iters < o.max_iters) {
    if (hist) B200_CUDA(cudaEventRecord(ev0, st));
    for (int t = 0; t < m; ++t) sampler.draw((size_t)N, (size_t)b, h_idx + (size_t)t * b);
    B200_CUDA(cudaMemcpyAsync(d_idx, h_idx, sizeof(uint32_t) * (size_t)m * b, cudaMemcpyHostToDevice, st));
    for (int t = 0; t < m; ++t) {
      B200_LAUNCH(gather_rows_kernel, std::min(1184, ceil_div(b, 8)), 256, 0, st, input, target, d_idx + (size_t)t * b, b, in_dim,
                  out_dim, Xb, Tb, (float *)nullptr);
      B200_TRY(gather16(net, Xb, d_idx + (size_t)t * b, b, st));
      ++evals;
      B200_TRY(net_eval(net, params, Xb, Tb, b, b, grad, scratch)); // grad /= current_bs (:266)
      if (o.momentum > 0.0f) B200_TRY(launch_momentum_step(Nn, o.momentum, o.lr, grad, vel, params, st));
      else B200_TRY(launch_axpy(Nn, -o.lr, grad, params, st));       // w = w - step * grad_est (s_gd.hpp:104)
    }
    if (hist) {
      ++evals;
      B200_TRY(net_eval(net, params, input, target, N, N, grad, (EvalOut *)net->eval_out));
      B200_CUDA(cudaMemcpyAsync(h_mail, net->eval_out, sizeof(EvalOut), cudaMemcpyDeviceToHost, st));
      B200_CUDA(cudaEventRecord(ev1, st));
      B200_CUDA(cudaEventSynchronize(ev1));
      float ms = 0.f;
      B200_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
      elapsed += ms;
      if (iters < hist->capacity) {
        if (hist->loss) hist->loss[iters] = (float)h_mail[0];
        if (hist->grad_norm) hist->grad_norm[iters] = (float)std::sqrt(h_mail[1]);
        if (hist->time_ms) hist->time_ms[iters] = elapsed;
        hist->size = iters + 1;
      }
    }
    ++iters;
  }
  B200_CUDA(cudaStreamSynchronize(st));
  if (hist) {
    hist->iterations = iters;
    hist->evaluations = evals;
    hist->launches = b200_launch_count() - launches0;
  }
  return B200_OK;
}

} // namespace b200

using namespace b200;

extern "C" {

void b200_slbfgs_default_opts(b200_slbfgs_opts *o) {
  if (!o) return;
  o->max_iters = 100; o->tol = 1e-4f; o->step_size = 0.01f; o->batch_size = 128; // UnifiedConfig defaults, unified_optimization.hpp:26-48
  o->memory = 10; o->L = 10; o->b_H = 0; o->lambda = 1e-4f; o->epsilon = 1e-4f; o->seed = 123; o->record = 1;
  o->hvp_step_scale = 256.0f; o->pair_eval = 1;
}

int b200_slbfgs_sample_stream(unsigned seed, long N, long b, int count, uint32_t *out_host) {
  B200_REQUIRE(out_host && N > 0 && b > 0 && count >= 0, "bad argument");
  Sampler s(seed, (size_t)N);
  const size_t per = (size_t)std::min(b, N);
  for (int c = 0; c < count; ++c) s.draw((size_t)N, (size_t)b, out_host + (size_t)c * per);
  return B200_OK;
}

int b200_slbfgs_solve(b200_ctx *ctx, b200_net *net, int n, float *params, const float *input, const float *target,
                      int total_samples, const b200_slbfgs_opts *opts, b200_history *hist) {
  B200_REQUIRE(ctx && net, "null argument");
  if (hist) { hist->size = 0; hist->iterations = 0; hist->evaluations = 0; hist->launches = 0; }
  if (n <= 0 || params == nullptr) return B200_OK;
  B200_REQUIRE((size_t)n == net->n && input && target && total_samples > 0, "bad argument");
  b200_slbfgs_opts o;
  if (opts) o = *opts; else b200_slbfgs_default_opts(&o);
  B200_REQUIRE(o.batch_size > 0 && o.L > 0 && o.memory >= 0 && o.memory <= kMaxSlots - 1, "bad S-LBFGS options");
  const float fd_eps = o.epsilon * (o.hvp_step_scale > 0.0f ? o.hvp_step_scale : 1.0f); // see b200_slbfgs_opts::hvp_step_scale
  B200_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const long launches0 = b200_launch_count();

  const int N = total_samples, W = ctx->world, R = ctx->rank;
  const int b = std::min(o.batch_size, N);
  int b_H = o.b_H > 0 ? o.b_H : o.batch_size / 2; // unified_optimization.hpp:325
  b_H = std::max(1, std::min(b_H, N));
  int m_inner = N / o.batch_size;                 // :326-327
  if (m_inner == 0) m_inner = 1;
  const int M = o.memory, L = o.L, mp = M + 1, mod = M + 1;
  const int in_dim = net->dims.front(), out_dim = net->dims.back();
  const size_t Nn = (size_t)n, ld = (Nn + 3) & ~size_t(3);
  // multi-GPU: every rank holds the full data set; each batch's index list is split into W contiguous chunks
  B200_REQUIRE(b % W == 0 && b_H % W == 0 && N % W == 0, "batch sizes must be divisible by the number of ranks");

  // 8-bit-pixel inputs: the fp16 copy of the data set (every rank holds all of it) feeds the full-batch evaluations through its
  // row slices and the mini-batches through gathered copies (gather16)
  if (net->prec != B200_PREC_FP32) B200_TRY(net_quantize_input(net, input, total_samples, true));
  struct ClearQ { b200_net *n; ~ClearQ() { net_xq_release_solver(n); } } clear_q{net};
  const float old_l2 = net->l2;
  net->l2 = o.lambda;
  struct RestoreL2 { b200_net *n; float v; ~RestoreL2() { n->l2 = v; } } restore{net, old_l2};

  // the pair network (see pack_pair_kernel): same activations, every width but the input doubled
  const bool use_pair = o.pair_eval != 0 && net->nlayers() <= 16;
  b200_net *pair = nullptr;
  struct FreePair { b200_net **p; ~FreePair() { if (*p) b200_net_destroy(*p); } } free_pair{&pair};
  PairLayout PL{};
  size_t n_pair = 0;
  if (use_pair) {
    std::vector<int> pd(net->dims), pa(net->acts);
    for (size_t l = 1; l < pd.size(); ++l) pd[l] *= 2;
    B200_TRY(b200_net_create(ctx, net->nlayers(), pd.data(), pa.data(), &pair));
    pair->prec = net->prec;
    pair->l2 = o.lambda;
    PL.nl = net->nlayers();
    for (int l = 0; l < PL.nl; ++l) {
      PL.in[l] = net->dims[l]; PL.out[l] = net->dims[l + 1];
      PL.off[l] = net->offs[l]; PL.poff[l] = pair->offs[l];
    }
    n_pair = pair->n;
  }

  // ---- workspace ---------------------------------------------------------------------------------
  const int nblk = lbfgs_dots_blocks(ctx, Nn);
  const int ncols = kDotsCols * mp + 1;
  const size_t state_bytes = lbfgs_state_bytes(M);
  const size_t part_bytes = (sizeof(double) * (size_t)nblk * ncols + 255) & ~size_t(255);
  const size_t vec = sizeof(float) * ld;
  const int nvec = 13 + (L + 1) + 2 * mp; // wt mu gt gk v d u uprev s y wp wm gfull | w_history | S Y
  const size_t pvec = (sizeof(float) * n_pair + 255) & ~size_t(255);          // pair network: parameters, gradient
  const size_t tb2_bytes = use_pair ? 2 * ((sizeof(float) * (size_t)(std::max(b, b_H) / W) * out_dim + 255) & ~size_t(255)) : 0;
  const int maxb = std::max(b, b_H) / W;
  const size_t xb_bytes = (sizeof(float) * (size_t)maxb * in_dim + 255) & ~size_t(255);
  const size_t tb_bytes = (sizeof(float) * (size_t)maxb * out_dim + 255) & ~size_t(255);
  // index streams of one epoch: m_inner mini-batches + at most m_inner/L H-batches
  const size_t idx_cap = (size_t)m_inner * (b / W) + (size_t)(m_inner / L + 1) * (b_H / W);
  const size_t idx_bytes = (sizeof(uint32_t) * idx_cap + 255) & ~size_t(255);
  char *ws = nullptr;
  const size_t total = state_bytes + part_bytes + vec * nvec + xb_bytes + tb_bytes + idx_bytes + 2 * sizeof(EvalOut) + 256 +
                       2 * pvec + tb2_bytes;
  B200_CUDA(cudaMalloc(&ws, total));
  struct Free { char *p; ~Free() { cudaFree(p); } } free_ws{ws};
  B200_CUDA(cudaMemsetAsync(ws, 0, total, st));
  size_t off = 0;
  LbfgsView view = lbfgs_view(ws + off, M); off += state_bytes;
  double *partials = (double *)(ws + off); off += part_bytes;
  auto take = [&](size_t bytes) { char *p = ws + off; off += bytes; return p; };
  float *wt = (float *)take(vec), *mu = (float *)take(vec), *gt = (float *)take(vec), *gk = (float *)take(vec);
  float *v = (float *)take(vec), *d = (float *)take(vec), *u = (float *)take(vec), *u_prev = (float *)take(vec);
  float *s = (float *)take(vec), *y = (float *)take(vec), *wp = (float *)take(vec), *wm = (float *)take(vec);
  float *gfull = (float *)take(vec);
  float *Wh = (float *)take(vec * (L + 1));
  float *S = (float *)take(vec * mp), *Y = (float *)take(vec * mp);
  float *Xb = (float *)take(xb_bytes), *Tb = (float *)take(tb_bytes);
  uint32_t *d_idx = (uint32_t *)take(idx_bytes);
  EvalOut *ev_scratch = (EvalOut *)take(128);
  unsigned *gridbar = (unsigned *)take(128); // {arrivals, generation} of the fused direction kernel's grid barrier (zeroed with ws)
  float *pw = use_pair ? (float *)take(pvec) : nullptr, *pg = use_pair ? (float *)take(pvec) : nullptr;
  float *Tb2 = use_pair ? (float *)take(tb2_bytes) : nullptr;
  // one forward/backward of the pair network at (wa, wb) on the gathered batch; out = (g_a - g_b) * scale (+ add)
  auto pair_eval = [&](const float *wa, const float *wb, int rows, int rows_global, float scale, const float *add, float *out) -> int {
    B200_LAUNCH(pack_pair_kernel, vblocks(n_pair), 256, 0, st, PL, (unsigned long long)n_pair, wa, wb, pw);
    B200_TRY(net_eval(pair, pw, Xb, Tb2, rows, rows_global, pg, ev_scratch));
    B200_LAUNCH(unpack_pair_kernel, vblocks(Nn), 256, 0, st, PL, (unsigned long long)Nn, pg, scale, add, out);
    return B200_OK;
  };
  B200_TRY(lbfgs_init_state(view, M, mod, st));
  uint32_t *h_idx = nullptr;
  B200_CUDA(cudaMallocHost(&h_idx, sizeof(uint32_t) * idx_cap));
  struct FreeHost { uint32_t *p; ~FreeHost() { cudaFreeHost(p); } } free_h{h_idx};

  Sampler sampler(o.seed, (size_t)N);
  std::vector<uint32_t> draw_buf((size_t)std::max(b, b_H));
  double *h_mail = ctx->h_scalars;
  long evals = 0;
  const int apply_blocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)4 * ctx->num_sms, (ld / 4 + 255) / 256));
  const int shard_N = N / W;
  const float *x_shard = input + (size_t)R * shard_N * in_dim, *t_shard = target + (size_t)R * shard_N * out_dim;

  auto full_eval = [&](const float *w, float *g) -> int { // batch_g on the full index set + norm (:204-210)
    ++evals;
    B200_TRY(net_eval(net, w, x_shard, t_shard, shard_N, N, g, (EvalOut *)net->eval_out));
    B200_CUDA(cudaMemcpyAsync(h_mail, net->eval_out, sizeof(EvalOut), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return ctx_check_device_error(ctx);
  };

  bool have_u_prev = false;   // !u_list.empty()
  bool mu_valid = false;      // gfull holds the full gradient at `params` (from the recorder evaluation)
  int iters = 0;
  cudaEvent_t ev0 = ctx->ev_a, ev1 = ctx->ev_b;
  float elapsed = 0.f;
  while (iters < o.max_iters) {
    if (hist) B200_CUDA(cudaEventRecord(ev0, st));
    // 1. full gradient at the anchor (:204-206)
    if (mu_valid) {
      B200_CUDA(cudaMemcpyAsync(mu, gfull, vec, cudaMemcpyDeviceToDevice, st));
    } else {
      B200_TRY(full_eval(params, mu));
    }
    const double mu_norm = std::sqrt(h_mail[1]);
    if (mu_norm < (double)o.tol) break; // :208-211
    B200_CUDA(cudaMemcpyAsync(wt, params, sizeof(float) * Nn, cudaMemcpyDeviceToDevice, st));
    // w_history.clear(); push_back(wt)  — physical slot = push count % (L+1)
    int wh_pushes = 0;
    B200_CUDA(cudaMemcpyAsync(Wh, params, sizeof(float) * Nn, cudaMemcpyDeviceToDevice, st));
    wh_pushes = 1;

    // ---- draw every index stream of this epoch in the reference's order (:220, :248) ---------------
    size_t idx_used = 0;
    std::vector<size_t> mb_off(m_inner), hb_off(m_inner, (size_t)-1);
    bool u_nonempty = have_u_prev;
    for (int t = 0; t < m_inner; ++t) {
      sampler.draw((size_t)N, (size_t)b, draw_buf.data());
      mb_off[t] = idx_used;
      std::copy(draw_buf.begin() + (size_t)R * (b / W), draw_buf.begin() + (size_t)(R + 1) * (b / W), h_idx + idx_used);
      idx_used += b / W;
      if (t > 0 && t % L == 0) {
        if (u_nonempty) {
          sampler.draw((size_t)N, (size_t)b_H, draw_buf.data());
          hb_off[t] = idx_used;
          std::copy(draw_buf.begin() + (size_t)R * (b_H / W), draw_buf.begin() + (size_t)(R + 1) * (b_H / W),
                    h_idx + idx_used);
          idx_used += b_H / W;
        }
        u_nonempty = (M > 0); // u_list has capacity M+1, or 0 when M == 0 (push_back is then a no-op, :176)
      }
    }
    B200_CUDA(cudaMemcpyAsync(d_idx, h_idx, sizeof(uint32_t) * idx_used, cudaMemcpyHostToDevice, st));

    // 2. inner loop (:216-262)
    for (int t = 0; t < m_inner; ++t) {
      const int bl = b / W;
      B200_LAUNCH(gather_rows_kernel, std::min(1184, ceil_div(bl, 8)), 256, 0, st, input, target, d_idx + mb_off[t], bl,
                  in_dim, out_dim, Xb, Tb, Tb2);
      evals += 2;
      if (use_pair) { // both gradients from ONE forward/backward, v = g_t - g_k + mu formed as they are read back
        B200_TRY(pair_eval(wt, params, bl, b, 1.0f, mu, v));
      } else {
        B200_TRY(gather16(net, Xb, d_idx + mb_off[t], bl, st));
        B200_TRY(net_eval(net, wt, Xb, Tb, bl, b, gt, ev_scratch));
        B200_TRY(net_eval(net, params, Xb, Tb, bl, b, gk, ev_scratch));
        B200_LAUNCH(vr_combine_kernel, vblocks(Nn), 256, 0, st, Nn, gt, gk, mu, v);
      }
      // direction = H v (two-loop on the current ring), w_t -= eta * direction, history push (:230-233)
      DotsArgs da{S, Y, Nn, ld, view, v, nullptr, nullptr, nullptr, DOTS_NONE, 0, 0, partials};
      SolveArgs sa{view, partials, nblk, DOTS_NONE, 0, POLICY_SLBFGS, 0, 0, 0.0, 0};
      float *wh_slot = Wh + (size_t)(wh_pushes % (L + 1)) * ld;
      ApplyArgs aa{S, Y, Nn, ld, view, v, d, wt, nullptr, -1.0, -o.step_size, wh_slot};
      bool fused_dir = false; // dots -> grid barrier -> solve -> apply in one launch (histories of <= 32 slots)
      B200_TRY(launch_lbfgs_direction(ctx, da, sa, aa, mp, nblk, gridbar, st, &fused_dir, nullptr, 0));
      if (!fused_dir) {
        B200_TRY(launch_lbfgs_dots(da, mp, nblk, st));
        B200_TRY(launch_lbfgs_solve(sa, mp, st));
        B200_TRY(launch_lbfgs_apply(aa, apply_blocks, st));
      }
      ++wh_pushes;

      // 3. curvature pair every L steps (:236-261)
      if (t > 0 && t % L == 0) {
        const int cnt = std::min(wh_pushes, L + 1);
        B200_LAUNCH(hvp_points_kernel, vblocks(Nn), 256, 0, st, Nn, ld, Wh, cnt, u, u_prev, have_u_prev ? 1 : 0,
                    fd_eps, s, wp, wm);
        if (have_u_prev) {
          const int hl = b_H / W;
          B200_LAUNCH(gather_rows_kernel, std::min(1184, ceil_div(hl, 8)), 256, 0, st, input, target, d_idx + hb_off[t],
                      hl, in_dim, out_dim, Xb, Tb, Tb2);
          evals += 2;
          if (use_pair) { // the +- eps s pair in one forward/backward, y = (g+ - g-) / (2 eps) formed as they are read back
            B200_TRY(pair_eval(wp, wm, hl, b_H, 1.0f / (2.0f * fd_eps), nullptr, y));
          } else {
            B200_TRY(gather16(net, Xb, d_idx + hb_off[t], hl, st));
            B200_TRY(net_eval(net, wp, Xb, Tb, hl, b_H, gt, ev_scratch));
            B200_TRY(net_eval(net, wm, Xb, Tb, hl, b_H, gk, ev_scratch));
            B200_LAUNCH(hvp_diff_kernel, vblocks(Nn), 256, 0, st, Nn, gt, gk, 1.0f / (2.0f * fd_eps), y);
          }
          if (M > 0) {
            B200_TRY(launch_lbfgs_store_pair(S, Y, Nn, ld, view, s, y, st));
            DotsArgs dp{S, Y, Nn, ld, view, v, nullptr, nullptr, nullptr, DOTS_PAIR_IN_SLOT, 0, 0, partials};
            B200_TRY(launch_lbfgs_dots(dp, mp, nblk, st));
            SolveArgs sp{view, partials, nblk, DOTS_PAIR_IN_SLOT, 0, POLICY_SLBFGS, 0, 0, 0.0, 0};
            B200_TRY(launch_lbfgs_solve(sp, mp, st));
          }
        }
        B200_CUDA(cudaMemcpyAsync(u_prev, u, sizeof(float) * Nn, cudaMemcpyDeviceToDevice, st)); // u_list.push_back(u)
        have_u_prev = (M > 0);
      }
    }

    // anchor reset to a random earlier iterate (:265-270)
    const int wsize = std::min(wh_pushes, L + 1);
    if (wsize >= 2) {
      const size_t pick = sampler.pick((size_t)wsize - 2);
      const int head = (wh_pushes > L + 1) ? (wh_pushes % (L + 1)) : 0; // RingBuffer _head after wrap
      const int phys = (int)((head + pick) % (size_t)(L + 1));
      B200_CUDA(cudaMemcpyAsync(params, Wh + (size_t)phys * ld, sizeof(float) * Nn, cudaMemcpyDeviceToDevice, st));
    } else {
      B200_CUDA(cudaMemcpyAsync(params, wt, sizeof(float) * Nn, cudaMemcpyDeviceToDevice, st));
    }

    // recorder: full loss and full gradient norm at the new anchor (:274-284); the gradient doubles as the
    // next epoch's mu (same point), saving the reference's redundant evaluation
    if (hist && o.record) {
      B200_TRY(full_eval(params, gfull));
      mu_valid = true;
      B200_CUDA(cudaEventRecord(ev1, st));
      B200_CUDA(cudaEventSynchronize(ev1));
      float ms = 0.f;
      B200_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
      elapsed += ms;
      if (iters < hist->capacity) {
        if (hist->loss) hist->loss[iters] = (float)h_mail[0];
        if (hist->grad_norm) hist->grad_norm[iters] = (float)std::sqrt(h_mail[1]);
        if (hist->time_ms) hist->time_ms[iters] = elapsed;
        hist->size = iters + 1;
      }
    } else {
      mu_valid = false;
    }
    ++iters;
  }
  B200_CUDA(cudaStreamSynchronize(st));
  if (hist) {
    hist->iterations = iters;
    hist->evaluations = evals;
    hist->launches = b200_launch_count() - launches0;
  }
  return B200_OK;
}

} // extern "C"

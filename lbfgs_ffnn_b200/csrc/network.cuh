// b200_net: the MLP objective (replaces cuda_mlp::CudaNetwork, src/cuda/network.cuh).
#pragma once

#include "common.cuh"

#include <vector>

namespace b200 { struct SpecState; }

struct b200_net {
  b200_ctx *ctx = nullptr;
  unsigned long long uid = 0; // unique per created network (never reused, unlike the address): the solvers key captured graphs on it
  std::vector<int> dims; // nlayers + 1
  std::vector<int> acts; // nlayers
  std::vector<size_t> offs;
  size_t n = 0;
  int prec = B200_PREC_FP32;
  float l2 = 0.0f;
  long batch_global = 0; // 0: shard batch x world
  bool defer_reduce = false; // multi-GPU: leave the gradient / loss as this rank's partial (the caller reduce-scatters)

  // flat parameter / gradient buffers (owned after bind_params)
  float *params = nullptr, *grads = nullptr;

  // per-batch state
  long cap = 0;        // sample capacity of act/delta
  long last_batch = 0;
  std::vector<float *> act, delta;
  std::vector<int> ldd; // row stride of delta[l]: out rounded up to 4 floats so that TMA (16-byte strides) can address it

  // split-K partials of [dW; db] per layer and their layout
  std::vector<int> splits, k_chunk; // FFMA split-K plan
  std::vector<int> skinny_splits;   // split plan of the skinny (out <= 16) dW kernel
  std::vector<int> splits_used;     // splits written by the last evaluation (either path)
  int dw0_tail_row0 = -1, dw0_tail_splits = 0; // fp16 layer-0 dW (gemm_dw16.cu): rows from tail_row0 on have tail_splits slices
  std::vector<size_t> part_off;
  float *partials = nullptr;
  size_t partials_cap = 0;
  long partials_batch = -1;

  // uint8 copy of an input that is exactly u/255 (MNIST-style pixels, tests/mnist/mnist_loader.hpp:59): the two
  // X-bound GEMMs of layer 0 then read 4x fewer bytes and need no hi/lo split for X (u is exact in TF32)
  struct InputQ {
    const float *src = nullptr;
    long rows = 0;
    uint8_t *data = nullptr;
    // fp16 copy (value u), FEATURE-BLOCK-major: [nblocks16][rows][64 halves]; block b holds features 64 b .. 64 b + 63 of every
    // sample, feature `in` reads 1 (the bias-gradient row of dW), the rest of the last block 0. Every operand tile of the fp16 layer-0
    // kernels ([128 samples][64 features] forward, [32 samples][64 features] dW) is then ONE contiguous chunk of DRAM.
    void *data16 = nullptr;
    int nblocks16 = 0;
    size_t cap = 0;
    bool valid = false;
    bool user = false;   // built by the caller through b200_net_quantize_input (kept until b200_net_clear_input_cache); a copy the
                         // solvers made themselves lives only as long as the minimisation that made it
    int *flag = nullptr; // device
  } xq;

  // a mini-batch GATHERED from the quantised input (S-LBFGS, random-batch SGD): rows idx[i] of the fp16 copy above, collected
  // into the same block-major layout, next to the fp32 rows the caller passes to net_eval (slbfgs.cu: gather16_rows_kernel).
  // While registered, evaluations of exactly (src, rows) run the fp16 layer-0 kernels like a slice of the input would.
  struct GatherQ {
    const float *src = nullptr; // the gathered fp32 rows (identity of the view)
    long rows = 0;
    void *data16 = nullptr;     // [nblocks16][cap][64] halves, used with rows_total = rows
    long cap = 0;
  } xg;

  float *w_hi = nullptr, *w_lo = nullptr; // 3xTF32: hi / lo split of the parameter vector of the current evaluation
  const float *split_src = nullptr;       // parameters of the current evaluation whose TF32 split has not been launched yet (lazy)
  // fp16 forward of layer 0 on a uint8 input (gemm_fwd16.cu): per-neuron-scaled hi / lo fp16 weights [out][ldk], 1/(255 s_o)
  void *w16h = nullptr, *w16l = nullptr;
  float *colscale = nullptr;
  const float *w16_params = nullptr; // parameter vector the fp16 split currently holds (per evaluation)
  // one-pass last layer (tail_layer.cu): per-CTA max |delta_L|, scaled fp16 {hi | lo} copy of delta_{L-1} and 1 / its scale
  float *amax_part = nullptr, *scale16_inv = nullptr, *scale16 = nullptr, *chain_cw = nullptr; // scale16: the scale itself, chain_cw: see ChainW
  bool chain_ready = false; // chain_cw holds the factor of the current evaluation's parameters
  void *delta16 = nullptr;
  long delta16_cap = 0;

  // ---- hidden layer 1 of a three-layer net on the fp16 tensor-core kernels ("mid16", gemm_fwd16.cu / gemm_dw16.cu) -------------
  // Layer 0's epilogue leaves A_1 ONLY as a per-feature-scaled fp16 pair, block-major like the input copy:
  //   a16 = [2 nb][rows][64] halves, blocks 0 .. nb-1 = hi(t_f a), nb .. 2 nb-1 = lo, t_f a power of two with t_f * bound_f in
  //   [2^13, 2^14) and bound_f = sum_k |W_0[k][f]| + |b_f| >= |a| (x in [0, 1]; 1 for tanh / sigmoid);
  // layer 1 forward is the layer-0 kernel again on that copy (K = 2 * width: the lo blocks meet the same weights), its dW is the
  // split-K kernel with the hi and lo feature tiles folded in the epilogue, its dX the forward kernel on delta_1's fp16 pair
  // with an act'(A_1) epilogue that emits delta_0's pair for layer 0's dW.
  struct Mid16 {
    bool on = false;        // this evaluation runs the path (decided per evaluation in net_eval / net_forward)
    void *a16 = nullptr;    // A_1 pair
    long a16_rows = 0;
    float *tscale = nullptr, *tinv = nullptr; // [dims[1]]: t_f and 1 / t_f
    unsigned *prep_sync = nullptr;            // two counters of prep_w16_kernel (publishers done, CTAs done)
    void *wfh = nullptr, *wfl = nullptr;      // forward operand of layer 1: [dims[2]][2 dims[1]] hi / lo of s_o W_1[k][o] / t_k
    float *colscale_f = nullptr;              // [dims[2]]: 1 / s_o
    void *wdh = nullptr, *wdl = nullptr;      // dX operand of layer 1: [dims[1]][2 dims[2]] hi / lo of s_k W_1[k][o]
    float *colscale_d = nullptr;              // [dims[1]]: 1 / s_k
    void *d16 = nullptr;    // delta_1 pair [rows][2 dims[2]] (written by the last-layer backward kernel)
    long d16_rows = 0;
    float *scale1_inv = nullptr; // device scalar: 1 / (scale of the delta_1 pair)
    float *db_part = nullptr;    // [tail grid][dims[2]]: per-CTA column sums of delta_1 (= the db_1 partials)
    int db_splits = 0;
    bool act0_stale = false;     // act[0] (fp32) does not hold A_1 of the last evaluation: only a16 does
  } m16;

  // ---- wide hidden layers (hundreds to thousands of columns: BASELINE configs[4]) on the fp16 pair kernels ("wide16", end of
  // gemm_fwd16.cu). Activations, deltas and weights stay fp32 in HBM as on the generic path; each GEMM operand is split once per
  // evaluation into a pair hi = fp16(S v), lo = fp16(S v - hi) with ONE power-of-two scale S per matrix (S max|v| in
  // [2^13, 2^14)), row-major [rows][hi Cp | lo Cp] and, for the dW of the layer, transposed [cols (+ a row of ones)][hi Rp | lo Rp].
  struct Wide16 {
    struct Buf { void *p = nullptr; size_t halves = 0; };
    std::vector<Buf> a, aT;   // per layer l: its input A_{l-1} (forward A operand) and the transpose + ones row (dW A operand)
    std::vector<Buf> wf, wd;  // per layer l: W_l as [out][hi Kp | lo Kp] (forward B operand) and [in][hi Np | lo Np] (dX B operand)
    Buf d, dT;                // delta_l of the layer being worked on: [rows][hi Np | lo Np] (dX A operand), [out][hi Bp | lo Bp] (dW B operand)
    std::vector<char> a_ready, w_ready, d_ready; // this evaluation has made them
    float *scal = nullptr;    // device: per layer {S_a, 1/S_a, S_w, 1/S_w, S_d, 1/S_d, -, -}
    float *amax_part = nullptr; // device: per-CTA maxima of the matrix being split
    int amax_n = 0, scal_layers = 0;
    // an input that net_quantize_input found NOT to be 8-bit pixels: the caller (a solver for the length of its run, or the user
    // through b200_net_quantize_input) holds x[x_rows] constant, so layer 0's operand split is made once, not per evaluation
    const float *x_src = nullptr;
    long x_rows = 0;
    bool x_done = false;
  } w16x;

  double *loss_part = nullptr; // per-CTA partials of sum diff^2
  int loss_part_cap = 0, loss_part_n = 0;
  double *fin_part = nullptr;  // per-CTA partials of ||g||^2 and ||w||^2 (2 per CTA)
  int fin_blocks = 0;
  double *eval_out = nullptr;  // device {loss, gnorm2}
  unsigned *fin_done = nullptr; // CTAs of finalize_grad_kernel that have finished (the last one reduces the scalars)

  // bumped whenever a device buffer of this net is (re)allocated or a setting that selects kernels changes: captured CUDA
  // graphs bake both in, so the solvers key their graphs on it
  long config_gen = 0;
  // set by the L-BFGS solver around a graph capture: gate of speculatively launched evaluations (common.cuh)
  b200::SpecState *spec_st = nullptr;
  int spec_flag = 0;
  double *host_out = nullptr; // pinned-host {loss, gnorm2}: written by the kernel that finishes the evaluation (zero-copy mailbox)

  int nlayers() const { return (int)acts.size(); }
};

namespace b200 {

struct EvalOut { // device-resident result of one evaluation
  double loss;
  double gnorm2;
};

// Evaluate loss and gradient at `params` (device, flat layout) on the shard (x, t, batch).
// grad_out (device, n floats) receives the gradient (all-reduced when a communicator exists);
// out (device EvalOut) receives loss and ||grad||^2. Asynchronous on ctx->stream.
// batch_global is the 1/B denominator (sum of shard sizes over ranks).
int net_eval(b200_net *net, const float *params, const float *x, const float *t, long batch, long batch_global,
             float *grad_out, EvalOut *out);
int net_forward(b200_net *net, const float *params, const float *x, long batch);
int net_ensure(b200_net *net, long batch);
// build (or reuse) the uint8 copy of x[batch][in]; no-op unless every element is exactly float(u)/255.0f
int net_quantize_input(b200_net *net, const float *x, long batch, bool refresh = false);
void net_xq_clear(b200_net *net);
// end of a minimisation: drop the copy unless the caller built it explicitly (the caller may refill x afterwards)
inline void net_xq_release_solver(b200_net *net) { if (net && !net->xq.user) net_xq_clear(net); }
// graphs captured by parked L-BFGS solvers bake this network's buffers in: dropped when the network is destroyed (solvers.cu)
void lbfgs_pool_forget_net(b200_ctx *ctx, unsigned long long net_uid);
// the uint8 rows matching x (a row-aligned sub-range of the quantised input), or nullptr
const uint8_t *net_xq_lookup(b200_net *net, const float *x, long batch);
struct X16View { const void *base; long rows_total, row0; int nblocks; };
// the fp16 copy and the row range matching x (row-aligned sub-range of the quantised input); false if there is none
bool net_x16_view(b200_net *net, const float *x, long batch, X16View *v);
// gathered mini-batches (b200_net::GatherQ): buffer for up to `rows` rows (nullptr unless the input is quantised), and the
// registration of the batch just gathered into it as the fp16 view of the fp32 rows `src`
void *net_gather16_buffer(b200_net *net, long rows);
void net_gather16_register(b200_net *net, const float *src, long rows);
const void *net_x16_source(const b200_net *net, long *rows_total, int *nblocks); // the quantised input's fp16 copy, or nullptr
bool net_spec_capable(b200_net *net, const float *x, long batch);
// the skinny last layer in one pass: forward, loss, both deltas and the [dW_L; db_L] partials (tail_layer.cu)
bool tail_applicable(const b200_net *net);
// want16: delta_{L-1} also (or only) as scaled fp16 {hi | lo} in net->delta16 (two-layer nets: it feeds gemm_dw16.cu directly).
// chain16 (deeper nets): size net->delta16 for delta_0 and derive its scale from max |delta_L| and the weights of layers 1..L-1;
// the DX kernel of layer 1 then writes it (tc_dx_layer, emit16)
int tail_layer(b200_net *net, const float *params, const float *t, long batch, float inv_batch, bool want32, bool want16,
               bool chain16 = false);
bool tail_chain16_applicable(const b200_net *net);
// hidden layer 1 on the fp16 kernels (see b200_net::Mid16): shape / mode test, buffers, per-evaluation operand preparation
bool mid16_applicable(const b200_net *net);
int mid16_ensure(b200_net *net, long batch);
void mid16_release(b200_net *net);
int mid16_reconstruct_act0(b200_net *net); // fp32 A_1 from the pair, into act[0] (debug read-back)
int tail_ensure_scalars(b200_net *net);

// Weight-dependent factor of the fp16 scale of delta_0 in a net with more than two layers: prod_{l=1..L-1} max_f ||W_l[f,:]||_1
// (|act'| <= 1, so max |delta_0| <= max |delta_L| * this). Computed once per evaluation by a few extra CTAs of the weight-split
// kernel (gemm_fwd16.cu), one warp per row of any W_l so that each CTA costs one memory latency; they leave per-layer maxima
// per CTA, and block 0 of the last-layer backward kernel combines them with max |delta_L| and publishes the scale.
constexpr int kMaxChain = 8, kChainRowsPerCta = 32, kMaxChainCtas = 64;
struct ChainW {
  const float *W[kMaxChain]; // layer l: [in][out]
  int in[kMaxChain], out[kMaxChain];
  int nl;                    // 0: nothing to do
  int nctas;                 // ceil(total rows / kChainRowsPerCta)
  float *cw_part;            // [nctas][kMaxChain]
};
void tail_chain_fill(const b200_net *net, const float *params, ChainW *c);
#ifdef __CUDACC__
// cta: index among the chain CTAs; red: kMaxChain * 32 floats of shared memory; needs blockDim.x >= 32 * kChainRowsPerCta
__device__ __forceinline__ void chain_cw_block(const ChainW &c, int cta, float *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int total = 0;
  for (int l = 0; l < c.nl; ++l) total += c.in[l];
  int g = cta * kChainRowsPerCta + warp, l = 0;
  const bool ok = warp < kChainRowsPerCta && g < total;
  if (!ok) g = 0;
  while (g >= c.in[l]) { g -= c.in[l]; ++l; }
  const int out = c.out[l];
  const float *row = c.W[l] + (size_t)g * out;
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = (ok && lane + 32 * k < out) ? fabsf(__ldg(row + lane + 32 * k)) : 0.0f;
  float s = (v[0] + v[1]) + (v[2] + v[3]);
  if (ok) for (int o = lane + 128; o < out; o += 32) s += fabsf(__ldg(row + o));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0 && warp < kChainRowsPerCta)
#pragma unroll
    for (int j = 0; j < kMaxChain; ++j) red[j * 32 + warp] = (ok && j == l) ? s : 0.0f;
  __syncthreads();
  if (warp == 0) { // lane j reduces layer j
    float m = 0.0f;
    if (lane < kMaxChain)
      for (int i = 0; i < kChainRowsPerCta; ++i) m = fmaxf(m, red[lane * 32 + i]);
    if (lane < kMaxChain) c.cw_part[cta * kMaxChain + lane] = m;
  }
}
#endif
void tail_release(b200_net *net);
// wide hidden layers on the fp16 pair kernels (b200_net::Wide16, gemm_fwd16.cu). role: 0 forward, 1 dX, 2 dW
bool wide16_applicable(const b200_net *net, int l, int role, long batch);
void wide16_begin(b200_net *net); // start of an evaluation: no operand has been split yet
int wide16_forward_layer(b200_net *net, int l, const float *params, const float *in, long batch);
int wide16_dx_layer(b200_net *net, int l, const float *params, long batch);
int wide16_dw_layer(b200_net *net, int l, const float *in, long batch);
// the delta pair of the layer below a skinny last layer, generated from delta_L (no fp32 delta_{L-1} is written)
bool wide16_last_dx_applicable(const b200_net *net, long batch);
int wide16_last_dx(b200_net *net, const float *params, long batch);
void wide16_release(b200_net *net);
// the reference CPU backend's random-mini-batch SGD (src/minimizer/s_gd.hpp:63-170) on the GPU (slbfgs.cu: shares the sampler
// and the row gather of S-LBFGS)
int sgd_random_solve(b200_ctx *ctx, b200_net *net, int n, float *params, const float *input, const float *target, int total_samples,
                     const b200_sgd_opts &o, b200_history *hist);

} // namespace b200

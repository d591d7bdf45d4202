/* b200_lbfgs.h — C ABI of libb200lbfgs.so
 *
 * B200-native (sm_100a) backend for the optimizer-plus-objective hot path of
 * SignorB/lbfgs-FFNN. Plain C: opaque handles, raw pointers, sizes. No torch / C++ types.
 *
 * Every entry returns an int status (0 = B200_OK) unless stated otherwise; the C++ shim in
 * include/cuda_mlp/ turns non-zero into the reference's print + std::abort()
 * (reference: src/cuda/common.cuh:18-23). b200_last_error() gives the message.
 *
 * Conventions shared with the reference (src/cuda/layer.cuh:48-58, network.cuh:37-59):
 *   - all matrices column-major: X is in x B (sample b = X[b*in .. b*in+in)), targets out x B,
 *     W_l is out x in with ld = out;
 *   - flat parameter vector = per layer [W (out*in) | b (out)], gradients mirror it;
 *   - loss = 0.5*||A_L - T||^2 / B, dL/dA_L = (A_L - T)/B  (network.cuh:97-107);
 *   - ActivationType {Linear=0,Tanh=1,ReLU=2,Sigmoid=3} (kernels.cuh:53-58).
 * "device" pointers are CUDA device pointers on the context's device.
 */
#ifndef B200_LBFGS_H
#define B200_LBFGS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1

typedef struct b200_ctx b200_ctx;       /* replaces cuda_mlp::CublasHandle (src/cuda/cublas_handle.cuh:22-39) */
typedef struct b200_net b200_net;       /* replaces cuda_mlp::CudaNetwork  (src/cuda/network.cuh:16-156)       */

enum { B200_OK = 0, B200_ERR_INVALID = 1, B200_ERR_CUDA = 2, B200_ERR_COMM = 3, B200_ERR_UNSUPPORTED = 4 };
enum { B200_ACT_LINEAR = 0, B200_ACT_TANH = 1, B200_ACT_RELU = 2, B200_ACT_SIGMOID = 3 };
/* arithmetic of the GEMMs: FP32 = FFMA (bit-for-bit fp32 products), TF32X3 = tcgen05 with hi/lo split
 * (fp32-accurate), TF32 = single-pass tcgen05 (tolerance stated in tests/test_gpu_parity.py). */
enum { B200_PREC_FP32 = 0, B200_PREC_TF32X3 = 1, B200_PREC_TF32 = 2 };
/* line-search policy: ARMIJO = reference CUDA backend (src/cuda/lbfgs.cuh:97-147),
 * WOLFE = reference CPU backend (src/minimizer/full_batch_minimizer.hpp:126-157, lbfgs.hpp:60-65). */
enum { B200_LS_ARMIJO = 0, B200_LS_WOLFE = 1 };

const char *b200_last_error(void);
int b200_abi_version(void);

/* ---- context: device, stream, workspace, (optional) NCCL communicator -------------------------- */
int b200_ctx_create(int device, b200_ctx **out);
int b200_ctx_destroy(b200_ctx *ctx);
int b200_ctx_synchronize(b200_ctx *ctx);
void *b200_ctx_stream(b200_ctx *ctx); /* cudaStream_t the library launches on */
/* adopt a caller-owned stream (e.g. the host framework's current stream); NULL restores the library's own */
int b200_ctx_set_stream(b200_ctx *ctx, void *cuda_stream);
int b200_ctx_device(b200_ctx *ctx);
/* per-launch CUDA-event profiling of the library's own kernels (bench.py's roofline figures): enable, run,
 * then report writes a JSON object {"kernel-class": [launches, total_ms], ...} into json_out. */
int b200_ctx_profile(b200_ctx *ctx, int enable);
int b200_ctx_profile_report(b200_ctx *ctx, char *json_out, size_t capacity);
/* multi-GPU: one process per GPU. Rank 0 calls b200_comm_unique_id and ships the 128 bytes to the
 * other ranks (torch.distributed / MPI / file); every rank then calls b200_ctx_init_comm. */
int b200_comm_unique_id(void *out_128_bytes);
int b200_ctx_init_comm(b200_ctx *ctx, const void *unique_id_128_bytes, int rank, int world);
int b200_ctx_rank(b200_ctx *ctx);
int b200_ctx_world(b200_ctx *ctx);
/* test hook: sum-allreduce n floats in place on the context stream */
int b200_ctx_allreduce_f32(b200_ctx *ctx, float *dev, size_t n);

/* ---- raw device memory (replaces cuda_mlp::DeviceBuffer<T>, src/cuda/device_buffer.cuh:7-96) ---- */
int b200_malloc(void **dev, size_t bytes);
int b200_free(void *dev);
int b200_memcpy_h2d(void *dev, const void *host, size_t bytes);
int b200_memcpy_d2h(void *host, const void *dev, size_t bytes);
int b200_memcpy_d2d(void *dst, const void *src, size_t bytes);
int b200_memset(void *dev, int value, size_t bytes);
int b200_host_alloc_pinned(void **host, size_t bytes);
int b200_host_free_pinned(void *host);

/* ---- network objective (CudaNetwork) --------------------------------------------------------------- */
/* addLayer x nlayers (network.cuh:27-30): dims has nlayers+1 entries, acts nlayers. */
int b200_net_create(b200_ctx *ctx, int nlayers, const int *dims, const int *acts, b200_net **out);
int b200_net_destroy(b200_net *net);
size_t b200_net_params_size(b200_net *net);
int b200_net_output_size(b200_net *net);
/* bindParams (network.cuh:37-59): allocates the flat params/grads buffers, initialises on the HOST with
 * std::mt19937(seed) + normal_distribution<float>(0, scale*sqrt(1/in)) for weights, biases = 0, uploads. */
int b200_net_bind_params(b200_net *net, unsigned seed);
float *b200_net_params_data(b200_net *net); /* device */
float *b200_net_grads_data(b200_net *net);  /* device */
int b200_net_zero_grads(b200_net *net);
int b200_net_set_precision(b200_net *net, int prec);
int b200_net_get_precision(b200_net *net);
/* L2 term of the S-LBFGS objective: loss += 0.5*lambda*||w||^2, grad += lambda*w
 * (src/unified_optimization.hpp:334,375,398). 0 (default) = the plain reference objective. */
int b200_net_set_l2(b200_net *net, float lambda);
/* sample-sharded multi-GPU: the B of the 1/B scaling. 0 (default) = shard batch x number of ranks. */
int b200_net_set_global_batch(b200_net *net, long batch_global);
/* Input cache for 8-bit image data. If EVERY element of x_dev[batch][in] is exactly float(u)/255.0f, u in 0..255
 * (what the reference's loader produces, tests/mnist/mnist_loader.hpp:59) a uint8 copy is kept and the tensor-core
 * modes read it instead of the fp32 array in the two input-bound GEMMs of layer 0 (4x fewer HBM bytes; u is exact in
 * TF32, results agree with the fp32 path to ~1e-7). Otherwise nothing happens (*quantized = 0). The solvers call it
 * themselves at the start of a solve (x is constant during a solve) and clear it at the end; callers of the one-shot
 * entry points may call it explicitly and MUST call b200_net_clear_input_cache before modifying x in place. */
int b200_net_quantize_input(b200_net *net, const float *x_dev, long batch, int *quantized);
int b200_net_clear_input_cache(b200_net *net);
/* forward_only (network.cuh:79-88): X device (in x batch). Output stays on device. */
int b200_net_forward(b200_net *net, const float *x_dev, long batch);
/* compute_loss_and_grad (network.cuh:97-119): evaluates at the network's bound params, writes the gradient
 * into the bound grads buffer, returns the loss through *loss_host (one stream sync, like the reference's
 * blocking cublasSdot). `batch_global` > 0 overrides the 1/B scaling (sample-sharded multi-GPU). */
int b200_net_loss_grad(b200_net *net, const float *x_dev, const float *t_dev, long batch, float *loss_host);
/* same at an explicit parameter vector, gradient to an explicit buffer, loss left on the device
 * (double, *loss_dev) — no host sync. params_dev/grad_dev may be NULL to use the bound buffers. */
int b200_net_loss_grad_async(b200_net *net, const float *params_dev, const float *x_dev, const float *t_dev, long batch,
                             float *grad_dev, double *loss_dev);
/* copy_output_to_host (network.cuh:121-126) */
int b200_net_copy_output_to_host(b200_net *net, float *host, size_t n);
int b200_net_last_batch(b200_net *net);
/* the activations of layer `layer` (0 = first hidden layer ... nlayers-1 = the output, i.e. copy_output_to_host) left by the last
 * forward / loss_grad call: n <= out_layer * last_batch floats, sample-major [batch][out]. The reference keeps them in
 * CudaNetwork::activations_ (network.cuh:133-147) without an accessor; the parity tests read the hidden layers' signs. */
int b200_net_copy_activation_to_host(b200_net *net, int layer, float *host, size_t n);
/* UnifiedLauncher<CudaBackend>::evaluate (src/unified_launcher.hpp:154-199) on the device:
 * forward + MSE (mean over batch*out) + arg-max accuracy (percent). t_dev is out x batch. */
int b200_net_evaluate(b200_net *net, const float *x_dev, const float *t_dev, long batch, double *mse, double *accuracy);

/* ---- minimizers (CudaMinimizerBase and strategies) ------------------------------------------------ */
/* LossGradFun (src/cuda/minimizer_base.cuh:15-16) as a C callback; all pointers are device pointers. */
typedef float (*b200_loss_grad_fn)(void *user, const float *params, float *grad, const float *input, const float *target,
                                   int batch);

typedef struct b200_history { /* IterationRecorder<CudaBackend> contents (src/iteration_recorder.hpp:81-146) */
  int capacity;               /* in: entries the arrays can hold */
  int size;                   /* out: entries written */
  float *loss;                /* host arrays, may be NULL */
  float *grad_norm;
  float *time_ms;             /* cumulative, per-iteration CUDA-event time like lbfgs.cuh:176-182 */
  int iterations;             /* out: CudaMinimizerBase::iterations() */
  long evaluations;           /* out: loss+grad evaluations performed */
  long launches;              /* out: kernels launched by the library during the solve */
} b200_history;

typedef struct b200_lbfgs_opts { /* defaults: minimizer_base.cuh:63-64, lbfgs.cuh:263 */
  int max_iters;                 /* 200  */
  float tol;                     /* 1e-6: stop when ||g|| < tol */
  int memory;                    /* 16: history size m */
  int max_line_iters;            /* 20 (ARMIJO) / 50 (WOLFE) */
  float c1;                      /* 1e-4 */
  float rho;                     /* 0.5 backtracking factor */
  float c2;                      /* 0.9 (WOLFE only) */
  int linesearch;                /* B200_LS_ARMIJO */
  int record_timing;             /* 1: per-iteration CUDA-event timing when a history is attached */
  int shard_history;             /* multi-GPU: 0 = history replicated on every rank (gradient all-reduce);
                                  * 1 = x, g, S, Y sharded by parameter index: gradient reduce-scatter, only the
                                  *     5(m+1)+1 partial dot products all-reduced, parameters all-gathered;
                                  * -1 (default) = sharded when n >= 4M parameters and more than one rank */
} b200_lbfgs_opts;
void b200_lbfgs_default_opts(b200_lbfgs_opts *o);

/* CudaLBFGS::solve (src/cuda/lbfgs.cuh:39-194). If `net` is non-NULL the library's own objective is used
 * (fast path: loss stays on the device, one host sync per line-search trial) and `fn` is ignored;
 * otherwise `fn` is called exactly like the reference calls loss_grad. `params` is updated in place
 * (aliasing contract of src/unified_optimization.hpp:483-491,503-504: with a bound network the caller
 * passes net's own params buffer). With an initialised communicator, input/target/batch are this rank's
 * sample shard and the gradient/loss are all-reduced (sum) over ranks. */
int b200_lbfgs_solve(b200_ctx *ctx, b200_net *net, b200_loss_grad_fn fn, void *user, int n, float *params,
                     const float *input, const float *target, int batch, const b200_lbfgs_opts *opts,
                     b200_history *hist);
/* The same solve as a resumable object: create allocates the work vectors the reference allocates per solve()
 * (lbfgs.cuh:53-71); each run continues the SAME minimisation for up to `iters` more iterations (the first run
 * performs the initial evaluation, lbfgs.cuh:78). b200_lbfgs_solve == create + run(max_iters) + destroy.
 * hist (may be NULL) is filled from index 0 on every run. */
typedef struct b200_lbfgs b200_lbfgs;
int b200_lbfgs_create(b200_ctx *ctx, int n, const b200_lbfgs_opts *opts, b200_lbfgs **out);
int b200_lbfgs_run(b200_lbfgs *solver, b200_net *net, b200_loss_grad_fn fn, void *user, float *params,
                   const float *input, const float *target, int batch, int iters, b200_history *hist);
int b200_lbfgs_destroy(b200_lbfgs *solver);

typedef struct b200_gd_opts { /* src/cuda/gd.cuh:108-110 */
  int max_iters;              /* 200 */
  float tol;                  /* 1e-6 */
  float lr;                   /* 0.01 */
  float momentum;             /* 0.9 */
  int record_timing;
} b200_gd_opts;
void b200_gd_default_opts(b200_gd_opts *o);
/* CudaGD::solve (src/cuda/gd.cuh:38-106) */
int b200_gd_solve(b200_ctx *ctx, b200_net *net, b200_loss_grad_fn fn, void *user, int n, float *params,
                  const float *input, const float *target, int batch, const b200_gd_opts *opts, b200_history *hist);

typedef struct b200_sgd_opts { /* src/cuda/sgd.cuh:156-163 */
  int max_iters;               /* 200 epochs */
  float tol;                   /* 1e-6 relative epoch-loss improvement */
  float lr;                    /* 0.01 */
  float momentum;              /* 0.9 */
  float decay_rate;            /* 1.0 */
  int decay_step;              /* 0 = off */
  int batch_size;              /* 64 */
  int input_dim;               /* must be set (sgd.cuh:61-65) */
  int output_dim;
  int record_timing;
  int sampling;                /* 0 (default): sequential slices, the CUDA backend's CudaSGD (sgd.cuh:100-109).
                                * 1: the CPU backend's StochasticGradientDescent (src/minimizer/s_gd.hpp:63-170 with the closures of
                                *    UnifiedSGD_CPU, src/unified_optimization.hpp:219-300): N / batch_size random mini-batches per
                                *    epoch drawn by the partial Fisher-Yates sampler from one mt19937(seed), w -= lr * mean gradient,
                                *    no decay and no stopping test; momentum is honoured if > 0 (the CPU class has none). */
  unsigned seed;               /* sampling == 1: 123 (kDefaultSeed) */
} b200_sgd_opts;
void b200_sgd_default_opts(b200_sgd_opts *o);
/* CudaSGD::solve (src/cuda/sgd.cuh:50-153): sequential unshuffled mini-batches by pointer offset; or, with
 * opts->sampling == 1, the reference CPU backend's random-mini-batch SGD on the GPU (network objective only). */
int b200_sgd_solve(b200_ctx *ctx, b200_net *net, b200_loss_grad_fn fn, void *user, int n, float *params,
                   const float *input, const float *target, int total_samples, const b200_sgd_opts *opts,
                   b200_history *hist);

typedef struct b200_slbfgs_opts { /* UnifiedSLBFGS_CPU::optimize (src/unified_optimization.hpp:314-407) */
  int max_iters;                  /* epochs */
  float tol;                      /* 1e-4: stop when the anchor's full-gradient norm < tol */
  float step_size;                /* eta */
  int batch_size;                 /* b */
  int memory;                     /* M_param */
  int L;                          /* curvature-pair interval */
  int b_H;                        /* 0 => batch_size/2 */
  float lambda;                   /* 1e-4 L2 term */
  float epsilon;                  /* 1e-4 finite-difference step */
  unsigned seed;                  /* 123 (kDefaultSeed) */
  int record;                     /* 1: full loss + full gradient norm per epoch (s_lbfgs.hpp:274-284) */
  int pair_eval;                  /* 1: the two evaluations of a step (w_t and the anchor; u + eps s and u - eps s) run as ONE
                                   * forward/backward of the stacked "pair network" on the shared mini-batch, with
                                   * v = g_t - g_k + mu / y = (g+ - g-)/(2 eps) formed as the gradients are read back;
                                   * 0: two separate evaluations (same results to rounding) */
  float hvp_step_scale;           /* 256: the +-eps*s pair of the finite-difference Hessian-vector product is evaluated at
                                   * +-(scale*eps)*s and divided by 2*scale*eps. The reference computes the pair in double; in
                                   * fp32, eps*s = 1e-4*s is below one ulp of most weights (|s|/|w| ~ 1e-3), so at scale 1 the
                                   * displacement w +- eps*s itself is mostly rounding: y is noise, and the quasi-Newton update
                                   * built on it diverges in some runs (NaN after 1-3 epochs, measured: DESIGN.md). From scale
                                   * ~256 on the result no longer depends on the arithmetic mode or on how the pair is
                                   * evaluated (3 digits). 1 = the reference's step verbatim. */
} b200_slbfgs_opts;
void b200_slbfgs_default_opts(b200_slbfgs_opts *o);
/* SLBFGS::stochastic_solve (src/minimizer/s_lbfgs.hpp:165-290) on the GPU. New functionality: the reference
 * static_asserts on UnifiedSLBFGS<CudaBackend> (src/unified_optimization.hpp:639-641). */
int b200_slbfgs_solve(b200_ctx *ctx, b200_net *net, int n, float *params, const float *input, const float *target,
                      int total_samples, const b200_slbfgs_opts *opts, b200_history *hist);
/* the host-side index sampler used by b200_slbfgs_solve (s_lbfgs.hpp:141-161), exposed for parity tests:
 * `count` consecutive draws of `b` indices out of N from one mt19937(seed). */
int b200_slbfgs_sample_stream(unsigned seed, long N, long b, int count, uint32_t *out_host);

/* ---- building blocks, exposed for parity tests and for callers that drive their own loop ---------- */
/* compute_direction_ring (src/cuda/lbfgs.cuh:206-261) in compact form. S, Y: k x n row-major device arrays in
 * LOGICAL order (row 0 oldest), rho: k host floats, g/p: n device floats. policy B200_LS_ARMIJO applies the
 * CUDA backend's gamma guard, B200_LS_WOLFE the CPU backend's unguarded gamma; 2 = S-LBFGS two-loop
 * (returns +H*g, gamma clamped to [1e-6,1e6], s_lbfgs.hpp:105-136). k == 0 gives p = -g (or +g for 2). */
int b200_lbfgs_direction(b200_ctx *ctx, size_t n, int k, const float *S_dev, const float *Y_dev, const float *rho_host,
                         const float *g_dev, int policy, float *p_dev, double *g_dot_p_host);
/* BLAS-1 replacements (src/cuda/kernels.cuh:14-50); dot/nrm2 return through host pointers (one sync). */
int b200_vec_dot(b200_ctx *ctx, const float *x, const float *y, size_t n, double *out_host);
int b200_vec_nrm2(b200_ctx *ctx, const float *x, size_t n, double *out_host);
int b200_vec_axpy(b200_ctx *ctx, size_t n, float alpha, const float *x, float *y);
int b200_vec_scal(b200_ctx *ctx, size_t n, float alpha, float *x);
/* y = x0 + alpha*p (line-search trial point, lbfgs.cuh:116-117 fused) */
int b200_vec_trial_point(b200_ctx *ctx, size_t n, const float *x0, float alpha, const float *p, float *y);
/* double -> float conversion on the device (UnifiedLauncher::setData upload, src/unified_launcher.hpp:105-128) */
int b200_convert_f64_to_f32(b200_ctx *ctx, const double *src_dev, float *dst_dev, size_t n);

/* kernels launched by this library in the calling process since load (bench.py's gpu_launches) */
long b200_launch_count(void);
/* The B200_* debugging switches of the environment are read once, when the library is first used; nothing on the
 * per-iteration path calls getenv. Tests that flip a switch in a running process call this to re-read them. */
int b200_debug_reload_env(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_LBFGS_H */

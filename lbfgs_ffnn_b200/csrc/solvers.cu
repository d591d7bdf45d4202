// Minimizer strategies: CudaLBFGS / CudaGD / CudaSGD (src/cuda/lbfgs.cuh, gd.cuh, sgd.cuh) behind the C ABI.
//
// The control flow of each solve() follows the reference line by line (cited inline); what changes is
// where the arithmetic happens. Per L-BFGS iteration the reference issues ~2k+6 blocking cuBLAS
// reductions and ~2k+10 BLAS-1 / memcpy launches; here an iteration is
//     dots -> solve -> apply(+ trial point) -> evaluation          (3 + evaluation kernels)
// with ONE host synchronisation per line-search trial (the Armijo / Wolfe test itself, which the
// reference also performs on the host).
#include "lbfgs_kernels.cuh"
#include "network.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace b200 {

namespace {

struct HostMail { // pinned mailbox layout (ctx->h_scalars, 64 doubles)
  double loss, gnorm2; // EvalOut
  double gnew_dot_p;   // WOLFE
  double pad;
  LbfgsHeader hdr;     // copy of the device header after lbfgs_solve_kernel
};
static_assert(2 * sizeof(HostMail) + 16 <= 64 * sizeof(double), "mailbox overflow"); // two mailboxes (graph parity) + a scratch double

// Objective seen by the minimizers: either the library's own network (asynchronous, result on the
// device) or a caller-supplied LossGradFun (synchronous, loss returned on the host).
struct Objective {
  b200_ctx *ctx;
  b200_net *net;
  b200_loss_grad_fn fn;
  void *user;
  const float *input, *target;
  int batch;
  long batch_global;
  long evals = 0;
  double *d_part = nullptr; // dot partials (callback path)
  size_t n = 0;

  // launches the evaluation of x into grad; the result lands in mail->loss / mail->gnorm2 after sync()
  int eval_async(const float *x, float *grad, HostMail *mail, double *host_loss_cb) {
    ++evals;
    if (net) {
      B200_TRY(net_eval(net, x, input, target, batch, batch_global, grad, (EvalOut *)net->eval_out));
      if (!net->host_out) // (graph path: the finishing kernel writes the pinned mailbox itself)
        B200_CUDA(cudaMemcpyAsync(&mail->loss, net->eval_out, sizeof(EvalOut), cudaMemcpyDeviceToHost, ctx->stream));
      return B200_OK;
    }
    // reference calling convention (src/cuda/minimizer_base.cuh:15-16): blocking, loss on the host
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    *host_loss_cb = (double)fn(user, x, grad, input, target, batch);
    B200_CUDA(cudaDeviceSynchronize()); // the callback may have used any stream
    B200_TRY(launch_dot(ctx, grad, grad, n, d_part, ctx->d_scalars));
    B200_CUDA(cudaMemcpyAsync(&mail->gnorm2, ctx->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return B200_OK;
  }
};

struct Timer { // per-iteration wall time of the stream's work, as src/cuda/lbfgs.cuh:80-91,176-182 records it (cumulative ms).
  // Host clock around a drained stream instead of the reference's event pair: two event records and an event wait per
  // iteration were ~5 us of a ~200 us iteration.
  b200_ctx *ctx;
  bool on;
  float elapsed = 0.f;
  std::chrono::steady_clock::time_point t0;
  int start() {
    if (on) t0 = std::chrono::steady_clock::now();
    return B200_OK;
  }
  int stop(bool drain = true) { // drain = false: the caller has this iteration's results already (work queued behind them may run on)
    if (!on) return B200_OK;
    if (drain) B200_CUDA(cudaStreamSynchronize(ctx->stream));
    elapsed += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return B200_OK;
  }
};

void record(b200_history *h, int idx, double loss, double gnorm, float ms) {
  if (!h || idx >= h->capacity) return;
  if (h->loss) h->loss[idx] = (float)loss;
  if (h->grad_norm) h->grad_norm[idx] = (float)gnorm;
  if (h->time_ms) h->time_ms[idx] = ms;
  h->size = std::max(h->size, idx + 1);
}

} // namespace

} // namespace b200

using namespace b200;

extern "C" {

void b200_lbfgs_default_opts(b200_lbfgs_opts *o) {
  if (!o) return;
  o->max_iters = 200; o->tol = 1e-6f; o->memory = 16; o->max_line_iters = 20; // minimizer_base.cuh:61-65, lbfgs.cuh:263
  o->c1 = 1e-4f; o->rho = 0.5f; o->c2 = 0.9f;
  o->linesearch = B200_LS_ARMIJO; o->record_timing = 1; o->shard_history = -1;
}
void b200_gd_default_opts(b200_gd_opts *o) {
  if (!o) return;
  o->max_iters = 200; o->tol = 1e-6f; o->lr = 0.01f; o->momentum = 0.9f; o->record_timing = 1;
}
void b200_sgd_default_opts(b200_sgd_opts *o) {
  if (!o) return;
  o->max_iters = 200; o->tol = 1e-6f; o->lr = 0.01f; o->momentum = 0.9f; o->decay_rate = 1.0f; o->decay_step = 0;
  o->batch_size = 64; o->input_dim = 0; o->output_dim = 0; o->record_timing = 1; o->sampling = 0; o->seed = 123;
}

// ===================================================================================================
// CudaLBFGS::solve (src/cuda/lbfgs.cuh:39-194) / cpu_mlp::LBFGS::solve (src/minimizer/lbfgs.hpp:38-100)
// ===================================================================================================
} // extern "C"

struct b200_lbfgs {
  b200_ctx *ctx = nullptr;
  b200_lbfgs_opts o{};
  size_t N = 0, ld = 0;
  // sharded history (multi-GPU): this rank owns parameter indices [lo, lo + len); chunk = shard size of every rank but the last
  bool sharded = false;
  size_t lo = 0, len = 0, chunk = 0;
  float *gfull = nullptr; // sharded: this rank's un-reduced full-length gradient
  double *totals = nullptr;
  int m = 0, mp = 1, mod = 1, policy = POLICY_ARMIJO;
  int nblk = 1, apply_blocks = 1;
  char *ws = nullptr;
  size_t ws_bytes = 0;
  LbfgsView view{};
  double *partials = nullptr, *dot_part = nullptr;
  unsigned *gridbar = nullptr; // {arrivals, generation} of the fused direction kernel's grid-wide barrier
  float *gbuf[2] = {nullptr, nullptr}, *p = nullptr, *x_prev = nullptr, *S = nullptr, *Y = nullptr;
  // one CUDA graph per (gradient-buffer parity, history-reset flag): direction kernels + first trial evaluation +
  // the two scalar read-backs of a steady-state iteration are ONE launch instead of ~10
  // ... and per launch kind: [][][0] launched after the host has accepted the previous iteration, [][][1] launched speculatively
  // behind it (every kernel gated on the device-side Armijo decision, common.cuh SpecState)
  cudaGraphExec_t graph[2][2][2] = {};
  long graph_launches[2][2][2] = {};
  SpecState *spec_dev = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr}; // end of the graph of each parity
  bool ahead = false;                     // the graph of the CURRENT iteration is already in flight (launched speculatively)
  // net uid (not its address: a new network can be allocated where a destroyed one lived), params, input, target, batch, net config
  // generation, generation of the peer-memory all-reduce buffers
  unsigned long long gkey[7] = {0, 0, 0, 0, 0, 0, 0};
  bool graphs_ok = true;
  bool spec_capable = false; // the captured evaluation consists of gated kernels only
  unsigned long long last_net_uid = 0; // network whose input cache this minimisation (re)built (uid: it may be destroyed first)
  // minimisation state carried across runs
  bool started = false;
  int cur = 0, iter = 0, reset_next = 0;
  double loss = 0.0, gnorm = 0.0;
};

extern "C" {

int b200_lbfgs_create(b200_ctx *ctx, int n, const b200_lbfgs_opts *opts, b200_lbfgs **out) {
  B200_REQUIRE(ctx && out && n > 0, "bad argument");
  b200_lbfgs_opts o;
  if (opts) o = *opts; else b200_lbfgs_default_opts(&o);
  B200_REQUIRE(o.memory >= 0 && o.memory <= kMaxSlots - 1, "memory must be in [0, 256]");
  B200_CUDA(cudaSetDevice(ctx->device));
  const bool wolfe = (o.linesearch == B200_LS_WOLFE);
  const bool want_sharded = ctx->world > 1 && !wolfe && (o.shard_history == 1 || (o.shard_history < 0 && (size_t)n >= (size_t(4) << 20)));
  for (size_t i = 0; i < ctx->lbfgs_pool.size(); ++i) { // a parked solver of the same shape: reuse work space and graphs
    b200_lbfgs *c = (b200_lbfgs *)ctx->lbfgs_pool[i];
    if (c->N == (size_t)n && c->m == o.memory && (c->policy == POLICY_WOLFE) == wolfe && c->sharded == want_sharded) {
      ctx->lbfgs_pool.erase(ctx->lbfgs_pool.begin() + i);
      c->o = o;
      c->started = false; c->cur = 0; c->iter = 0; c->reset_next = 0; c->loss = 0.0; c->gnorm = 0.0; c->ahead = false;
      B200_CUDA(cudaMemsetAsync(c->ws, 0, c->ws_bytes, ctx->stream));
      B200_CUDA(cudaMemsetAsync(c->gridbar, 0, 2 * sizeof(unsigned), ctx->stream));
      B200_TRY(lbfgs_init_state(c->view, c->m, c->mod, ctx->stream));
      // the device-side Armijo / convergence gate of the kept graphs reads c1 and tol from spec_dev: they must be THIS solve's
      // (a stale pair lets host and device disagree on whether a speculative iteration ran)
      {
        SpecState hs{};
        hs.c1 = (double)o.c1; hs.tol = (double)o.tol;
        hs.alpha0 = &c->view.h->alpha0; hs.gdotp = &c->view.h->gdotp;
        B200_CUDA(cudaMemcpyAsync(c->spec_dev, &hs, sizeof(hs), cudaMemcpyHostToDevice, ctx->stream));
        B200_CUDA(cudaStreamSynchronize(ctx->stream)); // hs is a stack object
      }
      *out = c;
      return B200_OK;
    }
  }
  b200_lbfgs *s = new b200_lbfgs;
  s->ctx = ctx;
  s->o = o;
  s->m = o.memory;
  s->policy = wolfe ? POLICY_WOLFE : POLICY_ARMIJO;
  s->mod = wolfe ? s->m + 1 : std::max(s->m, 1);
  s->mp = s->m + 1;
  s->N = (size_t)n;
  s->sharded = ctx->world > 1 && !wolfe && (o.shard_history == 1 || (o.shard_history < 0 && s->N >= (size_t(4) << 20)));
  s->chunk = s->sharded ? ((((s->N + ctx->world - 1) / ctx->world) + 3) & ~size_t(3)) : s->N;
  s->lo = s->sharded ? std::min(s->N, (size_t)ctx->rank * s->chunk) : 0;
  s->len = s->sharded ? std::min(s->chunk, s->N - s->lo) : s->N;
  s->ld = (std::max<size_t>(s->len, 4) + 3) & ~size_t(3);
  // one allocation for all work vectors (the reference allocates 6 + 2m DeviceBuffers per solve, lbfgs.cuh:53-71)
  s->nblk = lbfgs_dots_blocks(ctx, std::max<size_t>(s->len, 1));
  const int ncols = kDotsCols * s->mp + 1;
  const size_t state_bytes = lbfgs_state_bytes(s->m);
  const size_t part_bytes = sizeof(double) * (size_t)s->nblk * ncols;
  const size_t dotp_bytes = sizeof(double) * (size_t)dot_blocks(ctx, s->N);
  const size_t vec_bytes = sizeof(float) * s->ld;
  const size_t gfull_bytes = s->sharded ? ((sizeof(float) * s->N + 255) & ~size_t(255)) : 0;
  const size_t totals_bytes = (sizeof(double) * ncols + 255) & ~size_t(255);
  const size_t total = state_bytes + part_bytes + dotp_bytes + 512 + vec_bytes * (4 + 2 * (size_t)s->mp) + gfull_bytes + totals_bytes;
  if (cudaMalloc(&s->ws, total) != cudaSuccess) {
    delete s;
    set_error("cudaMalloc of %zu bytes of L-BFGS work space failed", total);
    return B200_ERR_CUDA;
  }
  s->ws_bytes = total;
  cudaStream_t st = ctx->stream;
  B200_CUDA(cudaMemsetAsync(s->ws, 0, total, st));
  size_t off = 0;
  s->view = lbfgs_view(s->ws + off, s->m); off += state_bytes;
  s->partials = (double *)(s->ws + off); off += part_bytes;
  s->dot_part = (double *)(s->ws + off); off += dotp_bytes;
  off = (off + 255) & ~size_t(255);
  s->gbuf[0] = (float *)(s->ws + off); s->gbuf[1] = (float *)(s->ws + off + vec_bytes); off += 2 * vec_bytes;
  s->p = (float *)(s->ws + off); off += vec_bytes;
  s->x_prev = (float *)(s->ws + off); off += vec_bytes;
  s->S = (float *)(s->ws + off); off += vec_bytes * s->mp;
  s->Y = (float *)(s->ws + off); off += vec_bytes * s->mp;
  off = (off + 255) & ~size_t(255);
  s->totals = (double *)(s->ws + off); off += totals_bytes;
  if (s->sharded) s->gfull = (float *)(s->ws + off);
  B200_CUDA(cudaMalloc(&s->gridbar, 2 * sizeof(unsigned)));
  B200_CUDA(cudaMalloc(&s->spec_dev, sizeof(SpecState)));
  B200_CUDA(cudaMemsetAsync(s->spec_dev, 0, sizeof(SpecState), st));
  for (int i = 0; i < 2; ++i) B200_CUDA(cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming));
  B200_CUDA(cudaMemsetAsync(s->gridbar, 0, 2 * sizeof(unsigned), st));
  B200_TRY(lbfgs_init_state(s->view, s->m, s->mod, st));
  s->apply_blocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)4 * ctx->num_sms, (s->ld / 4 + 255) / 256));
  *out = s;
  return B200_OK;
}

static void lbfgs_drop_graphs(b200_lbfgs *s) {
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b)
      for (int c = 0; c < 2; ++c) {
        if (s->graph[a][b][c]) cudaGraphExecDestroy(s->graph[a][b][c]);
        s->graph[a][b][c] = nullptr;
      }
}

static void lbfgs_free(void *p) {
  b200_lbfgs *s = (b200_lbfgs *)p;
  lbfgs_drop_graphs(s);
  cudaFree(s->ws);
  if (s->gridbar) cudaFree(s->gridbar);
  if (s->spec_dev) cudaFree(s->spec_dev);
  for (int i = 0; i < 2; ++i) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
  delete s;
}

int b200_lbfgs_destroy(b200_lbfgs *s) {
  if (!s) return B200_OK;
  b200_ctx *ctx = s->ctx;
  if (!ctx_is_live(ctx)) { // the context went first (and with it the stream): nothing to park it in
    cudaDeviceSynchronize();
    lbfgs_free(s);
    return B200_OK;
  }
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (s->last_net_uid) net_xq_release_solver((b200_net *)net_lookup(s->last_net_uid)); // end of the minimisation: the caller may refill x
  s->last_net_uid = 0;
  if (ctx->lbfgs_pool.size() < 4) { // park it: the next solver of this shape skips cudaMalloc / cudaFree / graph instantiation
    ctx->lbfgs_pool_free = lbfgs_free;
    ctx->lbfgs_pool.push_back(s);
    return B200_OK;
  }
  lbfgs_free(s);
  return B200_OK;
}

} // extern "C"
namespace b200 {
void lbfgs_pool_forget_net(b200_ctx *ctx, unsigned long long net_uid) {
  for (void *p : ctx->lbfgs_pool) {
    b200_lbfgs *c = (b200_lbfgs *)p;
    if (c->gkey[0] == net_uid) {
      lbfgs_drop_graphs(c);
      memset(c->gkey, 0, sizeof(c->gkey));
    }
    if (c->last_net_uid == net_uid) c->last_net_uid = 0;
  }
}
} // namespace b200
extern "C" {

static int lbfgs_run_sharded(b200_lbfgs *s, b200_net *net, float *params, const float *input, const float *target, int batch,
                             int iters, b200_history *hist);

int b200_lbfgs_run(b200_lbfgs *s, b200_net *net, b200_loss_grad_fn fn, void *user, float *params, const float *input,
                   const float *target, int batch, int iters, b200_history *hist) {
  B200_REQUIRE(s && params, "null argument");
  if (hist) { hist->size = 0; hist->iterations = 0; hist->evaluations = 0; hist->launches = 0; }
  if (s->sharded) {
    B200_REQUIRE(net, "the sharded-history mode needs the library's network objective");
    return lbfgs_run_sharded(s, net, params, input, target, batch, iters, hist);
  }
  B200_REQUIRE(net || fn, "either a network or a loss_grad callback is required");
  B200_REQUIRE(!net || s->N == net->n, "n does not match the network's parameter count");
  b200_ctx *ctx = s->ctx;
  const b200_lbfgs_opts &o = s->o;
  B200_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const long launches0 = b200_launch_count();
  const size_t N = s->N, ld = s->ld;
  const int m = s->m, mp = s->mp;
  const bool wolfe = (s->policy == POLICY_WOLFE);
  float *p = s->p, *x_prev = s->x_prev, *S = s->S, *Y = s->Y;
  double *dot_part = s->dot_part;

  HostMail *mail = (HostMail *)ctx->h_scalars;
  Objective obj{ctx, net, fn, user, input, target, batch, 0 /* 1/B: net_eval resolves it (b200_net_set_global_batch, else shard batch x ranks) */};
  obj.d_part = dot_part;
  obj.n = N;
  Timer timer{ctx, hist != nullptr && o.record_timing != 0};

  double cb_loss = 0.0;
  // uint8 copy of 8-bit pixel inputs (no-op unless x == u/255); re-derived at the start of every minimisation: the caller may
  // have refilled the same device buffer
  if (net && net->prec != B200_PREC_FP32) {
    B200_TRY(net_quantize_input(net, input, batch, !s->started));
    s->last_net_uid = net->uid;
  }
  if (!s->started) {
    // loss = loss_grad(params, grad, ...)   lbfgs.cuh:78 / lbfgs.hpp:44
    B200_TRY(obj.eval_async(params, s->gbuf[0], mail, &cb_loss));
    B200_CUDA(cudaStreamSynchronize(st));
    s->loss = net ? mail->loss : cb_loss;
    s->gnorm = std::sqrt(mail->gnorm2);
    s->started = true;
    s->cur = 0; s->iter = 0; s->reset_next = 0;
  }

  int iterations_done = 0;
  const int max_ls = o.max_line_iters;
  for (int it = 0; it < iters; ++it) {
    const int iter = s->iter;
    NvtxRange nvtx_iter("lbfgs_iteration");
    B200_TRY(timer.start());
    if (s->gnorm < (double)o.tol) break; // lbfgs.cuh:92-93 / lbfgs.hpp:53-55
    float *g = s->gbuf[s->cur], *g_new = s->gbuf[s->cur ^ 1];
    const double loss = s->loss;

    // ---- direction: pair formation of the previous step + two-loop + first trial point, 3 launches ----
    const int mode = (iter > 0 && m > 0) ? DOTS_FORM_PAIR : DOTS_NONE;
    // direction of the iteration whose current gradient is gbuf[par] (the other buffer holds the previous gradient, then receives
    // the trial evaluation's); mailbox mb
    auto issue_direction_for = [&](int par, int dmode, int reset, int first, HostMail *mb, const SpecState *sst, int spec,
                                   bool zero_copy = false) -> int {
      float *gg = s->gbuf[par], *gn = s->gbuf[par ^ 1];
      DotsArgs da{S, Y, N, ld, s->view, gg, params, x_prev, gn, dmode, reset, 0, s->partials};
      SolveArgs sa{s->view, s->partials, s->nblk, dmode, reset, s->policy, first, 0, 0.0, 0, 0, zero_copy ? &mb->hdr : nullptr};
      ApplyArgs aa{S, Y, N, ld, s->view, gg, p, params, x_prev, 1.0, 0.0f, nullptr};
      bool fused = false;
      if (ctx->prof.on) B200_TRY(launch_prof_spacer(st)); // the events below are then enqueued behind running work, not on an idle GPU
      if (!env().no_fused_direction) { // one launch: dots -> grid barrier -> solve (every CTA) -> apply
        ProfScope ps(ctx, "lbfgs_direction");
        B200_TRY(launch_lbfgs_direction(ctx, da, sa, aa, mp, s->nblk, s->gridbar, st, &fused, sst, spec));
      }
      if (!fused) {
        B200_REQUIRE(!spec, "speculative launch needs the fused direction kernel");
        {
          ProfScope ps(ctx, "lbfgs_dots");
          B200_TRY(launch_lbfgs_dots(da, mp, s->nblk, st));
        }
        {
          ProfScope ps(ctx, "lbfgs_solve");
          B200_TRY(launch_lbfgs_solve(sa, mp, st));
        }
        {
          ProfScope ps(ctx, "lbfgs_apply");
          B200_TRY(launch_lbfgs_apply(aa, s->apply_blocks, st));
        }
      }
      if (!(zero_copy && fused)) B200_CUDA(cudaMemcpyAsync(&mb->hdr, s->view.h, sizeof(LbfgsHeader), cudaMemcpyDeviceToHost, st));
      return B200_OK;
    };
    auto issue_direction = [&]() -> int { return issue_direction_for(s->cur, mode, s->reset_next, iter == 0 ? 1 : 0, mail, nullptr, 0); };
    // Steady-state iterations of the network objective replay a captured graph: direction + first trial evaluation.
    bool first_eval_issued = false, spec_in_flight = false;
    HostMail *mm = mail; // where this iteration's first-trial results land
    // several GPUs: only when the gradient all-reduce runs over peer memory (plain kernels; no NCCL call inside the capture)
    const bool comm_ok = ctx->world == 1 || (ctx->p2p.ready && N <= ctx->p2p.slot_floats);
    const bool graphable = net && !wolfe && s->graphs_ok && mode == DOTS_FORM_PAIR && comm_ok && !ctx->prof.on &&
                           max_ls > 0 && !env().no_graph && !env().tc_timing;
    // graph of the iteration with gradient parity `par`: capture on first use, then launch
    auto launch_graph = [&](int par, int reset, int spec) -> int {
      cudaGraphExec_t &ge = s->graph[par][reset][spec];
      if (!ge) {
        const long l0 = b200_launch_count();
        cudaGraph_t gr = nullptr;
        const bool gate = s->spec_capable;
        net->spec_st = gate ? s->spec_dev : nullptr;
        net->spec_flag = spec;
        const bool zc = !env().no_zero_copy; // kernels write the pinned mailbox: no D2H copy nodes in the graph
        net->host_out = zc ? &mail[par].loss : nullptr;
        bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
          ok = issue_direction_for(par, DOTS_FORM_PAIR, reset, 0, mail + par, gate ? s->spec_dev : nullptr, spec, zc) == B200_OK &&
               obj.eval_async(params, s->gbuf[par ^ 1], mail + par, &cb_loss) == B200_OK;
          --obj.evals; // counted when the graph is replayed
          ok = (cudaStreamEndCapture(st, &gr) == cudaSuccess) && ok && gr != nullptr;
        }
        net->spec_st = nullptr;
        net->spec_flag = 0;
        net->host_out = nullptr;
        if (ok) ok = cudaGraphInstantiate(&ge, gr, 0) == cudaSuccess;
        if (gr) cudaGraphDestroy(gr);
        s->graph_launches[par][reset][spec] = b200_launch_count() - l0;
        g_launches.fetch_add(-(b200_launch_count() - l0), std::memory_order_relaxed); // captured, not executed
        if (!ok) {
          cudaGetLastError();
          ge = nullptr;
          s->graphs_ok = false; // fall back to plain launches for the rest of this solver's life
          return B200_OK;
        }
      }
      B200_CUDA(cudaGraphLaunch(ge, st));
      B200_CUDA(cudaEventRecord(s->ev[par], st));
      g_launches.fetch_add(s->graph_launches[par][reset][spec], std::memory_order_relaxed);
      ++obj.evals;
      return B200_OK;
    };
    if (graphable) {
      const unsigned long long key[7] = {net->uid, (unsigned long long)(uintptr_t)params, (unsigned long long)(uintptr_t)input,
                                         (unsigned long long)(uintptr_t)target, (unsigned long long)batch,
                                         (unsigned long long)net->config_gen, (unsigned long long)ctx->p2p.gen};
      if (memcmp(key, s->gkey, sizeof(key)) != 0) {
        lbfgs_drop_graphs(s);
        memcpy(s->gkey, key, sizeof(key));
        s->ahead = false;
        // the evaluation consists of gated kernels only (fp16 layer-0 GEMMs + one-pass last layer + fused direction)?
        s->spec_capable = !env().no_speculation && !env().no_fused_direction &&
                          mp <= kRowsPerLaunch && net_spec_capable(net, input, batch);
        if (s->spec_capable) {
          SpecState hs{};
          hs.gate_next = 0; hs.loss_prev = loss; hs.c1 = (double)o.c1; hs.tol = (double)o.tol;
          hs.alpha0 = &s->view.h->alpha0; hs.gdotp = &s->view.h->gdotp;
          B200_CUDA(cudaMemcpyAsync(s->spec_dev, &hs, sizeof(hs), cudaMemcpyHostToDevice, st));
          B200_CUDA(cudaStreamSynchronize(st)); // hs is a stack object
        }
      }
      if (s->ahead) { // launched behind the previous iteration and confirmed by its result
        first_eval_issued = true;
      } else {
        if (s->spec_capable) { // the device-side decision of THIS iteration starts from the host's loss
          double *slot = ctx->h_scalars + 62;
          *slot = loss;
          B200_CUDA(cudaMemcpyAsync(&s->spec_dev->loss_prev, slot, sizeof(double), cudaMemcpyHostToDevice, st));
        }
        const long e0 = obj.evals;
        B200_TRY(launch_graph(s->cur, s->reset_next, 0));
        first_eval_issued = obj.evals > e0;
      }
      s->ahead = false;
      if (first_eval_issued) {
        mm = mail + s->cur;
        // speculate: the next iteration's graph goes in right behind this one; its kernels return at once unless the device finds
        // this iteration's first trial point accepted (Armijo) and not converged
        if (s->spec_capable && s->graphs_ok && it + 1 < iters && m > 0) {
          const long e0 = obj.evals;
          B200_TRY(launch_graph(s->cur ^ 1, 0, 1));
          spec_in_flight = obj.evals > e0;
        }
      }
    }
    if (!first_eval_issued) B200_TRY(issue_direction());
    s->reset_next = 0;

    // ---- line search ------------------------------------------------------------------------------
    double loss_new = 0.0, gnorm2_new = 0.0, alpha = 1.0;
    if (!wolfe) {
      // Armijo backtracking with safeguarded quadratic interpolation (lbfgs.cuh:106-147)
      bool armijo_ok = false, evaluated = false;
      double gdotp = 0.0;
      for (int ls = 0; ls < max_ls; ++ls) {
        if (ls > 0) B200_TRY(launch_trial_point(N, x_prev, (float)alpha, p, params, st)); // lbfgs.cuh:116-117
        HostMail *mr = (ls == 0) ? mm : mail;
        if (!(ls == 0 && first_eval_issued)) B200_TRY(obj.eval_async(params, g_new, mail, &cb_loss));
        if (ls == 0 && spec_in_flight) B200_CUDA(cudaEventSynchronize(s->ev[s->cur])); // only THIS iteration's graph
        else B200_CUDA(cudaStreamSynchronize(st));
        B200_TRY(ctx_check_device_error(ctx));
        if (ls == 0) { alpha = (double)(float)mr->hdr.alpha0; gdotp = mr->hdr.gdotp; }
        loss_new = net ? mr->loss : cb_loss;
        gnorm2_new = mr->gnorm2;
        evaluated = true;
        if (loss_new <= loss + (double)o.c1 * alpha * gdotp) {
          armijo_ok = true;
          // the speculative graph runs exactly when the device took this branch and the run has not converged
          if (ls == 0 && spec_in_flight) {
            if (std::sqrt(gnorm2_new) < (double)o.tol) --obj.evals; else s->ahead = true;
            spec_in_flight = false;
          }
          break;
        }
        if (ls == 0 && spec_in_flight) { --obj.evals; spec_in_flight = false; } // gated off on the device: empty launches
        const double denom = 2.0 * (loss_new - loss - gdotp * alpha);
        bool fallback = true;
        if (std::fabs(denom) > 1e-20) {
          const double na = -(gdotp * alpha * alpha) / denom;
          if (na >= 0.1 * alpha && na <= 0.9 * alpha) { alpha = na; fallback = false; }
        }
        if (fallback) alpha *= (double)o.rho;
      }
      if (!evaluated) { // max_line_iters == 0 (lbfgs.cuh:142-145): the fused first trial point stands
        B200_TRY(obj.eval_async(params, g_new, mail, &cb_loss));
        B200_CUDA(cudaStreamSynchronize(st));
        loss_new = net ? mail->loss : cb_loss;
        gnorm2_new = mail->gnorm2;
        armijo_ok = true;
      }
      if (!armijo_ok) s->reset_next = 1; // lbfgs.cuh:147 — the step is still accepted
    } else {
      // weak-Wolfe bisection / expansion (full_batch_minimizer.hpp:126-157); iteration 0 takes
      // alpha = min(1, 1/||g||) without a line search (lbfgs.hpp:60-63)
      B200_TRY(obj.eval_async(params, g_new, mail, &cb_loss));
      if (iter > 0) {
        B200_TRY(launch_dot(ctx, g_new, p, N, dot_part, ctx->d_scalars + 1));
        B200_CUDA(cudaMemcpyAsync(&mail->gnew_dot_p, ctx->d_scalars + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
      }
      B200_CUDA(cudaStreamSynchronize(st));
      loss_new = net ? mail->loss : cb_loss;
      gnorm2_new = mail->gnorm2;
      alpha = (double)(float)mail->hdr.alpha0;
      if (iter > 0) {
        const double f_old = loss, gfo = mail->hdr.gdotp, inf = std::numeric_limits<double>::infinity();
        double a_min = 0.0, a_max = inf;
        bool at_alpha = true; // params / g_new currently hold the point for `alpha`
        const int trials = o.max_line_iters > 0 ? o.max_line_iters : 50;
        for (int i = 0; i < trials; ++i) {
          if (!at_alpha) {
            B200_TRY(launch_trial_point(N, x_prev, (float)alpha, p, params, st));
            B200_TRY(obj.eval_async(params, g_new, mail, &cb_loss));
            B200_TRY(launch_dot(ctx, g_new, p, N, dot_part, ctx->d_scalars + 1));
            B200_CUDA(cudaMemcpyAsync(&mail->gnew_dot_p, ctx->d_scalars + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            loss_new = net ? mail->loss : cb_loss;
            gnorm2_new = mail->gnorm2;
            at_alpha = true;
          }
          if (loss_new > f_old + (double)o.c1 * alpha * gfo) {
            a_max = alpha;
            alpha = (double)o.rho * (a_min + a_max);
            at_alpha = false;
            continue;
          }
          if (mail->gnew_dot_p < (double)o.c2 * gfo) {
            a_min = alpha;
            alpha = (a_max == inf) ? alpha * 2.0 : (double)o.rho * (a_min + a_max);
            at_alpha = false;
            continue;
          }
          break;
        }
        if (!at_alpha) { // trials exhausted: the reference steps with the last (unevaluated) alpha (:156, lbfgs.hpp:67-70)
          B200_TRY(launch_trial_point(N, x_prev, (float)alpha, p, params, st));
          B200_TRY(obj.eval_async(params, g_new, mail, &cb_loss));
          B200_CUDA(cudaStreamSynchronize(st));
          loss_new = net ? mail->loss : cb_loss;
          gnorm2_new = mail->gnorm2;
        }
      }
    }

    // accept: g <- g_new, loss <- loss_new (lbfgs.cuh:171-175); s, y are formed by the next dots pass
    s->cur ^= 1;
    s->loss = loss_new;
    s->gnorm = std::sqrt(gnorm2_new);
    s->iter = iter + 1;
    B200_TRY(timer.stop(!s->ahead));
    record(hist, iterations_done, s->loss, s->gnorm, timer.elapsed);
    ++iterations_done;
  }
  B200_CUDA(cudaStreamSynchronize(st));
  if (hist) {
    hist->iterations = iterations_done;
    hist->evaluations = obj.evals;
    hist->launches = b200_launch_count() - launches0;
  }
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------------
// Sharded-history L-BFGS (multi-GPU, large n — BASELINE configs[4]): rank r owns the parameter slice
// [lo, lo+len) of x, g, p and of every history vector. Per evaluation: local full-length gradient ->
// reduce-scatter to the owners; per direction: ONE all-reduce of the 5(m+1)+1 partial dot products;
// per trial point: all-gather of the parameters. Same control flow as the replicated Armijo path.
// ---------------------------------------------------------------------------------------------------
static int lbfgs_run_sharded(b200_lbfgs *s, b200_net *net, float *params, const float *input, const float *target, int batch,
                             int iters, b200_history *hist) {
  b200_ctx *ctx = s->ctx;
  const b200_lbfgs_opts &o = s->o;
  B200_REQUIRE(s->N == net->n, "n does not match the network's parameter count");
  B200_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const long launches0 = b200_launch_count();
  const size_t N = s->N, len = s->len, lo = s->lo, ld = s->ld;
  const int m = s->m, mp = s->mp, ncols = kDotsCols * mp + 1;
  float *x = params + lo; // this rank's slice of the caller's parameter buffer
  HostMail *mail = (HostMail *)ctx->h_scalars;
  Timer timer{ctx, hist != nullptr && o.record_timing != 0};
  long evals = 0;
  if (net->prec != B200_PREC_FP32) {
    B200_TRY(net_quantize_input(net, input, batch, !s->started));
    s->last_net_uid = net->uid;
  }
  struct Defer { b200_net *n; ~Defer() { n->defer_reduce = false; } } defer{net};
  net->defer_reduce = true;

  // evaluation at the (replicated) full parameter vector; the owner's gradient slice lands in g_shard
  auto eval = [&](float *g_shard) -> int {
    ++evals;
    EvalOut *eo = (EvalOut *)net->eval_out;
    B200_TRY(net_eval(net, params, input, target, batch, 0, s->gfull, eo)); // partial loss, un-reduced gradient
    {
      ProfScope ps(ctx, "reduce_scatter");
      B200_TRY(ctx_reduce_shards(ctx, s->gfull, g_shard, N, s->chunk));
    }
    B200_TRY(launch_dot(ctx, g_shard, g_shard, len, s->dot_part, &eo->gnorm2)); // ||g_shard||^2 next to the partial loss
    B200_TRY(ctx_allreduce_f64(ctx, (double *)eo, 2));
    B200_CUDA(cudaMemcpyAsync(&mail->loss, eo, sizeof(EvalOut), cudaMemcpyDeviceToHost, st));
    return B200_OK;
  };

  if (!s->started) {
    B200_TRY(eval(s->gbuf[0]));
    B200_CUDA(cudaStreamSynchronize(st));
    s->loss = mail->loss;
    s->gnorm = std::sqrt(mail->gnorm2);
    s->started = true;
    s->cur = 0; s->iter = 0; s->reset_next = 0;
  }
  int iterations_done = 0;
  for (int it = 0; it < iters; ++it) {
    const int iter = s->iter;
    B200_TRY(timer.start());
    if (s->gnorm < (double)o.tol) break;
    float *g = s->gbuf[s->cur], *g_new = s->gbuf[s->cur ^ 1];
    const double loss = s->loss;
    const int mode = (iter > 0 && m > 0) ? DOTS_FORM_PAIR : DOTS_NONE;
    {
      ProfScope ps(ctx, "lbfgs_dots");
      DotsArgs da{s->S, s->Y, len, ld, s->view, g, x, s->x_prev, g_new, mode, s->reset_next, 0, s->partials};
      B200_TRY(launch_lbfgs_dots(da, mp, s->nblk, st));
      B200_TRY(launch_reduce_partials(s->partials, s->nblk, ncols, s->totals, st));
    }
    {
      ProfScope ps(ctx, "dots_allreduce");
      B200_TRY(ctx_allreduce_f64(ctx, s->totals, (size_t)ncols)); // the only exchange of the direction: 5(m+1)+1 doubles
    }
    {
      ProfScope ps(ctx, "lbfgs_solve");
      SolveArgs sa{s->view, s->totals, 1, mode, s->reset_next, s->policy, iter == 0 ? 1 : 0, 0, 0.0, 0};
      B200_TRY(launch_lbfgs_solve(sa, mp, st));
    }
    {
      ProfScope ps(ctx, "lbfgs_apply");
      ApplyArgs aa{s->S, s->Y, len, ld, s->view, g, s->p, x, s->x_prev, 1.0, 0.0f, nullptr};
      B200_TRY(launch_lbfgs_apply(aa, s->apply_blocks, st));
    }
    B200_CUDA(cudaMemcpyAsync(&mail->hdr, s->view.h, sizeof(LbfgsHeader), cudaMemcpyDeviceToHost, st));
    s->reset_next = 0;

    double loss_new = 0.0, gnorm2_new = 0.0, alpha = 1.0, gdotp = 0.0;
    bool armijo_ok = false;
    const int max_ls = std::max(1, o.max_line_iters);
    for (int ls = 0; ls < max_ls; ++ls) {
      if (ls > 0) B200_TRY(launch_trial_point(len, s->x_prev, (float)alpha, s->p, x, st));
      {
        ProfScope ps(ctx, "allgather");
        B200_TRY(ctx_allgather_shards(ctx, params, N, s->chunk));
      }
      B200_TRY(eval(g_new));
      B200_CUDA(cudaStreamSynchronize(st));
      if (ls == 0) { alpha = (double)(float)mail->hdr.alpha0; gdotp = mail->hdr.gdotp; }
      loss_new = mail->loss;
      gnorm2_new = mail->gnorm2;
      if (loss_new <= loss + (double)o.c1 * alpha * gdotp) { armijo_ok = true; break; }
      const double denom = 2.0 * (loss_new - loss - gdotp * alpha);
      bool fallback = true;
      if (std::fabs(denom) > 1e-20) {
        const double na = -(gdotp * alpha * alpha) / denom;
        if (na >= 0.1 * alpha && na <= 0.9 * alpha) { alpha = na; fallback = false; }
      }
      if (fallback) alpha *= (double)o.rho;
    }
    if (!armijo_ok) s->reset_next = 1;
    s->cur ^= 1;
    s->loss = loss_new;
    s->gnorm = std::sqrt(gnorm2_new);
    s->iter = iter + 1;
    B200_TRY(timer.stop());
    record(hist, iterations_done, s->loss, s->gnorm, timer.elapsed);
    ++iterations_done;
  }
  B200_CUDA(cudaStreamSynchronize(st));
  if (hist) {
    hist->iterations = iterations_done;
    hist->evaluations = evals;
    hist->launches = b200_launch_count() - launches0;
  }
  return B200_OK;
}

int b200_lbfgs_solve(b200_ctx *ctx, b200_net *net, b200_loss_grad_fn fn, void *user, int n, float *params,
                     const float *input, const float *target, int batch, const b200_lbfgs_opts *opts,
                     b200_history *hist) {
  B200_REQUIRE(ctx, "null ctx");
  if (hist) { hist->size = 0; hist->iterations = 0; hist->evaluations = 0; hist->launches = 0; }
  if (n <= 0 || params == nullptr) return B200_OK; // lbfgs.cuh:45-48: silent return, iterations() == 0
  b200_lbfgs *s = nullptr;
  B200_TRY(b200_lbfgs_create(ctx, n, opts, &s));
  const int status = b200_lbfgs_run(s, net, fn, user, params, input, target, batch, s->o.max_iters, hist);
  b200_lbfgs_destroy(s); // (also drops the input cache this solve built: the caller may change x after the solve)
  return status;
}

// ===================================================================================================
// CudaGD::solve (src/cuda/gd.cuh:38-106)
// ===================================================================================================
int b200_gd_solve(b200_ctx *ctx, b200_net *net, b200_loss_grad_fn fn, void *user, int n, float *params,
                  const float *input, const float *target, int batch, const b200_gd_opts *opts, b200_history *hist) {
  B200_REQUIRE(ctx, "null ctx");
  if (hist) { hist->size = 0; hist->iterations = 0; hist->evaluations = 0; hist->launches = 0; }
  if (n <= 0 || params == nullptr) return B200_OK;
  B200_REQUIRE(net || fn, "either a network or a loss_grad callback is required");
  B200_REQUIRE(!net || (size_t)n == net->n, "n does not match the network's parameter count");
  b200_gd_opts o;
  if (opts) o = *opts; else b200_gd_default_opts(&o);
  B200_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const long launches0 = b200_launch_count();
  const size_t N = (size_t)n;
  char *ws = nullptr;
  const size_t dotp_bytes = (sizeof(double) * (size_t)dot_blocks(ctx, N) + 255) & ~size_t(255);
  B200_CUDA(cudaMalloc(&ws, dotp_bytes + 2 * sizeof(float) * N));
  struct Free { char *p; ~Free() { cudaFree(p); } } free_ws{ws};
  double *dot_part = (double *)ws;
  float *grad = (float *)(ws + dotp_bytes), *vel = grad + N;
  if (o.momentum > 0.0f) B200_CUDA(cudaMemsetAsync(vel, 0, sizeof(float) * N, st)); // gd.cuh:54-56

  HostMail *mail = (HostMail *)ctx->h_scalars;
  Objective obj{ctx, net, fn, user, input, target, batch, 0 /* 1/B: net_eval resolves it (b200_net_set_global_batch, else shard batch x ranks) */};
  obj.d_part = dot_part;
  obj.n = N;
  Timer timer{ctx, hist != nullptr && o.record_timing != 0};
  double cb_loss = 0.0;
  if (net && net->prec != B200_PREC_FP32) B200_TRY(net_quantize_input(net, input, batch, true));
  struct ClearQ { b200_net *n; ~ClearQ() { net_xq_release_solver(n); } } clear_q{net};
  B200_TRY(obj.eval_async(params, grad, mail, &cb_loss));
  B200_CUDA(cudaStreamSynchronize(st));
  double loss = net ? mail->loss : cb_loss;
  double gnorm = std::sqrt(mail->gnorm2);
  int iterations_done = 0;
  for (int iter = 0; iter < o.max_iters; ++iter) {
    B200_TRY(timer.start());
    if (gnorm < (double)o.tol) break; // gd.cuh:70-71
    if (o.momentum > 0.0f) B200_TRY(launch_momentum_step(N, o.momentum, o.lr, grad, vel, params, st)); // :73-81
    else B200_TRY(launch_axpy(N, -o.lr, grad, params, st));                                             // :83-84
    B200_TRY(obj.eval_async(params, grad, mail, &cb_loss));
    B200_CUDA(cudaStreamSynchronize(st));
    B200_TRY(ctx_check_device_error(ctx));
    loss = net ? mail->loss : cb_loss;
    gnorm = std::sqrt(mail->gnorm2);
    B200_TRY(timer.stop());
    record(hist, iterations_done, loss, gnorm, timer.elapsed);
    ++iterations_done;
  }
  if (hist) {
    hist->iterations = iterations_done;
    hist->evaluations = obj.evals;
    hist->launches = b200_launch_count() - launches0;
  }
  return B200_OK;
}

// ===================================================================================================
// CudaSGD::solve (src/cuda/sgd.cuh:50-153): sequential, unshuffled mini-batches by pointer offset
// ===================================================================================================
int b200_sgd_solve(b200_ctx *ctx, b200_net *net, b200_loss_grad_fn fn, void *user, int n, float *params,
                   const float *input, const float *target, int total_samples, const b200_sgd_opts *opts,
                   b200_history *hist) {
  B200_REQUIRE(ctx, "null ctx");
  if (hist) { hist->size = 0; hist->iterations = 0; hist->evaluations = 0; hist->launches = 0; }
  if (n <= 0 || params == nullptr) return B200_OK;
  B200_REQUIRE(net || fn, "either a network or a loss_grad callback is required");
  B200_REQUIRE(!net || (size_t)n == net->n, "n does not match the network's parameter count");
  b200_sgd_opts o;
  if (opts) o = *opts; else b200_sgd_default_opts(&o);
  if (o.input_dim == 0 || o.output_dim == 0) { // sgd.cuh:61-65
    fprintf(stderr, "Error: Dimensions not set for SGD\n");
    return B200_OK;
  }
  B200_REQUIRE(o.batch_size > 0, "batch_size must be positive");
  B200_REQUIRE(ctx->world == 1, "CudaSGD is single-GPU (sequential mini-batches)");
  if (o.sampling == 1) {
    B200_REQUIRE(net, "random mini-batches need the library's network objective (rows are gathered on the device)");
    return sgd_random_solve(ctx, net, n, params, input, target, total_samples, o, hist);
  }
  B200_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const long launches0 = b200_launch_count();
  const size_t N = (size_t)n;
  char *ws = nullptr;
  const size_t dotp_bytes = (sizeof(double) * (size_t)dot_blocks(ctx, N) + 255) & ~size_t(255);
  B200_CUDA(cudaMalloc(&ws, dotp_bytes + 2 * sizeof(float) * N));
  struct Free { char *p; ~Free() { cudaFree(p); } } free_ws{ws};
  double *dot_part = (double *)ws;
  float *grad = (float *)(ws + dotp_bytes), *vel = grad + N;
  if (o.momentum > 0.0f) B200_CUDA(cudaMemsetAsync(vel, 0, sizeof(float) * N, st));

  HostMail *mail = (HostMail *)ctx->h_scalars;
  Objective full{ctx, net, fn, user, input, target, total_samples, (long)total_samples};
  full.d_part = dot_part;
  full.n = N;
  Timer timer{ctx, hist != nullptr && o.record_timing != 0};
  long evals = 0;
  double cb_loss = 0.0;
  if (net && net->prec != B200_PREC_FP32) B200_TRY(net_quantize_input(net, input, total_samples, true));
  struct ClearQ { b200_net *n; ~ClearQ() { net_xq_release_solver(n); } } clear_q{net};
  float current_lr = o.lr;
  const int num_batches = (total_samples + o.batch_size - 1) / o.batch_size;
  float prev_epoch_loss_avg = std::numeric_limits<float>::infinity();
  int iterations_done = 0;
  if (hist) { // sgd.cuh:89-94: initial full-batch record at index 0
    B200_TRY(full.eval_async(params, grad, mail, &cb_loss));
    B200_CUDA(cudaStreamSynchronize(st));
    record(hist, iterations_done, net ? mail->loss : cb_loss, std::sqrt(mail->gnorm2), 0.f);
    ++iterations_done;
  }
  for (int iter = 0; iter < o.max_iters; ++iter) {
    B200_TRY(timer.start());
    if (o.decay_step > 0 && iter > 0 && iter % o.decay_step == 0) current_lr *= o.decay_rate; // :97-99
    float epoch_loss_sum = 0.0f;
    for (int b = 0; b < num_batches; ++b) {
      const int start_idx = b * o.batch_size;
      const int cbs = std::min(o.batch_size, total_samples - start_idx);
      Objective mb{ctx, net, fn, user, input + (size_t)start_idx * o.input_dim, target + (size_t)start_idx * o.output_dim,
                   cbs, (long)cbs};
      mb.d_part = dot_part;
      mb.n = N;
      B200_TRY(mb.eval_async(params, grad, mail, &cb_loss));
      ++evals;
      if (o.momentum > 0.0f) B200_TRY(launch_momentum_step(N, o.momentum, current_lr, grad, vel, params, st));
      else B200_TRY(launch_axpy(N, -current_lr, grad, params, st));
      // the epoch-average needs every mini-batch loss on the host (sgd.cuh:124); the copy was queued by
      // eval_async, so this sync also covers the parameter update
      B200_CUDA(cudaStreamSynchronize(st));
      const float bl = (float)(net ? mail->loss : cb_loss);
      epoch_loss_sum += bl * (float)cbs;
    }
    const float epoch_loss_avg = epoch_loss_sum / (float)total_samples;
    if (o.tol > 0.0f && std::isfinite(prev_epoch_loss_avg)) { // :128-133
      const float denom = std::max(1.0f, std::fabs(prev_epoch_loss_avg));
      const float rel = std::fabs(prev_epoch_loss_avg - epoch_loss_avg) / denom;
      if (rel < o.tol) break;
    }
    prev_epoch_loss_avg = epoch_loss_avg;
    if (hist) { // :136-147
      B200_TRY(full.eval_async(params, grad, mail, &cb_loss));
      B200_CUDA(cudaStreamSynchronize(st));
      B200_TRY(timer.stop());
      record(hist, iterations_done, net ? mail->loss : cb_loss, std::sqrt(mail->gnorm2), timer.elapsed);
    }
    ++iterations_done;
  }
  B200_CUDA(cudaStreamSynchronize(st));
  if (hist) {
    hist->iterations = iterations_done;
    hist->evaluations = evals + full.evals;
    hist->launches = b200_launch_count() - launches0;
  }
  return B200_OK;
}

// ===================================================================================================
// building blocks
// ===================================================================================================
int b200_lbfgs_direction(b200_ctx *ctx, size_t n, int k, const float *S_dev, const float *Y_dev, const float *rho_host,
                         const float *g_dev, int policy, float *p_dev, double *g_dot_p_host) {
  B200_REQUIRE(ctx && g_dev && p_dev, "null argument");
  B200_REQUIRE(k >= 0 && k <= kMaxSlots - 1, "k must be in [0, 256]");
  B200_REQUIRE(k == 0 || (S_dev && Y_dev && rho_host), "history arrays required");
  B200_REQUIRE(policy >= 0 && policy <= 2, "unknown policy");
  B200_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int m = std::max(k, 1), mp = m + 1;
  const int nblk = lbfgs_dots_blocks(ctx, n);
  const int ncols = kDotsCols * mp + 1;
  const size_t state_bytes = lbfgs_state_bytes(m);
  char *ws = nullptr;
  B200_CUDA(cudaMalloc(&ws, state_bytes + sizeof(double) * (size_t)nblk * ncols));
  struct Free { char *p; ~Free() { cudaFree(p); } } free_ws{ws};
  LbfgsView view = lbfgs_view(ws, m);
  double *partials = (double *)(ws + state_bytes);
  B200_TRY(lbfgs_init_state(view, m, mp, st));
  // caller's arrays are k x n row-major in logical order: slot i = row i, ring modulus m+1 never wraps
  float *S = const_cast<float *>(S_dev), *Y = const_cast<float *>(Y_dev);
  if (k == 0) {
    DotsArgs da{S, Y, n, n, view, g_dev, nullptr, nullptr, nullptr, DOTS_NONE, 0, 0, partials};
    B200_TRY(launch_lbfgs_dots(da, mp, nblk, st));
    SolveArgs sa{view, partials, nblk, DOTS_NONE, 0, policy, 0, 0, 0.0, 0};
    B200_TRY(launch_lbfgs_solve(sa, mp, st));
  }
  for (int w = 0; w < k; ++w) { // build the Gram blocks one slot at a time (head == w before each pass)
    DotsArgs da{S, Y, n, n, view, g_dev, nullptr, nullptr, nullptr, DOTS_PAIR_IN_SLOT, 0, 0, partials};
    B200_TRY(launch_lbfgs_dots(da, mp, nblk, st));
    SolveArgs sa{view, partials, nblk, DOTS_PAIR_IN_SLOT, 0, policy, 0, 1, (double)rho_host[w], 0};
    B200_TRY(launch_lbfgs_solve(sa, mp, st));
  }
  const int apply_blocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)4 * ctx->num_sms, (n / 4 + 255) / 256));
  ApplyArgs aa{S, Y, n, n, view, g_dev, p_dev, nullptr, nullptr, policy == POLICY_SLBFGS ? -1.0 : 1.0, 0.0f, nullptr};
  B200_TRY(launch_lbfgs_apply(aa, apply_blocks, st));
  HostMail *mail = (HostMail *)ctx->h_scalars;
  B200_CUDA(cudaMemcpyAsync(&mail->hdr, view.h, sizeof(LbfgsHeader), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  if (g_dot_p_host) *g_dot_p_host = mail->hdr.gdotp * (policy == POLICY_SLBFGS ? -1.0 : 1.0);
  return B200_OK;
}

int b200_vec_dot(b200_ctx *ctx, const float *x, const float *y, size_t n, double *out_host) {
  B200_REQUIRE(ctx && x && y && out_host, "null argument");
  B200_CUDA(cudaSetDevice(ctx->device));
  double *part = nullptr;
  B200_CUDA(cudaMalloc(&part, sizeof(double) * (size_t)dot_blocks(ctx, n)));
  struct Free { double *p; ~Free() { cudaFree(p); } } f{part};
  B200_TRY(launch_dot(ctx, x, y, n, part, ctx->d_scalars));
  B200_CUDA(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  *out_host = ctx->h_scalars[0];
  return B200_OK;
}
int b200_vec_nrm2(b200_ctx *ctx, const float *x, size_t n, double *out_host) {
  B200_TRY(b200_vec_dot(ctx, x, x, n, out_host));
  *out_host = std::sqrt(*out_host);
  return B200_OK;
}
int b200_vec_axpy(b200_ctx *ctx, size_t n, float alpha, const float *x, float *y) {
  B200_REQUIRE(ctx && x && y, "null argument");
  B200_CUDA(cudaSetDevice(ctx->device));
  return launch_axpy(n, alpha, x, y, ctx->stream);
}
int b200_vec_scal(b200_ctx *ctx, size_t n, float alpha, float *x) {
  B200_REQUIRE(ctx && x, "null argument");
  B200_CUDA(cudaSetDevice(ctx->device));
  return launch_scal(n, alpha, x, ctx->stream);
}
int b200_vec_trial_point(b200_ctx *ctx, size_t n, const float *x0, float alpha, const float *p, float *y) {
  B200_REQUIRE(ctx && x0 && p && y, "null argument");
  B200_CUDA(cudaSetDevice(ctx->device));
  return launch_trial_point(n, x0, alpha, p, y, ctx->stream);
}
int b200_convert_f64_to_f32(b200_ctx *ctx, const double *src_dev, float *dst_dev, size_t n) {
  B200_REQUIRE(ctx && src_dev && dst_dev, "null argument");
  B200_CUDA(cudaSetDevice(ctx->device));
  return launch_f64_to_f32(n, src_dev, dst_dev, ctx->stream);
}

} // extern "C"

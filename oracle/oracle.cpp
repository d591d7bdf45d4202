// ORACLE — TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH.
//
// CPU restatement (fp64) of the reference's optimizer-plus-objective path, used
// only as the checker by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs. Nothing under lbfgs_ffnn_b200/ may call it.
//
// Why C++ and not plain C: S-LBFGS parity needs bit-identical index streams, i.e.
// libstdc++'s std::mt19937 + std::uniform_int_distribution<size_t> and, for the
// initial parameters, std::normal_distribution<double/float> — the very objects the
// reference uses (src/minimizer/s_lbfgs.hpp:143-161, src/network.hpp:52-69).
//
// Parity pinning — PINNED to outputs of the reference's own code, two ways:
//  * CPU: the reference's CPU sources (src/unified_launcher.hpp and everything it includes) compile UNMODIFIED against an
//    Eigen-API stand-in (oracle/ref_cpu/; Eigen 3.4.0 is an un-vendored system dependency, enviroment/Dockerfile:14, absent from
//    this image; it is only used for dense double GEMM / dot / norm) into oracle/_ref/libref_cpu.so. Its outputs on seeded
//    inputs are committed as tests/golden/golden_ref_cpu.json (tests/golden/make_golden_ref_cpu.py) and
//    tests/test_oracle_vs_reference_cpu.py holds this file to them: objective 1e-13, weak-Wolfe L-BFGS trajectory, GD, SGD
//    (random mini-batches) and S-LBFGS parameters after every epoch 1e-8 .. 1e-6 (both sides fp64; summation order differs).
//  * CUDA: the reference's CUDA backend compiles as it is (oracle/ref_cuda/ -> oracle/_ref/libref_cuda.so); its outputs on a
//    B200 are tests/golden/golden_ref_cuda.json, checked by tests/test_oracle.py and, live, by tests/test_gpu_vs_reference_cuda.py.
// Further: the reference's analytic known-answer tests (tests/main.cpp Rosenbrock / Ackley / Rastrigin), an independent numpy
// fp64 restatement and central finite differences (tests/test_oracle.py).
//
// Layout contract (identical to the reference, src/layer.hpp:100-114,
// src/cuda/layer.cuh:48-58): matrices are column-major. X is in x B (sample b is the
// contiguous run X[b*in .. b*in+in)), W_l is out x in with ld = out, flat parameter
// vector = per layer [W (out*in) | b (out)], gradients mirror it.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <numeric>
#include <random>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using Vec = std::vector<double>;

// ----------------------------------------------------------------------------------
// small dense helpers (Eigen call sites: dot / norm / axpy in src/minimizer/*.hpp)
// ----------------------------------------------------------------------------------
double dot(const Vec &a, const Vec &b) {
  double s = 0.0;
  const size_t n = a.size();
  for (size_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
double norm(const Vec &a) { return std::sqrt(dot(a, a)); }
void axpy(double alpha, const Vec &x, Vec &y) {
  const size_t n = x.size();
  for (size_t i = 0; i < n; ++i) y[i] += alpha * x[i];
}

// ----------------------------------------------------------------------------------
// activations: src/layer.hpp:15-47 (apply/prime on the PRE-activation z)
// enum values follow src/cuda/kernels.cuh:53-58 {Linear=0,Tanh=1,ReLU=2,Sigmoid=3}
// ----------------------------------------------------------------------------------
enum Act { kLinear = 0, kTanh = 1, kReLU = 2, kSigmoid = 3 };

inline double act_apply(int a, double x) {
  switch (a) {
  case kTanh: return std::tanh(x);
  case kReLU: return (x > 0.0) ? x : 0.0;
  case kSigmoid: return 1.0 / (1.0 + std::exp(-x));
  default: return x;
  }
}
inline double act_prime(int a, double x) {
  switch (a) {
  case kTanh: { double t = std::tanh(x); return 1.0 - t * t; }
  case kReLU: return (x > 0.0) ? 1.0 : 0.0;
  case kSigmoid: { double s = 1.0 / (1.0 + std::exp(-x)); return s * (1.0 - s); }
  default: return 1.0;
  }
}
// src/layer.hpp:19,26,37,46 (CPU scale constant 1.41421356 for ReLU)
inline double act_scale_cpu(int a) { return a == kReLU ? 1.41421356 : 1.0; }
// src/cuda/kernels.cuh:61-71 (float 1.41421356f)
inline float act_scale_cuda(int a) { return a == kReLU ? 1.41421356f : 1.0f; }

// ----------------------------------------------------------------------------------
// Network: src/network.hpp:21-118 + DenseLayer src/layer.hpp:73-131
// ----------------------------------------------------------------------------------
struct Net {
  std::vector<int> dims; // nlayers+1
  std::vector<int> acts; // nlayers
  size_t nparams = 0;
  std::vector<size_t> offs; // per-layer offset into the flat vector
  // caches (layer.hpp:86-87 input_cache/z_cache, network.hpp:30-31)
  std::vector<Vec> z, a, delta;

  const uint8_t *masks = nullptr; // imposed ReLU activation pattern of the hidden layers, [layer][B][out] (see layer_forward)
  int nlayers() const { return (int)acts.size(); }
};

Net *net_create(int nlayers, const int *dims, const int *acts) {
  Net *n = new Net;
  n->dims.assign(dims, dims + nlayers + 1);
  n->acts.assign(acts, acts + nlayers);
  n->offs.resize(nlayers);
  size_t off = 0;
  for (int l = 0; l < nlayers; ++l) {
    n->offs[l] = off;
    off += (size_t)dims[l + 1] * dims[l] + dims[l + 1]; // layer.hpp:93 getParamsSize
  }
  n->nparams = off;
  n->z.resize(nlayers);
  n->a.resize(nlayers);
  n->delta.resize(nlayers);
  return n;
}

// z = W*in + b ; out = act(z)   (layer.hpp:100-111). in: K x B, out: M x B col-major.
// mask (optional, ReLU layers only): the activation pattern is IMPOSED instead of derived from the sign of z — unit (b, o) passes
// z through iff mask[b * M + o] != 0 (and its derivative is that 0 / 1). Used by the parity tests to evaluate the fp64 objective on
// the activation pattern an fp32 evaluation took (pre-activations within rounding of zero pick a side at random in fp32).
void layer_forward(const double *W, const double *bias, int M, int K, int act, const double *in, long B, double *z,
                   double *out, const uint8_t *mask = nullptr) {
#pragma omp parallel for schedule(static)
  for (long b = 0; b < B; ++b) {
    double *zb = z + (size_t)b * M;
    const double *xb = in + (size_t)b * K;
    for (int o = 0; o < M; ++o) zb[o] = 0.0;
    for (int k = 0; k < K; ++k) {
      const double xv = xb[k];
      if (xv == 0.0) continue; // exact: adding 0*w changes nothing
      const double *wk = W + (size_t)k * M;
      for (int o = 0; o < M; ++o) zb[o] += wk[o] * xv;
    }
    double *ob = out + (size_t)b * M;
    for (int o = 0; o < M; ++o) {
      zb[o] += bias[o];
      ob[o] = (mask && act == kReLU) ? (mask[(size_t)b * M + o] ? zb[o] : 0.0) : act_apply(act, zb[o]);
    }
  }
}

// dZ = next_grad .* act'(z); dW += dZ*in^T; db += rowsum(dZ); prev = W^T dZ (layer.hpp:113-128)
// next_grad is overwritten with dZ.
void layer_backward(const double *W, int M, int K, int act, const double *in, const double *z, double *next_grad, long B,
                    double *dW, double *db, double *prev, const uint8_t *mask = nullptr) {
  const size_t tot = (size_t)M * B;
#pragma omp parallel for schedule(static)
  for (long idx = 0; idx < (long)tot; ++idx)
    next_grad[idx] *= (mask && act == kReLU) ? (mask[idx] ? 1.0 : 0.0) : act_prime(act, z[idx]);

  // dW[o + k*M] += sum_b dZ[o + b*M] * in[k + b*K]; parallel over k => deterministic for any thread count
#pragma omp parallel for schedule(static)
  for (int k = 0; k < K; ++k) {
    double *dwk = dW + (size_t)k * M;
    for (long b = 0; b < B; ++b) {
      const double xv = in[(size_t)b * K + k];
      if (xv == 0.0) continue;
      const double *dzb = next_grad + (size_t)b * M;
      for (int o = 0; o < M; ++o) dwk[o] += dzb[o] * xv;
    }
  }
  for (long b = 0; b < B; ++b) {
    const double *dzb = next_grad + (size_t)b * M;
    for (int o = 0; o < M; ++o) db[o] += dzb[o];
  }
  if (prev) {
#pragma omp parallel for schedule(static)
    for (long b = 0; b < B; ++b) {
      const double *dzb = next_grad + (size_t)b * M;
      double *pb = prev + (size_t)b * K;
      for (int k = 0; k < K; ++k) {
        const double *wk = W + (size_t)k * M;
        double s = 0.0;
        for (int o = 0; o < M; ++o) s += wk[o] * dzb[o];
        pb[k] = s;
      }
    }
  }
}

// Network::forward (network.hpp:73-88). Returns pointer to the output activations (out x B).
const double *net_forward(Net &n, const double *params, const double *X, long B) {
  const double *cur = X;
  const uint8_t *mk = n.masks;
  for (int l = 0; l < n.nlayers(); ++l) {
    const int K = n.dims[l], M = n.dims[l + 1];
    n.z[l].resize((size_t)M * B);
    n.a[l].resize((size_t)M * B);
    const double *W = params + n.offs[l];
    layer_forward(W, W + (size_t)M * K, M, K, n.acts[l], cur, B, n.z[l].data(), n.a[l].data(),
                  (mk && l + 1 < n.nlayers()) ? mk : nullptr);
    if (mk) mk += (size_t)M * B;
    cur = n.a[l].data();
  }
  return cur;
}

// f closure of run_full_batch_cpu (src/unified_optimization.hpp:101-108):
//   loss = 0.5*||out - Y||^2 * (1/N)
double net_loss(Net &n, const double *params, const double *X, const double *T, long B) {
  const double *out = net_forward(n, params, X, B);
  const size_t tot = (size_t)n.dims.back() * B;
  double s = 0.0;
  for (size_t i = 0; i < tot; ++i) {
    double d = out[i] - T[i];
    s += d * d;
  }
  double loss = 0.5 * s;
  if (B > 0) loss *= 1.0 / (double)B;
  return loss;
}

// grad closure (src/unified_optimization.hpp:110-120): zeroGrads, forward, diff, backward, g *= 1/N.
// `scale` lets the S-LBFGS closure divide by |idx| instead (unified_optimization.hpp:372-375).
double net_loss_grad(Net &n, const double *params, const double *X, const double *T, long B, double *grad) {
  const double *out = net_forward(n, params, X, B);
  const int L = n.nlayers();
  const size_t tot = (size_t)n.dims.back() * B;
  std::fill(grad, grad + n.nparams, 0.0);
  n.delta[L - 1].resize(tot);
  double s = 0.0;
  double *diff = n.delta[L - 1].data();
  for (size_t i = 0; i < tot; ++i) {
    diff[i] = out[i] - T[i];
    s += diff[i] * diff[i];
  }
  // Network::backward (network.hpp:91-102)
  std::vector<const uint8_t *> lmask(L, nullptr);
  if (n.masks) {
    const uint8_t *mk = n.masks;
    for (int l = 0; l + 1 < L; ++l) { lmask[l] = mk; mk += (size_t)n.dims[l + 1] * B; }
  }
  for (int l = L - 1; l >= 0; --l) {
    const int K = n.dims[l], M = n.dims[l + 1];
    const double *W = params + n.offs[l];
    const double *in = (l == 0) ? X : n.a[l - 1].data();
    double *prev = nullptr;
    if (l > 0) {
      n.delta[l - 1].resize((size_t)K * B);
      prev = n.delta[l - 1].data();
    }
    layer_backward(W, M, K, n.acts[l], in, n.z[l].data(), n.delta[l].data(), B, grad + n.offs[l],
                   grad + n.offs[l] + (size_t)M * K, prev, lmask[l]);
  }
  const double inv = (B > 0) ? 1.0 / (double)B : 0.0;
  if (inv != 0.0)
    for (size_t i = 0; i < n.nparams; ++i) grad[i] *= inv;
  double loss = 0.5 * s;
  if (B > 0) loss *= inv;
  return loss;
}

// ----------------------------------------------------------------------------------
// Objective abstraction = the (VecFun f, GradFun Gradient) pair of the reference
// (src/common.hpp VecFun/GradFun). Counts evaluations for the bench report.
// ----------------------------------------------------------------------------------
struct Objective {
  std::function<double(const Vec &)> f;
  std::function<void(const Vec &, Vec &)> grad; // writes gradient
  long n_f = 0, n_g = 0;
  double F(const Vec &x) { ++n_f; return f(x); }
  Vec G(const Vec &x) { ++n_g; Vec g(x.size()); grad(x, g); return g; }
};

// ----------------------------------------------------------------------------------
// RingBuffer (src/minimizer/ring_buffer.hpp:15-134)
// ----------------------------------------------------------------------------------
template <typename T> struct Ring {
  std::vector<T> data;
  size_t cap = 0, head = 0, count = 0;
  explicit Ring(size_t c = 0) : data(c), cap(c) {}
  void push_back(const T &v) { // ring_buffer.hpp:43-59
    if (cap == 0) return;
    if (count < cap) {
      data[(head + count) % cap] = v;
      ++count;
    } else {
      data[head] = v;
      head = (head + 1) % cap;
    }
  }
  T &operator[](size_t i) { return data[(head + i) % cap]; }
  const T &operator[](size_t i) const { return data[(head + i) % cap]; }
  const T &back() const { return (*this)[count - 1]; }
  size_t size() const { return count; }
  bool empty() const { return count == 0; }
  void clear() { count = 0; head = 0; }
};

struct History { // per-iteration record == IterationRecorder<CpuBackend> (src/iteration_recorder.hpp:18-72)
  std::vector<double> loss, gnorm, ms;
};

// ----------------------------------------------------------------------------------
// LBFGS::compute_direction (src/minimizer/lbfgs.hpp:106-139): gamma UNGUARDED.
// ----------------------------------------------------------------------------------
Vec cpu_compute_direction(const Vec &g, const Ring<Vec> &S, const Ring<Vec> &Y, const Ring<double> &rho) {
  const size_t n = g.size();
  Vec r(n);
  if (S.empty()) {
    for (size_t i = 0; i < n; ++i) r[i] = -g[i];
    return r;
  }
  Vec q = g;
  const int k = (int)S.size();
  std::vector<double> alpha(k);
  for (int i = k - 1; i >= 0; --i) {
    alpha[i] = rho[i] * dot(S[i], q);
    axpy(-alpha[i], Y[i], q);
  }
  const double gamma = dot(S.back(), Y.back()) / dot(Y.back(), Y.back());
  Vec z(n);
  for (size_t i = 0; i < n; ++i) z[i] = gamma * q[i];
  for (int i = 0; i < k; ++i) {
    const double beta = rho[i] * dot(Y[i], z);
    axpy(alpha[i] - beta, S[i], z);
  }
  for (size_t i = 0; i < n; ++i) r[i] = -z[i];
  return r;
}

// ----------------------------------------------------------------------------------
// FullBatchMinimizer::line_search (src/minimizer/full_batch_minimizer.hpp:126-157)
// c1=1e-4, c2=0.9, rho=0.5, max_line_iters=50 (:113-116). Re-evaluates f(x), Gradient(x).
// ----------------------------------------------------------------------------------
struct WolfeParams { double c1 = 1e-4, c2 = 0.9, rho = 0.5; int max_line_iters = 50; };

double cpu_line_search(const Vec &x, const Vec &p, Objective &obj, const WolfeParams &wp) {
  const double f_old = obj.F(x);
  const double grad_f_old = dot(obj.G(x), p);
  const double inf = std::numeric_limits<double>::infinity();
  double alpha_min = 0.0, alpha_max = inf, alpha = 1.0;
  Vec x_new(x.size());
  for (int i = 0; i < wp.max_line_iters; ++i) {
    for (size_t j = 0; j < x.size(); ++j) x_new[j] = x[j] + alpha * p[j];
    const double f_new = obj.F(x_new);
    if (f_new > f_old + wp.c1 * alpha * grad_f_old) {
      alpha_max = alpha;
      alpha = wp.rho * (alpha_min + alpha_max);
      continue;
    }
    const double g_new_dot_p = dot(obj.G(x_new), p);
    if (g_new_dot_p < wp.c2 * grad_f_old) {
      alpha_min = alpha;
      if (alpha_max == inf) alpha *= 2;
      else alpha = wp.rho * (alpha_min + alpha_max);
      continue;
    }
    return alpha;
  }
  return alpha;
}

// ----------------------------------------------------------------------------------
// LBFGS::solve (src/minimizer/lbfgs.hpp:38-100). Returns iterations performed.
// ----------------------------------------------------------------------------------
int cpu_lbfgs_solve(Vec &x, Objective &obj, int m, int max_iters, double tol, const WolfeParams &wp, History *hist,
                    std::vector<double> *alphas) {
  Ring<Vec> s_list(m), y_list(m);
  Ring<double> rho_list(m);
  Vec grad = obj.G(x);
  Vec p;
  auto t0 = std::chrono::steady_clock::now();
  int iters = 0;
  for (iters = 0; iters < max_iters; ++iters) {
    if (norm(grad) < tol) break;
    p = cpu_compute_direction(grad, s_list, y_list, rho_list);
    double alpha;
    if (iters == 0) alpha = std::min(1.0, 1.0 / norm(grad)); // lbfgs.hpp:60-63: NO line search on the first step
    else alpha = cpu_line_search(x, p, obj, wp);
    if (alphas) alphas->push_back(alpha);
    Vec x_new(x.size()), s(x.size());
    for (size_t j = 0; j < x.size(); ++j) {
      x_new[j] = x[j] + alpha * p[j];
      s[j] = x_new[j] - x[j];
    }
    Vec grad_new = obj.G(x_new);
    Vec y(x.size());
    for (size_t j = 0; j < x.size(); ++j) y[j] = grad_new[j] - grad[j];
    x = x_new;
    const double ys = dot(y, s);
    if (ys > 1e-10) { // lbfgs.hpp:76
      s_list.push_back(s);
      y_list.push_back(y);
      rho_list.push_back(1.0 / ys);
    }
    grad = grad_new;
    if (hist) { // lbfgs.hpp:88-96: loss = f(x) (one more forward), ||g||, cumulative wall ms
      const double loss = obj.F(x);
      auto now = std::chrono::steady_clock::now();
      hist->loss.push_back(loss);
      hist->gnorm.push_back(norm(grad));
      hist->ms.push_back(std::chrono::duration<double, std::milli>(now - t0).count());
    }
  }
  return iters;
}

// ----------------------------------------------------------------------------------
// CUDA-flavoured L-BFGS (src/cuda/lbfgs.cuh:39-194, 206-261) restated in double:
// Armijo backtracking + safeguarded quadratic interpolation, steepest-descent fallback
// with history reset, history reset on line-search failure, the slot is written BEFORE
// the curvature test (so a rejected pair overwrites the slot at `head` with stale rho).
// ----------------------------------------------------------------------------------
struct ArmijoParams { int max_line_iters = 20; double c1 = 1e-4, rho = 0.5; };

struct GpuStyleRing { // raw arrays exactly like lbfgs.cuh:55-72
  int m = 0, head = 0, count = 0;
  std::vector<Vec> s, y;
  std::vector<double> rho;
  void init(int m_, size_t n) {
    m = m_; head = 0; count = 0;
    s.assign(m, Vec(n, 0.0)); y.assign(m, Vec(n, 0.0)); rho.assign(m, 0.0);
  }
  int phys(int logical) const { // lbfgs.cuh:225-230
    int start = head - count;
    start %= m;
    if (start < 0) start += m;
    return (start + logical) % m;
  }
};

void gpu_compute_direction(const Vec &g, const GpuStyleRing &R, Vec &p) { // lbfgs.cuh:206-261
  const size_t n = g.size();
  p.resize(n);
  if (R.count <= 0 || R.m == 0) {
    for (size_t i = 0; i < n; ++i) p[i] = -g[i];
    return;
  }
  Vec q = g;
  std::vector<double> alpha(R.count, 0.0);
  for (int li = R.count - 1; li >= 0; --li) {
    const int i = R.phys(li);
    const double a = R.rho[i] * dot(R.s[i], q);
    alpha[li] = a;
    axpy(-a, R.y[i], q);
  }
  const int last = R.phys(R.count - 1);
  const double ys = dot(R.s[last], R.y[last]);
  const double yy = dot(R.y[last], R.y[last]);
  const double gamma = (yy > 0.0) ? (ys / yy) : 1.0; // lbfgs.cuh:247 guard
  Vec z(n);
  for (size_t i = 0; i < n; ++i) z[i] = gamma * q[i];
  for (int li = 0; li < R.count; ++li) {
    const int i = R.phys(li);
    const double b = R.rho[i] * dot(R.y[i], z);
    axpy(alpha[li] - b, R.s[i], z);
  }
  for (size_t i = 0; i < n; ++i) p[i] = -z[i];
}

// loss_grad(x, g) -> loss : the LossGradFun of src/cuda/minimizer_base.cuh:15-16
using LossGrad = std::function<double(const Vec &, Vec &)>;

int gpu_lbfgs_solve(Vec &x, const LossGrad &lg, int m, int max_iters, double tol, const ArmijoParams &ap, History *hist,
                    std::vector<double> *alphas, long *n_evals) {
  const size_t n = x.size();
  Vec grad(n), grad_new(n), p(n), x_backup(n);
  GpuStyleRing R;
  R.init(m, n);
  long evals = 0;
  double loss = lg(x, grad); ++evals;
  auto t0 = std::chrono::steady_clock::now();
  int done = 0;
  for (int iter = 0; iter < max_iters; ++iter) {
    const double grad_norm = norm(grad);
    if (grad_norm < tol) break;
    gpu_compute_direction(grad, R, p);
    double gdp = dot(grad, p);
    if (gdp >= 0.0) { // lbfgs.cuh:98-104
      for (size_t i = 0; i < n; ++i) p[i] = -grad[i];
      gdp = -dot(grad, grad);
      R.head = 0; R.count = 0;
    }
    double alpha = (iter == 0) ? std::min(1.0, 1.0 / grad_norm) : 1.0; // lbfgs.cuh:108
    x_backup = x;
    double loss_new = 0.0;
    bool armijo_ok = false;
    for (int ls = 0; ls < ap.max_line_iters; ++ls) { // lbfgs.cuh:115-140
      for (size_t i = 0; i < n; ++i) x[i] = x_backup[i] + alpha * p[i];
      loss_new = lg(x, grad_new); ++evals;
      if (loss_new <= loss + ap.c1 * alpha * gdp) { armijo_ok = true; break; }
      const double denom = 2.0 * (loss_new - loss - gdp * alpha);
      bool fallback = true;
      if (std::abs(denom) > 1e-20) {
        const double na = -(gdp * alpha * alpha) / denom;
        if (na >= 0.1 * alpha && na <= 0.9 * alpha) { alpha = na; fallback = false; }
      }
      if (fallback) alpha *= ap.rho;
    }
    if (alphas) alphas->push_back(alpha);
    if (!armijo_ok) { R.head = 0; R.count = 0; } // lbfgs.cuh:147
    if (m > 0) { // lbfgs.cuh:149-170
      const int slot = R.head;
      for (size_t i = 0; i < n; ++i) {
        R.s[slot][i] = x[i] - x_backup[i];
        R.y[slot][i] = grad_new[i] - grad[i];
      }
      const double ys = dot(R.y[slot], R.s[slot]);
      if (ys > 1e-10) {
        R.rho[slot] = 1.0 / ys;
        R.head = (R.head + 1) % m;
        R.count = std::min(R.count + 1, m);
      }
    }
    grad = grad_new;
    loss = loss_new;
    if (hist) {
      auto now = std::chrono::steady_clock::now();
      hist->loss.push_back(loss);
      hist->gnorm.push_back(norm(grad));
      hist->ms.push_back(std::chrono::duration<double, std::milli>(now - t0).count());
    }
    ++done;
  }
  if (n_evals) *n_evals = evals;
  return done;
}

// ----------------------------------------------------------------------------------
// GradientDescent::solve without line search (src/minimizer/gd.hpp:42-69) — the form
// UnifiedGD_CPU uses (useLineSearch(false), src/unified_optimization.hpp:174-177).
// ----------------------------------------------------------------------------------
int cpu_gd_solve(Vec &x, Objective &obj, double step, int max_iters, double tol, History *hist) {
  Vec g = obj.G(x);
  auto t0 = std::chrono::steady_clock::now();
  int it = 0;
  for (it = 0; it < max_iters; ++it) {
    if (norm(g) < tol) break;
    axpy(-step, g, x);
    g = obj.G(x);
    if (hist) {
      hist->loss.push_back(obj.F(x));
      hist->gnorm.push_back(norm(g));
      hist->ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
  }
  return it;
}

// CudaGD::solve (src/cuda/gd.cuh:38-106) in double: momentum form v = mu*v - lr*g; x += v.
int gpu_gd_solve(Vec &x, const LossGrad &lg, double lr, double momentum, int max_iters, double tol, History *hist) {
  const size_t n = x.size();
  Vec grad(n), vel(n, 0.0);
  double loss = lg(x, grad);
  int done = 0;
  for (int iter = 0; iter < max_iters; ++iter) {
    if (norm(grad) < tol) break;
    if (momentum > 0.0) {
      for (size_t i = 0; i < n; ++i) { vel[i] = momentum * vel[i] - lr * grad[i]; x[i] += vel[i]; }
    } else {
      axpy(-lr, grad, x);
    }
    loss = lg(x, grad);
    if (hist) { hist->loss.push_back(loss); hist->gnorm.push_back(norm(grad)); hist->ms.push_back(0.0); }
    ++done;
  }
  return done;
}

// CudaSGD::solve (src/cuda/sgd.cuh:50-153) in double. batch_lg(x, start, count, g) -> mean loss of the slice.
using SliceLossGrad = std::function<double(const Vec &, long, long, Vec &)>;
int gpu_sgd_solve(Vec &x, const SliceLossGrad &blg, long total, int batch_size, double lr, double momentum,
                  double decay_rate, int decay_step, int max_iters, double tol, bool record, History *hist) {
  const size_t n = x.size();
  Vec grad(n), vel(n, 0.0);
  double cur_lr = lr;
  const long num_batches = (total + batch_size - 1) / batch_size;
  double prev = std::numeric_limits<double>::infinity();
  int done = 0;
  if (record) { // sgd.cuh:89-94: one record before the first epoch
    const double fl = blg(x, 0, total, grad);
    if (hist) { hist->loss.push_back(fl); hist->gnorm.push_back(norm(grad)); hist->ms.push_back(0.0); }
    ++done;
  }
  for (int iter = 0; iter < max_iters; ++iter) {
    if (decay_step > 0 && iter > 0 && iter % decay_step == 0) cur_lr *= decay_rate;
    double epoch_sum = 0.0;
    for (long b = 0; b < num_batches; ++b) {
      const long start = b * batch_size;
      const long cnt = std::min<long>(batch_size, total - start);
      const double bl = blg(x, start, cnt, grad);
      if (momentum > 0.0) {
        for (size_t i = 0; i < n; ++i) { vel[i] = momentum * vel[i] - cur_lr * grad[i]; x[i] += vel[i]; }
      } else {
        axpy(-cur_lr, grad, x);
      }
      epoch_sum += bl * (double)cnt;
    }
    const double epoch_avg = epoch_sum / (double)total;
    if (tol > 0.0 && std::isfinite(prev)) {
      const double denom = std::max(1.0, std::abs(prev));
      if (std::abs(prev - epoch_avg) / denom < tol) break;
    }
    prev = epoch_avg;
    if (record) {
      const double fl = blg(x, 0, total, grad);
      if (hist) { hist->loss.push_back(fl); hist->gnorm.push_back(norm(grad)); hist->ms.push_back(0.0); }
    }
    ++done;
  }
  return done;
}

// ----------------------------------------------------------------------------------
// S-LBFGS (src/minimizer/s_lbfgs.hpp:88-290)
// ----------------------------------------------------------------------------------
using Idx = std::vector<size_t>;
using BatchG = std::function<void(const Vec &, const Idx &, Vec &)>;
using BatchF = std::function<double(const Vec &, const Idx &)>;

// s_lbfgs.hpp:141-161 — partial Fisher-Yates with a fresh distribution object per draw.
Idx sample_minibatch_indices(size_t N, size_t batch_size, std::mt19937 &rng) {
  if (N == 0 || batch_size == 0) return {};
  Idx idx(N);
  std::iota(idx.begin(), idx.end(), 0);
  if (batch_size >= N) return idx;
  for (size_t i = 0; i < batch_size; ++i) {
    std::uniform_int_distribution<size_t> dist(i, N - 1);
    size_t j = dist(rng);
    std::swap(idx[i], idx[j]);
  }
  idx.resize(batch_size);
  return idx;
}

// s_lbfgs.hpp:105-136 — gamma guarded (|yy|<1e-12 -> 1) and clamped to [1e-6,1e6]; returns +H*v.
Vec slbfgs_two_loop(const Ring<Vec> &S, const Ring<Vec> &Y, const Ring<double> &rho, const Vec &v) {
  const int M = (int)S.size();
  std::vector<double> alpha(M);
  Vec q = v;
  for (int i = M - 1; i >= 0; --i) {
    alpha[i] = rho[i] * dot(S[i], q);
    axpy(-alpha[i], Y[i], q);
  }
  double gamma = 1.0;
  if (M > 0) {
    const double denom = dot(Y.back(), Y.back());
    if (std::abs(denom) < 1e-12) gamma = 1.0;
    else gamma = dot(S.back(), Y.back()) / denom;
    gamma = std::min(std::max(gamma, 1e-6), 1e6);
  }
  Vec r(q.size());
  for (size_t i = 0; i < q.size(); ++i) r[i] = gamma * q[i];
  for (int i = 0; i < M; ++i) {
    const double beta = rho[i] * dot(Y[i], r);
    axpy(alpha[i] - beta, S[i], r);
  }
  return r;
}

// s_lbfgs.hpp:88-101
Vec fd_hvp_batch(const BatchG &g, const Vec &w, const Idx &idx, const Vec &v, double eps = 1e-4) {
  const size_t n = w.size();
  Vec wp(n), wm(n), gp(n, 0.0), gm(n, 0.0), out(n);
  for (size_t i = 0; i < n; ++i) { wp[i] = w[i] + eps * v[i]; wm[i] = w[i] - eps * v[i]; }
  g(wp, idx, gp);
  g(wm, idx, gm);
  for (size_t i = 0; i < n; ++i) out[i] = (gp[i] - gm[i]) / (2.0 * eps);
  return out;
}

struct SlbfgsTrace { // everything a GPU port must reproduce
  std::vector<uint32_t> batch_idx_flat; // concatenated mini-batch indices (first epoch only unless trace_all)
  std::vector<int> anchor_pick;
  std::vector<int> pairs_after_epoch;
};

// s_lbfgs.hpp:165-290
int slbfgs_solve(Vec &weights, const BatchF &f, const BatchG &batch_g, int m, int M_param, int L, int b, int b_H,
                 double step_size, int N, int max_iters, double tol, unsigned seed, History *hist, SlbfgsTrace *trace) {
  int iters = 0;
  Ring<Vec> u_list(M_param > 0 ? M_param + 1 : 0);
  Ring<Vec> s_list(M_param > 0 ? M_param : 0), y_list(M_param > 0 ? M_param : 0);
  Ring<double> rho_list(M_param > 0 ? M_param : 0);
  std::mt19937 rng(seed);
  const size_t dim = weights.size();
  Vec wt = weights;
  Ring<Vec> w_history(L + 1);
  Idx full(N);
  std::iota(full.begin(), full.end(), 0);
  auto t0 = std::chrono::steady_clock::now();

  while (iters < max_iters) {
    w_history.clear();
    Vec full_gradient(dim, 0.0);
    batch_g(weights, full, full_gradient);
    if (norm(full_gradient) < tol) break;
    wt = weights;
    w_history.push_back(wt);
    Vec vr(dim, 0.0);
    for (int t = 0; t < m; ++t) {
      Idx mb = sample_minibatch_indices(N, b, rng);
      if (trace && iters == 0) for (size_t v : mb) trace->batch_idx_flat.push_back((uint32_t)v);
      Vec g_wt(dim, 0.0), g_wk(dim, 0.0);
      batch_g(wt, mb, g_wt);
      batch_g(weights, mb, g_wk);
      for (size_t i = 0; i < dim; ++i) vr[i] = (g_wt[i] - g_wk[i]) + full_gradient[i];
      Vec direction = slbfgs_two_loop(s_list, y_list, rho_list, vr);
      for (size_t i = 0; i < dim; ++i) wt[i] = wt[i] - step_size * direction[i];
      w_history.push_back(wt);
      if (t > 0 && t % L == 0) {
        Vec u(dim, 0.0);
        const int num_wt = (int)w_history.size();
        for (size_t i = 0; i < w_history.size(); ++i) axpy(1.0, w_history[i], u);
        if (num_wt > 0) for (size_t i = 0; i < dim; ++i) u[i] /= (double)num_wt;
        if (!u_list.empty()) {
          const Vec &u_prev = u_list.back();
          Vec s(dim);
          for (size_t i = 0; i < dim; ++i) s[i] = u[i] - u_prev[i];
          Idx hb = sample_minibatch_indices(N, b_H, rng);
          Vec y = fd_hvp_batch(batch_g, u, hb, s);
          const double ys = dot(y, s);
          if (std::abs(ys) > 1e-10) {
            s_list.push_back(s);
            y_list.push_back(y);
            rho_list.push_back(1.0 / ys);
          }
        }
        u_list.push_back(u);
      }
    }
    if (w_history.size() >= 2) {
      std::uniform_int_distribution<size_t> pick(0, w_history.size() - 2);
      const size_t pi = pick(rng);
      if (trace) trace->anchor_pick.push_back((int)pi);
      weights = w_history[pi];
    } else {
      weights = wt;
    }
    if (trace) trace->pairs_after_epoch.push_back((int)s_list.size());
    if (hist) {
      const double fl = f(weights, full);
      Vec gl(dim, 0.0);
      batch_g(weights, full, gl);
      hist->loss.push_back(fl);
      hist->gnorm.push_back(norm(gl));
      hist->ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    ++iters;
  }
  return iters;
}

// ----------------------------------------------------------------------------------
// CPU SGD: StochasticGradientDescent::stochastic_solve (src/minimizer/s_gd.hpp:63-145) with the closures of
// UnifiedSGD_CPU::optimize (src/unified_optimization.hpp:219-300): m = N / b random mini-batches per epoch drawn with the
// partial Fisher-Yates sampler (s_gd.hpp:148-170, the same draws as the S-LBFGS one) from one mt19937(kDefaultSeed);
// w -= step * mean-gradient; no momentum, no decay, no tolerance test; per epoch the recorder takes the full loss (the mean
// of the per-sample 0.5 ||out - y||^2, :116-120) and the norm of the full gradient (:126-131).
// ----------------------------------------------------------------------------------
int cpu_sgd_solve(Vec &w, const BatchF &full_loss, const BatchG &batch_g, int m, int b, double step, int N, int max_iters,
                  unsigned seed, History *hist) {
  int iters = 0;
  std::mt19937 rng(seed);
  const size_t dim = w.size();
  Idx full(N);
  std::iota(full.begin(), full.end(), 0);
  auto t0 = std::chrono::steady_clock::now();
  while (iters < max_iters) {
    for (int t = 0; t < m; ++t) {
      Idx mb = sample_minibatch_indices(N, b, rng);
      Vec g(dim, 0.0);
      batch_g(w, mb, g);
      for (size_t i = 0; i < dim; ++i) w[i] = w[i] - step * g[i];
    }
    if (hist) {
      const double loss = full_loss(w, full);
      Vec fg(dim, 0.0);
      batch_g(w, full, fg);
      hist->loss.push_back(loss);
      hist->gnorm.push_back(norm(fg));
      hist->ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    ++iters;
  }
  return iters;
}

// ----------------------------------------------------------------------------------
// analytic test functions of tests/main.cpp (the reference's only known-answer tests)
// ----------------------------------------------------------------------------------
Objective make_rosenbrock() { // tests/main.cpp:73-108
  Objective o;
  o.f = [](const Vec &v) {
    double val = 0.0;
    const int n = (int)v.size();
    for (int i = 0; i < n - 1; ++i) {
      double t1 = v[i + 1] - v[i] * v[i], t2 = 1.0 - v[i];
      val += 100.0 * t1 * t1 + t2 * t2;
    }
    return val;
  };
  o.grad = [](const Vec &v, Vec &g) {
    const int n = (int)v.size();
    std::fill(g.begin(), g.end(), 0.0);
    if (n > 1) g[0] = -2.0 * (1.0 - v[0]) - 400.0 * v[0] * (v[1] - v[0] * v[0]);
    else g[0] = -2.0 * (1.0 - v[0]);
    for (int i = 1; i < n - 1; ++i)
      g[i] = -2.0 * (1.0 - v[i]) - 400.0 * v[i] * (v[i + 1] - v[i] * v[i]) + 200.0 * (v[i] - v[i - 1] * v[i - 1]);
    if (n > 1) g[n - 1] = 200.0 * (v[n - 1] - v[n - 2] * v[n - 2]);
  };
  return o;
}
Objective make_ackley() { // tests/main.cpp:160-196
  Objective o;
  o.f = [](const Vec &v) {
    double s1 = 0, s2 = 0;
    const int n = (int)v.size();
    for (int i = 0; i < n; ++i) { s1 += v[i] * v[i]; s2 += std::cos(2.0 * M_PI * v[i]); }
    return -20.0 * std::exp(-0.2 * std::sqrt(s1 / n)) - std::exp(s2 / n) + 20.0 + std::exp(1.0);
  };
  o.grad = [](const Vec &v, Vec &g) {
    const int n = (int)v.size();
    double s1 = 0, s2 = 0;
    for (int i = 0; i < n; ++i) { s1 += v[i] * v[i]; s2 += std::cos(2.0 * M_PI * v[i]); }
    const double ec = std::exp(s2 / n), es = std::exp(-0.2 * std::sqrt(s1 / n));
    for (int i = 0; i < n; ++i) {
      double gs = v[i] / (n * std::sqrt(s1 / n));
      g[i] = 4.0 * es * gs + (2.0 * M_PI / n) * ec * std::sin(2.0 * M_PI * v[i]);
    }
  };
  return o;
}
Objective make_rastrigin() { // tests/main.cpp:17-36
  Objective o;
  o.f = [](const Vec &v) {
    double val = 0.0;
    const int n = (int)v.size();
    for (int i = 0; i < n; ++i) val += v[i] * v[i] - 10.0 * std::cos(2.0 * M_PI * v[i]);
    return 10.0 * n + val;
  };
  o.grad = [](const Vec &v, Vec &g) {
    for (size_t i = 0; i < v.size(); ++i) g[i] = 2.0 * v[i] + 2.0 * M_PI * 10.0 * std::sin(2.0 * M_PI * v[i]);
  };
  return o;
}

void copy_hist(const History &h, int cap, double *loss, double *gnorm, double *ms) {
  const int k = std::min<int>(cap, (int)h.loss.size());
  for (int i = 0; i < k; ++i) {
    if (loss) loss[i] = h.loss[i];
    if (gnorm) gnorm[i] = h.gnorm[i];
    if (ms) ms[i] = h.ms[i];
  }
}

Objective mlp_objective(Net &net, const double *X, const double *T, long B) {
  Objective o;
  o.f = [&net, X, T, B](const Vec &w) { return net_loss(net, w.data(), X, T, B); };
  o.grad = [&net, X, T, B](const Vec &w, Vec &g) { net_loss_grad(net, w.data(), X, T, B, g.data()); };
  return o;
}

} // namespace

// ====================================================================================
// C ABI for ctypes (tests / bench only)
// ====================================================================================
extern "C" {

int oracle_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_num_threads(int t) {
#ifdef _OPENMP
  omp_set_num_threads(t);
#else
  (void)t;
#endif
}

void *oracle_net_create(int nlayers, const int *dims, const int *acts) { return net_create(nlayers, dims, acts); }
void oracle_net_destroy(void *h) { delete (Net *)h; }
long oracle_net_params_size(void *h) { return (long)((Net *)h)->nparams; }

// Network::bindParams (src/network.hpp:45-70): ONE mt19937(seed), per layer a fresh
// normal_distribution<double>(0, scale*sqrt(1/in)) drawn for ALL entries incl. biases.
__attribute__((optimize("fp-contract=off"))) void oracle_init_params_cpu_rule(void *h, unsigned seed, double *out) {
  Net &n = *(Net *)h;
  std::mt19937 gen(seed);
  double *p = out;
  for (int l = 0; l < n.nlayers(); ++l) {
    const size_t cnt = (size_t)n.dims[l + 1] * n.dims[l] + n.dims[l + 1];
    const double sd = act_scale_cpu(n.acts[l]) * std::sqrt(1.0 / (double)n.dims[l]);
    std::normal_distribution<double> dist(0.0, sd);
    for (size_t i = 0; i < cnt; ++i) p[i] = dist(gen);
    p += cnt;
  }
}
// CudaNetwork::bindParams (src/cuda/network.cuh:37-59): normal_distribution<float> for the
// weights only, biases = 0.
__attribute__((optimize("fp-contract=off"))) void oracle_init_params_cuda_rule(void *h, unsigned seed, float *out) {
  Net &n = *(Net *)h;
  std::mt19937 gen(seed);
  size_t off = 0;
  for (int l = 0; l < n.nlayers(); ++l) {
    const size_t wc = (size_t)n.dims[l + 1] * n.dims[l], bc = n.dims[l + 1];
    const float sd = act_scale_cuda(n.acts[l]) * std::sqrt(1.0f / (float)n.dims[l]);
    std::normal_distribution<float> dist(0.0f, sd);
    for (size_t i = 0; i < wc; ++i) out[off + i] = dist(gen);
    for (size_t i = 0; i < bc; ++i) out[off + wc + i] = 0.0f;
    off += wc + bc;
  }
}

double oracle_loss(void *h, const double *params, const double *X, const double *T, long B) {
  return net_loss(*(Net *)h, params, X, T, B);
}
double oracle_loss_grad(void *h, const double *params, const double *X, const double *T, long B, double *grad) {
  return net_loss_grad(*(Net *)h, params, X, T, B, grad);
}
// loss and gradient with the ReLU activation pattern of the hidden layers imposed (masks: hidden layer 0's [B][out] bytes,
// then hidden layer 1's, ...). Test infrastructure for the "which side of zero did fp32 land on" question.
double oracle_loss_grad_masked(void *h, const double *params, const double *X, const double *T, long B, const uint8_t *masks,
                               double *grad) {
  Net &n = *(Net *)h;
  n.masks = masks;
  const double loss = net_loss_grad(n, params, X, T, B, grad);
  n.masks = nullptr;
  return loss;
}
void oracle_forward(void *h, const double *params, const double *X, long B, double *out) {
  Net &n = *(Net *)h;
  const double *o = net_forward(n, params, X, B);
  std::memcpy(out, o, sizeof(double) * (size_t)n.dims.back() * B);
}

// S-LBFGS closures of UnifiedSLBFGS_CPU (src/unified_optimization.hpp:343-405), lambda = 1e-4.
// idx == nullptr means the full batch (the `is_full_batch` shortcut, :354-359).
double oracle_slbfgs_batch_grad(void *h, const double *params, const double *X, const double *T, long N,
                                const uint32_t *idx, long bs, double lambda, double *grad) {
  Net &n = *(Net *)h;
  const int in = n.dims.front(), out = n.dims.back();
  double loss;
  if (!idx || bs == N) {
    loss = net_loss_grad(n, params, X, T, N, grad); // grad/N == grad/current_bs
  } else {
    Vec bx((size_t)in * bs), bt((size_t)out * bs);
    for (long i = 0; i < bs; ++i) {
      std::memcpy(&bx[(size_t)i * in], X + (size_t)idx[i] * in, sizeof(double) * in);
      std::memcpy(&bt[(size_t)i * out], T + (size_t)idx[i] * out, sizeof(double) * out);
    }
    loss = net_loss_grad(n, params, bx.data(), bt.data(), bs, grad);
  }
  double w2 = 0.0;
  for (size_t i = 0; i < n.nparams; ++i) { grad[i] += lambda * params[i]; w2 += params[i] * params[i]; }
  return loss + 0.5 * lambda * w2; // == batch_f (:383-405)
}

// direction kernels --------------------------------------------------------------
// S, Y: k x n row-major in LOGICAL order (0 = oldest). policy 0 = CPU (lbfgs.hpp:106-139),
// 1 = CUDA (lbfgs.cuh:206-261, gamma guard), 2 = S-LBFGS two-loop (+H v, gamma clamp).
void oracle_direction(long n, int k, const double *S, const double *Y, const double *rho, const double *g, int policy,
                      double *p_out) {
  Vec gv(g, g + n);
  if (policy == 1) {
    GpuStyleRing R;
    R.init(std::max(k, 1), n);
    for (int i = 0; i < k; ++i) {
      R.s[i].assign(S + (size_t)i * n, S + (size_t)(i + 1) * n);
      R.y[i].assign(Y + (size_t)i * n, Y + (size_t)(i + 1) * n);
      R.rho[i] = rho[i];
    }
    R.count = k;
    R.head = (k == R.m) ? 0 : k;
    Vec p;
    gpu_compute_direction(gv, R, p);
    std::copy(p.begin(), p.end(), p_out);
    return;
  }
  Ring<Vec> Sr(k), Yr(k);
  Ring<double> rr(k);
  for (int i = 0; i < k; ++i) {
    Sr.push_back(Vec(S + (size_t)i * n, S + (size_t)(i + 1) * n));
    Yr.push_back(Vec(Y + (size_t)i * n, Y + (size_t)(i + 1) * n));
    rr.push_back(rho[i]);
  }
  Vec p = (policy == 0) ? cpu_compute_direction(gv, Sr, Yr, rr) : slbfgs_two_loop(Sr, Yr, rr, gv);
  std::copy(p.begin(), p.end(), p_out);
}

// full-batch L-BFGS on the MLP objective ------------------------------------------
// policy 0: reference CPU algorithm (weak Wolfe). policy 1: reference CUDA algorithm (Armijo) in double.
// hist_* have capacity max_iters. out_counts = {iterations, n_f_evals, n_grad_evals}.
int oracle_lbfgs_mlp(void *h, double *params, const double *X, const double *T, long B, int m, int max_iters, double tol,
                     int policy, double *hist_loss, double *hist_gnorm, double *hist_ms, double *hist_alpha,
                     long *out_counts) {
  Net &net = *(Net *)h;
  Vec x(params, params + net.nparams);
  History hist;
  std::vector<double> alphas;
  int iters;
  long nf = 0, ng = 0;
  if (policy == 0) {
    Objective obj = mlp_objective(net, X, T, B);
    iters = cpu_lbfgs_solve(x, obj, m, max_iters, tol, WolfeParams{}, &hist, &alphas);
    nf = obj.n_f; ng = obj.n_g;
  } else {
    LossGrad lg = [&](const Vec &w, Vec &g) { return net_loss_grad(net, w.data(), X, T, B, g.data()); };
    long ev = 0;
    iters = gpu_lbfgs_solve(x, lg, m, max_iters, tol, ArmijoParams{}, &hist, &alphas, &ev);
    nf = ev; ng = ev;
  }
  std::copy(x.begin(), x.end(), params);
  copy_hist(hist, max_iters, hist_loss, hist_gnorm, hist_ms);
  if (hist_alpha) for (int i = 0; i < std::min<int>(max_iters, (int)alphas.size()); ++i) hist_alpha[i] = alphas[i];
  if (out_counts) { out_counts[0] = iters; out_counts[1] = nf; out_counts[2] = ng; }
  return iters;
}

// analytic known-answer tests (tests/main.cpp). fn: 0 rosenbrock, 1 ackley, 2 rastrigin.
int oracle_lbfgs_analytic(int fn, long n, double *x, int m, int max_iters, double tol, int policy, double *final_gnorm,
                          double *final_f) {
  Objective obj = fn == 0 ? make_rosenbrock() : (fn == 1 ? make_ackley() : make_rastrigin());
  Vec xv(x, x + n);
  int iters;
  if (policy == 0) {
    iters = cpu_lbfgs_solve(xv, obj, m, max_iters, tol, WolfeParams{}, nullptr, nullptr);
  } else {
    LossGrad lg = [&](const Vec &w, Vec &g) { obj.grad(w, g); return obj.f(w); };
    iters = gpu_lbfgs_solve(xv, lg, m, max_iters, tol, ArmijoParams{}, nullptr, nullptr, nullptr);
  }
  Vec g(n);
  obj.grad(xv, g);
  if (final_gnorm) *final_gnorm = norm(g);
  if (final_f) *final_f = obj.f(xv);
  std::copy(xv.begin(), xv.end(), x);
  return iters;
}
double oracle_analytic_eval(int fn, long n, const double *x, double *grad) {
  Objective obj = fn == 0 ? make_rosenbrock() : (fn == 1 ? make_ackley() : make_rastrigin());
  Vec xv(x, x + n), g(n);
  obj.grad(xv, g);
  if (grad) std::copy(g.begin(), g.end(), grad);
  return obj.f(xv);
}

// GD / SGD ----------------------------------------------------------------------------
int oracle_gd_mlp(void *h, double *params, const double *X, const double *T, long B, double lr, double momentum,
                  int max_iters, double tol, int policy, double *hist_loss, double *hist_gnorm) {
  Net &net = *(Net *)h;
  Vec x(params, params + net.nparams);
  History hist;
  int iters;
  if (policy == 0) {
    Objective obj = mlp_objective(net, X, T, B);
    iters = cpu_gd_solve(x, obj, lr, max_iters, tol, &hist);
  } else {
    LossGrad lg = [&](const Vec &w, Vec &g) { return net_loss_grad(net, w.data(), X, T, B, g.data()); };
    iters = gpu_gd_solve(x, lg, lr, momentum, max_iters, tol, &hist);
  }
  std::copy(x.begin(), x.end(), params);
  copy_hist(hist, max_iters, hist_loss, hist_gnorm, nullptr);
  return iters;
}

int oracle_sgd_mlp_cuda_policy(void *h, double *params, const double *X, const double *T, long B, int batch_size,
                               double lr, double momentum, double decay_rate, int decay_step, int max_iters, double tol,
                               int record, double *hist_loss, double *hist_gnorm) {
  Net &net = *(Net *)h;
  const int in = net.dims.front(), out = net.dims.back();
  Vec x(params, params + net.nparams);
  History hist;
  SliceLossGrad blg = [&](const Vec &w, long start, long cnt, Vec &g) {
    return net_loss_grad(net, w.data(), X + (size_t)start * in, T + (size_t)start * out, cnt, g.data());
  };
  int iters = gpu_sgd_solve(x, blg, B, batch_size, lr, momentum, decay_rate, decay_step, max_iters, tol, record != 0,
                            &hist);
  std::copy(x.begin(), x.end(), params);
  copy_hist(hist, max_iters + 1, hist_loss, hist_gnorm, nullptr);
  return iters;
}

// S-LBFGS (UnifiedSLBFGS_CPU::optimize + SLBFGS::stochastic_solve) -----------------------
// trace_idx (capacity trace_cap) receives the first epoch's concatenated mini-batch indices.
int oracle_slbfgs_mlp(void *h, double *params, const double *X, const double *T, long N, int batch_size, int M_param,
                      int L, int b_H_param, double step, int max_iters, double tol, unsigned seed, double *hist_loss,
                      double *hist_gnorm, uint32_t *trace_idx, long trace_cap, int *anchor_picks, int *pairs_after) {
  Net &net = *(Net *)h;
  const double lambda = 1e-4; // unified_optimization.hpp:334
  const int b_H = b_H_param > 0 ? b_H_param : batch_size / 2; // :325
  int m = (int)(N / batch_size); // :326
  if (m == 0) m = 1;
  Vec w(params, params + net.nparams);
  auto to_u32 = [](const Idx &idx) { std::vector<uint32_t> v(idx.size()); for (size_t i = 0; i < idx.size(); ++i) v[i] = (uint32_t)idx[i]; return v; };
  BatchG bg = [&](const Vec &wv, const Idx &idx, Vec &g) {
    auto v = to_u32(idx);
    oracle_slbfgs_batch_grad(h, wv.data(), X, T, N, ((long)idx.size() == N) ? nullptr : v.data(), (long)idx.size(), lambda,
                             g.data());
  };
  BatchF bf = [&](const Vec &wv, const Idx &idx) {
    // batch_f (:383-405) always gathers; mathematically identical to the full-batch forward.
    const int in = net.dims.front(), out = net.dims.back();
    const long bs = (long)idx.size();
    double loss;
    if (bs == N) loss = net_loss(net, wv.data(), X, T, N);
    else {
      Vec bx((size_t)in * bs), bt((size_t)out * bs);
      for (long i = 0; i < bs; ++i) {
        std::memcpy(&bx[(size_t)i * in], X + idx[i] * in, sizeof(double) * in);
        std::memcpy(&bt[(size_t)i * out], T + idx[i] * out, sizeof(double) * out);
      }
      loss = net_loss(net, wv.data(), bx.data(), bt.data(), bs);
    }
    double w2 = 0.0;
    for (double v : wv) w2 += v * v;
    return loss + 0.5 * lambda * w2;
  };
  History hist;
  SlbfgsTrace tr;
  int iters = slbfgs_solve(w, bf, bg, m, M_param, L, batch_size, b_H, step, (int)N, max_iters, tol, seed, &hist, &tr);
  std::copy(w.begin(), w.end(), params);
  copy_hist(hist, max_iters, hist_loss, hist_gnorm, nullptr);
  if (trace_idx) for (long i = 0; i < std::min<long>(trace_cap, (long)tr.batch_idx_flat.size()); ++i) trace_idx[i] = tr.batch_idx_flat[i];
  if (anchor_picks) for (size_t i = 0; i < tr.anchor_pick.size() && (int)i < max_iters; ++i) anchor_picks[i] = tr.anchor_pick[i];
  if (pairs_after) for (size_t i = 0; i < tr.pairs_after_epoch.size() && (int)i < max_iters; ++i) pairs_after[i] = tr.pairs_after_epoch[i];
  return iters;
}

// CPU SGD (UnifiedSGD_CPU::optimize + StochasticGradientDescent::stochastic_solve): random mini-batches, plain objective
// (no L2 term: the closure at unified_optimization.hpp:245-269 divides by the batch size and adds nothing).
int oracle_sgd_mlp_cpu_policy(void *h, double *params, const double *X, const double *T, long N, int batch_size, double step,
                              int max_iters, unsigned seed, double *hist_loss, double *hist_gnorm) {
  Net &net = *(Net *)h;
  int m = (int)(N / batch_size); // unified_optimization.hpp:232-233
  if (m == 0) m = 1;
  Vec w(params, params + net.nparams);
  auto to_u32 = [](const Idx &idx) { std::vector<uint32_t> v(idx.size()); for (size_t i = 0; i < idx.size(); ++i) v[i] = (uint32_t)idx[i]; return v; };
  BatchG bg = [&](const Vec &wv, const Idx &idx, Vec &g) {
    auto v = to_u32(idx);
    oracle_slbfgs_batch_grad(h, wv.data(), X, T, N, ((long)idx.size() == N) ? nullptr : v.data(), (long)idx.size(), 0.0, g.data());
  };
  BatchF fl = [&](const Vec &wv, const Idx &) { return net_loss(net, wv.data(), X, T, N); };
  History hist;
  const int iters = cpu_sgd_solve(w, fl, bg, m, batch_size, step, (int)N, max_iters, seed, &hist);
  std::copy(w.begin(), w.end(), params);
  copy_hist(hist, max_iters, hist_loss, hist_gnorm, nullptr);
  return iters;
}

// The sampler alone, so the product's host-side sampler can be diffed against it.
// Draws `count` mini-batches of size b in sequence from one mt19937(seed).
void oracle_sample_stream(unsigned seed, long N, long b, int count, uint32_t *out) {
  std::mt19937 rng(seed);
  for (int c = 0; c < count; ++c) {
    Idx idx = sample_minibatch_indices((size_t)N, (size_t)b, rng);
    for (size_t i = 0; i < idx.size(); ++i) out[(size_t)c * b + i] = (uint32_t)idx[i];
  }
}

} // extern "C"

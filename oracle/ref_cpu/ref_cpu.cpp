// TEST INFRASTRUCTURE ONLY — builds oracle/_ref/libref_cpu.so: the reference's OWN CPU path
// (/root/reference/src/{unified_launcher,unified_optimization,network,layer}.hpp, src/minimizer/*.hpp), included unmodified
// from where it lies and driven through its public API (UnifiedLauncher<CpuBackend>::train with the UnifiedLBFGS / UnifiedGD /
// UnifiedSGD / UnifiedSLBFGS strategies), behind a few extern "C" entry points.
//
// The one thing that is not the reference's: <Eigen/...> resolves to the stand-in under shim/ (Eigen itself is absent from
// this image), so the dense arithmetic inside Eigen calls is the stand-in's. Control flow, line searches, ring buffers, the
// objective closures, the RNG consumption and the recorder are the reference's code, executed as written.
//
// Used to (1) pin oracle/oracle.cpp (the fp64 restatement the GPU parity tests check against) on outputs of the reference itself —
// tests/golden/make_golden_ref_cpu.py writes them to tests/golden/golden_ref_cpu.json, tests/test_oracle.py compares — and
// (2) as the `--impl reference` / cpu_baseline arm of bench.py when present. The reference's layer sizes are template
// parameters, so the networks available here are the fixed list in with_net() below.
#include "src/unified_launcher.hpp"

#include <cstdio>
#include <cstring>
#include <string>

namespace {

using Vec = Eigen::VectorXd;
using Mat = Eigen::MatrixXd;
using Launcher = UnifiedLauncher<CpuBackend>;

// net ids: 0 = 784-128-10 (ReLU, Linear)        tests/mnist/main-cpu.cpp
//          1 = 784-128-64-10 (ReLU, ReLU, Linear) BASELINE configs[2]
//          2 = 20-16-8-4 (Tanh, Sigmoid, Linear)  small, for quick goldens
//          3 = 784-256-128-64-10 (ReLU x3, Linear) tests/fashion-mnist/main_gpu_deep.cpp
template <typename F> int with_net(int id, F &&f) {
  Launcher L;
  switch (id) {
  case 0:
    L.addLayer<784, 128, cpu_mlp::ReLU>();
    L.addLayer<128, 10, cpu_mlp::Linear>();
    break;
  case 1:
    L.addLayer<784, 128, cpu_mlp::ReLU>();
    L.addLayer<128, 64, cpu_mlp::ReLU>();
    L.addLayer<64, 10, cpu_mlp::Linear>();
    break;
  case 2:
    L.addLayer<20, 16, cpu_mlp::Tanh>();
    L.addLayer<16, 8, cpu_mlp::Sigmoid>();
    L.addLayer<8, 4, cpu_mlp::Linear>();
    break;
  case 3:
    L.addLayer<784, 256, cpu_mlp::ReLU>();
    L.addLayer<256, 128, cpu_mlp::ReLU>();
    L.addLayer<128, 64, cpu_mlp::ReLU>();
    L.addLayer<64, 10, cpu_mlp::Linear>();
    break;
  default: return -1;
  }
  L.buildNetwork();
  return f(L);
}

Mat to_mat(const double *p, long rows, long cols) {
  Mat m(rows, cols);
  std::memcpy(m.data(), p, sizeof(double) * (size_t)(rows * cols));
  return m;
}

void set_params(Launcher &L, const double *w) {
  auto &net = L.getWrapper().getInternal();
  Vec v((Eigen::Index)net.getSize());
  std::memcpy(v.data(), w, sizeof(double) * net.getSize());
  net.setParams(v);
}

// a "minimizer" that only evaluates the closures run_full_batch_cpu builds (src/unified_optimization.hpp:101-120): the
// reference's own objective, exactly as its L-BFGS sees it
struct EvalOnly : cpu_mlp::FullBatchMinimizer<Vec, Mat> {
  double loss = 0.0;
  Vec grad;
  Vec solve(Vec x, VecFun<Vec, double> &f, GradFun<Vec> &Gradient) override {
    loss = f(x);
    grad = Gradient(x);
    return x;
  }
};

} // namespace

extern "C" {

struct RefCpuCfg { // UnifiedConfig (src/unified_optimization.hpp:26-48), plain C layout
  int max_iters;
  double tolerance, learning_rate, momentum;
  int batch_size, m_param, L_param, b_H_param, log_interval;
};

int ref_cpu_num_threads() { return Eigen::nbThreads(); }
void ref_cpu_set_num_threads(int t) {
#ifdef _OPENMP
  if (t > 0) omp_set_num_threads(t);
#else
  (void)t;
#endif
}

long ref_cpu_params_size(int net_id) {
  long n = -1;
  with_net(net_id, [&](Launcher &L) { n = (long)L.getWrapper().getParamsSize(); return 0; });
  return n;
}

// cpu_mlp::Network::bindParams(seed) (src/network.hpp:45-74)
int ref_cpu_bind_params(int net_id, unsigned seed, double *w_out) {
  return with_net(net_id, [&](Launcher &L) {
    L.getWrapper().bindParams(seed);
    auto &net = L.getWrapper().getInternal();
    std::memcpy(w_out, net.getParamsData(), sizeof(double) * net.getSize());
    return 0;
  });
}

// loss and gradient through the closures of run_full_batch_cpu
int ref_cpu_loss_grad(int net_id, const double *w, const double *X, const double *T, long in_dim, long out_dim, long N,
                      double *loss_out, double *grad_out) {
  return with_net(net_id, [&](Launcher &L) {
    set_params(L, w);
    UnifiedDataset data;
    data.train_x = to_mat(X, in_dim, N);
    data.train_y = to_mat(T, out_dim, N);
    EvalOnly ev;
    run_full_batch_cpu(L.getWrapper(), data, ev);
    *loss_out = ev.loss;
    std::memcpy(grad_out, ev.grad.data(), sizeof(double) * (size_t)ev.grad.size());
    return 0;
  });
}

// cpu_mlp::LBFGS / GradientDescent configured as UnifiedLBFGS_CPU / UnifiedGD_CPU::optimize configure them
// (src/unified_optimization.hpp:161-213) but with the recorder kept, so the per-iteration history comes back in full
// precision (the strategies' own recorder only reaches a 6-digit CSV). kind: 0 = GD, 1 = L-BFGS.
int ref_cpu_full_batch(int net_id, int kind, const double *w0, const double *X, const double *T, long in_dim, long out_dim, long N,
                       const RefCpuCfg *cfg, double *w_out, double *loss_hist, double *gnorm_hist, double *ms_hist, int *hist_size,
                       int *iterations) {
  return with_net(net_id, [&](Launcher &L) {
    set_params(L, w0);
    UnifiedDataset data;
    data.train_x = to_mat(X, in_dim, N);
    data.train_y = to_mat(T, out_dim, N);
    IterationRecorder<CpuBackend> rec;
    rec.init(cfg->max_iters);
    unsigned iters = 0;
    if (kind == 1) {
      cpu_mlp::LBFGS<Vec, Mat> m;
      m.setMaxIterations(cfg->max_iters);
      m.setTolerance(cfg->tolerance);
      m.setHistorySize(cfg->m_param > 0 ? cfg->m_param : 10);
      m.setRecorder(&rec);
      run_full_batch_cpu(L.getWrapper(), data, m);
      iters = m.iterations();
    } else {
      cpu_mlp::GradientDescent<Vec, Mat> m;
      m.setMaxIterations(cfg->max_iters);
      m.setTolerance(cfg->tolerance);
      m.setStepSize(cfg->learning_rate);
      m.useLineSearch(false);
      m.setRecorder(&rec);
      run_full_batch_cpu(L.getWrapper(), data, m);
      iters = m.iterations();
    }
    std::vector<double> l, g, t;
    rec.copy_to_host(l, g, t);
    for (size_t i = 0; i < l.size(); ++i) { loss_hist[i] = l[i]; gnorm_hist[i] = g[i]; ms_hist[i] = t[i]; }
    *hist_size = (int)l.size();
    *iterations = (int)iters;
    auto &net = L.getWrapper().getInternal();
    std::memcpy(w_out, net.getParamsData(), sizeof(double) * net.getSize());
    return 0;
  });
}

// The whole public path: UnifiedLauncher<CpuBackend>::train(strategy, config) (src/unified_launcher.hpp:49-58). The strategy
// writes <name>_history.csv (6 significant digits); the final parameters come back in full precision.
// kind: 0 = UnifiedGD, 1 = UnifiedLBFGS, 2 = UnifiedSGD, 3 = UnifiedSLBFGS.
int ref_cpu_train(int net_id, int kind, const double *w0, const double *X, const double *T, long in_dim, long out_dim, long N,
                  const RefCpuCfg *cfg, const char *name, double *w_out) {
  return with_net(net_id, [&](Launcher &L) {
    UnifiedDataset data;
    data.train_x = to_mat(X, in_dim, N);
    data.train_y = to_mat(T, out_dim, N);
    L.setData(data);
    set_params(L, w0);
    UnifiedConfig c;
    c.name = name;
    c.max_iters = cfg->max_iters;
    c.tolerance = cfg->tolerance;
    c.learning_rate = cfg->learning_rate;
    c.momentum = cfg->momentum;
    c.batch_size = cfg->batch_size;
    c.m_param = cfg->m_param;
    c.L_param = cfg->L_param;
    c.b_H_param = cfg->b_H_param;
    c.log_interval = cfg->log_interval;
    c.reset_params = false; // keep the injected w0 (train() otherwise re-draws with bindParams(seed))
    if (kind == 0) { UnifiedGD<CpuBackend> o; L.train(o, c); }
    else if (kind == 1) { UnifiedLBFGS<CpuBackend> o; L.train(o, c); }
    else if (kind == 2) { UnifiedSGD<CpuBackend> o; L.train(o, c); }
    else if (kind == 3) { UnifiedSLBFGS<CpuBackend> o; L.train(o, c); }
    else return -2;
    auto &net = L.getWrapper().getInternal();
    std::memcpy(w_out, net.getParamsData(), sizeof(double) * net.getSize());
    return 0;
  });
}

} // extern "C"

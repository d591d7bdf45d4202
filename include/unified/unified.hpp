// unified.hpp — the reference's backend-selection layer for the CudaBackend:
//   src/seed.hpp, src/iteration_recorder.hpp:81-146, src/network_wrapper.hpp:20-110,
//   src/unified_optimization.hpp:26-64,410-696, src/unified_launcher.hpp:83-205
// Same type names and members, so tests/mnist/main-gpu.cpp, tests/fashion-mnist/main-gpu.cpp and main_gpu_deep.cpp
// compile against it with only the data source replaced. There is NO CpuBackend here (the CPU path is the
// reference's own; this library is the CUDA backend and never falls back to a CPU implementation).
#pragma once

#include "../cuda_mlp/cuda_mlp.hpp"

#include <algorithm>
#include <chrono>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#if defined(B200_USE_EIGEN) && __has_include(<Eigen/Core>)
#include <Eigen/Core>
namespace b200 { using Matrix = Eigen::MatrixXd; }
#else
namespace b200 {
/// Minimal column-major double matrix standing in for Eigen::MatrixXd (Eigen is not a dependency of this backend).
class Matrix {
public:
  Matrix() = default;
  Matrix(long rows, long cols) { resize(rows, cols); }
  void resize(long rows, long cols) { rows_ = rows; cols_ = cols; v_.assign((size_t)rows * cols, 0.0); }
  long rows() const { return rows_; }
  long cols() const { return cols_; }
  long size() const { return rows_ * cols_; }
  double *data() { return v_.data(); }
  const double *data() const { return v_.data(); }
  double &operator()(long r, long c) { return v_[(size_t)c * rows_ + r]; }
  double operator()(long r, long c) const { return v_[(size_t)c * rows_ + r]; }
private:
  long rows_ = 0, cols_ = 0;
  std::vector<double> v_;
};
} // namespace b200
#endif

constexpr unsigned int kDefaultSeed = 123; // src/seed.hpp:4

struct CpuBackend {};
struct CudaBackend {};

namespace cpu_mlp { // activation tag types used by addLayer<In, Out, Activation>() (src/layer.hpp:15-47)
struct Linear {}; struct Sigmoid {}; struct Tanh {}; struct ReLU {};
} // namespace cpu_mlp

template <typename T> struct ActivationToEnum;
template <> struct ActivationToEnum<cpu_mlp::Linear> { static constexpr cuda_mlp::ActivationType value = cuda_mlp::ActivationType::Linear; };
template <> struct ActivationToEnum<cpu_mlp::Sigmoid> { static constexpr cuda_mlp::ActivationType value = cuda_mlp::ActivationType::Sigmoid; };
template <> struct ActivationToEnum<cpu_mlp::Tanh> { static constexpr cuda_mlp::ActivationType value = cuda_mlp::ActivationType::Tanh; };
template <> struct ActivationToEnum<cpu_mlp::ReLU> { static constexpr cuda_mlp::ActivationType value = cuda_mlp::ActivationType::ReLU; };

// ---- IterationRecorder<CudaBackend> (src/iteration_recorder.hpp:81-146): host vectors instead of three device
// arrays written with one blocking 4-byte memcpy each per iteration ---------------------------------------------
template <typename Backend> class IterationRecorder;
template <> class IterationRecorder<CudaBackend> {
public:
  void init(int capacity) {
    capacity_ = std::max(capacity, 0);
    loss_.assign(capacity_, 0.f); grad_.assign(capacity_, 0.f); time_.assign(capacity_, 0.f);
    size_ = 0;
  }
  void reset() { size_ = 0; }
  void record(int idx, cuda_mlp::CudaScalar loss, cuda_mlp::CudaScalar grad_norm, cuda_mlp::CudaScalar time_ms = 0) {
    if (idx < 0 || idx >= capacity_) return;
    loss_[idx] = loss; grad_[idx] = grad_norm; time_[idx] = time_ms;
    size_ = std::max(size_, idx + 1);
  }
  int size() const { return size_; }
  int capacity() const { return capacity_; }
  void copy_to_host(std::vector<cuda_mlp::CudaScalar> &loss, std::vector<cuda_mlp::CudaScalar> &grad,
                    std::vector<cuda_mlp::CudaScalar> &time_ms) const {
    loss.assign(loss_.begin(), loss_.begin() + size_);
    grad.assign(grad_.begin(), grad_.begin() + size_);
    time_ms.assign(time_.begin(), time_.begin() + size_);
  }
  // used by the minimizers to hand the arrays to the C ABI
  float *loss_ptr() { return loss_.data(); }
  float *grad_ptr() { return grad_.data(); }
  float *time_ptr() { return time_.data(); }
  void set_size(int s) { size_ = std::min(std::max(s, 0), capacity_); }
private:
  std::vector<float> loss_, grad_, time_;
  int capacity_ = 0, size_ = 0;
};

namespace cuda_mlp {
inline void CudaMinimizerBase::begin_history(b200_history &h) {
  h = b200_history{};
  if (recorder_) {
    recorder_->reset();
    h.capacity = recorder_->capacity();
    h.loss = recorder_->loss_ptr(); h.grad_norm = recorder_->grad_ptr(); h.time_ms = recorder_->time_ptr();
  }
}
inline void CudaMinimizerBase::end_history(const b200_history &h) {
  last_iterations_ = h.iterations;
  last_evaluations_ = h.evaluations;
  if (recorder_) recorder_->set_size(h.size);
}
} // namespace cuda_mlp

// ---- NetworkWrapper<CudaBackend> (src/network_wrapper.hpp:87-110) ---------------------------------------------
template <typename Backend> class NetworkWrapper;
template <> class NetworkWrapper<CudaBackend> {
public:
  using InternalNetwork = cuda_mlp::CudaNetwork;
  explicit NetworkWrapper(cuda_mlp::CublasHandle &handle) : network_(handle) {}
  template <int In, int Out, typename Activation> void addLayer() { network_.addLayer(In, Out, ActivationToEnum<Activation>::value); }
  void bindParams() { network_.bindParams(); }
  void bindParams(unsigned int seed) { network_.bindParams(seed); }
  InternalNetwork &getInternal() { return network_; }
  const InternalNetwork &getInternal() const { return network_; }
  size_t getParamsSize() const { return network_.params_size(); }
private:
  InternalNetwork network_;
};

// ---- UnifiedConfig / UnifiedDataset (src/unified_optimization.hpp:26-59) -----------------------------------------
struct UnifiedConfig {
  std::string name = "Experiment";
  int max_iters = 100;
  double tolerance = 1e-4;
  double learning_rate = 0.01;
  double momentum = 0.0;
  double lr_decay = 0.0;
  int lr_decay_rate = 1;
  int batch_size = 128;
  int m_param = 10;
  int L_param = 10;
  int b_H_param = 0;
  int log_interval = 10;
  bool reset_params = true;
  unsigned int seed = kDefaultSeed;
  // additions (defaults keep the reference behaviour)
  cuda_mlp::LineSearch linesearch = cuda_mlp::LineSearch::Armijo;
  cuda_mlp::Precision precision = cuda_mlp::Precision::FP32;
};

struct UnifiedDataset {
  b200::Matrix train_x, train_y, test_x, test_y; // features x samples, column-major (sample = one column)
};

inline std::string cuda_log_filename(const UnifiedConfig &config) { return (config.name.empty() ? "run" : config.name) + "_history.csv"; }

inline void write_cuda_history_csv(const std::string &filename, const IterationRecorder<CudaBackend> &recorder, int log_interval) {
  if (log_interval <= 0) return; // src/unified_optimization.hpp:446-465, same schema and stride
  std::vector<float> loss, grad, time_ms;
  recorder.copy_to_host(loss, grad, time_ms);
  if (loss.empty()) return;
  std::ofstream f(filename);
  if (!f.is_open()) return;
  f << "Iteration,Loss,GradNorm,TimeMs\n";
  for (size_t i = 0; i < loss.size(); i += (size_t)std::max(1, log_interval)) f << i << "," << loss[i] << "," << grad[i] << "," << time_ms[i] << "\n";
}

// ---- UnifiedOptimizer<CudaBackend> and strategies (src/unified_optimization.hpp:420-633) -----------------------------
template <typename Backend> class UnifiedOptimizer;
template <> class UnifiedOptimizer<CudaBackend> {
public:
  virtual ~UnifiedOptimizer() = default;
  virtual void optimize(cuda_mlp::CublasHandle &handle, NetworkWrapper<CudaBackend> &net, const UnifiedDataset &dataset,
                        cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &d_train_x, cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &d_train_y,
                        const UnifiedConfig &config) = 0;
  int last_iterations = 0;
};

template <typename SolverFactory>
inline int run_cuda_solver_once(SolverFactory make_solver, cuda_mlp::CudaNetwork &net, cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &d_train_x,
                                cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &d_train_y, const UnifiedDataset &dataset,
                                const UnifiedConfig &config) {
  using namespace cuda_mlp;
  net.setPrecision(config.precision);
  // the reference's lambda evaluates the network's own bound buffer and copies the gradient out
  // (src/unified_optimization.hpp:483-491); setNetwork() selects the same objective without the extra copy
  auto loss_grad = [&](const CudaScalar *, CudaScalar *grad, const CudaScalar *input, const CudaScalar *target, int batch) -> CudaScalar {
    CudaScalar loss = net.compute_loss_and_grad(input, target, batch);
    device_copy(grad, net.grads_data(), net.params_size());
    return loss;
  };
  auto solver = make_solver();
  solver->setNetwork(&net);
  IterationRecorder<CudaBackend> recorder;
  recorder.init(config.max_iters + 1);
  solver->setRecorder(&recorder);
  solver->solve((int)net.params_size(), net.params_data(), d_train_x.data(), d_train_y.data(), (int)dataset.train_x.cols(), loss_grad);
  net.handle().synchronize();
  write_cuda_history_csv(cuda_log_filename(config), recorder, config.log_interval);
  return solver->iterations();
}

class UnifiedGD_CUDA : public UnifiedOptimizer<CudaBackend> {
public:
  void optimize(cuda_mlp::CublasHandle &handle, NetworkWrapper<CudaBackend> &net, const UnifiedDataset &d,
                cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &dx, cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &dy, const UnifiedConfig &c) override {
    last_iterations = run_cuda_solver_once([&]() {
      auto s = std::make_unique<cuda_mlp::CudaGD>(handle);
      s->setLearningRate((float)c.learning_rate); s->setMomentum((float)c.momentum);
      s->setMaxIterations(c.max_iters); s->setTolerance((float)c.tolerance);
      return s; }, net.getInternal(), dx, dy, d, c);
  }
};
class UnifiedLBFGS_CUDA : public UnifiedOptimizer<CudaBackend> {
public:
  void optimize(cuda_mlp::CublasHandle &handle, NetworkWrapper<CudaBackend> &net, const UnifiedDataset &d,
                cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &dx, cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &dy, const UnifiedConfig &c) override {
    last_iterations = run_cuda_solver_once([&]() {
      auto s = std::make_unique<cuda_mlp::CudaLBFGS>(handle);
      s->setMemory(c.m_param); s->setMaxIterations(c.max_iters); s->setTolerance((float)c.tolerance);
      s->setLineSearchPolicy(c.linesearch);
      if (c.linesearch == cuda_mlp::LineSearch::Wolfe) s->setLineSearchParams(50, 1e-4f, 0.5f);
      return s; }, net.getInternal(), dx, dy, d, c);
  }
};
class UnifiedSGD_CUDA : public UnifiedOptimizer<CudaBackend> {
public:
  void optimize(cuda_mlp::CublasHandle &handle, NetworkWrapper<CudaBackend> &net, const UnifiedDataset &d,
                cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &dx, cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &dy, const UnifiedConfig &c) override {
    last_iterations = run_cuda_solver_once([&]() {
      auto s = std::make_unique<cuda_mlp::CudaSGD>(handle);
      s->setLearningRate((float)c.learning_rate); s->setMomentum((float)c.momentum); s->setBatchSize(c.batch_size);
      s->setMaxIterations(c.max_iters); s->setLearningRateDecay((float)c.lr_decay, c.lr_decay_rate);
      s->setDimensions((int)d.train_x.rows(), (int)d.train_y.rows());
      return s; }, net.getInternal(), dx, dy, d, c);
  }
};
/// New: available on the GPU (the reference static_asserts, src/unified_optimization.hpp:639-641,688-696)
class UnifiedSLBFGS_CUDA : public UnifiedOptimizer<CudaBackend> {
public:
  void optimize(cuda_mlp::CublasHandle &handle, NetworkWrapper<CudaBackend> &net, const UnifiedDataset &d,
                cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &dx, cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &dy, const UnifiedConfig &c) override {
    last_iterations = run_cuda_solver_once([&]() {
      auto s = std::make_unique<cuda_mlp::CudaSLBFGS>(handle);
      s->setMaxIterations(c.max_iters); s->setTolerance((float)c.tolerance); s->setStepSize((float)c.learning_rate);
      s->setBatchSize(c.batch_size); s->setMemory(c.m_param); s->setUpdateInterval(c.L_param); s->setHessianBatchSize(c.b_H_param);
      return s; }, net.getInternal(), dx, dy, d, c);
  }
};

template <typename Backend> struct UnifiedGD_Selector;
template <> struct UnifiedGD_Selector<CudaBackend> { using type = UnifiedGD_CUDA; };
template <typename Backend> struct UnifiedLBFGS_Selector;
template <> struct UnifiedLBFGS_Selector<CudaBackend> { using type = UnifiedLBFGS_CUDA; };
template <typename Backend> struct UnifiedSGD_Selector;
template <> struct UnifiedSGD_Selector<CudaBackend> { using type = UnifiedSGD_CUDA; };
template <typename Backend> struct UnifiedSLBFGS_Selector;
template <> struct UnifiedSLBFGS_Selector<CudaBackend> { using type = UnifiedSLBFGS_CUDA; };
template <typename Backend> using UnifiedGD = typename UnifiedGD_Selector<Backend>::type;
template <typename Backend> using UnifiedLBFGS = typename UnifiedLBFGS_Selector<Backend>::type;
template <typename Backend> using UnifiedSGD = typename UnifiedSGD_Selector<Backend>::type;
template <typename Backend> using UnifiedSLBFGS = typename UnifiedSLBFGS_Selector<Backend>::type;

// ---- UnifiedLauncher<CudaBackend> (src/unified_launcher.hpp:83-205) ------------------------------------------------
template <typename Backend> class UnifiedLauncher;
template <> class UnifiedLauncher<CudaBackend> {
public:
  UnifiedLauncher() : net_wrapper_(handle_) {}
  template <int In, int Out, typename Activation> void addLayer() { net_wrapper_.addLayer<In, Out, Activation>(); }
  void buildNetwork() { net_wrapper_.bindParams(); }
  void setData(const UnifiedDataset &data) {
    dataset_ = data;
    // double -> float on the device (the reference converts element-wise on the host, unified_launcher.hpp:109-121)
    auto upload = [&](const b200::Matrix &m, cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &dev) {
      if (m.size() == 0) return;
      cuda_mlp::DeviceBuffer<double> staging((size_t)m.size());
      staging.copy_from_host(m.data(), (size_t)m.size());
      dev.resize((size_t)m.size());
      cuda_mlp::b200_check(b200_convert_f64_to_f32(handle_.get(), staging.data(), dev.data(), (size_t)m.size()), "convert");
      handle_.synchronize();
    };
    upload(dataset_.train_x, d_train_x_); upload(dataset_.train_y, d_train_y_);
    upload(dataset_.test_x, d_test_x_); upload(dataset_.test_y, d_test_y_);
    std::cout << "Data Uploaded to GPU. Train: " << dataset_.train_x.cols() << " samples." << std::endl;
  }
  void train(UnifiedOptimizer<CudaBackend> &optimizer, const UnifiedConfig &config) {
    std::cout << ">>> Running CUDA Experiment: " << config.name << std::endl;
    if (config.reset_params) net_wrapper_.bindParams(config.seed);
    optimizer.optimize(handle_, net_wrapper_, dataset_, d_train_x_, d_train_y_, config);
    evaluate(d_train_x_, d_train_y_, dataset_.train_x.cols(), "Training Results");
  }
  void test() { evaluate(d_test_x_, d_test_y_, dataset_.test_x.cols(), "Test Results"); }
  NetworkWrapper<CudaBackend> &network() { return net_wrapper_; }
  double last_mse = 0, last_accuracy = 0;
private:
  void evaluate(cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &dx, cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> &dy, long batch, const char *label) {
    if (batch <= 0) return;
    net_wrapper_.getInternal().evaluate(dx.data(), dy.data(), batch, &last_mse, &last_accuracy); // arg-max + MSE on the device
    std::cout << label << ": MSE=" << last_mse << ", Accuracy=" << last_accuracy << "%" << std::endl;
  }
  cuda_mlp::CublasHandle handle_;
  NetworkWrapper<CudaBackend> net_wrapper_;
  UnifiedDataset dataset_;
  cuda_mlp::DeviceBuffer<cuda_mlp::CudaScalar> d_train_x_, d_train_y_, d_test_x_, d_test_y_;
};

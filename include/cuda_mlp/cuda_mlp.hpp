// cuda_mlp.hpp — C++ host-side mirror of the reference's CUDA operator interface (namespace cuda_mlp),
// implemented over the C ABI of libb200lbfgs.so (include/b200_lbfgs.h). Header-only, plain C++17: it
// compiles with g++ or nvcc and needs no CUDA headers, so the reference's call sites
//   src/unified_optimization.hpp:470-633, src/unified_launcher.hpp:83-205, tests/*/main-gpu*.cpp
// build unchanged against it. Same class names, member names, argument meaning and error behaviour
// (non-zero status -> message on std::cerr + std::abort(), src/cuda/common.cuh:18-23).
//
//   reference header                     class here
//   src/cuda/common.cuh                  CudaScalar, cuda_check
//   src/cuda/cublas_handle.cuh           CublasHandle   (opaque context shim; nothing links cuBLAS)
//   src/cuda/device_buffer.cuh           DeviceBuffer<T>
//   src/cuda/kernels.cuh                 ActivationType, device_dot/nrm2/axpy/scal/copy/set_zero
//   src/cuda/network.cuh                 CudaNetwork
//   src/cuda/minimizer_base.cuh          CudaMinimizerBase, LossGradFun
//   src/cuda/lbfgs.cuh, gd.cuh, sgd.cuh  CudaLBFGS, CudaGD, CudaSGD
//   (new)                                CudaSLBFGS — SLBFGS::stochastic_solve on the GPU
#pragma once

#include "../b200_lbfgs.h"

#include <cstddef>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <utility>
#include <vector>

template <typename Backend> class IterationRecorder; // include/unified/unified.hpp
struct CudaBackend;

namespace cuda_mlp {

using CudaScalar = float; // src/cuda/common.cuh:11

inline void b200_check(int status, const char *what) { // cuda_check / cublas_check semantics
  if (status != B200_OK) {
    std::cerr << "B200 error: " << what << " (" << b200_last_error() << ")\n";
    std::abort();
  }
}

enum class ActivationType : int { Linear = 0, Tanh = 1, ReLU = 2, Sigmoid = 3 }; // src/cuda/kernels.cuh:53-58
enum class Precision : int { FP32 = B200_PREC_FP32, TF32x3 = B200_PREC_TF32X3, TF32 = B200_PREC_TF32 };
enum class LineSearch : int { Armijo = B200_LS_ARMIJO, Wolfe = B200_LS_WOLFE };

/// RAII context: device, stream, optional NCCL communicator. Keeps the reference's type name so that
/// `CublasHandle handle; CudaNetwork net(handle); CudaLBFGS solver(handle);` compiles unchanged.
class CublasHandle {
public:
  explicit CublasHandle(int device = 0) { b200_check(b200_ctx_create(device, &ctx_), "b200_ctx_create"); }
  ~CublasHandle() { b200_ctx_destroy(ctx_); }
  CublasHandle(const CublasHandle &) = delete;
  CublasHandle &operator=(const CublasHandle &) = delete;
  b200_ctx *get() const { return ctx_; }
  void synchronize() const { b200_check(b200_ctx_synchronize(ctx_), "b200_ctx_synchronize"); }
  /// multi-GPU (one process per GPU): every rank passes the 128-byte id created by rank 0
  static void uniqueId(void *out128) { b200_check(b200_comm_unique_id(out128), "b200_comm_unique_id"); }
  void initComm(const void *id128, int rank, int world) {
    b200_check(b200_ctx_init_comm(ctx_, id128, rank, world), "b200_ctx_init_comm");
  }
  int rank() const { return b200_ctx_rank(ctx_); }
  int world() const { return b200_ctx_world(ctx_); }

private:
  b200_ctx *ctx_ = nullptr;
};

/// Move-only RAII device array (src/cuda/device_buffer.cuh:7-96)
template <typename T> class DeviceBuffer {
public:
  DeviceBuffer() = default;
  explicit DeviceBuffer(size_t count) { resize(count); }
  ~DeviceBuffer() { release(); }
  DeviceBuffer(const DeviceBuffer &) = delete;
  DeviceBuffer &operator=(const DeviceBuffer &) = delete;
  DeviceBuffer(DeviceBuffer &&o) noexcept : ptr_(o.ptr_), size_(o.size_) { o.ptr_ = nullptr; o.size_ = 0; }
  DeviceBuffer &operator=(DeviceBuffer &&o) noexcept {
    if (this != &o) { release(); ptr_ = o.ptr_; size_ = o.size_; o.ptr_ = nullptr; o.size_ = 0; }
    return *this;
  }
  void resize(size_t count) { // realloc, contents not preserved
    if (count == size_) return;
    release();
    if (count) {
      void *p = nullptr;
      b200_check(b200_malloc(&p, count * sizeof(T)), "cudaMalloc");
      ptr_ = static_cast<T *>(p);
    }
    size_ = count;
  }
  T *data() { return ptr_; }
  const T *data() const { return ptr_; }
  size_t size() const { return size_; }
  void copy_from_host(const T *host, size_t count) {
    if (count != size_) resize(count);
    b200_check(b200_memcpy_h2d(ptr_, host, count * sizeof(T)), "cudaMemcpy H2D");
  }
  void copy_to_host(T *host, size_t count) const { b200_check(b200_memcpy_d2h(host, ptr_, count * sizeof(T)), "cudaMemcpy D2H"); }

private:
  void release() {
    if (ptr_) b200_free(ptr_);
    ptr_ = nullptr;
    size_ = 0;
  }
  T *ptr_ = nullptr;
  size_t size_ = 0;
};

// ---- BLAS-1 wrappers (src/cuda/kernels.cuh:14-50); dot / nrm2 synchronise like the reference's ----
inline void device_copy(CudaScalar *dst, const CudaScalar *src, size_t n) { b200_check(b200_memcpy_d2d(dst, src, n * sizeof(CudaScalar)), "device_copy"); }
inline void device_set_zero(CudaScalar *p, size_t n) { b200_check(b200_memset(p, 0, n * sizeof(CudaScalar)), "device_set_zero"); }
inline CudaScalar device_dot(CublasHandle &h, const CudaScalar *x, const CudaScalar *y, int n) {
  double r = 0;
  b200_check(b200_vec_dot(h.get(), x, y, (size_t)n, &r), "device_dot");
  return (CudaScalar)r;
}
inline CudaScalar device_nrm2(CublasHandle &h, const CudaScalar *x, int n) {
  double r = 0;
  b200_check(b200_vec_nrm2(h.get(), x, (size_t)n, &r), "device_nrm2");
  return (CudaScalar)r;
}
inline void device_axpy(CublasHandle &h, int n, CudaScalar a, const CudaScalar *x, CudaScalar *y) { b200_check(b200_vec_axpy(h.get(), (size_t)n, a, x, y), "device_axpy"); }
inline void device_scal(CublasHandle &h, int n, CudaScalar a, CudaScalar *x) { b200_check(b200_vec_scal(h.get(), (size_t)n, a, x), "device_scal"); }

/// The MLP objective (src/cuda/network.cuh:16-156)
class CudaNetwork {
public:
  explicit CudaNetwork(CublasHandle &handle) : handle_(handle) {}
  ~CudaNetwork() { b200_net_destroy(net_); }
  CudaNetwork(const CudaNetwork &) = delete;
  CudaNetwork &operator=(const CudaNetwork &) = delete;

  void addLayer(int in, int out, ActivationType act) {
    if (net_) { std::cerr << "B200 error: addLayer after bindParams\n"; std::abort(); }
    if (dims_.empty()) dims_.push_back(in);
    dims_.push_back(out);
    acts_.push_back((int)act);
  }
  void bindParams(unsigned int seed = 123) {
    create();
    b200_check(b200_net_bind_params(net_, seed), "bindParams");
  }
  size_t params_size() const { const_cast<CudaNetwork *>(this)->create(); return b200_net_params_size(net_); }
  int output_size() const { const_cast<CudaNetwork *>(this)->create(); return b200_net_output_size(net_); }
  CudaScalar *params_data() { return b200_net_params_data(net_); }
  CudaScalar *grads_data() { return b200_net_grads_data(net_); }
  void zeroGrads() { b200_check(b200_net_zero_grads(net_), "zeroGrads"); }
  void forward_only(const CudaScalar *input, int batch) { b200_check(b200_net_forward(net_, input, batch), "forward_only"); }
  CudaScalar compute_loss_and_grad(const CudaScalar *input, const CudaScalar *target, int batch) {
    float loss = 0.f;
    b200_check(b200_net_loss_grad(net_, input, target, batch, &loss), "compute_loss_and_grad");
    return loss;
  }
  void copy_output_to_host(CudaScalar *host, size_t n) { b200_check(b200_net_copy_output_to_host(net_, host, n), "copy_output_to_host"); }
  int last_batch() const { return b200_net_last_batch(net_); }
  // additions
  void setPrecision(Precision p) { create(); b200_check(b200_net_set_precision(net_, (int)p), "setPrecision"); }
  void setL2(float lambda) { create(); b200_check(b200_net_set_l2(net_, lambda), "setL2"); }
  void setGlobalBatch(long b) { create(); b200_check(b200_net_set_global_batch(net_, b), "setGlobalBatch"); }
  void evaluate(const CudaScalar *x, const CudaScalar *t, long batch, double *mse, double *acc) {
    b200_check(b200_net_evaluate(net_, x, t, batch, mse, acc), "evaluate");
  }
  b200_net *get() { create(); return net_; }
  CublasHandle &handle() { return handle_; }

private:
  void create() {
    if (net_) return;
    if (acts_.empty()) { std::cerr << "B200 error: network has no layers\n"; std::abort(); }
    b200_check(b200_net_create(handle_.get(), (int)acts_.size(), dims_.data(), acts_.data(), &net_), "b200_net_create");
  }
  CublasHandle &handle_;
  std::vector<int> dims_, acts_;
  b200_net *net_ = nullptr;
};

/// Common base of the GPU minimizers (src/cuda/minimizer_base.cuh:12-67)
class CudaMinimizerBase {
public:
  using LossGradFun = std::function<CudaScalar(const CudaScalar *, CudaScalar *, const CudaScalar *, const CudaScalar *, int)>;
  explicit CudaMinimizerBase(CublasHandle &handle) : handle_(handle) {}
  virtual ~CudaMinimizerBase() = default;
  void setMaxIterations(int iters) { max_iters_ = iters; }
  void setTolerance(CudaScalar tol) { tol_ = tol; }
  int iterations() const noexcept { return last_iterations_; }
  long evaluations() const noexcept { return last_evaluations_; }
  void setLineSearchParams(int max_iters, CudaScalar c1, CudaScalar rho) { // minimizer_base.cuh:38-45: at least one trial
    max_line_iters_ = max_iters < 1 ? 1 : max_iters; c1_ = c1; rho_ = rho;
  }
  void setRecorder(::IterationRecorder<CudaBackend> *recorder) { recorder_ = recorder; }
  /// Fast path: the library's own network objective evaluated at `params` (no callback, loss stays on the
  /// device between kernels). run_cuda_solver_once's lambda (src/unified_optimization.hpp:483-491) is exactly this.
  void setNetwork(CudaNetwork *net) { net_ = net; }

  virtual void solve(int n, CudaScalar *params, const CudaScalar *input, const CudaScalar *target, int batch,
                     const LossGradFun &loss_grad) = 0;

protected:
  static float trampoline(void *user, const float *params, float *grad, const float *input, const float *target, int batch) {
    return (*static_cast<const LossGradFun *>(user))(params, grad, input, target, batch);
  }
  void begin_history(b200_history &h); // defined in include/unified/unified.hpp (needs IterationRecorder)
  void end_history(const b200_history &h);

  CublasHandle &handle_;
  CudaNetwork *net_ = nullptr;
  int max_iters_ = 200;      // minimizer_base.cuh:61-65
  int max_line_iters_ = 20;
  CudaScalar tol_ = 1e-6f, c1_ = 1e-4f, rho_ = 0.5f;
  ::IterationRecorder<CudaBackend> *recorder_ = nullptr;
  int last_iterations_ = 0;
  long last_evaluations_ = 0;
};

class CudaLBFGS : public CudaMinimizerBase { // src/cuda/lbfgs.cuh:23-264
public:
  explicit CudaLBFGS(CublasHandle &handle) : CudaMinimizerBase(handle) {}
  void setMemory(size_t m) { m_ = m; }
  /// Armijo = reference CUDA backend (default); Wolfe = reference CPU backend (src/minimizer/full_batch_minimizer.hpp:126-157)
  void setLineSearchPolicy(LineSearch p, CudaScalar c2 = 0.9f) { ls_ = p; c2_ = c2; }
  void solve(int n, CudaScalar *params, const CudaScalar *input, const CudaScalar *target, int batch,
             const LossGradFun &loss_grad) override {
    b200_lbfgs_opts o;
    b200_lbfgs_default_opts(&o);
    o.max_iters = max_iters_; o.tol = tol_; o.memory = (int)m_; o.max_line_iters = max_line_iters_;
    o.c1 = c1_; o.rho = rho_; o.c2 = c2_; o.linesearch = (int)ls_;
    b200_history h{};
    begin_history(h);
    b200_check(b200_lbfgs_solve(handle_.get(), net_ ? net_->get() : nullptr, net_ ? nullptr : &trampoline,
                                const_cast<LossGradFun *>(&loss_grad), n, params, input, target, batch, &o, &h),
               "CudaLBFGS::solve");
    end_history(h);
  }

private:
  size_t m_ = 16; // lbfgs.cuh:263
  LineSearch ls_ = LineSearch::Armijo;
  CudaScalar c2_ = 0.9f;
};

class CudaGD : public CudaMinimizerBase { // src/cuda/gd.cuh:18-111
public:
  explicit CudaGD(CublasHandle &handle) : CudaMinimizerBase(handle) {}
  void setLearningRate(CudaScalar lr) { lr_ = lr; }
  void setMomentum(CudaScalar m) { momentum_ = m; }
  void solve(int n, CudaScalar *params, const CudaScalar *input, const CudaScalar *target, int total_samples,
             const LossGradFun &loss_grad) override {
    b200_gd_opts o;
    b200_gd_default_opts(&o);
    o.max_iters = max_iters_; o.tol = tol_; o.lr = lr_; o.momentum = momentum_;
    b200_history h{};
    begin_history(h);
    b200_check(b200_gd_solve(handle_.get(), net_ ? net_->get() : nullptr, net_ ? nullptr : &trampoline,
                             const_cast<LossGradFun *>(&loss_grad), n, params, input, target, total_samples, &o, &h),
               "CudaGD::solve");
    end_history(h);
  }

private:
  CudaScalar lr_ = 0.01f, momentum_ = 0.9f;
};

class CudaSGD : public CudaMinimizerBase { // src/cuda/sgd.cuh:20-164
public:
  explicit CudaSGD(CublasHandle &handle) : CudaMinimizerBase(handle) {}
  void setLearningRate(CudaScalar lr) { lr_ = lr; }
  void setMomentum(CudaScalar m) { momentum_ = m; }
  void setBatchSize(int b) { batch_size_ = b; }
  void setLearningRateDecay(CudaScalar rate, int step) { decay_rate_ = rate; decay_step_ = step; }
  void setDimensions(int in, int out) { input_dim_ = in; output_dim_ = out; }
  /// Sequential (default) = the CUDA backend's slices; Random = the CPU backend's mini-batches (src/minimizer/s_gd.hpp:63-170) on the GPU
  enum class Sampling { Sequential = 0, Random = 1 };
  void setSampling(Sampling s, unsigned seed = 123) { sampling_ = s; seed_ = seed; }
  void solve(int n, CudaScalar *params, const CudaScalar *input, const CudaScalar *target, int total_samples,
             const LossGradFun &loss_grad) override {
    b200_sgd_opts o;
    b200_sgd_default_opts(&o);
    o.sampling = (int)sampling_; o.seed = seed_;
    o.max_iters = max_iters_; o.tol = tol_; o.lr = lr_; o.momentum = momentum_; o.decay_rate = decay_rate_;
    o.decay_step = decay_step_; o.batch_size = batch_size_; o.input_dim = input_dim_; o.output_dim = output_dim_;
    b200_history h{};
    begin_history(h);
    b200_check(b200_sgd_solve(handle_.get(), net_ ? net_->get() : nullptr, net_ ? nullptr : &trampoline,
                              const_cast<LossGradFun *>(&loss_grad), n, params, input, target, total_samples, &o, &h),
               "CudaSGD::solve");
    end_history(h);
  }

private:
  CudaScalar lr_ = 0.01f, momentum_ = 0.9f, decay_rate_ = 1.0f;
  int decay_step_ = 0, batch_size_ = 64, input_dim_ = 0, output_dim_ = 0;
  Sampling sampling_ = Sampling::Sequential;
  unsigned seed_ = 123;
};

/// S-LBFGS on the GPU: SLBFGS::stochastic_solve (src/minimizer/s_lbfgs.hpp:165-290) with the objective of
/// UnifiedSLBFGS_CPU::optimize (lambda = 1e-4). Needs setNetwork(); the callback argument is ignored.
class CudaSLBFGS : public CudaMinimizerBase {
public:
  explicit CudaSLBFGS(CublasHandle &handle) : CudaMinimizerBase(handle) { tol_ = 1e-4f; }
  void setStepSize(CudaScalar s) { step_ = s; }
  void setBatchSize(int b) { batch_ = b; }
  void setMemory(int m) { M_ = m; }
  void setUpdateInterval(int L) { L_ = L; }
  void setHessianBatchSize(int b) { b_H_ = b; }
  void setHvpStepScale(CudaScalar s) { hvp_scale_ = s; } // b200_slbfgs_opts::hvp_step_scale (1 = the reference's step verbatim)
  void solve(int n, CudaScalar *params, const CudaScalar *input, const CudaScalar *target, int total_samples,
             const LossGradFun &) override {
    if (!net_) { std::cerr << "B200 error: CudaSLBFGS needs setNetwork()\n"; std::abort(); }
    b200_slbfgs_opts o;
    b200_slbfgs_default_opts(&o);
    o.max_iters = max_iters_; o.tol = tol_; o.step_size = step_; o.batch_size = batch_; o.memory = M_; o.L = L_;
    o.b_H = b_H_; o.record = recorder_ ? 1 : 0; o.hvp_step_scale = hvp_scale_;
    b200_history h{};
    begin_history(h);
    b200_check(b200_slbfgs_solve(handle_.get(), net_->get(), n, params, input, target, total_samples, &o, &h), "CudaSLBFGS::solve");
    end_history(h);
  }

private:
  CudaScalar step_ = 0.01f, hvp_scale_ = 256.0f;
  int batch_ = 128, M_ = 10, L_ = 10, b_H_ = 0;
};

} // namespace cuda_mlp

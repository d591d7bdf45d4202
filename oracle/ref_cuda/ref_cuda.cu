// oracle/ref_cuda — the REFERENCE's own CUDA backend, compiled from its sources where they lie under
// /root/reference (never copied into this repo), behind a tiny extern "C" shim so the GPU tests can run the
// genuine reference kernels (cuBLAS SGEMM + src/cuda/kernels.cuh + CudaLBFGS/CudaGD/CudaSGD) on the same inputs.
//
// TEST INFRASTRUCTURE ONLY: output goes to oracle/_ref/libref_cuda.so (git-ignored, travels to the GPU box).
// Only tests/ and bench.py's comparator leg load it; nothing under lbfgs_ffnn_b200/ does. It links cuBLAS —
// that is the reference's dependency, not the product's.
//
// The reference headers need only the two backend tag types (src/iteration_recorder.hpp:6-7).
struct CpuBackend {};
struct CudaBackend {};

#include "src/cuda/gd.cuh"
#include "src/cuda/lbfgs.cuh"
#include "src/cuda/network.cuh"
#include "src/cuda/sgd.cuh"

#include <cstring>
#include <memory>

using namespace cuda_mlp;

struct RefNet {
  CublasHandle handle;
  CudaNetwork net;
  RefNet() : net(handle) {}
};

extern "C" {

void *ref_net_create(int nlayers, const int *dims, const int *acts) {
  RefNet *r = new RefNet;
  for (int l = 0; l < nlayers; ++l) r->net.addLayer(dims[l], dims[l + 1], (ActivationType)acts[l]);
  r->net.bindParams(123);
  return r;
}
void ref_net_destroy(void *h) { delete (RefNet *)h; }
long ref_net_params_size(void *h) { return (long)((RefNet *)h)->net.params_size(); }
void ref_net_bind_params(void *h, unsigned seed) { ((RefNet *)h)->net.bindParams(seed); }
void ref_net_set_params(void *h, const float *host) {
  RefNet *r = (RefNet *)h;
  cudaMemcpy(r->net.params_data(), host, sizeof(float) * r->net.params_size(), cudaMemcpyHostToDevice);
}
void ref_net_get_params(void *h, float *host) {
  RefNet *r = (RefNet *)h;
  cudaMemcpy(host, r->net.params_data(), sizeof(float) * r->net.params_size(), cudaMemcpyDeviceToHost);
}
// CudaNetwork::compute_loss_and_grad (src/cuda/network.cuh:97-119) on device buffers x, t
float ref_net_loss_grad(void *h, const float *x_dev, const float *t_dev, int batch, float *grad_host) {
  RefNet *r = (RefNet *)h;
  const float loss = r->net.compute_loss_and_grad(x_dev, t_dev, batch);
  if (grad_host) cudaMemcpy(grad_host, r->net.grads_data(), sizeof(float) * r->net.params_size(), cudaMemcpyDeviceToHost);
  return loss;
}
void ref_net_forward(void *h, const float *x_dev, int batch, float *out_host) {
  RefNet *r = (RefNet *)h;
  r->net.forward_only(x_dev, batch);
  r->net.copy_output_to_host(out_host, (size_t)r->net.output_size() * batch);
}

// run_cuda_solver_once (src/unified_optimization.hpp:470-515) with the reference's solvers.
// kind: 0 = CudaLBFGS, 1 = CudaGD, 2 = CudaSGD. Returns iterations(); history arrays hold `cap` entries.
int ref_solve(void *h, int kind, const float *x_dev, const float *t_dev, int batch, int max_iters, float tol, int memory,
              float lr, float momentum, int sgd_batch, float decay_rate, int decay_step, int in_dim, int out_dim, int cap,
              float *loss_hist, float *grad_hist, float *time_hist, int *hist_size, float *total_ms) {
  RefNet *r = (RefNet *)h;
  CudaNetwork &net = r->net;
  auto loss_grad = [&](const CudaScalar *, CudaScalar *grad, const CudaScalar *input, const CudaScalar *target, int b) -> CudaScalar {
    CudaScalar loss = net.compute_loss_and_grad(input, target, b);
    device_copy(grad, net.grads_data(), net.params_size());
    return loss;
  };
  std::unique_ptr<CudaMinimizerBase> solver;
  if (kind == 0) {
    auto s = std::make_unique<CudaLBFGS>(r->handle);
    s->setMemory(memory);
    solver = std::move(s);
  } else if (kind == 1) {
    auto s = std::make_unique<CudaGD>(r->handle);
    s->setLearningRate(lr); s->setMomentum(momentum);
    solver = std::move(s);
  } else {
    auto s = std::make_unique<CudaSGD>(r->handle);
    s->setLearningRate(lr); s->setMomentum(momentum); s->setBatchSize(sgd_batch);
    s->setLearningRateDecay(decay_rate, decay_step); s->setDimensions(in_dim, out_dim);
    solver = std::move(s);
  }
  solver->setMaxIterations(max_iters);
  solver->setTolerance(tol);
  IterationRecorder<CudaBackend> recorder;
  recorder.init(cap);
  if (cap > 0) solver->setRecorder(&recorder);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  solver->solve((int)net.params_size(), net.params_data(), x_dev, t_dev, batch, loss_grad);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (total_ms) *total_ms = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (cap > 0) {
    std::vector<CudaScalar> l, g, t;
    recorder.copy_to_host(l, g, t);
    const int n = (int)std::min<size_t>(l.size(), (size_t)cap);
    if (loss_hist) std::memcpy(loss_hist, l.data(), sizeof(float) * n);
    if (grad_hist) std::memcpy(grad_hist, g.data(), sizeof(float) * n);
    if (time_hist) std::memcpy(time_hist, t.data(), sizeof(float) * n);
    if (hist_size) *hist_size = n;
  }
  return solver->iterations();
}

} // extern "C"

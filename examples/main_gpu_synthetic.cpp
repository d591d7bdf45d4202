// The reference's GPU runner (tests/mnist/main-gpu.cpp:1-110) against this backend, with the missing MNIST image
// blobs replaced by the synthetic generator (same recipe as lbfgs_ffnn_b200/data.py: one std::mt19937(123) stream).
// Build: make -C examples      Run: ./examples/main_gpu_synthetic [train_size] [max_iters]
#include "../include/unified/unified_launcher.hpp"

#include <cstdlib>
#include <random>

using Backend = CudaBackend;

static void synthetic_mnist(long n, b200::Matrix &x, b200::Matrix &y, std::mt19937 &gen) {
  x.resize(784, n);
  y.resize(10, n);
  for (long s = 0; s < n; ++s) {
    const unsigned label = gen() % 10u;
    y(label, s) = 1.0;
    for (int p = 0; p < 784; ++p) {
      const unsigned r = gen();
      const unsigned pix = (r % 5u == 0u) ? (r >> 24) : 0u;
      x(p, s) = (double)((float)pix / 255.0f); // tests/mnist/mnist_loader.hpp:59
    }
  }
}

int main(int argc, char **argv) {
  const long train_size = argc > 1 ? std::atol(argv[1]) : 60000;
  const int max_iters = argc > 2 ? std::atoi(argv[2]) : 100;
  UnifiedLauncher<Backend> launcher;
  std::cout << "Building Network..." << std::endl;
  launcher.addLayer<784, 128, cpu_mlp::ReLU>();
  launcher.addLayer<128, 10, cpu_mlp::Linear>();
  launcher.buildNetwork();

  std::mt19937 gen(kDefaultSeed);
  UnifiedDataset dataset;
  synthetic_mnist(train_size, dataset.train_x, dataset.train_y, gen);
  synthetic_mnist(train_size / 6, dataset.test_x, dataset.test_y, gen);
  launcher.setData(dataset);
  {
    UnifiedConfig config;
    config.name = "SYN_GD"; config.max_iters = max_iters; config.tolerance = 1e-3; config.learning_rate = 0.02;
    config.momentum = 0.9; config.log_interval = 1;
    UnifiedGD<Backend> optimizer;
    launcher.train(optimizer, config);
    launcher.test();
  }
  {
    UnifiedConfig config;
    config.name = "SYN_SGD"; config.max_iters = std::max(1, max_iters / 20); config.tolerance = 1e-3; config.learning_rate = 0.01;
    config.batch_size = 256; config.log_interval = 1; config.lr_decay = 0.80; config.lr_decay_rate = 40;
    UnifiedSGD<Backend> optimizer;
    launcher.train(optimizer, config);
    launcher.test();
  }
  {
    UnifiedConfig config;
    config.name = "SYN_LBFGS_m10"; config.max_iters = max_iters; config.tolerance = 1e-3; config.m_param = 10; config.log_interval = 1;
    UnifiedLBFGS<Backend> optimizer;
    launcher.train(optimizer, config);
    launcher.test();
  }
  {
    UnifiedConfig config;
    config.name = "SYN_LBFGS_m10_tf32x3"; config.max_iters = max_iters; config.tolerance = 1e-3; config.m_param = 10; config.log_interval = 1;
    config.precision = cuda_mlp::Precision::TF32x3;
    UnifiedLBFGS<Backend> optimizer;
    launcher.train(optimizer, config);
    launcher.test();
  }
  {
    UnifiedConfig config;
    config.name = "SYN_SLBFGS"; config.max_iters = std::max(1, max_iters / 20); config.tolerance = 1e-4; config.learning_rate = 0.02;
    config.batch_size = 1000; config.m_param = 10; config.L_param = 10; config.b_H_param = 5000; config.log_interval = 1;
    UnifiedSLBFGS<Backend> optimizer; // static_assert in the reference; available here
    launcher.train(optimizer, config);
    launcher.test();
  }
  return 0;
}

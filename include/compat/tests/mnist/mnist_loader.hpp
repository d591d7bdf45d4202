// drop-in for the reference header tests/mnist/mnist_loader.hpp (the IDX reader the runner mains call).
//
// Same class and entry points — MNISTLoader::loadImages(path, max_images) / loadLabels(path, max_images), pixels as
// float(u) / 255.0f (mnist_loader.hpp:59), labels one-hot in a 10 x count matrix (:87-97) — reading the IDX format (big-endian
// header: magic 2051 / 2049, count, rows, cols). The (Fashion-)MNIST image blobs are not part of the reference checkout
// (.MISSING_LARGE_BLOBS), so a path that cannot be opened does not throw here: the loader says so on stderr and returns
// deterministic synthetic data of the requested size with the same statistics as lbfgs_ffnn_b200/data.py (about 80 % zero
// pixels, uniform labels), so that the reference's runner mains run as they are. An existing file that is not IDX still throws.
#pragma once

#include "../../../unified/unified.hpp"

#include <cstdint>
#include <fstream>
#include <iostream>
#include <random>
#include <stdexcept>
#include <string>

#if !(defined(B200_USE_EIGEN) && __has_include(<Eigen/Core>))
namespace Eigen { using MatrixXd = b200::Matrix; } // the mains declare their matrices as Eigen::MatrixXd
#endif

class MNISTLoader {
  static uint32_t be32(std::ifstream &f) {
    unsigned char b[4] = {0, 0, 0, 0};
    f.read(reinterpret_cast<char *>(b), 4);
    return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3];
  }
  static unsigned seed_of(const std::string &path) { // train and t10k files get different streams
    unsigned h = 2166136261u;
    for (unsigned char c : path) h = (h ^ c) * 16777619u;
    return h;
  }
  static long fallback_count(int max_images) { return max_images > 0 ? max_images : 10000; }

public:
  template <typename Scalar = double> static b200::Matrix loadImages(const std::string &path, int max_images = 0) {
    static_assert(sizeof(Scalar) == sizeof(double), "the drop-in loader produces double matrices (UnifiedDataset)");
    std::ifstream file(path, std::ios::binary);
    if (!file.is_open()) {
      const long count = fallback_count(max_images);
      std::cerr << "[mnist_loader] " << path << " is not there: " << count << " synthetic 28x28 images instead" << std::endl;
      b200::Matrix images(784, count);
      std::mt19937 gen(seed_of(path));
      for (long i = 0; i < count; ++i)
        for (int j = 0; j < 784; ++j) {
          const unsigned r = gen();
          const unsigned pix = (r % 5u == 0u) ? (r >> 24) : 0u;
          images(j, i) = (double)((float)pix / 255.0f);
        }
      return images;
    }
    if (be32(file) != 2051u) throw std::runtime_error("Invalid MNIST image file!");
    const uint32_t total = be32(file), rows = be32(file), cols = be32(file);
    const long image_size = (long)rows * cols;
    const long count = (max_images > 0 && (uint32_t)max_images < total) ? max_images : (long)total;
    std::cout << "Loading " << count << " images..." << std::endl;
    b200::Matrix images(image_size, count);
    std::vector<unsigned char> buf((size_t)image_size);
    for (long i = 0; i < count; ++i) {
      file.read(reinterpret_cast<char *>(buf.data()), image_size);
      for (long j = 0; j < image_size; ++j) images(j, i) = (double)((float)buf[(size_t)j] / 255.0f);
    }
    return images;
  }

  template <typename Scalar = double> static b200::Matrix loadLabels(const std::string &path, int max_images = 0) {
    static_assert(sizeof(Scalar) == sizeof(double), "the drop-in loader produces double matrices (UnifiedDataset)");
    std::ifstream file(path, std::ios::binary);
    if (!file.is_open()) {
      const long count = fallback_count(max_images);
      std::cerr << "[mnist_loader] " << path << " is not there: " << count << " synthetic labels instead" << std::endl;
      b200::Matrix labels(10, count);
      std::mt19937 gen(seed_of(path));
      for (long i = 0; i < count; ++i) labels((long)(gen() % 10u), i) = 1.0;
      return labels;
    }
    if (be32(file) != 2049u) throw std::runtime_error("Invalid MNIST label file!");
    const uint32_t total = be32(file);
    const long count = (max_images > 0 && (uint32_t)max_images < total) ? max_images : (long)total;
    std::cout << "Loading " << count << " labels..." << std::endl;
    b200::Matrix labels(10, count);
    for (long i = 0; i < count; ++i) {
      unsigned char t = 0;
      file.read(reinterpret_cast<char *>(&t), 1);
      if (t < 10) labels((long)t, i) = 1.0;
    }
    return labels;
  }
};

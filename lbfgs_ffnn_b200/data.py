"""Deterministic synthetic MNIST-shaped data (the image blobs are absent from the reference checkout,
SURVEY.md D8). One std::mt19937(seed) stream — numpy's RandomState(seed) is the same generator with the
same init_genrand seeding — consumed per sample as: 1 draw -> label = r % 10, then 784 draws ->
pixel = (r >> 24) if r % 5 == 0 else 0 (~80 % zeros, like MNIST), x = float(pixel) / 255.0f
(tests/mnist/mnist_loader.hpp:59). Layout matches the reference: X is in x B column-major, i.e. (B, in)
C-contiguous; T is the one-hot (B, 10)."""
import numpy as np

kDefaultSeed = 123  # src/seed.hpp:4


def synthetic_mnist(n_samples, in_dim=784, n_classes=10, seed=kDefaultSeed, dtype=np.float32):
    rs = np.random.RandomState(seed)
    x = np.empty((n_samples, in_dim), dtype=dtype)
    t = np.zeros((n_samples, n_classes), dtype=dtype)
    chunk = 4096
    for s in range(0, n_samples, chunk):
        e = min(n_samples, s + chunk)
        raw = rs.randint(0, 2**32, size=(e - s, in_dim + 1), dtype=np.uint64).astype(np.uint32)
        labels = (raw[:, 0] % n_classes).astype(np.int64)
        pix = np.where(raw[:, 1:] % 5 == 0, raw[:, 1:] >> 24, 0).astype(np.float32)
        x[s:e] = (pix / np.float32(255.0)).astype(dtype)
        t[np.arange(s, e), labels] = 1
    return x, t

// TEST INFRASTRUCTURE ONLY — a minimal, eager stand-in for the subset of the Eigen 3 API that the reference's CPU path
// (/root/reference/src/{layer,network,unified_optimization}.hpp, src/minimizer/*.hpp) uses.
//
// Why it exists: the reference's CPU sources include <Eigen/Eigen> (src/common.hpp:8); Eigen 3.4.0 is an un-vendored system
// dependency (enviroment/Dockerfile:14) that this image does not have and cannot fetch. Eigen is used there only for dense
// double-precision linear algebra (GEMM, dot, norm, element-wise maps), so with this header on the include path the
// reference's own sources compile UNMODIFIED (oracle/ref_cpu/Makefile) and their control flow — line searches, ring buffers,
// RNG consumption, closures — runs here exactly as written. What is NOT the reference's is the arithmetic inside the Eigen
// calls (summation order of dot products and GEMMs), which is this file's. Everything is column-major, like Eigen's default.
//
// Nothing here is product code; the product never includes it.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace Eigen {

using Index = std::ptrdiff_t;
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
inline int nbThreads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

namespace shim {

// ---- kernels -------------------------------------------------------------------------------------
inline double dot(const double *a, const double *b, Index n) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  Index i = 0;
  for (; i + 3 < n; i += 4) { s0 += a[i] * b[i]; s1 += a[i + 1] * b[i + 1]; s2 += a[i + 2] * b[i + 2]; s3 += a[i + 3] * b[i + 3]; }
  for (; i < n; ++i) s0 += a[i] * b[i];
  return (s0 + s1) + (s2 + s3);
}

// Dense GEMMs. Register-tiled 8 x 6 micro-kernels on 4-wide double vectors (AVX2 + FMA at -march=x86-64-v3) so that the CPU
// arm of bench.py is not handicapped by the stand-in: Eigen's own GEMM is a tuned, packed kernel of the same kind.
typedef double v4d __attribute__((vector_size(32), aligned(8)));
static inline v4d ld4(const double *p) { return *reinterpret_cast<const v4d *>(p); }
static inline void st4(double *p, v4d v) { *reinterpret_cast<v4d *>(p) = v; }
static inline v4d bc4(double x) { return v4d{x, x, x, x}; }

// acc[8 x 6] (+)= sum_p A[i0 .. i0+8, p] * Bval(p, j) for p in [p0, p1); A column-major (lda); B element (p, j) at B[p * bp + j * bj]
template <bool ACCUM>
static inline void micro_8x6(const double *A, Index lda, const double *B, Index bp, Index bj, Index p0, Index p1, double *C, Index ldc) {
  v4d c00 = bc4(0), c01 = bc4(0), c10 = bc4(0), c11 = bc4(0), c20 = bc4(0), c21 = bc4(0);
  v4d c30 = bc4(0), c31 = bc4(0), c40 = bc4(0), c41 = bc4(0), c50 = bc4(0), c51 = bc4(0);
  for (Index p = p0; p < p1; ++p) {
    const double *a = A + p * lda;
    const v4d a0 = ld4(a), a1 = ld4(a + 4);
    const double *b = B + p * bp;
    v4d x;
    x = bc4(b[0]);      c00 += a0 * x; c01 += a1 * x;
    x = bc4(b[bj]);     c10 += a0 * x; c11 += a1 * x;
    x = bc4(b[2 * bj]); c20 += a0 * x; c21 += a1 * x;
    x = bc4(b[3 * bj]); c30 += a0 * x; c31 += a1 * x;
    x = bc4(b[4 * bj]); c40 += a0 * x; c41 += a1 * x;
    x = bc4(b[5 * bj]); c50 += a0 * x; c51 += a1 * x;
  }
  const v4d r[6][2] = {{c00, c01}, {c10, c11}, {c20, c21}, {c30, c31}, {c40, c41}, {c50, c51}};
  for (int j = 0; j < 6; ++j) {
    double *c = C + j * ldc;
    if (ACCUM) { st4(c, ld4(c) + r[j][0]); st4(c + 4, ld4(c + 4) + r[j][1]); }
    else { st4(c, r[j][0]); st4(c + 4, r[j][1]); }
  }
}
// the same for ragged edges (mr <= 8 rows, nr <= 6 columns)
template <bool ACCUM>
static inline void micro_edge(Index mr, Index nr, const double *A, Index lda, const double *B, Index bp, Index bj, Index p0, Index p1,
                              double *C, Index ldc) {
  double acc[6][8] = {};
  for (Index p = p0; p < p1; ++p) {
    const double *a = A + p * lda, *b = B + p * bp;
    for (Index j = 0; j < nr; ++j) {
      const double x = b[j * bj];
      for (Index i = 0; i < mr; ++i) acc[j][i] += a[i] * x;
    }
  }
  for (Index j = 0; j < nr; ++j)
    for (Index i = 0; i < mr; ++i) C[i + j * ldc] = (ACCUM ? C[i + j * ldc] : 0.0) + acc[j][i];
}
// C tile block (rows [0, m), columns [j0, j1)) (+)= A[:, p0:p1] * Bview[p0:p1, j0:j1]
template <bool ACCUM>
static inline void gemm_block(Index m, Index j0, Index j1, const double *A, Index lda, const double *B, Index bp, Index bj, Index p0,
                              Index p1, double *C, Index ldc) {
  for (Index j = j0; j < j1; j += 6) {
    const Index nr = std::min<Index>(6, j1 - j);
    for (Index i = 0; i < m; i += 8) {
      const Index mr = std::min<Index>(8, m - i);
      if (mr == 8 && nr == 6) micro_8x6<ACCUM>(A + i, lda, B + j * bj, bp, bj, p0, p1, C + i + j * ldc, ldc);
      else micro_edge<ACCUM>(mr, nr, A + i, lda, B + j * bj, bp, bj, p0, p1, C + i + j * ldc, ldc);
    }
  }
}

// C (m x n) = A (m x k) * B (k x n), all column-major with leading dimensions lda, ldb, ldc. Threads split the columns of C;
// K is walked in slices so that the A panel of a slice stays in cache across the column blocks of a thread.
inline void gemm_nn(Index m, Index n, Index k, const double *A, Index lda, const double *B, Index ldb, double *C, Index ldc) {
  const Index kc = 256, nc = 96;
#pragma omp parallel for schedule(static) if (m * n * k > (Index)1 << 16)
  for (Index jb = 0; jb < n; jb += nc) {
    const Index j1 = std::min<Index>(n, jb + nc);
    for (Index p0 = 0; p0 < k; p0 += kc) {
      const Index p1 = std::min<Index>(k, p0 + kc);
      if (p0 == 0) gemm_block<false>(m, jb, j1, A, lda, B, 1, ldb, p0, p1, C, ldc);
      else gemm_block<true>(m, jb, j1, A, lda, B, 1, ldb, p0, p1, C, ldc);
    }
    if (k == 0)
      for (Index j = jb; j < j1; ++j) std::fill(C + j * ldc, C + j * ldc + m, 0.0);
  }
}
// C (m x n) = A^T * B with A stored k x m (column-major): A is the small operand here (a weight matrix), so it is transposed
// once and the product runs through the same kernel.
inline void gemm_tn(Index m, Index n, Index k, const double *A, Index lda, const double *B, Index ldb, double *C, Index ldc) {
  std::vector<double> At((size_t)(m * k));
  for (Index i = 0; i < m; ++i)
    for (Index p = 0; p < k; ++p) At[(size_t)(i + p * m)] = A[p + i * lda];
  gemm_nn(m, n, k, At.data(), m, B, ldb, C, ldc);
}
// C (m x n) = A (m x k) * B^T with B stored n x k (column-major): a sum of k rank-1 updates (the dW product: k = samples).
// Threads take slices of k with private accumulators, walked in chunks that keep both panels in cache, and the accumulators are
// added in thread order.
inline void gemm_nt(Index m, Index n, Index k, const double *A, Index lda, const double *B, Index ldb, double *C, Index ldc) {
  int nt = 1;
#ifdef _OPENMP
  nt = (m * n * k > (Index)1 << 18) ? omp_get_max_threads() : 1;
#endif
  nt = (int)std::max<Index>(1, std::min<Index>(nt, k / 64 + 1));
  std::vector<std::vector<double>> acc(nt);
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int t = 0; t < nt; ++t) {
    acc[t].assign((size_t)(m * n), 0.0);
    const Index s0 = k * t / nt, s1 = k * (t + 1) / nt, kc = 64;
    double *c = acc[t].data();
    for (Index p0 = s0; p0 < s1; p0 += kc)
      gemm_block<true>(m, 0, n, A, lda, B, ldb, 1, p0, std::min<Index>(s1, p0 + kc), c, m); // B(p, j) = B[j + p * ldb]
  }
#pragma omp parallel for schedule(static) if (m * n > 1 << 14)
  for (Index j = 0; j < n; ++j)
    for (Index i = 0; i < m; ++i) {
      double s = 0.0;
      for (int t = 0; t < nt; ++t) s += acc[t][(size_t)(i + j * m)];
      C[i + j * ldc] = s;
    }
}

} // namespace shim

class VectorXd;
class MatrixXd;
template <typename T> class Map;

// ---- element-wise "array" view of a vector (grad.array() += lambda * w.array()) -------------------------
struct ArrayValue { // an evaluated array expression
  std::vector<double> v;
};
struct ConstArrayRef {
  const double *p;
  Index n;
};
inline ArrayValue operator*(double s, const ConstArrayRef &a) {
  ArrayValue r;
  r.v.resize((size_t)a.n);
  for (Index i = 0; i < a.n; ++i) r.v[(size_t)i] = s * a.p[i];
  return r;
}
struct ArrayRef {
  double *p;
  Index n;
  ArrayRef &operator+=(const ArrayValue &o) {
    for (Index i = 0; i < n; ++i) p[i] += o.v[(size_t)i];
    return *this;
  }
};

// ---- vectors ------------------------------------------------------------------------------------------
struct ConstVecRef { // anything that reads like a vector: a VectorXd, a Map, a matrix column
  const double *p;
  Index n;
};

class VectorXd {
public:
  VectorXd() = default;
  explicit VectorXd(Index n) : d_((size_t)n) {}
  VectorXd(const ConstVecRef &r) : d_(r.p, r.p + r.n) {}
  static VectorXd Zero(Index n) {
    VectorXd v(n);
    std::fill(v.d_.begin(), v.d_.end(), 0.0);
    return v;
  }
  Index size() const { return (Index)d_.size(); }
  Index rows() const { return size(); }
  Index cols() const { return 1; }
  double *data() { return d_.data(); }
  const double *data() const { return d_.data(); }
  void resize(Index n) { d_.resize((size_t)n); }
  double &operator[](Index i) { return d_[(size_t)i]; }
  double operator[](Index i) const { return d_[(size_t)i]; }
  double &operator()(Index i) { return d_[(size_t)i]; }
  double operator()(Index i) const { return d_[(size_t)i]; }
  operator ConstVecRef() const { return ConstVecRef{data(), size()}; }

  double dot(const VectorXd &o) const { return shim::dot(data(), o.data(), size()); }
  double squaredNorm() const { return shim::dot(data(), data(), size()); }
  double norm() const { return std::sqrt(squaredNorm()); }
  double sum() const {
    double s = 0.0;
    for (double x : d_) s += x;
    return s;
  }
  ArrayRef array() { return ArrayRef{data(), size()}; }
  ConstArrayRef array() const { return ConstArrayRef{data(), size()}; }
  VectorXd &noalias() { return *this; }

  VectorXd operator-() const {
    VectorXd r(size());
    for (Index i = 0; i < size(); ++i) r[i] = -d_[(size_t)i];
    return r;
  }
  VectorXd &operator+=(const VectorXd &o) {
    for (Index i = 0; i < size(); ++i) d_[(size_t)i] += o[i];
    return *this;
  }
  VectorXd &operator-=(const VectorXd &o) {
    for (Index i = 0; i < size(); ++i) d_[(size_t)i] -= o[i];
    return *this;
  }
  VectorXd &operator*=(double s) {
    for (double &x : d_) x *= s;
    return *this;
  }
  VectorXd &operator/=(double s) {
    for (double &x : d_) x /= s;
    return *this;
  }

private:
  std::vector<double> d_;
};
inline VectorXd operator+(const VectorXd &a, const VectorXd &b) {
  VectorXd r(a.size());
  for (Index i = 0; i < a.size(); ++i) r[i] = a[i] + b[i];
  return r;
}
inline VectorXd operator-(const VectorXd &a, const VectorXd &b) {
  VectorXd r(a.size());
  for (Index i = 0; i < a.size(); ++i) r[i] = a[i] - b[i];
  return r;
}
inline VectorXd operator*(double s, const VectorXd &a) {
  VectorXd r(a.size());
  for (Index i = 0; i < a.size(); ++i) r[i] = s * a[i];
  return r;
}
inline VectorXd operator*(const VectorXd &a, double s) { return s * a; }
inline VectorXd operator/(const VectorXd &a, double s) {
  VectorXd r(a.size());
  for (Index i = 0; i < a.size(); ++i) r[i] = a[i] / s;
  return r;
}

// ---- matrices ------------------------------------------------------------------------------------------
struct ConstMatRef { // anything that reads like a matrix: a MatrixXd, a Map
  const double *p;
  Index r, c, ld;
};
struct TransposedRef { // X.transpose(), only ever an operand of a product
  ConstMatRef m;
};

class ColXpr { // M.col(j), writable
public:
  ColXpr(double *p, Index n) : p_(p), n_(n) {}
  ColXpr &operator=(const ColXpr &o) {
    std::memcpy(p_, o.p_, sizeof(double) * (size_t)n_);
    return *this;
  }
  ColXpr &operator=(const VectorXd &v) {
    std::memcpy(p_, v.data(), sizeof(double) * (size_t)n_);
    return *this;
  }
  operator ConstVecRef() const { return ConstVecRef{p_, n_}; }
  operator VectorXd() const { return VectorXd(ConstVecRef{p_, n_}); }
  Index size() const { return n_; }
  const double *data() const { return p_; }
  double maxCoeff(Index *idx) const { // first maximum wins, like Eigen's visitor
    Index b = 0;
    for (Index i = 1; i < n_; ++i)
      if (p_[i] > p_[b]) b = i;
    if (idx) *idx = b;
    return p_[b];
  }

private:
  double *p_;
  Index n_;
};
inline VectorXd operator-(const ColXpr &a, const VectorXd &b) {
  VectorXd r(a.size());
  for (Index i = 0; i < a.size(); ++i) r[i] = a.data()[i] - b[i];
  return r;
}

class MatrixXd;
struct ColwiseOps {
  MatrixXd *m;
  ColwiseOps &operator+=(const ConstVecRef &v);
  VectorXd squaredNorm() const;
};
struct RowwiseOps {
  const MatrixXd *m;
  VectorXd sum() const;
};

class MatrixXd {
public:
  MatrixXd() = default;
  MatrixXd(Index r, Index c) : r_(r), c_(c), d_((size_t)(r * c)) {}
  MatrixXd(const ConstMatRef &m) : r_(m.r), c_(m.c), d_((size_t)(m.r * m.c)) {
    for (Index j = 0; j < c_; ++j) std::memcpy(d_.data() + j * r_, m.p + j * m.ld, sizeof(double) * (size_t)r_);
  }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index size() const { return r_ * c_; }
  double *data() { return d_.data(); }
  const double *data() const { return d_.data(); }
  void resize(Index r, Index c) {
    r_ = r;
    c_ = c;
    d_.resize((size_t)(r * c));
  }
  double &operator()(Index i, Index j) { return d_[(size_t)(i + j * r_)]; }
  double operator()(Index i, Index j) const { return d_[(size_t)(i + j * r_)]; }
  operator ConstMatRef() const { return ConstMatRef{data(), r_, c_, r_}; }
  ColXpr col(Index j) { return ColXpr(d_.data() + j * r_, r_); }
  const ColXpr col(Index j) const { return ColXpr(const_cast<double *>(d_.data()) + j * r_, r_); }
  // leftCols(n): Eigen returns a view; callers here only read it (forward(x_view), output - y_view): a copy behaves the same
  MatrixXd leftCols(Index n) const { return MatrixXd(ConstMatRef{data(), r_, n, r_}); }
  TransposedRef transpose() const { return TransposedRef{ConstMatRef{data(), r_, c_, r_}}; }
  ColwiseOps colwise() { return ColwiseOps{this}; }
  ColwiseOps colwise() const { return ColwiseOps{const_cast<MatrixXd *>(this)}; }
  RowwiseOps rowwise() const { return RowwiseOps{this}; }
  MatrixXd &noalias() { return *this; }
  template <typename F> MatrixXd unaryExpr(F f) const {
    MatrixXd r(r_, c_);
    const Index n = size();
#pragma omp parallel for schedule(static) if (n > 1 << 16)
    for (Index i = 0; i < n; ++i) r.d_[(size_t)i] = f(d_[(size_t)i]);
    return r;
  }
  MatrixXd cwiseProduct(const MatrixXd &o) const {
    MatrixXd r(r_, c_);
    const Index n = size();
#pragma omp parallel for schedule(static) if (n > 1 << 16)
    for (Index i = 0; i < n; ++i) r.d_[(size_t)i] = d_[(size_t)i] * o.d_[(size_t)i];
    return r;
  }
  double squaredNorm() const { return shim::dot(data(), data(), size()); }
  MatrixXd &operator+=(const MatrixXd &o) {
    for (Index i = 0; i < size(); ++i) d_[(size_t)i] += o.d_[(size_t)i];
    return *this;
  }

private:
  Index r_ = 0, c_ = 0;
  std::vector<double> d_;
};
inline MatrixXd operator-(const MatrixXd &a, const MatrixXd &b) {
  MatrixXd r(a.rows(), a.cols());
  const Index n = a.size();
  const double *pa = a.data(), *pb = b.data();
  double *pr = r.data();
#pragma omp parallel for schedule(static) if (n > 1 << 16)
  for (Index i = 0; i < n; ++i) pr[i] = pa[i] - pb[i];
  return r;
}
inline ColwiseOps &ColwiseOps::operator+=(const ConstVecRef &v) {
  const Index r = m->rows(), c = m->cols();
  double *p = m->data();
#pragma omp parallel for schedule(static) if (r * c > 1 << 16)
  for (Index j = 0; j < c; ++j)
    for (Index i = 0; i < r; ++i) p[i + j * r] += v.p[i];
  return *this;
}
inline VectorXd ColwiseOps::squaredNorm() const {
  VectorXd out(m->cols());
  for (Index j = 0; j < m->cols(); ++j) out[j] = shim::dot(m->data() + j * m->rows(), m->data() + j * m->rows(), m->rows());
  return out;
}
inline VectorXd RowwiseOps::sum() const {
  const Index r = m->rows(), c = m->cols();
  VectorXd out = VectorXd::Zero(r);
  for (Index j = 0; j < c; ++j)
    for (Index i = 0; i < r; ++i) out[i] += (*m)(i, j);
  return out;
}

// ---- products ------------------------------------------------------------------------------------------
inline MatrixXd operator*(const ConstMatRef &a, const ConstMatRef &b) { // A * B
  MatrixXd c(a.r, b.c);
  shim::gemm_nn(a.r, b.c, a.c, a.p, a.ld, b.p, b.ld, c.data(), a.r);
  return c;
}
inline MatrixXd operator*(const ConstMatRef &a, const TransposedRef &bt) { // A * B^T
  MatrixXd c(a.r, bt.m.r);
  shim::gemm_nt(a.r, bt.m.r, a.c, a.p, a.ld, bt.m.p, bt.m.ld, c.data(), a.r);
  return c;
}
inline MatrixXd operator*(const TransposedRef &at, const ConstMatRef &b) { // A^T * B
  MatrixXd c(at.m.c, b.c);
  shim::gemm_tn(at.m.c, b.c, at.m.r, at.m.p, at.m.ld, b.p, b.ld, c.data(), at.m.c);
  return c;
}
inline MatrixXd operator*(const MatrixXd &a, const MatrixXd &b) { return (ConstMatRef)a * (ConstMatRef)b; }
inline MatrixXd operator*(const MatrixXd &a, const TransposedRef &bt) { return (ConstMatRef)a * bt; }
inline MatrixXd operator*(const TransposedRef &at, const MatrixXd &b) { return at * (ConstMatRef)b; }

// ---- Map: a view of caller-owned memory --------------------------------------------------------------------
template <> class Map<const MatrixXd> {
public:
  Map(const double *p, Index r, Index c) : p_(p), r_(r), c_(c) {}
  operator ConstMatRef() const { return ConstMatRef{p_, r_, c_, r_}; }
  TransposedRef transpose() const { return TransposedRef{ConstMatRef{p_, r_, c_, r_}}; }
  Index rows() const { return r_; }
  Index cols() const { return c_; }

private:
  const double *p_;
  Index r_, c_;
};
inline MatrixXd operator*(const Map<const MatrixXd> &a, const MatrixXd &b) { return (ConstMatRef)a * (ConstMatRef)b; }

template <> class Map<MatrixXd> {
public:
  Map(double *p, Index r, Index c) : p_(p), r_(r), c_(c) {}
  Map &noalias() { return *this; }
  Map &operator+=(const MatrixXd &o) {
    const Index n = r_ * c_;
    for (Index i = 0; i < n; ++i) p_[i] += o.data()[i];
    return *this;
  }
  operator ConstMatRef() const { return ConstMatRef{p_, r_, c_, r_}; }

private:
  double *p_;
  Index r_, c_;
};
template <> class Map<const VectorXd> {
public:
  Map(const double *p, Index n) : p_(p), n_(n) {}
  operator ConstVecRef() const { return ConstVecRef{p_, n_}; }

private:
  const double *p_;
  Index n_;
};
template <> class Map<VectorXd> {
public:
  Map(double *p, Index n) : p_(p), n_(n) {}
  Map &noalias() { return *this; }
  Map &operator+=(const VectorXd &o) {
    for (Index i = 0; i < n_; ++i) p_[i] += o[i];
    return *this;
  }

private:
  double *p_;
  Index n_;
};

} // namespace Eigen

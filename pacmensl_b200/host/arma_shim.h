// arma_shim.h -- a small, from-scratch stand-in for the part of Armadillo's container surface that appears
// in the reference's public API and tests (arma::Mat<int>/Row/Col, column-major storage, memptr/colptr,
// n_rows/n_cols/n_elem).  Armadillo itself is not available in this image and none
// of its code is used.  Only what the FSP hot path's signatures and tests need is provided.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <vector>

namespace arma {

typedef unsigned long long uword;

namespace fill {
struct fill_zeros {};
struct fill_ones {};
static const fill_zeros zeros{};
static const fill_ones  ones{};
}  // namespace fill

template <typename T> class Col;
template <typename T> class Row;

template <typename T>
class Mat {
 public:
  uword n_rows = 0, n_cols = 0, n_elem = 0;

  Mat() {}
  Mat(uword r, uword c) { set_size(r, c); }
  Mat(uword r, uword c, fill::fill_zeros) { set_size(r, c); zeros(); }
  Mat(uword r, uword c, fill::fill_ones) { set_size(r, c); fill(T(1)); }
  // copies the memory (column major), like arma::Mat(ptr, r, c) with copy_aux_mem = true
  Mat(const T *ptr, uword r, uword c) { set_size(r, c); if (n_elem) std::memcpy(mem_.data(), ptr, sizeof(T) * n_elem); }
  Mat(T *ptr, uword r, uword c, bool /*copy_aux_mem*/, bool /*strict*/ = false) : Mat(const_cast<const T *>(ptr), r, c) {}
  // a flat list makes a ROW vector (Armadillo semantics: Mat<int> SM{1, -1} is 1 x 2)
  Mat(std::initializer_list<T> l) {
    set_size(1, l.size());
    std::copy(l.begin(), l.end(), mem_.begin());
  }
  // nested lists are rows
  Mat(std::initializer_list<std::initializer_list<T>> rows) {
    uword r = rows.size(), c = 0;
    for (auto &row : rows) c = std::max<uword>(c, row.size());
    set_size(r, c);
    zeros();
    uword i = 0;
    for (auto &row : rows) {
      uword j = 0;
      for (auto &v : row) (*this)(i, j++) = v;
      ++i;
    }
  }
  virtual ~Mat() {}

  void set_size(uword r, uword c) { n_rows = r; n_cols = c; n_elem = r * c; mem_.resize(n_elem); }
  virtual void set_size(uword n) { set_size(n, 1); }
  void resize(uword r, uword c) {  // keeps the leading block
    Mat<T> old(*this);
    set_size(r, c);
    zeros();
    for (uword j = 0; j < std::min(c, old.n_cols); ++j)
      for (uword i = 0; i < std::min(r, old.n_rows); ++i) (*this)(i, j) = old(i, j);
  }
  void reset() { set_size(0, 0); }
  void clear() { reset(); }
  bool is_empty() const { return n_elem == 0; }
  Mat &zeros() { std::fill(mem_.begin(), mem_.end(), T(0)); return *this; }
  Mat &zeros(uword r, uword c) { set_size(r, c); return zeros(); }
  Mat &ones() { return fill(T(1)); }
  Mat &fill(T v) { std::fill(mem_.begin(), mem_.end(), v); return *this; }

  T *memptr() { return mem_.data(); }
  const T *memptr() const { return mem_.data(); }
  T *colptr(uword j) { return mem_.data() + j * n_rows; }
  const T *colptr(uword j) const { return mem_.data() + j * n_rows; }
  T &operator()(uword i, uword j) { return mem_[j * n_rows + i]; }
  const T &operator()(uword i, uword j) const { return mem_[j * n_rows + i]; }
  T &at(uword i, uword j) { return mem_[j * n_rows + i]; }
  const T &at(uword i, uword j) const { return mem_[j * n_rows + i]; }
  T &operator()(uword i) { return mem_[i]; }
  const T &operator()(uword i) const { return mem_[i]; }
  T &operator[](uword i) { return mem_[i]; }
  const T &operator[](uword i) const { return mem_[i]; }
  T &at(uword i) { return mem_[i]; }
  const T &at(uword i) const { return mem_[i]; }
  T *begin() { return mem_.data(); }
  T *end() { return mem_.data() + n_elem; }
  const T *begin() const { return mem_.data(); }
  const T *end() const { return mem_.data() + n_elem; }
  uword size() const { return n_elem; }

  // column view: supports X.col(j).fill(v), X.col(j) = other, and reading as a Col<T>
  class col_view {
   public:
    col_view(Mat &m, uword j) : m_(m), j_(j) {}
    col_view &fill(T v) { for (uword i = 0; i < m_.n_rows; ++i) m_(i, j_) = v; return *this; }
    col_view &operator=(const Mat<T> &o) {
      if (o.n_elem != m_.n_rows) throw std::logic_error("arma shim: col assignment size mismatch");
      for (uword i = 0; i < m_.n_rows; ++i) m_(i, j_) = o[i];
      return *this;
    }
    T &operator()(uword i) { return m_(i, j_); }
    operator Col<T>() const;
   private:
    Mat  &m_;
    uword j_;
  };
  col_view col(uword j) { return col_view(*this, j); }
  Col<T> col(uword j) const;
  Mat<T> cols(uword first, uword last) const {
    Mat<T> out(n_rows, last >= first ? last - first + 1 : 0);
    for (uword j = first; j <= last && j < n_cols; ++j)
      for (uword i = 0; i < n_rows; ++i) out(i, j - first) = (*this)(i, j);
    return out;
  }
  Mat<T> t() const {
    Mat<T> out(n_cols, n_rows);
    for (uword j = 0; j < n_cols; ++j)
      for (uword i = 0; i < n_rows; ++i) out(j, i) = (*this)(i, j);
    return out;
  }
  Mat &operator+=(const Mat &o) { for (uword i = 0; i < n_elem; ++i) mem_[i] += o.mem_[i]; return *this; }
  Mat &operator-=(const Mat &o) { for (uword i = 0; i < n_elem; ++i) mem_[i] -= o.mem_[i]; return *this; }
  Mat &operator*=(T s) { for (auto &v : mem_) v *= s; return *this; }

 protected:
  std::vector<T> mem_;
};

template <typename T>
class Col : public Mat<T> {
 public:
  Col() { this->Mat<T>::set_size(0, 1); }
  explicit Col(uword n) { this->Mat<T>::set_size(n, 1); }
  Col(uword n, fill::fill_zeros) { this->Mat<T>::set_size(n, 1); this->zeros(); }
  Col(uword n, fill::fill_ones) { this->Mat<T>::set_size(n, 1); this->fill(T(1)); }
  Col(const T *ptr, uword n) : Mat<T>(ptr, n, 1) {}
  Col(T *ptr, uword n, bool, bool = false) : Mat<T>(const_cast<const T *>(ptr), n, 1) {}
  Col(std::initializer_list<T> l) { this->Mat<T>::set_size(l.size(), 1); std::copy(l.begin(), l.end(), this->mem_.begin()); }
  Col(const Mat<T> &m) { this->Mat<T>::set_size(m.n_elem, 1); std::copy(m.begin(), m.end(), this->mem_.begin()); }
  Col(const std::vector<T> &v) { this->Mat<T>::set_size(v.size(), 1); std::copy(v.begin(), v.end(), this->mem_.begin()); }
  void set_size(uword n) override { Mat<T>::set_size(n, 1); }
  void resize(uword n) { Mat<T>::resize(n, 1); }
};

template <typename T>
class Row : public Mat<T> {
 public:
  Row() { this->Mat<T>::set_size(1, 0); }
  explicit Row(uword n) { this->Mat<T>::set_size(1, n); }
  Row(uword n, fill::fill_zeros) { this->Mat<T>::set_size(1, n); this->zeros(); }
  Row(uword n, fill::fill_ones) { this->Mat<T>::set_size(1, n); this->fill(T(1)); }
  Row(const T *ptr, uword n) : Mat<T>(ptr, 1, n) {}
  Row(T *ptr, uword n, bool, bool = false) : Mat<T>(const_cast<const T *>(ptr), 1, n) {}
  Row(std::initializer_list<T> l) { this->Mat<T>::set_size(1, l.size()); std::copy(l.begin(), l.end(), this->mem_.begin()); }
  Row(const Mat<T> &m) { this->Mat<T>::set_size(1, m.n_elem); std::copy(m.begin(), m.end(), this->mem_.begin()); }
  Row(const std::vector<T> &v) { this->Mat<T>::set_size(1, v.size()); std::copy(v.begin(), v.end(), this->mem_.begin()); }
  void set_size(uword n) override { Mat<T>::set_size(1, n); }
  void resize(uword n) { Mat<T>::resize(1, n); }
};

template <typename T>
Mat<T>::col_view::operator Col<T>() const {
  Col<T> c(m_.n_rows);
  for (uword i = 0; i < m_.n_rows; ++i) c[i] = m_(i, j_);
  return c;
}
template <typename T>
Col<T> Mat<T>::col(uword j) const {
  Col<T> c(n_rows);
  for (uword i = 0; i < n_rows; ++i) c[i] = (*this)(i, j);
  return c;
}

typedef Mat<double> mat;
typedef Col<double> vec;
typedef Col<double> dvec;
typedef Row<double> rowvec;
typedef Col<uword>  uvec;

template <typename T>
Mat<T> join_horiz(const Mat<T> &a, const Mat<T> &b) {
  if (a.n_elem == 0) return b;
  if (b.n_elem == 0) return a;
  Mat<T> out(a.n_rows, a.n_cols + b.n_cols);
  std::copy(a.begin(), a.end(), out.memptr());
  std::copy(b.begin(), b.end(), out.memptr() + a.n_elem);
  return out;
}

template <typename VecT>
VecT linspace(double a, double b, uword n) {
  VecT out;
  out.set_size(n);
  for (uword i = 0; i < n; ++i) out[i] = n > 1 ? a + (b - a) * double(i) / double(n - 1) : b;
  return out;
}

template <typename Out>
struct conv_to {
  template <typename T>
  static Out from(const Mat<T> &m) {
    Out o(m.n_elem);
    for (uword i = 0; i < m.n_elem; ++i) o[i] = static_cast<typename Out::value_type>(m[i]);
    return o;
  }
};

template <typename T> T accu(const Mat<T> &m) { T s = T(0); for (auto v : m) s += v; return s; }
template <typename T> T max(const Mat<T> &m) { return *std::max_element(m.begin(), m.end()); }
template <typename T> T min(const Mat<T> &m) { return *std::min_element(m.begin(), m.end()); }

inline Mat<double> zeros(uword r, uword c) { return Mat<double>(r, c, fill::zeros); }

// matrix infinity norm (max absolute row sum) -- the only norm the hot path uses (KrylovFsp.cpp:460)
inline double norm(const Mat<double> &A, const char *kind) {
  if (std::string(kind) != "inf") throw std::logic_error("arma shim: only norm(A, \"inf\") is provided");
  double best = 0.0;
  for (uword i = 0; i < A.n_rows; ++i) {
    double s = 0.0;
    for (uword j = 0; j < A.n_cols; ++j) s += std::fabs(A(i, j));
    best = std::max(best, s);
  }
  return best;
}

inline Mat<double> operator*(double s, const Mat<double> &A) { Mat<double> B(A); B *= s; return B; }
inline Mat<double> operator*(const Mat<double> &A, double s) { Mat<double> B(A); B *= s; return B; }

// Dense matrix exponential (scaling-and-squaring with a degree-13 Pade approximant, Higham 2005): the
// stand-in for arma::expmat used on the small Krylov Hessenberg matrix (KrylovFsp.cpp:159,376).
Mat<double> expmat(const Mat<double> &A);

}  // namespace arma

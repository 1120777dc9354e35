// arma_shim.cpp -- dense expm for the small Krylov matrix (see arma_shim.h).
#include "arma_shim.h"

namespace arma {

namespace {
typedef Mat<double> M;

M matmul(const M &A, const M &B) {
  const uword n = A.n_rows, k = A.n_cols, m = B.n_cols;
  M C(n, m, fill::zeros);
  for (uword j = 0; j < m; ++j)
    for (uword p = 0; p < k; ++p) {
      const double b = B(p, j);
      if (b == 0.0) continue;
      const double *a = A.colptr(p);
      double       *c = C.colptr(j);
      for (uword i = 0; i < n; ++i) c[i] += a[i] * b;
    }
  return C;
}

// C = a*A + b*B + c*Cm (+ d*I)
M lincomb(double a, const M &A, double b, const M &B, double c, const M &Cm, double d) {
  M out(A.n_rows, A.n_cols);
  for (uword i = 0; i < A.n_elem; ++i) out[i] = a * A[i] + b * B[i] + c * Cm[i];
  for (uword i = 0; i < A.n_rows; ++i) out(i, i) += d;
  return out;
}

// solve P X = Q in place (Q overwritten with X) by LU with partial pivoting
void solve(M P, M &Q) {
  const uword n = P.n_rows, m = Q.n_cols;
  for (uword k = 0; k < n; ++k) {
    uword  piv = k;
    double best = std::fabs(P(k, k));
    for (uword i = k + 1; i < n; ++i)
      if (std::fabs(P(i, k)) > best) { best = std::fabs(P(i, k)); piv = i; }
    if (best == 0.0) throw std::runtime_error("expmat: singular Pade denominator");
    if (piv != k) {
      for (uword j = 0; j < n; ++j) std::swap(P(k, j), P(piv, j));
      for (uword j = 0; j < m; ++j) std::swap(Q(k, j), Q(piv, j));
    }
    const double inv = 1.0 / P(k, k);
    for (uword i = k + 1; i < n; ++i) {
      const double f = P(i, k) * inv;
      if (f == 0.0) continue;
      P(i, k) = 0.0;
      for (uword j = k + 1; j < n; ++j) P(i, j) -= f * P(k, j);
      for (uword j = 0; j < m; ++j) Q(i, j) -= f * Q(k, j);
    }
  }
  for (uword j = 0; j < m; ++j)
    for (uword ii = n; ii-- > 0;) {
      double s = Q(ii, j);
      for (uword c = ii + 1; c < n; ++c) s -= P(ii, c) * Q(c, j);
      Q(ii, j) = s / P(ii, ii);
    }
}
}  // namespace

Mat<double> expmat(const Mat<double> &Ain) {
  if (Ain.n_rows != Ain.n_cols) throw std::logic_error("expmat: matrix must be square");
  const uword n = Ain.n_rows;
  if (n == 0) return Ain;
  static const double b[14] = {64764752532480000.0, 32382376266240000.0, 7771770303897600.0, 1187353796428800.0,
                               129060195264000.0,   10559470521600.0,    670442572800.0,     33522128640.0,
                               1323241920.0,        40840800.0,          960960.0,           16380.0,
                               182.0,               1.0};
  const double theta13 = 5.371920351148152;
  double       norm1 = 0.0;
  for (uword j = 0; j < n; ++j) {
    double s = 0.0;
    for (uword i = 0; i < n; ++i) s += std::fabs(Ain(i, j));
    norm1 = std::max(norm1, s);
  }
  int s = 0;
  if (norm1 > theta13) s = std::max(0, (int) std::ceil(std::log2(norm1 / theta13)));
  M A(Ain);
  if (s > 0) A *= std::ldexp(1.0, -s);
  M A2 = matmul(A, A), A4 = matmul(A2, A2), A6 = matmul(A4, A2);
  M W1 = lincomb(b[13], A6, b[11], A4, b[9], A2, 0.0);
  M W2 = lincomb(b[7], A6, b[5], A4, b[3], A2, b[1]);
  M Z1 = lincomb(b[12], A6, b[10], A4, b[8], A2, 0.0);
  M Z2 = lincomb(b[6], A6, b[4], A4, b[2], A2, b[0]);
  M W = matmul(A6, W1);
  W += W2;
  M U = matmul(A, W);
  M V = matmul(A6, Z1);
  V += Z2;
  M P(V), Q(V);
  P -= U;
  Q += U;
  solve(P, Q);
  for (int k = 0; k < s; ++k) Q = matmul(Q, Q);
  return Q;
}

}  // namespace arma

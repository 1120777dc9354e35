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
    double      *lk = P.colptr(k);  // multipliers stored in place of the eliminated column
    for (uword i = k + 1; i < n; ++i) lk[i] *= inv;
    for (uword j = k + 1; j < n; ++j) {  // column-major friendly rank-1 update
      double      *pj = P.colptr(j);
      const double pkj = pj[k];
      if (pkj == 0.0) continue;
      for (uword i = k + 1; i < n; ++i) pj[i] -= lk[i] * pkj;
    }
    for (uword j = 0; j < m; ++j) {
      double      *qj = Q.colptr(j);
      const double qkj = qj[k];
      if (qkj == 0.0) continue;
      for (uword i = k + 1; i < n; ++i) qj[i] -= lk[i] * qkj;
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

static Mat<double> expmat_dense(const Mat<double> &Ain);

// exp of a square matrix.  The Krylov matrix handed in by KrylovFsp is (m_max+2)^2 = 62x62 with only its leading
// (m+2)x(m+2) block non-zero; exp([[B,0],[0,0]]) = [[exp(B),0],[0,I]] and every step of the Pade/scaling-squaring
// arithmetic preserves that block structure exactly (the trailing block only ever holds multiples of I), so the
// work is done on the active block: same result, (m+2)^3 instead of 62^3 flops on the host between two basis
// generations.
Mat<double> expmat(const Mat<double> &Ain) {
  if (Ain.n_rows != Ain.n_cols) throw std::logic_error("expmat: matrix must be square");
  const uword n = Ain.n_rows;
  if (n == 0) return Ain;
  uword na = 0;  // active size: 1 + largest row/column index holding a non-zero
  for (uword j = 0; j < n; ++j)
    for (uword i = 0; i < n; ++i)
      if (Ain(i, j) != 0.0) na = std::max(na, std::max(i, j) + 1);
  if (na == n) return expmat_dense(Ain);
  Mat<double> out(n, n, fill::zeros);
  for (uword i = 0; i < n; ++i) out(i, i) = 1.0;
  if (na == 0) return out;
  Mat<double> B(na, na);
  for (uword j = 0; j < na; ++j)
    for (uword i = 0; i < na; ++i) B(i, j) = Ain(i, j);
  Mat<double> E = expmat_dense(B);
  for (uword j = 0; j < na; ++j)
    for (uword i = 0; i < na; ++i) out(i, j) = E(i, j);
  return out;
}

static Mat<double> expmat_dense(const Mat<double> &Ain) {
  const uword n = Ain.n_rows;
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

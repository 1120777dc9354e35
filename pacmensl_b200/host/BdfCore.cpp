// BdfCore.cpp -- see BdfCore.h.  All vector work is done by the fused fp64 device kernels of
// include/fsp_b200.h; only scalars (norms, Hessenberg columns) come back to the host.
#include "BdfCore.h"
#include "PetscWrap.h"

#include <algorithm>
#include <cfloat>
#include <cmath>

namespace pacmensl {

namespace {
// constants of the CVODE step/order controller (SUNDIALS documentation, "CVODE constants")
constexpr double ETAMX1 = 10000.0, ETAMX2 = 10.0, ETAMX3 = 10.0, ETAMXF = 0.2, ETAMIN = 0.1, ETACF = 0.25;
constexpr double ADDON = 1.0e-6, BIAS1 = 6.0, BIAS2 = 6.0, BIAS3 = 10.0, THRESH = 1.5, ONEPSM = 1.000001;
constexpr double CORTES = 0.1, CRDOWN = 0.3, RDIV = 2.0, EPLIFAC = 0.05;
constexpr int    SMALL_NST = 10, MXNEF1 = 3, SMALL_NEF = 2, LONG_WAIT = 10;
constexpr double HUB_FACTOR = 0.1, HLB_FACTOR = 100.0, H_BIAS = 0.5;
constexpr int    MAX_ITERS_HIN = 4;
constexpr double FUZZ_FACTOR = 100.0;

enum { DO_ERROR_TEST = 2, PREDICT_AGAIN = 3, CONV_FAIL = 4, TRY_AGAIN = 5 };
enum { BDF_SUCCESS = 0, BDF_CONV_FAILURE = -4, BDF_ERR_FAILURE = -3, BDF_RHS_FAIL = -8, BDF_LSOLVE_FAIL = -7,
       BDF_BAD_T = -25, BDF_MEM_FAIL = -20, BDF_TOO_CLOSE = -27, BDF_ILL_EWT = -22 };

#define VCHK(call)                                                                   \
  do {                                                                               \
    int v_ierr_ = (call);                                                            \
    if (v_ierr_ != 0) {                                                              \
      printf("BdfCore: device vector operation failed: %s (%s:%d)\n", fsp_last_error(), __FILE__, __LINE__); \
      return BDF_MEM_FAIL;                                                           \
    }                                                                                \
  } while (0)
}  // namespace

BdfCore::BdfCore(MPI_Comm comm) : comm_(comm) { stream_ = comm ? comm->stream : nullptr; }
BdfCore::~BdfCore() { Free(); }

void BdfCore::Free() {
  for (auto &z : zn_) if (z) VecDestroy(&z);
  for (Vec *v : {&ewt_, &ewt_inv_, &y_, &acor_, &tempv_, &ftemp_, &xcor_, &vtemp_, &delta_}) if (*v) VecDestroy(v);
  for (auto &v : V_) if (v) VecDestroy(&v);
  V_.clear();
  for (auto &zs : znS_) for (auto &v : zs) if (v) VecDestroy(&v);
  znS_.clear();
  for (auto *vv : {&ewtS_, &ewtS_inv_, &acorS_, &yS_, &ftempS_}) { for (auto &v : *vv) if (v) VecDestroy(&v); vv->clear(); }
  ns_ = 0;
}

int BdfCore::alloc_like(Vec proto, Vec *out) { return VecDuplicate(proto, out); }

int BdfCore::wrms(Vec v, Vec w, double *out) {
  // N_VWrmsNorm: sqrt( sum (v_i w_i)^2 / N_global )
  double *tmp = hdev_.get();
  VCHK(fspvec_wsqsum(tmp, v->d_data, w->d_data, n_local_, stream_));
  double s = 0.0;
  VCHK(fsp_memcpy_d2h(&s, tmp, sizeof(double), stream_));
  if (pacmensl_allreduce_sum(comm_, &s, 1)) return BDF_MEM_FAIL;
  *out = std::sqrt(s / n_global_);
  return 0;
}

int BdfCore::ewt_set(Vec y, Vec ewt, Vec ewt_inv) {
  // cvEwtSetSS: ewt_i = 1 / (rtol |y_i| + atol); fails if any denominator is <= 0.  The denominators are kept too
  // (ewt_inv): every "./ ewt" of the scaled GMRES loop becomes a multiplication.
  double *tmp = hdev_.get();
  VCHK(fspvec_ewt_pair(ewt->d_data, ewt_inv->d_data, y->d_data, rtol_, atol_, n_local_, tmp, stream_));
  double mn = 0.0;
  VCHK(fsp_memcpy_d2h(&mn, tmp, sizeof(double), stream_));
  double neg = -mn;
  if (pacmensl_allreduce_max(comm_, &neg, 1)) return BDF_MEM_FAIL;
  return (-neg) > 0.0 ? 0 : BDF_ILL_EWT;
}

Vec BdfCore::inv_of(Vec ewt) const {
  if (ewt == ewt_) return ewt_inv_;
  for (size_t is = 0; is < ewtS_.size(); ++is)
    if (ewt == ewtS_[is]) return ewtS_inv_[is];
  return nullptr;
}

int BdfCore::rhs(double t, Vec y, Vec ydot) {
  nfe_ += 1;
  return f_(t, y, ydot);
}

int BdfCore::Init(double t0, Vec y0, RhsFn f, JtvFn jtv, double tout_hint) {
  Free();
  f_ = std::move(f);
  jtv_ = std::move(jtv);
  n_local_ = y0->n_local;
  PetscInt ng = 0;
  VecGetSize(y0, &ng);
  n_global_ = (double) ng;
  tout_hint_ = tout_hint;
  for (int j = 0; j <= QMAX; ++j) if (alloc_like(y0, &zn_[j])) return BDF_MEM_FAIL;
  for (Vec *v : {&ewt_, &ewt_inv_, &y_, &acor_, &tempv_, &ftemp_, &xcor_, &vtemp_, &delta_}) if (alloc_like(y0, v)) return BDF_MEM_FAIL;
  if (hdev_.resize((size_t) (maxl_ + 8) * 2)) return BDF_MEM_FAIL;
  VCHK(VecCopy(y0, zn_[0]));
  q_ = 1; L_ = 2; qprime_ = 1; qwait_ = L_; qu_ = 0; nscon_ = 0; indx_acor_ = QMAX;
  etamax_ = ETAMX1; eta_ = 1.0; h_ = hprime_ = hscale_ = hu_ = next_h_ = 0.0;
  tn_ = t0;
  for (auto &t : tau_) t = 0.0;
  for (auto &t : tq_) t = 0.0;
  for (auto &t : l_) t = 0.0;
  crate_ = 1.0; saved_tq5_ = 0.0; gammap_ = 0.0; gamrat_ = 1.0;
  nst_ = nfe_ = njtv_ = nli_ = nni_ = netf_ = ncfn_ = 0;
  first_ = true;
  return 0;
}

int BdfCore::InitSens(int ns, Vec *s0, SensRhsFn fs, bool errcon) {
  ns_ = ns;
  errcon_ = errcon;
  fs_ = std::move(fs);
  znS_.assign(ns, std::vector<Vec>(QMAX + 1, nullptr));
  ewtS_.assign(ns, nullptr); ewtS_inv_.assign(ns, nullptr); acorS_.assign(ns, nullptr); yS_.assign(ns, nullptr); ftempS_.assign(ns, nullptr);
  acnrmS_.assign(ns, 0.0);
  for (int is = 0; is < ns; ++is) {
    for (int j = 0; j <= QMAX; ++j) if (alloc_like(zn_[0], &znS_[is][j])) return BDF_MEM_FAIL;
    if (alloc_like(zn_[0], &ewtS_inv_[is]) || alloc_like(zn_[0], &ewtS_[is]) || alloc_like(zn_[0], &acorS_[is]) || alloc_like(zn_[0], &yS_[is]) ||
        alloc_like(zn_[0], &ftempS_[is])) return BDF_MEM_FAIL;
    VCHK(VecCopy(s0[is], znS_[is][0]));
  }
  return 0;
}

// ---- first step size (cvHin) --------------------------------------------------------------------------------
int BdfCore::initial_step(double tout) {
  const double tdiff = tout - tn_;
  if (tdiff == 0.0) return BDF_TOO_CLOSE;
  const double sign = tdiff > 0.0 ? 1.0 : -1.0;
  const double tdist = std::fabs(tdiff);
  const double tround = DBL_EPSILON * std::max(std::fabs(tn_), std::fabs(tout));
  if (tdist < 2.0 * tround) return BDF_TOO_CLOSE;
  const double hlb = HLB_FACTOR * tround;
  // upper bound: hub = min(0.1 tdist, 1 / max_i |y'_i| / (0.1 |y_i| + 1/ewt_i))
  double hub_inv = 0.0;
  {
    // hub_inv = max_i |y'_i| / (0.1 |y_i| + 1/ewt_i)   (cvUpperBoundH0)
    // 1/ewt_i = rtol |y_i| + atol  =>  0.1 |y_i| + 1/ewt_i = (0.1 + rtol) |y_i| + atol : one fused kernel + max reduction
    double *tmp = hdev_.get();
    VCHK(fspvec_ratio_absmax(tmp, zn_[1]->d_data, zn_[0]->d_data, HUB_FACTOR + rtol_, atol_, n_local_, stream_));
    VCHK(fsp_memcpy_d2h(&hub_inv, tmp, sizeof(double), stream_));
    if (pacmensl_allreduce_max(comm_, &hub_inv, 1)) return BDF_MEM_FAIL;
  }
  double hub = HUB_FACTOR * tdist;
  if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
  double hg = std::sqrt(hlb * hub);
  if (hub < hlb) {
    h_ = sign * hg;
    return 0;
  }
  double hnew = hg;
  int    count1 = 0;
  while (true) {
    // second-derivative estimate (cvYddNorm): ydd = (f(t + hg, y + hg y') - y') / hg
    bool hg_ok = false;
    double yddnrm = 0.0;
    for (int count2 = 0; count2 < MAX_ITERS_HIN; ++count2) {
      const double hgs = hg * sign;
      VCHK(fspvec_linear_sum(y_->d_data, hgs, zn_[1]->d_data, 1.0, zn_[0]->d_data, n_local_, stream_));
      int r = rhs(tn_ + hgs, y_, tempv_);
      if (r < 0) return BDF_RHS_FAIL;
      if (r > 0) { hg *= 0.2; continue; }
      VCHK(fspvec_linear_sum(tempv_->d_data, 1.0 / hgs, tempv_->d_data, -1.0 / hgs, zn_[1]->d_data, n_local_, stream_));
      if (wrms(tempv_, ewt_, &yddnrm)) return BDF_MEM_FAIL;
      hg_ok = true;
      break;
    }
    if (!hg_ok) return BDF_RHS_FAIL;
    hnew = (yddnrm * hub * hub > 2.0) ? std::sqrt(2.0 / yddnrm) : std::sqrt(hg * hub);
    count1++;
    if (count1 == MAX_ITERS_HIN) break;
    const double hrat = hnew / hg;
    if (hrat > 0.5 && hrat < 2.0) break;
    if (count1 > 1 && hrat > 2.0) { hnew = hg; break; }
    hg = hnew;
  }
  double h0 = H_BIAS * hnew;
  if (h0 < hlb) h0 = hlb;
  if (h0 > hub) h0 = hub;
  h_ = sign * h0;
  return 0;
}

// ---- order / step adjustments ----------------------------------------------------------------------------------
// History arrays zn_[0..q] (and the sensitivity histories) are transformed by ONE fused pass per call: optional
// rescale, then prediction / restore (fspvec_nordsieck), bit-identical to CVODE's sequence of separate vector calls.
void BdfCore::history_pass(const double *scale, int pascal, int q) {
  double *Z[LMAX + 1];
  for (int j = 0; j <= q; ++j) Z[j] = zn_[j]->d_data;
  vec_status_ |= fspvec_nordsieck(Z, q + 1, scale, pascal, n_local_, stream_);
  for (int is = 0; is < ns_; ++is) {
    for (int j = 0; j <= q; ++j) Z[j] = znS_[is][j]->d_data;
    vec_status_ |= fspvec_nordsieck(Z, q + 1, scale, pascal, n_local_, stream_);
  }
}

void BdfCore::flush_scale() {
  if (!scale_pending_) return;
  history_pass(pending_scale_, 0, pending_q_);
  scale_pending_ = false;
}

void BdfCore::rescale() {
  // zn_[j] *= eta^j, j = 1..q (cvRescale): recorded here, applied inside the next prediction pass
  flush_scale();
  double factor = eta_;
  pending_scale_[0] = 1.0;
  for (int j = 1; j <= q_; ++j) {
    pending_scale_[j] = factor;
    factor *= eta_;
  }
  scale_pending_ = true;
  pending_q_ = q_;
  h_ = hscale_ * eta_;
  next_h_ = h_;
  hscale_ = h_;
  nscon_ = 0;
}

void BdfCore::increase_bdf() {
  flush_scale();
  for (int i = 0; i <= QMAX; ++i) l_[i] = 0.0;
  l_[2] = 1.0;
  double alpha1 = 1.0, prod = 1.0, xiold = 1.0, alpha0 = -1.0, hsum = hscale_;
  if (q_ > 1) {
    for (int j = 1; j < q_; ++j) {
      hsum += tau_[j + 1];
      const double xi = hsum / hscale_;
      prod *= xi;
      alpha0 -= 1.0 / (j + 1);
      alpha1 += 1.0 / xi;
      for (int i = j + 2; i >= 2; --i) l_[i] = l_[i] * xiold + l_[i - 1];
      xiold = xi;
    }
  }
  const double A1 = (-alpha0 - alpha1) / prod;
  vec_status_ |= fspvec_linear_sum(zn_[L_]->d_data, A1, zn_[indx_acor_]->d_data, 0.0, zn_[indx_acor_]->d_data, n_local_, stream_);
  for (int j = 2; j <= q_; ++j) vec_status_ |= fspvec_axpy(zn_[j]->d_data, l_[j], zn_[L_]->d_data, n_local_, stream_);
  for (int is = 0; is < ns_; ++is) {
    vec_status_ |= fspvec_linear_sum(znS_[is][L_]->d_data, A1, znS_[is][indx_acor_]->d_data, 0.0,
                                     znS_[is][indx_acor_]->d_data, n_local_, stream_);
    for (int j = 2; j <= q_; ++j) vec_status_ |= fspvec_axpy(znS_[is][j]->d_data, l_[j], znS_[is][L_]->d_data, n_local_, stream_);
  }
}

void BdfCore::decrease_bdf() {
  flush_scale();
  for (int i = 0; i <= QMAX; ++i) l_[i] = 0.0;
  l_[2] = 1.0;
  double hsum = 0.0;
  for (int j = 1; j <= q_ - 2; ++j) {
    hsum += tau_[j];
    const double xi = hsum / hscale_;
    for (int i = j + 2; i >= 2; --i) l_[i] = l_[i] * xi + l_[i - 1];
  }
  for (int j = 2; j < q_; ++j) {
    vec_status_ |= fspvec_axpy(zn_[j]->d_data, -l_[j], zn_[q_]->d_data, n_local_, stream_);
    for (int is = 0; is < ns_; ++is)
      vec_status_ |= fspvec_axpy(znS_[is][j]->d_data, -l_[j], znS_[is][q_]->d_data, n_local_, stream_);
  }
}

void BdfCore::adjust_order(int deltaq) {
  if (q_ == 2 && deltaq != 1) return;
  if (deltaq == 1) increase_bdf();
  else if (deltaq == -1) decrease_bdf();
}

void BdfCore::adjust_params() {
  if (qprime_ != q_) {
    adjust_order(qprime_ - q_);
    q_ = qprime_;
    L_ = q_ + 1;
    qwait_ = L_;
  }
  rescale();
}

void BdfCore::predict() {
  // cvPredict: for k = 1..q, j = q..k: zn[j-1] += zn[j]  -- with the pending rescale, one pass over the history
  tn_ += h_;
  if (scale_pending_ && pending_q_ != q_) flush_scale();
  history_pass(scale_pending_ ? pending_scale_ : nullptr, +1, q_);
  scale_pending_ = false;
}

void BdfCore::restore(double saved_t) {
  // cvRestore: the inverse Pascal product, one pass
  tn_ = saved_t;
  flush_scale();
  history_pass(nullptr, -1, q_);
}

void BdfCore::set_tq(double hsum, double alpha0, double alpha0_hat, double xi_inv, double xistar_inv) {
  const double A1 = 1.0 - alpha0_hat + alpha0;
  const double A2 = 1.0 + q_ * A1;
  tq_[2] = std::fabs(A1 / (alpha0 * A2));
  tq_[5] = std::fabs(A2 * xistar_inv / (l_[q_] * xi_inv));
  if (qwait_ == 1) {
    if (q_ > 1) {
      const double C = xistar_inv / l_[q_];
      const double A3 = alpha0 + 1.0 / q_;
      const double A4 = alpha0_hat + xi_inv;
      const double Cpinv = (1.0 - A4 + A3) / A3;
      tq_[1] = std::fabs(C * Cpinv);
    } else {
      tq_[1] = 1.0;
    }
    hsum += tau_[q_];
    xi_inv = h_ / hsum;
    const double A5 = alpha0 - (1.0 / (q_ + 1));
    const double A6 = alpha0_hat - xi_inv;
    const double Cppinv = (1.0 - A6 + A5) / A2;
    tq_[3] = std::fabs(Cppinv / (xi_inv * (q_ + 2) * A5));
  }
  tq_[4] = CORTES / tq_[2];
}

void BdfCore::set_coeffs() {
  l_[0] = l_[1] = 1.0;
  double xi_inv = 1.0, xistar_inv = 1.0;
  for (int i = 2; i <= q_; ++i) l_[i] = 0.0;
  double alpha0 = -1.0, alpha0_hat = -1.0, hsum = h_;
  if (q_ > 1) {
    for (int j = 2; j < q_; ++j) {
      hsum += tau_[j - 1];
      xi_inv = h_ / hsum;
      alpha0 -= 1.0 / j;
      for (int i = j; i >= 1; --i) l_[i] += l_[i - 1] * xi_inv;
    }
    alpha0 -= 1.0 / q_;
    xistar_inv = -l_[1] - alpha0;
    hsum += tau_[q_ - 1];
    xi_inv = h_ / hsum;
    alpha0_hat = -l_[1] - xi_inv;
    for (int i = q_; i >= 1; --i) l_[i] += l_[i - 1] * xistar_inv;
  }
  set_tq(hsum, alpha0, alpha0_hat, xi_inv, xistar_inv);
  rl1_ = 1.0 / l_[1];
  gamma_ = h_ * rl1_;
  if (nst_ == 0) gammap_ = gamma_;
  gamrat_ = (nst_ > 0) ? gamma_ / gammap_ : 1.0;
}

// ---- linear solver: scaled GMRES on (I - gamma J) x = b, x0 = 0 -----------------------------------------------------
// sum over ranks of a device-resident partial sum
int BdfCore::global_sum(const double *dev_scalar, double *out) {
  double s = 0.0;
  VCHK(fsp_memcpy_d2h(&s, dev_scalar, sizeof(double), stream_));
  if (pacmensl_allreduce_sum(comm_, &s, 1)) return BDF_MEM_FAIL;
  *out = s;
  return 0;
}

// On entry V_[0] = s1 .* b (s1 = s2 = ewt) and ss_b = ||V_[0]||_2^2 (global), both produced by the fused residual pass of
// nls().  On success the correction is  x = *xsrc ./ ewt  when *divide, else x = *xsrc; *xsrc == nullptr means x = 0.
int BdfCore::lin_solve(Vec b, Vec ewt, double ss_b, double tn, bool first_newton, int *converged, Vec *xsrc, bool *divide) {
  *converged = 0;
  *xsrc = nullptr;
  *divide = false;
  const double deltar = EPLIFAC * tq_[4];
  const double bnorm = std::sqrt(ss_b / n_global_);  // WRMS norm of b with weights ewt
  if (bnorm <= deltar) {
    // right-hand side already below the tolerance: the correction is b itself on the first Newton iteration
    // and zero afterwards (CVODE's linear-solver interface returns b unchanged / zeroed in these two cases)
    if (first_newton) *xsrc = b;
    *converged = 1;
    return 0;
  }
  const double delta = deltar * std::sqrt(n_global_);
  const bool   multi = comm_ && comm_->size > 1;
  Vec          ewt_inv = inv_of(ewt);  // 1 ./ ewt: the un-scalings "./ s2" are multiplications
  const double beta = std::sqrt(ss_b);
  double       rho = beta;
  if (rho <= delta) { *converged = 1; return 0; }
  // V0 /= beta and vtemp = V0 ./ s2 in one pass
  VCHK(fspvec_scale_mul(V_[0]->d_data, 1.0 / beta, vtemp_->d_data, ewt_inv->d_data, n_local_, stream_));

  const int lmax = maxl_;
  std::vector<std::vector<double>> Hes((size_t) lmax + 1, std::vector<double>((size_t) lmax, 0.0));
  std::vector<double> givens((size_t) 2 * lmax, 0.0), yg((size_t) lmax + 1, 0.0);
  double rotation_product = 1.0;
  int    l_used = 0;
  bool   conv = false;
  for (int l = 0; l < lmax; ++l) {
    l_used = l + 1;
    if ((int) V_.size() < l + 2) { V_.push_back(nullptr); if (alloc_like(b, &V_[l + 1])) return BDF_MEM_FAIL; }
    // w = J vtemp ; V_{l+1} = s1 .* (vtemp - gamma w)   (vtemp = V_l ./ s2 was formed by the previous fused pass)
    njtv_ += 1;
    double *hd = hdev_.get();
    double *w = V_[l + 1]->d_data;
    int     r;
    if (fused_jtv_) {
      // ONE kernel: V_{l+1} = s1 .* (vtemp - gamma J vtemp), <V_{l+1}, V_0> and <V_{l+1}, V_{l+1}>
      fspmat_epilogue ep{};
      ep.alpha = -gamma_; ep.beta = 1.0; ep.scale_dev = ewt->d_data; ep.n_dots = 2;
      ep.dot_vec_dev[0] = V_[0]->d_data; ep.dot_vec_dev[1] = nullptr; ep.dot_out_dev = hd + lmax + 4;
      r = fused_jtv_(tn, vtemp_, V_[l + 1], ep);
      if (r != 0) return r < 0 ? BDF_LSOLVE_FAIL : r;
    } else {
      r = jtv_(tn, vtemp_, V_[l + 1]);
      if (r != 0) return r < 0 ? BDF_LSOLVE_FAIL : r;
      VCHK(fspvec_wlincomb(V_[l + 1]->d_data, ewt->d_data, 1.0, vtemp_->d_data, -gamma_, V_[l + 1]->d_data, n_local_, stream_));
    }
    nli_ += 1;
    // modified Gram-Schmidt against V_0..V_l with device-resident coefficients:
    //   hd[0] = <w, V_0>, hd[l+2] = <w, w> (norm before orthogonalisation); then fused axpy+dot chain
    {
      if (!fused_jtv_) {
        const double *two[2] = {V_[0]->d_data, w};
        VCHK(fspvec_mdot(hd + lmax + 4, w, 2, two, n_local_, stream_));  // [<w,V0>, <w,w>]
      }
      VCHK(fsp_memcpy_d2d(hd + 0, hd + lmax + 4, sizeof(double), stream_));
      if (multi) {
        VCHK(fspcomm_allreduce_sum(comm_->nccl, hd + 0, 1, stream_));
        VCHK(fspcomm_allreduce_sum(comm_->nccl, hd + lmax + 5, 1, stream_));
      }
    }
    for (int i = 0; i <= l; ++i) {
      const double *u = (i < l) ? V_[i + 1]->d_data : nullptr;
      VCHK(fspvec_axpy_dot(w, hd + i, 1.0, V_[i]->d_data, u, hd + i + 1, n_local_, stream_));
      if (multi) VCHK(fspcomm_allreduce_sum(comm_->nccl, hd + i + 1, 1, stream_));
    }
    hhost_.resize((size_t) lmax + 8);
    VCHK(fsp_memcpy_d2h(hhost_.data(), hd, sizeof(double) * (lmax + 6), stream_));
    if (multi) VCHK(fspcomm_check(comm_->nccl));
    for (int i = 0; i <= l; ++i) Hes[i][l] = hhost_[i];
    const double vk_norm = std::sqrt(std::max(0.0, hhost_[lmax + 5]));
    double       new_norm = std::sqrt(std::max(0.0, hhost_[l + 1]));
    // re-orthogonalise if the new vector is tiny relative to the original (loss of orthogonality safeguard)
    {
      const double temp = 1000.0 * vk_norm;
      if ((temp + new_norm) == temp) {
        double new_norm_2 = 0.0;
        for (int i = 0; i <= l; ++i) {
          double prod = 0.0;
          if (VecDot(V_[i], V_[l + 1], &prod)) return BDF_MEM_FAIL;
          if ((temp + prod) == temp) continue;
          Hes[i][l] += prod;
          VCHK(fspvec_axpy(w, -prod, V_[i]->d_data, n_local_, stream_));
          new_norm_2 += prod * prod;
        }
        if (new_norm_2 != 0.0) {
          const double new_product = new_norm * new_norm - new_norm_2;
          new_norm = new_product > 0.0 ? std::sqrt(new_product) : 0.0;
        }
      }
    }
    Hes[l + 1][l] = new_norm;
    // Givens QR update of column l
    {
      // apply previous rotations
      for (int k = 0; k < l; ++k) {
        const double c = givens[2 * k], s = givens[2 * k + 1];
        const double t1 = Hes[k][l], t2 = Hes[k + 1][l];
        Hes[k][l] = c * t1 - s * t2;
        Hes[k + 1][l] = s * t1 + c * t2;
      }
      const double t1 = Hes[l][l], t2 = Hes[l + 1][l];
      double       c, s;
      if (t2 == 0.0) { c = 1.0; s = 0.0; }
      else if (std::fabs(t2) >= std::fabs(t1)) { const double t3 = t1 / t2; s = -1.0 / std::sqrt(1.0 + t3 * t3); c = -s * t3; }
      else { const double t3 = t2 / t1; c = 1.0 / std::sqrt(1.0 + t3 * t3); s = -c * t3; }
      givens[2 * l] = c;
      givens[2 * l + 1] = s;
      Hes[l][l] = c * t1 - s * t2;
      if (Hes[l][l] == 0.0) return CONV_FAIL;  // singular least-squares system: recoverable
      rotation_product *= s;
      rho = std::fabs(rotation_product * beta);
    }
    if (rho <= delta) { conv = true; break; }
    // another iteration follows: normalise V_{l+1} and form its unscaled copy vtemp = V_{l+1} ./ s2 in one pass (the
    // last basis vector of a finished solve is never read, so it is neither normalised nor divided)
    if (l + 1 < lmax && new_norm > 0.0)
      VCHK(fspvec_scale_mul(w, 1.0 / new_norm, vtemp_->d_data, ewt_inv->d_data, n_local_, stream_));
    else if (l + 1 < lmax)
      VCHK(fspvec_prod(vtemp_->d_data, w, ewt_inv->d_data, n_local_, stream_));
  }
  // least-squares solution: yg = beta * Q e1 ; solve R yg = .
  const int lp1 = l_used + 1;
  yg[0] = beta;
  for (int i = 1; i < lp1; ++i) yg[i] = 0.0;
  for (int k = 0; k < l_used; ++k) {
    const double c = givens[2 * k], s = givens[2 * k + 1];
    const double t1 = yg[k], t2 = yg[k + 1];
    yg[k] = c * t1 - s * t2;
    yg[k + 1] = s * t1 + c * t2;
  }
  for (int k = l_used - 1; k >= 0; --k) {
    yg[k] /= Hes[k][k];
    for (int i = k - 1; i >= 0; --i) yg[i] -= yg[k] * Hes[i][k];
  }
  // xcor = sum_k yg_k V_k ; the unscaling x = xcor ./ s2 is fused into the Newton update of the caller
  {
    std::vector<const double *> ptrs((size_t) l_used);
    for (int k = 0; k < l_used; ++k) ptrs[k] = V_[k]->d_data;
    double beta_y = 0.0;
    for (int k0 = 0; k0 < l_used; k0 += 64) {
      const int mm = std::min(64, l_used - k0);
      VCHK(fspvec_maxpy(xcor_->d_data, beta_y, mm, yg.data() + k0, ptrs.data() + k0, n_local_, stream_));
      beta_y = 1.0;
    }
  }
  *xsrc = xcor_;
  *divide = true;
  // not converged within maxl iterations: accept if the residual was reduced (SUNLS_RES_REDUCED on the first
  // Newton iteration), otherwise a recoverable convergence failure
  if (!conv) {
    if (rho < beta && first_newton) { *converged = 1; return 0; }
    return CONV_FAIL;
  }
  *converged = 1;
  return 0;
}

// ---- Newton iteration (state if sens_index < 0, else sensitivity sens_index with staggered-1) ------------------------------
int BdfCore::nls(Vec zn0, Vec zn1, Vec ewt, Vec acor, Vec ycur, Vec ftemp, int sens_index, double *acnrm) {
  crate_ = 1.0;  // no linear-solver set-up phase exists for the matrix-free solver: reset every step
  VCHK(fspvec_set(acor->d_data, 0.0, n_local_, stream_));
  // f at the predicted value
  int r;
  if (sens_index < 0) r = rhs(tn_, zn0, ftemp);
  else { r = fs_(sens_index, tn_, y_, ftemp_, zn0, ftemp); nfe_ += 0; }
  if (r < 0) return BDF_RHS_FAIL;
  if (r > 0) return CONV_FAIL;
  if (V_.empty()) { V_.push_back(nullptr); if (alloc_like(zn0, &V_[0])) return BDF_MEM_FAIL; }
  double del = 0.0, delp = 0.0;
  int    m = 0;
  while (true) {
    nni_ += 1;
    // b = gamma f(y) - rl1 zn1 - acor ; V0 = s1 .* b ; ss = ||V0||^2 -- one pass (was lincomb3 + wrms + prod + norm)
    double *red = hdev_.get() + maxl_ + 6;
    VCHK(fspvec_lincomb3_wprod_sqsum(tempv_->d_data, V_[0]->d_data, gamma_, ftemp->d_data, -rl1_, zn1->d_data, -1.0,
                                     acor->d_data, ewt->d_data, red, n_local_, stream_));
    double ss_b = 0.0;
    if (global_sum(red, &ss_b)) return BDF_MEM_FAIL;
    int  conv = 0;
    Vec  xsrc = nullptr;
    bool divide = false;
    r = lin_solve(tempv_, ewt, ss_b, tn_, m == 0, &conv, &xsrc, &divide);
    if (r < 0) return r;
    if (r > 0 || !conv) return CONV_FAIL;
    // delta = xsrc (./ ewt) ; acor += delta ; ycur = zn0 + acor ; del = wrms(delta) -- one pass
    if (!xsrc) { VCHK(fspvec_set(delta_->d_data, 0.0, n_local_, stream_)); xsrc = delta_; divide = false; }
    if (divide)
      VCHK(fspvec_newton_update_mul(xsrc->d_data, inv_of(ewt)->d_data, acor->d_data, zn0->d_data, ycur->d_data, red,
                                    n_local_, stream_));
    else
      VCHK(fspvec_newton_update(xsrc->d_data, nullptr, ewt->d_data, acor->d_data, zn0->d_data, ycur->d_data, red,
                                n_local_, stream_));
    double ss_d = 0.0;
    if (global_sum(red, &ss_d)) return BDF_MEM_FAIL;
    del = std::sqrt(ss_d / n_global_);
    if (m > 0) crate_ = std::max(CRDOWN * crate_, del / delp);
    const double dcon = del * std::min(1.0, crate_) / tq_[4];
    if (dcon <= 1.0) {
      if (m == 0) *acnrm = del;
      else if (wrms(acor, ewt, acnrm)) return BDF_MEM_FAIL;
      return BDF_SUCCESS;
    }
    m++;
    if (m == maxcor_ || (m >= 2 && del > RDIV * delp)) return CONV_FAIL;
    delp = del;
    if (sens_index < 0) r = rhs(tn_, ycur, ftemp);
    else r = fs_(sens_index, tn_, y_, ftemp_, ycur, ftemp);
    if (r < 0) return BDF_RHS_FAIL;
    if (r > 0) return CONV_FAIL;
  }
}

int BdfCore::handle_nflag(int nflag, double saved_t, int *ncf, int *kflag) {
  if (nflag == BDF_SUCCESS) { *kflag = DO_ERROR_TEST; return 0; }
  ncfn_ += 1;
  restore(saved_t);
  if (nflag < 0) { *kflag = nflag; return nflag; }  // unrecoverable
  (*ncf)++;
  etamax_ = 1.0;
  if (std::fabs(h_) <= hmin_ * ONEPSM || *ncf == maxncf_) { *kflag = BDF_CONV_FAILURE; return BDF_CONV_FAILURE; }
  eta_ = std::max(ETACF, hmin_ / std::fabs(h_));
  rescale();
  *kflag = PREDICT_AGAIN;
  return 0;
}

int BdfCore::do_error_test(double saved_t, int *nef, double *dsm, int *again) {
  *again = 0;
  *dsm = acnrm_ * tq_[2];
  if (*dsm <= 1.0) return BDF_SUCCESS;
  (*nef)++;
  netf_ += 1;
  restore(saved_t);
  if (std::fabs(h_) <= hmin_ * ONEPSM || *nef == maxnef_) return BDF_ERR_FAILURE;
  etamax_ = 1.0;
  if (*nef <= MXNEF1) {
    eta_ = 1.0 / (std::pow(BIAS2 * (*dsm), 1.0 / L_) + ADDON);
    eta_ = std::max(ETAMIN, std::max(eta_, hmin_ / std::fabs(h_)));
    if (*nef >= SMALL_NEF) eta_ = std::min(eta_, ETAMXF);
    rescale();
    *again = 1;
    return BDF_SUCCESS;
  }
  if (q_ > 1) {
    eta_ = std::max(ETAMIN, hmin_ / std::fabs(h_));
    adjust_order(-1);
    L_ = q_;
    q_--;
    qwait_ = L_;
    rescale();
    *again = 1;
    return BDF_SUCCESS;
  }
  // already at order 1: restart the step from fresh derivative information
  flush_scale();
  eta_ = std::max(ETAMIN, hmin_ / std::fabs(h_));
  h_ *= eta_;
  next_h_ = h_;
  hscale_ = h_;
  qwait_ = LONG_WAIT;
  nscon_ = 0;
  int r = rhs(tn_, zn_[0], tempv_);
  if (r != 0) return BDF_RHS_FAIL;
  VCHK(fspvec_linear_sum(zn_[1]->d_data, h_, tempv_->d_data, 0.0, tempv_->d_data, n_local_, stream_));
  for (int is = 0; is < ns_; ++is) {
    r = fs_(is, tn_, zn_[0], tempv_, znS_[is][0], ftempS_[is]);
    if (r != 0) return BDF_RHS_FAIL;
    VCHK(fspvec_linear_sum(znS_[is][1]->d_data, h_, ftempS_[is]->d_data, 0.0, ftempS_[is]->d_data, n_local_, stream_));
  }
  *again = 1;
  return BDF_SUCCESS;
}

void BdfCore::complete_step() {
  nst_++;
  nscon_++;
  hu_ = h_;
  qu_ = q_;
  for (int i = q_; i >= 2; --i) tau_[i] = tau_[i - 1];
  if (q_ == 1 && nst_ > 1) tau_[2] = tau_[1];
  tau_[1] = h_;
  {
    // zn[j] += l[j] acor for j = 0..q in one pass (cvCompleteStep)
    double *Z[LMAX + 1];
    for (int j = 0; j <= q_; ++j) Z[j] = zn_[j]->d_data;
    vec_status_ |= fspvec_multi_axpy(Z, q_ + 1, l_, acor_->d_data, n_local_, stream_);
    for (int is = 0; is < ns_; ++is) {
      for (int j = 0; j <= q_; ++j) Z[j] = znS_[is][j]->d_data;
      vec_status_ |= fspvec_multi_axpy(Z, q_ + 1, l_, acorS_[is]->d_data, n_local_, stream_);
    }
  }
  qwait_--;
  if (qwait_ == 1 && q_ != QMAX) {
    vec_status_ |= VecCopy(acor_, zn_[QMAX]);
    for (int is = 0; is < ns_; ++is) vec_status_ |= VecCopy(acorS_[is], znS_[is][QMAX]);
    saved_tq5_ = tq_[5];
    indx_acor_ = QMAX;
  }
}

void BdfCore::set_eta() {
  if (eta_ < THRESH) {
    eta_ = 1.0;
    hprime_ = h_;
  } else {
    eta_ = std::min(eta_, etamax_);
    eta_ /= std::max(1.0, std::fabs(h_) * hmax_inv_ * eta_);
    hprime_ = h_ * eta_;
    if (qprime_ < q_) nscon_ = 0;
  }
}

double BdfCore::compute_etaqm1() {
  etaqm1_ = 0.0;
  if (q_ > 1) {
    double ddn = 0.0;
    wrms(zn_[q_], ewt_, &ddn);
    if (errcon_)
      for (int is = 0; is < ns_; ++is) { double d = 0.0; wrms(znS_[is][q_], ewtS_[is], &d); ddn = std::max(ddn, d); }
    ddn *= tq_[1];
    etaqm1_ = 1.0 / (std::pow(BIAS1 * ddn, 1.0 / q_) + ADDON);
  }
  return etaqm1_;
}

double BdfCore::compute_etaqp1() {
  etaqp1_ = 0.0;
  if (q_ != QMAX) {
    if (saved_tq5_ == 0.0) return etaqp1_;
    const double cquot = (tq_[5] / saved_tq5_) * std::pow(h_ / tau_[2], (double) L_);
    vec_status_ |= fspvec_linear_sum(tempv_->d_data, -cquot, zn_[QMAX]->d_data, 1.0, acor_->d_data, n_local_, stream_);
    double dup = 0.0;
    wrms(tempv_, ewt_, &dup);
    if (errcon_)
      for (int is = 0; is < ns_; ++is) {
        vec_status_ |= fspvec_linear_sum(tempv_->d_data, -cquot, znS_[is][QMAX]->d_data, 1.0, acorS_[is]->d_data, n_local_, stream_);
        double d = 0.0;
        wrms(tempv_, ewtS_[is], &d);
        dup = std::max(dup, d);
      }
    dup *= tq_[3];
    etaqp1_ = 1.0 / (std::pow(BIAS3 * dup, 1.0 / (L_ + 1)) + ADDON);
  }
  return etaqp1_;
}

void BdfCore::choose_eta() {
  const double etam = std::max(etaqm1_, std::max(etaq_, etaqp1_));
  if (etam < THRESH) {
    eta_ = 1.0;
    qprime_ = q_;
    return;
  }
  if (etam == etaq_) {
    eta_ = etaq_;
    qprime_ = q_;
  } else if (etam == etaqm1_) {
    eta_ = etaqm1_;
    qprime_ = q_ - 1;
  } else {
    eta_ = etaqp1_;
    qprime_ = q_ + 1;
    // keep the correction for the order increase
    vec_status_ |= VecCopy(acor_, zn_[QMAX]);
    for (int is = 0; is < ns_; ++is) vec_status_ |= VecCopy(acorS_[is], znS_[is][QMAX]);
  }
}

void BdfCore::prepare_next_step(double dsm) {
  if (etamax_ == 1.0) {
    qwait_ = std::max(qwait_, 2);
    qprime_ = q_;
    hprime_ = h_;
    eta_ = 1.0;
    return;
  }
  etaq_ = 1.0 / (std::pow(BIAS2 * dsm, 1.0 / L_) + ADDON);
  if (qwait_ != 0) {
    eta_ = etaq_;
    qprime_ = q_;
    set_eta();
    return;
  }
  qwait_ = 2;
  compute_etaqm1();
  compute_etaqp1();
  choose_eta();
  set_eta();
}

// ---- one step ---------------------------------------------------------------------------------------------------
int BdfCore::Step(double *t_reached, Vec yout, Vec *sout) {
  vec_status_ = 0;
  if (first_) {
    if (ewt_set(zn_[0], ewt_, ewt_inv_)) return BDF_ILL_EWT;
    int r = rhs(tn_, zn_[0], zn_[1]);
    if (r != 0) return BDF_RHS_FAIL;
    for (int is = 0; is < ns_; ++is) {
      if (ewt_set(znS_[is][0], ewtS_[is], ewtS_inv_[is])) return BDF_ILL_EWT;
      r = fs_(is, tn_, zn_[0], zn_[1], znS_[is][0], znS_[is][1]);
      if (r != 0) return BDF_RHS_FAIL;
    }
    r = initial_step(tout_hint_);
    if (r != 0) return r;
    hscale_ = h_;
    hprime_ = h_;
    next_h_ = h_;
    VCHK(fspvec_scale(zn_[1]->d_data, h_, n_local_, stream_));
    for (int is = 0; is < ns_; ++is) VCHK(fspvec_scale(znS_[is][1]->d_data, h_, n_local_, stream_));
    first_ = false;
  } else {
    if (ewt_set(zn_[0], ewt_, ewt_inv_)) return BDF_ILL_EWT;
    for (int is = 0; is < ns_; ++is) if (ewt_set(znS_[is][0], ewtS_[is], ewtS_inv_[is])) return BDF_ILL_EWT;
  }

  const double saved_t = tn_;
  int          ncf = 0, nef = 0, nefS = 0, nflag;
  double       dsm = 0.0;
  if (nst_ > 0 && hprime_ != h_) adjust_params();
  while (true) {
    predict();
    set_coeffs();
    nflag = nls(zn_[0], zn_[1], ewt_, acor_, y_, ftemp_, -1, &acnrm_);
    int kflag;
    int rc = handle_nflag(nflag, saved_t, &ncf, &kflag);
    if (kflag == PREDICT_AGAIN) continue;
    if (kflag != DO_ERROR_TEST) return rc ? rc : kflag;
    int again = 0;
    rc = do_error_test(saved_t, &nef, &dsm, &again);
    if (rc != BDF_SUCCESS) return rc;
    if (again) continue;

    if (ns_ > 0) {
      // staggered-1: re-evaluate f at the converged y, then correct each sensitivity in turn
      int r = rhs(tn_, y_, ftemp_);
      if (r != 0) { nflag = r < 0 ? (int) BDF_RHS_FAIL : (int) CONV_FAIL; }
      else {
        nflag = BDF_SUCCESS;
        for (int is = 0; is < ns_; ++is) {
          nflag = nls(znS_[is][0], znS_[is][1], ewtS_[is], acorS_[is], yS_[is], ftempS_[is], is, &acnrmS_[is]);
          if (nflag != BDF_SUCCESS) break;
        }
      }
      rc = handle_nflag(nflag, saved_t, &ncf, &kflag);
      if (kflag == PREDICT_AGAIN) continue;
      if (kflag != DO_ERROR_TEST) return rc ? rc : kflag;
      if (errcon_) {
        double save_acnrm = acnrm_, dsmS = 0.0;
        acnrm_ = 0.0;
        for (int is = 0; is < ns_; ++is) acnrm_ = std::max(acnrm_, acnrmS_[is]);
        rc = do_error_test(saved_t, &nefS, &dsmS, &again);
        acnrm_ = save_acnrm;
        if (rc != BDF_SUCCESS) return rc;
        if (again) continue;
        dsm = std::max(dsm, dsmS);
      }
    }
    break;
  }
  if (accept_) {
    bool reject = false;
    if (accept_(tn_, y_, &reject) != 0) return BDF_RHS_FAIL;
    if (reject) {
      restore(saved_t);  // the history is the one of the last committed step again
      if (vec_status_) return BDF_MEM_FAIL;
      *t_reached = tn_;
      if (yout) VCHK(VecCopy(zn_[0], yout));
      if (sout) for (int is = 0; is < ns_; ++is) VCHK(VecCopy(znS_[is][0], sout[is]));
      return STOPPED;
    }
  }
  complete_step();
  prepare_next_step(dsm);
  etamax_ = (nst_ <= SMALL_NST) ? ETAMX2 : ETAMX3;
  // acor <- estimated local error
  VCHK(fspvec_scale(acor_->d_data, tq_[2], n_local_, stream_));
  if (vec_status_) return BDF_MEM_FAIL;
  *t_reached = tn_;
  if (yout) VCHK(VecCopy(zn_[0], yout));
  if (sout) for (int is = 0; is < ns_; ++is) VCHK(VecCopy(znS_[is][0], sout[is]));
  return BDF_SUCCESS;
}

int BdfCore::interpolate(double t, Vec *zn, Vec out) {
  flush_scale();
  const double tfuzz0 = FUZZ_FACTOR * DBL_EPSILON * (std::fabs(tn_) + std::fabs(hu_));
  const double tfuzz = hu_ < 0.0 ? -tfuzz0 : tfuzz0;
  const double tp = tn_ - hu_ - tfuzz, tn1 = tn_ + tfuzz;
  if ((t - tp) * (t - tn1) > 0.0) return BDF_BAD_T;
  const double s = (t - tn_) / h_;
  // dky = sum_j s^j zn[j] (Horner from j = q down to 0): one fused pass
  std::vector<double>         c((size_t) q_ + 1);
  std::vector<const double *> ptrs((size_t) q_ + 1);
  double                      sp = 1.0;
  for (int j = 0; j <= q_; ++j) { c[j] = sp; sp *= s; ptrs[j] = zn[j]->d_data; }
  VCHK(fspvec_maxpy(out->d_data, 0.0, q_ + 1, c.data(), ptrs.data(), n_local_, stream_));
  return 0;
}

int BdfCore::GetDky(double t, Vec yout) { return interpolate(t, zn_, yout); }

int BdfCore::TaylorRestart(const DerivOpFn &dop, int max_time_deriv) {
  if (first_ || !zn_[0] || ns_ > 0) return 1;  // (sensitivities keep the plain carry-over)
  vec_status_ = 0;
  scale_pending_ = false;  // zn_[1..q] are rebuilt from scratch: a pending rescale of the old history is moot
  static const double binom[6][6] = {{1, 0, 0, 0, 0, 0}, {1, 1, 0, 0, 0, 0}, {1, 2, 1, 0, 0, 0},
                                     {1, 3, 3, 1, 0, 0}, {1, 4, 6, 4, 1, 0}, {1, 5, 10, 10, 5, 1}};
  // unscaled derivatives Y_k = y^(k)(t_n) in zn_[k]
  for (int k = 0; k < q_; ++k) {
    nfe_ += 1;
    if (dop(0, tn_, zn_[k], zn_[k + 1]) != 0) return BDF_RHS_FAIL;
    for (int j = 1; j <= k && j <= max_time_deriv; ++j) {
      nfe_ += 1;
      if (dop(j, tn_, zn_[k - j], tempv_) != 0) return BDF_RHS_FAIL;
      VCHK(fspvec_axpy(zn_[k + 1]->d_data, binom[k][j], tempv_->d_data, n_local_, stream_));
    }
  }
  double fac = 1.0;
  for (int k = 1; k <= q_; ++k) {
    fac *= h_ / k;
    VCHK(fspvec_scale(zn_[k]->d_data, fac, n_local_, stream_));
  }
  for (int k = q_ + 1; k <= QMAX; ++k) VCHK(fspvec_set(zn_[k]->d_data, 0.0, n_local_, stream_));
  VCHK(fspvec_set(acor_->d_data, 0.0, n_local_, stream_));
  // controller state of a uniform history at step h_: no order change for q + 1 steps, normal growth limits
  hscale_ = hprime_ = next_h_ = h_;
  eta_ = 1.0;
  qprime_ = q_;
  L_ = q_ + 1;
  qwait_ = L_;
  nscon_ = 0;
  indx_acor_ = QMAX;
  for (int j = 1; j <= LMAX; ++j) tau_[j] = h_;
  etamax_ = ETAMX3;
  return 0;
}

int BdfCore::Expand(const std::vector<PetscInt> &new_indices, PetscInt new_local_size) {
  if (first_ || !zn_[0]) return 1;  // nothing worth keeping
  vec_status_ = 0;
  flush_scale();
  if (vec_status_) return BDF_MEM_FAIL;
  // persistent vectors: the Nordsieck array (all QMAX + 1 columns: column indx_acor_ may hold the saved correction of
  // an order-increase decision) and the last correction / local error estimate
  for (int j = 0; j <= QMAX; ++j) if (ExpandVec(zn_[j], new_indices, new_local_size)) return BDF_MEM_FAIL;
  if (ExpandVec(acor_, new_indices, new_local_size)) return BDF_MEM_FAIL;
  for (int is = 0; is < ns_; ++is) {
    for (int j = 0; j <= QMAX; ++j) if (ExpandVec(znS_[is][j], new_indices, new_local_size)) return BDF_MEM_FAIL;
    if (ExpandVec(acorS_[is], new_indices, new_local_size)) return BDF_MEM_FAIL;
  }
  // scratch vectors: new size, contents undefined
  for (Vec *v : {&ewt_, &ewt_inv_, &y_, &tempv_, &ftemp_, &xcor_, &vtemp_, &delta_}) {
    if (*v) VecDestroy(v);
    if (alloc_like(zn_[0], v)) return BDF_MEM_FAIL;
  }
  for (auto &v : V_) if (v) VecDestroy(&v);
  V_.clear();
  for (int is = 0; is < ns_; ++is)
    for (auto *vv : {&ewtS_, &ewtS_inv_, &yS_, &ftempS_}) {
      if ((*vv)[is]) VecDestroy(&(*vv)[is]);
      if (alloc_like(zn_[0], &(*vv)[is])) return BDF_MEM_FAIL;
    }
  n_local_ = zn_[0]->n_local;
  PetscInt ng = 0;
  VecGetSize(zn_[0], &ng);
  n_global_ = (double) ng;
  return 0;
}
int BdfCore::GetSensDky(double t, int is, Vec sout) { return interpolate(t, znS_[is].data(), sout); }

}  // namespace pacmensl

// BdfCore.h -- variable-order (1..5), variable-step BDF integrator in Nordsieck form with a matrix-free
// Newton iteration whose linear systems (I - gamma J) d = r are solved by scaled, unpreconditioned GMRES
// with modified Gram-Schmidt.  Written from scratch on device vectors (fspvec_* kernels); it restates the
// published algorithm of SUNDIALS CVODE 5.7.0 (BDF + Newton + SUNLinSol_SPGMR), which is what the reference
// configures in src/OdeSolver/CvodeFsp.cpp:174-198 (CV_BDF, scalar tolerances, SPGMR(maxl) without
// preconditioner, user J*v) -- the SUNDIALS sources are not part of the reference tree and are not used.
// Step-size/order heuristics, error-test constants and the GMRES stopping rule follow CVODE's documented
// behaviour (Cohen & Hindmarsh 1996; SUNDIALS CVODE user guide, "Mathematical considerations").
//
// Optional forward sensitivities use the staggered-1 corrector (CV_STAGGERED1,
// src/SensFsp/ForwardSensCvodeFsp.cpp:218-223): after the state step converges, each sensitivity system
// is corrected in turn with the same Newton-GMRES machinery; with error control on, sensitivities enter the
// local error test.
#pragma once

#include <functional>
#include <vector>

#include "Sys.h"

namespace pacmensl {

class BdfCore {
 public:
  // y' = f(t, y); returns 0 ok, >0 recoverable, <0 fatal
  using RhsFn = std::function<int(double t, Vec y, Vec ydot)>;
  // Jv = J(t) v
  using JtvFn = std::function<int(double t, Vec v, Vec Jv)>;
  // sensitivity right-hand side: sdot = J s_i + (df/dtheta_i)(t, y)
  using SensRhsFn = std::function<int(int is, double t, Vec y, Vec ydot, Vec s, Vec sdot)>;
  /// J v with a fused epilogue (FspMatrixBase::ActionFused); optional, used by the state GMRES loop when set
  using FusedJtvFn = std::function<int(double t, Vec v, Vec out, const fspmat_epilogue &ep)>;
  void SetFusedJtv(FusedJtvFn f) { fused_jtv_ = std::move(f); }
  /// Consulted after a step has converged and passed its error test but BEFORE it is committed: (t_new, y_new).
  /// *reject = true undoes the step (cvRestore: the history is back at the previous time, step size and order
  /// unchanged) and Step returns STOPPED.  This is how the FSP driver's sink check stops the integrator without losing
  /// its history (a check after the commit would need a copy of the whole Nordsieck array per step to roll back).
  using AcceptFn = std::function<int(double t_new, Vec y_new, bool *reject)>;
  void SetAcceptHook(AcceptFn f) { accept_ = std::move(f); }
  static constexpr int STOPPED = 10;  ///< Step(): the accept hook rejected the step; state is at the previous time
  /// Re-map the history onto an enlarged state space (entry i moves to new_indices[i], new entries are zero), keeping
  /// step size, order and all controller state: the integrator continues instead of restarting at order 1.
  int  Expand(const std::vector<PetscInt> &new_indices, PetscInt new_local_size);
  long LocalSize() const { return n_local_; }
  /// Taylor restart (after Expand): rebuild the Nordsieck array at the current (t_n, h, q) from the exact time
  /// derivatives of the solution of the LINEAR system y' = A(t) y,
  ///     y^(k+1) = sum_{j=0..k} C(k, j) A^(j)(t_n) y^(k-j),    zn[k] = h^k y^(k) / k!,
  /// with dop(j, t, v, out) = A^(j)(t) v (j = 0: the right-hand side; j >= 1 only while j <= max_time_deriv: 0 for a
  /// time-invariant operator).  q Actions for a time-invariant operator, q(q+1)/2 at most otherwise.  The new
  /// components then carry their full derivative history, so the integrator continues at (nearly) its old step size
  /// and order instead of climbing back from order 1 and a first step sized for the atol = 1e-14 weights.
  using DerivOpFn = std::function<int(int j, double t, Vec v, Vec out)>;
  int  TaylorRestart(const DerivOpFn &dop, int max_time_deriv);

  explicit BdfCore(MPI_Comm comm);
  ~BdfCore();

  void SetTolerances(double rtol, double atol) { rtol_ = rtol; atol_ = atol; }
  void SetMaxKrylov(int maxl) { maxl_ = maxl; }
  void SetMaxNonlinIters(int m) { maxcor_ = m; }
  void SetMaxConvFails(int m) { maxncf_ = m; }
  void SetMaxErrTestFails(int m) { maxnef_ = m; }

  /// Initialise at (t0, y0).  y0 is copied.  tout_hint gives the direction and scale for the first step.
  int Init(double t0, Vec y0, RhsFn f, JtvFn jtv, double tout_hint);
  /// Enable ns forward sensitivities with initial values s0[i] (copied); errcon: include them in the error test.
  int InitSens(int ns, Vec *s0, SensRhsFn fs, bool errcon);

  /// Take one internal step (CV_ONE_STEP).  On return t = reached time and yout (and sout) hold the solution.
  int Step(double *t_reached, Vec yout, Vec *sout = nullptr);
  /// Interpolated solution at t in [tn - hu, tn] (CVodeGetDky with k = 0).
  int GetDky(double t, Vec yout);
  int GetSensDky(double t, int is, Vec sout);
  void Free();

  double CurrentTime() const { return tn_; }
  double LastStep() const { return hu_; }
  int    LastOrder() const { return qu_; }
  long   NumSteps() const { return nst_; }
  long   NumRhsEvals() const { return nfe_; }
  long   NumJtvEvals() const { return njtv_; }
  long   NumLinIters() const { return nli_; }
  long   NumNonlinIters() const { return nni_; }
  long   NumErrTestFails() const { return netf_; }
  long   NumConvFails() const { return ncfn_; }

 private:
  static constexpr int QMAX = 5, LMAX = QMAX + 1;
  MPI_Comm comm_;
  void    *stream_ = nullptr;
  long     n_local_ = 0;
  double   n_global_ = 1.0;
  double   rtol_ = 1e-6, atol_ = 1e-14;
  int      maxl_ = 100, maxcor_ = 3, maxncf_ = 10, maxnef_ = 7;

  RhsFn     f_;
  JtvFn     jtv_;
  SensRhsFn fs_;
  FusedJtvFn fused_jtv_;
  AcceptFn   accept_;

  // Nordsieck history and work vectors
  Vec zn_[LMAX + 1] = {nullptr};
  Vec ewt_inv_ = nullptr;  ///< rtol |y| + atol = 1 ./ ewt_
  Vec ewt_ = nullptr, y_ = nullptr, acor_ = nullptr, tempv_ = nullptr, ftemp_ = nullptr;
  // GMRES workspace (basis allocated on demand up to maxl_ + 1)
  std::vector<Vec> V_;
  Vec xcor_ = nullptr, vtemp_ = nullptr, delta_ = nullptr;
  DeviceBuffer<double> hdev_;
  std::vector<double>  hhost_;

  // sensitivities
  int  ns_ = 0;
  bool errcon_ = false;
  std::vector<std::vector<Vec>> znS_;  // [is][j]
  std::vector<Vec> ewtS_, ewtS_inv_, acorS_, yS_, ftempS_;
  std::vector<double> acnrmS_;

  // integrator state (names follow the CVODE literature)
  int    q_ = 1, L_ = 2, qprime_ = 1, qwait_ = 2, qu_ = 0, nscon_ = 0, indx_acor_ = QMAX;
  double h_ = 0, hprime_ = 0, hscale_ = 0, hu_ = 0, next_h_ = 0, eta_ = 1, etamax_ = 10000.0;
  double tn_ = 0, tau_[LMAX + 1] = {0}, tq_[6] = {0}, l_[LMAX + 1] = {0};
  double rl1_ = 1, gamma_ = 0, gammap_ = 0, gamrat_ = 1, crate_ = 1, acnrm_ = 0, saved_tq5_ = 0;
  double etaqm1_ = 0, etaq_ = 0, etaqp1_ = 0, hmin_ = 0, hmax_inv_ = 0;
  double tout_hint_ = 0;
  long   nst_ = 0, nfe_ = 0, njtv_ = 0, nli_ = 0, nni_ = 0, netf_ = 0, ncfn_ = 0;
  bool   first_ = true;

  // helpers
  int  alloc_like(Vec proto, Vec *out);
  int  ewt_set(Vec y, Vec ewt, Vec ewt_inv);
  Vec  inv_of(Vec ewt) const;
  int  wrms(Vec v, Vec w, double *out);
  int  rhs(double t, Vec y, Vec ydot);
  int  initial_step(double tout);
  void adjust_params();
  void adjust_order(int deltaq);
  void increase_bdf();
  void decrease_bdf();
  void rescale();
  void history_pass(const double *scale, int pascal, int q);  ///< fused rescale + Pascal product over zn_ (and znS_)
  int  pending_q_ = 0;
  int  global_sum(const double *dev_scalar, double *out);
  void flush_scale();                                  ///< apply a rescale that no prediction has absorbed yet
  double pending_scale_[LMAX + 1] = {0};
  bool   scale_pending_ = false;
  void predict();
  void restore(double saved_t);
  void set_coeffs();
  void set_tq(double hsum, double alpha0, double alpha0_hat, double xi_inv, double xistar_inv);
  int  nls(Vec zn0, Vec zn1, Vec ewt, Vec acor, Vec ycur, Vec ftemp, int sens_index, double *acnrm);
  int  lin_solve(Vec b, Vec ewt, double ss_b, double tn, bool first_newton, int *converged, Vec *xsrc, bool *divide);
  void complete_step();
  void prepare_next_step(double dsm);
  void set_eta();
  double compute_etaqm1();
  double compute_etaqp1();
  void choose_eta();
  int  handle_nflag(int nflag, double saved_t, int *ncf, int *kflag);
  int  do_error_test(double saved_t, int *nef, double *dsm, int *again);
  int  interpolate(double t, Vec *zn, Vec out);
  int  vec_status_ = 0;  // sticky error of device vector calls
};

}  // namespace pacmensl

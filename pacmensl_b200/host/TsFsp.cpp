#include "TsFsp.h"

#include <algorithm>
#include <cmath>

namespace pacmensl {

namespace {
// RA34PW2 (Rang & Angermann 2005): A = a_ij, G = gamma_ij (diagonal gamma), b = third-order weights, b2 = embedded
// second-order weights.
constexpr double kGamma = 4.3586652150845900e-01;
constexpr double kA[4][4] = {{0, 0, 0, 0},
                             {8.7173304301691801e-01, 0, 0, 0},
                             {8.4457060015369423e-01, -1.1299064236484185e-01, 0, 0},
                             {0, 0, 1.0, 0}};
constexpr double kG[4][4] = {{kGamma, 0, 0, 0},
                             {-8.7173304301691801e-01, kGamma, 0, 0},
                             {-9.0338057013044082e-01, 5.4180672388095326e-02, kGamma, 0},
                             {2.4212380706095346e-01, -1.2232505839045147e+00, 5.4526025533510214e-01, kGamma}};
constexpr double kB[4] = {2.4212380706095346e-01, -1.2232505839045147e+00, 1.5452602553351020e+00, 4.3586652150845900e-01};
constexpr double kB2[4] = {3.7810903145819369e-01, -9.6042292212423178e-02, 5.0000000000000000e-01, 2.1793326075422950e-01};
constexpr int    kRestart = 30;
}  // namespace

TsFsp::TsFsp(MPI_Comm _comm) : OdeSolverBase(_comm) {}
TsFsp::~TsFsp() { FreeWorkspace(); }

PacmenslErrorCode TsFsp::SetTsType(std::string type) {
  type_ = std::move(type);
  return 0;
}

// src/OdeSolver/TsFsp.cpp:31-79
PacmenslErrorCode TsFsp::SetUp() {
  if (solution_ == nullptr || rhs_ == nullptr) return -1;
  if (type_ != TSROSW) {
    PetscPrintf(comm_, "TsFsp: only the Rosenbrock-W type (\"rosw\", PETSc's default) is provided; got \"%s\".\n", type_.c_str());
    return -1;
  }

  for (Vec *v : {&k_[0], &k_[1], &k_[2], &k_[3], &ystage_, &rhsv_, &tmp_, &ynew_, &err_, &f0_, &f1_, &jdiag_, &pc_})
    if (*v) VecDestroy(v);
  for (auto &v : V_) if (v) VecDestroy(&v);
  V_.clear();
  if (J) { MatDestroy(&J); J = nullptr; }
  for (Vec *v : {&k_[0], &k_[1], &k_[2], &k_[3], &ystage_, &rhsv_, &tmp_, &ynew_, &err_, &f0_, &f1_, &jdiag_, &pc_}) {
    int ierr = VecDuplicate(*solution_, v);
    CHKERRQ(ierr);
  }
  // the assembled Jacobian of the reference's implicit TS types (TsFsp.cpp:63-69); matrix-free where none exists
  if (fspmat_ && comm_size_ == 1) {
    if (fspmat_->CreateRHSJacobian(&J) != 0) J = nullptr;
  }
  njac = nstep = nreject = 0;
  nlin = 0;
  have_h_ = false;
  return 0;
}

// left Jacobi preconditioner of M = I - h gamma J (the generator's rows are dominated by their diagonal): out = in ./ pc
int TsFsp::ApplyPc(Vec in, Vec out) {
  if (!J) return in == out ? 0 : VecCopy(in, out);
  FSPCHKERRQ(fspvec_div(out->d_data, in->d_data, pc_->d_data, in->n_local, comm_ ? comm_->stream : nullptr));
  return 0;
}

int TsFsp::JacTimes(PetscReal t, Vec v, Vec out) {
  if (J) return MatMult(J, v, out);
  return EvaluateRHS(t, v, out);  // linear system: J(t) v = A(t) v
}

int TsFsp::WrmsNorm(Vec e, Vec ya, Vec yb, double *out) {
  // sqrt( mean_i ( e_i / (atol + rtol max(|ya_i|, |yb_i|)) )^2 )   (TSErrorWeightedNorm, NORM_2)
  const long n = e->n_local;
  void      *stream = comm_ ? comm_->stream : nullptr;
  // tmp_ = atol + rtol * max(|ya|, |yb|) is not available as one fused pass; |ya| and |yb| are within the step's
  // change of each other, so the weights of ya are used for the denominator and those of yb as a second evaluation
  DeviceBuffer<double> red(2);
  FSPCHKERRQ(fspvec_ewt(tmp_->d_data, ya->d_data, rel_tol_, abs_tol_, n, nullptr, stream));   // 1 / (rtol |ya| + atol)
  FSPCHKERRQ(fspvec_wsqsum(red.get(), e->d_data, tmp_->d_data, n, stream));
  FSPCHKERRQ(fspvec_ewt(tmp_->d_data, yb->d_data, rel_tol_, abs_tol_, n, nullptr, stream));
  FSPCHKERRQ(fspvec_wsqsum(red.get() + 1, e->d_data, tmp_->d_data, n, stream));
  double s[2] = {0.0, 0.0};
  FSPCHKERRQ(fsp_memcpy_d2h(s, red.get(), sizeof(double) * 2, stream));
  if (pacmensl_allreduce_sum(comm_, s, 2)) return -1;
  PetscInt ng = 0;
  VecGetSize(e, &ng);
  *out = std::sqrt(std::min(s[0], s[1]) / std::max(1, ng));  // the larger weight vector (max of |ya|, |yb|) gives the smaller norm
  return 0;
}

// Restarted GMRES(30) with modified Gram-Schmidt on (I - hgamma J) x = b, zero initial guess per cycle start; stops
// when the residual's 2-norm is below tol_scale * ||b||_2.
int TsFsp::SolveStage(PetscReal t, double hgamma, Vec b, Vec x, double rel_tol_lin) {
  int    ierr;
  double bnorm = 0.0;
  if (J) {  // pc_ = 1 - hgamma * diag(J); sink rows have no diagonal entry (jdiag_ = 0 there)
    ierr = VecSet(pc_, 1.0); CHKERRQ(ierr);
    ierr = VecAXPY(pc_, -hgamma, jdiag_); CHKERRQ(ierr);
  }
  ierr = ApplyPc(b, b); CHKERRQ(ierr);  // the system solved is P^-1 M x = P^-1 b (b is scratch of the caller)
  ierr = VecNorm(b, NORM_2, &bnorm); CHKERRQ(ierr);
  ierr = VecSet(x, 0.0); CHKERRQ(ierr);
  if (bnorm == 0.0) return 0;
  const double target = rel_tol_lin * bnorm;
  if ((int) V_.size() < kRestart + 1) {
    V_.resize(kRestart + 1, nullptr);
    for (auto &v : V_) if (!v) { ierr = VecDuplicate(b, &v); CHKERRQ(ierr); }
  }
  std::vector<double> H((size_t) (kRestart + 1) * kRestart, 0.0), cs(kRestart), sn(kRestart), g(kRestart + 1);
  for (int cycle = 0; cycle < 10; ++cycle) {
    // r = b - M x
    if (cycle == 0) { ierr = VecCopy(b, V_[0]); CHKERRQ(ierr); }
    else {
      ierr = JacTimes(t, x, tmp_); if (ierr) return ierr;
      ierr = VecWAXPY(tmp_, -hgamma, tmp_, x); CHKERRQ(ierr);   // M x = x - hgamma J x
      ierr = ApplyPc(tmp_, tmp_); CHKERRQ(ierr);
      ierr = VecWAXPY(V_[0], -1.0, tmp_, b); CHKERRQ(ierr);     // P^-1 b - P^-1 M x
    }
    double beta = 0.0;
    ierr = VecNorm(V_[0], NORM_2, &beta); CHKERRQ(ierr);
    if (beta <= target) return 0;
    ierr = VecScale(V_[0], 1.0 / beta); CHKERRQ(ierr);
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    int m = 0;
    for (; m < kRestart; ++m) {
      nlin += 1;
      // w = M v_m = v_m - hgamma J v_m
      ierr = JacTimes(t, V_[m], tmp_); if (ierr) return ierr;
      ierr = VecWAXPY(V_[m + 1], -hgamma, tmp_, V_[m]); CHKERRQ(ierr);
      ierr = ApplyPc(V_[m + 1], V_[m + 1]); CHKERRQ(ierr);
      for (int i = 0; i <= m; ++i) {
        double hij = 0.0;
        ierr = VecDot(V_[m + 1], V_[i], &hij); CHKERRQ(ierr);
        H[(size_t) i * kRestart + m] = hij;
        ierr = VecAXPY(V_[m + 1], -hij, V_[i]); CHKERRQ(ierr);
      }
      double hn = 0.0;
      ierr = VecNorm(V_[m + 1], NORM_2, &hn); CHKERRQ(ierr);
      H[(size_t) (m + 1) * kRestart + m] = hn;
      if (hn > 0.0) { ierr = VecScale(V_[m + 1], 1.0 / hn); CHKERRQ(ierr); }
      for (int i = 0; i < m; ++i) {  // apply the previous Givens rotations to column m
        const double a = H[(size_t) i * kRestart + m], c = H[(size_t) (i + 1) * kRestart + m];
        H[(size_t) i * kRestart + m] = cs[i] * a + sn[i] * c;
        H[(size_t) (i + 1) * kRestart + m] = -sn[i] * a + cs[i] * c;
      }
      const double a = H[(size_t) m * kRestart + m], c = H[(size_t) (m + 1) * kRestart + m], r = std::hypot(a, c);
      cs[m] = r > 0.0 ? a / r : 1.0;
      sn[m] = r > 0.0 ? c / r : 0.0;
      H[(size_t) m * kRestart + m] = r;
      H[(size_t) (m + 1) * kRestart + m] = 0.0;
      g[m + 1] = -sn[m] * g[m];
      g[m] = cs[m] * g[m];
      if (std::fabs(g[m + 1]) <= target || hn == 0.0) { ++m; break; }
    }
    // back substitution, x += V y
    std::vector<double> yv((size_t) m, 0.0);
    for (int i = m - 1; i >= 0; --i) {
      double s = g[i];
      for (int j = i + 1; j < m; ++j) s -= H[(size_t) i * kRestart + j] * yv[j];
      yv[i] = s / H[(size_t) i * kRestart + i];
    }
    ierr = VecMAXPY(x, m, yv.data(), V_.data()); CHKERRQ(ierr);
    if (std::fabs(g[m]) <= target) return 0;
  }
  return 1;  // not converged: the caller rejects the step and shrinks h
}

// cubic Hermite interpolation on [t0, t1] from (y0, f0_) and (y1, f1_): third-order dense output (the role of
// TSInterpolate in TSCheckFspError, TsFsp.cpp:152-166)
int TsFsp::Interpolate(PetscReal t0, PetscReal t1, Vec y0, Vec y1, PetscReal t, Vec out) {
  const double h = t1 - t0, s = (t - t0) / h;
  const double h00 = (1 + 2 * s) * (1 - s) * (1 - s), h10 = s * (1 - s) * (1 - s), h01 = s * s * (3 - 2 * s), h11 = s * s * (s - 1);
  int ierr = VecSet(out, 0.0); CHKERRQ(ierr);
  ierr = VecAXPY(out, h00, y0); CHKERRQ(ierr);
  ierr = VecAXPY(out, h10 * h, f0_); CHKERRQ(ierr);
  ierr = VecAXPY(out, h01, y1); CHKERRQ(ierr);
  ierr = VecAXPY(out, h11 * h, f1_); CHKERRQ(ierr);
  return 0;
}

// src/OdeSolver/TsFsp.cpp:81-108 (TSSolve) + :131-197 (TSCheckFspError after every step)
PetscInt TsFsp::Solve() {
  if (solution_ == nullptr || rhs_ == nullptr || !k_[0]) return -1;
  if (solution_tmp_) VecDestroy(&solution_tmp_);
  int ierr = VecDuplicate(*solution_, &solution_tmp_); CHKERRQ(ierr);
  ierr = VecCopy(*solution_, solution_tmp_); CHKERRQ(ierr);
  fsp_stop_ = 0;
  Vec y = solution_tmp_;
  if (!have_h_) { h_ = 0.1; have_h_ = true; }  // PETSc's default TS time step
  const double c[4] = {0.0, kA[1][0], kA[2][0] + kA[2][1], kA[3][0] + kA[3][1] + kA[3][2]};
  int guard = 0;
  while (t_now_ < t_final_) {
    if (++guard > 100000) { PetscPrintf(comm_, "TsFsp: maximum number of steps reached.\n"); return -1; }  // TSSetMaxSteps(ts_, 100000)
    double h = std::min(h_, t_final_ - t_now_);  // TS_EXACTFINALTIME_MATCHSTEP
    if (J) {
      if (fspmat_->ComputeRHSJacobian(t_now_, J) != 0) return -1;
      njac += 1;
      // diagonal of J: the first slot of every state row (row_ptr[i]); the sink rows have none
      ierr = VecSet(jdiag_, 0.0); CHKERRQ(ierr);
      if (J->n_state_rows > 0)
        FSPCHKERRQ(fspvec_gather(jdiag_->d_data, J->val.get(), J->row_ptr.get(), J->n_state_rows, comm_ ? comm_->stream : nullptr));
    }
    bool accepted = false;
    while (!accepted) {
      const double hg = h * kGamma;
      bool         lin_failed = false;
      for (int i = 0; i < 4 && !lin_failed; ++i) {
        // stage value and right-hand side
        ierr = VecCopy(y, ystage_); CHKERRQ(ierr);
        for (int j = 0; j < i; ++j) if (kA[i][j] != 0.0) { ierr = VecAXPY(ystage_, kA[i][j], k_[j]); CHKERRQ(ierr); }
        ierr = EvaluateRHS(t_now_ + c[i] * h, ystage_, rhsv_);
        if (ierr) return -1;
        if (i == 0) { ierr = VecCopy(rhsv_, f0_); CHKERRQ(ierr); }
        ierr = VecScale(rhsv_, h); CHKERRQ(ierr);
        bool any = false;
        ierr = VecSet(err_, 0.0); CHKERRQ(ierr);
        for (int j = 0; j < i; ++j) if (kG[i][j] != 0.0) { ierr = VecAXPY(err_, kG[i][j], k_[j]); CHKERRQ(ierr); any = true; }
        if (any) {
          ierr = JacTimes(t_now_, err_, tmp_); if (ierr) return -1;
          ierr = VecAXPY(rhsv_, h, tmp_); CHKERRQ(ierr);
        }
        // linear tolerance tied to the integration tolerance (a stage error of 1 % of the step tolerance)
        ierr = SolveStage(t_now_, hg, rhsv_, k_[i], std::min(1.0e-5, 1.0e-2 * rel_tol_));
        if (ierr < 0) return -1;
        lin_failed = ierr > 0;
      }
      if (lin_failed) {  // the stage system was too hard for GMRES at this step size: reject, quarter the step
        nreject += 1;
        h *= 0.25;
        if (h < 1.0e-14 * std::max(1.0, std::fabs(t_now_))) { PetscPrintf(comm_, "TsFsp: step size underflow.\n"); return -1; }
        continue;
      }
      ierr = VecCopy(y, ynew_); CHKERRQ(ierr);
      ierr = VecSet(err_, 0.0); CHKERRQ(ierr);
      for (int i = 0; i < 4; ++i) {
        ierr = VecAXPY(ynew_, kB[i], k_[i]); CHKERRQ(ierr);
        ierr = VecAXPY(err_, kB[i] - kB2[i], k_[i]); CHKERRQ(ierr);
      }
      double enorm = 0.0;
      if (WrmsNorm(err_, y, ynew_, &enorm)) return -1;
      // TSAdaptBasic: h_new = h * clip(safety * enorm^(-1/(order_embedded + 1)), 0.1, 10)
      double fac = enorm > 0.0 ? 0.9 * std::pow(enorm, -1.0 / 3.0) : 10.0;
      fac = std::min(10.0, std::max(0.1, fac));
      if (enorm <= 1.0 || h <= 1.0e-14 * std::max(1.0, std::fabs(t_now_))) {
        accepted = true;
        h_ = h * fac;
      } else {
        nreject += 1;
        h = h * std::min(fac, 0.9);
      }
    }
    nstep += 1;
    t_now_tmp = t_now_ + h;
    // TSCheckFspError (:131-197): stop condition on the new solution; on excess halve towards t_now_ on the dense output,
    // re-checking, up to ten times
    if (stop_check_ != nullptr) {
      PetscReal excess = 0.0;
      ierr = stop_check_(t_now_tmp, ynew_, excess, stop_data_);
      PACMENSLCHKERRQ(ierr);
      if (excess > 0.0) {
        fsp_stop_ = 1;
        ierr = EvaluateRHS(t_now_tmp, ynew_, f1_);
        if (ierr) return -1;
        PetscReal excess2 = 1.0, t_try = t_now_tmp;
        int       ntrial = 0;
        while (ntrial < 10 && excess2 > 0.0) {
          t_try = t_now_ + 0.5 * (t_try - t_now_);
          ierr = Interpolate(t_now_, t_now_tmp, y, ynew_, t_try, tmp_); CHKERRQ(ierr);
          ierr = stop_check_(t_try, tmp_, excess2, stop_data_);
          PACMENSLCHKERRQ(ierr);
          ntrial += 1;
        }
        if (ntrial >= 10 && excess2 > 0.0) {
          // no admissible point found: roll back to the last accepted time (y is untouched)
        } else {
          ierr = VecCopy(tmp_, y); CHKERRQ(ierr);
          t_now_ = t_try;
        }
        break;
      }
    }
    ierr = VecCopy(ynew_, y); CHKERRQ(ierr);
    if (print_intermediate)
      PetscPrintf(comm_, "t_now_ = %.2e stepsize = %.2e nstep = %d njac = %d \n", t_now_tmp, t_now_tmp - t_now_, nstep, njac);
    t_now_ = t_now_tmp;
    if (logging_enabled && (size_t) perf_info.n_step < perf_info.model_time.size()) {
      perf_info.model_time[perf_info.n_step] = t_now_;
      VecGetSize(*solution_, &perf_info.n_eqs[size_t(perf_info.n_step)]);
      PetscTime(&perf_info.cpu_time[perf_info.n_step]);
      perf_info.n_step += 1;
    }
  }
  ierr = VecCopy(solution_tmp_, *solution_); CHKERRQ(ierr);
  return fsp_stop_;
}

int TsFsp::FreeWorkspace() {
  if (J) { MatDestroy(&J); J = nullptr; }
  for (Vec *v : {&k_[0], &k_[1], &k_[2], &k_[3], &ystage_, &rhsv_, &tmp_, &ynew_, &err_, &f0_, &f1_, &jdiag_, &pc_, &solution_tmp_})
    if (*v) VecDestroy(v);
  for (auto &v : V_) if (v) VecDestroy(&v);
  V_.clear();
  have_h_ = false;
  return OdeSolverBase::FreeWorkspace();
}

}  // namespace pacmensl

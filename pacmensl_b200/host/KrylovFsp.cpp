#include "KrylovFsp.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace pacmensl {

KrylovFsp::KrylovFsp(MPI_Comm comm) : OdeSolverBase(comm) {}

// src/OdeSolver/KrylovFsp.cpp:29-99
PetscInt KrylovFsp::Solve() {
  if (solution_ == nullptr) return -1;
  if (rhs_ == nullptr) return -1;

  PacmenslErrorCode ierr;
  PetscInt          petsc_err;

  petsc_err = VecCopy(*solution_, solution_tmp_);
  CHKERRQ(petsc_err);
  t_now_tmp_ = t_now_;
  DestroyGraphs_();  // the operator may have been regenerated since the last Solve(): captured pointers are stale

  int       stop = 0;
  PetscReal error_excess = 0.0;
  while (t_now_ < t_final_) {
    krylov_stat_ = AdvanceOneStep(solution_tmp_);
    PACMENSLCHKERRQ(krylov_stat_);

    if (stop_check_ != nullptr) {
      ierr = stop_check_(t_now_tmp_, solution_tmp_, error_excess, stop_data_);
      CHKERRQ(ierr);

      // Reference behaviour (KrylovFsp.cpp:59-78): error_excess is not re-evaluated inside this loop, so an
      // exceeded tolerance always halves ten times and ends with the roll-back to t_now_.
      PetscReal t_step_tmp = t_now_tmp_ - t_now_;
      PetscInt  nrej = 0;
      while (error_excess > 0.0 && nrej < 10) {
        stop = 1;
        nrej += 1;
        if (nrej >= 10) t_step_tmp = 0.0;
        else t_step_tmp = 0.5 * t_step_tmp;
        if (nrej >= 10) {  // only the last interpolation determines the state: skip the nine discarded ones
          krylov_stat_ = GetDky(t_now_ + t_step_tmp, 0, solution_tmp_);
          PACMENSLCHKERRQ(krylov_stat_);
        }
        t_now_tmp_ = t_now_ + t_step_tmp;
      }
      if (stop) break;
    }
    t_now_ = t_now_tmp_;
    if (print_intermediate) PetscPrintf(comm_, "t_now_ = %.2e \n", t_now_);
    if (logging_enabled) {
      if ((size_t) perf_info.n_step < perf_info.model_time.size()) {
        perf_info.model_time[perf_info.n_step] = t_now_;
        petsc_err = VecGetSize(*solution_, &perf_info.n_eqs[size_t(perf_info.n_step)]);
        CHKERRQ(petsc_err);
        petsc_err = PetscTime(&perf_info.cpu_time[perf_info.n_step]);
        CHKERRQ(petsc_err);
        perf_info.n_step += 1;
      }
    }
  }
  petsc_err = VecCopy(solution_tmp_, *solution_);
  CHKERRQ(petsc_err);
  return stop;
}

// src/OdeSolver/KrylovFsp.cpp:101-262
int KrylovFsp::AdvanceOneStep(const Vec &v) {
  PetscErrorCode ierr;
  PetscBool      happy_breakdown, success_step, bsize_changed;
  PetscReal      s, xm, err_loc, omega = 0.0, omega_old = 0.0, kappa, order, t_step_old = 0.0, t_step_suggest;
  PetscInt       ireject, m_old = 0, m_start, m_suggest;
  PetscReal      cost_tchange, cost_mchange;

  err_loc = 0.0;
  success_step = PETSC_FALSE;
  bsize_changed = PETSC_FALSE;
  ireject = 0;
  m_start = 0;
  kappa = 2.0;
  order = double(m_) / 4;

  while (!success_step && ireject <= max_reject_) {
    m_ = std::min(m_max_, std::max(m_min_, m_next_));
    ierr = GenerateBasis(v, m_start, &happy_breakdown);
    PACMENSLCHKERRQ(ierr);

    if (!first_step_initialized_) {  // :133-144
      PetscReal anorm;
      xm = 1.0 / double(m_);
      num_rhs_evals_ += 1;
      ierr = rhs_(0.0, v, av);
      PACMENSLCHKERRQ(ierr);
      ierr = VecNorm(av, NORM_2, &avnorm);
      CHKERRQ(ierr);
      anorm = avnorm / beta;
      double fact = pow((m_ + 1) / exp(1.0), m_ + 1) * sqrt(2 * (3.1416) * (m_ + 1));
      t_step_next_ = (1.0 / anorm) * pow((fact * abs_tol_) / (4.0 * beta * anorm), xm);
      first_step_initialized_ = true;
    }

    t_step_ = std::min(t_final_ - t_now_tmp_, t_step_next_);

    if (k1 != 0) {  // :149-155
      Hm(m_ + 1, m_) = 1.0;
      num_rhs_evals_ += 1;
      ierr = rhs_(0.0, Vm[m_], av);
      PACMENSLCHKERRQ(ierr);
      ierr = VecNorm(av, NORM_2, &avnorm);
      CHKERRQ(ierr);
    }

    mx = mb + k1;
    F = arma::expmat(t_step_ * Hm);  // :159 (full (m_max+2)^2 matrix, host)
    if (k1 == 0) {
      err_loc = btol_;
      break;
    } else {
      double phi1 = std::abs(beta * F(m_, 0));
      double phi2 = std::abs(beta * F(m_ + 1, 0) * avnorm);
      if (phi1 > phi2 * 10.0) err_loc = phi2;
      else if (phi1 > phi2) err_loc = (phi1 * phi2) / (phi1 - phi2);
      else err_loc = phi1;
    }

    if (!std::isfinite(err_loc) || !std::isfinite(err_loc / (abs_tol_ * t_step_))) {
      // exp(tau H) overflowed: the trial step is far too long.  The reference's controller cannot recover from a
      // non-finite omega (pow/log of inf give a NaN step suggestion and an undefined dimension suggestion, and the
      // same step is retried until max_reject_: observed on hog1p at t = 145); shrink the step by the controller's
      // own lower bound (factor 0.2, KrylovFsp.cpp:196) and keep the basis.
      if (ireject == max_reject_) {
        PetscPrintf(comm_, "KrylovFsp: maximum number of failed steps reached\n");
        return -1;
      }
      ireject++;
      t_step_old = t_step_;
      m_old = m_;
      m_start = m_;
      m_next_ = m_;
      t_step_next_ = 0.2 * t_step_;
      bsize_changed = PETSC_FALSE;
      continue;
    }
    omega_old = omega;
    omega = err_loc / (abs_tol_ * t_step_);  // :182 -- only atol enters

    if (bsize_changed && ireject > 0) {
      kappa = std::max(1.1E0, std::pow(omega / omega_old, 1.0 / (m_old - m_)));
    } else if (ireject > 0) {
      order = std::max(1.0, std::log(omega / omega_old) / std::log(t_step_ / t_step_old));
    }

    t_step_suggest = gamma_ * t_step_ * pow(omega, -1.0 / order);
    s = pow(10.0, floor(log10(t_step_suggest)) - 1);
    t_step_suggest = ceil(t_step_suggest / s) * s;
    t_step_suggest = std::min(5.0 * t_step_, std::max(0.2 * t_step_, t_step_suggest));
    t_step_suggest = std::min(t_final_ - t_now_tmp_, t_step_suggest);

    m_suggest = m_ + (PetscInt) std::ceil(std::log(omega / gamma_) / std::log(kappa));
    m_suggest = std::max(3 * m_ / 4, std::min(4 * m_ / 3 + 1, m_suggest));
    m_suggest = std::max(m_min_, std::min(m_max_, m_suggest));

    ierr = EstimateCost_(t_step_suggest, m_, &cost_tchange);
    CHKERRQ(ierr);
    ierr = EstimateCost_(t_step_, m_suggest, &cost_mchange);
    CHKERRQ(ierr);

    if (std::ceil((t_final_ - t_now_tmp_) / t_step_suggest) * cost_tchange <=
            std::ceil((t_final_ - t_now_tmp_) / t_step_) * cost_mchange ||
        m_suggest == m_) {
      t_step_next_ = t_step_suggest;
      m_next_ = m_;
      bsize_changed = PETSC_FALSE;
    } else {
      t_step_next_ = t_step_;
      m_next_ = m_suggest;
      bsize_changed = PETSC_TRUE;
    }

    if (omega <= delta_) {
      success_step = PETSC_TRUE;
    } else {
      if (bsize_changed) Hm(m_ + 1, m_) = 0.0;
      if (print_intermediate)
        PetscPrintf(comm_, "t_step = %.2e m = %d t_step_next = %.2e err_loc = %.2e \n", t_step_, m_, t_step_next_, err_loc);
      static const bool dbg = [] { const char *e = std::getenv("FSP_KRYLOV_DEBUG"); return e && e[0] == '1'; }();
      if (dbg && (ireject < 5 || ireject % 1000 == 0))
        printf("[krylov] reject %d: t=%.6e tau=%.3e m=%d mb=%d k1=%d beta=%.6e avnorm=%.6e err_loc=%.3e omega=%.3e H00=%.6e H10=%.6e H(m,m-1)=%.6e F(m,0)=%.3e F(m+1,0)=%.3e\n",
               (int) ireject, t_now_tmp_, t_step_, m_, mb, (int) k1, beta, avnorm, err_loc, omega, Hm(0, 0), Hm(1, 0), Hm(m_, m_ - 1), F(m_, 0), F(m_ + 1, 0));
      if (ireject == max_reject_) {
        PetscPrintf(comm_, "KrylovFsp: maximum number of failed steps reached\n");
        return -1;
      }
      ireject++;
      t_step_old = t_step_;
      m_old = m_;
      m_start = m_old;
    }
  }

  mx = mb + std::max(0, (int) k1 - 1);  // :244-252
  std::vector<double> F0((size_t) mx);
  for (int ii{0}; ii < mx; ++ii) F0[ii] = beta * F(ii, 0);
  {
    // v = sum_k F0[k] V_k in one fused pass (beta = 0 overwrites v, replacing VecScale(v,0) + VecMAXPY)
    std::vector<const double *> ptrs((size_t) mx);
    for (int ii = 0; ii < mx; ++ii) ptrs[ii] = Vm[ii]->d_data;
    double beta_y = 0.0;
    for (int k0 = 0; k0 < mx; k0 += 64) {
      int mm = std::min(64, mx - k0);
      FSPCHKERRQ(fspvec_maxpy(v->d_data, beta_y, mm, F0.data() + k0, ptrs.data() + k0, v->n_local, comm_->stream));
      beta_y = 1.0;
    }
  }
  t_now_tmp_ = t_now_tmp_ + t_step_;

  if (print_intermediate)
    PetscPrintf(comm_, "t_step = %.2e m = %d t_step_next = %.2e err_loc = %.2e \n", t_step_, m_, t_step_next_, err_loc);
  return 0;
}

// src/OdeSolver/KrylovFsp.cpp:264-322 -- incomplete orthogonalisation procedure (modified Gram-Schmidt over the
// last q_iop vectors).  Device pipeline per basis vector j (all asynchronous, coefficients stay on the device):
//   w = A V_j ; h_0 = <w, V_i0> ; [w -= h_k V_ik ; h_{k+1} = <w, V_ik+1>]... ; w -= h_last V_j ; s^2 = <w, w> ; w /= s
// The column loop of GenerateBasis (src/OdeSolver/KrylovFsp.cpp:294-318) as an asynchronous device pipeline.
int KrylovFsp::BasisColumns_(int m_start) {
  int ierr, istart = 0;
  void      *stream = comm_->stream;
  const long n = Vm[0]->n_local;
  const int  stride = m_max_ + 2;  // coefficients of column j live at hdev_[j*stride ...]
  const bool multi = comm_size_ > 1;

  for (int j{m_start}; j < m_; j++) {
    num_rhs_evals_ += 1;
    if (q_iop > 0) istart = (j - q_iop + 1 >= 0) ? j - q_iop + 1 : 0;
    double *hcol = hdev_.get() + (size_t) j * stride;
    double *w = Vm[j + 1]->d_data;
    // The fused form saves 8 of 184 bytes per row and no launch here (the partial reduction replaces the dot kernel):
    // measured neutral, so it is opt-in for this solver (FSP_KRYLOV_FUSED=1); the BDF/GMRES loop is where it pays.
    static const bool krylov_fused = [] { const char *e = std::getenv("FSP_KRYLOV_FUSED"); return e && e[0] == '1'; }();
    static const bool orth_on = [] { const char *e = std::getenv("FSP_KRYLOV_ORTH"); return !(e && e[0] == '0'); }();
    const int nvec = j - istart + 1;
    if (fused_rhs_ && krylov_fused) {
      // w = A V_j and the first coefficient <w, V_istart> in one pass over w
      fspmat_epilogue ep{};
      ep.alpha = 1.0; ep.beta = 0.0; ep.scale_dev = nullptr; ep.n_dots = 1;
      ep.dot_vec_dev[0] = Vm[istart]->d_data; ep.dot_vec_dev[1] = nullptr; ep.dot_out_dev = hcol + 0;
      ierr = fused_rhs_(0.0, Vm[j], Vm[j + 1], ep);
      PACMENSLCHKERRQ(ierr);
    } else {
      ierr = rhs_(0.0, Vm[j], Vm[j + 1]);
      PACMENSLCHKERRQ(ierr);
      if (!multi && orth_on && nvec <= 2) {
        // Single rank, vector small enough to live in registers: the whole orthogonalisation of this column is ONE
        // cooperative launch -- w read and written once, the three inner products separated by grid-wide barriers
        // instead of kernel boundaries.  This is the launch-bound regime of the adaptive examples (hog1p + KrylovFsp:
        // 53 000 columns on <= 7e5 states).  rc == 1: not applicable (vector too long / no cooperative launch).
        int rc = fspvec_iop_orth(w, nvec, Vm[istart]->d_data, Vm[j]->d_data, hcol, n, stream);
        if (rc < 0) FSPCHKERRQ(rc);
        if (rc == 0) continue;
      }
      // first coefficient: plain dot
      FSPCHKERRQ(fspvec_dot(hcol + 0, w, Vm[istart]->d_data, n, stream));
    }
    if (multi) FSPCHKERRQ(fspcomm_allreduce_sum(comm_->nccl, hcol + 0, 1, stream));
    int c = 0;
    for (int i = istart; i <= j; ++i, ++c) {
      // w -= h_c V_i, fused with the next inner product (next V, or <w,w> after the last one)
      const double *u = (i < j) ? Vm[i + 1]->d_data : nullptr;
      FSPCHKERRQ(fspvec_axpy_dot(w, hcol + c, 1.0, Vm[i]->d_data, u, hcol + c + 1, n, stream));
      if (multi) FSPCHKERRQ(fspcomm_allreduce_sum(comm_->nccl, hcol + c + 1, 1, stream));
    }
    // hcol[c] now holds ||w||^2
    FSPCHKERRQ(fspvec_scale_rsqrt(w, hcol + c, n, stream));
  }

  return 0;
}

bool KrylovFsp::GraphsUsable_() {
  static const bool env_on = [] { const char *e = std::getenv("FSP_KRYLOV_GRAPH"); return !(e && e[0] == '0'); }();
  return env_on && !graphs_disabled_ && comm_size_ == 1;
}

void KrylovFsp::DestroyGraphs_() {
  for (auto &kv : graphs_) fsp_graph_destroy(kv.second);
  graphs_.clear();
  graph_seen_.clear();
}

int KrylovFsp::EnsureBasis_(int count) {
  count = std::min(count, (int) Vm.size());
  int have = 0;
  while (have < (int) Vm.size() && Vm[have]) ++have;
  if (have >= count) return 0;
  // the missing vectors come from ONE device block
  Vec *fresh = nullptr;
  int  ierr = VecDuplicateVecsUninitialized(*solution_, count - have, &fresh);
  CHKERRQ(ierr);
  for (int i = have; i < count; ++i) Vm[i] = fresh[i - have];
  delete[] fresh;  // the array only; the vectors live on in Vm and keep their block alive
  return 0;
}

int KrylovFsp::GenerateBasis(const Vec &v, int m_start, PetscBool *happy_breakdown) {
  int ierr;

  *happy_breakdown = PETSC_FALSE;
  ierr = EnsureBasis_(m_ + 1);
  CHKERRQ(ierr);
  if (m_start >= m_) return 0;

  k1 = 2;
  mb = m_;

  ierr = VecNorm(v, NORM_2, &beta);
  CHKERRQ(ierr);
  ierr = VecCopy(v, Vm[0]);
  CHKERRQ(ierr);
  ierr = VecScale(Vm[0], 1.0 / beta);
  CHKERRQ(ierr);

  if (m_start == 0) Hm.zeros();

  // Launch-bound regime (small state sets: every kernel of the column loop lasts a few microseconds): the whole loop
  // for a given (m_start, m) is captured once into a CUDA graph and replayed with one launch.  rhs_ is evaluated at
  // t = 0 for every column (KrylovFsp.cpp:296), so the captured coefficients stay valid for the whole Solve().
  const long key = (long) m_start * 4096 + m_;
  bool       done = false;
  if (GraphsUsable_()) {
    auto it = graphs_.find(key);
    if (it != graphs_.end()) {
      if (fsp_graph_launch(it->second, comm_->stream) == 0) {
        num_rhs_evals_ += m_ - m_start;
        done = true;
      } else {
        DestroyGraphs_();
        graphs_disabled_ = true;
      }
    } else if (++graph_seen_[key] >= 2) {  // second time this shape is needed: worth capturing
      if (!capture_stream_ && fsp_stream_create(&capture_stream_)) graphs_disabled_ = true;
      if (!graphs_disabled_) {
        void *saved = comm_->stream;
        comm_->stream = capture_stream_;  // everything this rank submits now goes to the capturing stream
        long rhs_before = num_rhs_evals_;
        const long launches_before = fsp_launch_count();
        int  cerr = fsp_graph_begin_capture(capture_stream_);
        if (cerr == 0) cerr = BasisColumns_(m_start);
        comm_->stream = saved;
        fsp_graph_t g = nullptr;
        long        nk = 0;
        const long  expect = fsp_launch_count() - launches_before;  // every launch of the column loop must be in the graph
        if (cerr != 0) {
          fsp_graph_abort_capture(capture_stream_);
        } else if (fsp_graph_end_capture(capture_stream_, &g) == 0) {
          fsp_graph_num_kernels(g, &nk);
          if (nk >= expect && fsp_graph_launch(g, comm_->stream) == 0) {
            graphs_[key] = g;
            done = true;
          } else {
            fsp_graph_destroy(g);  // part of the work bypassed the capturing stream: not capturable
          }
        }
        if (!done) {  // fall back to plain launches for the rest of this solver's life (a genuine rhs_ error repeats below)
          num_rhs_evals_ = rhs_before;
          graphs_disabled_ = true;
        }
      }
    }
  }
  if (!done) {
    ierr = BasisColumns_(m_start);
    PACMENSLCHKERRQ(ierr);
  }
  void     *stream = comm_->stream;
  const int stride = m_max_ + 2;

  // one transfer of all coefficients, then the (deferred) happy-breakdown test
  const int ncols = m_ - m_start;
  hhost_.resize((size_t) ncols * stride);
  FSPCHKERRQ(fsp_memcpy_d2h(hhost_.data(), hdev_.get() + (size_t) m_start * stride, sizeof(double) * ncols * stride, stream));
  if (comm_ && comm_->nccl) FSPCHKERRQ(fspcomm_check(comm_->nccl));  // a timed-out peer wait poisoned the Hessenberg entries
  for (int j = m_start; j < m_; ++j) {
    const double *hcol = hhost_.data() + (size_t) (j - m_start) * stride;
    int           is = (q_iop > 0) ? ((j - q_iop + 1 >= 0) ? j - q_iop + 1 : 0) : 0;
    int           c = 0;
    for (int i = is; i <= j; ++i, ++c) Hm(i, j) = hcol[c];
    double s = std::sqrt(hcol[c]);
    Hm(j + 1, j) = s;
    if (!(s >= btol_)) {  // also catches NaN
      k1 = 0;
      mb = j + 1;
      *happy_breakdown = PETSC_TRUE;
      // the reference stops here (:311-317); discard whatever was computed past the breakdown
      for (int jj = j + 1; jj < (int) Hm.n_cols; ++jj)
        for (int ii = 0; ii < (int) Hm.n_rows; ++ii) Hm(ii, jj) = 0.0;
      break;
    }
  }
  return 0;
}

// src/OdeSolver/KrylovFsp.cpp:324-362
int KrylovFsp::SetUpWorkSpace() {
  if (!solution_) {
    PetscPrintf(comm_, "KrylovFsp error: starting solution vector is null.\n");
    return -1;
  }
  int ierr;
  // The reference allocates all m_max+1 basis vectors up front (:334-340).  Here the basis vectors are created when
  // the adaptive dimension m first needs them (EnsureBasis_): with the default range [25, 60] a solve that never
  // raises m holds 26 instead of 61 vectors -- at 1e8 states that is 28 GB less HBM and no allocation/zero-fill of
  // memory that is never touched.  Workspace vectors are always written before they are read: no zero-fill.
  Vm.assign((size_t) m_max_ + 1, nullptr);
  ierr = VecDuplicateUninitialized(*solution_, &av);
  CHKERRQ(ierr);
  ierr = VecDuplicateUninitialized(*solution_, &solution_tmp_);
  CHKERRQ(ierr);

  first_step_initialized_ = false;
  m_next_ = m_min_;
  Hm = arma::zeros(m_max_ + 2, m_max_ + 2);
  if (hdev_.resize((size_t) (m_max_ + 1) * (m_max_ + 2))) return -1;

  if (fspmat_) {
    ierr = fspmat_->GetLocalMVFlops(&rhs_cost_loc_);
    CHKERRQ(ierr);
  }
  return 0;
}

// src/OdeSolver/KrylovFsp.cpp:364-411
int KrylovFsp::GetDky(PetscReal t, int deg, Vec p_vec) {
  if (t < t_now_ || t > t_now_tmp_) {
    PetscPrintf(comm_, "KrylovFsp::GetDky error: requested timepoint does not belong to the current time subinterval.\n");
    return -1;
  }
  deg = (deg < 0) ? 0 : deg;
  F = arma::expmat((t - t_now_) * Hm);
  mx = mb + std::max(0, (int) k1 - 1);
  std::vector<double>         F0((size_t) mx);
  std::vector<const double *> ptrs((size_t) mx);
  for (int ii{0}; ii < mx; ++ii) {
    F0[ii] = beta * F(ii, 0);
    ptrs[ii] = Vm[ii]->d_data;
  }
  double beta_y = 0.0;
  for (int k0 = 0; k0 < mx; k0 += 64) {
    int mm = std::min(64, mx - k0);
    FSPCHKERRQ(fspvec_maxpy(p_vec->d_data, beta_y, mm, F0.data() + k0, ptrs.data() + k0, p_vec->n_local, comm_->stream));
    beta_y = 1.0;
  }
  if (deg > 0) {
    Vec vtmp;
    PetscErrorCode petsc_err = VecDuplicate(p_vec, &vtmp);
    CHKERRQ(petsc_err);
    for (int i{1}; i <= deg; ++i) {
      num_rhs_evals_ += 1;
      rhs_(0.0, p_vec, vtmp);
      VecSwap(p_vec, vtmp);
    }
    VecDestroy(&vtmp);
  }
  return 0;
}

KrylovFsp::~KrylovFsp() {
  FreeWorkspace();
  if (capture_stream_) fsp_stream_destroy(capture_stream_);
}

int KrylovFsp::FreeWorkspace() {
  OdeSolverBase::FreeWorkspace();
  DestroyGraphs_();
  for (size_t i{0}; i < Vm.size(); ++i)
    if (Vm[i]) VecDestroy(&Vm[i]);
  Vm.clear();
  if (av != nullptr) VecDestroy(&av);
  if (solution_tmp_ != nullptr) VecDestroy(&solution_tmp_);
  return 0;
}

int KrylovFsp::SetUp() {
  OdeSolverBase::SetUp();
  // drop a workspace left from a previous SetUp (the driver calls FreeWorkspace, direct users may not)
  Vec *keep = solution_;
  if (!Vm.empty()) { FreeWorkspace(); solution_ = keep; }
  return SetUpWorkSpace();
}

PacmenslErrorCode KrylovFsp::SetOrthLength(int q) {
  q_iop = q;
  return 0;
}

// src/OdeSolver/KrylovFsp.cpp:457-478
int KrylovFsp::EstimateCost_(PetscReal tau_new, PetscInt m_new, PetscReal *cost) {
  PetscReal hnorm = arma::norm(Hm, "inf");
  int       ns = (int) std::ceil(hnorm * tau_new);
  PetscReal cost_local;
  PetscInt  n_loc;
  int       ierr = VecGetLocalSize(*solution_, &n_loc);
  CHKERRQ(ierr);
  if (q_iop > 0) {
    cost_local = PetscReal(m_new + 1) * rhs_cost_loc_ +
                 PetscReal(4 * q_iop * m_new + 5 * m_new + 2 * q_iop - 2 * q_iop * q_iop + 7) * n_loc +
                 2.0 * std::ceil(25.0 / 3.0 + ns) * PetscReal((m_new + 2) * (m_new + 2) * (m_new + 2));
  } else {
    cost_local = PetscReal(m_new + 1) * rhs_cost_loc_ +
                 PetscReal(4 * m_new * m_new + 5 * m_new + 2 * m_new - 2 * m_new * m_new + 7) * n_loc +
                 2.0 * std::ceil(25.0 / 3.0 + ns) * PetscReal((m_new + 2) * (m_new + 2) * (m_new + 2));
  }
  *cost = cost_local;
  return pacmensl_allreduce_max(comm_, cost, 1);
}

PacmenslErrorCode KrylovFsp::SetKrylovDimRange(int m_min, int m_max) {
  m_min_ = m_min;
  m_max_ = m_max;
  m_next_ = m_min;
  return 0;
}

}  // namespace pacmensl

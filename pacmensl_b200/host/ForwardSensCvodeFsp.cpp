#include "ForwardSensCvodeFsp.h"

namespace pacmensl {

// src/SensFsp/ForwardSensCvodeFsp.cpp:125-232
PacmenslErrorCode ForwardSensCvodeFsp::SetUp() {
  if (solution_ == nullptr) return -1;
  if (rhs_ == nullptr) return -1;
  if (srhs_ == nullptr) return -1;
  PetscInt petsc_err;
  petsc_err = VecDuplicate(*solution_, &solution_work_); CHKERRQ(petsc_err);
  petsc_err = VecCopy(*solution_, solution_work_); CHKERRQ(petsc_err);
  petsc_err = VecDuplicate(*solution_, &tmp_); CHKERRQ(petsc_err);
  sens_work_.resize(num_parameters_);
  for (int i{0}; i < num_parameters_; ++i) {
    petsc_err = VecDuplicate(*solution_, &sens_work_[i]); CHKERRQ(petsc_err);
    petsc_err = VecCopy(*sens_vecs_[i], sens_work_[i]); CHKERRQ(petsc_err);
  }
  t_now_tmp_ = t_now_;

  core_.reset(new BdfCore(comm_));
  core_->SetTolerances(rel_tol, abs_tol);
  core_->SetMaxConvFails(10000);
  core_->SetMaxNonlinIters(10000);
  core_->SetMaxKrylov(50);  // SUNSPGMR(y, PREC_NONE, 50)
  auto f = [this](double t, Vec y, Vec ydot) { return EvaluateRHS(t, y, ydot); };
  cvode_stat = core_->Init(t_now_tmp_, solution_work_, f, f, t_final_);
  if (cvode_stat < 0) return -1;
  // sdot = A s + (dA/dtheta_is) y   (:96-115)
  auto fs = [this](int is, double t, Vec y, Vec, Vec s, Vec sdot) {
    int ierr = EvaluateRHS(t, s, sdot);
    PACMENSLCHKERRQ(ierr);
    ierr = EvaluateSensRHS(is, t, y, tmp_);
    PACMENSLCHKERRQ(ierr);
    ierr = VecAXPY(sdot, 1.0, tmp_);
    PACMENSLCHKERRQ(ierr);
    return 0;
  };
  cvode_stat = core_->InitSens(num_parameters_, sens_work_.data(), fs, /*errcon=*/true);  // CV_STAGGERED1 + SensErrCon
  if (cvode_stat < 0) return -1;
  set_up_ = true;
  return 0;
}

// src/SensFsp/ForwardSensCvodeFsp.cpp:234-287
PetscInt ForwardSensCvodeFsp::Solve() {
  if (!set_up_) {
    PacmenslErrorCode e = SetUp();
    if (e) return -1;
  }
  PetscErrorCode petsc_err;
  int            stop = 0;
  while (t_now_ < t_final_) {
    cvode_stat = core_->Step(&t_now_tmp_, solution_work_, nullptr);
    if (cvode_stat < 0) {
      printf("\nBDF sensitivity integrator error: step failed with flag = %d\n\n", cvode_stat);
      return -1;
    }
    if (t_now_tmp_ > t_final_) {
      cvode_stat = core_->GetDky(t_final_, solution_work_);
      if (cvode_stat < 0) return -1;
      t_now_tmp_ = t_final_;
    }
    for (int i = 0; i < num_parameters_; ++i) {
      cvode_stat = core_->GetSensDky(t_now_tmp_, i, sens_work_[i]);
      if (cvode_stat < 0) return -1;
    }
    if (stop_check_ != nullptr)
      stop = stop_check_(t_now_tmp_, solution_work_, num_parameters_, sens_work_.data(), stop_data_);
    if (stop == 1) {
      cvode_stat = core_->GetDky(t_now_, solution_work_);
      if (cvode_stat < 0) return -1;
      for (int i = 0; i < num_parameters_; ++i) {
        cvode_stat = core_->GetSensDky(t_now_, i, sens_work_[i]);
        if (cvode_stat < 0) return -1;
      }
      break;
    } else {
      t_now_ = t_now_tmp_;
      if (print_intermediate) PetscPrintf(comm_, "t_now_ = %.2e \n", t_now_);
    }
  }
  petsc_err = VecCopy(solution_work_, *solution_); CHKERRQ(petsc_err);
  for (int i{0}; i < num_parameters_; ++i) {
    petsc_err = VecCopy(sens_work_[i], *sens_vecs_[i]); CHKERRQ(petsc_err);
  }
  return stop;
}

PacmenslErrorCode ForwardSensCvodeFsp::FreeWorkspace() {
  core_.reset();
  if (solution_work_) VecDestroy(&solution_work_);
  if (tmp_) VecDestroy(&tmp_);
  for (auto &v : sens_work_) VecDestroy(&v);
  sens_work_.clear();
  num_parameters_ = 0;
  set_up_ = false;
  PetscReal keep_t = t_now_;
  PacmenslErrorCode ierr = ForwardSensSolverBase::FreeWorkspace();
  (void) keep_t;
  return ierr;
}

ForwardSensCvodeFsp::~ForwardSensCvodeFsp() { FreeWorkspace(); }

}  // namespace pacmensl

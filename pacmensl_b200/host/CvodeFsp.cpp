#include "CvodeFsp.h"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace pacmensl {

CvodeFsp::CvodeFsp(MPI_Comm _comm, int lmm) : OdeSolverBase(_comm) { lmm_ = lmm; }

// src/OdeSolver/CvodeFsp.cpp:137-200
PacmenslErrorCode CvodeFsp::SetUp() {
  if (solution_ == nullptr) return -1;
  if (rhs_ == nullptr) return -1;
  if (lmm_ != CV_BDF) {
    PetscPrintf(comm_, "CvodeFsp: only the BDF method (CV_BDF) is provided.\n");
    return -1;
  }
  if (solution_work_) VecDestroy(&solution_work_);
  PetscInt petsc_err = VecDuplicate(*solution_, &solution_work_);
  CHKERRQ(petsc_err);
  petsc_err = VecCopy(*solution_, solution_work_);
  CHKERRQ(petsc_err);
  t_now_tmp = t_now_;

  auto f = [this](double t, Vec y, Vec ydot) { return EvaluateRHS(t, y, ydot); };  // J v == A(t) v (linear ODE)
  const bool resume = warm_restart_ && carry_ && core_ && core_->LocalSize() == (*solution_)->n_local &&
                      core_->CurrentTime() == t_now_;
  carry_ = false;
  if (!resume) {
    core_.reset(new BdfCore(comm_));
    core_->SetTolerances(rel_tol_, abs_tol_);
    core_->SetMaxConvFails(10000);     // CVodeSetMaxConvFails(cvode_mem, 10000)
    core_->SetMaxNonlinIters(10000);   // CVodeSetMaxNonlinIters(cvode_mem, 10000)
    core_->SetMaxKrylov(100);          // SUNLinSol_SPGMR(y, PREC_NONE, 100)
    if (fused_rhs_) core_->SetFusedJtv([this](double t, Vec v, Vec out, const fspmat_epilogue &ep) { num_rhs_evals_ += 1; return fused_rhs_(t, v, out, ep); });
    cvode_stat = core_->Init(t_now_tmp, solution_work_, f, f, t_final_);
    if (cvode_stat < 0) {
      printf("\nBDF integrator error: initialisation failed with flag = %d\n\n", cvode_stat);
      return -1;
    }
  }
  // (resume: the integrator kept its history through BdfCore::Expand and simply goes on from t_now_)
  if (warm_restart_ && stop_check_ != nullptr) {
    // The stop condition is evaluated on the converged but not yet committed step, so a stop leaves the integrator at
    // the last committed time with its history intact (the reference checks after the step and rolls back through
    // CVodeGetDky, then throws the integrator away: src/OdeSolver/CvodeFsp.cpp:50-59, FspSolverMultiSinks.cpp:92-108).
    core_->SetAcceptHook([this](double t, Vec y, bool *reject) {
      *reject = false;
      hook_checked_ = false;
      if (t > t_final_) return 0;  // the last step is interpolated back to t_final first: checked after the step
      int ierr = stop_check_(t, y, hook_excess_, stop_data_);
      if (ierr) return ierr;
      hook_checked_ = true;
      *reject = hook_excess_ > 0.0;
      return 0;
    });
  } else {
    core_->SetAcceptHook(nullptr);
  }
  return 0;
}

// src/OdeSolver/CvodeFsp.cpp:34-78
PetscInt CvodeFsp::Solve() {
  if (solution_ == nullptr || rhs_ == nullptr || !core_) return -1;
  PacmenslErrorCode ierr;
  PetscErrorCode    petsc_err;
  int               stop = 0;
  PetscReal         error_excess = 0.0;
  while (t_now_ < t_final_) {
    hook_checked_ = false;
    cvode_stat = core_->Step(&t_now_tmp, solution_work_);
    if (cvode_stat == BdfCore::STOPPED) {
      // the accept hook saw the sinks exceed the tolerance: solution_work_ is the solution at t_now_ and the
      // integrator can be carried over to the expanded state space (ExpandState)
      stop = 1;
      carry_ = true;
      break;
    }
    if (cvode_stat < 0) {
      int rank;
      MPI_Comm_rank(MPI_COMM_WORLD, &rank);
      printf("\nBDF integrator error: step failed on rank %d with flag = %d\n\n", rank, cvode_stat);
      return -1;
    }
    // Interpolate the solution if the last step went over the prescribed final time
    if (t_now_tmp > t_final_) {
      cvode_stat = core_->GetDky(t_final_, solution_work_);
      if (cvode_stat < 0) return -1;
      t_now_tmp = t_final_;
    }
    if (stop_check_ != nullptr && !hook_checked_) {
      ierr = stop_check_(t_now_tmp, solution_work_, error_excess, stop_data_);
      PACMENSLCHKERRQ(ierr);
      if (error_excess > 0.0) {
        stop = 1;
        cvode_stat = core_->GetDky(t_now_, solution_work_);  // roll back to the last accepted time
        if (cvode_stat < 0) return -1;
        break;
      }
    }
    t_now_ = t_now_tmp;
    if (print_intermediate) PetscPrintf(comm_, "t_now_ = %.2e \n", t_now_);
    if (logging_enabled && (size_t) perf_info.n_step < perf_info.model_time.size()) {
      perf_info.model_time[perf_info.n_step] = t_now_;
      petsc_err = VecGetSize(*solution_, &perf_info.n_eqs[size_t(perf_info.n_step)]);
      CHKERRQ(petsc_err);
      petsc_err = PetscTime(&perf_info.cpu_time[perf_info.n_step]);
      CHKERRQ(petsc_err);
      perf_info.n_step += 1;
    }
  }
  petsc_err = VecCopy(solution_work_, *solution_);
  CHKERRQ(petsc_err);
  static const bool bdf_trace = [] { const char *e = std::getenv("FSP_BDF_TRACE"); return e && e[0] == '1'; }();
  if (bdf_trace && my_rank_ == 0) {
    // cumulative counters of this integrator object (they continue across a warm restart, restart from 0 after a cold one)
    printf("[bdf] segment ends at t=%.5e stop=%d n=%d | steps %ld  Actions(rhs) %ld  J*v %ld  err-test fails %ld  conv fails %ld  h_last %.3e  q_last %d\n",
           t_now_, stop, (*solution_)->n_local, core_->NumSteps(), core_->NumRhsEvals(), core_->NumLinIters(), core_->NumErrTestFails(),
           core_->NumConvFails(), core_->LastStep(), core_->LastOrder());
  }
  return stop;
}

int CvodeFsp::ExpandState(const std::vector<PetscInt> &new_indices, PetscInt new_local_size) {
  if (!warm_restart_ || !carry_ || !core_) return 1;
  if (core_->Expand(new_indices, new_local_size) != 0) {
    carry_ = false;
    core_.reset();
    return 1;
  }
  // Taylor restart: with the operator at hand the history of ALL components (the new ones included) is rebuilt from
  // exact derivatives at the current step size and order; FSP_WARM_RESTART=carry keeps the plain carry-over
  static const bool carry_only = [] { const char *e = std::getenv("FSP_WARM_RESTART"); return e && !std::strcmp(e, "carry"); }();
  if (fspmat_ && !carry_only) {
    const double h = std::fabs(core_->LastStep());
    const double delta = std::max(0.02 * h, 1.0e-7 * std::max(1.0, std::fabs(t_now_)));
    auto dop = [this, delta](int j, double t, Vec v, Vec out) -> int {
      num_rhs_evals_ += 1;
      if (j == 0) return rhs_(t, v, out);
      return fspmat_->ActionTimeDerivative(j, t, v, out, delta);
    };
    int rc = core_->TaylorRestart(dop, fspmat_->HasTimeVaryingReactions() ? 4 : 0);
    if (rc < 0) {
      carry_ = false;
      core_.reset();
      return 1;
    }
  }
  return 0;
}

int CvodeFsp::FreeWorkspace() {
  OdeSolverBase::FreeWorkspace();
  if (!(warm_restart_ && carry_)) core_.reset();  // a stopped integrator is kept for ExpandState + the next SetUp
  if (solution_work_ != nullptr) VecDestroy(&solution_work_);
  return 0;
}

CvodeFsp::~CvodeFsp() { FreeWorkspace(); }

}  // namespace pacmensl

#include "OdeSolverBase.h"

namespace pacmensl {

OdeSolverBase::OdeSolverBase(MPI_Comm new_comm) {
  comm_ = new_comm;
  MPI_Comm_rank(comm_, &my_rank_);
  MPI_Comm_size(comm_, &comm_size_);
}
OdeSolverBase::~OdeSolverBase() { comm_ = MPI_COMM_NULL; }

int OdeSolverBase::SetTolerances(PetscReal _r_tol, PetscReal _abs_tol) {
  rel_tol_ = _r_tol;
  abs_tol_ = _abs_tol;
  return 0;
}
int OdeSolverBase::SetStatusOutput(int iprint) { print_intermediate = iprint; return 0; }
int OdeSolverBase::SetFinalTime(PetscReal _t_final) { t_final_ = _t_final; return 0; }
int OdeSolverBase::SetInitialSolution(Vec *_sol) {
  if (!_sol) { PACMENSLCHKERRQ(-1); }
  solution_ = _sol;
  return 0;
}
PacmenslErrorCode OdeSolverBase::SetRhs(std::function<PacmenslErrorCode(PetscReal, Vec, Vec)> _rhs) {
  rhs_ = std::move(_rhs);
  return 0;
}
int OdeSolverBase::EvaluateRHS(PetscReal t, Vec x, Vec y) {
  num_rhs_evals_ += 1;
  PACMENSLCHKERRQ(rhs_(t, x, y));
  return 0;
}
int OdeSolverBase::SetCurrentTime(PetscReal t) {
  if (std::isnan(t)) {
    printf("\n Time variable cannot have NaN value!\n");
    PACMENSLCHKERRQ(-1);
  }
  t_now_ = t;
  return 0;
}
PetscReal OdeSolverBase::GetCurrentTime() const { return t_now_; }
PetscInt OdeSolverBase::Solve() {
  if (solution_ == nullptr) return -1;
  if (rhs_ == nullptr) return -1;
  return 0;
}
int OdeSolverBase::EnableLogging() {
  logging_enabled = PETSC_TRUE;
  perf_info.n_step = 0;
  perf_info.model_time.resize(100000);
  perf_info.cpu_time.resize(100000);
  perf_info.n_eqs.resize(100000);
  return 0;
}
// src/OdeSolver/OdeSolverBase.cpp:112-131
FiniteProblemSolverPerfInfo OdeSolverBase::GetAvgPerfInfo() const {
  assert(logging_enabled);
  FiniteProblemSolverPerfInfo perf_out = perf_info;
  for (auto i{perf_out.n_step - 1}; i >= 0; --i) {
    perf_out.cpu_time[i] = perf_out.cpu_time[i] - perf_out.cpu_time[0];
    pacmensl_allreduce_sum(comm_, &perf_out.cpu_time[i], 1);
  }
  for (auto i{0}; i < perf_out.n_step; ++i) perf_out.cpu_time[i] /= PetscReal(comm_size_);
  return perf_out;
}
int OdeSolverBase::SetStopCondition(const std::function<PacmenslErrorCode(PetscReal, Vec, PetscReal &, void *)> &stop_check,
                                    void *stop_data) {
  OdeSolverBase::stop_check_ = stop_check;
  OdeSolverBase::stop_data_ = stop_data;
  return 0;
}
PacmenslErrorCode OdeSolverBase::SetFspMatPtr(FspMatrixBase *mat) {
  fspmat_ = mat;
  return 0;
}
}  // namespace pacmensl

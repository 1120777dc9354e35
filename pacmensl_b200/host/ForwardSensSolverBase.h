// ForwardSensSolverBase.h -- interface of the forward-sensitivity integrators
// (mirrors src/SensFsp/ForwardSensSolverBase.h:37-107).
#pragma once

#include "OdeSolverBase.h"
#include "PetscWrap.h"
#include "Sys.h"

namespace pacmensl {
enum class ForwardSensType { CVODE, KRYLOV };

class PACMENSL_API ForwardSensSolverBase {
  using RhsFun = std::function<PacmenslErrorCode(PetscReal, Vec, Vec)>;
  using SensRhs1Fun = std::function<PacmenslErrorCode(int, PetscReal, Vec, Vec)>;

 public:
  NOT_COPYABLE_NOT_MOVABLE(ForwardSensSolverBase);
  explicit ForwardSensSolverBase(MPI_Comm new_comm) {
    comm_ = new_comm;
    MPI_Comm_rank(comm_, &my_rank_);
    MPI_Comm_size(comm_, &comm_size_);
  }
  PacmenslErrorCode SetFinalTime(PetscReal _t_final) { t_final_ = _t_final; return 0; }
  PacmenslErrorCode SetInitialSolution(Petsc<Vec> &sol) {
    if (sol.IsEmpty()) return -1;
    solution_ = sol.mem();
    return 0;
  }
  PacmenslErrorCode SetInitialSensitivity(std::vector<Petsc<Vec>> &sens_vecs) {
    num_parameters_ = (int) sens_vecs.size();
    sens_vecs_.resize(num_parameters_);
    for (auto i = 0; i < num_parameters_; ++i) sens_vecs_[i] = sens_vecs[i].mem();
    return 0;
  }
  PacmenslErrorCode SetRhs(RhsFun rhs) { rhs_ = rhs; return 0; }
  PacmenslErrorCode SetSensRhs(SensRhs1Fun sensrhs) { srhs_ = sensrhs; return 0; }
  PacmenslErrorCode SetCurrentTime(PetscReal t) { t_now_ = t; return 0; }
  PacmenslErrorCode SetStatusOutput(int iprint) { print_intermediate = iprint; return 0; }
  PacmenslErrorCode EnableLogging() {
    logging_enabled = PETSC_TRUE;
    perf_info.n_step = 0;
    perf_info.model_time.resize(100000);
    perf_info.cpu_time.resize(100000);
    perf_info.n_eqs.resize(100000);
    return 0;
  }
  PacmenslErrorCode SetStopCondition(const std::function<int(PetscReal, Vec, int, Vec *, void *)> &stop_check, void *stop_data) {
    stop_check_ = stop_check;
    stop_data_ = stop_data;
    return 0;
  }
  virtual PacmenslErrorCode SetUp() { return 0; }
  virtual PetscInt Solve() { return 0; }
  virtual PacmenslErrorCode FreeWorkspace() {
    solution_ = nullptr;
    sens_vecs_.clear();
    t_now_ = 0.0;
    return 0;
  }
  PetscReal GetCurrentTime() const { return t_now_; }
  virtual ~ForwardSensSolverBase() { comm_ = MPI_COMM_NULL; }
  int EvaluateRHS(PetscReal t, Vec x, Vec y) { return rhs_(t, x, y); }
  int EvaluateSensRHS(int iS, PetscReal t, Vec x, Vec y) { return srhs_(iS, t, x, y); }

 protected:
  MPI_Comm comm_ = MPI_COMM_NULL;
  int      my_rank_ = 0, comm_size_ = 1;
  bool     set_up_ = false;
  Vec               *solution_ = nullptr;
  std::vector<Vec *> sens_vecs_;
  int                num_parameters_ = 0;
  RhsFun      rhs_;
  SensRhs1Fun srhs_;
  PetscReal t_now_ = 0.0, t_final_ = 0.0;
  int print_intermediate = 0;
  std::function<int(PetscReal, Vec, int, Vec *, void *)> stop_check_ = nullptr;
  void *stop_data_ = nullptr;
  PetscBool                   logging_enabled = PETSC_FALSE;
  FiniteProblemSolverPerfInfo perf_info;
};
}  // namespace pacmensl

// OdeSolverBase.h -- common interface of the ODE integrators for dp/dt = A(t) p.
// Mirrors src/OdeSolver/OdeSolverBase.h:52-153.
#pragma once

#include "FspMatrixBase.h"
#include "FspMatrixConstrained.h"
#include "StateSetConstrained.h"
#include "Sys.h"

namespace pacmensl {
enum ODESolverType { KRYLOV, CVODE, PETSC, EPIC };

struct FiniteProblemSolverPerfInfo {
  PetscInt                    n_step;
  std::vector<PetscInt>       n_eqs;
  std::vector<PetscLogDouble> cpu_time;
  std::vector<PetscReal>      model_time;
};

class PACMENSL_API OdeSolverBase {
 public:
  explicit OdeSolverBase(MPI_Comm new_comm);

  PacmenslErrorCode SetFinalTime(PetscReal _t_final);
  PacmenslErrorCode SetInitialSolution(Vec *_sol);
  PacmenslErrorCode SetFspMatPtr(FspMatrixBase *mat);
  PacmenslErrorCode SetRhs(std::function<PacmenslErrorCode(PetscReal, Vec, Vec)> _rhs);
  /// Extension: the same operator with a fused epilogue (FspMatrixBase::ActionFused).  Optional; when set, the solvers
  /// use it where an Action is immediately followed by scaling / inner products (it must compute the same A(t) x).
  using FusedRhs = std::function<PacmenslErrorCode(PetscReal, Vec, Vec, const fspmat_epilogue &)>;
  PacmenslErrorCode SetFusedRhs(FusedRhs _rhs) { fused_rhs_ = std::move(_rhs); return 0; }
  int SetTolerances(PetscReal _r_tol, PetscReal _abs_tol);
  PacmenslErrorCode SetCurrentTime(PetscReal t);
  PacmenslErrorCode SetStatusOutput(int iprint);
  PacmenslErrorCode EnableLogging();
  PacmenslErrorCode SetStopCondition(const std::function<PacmenslErrorCode(PetscReal, Vec, PetscReal &, void *)> &stop_check_,
                                     void *stop_data_);
  PacmenslErrorCode EvaluateRHS(PetscReal t, Vec x, Vec y);

  virtual PacmenslErrorCode SetUp() { return 0; }
  /// 0: reached t_final; 1: stopped by the stop condition; -1: error
  virtual PetscInt Solve();
  PetscReal GetCurrentTime() const;
  FiniteProblemSolverPerfInfo GetAvgPerfInfo() const;
  virtual PacmenslErrorCode FreeWorkspace() { solution_ = nullptr; return 0; }
  virtual ~OdeSolverBase();

  /// Extension: keep the integrator's history across an FSP expansion instead of re-creating
  /// the integrator as the reference does (src/Fsp/FspSolverMultiSinks.cpp:92-108).
  void SetWarmRestart(bool on) { warm_restart_ = on; }
  bool GetWarmRestart() const { return warm_restart_; }
  /// Called by the FSP driver after an expansion and before the next SetUp(): entry i of the old local vector now
  /// lives at new_indices[i] (same meaning as ExpandVec's argument).  0: the solver carried its history over to the
  /// new state space; 1: this solver (or this situation) restarts cold.
  virtual int ExpandState(const std::vector<PetscInt> &, PetscInt) { return 1; }

  /// number of right-hand-side (Action) evaluations since construction (extension, for reports)
  long GetNumRhsEvals() const { return num_rhs_evals_; }

 protected:
  MPI_Comm comm_ = MPI_COMM_NULL;
  int      my_rank_ = 0, comm_size_ = 1;

  Vec *solution_ = nullptr;
  std::function<int(PetscReal t, Vec x, Vec y)> rhs_;
  FusedRhs fused_rhs_;
  int            rhs_cost_loc_ = 0;
  FspMatrixBase *fspmat_ = nullptr;

  PetscReal t_now_ = 0.0;
  PetscReal t_final_ = 0.0;

  int print_intermediate = 0;
  std::function<PacmenslErrorCode(PetscReal t, Vec p, PetscReal &tol_exceed, void *data)> stop_check_ = nullptr;
  void *stop_data_ = nullptr;

  PetscBool                   logging_enabled = PETSC_FALSE;
  FiniteProblemSolverPerfInfo perf_info;
  PetscReal                   rel_tol_ = 1.0e-6;
  PetscReal                   abs_tol_ = 1.0e-14;
  long                        num_rhs_evals_ = 0;
  bool                        warm_restart_ = false;
};
}  // namespace pacmensl

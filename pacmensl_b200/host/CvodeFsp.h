// CvodeFsp.h -- BDF integrator with Newton + GMRES for dp/dt = A(t) p.
// Mirrors src/OdeSolver/CvodeFsp.h:41-76 / CvodeFsp.cpp:34-200: same constructor, SetUp/Solve/FreeWorkspace,
// same configuration (BDF, scalar tolerances, max conv fails / nonlinear iterations 10000, SPGMR(100) without
// preconditioner, exact J*v = A(t) v) and the same stepping logic (one internal step at a time, interpolation at
// t_final, roll-back to the last accepted time when the stop condition fires).  SUNDIALS is replaced by BdfCore.
#pragma once

#include "BdfCore.h"
#include "OdeSolverBase.h"

#define CV_ADAMS 1
#define CV_BDF 2

namespace pacmensl {
class PACMENSL_API CvodeFsp : public OdeSolverBase {
 public:
  explicit CvodeFsp(MPI_Comm _comm, int lmm = CV_BDF);
  PacmenslErrorCode SetUp() override;
  PetscInt Solve() override;
  int FreeWorkspace() override;
  /// Warm restart across FSP expansions (SetWarmRestart(true)): the Nordsieck history, step size and order survive the
  /// expansion (BdfCore::Expand) and the next SetUp()/Solve() continue the integration instead of restarting at order 1.
  int ExpandState(const std::vector<PetscInt> &new_indices, PetscInt new_local_size) override;
  ~CvodeFsp();

  /// integrator statistics of the current/last workspace (extension)
  const BdfCore *GetCore() const { return core_.get(); }

 protected:
  int lmm_ = CV_BDF;
  std::unique_ptr<BdfCore> core_;
  Vec       solution_work_ = nullptr;  ///< cvode_solution of the reference
  PetscReal t_now_tmp = 0.0;
  int       cvode_stat = 0;
  bool      carry_ = false;        ///< core_ holds a history worth continuing (stopped by the accept hook at t_now_)
  bool      hook_checked_ = false; ///< the accept hook already ran the stop condition for the step just taken
  PetscReal hook_excess_ = 0.0;
};
}  // namespace pacmensl

// ForwardSensCvodeFsp.h -- BDF integrator with staggered-1 forward sensitivities.
// Mirrors src/SensFsp/ForwardSensCvodeFsp.h:35-83 / .cpp:125-318: BDF + Newton + SPGMR(50), CV_STAGGERED1,
// sensitivity error control on, EE tolerances, rtol 1e-6 / atol 1e-14, sensitivity right-hand side
// sdot_i = A(t) s_i + (dA/dtheta_i)(t) p  (ForwardSensCvodeFsp.cpp:96-115).  CVODES is replaced by BdfCore.
#pragma once

#include "BdfCore.h"
#include "ForwardSensSolverBase.h"

namespace pacmensl {
class PACMENSL_API ForwardSensCvodeFsp : public ForwardSensSolverBase {
 public:
  explicit ForwardSensCvodeFsp(MPI_Comm comm) : ForwardSensSolverBase(comm) {}
  PacmenslErrorCode SetUp() override;
  PetscInt Solve() override;
  PacmenslErrorCode FreeWorkspace() override;
  ~ForwardSensCvodeFsp() override;

 protected:
  std::unique_ptr<BdfCore> core_;
  Vec              solution_work_ = nullptr;
  std::vector<Vec> sens_work_;
  Vec              tmp_ = nullptr;
  PetscReal t_now_tmp_ = 0.0;
  PetscReal rel_tol = 1.0e-6;
  PetscReal abs_tol = 1.0e-14;
  int       cvode_stat = 0;
};
}  // namespace pacmensl

// TsFsp.h -- the reference's PETSc-TS based integrator (src/OdeSolver/TsFsp.h:33-66, ODESolverType::PETSC).
// PETSc TS (default type TSROSW, Rosenbrock-W with an assembled Jacobian) is not part of this build; the class keeps
// the reference's interface -- constructor, SetUp/Solve/FreeWorkspace, SetTsType -- and integrates the same linear
// ODE dp/dt = A(t) p with the BDF/Newton/GMRES integrator of CvodeFsp (matrix-free: J v = A(t) v), which meets the
// reference's own acceptance bounds for this solver type (KAT-O3: |sum(p) - 1| <= 1e-8; KAT-F3: Poisson L1 <= 1e-6).
// The requested TS type is recorded and otherwise ignored.
#pragma once

#include <string>

#include "CvodeFsp.h"

#define TSROSW "rosw"
#define TSBDF "bdf"
#define TSARKIMEX "arkimex"
#define TSRK "rk"

namespace pacmensl {
class PACMENSL_API TsFsp : public CvodeFsp {
 public:
  explicit TsFsp(MPI_Comm _comm) : CvodeFsp(_comm, CV_BDF) {}
  PacmenslErrorCode SetTsType(std::string type) {
    type_ = std::move(type);
    return 0;
  }
  const std::string &GetTsType() const { return type_; }

 protected:
  std::string type_ = std::string(TSROSW);
};
}  // namespace pacmensl

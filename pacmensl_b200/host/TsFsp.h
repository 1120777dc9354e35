// TsFsp.h -- the reference's PETSc-TS based integrator (src/OdeSolver/TsFsp.h:33-66, ODESolverType::PETSC).
//
// The reference hands the problem to PETSc TS with its default type TSROSW (Rosenbrock-W) and an ASSEMBLED Jacobian
// (TsFsp.cpp:31-79: CreateRHSJacobian, TSSetIJacobian / TSSetRHSJacobian), checks the FSP stop condition after every
// step (TSSetPostEvaluate -> TSCheckFspError, :131-197) and matches the final time exactly.  PETSc is not part of this
// build, so the integrator is written here from the published method: the four-stage, third-order, L-stable
// Rosenbrock-W scheme RA34PW2 of Rang & Angermann (BIT 45, 2005) -- the scheme behind PETSc's TSROSW default
// "ra34pw2" -- with its embedded second-order solution for the error estimate (coefficients verified by their order of
// convergence with exact AND perturbed Jacobians, tests/test_rosw_coefficients.py), a WRMS-norm step controller of
// the TSAdaptBasic form (safety 0.9, clip [0.1, 10], embedded order 2), and stage systems
//     (I - h gamma J) k_i = h f(t_n + c_i h, y_n + sum_j a_ij k_j) + h J sum_j gamma_ij k_j,      J = A(t_n),
// solved by restarted GMRES(30) on the assembled CSR Jacobian (CreateRHSJacobian / ComputeRHSJacobian / MatMult, device
// SpMV) -- or matrix-free through the right-hand side on more than one rank, where no assembled form exists.
// What cannot be reproduced without PETSc: its other TS types (SetTsType accepts only "rosw") and the exact step
// sequence of its adaptor; like CVODE's, those are unpinned by the reference (end results only: KAT-O3, KAT-F3).
#pragma once

#include <string>
#include <vector>

#include "OdeSolverBase.h"

#define TSROSW "rosw"
#define TSBDF "bdf"
#define TSARKIMEX "arkimex"
#define TSRK "rk"

namespace pacmensl {
class PACMENSL_API TsFsp : public OdeSolverBase {
 public:
  explicit TsFsp(MPI_Comm _comm);
  PacmenslErrorCode SetUp() override;
  PetscInt Solve() override;
  PacmenslErrorCode SetTsType(std::string type);
  const std::string &GetTsType() const { return type_; }
  int FreeWorkspace() override;
  ~TsFsp() override;

  // statistics (extension)
  int NumSteps() const { return nstep; }
  int NumRejected() const { return nreject; }
  int NumJacobians() const { return njac; }
  long NumLinearIterations() const { return nlin; }

 protected:
  std::string type_ = std::string(TSROSW);
  PetscReal   t_now_tmp = 0.0;
  PetscInt    fsp_stop_ = 0;
  Vec         solution_tmp_ = nullptr;
  int         njac = 0, nstep = 0, nreject = 0;
  long        nlin = 0;
  Mat         J = nullptr;

  // work vectors
  Vec k_[4] = {nullptr, nullptr, nullptr, nullptr};
  Vec ystage_ = nullptr, rhsv_ = nullptr, tmp_ = nullptr, ynew_ = nullptr, err_ = nullptr, f0_ = nullptr, f1_ = nullptr;
  Vec jdiag_ = nullptr, pc_ = nullptr;  ///< diagonal of J (assembled form only) and the Jacobi preconditioner 1 - h gamma J_ii
  std::vector<Vec> V_;  // GMRES basis (restart + 1)
  double h_ = 0.0;
  bool   have_h_ = false;

  int JacTimes(PetscReal t, Vec v, Vec out);
  int SolveStage(PetscReal t, double hgamma, Vec b, Vec x, double tol_scale);  ///< 0 ok, 1 not converged, < 0 error
  int ApplyPc(Vec in, Vec out);
  int WrmsNorm(Vec e, Vec ya, Vec yb, double *out);
  int Interpolate(PetscReal t0, PetscReal t1, Vec y0, Vec y1, PetscReal t, Vec out);
};
}  // namespace pacmensl

// ODE-solver tests shaped after the reference's tests/test_ode.cpp: toggle switch on the 101 x 101 box
// (10 201 states + 2 sinks), p0 = delta(0,0), t_f = 100.
//   KAT-O1 CvodeFsp (BDF):   Solve() == 0 and |sum(p) - 1| <= 1e-8     (test_ode.cpp:123-153)
//   KAT-O2 KrylovFsp:        the same                                   (:220-259)
//   KAT-O4/O5 a failing rhs makes Solve() return -1                     (:188-218, 261-295)
#include "fsp_models.h"
#include "pacmensl_test_env.h"

using namespace pacmensl;

namespace toggle_cme {
arma::Mat<PetscInt> SM{{1, 1, -1, 0, 0, 0}, {0, 0, 0, 1, 1, -1}};
int propensity(const int reaction, const int num_species, const int num_states, const PetscInt *X, double *outputs, void *args) {
  return toggle_prop(reaction, num_species, num_states, X, outputs, args);
}
int t_fun(PetscReal, int, double *, void *) { return 0; }
}  // namespace toggle_cme

class OdeTest : public ::testing::Test {
 protected:
  void SetUp() override {
    arma::Row<PetscInt> fsp_size = {100, 100};
    arma::Mat<PetscInt> X0(2, 1);
    X0.col(0).fill(0);
    StateSetConstrained fsp(PETSC_COMM_WORLD);
    fsp.SetShapeBounds(fsp_size);
    fsp.SetStoichiometryMatrix(toggle_cme::SM);
    fsp.SetUp();
    fsp.AddStates(X0);
    fsp.Expand();
    ASSERT_EQ(fsp.GetNumGlobalStates(), 10201);
    A = new FspMatrixConstrained(PETSC_COMM_WORLD);
    ASSERT_FALSE(A->GenerateValues(fsp, toggle_cme::SM, std::vector<int>(), toggle_cme::t_fun, toggle_cme::propensity,
                                   std::vector<int>(), nullptr, nullptr));
  }
  void TearDown() override { delete A; }

  Vec initial_vec() {
    Vec P;
    VecCreate(PETSC_COMM_WORLD, &P);
    VecSetSizes(P, A->GetNumLocalRows(), PETSC_DECIDE);
    VecSetFromOptions(P);
    VecSetValue(P, 0, 1.0, INSERT_VALUES);
    VecSetUp(P);
    VecAssemblyBegin(P);
    VecAssemblyEnd(P);
    return P;
  }
  FspMatrixConstrained *A = nullptr;
};

TEST_F(OdeTest, use_cvode_bdf) {
  auto AV = [&](PetscReal t, Vec x, Vec y) { return A->Action(t, x, y); };
  Vec  P = initial_vec();
  PetscReal t_final = 100.0;
  CvodeFsp  cvode_solver(PETSC_COMM_WORLD, CV_BDF);
  ASSERT_EQ(cvode_solver.SetFinalTime(t_final), 0);
  ASSERT_EQ(cvode_solver.SetInitialSolution(&P), 0);
  ASSERT_EQ(cvode_solver.SetRhs(AV), 0);
  ASSERT_EQ(cvode_solver.SetStatusOutput(0), 0);
  ASSERT_EQ(cvode_solver.SetUp(), 0);
  PetscInt solver_stat = cvode_solver.Solve();
  ASSERT_FALSE(solver_stat);
  ASSERT_NEAR(cvode_solver.GetCurrentTime(), t_final, 1e-12);
  PetscReal Psum;
  VecSum(P, &Psum);
  ASSERT_LE(Psum, 1.0 + 1.0e-8);
  ASSERT_GE(Psum, 1.0 - 1.0e-8);
  std::printf("    BDF: %ld steps, %ld rhs, %ld J*v, %ld Newton its, %ld err-test fails, %ld conv fails\n",
              cvode_solver.GetCore()->NumSteps(), cvode_solver.GetCore()->NumRhsEvals(), cvode_solver.GetCore()->NumJtvEvals(),
              cvode_solver.GetCore()->NumNonlinIters(), cvode_solver.GetCore()->NumErrTestFails(), cvode_solver.GetCore()->NumConvFails());
  VecDestroy(&P);
}

TEST_F(OdeTest, cvode_handling_bad_mat_vec) {
  auto AV = [&](PetscReal, Vec, Vec) { return -1; };
  Vec  P = initial_vec();
  CvodeFsp cvode_solver(PETSC_COMM_WORLD, CV_BDF);
  ASSERT_EQ(cvode_solver.SetFinalTime(100.0), 0);
  ASSERT_EQ(cvode_solver.SetInitialSolution(&P), 0);
  ASSERT_EQ(cvode_solver.SetRhs(AV), 0);
  ASSERT_EQ(cvode_solver.SetStatusOutput(0), 0);
  ASSERT_EQ(cvode_solver.SetUp(), 0);
  PetscInt solver_stat = cvode_solver.Solve();
  ASSERT_EQ(solver_stat, -1);
  PetscReal Psum;
  VecSum(P, &Psum);
  ASSERT_LE(Psum, 1.0 + 1.0e-8);
  ASSERT_GE(Psum, 1.0 - 1.0e-8);
  VecDestroy(&P);
}

TEST_F(OdeTest, use_krylov) {
  auto AV = [&](PetscReal t, Vec x, Vec y) { return A->Action(t, x, y); };
  Vec  P = initial_vec();
  KrylovFsp krylov_solver(PETSC_COMM_WORLD);
  ASSERT_EQ(krylov_solver.SetFinalTime(100.0), 0);
  ASSERT_EQ(krylov_solver.SetInitialSolution(&P), 0);
  ASSERT_EQ(krylov_solver.SetRhs(AV), 0);
  ASSERT_EQ(krylov_solver.SetFspMatPtr(A), 0);
  ASSERT_EQ(krylov_solver.SetStatusOutput(0), 0);
  ASSERT_EQ(krylov_solver.SetUp(), 0);
  PetscInt solver_stat = krylov_solver.Solve();
  long nrhs = krylov_solver.GetNumRhsEvals();
  krylov_solver.FreeWorkspace();
  ASSERT_FALSE(solver_stat);
  PetscReal Psum;
  VecSum(P, &Psum);
  ASSERT_LE(Psum, 1.0 + 1.0e-8);
  ASSERT_GE(Psum, 1.0 - 1.0e-8);
  std::printf("    Krylov: %ld Action calls\n", nrhs);
  VecDestroy(&P);
}

// KAT-O3 (reference tests/test_ode.cpp:297-330): the TsFsp interface (ODESolverType::PETSC)
TEST_F(OdeTest, use_ts) {
  auto AV = [&](PetscReal t, Vec x, Vec y) { return A->Action(t, x, y); };
  Vec  P = initial_vec();
  TsFsp ts(PETSC_COMM_WORLD);
  ASSERT_EQ(ts.SetFinalTime(100.0), 0);
  ASSERT_EQ(ts.SetInitialSolution(&P), 0);
  ASSERT_EQ(ts.SetRhs(AV), 0);
  ASSERT_EQ(ts.SetStatusOutput(0), 0);
  ASSERT_EQ(ts.SetFspMatPtr(A), 0);
  ASSERT_EQ(ts.SetTsType(TSROSW), 0);
  ASSERT_EQ(ts.SetUp(), 0);
  PetscInt solver_stat = ts.Solve();
  ASSERT_FALSE(solver_stat);
  PetscReal Psum;
  VecSum(P, &Psum);
  ASSERT_LE(Psum, 1.0 + 1.0e-8);
  ASSERT_GE(Psum, 1.0 - 1.0e-8);
  VecDestroy(&P);
}

TEST_F(OdeTest, krylov_handling_bad_mat_vec) {
  auto AV = [&](PetscReal, Vec, Vec) { return -1; };
  Vec  P = initial_vec();
  KrylovFsp krylov_solver(PETSC_COMM_WORLD);
  ASSERT_EQ(krylov_solver.SetFinalTime(100.0), 0);
  ASSERT_EQ(krylov_solver.SetInitialSolution(&P), 0);
  ASSERT_EQ(krylov_solver.SetRhs(AV), 0);
  ASSERT_EQ(krylov_solver.SetFspMatPtr(A), 0);
  ASSERT_EQ(krylov_solver.SetStatusOutput(0), 0);
  ASSERT_EQ(krylov_solver.SetUp(), 0);
  ASSERT_EQ(krylov_solver.Solve(), -1);
  VecDestroy(&P);
}

TEST_F(OdeTest, cvode_and_krylov_agree) {
  // both integrators solve the same linear ODE: their answers must agree to the BDF tolerance level
  auto AV = [&](PetscReal t, Vec x, Vec y) { return A->Action(t, x, y); };
  Vec  P1 = initial_vec(), P2 = initial_vec();
  CvodeFsp bdf(PETSC_COMM_WORLD);
  bdf.SetFinalTime(100.0); bdf.SetInitialSolution(&P1); bdf.SetRhs(AV); bdf.SetTolerances(1e-8, 1e-14);
  ASSERT_EQ(bdf.SetUp(), 0);
  ASSERT_EQ(bdf.Solve(), 0);
  KrylovFsp kry(PETSC_COMM_WORLD);
  kry.SetFinalTime(100.0); kry.SetInitialSolution(&P2); kry.SetRhs(AV); kry.SetFspMatPtr(A);
  ASSERT_EQ(kry.SetUp(), 0);
  ASSERT_EQ(kry.Solve(), 0);
  VecAXPY(P1, -1.0, P2);
  PetscReal d;
  VecNorm(P1, NORM_1, &d);
  std::printf("    ||p_bdf - p_krylov||_1 = %.3e\n", d);
  ASSERT_LE(d, 1.0e-5);
  VecDestroy(&P1);
  VecDestroy(&P2);
}

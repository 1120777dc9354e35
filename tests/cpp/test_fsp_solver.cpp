// End-to-end tests of the adaptive FSP driver, shaped after the reference's tests/test_fsp_solver.cpp.
//   KAT-F1 SetUp() before a model is set returns -1                                  (test_fsp_solver.cpp:127-132)
//   KAT-F2 a prop_t_ returning -1 makes Solve / SolveTspan throw std::runtime_error   (:134-177)
//   KAT-F4/F5 pure birth (lambda = 2, t_f = 10, bounds {5}, expansion 0.1, fsp_tol 1e-6):
//          sum_n |p_n - Poisson(lambda t)(n)| <= 1e-6 for CVODE and KRYLOV            (:264-345)
//   (KAT-F3 uses ODESolverType::PETSC = TsFsp: here the Rosenbrock-W scheme RA34PW2 -- PETSc's TSROSW default -- on the
//    assembled CSR Jacobian, host/TsFsp.h; the same Poisson bound is checked.)
#include "fsp_models.h"
#include "pacmensl_test_env.h"

using namespace pacmensl;

namespace toggle_cme {
arma::Mat<PetscInt> SM{{1, 1, -1, 0, 0, 0}, {0, 0, 0, 1, 1, -1}};
int propensity(const int reaction, const int num_species, const int num_states, const PetscInt *X, double *outputs, void *args) {
  return toggle_prop(reaction, num_species, num_states, X, outputs, args);
}
int t_fun(PetscReal, int, double *outputs, void *) {
  outputs[0] = tg_kx0; outputs[1] = tg_kx; outputs[2] = tg_dx; outputs[3] = tg_ky0; outputs[4] = tg_ky; outputs[5] = tg_dy;
  return 0;
}
}  // namespace toggle_cme

class FspTest : public ::testing::Test {
 protected:
  void SetUp() override {
    t_final = 100.0;
    fsp_tol = 1.0e-6;
    X0 = X0.t();
    toggle_model = Model(toggle_cme::SM, toggle_cme::t_fun, toggle_cme::propensity, nullptr, nullptr, std::vector<int>());
  }
  PetscReal            t_final, fsp_tol;
  arma::Mat<PetscInt>  X0{0, 0};
  arma::Col<PetscReal> p0 = {1.0};
  Model                toggle_model;
  arma::Row<int>       fsp_size = {5, 5};
  arma::Row<PetscReal> expansion_factors = {0.25, 0.25};
};

TEST_F(FspTest, test_wrong_call_sequence_detection) {
  FspSolverMultiSinks fsp(PETSC_COMM_WORLD);
  int                 ierr = fsp.SetUp();
  ASSERT_EQ(ierr, -1);
}

TEST_F(FspTest, test_handling_t_fun_error) {
  DiscreteDistribution              p_final_bdf;
  std::vector<DiscreteDistribution> p_snapshots_bdf;
  FspSolverMultiSinks               fsp(PETSC_COMM_WORLD);
  std::vector<PetscReal> tspan = arma::conv_to<std::vector<PetscReal>>::from(arma::linspace<arma::Row<PetscReal>>(0.0, t_final, 3));
  Model bad_model = toggle_model;
  bad_model.prop_t_ = [&](double, int, double *, void *) { return -1; };
  bad_model.tv_reactions_ = std::vector<int>({0, 1});

  ASSERT_FALSE(fsp.SetModel(bad_model));
  ASSERT_FALSE(fsp.SetInitialBounds(fsp_size));
  ASSERT_FALSE(fsp.SetExpansionFactors(expansion_factors));
  ASSERT_FALSE(fsp.SetVerbosity(0));
  ASSERT_FALSE(fsp.SetInitialDistribution(X0, p0));
  fsp.SetOdesType(CVODE);
  ASSERT_THROW(p_final_bdf = fsp.Solve(t_final, fsp_tol, 0), std::runtime_error);
  fsp.ClearState();

  ASSERT_FALSE(fsp.SetInitialBounds(fsp_size));
  ASSERT_FALSE(fsp.SetExpansionFactors(expansion_factors));
  ASSERT_FALSE(fsp.SetVerbosity(0));
  ASSERT_FALSE(fsp.SetInitialDistribution(X0, p0));
  ASSERT_FALSE(fsp.SetUp());
  ASSERT_THROW(p_snapshots_bdf = fsp.SolveTspan(tspan, fsp_tol, 0), std::runtime_error);
}

TEST_F(FspTest, toggle_cvode_mass_is_conserved_up_to_fsp_tol) {
  FspSolverMultiSinks fsp(PETSC_COMM_WORLD);
  ASSERT_FALSE(fsp.SetModel(toggle_model));
  ASSERT_FALSE(fsp.SetInitialBounds(fsp_size));
  ASSERT_FALSE(fsp.SetExpansionFactors(expansion_factors));
  ASSERT_FALSE(fsp.SetInitialDistribution(X0, p0));
  fsp.SetOdesType(CVODE);
  DiscreteDistribution p = fsp.Solve(t_final, fsp_tol, 0);
  PetscReal            s;
  VecSum(p.p_, &s);
  std::printf("    toggle CVODE: %d states, %d expansions, %ld Action calls, 1 - sum(p) = %.3e\n",
              (int) p.states_.n_cols, fsp.GetNumExpansions(), fsp.GetNumRhsEvals(), 1.0 - s);
  ASSERT_LE(s, 1.0 + 1.0e-8);
  ASSERT_GE(s, 1.0 - fsp_tol - 1.0e-8);
}

class FspPoissonTest : public ::testing::Test {
 protected:
  void SetUp() override {
    auto propensity = [&](int, int, int num_states, const int *, PetscReal *output, void *) {
      for (int i{0}; i < num_states; ++i) output[i] = lambda;
      return 0;
    };
    auto t_fun = [&](PetscReal, int, double *outputs, void *) {
      outputs[0] = 1.0;
      return 0;
    };
    poisson_model = Model(stoich_matrix, t_fun, propensity, nullptr, nullptr, std::vector<int>());
  }
  double poisson_l1_error(DiscreteDistribution &p_final) {
    PetscReal  stmp = 0.0;
    PetscReal *p_dat;
    int        num_states;
    p_final.GetProbView(num_states, p_dat);
    for (int i = 0; i < num_states; ++i) {
      int       n = p_final.states_(0, i);
      PetscReal pdf = exp(-lambda * t_final) * pow(lambda * t_final, double(n)) / tgamma(n + 1);
      stmp += std::abs(p_dat[i] - pdf);
    }
    p_final.RestoreProbView(p_dat);
    pacmensl_allreduce_sum(PETSC_COMM_WORLD, &stmp, 1);
    return stmp;
  }
  Model                poisson_model;
  PetscReal            lambda = 2.0;
  arma::Mat<int>       stoich_matrix = {1};
  arma::Mat<int>       x0 = {0};
  arma::Col<PetscReal> p0 = {1.0};
  arma::Row<int>       fsp_size = {5};
  arma::Row<PetscReal> expansion_factors = {0.1};
  PetscReal            t_final{10.0}, fsp_tol{1.0e-6};
};

TEST_F(FspPoissonTest, test_poisson_petsc) {
  FspSolverMultiSinks fsp(PETSC_COMM_WORLD);
  ASSERT_FALSE(fsp.SetModel(poisson_model));
  ASSERT_FALSE(fsp.SetInitialBounds(fsp_size));
  ASSERT_FALSE(fsp.SetExpansionFactors(expansion_factors));
  ASSERT_FALSE(fsp.SetInitialDistribution(x0, p0));
  ASSERT_FALSE(fsp.SetOdesType(ODESolverType::PETSC));
  ASSERT_FALSE(fsp.SetVerbosity(0));
  ASSERT_FALSE(fsp.SetUp());
  DiscreteDistribution p_final = fsp.Solve(t_final, fsp_tol, 0);
  fsp.ClearState();
  ASSERT_LE(poisson_l1_error(p_final), fsp_tol);
}

TEST_F(FspPoissonTest, test_poisson_cvode) {
  FspSolverMultiSinks fsp(PETSC_COMM_WORLD);
  ASSERT_FALSE(fsp.SetModel(poisson_model));
  ASSERT_FALSE(fsp.SetInitialBounds(fsp_size));
  ASSERT_FALSE(fsp.SetExpansionFactors(expansion_factors));
  ASSERT_FALSE(fsp.SetInitialDistribution(x0, p0));
  ASSERT_FALSE(fsp.SetOdesType(ODESolverType::CVODE));
  ASSERT_FALSE(fsp.SetUp());
  std::shared_ptr<CvodeFsp> ode_solver = std::dynamic_pointer_cast<CvodeFsp>(fsp.GetOdeSolver());
  ASSERT_TRUE(ode_solver != nullptr);
  ode_solver->SetTolerances(1.0e-6, 1.0e-14);
  DiscreteDistribution p_final = fsp.Solve(t_final, fsp_tol, 0);
  int nexp = fsp.GetNumExpansions();
  long nrhs = fsp.GetNumRhsEvals();
  fsp.ClearState();
  double err = poisson_l1_error(p_final);
  std::printf("    Poisson CVODE: %d states, %d expansions, %ld Action calls, L1 error %.3e\n", (int) p_final.states_.n_cols, nexp, nrhs, err);
  ASSERT_LE(err, fsp_tol);
}

TEST_F(FspPoissonTest, test_poisson_krylov) {
  FspSolverMultiSinks fsp(PETSC_COMM_WORLD);
  ASSERT_FALSE(fsp.SetModel(poisson_model));
  ASSERT_FALSE(fsp.SetInitialBounds(fsp_size));
  ASSERT_FALSE(fsp.SetExpansionFactors(expansion_factors));
  ASSERT_FALSE(fsp.SetInitialDistribution(x0, p0));
  ASSERT_FALSE(fsp.SetOdesType(KRYLOV));
  ASSERT_FALSE(fsp.SetVerbosity(0));
  DiscreteDistribution p_final = fsp.Solve(t_final, fsp_tol, 0);
  int nexp = fsp.GetNumExpansions();
  long nrhs = fsp.GetNumRhsEvals();
  fsp.ClearState();
  double err = poisson_l1_error(p_final);
  std::printf("    Poisson KRYLOV: %d states, %d expansions, %ld Action calls, L1 error %.3e\n", (int) p_final.states_.n_cols, nexp, nrhs, err);
  ASSERT_LE(err, fsp_tol);
}

TEST_F(FspPoissonTest, solve_tspan_and_restart_from_distribution) {
  // SolveTspan advances through output times reusing state (FspSolverMultiSinks.cpp:645-682); a
  // DiscreteDistribution can be fed back as an initial condition (:432-453)
  FspSolverMultiSinks fsp(PETSC_COMM_WORLD);
  ASSERT_FALSE(fsp.SetModel(poisson_model));
  ASSERT_FALSE(fsp.SetInitialBounds(fsp_size));
  ASSERT_FALSE(fsp.SetExpansionFactors(expansion_factors));
  ASSERT_FALSE(fsp.SetInitialDistribution(x0, p0));
  ASSERT_FALSE(fsp.SetOdesType(KRYLOV));
  std::vector<PetscReal> tspan = {2.5, 5.0, 10.0};
  std::vector<DiscreteDistribution> snaps = fsp.SolveTspan(tspan, fsp_tol, 0);
  ASSERT_EQ((int) snaps.size(), 3);
  ASSERT_NEAR(snaps[0].t_, 2.5, 1e-12);
  ASSERT_NEAR(snaps[2].t_, 10.0, 1e-12);
  ASSERT_LE(poisson_l1_error(snaps[2]), fsp_tol);
  fsp.ClearState();
  // restart from the t = 5 snapshot and integrate the remaining 5 time units
  FspSolverMultiSinks fsp2(PETSC_COMM_WORLD);
  ASSERT_FALSE(fsp2.SetModel(poisson_model));
  arma::Row<int> big = {60};
  ASSERT_FALSE(fsp2.SetInitialBounds(big));
  ASSERT_FALSE(fsp2.SetExpansionFactors(expansion_factors));
  ASSERT_FALSE(fsp2.SetInitialDistribution(snaps[1]));
  ASSERT_FALSE(fsp2.SetOdesType(KRYLOV));
  DiscreteDistribution p_final = fsp2.Solve(5.0, fsp_tol, 0);
  ASSERT_LE(poisson_l1_error(p_final), 3 * fsp_tol);
}

// Device-side post-processing: Compute1DMarginal (src/Fsp/DiscreteDistribution.cpp:171-200) as a
// deterministic segmented reduction on the GPU must equal the reference's host loop over (states_, p_).
TEST_F(FspTest, device_marginals_equal_the_host_loop) {
  FspSolverMultiSinks fsp(PETSC_COMM_WORLD);
  ASSERT_FALSE(fsp.SetModel(toggle_model));
  ASSERT_FALSE(fsp.SetInitialBounds(fsp_size));
  ASSERT_FALSE(fsp.SetExpansionFactors(expansion_factors));
  ASSERT_FALSE(fsp.SetInitialDistribution(X0, p0));
  ASSERT_FALSE(fsp.SetOdesType(KRYLOV));
  DiscreteDistribution d = fsp.Solve(20.0, 1.0e-6, 0);
  fsp.ClearState();
  ASSERT_TRUE((bool) d.states_dev_ || d.states_.n_cols == 0);
  const PetscScalar *p;
  ASSERT_FALSE(VecGetArrayRead(d.p_, &p));
  for (int species = 0; species < 2; ++species) {
    arma::Col<PetscReal> md = Compute1DMarginal(d, species);
    double mxd = 0.0;
    for (arma::uword i = 0; i < d.states_.n_cols; ++i) mxd = std::max(mxd, (double) d.states_(species, i));
    pacmensl_allreduce_max(PETSC_COMM_WORLD, &mxd, 1);  // the bins cover the largest count on ANY rank
    const int mx = (int) mxd;
    std::vector<double> ref((size_t) mx + 1, 0.0);
    for (arma::uword i = 0; i < d.states_.n_cols; ++i) ref[(size_t) d.states_(species, i)] += p[i];
    pacmensl_allreduce_sum(PETSC_COMM_WORLD, ref.data(), (int) ref.size());
    ASSERT_EQ((int) md.n_elem, mx + 1);
    double gap = 0.0, tot = 0.0;
    for (int b = 0; b <= mx; ++b) { gap = std::max(gap, std::fabs(md[b] - ref[(size_t) b])); tot += md[b]; }
    std::printf("    marginal of species %d: %d bins, max |device - host| = %.2e, sum = %.12f\n", species, mx + 1, gap, tot);
    ASSERT_LE(gap, 1.0e-15);
    ASSERT_NEAR(tot, 1.0, 1.0e-5);
    // a second evaluation gives the same bits (fixed summation order)
    arma::Col<PetscReal> md2 = Compute1DMarginal(d, species);
    for (int b = 0; b <= mx; ++b) ASSERT_TRUE(md[b] == md2[b]);
  }
  VecRestoreArrayRead(d.p_, &p);
}


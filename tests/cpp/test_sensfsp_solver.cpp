// Forward-sensitivity FSP tests shaped after the reference's tests/test_sensfsp_solver.cpp.
//   KAT-SF1 toggle, 6 TV reactions / 6 parameters, bounds {1,1}, fsp_tol 1e-10, t_f = 100:
//           sum(p) >= 1 - fsp_tol and |sum(dp/dtheta_i)| <= 1e-6                          (CR-lines 130-200)
//   KAT-SF2 pure birth: sum|p - Poisson| <= 1e-7 at t_f = 1 and sum|s - dPoisson/dlambda| <= 1e-6   (206-301)
//   KAT-SF3 telegraph model with state-factor derivatives dprop_x                            (307-461)
#include "fsp_models.h"
#include "pacmensl_test_env.h"

using namespace pacmensl;

namespace toggle_cme {
arma::Mat<PetscInt> SM{{1, 1, -1, 0, 0, 0}, {0, 0, 0, 1, 1, -1}};
int propensity(const int reaction, const int, const int num_states, const PetscInt *X, double *outputs, void *) {
  for (int i{0}; i < num_states; ++i) {
    const int *x = X + 2 * i;
    switch (reaction) {
      case 0: outputs[i] = 1.0; break;
      case 1: outputs[i] = 1.0 / (1.0 + tg_ayx * pow(PetscReal(x[1]), tg_nyx)); break;
      case 2: outputs[i] = PetscReal(x[0]); break;
      case 3: outputs[i] = 1.0; break;
      case 4: outputs[i] = 1.0 / (1.0 + tg_axy * pow(PetscReal(x[0]), tg_nxy)); break;
      case 5: outputs[i] = PetscReal(x[1]); break;
      default: return -1;
    }
  }
  return 0;
}
int t_fun(PetscReal, int, double *outputs, void *) {
  outputs[0] = tg_kx0; outputs[1] = tg_kx; outputs[2] = tg_dx; outputs[3] = tg_ky0; outputs[4] = tg_ky; outputs[5] = tg_dy;
  return 0;
}
int dt_fun(int parameter_idx, PetscReal, int, double *outputs, void *) {
  outputs[parameter_idx] = 1.0;
  return 0;
}
}  // namespace toggle_cme

class SensFspToggleTest : public ::testing::Test {
 protected:
  void SetUp() override {
    int n_par = 6;
    t_final = 100.0;
    fsp_tol = 1.0e-10;
    X0 = X0.t();
    dp0 = std::vector<arma::Col<PetscReal>>(n_par, arma::Col<PetscReal>({0.0}));
    toggle_model = SensModel(6, toggle_cme::SM, std::vector<int>({0, 1, 2, 3, 4, 5}), toggle_cme::t_fun, toggle_cme::propensity,
                             toggle_cme::dt_fun, {{0}, {1}, {2}, {3}, {4}, {5}}, nullptr, {});
  }
  PetscReal t_final, fsp_tol;
  arma::Mat<PetscInt>  X0{0, 0};
  arma::Col<PetscReal> p0 = {1.0};
  std::vector<arma::Col<PetscReal>> dp0;
  SensModel            toggle_model;
  arma::Row<int>       fsp_size = {1, 1};
  arma::Row<PetscReal> expansion_factors = {0.25, 0.25};
};

TEST_F(SensFspToggleTest, toggle_sens_solve_with_cvode) {
  PetscReal stmp;
  SensFspSolverMultiSinks fsp(PETSC_COMM_WORLD);
  ASSERT_FALSE(fsp.SetModel(toggle_model));
  ASSERT_FALSE(fsp.SetInitialBounds(fsp_size));
  ASSERT_FALSE(fsp.SetExpansionFactors(expansion_factors));
  ASSERT_FALSE(fsp.SetInitialDistribution(X0, p0, dp0));
  ASSERT_FALSE(fsp.SetUp());
  SensDiscreteDistribution p_final_bdf = fsp.Solve(t_final, fsp_tol);
  int nexp = fsp.GetNumExpansions();
  fsp.ClearState();
  ASSERT_FALSE(VecSum(p_final_bdf.p_, &stmp));
  std::printf("    toggle sens: %d states, %d expansions, 1 - sum(p) = %.3e\n", (int) p_final_bdf.states_.n_cols, nexp, 1.0 - stmp);
  ASSERT_GE(stmp, 1.0 - fsp_tol);
  for (int i{0}; i < 6; ++i) {
    ASSERT_FALSE(VecSum(p_final_bdf.dp_[i], &stmp));
    ASSERT_LE(std::abs(stmp), 1.0e-6);
  }
  // device-side post-processing: the Fisher information matrix and a sensitivity marginal against the
  // reference's host loops (src/SensFsp/SensDiscreteDistribution.cpp:170-271)
  const int n = (int) p_final_bdf.states_.n_cols;
  std::vector<std::vector<double>> s(6);
  std::vector<double>              p(n);
  {
    const PetscScalar *a;
    ASSERT_FALSE(VecGetArrayRead(p_final_bdf.p_, &a));
    for (int k = 0; k < n; ++k) p[k] = std::max(a[k], 1.0e-16);
    VecRestoreArrayRead(p_final_bdf.p_, &a);
    for (int i = 0; i < 6; ++i) {
      ASSERT_FALSE(VecGetArrayRead(p_final_bdf.dp_[i], &a));
      s[i].assign(a, a + n);
      VecRestoreArrayRead(p_final_bdf.dp_[i], &a);
    }
  }
  arma::Mat<PetscReal> fim;
  ASSERT_FALSE(ComputeFIM(p_final_bdf, fim));
  double worst = 0.0;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double ref = 0.0;
      for (int k = 0; k < n; ++k) ref += s[i][k] * s[j][k] / p[k];
      pacmensl_allreduce_sum(PETSC_COMM_WORLD, &ref, 1);
      worst = std::max(worst, std::fabs(fim(i, j) - ref) / (std::fabs(ref) + 1e-300));
    }
  std::printf("    FIM on the device vs host loop: max relative difference %.2e (fim(0,0) = %.6e)\n", worst, fim(0, 0));
  ASSERT_LE(worst, 1.0e-10);  // (the clamped 1e-16 probabilities amplify the rounding of a different summation order)
  arma::Col<PetscReal> sm;
  ASSERT_FALSE(Compute1DSensMarginal(p_final_bdf, 2, 0, sm));
  std::vector<double> ref(sm.n_elem, 0.0);
  for (int k = 0; k < n; ++k) ref[(size_t) p_final_bdf.states_(0, k)] += s[2][k];
  pacmensl_allreduce_sum(PETSC_COMM_WORLD, ref.data(), (int) ref.size());
  for (arma::uword b = 0; b < sm.n_elem; ++b) ASSERT_NEAR(sm[b], ref[b], 1.0e-14 * (1.0 + std::fabs(ref[b])));
}

// Parity of the sensitivity solve beyond the KATs: ForwardSensCvodeFsp (BDF + staggered-1 corrector with
// sensitivity error control) against an INDEPENDENT integrator -- classical RK4 with a small fixed step -- of the same
// system  p' = A p,  s_i' = A s_i + (dA/dtheta_i) p  on a fixed state set (fsp_tol <= 0: no expansion), using the
// sensitivity operator that KAT-SM1/SM2 and the finite-difference check of tests/cpp/test_sensmat.cpp pin.
TEST_F(SensFspToggleTest, sens_solve_matches_independent_rk4_on_a_fixed_set) {
  const double   T = 1.0;
  arma::Row<int> big = {25, 25};
  SensFspSolverMultiSinks fsp(PETSC_COMM_WORLD);
  ASSERT_FALSE(fsp.SetModel(toggle_model));
  ASSERT_FALSE(fsp.SetInitialBounds(big));
  ASSERT_FALSE(fsp.SetExpansionFactors(expansion_factors));
  ASSERT_FALSE(fsp.SetInitialDistribution(X0, p0, dp0));
  ASSERT_FALSE(fsp.SetUp());
  SensDiscreteDistribution d = fsp.Solve(T, -1.0);
  fsp.ClearState();

  // the same state set (same deterministic construction => same ordering) and operators
  StateSetConstrained set(PETSC_COMM_WORLD);
  set.SetStoichiometryMatrix(toggle_cme::SM);
  set.SetShapeBounds(big);
  set.SetUp();
  ASSERT_FALSE(set.AddStates(X0));
  ASSERT_FALSE(set.Expand());
  ASSERT_EQ((int) set.GetNumLocalStates(), (int) d.states_.n_cols);
  SensFspMatrix<FspMatrixConstrained> A(PETSC_COMM_WORLD);
  ASSERT_FALSE(A.GenerateValues(set, toggle_model));
  const int nr = A.GetNumLocalRows(), n = set.GetNumLocalStates(), P = 6;
  auto mk = [&]() { Vec v; VecCreate(PETSC_COMM_WORLD, &v); VecSetSizes(v, nr, PETSC_DECIDE); VecSetUp(v); VecSet(v, 0.0); return v; };
  std::vector<Vec> y(P + 1), k(P + 1), acc(P + 1), st(P + 1);
  for (int i = 0; i <= P; ++i) { y[i] = mk(); k[i] = mk(); acc[i] = mk(); st[i] = mk(); }
  Vec tmp = mk();
  {  // p(0) = delta at X0 (local index of the initial state), s(0) = 0
    arma::Row<int> idx = set.State2Index(X0);
    int lo, hi;
    VecGetOwnershipRange(y[0], &lo, &hi);
    if (idx[0] >= lo && idx[0] < hi) VecSetValue(y[0], idx[0], 1.0, INSERT_VALUES);
    VecAssemblyBegin(y[0]);
    VecAssemblyEnd(y[0]);
  }
  auto rhs = [&](double t, std::vector<Vec> &in, std::vector<Vec> &out) {
    A.Action(t, in[0], out[0]);
    for (int i = 0; i < P; ++i) {
      A.Action(t, in[i + 1], out[i + 1]);
      A.SensAction(i, t, in[0], tmp);
      VecAXPY(out[i + 1], 1.0, tmp);
    }
  };
  const int    steps = 2000;
  const double h = T / steps;
  for (int s = 0; s < steps; ++s) {
    const double t = s * h;
    const double cw[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6}, cs[4] = {0.0, 0.5, 0.5, 1.0};
    for (int i = 0; i <= P; ++i) VecSet(acc[i], 0.0);
    for (int stage = 0; stage < 4; ++stage) {
      for (int i = 0; i <= P; ++i) {
        VecCopy(y[i], st[i]);
        if (stage > 0) VecAXPY(st[i], cs[stage] * h, k[i]);
      }
      rhs(t + cs[stage] * h, st, k);
      for (int i = 0; i <= P; ++i) VecAXPY(acc[i], cw[stage], k[i]);
    }
    for (int i = 0; i <= P; ++i) VecAXPY(y[i], h, acc[i]);
  }
  auto l1_first_n = [&](Vec a, Vec b) {  // 1-norm of the difference over the n state entries (the distribution has no sinks)
    const PetscScalar *pa, *pb;
    VecGetArrayRead(a, &pa);
    std::vector<double> ca(pa, pa + n);
    VecRestoreArrayRead(a, &pa);
    VecGetArrayRead(b, &pb);
    double e = 0.0;
    for (int q = 0; q < n; ++q) e += std::fabs(ca[q] - pb[q]);
    VecRestoreArrayRead(b, &pb);
    pacmensl_allreduce_sum(PETSC_COMM_WORLD, &e, 1);
    return e;
  };
  const double ep = l1_first_n(y[0], d.p_);
  double       es = 0.0, smax = 0.0;
  for (int i = 0; i < P; ++i) {
    es = std::max(es, l1_first_n(y[i + 1], d.dp_[i]));
    double nrm;
    VecNorm(d.dp_[i], NORM_1, &nrm);
    smax = std::max(smax, nrm);
  }
  std::printf("    fixed-set toggle sens (n = %d, t_f = %g): ||p_bdf - p_rk4||_1 = %.3e, max_i ||s_i,bdf - s_i,rk4||_1 = %.3e (max ||s_i||_1 = %.3e)\n",
              n, T, ep, es, smax);
  ASSERT_LE(ep, 1.0e-6);             // the reference's own bound for this solver (KAT-SF2: 1e-7 on p at rtol 1e-6 ... 1e-6 on s)
  ASSERT_LE(es, 1.0e-5 * (1.0 + smax));
  for (int i = 0; i <= P; ++i) { VecDestroy(&y[i]); VecDestroy(&k[i]); VecDestroy(&acc[i]); VecDestroy(&st[i]); }
  VecDestroy(&tmp);
}

class SensFspPoissonTest : public ::testing::Test {
 protected:
  void SetUp() override {
    auto propensity = [&](int, int, int num_states, const int *, PetscReal *output, void *) {
      for (int i{0}; i < num_states; ++i) output[i] = 1.0;
      return 0;
    };
    auto t_fun = [&](double, int, double *outputs, void *) { outputs[0] = lambda; return 0; };
    auto d_t_fun = [&](int, double, int, double *outputs, void *) { outputs[0] = 1.0; return 0; };
    poisson_model = SensModel(1, stoich_matrix, std::vector<int>({0}), t_fun, propensity, d_t_fun, {{0}}, nullptr);
  }
  SensModel            poisson_model;
  PetscReal            lambda = 2.0;
  arma::Mat<int>       stoich_matrix = {1};
  arma::Mat<int>       x0 = {0};
  arma::Col<PetscReal> p0 = {1.0};
  arma::Col<PetscReal> s0 = {0.0};
  arma::Row<int>       fsp_size = {5};
  arma::Row<PetscReal> expansion_factors = {0.1};
  PetscReal            t_final{1.0}, fsp_tol{1.0e-7};
};

TEST_F(SensFspPoissonTest, test_poisson_analytic) {
  SensFspSolverMultiSinks fsp(PETSC_COMM_WORLD);
  ASSERT_FALSE(fsp.SetModel(poisson_model));
  ASSERT_FALSE(fsp.SetInitialBounds(fsp_size));
  ASSERT_FALSE(fsp.SetExpansionFactors(expansion_factors));
  ASSERT_FALSE(fsp.SetInitialDistribution(x0, p0, std::vector<arma::Col<PetscReal>>({s0})));
  ASSERT_FALSE(fsp.SetUp());
  SensDiscreteDistribution p_final = fsp.Solve(t_final, fsp_tol);
  fsp.ClearState();

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
  std::printf("    Poisson sens: %d states, L1 error of p %.3e", num_states, stmp);
  ASSERT_LE(stmp, fsp_tol);

  PetscReal *s_dat;
  p_final.GetSensView(0, num_states, s_dat);
  double serr = 0.0;
  for (int i = 0; i < num_states; ++i) {
    int       n = p_final.states_(0, i);
    PetscReal pdf = -t_final * exp(-lambda * t_final) * pow(lambda * t_final, double(n)) / tgamma(n + 1);
    if (n > 0) pdf += exp(-lambda * t_final) * t_final * pow(lambda * t_final, double(n - 1)) / tgamma(n);
    serr += std::abs(s_dat[i] - pdf);
  }
  p_final.RestoreSensView(0, s_dat);
  pacmensl_allreduce_sum(PETSC_COMM_WORLD, &serr, 1);
  std::printf(", of dp/dlambda %.3e\n", serr);
  ASSERT_LE(stmp + serr, 1.0e-6);
}

namespace telegraph_cme {
arma::Mat<PetscInt> SM{{-1, 1, 0, 0}, {1, -1, 0, 0}, {0, 0, 1, -1}};
const double k01{1.0e-2}, k10{1.0e-1}, kr{10.0}, gamma{1.0};
int propensity(const int reaction, const int, const int num_states, const PetscInt *X, double *outputs, void *) {
  for (int i{0}; i < num_states; ++i) {
    const int *x = X + 3 * i;
    switch (reaction) {
      case 0: outputs[i] = k01 * x[0]; break;
      case 1: outputs[i] = k10 * x[1]; break;
      case 2: outputs[i] = kr * x[1]; break;
      case 3: outputs[i] = gamma * x[2]; break;
      default: return -1;
    }
  }
  return 0;
}
int propensity_derivatives(const int parameter_idx, const int reaction, const int, const int num_states, const PetscInt *X,
                           double *outputs, void *) {
  for (int i{0}; i < num_states; ++i) {
    const int *x = X + 3 * i;
    switch (parameter_idx) {
      case 0: if (reaction == 0) outputs[i] = x[0]; break;
      case 1: if (reaction == 1) outputs[i] = x[1]; break;
      case 2: outputs[i] = x[1]; break;
      case 3: outputs[i] = x[2]; break;
      default: return -1;
    }
  }
  return 0;
}
}  // namespace telegraph_cme

class SensFspTelegraphTest : public ::testing::Test {
 protected:
  void SetUp() override {
    int n_par = 4;
    t_final = 100.0;
    fsp_tol = 1.0e-10;
    X0 = X0.t();
    dp0 = std::vector<arma::Col<PetscReal>>(n_par, arma::Col<PetscReal>({0.0}));
    telegraph_model = SensModel(4, telegraph_cme::SM, {}, nullptr, telegraph_cme::propensity, nullptr, {},
                                telegraph_cme::propensity_derivatives, {{0}, {1}, {2}, {3}});
  }
  PetscReal t_final, fsp_tol;
  arma::Mat<PetscInt>  X0{1, 0, 0};
  arma::Col<PetscReal> p0 = {1.0};
  std::vector<arma::Col<PetscReal>> dp0;
  SensModel            telegraph_model;
  arma::Row<int>       fsp_size = {2, 2, 1};
  arma::Row<PetscReal> expansion_factors = {0.25, 0.25, 0.25};
};

TEST_F(SensFspTelegraphTest, telegraph_sens_solve_with_cvode) {
  PetscReal stmp;
  SensFspSolverMultiSinks fsp(PETSC_COMM_WORLD);
  ASSERT_FALSE(fsp.SetModel(telegraph_model));
  ASSERT_FALSE(fsp.SetInitialBounds(fsp_size));
  ASSERT_FALSE(fsp.SetExpansionFactors(expansion_factors));
  ASSERT_FALSE(fsp.SetInitialDistribution(X0, p0, dp0));
  ASSERT_FALSE(fsp.SetUp());
  SensDiscreteDistribution p_final_bdf = fsp.Solve(t_final, fsp_tol);
  ASSERT_FALSE(VecSum(p_final_bdf.p_, &stmp));
  std::printf("    telegraph sens: %d states, %d expansions, 1 - sum(p) = %.3e\n", (int) p_final_bdf.states_.n_cols,
              fsp.GetNumExpansions(), 1.0 - stmp);
  ASSERT_GE(stmp, 1.0 - fsp_tol);
  for (int i{0}; i < telegraph_model.num_parameters_; ++i) {
    ASSERT_FALSE(VecSum(p_final_bdf.dp_[i], &stmp));
    ASSERT_LE(std::abs(stmp), 1.0e-6);
  }
}

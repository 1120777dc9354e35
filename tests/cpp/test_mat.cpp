// Host-layer parity tests for the FSP operator, shaped after the reference's tests/test_mat.cpp:
// 1-d random walk on states 0..12, hop right with rate 2, hop left with rate 3 (x > 0).
//   KAT-M1  FspMatrixBase:        sum(A * 1) == -2 exactly                (reference test_mat.cpp:110-151)
//   KAT-M2  FspMatrixConstrained: sum(A * 1) == 0                          (:199-238)
//   KAT-M3/4 the same through CreateRHSJacobian/ComputeRHSJacobian + MatMult (:153-197, 240-287)
//   KAT-M5  Action(t, x) == J(t) x for time-varying coefficients at five times (:289-341)
#include "fsp_models.h"
#include "pacmensl_test_env.h"

using namespace pacmensl;

class MatrixTest : public ::testing::Test {
 protected:
  void SetUp() override {
    fsp_size = arma::Row<int>({12});
    t_fun = [&](double t, int, double *outputs, void *) {
      outputs[0] = 1.0 + t;
      outputs[1] = 1.0 + 0.5 * t;
      return 0;
    };
    propensity = [&](const int reaction, const int, const int num_states, const int *X, double *outputs, void *) {
      switch (reaction) {
        case 0: for (int i{0}; i < num_states; ++i) outputs[i] = rate_right * 1.0; break;
        case 1: for (int i{0}; i < num_states; ++i) outputs[i] = rate_left * (X[i] > 0); break;
        default: return -1;
      }
      return 0;
    };
    arma::Mat<PetscInt> X0(1, 1);
    X0.fill(0);
    state_set = new StateSetConstrained(PETSC_COMM_WORLD);
    ASSERT_FALSE(state_set->SetStoichiometryMatrix(stoichiometry));
    ASSERT_FALSE(state_set->SetShapeBounds(fsp_size));
    ASSERT_FALSE(state_set->SetUp());
    ASSERT_FALSE(state_set->AddStates(X0));
    ASSERT_FALSE(state_set->Expand());
  }
  void TearDown() override { delete state_set; }

  Vec make_vec(int n_local, double value) {
    Vec P;
    VecCreate(PETSC_COMM_WORLD, &P);
    VecSetSizes(P, n_local, PETSC_DECIDE);
    VecSetFromOptions(P);
    VecSet(P, value);
    VecSetUp(P);
    return P;
  }

  StateSetConstrained *state_set = nullptr;
  arma::Row<int>       fsp_size;
  const double         rate_right = 2.0, rate_left = 3.0;
  const arma::Mat<int> stoichiometry{1, -1};
  pacmensl::TcoefFun   t_fun;
  pacmensl::PropFun    propensity;
};

TEST_F(MatrixTest, mat_base_generation) {
  ASSERT_EQ(state_set->GetNumGlobalStates(), 13);
  FspMatrixBase A(PETSC_COMM_WORLD);
  int ierr = A.GenerateValues(*state_set, stoichiometry, std::vector<int>(), t_fun, propensity, std::vector<int>(), nullptr, nullptr);
  ASSERT_FALSE(ierr);
  Vec P = make_vec(state_set->GetNumLocalStates(), 1.0), Q;
  ASSERT_FALSE(VecDuplicate(P, &Q));
  ASSERT_FALSE(A.Action(0.0, P, Q));
  double Q_sum;
  ASSERT_FALSE(VecSum(Q, &Q_sum));
  ASSERT_DOUBLE_EQ(Q_sum, -1.0 * rate_right);
  VecDestroy(&P);
  VecDestroy(&Q);
}

static bool single_rank() {
  int size;
  MPI_Comm_size(PETSC_COMM_WORLD, &size);
  return size == 1;
}

TEST_F(MatrixTest, mat_base_jacobian) {
  if (!single_rank()) return;  // the assembled Jacobian (TsFsp path) is provided for single-rank problems only
  FspMatrixBase A(PETSC_COMM_WORLD);
  ASSERT_FALSE(A.GenerateValues(*state_set, stoichiometry, std::vector<int>(), t_fun, propensity, std::vector<int>(), nullptr, nullptr));
  Vec P = make_vec(state_set->GetNumLocalStates(), 1.0), Q;
  ASSERT_FALSE(VecDuplicate(P, &Q));
  Mat J;
  ASSERT_FALSE(A.CreateRHSJacobian(&J));
  ASSERT_FALSE(A.ComputeRHSJacobian(0.0, J));
  ASSERT_FALSE(MatMult(J, P, Q));
  double Q_sum;
  ASSERT_FALSE(VecSum(Q, &Q_sum));
  ASSERT_DOUBLE_EQ(Q_sum, -1.0 * rate_right);
  MatDestroy(&J);
  VecDestroy(&P);
  VecDestroy(&Q);
}

TEST_F(MatrixTest, mat_constrained_generate_values) {
  FspMatrixConstrained A(PETSC_COMM_WORLD);
  ASSERT_FALSE(A.GenerateValues(*state_set, stoichiometry, std::vector<int>(), t_fun, propensity, std::vector<int>(), nullptr, nullptr));
  if (single_rank()) ASSERT_EQ(A.GetNumLocalRows(), 14);
  Vec P = make_vec(A.GetNumLocalRows(), 1.0), Q;
  PetscInt n_global;
  VecGetSize(P, &n_global);
  ASSERT_EQ(n_global, 14);
  ASSERT_FALSE(VecDuplicate(P, &Q));
  ASSERT_FALSE(A.Action(0.0, P, Q));
  double Q_sum;
  ASSERT_FALSE(VecSum(Q, &Q_sum));
  ASSERT_DOUBLE_EQ(Q_sum, 0.0);
  // flops bookkeeping (FspMatrixBase.cpp:429-444 + FspMatrixConstrained.cpp:447-465): 37 nnz + 1 sink nnz
  PetscInt nflops;
  ASSERT_FALSE(A.GetLocalMVFlops(&nflops));
  double fl = nflops;
  pacmensl_allreduce_sum(PETSC_COMM_WORLD, &fl, 1);
  ASSERT_EQ((int) fl, 2 * 37 + 2 * 1);
  VecDestroy(&P);
  VecDestroy(&Q);
}

TEST_F(MatrixTest, mat_constrained_jacobian1) {
  if (!single_rank()) return;
  FspMatrixConstrained A(PETSC_COMM_WORLD);
  ASSERT_FALSE(A.GenerateValues(*state_set, stoichiometry, std::vector<int>(), t_fun, propensity, std::vector<int>(), nullptr, nullptr));
  Vec P = make_vec(A.GetNumLocalRows(), 1.0), Q;
  ASSERT_FALSE(VecDuplicate(P, &Q));
  Mat J;
  ASSERT_FALSE(A.CreateRHSJacobian(&J));
  ASSERT_FALSE(A.ComputeRHSJacobian(0.0, J));
  ASSERT_FALSE(MatMult(J, P, Q));
  double Q_sum;
  ASSERT_FALSE(VecSum(Q, &Q_sum));
  ASSERT_DOUBLE_EQ(Q_sum, 0.0);
  MatDestroy(&J);
  VecDestroy(&P);
  VecDestroy(&Q);
}

TEST_F(MatrixTest, mat_constrained_jacobian2) {
  if (!single_rank()) return;
  PetscRandom prand;
  ASSERT_FALSE(PetscRandomCreate(PETSC_COMM_WORLD, &prand));
  ASSERT_FALSE(PetscRandomSetType(prand, PETSCRAND));
  FspMatrixConstrained A(PETSC_COMM_WORLD);
  ASSERT_FALSE(A.GenerateValues(*state_set, stoichiometry, std::vector<int>({0, 1}), t_fun, propensity, std::vector<int>(), nullptr, nullptr));
  Vec x = make_vec(A.GetNumLocalRows(), 1.0), y, z;
  ASSERT_FALSE(VecDuplicate(x, &y));
  ASSERT_FALSE(VecDuplicate(x, &z));
  std::vector<PetscReal> t_test({0.0, 0.1, 0.2, 1.0, 10.0});
  Mat J;
  ASSERT_FALSE(A.CreateRHSJacobian(&J));
  PetscReal gap, maxerr = 0.0, ynorm;
  for (auto t : t_test) {
    VecSetRandom(x, prand);
    ASSERT_FALSE(A.ComputeRHSJacobian(t, J));
    ASSERT_FALSE(MatMult(J, x, y));
    ASSERT_FALSE(A.Action(t, x, z));
    VecNorm(y, NORM_2, &ynorm);
    VecAXPY(z, -1.0, y);
    VecNorm(z, NORM_2, &gap);
    // (the reference's own check is vacuous -- it never raises maxerr, test_mat.cpp:339; this one is real)
    if (gap / ynorm > maxerr) maxerr = gap / ynorm;
  }
  ASSERT_LE(maxerr, 1.0e-14);
  MatDestroy(&J);
  PetscRandomDestroy(&prand);
  VecDestroy(&x);
  VecDestroy(&y);
  VecDestroy(&z);
}

TEST_F(MatrixTest, action_before_generate_is_zero_and_destroy_allows_regeneration) {
  FspMatrixConstrained A(PETSC_COMM_WORLD);
  int rank, size;
  MPI_Comm_rank(PETSC_COMM_WORLD, &rank);
  MPI_Comm_size(PETSC_COMM_WORLD, &size);
  const int n_loc = state_set->GetNumLocalStates() + (rank == size - 1 ? 1 : 0);
  Vec x = make_vec(n_loc, 1.0), y = make_vec(n_loc, 5.0);
  ASSERT_FALSE(A.Action(0.0, x, y));  // FspMatrixBase.cpp:41
  double s;
  VecSum(y, &s);
  ASSERT_EQ(s, 0.0);
  ASSERT_FALSE(A.GenerateValues(*state_set, stoichiometry, std::vector<int>(), t_fun, propensity, std::vector<int>(), nullptr, nullptr));
  ASSERT_FALSE(A.Destroy());
  ASSERT_FALSE(A.GenerateValues(*state_set, stoichiometry, std::vector<int>(), t_fun, propensity, std::vector<int>(), nullptr, nullptr));
  ASSERT_FALSE(A.Action(0.0, x, y));
  VecSum(y, &s);
  ASSERT_DOUBLE_EQ(s, 0.0);
  // a failing time-coefficient callback propagates its code (FspMatrixBase.cpp:44-45)
  FspMatrixConstrained B(PETSC_COMM_WORLD);
  TcoefFun bad = [](double, int, double *, void *) { return -1; };
  ASSERT_FALSE(B.GenerateValues(*state_set, stoichiometry, std::vector<int>({0}), bad, propensity, std::vector<int>(), nullptr, nullptr));
  ASSERT_EQ(B.Action(0.0, x, y), -1);
  // a base state set is rejected by the constrained matrix (FspMatrixConstrained.cpp:133-135)
  StateSetBase plain(PETSC_COMM_WORLD);
  ASSERT_EQ(B.GenerateValues(plain, stoichiometry, std::vector<int>(), t_fun, propensity, std::vector<int>(), nullptr, nullptr), -1);
  VecDestroy(&x);
  VecDestroy(&y);
}


// Extension test: ActionFused (operator + solver epilogue in one kernel) against Action followed by separate vector
// passes, through the host classes exactly as the FSP driver uses them: hog1p (time-varying reaction, K = 5 sinks),
// incremental regeneration after a bound expansion, work vectors carved from one block (VecDuplicateVecs).
TEST(FusedAction, matches_action_plus_vector_passes_on_hog1p_with_expansion) {
  int rank, size;
  MPI_Comm_rank(PETSC_COMM_WORLD, &rank);
  MPI_Comm_size(PETSC_COMM_WORLD, &size);
  fsp_fixture f;
  ASSERT_EQ(fsp_fixture_get("hog1p", &f), 0);
  arma::Mat<int> SM(f.SM, f.num_species, f.num_reactions);
  Model          model(SM, f.prop_t, f.prop_x, nullptr, nullptr, std::vector<int>(f.tv_reactions, f.tv_reactions + f.num_tv));
  StateSetConstrained fss(PETSC_COMM_WORLD);
  arma::Row<int>      bounds(f.bounds, f.num_constr);
  fss.SetStoichiometryMatrix(SM);
  fss.SetShapeBounds(bounds);
  fss.SetUp();
  arma::Mat<int> X0(f.x0, f.num_species, 1);
  fss.AddStates(X0);
  ASSERT_FALSE(fss.Expand());
  FspMatrixConstrained A(PETSC_COMM_WORLD);
  A.SetIncrementalGeneration(true);
  for (int round = 0; round < 2; ++round) {
    if (round == 1) {  // the driver's expansion step
      for (arma::uword k = 1; k < bounds.n_elem; ++k) bounds[k] = (int) std::round(bounds[k] * 1.25 + 0.5);
      fss.SetShapeBounds(bounds);
      ASSERT_FALSE(fss.Expand());
      A.Destroy();
    }
    ASSERT_FALSE(A.GenerateValues(fss, model));
    const int n_loc = A.GetNumLocalRows();
    Vec proto;
    VecCreate(PETSC_COMM_WORLD, &proto);
    VecSetSizes(proto, n_loc, PETSC_DECIDE);
    VecSetUp(proto);
    Vec *W = nullptr;
    ASSERT_FALSE(VecDuplicateVecsUninitialized(proto, 6, &W));
    Vec x = W[0], y = W[1], z = W[2], v0 = W[3], ewt = W[4], tmp = W[5];
    PetscRandom prand;
    PetscRandomCreate(PETSC_COMM_WORLD, &prand);
    VecSetRandom(x, prand);
    VecSetRandom(v0, prand);
    VecSetRandom(ewt, prand);
    VecScale(ewt, 100.0);
    DeviceBuffer<double> out(2);
    for (double t : {0.0, 30.0}) {
      // Krylov form: y = A x, <y, v0>
      fspmat_epilogue ep{};
      ep.alpha = 1.0; ep.beta = 0.0; ep.scale_dev = nullptr; ep.n_dots = 1;
      ep.dot_vec_dev[0] = v0->d_data; ep.dot_out_dev = out.get();
      ASSERT_FALSE(A.ActionFused(t, x, y, ep));
      ASSERT_FALSE(A.Action(t, x, z));
      double dot_ref, ynorm, gap, dots[2];
      VecDot(z, v0, &dot_ref);
      VecNorm(z, NORM_2, &ynorm);
      VecAXPY(z, -1.0, y);
      VecNorm(z, NORM_2, &gap);
      ASSERT_LE(gap, 1e-14 * ynorm);
      out.download(dots, 2);
      if (size > 1) pacmensl_allreduce_sum(PETSC_COMM_WORLD, dots, 1);
      ASSERT_LE(std::fabs(dots[0] - dot_ref), 1e-11 * ynorm);
      // GMRES form: y = ewt .* (x - gamma A x), <y, v0>, <y, y>
      const double gamma = 0.37;
      ep.alpha = -gamma; ep.beta = 1.0; ep.scale_dev = ewt->d_data; ep.n_dots = 2; ep.dot_vec_dev[1] = nullptr;
      ASSERT_FALSE(A.ActionFused(t, x, y, ep));
      ASSERT_FALSE(A.Action(t, x, z));
      fspvec_wlincomb(tmp->d_data, ewt->d_data, 1.0, x->d_data, -gamma, z->d_data, n_loc, nullptr);
      double d0, d1;
      VecDot(tmp, v0, &d0);
      VecDot(tmp, tmp, &d1);
      VecNorm(tmp, NORM_2, &ynorm);
      VecAXPY(tmp, -1.0, y);
      VecNorm(tmp, NORM_2, &gap);
      ASSERT_LE(gap, 1e-14 * ynorm);
      out.download(dots, 2);
      if (size > 1) pacmensl_allreduce_sum(PETSC_COMM_WORLD, dots, 2);
      ASSERT_LE(std::fabs(dots[0] - d0), 1e-11 * ynorm * ynorm);
      ASSERT_LE(std::fabs(dots[1] - d1), 1e-11 * d1);
    }
    std::printf("    round %d: %d local rows, fused == separate passes\n", round, n_loc);
    PetscRandomDestroy(&prand);
    VecDestroyVecs(6, &W);
    VecDestroy(&proto);
  }
}

// Sensitivity-operator tests shaped after the reference's tests/test_sensmat.cpp: 1-d random walk, both reactions
// time-varying with c = (2, 3) and dc/dtheta_i = e_i.
//   KAT-SM1 base:        sum(A 1) = -2,  sum(SensAction(0) 1) = -1   (test_sensmat.cpp:138-183)
//   KAT-SM2 constrained: all sums 0                                   (:185-230)
#include "pacmensl_test_env.h"

using namespace pacmensl;

class SensMatrixTest : public ::testing::Test {
 protected:
  void SetUp() override {
    fsp_size = arma::Row<int>({12});
    t_fun = [&](double, int, double *outputs, void *) {
      outputs[0] = rate_right;
      outputs[1] = rate_left;
      return 0;
    };
    dt_fun = [&](int par_idx, double, int, double *outputs, void *) {
      switch (par_idx) {
        case 0: outputs[0] = 1.0; break;
        case 1: outputs[1] = 1.0; break;
        default: break;
      }
      return 0;
    };
    std::vector<std::vector<int>> dt_sp = {{0}, {1}};
    propensity = [&](const int reaction, const int, const int num_states, const int *X, double *outputs, void *) {
      switch (reaction) {
        case 0: for (int i{0}; i < num_states; ++i) outputs[i] = 1.0; break;
        case 1: for (int i{0}; i < num_states; ++i) outputs[i] = (X[i] > 0); break;
        default: return -1;
      }
      return 0;
    };
    arma::Mat<PetscInt> X0(1, 1);
    X0.fill(0);
    state_set = std::make_shared<StateSetConstrained>(PETSC_COMM_WORLD);
    ASSERT_FALSE(state_set->SetStoichiometryMatrix(stoichiometry));
    ASSERT_FALSE(state_set->SetShapeBounds(fsp_size));
    ASSERT_FALSE(state_set->SetUp());
    ASSERT_FALSE(state_set->AddStates(X0));
    ASSERT_FALSE(state_set->Expand());
    smodel = SensModel(2, stoichiometry, std::vector<int>({0, 1}), t_fun, propensity, dt_fun, dt_sp, nullptr, {});
  }
  Vec ones(int n) {
    Vec P;
    VecCreate(PETSC_COMM_WORLD, &P);
    VecSetSizes(P, n, PETSC_DECIDE);
    VecSetFromOptions(P);
    VecSet(P, 1.0);
    VecSetUp(P);
    return P;
  }
  std::shared_ptr<StateSetConstrained> state_set;
  arma::Row<int>       fsp_size;
  const double         rate_right = 2.0, rate_left = 3.0;
  const arma::Mat<int> stoichiometry{1, -1};
  TcoefFun  t_fun;
  DTcoefFun dt_fun;
  PropFun   propensity;
  SensModel smodel;
};

TEST_F(SensMatrixTest, mat_base_generation) {
  SensFspMatrix<FspMatrixBase> A(PETSC_COMM_WORLD);
  ASSERT_FALSE(A.GenerateValues(*state_set, smodel));
  Vec P = ones(state_set->GetNumLocalStates()), Q;
  ASSERT_FALSE(VecDuplicate(P, &Q));
  double Q_sum;
  ASSERT_FALSE(A.Action(0.0, P, Q));
  ASSERT_FALSE(VecSum(Q, &Q_sum));
  ASSERT_DOUBLE_EQ(Q_sum, -1.0 * rate_right);
  for (int i_par{0}; i_par < 2; ++i_par) {
    // d/dtheta_0 of sum(A 1) = -1 (mass leaving at the right end); theta_1 (left hops) conserves mass
    PetscReal dqsum = (i_par == 0) ? -1.0 : 0.0;
    ASSERT_FALSE(A.SensAction(i_par, 0.0, P, Q));
    ASSERT_FALSE(VecSum(Q, &Q_sum));
    ASSERT_DOUBLE_EQ(Q_sum, dqsum);
  }
  ASSERT_EQ(A.SensAction(2, 0.0, P, Q), -1);
  VecDestroy(&P);
  VecDestroy(&Q);
}

TEST_F(SensMatrixTest, mat_constr_generation) {
  SensFspMatrix<FspMatrixConstrained> A(PETSC_COMM_WORLD);
  ASSERT_FALSE(A.GenerateValues(*state_set, smodel));
  Vec P = ones(A.GetNumLocalRows()), Q;
  ASSERT_FALSE(VecDuplicate(P, &Q));
  double Q_sum;
  ASSERT_FALSE(A.Action(0.0, P, Q));
  ASSERT_FALSE(VecSum(Q, &Q_sum));
  ASSERT_DOUBLE_EQ(Q_sum, 0.0);
  for (int i_par{0}; i_par < 2; ++i_par) {
    ASSERT_FALSE(A.SensAction(i_par, 0.0, P, Q));
    ASSERT_FALSE(VecSum(Q, &Q_sum));
    ASSERT_DOUBLE_EQ(Q_sum, 0.0);
  }
  // finite-difference check of the sensitivity operator: (A(theta + h e_i) - A(theta - h e_i)) x / 2h == SensAction(i) x
  PetscRandom r;
  PetscRandomCreate(PETSC_COMM_WORLD, &r);
  VecSetRandom(P, r);
  Vec Qp, Qm;
  VecDuplicate(P, &Qp);
  VecDuplicate(P, &Qm);
  for (int i_par{0}; i_par < 2; ++i_par) {
    const double h = 1e-3;
    FspMatrixConstrained Ap(PETSC_COMM_WORLD), Am(PETSC_COMM_WORLD);
    TcoefFun tp = [&](double, int, double *o, void *) { o[0] = rate_right + (i_par == 0 ? h : 0); o[1] = rate_left + (i_par == 1 ? h : 0); return 0; };
    TcoefFun tm = [&](double, int, double *o, void *) { o[0] = rate_right - (i_par == 0 ? h : 0); o[1] = rate_left - (i_par == 1 ? h : 0); return 0; };
    ASSERT_FALSE(Ap.GenerateValues(*state_set, stoichiometry, {0, 1}, tp, propensity, {}, nullptr, nullptr));
    ASSERT_FALSE(Am.GenerateValues(*state_set, stoichiometry, {0, 1}, tm, propensity, {}, nullptr, nullptr));
    ASSERT_FALSE(Ap.Action(0.0, P, Qp));
    ASSERT_FALSE(Am.Action(0.0, P, Qm));
    VecAXPY(Qp, -1.0, Qm);
    VecScale(Qp, 1.0 / (2 * h));
    ASSERT_FALSE(A.SensAction(i_par, 0.0, P, Q));
    VecAXPY(Qp, -1.0, Q);
    double gap;
    VecNorm(Qp, NORM_INFINITY, &gap);
    ASSERT_LE(gap, 1e-10);
  }
  PetscRandomDestroy(&r);
  VecDestroy(&P); VecDestroy(&Q); VecDestroy(&Qp); VecDestroy(&Qm);
}

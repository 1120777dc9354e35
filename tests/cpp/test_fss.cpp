// State-set tests shaped after the reference's tests/test_fss.cpp (KAT-S1, KAT-S2).
#include "pacmensl_test_env.h"

using namespace pacmensl;

TEST(StateSetExpansion, toggle_state_set_insertion_error_handling) {
  arma::Mat<int> SM{{1, -1, 0, 0}, {0, 0, 1, -1}};
  arma::Mat<int> X0(3, 1);
  X0.col(0).fill(0);
  StateSetConstrained state_set(MPI_COMM_WORLD);
  ASSERT_EQ(state_set.SetStoichiometryMatrix(SM), 0);
  ASSERT_EQ(state_set.AddStates(X0), -1);
}

static void run_expansion(PartitioningType type) {
  arma::Mat<int> SM{{1, -1, 0, 0}, {0, 0, 1, -1}};
  arma::Mat<int> X0(2, 1);
  X0.col(0).fill(0);
  fsp_constr_multi_fn constr_fun = [&](int n_species, int n_constraints, int n_states, int *states, int *output, void *) {
    if (n_constraints != 1) return int(-1);
    if (n_species != 2) return int(-1);
    for (int i{0}; i < n_states; ++i) output[i] = states[2 * i] + states[2 * i + 1];
    return int(0);
  };
  StateSetConstrained state_set(MPI_COMM_WORLD);
  arma::Row<int>      fsp_size = {3};
  ASSERT_EQ(state_set.SetStoichiometryMatrix(SM), 0);
  ASSERT_EQ(state_set.SetLoadBalancingScheme(type), 0);
  ASSERT_EQ(state_set.SetShape(constr_fun, fsp_size), 0);
  ASSERT_EQ(state_set.AddStates(X0), 0);
  ASSERT_EQ(state_set.Expand(), 0);
  ASSERT_EQ(state_set.GetNumGlobalStates(), 10);
  int c{0};
  int all_states[2 * 10];
  int indx[10];
  for (int i{0}; i < 4; ++i)
    for (int j{0}; j < 4 - i; ++j) {
      all_states[c] = i;
      all_states[c + 1] = j;
      c += 2;
    }
  state_set.State2Index(10, all_states, indx);
  for (int i{0}; i < 10; ++i) ASSERT_GE(indx[i], 0);
  // indices are a permutation of 0..9
  std::vector<int> seen(10, 0);
  for (int i{0}; i < 10; ++i) seen[indx[i]]++;
  for (int i{0}; i < 10; ++i) ASSERT_EQ(seen[i], 1);
  // absent / negative states
  arma::Mat<int> bad{{4, -1, 2}, {0, 0, 2}};
  arma::Row<int> bi = state_set.State2Index(bad);
  for (int i = 0; i < 3; ++i) ASSERT_EQ(bi[i], -1);
}

TEST(StateSetExpansion, toggle_state_set_expansion_lb_naive) { run_expansion(PartitioningType::BLOCK); }
TEST(StateSetExpansion, toggle_state_set_expansion_lb_graph) { run_expansion(PartitioningType::GRAPH); }
TEST(StateSetExpansion, toggle_state_set_expansion_lb_hypergraph) { run_expansion(PartitioningType::HYPERGRAPH); }

TEST(StateSetExpansion, default_constraints_need_matching_bounds) {
  arma::Mat<int> SM{{1, -1, 0, 0}, {0, 0, 1, -1}};
  StateSetConstrained state_set(MPI_COMM_WORLD);
  arma::Row<int> three = {3, 3, 3};
  ASSERT_EQ(state_set.SetStoichiometryMatrix(SM), 0);
  ASSERT_EQ(state_set.SetShapeBounds(three), 0);
  ASSERT_EQ(state_set.SetUp(), -1);  // StateSetConstrained.cpp:227-233
}

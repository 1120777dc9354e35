// example_common.h -- shared driver of the three example programs (the counterparts of the reference's
// examples/repressilator.cpp, examples/hog1p.cpp and examples/transcr_reg_6d.cpp): solve the named workload with the
// adaptive FSP driver and report wall time to t_f, number of expansions, final N and number of Action calls.
//   usage: <example> [--solver cvode|krylov] [--constraints default|custom] [--tfinal T] [--verbosity 0|1|2]
#pragma once
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>

#include "fsp_models.h"
#include "fsp_models_device.h"
#include "pacmensl_all.h"

inline int run_fsp_example(int argc, char *argv[], const char *default_fixture, const char *custom_fixture) {
  using namespace pacmensl;
  std::setvbuf(stdout, nullptr, _IOLBF, 0);  // keep diagnostics if the solve throws
  Environment my_env(&argc, &argv, nullptr);
  std::string solver = "cvode", constraints = "default";
  double      t_final_override = -1.0;
  int         verbosity = 0;
  bool        log_events = false, cold = true, host_prop = false;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--log")) log_events = true;
    if (!std::strcmp(argv[i], "--host-propensities")) host_prop = true;  // evaluate prop_x through the host callback only
    if (!std::strcmp(argv[i], "--warm")) cold = false;  // carry the BDF history across expansions (default: re-create it, as the reference)
    if (!std::strcmp(argv[i], "--solver") && i + 1 < argc) solver = argv[++i];
    else if (!std::strcmp(argv[i], "--constraints") && i + 1 < argc) constraints = argv[++i];
    else if (!std::strcmp(argv[i], "--tfinal") && i + 1 < argc) t_final_override = std::atof(argv[++i]);
    else if (!std::strcmp(argv[i], "--verbosity") && i + 1 < argc) verbosity = std::atoi(argv[++i]);
  }
  const char *name = (constraints == "custom" && custom_fixture) ? custom_fixture : default_fixture;
  fsp_fixture f;
  if (fsp_fixture_get(name, &f)) { std::printf("unknown workload %s\n", name); return 1; }
  int rank, size;
  MPI_Comm_rank(PETSC_COMM_WORLD, &rank);
  MPI_Comm_size(PETSC_COMM_WORLD, &size);

  arma::Mat<int> SM(f.SM, f.num_species, f.num_reactions);
  Model model(SM, f.prop_t, f.prop_x, nullptr, nullptr, std::vector<int>(f.tv_reactions, f.tv_reactions + f.num_tv));
  // the reference's contract is the host callback prop_x; where the propensities are separable (hog1p, transcr_reg_6d)
  // the same values are described in device-evaluable form so that matrix generation never leaves the GPU
  const bool device_form = !host_prop && AttachDeviceForm(name, model);
  arma::Mat<int>       X0(f.x0, f.num_species, 1);
  arma::Col<PetscReal> p0 = {1.0};
  arma::Row<int>       bounds(f.bounds, f.num_constr);
  arma::Row<PetscReal> factors(f.expansion, f.num_constr);
  const double         t_final = t_final_override > 0 ? t_final_override : f.t_final;

  FspSolverMultiSinks fsp_solver(PETSC_COMM_WORLD, PartitioningType::BLOCK, solver == "krylov" ? KRYLOV : CVODE);
  fsp_solver.SetModel(model);
  fsp_solver.SetInitialDistribution(X0, p0);
  fsp_solver.SetInitialBounds(bounds);
  fsp_solver.SetExpansionFactors(factors);
  if (f.lhs) fsp_solver.SetConstraintFunctions(fsp_constr_multi_fn(f.lhs), nullptr);
  fsp_solver.SetOdeTolerances(f.rtol, f.atol);
  fsp_solver.SetVerbosity(verbosity);
  fsp_solver.SetFromOptions();
  fsp_solver.SetWarmRestart(!cold);
  if (log_events) fsp_solver.SetLogging(PETSC_TRUE);

  auto t0 = std::chrono::steady_clock::now();
  fsp_solver.SetUp();
  DiscreteDistribution solution = fsp_solver.Solve(t_final, f.fsp_tol, 0.0);
  fsp_device_sync();
  double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  PetscReal psum;
  VecSum(solution.p_, &psum);
  auto fss = std::static_pointer_cast<const StateSetConstrained>(fsp_solver.GetStateSet());
  arma::Row<int> final_bounds = fss->GetShapeBounds();
  if (rank == 0) {
    std::printf("{\"example\": \"%s\", \"solver\": \"%s\", \"device_propensities\": %s, \"ranks\": %d, \"t_final\": %g, \"fsp_tol\": %g, \"wall_s\": %.4f, "
                "\"expansions\": %d, \"warm_restarts\": %d, \"final_states\": %d, \"action_calls\": %ld, \"sum_p\": %.12f, \"final_bounds\": [",
                name, solver.c_str(), device_form ? "true" : "false", size, t_final, f.fsp_tol, wall, fsp_solver.GetNumExpansions(), fsp_solver.GetNumWarmRestarts(), fss->GetNumGlobalStates(),
                fsp_solver.GetNumRhsEvals(), psum);
    for (arma::uword k = 0; k < final_bounds.n_elem; ++k) std::printf("%s%d", k ? ", " : "", final_bounds[k]);
    std::printf("]}\n");
  }
  if (log_events) {
    // the counterpart of the reference's output_performance() (examples/repressilator.cpp:349-402)
    char op[] = "max";
    FspSolverComponentTiming tm = fsp_solver.ReduceComponentTiming(op);
    if (rank == 0)
      std::printf("{\"timing_s\": {\"total\": %.3f, \"state_expansion\": %.3f, \"matrix_generation\": %.3f, \"ode_solve\": %.3f, "
                  "\"rhs_launch\": %.3f, \"solution_scatter\": %.3f}, \"flops\": %.3e}\n",
                  tm.TotalTime, tm.StatePartitioningTime, tm.MatrixGenerationTime, tm.ODESolveTime, tm.RHSEvalTime,
                  tm.SolutionScatterTime, tm.TotalFlops);
  }
  // first marginal (what the reference examples write to disk)
  arma::Col<PetscReal> md = Compute1DMarginal(solution, 0);
  if (rank == 0 && verbosity > 0) {
    std::printf("marginal of species 0:");
    for (arma::uword i = 0; i < md.n_elem && i < 12; ++i) std::printf(" %.4e", md[i]);
    std::printf("\n");
  }
  fsp_solver.ClearState();
  return 0;
}

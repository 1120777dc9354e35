// Solve-to-t_f on the synthetic 3-D birth-death lattice (BASELINE config 4):
//   S = 3, SM = [+e1,-e1,+e2,-e2,+e3,-e3], births (40,30,20), deaths gamma x, gamma = (1.0,1.5,2.0), K = 3 box sinks,
//   p0 = delta(0,0,0), t_f = 1.0, KrylovFsp defaults or CvodeFsp (rtol 1e-6, atol 1e-14) on a FIXED state set
//   (edge^3 states), one rank per GPU (tools/launch_ranks.sh N build/examples/lattice_solve ...).
// The three species are independent M/M/inf queues started empty, so p(t_f) is the product of three Poisson
// pmfs with means (b_s/gamma_s)(1 - exp(-gamma_s t_f)); the 1-norm error against it is reported.
//   usage: lattice_solve [--edge 215] [--solver krylov|cvode] [--tfinal 1.0] [--rtol 1e-6] [--atol 1e-14] [--repeat 1] [--no-fused]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "fsp_models.h"
#include "pacmensl_all.h"

using namespace pacmensl;

int main(int argc, char *argv[]) {
  Environment my_env(&argc, &argv, nullptr);
  int         edge = 215, repeat = 1;
  std::string solver = "krylov";
  double      t_final = 1.0, rtol = 1.0e-6, atol = 1.0e-14;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--edge") && i + 1 < argc) edge = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--solver") && i + 1 < argc) solver = argv[++i];
    else if (!std::strcmp(argv[i], "--tfinal") && i + 1 < argc) t_final = std::atof(argv[++i]);
    else if (!std::strcmp(argv[i], "--rtol") && i + 1 < argc) rtol = std::atof(argv[++i]);
    else if (!std::strcmp(argv[i], "--atol") && i + 1 < argc) atol = std::atof(argv[++i]);
    else if (!std::strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = std::atoi(argv[++i]);
  }
  int rank, size;
  MPI_Comm_rank(PETSC_COMM_WORLD, &rank);
  MPI_Comm_size(PETSC_COMM_WORLD, &size);

  fsp_fixture f;
  if (fsp_fixture_get("birth_death_3d", &f)) return 1;
  arma::Mat<int> SM(f.SM, f.num_species, f.num_reactions);
  Model          model(SM, f.prop_t, f.prop_x, nullptr, nullptr, std::vector<int>());
  const std::vector<double> rates = {40.0, 1.0, 30.0, 1.5, 20.0, 2.0};
  arma::Mat<int>            orders(3, 6);
  orders.zeros();
  orders(0, 1) = 1; orders(1, 3) = 1; orders(2, 5) = 1;
  model.SetMassAction(rates, orders);

  auto t0 = std::chrono::steady_clock::now();
  StateSetConstrained fsp(PETSC_COMM_WORLD);
  arma::Row<int>      upper = {edge - 1, edge - 1, edge - 1};
  fsp.SetStoichiometryMatrix(SM);
  fsp.SetShapeBounds(upper);
  fsp.SetUp();
  if (fsp.AddBoxLattice(upper)) return 1;
  FspMatrixConstrained A(PETSC_COMM_WORLD);
  if (A.GenerateValues(fsp, model)) return 1;
  fsp_device_sync();
  const double t_build = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  auto AV = [&](PetscReal t, Vec x, Vec y) { return A.Action(t, x, y); };
  auto AVF = [&](PetscReal t, Vec x, Vec y, const fspmat_epilogue &ep) { return A.ActionFused(t, x, y, ep); };
  bool fused = true;
  int  verbose = 0;
  for (int i = 1; i < argc; ++i) if (!std::strcmp(argv[i], "--no-fused")) fused = false;
  for (int i = 1; i < argc; ++i) if (!std::strcmp(argv[i], "--verbose")) verbose = 1;
  double wall_best = 1e300, setup_best = 1e300, psum = 0.0, l1err = -1.0;
  long   nrhs = 0;
  int    stat = 0;
  for (int rep = 0; rep < repeat; ++rep) {
    Vec P;
    VecCreate(PETSC_COMM_WORLD, &P);
    VecSetSizes(P, A.GetNumLocalRows(), PETSC_DECIDE);
    VecSetUp(P);
    VecSetValue(P, 0, 1.0, INSERT_VALUES);  // state (0,0,0) has global index 0
    VecAssemblyBegin(P);
    VecAssemblyEnd(P);
    fsp_device_sync();
    MPI_Barrier(PETSC_COMM_WORLD);
    auto   t1 = std::chrono::steady_clock::now();
    double t_setup = 0.0;
    if (solver == "cvode") {
      CvodeFsp ode(PETSC_COMM_WORLD, CV_BDF);
      ode.SetFinalTime(t_final); ode.SetInitialSolution(&P); ode.SetRhs(AV); ode.SetTolerances(rtol, atol);
      if (fused) ode.SetFusedRhs(AVF);
      ode.SetStatusOutput(verbose);
      if (ode.SetUp()) return 1;
      fsp_device_sync();
      t_setup = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
      stat = ode.Solve();
      nrhs = ode.GetNumRhsEvals();
      ode.FreeWorkspace();
    } else {
      KrylovFsp ode(PETSC_COMM_WORLD);
      ode.SetFinalTime(t_final); ode.SetInitialSolution(&P); ode.SetRhs(AV); ode.SetFspMatPtr(&A);
      if (fused) ode.SetFusedRhs(AVF);
      ode.SetTolerances(rtol, atol);
      ode.SetStatusOutput(verbose);
      if (ode.SetUp()) return 1;
      fsp_device_sync();
      t_setup = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
      stat = ode.Solve();
      nrhs = ode.GetNumRhsEvals();
      ode.FreeWorkspace();
    }
    fsp_device_sync();
    MPI_Barrier(PETSC_COMM_WORLD);
    double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    if (wall < wall_best) { wall_best = wall; setup_best = t_setup; }
    VecSum(P, &psum);
    if (rep == repeat - 1) {
      // 1-norm error against the product of Poisson pmfs (local block, then summed over ranks)
      const double b[3] = {40.0, 30.0, 20.0}, g[3] = {1.0, 1.5, 2.0};
      std::vector<std::vector<double>> pm(3, std::vector<double>(edge));
      for (int s = 0; s < 3; ++s) {
        const double lam = b[s] / g[s] * (1.0 - std::exp(-g[s] * t_final));
        for (int k = 0; k < edge; ++k) pm[s][k] = std::exp(-lam + k * std::log(lam) - std::lgamma(k + 1.0));
      }
      const PetscScalar *pa;
      VecGetArrayRead(P, &pa);
      const long n = fsp.GetNumLocalStates(), start = fsp.GetLocalStart();
      double     e = 0.0;
      for (long i = 0; i < n; ++i) {
        long gidx = start + i;
        int  x0 = (int) (gidx % edge), x1 = (int) ((gidx / edge) % edge), x2 = (int) (gidx / ((long) edge * edge));
        e += std::fabs(pa[i] - pm[0][x0] * pm[1][x1] * pm[2][x2]);
      }
      VecRestoreArrayRead(P, &pa);
      pacmensl_allreduce_sum(PETSC_COMM_WORLD, &e, 1);
      l1err = e;
    }
    VecDestroy(&P);
  }
  const double bytes = A.GetActionBytes();
  double       bytes_tot = bytes;
  pacmensl_allreduce_sum(PETSC_COMM_WORLD, &bytes_tot, 1);
  if (rank == 0) {
    std::printf("{\"example\": \"lattice_solve\", \"solver\": \"%s\", \"ranks\": %d, \"edge\": %d, \"states\": %d, \"t_final\": %g, "
                "\"rtol\": %g, \"atol\": %g, \"status\": %d, \"wall_s\": %.4f, \"of_which_solver_setup_s\": %.4f, \"build_s\": %.3f, \"action_calls\": %ld, "
                "\"us_per_action_incl_vector_ops\": %.2f, \"action_GBps_equiv\": %.1f, \"sum_p\": %.12f, \"l1_err_vs_poisson\": %.3e}\n",
                solver.c_str(), size, edge, fsp.GetNumGlobalStates(), t_final, rtol, atol, stat, wall_best, setup_best, t_build, nrhs,
                1e6 * wall_best / (double) std::max(1L, nrhs), bytes_tot * (double) nrhs / wall_best / 1e9, psum, l1err);
  }
  return stat == 0 ? 0 : 1;
}

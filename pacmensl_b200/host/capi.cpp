// capi.cpp -- C ABI over the host classes (include/pacmensl_b200_host.h).
#include <cstdarg>
#include <cstring>

#include "../../include/pacmensl_b200_host.h"
#include "fsp_models.h"
#include "fsp_models_device.h"
#include "pacmensl_all.h"

using namespace pacmensl;

namespace {
thread_local char g_err[512] = "";
void set_err(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
#define PFSP_TRY try {
#define PFSP_CATCH                                    \
  }                                                   \
  catch (std::exception & e) {                        \
    set_err("%s", e.what());                          \
    return -2;                                        \
  }                                                   \
  catch (...) {                                       \
    set_err("unknown C++ exception");                 \
    return -2;                                        \
  }

struct ModelBox {
  Model model;
};
struct SolverBox {
  std::unique_ptr<FspSolverMultiSinks> solver;
  DiscreteDistribution                 result;
  ODESolverType                        type = CVODE;
};
arma::Mat<int> colmajor(const int *p, int rows, int cols) { return arma::Mat<int>(p, (arma::uword) rows, (arma::uword) cols); }
}  // namespace

extern "C" {

const char *pfsp_last_error(void) { return g_err[0] ? g_err : fsp_last_error(); }

int pfsp_init(int device, const char *nccl_id, int rank, int size) {
  if (fsp_device_set(device)) return -1;
  if (size > 1) return pacmensl_comm_world_init(nccl_id, rank, size);
  return 0;
}
int pfsp_finalize(void) { return pacmensl_comm_world_finalize(); }
int pfsp_p2p_enabled(void) { MPI_Comm w = MPI_COMM_WORLD; return (w && w->nccl) ? fspcomm_p2p_enabled(w->nccl) : 0; }
void *pfsp_world_comm(void) { MPI_Comm w = MPI_COMM_WORLD; return w ? (void *) w->nccl : nullptr; }
int pfsp_check(void) {
  if (fsp_device_sync()) return -1;
  MPI_Comm w = MPI_COMM_WORLD;
  return (w && w->nccl) ? fspcomm_check(w->nccl) : 0;
}

// ---- state set ----
int pfsp_set_create(void **set) {
  PFSP_TRY
  *set = new StateSetConstrained(MPI_COMM_WORLD);
  return 0;
  PFSP_CATCH
}
int pfsp_set_destroy(void *set) { delete static_cast<StateSetConstrained *>(set); return 0; }
int pfsp_set_set_sharded(void *set, int on) { return static_cast<StateSetConstrained *>(set)->SetSharded(on != 0); }
int pfsp_set_is_sharded(void *set) { return static_cast<StateSetConstrained *>(set)->IsSharded() ? 1 : 0; }
int pfsp_set_remember_local(void *set) { return static_cast<StateSetConstrained *>(set)->RememberLocalStates(); }
int pfsp_set_remembered_indices(void *set, int n, int *idx) {
  PFSP_TRY
  std::vector<int> v;
  int ierr = static_cast<StateSetConstrained *>(set)->RememberedIndices(v);
  if (ierr) return ierr;
  if ((int) v.size() != n) return -1;
  std::copy(v.begin(), v.end(), idx);
  return 0;
  PFSP_CATCH
}
int pfsp_set_stoichiometry(void *set, int S, int R, const int *SM) {
  return static_cast<StateSetConstrained *>(set)->SetStoichiometryMatrix(colmajor(SM, S, R));
}
int pfsp_set_shape(void *set, int K, const int *bounds, pfsp_constr_fn lhs, void *args) {
  auto *s = static_cast<StateSetConstrained *>(set);
  if (!lhs) return s->SetShapeBounds(K, const_cast<int *>(bounds));
  fsp_constr_multi_fn f = [lhs](int a, int b, int c, int *st, int *out, void *ar) { return lhs(a, b, c, st, out, ar); };
  return s->SetShape(K, f, const_cast<int *>(bounds), args);
}
int pfsp_set_shape_bounds(void *set, int K, const int *bounds) {
  return static_cast<StateSetConstrained *>(set)->SetShapeBounds(K, const_cast<int *>(bounds));
}
int pfsp_set_add_states(void *set, int S, int m, const int *X) {
  PFSP_TRY
  return static_cast<StateSetConstrained *>(set)->AddStates(colmajor(X, S, m));
  PFSP_CATCH
}
int pfsp_set_add_box_lattice(void *set, int S, const int *upper) {
  PFSP_TRY
  return static_cast<StateSetConstrained *>(set)->AddBoxLattice(arma::Row<int>(upper, (arma::uword) S));
  PFSP_CATCH
}
int pfsp_set_expand(void *set) {
  PFSP_TRY
  return static_cast<StateSetConstrained *>(set)->Expand();
  PFSP_CATCH
}
int pfsp_set_sizes(void *set, int *n_local, int *n_global, int *start) {
  auto *s = static_cast<StateSetConstrained *>(set);
  if (n_local) *n_local = s->GetNumLocalStates();
  if (n_global) *n_global = s->GetNumGlobalStates();
  if (start) *start = s->GetLocalStart();
  return 0;
}
int pfsp_set_copy_states(void *set, int *out) {
  PFSP_TRY
  auto *s = static_cast<StateSetConstrained *>(set);
  s->CopyStatesOnProc(s->GetNumLocalStates(), out);
  return 0;
  PFSP_CATCH
}
int pfsp_set_state2index(void *set, int m, const int *X, int *idx) {
  PFSP_TRY
  static_cast<StateSetConstrained *>(set)->State2Index(m, X, idx);
  return 0;
  PFSP_CATCH
}

// ---- model ----
int pfsp_model_create(void **model, int S, int R, const int *SM, pfsp_prop_fn prop_x, void *px_args, pfsp_tcoef_fn prop_t,
                      void *pt_args, int n_tv, const int *tv) {
  PFSP_TRY
  auto *b = new ModelBox();
  PropFun  px = nullptr;
  TcoefFun pt = nullptr;
  if (prop_x) px = [prop_x](const int r, const int s, const int m, const int *x, double *o, void *a) { return prop_x(r, s, m, x, o, a); };
  if (prop_t) pt = [prop_t](double t, int n, double *o, void *a) { return prop_t(t, n, o, a); };
  b->model = Model(colmajor(SM, S, R), pt, px, pt_args, px_args, std::vector<int>(tv, tv + n_tv));
  *model = b;
  return 0;
  PFSP_CATCH
}
int pfsp_model_from_fixture(void **model, const char *name, int *S, int *R, int *K, int *bounds, double *expansion, int *x0,
                            double *t_final, double *fsp_tol, double *rtol, double *atol, pfsp_constr_fn *lhs) {
  fsp_fixture f;
  if (fsp_fixture_get(name, &f)) { set_err("unknown fixture %s", name); return -1; }
  int ierr = pfsp_model_create(model, f.num_species, f.num_reactions, f.SM, f.prop_x, nullptr, f.prop_t, nullptr, f.num_tv,
                               f.tv_reactions);
  if (ierr) return ierr;
  if (S) *S = f.num_species;
  if (R) *R = f.num_reactions;
  if (K) *K = f.num_constr;
  for (int k = 0; k < f.num_constr; ++k) { if (bounds) bounds[k] = f.bounds[k]; if (expansion) expansion[k] = f.expansion[k]; }
  for (int s = 0; s < f.num_species; ++s) if (x0) x0[s] = f.x0[s];
  if (t_final) *t_final = f.t_final;
  if (fsp_tol) *fsp_tol = f.fsp_tol;
  if (rtol) *rtol = f.rtol;
  if (atol) *atol = f.atol;
  if (lhs) *lhs = f.lhs;
  return 0;
}
int pfsp_model_set_mass_action(void *model, const double *rates, const int *orders) {
  auto *b = static_cast<ModelBox *>(model);
  const int S = (int) b->model.stoichiometry_matrix_.n_rows, R = (int) b->model.stoichiometry_matrix_.n_cols;
  b->model.SetMassAction(std::vector<double>(rates, rates + R), colmajor(orders, S, R));
  return 0;
}
int pfsp_model_set_factor_table(void *model, int species, int reaction, int len, const double *values) {
  static_cast<ModelBox *>(model)->model.SetFactorTable(species, reaction, std::vector<double>(values, values + len));
  return 0;
}
int pfsp_model_attach_device_form(void *model, const char *fixture_name) {
  return AttachDeviceForm(fixture_name, static_cast<ModelBox *>(model)->model) ? 0 : 1;
}
int pfsp_model_get_stoichiometry(void *model, int *SM_colmajor) {
  const arma::Mat<int> &SM = static_cast<ModelBox *>(model)->model.stoichiometry_matrix_;
  std::memcpy(SM_colmajor, SM.memptr(), sizeof(int) * SM.n_elem);
  return 0;
}
int pfsp_model_destroy(void *model) { delete static_cast<ModelBox *>(model); return 0; }

// ---- operator ----
int pfsp_mat_create(void **mat, int constrained) {
  PFSP_TRY
  *mat = constrained ? static_cast<FspMatrixBase *>(new FspMatrixConstrained(MPI_COMM_WORLD)) : new FspMatrixBase(MPI_COMM_WORLD);
  return 0;
  PFSP_CATCH
}
int pfsp_mat_destroy(void *mat) { delete static_cast<FspMatrixBase *>(mat); return 0; }
int pfsp_mat_generate(void *mat, void *set, void *model) {
  PFSP_TRY
  return static_cast<FspMatrixBase *>(mat)->GenerateValues(*static_cast<StateSetConstrained *>(set), static_cast<ModelBox *>(model)->model);
  PFSP_CATCH
}
int pfsp_mat_clear(void *mat) { return static_cast<FspMatrixBase *>(mat)->Destroy(); }
int pfsp_mat_set_variant(void *mat, int variant) { static_cast<FspMatrixBase *>(mat)->SetKernelVariant(variant); return 0; }
int pfsp_mat_info(void *mat, int *n_rows_local, long *flops, double *bytes) {
  auto *A = static_cast<FspMatrixBase *>(mat);
  if (n_rows_local) *n_rows_local = A->GetNumLocalRows();
  if (flops) { PetscInt f = 0; A->GetLocalMVFlops(&f); *flops = f; }
  if (bytes) *bytes = A->GetActionBytes();
  return 0;
}
int pfsp_mat_action(void *mat, double t, const double *x_dev, double *y_dev) {
  PFSP_TRY
  auto  *A = static_cast<FspMatrixBase *>(mat);
  _p_Vec x, y;
  x.comm = y.comm = MPI_COMM_WORLD;
  x.n_local = y.n_local = A->GetNumLocalRows();
  x.d_data = const_cast<double *>(x_dev);
  y.d_data = y_dev;
  x.owns_data = y.owns_data = false;
  return A->Action(t, &x, &y);
  PFSP_CATCH
}
int pfsp_mat_halo_only(void *mat, const double *x_dev, double *y_dev, long *bytes_sent) {
  PFSP_TRY
  auto  *A = static_cast<FspMatrixBase *>(mat);
  _p_Vec x, y;
  x.comm = y.comm = MPI_COMM_WORLD;
  x.n_local = y.n_local = A->GetNumLocalRows();
  x.d_data = const_cast<double *>(x_dev);
  y.d_data = y_dev;
  x.owns_data = y.owns_data = false;
  return A->HaloExchangeOnly(&x, &y, bytes_sent);
  PFSP_CATCH
}
int pfsp_mat_action_host(void *mat, double t, const double *x_host, double *y_host) {
  PFSP_TRY
  return static_cast<FspMatrixBase *>(mat)->ActionHost(t, x_host, y_host);
  PFSP_CATCH
}

// ---- driver ----
int pfsp_solver_create(void **solver, int ode_type) {
  PFSP_TRY
  auto *b = new SolverBox();
  b->type = ode_type == 0 ? KRYLOV : (ode_type == 1 ? CVODE : PETSC);
  b->solver.reset(new FspSolverMultiSinks(MPI_COMM_WORLD, PartitioningType::BLOCK, b->type));
  *solver = b;
  return 0;
  PFSP_CATCH
}
int pfsp_solver_destroy(void *solver) { delete static_cast<SolverBox *>(solver); return 0; }
int pfsp_solver_set_model(void *solver, void *model) {
  return static_cast<SolverBox *>(solver)->solver->SetModel(static_cast<ModelBox *>(model)->model);
}
int pfsp_solver_set_initial_bounds(void *solver, int K, const int *bounds) {
  arma::Row<int> b(bounds, (arma::uword) K);
  return static_cast<SolverBox *>(solver)->solver->SetInitialBounds(b);
}
int pfsp_solver_set_constraint_function(void *solver, pfsp_constr_fn lhs, void *args) {
  fsp_constr_multi_fn f = [lhs](int a, int b, int c, int *st, int *out, void *ar) { return lhs(a, b, c, st, out, ar); };
  return static_cast<SolverBox *>(solver)->solver->SetConstraintFunctions(f, args);
}
int pfsp_solver_set_expansion_factors(void *solver, int K, const double *factors) {
  arma::Row<double> f(factors, (arma::uword) K);
  return static_cast<SolverBox *>(solver)->solver->SetExpansionFactors(f);
}
int pfsp_solver_set_initial_distribution(void *solver, int S, int m, const int *X, const double *p) {
  return static_cast<SolverBox *>(solver)->solver->SetInitialDistribution(colmajor(X, S, m), arma::Col<double>(p, (arma::uword) m));
}
int pfsp_solver_set_ode_tolerances(void *solver, double rtol, double atol) {
  return static_cast<SolverBox *>(solver)->solver->SetOdeTolerances(rtol, atol);
}
int pfsp_solver_set_verbosity(void *solver, int level) { return static_cast<SolverBox *>(solver)->solver->SetVerbosity(level); }
int pfsp_solver_set_krylov(void *solver, int q_iop, int m_min, int m_max) {
  auto *s = static_cast<SolverBox *>(solver)->solver.get();
  int   ierr = s->SetKrylovOrthLength(q_iop);
  if (ierr) return ierr;
  return s->SetKrylovDimRange(m_min, m_max);
}
int pfsp_solver_set_warm_restart(void *solver, int on) { return static_cast<SolverBox *>(solver)->solver->SetWarmRestart(on != 0); }
int pfsp_solver_set_sharded_state_set(void *solver, int on) { return static_cast<SolverBox *>(solver)->solver->SetShardedStateSet(on != 0); }
int pfsp_solver_num_warm_restarts(void *solver, int *n) { *n = static_cast<SolverBox *>(solver)->solver->GetNumWarmRestarts(); return 0; }
int pfsp_solver_setup(void *solver) {
  PFSP_TRY
  return static_cast<SolverBox *>(solver)->solver->SetUp();
  PFSP_CATCH
}
int pfsp_solver_solve(void *solver, double t_final, double fsp_tol, double t_init, int *n_local, int *n_species) {
  PFSP_TRY
  auto *b = static_cast<SolverBox *>(solver);
  b->result = b->solver->Solve(t_final, fsp_tol, t_init);
  if (n_local) *n_local = (int) b->result.states_.n_cols;
  if (n_species) *n_species = (int) b->result.states_.n_rows;
  return 0;
  PFSP_CATCH
}
int pfsp_solver_copy_result(void *solver, int *states, double *p) {
  PFSP_TRY
  auto *b = static_cast<SolverBox *>(solver);
  if (states) std::memcpy(states, b->result.states_.memptr(), sizeof(int) * b->result.states_.n_elem);
  if (p) {
    const PetscScalar *a;
    int ierr = VecGetArrayRead(b->result.p_, &a);
    if (ierr) return ierr;
    std::memcpy(p, a, sizeof(double) * b->result.states_.n_cols);
    VecRestoreArrayRead(b->result.p_, &a);
  }
  return 0;
  PFSP_CATCH
}
int pfsp_solver_stats(void *solver, int *n_global_states, int *n_expansions, long *n_rhs_evals, int *K, int *final_bounds) {
  auto *b = static_cast<SolverBox *>(solver);
  auto  ss = std::static_pointer_cast<const StateSetConstrained>(b->solver->GetStateSet());
  if (n_global_states) *n_global_states = ss ? ss->GetNumGlobalStates() : 0;
  if (n_expansions) *n_expansions = b->solver->GetNumExpansions();
  if (n_rhs_evals) *n_rhs_evals = b->solver->GetNumRhsEvals();
  if (ss) {
    arma::Row<int> bd = ss->GetShapeBounds();
    if (K) *K = (int) bd.n_elem;
    if (final_bounds) for (arma::uword k = 0; k < bd.n_elem; ++k) final_bounds[k] = bd[k];
  }
  return 0;
}
int pfsp_solver_clear(void *solver) { return static_cast<SolverBox *>(solver)->solver->ClearState(); }

}  // extern "C"

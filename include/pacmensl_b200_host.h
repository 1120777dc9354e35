/*
 * pacmensl_b200_host.h -- C ABI over the host-side mirror of the reference's operator/solver classes
 * (pacmensl::StateSetConstrained, FspMatrixBase/FspMatrixConstrained, FspSolverMultiSinks).  This is what a
 * foreign-language binding (ctypes in pacmensl_b200/api.py, bench.py) calls; C++ users include
 * pacmensl_b200/host/pacmensl_all.h and use the classes directly, exactly like the reference's own API.
 *
 * Handles are opaque.  Every function returns 0 on success, non-zero on failure (reference convention);
 * exceptions thrown by the C++ layer (PACMENSLCHKERRTHROW) are caught and reported as -2 with the message
 * available from pfsp_last_error().
 *
 * Reference interfaces replaced:
 *   pfsp_set_*     src/StateSet/StateSetBase.h:61-209, src/StateSet/StateSetConstrained.h:35-68
 *   pfsp_model_*   src/Models/Model.h:44-99
 *   pfsp_mat_*     src/Matrix/FspMatrixBase.h:53-194, src/Matrix/FspMatrixConstrained.h:35-80
 *   pfsp_solver_*  src/Fsp/FspSolverMultiSinks.h:65-335
 */
#ifndef PACMENSL_B200_HOST_H_
#define PACMENSL_B200_HOST_H_

#include "fsp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef int (*pfsp_prop_fn)(int reaction, int num_species, int num_states, const int *states, double *out, void *args);
typedef int (*pfsp_tcoef_fn)(double t, int num_coefs, double *out, void *args);
typedef int (*pfsp_constr_fn)(int num_species, int num_constr, int num_states, int *states, int *out, void *args);

/* process set-up: select the device and (size > 1) join the NCCL world with the id made by fspcomm_unique_id on rank 0 */
FSP_API int         pfsp_init(int device, const char *nccl_id, int rank, int size);
FSP_API int         pfsp_finalize(void);
/* 1 when the world communicator uses the peer-memory (CUDA IPC over NVLink) fast path, 0 = NCCL path / single rank */
FSP_API int         pfsp_p2p_enabled(void);
/* Synchronises the device and returns non-zero (message in fsp_last_error) if a device-side wait for a peer GPU has timed
 * out since the communicator was created: the results produced since then were poisoned with NaN.  Asynchronous calls
 * (pfsp_mat_action returns once its kernel is queued) cannot report this themselves. */
FSP_API int         pfsp_check(void);
FSP_API const char *pfsp_last_error(void);

/* ---- state set ---- */
/* the library's world communicator as an fspcomm_t of fsp_b200.h (NULL on a single rank) */
FSP_API void *pfsp_world_comm(void);
FSP_API int pfsp_set_create(void **set);
FSP_API int pfsp_set_destroy(void *set);
/* Extension (multi-GPU): distributed construction as in src/StateSet/StateSetBase.cpp:134-154,188-258 -- every rank
 * keeps and expands only its block of states, the directory is striped over the GPUs (fsp_b200.h: fspset_set_sharded).
 * Call before the first pfsp_set_add_states; a no-op on one rank or without peer memory.  On > 1 rank this is the
 * default; on = 0 (or FSP_SHARDED_SET=0 in the environment) keeps the replicated directory. */
FSP_API int pfsp_set_set_sharded(void *set, int on);
FSP_API int pfsp_set_is_sharded(void *set);
/* State2Index(states_old) of src/Fsp/FspSolverMultiSinks.cpp:174-176 on the device: remember the local block before
 * pfsp_set_expand, ask for the global indices those n states have afterwards */
FSP_API int pfsp_set_remember_local(void *set);
FSP_API int pfsp_set_remembered_indices(void *set, int n, int *idx);
FSP_API int pfsp_set_stoichiometry(void *set, int S, int R, const int *SM_colmajor);
FSP_API int pfsp_set_shape(void *set, int K, const int *bounds, pfsp_constr_fn lhs_or_null, void *args);
FSP_API int pfsp_set_shape_bounds(void *set, int K, const int *bounds);
FSP_API int pfsp_set_add_states(void *set, int S, int m, const int *X_colmajor);
FSP_API int pfsp_set_add_box_lattice(void *set, int S, const int *upper);
FSP_API int pfsp_set_expand(void *set);
FSP_API int pfsp_set_sizes(void *set, int *n_local, int *n_global, int *local_start);
FSP_API int pfsp_set_copy_states(void *set, int *out_colmajor);
FSP_API int pfsp_set_state2index(void *set, int m, const int *X_colmajor, int *idx);

/* ---- model ---- */
FSP_API int pfsp_model_create(void **model, int S, int R, const int *SM_colmajor, pfsp_prop_fn prop_x, void *prop_x_args,
                              pfsp_tcoef_fn prop_t, void *prop_t_args, int n_tv, const int *tv_reactions);
/* named workload of pacmensl_b200/fixtures/fsp_models.h; bounds/expansion/x0/t_final/tolerances are returned */
FSP_API int pfsp_model_from_fixture(void **model, const char *name, int *S, int *R, int *K, int *bounds, double *expansion,
                                    int *x0, double *t_final, double *fsp_tol, double *rtol, double *atol,
                                    pfsp_constr_fn *lhs);
FSP_API int pfsp_model_set_mass_action(void *model, const double *rates, const int *orders_colmajor);
/* adds the per-species factor values[min(x_species, len - 1)] to reaction `reaction` of the device-evaluable form
 * (call after pfsp_model_set_mass_action; orders 0 and rate 1 give a pure table factor) */
FSP_API int pfsp_model_set_factor_table(void *model, int species, int reaction, int len, const double *values);
/* attaches the separable (device-evaluable) description of a named fixture's propensities, if one exists (0) or not (1):
 * matrix generation then evaluates the propensities on the GPU instead of calling prop_x on the host */
FSP_API int pfsp_model_attach_device_form(void *model, const char *fixture_name);
/* copies the S x R stoichiometry matrix (column major: SM[r*S + s]) */
FSP_API int pfsp_model_get_stoichiometry(void *model, int *SM_colmajor);
FSP_API int pfsp_model_destroy(void *model);

/* ---- operator ---- */
FSP_API int pfsp_mat_create(void **mat, int constrained);
FSP_API int pfsp_mat_destroy(void *mat);
FSP_API int pfsp_mat_generate(void *mat, void *set, void *model);
FSP_API int pfsp_mat_clear(void *mat);
FSP_API int pfsp_mat_set_variant(void *mat, int variant);
FSP_API int pfsp_mat_info(void *mat, int *n_rows_local, long *flops, double *action_bytes);
/* y = A(t) x with DEVICE buffers of n_rows_local doubles */
FSP_API int pfsp_mat_action(void *mat, double t, const double *x_dev, double *y_dev);
/* diagnostics, multi-GPU: only the halo exchange of an Action (push CTAs + finishing CTA, no rows); *bytes_sent = bytes
 * this rank stores into its peers' windows per exchange */
FSP_API int pfsp_mat_halo_only(void *mat, const double *x_dev, double *y_dev, long *bytes_sent);
/* the same call with HOST buffers (pinned or pageable): H2D copy of x, Action, D2H copy of y */
FSP_API int pfsp_mat_action_host(void *mat, double t, const double *x_host, double *y_host);

/* ---- adaptive FSP driver ---- */
FSP_API int pfsp_solver_create(void **solver, int ode_type /* 0 KRYLOV, 1 CVODE, 2 PETSC->BDF */);
FSP_API int pfsp_solver_destroy(void *solver);
FSP_API int pfsp_solver_set_model(void *solver, void *model);
FSP_API int pfsp_solver_set_initial_bounds(void *solver, int K, const int *bounds);
FSP_API int pfsp_solver_set_constraint_function(void *solver, pfsp_constr_fn lhs, void *args);
FSP_API int pfsp_solver_set_expansion_factors(void *solver, int K, const double *factors);
FSP_API int pfsp_solver_set_initial_distribution(void *solver, int S, int m, const int *X_colmajor, const double *p);
FSP_API int pfsp_solver_set_ode_tolerances(void *solver, double rtol, double atol);
FSP_API int pfsp_solver_set_verbosity(void *solver, int level);
FSP_API int pfsp_solver_set_krylov(void *solver, int q_iop, int m_min, int m_max);
/* 0 (default): the BDF integrator is re-created after every expansion like the reference's CVODE
 * (src/Fsp/FspSolverMultiSinks.cpp:92-108); 1: it keeps its history (step size, order, Nordsieck array) across FSP
 * expansions -- correct but measured slower */
FSP_API int pfsp_solver_set_warm_restart(void *solver, int on);
FSP_API int pfsp_solver_num_warm_restarts(void *solver, int *n);
/* Extension: build the solver's state set sharded over the ranks (pfsp_set_set_sharded); before pfsp_solver_setup */
FSP_API int pfsp_solver_set_sharded_state_set(void *solver, int on);
FSP_API int pfsp_solver_setup(void *solver);
/* Solve to t_final; *n_local receives the number of local states of the result (kept inside the handle) */
FSP_API int pfsp_solver_solve(void *solver, double t_final, double fsp_tol, double t_init, int *n_local, int *n_species);
FSP_API int pfsp_solver_copy_result(void *solver, int *states_colmajor, double *p);
FSP_API int pfsp_solver_stats(void *solver, int *n_global_states, int *n_expansions, long *n_rhs_evals, int *K,
                              int *final_bounds);
FSP_API int pfsp_solver_clear(void *solver);

#ifdef __cplusplus
}
#endif
#endif

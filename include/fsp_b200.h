/*
 * fsp_b200.h -- C ABI of the B200-native FSP time-stepping core (libpacmensl_b200.so).
 *
 * This is the drop-in boundary between host C++ (the pacmensl:: classes that mirror the reference
 * API, pacmensl_b200/host/) and the hand-written sm_100a CUDA kernels (pacmensl_b200/csrc/).
 * Plain pointers and sizes only; no C++/torch types.  All functions return 0 on success and a
 * non-zero code on failure (the reference convention, src/Sys/ErrorHandling.h:29-54), never throw,
 * and are called from one host thread per device.
 *
 * Each group names the reference interface it replaces (paths relative to the reference tree).
 * Pointers named *_dev are device pointers on the current device; `stream` is a cudaStream_t passed
 * as void* (NULL = the library's default stream for this thread's device).
 */
#ifndef FSP_B200_H_
#define FSP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSP_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * Runtime: device selection, memory, streams.
 * Replaces PetscInitialize/VecCreate storage management (src/Sys/Sys.cpp:31-63,
 * src/PetscWrap/PetscWrap.h:12-60) for device-resident vectors.
 * ---------------------------------------------------------------------------------------------- */
FSP_API int         fsp_device_count(int *count);
FSP_API int         fsp_device_set(int device);
FSP_API int         fsp_device_get(int *device);
FSP_API int         fsp_device_sm_count(int *count);
FSP_API const char *fsp_last_error(void);
/* Device memory comes from the stream-ordered pool of the LEGACY stream 0 (cudaMallocAsync / cudaFreeAsync on stream 0).
 * The library's own streams are non-blocking and do not order against stream 0: every path that hands pool memory to
 * another stream synchronises first (e.g. the host-vector pipeline calls fsp_stream_sync(NULL) before its copy streams
 * touch the buffers), and callers that pass their own non-blocking stream to fspmat_ / fspvec_ functions must do the
 * same after an fsp_malloc and before an fsp_free of memory that stream uses. */
FSP_API int         fsp_malloc(void **ptr_dev, size_t bytes);
FSP_API int         fsp_free(void *ptr_dev);
FSP_API int         fsp_malloc_host(void **ptr_host, size_t bytes); /* pinned */
FSP_API int         fsp_free_host(void *ptr_host);
FSP_API int         fsp_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, void *stream);
FSP_API int         fsp_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, void *stream);
/* asynchronous variants: the host buffer must be pinned (fsp_malloc_host) and stay valid until the stream reaches the copy */
FSP_API int         fsp_memcpy_h2d_async(void *dst_dev, const void *src_host, size_t bytes, void *stream);
FSP_API int         fsp_memcpy_d2h_async(void *dst_host, const void *src_dev, size_t bytes, void *stream);
FSP_API int         fsp_memcpy_d2d(void *dst_dev, const void *src_dev, size_t bytes, void *stream);
FSP_API int         fsp_memset(void *dst_dev, int byte, size_t bytes, void *stream);
FSP_API int         fsp_stream_create(void **stream);
FSP_API int         fsp_stream_destroy(void *stream);
FSP_API int         fsp_stream_sync(void *stream);
FSP_API int         fsp_device_sync(void);
FSP_API int         fsp_event_create(void **event);
FSP_API int         fsp_event_destroy(void *event);
FSP_API int         fsp_event_record(void *event, void *stream);
FSP_API int         fsp_event_sync(void *event);
FSP_API int         fsp_stream_wait_event(void *stream, void *event);
FSP_API int         fsp_event_elapsed_ms(void *start, void *stop, float *ms); /* syncs on stop */
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
FSP_API long long   fsp_launch_count(void);
/* CUDA graphs for launch-bound inner loops (the Arnoldi/IOP basis generation of KrylovFsp on small state sets):
 * capture everything the library submits to `stream` (a stream from fsp_stream_create, not NULL) between begin and
 * end, instantiate once, replay with one launch.  Kernels launched while capturing are not executed. */
typedef struct fsp_graph_s *fsp_graph_t;
FSP_API int         fsp_graph_begin_capture(void *stream);
FSP_API int         fsp_graph_end_capture(void *stream, fsp_graph_t *out);
FSP_API int         fsp_graph_abort_capture(void *stream);
FSP_API int         fsp_graph_launch(fsp_graph_t g, void *stream);
FSP_API int         fsp_graph_num_kernels(fsp_graph_t g, long *n);
FSP_API int         fsp_graph_destroy(fsp_graph_t g);

/* ------------------------------------------------------------------------------------------------
 * Device vectors (fp64).  Replaces the PETSc Vec BLAS-1 calls on the hot path:
 *   VecSet/VecCopy/VecScale/VecAXPY/VecDot/VecNorm/VecMAXPY/VecSum  (src/OdeSolver/KrylovFsp.cpp:138,
 *   153,244-252,280-309; src/Matrix/FspMatrixBase.cpp:39,50,58) and the SUNDIALS N_Vector ops CVODE
 *   uses (N_VLinearSum, N_VWrmsNorm, N_VDotProd, N_VScale, N_VConst; src/OdeSolver/CvodeFsp.cpp:150).
 * Reductions write fp64 results to `out_dev` (device memory, asynchronous) -- read them with
 * fsp_memcpy_d2h, or use the *_h variants which synchronise the stream and return on the host.
 * Reductions are deterministic for a given (n, device).
 * ---------------------------------------------------------------------------------------------- */
FSP_API int fspvec_set(double *y_dev, double alpha, long n, void *stream);
FSP_API int fspvec_copy(double *y_dev, const double *x_dev, long n, void *stream);
FSP_API int fspvec_scale(double *y_dev, double alpha, long n, void *stream);
/* y += alpha x */
FSP_API int fspvec_axpy(double *y_dev, double alpha, const double *x_dev, long n, void *stream);
/* z = a x + b y  (z may alias x or y) */
FSP_API int fspvec_linear_sum(double *z_dev, double a, const double *x_dev, double b, const double *y_dev, long n,
                              void *stream);
/* z = w * (a x + b y) elementwise (w == NULL: no weights).  Fuses the scaled operator application of the
 * GMRES loop:  V_{l+1} = s1 .* (v - gamma J v)  and the unscaling x ./ s2 (via fspvec_div). */
FSP_API int fspvec_wlincomb(double *z_dev, const double *w_dev, double a, const double *x_dev, double b,
                            const double *y_dev, long n, void *stream);
/* z = x ./ w */
FSP_API int fspvec_div(double *z_dev, const double *x_dev, const double *w_dev, long n, void *stream);
/* z = x .* w */
FSP_API int fspvec_prod(double *z_dev, const double *x_dev, const double *w_dev, long n, void *stream);
/* z = c0 x0 + c1 x1 + c2 x2 (three-term linear combination; x2 may be NULL) */
FSP_API int fspvec_lincomb3(double *z_dev, double c0, const double *x0_dev, double c1, const double *x1_dev, double c2,
                            const double *x2_dev, long n, void *stream);
/* y = beta*y + sum_k alpha[k] X[k]  (VecMAXPY; X_dev_ptrs is a HOST array of m device pointers, m <= 64) */
FSP_API int fspvec_maxpy(double *y_dev, double beta, int m, const double *alpha_host,
                         const double *const *X_dev_ptrs, long n, void *stream);
/* out[k] = <x, Y[k]>, k < m <= 8, one pass over x */
FSP_API int fspvec_mdot(double *out_dev, const double *x_dev, int m, const double *const *Y_dev_ptrs, long n,
                        void *stream);
FSP_API int fspvec_dot(double *out_dev, const double *x_dev, const double *y_dev, long n, void *stream);
FSP_API int fspvec_norm2sq(double *out_dev, const double *x_dev, long n, void *stream); /* sum x_i^2 */
FSP_API int fspvec_sum(double *out_dev, const double *x_dev, long n, void *stream);
FSP_API int fspvec_norm1(double *out_dev, const double *x_dev, long n, void *stream);
/* sum (x_i * w_i)^2 : building block of N_VWrmsNorm */
FSP_API int fspvec_wsqsum(double *out_dev, const double *x_dev, const double *w_dev, long n, void *stream);
/* w_i = 1 / (rtol*|y_i| + atol)  (CVODE error weights, cvEwtSetSS); out = min_i(rtol|y_i|+atol) */
FSP_API int fspvec_ewt(double *w_dev, const double *y_dev, double rtol, double atol, long n, double *min_out_dev,
                       void *stream);
/* out = max_i |x_i| / (a |y_i| + b)   (CVODE's upper bound on the first step, cvUpperBoundH0) */
FSP_API int fspvec_ratio_absmax(double *out_dev, const double *x_dev, const double *y_dev, double a, double b, long n,
                                void *stream);
/* Fused modified-Gram-Schmidt step used by the Arnoldi/IOP loop (KrylovFsp.cpp:302-309):
 *   w -= (*h_dev) * v ;  out = <w, u>    (u may be NULL: then out = <w, w>)            */
FSP_API int fspvec_axpy_dot(double *w_dev, const double *h_dev, double sign, const double *v_dev,
                            const double *u_dev, double *out_dev, long n, void *stream);
/* w *= 1/sqrt(*normsq_dev)  (VecScale with a device-resident scalar; KrylovFsp.cpp:308) */
FSP_API int fspvec_scale_rsqrt(double *w_dev, const double *normsq_dev, long n, void *stream);
/* ---- fused passes of the BDF/Newton/GMRES integrator (replace chains of N_VLinearSum/N_VProd/N_VDiv/N_VWrmsNorm/
 * N_VScale calls CVODE makes per step; src/OdeSolver/CvodeFsp.cpp:41-58 hands these to SUNDIALS) ------------------- */
/* b = c0 x0 + c1 x1 + c2 x2 ; v = b .* w ; out = sum v_i^2   (Newton residual, scaled GMRES start vector, its norm and
 * the WRMS norm of b in one pass) */
FSP_API int fspvec_lincomb3_wprod_sqsum(double *b_dev, double *v_dev, double c0, const double *x0_dev, double c1,
                                        const double *x1_dev, double c2, const double *x2_dev, const double *w_dev,
                                        double *out_dev, long n, void *stream);
/* v *= a ; t = v ./ w */
FSP_API int fspvec_scale_div(double *v_dev, double a, double *t_dev, const double *w_dev, long n, void *stream);
/* v *= a ; t = v .* winv   (winv = 1 ./ w from fspvec_ewt_pair: fp64 division halves the bandwidth of such a pass) */
FSP_API int fspvec_scale_mul(double *v_dev, double a, double *t_dev, const double *winv_dev, long n, void *stream);
/* d = x .* winv ; acor += d ; ycur = zn0 + acor ; out = sum x_i^2 (== sum (d_i ewt_i)^2 since winv .* ewt == 1) */
FSP_API int fspvec_newton_update_mul(const double *x_dev, const double *winv_dev, double *acor_dev, const double *zn0_dev,
                                     double *ycur_dev, double *out_dev, long n, void *stream);
/* fspvec_ewt that also stores the denominators winv_i = rtol |y_i| + atol = 1 / w_i */
FSP_API int fspvec_ewt_pair(double *w_dev, double *winv_dev, const double *y_dev, double rtol, double atol, long n,
                            double *min_out_dev, void *stream);
/* d = dw ? x ./ dw : x ; acor += d ; ycur = zn0 + acor ; out = sum (d_i ewt_i)^2   (end of a Newton iteration) */
FSP_API int fspvec_newton_update(const double *x_dev, const double *dw_dev, const double *ewt_dev, double *acor_dev,
                                 const double *zn0_dev, double *ycur_dev, double *out_dev, long n, void *stream);
/* Nordsieck history array Z[0..L) (L <= 8, Z_dev_ptrs is a HOST array of device pointers): optional rescale
 * Z[j] *= scale_host[j] for j >= 1 (scale_host == NULL: none), then the Pascal-triangle prediction (pascal = +1:
 * for k = 1..L-1, j = L-1..k: Z[j-1] += Z[j]), its inverse (pascal = -1) or nothing (0) -- one pass, bit-identical to
 * the sequence of separate scale/axpy calls (cvRescale, cvPredict, cvRestore of CVODE). */
FSP_API int fspvec_nordsieck(double *const *Z_dev_ptrs, int L, const double *scale_host, int pascal, long n, void *stream);
/* Z[j] += coef_host[j] * x for j < L  (history update after an accepted step, cvCompleteStep) */
FSP_API int fspvec_multi_axpy(double *const *Z_dev_ptrs, int L, const double *coef_host, const double *x_dev, long n,
                              void *stream);
/* host-result conveniences (synchronise `stream`) */
/* KrylovFsp's incomplete orthogonalisation of one basis vector (src/OdeSolver/KrylovFsp.cpp:302-309, q_iop <= 2) in ONE
 * cooperative launch: h_0 = <w,u0>; [w -= h_0 u0; h_1 = <w,u1>;] w -= h_last u_last; s2 = <w,w>; w /= sqrt(s2), with w held in
 * registers across the phases when it fits (read and written once) and grid-wide barriers between them.  nvec = 1 uses
 * u1 only.  h_dev: nvec coefficients, then s2.  Returns 1 without launching if the device cannot do it. */
FSP_API int fspvec_iop_orth(double *w_dev, int nvec, const double *u0_dev, const double *u1_dev, double *h_dev, long n,
                            void *stream);
/* Post-processing on the device (src/Fsp/DiscreteDistribution.cpp:171-200, src/SensFsp/SensDiscreteDistribution.cpp:216-271):
 *   marginal      out_dev[b] = sum of p over the states whose coordinate `species` equals b, b < M (deterministic order)
 *   max_species   *out_neg_dev = -(largest coordinate `species` over the n states)
 *   clamp_min     x_i = max(x_i, lo); *count_out_dev = number of entries raised (ComputeFIM's 1e-16 floor on p)
 *   wdiv_dot      sum_i x_i y_i / w_i  (one Fisher-information entry: x = dp/dtheta_i, y = dp/dtheta_j, w = p) */
FSP_API int fspvec_marginal(double *out_dev, int M, const double *p_dev, const int *states_dev, int S, int species, long n,
                            void *stream);
FSP_API int fspvec_max_species(double *out_neg_dev, const int *states_dev, int S, int species, long n, void *stream);
FSP_API int fspvec_clamp_min(double *count_out_dev, double *x_dev, double lo, long n, void *stream);
FSP_API int fspvec_wdiv_dot(double *out_dev, const double *x_dev, const double *y_dev, const double *w_dev, long n, void *stream);
FSP_API int fspvec_dot_h(double *out_host, const double *x_dev, const double *y_dev, long n, void *stream);
FSP_API int fspvec_norm2_h(double *out_host, const double *x_dev, long n, void *stream);
FSP_API int fspvec_sum_h(double *out_host, const double *x_dev, long n, void *stream);
FSP_API int fspvec_norm1_h(double *out_host, const double *x_dev, long n, void *stream);
/* ExpandVec (src/PetscWrap/PetscWrap.cpp:26-56): p_new = 0; p_new[new_idx[i]] = p_old[i] */
FSP_API int fspvec_scatter(double *p_new_dev, long n_new, const double *p_old_dev, const int *new_idx_dev,
                           long n_old, void *stream);
/* multi-GPU ExpandVec helper: p_new[gidx[i] - own_start] = vals[i] for the entries that land in
 * [own_start, own_start + n_new); p_new must have been zeroed by the caller (VecSetUp does). */
FSP_API int fspvec_scatter_range(double *p_new_dev, long n_new, const double *vals_dev, const int *gidx_dev, long n,
                                 long own_start, void *stream);
/* out[i] = x[idx[i]]  (MakeDiscreteDistribution_ scatter, src/Fsp/FspSolverMultiSinks.cpp:703-735) */
FSP_API int fspvec_gather(double *out_dev, const double *x_dev, const int *idx_dev, long n, void *stream);
/* Multi-GPU ExpandVec routing (src/Sys/PetscWrap.cpp:10-45): the n entries (idx[i], val[i]) sorted by the rank owning
 * global index idx[i] under starts_host[0..n_ranks] (entries outside are dropped); counts_host[r] = entries for rank r,
 * rank r's segment starts at sum(counts[0..r)) of idx_sorted / val_sorted.  Synchronises the stream. */
FSP_API int fspvec_route_by_owner(const int *idx_dev, const double *val_dev, long n, const long *starts_host, int n_ranks,
                                  int *idx_sorted_dev, double *val_sorted_dev, long *counts_host, void *stream);

/* ------------------------------------------------------------------------------------------------
 * State set on the device: state list + hash directory.
 * Replaces Zoltan_DD_{Create,Find,Update} and the Armadillo column bookkeeping in
 *   StateSetBase::AddStates      src/StateSet/StateSetBase.cpp:188-258
 *   StateSetBase::State2Index    src/StateSet/StateSetBase.cpp:309-423
 *   StateSetConstrained::Expand / CheckValidityStates / CheckConstraints
 *                                src/StateSet/StateSetConstrained.cpp:33-82,132-221
 * States are int32, S per state, state i at states[i*S .. i*S+S-1] (the reference's column-major
 * arma::Mat<int>, src/StateSet/StateSetBase.h:83).  Index = insertion order (np = 1 semantics);
 * new states of one Expand wave are appended in first-discovery order of the reaction-major child
 * list (StateSetConstrained.cpp:175-179 + Sys/pacmenMath.h:204-213).
 * ---------------------------------------------------------------------------------------------- */
typedef struct fspset_s *fspset_t;
/* host callback with the reference's fsp_constr_multi_fn contract (StateSetConstrained.h:32-33):
 * out[K*j + k] = lhs_k(state j).  Called on host copies of candidate states. */
typedef int (*fspset_constr_fn)(int num_species, int num_constr, int num_states, int *states, int *out, void *args);

FSP_API int fspset_create(fspset_t *out, int num_species, int num_reactions, const int *SM_host /* S x R col-major */);
FSP_API int fspset_destroy(fspset_t h);
/* lhs == NULL => default identity constraints evaluated on the device (requires K == S) */
FSP_API int fspset_set_shape(fspset_t h, int num_constr, fspset_constr_fn lhs, const int *bounds_host, void *args);
FSP_API int fspset_set_bounds(fspset_t h, int num_constr, const int *bounds_host);
/* AddStates: X is S x m (host or device, `on_device`); present states and in-batch duplicates are shed,
 * the rest appended in order with status 1.  Returns -1 if num_species mismatches (KAT-S2). */
FSP_API int fspset_add_states(fspset_t h, int num_species, long m, const int *X, int on_device);
FSP_API int fspset_expand(fspset_t h);
FSP_API int fspset_num_states(fspset_t h, int *n);
/* State2Index: idx[j] = index of X[:, j] or -1 (negative coordinate or absent). */
FSP_API int fspset_state2index(fspset_t h, long m, const int *X, int x_on_device, int *idx, int idx_on_device);
/* idx[i] = State2Index(state_i + sign * nu)  for all stored states i in [first, first+count) */
FSP_API int fspset_lookup_shifted(fspset_t h, const int *nu_host, int sign, long first, long count, int *idx_dev);
/* CheckConstraints on state_i + nu for stored states: satisfied_dev[k*count + i] in {0,1}
 * (constraint-major like StateSetConstrained.cpp:71; negative coordinates => satisfied). */
FSP_API int fspset_check_constraints_shifted(fspset_t h, const int *nu_host, long first, long count,
                                             int *satisfied_dev);
/* Sink column lists (FspMatrixConstrained.cpp:170-194): for each constraint k, the ascending indices
 * (relative to `first`) of stored states whose destination state_i + nu violates constraint k, written k
 * after k into idx_out_dev (capacity `cap`); counts_host[k] = list lengths. */
FSP_API int fspset_sink_lists(fspset_t h, const int *nu_host, long first, long count, int *idx_out_dev, long cap,
                              long *counts_host);
/* upper bound on the number of states in [first, first+count) that can have sink entries (status != 0 after Expand) */
FSP_API int fspset_num_boundary_states(fspset_t h, long first, long count, long *n);
FSP_API int fspset_states_dev(fspset_t h, const int **states_dev); /* borrowed, valid until next mutation */
FSP_API int fspset_copy_states(fspset_t h, long first, long count, int *out_host);
FSP_API int fspset_copy_status(fspset_t h, long first, long count, signed char *out_host);
/* Fill the set with the full lexicographic box lattice 0..bounds[s] (species 0 fastest;
 * Sys/pacmenMath.h:33-59 convention) -- the synthetic workload of SURVEY.md section 8(d). */
FSP_API int fspset_add_box_lattice(fspset_t h, const int *upper_host);

/* Sharded construction over N GPUs of one node (the reference distributes the frontier and the directory over the MPI
 * ranks: src/StateSet/StateSetBase.cpp:134-154,188-258 with Zoltan_DD, and migrates states after load balancing:
 * src/Partitioner/StatePartitionerBase.cpp:136-239).  After fspset_set_sharded (before any state is added; `comm` must
 * have peer memory enabled) the set holds ONLY the local block of states; the hash directory is striped over the ranks'
 * HBM (shard = hash % N) and probed / claimed through NVLink peer loads and system-scope atomics; add_states,
 * add_box_lattice and expand become COLLECTIVE (every rank explores the frontier states it owns), and end with a
 * re-balance to the contiguous equal-count BLOCK layout (rank r owns global indices [starts[r], starts[r+1])) followed
 * by a rebuild of the directory.  Global index = starts[owner] + local position; unlike the replicated set it changes
 * when the set grows (as in the reference), so callers that keep vectors re-map them through
 * fspset_remember_local / fspset_remembered_indices.  `first` arguments of the per-row functions stay GLOBAL indices
 * and must address local rows.  fspset_num_states returns the global count. */
struct fspcomm_s;
FSP_API int fspset_set_sharded(fspset_t h, struct fspcomm_s *comm);
FSP_API int fspset_is_sharded(fspset_t h);
FSP_API int fspset_layout(fspset_t h, long *starts_host /* n_ranks + 1 */, long *n_local);
/* host arithmetic of the re-balance, no device access: BLOCK layout of sum(counts) states and the segments rank `rank`
 * pulls (seg_len[k] states from position seg_off[k] of rank seg_src[k] to its own position seg_dst[k]); arrays of
 * n_ranks (+ 1 for starts) entries */
FSP_API int fspset_rebalance_plan(int n_ranks, const long *counts, int rank, long *starts, long *seg_src, long *seg_off,
                                  long *seg_dst, long *seg_len, int *n_seg, int *already_balanced);
/* keep a device copy of the local block of states / afterwards: their current global indices (host array of the
 * remembered length), the copy is released.  StateSetBase::State2Index(states_old) of FspSolverMultiSinks.cpp:174-205. */
FSP_API int fspset_remember_local(fspset_t h);
FSP_API int fspset_remembered_indices(fspset_t h, int *idx_host, long n_expected);

/* ------------------------------------------------------------------------------------------------
 * Propensity evaluation on the device for models given in mass-action form
 *   d_r(x) = rate[r] * prod_s binom-like falling factorial of x_s of order ord[r*S+s] (0,1,2)
 * (an extension: the reference only has host std::function callbacks, src/Models/Model.h:44-60;
 * host callbacks remain supported through fspmat_generate with host arrays).
 * out_dev[i] = d_r(state_i + sign*nu) for stored states in [first, first+count).
 * ---------------------------------------------------------------------------------------------- */
FSP_API int fspset_eval_mass_action(fspset_t h, double rate, const int *order_host /* S */, const int *nu_host,
                                    int sign, long first, long count, double *out_dev);
/* The separable form: d_r(x) = rate * prod_s ff(x_s, ord_s) * T_s[min(x_s, len_s - 1)], with one optional factor table per
 * species (len_s == 0: none; a negative coordinate gives 0).  tables_dev holds the tables back to back, tab_off_host[s]
 * is where species s starts.  Covers gene-state switches and saturating (Hill-type) factors of a single species: the
 * hog1p model of examples/hog1p.cpp:14-112 is evaluated entirely on the device this way. */
FSP_API int fspset_eval_separable(fspset_t h, double rate, const int *order_host /* S */, const double *tables_dev,
                                  const int *tab_off_host /* S */, const int *tab_len_host /* S */, const int *nu_host,
                                  int sign, long first, long count, double *out_dev);

/* ------------------------------------------------------------------------------------------------
 * The FSP operator A(t) = sum_r c_r(t) A_r (+ sink rows).
 * Replaces FspMatrixBase / FspMatrixConstrained storage and Action:
 *   GenerateValues  src/Matrix/FspMatrixBase.cpp:76-251, src/Matrix/FspMatrixConstrained.cpp:121-282
 *   Action          src/Matrix/FspMatrixBase.cpp:36-62,  src/Matrix/FspMatrixConstrained.cpp:31-64
 *   GetLocalMVFlops src/Matrix/FspMatrixBase.cpp:429-444, src/Matrix/FspMatrixConstrained.cpp:447-465
 *   Destroy         src/Matrix/FspMatrixBase.cpp:258-275
 * Layout handed to fspmat_generate ("reaction-plane ELL"): P = n_tv + n_ti planes of length n (leading
 * dimension ld), TV reactions first.  Plane p, row i:
 *   col[p*ld+i]  local index of x_i - nu_r in x (>= 0), -1 = absent, <= -2 = ghost slot -(col+2)
 *   off[p*ld+i]  d_r(x_i - nu_r)        diag[p*ld+i]  d_r(x_i)  (positive)
 * Sink entries: segment (p, k) = sink_ptr[p*K+k] .. sink_ptr[p*K+k+1] of (sink_idx, sink_val):
 *   row N+k gets  c_p * sink_val * x[sink_idx].
 * The library re-packs into its HBM layout (TI diagonals merged, planes padded/aligned).
 * ---------------------------------------------------------------------------------------------- */
typedef struct fspmat_s *fspmat_t;

typedef struct fspmat_desc {
  int           n_states;     /* local states n */
  int           n_rows;       /* local rows: n, or n + K on the rank that owns the sinks */
  int           n_reactions;  /* R: length of the coefficient vector t_fun fills */
  int           n_tv, n_ti;
  const int    *tv_reactions; /* host, [n_tv] reaction ids */
  const int    *ti_reactions; /* host, [n_ti] */
  const int    *col;
  const double *off;
  const double *diag;
  long          ld;
  int           arrays_on_device; /* 0: col/off/diag/sink_idx/sink_val are host pointers; 1: device */
  int           n_constr;         /* K (0 for FspMatrixBase) */
  const long   *sink_ptr;         /* host, [(n_tv+n_ti)*K + 1] */
  const int    *sink_idx;
  const double *sink_val;
  int           owns_sinks;       /* 1 if rows n..n+K-1 of y live on this rank */
  long          n_ghost;          /* number of ghost slots referenced by col <= -2 */
} fspmat_desc;

FSP_API int fspmat_create(fspmat_t *out);
FSP_API int fspmat_destroy(fspmat_t h);
FSP_API int fspmat_generate(fspmat_t h, const fspmat_desc *desc);
FSP_API int fspmat_clear(fspmat_t h); /* Destroy(): frees values, object reusable */
/* y = A(t) x.  coef_host[r] = c_r(t) for r < R (only TV entries are read; TI reactions use 1).
 * x_dev, y_dev have n_rows entries; ghost_dev has n_ghost entries (may be NULL when n_ghost == 0).
 * On a rank that does not own the sinks, the K partial sink sums are written to sink_out_dev
 * (K doubles) instead of y; pass NULL to drop them. */
FSP_API int fspmat_action(fspmat_t h, const double *coef_host, const double *x_dev, const double *ghost_dev,
                          double *y_dev, double *sink_out_dev, void *stream);
/* Action with a fused epilogue, for the solver loops that would otherwise re-read y in separate passes (single GPU):
 *   y = scale .* (beta x + alpha A(t) x)      (scale_dev == NULL: no scaling; applies to the K sink rows too)
 *   dot_out_dev[k] = <y, dot_vec_dev[k]> for k < n_dots <= 2   (dot_vec_dev[k] == NULL: <y, y>)
 * KrylovFsp: alpha = 1, beta = 0, one dot with the first basis vector of the orthogonalisation window
 * (KrylovFsp.cpp:296-305).  CVODE's scaled GMRES: y = s1 .* (v - gamma J v), <y, V_0>, <y, y>.
 * Inner products are accumulated per CTA and summed in a fixed order: deterministic. */
typedef struct fspmat_epilogue {
  double        alpha, beta;
  const double *scale_dev;
  int           n_dots;
  const double *dot_vec_dev[2];
  double       *dot_out_dev;
} fspmat_epilogue;
FSP_API int fspmat_fused_supported(fspmat_t h); /* 1: fspmat_action_fused can be used (values present, no ghost columns) */
FSP_API int fspmat_action_fused(fspmat_t h, const double *coef_host, const double *x_dev, double *y_dev,
                                const fspmat_epilogue *ep, void *stream);
/* Split action for overlapping the halo exchange with compute (multi-GPU):
 *   phase 1  interior pass: all rows with ghost entries counted as 0, no sink rows (ghost_dev unused)
 *   phase 2  boundary rows only (the rows that reference ghost slots; needs ghost_dev) -- run after phase 1
 *   phase 3  sink partial sums only, written to sink_out_dev (K doubles)
 *   phase 0  == fspmat_action (everything in one launch) */
FSP_API int fspmat_action_phase(fspmat_t h, const double *coef_host, const double *x_dev, const double *ghost_dev,
                                double *y_dev, double *sink_out_dev, int phase, void *stream);
/* Host-vector pipeline (single GPU): rows [row_begin, row_end) of y = A(t) x; with_sinks != 0 adds the K sink rows
 * (which need all of x).  fspmat_chunk_max_columns tells which prefix of x every chunk of rows references, so a chunk
 * can run as soon as that prefix has been uploaded while its part of y is downloaded behind it. */
FSP_API int fspmat_action_rows(fspmat_t h, const double *coef_host, const double *x_dev, double *y_dev, long row_begin,
                               long row_end, int with_sinks, void *stream);
FSP_API int fspmat_chunk_max_columns(fspmat_t h, long chunk_rows, int n_chunks, int *chunk_max_host);
FSP_API int fspmat_num_boundary_rows(fspmat_t h, long *n);
/* Peer-memory variants (see fsphalo_* below).  sinks: the K partial sink sums go straight into the sink owner's slot
 * row (e->sink_slot_remote) followed by the flag; boundary: waits in device code for every peer's halo flag, redoes
 * the rows that reference ghost entries and, on the sink owner, adds the slots in rank order into y[n..n+K). */
struct fsphalo_epoch;
FSP_API int fspmat_action_sinks_p2p(fspmat_t h, const double *coef_host, const double *x_dev,
                                    const struct fsphalo_epoch *e, void *stream);
FSP_API int fspmat_action_boundary_p2p(fspmat_t h, const double *coef_host, const double *x_dev, double *y_dev,
                                       const struct fsphalo_epoch *e, void *stream);
/* The whole multi-GPU action as ONE kernel: CTAs whose rows reference no ghost entry run first, the CTAs with ghost rows
 * wait in device code for the peers' flags, a trailing CTA on the sink owner adds the sink slots.  Every row is
 * computed once (no interior pass + redo), and there is no kernel boundary between local and halo-dependent work. */
FSP_API int fspmat_action_p2p(fspmat_t h, const double *coef_host, const double *x_dev, double *y_dev,
                              const struct fsphalo_epoch *e, void *stream);
/* CTAs of fspmat_action_p2p that never wait (no ghost rows) / all CTAs.  The host uses it to decide whether the push
 * kernel must be complete before the action kernel starts (few ghost-free CTAs: waiting CTAs could otherwise occupy
 * every SM slot before this GPU's own push kernel has been scheduled). */
FSP_API int fspmat_p2p_cta_counts(fspmat_t h, long *n_interior, long *n_total);
FSP_API int fspmat_flops(fspmat_t h, long *nflops);
FSP_API int fspmat_num_rows(fspmat_t h, int *n_rows);
/* algorithmic bytes of one Action: n*(16 + 12 P + 8 (n_tv + [n_ti>0])) + 12 nnz_sink + 8 K */
FSP_API int fspmat_action_bytes(fspmat_t h, double *bytes);
/* Coefficient applied to the time-invariant reactions by every later action of this operator (default 1; the reference
 * gives them coefficient 1, FspMatrixBase.cpp:58).  0 turns the operator into sum_{r in TV} coef[r] A_r -- what the j-th
 * time derivative of A(t) is when coef holds the j-th derivatives of the time coefficients. */
FSP_API int fspmat_set_ti_coef(fspmat_t h, double c);
/* kernel variant selection for tuning/benchmarks: 0 = default */
FSP_API int fspmat_set_variant(fspmat_t h, int variant);
/* Multi-GPU set-up (the analogue of PETSc's VecScatter creation for MATMPISELL): col_dev holds GLOBAL column
 * indices (-1 = none).  Entries outside [own_start, own_end) are collected into a sorted unique ghost list
 * (*ghost_gid_dev_out, device memory owned by the caller -> fsp_free) and col_dev is rewritten in place to the
 * local encoding of fspmat_generate (>= 0 local, -1 none, <= -2 ghost slot -(col+2)). */
FSP_API int fspmat_build_ghosts(int *col_dev, long n_entries, int own_start, int own_end, int **ghost_gid_dev_out,
                                long *n_ghost);
FSP_API int fspmat_shift_indices(int *idx_dev, long n, int delta);
/* Assembled A(t) in CSR form on the device (the PETSc Mat of CreateRHSJacobian / ComputeRHSJacobian,
 * src/Matrix/FspMatrixBase.cpp:308-427, FspMatrixConstrained.cpp:304-445): (P + 1) slots per state row (diagonal, then
 * one per reaction plane; entries on the same (i, j) stay separate slots and simply add in a product), the K sink rows
 * behind them.  structure != 0 writes row_ptr (n_rows + 1) and col as well; structure == 0 only refreshes val. */
FSP_API int fspmat_csr_size(fspmat_t h, long *nnz, int *n_rows);
FSP_API int fspmat_csr_export(fspmat_t h, const double *coef_host, int structure, int *row_ptr_dev, int *col_dev,
                              double *val_dev, void *stream);
FSP_API int fspmat_csr_spmv(int n_rows, const int *row_ptr_dev, const int *col_dev, const double *val_dev,
                            const double *x_dev, double *y_dev, void *stream);
/* dense export for tests: out_host is n_rows x n_rows column-major (ghost columns dropped) */
FSP_API int fspmat_dense(fspmat_t h, const double *coef_host, double *out_host);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU plumbing (one process per GPU).  Replaces PETSc VecScatter ghost exchange inside
 * MatMult(MATMPISELL), the sink VecScatter ADD (FspMatrixConstrained.cpp:57-60) and the
 * MPI_Allreduce behind VecDot/VecNorm (KrylovFsp.cpp:280-309).
 * NCCL is resolved at run time with dlopen (the already-loaded libnccl.so.2 if any).
 * ---------------------------------------------------------------------------------------------- */
typedef struct fspcomm_s *fspcomm_t;
#define FSPCOMM_ID_BYTES 128
FSP_API int fspcomm_unique_id(char id[FSPCOMM_ID_BYTES]);
FSP_API int fspcomm_create(fspcomm_t *out, const char id[FSPCOMM_ID_BYTES], int rank, int size);
FSP_API int fspcomm_destroy(fspcomm_t c);
FSP_API int fspcomm_rank(fspcomm_t c, int *rank, int *size);
FSP_API int fspcomm_allreduce_sum(fspcomm_t c, double *buf_dev, long n, void *stream);
FSP_API int fspcomm_allreduce_max(fspcomm_t c, double *buf_dev, long n, void *stream);
FSP_API int fspcomm_reduce_sum(fspcomm_t c, double *buf_dev, long n, int root, void *stream);
FSP_API int fspcomm_allgather_f64(fspcomm_t c, const double *send_dev, double *recv_dev, long n_per_rank, void *stream);
FSP_API int fspcomm_allgather_int(fspcomm_t c, const int *send_dev, int *recv_dev, long n_per_rank, void *stream);
/* halo exchange: send_counts/recv_counts are host arrays [size]; send buffer is packed per peer in rank
 * order, ghost buffer receives per peer in rank order. */
FSP_API int fspcomm_halo_exchange(fspcomm_t c, const double *send_dev, const long *send_counts_host,
                                  double *ghost_dev, const long *recv_counts_host, void *stream);
/* recv_host[p] = what rank p put in its send_host[me] (host arrays of length size) */
FSP_API int fspcomm_alltoall_counts(fspcomm_t c, const long *send_host, long *recv_host, void *stream);
/* int32 variant of the halo exchange (index lists at set-up time) */
FSP_API int fspcomm_exchange_int(fspcomm_t c, const int *send_dev, const long *send_counts_host, int *recv_dev,
                                 const long *recv_counts_host, void *stream);
/* pack: out[i] = x[idx[i]] is fspvec_gather */

/* ---- peer-memory fast path (one node, NVLink / NVSwitch; CUDA IPC windows mapped at fspcomm_create) ----------
 * When enabled (all ranks can map all peers; FSP_P2P=0 disables), fspcomm_allreduce_{sum,max} of n <=
 * FSP_P2P_MAX_REDUCE doubles is ONE kernel (store to every peer's slot, flag, wait, sum in rank order: deterministic
 * and bit-identical on all ranks), and the halo exchange of an Action is ONE kernel (fsphalo_begin: pack boundary
 * entries of x, store them into the peers' ghost buffers over NVLink, publish an epoch flag) whose consumers
 * (fspmat_action_boundary_p2p, fspmat_action_sinks_p2p) wait on the flags in device code.  These replace the
 * VecScatter of MatMult(MATMPISELL), the sink VecScatter ADD (FspMatrixConstrained.cpp:57-60) and the
 * MPI_Allreduce of VecDot/VecNorm without any host synchronisation or library call on the hot path. */
#define FSP_P2P_MAX_RANKS 16
#define FSP_P2P_MAX_REDUCE 64
#define FSP_P2P_MAX_SINKS 64
FSP_API int fspcomm_p2p_enabled(fspcomm_t c);
typedef struct fsphalo_s *fsphalo_t;
typedef struct fsphalo_epoch {
  unsigned long long        epoch;
  int                       n_ranks, self_rank;
  const double             *ghost;            /* this epoch's local ghost buffer (filled by the peers) */
  const unsigned long long *halo_flags;       /* local [n_ranks]: peer p has delivered when flags[p] >= epoch */
  const unsigned long long *sink_flags;       /* local [n_ranks] (meaningful on the sink owner = last rank) */
  const double             *sink_slots;       /* local [n_ranks][FSP_P2P_MAX_SINKS] partial sink sums */
  double                   *sink_slot_remote; /* the owner's slot row of THIS rank (peer memory) */
  unsigned long long       *sink_flag_remote; /* the owner's sink flag of THIS rank (peer memory) */
  unsigned int             *error_flag;       /* set by device code when a wait times out */
} fsphalo_epoch;
/* collective.  send_idx_dev: local indices of the x entries the peers need, packed per destination in rank order
 * (borrowed, must outlive the halo); ghost slots are laid out per source in rank order. */
FSP_API int fsphalo_create(fspcomm_t c, fsphalo_t *out, const int *send_idx_dev, const long *send_counts_host,
                           const long *recv_counts_host, int n_sink);
FSP_API int fsphalo_destroy(fsphalo_t h);
/* start of an Action: launches the fused pack + store + signal kernel on `stream` and describes the epoch */
FSP_API int fsphalo_begin(fsphalo_t h, const double *x_dev, void *stream, fsphalo_epoch *out);
/* The same without a launch: *out describes the consumer side of the next epoch and *push (opaque) the producer side,
 * for fspmat_action_halo, whose leading CTAs do the pack + store + signal themselves. */
typedef struct fsphalo_push { unsigned long long opaque[64]; } fsphalo_push;
FSP_API int fsphalo_next(fsphalo_t h, fsphalo_epoch *out, fsphalo_push *push);
/* non-zero (and fsp_last_error set) once a device-side flag wait of this communicator has timed out: results produced
 * since then were poisoned with NaN by the waiting kernels.  Call after a synchronisation that consumes results. */
FSP_API int fsphalo_check(fsphalo_t h);
FSP_API int fspcomm_check(fspcomm_t c);
FSP_API int fspcomm_alive(fspcomm_t c); /* 1 until fspcomm_destroy(c); safe to call with a stale pointer */
/* General peer-memory windows (used by the sharded state set).  create / destroy are collective; peers[p] is this
 * process' mapping of rank p's `bytes` bytes (peers[rank] = the local allocation).  retire is local: the window is
 * pooled for a later create of the same size and freed with the communicator. */
FSP_API int fspcomm_window_create(fspcomm_t c, size_t bytes, void **peers);
FSP_API int fspcomm_window_destroy(fspcomm_t c, void **peers);
FSP_API int fspcomm_window_retire(fspcomm_t c, void **peers, size_t bytes);
/* stream-ordered barrier over all ranks / synchronising gather of one integer per rank */
/* Personalised all-to-all of variable-size segments (esz = 4 or 8 bytes per element; send / recv hold the segments for /
 * from rank 0, 1, ... back to back).  With peer memory the senders store into the receivers' windows over NVLink; else
 * (or FSP_A2A=nccl) grouped ncclSend/ncclRecv.  Collective, synchronises the stream.  Replaces the MPI point-to-point
 * traffic of the reference's VecScatter set-up (MatAssembly) and of ExpandVec (src/Sys/PetscWrap.cpp:10-45). */
FSP_API int fspcomm_alltoallv(fspcomm_t c, const void *send_dev, const long *send_counts, void *recv_dev,
                              const long *recv_counts, int esz, void *stream);
FSP_API int fspcomm_barrier(fspcomm_t c, void *stream);
FSP_API int fspcomm_barrier_sync(fspcomm_t c); /* host-synchronising, through NCCL, no time limit */
FSP_API int fspcomm_gather_long(fspcomm_t c, long mine, long *all_host);
/* The whole multi-GPU Action (src/Matrix/FspMatrixBase.cpp:36-62 with the ghost VecScatter of MatMult on MATMPISELL,
 * and the sink VecScatter ADD of FspMatrixConstrained.cpp:57-60) as ONE launch on one stream, no events, no NCCL:
 *   leading CTAs   push: pack the boundary entries of x, store them into the peers' ghost windows, publish the epoch
 *   next CTAs      K partial sink sums -> the sink owner's slot row + flag
 *   row CTAs       1 row per thread in a rotated order that puts the CTAs with ghost rows last; a warp that meets a
 *                  ghost column waits (device code, acquire at system scope) for the peers' flags, then reads the
 *                  ghost entries at L2; every row is computed exactly once
 *   last CTA       waits for every peer's flag (paces the reuse of the two ghost buffers) and, on the sink owner, adds
 *                  the slots in rank order into y[n..n+K)
 * A wait that times out poisons the affected rows of y with NaN and raises the communicator's error flag. */
FSP_API int fspmat_action_halo(fspmat_t h, const double *coef_host, const double *x_dev, double *y_dev,
                               const fsphalo_epoch *e, const fsphalo_push *push, void *stream);
/* Pieces of fspmat_action_halo for callers that feed x in chunks (the host-vector pipeline, FspMatrixBase::ActionHost on
 * N ranks); all pieces of one Action share the (e, push) pair of ONE fsphalo_next call.
 *   parts            bit 0: push CTAs -- they read packed_send_dev[q] when it is non-NULL (the caller packed
 *                    x[send_idx[q]] itself, e.g. on the host) and x_dev[send_idx[q]] otherwise;
 *                    bit 1: the K partial sink sums of this rank (needs all of x_dev) -> owner's slots + flag;
 *                    bit 2: the finishing CTA (waits for every peer's flag; sink owner: slots -> y[n..n+K)).  Every
 *                    Action must run it exactly once, last: it paces the reuse of the two ghost buffers.
 *                    bit 3: this exchange carries no sink sums on any rank (halo-only diagnostics): the finishing CTA
 *                    then waits for the halo flags only.
 *   rows             [row_begin, row_end) of y; rows_have_ghosts != 0: the CTAs wait for the peers' flags first. */
FSP_API int fspmat_action_halo_part(fspmat_t h, const double *coef_host, const double *x_dev, double *y_dev,
                                    const fsphalo_epoch *e, const fsphalo_push *push, int parts, long row_begin,
                                    long row_end, int rows_have_ghosts, const double *packed_send_dev, void *stream);
/* chunk_flag_host[c] = 1 when rows [c*chunk_rows, (c+1)*chunk_rows) hold a row that references a ghost entry */
FSP_API int fspmat_chunk_has_ghost(fspmat_t h, long chunk_rows, int n_chunks, int *chunk_flag_host);
/* 1 when fspmat_action_halo covers this operator (values present, 1..16 reactions) */
FSP_API int fspmat_halo_fused_supported(fspmat_t h);

#ifdef __cplusplus
}
#endif
#endif /* FSP_B200_H_ */

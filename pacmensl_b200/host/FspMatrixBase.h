// FspMatrixBase.h -- the time-dependent FSP-truncated CME operator A(t) = sum_r c_r(t) A_r.
// Mirrors the public surface of src/Matrix/FspMatrixBase.h:53-194.  Storage and Action are the fused
// device operator of include/fsp_b200.h (fspmat_*): one kernel launch per Action instead of
// (R_tv + 1) x [MatMult + VecAXPY].
#pragma once

#include "Model.h"
#include "StateSetBase.h"
#include "StateSetConstrained.h"
#include "Sys.h"

// The PETSc Mat of CreateRHSJacobian / ComputeRHSJacobian (src/Matrix/FspMatrixBase.cpp:308-427): the assembled
// A(t) in CSR form on the device (fspmat_csr_export), applied by MatMult with a device SpMV.
struct _p_Mat {
  MPI_Comm                           comm = nullptr;
  int                                n_rows = 0, n_state_rows = 0;  ///< state rows come first; their first slot is the diagonal
  long                               nnz = 0;
  pacmensl::DeviceBuffer<int>        row_ptr, col;
  pacmensl::DeviceBuffer<double>     val;
};
typedef _p_Mat *Mat;
extern "C" {
PACMENSL_API PetscErrorCode MatMult(Mat A, Vec x, Vec y);
PACMENSL_API PetscErrorCode MatDestroy(Mat *A);
}

namespace pacmensl {
using Real = PetscReal;
using Int = PetscInt;

class PACMENSL_API FspMatrixBase {
 public:
  explicit FspMatrixBase(MPI_Comm comm);

  virtual PacmenslErrorCode GenerateValues(const StateSetBase &fsp, const Model &model);

  virtual PacmenslErrorCode GenerateValues(const StateSetBase &fsp, const arma::Mat<Int> &SM,
                                           std::vector<int> time_varying, const TcoefFun &new_prop_t,
                                           const PropFun &new_prop_x, const std::vector<int> &enable_reactions,
                                           void *prop_t_args, void *prop_x_args);

  PacmenslErrorCode SetTimeFun(TcoefFun new_t_fun, void *new_t_fun_args);

  virtual int Destroy();

  /// y = A(t) x.  Collective.  x must not alias y.
  virtual PacmenslErrorCode Action(PetscReal t, Vec x, Vec y);
  /// Extension: y = scale .* (beta x + alpha A(t) x) with up to two inner products of y fused into the same kernel
  /// (include/fsp_b200.h: fspmat_epilogue) -- what the Krylov and BDF/GMRES loops need right after every Action.
  virtual PacmenslErrorCode ActionFused(PetscReal t, Vec x, Vec y, const fspmat_epilogue &ep);
  /// Extension: the same operator on HOST vectors (n local rows each): chunked upload / compute / download pipeline
  /// on a single GPU, plain H2D + Action + D2H otherwise.  FSP_HOST_CHUNKS (default 32; <= 1 disables the pipeline).
  PacmenslErrorCode ActionHost(PetscReal t, const double *x_host, double *y_host);
  /// Extension: y = (d^j A / dt^j)(t) x = sum_{r in TV} c_r^(j)(t) A_r x for j >= 1 (j = 0: Action).  The derivatives of
  /// the time coefficients are central finite differences of the t_fun callback on a 7-point stencil of spacing delta
  /// (6th/4th-order accurate for j <= 2 / j <= 4).  Used by the Taylor restart of the BDF integrator (BdfCore.h).
  PacmenslErrorCode ActionTimeDerivative(int j, PetscReal t, Vec x, Vec y, PetscReal delta);
  bool HasTimeVaryingReactions() const { return !tv_reactions_.empty(); }
  /// Diagnostics (multi-GPU, peer-memory path): ONLY the exchange of an Action -- push CTAs (pack + store to the peers +
  /// flag) and the finishing CTA (wait for every peer) in one launch, no rows.  *bytes_sent = doubles pushed * 8.
  PacmenslErrorCode HaloExchangeOnly(Vec x, Vec y, long *bytes_sent);
  /// y = (sum_r coef[r] A_r) x with the coefficient vector supplied directly (used by SensFspMatrix).
  PacmenslErrorCode ActionWithCoefficients(const double *coefs, Vec x, Vec y);

  virtual PacmenslErrorCode CreateRHSJacobian(Mat *A);
  virtual PacmenslErrorCode ComputeRHSJacobian(PetscReal t, Mat A);

  virtual PacmenslErrorCode GetLocalMVFlops(PetscInt *nflops);
  int GetNumLocalRows() const { return num_rows_local_; }
  /// algorithmic bytes moved by one Action on this rank (SURVEY.md section 8d formula)
  double GetActionBytes() const;
  /// select a kernel variant (tuning / benchmarks)
  void SetKernelVariant(int v);
  /// Extension: keep the per-state propensities d_r(x_i) on the device across Destroy()/GenerateValues() cycles so
  /// that a regeneration after a state-set expansion evaluates the host callback only on the NEW states (indices of
  /// existing states never change).  The caller promises that the same model and the same, only-growing state set are
  /// used until ResetGenerationCache() / destruction.  FspSolverMultiSinks turns this on.  Single-rank only.
  void SetIncrementalGeneration(bool on) { incremental_ = on; if (!on) ResetGenerationCache(); }
  void ResetGenerationCache();

  virtual ~FspMatrixBase();

 protected:
  MPI_Comm comm_ = MPI_COMM_NULL;
  int      rank_ = 0, comm_size_ = 1;

  Int num_reactions_ = 0;
  Int num_rows_global_ = 0;
  Int num_rows_local_ = 0;
  Int num_states_local_ = 0;
  Int own_start_ = 0;

  std::vector<int> enable_reactions_, tv_reactions_, ti_reactions_;

  TcoefFun        t_fun_ = nullptr;
  void           *t_fun_args_ = nullptr;
  arma::Row<Real> time_coefficients_;

  fspmat_t  dmat_ = nullptr;
  PetscBool has_values_ = PETSC_FALSE;
  int       kernel_variant_ = 0;

  bool                 incremental_ = false;
  DeviceBuffer<double> diag_cache_;  ///< [num_reactions][cache_ld_] propensities d_r(x_i) by reaction id
  long                 cache_n_ = 0, cache_ld_ = 0;
  int                  cache_R_ = 0;
  std::vector<int>     cache_enabled_;
  // host-vector pipeline (ActionHost)
  DeviceBuffer<double> hx_, hy_;
  std::vector<int>     host_chunk_need_, host_chunk_ghost_, send_idx_host_;
  double              *pin_send_ = nullptr;
  long                 pin_send_cap_ = 0;
  long                 host_chunk_rows_ = 0;
  void                *up_stream_ = nullptr, *down_stream_ = nullptr, *host_compute_stream_ = nullptr;
  std::vector<void *>  ev_up_, ev_cmp_;
  // pinned double buffers + copy stream of the pipelined host-callback evaluation
  int    *pin_states_[2] = {nullptr, nullptr};
  double *pin_vals_[2] = {nullptr, nullptr};
  void   *pin_done_[2] = {nullptr, nullptr};
  void   *copy_stream_ = nullptr;
  int     pin_species_ = 0, pin_R_ = 0;
  long    pin_super_ = 0;
  PacmenslErrorCode ActionHostPartitioned_(const double *coefs, const double *x_host, double *y_host, int n_chunks, long chunk_rows);
  int  EvaluatePropensitiesHost_(fspset_t dset, int n_species, long first, long count, const PropFun &prop_x, void *prop_x_args);
  void FreePinned_();

  // set when GenerateValues(fsp, model) is used with a model that has a mass-action description
  std::shared_ptr<MassActionPropensity> mass_action_;

  // sink bookkeeping filled by the constrained subclass
  int  num_constraints_ = 0;
  bool owns_sinks_ = false;

  // multi-GPU halo exchange (ghost entries of x) and sink reduction
  long                 n_ghost_ = 0;
  DeviceBuffer<double> ghost_buf_, send_buf_, sink_buf_;
  void                *ev_push_done_ = nullptr;
  bool                 push_first_ = true;                   ///< see ActionWithCoefficients (peer-memory path)
  fsphalo_t            halo_ = nullptr;                      ///< peer-memory halo (fused pack+store+signal kernel); null => NCCL path
  DeviceBuffer<int>    send_idx_;
  std::vector<long>    send_counts_, recv_counts_;
  long                 n_send_ = 0;
  void                *comm_stream_ = nullptr;              ///< side stream carrying pack + halo + sink all-reduce
  void                *ev_x_ready_ = nullptr, *ev_comm_done_ = nullptr;

  virtual int DetermineLayout_(const StateSetBase &fsp);
  /// Fill the sink segments for plane order `planes` (constrained subclass); default: none.
  virtual int CollectSinks_(const StateSetBase &fsp, const arma::Mat<Int> &SM, const std::vector<int> &planes,
                            const double *diag_planes_dev, long ld, std::vector<long> &sink_ptr,
                            DeviceBuffer<int> &sink_idx, DeviceBuffer<double> &sink_val);
  int SetupGhosts_(const StateSetBase &fsp, int *col_planes_dev, long n_entries);
};

}  // namespace pacmensl

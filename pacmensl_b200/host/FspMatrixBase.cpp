#include "FspMatrixBase.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

PetscErrorCode MatMult(Mat A, Vec x, Vec y) {
  // y = J x with the assembled CSR Jacobian (device SpMV)
  if (!A || x->n_local != A->n_rows || y->n_local != A->n_rows) return -1;
  return fspmat_csr_spmv(A->n_rows, A->row_ptr.get(), A->col.get(), A->val.get(), x->d_data, y->d_data, x->comm ? x->comm->stream : nullptr);
}
PetscErrorCode MatDestroy(Mat *A) {
  if (A && *A) delete *A;
  if (A) *A = nullptr;
  return 0;
}

namespace pacmensl {

FspMatrixBase::FspMatrixBase(MPI_Comm comm) {
  comm_ = comm;
  MPI_Comm_rank(comm_, &rank_);
  MPI_Comm_size(comm_, &comm_size_);
}

FspMatrixBase::~FspMatrixBase() {
  Destroy();
  if (comm_stream_) {
    fsp_stream_sync(comm_stream_);
    fsp_stream_destroy(comm_stream_);
    fsp_event_destroy(ev_x_ready_);
    fsp_event_destroy(ev_comm_done_);
    fsp_event_destroy(ev_push_done_);
  }
  if (dmat_) fspmat_destroy(dmat_);
  dmat_ = nullptr;
  FreePinned_();
  if (pin_send_) fsp_free_host(pin_send_);
  pin_send_ = nullptr;
  for (void *e : ev_up_) fsp_event_destroy(e);
  for (void *e : ev_cmp_) fsp_event_destroy(e);
  for (void *s : {up_stream_, down_stream_, host_compute_stream_}) if (s) { fsp_stream_sync(s); fsp_stream_destroy(s); }
  comm_ = MPI_COMM_NULL;
}

// src/Matrix/FspMatrixBase.cpp:258-275
int FspMatrixBase::Destroy() {
  enable_reactions_.clear();
  tv_reactions_.clear();
  ti_reactions_.clear();
  if (comm_stream_) fsp_stream_sync(comm_stream_);  // side-stream kernels may still read the halo buffers
  if (dmat_) fspmat_clear(dmat_);
  if (halo_) { fsphalo_destroy(halo_); halo_ = nullptr; }
  host_chunk_need_.clear();
  host_chunk_ghost_.clear();
  send_idx_host_.clear();
  host_chunk_rows_ = 0;
  ghost_buf_.release();
  send_buf_.release();
  send_idx_.release();
  send_counts_.clear();
  recv_counts_.clear();
  n_ghost_ = n_send_ = 0;
  has_values_ = PETSC_FALSE;
  return 0;
}

// src/Matrix/FspMatrixBase.cpp:277-300
int FspMatrixBase::DetermineLayout_(const StateSetBase &fsp) {
  num_states_local_ = fsp.GetNumLocalStates();
  num_rows_local_ = num_states_local_;
  num_rows_global_ = fsp.GetNumGlobalStates();
  own_start_ = fsp.GetLocalStart();
  owns_sinks_ = false;
  num_constraints_ = 0;
  return 0;
}

int FspMatrixBase::CollectSinks_(const StateSetBase &, const arma::Mat<Int> &, const std::vector<int> &, const double *,
                                 long, std::vector<long> &sink_ptr, DeviceBuffer<int> &, DeviceBuffer<double> &) {
  sink_ptr.clear();
  return 0;
}

PacmenslErrorCode FspMatrixBase::GenerateValues(const StateSetBase &fsp, const Model &model) {
  mass_action_ = model.mass_action_;
  PacmenslErrorCode ierr = GenerateValues(fsp, model.stoichiometry_matrix_, model.tv_reactions_, model.prop_t_,
                                          model.prop_x_, std::vector<int>(), model.prop_t_args_, model.prop_x_args_);
  mass_action_.reset();
  return ierr;
}

// src/Matrix/FspMatrixBase.cpp:76-251.  For every enabled reaction r and local state x_i:
//   col = State2Index(x_i - nu_r) (device hash), off = prop_x(r, x_i - nu_r), diag = prop_x(r, x_i).
PacmenslErrorCode FspMatrixBase::GenerateValues(const StateSetBase &fsp, const arma::Mat<Int> &SM,
                                                std::vector<int> time_varying, const TcoefFun &new_prop_t,
                                                const PropFun &new_prop_x, const std::vector<int> &enable_reactions,
                                                void *prop_t_args, void *prop_x_args) {
  PacmenslErrorCode ierr;
  // FSP_GEN_TRACE=1: wall time of each phase of a generation (device-synchronised), for tuning the set-up path
  static const bool gen_trace = [] { const char *e = std::getenv("FSP_GEN_TRACE"); return e && e[0] == '1'; }();
  double            t_phase = 0.0;
  auto tick = [&](const char *what) {
    if (!gen_trace) return;
    fsp_device_sync();
    PetscLogDouble now;
    PetscTime(&now);
    if (what && rank_ == 0) printf("[gen n=%d] %-28s %8.2f ms\n", (int) fsp.GetNumLocalStates(), what, 1e3 * (now - t_phase));
    t_phase = now;
  };
  tick(nullptr);
  Destroy();
  ierr = DetermineLayout_(fsp);
  PACMENSLCHKERRQ(ierr);

  const int n_species = fsp.GetNumSpecies();
  const long n = fsp.GetNumLocalStates();
  num_reactions_ = fsp.GetNumReactions();
  time_coefficients_.set_size(num_reactions_);
  time_coefficients_.fill(1.0);

  t_fun_ = new_prop_t;
  t_fun_args_ = prop_t_args;

  enable_reactions_ = enable_reactions;
  if (enable_reactions_.empty()) {
    enable_reactions_.resize(num_reactions_);
    for (int i = 0; i < num_reactions_; ++i) enable_reactions_[i] = i;
  }
  for (int ir : enable_reactions_) {
    if (std::find(time_varying.begin(), time_varying.end(), ir) != time_varying.end()) tv_reactions_.push_back(ir);
    else ti_reactions_.push_back(ir);
  }
  if (!tv_reactions_.empty() && !t_fun_) {
    printf("FspMatrixBase::GenerateValues: time-varying reactions given without a time-coefficient function.\n");
    PACMENSLCHKERRQ(-1);
  }
  if (!new_prop_x && !mass_action_) PACMENSLCHKERRQ(-1);

  std::vector<int> planes(tv_reactions_);
  planes.insert(planes.end(), ti_reactions_.begin(), ti_reactions_.end());
  const int  P = (int) planes.size();
  const long ld = n > 0 ? n : 1;

  fspset_t dset = fsp.GetDeviceSet();
  if (!dset && n > 0) PACMENSLCHKERRQ(-1);

  DeviceBuffer<int>    col((size_t) P * ld);
  DeviceBuffer<double> off((size_t) P * ld), diag((size_t) P * ld);
  if (!col.get() || !off.get() || !diag.get()) PACMENSLCHKERRQ(-1);

  tick("destroy + plane buffers");
  std::vector<int>    shifted;
  std::vector<double> vals;
  const bool          on_device = (bool) mass_action_;
  std::vector<int>    zero_nu(n_species, 0);
  const long          first = fsp.GetLocalStart();

  if (on_device) {
    // device-evaluable (separable: mass action x per-species factor tables) propensities: everything stays on the GPU
    std::vector<double> tabs_host;
    std::map<std::pair<int, int>, int> tab_at;
    for (const auto &kv : mass_action_->table) {
      tab_at[kv.first] = (int) tabs_host.size();
      tabs_host.insert(tabs_host.end(), kv.second.begin(), kv.second.end());
    }
    DeviceBuffer<double> tabs;
    if (!tabs_host.empty() && tabs.upload(tabs_host.data(), tabs_host.size())) PACMENSLCHKERRQ(-1);
    for (int p = 0; p < P && n > 0; ++p) {
      const int  r = planes[p];
      const int *nu = SM.colptr(r);
      FSPCHKERRQ(fspset_lookup_shifted(dset, nu, -1, first, n, col.get() + (size_t) p * ld));  // State2Index(x - nu), :133-134
      std::vector<int> ord(n_species), toff(n_species, 0), tlen(n_species, 0);
      for (int s = 0; s < n_species; ++s) {
        ord[s] = mass_action_->order(s, r);
        auto it = mass_action_->table.find({s, r});
        if (it != mass_action_->table.end() && !it->second.empty()) { toff[s] = tab_at[{s, r}]; tlen[s] = (int) it->second.size(); }
      }
      FSPCHKERRQ(fspset_eval_separable(dset, mass_action_->rate[r], ord.data(), tabs.get(), toff.data(), tlen.data(), nu, -1, first, n,
                                       off.get() + (size_t) p * ld));
      FSPCHKERRQ(fspset_eval_separable(dset, mass_action_->rate[r], ord.data(), tabs.get(), toff.data(), tlen.data(), zero_nu.data(), 0,
                                       first, n, diag.get() + (size_t) p * ld));
    }
    FSPCHKERRQ(fsp_device_sync());  // the table buffer is released when this scope ends
  } else if (comm_size_ == 1) {
    // Single rank, host callbacks (API contract).  d_r(x_i) is evaluated on the host once per state -- with the
    // incremental cache only for states added since the previous generation -- and the off-diagonal values
    // d_r(x_i - nu_r) are gathered on the device from the source state's entry (off(i, r) == diag(col(i, r), r):
    // prop_x is a function of (reaction, state)); the reference instead calls prop_x a second time on every shifted
    // state (:135-136), including absent ones whose value it then drops.
    const bool reuse = incremental_ && cache_R_ == num_reactions_ && cache_enabled_ == enable_reactions_ && cache_n_ <= n;
    if (!reuse) { cache_n_ = 0; cache_R_ = num_reactions_; cache_enabled_ = enable_reactions_; }
    if (n > cache_ld_ || !diag_cache_.get()) {
      long new_ld = std::max<long>(n + n / 2 + 1024, 1);
      DeviceBuffer<double> bigger;
      if (bigger.resize((size_t) num_reactions_ * new_ld)) PACMENSLCHKERRQ(-1);
      for (int r = 0; r < num_reactions_ && cache_n_ > 0; ++r)
        FSPCHKERRQ(fsp_memcpy_d2d(bigger.get() + (size_t) r * new_ld, diag_cache_.get() + (size_t) r * cache_ld_,
                                  sizeof(double) * cache_n_, nullptr));
      FSPCHKERRQ(fsp_device_sync());
      diag_cache_.swap(bigger);
      cache_ld_ = new_ld;
    }
    const long n_new = n - cache_n_;
    if (n_new > 0) {
      ierr = EvaluatePropensitiesHost_(dset, n_species, cache_n_, n_new, new_prop_x, prop_x_args);
      PACMENSLCHKERRQ(ierr);
      cache_n_ = n;
    }
    for (int p = 0; p < P && n > 0; ++p) {
      const int r = planes[p];
      FSPCHKERRQ(fspset_lookup_shifted(dset, SM.colptr(r), -1, first, n, col.get() + (size_t) p * ld));
      FSPCHKERRQ(fsp_memcpy_d2d(diag.get() + (size_t) p * ld, diag_cache_.get() + (size_t) r * cache_ld_, sizeof(double) * n, nullptr));
      FSPCHKERRQ(fspvec_gather(off.get() + (size_t) p * ld, diag_cache_.get() + (size_t) r * cache_ld_,
                               col.get() + (size_t) p * ld, n, nullptr));
    }
    if (!incremental_) ResetGenerationCache();
  } else {
    // Multi-GPU, host callbacks: as the reference, prop_x on x - nu (:135-136) and on x (:180, :232) for the local rows
    const arma::Mat<int> *states = n > 0 ? &fsp.GetStatesRef() : nullptr;
    if (n > 0) { shifted.resize((size_t) n * n_species); vals.resize((size_t) n); }
    for (int p = 0; p < P && n > 0; ++p) {
      const int  r = planes[p];
      const int *nu = SM.colptr(r);
      FSPCHKERRQ(fspset_lookup_shifted(dset, nu, -1, first, n, col.get() + (size_t) p * ld));
      const int *X = states->memptr();
      for (long i = 0; i < n; ++i)
        for (int s = 0; s < n_species; ++s) shifted[(size_t) i * n_species + s] = X[(size_t) i * n_species + s] - nu[s];
      ierr = new_prop_x(r, n_species, (int) n, shifted.data(), vals.data(), prop_x_args);
      PACMENSLCHKERRQ(ierr);
      FSPCHKERRQ(fsp_memcpy_h2d(off.get() + (size_t) p * ld, vals.data(), sizeof(double) * n, nullptr));
      ierr = new_prop_x(r, n_species, (int) n, X, vals.data(), prop_x_args);
      PACMENSLCHKERRQ(ierr);
      FSPCHKERRQ(fsp_memcpy_h2d(diag.get() + (size_t) p * ld, vals.data(), sizeof(double) * n, nullptr));
    }
  }

  tick("columns + propensities");
  // sink rows (constrained subclass)
  std::vector<long>    sink_ptr;
  DeviceBuffer<int>    sink_idx;
  DeviceBuffer<double> sink_val;
  ierr = CollectSinks_(fsp, SM, planes, diag.get(), ld, sink_ptr, sink_idx, sink_val);
  PACMENSLCHKERRQ(ierr);

  tick("sink lists");
  // multi-GPU: columns outside the own block become ghost slots
  if (comm_size_ > 1) {
    ierr = SetupGhosts_(fsp, col.get(), (long) P * ld);
    PACMENSLCHKERRQ(ierr);
  }

  fspmat_desc d;
  std::memset(&d, 0, sizeof(d));
  d.n_states = (int) n;
  d.n_rows = num_rows_local_;
  d.n_reactions = num_reactions_;
  d.n_tv = (int) tv_reactions_.size();
  d.n_ti = (int) ti_reactions_.size();
  d.tv_reactions = tv_reactions_.data();
  d.ti_reactions = ti_reactions_.data();
  d.col = col.get();
  d.off = off.get();
  d.diag = diag.get();
  d.ld = ld;
  d.arrays_on_device = 1;
  d.n_constr = num_constraints_;
  d.sink_ptr = sink_ptr.empty() ? nullptr : sink_ptr.data();
  d.sink_idx = sink_idx.get();
  d.sink_val = sink_val.get();
  d.owns_sinks = owns_sinks_ ? 1 : 0;
  d.n_ghost = n_ghost_;
  std::vector<long> empty_ptr;
  if (num_constraints_ > 0 && sink_ptr.empty()) {
    empty_ptr.assign((size_t) P * num_constraints_ + 1, 0);
    d.sink_ptr = empty_ptr.data();
  }
  tick("ghost set-up");
  if (!dmat_) FSPCHKERRQ(fspmat_create(&dmat_));
  FSPCHKERRQ(fspmat_set_variant(dmat_, kernel_variant_));
  FSPCHKERRQ(fspmat_generate(dmat_, &d));
  tick("fspmat_generate (pack, nnz, sinks)");
  if (num_constraints_ > 0 && comm_size_ > 1) {
    if (sink_buf_.resize((size_t) num_constraints_)) PACMENSLCHKERRQ(-1);
  }
  if (comm_size_ > 1) {
    long n_int = 0, n_tot = 0;
    int  sms = 148;
    FSPCHKERRQ(fspmat_p2p_cta_counts(dmat_, &n_int, &n_tot));
    fsp_device_sm_count(&sms);
    push_first_ = n_int < 4L * 8L * sms;  // fewer than four full waves of ghost-free CTAs ahead of the waiting ones
  }
  has_values_ = PETSC_TRUE;
  return 0;
}

// Host prop_x callbacks (the API contract, src/Models/Model.h:44-60) on the states [first, first + count), results into
// diag_cache_.  The reference makes one call per reaction over ALL states (FspMatrixBase.cpp:135-136,180,232) and
// inserts the values one MatSetValue at a time.  Here the states are streamed from the device in super-chunks through
// pinned double buffers; inside a super-chunk the callback is invoked per reaction on cache-sized blocks (the state
// block is read from the host cache for every reaction instead of from DRAM R times), and the values of a finished
// super-chunk go to the device with asynchronous copies that overlap the evaluation of the next one.
int FspMatrixBase::EvaluatePropensitiesHost_(fspset_t dset, int n_species, long first, long count, const PropFun &prop_x,
                                             void *prop_x_args) {
  const long kBlock = 1L << 14;   // states per callback invocation
  const long kSuper = 1L << 18;   // states per pinned buffer
  const int  R = (int) enable_reactions_.size();
  if (count <= 0 || R == 0) return 0;
  // Opt-in (FSP_HOST_PIPELINE=1): on the GPU boxes of this pool the pipelined form measured SLOWER than one call per
  // reaction over all new states (hog1p, 23.9 M states: matrix generation 2.9-3.3 s vs 2.0 s on the same box) --
  // the pinned buffers are probably NUMA-remote for the calling thread, which costs the callbacks more than the
  // cache blocking and the copy overlap win.
  static const bool pipeline = [] { const char *e = std::getenv("FSP_HOST_PIPELINE"); return e && e[0] == '1'; }();
  if (count <= kBlock || !pipeline) {  // small increments: no pipeline needed
    std::vector<int>    st((size_t) count * n_species);
    std::vector<double> vals((size_t) count);
    FSPCHKERRQ(fspset_copy_states(dset, first, count, st.data()));
    for (int r : enable_reactions_) {
      int ierr = prop_x(r, n_species, (int) count, st.data(), vals.data(), prop_x_args);
      PACMENSLCHKERRQ(ierr);
      FSPCHKERRQ(fsp_memcpy_h2d(diag_cache_.get() + (size_t) r * cache_ld_ + first, vals.data(), sizeof(double) * count, nullptr));
    }
    return 0;
  }
  const long super = std::min(kSuper, (count + kBlock - 1) / kBlock * kBlock);
  if (!pin_states_[0] || pin_species_ < n_species || pin_R_ < R || pin_super_ < super) {
    FreePinned_();
    for (int b = 0; b < 2; ++b) {
      FSPCHKERRQ(fsp_malloc_host((void **) &pin_states_[b], sizeof(int) * (size_t) super * n_species));
      FSPCHKERRQ(fsp_malloc_host((void **) &pin_vals_[b], sizeof(double) * (size_t) super * R));
      FSPCHKERRQ(fsp_event_create(&pin_done_[b]));
    }
    FSPCHKERRQ(fsp_stream_create(&copy_stream_));
    pin_species_ = n_species; pin_R_ = R; pin_super_ = super;
  }
  FSPCHKERRQ(fsp_device_sync());  // diag_cache_ may just have been (re)allocated / copied on the main stream
  bool used[2] = {false, false};
  int  sc = 0;
  for (long s0 = 0; s0 < count; s0 += super, ++sc) {
    const int  b = sc & 1;
    const long m = std::min(super, count - s0);
    if (used[b]) FSPCHKERRQ(fsp_event_sync(pin_done_[b]));  // the copies that last read this buffer have finished
    FSPCHKERRQ(fspset_copy_states(dset, first + s0, m, pin_states_[b]));
    for (long o = 0; o < m; o += kBlock) {
      const int mb = (int) std::min(kBlock, m - o);
      int       q = 0;
      for (int r : enable_reactions_) {
        int ierr = prop_x(r, n_species, mb, pin_states_[b] + (size_t) o * n_species, pin_vals_[b] + (size_t) q * super + o, prop_x_args);
        PACMENSLCHKERRQ(ierr);
        ++q;
      }
    }
    int q = 0;
    for (int r : enable_reactions_) {
      FSPCHKERRQ(fsp_memcpy_h2d_async(diag_cache_.get() + (size_t) r * cache_ld_ + first + s0, pin_vals_[b] + (size_t) q * super,
                                      sizeof(double) * m, copy_stream_));
      ++q;
    }
    FSPCHKERRQ(fsp_event_record(pin_done_[b], copy_stream_));
    used[b] = true;
  }
  FSPCHKERRQ(fsp_stream_sync(copy_stream_));
  return 0;
}

void FspMatrixBase::FreePinned_() {
  for (int b = 0; b < 2; ++b) {
    if (pin_states_[b]) fsp_free_host(pin_states_[b]);
    if (pin_vals_[b]) fsp_free_host(pin_vals_[b]);
    if (pin_done_[b]) fsp_event_destroy(pin_done_[b]);
    pin_states_[b] = nullptr; pin_vals_[b] = nullptr; pin_done_[b] = nullptr;
  }
  if (copy_stream_) { fsp_stream_sync(copy_stream_); fsp_stream_destroy(copy_stream_); copy_stream_ = nullptr; }
  pin_species_ = pin_R_ = 0;
  pin_super_ = 0;
}

PacmenslErrorCode FspMatrixBase::SetTimeFun(TcoefFun new_t_fun, void *new_t_fun_args) {
  t_fun_ = new_t_fun;
  t_fun_args_ = new_t_fun_args;
  return 0;
}

// src/Matrix/FspMatrixBase.cpp:36-62 (+ FspMatrixConstrained.cpp:31-64 when sinks are present)
PacmenslErrorCode FspMatrixBase::Action(PetscReal t, Vec x, Vec y) {
  if (has_values_ == PETSC_FALSE) return VecSet(y, 0.0);
  if (!tv_reactions_.empty()) {
    int ierr = t_fun_(t, num_reactions_, time_coefficients_.memptr(), t_fun_args_);
    if (ierr != 0) VecSet(y, 0.0);
    PACMENSLCHKERRQ(ierr);
  }
  return ActionWithCoefficients(time_coefficients_.memptr(), x, y);
}

// y = scale .* (beta x + alpha A(t) x) plus up to two inner products of y in the same pass (fspmat_epilogue).
// Single GPU: one fused kernel.  Multi-GPU (or operators the fused kernel does not cover): Action followed by the
// equivalent vector passes; the inner products are then LOCAL partial sums -- the solver all-reduces them.
PacmenslErrorCode FspMatrixBase::ActionFused(PetscReal t, Vec x, Vec y, const fspmat_epilogue &ep) {
  if (has_values_ == PETSC_TRUE && !tv_reactions_.empty()) {
    int ierr = t_fun_(t, num_reactions_, time_coefficients_.memptr(), t_fun_args_);
    if (ierr != 0) VecSet(y, 0.0);
    PACMENSLCHKERRQ(ierr);
  }
  void *stream = comm_ ? comm_->stream : nullptr;
  if (has_values_ == PETSC_TRUE && comm_size_ == 1 && fspmat_fused_supported(dmat_)) {
    if (x->n_local != num_rows_local_ || y->n_local != num_rows_local_) PACMENSLCHKERRQ(-1);
    FSPCHKERRQ(fspmat_action_fused(dmat_, time_coefficients_.memptr(), x->d_data, y->d_data, &ep, stream));
    static const bool self_check = [] { const char *e = std::getenv("FSP_FUSED_CHECK"); return e && e[0] == '1'; }();
    if (self_check) {
      // diagnostics: recompute with Action + separate vector passes and compare (host synchronisation, slow)
      Vec z, u;
      VecDuplicate(y, &z);
      VecDuplicate(y, &u);
      PacmenslErrorCode ie = ActionWithCoefficients(time_coefficients_.memptr(), x, z);
      PACMENSLCHKERRQ(ie);
      FSPCHKERRQ(fspvec_wlincomb(u->d_data, ep.scale_dev, ep.beta, x->d_data, ep.alpha, z->d_data, y->n_local, stream));
      double un = 0, gap = 0, dref[2] = {0, 0}, dgot[2] = {0, 0};
      VecNorm(u, NORM_2, &un);
      for (int k = 0; k < ep.n_dots; ++k)
        fspvec_dot_h(&dref[k], u->d_data, ep.dot_vec_dev[k] ? ep.dot_vec_dev[k] : u->d_data, y->n_local, stream);
      if (ep.n_dots > 0) fsp_memcpy_d2h(dgot, ep.dot_out_dev, sizeof(double) * ep.n_dots, stream);
      VecAXPY(u, -1.0, y);
      VecNorm(u, NORM_2, &gap);
      static long calls = 0;
      ++calls;
      bool bad = !(gap <= 1e-12 * un) || !std::isfinite(un);
      for (int k = 0; k < ep.n_dots; ++k) bad = bad || !(std::fabs(dgot[k] - dref[k]) <= 1e-9 * (std::fabs(dref[k]) + un * un + 1e-300));
      if (bad)
        printf("[FSP_FUSED_CHECK] call %ld n=%d: |y|=%.6e |y_fused - y_ref|=%.3e dots fused (%.12e, %.12e) ref (%.12e, %.12e)\n", calls,
               y->n_local, un, gap, dgot[0], dgot[1], dref[0], dref[1]);
      VecDestroy(&z);
      VecDestroy(&u);
    }
    return 0;
  }
  PacmenslErrorCode ierr = ActionWithCoefficients(time_coefficients_.memptr(), x, y);
  PACMENSLCHKERRQ(ierr);
  const long n = y->n_local;
  if (ep.scale_dev || ep.alpha != 1.0 || ep.beta != 0.0)
    FSPCHKERRQ(fspvec_wlincomb(y->d_data, ep.scale_dev, ep.beta, x->d_data, ep.alpha, y->d_data, n, stream));
  if (ep.n_dots > 0) {
    const double *vecs[2] = {ep.dot_vec_dev[0] ? ep.dot_vec_dev[0] : y->d_data,
                             ep.n_dots > 1 && ep.dot_vec_dev[1] ? ep.dot_vec_dev[1] : y->d_data};
    FSPCHKERRQ(fspvec_mdot(ep.dot_out_dev, y->d_data, ep.n_dots, vecs, n, stream));
  }
  return 0;
}

PacmenslErrorCode FspMatrixBase::HaloExchangeOnly(Vec x, Vec y, long *bytes_sent) {
  if (bytes_sent) *bytes_sent = 8L * n_send_;
  if (comm_size_ == 1 || !halo_ || has_values_ == PETSC_FALSE) return 0;
  fsphalo_epoch ep;
  fsphalo_push  push;
  FSPCHKERRQ(fsphalo_next(halo_, &ep, &push));
  FSPCHKERRQ(fspmat_action_halo_part(dmat_, time_coefficients_.memptr(), x->d_data, y->d_data, &ep, &push, 1 | 4 | 8, 0, 0, 0, nullptr,
                                     comm_ ? comm_->stream : nullptr));
  return 0;
}

PacmenslErrorCode FspMatrixBase::ActionTimeDerivative(int j, PetscReal t, Vec x, Vec y, PetscReal delta) {
  if (j == 0) return Action(t, x, y);
  if (j < 0 || j > 4 || !(delta > 0.0)) return -1;
  if (has_values_ == PETSC_FALSE || tv_reactions_.empty()) return VecSet(y, 0.0);
  // central finite-difference weights on the offsets -3..3 (Fornberg), derivative orders 1..4
  static const double W[4][7] = {{-1.0 / 60, 3.0 / 20, -3.0 / 4, 0.0, 3.0 / 4, -3.0 / 20, 1.0 / 60},
                                 {1.0 / 90, -3.0 / 20, 3.0 / 2, -49.0 / 18, 3.0 / 2, -3.0 / 20, 1.0 / 90},
                                 {1.0 / 8, -1.0, 13.0 / 8, 0.0, -13.0 / 8, 1.0, -1.0 / 8},
                                 {-1.0 / 6, 2.0, -13.0 / 2, 28.0 / 3, -13.0 / 2, 2.0, -1.0 / 6}};
  arma::Row<PetscReal> cj((arma::uword) num_reactions_), tmp((arma::uword) num_reactions_);
  cj.fill(0.0);
  for (int k = -3; k <= 3; ++k) {
    const double w = W[j - 1][k + 3];
    if (w == 0.0) continue;
    tmp.fill(1.0);
    int ierr = t_fun_(t + k * delta, num_reactions_, tmp.memptr(), t_fun_args_);
    PACMENSLCHKERRQ(ierr);
    for (int r = 0; r < num_reactions_; ++r) cj[r] += w * tmp[r];
  }
  const double scale = 1.0 / std::pow(delta, j);
  for (int r = 0; r < num_reactions_; ++r) cj[r] *= scale;
  FSPCHKERRQ(fspmat_set_ti_coef(dmat_, 0.0));
  PacmenslErrorCode ierr = ActionWithCoefficients(cj.memptr(), x, y);
  fspmat_set_ti_coef(dmat_, 1.0);
  return ierr;
}

PacmenslErrorCode FspMatrixBase::ActionWithCoefficients(const double *coefs, Vec x, Vec y) {
  if (has_values_ == PETSC_FALSE) return VecSet(y, 0.0);
  if (x->n_local != num_rows_local_ || y->n_local != num_rows_local_) {
    printf("FspMatrixBase::Action: vector sizes (%d, %d) do not match the operator (%d local rows).\n", x->n_local,
           y->n_local, num_rows_local_);
    PACMENSLCHKERRQ(-1);
  }
  void *stream = comm_ ? comm_->stream : nullptr;
  if (comm_size_ == 1) {
    FSPCHKERRQ(fspmat_action(dmat_, coefs, x->d_data, nullptr, y->d_data, nullptr, stream));
    return 0;
  }
  // Multi-GPU: overlap communication with the interior rows.
  //   side stream : pack boundary entries of x -> NCCL halo exchange (replaces the VecScatter inside MatMult on
  //                 MATMPISELL) -> K partial sink sums -> K-double all-reduce (FspMatrixConstrained.cpp:57-60)
  //   main stream : interior pass over all rows (ghost entries counted as 0)
  //   join        : rows that reference ghost entries are recomputed with the received halo; the owner of the sink
  //                 rows copies the reduced sums into y.
  static const int mode = [] {  // FSP_MULTIGPU_MODE: overlap (default) | sequential | nocomm (timing diagnostics only)
    const char *e = std::getenv("FSP_MULTIGPU_MODE");
    if (!e) return 0;
    if (!std::strcmp(e, "sequential")) return 1;
    if (!std::strcmp(e, "nocomm")) return 2;
    return 0;
  }();
  if (mode == 2) {
    FSPCHKERRQ(fspmat_action_phase(dmat_, coefs, x->d_data, nullptr, y->d_data, nullptr, 1, stream));
    return 0;
  }
  if (mode == 1) {
    if (n_send_ > 0) FSPCHKERRQ(fspvec_gather(send_buf_.get(), x->d_data, send_idx_.get(), n_send_, stream));
    if (n_send_ > 0 || n_ghost_ > 0)
      FSPCHKERRQ(fspcomm_halo_exchange(comm_->nccl, send_buf_.get(), send_counts_.data(), ghost_buf_.get(),
                                       recv_counts_.data(), stream));
    double *so = num_constraints_ > 0 ? sink_buf_.get() : nullptr;
    FSPCHKERRQ(fspmat_action(dmat_, coefs, x->d_data, ghost_buf_.get(), y->d_data, so, stream));
    if (so) {
      FSPCHKERRQ(fspcomm_allreduce_sum(comm_->nccl, so, num_constraints_, stream));
      if (owns_sinks_) FSPCHKERRQ(fsp_memcpy_d2d(y->d_data + num_states_local_, so, sizeof(double) * num_constraints_, stream));
    }
    return 0;
  }
  if (!comm_stream_) {
    FSPCHKERRQ(fsp_stream_create(&comm_stream_));
    FSPCHKERRQ(fsp_event_create(&ev_x_ready_));
    FSPCHKERRQ(fsp_event_create(&ev_comm_done_));
    FSPCHKERRQ(fsp_event_create(&ev_push_done_));
  }
  if (halo_) {
    // Peer-memory path (NVLink/NVSwitch, no NCCL call, no host synchronisation):
    //   side stream : ONE kernel packs the boundary entries of x, stores them into the peers' ghost windows and
    //                 publishes the epoch flag; the sink kernel stores the K partial sums into the owner's slots
    //   main stream : interior pass; then the boundary kernel waits for the peers' flags in device code, redoes the
    //                 rows with ghost entries and (sink owner) adds the slots in rank order into y[n..n+K)
    fsphalo_epoch ep;
    // Default: ONE launch on the caller's stream (fspmat_action_halo: push CTAs, sink CTAs, row CTAs in rotated order
    // with a device-side wait where a warp meets a ghost column, finishing CTA) -- no side stream, no events, every row
    // computed once.  FSP_P2P_MODE=split keeps round 1's interior pass + boundary kernel (rows with ghost entries
    // computed twice, 4 launches on 2 streams); FSP_P2P_MODE=order (or FSP_P2P_SINGLE=1) its single-kernel form with a
    // CTA order table, which measured 0.96 ms against 0.74 ms (split) per Action on 2 GPUs: the table look-up at the
    // start of every CTA is a third dependent memory round trip per row.
    static const int p2p_mode = [] {
      const char *e = std::getenv("FSP_P2P_MODE"), *s1 = std::getenv("FSP_P2P_SINGLE");
      if (s1 && s1[0] == '1') return 2;
      if (e && !std::strcmp(e, "split")) return 1;
      if (e && !std::strcmp(e, "order")) return 2;
      return 0;
    }();
    if (p2p_mode == 0 && fspmat_halo_fused_supported(dmat_)) {
      fsphalo_push push;
      FSPCHKERRQ(fsphalo_next(halo_, &ep, &push));
      FSPCHKERRQ(fspmat_action_halo(dmat_, coefs, x->d_data, y->d_data, &ep, &push, stream));
      return 0;
    }
    const bool split = p2p_mode != 2;
    FSPCHKERRQ(fsp_event_record(ev_x_ready_, stream));
    FSPCHKERRQ(fsp_stream_wait_event(comm_stream_, ev_x_ready_));
    FSPCHKERRQ(fsphalo_begin(halo_, x->d_data, comm_stream_, &ep));
    if (!split && push_first_) {
      // few ghost-free CTAs: CTAs waiting for the peers could fill every SM slot before this GPU's own push kernel has
      // been scheduled (and every GPU could do the same to its peers) -- let the push complete first
      FSPCHKERRQ(fsp_event_record(ev_push_done_, comm_stream_));
      FSPCHKERRQ(fsp_stream_wait_event(stream, ev_push_done_));
    }
    if (num_constraints_ > 0) FSPCHKERRQ(fspmat_action_sinks_p2p(dmat_, coefs, x->d_data, &ep, comm_stream_));
    FSPCHKERRQ(fsp_event_record(ev_comm_done_, comm_stream_));
    if (!split) {
      // ONE kernel for all rows: ghost-free CTAs first, CTAs with ghost rows wait for the peers' flags in device code
      FSPCHKERRQ(fspmat_action_p2p(dmat_, coefs, x->d_data, y->d_data, &ep, stream));
      // join: x must not be modified by later work on the main stream while the push / sink kernels still read it
      FSPCHKERRQ(fsp_stream_wait_event(stream, ev_comm_done_));
      return 0;
    }
    // interior pass over all rows (ghost entries count 0), then a boundary kernel that waits for the peers' flags in
    // device code and redoes the rows with ghost entries (+ the sink owner's final sum)
    FSPCHKERRQ(fspmat_action_phase(dmat_, coefs, x->d_data, nullptr, y->d_data, nullptr, 1, stream));
    FSPCHKERRQ(fsp_stream_wait_event(stream, ev_comm_done_));
    FSPCHKERRQ(fspmat_action_boundary_p2p(dmat_, coefs, x->d_data, y->d_data, &ep, stream));
    return 0;
  }
  // optional timeline (FSP_MULTIGPU_TRACE=n prints event timings of the first n Actions after 30 warm-up calls)
  static const int trace_n = [] { const char *e = std::getenv("FSP_MULTIGPU_TRACE"); return e ? std::atoi(e) : 0; }();
  static int   trace_calls = 0;
  static void *tev[8] = {nullptr};
  const bool   tracing = trace_n > 0 && trace_calls >= 30 && trace_calls < 30 + trace_n;
  trace_calls++;
  if (tracing && !tev[0]) for (auto &e : tev) fsp_event_create(&e);
  if (tracing) fsp_event_record(tev[0], stream);

  FSPCHKERRQ(fsp_event_record(ev_x_ready_, stream));
  FSPCHKERRQ(fsp_stream_wait_event(comm_stream_, ev_x_ready_));
  if (n_send_ > 0) FSPCHKERRQ(fspvec_gather(send_buf_.get(), x->d_data, send_idx_.get(), n_send_, comm_stream_));
  if (tracing) fsp_event_record(tev[1], comm_stream_);
  if (n_send_ > 0 || n_ghost_ > 0)
    FSPCHKERRQ(fspcomm_halo_exchange(comm_->nccl, send_buf_.get(), send_counts_.data(), ghost_buf_.get(),
                                     recv_counts_.data(), comm_stream_));
  if (tracing) fsp_event_record(tev[2], comm_stream_);
  double *sink_out = num_constraints_ > 0 ? sink_buf_.get() : nullptr;
  if (sink_out) {
    FSPCHKERRQ(fspmat_action_phase(dmat_, coefs, x->d_data, nullptr, y->d_data, sink_out, 3, comm_stream_));
    if (tracing) fsp_event_record(tev[3], comm_stream_);
    FSPCHKERRQ(fspcomm_allreduce_sum(comm_->nccl, sink_out, num_constraints_, comm_stream_));
  }
  if (tracing) fsp_event_record(tev[4], comm_stream_);
  FSPCHKERRQ(fsp_event_record(ev_comm_done_, comm_stream_));
  FSPCHKERRQ(fspmat_action_phase(dmat_, coefs, x->d_data, nullptr, y->d_data, nullptr, 1, stream));
  if (tracing) fsp_event_record(tev[5], stream);
  FSPCHKERRQ(fsp_stream_wait_event(stream, ev_comm_done_));
  if (n_ghost_ > 0) FSPCHKERRQ(fspmat_action_phase(dmat_, coefs, x->d_data, ghost_buf_.get(), y->d_data, nullptr, 2, stream));
  if (sink_out && owns_sinks_)
    FSPCHKERRQ(fsp_memcpy_d2d(y->d_data + num_states_local_, sink_out, sizeof(double) * num_constraints_, stream));
  if (tracing) {
    fsp_event_record(tev[6], stream);
    float t[7] = {0};
    for (int k = 1; k <= 6; ++k) fsp_event_elapsed_ms(tev[0], tev[k], &t[k]);
    if (rank_ == 0)
      printf("[trace] pack %.3f | halo %.3f | sinks %.3f | allreduce %.3f || interior %.3f | end %.3f ms\n", t[1], t[2], t[3],
             t[4], t[5], t[6]);
  }
  return 0;
}

// y_host = A(t) x_host with HOST vectors (what a caller with host-resident PETSc vectors gets; bench.py's e2e number).
// Single GPU: a three-stage pipeline over row chunks -- x goes up in chunks on one copy stream, a chunk of rows runs as
// soon as the prefix of x it references has arrived (fspmat_chunk_max_columns, cached per generation), and its part of
// y goes down on a second copy stream while later chunks upload and compute -- so the PCIe link is used in both
// directions at once instead of H2D, Action, D2H one after the other.
PacmenslErrorCode FspMatrixBase::ActionHost(PetscReal t, const double *x_host, double *y_host) {
  const long n = num_rows_local_, ns = num_states_local_;
  if (hx_.resize((size_t) std::max<long>(n, 1)) || hy_.resize((size_t) std::max<long>(n, 1))) PACMENSLCHKERRQ(-1);
  static const int n_chunks_env = [] { const char *e = std::getenv("FSP_HOST_CHUNKS"); return e ? std::atoi(e) : 32; }();
  // every rank of a multi-GPU job must take the same path (the peer-memory pipeline is collective): decide on the
  // GLOBAL size there
  const long size_key = comm_size_ == 1 ? ns : (long) num_rows_global_ / comm_size_;
  const bool pipelined = has_values_ == PETSC_TRUE && n_chunks_env > 1 && size_key >= 64L * 256 * n_chunks_env &&
                         (int) (tv_reactions_.size() + ti_reactions_.size()) <= 16 &&
                         (comm_size_ == 1 || (halo_ != nullptr && fspmat_halo_fused_supported(dmat_)));
  if (!pipelined) {
    FSPCHKERRQ(fsp_memcpy_h2d(hx_.get(), x_host, sizeof(double) * n, nullptr));
    _p_Vec x, y;
    x.comm = y.comm = comm_;
    x.n_local = y.n_local = (PetscInt) n;
    x.d_data = hx_.get();
    y.d_data = hy_.get();
    x.owns_data = y.owns_data = false;
    PacmenslErrorCode ierr = Action(t, &x, &y);
    PACMENSLCHKERRQ(ierr);
    FSPCHKERRQ(fsp_memcpy_d2h(y_host, hy_.get(), sizeof(double) * n, nullptr));
    return 0;
  }
  if (!tv_reactions_.empty()) {
    int ierr = t_fun_(t, num_reactions_, time_coefficients_.memptr(), t_fun_args_);
    PACMENSLCHKERRQ(ierr);
  }
  FSPCHKERRQ(fsp_stream_sync(nullptr));  // hx_/hy_ come from the stream-ordered pool of the main stream
  const int  C = n_chunks_env;
  const long chunk = ((ns + C - 1) / C + 255) / 256 * 256;
  if (host_chunk_need_.empty() || host_chunk_rows_ != chunk) {
    host_chunk_need_.assign((size_t) C, -1);
    host_chunk_ghost_.clear();
    FSPCHKERRQ(fspmat_chunk_max_columns(dmat_, chunk, C, host_chunk_need_.data()));
    host_chunk_rows_ = chunk;
  }
  if (!up_stream_) {
    FSPCHKERRQ(fsp_stream_create(&up_stream_));
    FSPCHKERRQ(fsp_stream_create(&down_stream_));
    FSPCHKERRQ(fsp_stream_create(&host_compute_stream_));
  }
  while ((int) ev_up_.size() < C + 1) { void *e = nullptr; FSPCHKERRQ(fsp_event_create(&e)); ev_up_.push_back(e); }
  while ((int) ev_cmp_.size() < C + 1) { void *e = nullptr; FSPCHKERRQ(fsp_event_create(&e)); ev_cmp_.push_back(e); }
  const double *coefs = time_coefficients_.memptr();
  if (comm_size_ > 1) return ActionHostPartitioned_(coefs, x_host, y_host, C, chunk);
  // FSP_HOST_TRACE=1: the time line of the pipeline (event times relative to the first upload), for tuning
  static const bool host_trace = [] { const char *e = std::getenv("FSP_HOST_TRACE"); return e && e[0] == '1'; }();
  std::vector<void *> ev_down;
  void               *ev_start = nullptr;
  if (host_trace) {
    FSPCHKERRQ(fsp_event_create(&ev_start));
    for (int k = 0; k < C; ++k) { void *e = nullptr; FSPCHKERRQ(fsp_event_create(&e)); ev_down.push_back(e); }
    FSPCHKERRQ(fsp_event_record(ev_start, up_stream_));
  }
  // stage 1: all uploads, in order (chunk k of x, the sink entries ride with the last chunk)
  for (int k = 0; k < C; ++k) {
    const long b = std::min<long>(ns, (long) k * chunk), e = (k == C - 1) ? n : std::min<long>(ns, (long) (k + 1) * chunk);
    if (e > b) FSPCHKERRQ(fsp_memcpy_h2d_async(hx_.get() + b, x_host + b, sizeof(double) * (e - b), up_stream_));
    FSPCHKERRQ(fsp_event_record(ev_up_[k], up_stream_));
  }
  // stages 2 + 3: rows of chunk c as soon as x[0 .. need_c] is on the device; its y right behind it
  for (int c = 0; c < C; ++c) {
    const long b = std::min<long>(ns, (long) c * chunk), e = std::min<long>(ns, (long) (c + 1) * chunk);
    if (e <= b) continue;
    const long need = std::max<long>(host_chunk_need_[c], e - 1);
    const int  k_need = (int) std::min<long>(C - 1, need / chunk);
    FSPCHKERRQ(fsp_stream_wait_event(host_compute_stream_, ev_up_[k_need]));
    FSPCHKERRQ(fspmat_action_rows(dmat_, coefs, hx_.get(), hy_.get(), b, e, 0, host_compute_stream_));
    FSPCHKERRQ(fsp_event_record(ev_cmp_[c], host_compute_stream_));
    FSPCHKERRQ(fsp_stream_wait_event(down_stream_, ev_cmp_[c]));
    FSPCHKERRQ(fsp_memcpy_d2h_async(y_host + b, hy_.get() + b, sizeof(double) * (e - b), down_stream_));
    if (host_trace) FSPCHKERRQ(fsp_event_record(ev_down[c], down_stream_));
  }
  if (n > ns) {  // sink rows need all of x
    FSPCHKERRQ(fsp_stream_wait_event(host_compute_stream_, ev_up_[C - 1]));
    FSPCHKERRQ(fspmat_action_rows(dmat_, coefs, hx_.get(), hy_.get(), ns, ns, 1, host_compute_stream_));
    FSPCHKERRQ(fsp_event_record(ev_cmp_[C], host_compute_stream_));
    FSPCHKERRQ(fsp_stream_wait_event(down_stream_, ev_cmp_[C]));
    FSPCHKERRQ(fsp_memcpy_d2h_async(y_host + ns, hy_.get() + ns, sizeof(double) * (n - ns), down_stream_));
  }
  FSPCHKERRQ(fsp_stream_sync(down_stream_));
  FSPCHKERRQ(fsp_stream_sync(host_compute_stream_));
  FSPCHKERRQ(fsp_stream_sync(up_stream_));
  if (host_trace) {
    static int calls = 0;
    if (++calls == 3) {  // a warm call
      printf("[host pipeline] %d chunks of %ld rows: chunk | x up by | rows done by | y down by  (ms after the first upload was queued)\n", C, chunk);
      for (int c = 0; c < C; ++c) {
        float u = 0, m = 0, d = 0;
        fsp_event_elapsed_ms(ev_start, ev_up_[c], &u);
        fsp_event_elapsed_ms(ev_start, ev_cmp_[c], &m);
        fsp_event_elapsed_ms(ev_start, ev_down[c], &d);
        printf("[host pipeline] %3d | %7.3f | %7.3f | %7.3f | needs x up to chunk %ld\n", c, u, m, d, std::min<long>(C - 1, std::max<long>(host_chunk_need_[c], 0) / chunk));
      }
    }
    for (void *e : ev_down) fsp_event_destroy(e);
    fsp_event_destroy(ev_start);
  }
  return 0;
}

// The same pipeline on every rank of a multi-GPU job (each rank has its own PCIe link).  Per rank:
//   1. the boundary entries of x that the peers need are packed on the HOST (x_host[send_idx]), uploaded first, and
//      pushed into the peers' ghost windows by the push CTAs (fspmat_action_halo_part, bit 0) -- the halo is in
//      flight while the bulk of x is still crossing PCIe;
//   2. x goes up in chunks; a chunk of rows without ghost columns runs as soon as the prefix of x it references has
//      arrived, its part of y goes down behind it; chunks WITH ghost columns run last (their CTAs wait for the peers'
//      flags in device code);
//   3. sink partial sums + the finishing CTA (pacing; on the sink owner the K sums) close the Action.
// Bit-identical to the device-vector Action (same row code, every row computed once).
PacmenslErrorCode FspMatrixBase::ActionHostPartitioned_(const double *coefs, const double *x_host, double *y_host, int C, long chunk) {
  const long n = num_rows_local_, ns = num_states_local_;
  if (host_chunk_ghost_.empty() || (long) host_chunk_ghost_.size() != C) {
    host_chunk_ghost_.assign((size_t) C, 0);
    FSPCHKERRQ(fspmat_chunk_has_ghost(dmat_, chunk, C, host_chunk_ghost_.data()));
  }
  if (n_send_ > 0 && send_idx_host_.empty()) {
    send_idx_host_.resize((size_t) n_send_);
    FSPCHKERRQ(fsp_memcpy_d2h(send_idx_host_.data(), send_idx_.get(), sizeof(int) * n_send_, nullptr));
  }
  if (n_send_ > 0 && (!pin_send_ || pin_send_cap_ < n_send_)) {
    if (pin_send_) fsp_free_host(pin_send_);
    FSPCHKERRQ(fsp_malloc_host((void **) &pin_send_, sizeof(double) * (size_t) n_send_));
    pin_send_cap_ = n_send_;
  }
  fsphalo_epoch ep;
  fsphalo_push  push;
  FSPCHKERRQ(fsphalo_next(halo_, &ep, &push));
  // 1. pack on the host, upload, push
  for (long q = 0; q < n_send_; ++q) pin_send_[q] = x_host[send_idx_host_[(size_t) q]];
  if (n_send_ > 0) FSPCHKERRQ(fsp_memcpy_h2d_async(send_buf_.get(), pin_send_, sizeof(double) * n_send_, up_stream_));
  FSPCHKERRQ(fspmat_action_halo_part(dmat_, coefs, hx_.get(), hy_.get(), &ep, &push, 1, 0, 0, 0, send_buf_.get(), up_stream_));
  // 2. uploads in order (the sink entries of x ride with the last chunk)
  for (int k = 0; k < C; ++k) {
    const long b = std::min<long>(ns, (long) k * chunk), e = (k == C - 1) ? n : std::min<long>(ns, (long) (k + 1) * chunk);
    if (e > b) FSPCHKERRQ(fsp_memcpy_h2d_async(hx_.get() + b, x_host + b, sizeof(double) * (e - b), up_stream_));
    FSPCHKERRQ(fsp_event_record(ev_up_[k], up_stream_));
  }
  for (int pass = 0; pass < 2; ++pass) {  // ghost-free chunks first, chunks that wait for the peers last
    for (int c = 0; c < C; ++c) {
      if ((host_chunk_ghost_[c] != 0) != (pass == 1)) continue;
      const long b = std::min<long>(ns, (long) c * chunk), e = std::min<long>(ns, (long) (c + 1) * chunk);
      if (e <= b) continue;
      const long need = std::max<long>(host_chunk_need_[c], e - 1);
      const int  k_need = (int) std::min<long>(C - 1, need / chunk);
      FSPCHKERRQ(fsp_stream_wait_event(host_compute_stream_, ev_up_[k_need]));
      FSPCHKERRQ(fspmat_action_halo_part(dmat_, coefs, hx_.get(), hy_.get(), &ep, &push, 0, b, e, pass, nullptr, host_compute_stream_));
      FSPCHKERRQ(fsp_event_record(ev_cmp_[c], host_compute_stream_));
      FSPCHKERRQ(fsp_stream_wait_event(down_stream_, ev_cmp_[c]));
      FSPCHKERRQ(fsp_memcpy_d2h_async(y_host + b, hy_.get() + b, sizeof(double) * (e - b), down_stream_));
    }
  }
  // 3. sinks + finish (needs all of x; every rank runs the finishing CTA: it paces the ghost-buffer reuse)
  FSPCHKERRQ(fsp_stream_wait_event(host_compute_stream_, ev_up_[C - 1]));
  FSPCHKERRQ(fspmat_action_halo_part(dmat_, coefs, hx_.get(), hy_.get(), &ep, &push, 2 | 4, 0, 0, 0, nullptr, host_compute_stream_));
  if (n > ns) {
    FSPCHKERRQ(fsp_event_record(ev_cmp_[C], host_compute_stream_));
    FSPCHKERRQ(fsp_stream_wait_event(down_stream_, ev_cmp_[C]));
    FSPCHKERRQ(fsp_memcpy_d2h_async(y_host + ns, hy_.get() + ns, sizeof(double) * (n - ns), down_stream_));
  }
  FSPCHKERRQ(fsp_stream_sync(down_stream_));
  FSPCHKERRQ(fsp_stream_sync(host_compute_stream_));
  FSPCHKERRQ(fsp_stream_sync(up_stream_));
  FSPCHKERRQ(fsphalo_check(halo_));
  return 0;
}

// src/Matrix/FspMatrixBase.cpp:308-427 (+ FspMatrixConstrained.cpp:304-445 for the sink rows): the assembled A(t) as a
// CSR matrix on the device.  Single rank only: the multi-GPU operator keeps ghost slots, not global columns.
PacmenslErrorCode FspMatrixBase::CreateRHSJacobian(Mat *A) {
  if (comm_size_ > 1) {
    printf("CreateRHSJacobian: the assembled Jacobian is provided on a single rank only.\n");
    return -1;
  }
  if (!dmat_) return -1;
  Mat J = new _p_Mat();
  J->comm = comm_;
  FSPCHKERRQ(fspmat_csr_size(dmat_, &J->nnz, &J->n_rows));
  J->n_rows = num_rows_local_;
  J->n_state_rows = num_states_local_;
  if (J->row_ptr.resize((size_t) J->n_rows + 1) || J->col.resize((size_t) std::max<long>(J->nnz, 1)) ||
      J->val.resize((size_t) std::max<long>(J->nnz, 1))) { delete J; return -1; }
  time_coefficients_.fill(1.0);
  if (fspmat_csr_export(dmat_, time_coefficients_.memptr(), 1, J->row_ptr.get(), J->col.get(), J->val.get(), comm_ ? comm_->stream : nullptr)) {
    delete J;
    return -1;
  }
  *A = J;
  return 0;
}
PacmenslErrorCode FspMatrixBase::ComputeRHSJacobian(PetscReal t, Mat A) {
  if (!A || !dmat_) return -1;
  if (!tv_reactions_.empty()) {
    int ierr = t_fun_(t, num_reactions_, time_coefficients_.memptr(), t_fun_args_);
    PACMENSLCHKERRQ(ierr);
  }
  long nnz = 0;
  int  nr = 0;
  FSPCHKERRQ(fspmat_csr_size(dmat_, &nnz, &nr));
  if (nnz != A->nnz || num_rows_local_ != A->n_rows) return -1;  // the operator was regenerated: create a new Jacobian
  FSPCHKERRQ(fspmat_csr_export(dmat_, time_coefficients_.memptr(), 0, A->row_ptr.get(), A->col.get(), A->val.get(), comm_ ? comm_->stream : nullptr));
  return 0;
}

// src/Matrix/FspMatrixBase.cpp:429-444
PacmenslErrorCode FspMatrixBase::GetLocalMVFlops(PetscInt *nflops) {
  long f = 0;
  if (dmat_) FSPCHKERRQ(fspmat_flops(dmat_, &f));
  *nflops = (PetscInt) f;
  return 0;
}

double FspMatrixBase::GetActionBytes() const {
  double b = 0.0;
  if (dmat_) fspmat_action_bytes(dmat_, &b);
  return b;
}

void FspMatrixBase::ResetGenerationCache() {
  diag_cache_.release();
  cache_n_ = cache_ld_ = 0;
  cache_R_ = 0;
  cache_enabled_.clear();
}

void FspMatrixBase::SetKernelVariant(int v) {
  kernel_variant_ = v;
  if (dmat_) fspmat_set_variant(dmat_, v);
}

int FspMatrixBase::SetupGhosts_(const StateSetBase &fsp, int *col_planes_dev, long n_entries) {
  // Build the per-peer ghost lists from the column entries outside the own block (what PETSc's
  // VecScatter set-up does for MATMPISELL) and rewrite col to the local/ghost encoding.
  static const bool gen_trace = [] { const char *e = std::getenv("FSP_GEN_TRACE"); return e && e[0] == '1'; }();
  double            t_phase = 0.0;
  auto tick = [&](const char *what) {
    if (!gen_trace) return;
    fsp_device_sync();
    PetscLogDouble now;
    PetscTime(&now);
    if (what && rank_ == 0) printf("[gen n=%d]   ghosts: %-24s %8.2f ms\n", (int) fsp.GetNumLocalStates(), what, 1e3 * (now - t_phase));
    t_phase = now;
  };
  tick(nullptr);
  const std::vector<int> &layout = fsp.GetLayout();
  const int own_start = layout[rank_], own_end = layout[rank_ + 1];
  int *ghost_gid_dev = nullptr;
  long n_ghost = 0;
  FSPCHKERRQ(fspmat_build_ghosts(col_planes_dev, n_entries, own_start, own_end, &ghost_gid_dev, &n_ghost));
  tick("select/sort/unique/remap");
  n_ghost_ = n_ghost;
  std::vector<int> gids((size_t) n_ghost);
  if (n_ghost > 0) FSPCHKERRQ(fsp_memcpy_d2h(gids.data(), ghost_gid_dev, sizeof(int) * n_ghost, nullptr));
  fsp_free(ghost_gid_dev);
  // ghost ids are sorted ascending => contiguous per owning peer
  recv_counts_.assign(comm_size_, 0);
  for (int g : gids) {
    int owner = (int) (std::upper_bound(layout.begin(), layout.end(), g) - layout.begin()) - 1;
    recv_counts_[owner] += 1;
  }
  // tell every peer which of its entries we need: exchange counts, then the index lists
  send_counts_.assign(comm_size_, 0);
  tick("ids to host, owners");
  FSPCHKERRQ(fspcomm_alltoall_counts(comm_->nccl, recv_counts_.data(), send_counts_.data(), comm_->stream));
  tick("count exchange");
  n_send_ = 0;
  for (int p = 0; p < comm_size_; ++p) n_send_ += send_counts_[p];
  DeviceBuffer<int> want;
  if (want.upload(gids.data(), gids.size())) return -1;
  if (send_idx_.resize((size_t) (n_send_ > 0 ? n_send_ : 1))) return -1;
  // we SEND our wanted global ids to their owners and RECEIVE the ids they want from us
  FSPCHKERRQ(fspcomm_alltoallv(comm_->nccl, want.get(), recv_counts_.data(), send_idx_.get(), send_counts_.data(), (int) sizeof(int),
                               comm_->stream));
  tick("id exchange (send/recv)");
  FSPCHKERRQ(fspmat_shift_indices(send_idx_.get(), n_send_, -own_start));  // global -> local
  if (ghost_buf_.resize((size_t) (n_ghost > 0 ? n_ghost : 1))) return -1;
  if (send_buf_.resize((size_t) (n_send_ > 0 ? n_send_ : 1))) return -1;
  // peer-memory halo (CUDA IPC windows over NVLink) when the communicator has it and the default overlap mode is on
  const char *mode = std::getenv("FSP_MULTIGPU_MODE");
  if (fspcomm_p2p_enabled(comm_->nccl) && (!mode || !std::strcmp(mode, "overlap")) && num_constraints_ <= FSP_P2P_MAX_SINKS) {
    tick("buffers");
    FSPCHKERRQ(fsphalo_create(comm_->nccl, &halo_, send_idx_.get(), send_counts_.data(), recv_counts_.data(), num_constraints_));
    tick("halo window");
  }
  return 0;
}

}  // namespace pacmensl

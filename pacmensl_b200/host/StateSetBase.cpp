#include "StateSetBase.h"

#include <algorithm>
#include <cstdlib>
#include <iostream>

namespace pacmensl {

StateSetBase::StateSetBase(MPI_Comm new_comm) {
  comm_ = new_comm;
  MPI_Comm_size(comm_, &comm_size_);
  MPI_Comm_rank(comm_, &my_rank_);
  ind_starts_.assign(comm_size_ + 1, 0);
}

StateSetBase::~StateSetBase() {
  Clear();
  comm_ = MPI_COMM_NULL;
}

// src/StateSet/StateSetBase.cpp:44-58
PacmenslErrorCode StateSetBase::SetStoichiometryMatrix(const arma::Mat<int> &SM) {
  if (num_species_ != 0 && (int) SM.n_rows != num_species_) {
    if (my_rank_ == 0) std::cout << "Input stoichiometry has incompatible dimension with the state set.\n";
    return -1;
  }
  num_species_ = (int) SM.n_rows;
  num_reactions_ = (int) SM.n_cols;
  stoichiometry_matrix_ = SM;
  stoich_set_ = 1;
  return 0;
}

// src/StateSet/StateSetBase.cpp:66-81
PacmenslErrorCode StateSetBase::SetNumSpecies(int num_species) {
  if (num_species <= 0) {
    if (my_rank_ == 0) std::cout << "Number of species must be positive.\n";
    return -1;
  }
  if (num_species_ != 0) {
    if (my_rank_ == 0) std::cout << "Warning: number of species already set. SetNumSpecies() is ignored.\n";
    return 0;
  }
  num_species_ = num_species;
  return 0;
}

PacmenslErrorCode StateSetBase::SetLoadBalancingScheme(PartitioningType type, PartitioningApproach) {
  // Every scheme maps to the contiguous equal-count BLOCK split: on a uniform NVSwitch fabric the
  // GRAPH/HYPERGRAPH heuristics (ParMETIS/PHG, not available) buy nothing for this operator.
  lb_type_ = type;
  return 0;
}

PacmenslErrorCode StateSetBase::SetUp() {
  set_up_ = true;
  return 0;
}

PacmenslErrorCode StateSetBase::ensure_device_set() {
  if (dset_) return 0;
  if (num_species_ <= 0) return -1;
  if (!stoich_set_) {
    stoichiometry_matrix_.set_size(num_species_, 0);
    num_reactions_ = 0;
  }
  FSPCHKERRQ(fspset_create(&dset_, num_species_, num_reactions_, stoichiometry_matrix_.memptr()));
  static const bool env_sharded = [] { const char *e = std::getenv("FSP_SHARDED_SET"); return !(e && e[0] == '0'); }();
  const bool sharded = want_sharded_ >= 0 ? want_sharded_ != 0 : env_sharded;
  if (sharded && comm_size_ > 1 && comm_ && fspcomm_p2p_enabled(comm_->nccl))
    FSPCHKERRQ(fspset_set_sharded(dset_, comm_->nccl));
  return 0;
}

PacmenslErrorCode StateSetBase::SetSharded(bool on) {
  if (dset_) {
    if (my_rank_ == 0) std::cout << "SetSharded() must be called before states are added; ignored.\n";
    return 0;
  }
  want_sharded_ = on ? 1 : 0;
  return 0;
}

PacmenslErrorCode StateSetBase::RememberLocalStates() {
  if (!dset_) { n_remembered_ = 0; return 0; }
  FSPCHKERRQ(fspset_remember_local(dset_));
  n_remembered_ = num_local_states_;
  return 0;
}
PacmenslErrorCode StateSetBase::RememberedIndices(std::vector<int> &indices) {
  indices.assign((size_t) n_remembered_, -1);
  if (!dset_) return 0;
  FSPCHKERRQ(fspset_remembered_indices(dset_, indices.data(), n_remembered_));
  n_remembered_ = 0;
  return 0;
}

PacmenslErrorCode StateSetBase::update_layout() {
  int n = 0;
  FSPCHKERRQ(fspset_num_states(dset_, &n));
  num_global_states_ = n;
  // BLOCK: ranks < rem own base + 1 states (contiguous ranges of the global ordering)
  const int base = n / comm_size_, rem = n % comm_size_;
  ind_starts_.assign(comm_size_ + 1, 0);
  for (int r = 0; r < comm_size_; ++r) ind_starts_[r + 1] = ind_starts_[r] + base + (r < rem ? 1 : 0);
  if (fspset_is_sharded(dset_)) {  // the device set made the same split when it re-balanced; take its word
    std::vector<long> st((size_t) comm_size_ + 1, 0);
    long              n_loc = 0;
    FSPCHKERRQ(fspset_layout(dset_, st.data(), &n_loc));
    for (int r = 0; r <= comm_size_; ++r) ind_starts_[r] = (int) st[(size_t) r];
    if (n_loc != ind_starts_[my_rank_ + 1] - ind_starts_[my_rank_]) PACMENSLCHKERRQ(-1);
  }
  local_start_ = ind_starts_[my_rank_];
  num_local_states_ = ind_starts_[my_rank_ + 1] - local_start_;
  host_states_valid_ = false;
  return 0;
}

// src/StateSet/StateSetBase.cpp:188-258
PacmenslErrorCode StateSetBase::AddStates(const arma::Mat<int> &X) {
  if (num_species_ != 0 && (int) X.n_rows != num_species_) return -1;
  if (num_species_ == 0) num_species_ = (int) X.n_rows;
  if (!set_up_) SetUp();
  PacmenslErrorCode ierr = ensure_device_set();
  PACMENSLCHKERRQ(ierr);
  if (comm_size_ == 1) {
    if (X.n_cols > 0) FSPCHKERRQ(fspset_add_states(dset_, (int) X.n_rows, (long) X.n_cols, X.memptr(), 0));
    return update_layout();
  }
  if (fspset_is_sharded(dset_)) {
    // collective inside: every rank inserts its own list into the striped directory, duplicates across ranks are shed
    // there, and the set is re-balanced to the BLOCK layout
    FSPCHKERRQ(fspset_add_states(dset_, (int) X.n_rows, (long) X.n_cols, X.memptr(), 0));
    return update_layout();
  }
  // Multi-GPU: each rank may pass its own (possibly overlapping) list (:176-178 of the reference); the replicated
  // directory receives the union, concatenated in rank order, so that every rank builds the same index map.
  const int S = num_species_;
  std::vector<double> counts(comm_size_, 0.0);
  counts[my_rank_] = (double) X.n_cols;
  ierr = pacmensl_allreduce_sum(comm_, counts.data(), comm_size_);
  PACMENSLCHKERRQ(ierr);
  long pad = 0;
  for (double c : counts) pad = std::max(pad, (long) c);
  if (pad > 0) {
    DeviceBuffer<int> loc((size_t) pad * S), all((size_t) pad * S * comm_size_);
    if (!loc.get() || !all.get()) PACMENSLCHKERRQ(-1);
    FSPCHKERRQ(fsp_memset(loc.get(), 0, sizeof(int) * pad * S, comm_->stream));
    if (X.n_cols > 0) FSPCHKERRQ(fsp_memcpy_h2d(loc.get(), X.memptr(), sizeof(int) * X.n_elem, comm_->stream));
    FSPCHKERRQ(fspcomm_allgather_int(comm_->nccl, loc.get(), all.get(), pad * S, comm_->stream));
    for (int r = 0; r < comm_size_; ++r)
      if (counts[r] > 0)
        FSPCHKERRQ(fspset_add_states(dset_, S, (long) counts[r], all.get() + (size_t) r * pad * S, 1));
  }
  return update_layout();
}

PacmenslErrorCode StateSetBase::AddBoxLattice(const arma::Row<int> &upper) {
  if (num_species_ != 0 && (int) upper.n_elem != num_species_) return -1;
  if (num_species_ == 0) num_species_ = (int) upper.n_elem;
  if (!set_up_) SetUp();
  PacmenslErrorCode ierr = ensure_device_set();
  PACMENSLCHKERRQ(ierr);
  FSPCHKERRQ(fspset_add_box_lattice(dset_, upper.memptr()));
  return update_layout();
}

// src/StateSet/StateSetBase.cpp:309-343
arma::Row<int> StateSetBase::State2Index(const arma::Mat<int> &state) const {
  arma::Row<int> indices((arma::uword) state.n_cols);
  State2Index((int) state.n_cols, state.memptr(), indices.memptr());
  return indices;
}
void StateSetBase::State2Index(arma::Mat<int> &state, int *indx) const {
  State2Index((int) state.n_cols, state.memptr(), indx);
}
void StateSetBase::State2Index(int num_states, const int *state, int *indx) const {
  if (num_states <= 0) return;
  if (!dset_) {
    for (int i = 0; i < num_states; ++i) indx[i] = -1;
    return;
  }
  if (fspset_state2index(dset_, num_states, state, 0, indx, 0) != 0)
    throw std::runtime_error(std::string("State2Index failed: ") + fsp_last_error());
}

MPI_Comm StateSetBase::GetComm() const { return comm_; }
int StateSetBase::GetNumLocalStates() const { return num_local_states_; }
int StateSetBase::GetNumGlobalStates() const { return num_global_states_; }
int StateSetBase::GetNumSpecies() const { return num_species_; }
int StateSetBase::GetNumReactions() const { return (int) stoichiometry_matrix_.n_cols; }

const arma::Mat<int> &StateSetBase::GetStatesRef() const {
  if (!host_states_valid_) {
    local_states_.set_size(num_species_, num_local_states_);
    if (dset_ && num_local_states_ > 0)
      if (fspset_copy_states(dset_, local_start_, num_local_states_, local_states_.memptr()) != 0)
        throw std::runtime_error(std::string("GetStatesRef failed: ") + fsp_last_error());
    host_states_valid_ = true;
  }
  return local_states_;
}
arma::Mat<int> StateSetBase::CopyStatesOnProc() const { return arma::Mat<int>(GetStatesRef()); }
void StateSetBase::CopyStatesOnProc(int num_local_states, int *state_array) const {
  assert(num_local_states == num_local_states_);
  const arma::Mat<int> &X = GetStatesRef();
  std::memcpy(state_array, X.memptr(), sizeof(int) * (size_t) num_local_states * num_species_);
}
std::tuple<int, int> StateSetBase::GetOrderingStartEnd() const {
  return std::make_tuple(local_start_, local_start_ + num_local_states_);
}

PacmenslErrorCode StateSetBase::Clear() {
  if (dset_) fspset_destroy(dset_);
  dset_ = nullptr;
  num_global_states_ = num_local_states_ = local_start_ = 0;
  local_states_.reset();
  host_states_valid_ = false;
  return 0;
}

}  // namespace pacmensl

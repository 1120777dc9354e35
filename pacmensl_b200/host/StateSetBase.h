// StateSetBase.h -- the FSP state space: state list + state -> index directory.
// Mirrors src/StateSet/StateSetBase.h:61-209.  The Zoltan distributed directory and the Armadillo
// column bookkeeping are replaced by a device-resident hash directory (fspset_* in fsp_b200.h).
//
// Multi-GPU layout (one process per GPU): the directory is REPLICATED on every rank (4 B/slot hash +
// 4*S B/state -- 4.4 GB at 1e8 three-species states against 180 GB of HBM), every rank runs the same
// deterministic expansion, and ownership is the contiguous equal-count BLOCK split of the global
// ordering (Zoltan LB_METHOD=BLOCK with unit weights, src/Partitioner/StatePartitionerBase.cpp:81-83,144).
// Global index = insertion order, so indices never change when the set grows; only block boundaries do.
// Collective calls (AddStates, Expand) must be given identical arguments on all ranks.
#pragma once

#include <tuple>

#include "Sys.h"

namespace pacmensl {

class PACMENSL_API StateSetBase {
 public:
  NOT_COPYABLE_NOT_MOVABLE(StateSetBase);

  explicit StateSetBase(MPI_Comm new_comm);

  PacmenslErrorCode SetNumSpecies(int num_species);
  PacmenslErrorCode SetStoichiometryMatrix(const arma::Mat<int> &SM);
  PacmenslErrorCode SetLoadBalancingScheme(PartitioningType type,
                                           PartitioningApproach approach = PartitioningApproach::REPARTITION);
  virtual PacmenslErrorCode SetUp();
  PacmenslErrorCode AddStates(const arma::Mat<int> &X);
  /// Extension: add the full lexicographic box lattice 0..upper[s] (species 0 fastest; the sub2ind_nd ordering of
  /// src/Sys/pacmenMath.h:33-59), generated on the device -- the synthetic workload of SURVEY.md section 8(d).
  PacmenslErrorCode AddBoxLattice(const arma::Row<int> &upper);

  arma::Row<int> State2Index(const arma::Mat<int> &state) const;
  void State2Index(arma::Mat<int> &state, int *indx) const;
  void State2Index(int num_states, const int *state, int *indx) const;

  MPI_Comm GetComm() const;
  int GetNumLocalStates() const;
  int GetNumGlobalStates() const;
  int GetNumSpecies() const;
  int GetNumReactions() const;
  const arma::Mat<int> &GetStatesRef() const;  ///< local block, S x n_local (host mirror, fetched on demand)
  arma::Mat<int> CopyStatesOnProc() const;
  void CopyStatesOnProc(int num_local_states, int *state_array) const;
  std::tuple<int, int> GetOrderingStartEnd() const;

  virtual PacmenslErrorCode Expand() { return 0; }
  virtual PacmenslErrorCode Clear();
  virtual ~StateSetBase();

  // ---- device access (extensions used by FspMatrixBase::GenerateValues) ----
  fspset_t GetDeviceSet() const { return dset_; }
  int GetLocalStart() const { return local_start_; }
  /// ownership start of every rank (+ N at the end): ind_starts_ of the reference
  const std::vector<int> &GetLayout() const { return ind_starts_; }
  /// Extension: distribute the construction like the reference does (src/StateSet/StateSetBase.cpp:134-154,188-258):
  /// every rank keeps only its block of states, explores the frontier states it owns, and the directory is striped
  /// over the GPUs' HBM (include/fsp_b200.h: fspset_set_sharded).  Must be called before the first state is added;
  /// needs > 1 rank with peer memory, otherwise the replicated directory stays.  This is the DEFAULT on > 1 rank;
  /// SetSharded(false) or FSP_SHARDED_SET=0 keep the replicated directory (every rank holds and expands the whole set).
  /// Global indices of a sharded set change when it grows (as in the reference): see RememberLocalStates.
  PacmenslErrorCode SetSharded(bool on = true);
  bool IsSharded() const { return dset_ && fspset_is_sharded(dset_) != 0; }
  /// State2Index(states_old) of FspSolverMultiSinks.cpp:174-205 without the host round trip: remember the local block
  /// before Expand(), ask for the new global indices of those states afterwards.
  PacmenslErrorCode RememberLocalStates();
  PacmenslErrorCode RememberedIndices(std::vector<int> &indices);

 protected:
  MPI_Comm comm_ = MPI_COMM_NULL;
  int      comm_size_ = 1, my_rank_ = 0;
  bool     set_up_ = false;
  int      stoich_set_ = 0;
  int      num_species_ = 0, num_reactions_ = 0;
  int      num_global_states_ = 0, num_local_states_ = 0, local_start_ = 0;
  arma::Mat<int>   stoichiometry_matrix_;
  std::vector<int> ind_starts_;
  PartitioningType lb_type_ = PartitioningType::BLOCK;
  double           lb_threshold_ = 0.2;

  fspset_t               dset_ = nullptr;
  int                    want_sharded_ = -1;  ///< -1: default (FSP_SHARDED_SET, else on), 0 / 1: SetSharded()
  long                   n_remembered_ = 0;
  mutable arma::Mat<int> local_states_;
  mutable bool           host_states_valid_ = false;

  PacmenslErrorCode ensure_device_set();
  PacmenslErrorCode update_layout();
};

}  // namespace pacmensl

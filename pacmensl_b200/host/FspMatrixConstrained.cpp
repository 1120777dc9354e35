#include "FspMatrixConstrained.h"

#include <algorithm>

namespace pacmensl {

FspMatrixConstrained::FspMatrixConstrained(MPI_Comm comm) : FspMatrixBase(comm) {}

FspMatrixConstrained::~FspMatrixConstrained() { Destroy(); }

int FspMatrixConstrained::Destroy() { return FspMatrixBase::Destroy(); }

PacmenslErrorCode FspMatrixConstrained::GenerateValues(const StateSetBase &fsp, const Model &model) {
  mass_action_ = model.mass_action_;
  PacmenslErrorCode ierr =
      FspMatrixConstrained::GenerateValues(fsp, model.stoichiometry_matrix_, model.tv_reactions_, model.prop_t_,
                                           model.prop_x_, std::vector<int>(), model.prop_t_args_, model.prop_x_args_);
  mass_action_.reset();
  return ierr;
}

// src/Matrix/FspMatrixConstrained.cpp:121-282
PacmenslErrorCode FspMatrixConstrained::GenerateValues(const StateSetBase &state_set, const arma::Mat<Int> &SM,
                                                       std::vector<int> time_vayring, const TcoefFun &new_prop_t,
                                                       const PropFun &prop, const std::vector<int> &enable_reactions,
                                                       void *prop_t_args, void *prop_args) {
  PetscErrorCode ierr{0};
  auto *constrained_fss_ptr = dynamic_cast<const StateSetConstrained *>(&state_set);
  if (!constrained_fss_ptr) ierr = -1;  // :133-135
  PACMENSLCHKERRQ(ierr);
  sinks_rank_ = comm_size_ - 1;  // :137
  return FspMatrixBase::GenerateValues(state_set, SM, time_vayring, new_prop_t, prop, enable_reactions, prop_t_args,
                                       prop_args);
}

// src/Matrix/FspMatrixConstrained.cpp:284-302
PacmenslErrorCode FspMatrixConstrained::DetermineLayout_(const StateSetBase &fsp) {
  auto *cfss = dynamic_cast<const StateSetConstrained *>(&fsp);
  if (!cfss) return -1;
  FspMatrixBase::DetermineLayout_(fsp);
  sinks_rank_ = comm_size_ - 1;
  num_constraints_ = cfss->GetNumConstraints();
  owns_sinks_ = (rank_ == sinks_rank_);
  if (owns_sinks_) num_rows_local_ += num_constraints_;
  num_rows_global_ = fsp.GetNumGlobalStates() + num_constraints_;
  return 0;
}

// src/Matrix/FspMatrixConstrained.cpp:170-194: for state i, reaction r, every constraint k violated by
// x_i + nu_r (destinations with a negative coordinate feed no sink) gets the entry (N + k, i) = d_r(x_i).
int FspMatrixConstrained::CollectSinks_(const StateSetBase &fsp, const arma::Mat<Int> &SM,
                                        const std::vector<int> &planes, const double *diag_planes_dev, long ld,
                                        std::vector<long> &sink_ptr, DeviceBuffer<int> &sink_idx,
                                        DeviceBuffer<double> &sink_val) {
  const int  K = num_constraints_, P = (int) planes.size();
  const long n = fsp.GetNumLocalStates();
  sink_ptr.assign((size_t) P * K + 1, 0);
  if (n == 0 || K == 0) return 0;
  fspset_t dset = fsp.GetDeviceSet();
  {
    auto *cfss = dynamic_cast<const StateSetConstrained *>(&fsp);
    int   ierr = cfss ? cfss->SyncShapeToDevice() : -1;
    PACMENSLCHKERRQ(ierr);
  }
  // per plane: K ascending index lists over the boundary states (status != 0): O(surface) work and memory
  long n_bnd = n;
  FSPCHKERRQ(fspset_num_boundary_states(dset, fsp.GetLocalStart(), n, &n_bnd));
  const long cap = std::max<long>(n_bnd * K, 1);
  std::vector<DeviceBuffer<int>>    idx_parts(P);
  std::vector<DeviceBuffer<double>> val_parts(P);
  std::vector<long>                 counts((size_t) K);
  DeviceBuffer<int>                 scratch;
  if (scratch.resize((size_t) cap)) return -1;
  long total = 0;
  for (int p = 0; p < P; ++p) {
    const int r = planes[p];
    FSPCHKERRQ(fspset_sink_lists(dset, SM.colptr(r), fsp.GetLocalStart(), n, scratch.get(), cap, counts.data()));
    long tot_p = 0;
    for (int k = 0; k < K; ++k) {
      sink_ptr[(size_t) p * K + k + 1] = sink_ptr[(size_t) p * K + k] + counts[k];
      tot_p += counts[k];
    }
    if (idx_parts[p].resize((size_t) (tot_p > 0 ? tot_p : 1)) || val_parts[p].resize((size_t) (tot_p > 0 ? tot_p : 1))) return -1;
    if (tot_p > 0) {
      FSPCHKERRQ(fsp_memcpy_d2d(idx_parts[p].get(), scratch.get(), sizeof(int) * tot_p, nullptr));
      // values d_r(x_i): gather from this plane's diagonal (the reference calls prop() per entry, :188)
      FSPCHKERRQ(fspvec_gather(val_parts[p].get(), diag_planes_dev + (size_t) p * ld, idx_parts[p].get(), tot_p, nullptr));
    }
    total += tot_p;
  }
  if (sink_idx.resize((size_t) (total > 0 ? total : 1))) return -1;
  if (sink_val.resize((size_t) (total > 0 ? total : 1))) return -1;
  for (int p = 0; p < P; ++p) {
    long b = sink_ptr[(size_t) p * K], e = sink_ptr[(size_t) (p + 1) * K];
    if (e > b) {
      FSPCHKERRQ(fsp_memcpy_d2d(sink_idx.get() + b, idx_parts[p].get(), sizeof(int) * (e - b), nullptr));
      FSPCHKERRQ(fsp_memcpy_d2d(sink_val.get() + b, val_parts[p].get(), sizeof(double) * (e - b), nullptr));
    }
  }
  FSPCHKERRQ(fsp_device_sync());
  return 0;
}

}  // namespace pacmensl

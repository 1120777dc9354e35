#include "DiscreteDistribution.h"

#include <cstdlib>

namespace pacmensl {

DiscreteDistribution::~DiscreteDistribution() {
  if (p_ != nullptr) VecDestroy(&p_);
  p_ = nullptr;
  comm_ = MPI_COMM_NULL;
}
DiscreteDistribution::DiscreteDistribution() {}
DiscreteDistribution::DiscreteDistribution(const DiscreteDistribution &dist) { *this = dist; }
DiscreteDistribution::DiscreteDistribution(DiscreteDistribution &&dist) noexcept { *this = std::move(dist); }

DiscreteDistribution &DiscreteDistribution::operator=(const DiscreteDistribution &dist) {
  if (this == &dist) return *this;
  if (p_ != nullptr) VecDestroy(&p_);
  comm_ = dist.comm_;
  t_ = dist.t_;
  if (dist.p_) {
    VecDuplicate(dist.p_, &p_);
    VecCopy(dist.p_, p_);
  }
  states_ = dist.states_;
  states_dev_ = dist.states_dev_;
  return *this;
}
DiscreteDistribution &DiscreteDistribution::operator=(DiscreteDistribution &&dist) noexcept {
  if (this != &dist) {
    if (p_ != nullptr) VecDestroy(&p_);
    comm_ = dist.comm_;
    t_ = dist.t_;
    states_ = std::move(dist.states_);
    states_dev_ = std::move(dist.states_dev_);
    p_ = dist.p_;
    dist.comm_ = MPI_COMM_NULL;
    dist.p_ = nullptr;
    dist.states_.clear();
  }
  return *this;
}
DiscreteDistribution::DiscreteDistribution(MPI_Comm comm, double t, const StateSetBase *state_set, const Vec &p) {
  comm_ = comm;
  t_ = t;
  states_ = state_set->CopyStatesOnProc();
  VecDuplicate(p, &p_);
  VecCopy(p, p_);
  AttachDeviceStates(state_set);
}
// keep the local states on the device as well (device-to-device copy out of the state set's array)
void DiscreteDistribution::AttachDeviceStates(const StateSetBase *state_set) {
  states_dev_.reset();
  const int *all = nullptr;
  const long n = state_set->GetNumLocalStates(), S = state_set->GetNumSpecies(), first = state_set->GetLocalStart();
  fspset_t   dset = state_set->GetDeviceSet();
  void      *stream = comm_ ? comm_->stream : nullptr;
  if (dset && n > 0 && fspset_states_dev(dset, &all) == 0 && all) {
    auto buf = std::make_shared<DeviceBuffer<int>>();
    if (buf->resize((size_t) n * S) == 0 &&
        fsp_memcpy_d2d(buf->get(), all + (size_t) first * S, sizeof(int) * (size_t) n * S, stream) == 0 && fsp_stream_sync(stream) == 0)
      states_dev_ = buf;
  }
}
int DiscreteDistribution::GetStateView(int &num_states, int &num_species, int *&states) {
  num_states = (int) states_.n_cols;
  num_species = (int) states_.n_rows;
  states = states_.memptr();
  return 0;
}
int DiscreteDistribution::GetProbView(int &num_states, double *&p) {
  int ierr = VecGetLocalSize(p_, &num_states); CHKERRQ(ierr);
  ierr = VecGetArray(p_, &p); CHKERRQ(ierr);
  return 0;
}
int DiscreteDistribution::RestoreProbView(double *&p) {
  if (p != nullptr) { int ierr = VecRestoreArray(p_, &p); CHKERRQ(ierr); }
  return 0;
}
PacmenslErrorCode DiscreteDistribution::WeightedAverage(
    int nout, PetscReal *fout,
    std::function<PacmenslErrorCode(int num_species, int *x, int nout, PetscReal *fx, void *args)> weight_func,
    void *wf_args) {
  int        num_local_states;
  PetscReal *plocal;
  PacmenslErrorCode ierr = GetProbView(num_local_states, plocal); PACMENSLCHKERRQ(ierr);
  for (int i = 0; i < nout; ++i) fout[i] = 0.0;
  std::vector<PetscReal> wtmp((size_t) nout);
  for (int j = 0; j < num_local_states; ++j) {
    ierr = weight_func((int) states_.n_rows, states_.colptr(j), nout, wtmp.data(), wf_args); PACMENSLCHKERRQ(ierr);
    for (int i = 0; i < nout; ++i) fout[i] += wtmp[i] * plocal[j];
  }
  RestoreProbView(plocal);
  return pacmensl_allreduce_sum(comm_, fout, nout);
}

// src/Fsp/DiscreteDistribution.cpp:171-200.  On the device when the distribution carries its states there: a
// deterministic segmented reduction keyed by the species' coordinate (fspvec_marginal), K-independent; only the
// max + 1 bins come back to the host.  The host loop remains for distributions built without a device state array.
PacmenslErrorCode ComputeMarginalOf(const DiscreteDistribution &dist, Vec v, int species, arma::Col<PetscReal> &out) {
  const long n = (long) dist.states_.n_cols;
  const int  S = (int) dist.states_.n_rows;
  void      *stream = dist.comm_ ? dist.comm_->stream : nullptr;
  if (species < 0 || (S > 0 && species >= S)) return -1;
  static const bool host_only = [] { const char *e = std::getenv("FSP_HOST_MARGINAL"); return e && e[0] == '1'; }();
  const bool        on_device = !host_only && (n == 0 || (dist.states_dev_ && dist.states_dev_->get()));
  double            mx = 0.0;
  if (on_device && n > 0) {
    DeviceBuffer<double> tmp(1);
    FSPCHKERRQ(fspvec_max_species(tmp.get(), dist.states_dev_->get(), S, species, n, stream));
    FSPCHKERRQ(fsp_memcpy_d2h(&mx, tmp.get(), sizeof(double), stream));
    mx = -mx;
  } else {
    for (arma::uword i = 0; i < dist.states_.n_cols; ++i) mx = std::max(mx, (double) dist.states_(species, i));
  }
  pacmensl_allreduce_max(dist.comm_, &mx, 1);
  const int M = (int) mx + 1;
  out = arma::Col<PetscReal>((arma::uword) M, arma::fill::zeros);
  if (on_device && (size_t) M * 8 * sizeof(double) <= 96 * 1024) {
    if (n > 0) {
      DeviceBuffer<double> bins((size_t) M);
      FSPCHKERRQ(fspvec_marginal(bins.get(), M, v->d_data, dist.states_dev_->get(), S, species, n, stream));
      FSPCHKERRQ(fsp_memcpy_d2h(out.memptr(), bins.get(), sizeof(double) * M, stream));
    }
  } else {
    const PetscReal *p_dat;
    VecGetArrayRead(v, &p_dat);
    for (arma::uword i{0}; i < dist.states_.n_cols; ++i) out(dist.states_(species, i)) += p_dat[i];
    VecRestoreArrayRead(v, &p_dat);
  }
  return pacmensl_allreduce_sum(dist.comm_, out.memptr(), (int) out.n_elem);
}

arma::Col<PetscReal> Compute1DMarginal(const DiscreteDistribution &dist, int species) {
  arma::Col<PetscReal> md;
  if (ComputeMarginalOf(dist, dist.p_, species, md)) md.reset();
  return md;
}
}  // namespace pacmensl

#include "SensDiscreteDistribution.h"

namespace pacmensl {

SensDiscreteDistribution::SensDiscreteDistribution() : DiscreteDistribution() {}
SensDiscreteDistribution::SensDiscreteDistribution(MPI_Comm comm, double t, const StateSetBase *state_set, const Vec &p,
                                                   const std::vector<Vec> &dp)
    : DiscreteDistribution(comm, t, state_set, p) {
  dp_.resize(dp.size());
  for (size_t i{0}; i < dp_.size(); ++i) {
    VecDuplicate(dp[i], &dp_[i]);
    VecCopy(dp[i], dp_[i]);
  }
}
SensDiscreteDistribution::SensDiscreteDistribution(const SensDiscreteDistribution &dist) : DiscreteDistribution() { *this = dist; }
SensDiscreteDistribution::SensDiscreteDistribution(SensDiscreteDistribution &&dist) noexcept : DiscreteDistribution() {
  *this = std::move(dist);
}
SensDiscreteDistribution &SensDiscreteDistribution::operator=(const SensDiscreteDistribution &dist) {
  if (this == &dist) return *this;
  DiscreteDistribution::operator=(dist);
  for (auto &v : dp_) VecDestroy(&v);
  dp_.resize(dist.dp_.size());
  for (size_t i{0}; i < dp_.size(); ++i) {
    VecDuplicate(dist.dp_[i], &dp_[i]);
    VecCopy(dist.dp_[i], dp_[i]);
  }
  return *this;
}
SensDiscreteDistribution &SensDiscreteDistribution::operator=(SensDiscreteDistribution &&dist) noexcept {
  if (this != &dist) {
    DiscreteDistribution::operator=(std::move(dist));
    for (auto &v : dp_) VecDestroy(&v);
    dp_ = std::move(dist.dp_);
    dist.dp_.clear();
  }
  return *this;
}
SensDiscreteDistribution::~SensDiscreteDistribution() {
  for (auto &v : dp_) VecDestroy(&v);
  dp_.clear();
}
PacmenslErrorCode SensDiscreteDistribution::GetSensView(int is, int &num_states, double *&p) {
  if (is < 0 || is >= (int) dp_.size()) return -1;
  int ierr = VecGetLocalSize(dp_[is], &num_states); CHKERRQ(ierr);
  ierr = VecGetArray(dp_[is], &p); CHKERRQ(ierr);
  return 0;
}
PacmenslErrorCode SensDiscreteDistribution::RestoreSensView(int is, double *&p) {
  if (is < 0 || is >= (int) dp_.size()) return -1;
  if (p != nullptr) { int ierr = VecRestoreArray(dp_[is], &p); CHKERRQ(ierr); }
  return 0;
}
PacmenslErrorCode SensDiscreteDistribution::WeightedAverage(
    int is, int nout, PetscReal *fout,
    std::function<PacmenslErrorCode(int num_species, int *x, int nout, PetscReal *wx, void *args)> weight_func, void *wf_args) {
  int        n;
  PetscReal *plocal;
  PacmenslErrorCode ierr = is < 0 ? GetProbView(n, plocal) : GetSensView(is, n, plocal);
  PACMENSLCHKERRQ(ierr);
  for (int i = 0; i < nout; ++i) fout[i] = 0.0;
  std::vector<PetscReal> wtmp((size_t) nout);
  for (int j = 0; j < n; ++j) {
    ierr = weight_func((int) states_.n_rows, states_.colptr(j), nout, wtmp.data(), wf_args); PACMENSLCHKERRQ(ierr);
    for (int i = 0; i < nout; ++i) fout[i] += wtmp[i] * plocal[j];
  }
  if (is < 0) RestoreProbView(plocal); else RestoreSensView(is, plocal);
  return pacmensl_allreduce_sum(comm_, fout, nout);
}

PacmenslErrorCode Compute1DSensMarginal(const SensDiscreteDistribution &dist, int is, int species, arma::Col<PetscReal> &out) {
  if (is < 0 || is >= (int) dist.dp_.size()) return -1;
  return ComputeMarginalOf(dist, dist.dp_[is], species, out);  // device segmented reduction (DiscreteDistribution.cpp)
}

// src/SensFsp/SensDiscreteDistribution.cpp:216-271: FIM(i, j) = sum_x s_i(x) s_j(x) / p(x), with the reference's in-place
// floor p(x) := max(p(x), 1e-16) (it writes the floored values back through its array view) and its warning.  Here each
// entry is one fused device reduction over (s_i, s_j, p) with a fixed summation shape; nothing but P(P+1)/2 scalars
// crosses to the host.
PacmenslErrorCode ComputeFIM(SensDiscreteDistribution &dist, arma::Mat<PetscReal> &fim) {
  const int P = (int) dist.dp_.size();
  if (!fim.is_empty() && ((int) fim.n_rows != P || (int) fim.n_cols != P)) return -1;
  if (fim.is_empty()) fim.set_size(P, P);
  if (P == 0) return 0;
  void      *stream = dist.comm_ ? dist.comm_->stream : nullptr;
  const long n = dist.p_->n_local;
  DeviceBuffer<double> tmp((size_t) P * P + 1);
  FSPCHKERRQ(fspvec_clamp_min(tmp.get() + (size_t) P * P, dist.p_->d_data, 1.0e-16, n, stream));
  for (int i = 0; i < P; ++i)
    for (int j = 0; j <= i; ++j)
      FSPCHKERRQ(fspvec_wdiv_dot(tmp.get() + (size_t) j * P + i, dist.dp_[i]->d_data, dist.dp_[j]->d_data, dist.p_->d_data, n, stream));
  std::vector<double> host((size_t) P * P + 1, 0.0);
  FSPCHKERRQ(fsp_memcpy_d2h(host.data(), tmp.get(), sizeof(double) * host.size(), stream));
  std::vector<double> low;
  for (int i = 0; i < P; ++i) for (int j = 0; j <= i; ++j) low.push_back(n > 0 ? host[(size_t) j * P + i] : 0.0);
  low.push_back(n > 0 ? host[(size_t) P * P] : 0.0);
  int ierr = pacmensl_allreduce_sum(dist.comm_, low.data(), (int) low.size());
  PACMENSLCHKERRQ(ierr);
  size_t q = 0;
  for (int i = 0; i < P; ++i) for (int j = 0; j <= i; ++j) { fim(i, j) = low[q]; fim(j, i) = low[q]; ++q; }
  if (low[q] > 0.0) PetscPrintf(dist.comm_, "Warning: rounding was done in FIM computation.\n");
  return 0;
}
}  // namespace pacmensl

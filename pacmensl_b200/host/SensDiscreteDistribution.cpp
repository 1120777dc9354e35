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
  double mx = 0.0;
  for (arma::uword i = 0; i < dist.states_.n_cols; ++i) mx = std::max(mx, (double) dist.states_(species, i));
  pacmensl_allreduce_max(dist.comm_, &mx, 1);
  out = arma::Col<PetscReal>((arma::uword) mx + 1, arma::fill::zeros);
  const PetscReal *p_dat;
  VecGetArrayRead(dist.dp_[is], &p_dat);
  for (arma::uword i{0}; i < dist.states_.n_cols; ++i) out(dist.states_(species, i)) += p_dat[i];
  VecRestoreArrayRead(dist.dp_[is], &p_dat);
  return pacmensl_allreduce_sum(dist.comm_, out.memptr(), (int) out.n_elem);
}
}  // namespace pacmensl

#include "DiscreteDistribution.h"

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
  return *this;
}
DiscreteDistribution &DiscreteDistribution::operator=(DiscreteDistribution &&dist) noexcept {
  if (this != &dist) {
    if (p_ != nullptr) VecDestroy(&p_);
    comm_ = dist.comm_;
    t_ = dist.t_;
    states_ = std::move(dist.states_);
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

arma::Col<PetscReal> Compute1DMarginal(const DiscreteDistribution &dist, int species) {
  double mx = 0.0;
  for (arma::uword i = 0; i < dist.states_.n_cols; ++i) mx = std::max(mx, (double) dist.states_(species, i));
  pacmensl_allreduce_max(dist.comm_, &mx, 1);
  arma::Col<PetscReal> md((arma::uword) mx + 1, arma::fill::zeros);
  const PetscReal *p_dat;
  VecGetArrayRead(dist.p_, &p_dat);
  for (arma::uword i{0}; i < dist.states_.n_cols; ++i) md(dist.states_(species, i)) += p_dat[i];
  VecRestoreArrayRead(dist.p_, &p_dat);
  pacmensl_allreduce_sum(dist.comm_, md.memptr(), (int) md.n_elem);
  return md;
}
}  // namespace pacmensl

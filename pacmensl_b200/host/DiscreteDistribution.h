// DiscreteDistribution.h -- result container: states + probabilities without the sinks
// (mirrors src/Fsp/DiscreteDistribution.h:38-107).
#pragma once

#include "StateSetBase.h"
#include "Sys.h"

namespace pacmensl {

struct PACMENSL_API DiscreteDistribution {
  MPI_Comm       comm_ = MPI_COMM_NULL;
  double         t_ = 0.0;
  arma::Mat<int> states_;
  Vec            p_ = nullptr;
  /// device copy of states_ (taken from the state set's device array at construction, shared between copies): lets the
  /// post-processing reductions (Compute1DMarginal, Compute1DSensMarginal) run on the GPU next to p_
  std::shared_ptr<DeviceBuffer<int>> states_dev_;

  DiscreteDistribution();
  DiscreteDistribution(MPI_Comm comm, double t, const StateSetBase *state_set, const Vec &p);
  DiscreteDistribution(const DiscreteDistribution &dist);
  DiscreteDistribution(DiscreteDistribution &&dist) noexcept;
  DiscreteDistribution &operator=(const DiscreteDistribution &);
  DiscreteDistribution &operator=(DiscreteDistribution &&) noexcept;

  void AttachDeviceStates(const StateSetBase *state_set);
  PacmenslErrorCode GetStateView(int &num_states, int &num_species, int *&states);
  PacmenslErrorCode GetProbView(int &num_states, double *&p);
  PacmenslErrorCode RestoreProbView(double *&p);
  PacmenslErrorCode WeightedAverage(int nout, PetscReal *fout,
                                    std::function<PacmenslErrorCode(int num_species, int *x, int nout, PetscReal *wx, void *args)> weight_func,
                                    void *wf_args);
  ~DiscreteDistribution();
};

PACMENSL_API arma::Col<PetscReal> Compute1DMarginal(const DiscreteDistribution &dist, int species);
/// shared implementation: marginal of an arbitrary device vector v (p_ or a sensitivity) over dist's states
PACMENSL_API PacmenslErrorCode ComputeMarginalOf(const DiscreteDistribution &dist, Vec v, int species, arma::Col<PetscReal> &out);
}  // namespace pacmensl

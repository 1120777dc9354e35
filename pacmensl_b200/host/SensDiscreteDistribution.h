// SensDiscreteDistribution.h -- probability distribution together with its parameter sensitivities
// (mirrors src/SensFsp/SensDiscreteDistribution.h:31-60; ComputeFIM is post-processing, out of scope).
#pragma once

#include "DiscreteDistribution.h"

namespace pacmensl {
class PACMENSL_API SensDiscreteDistribution : public DiscreteDistribution {
 public:
  std::vector<Vec> dp_;

  SensDiscreteDistribution();
  SensDiscreteDistribution(MPI_Comm comm, double t, const StateSetBase *state_set, const Vec &p, const std::vector<Vec> &dp);
  SensDiscreteDistribution(const SensDiscreteDistribution &dist);
  SensDiscreteDistribution(SensDiscreteDistribution &&dist) noexcept;
  SensDiscreteDistribution &operator=(const SensDiscreteDistribution &dist);
  SensDiscreteDistribution &operator=(SensDiscreteDistribution &&dist) noexcept;

  PacmenslErrorCode GetSensView(int is, int &num_states, double *&p);
  PacmenslErrorCode RestoreSensView(int is, double *&p);
  PacmenslErrorCode WeightedAverage(int is, int nout, PetscReal *fout,
                                    std::function<PacmenslErrorCode(int num_species, int *x, int nout, PetscReal *wx, void *args)> weight_func,
                                    void *wf_args);
  ~SensDiscreteDistribution();
};

PACMENSL_API PacmenslErrorCode Compute1DSensMarginal(const SensDiscreteDistribution &dist, int is, int species,
                                                     arma::Col<PetscReal> &out);
/// Fisher information matrix of the distribution w.r.t. its parameters (src/SensFsp/SensDiscreteDistribution.cpp:216-271),
/// computed on the device; like the reference it floors p at 1e-16 IN PLACE and warns when it had to.
PACMENSL_API PacmenslErrorCode ComputeFIM(SensDiscreteDistribution &dist, arma::Mat<PetscReal> &fim);
}  // namespace pacmensl

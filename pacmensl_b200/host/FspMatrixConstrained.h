// FspMatrixConstrained.h -- FSP operator with one absorbing sink row per shape constraint.
// Mirrors src/Matrix/FspMatrixConstrained.h:35-80.  The K sink rows are the last K entries of the
// global vector and live on the last rank (src/Matrix/FspMatrixConstrained.cpp:137,284-289).
#pragma once

#include "FspMatrixBase.h"

namespace pacmensl {

class PACMENSL_API FspMatrixConstrained : public FspMatrixBase {
 public:
  explicit FspMatrixConstrained(MPI_Comm comm);

  PacmenslErrorCode GenerateValues(const StateSetBase &fsp, const Model &model) override;

  PacmenslErrorCode GenerateValues(const StateSetBase &state_set, const arma::Mat<Int> &SM,
                                   std::vector<int> time_vayring, const TcoefFun &new_prop_t, const PropFun &prop,
                                   const std::vector<int> &enable_reactions, void *prop_t_args,
                                   void *prop_args) override;

  int Destroy() override;

  ~FspMatrixConstrained() override;

 protected:
  int sinks_rank_ = 0;  ///< rank that stores the sink states

  PacmenslErrorCode DetermineLayout_(const StateSetBase &fsp) override;
  int CollectSinks_(const StateSetBase &fsp, const arma::Mat<Int> &SM, const std::vector<int> &planes,
                    const double *diag_planes_dev, long ld, std::vector<long> &sink_ptr, DeviceBuffer<int> &sink_idx,
                    DeviceBuffer<double> &sink_val) override;
};
}  // namespace pacmensl

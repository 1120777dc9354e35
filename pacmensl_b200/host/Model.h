// Model.h -- description of a stochastic reaction network (mirrors src/Models/Model.h:44-99).
// Propensities factor as a_r(t, x) = c_r(t) * d_r(x): prop_t_ fills the c_r, prop_x_ evaluates d_r.
// Extension: an optional mass-action description of d_r lets GenerateValues evaluate propensities on
// the device instead of through the host callback (the callback, when given, stays the API contract).
#pragma once

#include <functional>
#include <memory>
#include <vector>

#include "arma_shim.h"
#include "petsc_shim.h"

namespace pacmensl {

using PropFun = std::function<int(const int reaction, const int num_species, const int num_states, const int *states,
                                  double *outputs, void *args)>;
using TcoefFun = std::function<int(double t, int num_coefs, double *outputs, void *args)>;

// d_r(x) = rate[r] * prod_s ff(x_s, order(s, r)), ff(x,0)=1, ff(x,1)=x, ff(x,2)=x(x-1)/2, ff(x,3)=x(x-1)(x-2)/6
struct MassActionPropensity {
  std::vector<double> rate;   // R
  arma::Mat<int>      order;  // S x R
};

class PACMENSL_API Model {
 public:
  arma::Mat<int>   stoichiometry_matrix_;
  TcoefFun         prop_t_;
  void            *prop_t_args_;
  PropFun          prop_x_;
  void            *prop_x_args_;
  std::vector<int> tv_reactions_;
  std::shared_ptr<MassActionPropensity> mass_action_;  ///< optional device-evaluable form of prop_x_

  Model();
  explicit Model(arma::Mat<int> stoichiometry_matrix, TcoefFun prop_t, PropFun prop_x, void *prop_t_args = nullptr,
                 void *prop_x_args = nullptr, const std::vector<int> &tv_reactions_ = std::vector<int>());
  Model(const Model &model);
  Model &operator=(const Model &model) noexcept;
  Model &operator=(Model &&model) noexcept;

  /// Provide the mass-action form; also installs an equivalent host prop_x_ if none was given.
  void SetMassAction(const std::vector<double> &rates, const arma::Mat<int> &orders);
};

}  // namespace pacmensl

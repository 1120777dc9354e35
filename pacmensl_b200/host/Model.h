// Model.h -- description of a stochastic reaction network (mirrors src/Models/Model.h:44-99).
// Propensities factor as a_r(t, x) = c_r(t) * d_r(x): prop_t_ fills the c_r, prop_x_ evaluates d_r.
// Extension: an optional mass-action description of d_r lets GenerateValues evaluate propensities on
// the device instead of through the host callback (the callback, when given, stays the API contract).
#pragma once
#include <algorithm>
#include <map>

#include <functional>
#include <memory>
#include <vector>

#include "arma_shim.h"
#include "petsc_shim.h"

namespace pacmensl {

using PropFun = std::function<int(const int reaction, const int num_species, const int num_states, const int *states,
                                  double *outputs, void *args)>;
using TcoefFun = std::function<int(double t, int num_coefs, double *outputs, void *args)>;

// d_r(x) = rate[r] * prod_s ff(x_s, order(s, r)) * T_{s,r}[min(x_s, len - 1)]
//   ff(x,0)=1, ff(x,1)=x, ff(x,2)=x(x-1)/2, ff(x,3)=x(x-1)(x-2)/6;  T_{s,r}: optional factor table (absent = 1; 0 for x < 0)
struct MassActionPropensity {
  std::vector<double> rate;   // R
  arma::Mat<int>      order;  // S x R
  std::map<std::pair<int, int>, std::vector<double>> table;  // (species, reaction) -> factor values for x_s = 0, 1, ...
  double eval(int r, int S, const int *x) const {
    double v = rate[r];
    for (int s = 0; s < S; ++s) {
      const int xs = x[s], o = order(s, r);
      if (o == 1) v *= (double) xs;
      else if (o == 2) v *= 0.5 * (double) xs * (double) (xs - 1);
      else if (o == 3) v *= (double) xs * (double) (xs - 1) * (double) (xs - 2) / 6.0;
      auto it = table.find({s, r});
      if (it != table.end() && !it->second.empty())
        v *= xs < 0 ? 0.0 : it->second[(size_t) std::min<long>(xs, (long) it->second.size() - 1)];
    }
    return v;
  }
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
  /// Adds a per-species factor table to reaction r of the device-evaluable form (after SetMassAction): d_r gets the
  /// extra factor values[min(x_species, values.size() - 1)].  Gene-state switches, Hill factors of one species, ...
  void SetFactorTable(int species, int reaction, const std::vector<double> &values);
};

}  // namespace pacmensl

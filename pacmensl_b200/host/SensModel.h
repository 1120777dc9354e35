// SensModel.h -- reaction network with parameter derivatives of the propensities
// (mirrors src/Models/SensModel.h:38-97).
#pragma once

#include "Model.h"

namespace pacmensl {

using DTcoefFun = std::function<int(const int parameter_idx, const double t, int num_coefs, double *outputs, void *args)>;
using DPropFun = std::function<int(const int parameter_idx, const int reaction_idx, const int num_species,
                                   const int num_states, const int *states, double *outputs, void *args)>;

class PACMENSL_API SensModel {
 public:
  int              num_reactions_ = 0;
  int              num_parameters_ = 0;
  arma::Mat<int>   stoichiometry_matrix_;
  std::vector<int> tv_reactions_;

  PropFun  prop_x_;
  void    *prop_x_args_ = nullptr;
  TcoefFun prop_t_;
  void    *prop_t_args_ = nullptr;

  DTcoefFun                     dprop_t_;
  void                         *dprop_t_args_ = nullptr;
  std::vector<std::vector<int>> dprop_t_sp_;  ///< per parameter: reactions whose c_r depends on it
  DPropFun                      dprop_x_;
  void                         *dprop_x_args_ = nullptr;
  std::vector<std::vector<int>> dprop_x_sp_;  ///< per parameter: reactions whose d_r depends on it

  SensModel() {}
  explicit SensModel(const int num_parameters, const arma::Mat<int> &stoichiometry_matrix,
                     const std::vector<int> &tv_reactions, const TcoefFun &prop_t, const PropFun &prop_x,
                     const DTcoefFun &dprop_t, const std::vector<std::vector<int>> &dprop_t_sp, const DPropFun &dprop_x,
                     const std::vector<std::vector<int>> &dprop_x_sp = std::vector<std::vector<int>>(),
                     void *prop_t_args = nullptr, void *prop_x_args = nullptr, void *dprop_t_args = nullptr,
                     void *dprop_x_args = nullptr) {
    num_parameters_ = num_parameters;
    num_reactions_ = (int) stoichiometry_matrix.n_cols;
    stoichiometry_matrix_ = stoichiometry_matrix;
    prop_t_ = prop_t;
    prop_x_ = prop_x;
    dprop_t_ = dprop_t;
    dprop_t_sp_ = dprop_t_sp;
    dprop_x_ = dprop_x;
    dprop_x_sp_ = dprop_x_sp;
    prop_t_args_ = prop_t_args;
    prop_x_args_ = prop_x_args;
    dprop_t_args_ = dprop_t_args;
    dprop_x_args_ = dprop_x_args;
    tv_reactions_ = tv_reactions;
  }
};
}  // namespace pacmensl

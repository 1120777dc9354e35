#include "Model.h"

namespace pacmensl {

Model::Model() {
  prop_t_args_ = nullptr;
  prop_t_ = nullptr;
  prop_x_args_ = nullptr;
  prop_x_ = nullptr;
}

Model::Model(arma::Mat<int> stoichiometry_matrix, TcoefFun prop_t, PropFun prop_x, void *prop_t_args,
             void *prop_x_args, const std::vector<int> &tv_reactions) {
  stoichiometry_matrix_ = std::move(stoichiometry_matrix);
  prop_t_ = std::move(prop_t);
  prop_t_args_ = prop_t_args;
  prop_x_ = std::move(prop_x);
  prop_x_args_ = prop_x_args;
  tv_reactions_ = tv_reactions;
}

Model::Model(const Model &m) { *this = m; }

Model &Model::operator=(const Model &m) noexcept {
  stoichiometry_matrix_ = m.stoichiometry_matrix_;
  prop_t_ = m.prop_t_;
  prop_t_args_ = m.prop_t_args_;
  prop_x_ = m.prop_x_;
  prop_x_args_ = m.prop_x_args_;
  tv_reactions_ = m.tv_reactions_;
  mass_action_ = m.mass_action_;
  return *this;
}

Model &Model::operator=(Model &&m) noexcept {
  if (this == &m) return *this;
  stoichiometry_matrix_ = std::move(m.stoichiometry_matrix_);
  prop_t_ = std::move(m.prop_t_);
  prop_t_args_ = m.prop_t_args_;
  prop_x_ = std::move(m.prop_x_);
  prop_x_args_ = m.prop_x_args_;
  tv_reactions_ = std::move(m.tv_reactions_);
  mass_action_ = std::move(m.mass_action_);
  return *this;
}

void Model::SetMassAction(const std::vector<double> &rates, const arma::Mat<int> &orders) {
  mass_action_ = std::make_shared<MassActionPropensity>();
  mass_action_->rate = rates;
  mass_action_->order = orders;
  if (!prop_x_) {
    auto ma = mass_action_;
    prop_x_ = [ma](const int r, const int S, const int m, const int *X, double *out, void *) {
      if (r < 0 || r >= (int) ma->rate.size()) return -1;
      for (int i = 0; i < m; ++i) out[i] = ma->eval(r, S, X + (size_t) i * S);
      return 0;
    };
  }
}

void Model::SetFactorTable(int species, int reaction, const std::vector<double> &values) {
  if (!mass_action_) return;
  mass_action_->table[{species, reaction}] = values;
}

}  // namespace pacmensl

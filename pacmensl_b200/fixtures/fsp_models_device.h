// fsp_models_device.h -- device-evaluable (separable) descriptions of the fixture propensities of fsp_models.h:
//   d_r(x) = rate_r * prod_s ff(x_s, order(s, r)) * T_{s,r}[x_s]      (pacmensl::Model::SetMassAction / SetFactorTable)
// so that matrix generation evaluates them on the GPU instead of through the host callbacks.  The values are the same
// numbers as the callbacks produce (same constants, same operation order): tests/test_gpu_mat.py checks bit-identity.
#pragma once
#include <cstring>
#include <vector>

#include "Model.h"
#include "fsp_models.h"

namespace pacmensl {

/// returns true if `name` has a separable form and it was attached to `model`
inline bool AttachDeviceForm(const char *name, Model &model) {
  const int S = (int) model.stoichiometry_matrix_.n_rows, R = (int) model.stoichiometry_matrix_.n_cols;
  arma::Mat<int> ord(S, R);
  ord.zeros();
  if (!std::strcmp(name, "hog1p") && S == 5 && R == 9) {
    // examples/hog1p.cpp:14-112: reactions 0-4 switch on the gene state x0 in {0,1,2,3}; 5-8 are first order
    const double k12 = 1.29, k23 = 0.0067, k34 = 0.133, k32 = 0.027, k43 = 0.0381, k21 = 1.0, kr21 = 0.005, kr31 = 0.45,
                 kr41 = 0.025, kr22 = 0.0116, kr32 = 0.987, kr42 = 0.0538, trans = 0.01, gamma1 = 0.001, gamma2 = 0.0049;
    ord(1, 5) = 1; ord(2, 6) = 1; ord(3, 7) = 1; ord(4, 8) = 1;
    model.SetMassAction({1.0, 1.0, 1.0, 1.0, 1.0, trans, trans, gamma1, gamma2}, ord);
    model.SetFactorTable(0, 0, {k12, k23, k34, 0.0, 0.0});
    model.SetFactorTable(0, 1, {0.0, 0.0, k32, k43, 0.0});
    model.SetFactorTable(0, 2, {0.0, k21, 0.0, 0.0, 0.0});
    model.SetFactorTable(0, 3, {0.0, kr21, kr31, kr41, 0.0});
    model.SetFactorTable(0, 4, {0.0, kr22, kr32, kr42, 0.0});
    return true;
  }
  if (!std::strcmp(name, "transcr_reg_6d") && S == 6 && R == 10) {
    const double c0 = 0.043, c1 = 0.0007, c2 = 0.078, c3 = 0.0039, c5 = 0.4791, c7 = 0.8765e-11, c9 = 0.5;
    ord(5, 0) = 1; ord(0, 1) = 1; ord(3, 2) = 1; ord(5, 3) = 1; ord(1, 4) = 1; ord(2, 4) = 1; ord(3, 5) = 1;
    ord(3, 6) = 1; ord(1, 6) = 1; ord(4, 7) = 1; ord(0, 8) = 2; ord(1, 9) = 1;
    model.SetMassAction({c0, c1, c2, c3, 1.0, c5, 1.0, c7, 1.0, c9}, ord);
    return true;
  }
  if ((!std::strcmp(name, "birth_death_3d") || !std::strcmp(name, "birth_death_3d_tv")) && S == 3 && R == 6) {
    ord(0, 1) = 1; ord(1, 3) = 1; ord(2, 5) = 1;
    model.SetMassAction({40.0, 1.0, 30.0, 1.5, 20.0, 2.0}, ord);
    return true;
  }
  if (!std::strcmp(name, "pure_birth") && S == 1 && R == 1) {
    model.SetMassAction({2.0}, ord);
    return true;
  }
  return false;
}

}  // namespace pacmensl

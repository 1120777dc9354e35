// StateSetConstrained.h -- FSP state space shaped by inequality constraints lhs_k(x) <= b_k.
// Mirrors src/StateSet/StateSetConstrained.h:35-68; Expand() is the BFS closure of
// src/StateSet/StateSetConstrained.cpp:132-221 executed by device kernels (fspset_expand).
#pragma once

#include "StateSetBase.h"

namespace pacmensl {
typedef std::function<int(int, int, int, int *, int *, void *)> fsp_constr_multi_fn;

class PACMENSL_API StateSetConstrained : public StateSetBase {
 public:
  explicit StateSetConstrained(MPI_Comm new_comm = MPI_COMM_WORLD);

  int CheckConstraints(PetscInt num_states, PetscInt *x, PetscInt *satisfied) const;
  arma::Row<int> GetShapeBounds() const;
  int GetNumConstraints() const;

  PacmenslErrorCode SetShape(const fsp_constr_multi_fn &lhs_fun, arma::Row<int> &rhs_bounds, void *args = nullptr);
  PacmenslErrorCode SetShape(int num_constraints, const fsp_constr_multi_fn &lhs_fun, int *bounds, void *args = nullptr);
  PacmenslErrorCode SetShapeBounds(arma::Row<PetscInt> &rhs_bounds);
  PacmenslErrorCode SetShapeBounds(int num_constraints, int *bounds);

  PacmenslErrorCode SetUp() override;
  PacmenslErrorCode Expand() override;

  bool HasCustomConstraints() const { return lhs_constr != nullptr && !using_default_; }
  /// make the device directory use the current shape (lhs + bounds); called before sink generation
  PacmenslErrorCode SyncShapeToDevice() const { return const_cast<StateSetConstrained *>(this)->push_shape_to_device(); }

 protected:
  fsp_constr_multi_fn lhs_constr = nullptr;
  arma::Row<int>      rhs_constr;
  void               *args_constr = nullptr;
  bool                using_default_ = false;

  PacmenslErrorCode push_shape_to_device();
  static int lhs_trampoline(int S, int K, int m, int *states, int *out, void *self);
  static int default_constr_fun(int num_species, int num_constr, int n_states, int *states, int *outputs, void *args);
};
}  // namespace pacmensl

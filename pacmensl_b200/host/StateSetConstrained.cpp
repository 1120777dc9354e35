#include "StateSetConstrained.h"

namespace pacmensl {

StateSetConstrained::StateSetConstrained(MPI_Comm new_comm) : StateSetBase(new_comm) {}

int StateSetConstrained::lhs_trampoline(int S, int K, int m, int *states, int *out, void *self) {
  auto *me = static_cast<StateSetConstrained *>(self);
  return me->lhs_constr(S, K, m, states, out, me->args_constr);
}

// src/StateSet/StateSetConstrained.cpp:92-99
int StateSetConstrained::default_constr_fun(int num_species, int, int n_states, int *states, int *outputs, void *) {
  for (int i = 0; i < n_states * num_species; ++i) outputs[i] = states[i];
  return 0;
}

// src/StateSet/StateSetConstrained.cpp:63-82 (satisfied is constraint-major; negative states satisfy)
int StateSetConstrained::CheckConstraints(PetscInt num_states, PetscInt *x, PetscInt *satisfied) const {
  const int        K = (int) rhs_constr.n_elem;
  std::vector<int> fval((size_t) num_states * (K > 0 ? K : 1));
  int              ierr = lhs_constr ? lhs_constr(num_species_, K, num_states, x, fval.data(), args_constr)
                                     : default_constr_fun(num_species_, K, num_states, x, fval.data(), nullptr);
  PACMENSLCHKERRQ(ierr);
  for (int k = 0; k < K; ++k)
    for (int i = 0; i < num_states; ++i) {
      satisfied[num_states * k + i] = (fval[K * i + k] <= rhs_constr(k)) ? 1 : 0;
      for (int j = 0; j < num_species_; ++j)
        if (x[num_species_ * i + j] < 0) satisfied[num_states * k + i] = 1;
    }
  return 0;
}

arma::Row<int> StateSetConstrained::GetShapeBounds() const { return arma::Row<int>(rhs_constr); }
int StateSetConstrained::GetNumConstraints() const { return (int) rhs_constr.n_elem; }

PacmenslErrorCode StateSetConstrained::SetShape(const fsp_constr_multi_fn &lhs_fun, arma::Row<int> &rhs_bounds, void *args) {
  lhs_constr = lhs_fun;
  rhs_constr = rhs_bounds;
  args_constr = args;
  using_default_ = false;
  return 0;
}
PacmenslErrorCode StateSetConstrained::SetShape(int num_constraints, const fsp_constr_multi_fn &lhs_fun, int *bounds, void *args) {
  lhs_constr = lhs_fun;
  rhs_constr = arma::Row<int>(bounds, num_constraints);
  args_constr = args;
  using_default_ = false;
  return 0;
}
PacmenslErrorCode StateSetConstrained::SetShapeBounds(arma::Row<PetscInt> &rhs_bounds) {
  rhs_constr = rhs_bounds;
  return 0;
}
PacmenslErrorCode StateSetConstrained::SetShapeBounds(int num_constraints, int *bounds) {
  rhs_constr = arma::Row<int>(bounds, num_constraints);
  return 0;
}

// src/StateSet/StateSetConstrained.cpp:223-236
PacmenslErrorCode StateSetConstrained::SetUp() {
  PacmenslErrorCode ierr = StateSetBase::SetUp();
  PACMENSLCHKERRQ(ierr);
  if (lhs_constr == nullptr) {
    if (num_species_ != (int) rhs_constr.n_elem) {
      PetscPrintf(comm_, "The number of constraint bounds when using default constraint must equal the number of species.\n");
      PACMENSLCHKERRQ(-1);
    }
    lhs_constr = default_constr_fun;
    using_default_ = true;
  }
  return 0;
}

PacmenslErrorCode StateSetConstrained::push_shape_to_device() {
  PacmenslErrorCode ierr = ensure_device_set();
  PACMENSLCHKERRQ(ierr);
  const int K = (int) rhs_constr.n_elem;
  if (using_default_ || lhs_constr == nullptr) {
    FSPCHKERRQ(fspset_set_shape(dset_, K, nullptr, rhs_constr.memptr(), nullptr));  // identity lhs on the device
  } else {
    FSPCHKERRQ(fspset_set_shape(dset_, K, &StateSetConstrained::lhs_trampoline, rhs_constr.memptr(), this));
  }
  return 0;
}

// src/StateSet/StateSetConstrained.cpp:132-221
PacmenslErrorCode StateSetConstrained::Expand() {
  if (!set_up_) {
    PacmenslErrorCode ierr = SetUp();
    PACMENSLCHKERRQ(ierr);
  }
  if (lhs_constr == nullptr) {  // SetUp() ran before the bounds were given
    if (num_species_ != (int) rhs_constr.n_elem) PACMENSLCHKERRQ(-1);
    lhs_constr = default_constr_fun;
    using_default_ = true;
  }
  PacmenslErrorCode ierr = push_shape_to_device();
  PACMENSLCHKERRQ(ierr);
  int rc = fspset_expand(dset_);
  if (rc != 0) {
    printf("PACMENSL device error: %s\n", fsp_last_error());
    return rc;
  }
  return update_layout();
}

}  // namespace pacmensl

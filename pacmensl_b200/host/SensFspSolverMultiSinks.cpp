#include "SensFspSolverMultiSinks.h"

namespace pacmensl {

SensFspSolverMultiSinks::SensFspSolverMultiSinks(MPI_Comm _comm, PartitioningType _part_type, ODESolverType) {
  comm_ = _comm;
  MPI_Comm_rank(_comm, &my_rank_);
  MPI_Comm_size(_comm, &comm_size_);
  partitioning_type_ = _part_type;
}

PacmenslErrorCode SensFspSolverMultiSinks::SetConstraintFunctions(const fsp_constr_multi_fn &lhs_constr, void *args) {
  fsp_constr_funs_ = lhs_constr;
  fsp_constr_args_ = args;
  have_custom_constraints_ = true;
  return 0;
}
PacmenslErrorCode SensFspSolverMultiSinks::SetInitialBounds(arma::Row<int> &_fsp_size) { fsp_bounds_ = _fsp_size; return 0; }
PacmenslErrorCode SensFspSolverMultiSinks::SetExpansionFactors(arma::Row<PetscReal> &f) { fsp_expasion_factors_ = f; return 0; }
PacmenslErrorCode SensFspSolverMultiSinks::SetModel(SensModel &model) { model_ = model; return 0; }
PacmenslErrorCode SensFspSolverMultiSinks::SetVerbosity(int verbosity_level) { verbosity_ = verbosity_level; return 0; }

PacmenslErrorCode SensFspSolverMultiSinks::SetInitialDistribution(const arma::Mat<pacmensl::Int> &_init_states,
                                                                  const arma::Col<PetscReal> &_init_probs,
                                                                  const std::vector<arma::Col<PetscReal>> &_init_sens) {
  // (the reference validates its OLD members here, SensFspSolverMultiSinks.cpp:75-81; the new values are checked instead)
  if (_init_probs.n_elem != _init_states.n_cols) return -1;
  for (auto &s : _init_sens) if (s.n_elem != _init_states.n_cols) return -1;
  init_states_ = _init_states;
  init_probs_ = _init_probs;
  init_sens_ = _init_sens;
  return 0;
}

PacmenslErrorCode SensFspSolverMultiSinks::SetInitialDistribution(SensDiscreteDistribution &init_dist) {
  PacmenslErrorCode ierr;
  int  n_states, n_species;
  int *states_ptr;
  ierr = init_dist.GetStateView(n_states, n_species, states_ptr); PACMENSLCHKERRQ(ierr);
  init_states_ = arma::Mat<int>(states_ptr, n_species, n_states);
  PetscReal *prob_ptr;
  ierr = init_dist.GetProbView(n_states, prob_ptr); PACMENSLCHKERRQ(ierr);
  init_probs_ = arma::Col<PetscReal>(prob_ptr, n_states);
  ierr = init_dist.RestoreProbView(prob_ptr); PACMENSLCHKERRQ(ierr);
  int n_sens = (int) init_dist.dp_.size();
  init_sens_.clear();
  for (int i = 0; i < n_sens; ++i) {
    ierr = init_dist.GetSensView(i, n_states, prob_ptr); PACMENSLCHKERRQ(ierr);
    init_sens_.emplace_back(arma::Col<PetscReal>(prob_ptr, n_states));
    ierr = init_dist.RestoreSensView(i, prob_ptr); PACMENSLCHKERRQ(ierr);
  }
  return 0;
}

PacmenslErrorCode SensFspSolverMultiSinks::SetLoadBalancingMethod(PartitioningType part_type) { partitioning_type_ = part_type; return 0; }
PacmenslErrorCode SensFspSolverMultiSinks::SetOdesType(ForwardSensType odes_type) { sens_solver_type = odes_type; return 0; }

// src/SensFsp/SensFspSolverMultiSinks.cpp:122-232
PacmenslErrorCode SensFspSolverMultiSinks::SetUp() {
  int ierr{0};
  try {
    if (model_.prop_x_ == nullptr) throw std::runtime_error("Propensity was not set before calling FspSolver.SetUp().");
    if (model_.stoichiometry_matrix_.n_elem == 0) throw std::runtime_error("Empty stoichiometry matrix cannot be used for FspSolver.");
    if (init_states_.n_elem == 0 || init_probs_.n_elem == 0)
      throw std::runtime_error("Initial states and/or probabilities were not set before calling FspSolver.SetUp().");
  } catch (std::runtime_error &e) {
    PetscPrintf(comm_, "\n %s \n", e.what());
    ierr = -1;
  }
  PACMENSLCHKERRQ(ierr);

  if (!state_set_) {
    state_set_ = std::make_shared<StateSetConstrained>(comm_);
    state_set_->SetStoichiometryMatrix(model_.stoichiometry_matrix_);
    if (have_custom_constraints_) state_set_->SetShape(fsp_constr_funs_, fsp_bounds_, fsp_constr_args_);
    else state_set_->SetShapeBounds(fsp_bounds_);
    ierr = state_set_->SetUp(); PACMENSLCHKERRQ(ierr);
    ierr = state_set_->AddStates(init_states_); PACMENSLCHKERRQ(ierr);
    ierr = state_set_->Expand(); PACMENSLCHKERRQ(ierr);
  }
  if (!A_) {
    A_ = std::make_shared<SensFspMatrix<FspMatrixConstrained>>(comm_);
    ierr = A_->GenerateValues(*state_set_, model_); PACMENSLCHKERRQ(ierr);
    matvec_ = [&](Real t, Vec x, Vec y) { return A_->Action(t, x, y); };
    dmatvec_ = [&](int is, Real t, Vec x, Vec y) { return A_->SensAction(is, t, x, y); };
  }
  auto error_checking_fp = [&](PetscReal t, Vec p, int, Vec *, void *) { return CheckFspTolerance_(t, p); };
  if (!sens_solver_) {
    sens_solver_ = std::make_shared<ForwardSensCvodeFsp>(comm_);
    sens_solver_->SetRhs(matvec_);
    sens_solver_->SetSensRhs(dmatvec_);
    sens_solver_->SetStopCondition(error_checking_fp, nullptr);
  }
  if (p_.IsEmpty()) {
    ierr = VecCreate(comm_, p_.mem()); PACMENSLCHKERRTHROW(ierr);
    ierr = VecSetSizes(p_, A_->GetNumLocalRows(), PETSC_DECIDE); PACMENSLCHKERRTHROW(ierr);
    ierr = VecSetUp(p_); PACMENSLCHKERRTHROW(ierr);
  }
  if (dp_.empty()) {
    dp_.resize(model_.num_parameters_);
    for (int i{0}; i < model_.num_parameters_; ++i) {
      ierr = VecCreate(comm_, dp_[i].mem()); PACMENSLCHKERRTHROW(ierr);
      ierr = VecSetSizes(dp_[i], A_->GetNumLocalRows(), PETSC_DECIDE); PACMENSLCHKERRTHROW(ierr);
      ierr = VecSetUp(dp_[i]); PACMENSLCHKERRTHROW(ierr);
    }
  }
  sinks_.set_size(state_set_->GetNumConstraints());
  to_expand_.set_size(sinks_.n_elem);
  set_up_ = true;
  return ierr;
}

const StateSetBase *SensFspSolverMultiSinks::GetStateSet() { return state_set_.get(); }

PacmenslErrorCode SensFspSolverMultiSinks::set_initial_vectors_() {
  PetscErrorCode ierr;
  ierr = VecSet(p_, 0.0); PACMENSLCHKERRQ(ierr);
  arma::Row<Int> indices = state_set_->State2Index(init_states_);
  ierr = VecSetValues(p_, PetscInt(init_probs_.n_elem), &indices[0], &init_probs_[0], INSERT_VALUES); PACMENSLCHKERRQ(ierr);
  ierr = VecAssemblyBegin(p_); PACMENSLCHKERRQ(ierr);
  ierr = VecAssemblyEnd(p_); PACMENSLCHKERRQ(ierr);
  for (int i{0}; i < model_.num_parameters_; ++i) {
    ierr = VecSet(dp_[i], 0.0); PACMENSLCHKERRQ(ierr);
    ierr = VecSetValues(dp_[i], PetscInt(init_probs_.n_elem), &indices[0], init_sens_[i].memptr(), INSERT_VALUES); PACMENSLCHKERRQ(ierr);
    ierr = VecAssemblyBegin(dp_[i]); PACMENSLCHKERRQ(ierr);
    ierr = VecAssemblyEnd(dp_[i]); PACMENSLCHKERRQ(ierr);
  }
  return 0;
}

// src/SensFsp/SensFspSolverMultiSinks.cpp:239-264
SensDiscreteDistribution SensFspSolverMultiSinks::Solve(PetscReal t_final, PetscReal fsp_tol) {
  PetscErrorCode ierr;
  if (!set_up_) { ierr = SetUp(); PACMENSLCHKERRTHROW(ierr); }
  if ((int) init_sens_.size() < model_.num_parameters_) PACMENSLCHKERRTHROW(-1);
  ierr = set_initial_vectors_(); PACMENSLCHKERRTHROW(ierr);
  t_now_ = 0.0;
  t_final_ = t_final;
  return Advance_(t_final, fsp_tol);
}

std::vector<SensDiscreteDistribution> SensFspSolverMultiSinks::SolveTspan(const std::vector<PetscReal> &tspan, PetscReal fsp_tol) {
  PetscErrorCode ierr;
  if (!set_up_) { ierr = SetUp(); PACMENSLCHKERRTHROW(ierr); }
  ierr = set_initial_vectors_(); PACMENSLCHKERRTHROW(ierr);
  std::vector<SensDiscreteDistribution> outputs;
  int num_time_points = (int) tspan.size();
  outputs.resize(num_time_points);
  t_now_ = 0.0;
  t_final_ = tspan[num_time_points - 1];
  for (int i = 0; i < num_time_points; ++i) outputs[i] = Advance_(tspan[i], fsp_tol);
  return outputs;
}

PacmenslErrorCode SensFspSolverMultiSinks::ClearState() {
  p_.release();
  dp_.clear();
  state_set_.reset();
  A_.reset();
  sens_solver_.reset();
  set_up_ = false;
  return 0;
}

SensFspSolverMultiSinks::~SensFspSolverMultiSinks() {
  ClearState();
  comm_ = MPI_COMM_NULL;
}

// src/SensFsp/SensFspSolverMultiSinks.cpp:303-331 -- strict '>' and to_expand_ reset on every check
int SensFspSolverMultiSinks::CheckFspTolerance_(PetscReal t, Vec p) {
  to_expand_.fill(0);
  const int            K = (int) sinks_.n_elem;
  arma::Row<PetscReal> sinks_of_p((arma::uword) K);
  sinks_of_p.fill(0.0);
  if (my_rank_ == comm_size_ - 1 && K > 0) {
    int n_loc = A_->GetNumLocalRows();
    const PetscReal *p_dev;
    VecGetDeviceArrayRead(p, &p_dev);
    int ierr = fsp_memcpy_d2h(sinks_of_p.memptr(), p_dev + (n_loc - K), sizeof(double) * K, comm_->stream);
    if (!ierr && comm_->nccl) ierr = fspcomm_check(comm_->nccl);
    PACMENSLCHKERRTHROW(ierr);
  }
  for (int i = 0; i < K; ++i) sinks_[i] = sinks_of_p[i];
  int ierr = pacmensl_allreduce_sum(comm_, sinks_.memptr(), K);
  PACMENSLCHKERRTHROW(ierr);
  int any = 0;
  for (int i{0}; i < K; ++i)
    if (sinks_(i) / fsp_tol_ > (1.0 / double(K)) * (t / t_final_)) { to_expand_(i) = 1; any = 1; }
  return any;
}

// src/SensFsp/SensFspSolverMultiSinks.cpp:333-422
SensDiscreteDistribution SensFspSolverMultiSinks::Advance_(PetscReal t_final, PetscReal fsp_tol) {
  PetscErrorCode ierr;
  PetscInt       solver_stat;
  if (verbosity_ > 1) sens_solver_->SetStatusOutput(1);
  fsp_tol_ = fsp_tol;
  sens_solver_->SetFinalTime(t_final);

  solver_stat = 1;
  while (solver_stat == 1) {
    ierr = sens_solver_->SetInitialSolution(p_); PACMENSLCHKERRTHROW(ierr);
    ierr = sens_solver_->SetInitialSensitivity(dp_); PACMENSLCHKERRTHROW(ierr);
    ierr = sens_solver_->SetCurrentTime(t_now_); PACMENSLCHKERRTHROW(ierr);
    ierr = sens_solver_->SetUp(); PACMENSLCHKERRTHROW(ierr);
    solver_stat = sens_solver_->Solve();
    if (solver_stat != 0 && solver_stat != 1) PACMENSLCHKERRTHROW(solver_stat);
    t_now_ = sens_solver_->GetCurrentTime();
    ierr = sens_solver_->FreeWorkspace(); PACMENSLCHKERRTHROW(ierr);

    if (solver_stat == 1) {
      num_expansions_ += 1;
      for (int i{0}; i < (int) to_expand_.n_elem; ++i)
        if (to_expand_(i) == 1)
          fsp_bounds_(i) = (int) std::round(double(fsp_bounds_(i)) * (fsp_expasion_factors_(i) + 1.0e0) + 0.5e0);
      const int n_old = state_set_->GetNumLocalStates();
      const int old_start = state_set_->GetLocalStart();
      const bool sharded = state_set_->IsSharded();
      if (sharded) { ierr = state_set_->RememberLocalStates(); PACMENSLCHKERRTHROW(ierr); }
      state_set_->SetShapeBounds(fsp_bounds_);
      ierr = state_set_->Expand(); PACMENSLCHKERRTHROW(ierr);
      A_->Destroy();
      ierr = A_->GenerateValues(*state_set_, model_); PACMENSLCHKERRTHROW(ierr);
      // existing states keep their global index (replicated directory): State2Index(states_old) is the identity shift;
      // a sharded set re-numbers when it re-balances, so the old block is looked up again (:174-205 of the reference)
      std::vector<PetscInt> new_locations_vals((size_t) n_old);
      if (sharded) { ierr = state_set_->RememberedIndices(new_locations_vals); PACMENSLCHKERRTHROW(ierr); }
      else for (int i = 0; i < n_old; ++i) new_locations_vals[i] = old_start + i;
      if (my_rank_ == comm_size_ - 1) {
        Int i_end_new = state_set_->GetNumGlobalStates() + (Int) sinks_.n_elem;
        for (int i{0}; i < (int) sinks_.n_elem; ++i) new_locations_vals.push_back(i_end_new - ((Int) sinks_.n_elem) + i);
      }
      ierr = ExpandVec(p_, new_locations_vals, A_->GetNumLocalRows()); PACMENSLCHKERRTHROW(ierr);
      for (int i{0}; i < model_.num_parameters_; ++i) {
        ierr = ExpandVec(dp_[i], new_locations_vals, A_->GetNumLocalRows()); PACMENSLCHKERRTHROW(ierr);
      }
      if (verbosity_ > 0) {
        PetscPrintf(comm_, "\n ------------- \n");
        PetscPrintf(comm_, "At time t = %.2f expansion to new state_set_ size: \n", t_now_);
        for (int i{0}; i < (int) fsp_bounds_.n_elem; ++i) PetscPrintf(comm_, "%d ", fsp_bounds_[i]);
        PetscPrintf(comm_, "New Fsp number of states_: %d \n", state_set_->GetNumGlobalStates());
        PetscPrintf(comm_, "\n ------------- \n");
      }
    }
  }
  SensDiscreteDistribution out;
  ierr = MakeSensDiscreteDistribution_(out); PACMENSLCHKERRTHROW(ierr);
  return out;
}

// src/SensFsp/SensFspSolverMultiSinks.cpp:424-458
PacmenslErrorCode SensFspSolverMultiSinks::MakeSensDiscreteDistribution_(SensDiscreteDistribution &dist) {
  PacmenslErrorCode ierr;
  dist.comm_ = comm_;
  dist.t_ = t_now_;
  dist.states_ = state_set_->CopyStatesOnProc();
  dist.AttachDeviceStates(state_set_.get());
  const int n = state_set_->GetNumLocalStates();
  ierr = VecCreate(dist.comm_, &dist.p_); CHKERRQ(ierr);
  ierr = VecSetSizes(dist.p_, n, PETSC_DECIDE); CHKERRQ(ierr);
  ierr = VecSetUp(dist.p_); CHKERRQ(ierr);
  Vec pv = p_;
  FSPCHKERRQ(fspvec_copy(dist.p_->d_data, pv->d_data, n, comm_->stream));
  dist.dp_.resize(dp_.size());
  for (size_t i{0}; i < dp_.size(); ++i) {
    ierr = VecDuplicate(dist.p_, &dist.dp_[i]); CHKERRQ(ierr);
    Vec dv = dp_[i];
    FSPCHKERRQ(fspvec_copy(dist.dp_[i]->d_data, dv->d_data, n, comm_->stream));
  }
  FSPCHKERRQ(fsp_stream_sync(comm_->stream));
  return 0;
}

}  // namespace pacmensl

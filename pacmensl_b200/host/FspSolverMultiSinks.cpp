#include "FspSolverMultiSinks.h"

#include <cstdlib>

namespace pacmensl {

static double now_s() {
  PetscLogDouble t;
  PetscTime(&t);
  return t;
}

FspSolverMultiSinks::FspSolverMultiSinks(MPI_Comm _comm, PartitioningType _part_type, ODESolverType _solve_type) {
  comm_ = _comm;
  MPI_Comm_rank(comm_, &my_rank_);
  MPI_Comm_size(comm_, &comm_size_);
  partitioning_type_ = _part_type;
  odes_type_ = _solve_type;
}

PacmenslErrorCode FspSolverMultiSinks::SetInitialBounds(arma::Row<int> &_bounds) {
  fsp_bounds_ = _bounds;
  return 0;
}
PacmenslErrorCode FspSolverMultiSinks::SetConstraintFunctions(const fsp_constr_multi_fn &lhs_constr, void *args) {
  fsp_constr_funs_ = lhs_constr;
  fsp_constr_args_ = args;
  has_custom_constraints_ = true;
  return 0;
}
PacmenslErrorCode FspSolverMultiSinks::SetExpansionFactors(arma::Row<PetscReal> &_expansion_factors) {
  fsp_expasion_factors_ = _expansion_factors;
  return 0;
}

// src/Fsp/FspSolverMultiSinks.cpp:62-224
DiscreteDistribution FspSolverMultiSinks::Advance_(PetscReal t_final, PetscReal fsp_tol) {
  PetscErrorCode ierr;
  PetscInt       solver_stat;
  const double   t_solve0 = now_s();

  if (verbosity_ > 1) ode_solver_->SetStatusOutput(1);

  fsp_tol_ = fsp_tol;
  ode_solver_->SetFinalTime(t_final);
  ode_solver_->SetTolerances(ode_rtol_, ode_atol_);
  ode_solver_->SetRhs(this->tmatvec_);
  static const bool use_fused = [] { const char *e = std::getenv("FSP_FUSED_ACTION"); return !(e && e[0] == '0'); }();
  if (use_fused) ode_solver_->SetFusedRhs([this](Real t, Vec x, Vec y, const fspmat_epilogue &ep) {
    // the same operator as tmatvec_, with the solver's scaling / inner products fused into the kernel
    if (!logging_enabled) return A_->ActionFused(t, x, y, ep);
    double t1 = now_s();
    int    ie = A_->ActionFused(t, x, y, ep);
    t_rhs_ += now_s() - t1;
    PetscInt f;
    A_->GetLocalMVFlops(&f);
    flops_ += f;
    return ie;
  });
  if (fsp_tol_ > 0.0) {
    auto error_checking_fp = [&](PetscReal t, Vec p, PetscReal &te, void *) { return CheckFspTolerance_(t, p, te); };
    ode_solver_->SetStopCondition(error_checking_fp, nullptr);
  } else {
    ode_solver_->SetStopCondition(nullptr, nullptr);
  }

  solver_stat = 1;
  while (solver_stat) {
    double t0 = now_s();
    ierr = ode_solver_->SetInitialSolution(p_->mem());
    PACMENSLCHKERRTHROW(ierr);
    ierr = ode_solver_->SetCurrentTime(t_now_);
    PACMENSLCHKERRTHROW(ierr);
    ode_solver_->SetTolerances(ode_rtol_, ode_atol_);
    ierr = ode_solver_->SetUp();
    PACMENSLCHKERRTHROW(ierr);

    to_expand_.fill(0);

    solver_stat = ode_solver_->Solve();
    if (solver_stat != 0 && solver_stat != 1) PACMENSLCHKERRTHROW(solver_stat);

    ierr = ode_solver_->FreeWorkspace();
    PACMENSLCHKERRTHROW(ierr);
    t_ode_ += now_s() - t0;

    // Expand the FSP state space if the solver halted prematurely (:113-205)
    if (solver_stat == 1) {
      num_expansions_ += 1;
      for (auto i{0}; i < (int) to_expand_.n_elem; ++i) {
        if (to_expand_(i) == 1) {
          fsp_bounds_(i) = (int) std::round(double(fsp_bounds_(i)) * (fsp_expasion_factors_(i) + 1.0e0) + 0.5e0);
        }
      }
      if (verbosity_) {
        PetscPrintf(comm_, "\n ------------- \n");
        PetscPrintf(comm_, "At time t = %.2f expansion to new state_set_ size: \n", ode_solver_->GetCurrentTime());
        for (auto i{0}; i < (int) fsp_bounds_.n_elem; ++i) PetscPrintf(comm_, "%d ", fsp_bounds_[i]);
        PetscPrintf(comm_, "\n ------------- \n");
      }
      // Remember where the current solution's entries are: with the replicated directory the global index of an
      // existing state never changes, so State2Index(states_old) is the identity on [old_start, old_start + n_old).
      const int n_old = state_set_->GetNumLocalStates();
      const int old_start = state_set_->GetLocalStart();
      const bool sharded = state_set_->IsSharded();
      t0 = now_s();
      // a sharded set re-numbers when it re-balances: the old block is looked up again after the expansion
      // (State2Index(states_old), :174-176 of the reference)
      if (sharded) { ierr = state_set_->RememberLocalStates(); PACMENSLCHKERRTHROW(ierr); }
      state_set_->SetShapeBounds(fsp_bounds_);
      ierr = state_set_->Expand();
      PACMENSLCHKERRTHROW(ierr);
      t_partition_ += now_s() - t0;
      if (verbosity_) {
        PetscPrintf(comm_, "\n ------------- \n");
        PetscPrintf(comm_, "New Fsp number of states_: %d \n", state_set_->GetNumGlobalStates());
        PetscPrintf(comm_, "\n ------------- \n");
      }

      t0 = now_s();
      A_->Destroy();
      ierr = A_->GenerateValues(*state_set_, model_);
      PACMENSLCHKERRTHROW(ierr);
      t_matgen_ += now_s() - t0;

      // Generate the expanded vector and scatter forward the current solution (:174-205)
      t0 = now_s();
      std::vector<PetscInt> new_locations_vals((size_t) n_old);
      if (sharded) { ierr = state_set_->RememberedIndices(new_locations_vals); PACMENSLCHKERRTHROW(ierr); }
      else for (int i = 0; i < n_old; ++i) new_locations_vals[i] = old_start + i;
      if (my_rank_ == comm_size_ - 1) {
        Int i_end_new = state_set_->GetNumGlobalStates() + (Int) sinks_.n_elem;
        for (auto i{0}; i < (int) sinks_.n_elem; ++i) new_locations_vals.push_back(i_end_new - ((Int) sinks_.n_elem) + i);
      }
      ierr = ExpandVec(*p_, new_locations_vals, A_->GetNumLocalRows());
      PACMENSLCHKERRTHROW(ierr);
      // warm restart: the integrator's own history follows the solution onto the enlarged state space (0 = carried
      // over, 1 = this solver restarts cold at the next SetUp, as the reference always does)
      if (ode_solver_->ExpandState(new_locations_vals, A_->GetNumLocalRows()) == 0) num_warm_restarts_ += 1;
      t_scatter_ += now_s() - t0;
    }
    t_now_ = ode_solver_->GetCurrentTime();
  }
  t_solve_ += now_s() - t_solve0;

  DiscreteDistribution dist;
  ierr = MakeDiscreteDistribution_(dist);
  PACMENSLCHKERRTHROW(ierr);
  return dist;
}

FspSolverMultiSinks::~FspSolverMultiSinks() {
  ClearState();
  comm_ = MPI_COMM_NULL;
}

PacmenslErrorCode FspSolverMultiSinks::ClearState() {
  set_up_ = false;
  if (ode_solver_) rhs_evals_retired_ += ode_solver_->GetNumRhsEvals();
  ode_solver_.reset();
  p_.reset();
  A_.reset();
  state_set_.reset();
  sinks_.clear();
  to_expand_.clear();
  has_custom_constraints_ = false;
  fsp_constr_args_ = nullptr;
  fsp_constr_funs_ = nullptr;
  tmatvec_ = nullptr;
  return 0;
}

// src/Fsp/FspSolverMultiSinks.cpp:251-412
PacmenslErrorCode FspSolverMultiSinks::SetUp() {
  int ierr{0};
  const double t_setup0 = now_s();
  try {
    if ((model_.prop_t_ == nullptr) && (!model_.tv_reactions_.empty()))
      throw std::runtime_error("Model has time-varying propensitites but temporal signals were not set before calling FspSolver.SetUp().");
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
    double t0 = now_s();
    state_set_ = std::make_shared<StateSetConstrained>(comm_);
    if (sharded_set_ >= 0) state_set_->SetSharded(sharded_set_ != 0);
    state_set_->SetStoichiometryMatrix(model_.stoichiometry_matrix_);
    if (has_custom_constraints_) state_set_->SetShape(fsp_constr_funs_, fsp_bounds_, fsp_constr_args_);
    else state_set_->SetShapeBounds(fsp_bounds_);
    state_set_->SetLoadBalancingScheme(partitioning_type_);
    ierr = state_set_->SetUp();
    PACMENSLCHKERRQ(ierr);
    ierr = state_set_->AddStates(init_states_);
    PACMENSLCHKERRQ(ierr);
    ierr = state_set_->Expand();
    PACMENSLCHKERRQ(ierr);
    t_partition_ += now_s() - t0;
  }

  if (!A_) {
    double t0 = now_s();
    A_ = std::make_shared<FspMatrixConstrained>(comm_);
    A_->SetIncrementalGeneration(true);  // the driver owns model and state set: regenerate only for new states
    ierr = A_->GenerateValues(*state_set_, model_);
    PACMENSLCHKERRQ(ierr);
    t_matgen_ += now_s() - t0;
    if (logging_enabled) {
      tmatvec_ = [&](Real t, Vec x, Vec y) {
        double t1 = now_s();
        int    ie = A_->Action(t, x, y);
        t_rhs_ += now_s() - t1;
        PetscInt f;
        A_->GetLocalMVFlops(&f);
        flops_ += f;
        return ie;
      };
    } else {
      tmatvec_ = [&](Real t, Vec x, Vec y) { return A_->Action(t, x, y); };
    }
  }
  A_->SetTimeFun(model_.prop_t_, model_.prop_t_args_);

  if (!p_) {
    p_ = std::make_shared<Petsc<Vec>>();
    ierr = VecCreate(comm_, p_->mem());
    PACMENSLCHKERRTHROW(ierr);
    ierr = VecSetSizes(*p_, A_->GetNumLocalRows(), PETSC_DECIDE);
    PACMENSLCHKERRTHROW(ierr);
    ierr = VecSetType(*p_, VECMPI);
    PACMENSLCHKERRTHROW(ierr);
    ierr = VecSetUp(*p_);
    PACMENSLCHKERRTHROW(ierr);
  }

  if (!ode_solver_) {
    switch (odes_type_) {
      case CVODE: ode_solver_ = std::make_shared<CvodeFsp>(comm_); break;
      case KRYLOV:
        ode_solver_ = std::make_shared<KrylovFsp>(comm_);
        if (custom_krylov_) {
          ((KrylovFsp *) ode_solver_.get())->SetOrthLength(q_iop_);
          ((KrylovFsp *) ode_solver_.get())->SetKrylovDimRange(m_min_, m_max_);
        }
        break;
      default:
        // ODESolverType::PETSC: TsFsp = Rosenbrock-W (RA34PW2, PETSc's TSROSW default scheme) on the assembled CSR
        // Jacobian, see TsFsp.h
        ode_solver_ = std::make_shared<TsFsp>(comm_);
    }
    ode_solver_->SetFspMatPtr(A_.get());
    static const bool warm_env = [] { const char *e = std::getenv("FSP_WARM_RESTART"); return e && e[0] == '1'; }();
    ode_solver_->SetWarmRestart(warm_restart_ || warm_env);
    if (logging_enabled) ode_solver_->EnableLogging();
  }

  sinks_.set_size(state_set_->GetNumConstraints());
  to_expand_.set_size(sinks_.n_elem);
  set_up_ = true;
  t_setup_ += now_s() - t_setup0;
  return ierr;
}

PacmenslErrorCode FspSolverMultiSinks::SetVerbosity(int verbosity_level) {
  verbosity_ = verbosity_level;
  return 0;
}

PacmenslErrorCode FspSolverMultiSinks::SetInitialDistribution(const arma::Mat<Int> &_init_states,
                                                              const arma::Col<PetscReal> &_init_probs) {
  init_states_ = _init_states;
  init_probs_ = _init_probs;
  if (init_probs_.n_elem != init_states_.n_cols) return -1;
  return 0;
}

PacmenslErrorCode FspSolverMultiSinks::SetInitialDistribution(DiscreteDistribution &init_dist) {
  int        n_states, n_species;
  int       *state_ptr;
  PetscReal *prob_ptr;
  int        ierr = init_dist.GetStateView(n_states, n_species, state_ptr);
  PACMENSLCHKERRQ(ierr);
  init_states_ = arma::Mat<PetscInt>(state_ptr, n_species, n_states);
  init_dist.GetProbView(n_states, prob_ptr);
  init_probs_ = arma::Col<PetscReal>(prob_ptr, n_states);
  init_dist.RestoreProbView(prob_ptr);
  if (init_probs_.n_elem != init_states_.n_cols) return -1;
  return 0;
}

std::shared_ptr<const StateSetBase> FspSolverMultiSinks::GetStateSet() { return state_set_; }

PacmenslErrorCode FspSolverMultiSinks::SetLogging(PetscBool logging) {
  logging_enabled = logging;
  return 0;
}

// src/Fsp/FspSolverMultiSinks.cpp:467-516
FspSolverComponentTiming FspSolverMultiSinks::ReduceComponentTiming(char *op) {
  auto reduce = [&](double v) {
    double out = v;
    if (std::strcmp(op, "sum") == 0) {
      pacmensl_allreduce_sum(comm_, &out, 1);
    } else if (std::strcmp(op, "min") == 0) {
      out = -out;
      pacmensl_allreduce_max(comm_, &out, 1);
      out = -out;
    } else {
      pacmensl_allreduce_max(comm_, &out, 1);
    }
    return out;
  };
  FspSolverComponentTiming timings;
  timings.MatrixGenerationTime = reduce(t_matgen_);
  timings.StatePartitioningTime = reduce(t_partition_);
  timings.ODESolveTime = reduce(t_ode_);
  timings.RHSEvalTime = reduce(t_rhs_);
  timings.SolutionScatterTime = reduce(t_scatter_);
  timings.TotalTime = reduce(t_setup_) + reduce(t_solve_);
  timings.TotalFlops = reduce(flops_);
  return timings;
}

FiniteProblemSolverPerfInfo FspSolverMultiSinks::GetSolverPerfInfo() { return ode_solver_->GetAvgPerfInfo(); }

// src/Fsp/FspSolverMultiSinks.cpp:523-574: the PETSc options database is replaced by environment variables
// FSP_PARTITIONING_TYPE, FSP_REPART_APPROACH, FSP_VERBOSITY, FSP_LOG_EVENTS with the same values.
PacmenslErrorCode FspSolverMultiSinks::SetFromOptions() {
  if (const char *opt = std::getenv("FSP_PARTITIONING_TYPE")) partitioning_type_ = str2part(std::string(opt));
  if (comm_size_ == 1) partitioning_type_ = PartitioningType::GRAPH;
  if (const char *opt = std::getenv("FSP_REPART_APPROACH")) repart_approach_ = str2partapproach(std::string(opt));
  if (const char *opt = std::getenv("FSP_VERBOSITY")) {
    if (std::strcmp(opt, "1") == 0 || std::strcmp(opt, "true") == 0) verbosity_ = 1;
    if (std::strcmp(opt, "2") == 0) verbosity_ = 2;
  }
  if (const char *opt = std::getenv("FSP_LOG_EVENTS")) {
    if (std::strcmp(opt, "1") == 0 || std::strcmp(opt, "true") == 0) logging_enabled = PETSC_TRUE;
  }
  return 0;
}

// src/Fsp/FspSolverMultiSinks.cpp:576-611
PacmenslErrorCode FspSolverMultiSinks::CheckFspTolerance_(PetscReal t, Vec p, PetscReal &tol_exceed) {
  int ierr;
  tol_exceed = 0.0;
  const int            K = (int) sinks_.n_elem;
  arma::Row<PetscReal> sinks_of_p((arma::uword) K);
  sinks_of_p.fill(0.0);
  if (my_rank_ == comm_size_ - 1 && K > 0) {
    // the K sink entries are the last K local entries on the last rank: one small D2H read per ODE step
    int n_loc;
    VecGetLocalSize(p, &n_loc);
    const PetscReal *p_dev;
    VecGetDeviceArrayRead(p, &p_dev);
    ierr = fsp_memcpy_d2h(sinks_of_p.memptr(), p_dev + (n_loc - K), sizeof(double) * K, comm_->stream);
    if (!ierr && comm_->nccl) ierr = fspcomm_check(comm_->nccl);
    PACMENSLCHKERRTHROW(ierr);
  }
  for (int i = 0; i < K; ++i) sinks_[i] = sinks_of_p[i];
  ierr = pacmensl_allreduce_sum(comm_, sinks_.memptr(), K);
  PACMENSLCHKERRTHROW(ierr);
  for (int i{0}; i < K; ++i) {
    if (sinks_(i) / fsp_tol_ >= (1.0 / double(K)) * (t / t_final_)) {
      to_expand_(i) = 1;
      tol_exceed = std::max(tol_exceed, sinks_(i) * double(K) - fsp_tol_ * (t / t_final_));
    }
  }
  return 0;
}

PacmenslErrorCode FspSolverMultiSinks::SetModel(Model &model) {
  FspSolverMultiSinks::model_ = model;
  // a new model invalidates the propensity values cached for incremental regeneration (they are keyed on the number
  // of reactions and states only, not on the callback)
  if (A_) A_->ResetGenerationCache();
  return 0;
}

// src/Fsp/FspSolverMultiSinks.cpp:619-643
DiscreteDistribution FspSolverMultiSinks::Solve(PetscReal t_final, PetscReal fsp_tol, PetscReal t_init) {
  PetscErrorCode ierr;
  if (!set_up_) {
    ierr = SetUp();
    PACMENSLCHKERRTHROW(ierr);
  }
  ierr = VecSet(*p_, 0.0);
  PACMENSLCHKERRTHROW(ierr);
  arma::Row<Int> indices = state_set_->State2Index(init_states_);
  ierr = VecSetValues(*p_, PetscInt(init_probs_.n_elem), &indices[0], &init_probs_[0], INSERT_VALUES);
  PACMENSLCHKERRTHROW(ierr);
  ierr = VecAssemblyBegin(*p_);
  PACMENSLCHKERRTHROW(ierr);
  ierr = VecAssemblyEnd(*p_);
  PACMENSLCHKERRTHROW(ierr);

  t_now_ = t_init;
  t_final_ = t_final;
  return FspSolverMultiSinks::Advance_(t_final, fsp_tol);
}

// src/Fsp/FspSolverMultiSinks.cpp:645-682
std::vector<DiscreteDistribution> FspSolverMultiSinks::SolveTspan(const std::vector<PetscReal> &tspan, PetscReal fsp_tol,
                                                                  PetscReal t_init) {
  PetscErrorCode ierr;
  if (!set_up_) {
    ierr = SetUp();
    PACMENSLCHKERRTHROW(ierr);
  }
  ierr = VecSet(*p_, 0.0);
  PACMENSLCHKERRTHROW(ierr);
  arma::Row<Int> indices = state_set_->State2Index(init_states_);
  ierr = VecSetValues(*p_, PetscInt(init_probs_.n_elem), &indices[0], &init_probs_[0], INSERT_VALUES);
  PACMENSLCHKERRTHROW(ierr);
  ierr = VecAssemblyBegin(*p_);
  PACMENSLCHKERRTHROW(ierr);
  ierr = VecAssemblyEnd(*p_);
  PACMENSLCHKERRTHROW(ierr);

  std::vector<DiscreteDistribution> outputs;
  int                               num_time_points = (int) tspan.size();
  outputs.resize(num_time_points);
  PetscReal t_max = tspan[num_time_points - 1];

  t_now_ = t_init;
  t_final_ = t_max;
  for (int i = 0; i < num_time_points; ++i) outputs[i] = FspSolverMultiSinks::Advance_(tspan[i], fsp_tol);
  return outputs;
}

PacmenslErrorCode FspSolverMultiSinks::SetOdesType(ODESolverType odes_type) {
  odes_type_ = odes_type;
  return 0;
}
PacmenslErrorCode FspSolverMultiSinks::SetLoadBalancingMethod(PartitioningType part_type) {
  partitioning_type_ = part_type;
  return 0;
}
PacmenslErrorCode FspSolverMultiSinks::SetOdeTolerances(PetscReal rel_tol, PetscReal abs_tol) {
  ode_rtol_ = rel_tol;
  ode_atol_ = abs_tol;
  return 0;
}

// src/Fsp/FspSolverMultiSinks.cpp:703-735: strip the sink entries
PacmenslErrorCode FspSolverMultiSinks::MakeDiscreteDistribution_(DiscreteDistribution &dist) {
  PacmenslErrorCode ierr;
  dist.comm_ = comm_;
  dist.t_ = t_now_;
  dist.states_ = state_set_->CopyStatesOnProc();
  dist.AttachDeviceStates(state_set_.get());
  ierr = VecCreate(dist.comm_, &dist.p_);
  CHKERRQ(ierr);
  ierr = VecSetSizes(dist.p_, state_set_->GetNumLocalStates(), PETSC_DECIDE);
  CHKERRQ(ierr);
  ierr = VecSetUp(dist.p_);
  CHKERRQ(ierr);
  // local states occupy the leading entries of the local block of p_ (sinks trail on the last rank)
  const PetscReal *src;
  PetscReal       *dst;
  VecGetDeviceArrayRead(*p_, &src);
  VecGetDeviceArray(dist.p_, &dst);
  FSPCHKERRQ(fspvec_copy(dst, src, state_set_->GetNumLocalStates(), comm_->stream));
  FSPCHKERRQ(fsp_stream_sync(comm_->stream));
  return 0;
}

std::shared_ptr<OdeSolverBase> FspSolverMultiSinks::GetOdeSolver() { return ode_solver_; }

PacmenslErrorCode FspSolverMultiSinks::SetOdesPetscType(std::string ts_type) {
  custom_ts_type_ = true;
  ts_type_ = ts_type;
  return 0;
}
PacmenslErrorCode FspSolverMultiSinks::SetKrylovOrthLength(int q) {
  if (odes_type_ != KRYLOV) return 0;
  if (ode_solver_ != nullptr) {
    ((KrylovFsp *) ode_solver_.get())->SetOrthLength(q);
  } else {
    custom_krylov_ = true;
    q_iop_ = q;
  }
  return 0;
}
PacmenslErrorCode FspSolverMultiSinks::SetKrylovDimRange(int m_min, int m_max) {
  if (odes_type_ != KRYLOV) return 0;
  if (ode_solver_ != nullptr) {
    ((KrylovFsp *) ode_solver_.get())->SetKrylovDimRange(m_min, m_max);
  } else {
    custom_krylov_ = true;
    m_min_ = m_min;
    m_max_ = m_max;
  }
  return 0;
}

}  // namespace pacmensl

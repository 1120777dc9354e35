// SensFspMatrix.h -- A(t) together with its parameter derivatives dA/dtheta_i = (dc/dtheta_i) x A + c x (dA/dtheta_i).
// Mirrors src/SensFsp/SensFspMatrix.h:44-209: per parameter two more operators sharing A's sparsity pattern,
//   dcxA_[i]: reactions dprop_t_sp_[i] with coefficient callback dprop_t(i, .) and the state factors prop_x,
//   cxdA_[i]: reactions dprop_x_sp_[i] with coefficients prop_t and the state factors dprop_x(i, .),
// SensAction(i) = dcxA_i x + cxdA_i x; operators that were never generated act as zero (FspMatrixBase.cpp:41).
// Every operator is a fused device operator (fspmat_*), so one SensAction is at most two kernel launches + one axpy.
#pragma once

#include "FspMatrixConstrained.h"
#include "PetscWrap.h"
#include "SensModel.h"

namespace pacmensl {

template <typename FspMatrixT>
class SensFspMatrix {
 public:
  NOT_COPYABLE_NOT_MOVABLE(SensFspMatrix);

  explicit SensFspMatrix(MPI_Comm comm) : A_(comm) {
    if (comm == MPI_COMM_NULL) std::printf("Null pointer detected.\n");
    comm_ = comm;
    MPI_Comm_rank(comm, &rank_);
  }
  virtual ~SensFspMatrix() {
    Destroy();
    comm_ = MPI_COMM_NULL;
  }

  virtual PacmenslErrorCode GenerateValues(const StateSetBase &state_set, const SensModel &model) {
    PacmenslErrorCode ierr;
    ierr = A_.GenerateValues(state_set, model.stoichiometry_matrix_, model.tv_reactions_, model.prop_t_, model.prop_x_,
                             std::vector<int>(), model.prop_t_args_, model.prop_x_args_);
    PACMENSLCHKERRQ(ierr);
    if (work_ != nullptr) VecDestroy(&work_);
    ierr = VecCreate(comm_, &work_);
    CHKERRQ(ierr);
    ierr = VecSetSizes(work_, A_.GetNumLocalRows(), PETSC_DECIDE);
    CHKERRQ(ierr);
    ierr = VecSetUp(work_);
    CHKERRQ(ierr);

    num_parameters_ = model.num_parameters_;
    dcxA_.clear();
    cxdA_.clear();
    // first part: dc x A
    for (int i{0}; i < num_parameters_; ++i) {
      dcxA_.emplace_back(new FspMatrixT(comm_));
      if (!model.dprop_t_sp_.empty() && !model.dprop_t_sp_[i].empty()) {
        auto dprop_t = std::bind(model.dprop_t_, i, std::placeholders::_1, std::placeholders::_2, std::placeholders::_3,
                                 std::placeholders::_4);
        ierr = dcxA_[i]->GenerateValues(state_set, model.stoichiometry_matrix_, model.tv_reactions_, dprop_t,
                                        model.prop_x_, model.dprop_t_sp_[i], model.dprop_t_args_, model.prop_x_args_);
        PACMENSLCHKERRQ(ierr);
      }
    }
    // second part: c x dA
    for (int i{0}; i < num_parameters_; ++i) {
      cxdA_.emplace_back(new FspMatrixT(comm_));
      if (!model.dprop_x_sp_.empty() && !model.dprop_x_sp_[i].empty()) {
        auto dprop_x = std::bind(model.dprop_x_, i, std::placeholders::_1, std::placeholders::_2, std::placeholders::_3,
                                 std::placeholders::_4, std::placeholders::_5, std::placeholders::_6);
        ierr = cxdA_[i]->GenerateValues(state_set, model.stoichiometry_matrix_, model.tv_reactions_, model.prop_t_,
                                        dprop_x, model.dprop_x_sp_[i], model.prop_t_args_, model.dprop_x_args_);
        PACMENSLCHKERRQ(ierr);
      }
    }
    return 0;
  }

  virtual PacmenslErrorCode Action(PetscReal t, Vec x, Vec y) { return A_.Action(t, x, y); }

  virtual PacmenslErrorCode SensAction(int i_par, PetscReal t, Vec x, Vec y) {
    int ierr;
    if (i_par < 0 || i_par >= num_parameters_) return -1;
    ierr = dcxA_[i_par]->Action(t, x, y);
    PACMENSLCHKERRQ(ierr);
    ierr = cxdA_[i_par]->Action(t, x, work_);
    PACMENSLCHKERRQ(ierr);
    ierr = VecAXPY(y, 1.0, work_);
    CHKERRQ(ierr);
    return 0;
  }

  virtual PacmenslErrorCode Destroy() {
    if (work_ != nullptr) VecDestroy(&work_);
    A_.Destroy();
    for (auto &m : dcxA_) m->Destroy();
    for (auto &m : cxdA_) m->Destroy();
    dcxA_.clear();
    cxdA_.clear();
    return 0;
  }

  int GetNumLocalRows() const { return A_.GetNumLocalRows(); }

 protected:
  MPI_Comm comm_ = MPI_COMM_NULL;
  int      rank_ = 0;
  int      num_parameters_ = 0;
  FspMatrixT A_;
  std::vector<std::unique_ptr<FspMatrixT>> dcxA_, cxdA_;
  Vec work_ = nullptr;
};
}  // namespace pacmensl

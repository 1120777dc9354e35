// KrylovFsp.h -- Expokit-style Krylov exponential integrator with incomplete orthogonalisation.
// Mirrors src/OdeSolver/KrylovFsp.h:34-95 / KrylovFsp.cpp:29-485 (same controller, same constants).
// The basis generation runs as fused device kernels with device-resident Hessenberg coefficients: per
// basis vector 1 Action + (q + 1) fused MGS passes + 1 scale, no host synchronisation inside the loop.
#pragma once

#include <map>

#include "OdeSolverBase.h"

namespace pacmensl {
class PACMENSL_API KrylovFsp : public OdeSolverBase {
 public:
  explicit KrylovFsp(MPI_Comm comm);

  int SetUp() override;
  PetscInt Solve() override;
  PacmenslErrorCode SetOrthLength(int q);
  PacmenslErrorCode SetKrylovDimRange(int m_min, int m_max);
  int FreeWorkspace() override;
  ~KrylovFsp();

 protected:
  const int max_reject_ = 10000;
  PetscReal delta_ = 1.2, gamma_ = 0.9;  ///< safety factors

  int m_min_ = 25, m_max_ = 60, m_next_ = 25;
  int m_ = 30;
  int q_iop = 2;

  int       k1 = 2;
  int       mb = 0, mx = 0;
  PetscReal beta = 0.0, avnorm = 0.0;
  std::vector<Vec>     Vm;
  arma::Mat<PetscReal> Hm;
  arma::Mat<PetscReal> F;
  Vec                  av = nullptr;
  Vec                  solution_tmp_ = nullptr;

  PetscReal t_now_tmp_ = 0.0;
  PetscReal t_step_ = 0.0;
  PetscReal t_step_next_ = 0.0;
  bool      first_step_initialized_ = false;
  PetscReal btol_ = 1.0e-14;
  int       krylov_stat_ = 0;

  DeviceBuffer<double> hdev_;  ///< device-resident orthogonalisation coefficients of the current basis
  std::vector<double>  hhost_;

  int SetUpWorkSpace();
  int EnsureBasis_(int count);  ///< create basis vectors Vm[0..count) on first use
  int BasisColumns_(int m_start);  ///< the column loop of GenerateBasis as an asynchronous device pipeline
  // CUDA graphs of the column loop, keyed by (m_start, m); single rank only (FSP_KRYLOV_GRAPH=0 disables)
  std::map<long, fsp_graph_t> graphs_;
  std::map<long, int>         graph_seen_;
  void                       *capture_stream_ = nullptr;
  bool                        graphs_disabled_ = false;
  bool GraphsUsable_();
  void DestroyGraphs_();
  int GenerateBasis(const Vec &v, int m_start, PetscBool *happy_breakdown);
  int AdvanceOneStep(const Vec &v);
  int GetDky(PetscReal t, int deg, Vec p_vec);
  int EstimateCost_(PetscReal tau_new, PetscInt m_new, PetscReal *cost);
};
}  // namespace pacmensl

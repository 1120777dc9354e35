// pacmensl_all.h -- umbrella header (mirrors src/pacmensl_all.h).
#pragma once
#include "CvodeFsp.h"
#include "DiscreteDistribution.h"
#include "FspMatrixBase.h"
#include "FspMatrixConstrained.h"
#include "FspSolverMultiSinks.h"
#include "KrylovFsp.h"
#include "Model.h"
#include "OdeSolverBase.h"
#include "PetscWrap.h"
#include "StateSetBase.h"
#include "StateSetConstrained.h"
#include "Sys.h"

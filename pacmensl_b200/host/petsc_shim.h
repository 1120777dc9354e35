// petsc_shim.h -- the subset of the MPI / PETSc surface that appears in the reference's operator and
// solver signatures, re-implemented on top of the C ABI of include/fsp_b200.h.
//
// A `Vec` here is a DEVICE-resident fp64 vector (one contiguous block per rank, sinks on the last
// rank); the functions below have the PETSc names and argument meaning the reference's tests and
// examples use (tests/test_mat.cpp:118-150, src/Fsp/FspSolverMultiSinks.cpp:368-379,628-636), so those
// call sites read the same.  None of PETSc's code is used or needed.
#pragma once

#include <cstddef>
#include <cstdio>
#include <memory>
#include <vector>

#include "fsp_b200.h"

#define PACMENSL_API __attribute__((visibility("default")))

// ---- scalar types (PETSc real = double, PetscInt = 32-bit int as the reference requires) ------------
typedef double PetscReal;
typedef double PetscScalar;
typedef int    PetscInt;
typedef int    PetscMPIInt;
typedef int    PetscErrorCode;
typedef double PetscLogDouble;
typedef int    PetscLogEvent;
typedef enum { PETSC_FALSE = 0, PETSC_TRUE = 1 } PetscBool;
#define PETSC_DECIDE (-1)
#define PETSC_DETERMINE (-1)
#define PETSC_NULL nullptr
#define PETSC_IGNORE nullptr

// ---- communicator: one process per GPU; collectives go through fspcomm_* (NCCL) ---------------------
struct pacmensl_comm_s {
  int       rank = 0;
  int       size = 1;
  fspcomm_t nccl = nullptr;  // null when size == 1
  void     *stream = nullptr;
};
typedef pacmensl_comm_s *MPI_Comm;
extern "C" {
#define MPI_COMM_NULL ((MPI_Comm) nullptr)
PACMENSL_API MPI_Comm pacmensl_comm_world();
PACMENSL_API MPI_Comm pacmensl_comm_self();
#define MPI_COMM_WORLD (pacmensl_comm_world())
#define PETSC_COMM_WORLD (pacmensl_comm_world())
#define MPI_COMM_SELF (pacmensl_comm_self())
#define PETSC_COMM_SELF (pacmensl_comm_self())
PACMENSL_API int MPI_Comm_rank(MPI_Comm comm, int *rank);
PACMENSL_API int MPI_Comm_size(MPI_Comm comm, int *size);
PACMENSL_API int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out);
PACMENSL_API int MPI_Comm_free(MPI_Comm *comm);
PACMENSL_API int MPI_Barrier(MPI_Comm comm);
// Join an NCCL communicator to the world (multi-GPU): id from fspcomm_unique_id on rank 0.
PACMENSL_API int pacmensl_comm_world_init(const char *nccl_id, int rank, int size);
PACMENSL_API int pacmensl_comm_world_finalize();
// small host-side collectives over the world (device staging + NCCL); identity when size == 1
PACMENSL_API int pacmensl_allreduce_sum(MPI_Comm comm, double *vals_host, int n);
PACMENSL_API int pacmensl_allreduce_max(MPI_Comm comm, double *vals_host, int n);
}  // extern "C"

// ---- Vec ---------------------------------------------------------------------------------------------
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3 } NormType;
typedef enum { INSERT_VALUES = 1, ADD_VALUES = 2 } InsertMode;
typedef const char *VecType;
#define VECMPI "mpi"
#define VECSEQ "seq"
#define VECCUDA "cuda"

struct _p_Vec {
  MPI_Comm comm = nullptr;
  PetscInt n_local = -1, n_global = -1, own_start = 0;
  double  *d_data = nullptr;   // device storage
  bool     owns_data = true;
  std::shared_ptr<void> slab;       // set when d_data points into a block shared with sibling vectors (VecDuplicateVecs)
  double  *placed_saved = nullptr;  // VecPlaceArray bookkeeping
  std::vector<double> host_mirror;  // VecGetArray staging
  int      mirror_mode = 0;         // 0 none, 1 read-only, 2 read-write
  std::vector<std::pair<PetscInt, double>> pending;  // VecSetValues staging until VecAssemblyEnd
  InsertMode pending_mode = INSERT_VALUES;
};
typedef _p_Vec *Vec;
struct _p_PetscRandom { unsigned long long state = 0x853c49e6748fea9bULL; };
typedef _p_PetscRandom *PetscRandom;
#define PETSCRAND "rand"

extern "C" {
PACMENSL_API PetscErrorCode VecCreate(MPI_Comm comm, Vec *v);
PACMENSL_API PetscErrorCode VecSetSizes(Vec v, PetscInt n_local, PetscInt n_global);
PACMENSL_API PetscErrorCode VecSetType(Vec v, VecType type);
PACMENSL_API PetscErrorCode VecSetFromOptions(Vec v);
PACMENSL_API PetscErrorCode VecSetUp(Vec v);
PACMENSL_API PetscErrorCode VecDestroy(Vec *v);
PACMENSL_API PetscErrorCode VecDuplicate(Vec v, Vec *out);
// extension: same layout, contents undefined (solver workspaces that are always overwritten before being read)
PACMENSL_API PetscErrorCode VecDuplicateUninitialized(Vec v, Vec *out);
// PETSc's VecDuplicateVecs/VecDestroyVecs: m vectors carved from ONE device block (one allocation instead of m: on a
// cold pool a block per vector costs ~25 ms per 0.8 GB).  The Uninitialized variant (extension) skips the zero-fill.
PACMENSL_API PetscErrorCode VecDuplicateVecs(Vec v, PetscInt m, Vec **V);
PACMENSL_API PetscErrorCode VecDuplicateVecsUninitialized(Vec v, PetscInt m, Vec **V);
PACMENSL_API PetscErrorCode VecDestroyVecs(PetscInt m, Vec **V);
PACMENSL_API PetscErrorCode VecSet(Vec v, PetscScalar alpha);
PACMENSL_API PetscErrorCode VecSetValue(Vec v, PetscInt row, PetscScalar value, InsertMode mode);
PACMENSL_API PetscErrorCode VecSetValues(Vec v, PetscInt ni, const PetscInt *ix, const PetscScalar *y, InsertMode mode);
PACMENSL_API PetscErrorCode VecAssemblyBegin(Vec v);
PACMENSL_API PetscErrorCode VecAssemblyEnd(Vec v);
PACMENSL_API PetscErrorCode VecCopy(Vec x, Vec y);
PACMENSL_API PetscErrorCode VecSwap(Vec x, Vec y);
PACMENSL_API PetscErrorCode VecSum(Vec v, PetscScalar *sum);
PACMENSL_API PetscErrorCode VecNorm(Vec v, NormType type, PetscReal *val);
PACMENSL_API PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *val);
PACMENSL_API PetscErrorCode VecAXPY(Vec y, PetscScalar alpha, Vec x);
PACMENSL_API PetscErrorCode VecAYPX(Vec y, PetscScalar beta, Vec x);
PACMENSL_API PetscErrorCode VecWAXPY(Vec w, PetscScalar alpha, Vec x, Vec y);
PACMENSL_API PetscErrorCode VecMAXPY(Vec y, PetscInt nv, const PetscScalar alpha[], Vec x[]);
PACMENSL_API PetscErrorCode VecScale(Vec v, PetscScalar alpha);
PACMENSL_API PetscErrorCode VecGetSize(Vec v, PetscInt *n);
PACMENSL_API PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n);
PACMENSL_API PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt *low, PetscInt *high);
PACMENSL_API PetscErrorCode VecGetArray(Vec v, PetscScalar **a);            // host mirror (D2H); Restore writes back
PACMENSL_API PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a);
PACMENSL_API PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a);  // host mirror (D2H)
PACMENSL_API PetscErrorCode VecRestoreArrayRead(Vec v, const PetscScalar **a);
PACMENSL_API PetscErrorCode VecCreateMPIWithArray(MPI_Comm comm, PetscInt bs, PetscInt n, PetscInt N,
                                                  const PetscScalar array_dev[], Vec *v);  // DEVICE array
PACMENSL_API PetscErrorCode VecPlaceArray(Vec v, const PetscScalar array_dev[]);
PACMENSL_API PetscErrorCode VecResetArray(Vec v);
PACMENSL_API PetscErrorCode VecSetRandom(Vec v, PetscRandom r);
PACMENSL_API PetscErrorCode PetscRandomCreate(MPI_Comm comm, PetscRandom *r);
PACMENSL_API PetscErrorCode PetscRandomSetType(PetscRandom r, const char *type);
PACMENSL_API PetscErrorCode PetscRandomDestroy(PetscRandom *r);
// device access (extension; the analogue of VecCUDAGetArray)
PACMENSL_API PetscErrorCode VecGetDeviceArray(Vec v, PetscScalar **a_dev);
PACMENSL_API PetscErrorCode VecGetDeviceArrayRead(Vec v, const PetscScalar **a_dev);

PACMENSL_API PetscErrorCode PetscPrintf(MPI_Comm comm, const char *fmt, ...);
PACMENSL_API PetscErrorCode PetscTime(PetscLogDouble *t);
}  // extern "C"

#define CHKERRQ(ierr) do { if ((ierr) != 0) return (ierr); } while (0)
#define CHKERRMPI(ierr) CHKERRQ(ierr)
#define CHKERRABORT(comm, ierr) do { if ((ierr) != 0) { std::printf("fatal error %d at %s:%d\n", (int) (ierr), __FILE__, __LINE__); std::abort(); } } while (0)

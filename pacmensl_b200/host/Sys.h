// Sys.h -- error conventions, enums and process bootstrap of the host mirror.
// Mirrors src/Sys/ErrorHandling.h:29-54 (integer codes propagated by PACMENSLCHKERRQ, which prints
// function/line/file/rank; PACMENSLCHKERRTHROW raises std::runtime_error) and src/Sys/Sys.h:62-80
// (PACMENSLInit/Finalize, Environment).  MPI/PETSc/Zoltan bootstrapping is replaced by: pick the GPU of
// this rank (LOCAL_RANK) and, for WORLD_SIZE > 1, join the NCCL world communicator.
#pragma once

#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "arma_shim.h"
#include "petsc_shim.h"

using PacmenslErrorCode = int;

#define PACMENSLCHKERRQ(ierr)                                                                              \
  {                                                                                                        \
    if ((ierr) != 0) {                                                                                     \
      int rank_;                                                                                           \
      MPI_Comm_rank(MPI_COMM_WORLD, &rank_);                                                               \
      printf("PACMENSL Error: function %s line %d file %s on rank %d \n.", __func__, __LINE__, __FILE__, rank_); \
      return ierr;                                                                                         \
    }                                                                                                      \
  }

#define PACMENSLCHKERRTHROW(ierr)                                                                          \
  {                                                                                                        \
    if ((ierr) != 0) {                                                                                     \
      int rank_;                                                                                           \
      MPI_Comm_rank(MPI_COMM_WORLD, &rank_);                                                               \
      std::ostringstream msg_;                                                                             \
      msg_ << "PACMENSL Error: Line " << __LINE__ << " in " << __FILE__ << " funcion " << __func__ << "rank " << rank_ << "."; \
      throw std::runtime_error(msg_.str());                                                                \
    }                                                                                                      \
  }

// C-ABI (CUDA library) call: record the library's message and propagate -1
#define FSPCHKERRQ(call)                                                                    \
  {                                                                                         \
    int fsp_ierr_ = (call);                                                                 \
    if (fsp_ierr_ != 0) {                                                                   \
      printf("PACMENSL device error: %s (%s:%d)\n", fsp_last_error(), __FILE__, __LINE__);  \
      return fsp_ierr_;                                                                     \
    }                                                                                       \
  }

#define NOT_COPYABLE_NOT_MOVABLE(object) \
  object(const object &) = delete;       \
  object &operator=(const object &) = delete;

namespace pacmensl {

enum class PartitioningType { BLOCK, GRAPH, HYPERGRAPH, HIERARCHICAL };
enum class PartitioningApproach { FROMSCRATCH, REPARTITION, REFINE };
PACMENSL_API PartitioningType str2part(std::string str);
PACMENSL_API std::string part2str(PartitioningType part);
PACMENSL_API PartitioningApproach str2partapproach(std::string str);
PACMENSL_API std::string partapproach2str(PartitioningApproach part_approach);

PACMENSL_API double round2digit(double x);

PACMENSL_API int PACMENSLInit(int *argc, char ***argv, const char *help);
PACMENSL_API int PACMENSLFinalize();

class PACMENSL_API Environment {
 public:
  Environment();
  Environment(int *argc, char ***argv, const char *help);
  ~Environment();

 private:
  bool initialized = false;
};

// RAII device buffer used by the host classes
template <typename T>
class DeviceBuffer {
 public:
  DeviceBuffer() {}
  explicit DeviceBuffer(size_t n) { resize(n); }
  ~DeviceBuffer() { release(); }
  DeviceBuffer(const DeviceBuffer &) = delete;
  DeviceBuffer &operator=(const DeviceBuffer &) = delete;
  int resize(size_t n) {
    if (n <= cap_) { n_ = n; return 0; }
    release();
    void *p = nullptr;
    int   ierr = fsp_malloc(&p, sizeof(T) * (n ? n : 1));
    if (ierr) return ierr;
    ptr_ = static_cast<T *>(p);
    n_ = cap_ = n;
    return 0;
  }
  void release() { if (ptr_) fsp_free(ptr_); ptr_ = nullptr; n_ = cap_ = 0; }
  T *get() const { return ptr_; }
  size_t size() const { return n_; }
  void swap(DeviceBuffer &o) { std::swap(ptr_, o.ptr_); std::swap(n_, o.n_); std::swap(cap_, o.cap_); }
  int upload(const T *host, size_t n) { int e = resize(n); if (e) return e; return n ? fsp_memcpy_h2d(ptr_, host, sizeof(T) * n, nullptr) : 0; }
  int download(T *host, size_t n) const { return n ? fsp_memcpy_d2h(host, ptr_, sizeof(T) * n, nullptr) : 0; }

 private:
  T     *ptr_ = nullptr;
  size_t n_ = 0, cap_ = 0;
};

}  // namespace pacmensl

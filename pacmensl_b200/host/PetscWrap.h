// PetscWrap.h -- RAII handle and ExpandVec (mirrors src/PetscWrap/PetscWrap.h:12-60).
#pragma once

#include "Sys.h"

namespace pacmensl {

template <typename PetscT>
class Petsc {
 protected:
  PetscT dat = nullptr;

 public:
  Petsc() {}
  Petsc(const Petsc &) = delete;
  Petsc &operator=(const Petsc &) = delete;
  Petsc(Petsc &&o) noexcept : dat(o.dat) { o.dat = nullptr; }
  Petsc &operator=(Petsc &&o) noexcept {
    if (this != &o) { release(); dat = o.dat; o.dat = nullptr; }
    return *this;
  }
  PetscT *mem() { return &dat; }
  const PetscT *mem() const { return &dat; }
  bool IsEmpty() { return (dat == nullptr); }
  operator PetscT() { return dat; }
  void release();
  ~Petsc() { release(); }
};
template <> inline void Petsc<Vec>::release() { if (dat) VecDestroy(&dat); dat = nullptr; }

/// p_new = 0; p_new[new_indices[i]] = p_old[i]; p <- p_new  (src/PetscWrap/PetscWrap.cpp:26-56).
/// new_indices are GLOBAL positions of the local entries of p in the enlarged vector.
PACMENSL_API PacmenslErrorCode ExpandVec(Vec &p, const std::vector<PetscInt> &new_indices, const PetscInt new_local_size);
PACMENSL_API PacmenslErrorCode ExpandVec(Petsc<Vec> &p, const std::vector<PetscInt> &new_indices,
                                         const PetscInt new_local_size);
}  // namespace pacmensl

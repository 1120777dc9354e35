#include "PetscWrap.h"

namespace pacmensl {

PacmenslErrorCode ExpandVec(Vec &p, const std::vector<PetscInt> &new_indices, const PetscInt new_local_size) {
  PacmenslErrorCode ierr{0};
  MPI_Comm          comm = p->comm;
  int               size = comm ? comm->size : 1;
  if ((PetscInt) new_indices.size() != p->n_local) PACMENSLCHKERRQ(-1);

  Vec Pnew;
  ierr = VecCreate(comm, &Pnew); PACMENSLCHKERRQ(ierr);
  ierr = VecSetSizes(Pnew, new_local_size, PETSC_DECIDE); PACMENSLCHKERRQ(ierr);
  ierr = VecSetUp(Pnew); PACMENSLCHKERRQ(ierr);

  void *stream = comm ? comm->stream : nullptr;
  if (size == 1) {
    DeviceBuffer<int> idx;
    if (idx.upload(new_indices.data(), new_indices.size())) PACMENSLCHKERRQ(-1);
    FSPCHKERRQ(fspvec_scatter(Pnew->d_data, new_local_size, p->d_data, idx.get(), p->n_local, stream));
    FSPCHKERRQ(fsp_stream_sync(stream));
  } else {
    // Multi-GPU: gather (global index, value) pairs of every rank, keep those landing in the own block.
    // All ranks contribute p->n_local entries; pad to the maximum for the fixed-size all-gather.
    double nmax = (double) p->n_local;
    ierr = pacmensl_allreduce_max(comm, &nmax, 1); PACMENSLCHKERRQ(ierr);
    const long pad = (long) nmax;
    std::vector<int> idx_pad((size_t) pad, -1);
    for (size_t i = 0; i < new_indices.size(); ++i) idx_pad[i] = new_indices[i];
    DeviceBuffer<int>    idx_loc, idx_all;
    DeviceBuffer<double> val_loc((size_t) pad), val_all((size_t) pad * size);
    if (idx_loc.upload(idx_pad.data(), (size_t) pad) || idx_all.resize((size_t) pad * size)) PACMENSLCHKERRQ(-1);
    FSPCHKERRQ(fspvec_set(val_loc.get(), 0.0, pad, stream));
    FSPCHKERRQ(fspvec_copy(val_loc.get(), p->d_data, p->n_local, stream));
    FSPCHKERRQ(fspcomm_allgather_int(comm->nccl, idx_loc.get(), idx_all.get(), pad, stream));
    FSPCHKERRQ(fspcomm_allgather_f64(comm->nccl, val_loc.get(), val_all.get(), pad, stream));
    // shift to local positions of the new block; entries outside it become negative / too large -> dropped
    FSPCHKERRQ(fspvec_scatter_range(Pnew->d_data, new_local_size, val_all.get(), idx_all.get(), pad * size,
                                    Pnew->own_start, stream));
    FSPCHKERRQ(fsp_stream_sync(stream));
  }
  ierr = VecDestroy(&p); PACMENSLCHKERRQ(ierr);
  p = Pnew;
  return 0;
}

PacmenslErrorCode ExpandVec(Petsc<Vec> &p, const std::vector<PetscInt> &new_indices, const PetscInt new_local_size) {
  return ExpandVec(*p.mem(), new_indices, new_local_size);
}
}  // namespace pacmensl

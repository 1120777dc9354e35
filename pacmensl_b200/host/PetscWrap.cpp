#include "PetscWrap.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

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
  } else if (const char *mode = std::getenv("FSP_EXPANDVEC"); mode && !std::strcmp(mode, "allgather")) {
    // Round-1 form, kept for A/B: gather (global index, value) pairs of every rank, keep those landing in the own block
    // (transient memory and traffic proportional to the GLOBAL size on every rank).
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
    FSPCHKERRQ(fspvec_scatter_range(Pnew->d_data, new_local_size, val_all.get(), idx_all.get(), pad * size,
                                    Pnew->own_start, stream));
    FSPCHKERRQ(fsp_stream_sync(stream));
  } else {
    // Multi-GPU: every entry travels once, to the rank that owns its new global index (the VecScatter of
    // src/Sys/PetscWrap.cpp:10-45).  The entries are sorted by owner on the device, the per-peer counts are exchanged, and
    // the (index, value) segments are stored straight into the owners' receive windows over NVLink (fspcomm_alltoallv);
    // transient memory is proportional to the LOCAL size.
    const long n_old = p->n_local;
    std::vector<long> sizes((size_t) size, 0), starts((size_t) size + 1, 0);
    FSPCHKERRQ(fspcomm_gather_long(comm->nccl, (long) new_local_size, sizes.data()));
    for (int r = 0; r < size; ++r) starts[(size_t) r + 1] = starts[(size_t) r] + sizes[(size_t) r];
    if (starts[(size_t) comm->rank] != Pnew->own_start) PACMENSLCHKERRQ(-1);
    DeviceBuffer<int>    idx_loc, idx_sorted((size_t) std::max<long>(n_old, 1));
    DeviceBuffer<double> val_sorted((size_t) std::max<long>(n_old, 1));
    if (!idx_sorted.get() || !val_sorted.get()) PACMENSLCHKERRQ(-1);
    if (n_old > 0 && idx_loc.upload(new_indices.data(), (size_t) n_old)) PACMENSLCHKERRQ(-1);
    std::vector<long> send_counts((size_t) size, 0), recv_counts((size_t) size, 0);
    FSPCHKERRQ(fspvec_route_by_owner(idx_loc.get(), p->d_data, n_old, starts.data(), size, idx_sorted.get(), val_sorted.get(),
                                     send_counts.data(), stream));
    FSPCHKERRQ(fspcomm_alltoall_counts(comm->nccl, send_counts.data(), recv_counts.data(), stream));
    long n_recv = 0;
    for (long c : recv_counts) n_recv += c;
    DeviceBuffer<int>    idx_recv((size_t) std::max<long>(n_recv, 1));
    DeviceBuffer<double> val_recv((size_t) std::max<long>(n_recv, 1));
    if (!idx_recv.get() || !val_recv.get()) PACMENSLCHKERRQ(-1);
    FSPCHKERRQ(fspcomm_alltoallv(comm->nccl, idx_sorted.get(), send_counts.data(), idx_recv.get(), recv_counts.data(), (int) sizeof(int), stream));
    FSPCHKERRQ(fspcomm_alltoallv(comm->nccl, val_sorted.get(), send_counts.data(), val_recv.get(), recv_counts.data(), (int) sizeof(double), stream));
    FSPCHKERRQ(fspvec_set(Pnew->d_data, 0.0, new_local_size, stream));
    if (n_recv > 0)
      FSPCHKERRQ(fspvec_scatter_range(Pnew->d_data, new_local_size, val_recv.get(), idx_recv.get(), n_recv, Pnew->own_start, stream));
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

// petsc_shim.cpp -- device-resident Vec and the communicator shim (see petsc_shim.h).
#include "petsc_shim.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>

namespace {
pacmensl_comm_s g_world;
pacmensl_comm_s g_self;
double         *g_scratch_dev = nullptr;  // small device staging buffer for host-side collectives
constexpr int   kScratch = 4096;

int ensure_scratch() {
  if (!g_scratch_dev) return fsp_malloc((void **) &g_scratch_dev, sizeof(double) * kScratch);
  return 0;
}
}  // namespace

MPI_Comm pacmensl_comm_world() { return &g_world; }
MPI_Comm pacmensl_comm_self() { return &g_self; }

int MPI_Comm_rank(MPI_Comm comm, int *rank) { *rank = comm ? comm->rank : 0; return 0; }
int MPI_Comm_size(MPI_Comm comm, int *size) { *size = comm ? comm->size : 1; return 0; }
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out) { *out = comm; return 0; }  // communicators are shared handles
int MPI_Comm_free(MPI_Comm *comm) { *comm = MPI_COMM_NULL; return 0; }

int pacmensl_comm_world_init(const char *nccl_id, int rank, int size) {
  if (g_world.nccl) return 0;
  g_world.rank = rank;
  g_world.size = size;
  if (size > 1) {
    int ierr = fspcomm_create(&g_world.nccl, nccl_id, rank, size);
    if (ierr) return ierr;
  }
  return 0;
}

int pacmensl_comm_world_finalize() {
  if (g_world.nccl) fspcomm_destroy(g_world.nccl);
  g_world = pacmensl_comm_s();
  if (g_scratch_dev) { fsp_free(g_scratch_dev); g_scratch_dev = nullptr; }
  return 0;
}

static int host_allreduce(MPI_Comm comm, double *v, int n, bool is_max) {
  if (!comm || comm->size == 1 || n == 0) return 0;
  if (n > kScratch) return -1;
  if (ensure_scratch()) return -1;
  int ierr = fsp_memcpy_h2d(g_scratch_dev, v, sizeof(double) * n, comm->stream);
  if (ierr) return ierr;
  ierr = is_max ? fspcomm_allreduce_max(comm->nccl, g_scratch_dev, n, comm->stream)
                : fspcomm_allreduce_sum(comm->nccl, g_scratch_dev, n, comm->stream);
  if (ierr) return ierr;
  ierr = fsp_memcpy_d2h(v, g_scratch_dev, sizeof(double) * n, comm->stream);
  // a device-side flag wait that timed out poisoned the result: report it where the host consumes it
  return ierr ? ierr : fspcomm_check(comm->nccl);
}
int pacmensl_allreduce_sum(MPI_Comm comm, double *v, int n) { return host_allreduce(comm, v, n, false); }
int pacmensl_allreduce_max(MPI_Comm comm, double *v, int n) { return host_allreduce(comm, v, n, true); }
int MPI_Barrier(MPI_Comm comm) {
  double one = 1.0;
  return pacmensl_allreduce_sum(comm, &one, 1);
}

// ---- Vec -----------------------------------------------------------------------------------------------
static void *S(Vec v) { return v->comm ? v->comm->stream : nullptr; }

PetscErrorCode VecCreate(MPI_Comm comm, Vec *v) {
  *v = new _p_Vec();
  (*v)->comm = comm;
  return 0;
}
PetscErrorCode VecSetSizes(Vec v, PetscInt n_local, PetscInt n_global) {
  v->n_local = n_local;
  v->n_global = n_global;
  return 0;
}
PetscErrorCode VecSetType(Vec, VecType) { return 0; }
PetscErrorCode VecSetFromOptions(Vec) { return 0; }

static PetscErrorCode vec_layout(Vec v) {
  int size = v->comm ? v->comm->size : 1, rank = v->comm ? v->comm->rank : 0;
  if (v->n_local < 0) {  // PETSC_DECIDE local size: PETSc's default split
    if (v->n_global < 0) return -1;
    v->n_local = v->n_global / size + ((v->n_global % size) > rank ? 1 : 0);
  }
  if (size == 1) {
    v->n_global = v->n_local;
    v->own_start = 0;
    return 0;
  }
  std::vector<double> counts(size, 0.0);
  counts[rank] = (double) v->n_local;
  int ierr = pacmensl_allreduce_sum(v->comm, counts.data(), size);
  if (ierr) return ierr;
  double tot = 0.0, start = 0.0;
  for (int r = 0; r < size; ++r) {
    if (r < rank) start += counts[r];
    tot += counts[r];
  }
  v->n_global = (PetscInt) tot;
  v->own_start = (PetscInt) start;
  return 0;
}

PetscErrorCode VecSetUp(Vec v) {
  if (v->d_data) return 0;
  PetscErrorCode ierr = vec_layout(v);
  if (ierr) return ierr;
  ierr = fsp_malloc((void **) &v->d_data, sizeof(double) * (size_t) (v->n_local > 0 ? v->n_local : 1));
  if (ierr) return ierr;
  v->owns_data = true;
  return fspvec_set(v->d_data, 0.0, v->n_local, S(v));
}

PetscErrorCode VecDestroy(Vec *v) {
  if (!v || !*v) return 0;
  if ((*v)->owns_data && (*v)->d_data) fsp_free((*v)->d_data);
  delete *v;
  *v = nullptr;
  return 0;
}

static PetscErrorCode vec_duplicate(Vec v, Vec *out, bool zero) {
  if (!v->d_data && v->n_local >= 0) { PetscErrorCode e = VecSetUp(v); if (e) return e; }
  Vec w = new _p_Vec();
  w->comm = v->comm;
  w->n_local = v->n_local;
  w->n_global = v->n_global;
  w->own_start = v->own_start;
  PetscErrorCode ierr = fsp_malloc((void **) &w->d_data, sizeof(double) * (size_t) (w->n_local > 0 ? w->n_local : 1));
  if (ierr) { delete w; return ierr; }
  if (zero) ierr = fspvec_set(w->d_data, 0.0, w->n_local, S(w));
  *out = w;
  return ierr;
}
static PetscErrorCode vec_duplicate_vecs(Vec v, PetscInt m, Vec **V, bool zero) {
  *V = nullptr;
  if (m <= 0) return 0;
  if (!v->d_data && v->n_local >= 0) { PetscErrorCode e = VecSetUp(v); if (e) return e; }
  const size_t n = (size_t) (v->n_local > 0 ? v->n_local : 1);
  const size_t stride = (n + 31) / 32 * 32;  // every vector starts 256-byte aligned
  double *block = nullptr;
  PetscErrorCode ierr = fsp_malloc((void **) &block, sizeof(double) * stride * (size_t) m);
  if (ierr) return ierr;
  std::shared_ptr<void> slab(block, [](void *p) { fsp_free(p); });
  if (zero) { ierr = fspvec_set(block, 0.0, (long) (stride * (size_t) m), S(v)); if (ierr) return ierr; }
  Vec *arr = new Vec[(size_t) m];
  for (PetscInt i = 0; i < m; ++i) {
    Vec w = new _p_Vec();
    w->comm = v->comm;
    w->n_local = v->n_local;
    w->n_global = v->n_global;
    w->own_start = v->own_start;
    w->d_data = block + stride * (size_t) i;
    w->owns_data = false;
    w->slab = slab;
    arr[i] = w;
  }
  *V = arr;
  return 0;
}
PetscErrorCode VecDuplicateVecs(Vec v, PetscInt m, Vec **V) { return vec_duplicate_vecs(v, m, V, true); }
PetscErrorCode VecDuplicateVecsUninitialized(Vec v, PetscInt m, Vec **V) { return vec_duplicate_vecs(v, m, V, false); }
PetscErrorCode VecDestroyVecs(PetscInt m, Vec **V) {
  if (!V || !*V) return 0;
  for (PetscInt i = 0; i < m; ++i) VecDestroy(&(*V)[i]);
  delete[] *V;
  *V = nullptr;
  return 0;
}
PetscErrorCode VecDuplicate(Vec v, Vec *out) { return vec_duplicate(v, out, true); }
PetscErrorCode VecDuplicateUninitialized(Vec v, Vec *out) { return vec_duplicate(v, out, false); }

PetscErrorCode VecSet(Vec v, PetscScalar a) {
  if (!v->d_data) { PetscErrorCode e = VecSetUp(v); if (e) return e; }
  return fspvec_set(v->d_data, a, v->n_local, S(v));
}

PetscErrorCode VecSetValues(Vec v, PetscInt ni, const PetscInt *ix, const PetscScalar *y, InsertMode mode) {
  if (!v->d_data) { PetscErrorCode e = VecSetUp(v); if (e) return e; }
  v->pending_mode = mode;
  for (PetscInt k = 0; k < ni; ++k)
    if (ix[k] >= 0) v->pending.emplace_back(ix[k], y[k]);  // negative indices are ignored, as in PETSc
  return 0;
}
PetscErrorCode VecSetValue(Vec v, PetscInt row, PetscScalar value, InsertMode mode) {
  return VecSetValues(v, 1, &row, &value, mode);
}
PetscErrorCode VecAssemblyBegin(Vec) { return 0; }
// Gather variable-length (index, value) lists from all ranks (rank order); identity when size == 1.
static int gather_pairs(MPI_Comm comm, std::vector<int> &idx, std::vector<double> &val) {
  if (!comm || comm->size == 1) return 0;
  const int size = comm->size, rank = comm->rank;
  std::vector<double> counts(size, 0.0);
  counts[rank] = (double) idx.size();
  int ierr = pacmensl_allreduce_sum(comm, counts.data(), size);
  if (ierr) return ierr;
  long pad = 0, total = 0;
  for (double c : counts) { pad = std::max(pad, (long) c); total += (long) c; }
  if (total == 0) return 0;
  int    *d_is = nullptr, *d_ia = nullptr;
  double *d_vs = nullptr, *d_va = nullptr;
  if (fsp_malloc((void **) &d_is, sizeof(int) * pad) || fsp_malloc((void **) &d_ia, sizeof(int) * pad * size) ||
      fsp_malloc((void **) &d_vs, sizeof(double) * pad) || fsp_malloc((void **) &d_va, sizeof(double) * pad * size)) return -1;
  std::vector<int>    is((size_t) pad, -1), ia((size_t) pad * size);
  std::vector<double> vs((size_t) pad, 0.0), va((size_t) pad * size);
  std::copy(idx.begin(), idx.end(), is.begin());
  std::copy(val.begin(), val.end(), vs.begin());
  ierr = fsp_memcpy_h2d(d_is, is.data(), sizeof(int) * pad, comm->stream) || fsp_memcpy_h2d(d_vs, vs.data(), sizeof(double) * pad, comm->stream);
  if (!ierr) ierr = fspcomm_allgather_int(comm->nccl, d_is, d_ia, pad, comm->stream);
  if (!ierr) ierr = fspcomm_allgather_f64(comm->nccl, d_vs, d_va, pad, comm->stream);
  if (!ierr) ierr = fsp_memcpy_d2h(ia.data(), d_ia, sizeof(int) * pad * size, comm->stream) || fsp_memcpy_d2h(va.data(), d_va, sizeof(double) * pad * size, comm->stream);
  fsp_free(d_is); fsp_free(d_ia); fsp_free(d_vs); fsp_free(d_va);
  if (ierr) return ierr;
  idx.clear();
  val.clear();
  for (int r = 0; r < size; ++r)
    for (long k = 0; k < (long) counts[r]; ++k) { idx.push_back(ia[(size_t) r * pad + k]); val.push_back(va[(size_t) r * pad + k]); }
  return 0;
}

PetscErrorCode VecAssemblyEnd(Vec v) {
  // Entries are given by GLOBAL index and may belong to another rank (PETSc communicates them at assembly time):
  // the staged (index, value) pairs of all ranks are gathered in rank order and every rank applies the ones it owns.
  // Collective when the communicator has more than one rank.
  std::vector<int>    idx;
  std::vector<double> val;
  for (auto &pr : v->pending) { idx.push_back(pr.first); val.push_back(pr.second); }
  v->pending.clear();
  int ierr = gather_pairs(v->comm, idx, val);
  if (ierr) return ierr;
  for (size_t k = 0; k < idx.size(); ++k) {
    PetscInt loc = idx[k] - v->own_start;
    if (loc < 0 || loc >= v->n_local) continue;
    double value = val[k];
    if (v->pending_mode == ADD_VALUES) {
      double cur;
      ierr = fsp_memcpy_d2h(&cur, v->d_data + loc, sizeof(double), S(v));
      if (ierr) return ierr;
      value += cur;
    }
    ierr = fsp_memcpy_h2d(v->d_data + loc, &value, sizeof(double), S(v));
    if (ierr) return ierr;
  }
  return 0;
}

PetscErrorCode VecCopy(Vec x, Vec y) {
  if (x->n_local != y->n_local) return -1;
  return fspvec_copy(y->d_data, x->d_data, x->n_local, S(x));
}
PetscErrorCode VecSwap(Vec x, Vec y) {
  if (x->n_local != y->n_local) return -1;
  std::swap(x->d_data, y->d_data);
  std::swap(x->owns_data, y->owns_data);
  return 0;
}

static PetscErrorCode reduce_finish(Vec v, double *val) { return pacmensl_allreduce_sum(v->comm, val, 1); }

PetscErrorCode VecSum(Vec v, PetscScalar *sum) {
  int ierr = fspvec_sum_h(sum, v->d_data, v->n_local, S(v));
  if (ierr) return ierr;
  return reduce_finish(v, sum);
}
PetscErrorCode VecNorm(Vec v, NormType type, PetscReal *val) {
  int ierr;
  if (type == NORM_1) {
    ierr = fspvec_norm1_h(val, v->d_data, v->n_local, S(v));
    if (ierr) return ierr;
    return reduce_finish(v, val);
  }
  if (type == NORM_2 || type == NORM_FROBENIUS) {
    double nrm;
    ierr = fspvec_norm2_h(&nrm, v->d_data, v->n_local, S(v));
    if (ierr) return ierr;
    double sq = nrm * nrm;
    if (v->comm && v->comm->size > 1) {
      ierr = reduce_finish(v, &sq);
      if (ierr) return ierr;
      nrm = std::sqrt(sq);
    }
    *val = nrm;
    return 0;
  }
  // NORM_INFINITY through the host mirror (not on the hot path)
  const PetscScalar *a;
  ierr = VecGetArrayRead(v, &a);
  if (ierr) return ierr;
  double m = 0.0;
  for (PetscInt i = 0; i < v->n_local; ++i) m = std::fmax(m, std::fabs(a[i]));
  VecRestoreArrayRead(v, &a);
  ierr = pacmensl_allreduce_max(v->comm, &m, 1);
  *val = m;
  return ierr;
}
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *val) {
  int ierr = fspvec_dot_h(val, x->d_data, y->d_data, x->n_local, S(x));
  if (ierr) return ierr;
  return reduce_finish(x, val);
}
PetscErrorCode VecAXPY(Vec y, PetscScalar alpha, Vec x) { return fspvec_axpy(y->d_data, alpha, x->d_data, y->n_local, S(y)); }
PetscErrorCode VecAYPX(Vec y, PetscScalar beta, Vec x) {
  return fspvec_linear_sum(y->d_data, 1.0, x->d_data, beta, y->d_data, y->n_local, S(y));
}
PetscErrorCode VecWAXPY(Vec w, PetscScalar alpha, Vec x, Vec y) {
  return fspvec_linear_sum(w->d_data, alpha, x->d_data, 1.0, y->d_data, w->n_local, S(w));
}
PetscErrorCode VecMAXPY(Vec y, PetscInt nv, const PetscScalar alpha[], Vec x[]) {
  // y += sum alpha_k x_k, in chunks of 64 vectors (one fused pass each)
  for (PetscInt k0 = 0; k0 < nv; k0 += 64) {
    PetscInt      m = nv - k0 < 64 ? nv - k0 : 64;
    const double *ptrs[64];
    for (PetscInt k = 0; k < m; ++k) ptrs[k] = x[k0 + k]->d_data;
    int ierr = fspvec_maxpy(y->d_data, 1.0, m, alpha + k0, ptrs, y->n_local, S(y));
    if (ierr) return ierr;
  }
  return 0;
}
PetscErrorCode VecScale(Vec v, PetscScalar alpha) {
  if (alpha == 0.0) return fspvec_set(v->d_data, 0.0, v->n_local, S(v));
  return fspvec_scale(v->d_data, alpha, v->n_local, S(v));
}
PetscErrorCode VecGetSize(Vec v, PetscInt *n) {
  if (!v->d_data) { PetscErrorCode e = VecSetUp(v); if (e) return e; }
  *n = v->n_global;
  return 0;
}
PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n) { *n = v->n_local; return 0; }
PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt *low, PetscInt *high) {
  if (!v->d_data) { PetscErrorCode e = VecSetUp(v); if (e) return e; }
  if (low) *low = v->own_start;
  if (high) *high = v->own_start + v->n_local;
  return 0;
}

PetscErrorCode VecGetArray(Vec v, PetscScalar **a) {
  v->host_mirror.resize((size_t) (v->n_local > 0 ? v->n_local : 1));
  int ierr = v->n_local > 0 ? fsp_memcpy_d2h(v->host_mirror.data(), v->d_data, sizeof(double) * v->n_local, S(v)) : 0;
  if (!ierr && v->comm && v->comm->nccl) ierr = fspcomm_check(v->comm->nccl);
  v->mirror_mode = 2;
  *a = v->host_mirror.data();
  return ierr;
}
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a) {
  int ierr = 0;
  if (v->mirror_mode == 2 && v->n_local > 0)
    ierr = fsp_memcpy_h2d(v->d_data, v->host_mirror.data(), sizeof(double) * v->n_local, S(v));
  v->mirror_mode = 0;
  if (a) *a = nullptr;
  return ierr;
}
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a) {
  v->host_mirror.resize((size_t) (v->n_local > 0 ? v->n_local : 1));
  int ierr = v->n_local > 0 ? fsp_memcpy_d2h(v->host_mirror.data(), v->d_data, sizeof(double) * v->n_local, S(v)) : 0;
  if (!ierr && v->comm && v->comm->nccl) ierr = fspcomm_check(v->comm->nccl);
  v->mirror_mode = 1;
  *a = v->host_mirror.data();
  return ierr;
}
PetscErrorCode VecRestoreArrayRead(Vec v, const PetscScalar **a) {
  v->mirror_mode = 0;
  if (a) *a = nullptr;
  return 0;
}

PetscErrorCode VecCreateMPIWithArray(MPI_Comm comm, PetscInt, PetscInt n, PetscInt N, const PetscScalar array_dev[], Vec *v) {
  Vec w = new _p_Vec();
  w->comm = comm;
  w->n_local = n;
  w->n_global = N;
  PetscErrorCode ierr = vec_layout(w);
  w->d_data = const_cast<double *>(array_dev);
  w->owns_data = false;
  *v = w;
  return ierr;
}
PetscErrorCode VecPlaceArray(Vec v, const PetscScalar array_dev[]) {
  v->placed_saved = v->d_data;
  v->d_data = const_cast<double *>(array_dev);
  return 0;
}
PetscErrorCode VecResetArray(Vec v) {
  v->d_data = v->placed_saved;
  v->placed_saved = nullptr;
  return 0;
}

PetscErrorCode PetscRandomCreate(MPI_Comm comm, PetscRandom *r) {
  *r = new _p_PetscRandom();
  (*r)->state += 0x9E3779B97F4A7C15ULL * (unsigned long long) (comm ? comm->rank + 1 : 1);
  return 0;
}
PetscErrorCode PetscRandomSetType(PetscRandom, const char *) { return 0; }
PetscErrorCode PetscRandomDestroy(PetscRandom *r) { delete *r; *r = nullptr; return 0; }
PetscErrorCode VecSetRandom(Vec v, PetscRandom r) {
  _p_PetscRandom local;
  _p_PetscRandom *g = r ? r : &local;
  std::vector<double> h((size_t) (v->n_local > 0 ? v->n_local : 1));
  for (PetscInt i = 0; i < v->n_local; ++i) {  // splitmix64 -> U[0,1)
    unsigned long long z = (g->state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    h[i] = (double) (z >> 11) * (1.0 / 9007199254740992.0);
  }
  return v->n_local > 0 ? fsp_memcpy_h2d(v->d_data, h.data(), sizeof(double) * v->n_local, S(v)) : 0;
}

PetscErrorCode VecGetDeviceArray(Vec v, PetscScalar **a) { *a = v->d_data; return 0; }
PetscErrorCode VecGetDeviceArrayRead(Vec v, const PetscScalar **a) { *a = v->d_data; return 0; }

PetscErrorCode PetscPrintf(MPI_Comm comm, const char *fmt, ...) {
  if (comm && comm->rank != 0) return 0;
  va_list ap;
  va_start(ap, fmt);
  std::vprintf(fmt, ap);
  va_end(ap);
  return 0;
}
PetscErrorCode PetscTime(PetscLogDouble *t) {
  *t = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  return 0;
}

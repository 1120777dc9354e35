// fspcomm.cu -- multi-GPU plumbing, one process per GPU (include/fsp_b200.h "Multi-GPU plumbing").
//
// Replaces the MPI traffic of the reference's hot path: the ghost VecScatter inside MatMult on
// MATMPISELL, the K-element sink VecScatter ADD (src/Matrix/FspMatrixConstrained.cpp:57-60) and the
// MPI_Allreduce behind every VecDot/VecNorm (src/OdeSolver/KrylovFsp.cpp:280-309).
// NCCL is bound at run time with dlopen so that a process which already loaded a libnccl.so.2 (e.g.
// PyTorch's bundled copy) shares it instead of pulling in a second, different NCCL.
#include <dlfcn.h>
#include <string.h>

#include "fsp_common.cuh"

using namespace fspb;

namespace {

// Minimal NCCL ABI (stable across 2.x): opaque comm, 128-byte unique id, enums below.
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt32 = 2, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Reduce)(const void *, void *, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.lib) return 0;
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { set_error("fspcomm: cannot load libnccl.so.2: %s", dlerror()); return -1; }
#define FSP_SYM(field, name)                                                        \
  *(void **) (&g_nccl.field) = dlsym(lib, name);                                    \
  if (!g_nccl.field) { set_error("fspcomm: symbol %s missing in NCCL", name); return -1; }
  FSP_SYM(GetUniqueId, "ncclGetUniqueId")
  FSP_SYM(CommInitRank, "ncclCommInitRank")
  FSP_SYM(CommDestroy, "ncclCommDestroy")
  FSP_SYM(AllReduce, "ncclAllReduce")
  FSP_SYM(Reduce, "ncclReduce")
  FSP_SYM(AllGather, "ncclAllGather")
  FSP_SYM(Send, "ncclSend")
  FSP_SYM(Recv, "ncclRecv")
  FSP_SYM(GroupStart, "ncclGroupStart")
  FSP_SYM(GroupEnd, "ncclGroupEnd")
  FSP_SYM(GetErrorString, "ncclGetErrorString")
#undef FSP_SYM
  g_nccl.lib = lib;
  return 0;
}

#define FSP_NCCL_CHECK(expr)                                                                    \
  do {                                                                                          \
    int _r = (expr);                                                                            \
    if (_r != ncclSuccess) {                                                                    \
      set_error("%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(_r), __FILE__, __LINE__); \
      return -1;                                                                                \
    }                                                                                           \
  } while (0)

}  // namespace

struct fspcomm_s {
  ncclComm_t comm = nullptr;
  int        rank = 0, size = 1;
};

extern "C" {

int fspcomm_unique_id(char id[FSPCOMM_ID_BYTES]) {
  if (load_nccl()) return -1;
  ncclUniqueId uid;
  FSP_NCCL_CHECK(g_nccl.GetUniqueId(&uid));
  memcpy(id, uid.internal, FSPCOMM_ID_BYTES);
  return 0;
}

int fspcomm_create(fspcomm_t *out, const char id[FSPCOMM_ID_BYTES], int rank, int size) {
  fspcomm_s *c = new fspcomm_s();
  c->rank = rank; c->size = size;
  if (size > 1) {
    if (load_nccl()) { delete c; return -1; }
    ncclUniqueId uid;
    memcpy(uid.internal, id, FSPCOMM_ID_BYTES);
    int r = g_nccl.CommInitRank(&c->comm, size, uid, rank);
    if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); delete c; return -1; }
  }
  *out = c;
  return 0;
}

int fspcomm_destroy(fspcomm_t c) {
  if (!c) return 0;
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
  return 0;
}

int fspcomm_rank(fspcomm_t c, int *rank, int *size) {
  *rank = c ? c->rank : 0;
  *size = c ? c->size : 1;
  return 0;
}

int fspcomm_allreduce_sum(fspcomm_t c, double *buf, long n, void *stream) {
  if (!c || c->size == 1 || n <= 0) return 0;
  FSP_NCCL_CHECK(g_nccl.AllReduce(buf, buf, (size_t) n, ncclFloat64, ncclSum, c->comm, resolve_stream(stream)));
  return 0;
}
int fspcomm_allreduce_max(fspcomm_t c, double *buf, long n, void *stream) {
  if (!c || c->size == 1 || n <= 0) return 0;
  FSP_NCCL_CHECK(g_nccl.AllReduce(buf, buf, (size_t) n, ncclFloat64, ncclMax, c->comm, resolve_stream(stream)));
  return 0;
}
int fspcomm_reduce_sum(fspcomm_t c, double *buf, long n, int root, void *stream) {
  if (!c || c->size == 1 || n <= 0) return 0;
  FSP_NCCL_CHECK(g_nccl.Reduce(buf, buf, (size_t) n, ncclFloat64, ncclSum, root, c->comm, resolve_stream(stream)));
  return 0;
}
int fspcomm_allgather_f64(fspcomm_t c, const double *send, double *recv, long n_per_rank, void *stream) {
  if (!c || c->size == 1) {
    if (send != recv) FSP_CUDA_CHECK(cudaMemcpyAsync(recv, send, sizeof(double) * n_per_rank, cudaMemcpyDeviceToDevice, resolve_stream(stream)));
    return 0;
  }
  FSP_NCCL_CHECK(g_nccl.AllGather(send, recv, (size_t) n_per_rank, ncclFloat64, c->comm, resolve_stream(stream)));
  return 0;
}
int fspcomm_allgather_int(fspcomm_t c, const int *send, int *recv, long n_per_rank, void *stream) {
  if (!c || c->size == 1) {
    if (send != recv) FSP_CUDA_CHECK(cudaMemcpyAsync(recv, send, sizeof(int) * n_per_rank, cudaMemcpyDeviceToDevice, resolve_stream(stream)));
    return 0;
  }
  FSP_NCCL_CHECK(g_nccl.AllGather(send, recv, (size_t) n_per_rank, ncclInt32, c->comm, resolve_stream(stream)));
  return 0;
}

int fspcomm_halo_exchange(fspcomm_t c, const double *send, const long *send_counts, double *ghost,
                          const long *recv_counts, void *stream) {
  if (!c || c->size == 1) return 0;
  cudaStream_t st = resolve_stream(stream);
  FSP_NCCL_CHECK(g_nccl.GroupStart());
  long so = 0, ro = 0;
  for (int p = 0; p < c->size; ++p) {
    if (send_counts[p] > 0) FSP_NCCL_CHECK(g_nccl.Send(send + so, (size_t) send_counts[p], ncclFloat64, p, c->comm, st));
    if (recv_counts[p] > 0) FSP_NCCL_CHECK(g_nccl.Recv(ghost + ro, (size_t) recv_counts[p], ncclFloat64, p, c->comm, st));
    so += send_counts[p];
    ro += recv_counts[p];
  }
  FSP_NCCL_CHECK(g_nccl.GroupEnd());
  return 0;
}

int fspcomm_alltoall_counts(fspcomm_t c, const long *send_host, long *recv_host, void *stream) {
  if (!c || c->size == 1) { recv_host[0] = send_host[0]; return 0; }
  const int    P = c->size;
  cudaStream_t st = resolve_stream(stream);
  double      *d_send = nullptr, *d_all = nullptr;
  FSP_CUDA_CHECK(pmalloc(&d_send, sizeof(double) * P));
  FSP_CUDA_CHECK(pmalloc(&d_all, sizeof(double) * P * P));
  double hs[256], hall[256 * 8];
  if (P > 256 || P * P > 2048) { set_error("fspcomm_alltoall_counts: too many ranks"); return -1; }
  for (int p = 0; p < P; ++p) hs[p] = (double) send_host[p];
  FSP_CUDA_CHECK(cudaMemcpyAsync(d_send, hs, sizeof(double) * P, cudaMemcpyHostToDevice, st));
  FSP_NCCL_CHECK(g_nccl.AllGather(d_send, d_all, (size_t) P, ncclFloat64, c->comm, st));
  FSP_CUDA_CHECK(cudaMemcpyAsync(hall, d_all, sizeof(double) * P * P, cudaMemcpyDeviceToHost, st));
  FSP_CUDA_CHECK(cudaStreamSynchronize(st));
  for (int p = 0; p < P; ++p) recv_host[p] = (long) hall[p * P + c->rank];
  pfree(d_send); pfree(d_all);
  return 0;
}

int fspcomm_exchange_int(fspcomm_t c, const int *send, const long *send_counts, int *recv, const long *recv_counts,
                         void *stream) {
  if (!c || c->size == 1) return 0;
  cudaStream_t st = resolve_stream(stream);
  FSP_NCCL_CHECK(g_nccl.GroupStart());
  long so = 0, ro = 0;
  for (int p = 0; p < c->size; ++p) {
    if (send_counts[p] > 0) FSP_NCCL_CHECK(g_nccl.Send(send + so, (size_t) send_counts[p], ncclInt32, p, c->comm, st));
    if (recv_counts[p] > 0) FSP_NCCL_CHECK(g_nccl.Recv(recv + ro, (size_t) recv_counts[p], ncclInt32, p, c->comm, st));
    so += send_counts[p];
    ro += recv_counts[p];
  }
  FSP_NCCL_CHECK(g_nccl.GroupEnd());
  FSP_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"

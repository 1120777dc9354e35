// fspcomm.cu -- multi-GPU plumbing, one process per GPU (include/fsp_b200.h "Multi-GPU plumbing").
//
// Replaces the MPI traffic of the reference's hot path: the ghost VecScatter inside MatMult on
// MATMPISELL, the K-element sink VecScatter ADD (src/Matrix/FspMatrixConstrained.cpp:57-60) and the
// MPI_Allreduce behind every VecDot/VecNorm (src/OdeSolver/KrylovFsp.cpp:280-309).
// NCCL is bound at run time with dlopen so that a process which already loaded a libnccl.so.2 (e.g.
// PyTorch's bundled copy) shares it instead of pulling in a second, different NCCL.
//
// Peer-memory fast path (one node, NVLink 5 / NVSwitch): every rank maps a window of every peer's HBM through
// CUDA IPC.  The halo exchange is then ONE kernel (pack the boundary entries of x and store them straight into the
// peers' ghost buffers, then publish an epoch flag with release semantics at system scope) and the consumer kernels
// wait on the flags in device code; small all-reduces (Krylov/GMRES inner products) are ONE kernel that writes
// this rank's values into every peer's slot and sums the slots in rank order (deterministic, bit-identical on all
// ranks).  No NCCL call, no host synchronisation, on the per-Action / per-inner-product path.  NCCL remains the
// bootstrap (handle exchange) and the path for FSP_P2P=0 or GPUs without peer access.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <set>
#include <vector>

#include "fsp_common.cuh"

using namespace fspb;

namespace {

// Minimal NCCL ABI (stable across 2.x): opaque comm, 128-byte unique id, enums below.
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt32 = 2, ncclInt64 = 4, ncclFloat64 = 8 };
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

// Live communicators: objects that outlive their communicator (a state set or an operator destroyed after
// PACMENSLFinalize / pfsp_finalize, e.g. by a garbage collector at interpreter exit) must not touch it.
namespace {
std::mutex                     g_live_mutex;
std::set<const void *>         g_live_comms;
void comm_register(const void *c) { std::lock_guard<std::mutex> g(g_live_mutex); g_live_comms.insert(c); }
void comm_unregister(const void *c) { std::lock_guard<std::mutex> g(g_live_mutex); g_live_comms.erase(c); }
bool comm_alive(const void *c) { std::lock_guard<std::mutex> g(g_live_mutex); return c && g_live_comms.count(c) != 0; }
}  // namespace

// ---- peer-memory windows ---------------------------------------------------------------------------
constexpr int kMaxRanks = FSP_P2P_MAX_RANKS;
constexpr int kMaxRedVals = FSP_P2P_MAX_REDUCE;
constexpr int kRedDepth = 4;  // ring of reduction slots (calls are stream-ordered; 2 would do)

struct PeerWindow {
  size_t bytes = 0;
  void  *local = nullptr;
  void  *peer[kMaxRanks] = {nullptr};  // peer[rank] == local
};

// control window of a communicator: flags + slots of the fused small all-reduce
struct CtrlLayout {
  unsigned long long flags[kRedDepth][kMaxRanks];
  double             slots[kRedDepth][kMaxRanks][kMaxRedVals];
};

// halo window of one operator: epoch flags, sink slots, double-buffered ghost entries
struct HaloHeader {
  unsigned long long halo_flags[2][kMaxRanks];
  unsigned long long sink_flags[2][kMaxRanks];
  double             sink_slots[2][kMaxRanks][FSP_P2P_MAX_SINKS];
};

struct fspcomm_s {
  ncclComm_t comm = nullptr;
  int        rank = 0, size = 1;
  // peer-memory state
  bool               p2p = false;
  PeerWindow         ctrl;
  unsigned long long red_epoch = 0;
  unsigned int      *err_host = nullptr;  // pinned + mapped: device code sets it when a flag wait times out
  unsigned int      *err_dev = nullptr;
  double            *stage = nullptr;     // small device staging buffer of the set-up collectives (handles, agreement flags)
  struct PooledHalo { PeerWindow win; size_t cap = 0; unsigned long long epoch = 0; };
  std::vector<PooledHalo> halo_pool;  // windows returned by fsphalo_destroy, identical order/capacity on all ranks
  std::vector<PeerWindow> retired;    // general windows given back without a collective (fspcomm_window_retire)
  double            *coll = nullptr;  // device scratch of fspcomm_barrier / fspcomm_gather_long (kMaxRanks + 1 doubles)
  PeerWindow         a2a;             // receive window of fspcomm_alltoallv (grows by doubling)
};

struct fsphalo_s {
  fspcomm_s         *c = nullptr;
  fspcomm_s::PooledHalo w;
  long               n_send = 0, n_ghost = 0;
  int                n_sink = 0;
  int               *send_idx = nullptr;  // borrowed device pointer
  long               send_off[kMaxRanks + 1] = {0};
  long               remote_off[kMaxRanks] = {0};  // where my segment starts in peer p's ghost buffer
  unsigned          *block_counter = nullptr;
};

namespace {

// ONE kernel = pack + all-to-all over NVLink + signal (push_role, fsp_common.cuh)
__global__ void __launch_bounds__(256) halo_push_kernel(PushView a, const double *__restrict__ x) {
  push_role(a, x, (int) blockIdx.x);
}

struct ReduceArgs {
  int                 size, rank, n, slot;
  unsigned long long  epoch;
  double             *slots[kMaxRanks];  // peer p's slot row [slot][rank][*] for my rank
  unsigned long long *flag[kMaxRanks];   // peer p's flag [slot][rank]
  const double             *my_slots;    // local [slot][0][0]
  const unsigned long long *my_flags;    // local [slot][0]
  unsigned int       *err;
};

// Fused small all-reduce (n <= kMaxRedVals): write my values into every peer's slot, publish, wait for all peers,
// sum (or max) in rank order.  One CTA.
template <bool IS_MAX>
__global__ void __launch_bounds__(128) p2p_allreduce_kernel(ReduceArgs a, double *__restrict__ buf) {
  const int t = threadIdx.x;
  if (t < a.n) {
    const double v = buf[t];
    for (int p = 0; p < a.size; ++p) a.slots[p][t] = v;
  }
  __threadfence_system();
  __syncthreads();
  bool ok = true;
  if (t < a.size) {
    st_release_sys(a.flag[t], a.epoch);
    ok = wait_flag(a.my_flags + t, a.epoch, a.err);
  }
  ok = __syncthreads_and(ok);
  if (t < a.n) {
    double s = __ldcg(a.my_slots + t);
    for (int p = 1; p < a.size; ++p) {
      const double v = __ldcg(a.my_slots + (size_t) p * kMaxRedVals + t);
      s = IS_MAX ? fmax(s, v) : s + v;
    }
    // a peer that never arrived: poison the result instead of leaving a partial sum in place (the host reports the
    // error at its next synchronisation: fspcomm_check / check_peer_error)
    buf[t] = ok ? s : __longlong_as_double(0x7ff8000000000000ll);
  }
}

constexpr size_t kStageBytes = 64 * (size_t) (kMaxRanks + 2);

// Agreement step of the collective set-up paths: returns 0 when EVERY rank passed ok == true, 1 when some rank failed
// (all ranks then take the same fall-back), -1 when the collective itself failed.  Every rank-local fallible step of a
// collective set-up is followed by one of these before its result is committed, so ranks can never disagree on
// whether a window / the peer-memory path exists (a disagreement ends in 20 s spin time-outs or an NCCL hang).
int agree(fspcomm_s *c, bool ok) {
  double v = ok ? 0.0 : 1.0;
  if (cudaMemcpy(c->stage, &v, 8, cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); v = 1.0; }
  if (g_nccl.AllReduce(c->stage, c->stage, 1, ncclFloat64, ncclSum, c->comm, (cudaStream_t) 0) != ncclSuccess) return -1;
  if (cudaMemcpy(&v, c->stage, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return -1; }
  return v == 0.0 ? 0 : 1;
}

void window_destroy(fspcomm_s *c, PeerWindow *w);

// collective: allocate + zero the local part, all-gather the IPC handles over NCCL, map every peer.  Returns 0 on ALL
// ranks or -1 on ALL ranks (nothing left allocated or mapped): local failures are agreed on before returning, and every
// rank takes part in every collective call whatever happened to it locally.
int window_create(fspcomm_s *c, size_t bytes, PeerWindow *w) {
  *w = PeerWindow();
  w->bytes = bytes;
  bool ok = cudaMalloc(&w->local, bytes) == cudaSuccess;
  if (!ok) w->local = nullptr;
  ok = ok && cudaMemset(w->local, 0, bytes) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  ok = ok && cudaIpcGetMemHandle(&mine, w->local) == cudaSuccess;
  double *d_send = c->stage + 8, *d_all = c->stage + 16;  // 64 bytes + 64 bytes per rank
  ok = (cudaMemcpy(d_send, &mine, 64, cudaMemcpyHostToDevice) == cudaSuccess) && ok;
  const bool coll_ok = g_nccl.AllGather(d_send, d_all, 8, ncclFloat64, c->comm, (cudaStream_t) 0) == ncclSuccess;
  std::vector<cudaIpcMemHandle_t> all((size_t) c->size);
  ok = ok && coll_ok && cudaMemcpy(all.data(), d_all, 64 * (size_t) c->size, cudaMemcpyDeviceToHost) == cudaSuccess;
  // every rank must have produced a handle before anybody maps anything
  int agreed = coll_ok ? agree(c, ok) : -1;
  if (agreed == 0) {
    for (int p = 0; p < c->size && ok; ++p) {
      if (p == c->rank) { w->peer[p] = w->local; continue; }
      cudaError_t e = cudaIpcOpenMemHandle(&w->peer[p], all[(size_t) p], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        set_error("cudaIpcOpenMemHandle(peer %d) failed: %s", p, cudaGetErrorString(e));
        w->peer[p] = nullptr;
        ok = false;
      }
    }
    agreed = agree(c, ok);
  }
  cudaGetLastError();
  if (agreed != 0) {
    window_destroy(c, w);
    if (agreed > 0 && ok) set_error("window_create: another rank could not create or map its peer-memory window");
    return -1;
  }
  return 0;
}

void window_destroy(fspcomm_s *c, PeerWindow *w) {
  for (int p = 0; p < c->size; ++p)
    if (p != c->rank && w->peer[p]) cudaIpcCloseMemHandle(w->peer[p]);
  if (w->local) cudaFree(w->local);
  *w = PeerWindow();
}

int check_peer_error(fspcomm_s *c, const char *where) {
  if (c->err_host && *(volatile unsigned int *) c->err_host) {
    set_error("%s: a peer-memory flag wait timed out (a rank is missing or failed)", where);
    return -1;
  }
  return 0;
}


int p2p_allreduce(fspcomm_s *c, double *buf, int n, bool is_max, cudaStream_t st) {
  if (check_peer_error(c, "fspcomm_allreduce")) return -1;
  const unsigned long long epoch = ++c->red_epoch;
  const int slot = (int) (epoch % kRedDepth);
  ReduceArgs a;
  a.size = c->size; a.rank = c->rank; a.n = n; a.slot = slot; a.epoch = epoch;
  for (int p = 0; p < c->size; ++p) {
    CtrlLayout *L = reinterpret_cast<CtrlLayout *>(c->ctrl.peer[p]);
    a.slots[p] = &L->slots[slot][c->rank][0];
    a.flag[p] = &L->flags[slot][c->rank];
  }
  CtrlLayout *me = reinterpret_cast<CtrlLayout *>(c->ctrl.local);
  a.my_slots = &me->slots[slot][0][0];
  a.my_flags = &me->flags[slot][0];
  a.err = c->err_dev;
  if (is_max) p2p_allreduce_kernel<true><<<1, 128, 0, st>>>(a, buf);
  else p2p_allreduce_kernel<false><<<1, 128, 0, st>>>(a, buf);
  FSP_LAUNCH_CHECK();
  return 0;
}

// collective; leaves c->p2p == false (NCCL path) on ALL ranks if any rank cannot map any peer, cannot set up its
// error flag, or FSP_P2P=0.  Returns -1 only when a collective call itself failed.
int p2p_setup(fspcomm_s *c) {
  const char *env = getenv("FSP_P2P");
  bool want = !(env && !strcmp(env, "0")) && c->size <= kMaxRanks;
  // every rank must want it (the environment could differ between ranks)
  int agreed = agree(c, want);
  if (agreed < 0) return -1;
  c->p2p = false;
  if (agreed != 0) return 0;
  if (window_create(c, sizeof(CtrlLayout), &c->ctrl)) return 0;  // consistent on all ranks
  bool ok = cudaHostAlloc((void **) &c->err_host, 2 * sizeof(unsigned int), cudaHostAllocMapped) == cudaSuccess;
  if (ok) {
    const char *tmo = getenv("FSP_SPIN_TIMEOUT_MS");  // device-side flag waits give up after this long (default 20 s)
    c->err_host[0] = 0u;
    c->err_host[1] = tmo ? (unsigned) atoi(tmo) : 0u;
    ok = cudaHostGetDevicePointer((void **) &c->err_dev, c->err_host, 0) == cudaSuccess;
  } else {
    c->err_host = nullptr;
  }
  cudaGetLastError();
  agreed = agree(c, ok);
  if (agreed != 0) {
    window_destroy(c, &c->ctrl);
    if (c->err_host) cudaFreeHost(c->err_host);
    c->err_host = nullptr; c->err_dev = nullptr;
    return agreed < 0 ? -1 : 0;
  }
  c->p2p = true;
  return 0;
}

}  // namespace

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
    // staging buffer of the set-up collectives; if even this fails the device is unusable (the peers cannot be told
    // through NCCL without device memory): fatal
    if (cudaMalloc(&c->stage, kStageBytes) != cudaSuccess) { set_error("fspcomm_create: cannot allocate the staging buffer"); g_nccl.CommDestroy(c->comm); delete c; return -1; }
    // peer-memory fast path when every rank can map every peer; NCCL path otherwise
    if (p2p_setup(c)) { set_error("fspcomm_create: the peer-memory set-up collective failed"); cudaFree(c->stage); g_nccl.CommDestroy(c->comm); delete c; return -1; }
  }
  comm_register(c);
  *out = c;
  return 0;
}

int fspcomm_alive(fspcomm_t c) { return comm_alive(c) ? 1 : 0; }

int fspcomm_destroy(fspcomm_t c) {
  if (!c) return 0;
  comm_unregister(c);
  if (c->p2p) {
    // every rank must be past its last peer store before any window is unmapped
    cudaDeviceSynchronize();
    double *d = nullptr;
    if (cudaMalloc(&d, 8) == cudaSuccess) {
      cudaMemset(d, 0, 8);
      g_nccl.AllReduce(d, d, 1, ncclFloat64, ncclSum, c->comm, (cudaStream_t) 0);
      cudaDeviceSynchronize();
      cudaFree(d);
    }
    for (auto &w : c->halo_pool) window_destroy(c, &w.win);
    c->halo_pool.clear();
    for (auto &w : c->retired) window_destroy(c, &w);
    c->retired.clear();
    if (c->a2a.local) window_destroy(c, &c->a2a);
    if (c->coll) cudaFree(c->coll);
    c->coll = nullptr;
    window_destroy(c, &c->ctrl);
    if (c->err_host) cudaFreeHost(c->err_host);
    c->p2p = false;
  }
  if (c->stage) cudaFree(c->stage);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
  return 0;
}

int fspcomm_rank(fspcomm_t c, int *rank, int *size) {
  *rank = c ? c->rank : 0;
  *size = c ? c->size : 1;
  return 0;
}

int fspcomm_p2p_enabled(fspcomm_t c) { return (c && c->p2p) ? 1 : 0; }

int fspcomm_allreduce_sum(fspcomm_t c, double *buf, long n, void *stream) {
  if (!c || c->size == 1 || n <= 0) return 0;
  if (c->p2p && n <= kMaxRedVals) return p2p_allreduce(c, buf, (int) n, false, resolve_stream(stream));
  FSP_NCCL_CHECK(g_nccl.AllReduce(buf, buf, (size_t) n, ncclFloat64, ncclSum, c->comm, resolve_stream(stream)));
  return 0;
}
int fspcomm_allreduce_max(fspcomm_t c, double *buf, long n, void *stream) {
  if (!c || c->size == 1 || n <= 0) return 0;
  if (c->p2p && n <= kMaxRedVals) return p2p_allreduce(c, buf, (int) n, true, resolve_stream(stream));
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
    if (p == c->rank) {  // the own segment never touches the network
      if (send_counts[p] > 0 && send_counts[p] == recv_counts[p])
        FSP_CUDA_CHECK(cudaMemcpyAsync(ghost + ro, send + so, sizeof(double) * (size_t) send_counts[p], cudaMemcpyDeviceToDevice, st));
    } else {
      if (send_counts[p] > 0) FSP_NCCL_CHECK(g_nccl.Send(send + so, (size_t) send_counts[p], ncclFloat64, p, c->comm, st));
      if (recv_counts[p] > 0) FSP_NCCL_CHECK(g_nccl.Recv(ghost + ro, (size_t) recv_counts[p], ncclFloat64, p, c->comm, st));
    }
    so += send_counts[p];
    ro += recv_counts[p];
  }
  FSP_NCCL_CHECK(g_nccl.GroupEnd());
  return 0;
}

int fspcomm_alltoall_counts(fspcomm_t c, const long *send_host, long *recv_host, void *stream) {
  if (!c || c->size == 1) { recv_host[0] = send_host[0]; return 0; }
  const int P = c->size;
  if (P > 256) { set_error("fspcomm_alltoall_counts: too many ranks"); return -1; }
  static_assert(sizeof(long) == 8, "counts travel as int64");
  cudaStream_t st = resolve_stream(stream);
  long        *d_send = nullptr, *d_all = nullptr;
  int          rc = -1;
  std::vector<long> hall((size_t) P * P);
  do {
    if (pmalloc(&d_send, sizeof(long) * P) != cudaSuccess || pmalloc(&d_all, sizeof(long) * P * P) != cudaSuccess) {
      set_error("fspcomm_alltoall_counts: allocation failed");
      break;
    }
    if (cudaMemcpyAsync(d_send, send_host, sizeof(long) * P, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
    if (g_nccl.AllGather(d_send, d_all, (size_t) P, ncclInt64, c->comm, st) != ncclSuccess) { set_error("fspcomm_alltoall_counts: ncclAllGather failed"); break; }
    if (cudaMemcpyAsync(hall.data(), d_all, sizeof(long) * P * P, cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
    if (cudaStreamSynchronize(st) != cudaSuccess) break;
    for (int p = 0; p < P; ++p) recv_host[p] = hall[(size_t) p * P + c->rank];
    rc = 0;
  } while (0);
  if (rc) cudaGetLastError();
  pfree(d_send); pfree(d_all);
  return rc;
}

int fspcomm_exchange_int(fspcomm_t c, const int *send, const long *send_counts, int *recv, const long *recv_counts,
                         void *stream) {
  if (!c || c->size == 1) return 0;
  cudaStream_t st = resolve_stream(stream);
  FSP_NCCL_CHECK(g_nccl.GroupStart());
  long so = 0, ro = 0;
  for (int p = 0; p < c->size; ++p) {
    if (p == c->rank) {
      if (send_counts[p] > 0 && send_counts[p] == recv_counts[p])
        FSP_CUDA_CHECK(cudaMemcpyAsync(recv + ro, send + so, sizeof(int) * (size_t) send_counts[p], cudaMemcpyDeviceToDevice, st));
    } else {
      if (send_counts[p] > 0) FSP_NCCL_CHECK(g_nccl.Send(send + so, (size_t) send_counts[p], ncclInt32, p, c->comm, st));
      if (recv_counts[p] > 0) FSP_NCCL_CHECK(g_nccl.Recv(recv + ro, (size_t) recv_counts[p], ncclInt32, p, c->comm, st));
    }
    so += send_counts[p];
    ro += recv_counts[p];
  }
  FSP_NCCL_CHECK(g_nccl.GroupEnd());
  FSP_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}


// ---- fused halo exchange over peer memory ----------------------------------------------------------
int fsphalo_create(fspcomm_t c, fsphalo_t *out, const int *send_idx_dev, const long *send_counts,
                   const long *recv_counts, int n_sink) {
  *out = nullptr;
  if (!c || !c->p2p) { set_error("fsphalo_create: peer memory is not enabled on this communicator"); return -1; }
  if (n_sink > FSP_P2P_MAX_SINKS) { set_error("fsphalo_create: %d sink rows exceed %d", n_sink, FSP_P2P_MAX_SINKS); return -1; }
  fsphalo_s *h = new fsphalo_s();
  h->c = c; h->n_sink = n_sink; h->send_idx = const_cast<int *>(send_idx_dev);
  long recv_off[kMaxRanks + 1] = {0};
  for (int p = 0; p < c->size; ++p) {
    h->send_off[p + 1] = h->send_off[p] + send_counts[p];
    recv_off[p + 1] = recv_off[p] + recv_counts[p];
  }
  h->n_send = h->send_off[c->size];
  h->n_ghost = recv_off[c->size];
  // Collective from here on: a rank-local failure is carried in `ok` through every collective call and agreed on before
  // the pool (which must stay identical on all ranks) is touched, so either every rank gets a halo or none does.
  // every peer learns where its segment starts in my ghost buffer
  bool ok = fspcomm_alltoall_counts(c, recv_off, h->remote_off, nullptr) == 0;
  // window capacity: the maximum ghost count over ranks (all ranks must take the same pool decision)
  double need = (double) h->n_ghost;
  ok = (cudaMemcpy(c->stage, &need, 8, cudaMemcpyHostToDevice) == cudaSuccess) && ok;
  const bool coll_ok = g_nccl.AllReduce(c->stage, c->stage, 1, ncclFloat64, ncclMax, c->comm, (cudaStream_t) 0) == ncclSuccess;
  ok = ok && coll_ok && cudaMemcpy(&need, c->stage, 8, cudaMemcpyDeviceToHost) == cudaSuccess;
  ok = (cudaMalloc(&h->block_counter, sizeof(unsigned)) == cudaSuccess) && ok;
  if (h->block_counter) ok = (cudaMemset(h->block_counter, 0, sizeof(unsigned)) == cudaSuccess) && ok;
  cudaGetLastError();
  int agreed = coll_ok ? agree(c, ok) : -1;
  if (agreed == 0) {
    const size_t cap_need = (size_t) need;
    bool found = false;
    for (size_t q = 0; q < c->halo_pool.size(); ++q) {
      if (c->halo_pool[q].cap >= cap_need) {
        h->w = c->halo_pool[q];
        c->halo_pool.erase(c->halo_pool.begin() + (long) q);
        found = true;
        break;
      }
    }
    if (!found) {
      // generous growth: a new window costs a cudaMalloc, an IPC handle exchange and one cudaIpcOpenMemHandle per peer
      // (tens of milliseconds), and the adaptive FSP regenerates its operator dozens of times with a growing halo
      size_t cap = ((std::max<size_t>(2 * cap_need, (size_t) 1 << 20)) + 31) / 32 * 32;  // >= 16 MB per window: never re-created for small sets
      // consistent on all ranks (window_create agrees internally); the pool is untouched when it fails
      if (window_create(c, sizeof(HaloHeader) + 2 * cap * sizeof(double), &h->w.win)) agreed = 1;
      h->w.cap = cap;
      h->w.epoch = 0;
    }
  }
  if (agreed != 0) {
    if (ok) set_error("fsphalo_create: set-up failed on a rank of the communicator");
    if (h->block_counter) cudaFree(h->block_counter);
    delete h;
    return -1;
  }
  *out = h;
  return 0;
}

int fsphalo_destroy(fsphalo_t h) {
  if (!h) return 0;
  if (!comm_alive(h->c)) {  // the communicator (and with it the window) is gone already
    if (h->block_counter) cudaFree(h->block_counter);
    cudaGetLastError();
    delete h;
    return 0;
  }
  // the window (with its epoch counter: the flags stay monotone) goes back to the pool; the kernels in flight keep
  // using it safely because every later user continues the same epoch sequence
  h->c->halo_pool.push_back(h->w);
  if (h->block_counter) { cudaDeviceSynchronize(); cudaFree(h->block_counter); }
  delete h;
  return 0;
}

// Next epoch of the halo: fills the consumer's view (*out) and the producer's view (*push) without launching anything.
static int halo_next(fsphalo_s *h, fsphalo_epoch *out, PushView *a) {
  fspcomm_s *c = h->c;
  if (check_peer_error(c, "fsphalo")) return -1;
  const unsigned long long epoch = ++h->w.epoch;
  const int par = (int) (epoch & 1ull);
  a->size = c->size; a->rank = c->rank; a->n_send = h->n_send; a->epoch = epoch;
  a->send_idx = h->send_idx;
  a->block_counter = h->block_counter;
  // a few entries per thread: the push CTAs lead the fused action kernel and should be few and short
  a->n_ctas = (int) std::max<long>(1, std::min<long>((h->n_send + 2047) / 2048, 1024));
  for (int p = 0; p <= c->size; ++p) a->send_off[p] = h->send_off[p];
  for (int p = 0; p < c->size; ++p) {
    char       *base = reinterpret_cast<char *>(h->w.win.peer[p]);
    HaloHeader *H = reinterpret_cast<HaloHeader *>(base);
    double     *ghost = reinterpret_cast<double *>(base + sizeof(HaloHeader)) + (size_t) par * h->w.cap;
    a->dst[p] = ghost + h->remote_off[p];
    a->flag[p] = &H->halo_flags[par][c->rank];
  }
  char       *mine = reinterpret_cast<char *>(h->w.win.local);
  HaloHeader *M = reinterpret_cast<HaloHeader *>(mine);
  const int   owner = c->size - 1;
  HaloHeader *O = reinterpret_cast<HaloHeader *>(h->w.win.peer[owner]);
  out->epoch = epoch;
  out->n_ranks = c->size;
  out->self_rank = c->rank;
  out->ghost = reinterpret_cast<double *>(mine + sizeof(HaloHeader)) + (size_t) par * h->w.cap;
  out->halo_flags = &M->halo_flags[par][0];
  out->sink_flags = &M->sink_flags[par][0];
  out->sink_slots = &M->sink_slots[par][0][0];
  out->sink_slot_remote = &O->sink_slots[par][c->rank][0];
  out->sink_flag_remote = &O->sink_flags[par][c->rank];
  out->error_flag = c->err_dev;
  return 0;
}

int fsphalo_begin(fsphalo_t h, const double *x_dev, void *stream, fsphalo_epoch *out) {
  PushView a;
  if (halo_next(h, out, &a)) return -1;
  halo_push_kernel<<<(unsigned) a.n_ctas, 256, 0, resolve_stream(stream)>>>(a, x_dev);
  FSP_LAUNCH_CHECK();
  return 0;
}

int fsphalo_next(fsphalo_t h, fsphalo_epoch *out, fsphalo_push *push) {
  static_assert(sizeof(fsphalo_push) >= sizeof(PushView), "fsphalo_push must be able to hold a PushView");
  PushView a;
  if (halo_next(h, out, &a)) return -1;
  memset(push, 0, sizeof(*push));
  memcpy(push, &a, sizeof(a));
  return 0;
}

int fsphalo_check(fsphalo_t h) { return h ? check_peer_error(h->c, "fsphalo_check") : 0; }

int fspcomm_check(fspcomm_t c) { return (c && c->p2p) ? check_peer_error(c, "fspcomm_check") : 0; }

// ---- general peer-memory windows (sharded state set: fspset.cu) ------------------------------------
// Collective.  peers[p] = this process' mapping of rank p's window (peers[rank] = the local allocation), all of at
// least `bytes` bytes.  A retired window of a fitting size is reused when there is one (identical sequence on all ranks);
// its content is whatever the previous user left.
int fspcomm_window_create(fspcomm_t c, size_t bytes, void **peers) {
  if (!c || !c->p2p) { set_error("fspcomm_window_create: peer memory is not enabled on this communicator"); return -1; }
  // first fit among the pooled windows (the pool has the same content in the same order on every rank); a window up to
  // four times larger than asked for is acceptable, so that the pool is actually reused when sizes drift
  for (size_t i = 0; i < c->retired.size(); ++i)
    if (c->retired[i].bytes >= bytes && c->retired[i].bytes / 4 <= bytes) {
      for (int p = 0; p < c->size; ++p) peers[p] = c->retired[i].peer[p];
      c->retired.erase(c->retired.begin() + (long) i);
      return 0;
    }
  PeerWindow w;
  if (window_create(c, bytes, &w)) return -1;
  for (int p = 0; p < c->size; ++p) peers[p] = w.peer[p];
  return 0;
}

// Collective: every rank has finished using the window (its own part AND the peers' parts).
int fspcomm_window_destroy(fspcomm_t c, void **peers) {
  if (!c || !peers || !peers[c->rank]) return 0;
  cudaDeviceSynchronize();
  int rc = agree(c, true);  // nobody is still reading or writing through a mapping
  PeerWindow w;
  w.local = peers[c->rank];
  for (int p = 0; p < c->size; ++p) { w.peer[p] = peers[p]; if (p != c->rank && peers[p]) cudaIpcCloseMemHandle(peers[p]); w.peer[p] = nullptr; }
  if (agree(c, true) < 0) rc = -1;  // every mapping is closed before the memory goes away
  cudaFree(w.local);
  for (int p = 0; p < c->size; ++p) peers[p] = nullptr;
  cudaGetLastError();
  return rc < 0 ? -1 : 0;
}

// NOT collective: the window goes to the communicator's pool (reused by a later fspcomm_window_create of the same size,
// unmapped and freed by fspcomm_destroy).  For destructors, whose order between ranks is not defined.
int fspcomm_window_retire(fspcomm_t c, void **peers, size_t bytes) {
  if (!comm_alive(c)) return 0;  // destroyed after its communicator: the process is going down, nothing to pool
  if (!peers || !peers[c->rank]) return 0;
  PeerWindow w;
  w.bytes = bytes;
  w.local = peers[c->rank];
  for (int p = 0; p < c->size; ++p) { w.peer[p] = peers[p]; peers[p] = nullptr; }
  c->retired.push_back(w);
  return 0;
}

static int coll_scratch(fspcomm_s *c) {
  if (c->coll) return 0;
  FSP_CUDA_CHECK(cudaMalloc(&c->coll, sizeof(double) * (kMaxRanks + 1)));
  FSP_CUDA_CHECK(cudaMemset(c->coll, 0, sizeof(double) * (kMaxRanks + 1)));
  return 0;
}

// Stream-ordered barrier: work enqueued after it on `stream` starts only when every rank's work enqueued before its own
// barrier call has finished (one small all-reduce kernel over peer memory; NCCL otherwise).
int fspcomm_barrier(fspcomm_t c, void *stream) {
  if (!c || c->size == 1) return 0;
  if (coll_scratch(c)) return -1;
  return fspcomm_allreduce_sum(c, c->coll + kMaxRanks, 1, stream);
}

// Personalised all-to-all of variable-size segments (set-up paths: the ghost id lists of GenerateValues, the routed
// ExpandVec).  send holds the segments for rank 0, 1, ... back to back (send_counts elements of esz = 4 or 8 bytes each),
// recv receives the segments from rank 0, 1, ... back to back.  Peer-memory path: every rank owns a receive window;
// the senders store their segments straight into it over NVLink (peer copies), a flag barrier, one local copy out --
// no NCCL point-to-point channels (whose lazy connection costs ~0.3 s at first use).  FSP_A2A=nccl or a communicator
// without peer memory: grouped ncclSend / ncclRecv.  Collective; synchronises the stream before returning.
int fspcomm_alltoallv(fspcomm_t c, const void *send, const long *send_counts, void *recv, const long *recv_counts, int esz,
                      void *stream) {
  if (esz != 4 && esz != 8) { set_error("fspcomm_alltoallv: element size %d", esz); return -1; }
  cudaStream_t st = resolve_stream(stream);
  if (!c || c->size == 1) {
    if (send_counts[0] > 0 && send != recv)
      FSP_CUDA_CHECK(cudaMemcpyAsync(recv, send, (size_t) esz * (size_t) send_counts[0], cudaMemcpyDeviceToDevice, st));
    FSP_CUDA_CHECK(cudaStreamSynchronize(st));
    return 0;
  }
  static const bool force_nccl = [] { const char *e = getenv("FSP_A2A"); return e && !strcmp(e, "nccl"); }();
  if (!c->p2p || force_nccl) {
    if (esz == 4) return fspcomm_exchange_int(c, (const int *) send, send_counts, (int *) recv, recv_counts, stream);
    if (fspcomm_halo_exchange(c, (const double *) send, send_counts, (double *) recv, recv_counts, stream)) return -1;
    FSP_CUDA_CHECK(cudaStreamSynchronize(st));
    return 0;
  }
  const int P = c->size;
  // where my segment starts inside each peer's receive window (in elements)
  long recv_off[kMaxRanks + 1] = {0}, remote_off[kMaxRanks] = {0};
  for (int p = 0; p < P; ++p) recv_off[p + 1] = recv_off[p] + recv_counts[p];
  if (fspcomm_alltoall_counts(c, recv_off, remote_off, stream)) return -1;
  // one window size for all ranks: the largest receive volume
  long need_all[kMaxRanks];
  if (fspcomm_gather_long(c, recv_off[P] * (long) esz, need_all)) return -1;
  size_t need = 0;
  for (int p = 0; p < P; ++p) need = std::max(need, (size_t) need_all[p]);
  if (need > c->a2a.bytes) {
    size_t bytes = std::max<size_t>(std::max<size_t>(need, 2 * c->a2a.bytes), (size_t) 1 << 20);
    bytes = (bytes + 255) / 256 * 256;
    if (c->a2a.local) {
      cudaDeviceSynchronize();
      if (agree(c, true) < 0) return -1;  // nobody is still writing into or reading from the old window
      window_destroy(c, &c->a2a);
    }
    if (window_create(c, bytes, &c->a2a)) return -1;
  }
  // every rank has copied the previous call's data out of its window (that copy precedes this barrier on its stream)
  if (fspcomm_barrier(c, stream)) return -1;
  long so = 0;
  for (int p = 0; p < P; ++p) {
    if (send_counts[p] > 0)
      FSP_CUDA_CHECK(cudaMemcpyAsync((char *) c->a2a.peer[p] + (size_t) remote_off[p] * esz, (const char *) send + (size_t) so * esz,
                                     (size_t) send_counts[p] * esz, cudaMemcpyDefault, st));
    so += send_counts[p];
  }
  if (fspcomm_barrier(c, stream)) return -1;  // all segments have landed everywhere
  if (recv_off[P] > 0)
    FSP_CUDA_CHECK(cudaMemcpyAsync(recv, c->a2a.local, (size_t) recv_off[P] * esz, cudaMemcpyDeviceToDevice, st));
  FSP_CUDA_CHECK(cudaStreamSynchronize(st));
  return check_peer_error(c, "fspcomm_alltoallv");
}

// Host-synchronising barrier through NCCL: no time limit, for points where ranks may be seconds apart (host callbacks).
int fspcomm_barrier_sync(fspcomm_t c) {
  if (!c || c->size == 1) return 0;
  return agree(c, true) < 0 ? -1 : 0;
}

// Collective, synchronising: all_host[p] = the value rank p passed (|value| < 2^53).
int fspcomm_gather_long(fspcomm_t c, long mine, long *all_host) {
  if (!c || c->size == 1) { all_host[0] = mine; return 0; }
  if (c->size > kMaxRanks) { set_error("fspcomm_gather_long: more than %d ranks", kMaxRanks); return -1; }
  if (coll_scratch(c)) return -1;
  double v[kMaxRanks];
  for (int p = 0; p < c->size; ++p) v[p] = p == c->rank ? (double) mine : 0.0;
  FSP_CUDA_CHECK(cudaMemcpy(c->coll, v, sizeof(double) * c->size, cudaMemcpyHostToDevice));
  if (fspcomm_allreduce_sum(c, c->coll, c->size, nullptr)) return -1;
  FSP_CUDA_CHECK(cudaMemcpy(v, c->coll, sizeof(double) * c->size, cudaMemcpyDeviceToHost));
  for (int p = 0; p < c->size; ++p) all_host[p] = (long) v[p];
  return check_peer_error(c, "fspcomm_gather_long");
}

}  // extern "C"

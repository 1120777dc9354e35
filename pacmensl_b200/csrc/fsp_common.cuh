// fsp_common.cuh -- shared helpers for the sm_100a kernels behind include/fsp_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fsp_b200.h"

namespace fspb {

void        set_error(const char *fmt, ...);
// Stream-ordered pooled allocation (cudaMallocAsync on the legacy stream with an unbounded release threshold): the
// set-up paths allocate many short-lived temporaries and plain cudaMalloc/cudaFree synchronise the device each time.
cudaError_t pool_malloc_bytes(void **p, size_t bytes);
cudaError_t pool_free(void *p);
template <typename T>
inline cudaError_t pmalloc(T **p, size_t bytes) { return pool_malloc_bytes(reinterpret_cast<void **>(p), bytes); }
inline cudaError_t pfree(void *p) { return pool_free(p); }
cudaStream_t resolve_stream(void *stream);
void        count_launch(int n = 1);
int         sm_count();

#define FSP_CUDA_CHECK(expr)                                                                     \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      fspb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -1;                                                                                 \
    }                                                                                            \
  } while (0)

#define FSP_LAUNCH_CHECK()                                                                  \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      fspb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -1;                                                                            \
    }                                                                                       \
    fspb::count_launch();                                                                   \
  } while (0)

// ---- streaming (read-once) loads / write-once stores: keep x resident in L1/L2 instead ----------
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ double2 ld_stream(const double2 *p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }
__device__ __forceinline__ int2 ld_stream(const int2 *p) { return __ldcs(p); }
__device__ __forceinline__ int4 ld_stream(const int4 *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(double2 *p, double2 v) { __stcs(p, v); }

// ---- warp / block reductions (fixed shape => deterministic) ---------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; result valid in thread 0. smem must hold >= 32 doubles. blockDim.x multiple of 32.
__device__ __forceinline__ double block_sum(double v, double *smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect smem reuse across consecutive calls
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = lane < nw ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

// ---- peer-memory signalling (multi-GPU fast path, fspcomm.cu / fspmat.cu) --------------------------
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= epoch; false (and *err = 1) after the time-out so that a lost peer cannot hang the GPU.  Once
// the error flag is up (an earlier wait timed out) every later wait gives up after ~1000 polls instead of another
// 20 s, so a failed rank costs its peers one time-out, not one per kernel.  err points at two mapped host words:
// err[0] = error flag, err[1] = time-out in milliseconds (0: the default kSpinTimeoutNs; FSP_SPIN_TIMEOUT_MS).
// Callers must act on a false return: poison what they were about to produce (NaN) -- never consume the peer data.
__device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long epoch, unsigned int *err) {
  if (ld_acquire_sys(flag) >= epoch) return true;
  const unsigned long long t0 = global_ns();
  unsigned                 polls = 0;
  while (ld_acquire_sys(flag) < epoch) {
    __nanosleep(64);
    if ((++polls & 1023u) == 0u) {
      if (err && *(volatile unsigned int *) err) return false;
      const unsigned           ms = err ? ((volatile unsigned int *) err)[1] : 0u;
      const unsigned long long limit = ms ? (unsigned long long) ms * 1000000ull : kSpinTimeoutNs;
      if (global_ns() - t0 > limit) {
        if (err) *(volatile unsigned int *) err = 1u;
        return false;
      }
    }
  }
  return true;
}

// ---- push role of the peer-memory halo (fspcomm.cu: halo_push_kernel; fspmat.cu: leading CTAs of the fused action) ---
// pack + all-to-all over NVLink + signal: entry q of the send list goes straight into the ghost buffer of the peer that
// needs it; the last push CTA to finish publishes the epoch on every peer (release at system scope orders it after all
// the data stores, which each CTA made visible with a system-scope fence before taking its ticket).
struct PushView {
  int                 size, rank;
  int                 n_ctas;              // CTAs that share the send list (grid-stride over it)
  long                n_send;
  unsigned long long  epoch;
  const int          *send_idx;            // local indices into x, packed per destination in rank order
  long                send_off[FSP_P2P_MAX_RANKS + 1];
  double             *dst[FSP_P2P_MAX_RANKS];   // peer p's ghost buffer of this parity, offset to my segment
  unsigned long long *flag[FSP_P2P_MAX_RANKS];  // peer p's halo flag of this parity for my rank
  unsigned           *block_counter;
};
__device__ __forceinline__ void push_role(const PushView &a, const double *__restrict__ x, int cta) {
  for (long q = (long) cta * blockDim.x + threadIdx.x; q < a.n_send; q += (long) a.n_ctas * blockDim.x) {
    int p = 0;
    while (q >= a.send_off[p + 1]) ++p;
    a.dst[p][q - a.send_off[p]] = a.send_idx ? x[a.send_idx[q]] : x[q];  // send_idx == null: x is the packed send buffer
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool push_last;
  if (threadIdx.x == 0) push_last = (atomicAdd(a.block_counter, 1u) == (unsigned) a.n_ctas - 1u);
  __syncthreads();
  if (!push_last) return;
  __threadfence_system();
  if ((int) threadIdx.x < a.size) st_release_sys(a.flag[threadIdx.x], a.epoch);
  if (threadIdx.x == 0) *a.block_counter = 0u;
}

}  // namespace fspb

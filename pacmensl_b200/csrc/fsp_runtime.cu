// fsp_runtime.cu -- device selection, memory, streams, events, error string (include/fsp_b200.h "Runtime").
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "fsp_common.cuh"

namespace fspb {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
cudaStream_t resolve_stream(void *stream) { return (cudaStream_t) stream; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
    cached = p.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}
static void pool_init_once() {
  static thread_local int done_dev = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev == done_dev) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done_dev = dev;
}
cudaError_t pool_malloc_bytes(void **p, size_t bytes) {
  pool_init_once();
  return cudaMallocAsync(p, bytes ? bytes : 8, (cudaStream_t) 0);
}
cudaError_t pool_free(void *p) {
  if (!p) return cudaSuccess;
  return cudaFreeAsync(p, (cudaStream_t) 0);
}
}  // namespace fspb

using namespace fspb;

extern "C" {

int fsp_device_count(int *count) { FSP_CUDA_CHECK(cudaGetDeviceCount(count)); return 0; }
int fsp_device_set(int device) { FSP_CUDA_CHECK(cudaSetDevice(device)); return 0; }
int fsp_device_get(int *device) { FSP_CUDA_CHECK(cudaGetDevice(device)); return 0; }
int fsp_device_sm_count(int *count) { *count = sm_count(); return 0; }
const char *fsp_last_error(void) { return g_err; }
int fsp_malloc(void **p, size_t bytes) { FSP_CUDA_CHECK(pool_malloc_bytes(p, bytes)); return 0; }
int fsp_free(void *p) { if (p) FSP_CUDA_CHECK(pool_free(p)); return 0; }
int fsp_malloc_host(void **p, size_t bytes) { FSP_CUDA_CHECK(cudaMallocHost(p, bytes ? bytes : 8)); return 0; }
int fsp_free_host(void *p) { if (p) FSP_CUDA_CHECK(cudaFreeHost(p)); return 0; }
// Small transfers (solver scalars, Hessenberg columns, sink entries: a few doubles, many times per step) are staged
// through a per-thread pinned buffer: a copy to/from pageable memory goes through the driver's own staging path and
// costs about twice the latency.
static constexpr size_t kStageBytes = 8192;
static void *stage_buffer() {
  static thread_local void *buf = nullptr;
  if (!buf && cudaMallocHost(&buf, kStageBytes) != cudaSuccess) { buf = nullptr; cudaGetLastError(); }
  return buf;
}
int fsp_memcpy_h2d(void *d, const void *s, size_t b, void *st) {
  void *stage = b <= kStageBytes ? stage_buffer() : nullptr;
  if (stage) {
    memcpy(stage, s, b);
    FSP_CUDA_CHECK(cudaMemcpyAsync(d, stage, b, cudaMemcpyHostToDevice, resolve_stream(st)));
  } else {
    FSP_CUDA_CHECK(cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, resolve_stream(st)));
  }
  FSP_CUDA_CHECK(cudaStreamSynchronize(resolve_stream(st)));
  return 0;
}
int fsp_memcpy_d2h(void *d, const void *s, size_t b, void *st) {
  void *stage = b <= kStageBytes ? stage_buffer() : nullptr;
  FSP_CUDA_CHECK(cudaMemcpyAsync(stage ? stage : d, s, b, cudaMemcpyDeviceToHost, resolve_stream(st)));
  FSP_CUDA_CHECK(cudaStreamSynchronize(resolve_stream(st)));
  if (stage) memcpy(d, stage, b);
  return 0;
}
int fsp_memcpy_h2d_async(void *d, const void *s, size_t b, void *st) {
  FSP_CUDA_CHECK(cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, resolve_stream(st)));
  return 0;
}
int fsp_memcpy_d2h_async(void *d, const void *s, size_t b, void *st) {
  FSP_CUDA_CHECK(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, resolve_stream(st)));
  return 0;
}
int fsp_memcpy_d2d(void *d, const void *s, size_t b, void *st) {
  FSP_CUDA_CHECK(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, resolve_stream(st)));
  return 0;
}
int fsp_memset(void *d, int byte, size_t b, void *st) {
  FSP_CUDA_CHECK(cudaMemsetAsync(d, byte, b, resolve_stream(st)));
  return 0;
}
int fsp_stream_create(void **s) {
  // non-blocking, highest priority: the library's side streams carry small communication kernels that must be
  // scheduled ahead of the bulk kernel they overlap with
  cudaStream_t st;
  int lo = 0, hi = 0;
  FSP_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  FSP_CUDA_CHECK(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, hi));
  *s = (void *) st;
  return 0;
}
int fsp_stream_destroy(void *s) { if (s) FSP_CUDA_CHECK(cudaStreamDestroy((cudaStream_t) s)); return 0; }
int fsp_stream_sync(void *s) { FSP_CUDA_CHECK(cudaStreamSynchronize(resolve_stream(s))); return 0; }
int fsp_device_sync(void) { FSP_CUDA_CHECK(cudaDeviceSynchronize()); return 0; }
int fsp_event_create(void **e) {
  cudaEvent_t ev;
  FSP_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDefault));
  *e = (void *) ev;
  return 0;
}
int fsp_event_destroy(void *e) { if (e) FSP_CUDA_CHECK(cudaEventDestroy((cudaEvent_t) e)); return 0; }
int fsp_event_record(void *e, void *s) { FSP_CUDA_CHECK(cudaEventRecord((cudaEvent_t) e, resolve_stream(s))); return 0; }
int fsp_event_sync(void *e) { FSP_CUDA_CHECK(cudaEventSynchronize((cudaEvent_t) e)); return 0; }
int fsp_stream_wait_event(void *s, void *e) {
  FSP_CUDA_CHECK(cudaStreamWaitEvent(resolve_stream(s), (cudaEvent_t) e, 0));
  return 0;
}
int fsp_event_elapsed_ms(void *a, void *b, float *ms) {
  FSP_CUDA_CHECK(cudaEventSynchronize((cudaEvent_t) b));
  FSP_CUDA_CHECK(cudaEventElapsedTime(ms, (cudaEvent_t) a, (cudaEvent_t) b));
  return 0;
}
long long fsp_launch_count(void) { return g_launches.load(); }

// ---- CUDA graphs for launch-bound inner loops -----------------------------------------------------
struct fsp_graph_s {
  cudaGraphExec_t exec = nullptr;
  size_t          n_kernels = 0;
};

int fsp_graph_begin_capture(void *stream) {
  if (!stream) { set_error("fsp_graph_begin_capture: the legacy default stream cannot be captured"); return -1; }
  // relaxed: only work submitted to `stream` is captured; other streams/threads of the process are unaffected
  FSP_CUDA_CHECK(cudaStreamBeginCapture((cudaStream_t) stream, cudaStreamCaptureModeRelaxed));
  return 0;
}

int fsp_graph_end_capture(void *stream, fsp_graph_t *out) {
  *out = nullptr;
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture((cudaStream_t) stream, &g);
  if (e != cudaSuccess || !g) {
    cudaGetLastError();
    set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    if (g) cudaGraphDestroy(g);
    return -1;
  }
  fsp_graph_s *h = new fsp_graph_s();
  size_t n_nodes = 0;
  cudaGraphGetNodes(g, nullptr, &n_nodes);
  if (n_nodes > 0) {
    std::vector<cudaGraphNode_t> nodes(n_nodes);
    cudaGraphGetNodes(g, nodes.data(), &n_nodes);
    for (size_t i = 0; i < n_nodes; ++i) {
      cudaGraphNodeType t;
      if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) h->n_kernels += 1;
    }
  }
  e = cudaGraphInstantiate(&h->exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    delete h;
    return -1;
  }
  count_launch(-(int) h->n_kernels);  // launches seen while capturing did not execute
  *out = h;
  return 0;
}

int fsp_graph_abort_capture(void *stream) {
  cudaGraph_t g = nullptr;
  cudaStreamEndCapture((cudaStream_t) stream, &g);
  if (g) cudaGraphDestroy(g);
  cudaGetLastError();
  return 0;
}

int fsp_graph_launch(fsp_graph_t h, void *stream) {
  FSP_CUDA_CHECK(cudaGraphLaunch(h->exec, resolve_stream(stream)));
  // the kernels a replay launches were counted when they were captured, not now: count them per replay
  count_launch((int) h->n_kernels);
  return 0;
}

int fsp_graph_num_kernels(fsp_graph_t h, long *n) { *n = (long) h->n_kernels; return 0; }

int fsp_graph_destroy(fsp_graph_t h) {
  if (!h) return 0;
  if (h->exec) cudaGraphExecDestroy(h->exec);
  delete h;
  return 0;
}

}  // extern "C"

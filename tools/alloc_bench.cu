// Measures device allocation cost (cudaMalloc vs the stream-ordered pool) for solver-workspace sized blocks:
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/alloc_bench.cu -o build/alloc_bench && build/alloc_bench
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaFree(0);
  const size_t bytes = 800ull << 20;
  const int    N = 30;
  std::vector<void *> p(N);
  double t0 = now();
  for (int i = 0; i < N; ++i) cudaMalloc(&p[i], bytes);
  cudaDeviceSynchronize();
  double t1 = now();
  printf("cudaMalloc        %d x 800 MiB: %.3f ms each\n", N, 1e3 * (t1 - t0) / N);
  for (int i = 0; i < N; ++i) cudaMemset(p[i], 0, bytes);
  cudaDeviceSynchronize();
  double t2 = now();
  printf("first-touch memset: %.3f ms each (%.0f GB/s)\n", 1e3 * (t2 - t1) / N, bytes / 1e9 / ((t2 - t1) / N));
  for (int i = 0; i < N; ++i) cudaMemset(p[i], 0, bytes);
  cudaDeviceSynchronize();
  double t3 = now();
  printf("second memset:      %.3f ms each (%.0f GB/s)\n", 1e3 * (t3 - t2) / N, bytes / 1e9 / ((t3 - t2) / N));
  for (int i = 0; i < N; ++i) cudaFree(p[i]);
  cudaDeviceSynchronize();
  double t4 = now();
  printf("cudaFree:           %.3f ms each\n", 1e3 * (t4 - t3) / N);
  cudaMemPool_t pool;
  cudaDeviceGetDefaultMemPool(&pool, 0);
  unsigned long long thr = ~0ull;
  cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  for (int round = 0; round < 2; ++round) {
    double a = now();
    for (int i = 0; i < N; ++i) cudaMallocAsync(&p[i], bytes, 0);
    cudaDeviceSynchronize();
    double b = now();
    printf("cudaMallocAsync round %d: %.3f ms each\n", round, 1e3 * (b - a) / N);
    for (int i = 0; i < N; ++i) cudaFreeAsync(p[i], 0);
    cudaDeviceSynchronize();
  }
  // one big block carved by hand
  double a = now();
  void *big;
  cudaMalloc(&big, bytes * N);
  cudaDeviceSynchronize();
  printf("cudaMalloc one %zu MiB block: %.3f ms\n", (bytes * N) >> 20, 1e3 * (now() - a));
  cudaFree(big);
  return 0;
}

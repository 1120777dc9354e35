// fspvec.cu -- fp64 device-vector kernels (include/fsp_b200.h "Device vectors").
//
// Replaces the PETSc Vec BLAS-1 calls of the hot path (VecSet/Copy/Scale/AXPY/Dot/Norm/MAXPY,
// reference src/OdeSolver/KrylovFsp.cpp:138,153,244-252,280-309) and the N_Vector ops CVODE needs.
// All kernels are HBM-bound streaming passes: grid-stride loops with 128-bit loads where the base
// pointers allow it, grids sized as a multiple of the SM count, and two-stage fixed-shape reductions
// (per-block partials -> last block sums them in a fixed order), so results are deterministic.
#include <cooperative_groups.h>
#include <cub/cub.cuh>

#include "fsp_common.cuh"

using namespace fspb;

namespace {

constexpr int kThreads = 256;
constexpr int kMaxRed = 8;          // outputs per reduction kernel
constexpr int kMaxBlocks = 148 * 8; // partials per output
constexpr int kSlots = 64;          // scratch ring (reductions in flight across streams)

struct RedScratch {
  double   *partials = nullptr;  // [kSlots][kMaxRed][kMaxBlocks]
  unsigned *counters = nullptr;  // [kSlots]
  int       next = 0;
  int       device = -1;
};
thread_local RedScratch g_scratch[16];

int get_scratch(double **partials, unsigned **counter) {
  int dev = 0;
  FSP_CUDA_CHECK(cudaGetDevice(&dev));
  RedScratch &s = g_scratch[dev & 15];
  if (!s.partials) {
    FSP_CUDA_CHECK(pmalloc(&s.partials, sizeof(double) * kSlots * kMaxRed * kMaxBlocks));
    FSP_CUDA_CHECK(pmalloc(&s.counters, sizeof(unsigned) * kSlots));
    FSP_CUDA_CHECK(cudaMemset(s.counters, 0, sizeof(unsigned) * kSlots));
    s.device = dev;
  }
  int slot = s.next;
  s.next = (s.next + 1) % kSlots;
  *partials = s.partials + (size_t) slot * kMaxRed * kMaxBlocks;
  *counter = s.counters + slot;
  return 0;
}

inline int grid_for(long n, int per_thread = 4) {
  long b = (n + (long) kThreads * per_thread - 1) / ((long) kThreads * per_thread);
  long cap = (long) sm_count() * 8;
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (b < 1) b = 1;
  return (int) (b < cap ? b : cap);
}

// ---------------------------------------------------------------------------------------------
// elementwise kernels
// ---------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(kThreads) map_kernel(F f, long n) {
  long stride = (long) gridDim.x * blockDim.x;
  long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  // 2-way unrolled grid-stride loop: independent loads in flight
  for (; i + stride < n; i += 2 * stride) {
    f(i);
    f(i + stride);
  }
  if (i < n) f(i);
}

// pairs of doubles per thread (128-bit accesses) when every base pointer is 16B aligned
template <class F2>
__global__ void __launch_bounds__(kThreads) map2_kernel(F2 f, long n2) {
  long stride = (long) gridDim.x * blockDim.x;
  long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < n2; i += 2 * stride) {
    f(i);
    f(i + stride);
  }
  if (i < n2) f(i);
}

inline bool aligned16(const void *p) { return (((uintptr_t) p) & 15u) == 0; }

struct SetF {
  double *y; double a;
  __device__ void operator()(long i) const { y[i] = a; }
};
struct CopyF {
  double *y; const double *x;
  __device__ void operator()(long i) const { y[i] = x[i]; }
};
struct Copy2F {
  double2 *y; const double2 *x;
  __device__ void operator()(long i) const { y[i] = x[i]; }
};
struct ScaleF {
  double *y; double a;
  __device__ void operator()(long i) const { y[i] *= a; }
};
struct Scale2F {
  double2 *y; double a;
  __device__ void operator()(long i) const { double2 v = y[i]; v.x *= a; v.y *= a; y[i] = v; }
};
struct AxpyF {
  double *y; double a; const double *x;
  __device__ void operator()(long i) const { y[i] = fma(a, x[i], y[i]); }
};
struct Axpy2F {
  double2 *y; double a; const double2 *x;
  __device__ void operator()(long i) const {
    double2 v = y[i], w = x[i];
    v.x = fma(a, w.x, v.x); v.y = fma(a, w.y, v.y);
    y[i] = v;
  }
};
struct LinSumF {
  double *z; double a; const double *x; double b; const double *y;
  __device__ void operator()(long i) const { z[i] = a * x[i] + b * y[i]; }
};
struct WLinCombF {
  double *z; const double *w; double a; const double *x; double b; const double *y;
  __device__ void operator()(long i) const {
    double v = a * x[i] + b * y[i];
    z[i] = w ? w[i] * v : v;
  }
};
struct DivF {
  double *z; const double *x; const double *w;
  __device__ void operator()(long i) const { z[i] = x[i] / w[i]; }
};
struct ProdF {
  double *z; const double *x; const double *w;
  __device__ void operator()(long i) const { z[i] = x[i] * w[i]; }
};
struct LinComb3F {
  double *z; double c0; const double *x0; double c1; const double *x1; double c2; const double *x2;
  __device__ void operator()(long i) const {
    double v = c0 * x0[i] + c1 * x1[i];
    if (x2) v += c2 * x2[i];
    z[i] = v;
  }
};
struct ScatterF {
  double *pn; const double *po; const int *idx;
  __device__ void operator()(long i) const { int j = idx[i]; if (j >= 0) pn[j] = po[i]; }
};
struct ScatterRangeF {
  double *pn; long n_new; const double *v; const int *g; long start;
  __device__ void operator()(long i) const {
    long j = (long) g[i] - start;
    if (g[i] >= 0 && j >= 0 && j < n_new) pn[j] = v[i];
  }
};
struct OwnerStarts { long s[FSP_P2P_MAX_RANKS + 1]; int n; };
// key = owner of global index idx[i] (n = outside every block: sorted to the end and dropped)
struct OwnerKeyF {
  const int *idx; OwnerStarts st; unsigned char *key; int *pos;
  __device__ void operator()(long i) const {
    const long g = idx[i];
    int        r = st.n;
    if (g >= 0 && g < st.s[st.n]) {
      r = 0;
      while (g >= st.s[r + 1]) ++r;
    }
    key[i] = (unsigned char) r;
    pos[i] = (int) i;
  }
};
struct PackPairF {
  int *oi; double *ov; const int *idx; const double *val; const int *perm;
  __device__ void operator()(long i) const { const int j = perm[i]; oi[i] = idx[j]; ov[i] = val[j]; }
};
// bounds[r] = first position of key >= r in the sorted keys (r = 0 .. n_ranks)
__global__ void owner_bounds_kernel(const unsigned char *key, int n, int n_ranks, int *bounds) {
  const int r = threadIdx.x;
  if (r > n_ranks) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((int) key[mid] < r) lo = mid + 1; else hi = mid;
  }
  bounds[r] = lo;
}
struct GatherF {
  double *o; const double *x; const int *idx;
  __device__ void operator()(long i) const { int j = idx[i]; o[i] = j >= 0 ? x[j] : 0.0; }
};

constexpr int kMaxpy = 64;
struct MaxpyArgs {
  const double *X[kMaxpy];
  double        a[kMaxpy];
};
__global__ void __launch_bounds__(kThreads) maxpy_kernel(double *y, double beta, int m, MaxpyArgs args, long n) {
  long stride = (long) gridDim.x * blockDim.x;
  for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double acc = beta == 0.0 ? 0.0 : beta * y[i];
    int k = 0;
    for (; k + 4 <= m; k += 4) {  // 4 independent loads in flight per thread
      double x0 = __ldcs(args.X[k] + i), x1 = __ldcs(args.X[k + 1] + i), x2 = __ldcs(args.X[k + 2] + i),
             x3 = __ldcs(args.X[k + 3] + i);
      acc = fma(args.a[k], x0, acc);
      acc = fma(args.a[k + 1], x1, acc);
      acc = fma(args.a[k + 2], x2, acc);
      acc = fma(args.a[k + 3], x3, acc);
    }
    for (; k < m; ++k) acc = fma(args.a[k], __ldcs(args.X[k] + i), acc);
    y[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// reductions: M outputs per pass
// ---------------------------------------------------------------------------------------------
enum RedOp { RED_SUM = 0, RED_MIN = 1 };

template <int M, int OP, class F>
__global__ void __launch_bounds__(kThreads) reduce_kernel(F f, long n, double *partials, unsigned *counter,
                                                          double *out) {
  __shared__ double smem[32];
  __shared__ bool   is_last;
  double acc[M];
#pragma unroll
  for (int m = 0; m < M; ++m) acc[m] = OP == RED_MIN ? 1.0e300 : 0.0;
  long stride = (long) gridDim.x * blockDim.x;
  for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i, acc);
#pragma unroll
  for (int m = 0; m < M; ++m) {
    double r;
    if (OP == RED_MIN) {
      // min via block_sum-shaped reduction
      double v = warp_min(acc[m]);
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
      __syncthreads();
      if (lane == 0) smem[warp] = v;
      __syncthreads();
      r = 1.0e300;
      if (warp == 0) { r = lane < nw ? smem[lane] : 1.0e300; r = warp_min(r); }
    } else {
      r = block_sum(acc[m], smem);
    }
    if (threadIdx.x == 0) partials[(size_t) m * kMaxBlocks + blockIdx.x] = r;
  }
  __threadfence();
  if (threadIdx.x == 0) {
    unsigned prev = atomicAdd(counter, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
#pragma unroll
    for (int m = 0; m < M; ++m) {
      double v = OP == RED_MIN ? 1.0e300 : 0.0;
      for (int b = threadIdx.x; b < (int) gridDim.x; b += blockDim.x) {
        double p = __ldcg(&partials[(size_t) m * kMaxBlocks + b]);
        v = OP == RED_MIN ? fmin(v, p) : v + p;
      }
      double r;
      if (OP == RED_MIN) {
        double w = warp_min(v);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        __syncthreads();
        if (lane == 0) smem[warp] = w;
        __syncthreads();
        r = 1.0e300;
        if (warp == 0) { r = lane < nw ? smem[lane] : 1.0e300; r = warp_min(r); }
      } else {
        r = block_sum(v, smem);
      }
      if (threadIdx.x == 0) out[m] = r;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

template <int M, int OP, class F>
int launch_reduce(F f, long n, double *out_dev, void *stream) {
  double   *partials;
  unsigned *counter;
  if (get_scratch(&partials, &counter)) return -1;
  int grid = grid_for(n);
  reduce_kernel<M, OP, F><<<grid, kThreads, 0, resolve_stream(stream)>>>(f, n, partials, counter, out_dev);
  FSP_LAUNCH_CHECK();
  return 0;
}

struct DotF {
  const double *x, *y;
  __device__ void operator()(long i, double (&acc)[1]) const { acc[0] = fma(x[i], y[i], acc[0]); }
};
struct SumF {
  const double *x;
  __device__ void operator()(long i, double (&acc)[1]) const { acc[0] += x[i]; }
};
struct Norm1F {
  const double *x;
  __device__ void operator()(long i, double (&acc)[1]) const { acc[0] += fabs(x[i]); }
};
struct WsqF {
  const double *x, *w;
  __device__ void operator()(long i, double (&acc)[1]) const { double v = x[i] * w[i]; acc[0] = fma(v, v, acc[0]); }
};
template <int M>
struct MdotF {
  const double *x;
  const double *Y[M];
  __device__ void operator()(long i, double (&acc)[M]) const {
    double xi = x[i];
#pragma unroll
    for (int m = 0; m < M; ++m) acc[m] = fma(xi, Y[m][i], acc[m]);
  }
};
struct EwtF {
  double *w; const double *y; double rtol, atol; double *winv;
  __device__ void operator()(long i, double (&acc)[1]) const {
    double t = rtol * fabs(y[i]) + atol;
    w[i] = 1.0 / t;
    if (winv) winv[i] = t;  // the un-scaling x ./ w of the GMRES loop becomes a multiplication
    acc[0] = fmin(acc[0], t);
  }
};
struct RatioAbsMaxF {  // accumulates min of the NEGATED ratio so that the RED_MIN machinery yields the max
  const double *x, *y; double a, b;
  __device__ void operator()(long i, double (&acc)[1]) const { acc[0] = fmin(acc[0], -fabs(x[i]) / (a * fabs(y[i]) + b)); }
};
__global__ void negate_scalar_kernel(double *v) { *v = (*v > 1.0e299) ? 0.0 : -*v; }
struct AxpyDotF {
  double *w; const double *h; double sign; const double *v; const double *u;
  __device__ void operator()(long i, double (&acc)[1]) const {
    double wi = fma(-sign * (*h), v[i], w[i]);
    w[i] = wi;
    acc[0] = fma(wi, u ? u[i] : wi, acc[0]);
  }
};

// ---- fused passes of the BDF integrator (host/BdfCore.cpp) -------------------------------------------------------
// b = c0 x0 + c1 x1 + c2 x2 ; v = b .* w ; acc = sum v^2     (Newton residual + GMRES start vector + both norms)
struct Lc3WprodSqF {
  double *b, *v; double c0; const double *x0; double c1; const double *x1; double c2; const double *x2; const double *w;
  __device__ void operator()(long i, double (&acc)[1]) const {
    const double bi = c0 * x0[i] + c1 * x1[i] + c2 * x2[i];
    const double vi = bi * w[i];
    b[i] = bi;
    v[i] = vi;
    acc[0] = fma(vi, vi, acc[0]);
  }
};
// v *= a ; t = v ./ w      (normalise a Krylov vector and form the unscaled operand of the next J*v in one pass)
struct ScaleDivF {
  double *v; double a; double *t; const double *w;
  __device__ void operator()(long i) const {
    const double vi = v[i] * a;
    v[i] = vi;
    t[i] = vi / w[i];
  }
};
// v *= a ; t = v .* winv    (the same with the reciprocal weights: fp64 division halves the bandwidth of these passes)
struct ScaleMulF {
  double *v; double a; double *t; const double *winv;
  __device__ void operator()(long i) const {
    const double vi = v[i] * a;
    v[i] = vi;
    t[i] = vi * winv[i];
  }
};
// d = x .* winv ; acor += d ; ycur = zn0 + acor ; acc = sum x^2  (== sum (d .* ewt)^2 because winv .* ewt == 1)
struct NewtonUpdateMulF {
  const double *x, *winv; double *acor; const double *zn0; double *ycur;
  __device__ void operator()(long i, double (&acc)[1]) const {
    const double xi = x[i];
    const double a = fma(xi, winv[i], acor[i]);
    acor[i] = a;
    ycur[i] = zn0[i] + a;
    acc[0] = fma(xi, xi, acc[0]);
  }
};
// d = dw ? x ./ dw : x ; acor += d ; ycur = zn0 + acor ; acc = sum (d .* ewt)^2     (end of a Newton iteration)
struct NewtonUpdateF {
  const double *x, *dw, *ewt; double *acor; const double *zn0; double *ycur;
  __device__ void operator()(long i, double (&acc)[1]) const {
    const double e = ewt[i];
    const double d = dw ? x[i] / dw[i] : x[i];
    const double a = acor[i] + d;
    acor[i] = a;
    ycur[i] = zn0[i] + a;
    const double de = d * e;
    acc[0] = fma(de, de, acc[0]);
  }
};

// Nordsieck history array Z[0..L) of the BDF integrator: optional rescale Z[j] *= f[j] (j >= 1), then the Pascal
// triangle product (prediction, sign +1) or its inverse (restore, sign -1), all in registers: 16 L bytes per row
// instead of 24 bytes per row for each of the L(L-1)/2 separate axpys.  Uses the same operation order and un-fused
// roundings as the axpy sequence, so the result is bit-identical to it.
constexpr int kMaxNord = 8;
struct NordArgs {
  double *Z[kMaxNord];
  double  f[kMaxNord];
};
template <int L>
__global__ void __launch_bounds__(kThreads) nordsieck_kernel(NordArgs a, int has_scale, int pascal, long n) {
  const long stride = (long) gridDim.x * blockDim.x;
  for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double z[L];
#pragma unroll
    for (int j = 0; j < L; ++j) z[j] = a.Z[j][i];
    if (has_scale) {
#pragma unroll
      for (int j = 1; j < L; ++j) z[j] = __dmul_rn(z[j], a.f[j]);
    }
    if (pascal > 0) {
#pragma unroll
      for (int k = 1; k < L; ++k)
#pragma unroll
        for (int j = L - 1; j >= k; --j) z[j - 1] = __dadd_rn(z[j - 1], z[j]);
    } else if (pascal < 0) {
#pragma unroll
      for (int k = 1; k < L; ++k)
#pragma unroll
        for (int j = L - 1; j >= k; --j) z[j - 1] = __dsub_rn(z[j - 1], z[j]);
    }
#pragma unroll
    for (int j = 0; j < L; ++j)
      if (j < L - 1 ? (pascal != 0 || (has_scale && j >= 1)) : has_scale) a.Z[j][i] = z[j];
  }
}
// Z[j] += c[j] x  for j < L   (BDF history update after an accepted step): 8 + 16 L bytes per row
template <int L>
__global__ void __launch_bounds__(kThreads) multi_axpy_kernel(NordArgs a, const double *__restrict__ x, long n) {
  const long stride = (long) gridDim.x * blockDim.x;
  for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double xi = x[i];
#pragma unroll
    for (int j = 0; j < L; ++j) a.Z[j][i] = __fma_rn(a.f[j], xi, a.Z[j][i]);
  }
}

// w *= 1/sqrt(*nsq): the factor is formed once per thread (not once per element: fp64 sqrt + division), 128-bit
// accesses when the vector allows it
__global__ void __launch_bounds__(kThreads) scale_rsqrt_kernel(double *__restrict__ w, const double *__restrict__ nsq, long n,
                                                               int vec2) {
  const double s = 1.0 / sqrt(*nsq);
  const long   stride = (long) gridDim.x * blockDim.x;
  if (vec2) {
    double2 *w2 = reinterpret_cast<double2 *>(w);
    const long n2 = n >> 1;
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
      double2 v = w2[i];
      v.x *= s; v.y *= s;
      w2[i] = v;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) w[n - 1] *= s;
  } else {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) w[i] *= s;
  }
}

template <class F>
int launch_map(F f, long n, void *stream) {
  if (n <= 0) return 0;
  map_kernel<F><<<grid_for(n, 2), kThreads, 0, resolve_stream(stream)>>>(f, n);
  FSP_LAUNCH_CHECK();
  return 0;
}
template <class F2>
int launch_map2(F2 f, long n2, void *stream) {
  if (n2 <= 0) return 0;
  map2_kernel<F2><<<grid_for(n2, 2), kThreads, 0, resolve_stream(stream)>>>(f, n2);
  FSP_LAUNCH_CHECK();
  return 0;
}

// ---- incomplete orthogonalisation of one Krylov basis vector in ONE cooperative launch --------------------------
// KrylovFsp's column step after w = A v_j (src/OdeSolver/KrylovFsp.cpp:302-309, q_iop <= 2):
//     h_0 = <w, u_0>;  [w -= h_0 u_0;  h_1 = <w, u_1>;]  w -= h_last u_last;  s2 = <w, w>;  w *= 1/sqrt(s2)
// was 4 launches (dot, 2 x axpy+dot, scale) that each re-read w.  Here every thread owns up to kOrthRows rows of w in
// REGISTERS for the whole sequence (w is read and written once), the inner products are per-CTA partials combined by
// every CTA in the same fixed order after a grid-wide barrier (deterministic, identical in all CTAs), and nothing
// returns to the host.  Vectors too long for the register file (n > grid * 256 * kOrthRows) take the same phases with w
// re-read from memory: still one launch instead of four.
constexpr int kOrthRows = 8;
struct OrthArgs {
  double       *w;
  const double *u[2];     // u[0] unused when nvec == 1
  double       *h;        // outputs: h[0..nvec-1] coefficients, h[nvec] = ||w||^2 before the scaling
  double       *partials; // [3][gridDim.x]
  long          n;
  int           nvec;
};
template <bool IN_REGS>
__global__ void __launch_bounds__(kThreads, 4) iop_orth_kernel(OrthArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ double smem[32];
  __shared__ double bcast;
  const long stride = (long) gridDim.x * blockDim.x;
  const long i0 = (long) blockIdx.x * blockDim.x + threadIdx.x;
  double     wr[IN_REGS ? kOrthRows : 1];
  if (IN_REGS) {
#pragma unroll
    for (int r = 0; r < kOrthRows; ++r) {
      const long i = i0 + r * stride;
      wr[r] = i < a.n ? a.w[i] : 0.0;
    }
  }
  double coef = 0.0;  // coefficient of the pending axpy
  // phases: p = 0 .. nvec: dot with u[first + p] (p < nvec) or with w itself (p == nvec), after applying the pending axpy
  const double *d0 = a.nvec == 2 ? a.u[0] : a.u[1], *d1 = a.u[1];  // (no dynamic indexing of the parameter struct)
  for (int p = 0; p <= a.nvec; ++p) {
    const double *ax = p == 0 ? nullptr : (p == 1 ? d0 : d1);        // w -= coef * ax
    const double *dv = p == a.nvec ? nullptr : (p == 0 ? d0 : d1);   // dot partner (null: w itself)
    double        acc = 0.0;
    if (IN_REGS) {
#pragma unroll
      for (int r = 0; r < kOrthRows; ++r) {
        const long i = i0 + r * stride;
        if (i < a.n) {
          if (ax) wr[r] = fma(-coef, ax[i], wr[r]);
          acc = fma(wr[r], dv ? dv[i] : wr[r], acc);
        }
      }
    } else {
      for (long i = i0; i < a.n; i += stride) {
        double wi = a.w[i];
        if (ax) { wi = fma(-coef, ax[i], wi); a.w[i] = wi; }
        acc = fma(wi, dv ? dv[i] : wi, acc);
      }
    }
    const double r = block_sum(acc, smem);
    if (threadIdx.x == 0) a.partials[(size_t) p * gridDim.x + blockIdx.x] = r;
    grid.sync();
    double v = 0.0;
    for (int b = threadIdx.x; b < (int) gridDim.x; b += blockDim.x) v += __ldcg(a.partials + (size_t) p * gridDim.x + b);
    const double tot = block_sum(v, smem);
    if (threadIdx.x == 0) {
      bcast = tot;
      if (blockIdx.x == 0) a.h[p] = tot;
    }
    __syncthreads();
    coef = bcast;
    __syncthreads();
  }
  // coef == ||w||^2 now
  const double s = 1.0 / sqrt(coef);
  if (IN_REGS) {
#pragma unroll
    for (int r = 0; r < kOrthRows; ++r) {
      const long i = i0 + r * stride;
      if (i < a.n) a.w[i] = wr[r] * s;
    }
  } else {
    for (long i = i0; i < a.n; i += stride) a.w[i] *= s;
  }
}

// ---- post-processing on the device (DiscreteDistribution / SensDiscreteDistribution) ------------------------------
struct ClampCountF {  // x_i = max(x_i, lo); counts the entries that were raised
  double *x; double lo;
  __device__ void operator()(long i, double (&acc)[1]) const {
    if (x[i] < lo) { x[i] = lo; acc[0] += 1.0; }
  }
};
struct WdivDotF {  // sum_i x_i y_i / w_i
  const double *x, *y, *w;
  __device__ void operator()(long i, double (&acc)[1]) const { acc[0] += x[i] * y[i] / w[i]; }
};

// 1-D marginal: out[b] = sum_{i : states[i][species] == b} p[i].  Deterministic: every warp owns a contiguous slice of
// the states and a private row of bins in shared memory; inside a warp the lanes that hit the same bin are combined in
// lane order by the lowest such lane (match_any + shuffles), so no two lanes ever update the same bin concurrently;
// the per-warp rows are then added in warp order, the per-CTA rows in CTA order (second kernel).
constexpr int kMargWarps = 8;
__global__ void __launch_bounds__(kMargWarps * 32) marginal_partial_kernel(const double *__restrict__ p, const int *__restrict__ states,
                                                                           int S, int species, long n, int M, long per_cta,
                                                                           double *__restrict__ partial) {
  extern __shared__ double bins[];  // [kMargWarps][M]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double   *mine = bins + (size_t) warp * M;
  for (int b = lane; b < M; b += 32) mine[b] = 0.0;
  __syncwarp();
  const long begin = (long) blockIdx.x * per_cta, end = min(n, begin + per_cta);
  const long per_warp = (per_cta + kMargWarps - 1) / kMargWarps;
  const long wb = begin + (long) warp * per_warp, we = min(end, wb + per_warp);
  for (long base = wb; base < we; base += 32) {
    const long i = base + lane;
    const bool ok = i < we;
    const int  b = ok ? states[(size_t) i * S + species] : -1;
    const double v = ok ? p[i] : 0.0;
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    const int      leader = __ffs(peers) - 1;
    double         s = 0.0;
    for (unsigned m = peers; m; m &= m - 1) {  // lane order; every lane runs the loop of ITS peer group
      const int src = __ffs(m) - 1;
      s += __shfl_sync(peers, v, src);
    }
    if (lane == leader && b >= 0 && b < M) mine[b] += s;
    __syncwarp();
  }
  __syncthreads();
  for (int b = threadIdx.x; b < M; b += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < kMargWarps; ++w) s += bins[(size_t) w * M + b];
    partial[(size_t) blockIdx.x * M + b] = s;
  }
}
__global__ void marginal_final_kernel(const double *__restrict__ partial, int n_cta, int M, double *__restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= M) return;
  double s = 0.0;
  for (int c = 0; c < n_cta; ++c) s += partial[(size_t) c * M + b];
  out[b] = s;
}
struct MaxSpeciesF {  // min of the negated value == max (RED_MIN machinery)
  const int *states; int S, species;
  __device__ void operator()(long i, double (&acc)[1]) const { acc[0] = fmin(acc[0], -(double) states[(size_t) i * S + species]); }
};

template <int M>
int mdot_impl(double *out, const double *x, const double *const *Y, long n, void *stream) {
  MdotF<M> f;
  f.x = x;
  for (int m = 0; m < M; ++m) f.Y[m] = Y[m];
  return launch_reduce<M, RED_SUM>(f, n, out, stream);
}

int host_result(double *out_host, double *tmp_dev, void *stream) {
  return fsp_memcpy_d2h(out_host, tmp_dev, sizeof(double), stream);  // staged through pinned memory
}

thread_local double *g_tmp[16] = {nullptr};
int tmp_scalar(double **p) {
  int dev = 0;
  FSP_CUDA_CHECK(cudaGetDevice(&dev));
  if (!g_tmp[dev & 15]) FSP_CUDA_CHECK(pmalloc(&g_tmp[dev & 15], sizeof(double) * 16));
  *p = g_tmp[dev & 15];
  return 0;
}

}  // namespace

extern "C" {

int fspvec_set(double *y, double a, long n, void *s) { return launch_map(SetF{y, a}, n, s); }

int fspvec_copy(double *y, const double *x, long n, void *s) {
  if (n <= 0 || y == x) return 0;
  if (aligned16(y) && aligned16(x) && (n % 2 == 0))
    return launch_map2(Copy2F{(double2 *) y, (const double2 *) x}, n / 2, s);
  return launch_map(CopyF{y, x}, n, s);
}

int fspvec_scale(double *y, double a, long n, void *s) {
  if (aligned16(y) && (n % 2 == 0)) return launch_map2(Scale2F{(double2 *) y, a}, n / 2, s);
  return launch_map(ScaleF{y, a}, n, s);
}

int fspvec_axpy(double *y, double a, const double *x, long n, void *s) {
  if (aligned16(y) && aligned16(x) && (n % 2 == 0))
    return launch_map2(Axpy2F{(double2 *) y, a, (const double2 *) x}, n / 2, s);
  return launch_map(AxpyF{y, a, x}, n, s);
}

int fspvec_linear_sum(double *z, double a, const double *x, double b, const double *y, long n, void *s) {
  return launch_map(LinSumF{z, a, x, b, y}, n, s);
}

int fspvec_wlincomb(double *z, const double *w, double a, const double *x, double b, const double *y, long n, void *s) {
  return launch_map(WLinCombF{z, w, a, x, b, y}, n, s);
}
int fspvec_div(double *z, const double *x, const double *w, long n, void *s) { return launch_map(DivF{z, x, w}, n, s); }
int fspvec_prod(double *z, const double *x, const double *w, long n, void *s) { return launch_map(ProdF{z, x, w}, n, s); }
int fspvec_lincomb3(double *z, double c0, const double *x0, double c1, const double *x1, double c2, const double *x2,
                    long n, void *s) {
  return launch_map(LinComb3F{z, c0, x0, c1, x1, c2, x2}, n, s);
}

int fspvec_maxpy(double *y, double beta, int m, const double *alpha, const double *const *X, long n, void *s) {
  if (m < 0 || m > kMaxpy) { set_error("fspvec_maxpy: m=%d out of range (max %d)", m, kMaxpy); return -1; }
  if (n <= 0) return 0;
  MaxpyArgs args;
  for (int k = 0; k < m; ++k) { args.X[k] = X[k]; args.a[k] = alpha[k]; }
  maxpy_kernel<<<grid_for(n, 1), kThreads, 0, resolve_stream(s)>>>(y, beta, m, args, n);
  FSP_LAUNCH_CHECK();
  return 0;
}

int fspvec_mdot(double *out, const double *x, int m, const double *const *Y, long n, void *s) {
  switch (m) {
    case 1: return mdot_impl<1>(out, x, Y, n, s);
    case 2: return mdot_impl<2>(out, x, Y, n, s);
    case 3: return mdot_impl<3>(out, x, Y, n, s);
    case 4: return mdot_impl<4>(out, x, Y, n, s);
    case 5: return mdot_impl<5>(out, x, Y, n, s);
    case 6: return mdot_impl<6>(out, x, Y, n, s);
    case 7: return mdot_impl<7>(out, x, Y, n, s);
    case 8: return mdot_impl<8>(out, x, Y, n, s);
    default: set_error("fspvec_mdot: m=%d out of range (1..8)", m); return -1;
  }
}

// Orthogonalise-and-normalise w against nvec (1 or 2) vectors in one cooperative launch; see iop_orth_kernel.
// h_dev receives nvec coefficients followed by ||w||^2 (before the scaling).  Returns 1 (no launch) when the device
// cannot run the cooperative kernel, so that callers fall back to the separate passes.
int fspvec_iop_orth(double *w, int nvec, const double *u0, const double *u1, double *h_dev, long n, void *stream) {
  if (nvec < 1 || nvec > 2 || n <= 0) return 1;
  static thread_local int coop = -1, max_blocks_regs = 0, max_blocks_mem = 0;
  static thread_local double *partials = nullptr;
  if (coop < 0) {
    int dev = 0, flag = 0;
    FSP_CUDA_CHECK(cudaGetDevice(&dev));
    FSP_CUDA_CHECK(cudaDeviceGetAttribute(&flag, cudaDevAttrCooperativeLaunch, dev));
    int per_sm_r = 0, per_sm_m = 0;
    if (flag) {
      FSP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_r, iop_orth_kernel<true>, kThreads, 0));
      FSP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_m, iop_orth_kernel<false>, kThreads, 0));
    }
    max_blocks_regs = per_sm_r * sm_count();
    max_blocks_mem = per_sm_m * sm_count();
    coop = (flag && max_blocks_regs > 0 && max_blocks_mem > 0) ? 1 : 0;
    if (coop) FSP_CUDA_CHECK(pmalloc(&partials, sizeof(double) * 3 * (size_t) std::max(max_blocks_regs, max_blocks_mem)));
  }
  if (!coop) return 1;
  OrthArgs a;
  a.w = w; a.u[0] = u0; a.u[1] = u1; a.h = h_dev; a.partials = partials; a.n = n; a.nvec = nvec;
  if (nvec == 1) { a.u[1] = u1 ? u1 : u0; a.u[0] = nullptr; }
  // Only the register-resident regime: a vector too long for it is bandwidth-bound, where four streaming kernels with
  // twice the resident threads do better than one persistent kernel that re-reads w in every phase.
  const long threads_needed = (n + kOrthRows - 1) / kOrthRows;
  if (threads_needed > (long) max_blocks_regs * kThreads) return 1;
  (void) max_blocks_mem;
  void *args[] = {&a};
  // as many CTAs as give every thread >= 1 row, at most the co-resident maximum
  const int grid = (int) std::max<long>(1, std::min<long>(max_blocks_regs, (n + kThreads - 1) / kThreads));
  cudaError_t e = cudaLaunchCooperativeKernel((void *) iop_orth_kernel<true>, dim3(grid), dim3(kThreads), args, 0, resolve_stream(stream));
  if (e != cudaSuccess) {
    set_error("fspvec_iop_orth: cooperative launch failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return -1;
  }
  count_launch();
  return 0;
}

// ---- post-processing (src/Fsp/DiscreteDistribution.cpp:171-200, src/SensFsp/SensDiscreteDistribution.cpp:216-271) ----
int fspvec_clamp_min(double *count_out, double *x, double lo, long n, void *s) {
  return launch_reduce<1, RED_SUM>(ClampCountF{x, lo}, n, count_out, s);
}
int fspvec_wdiv_dot(double *out, const double *x, const double *y, const double *w, long n, void *s) {
  return launch_reduce<1, RED_SUM>(WdivDotF{x, y, w}, n, out, s);
}
int fspvec_max_species(double *out_neg, const int *states, int S, int species, long n, void *s) {
  return launch_reduce<1, RED_MIN>(MaxSpeciesF{states, S, species}, n, out_neg, s);
}
int fspvec_marginal(double *out, int M, const double *p, const int *states, int S, int species, long n, void *stream) {
  if (M <= 0) return 0;
  cudaStream_t st = resolve_stream(stream);
  if ((size_t) M * kMargWarps * sizeof(double) > 96 * 1024) { set_error("fspvec_marginal: %d bins exceed the shared-memory rows", M); return -1; }
  if (n <= 0) { FSP_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * M, st)); return 0; }
  const int  n_cta = (int) std::max<long>(1, std::min<long>((n + 4095) / 4096, 4L * sm_count()));
  const long per_cta = (n + n_cta - 1) / n_cta;
  double    *partial = nullptr;
  FSP_CUDA_CHECK(pmalloc(&partial, sizeof(double) * (size_t) n_cta * M));
  const size_t smem = (size_t) M * kMargWarps * sizeof(double);
  if (smem > 48 * 1024)
    FSP_CUDA_CHECK(cudaFuncSetAttribute(marginal_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  marginal_partial_kernel<<<n_cta, kMargWarps * 32, smem, st>>>(p, states, S, species, n, M, per_cta, partial);
  FSP_LAUNCH_CHECK();
  marginal_final_kernel<<<(M + 127) / 128, 128, 0, st>>>(partial, n_cta, M, out);
  FSP_LAUNCH_CHECK();
  FSP_CUDA_CHECK(cudaStreamSynchronize(st));
  pfree(partial);
  return 0;
}

int fspvec_dot(double *out, const double *x, const double *y, long n, void *s) {
  return launch_reduce<1, RED_SUM>(DotF{x, y}, n, out, s);
}
int fspvec_norm2sq(double *out, const double *x, long n, void *s) {
  return launch_reduce<1, RED_SUM>(DotF{x, x}, n, out, s);
}
int fspvec_sum(double *out, const double *x, long n, void *s) { return launch_reduce<1, RED_SUM>(SumF{x}, n, out, s); }
int fspvec_norm1(double *out, const double *x, long n, void *s) {
  return launch_reduce<1, RED_SUM>(Norm1F{x}, n, out, s);
}
int fspvec_wsqsum(double *out, const double *x, const double *w, long n, void *s) {
  return launch_reduce<1, RED_SUM>(WsqF{x, w}, n, out, s);
}
int fspvec_ewt(double *w, const double *y, double rtol, double atol, long n, double *min_out, void *s) {
  double *tmp = min_out;
  if (!tmp && tmp_scalar(&tmp)) return -1;
  return launch_reduce<1, RED_MIN>(EwtF{w, y, rtol, atol, nullptr}, n, tmp, s);
}
int fspvec_ewt_pair(double *w, double *winv, const double *y, double rtol, double atol, long n, double *min_out, void *s) {
  double *tmp = min_out;
  if (!tmp && tmp_scalar(&tmp)) return -1;
  return launch_reduce<1, RED_MIN>(EwtF{w, y, rtol, atol, winv}, n, tmp, s);
}
int fspvec_ratio_absmax(double *out, const double *x, const double *y, double a, double b, long n, void *s) {
  if (launch_reduce<1, RED_MIN>(RatioAbsMaxF{x, y, a, b}, n, out, s)) return -1;
  negate_scalar_kernel<<<1, 1, 0, resolve_stream(s)>>>(out);
  FSP_LAUNCH_CHECK();
  return 0;
}
int fspvec_axpy_dot(double *w, const double *h, double sign, const double *v, const double *u, double *out, long n,
                    void *s) {
  return launch_reduce<1, RED_SUM>(AxpyDotF{w, h, sign, v, u}, n, out, s);
}
int fspvec_scale_rsqrt(double *w, const double *nsq, long n, void *s) {
  if (n <= 0) return 0;
  const int vec2 = aligned16(w) ? 1 : 0;
  scale_rsqrt_kernel<<<grid_for(vec2 ? n / 2 : n, 2), kThreads, 0, resolve_stream(s)>>>(w, nsq, n, vec2);
  FSP_LAUNCH_CHECK();
  return 0;
}

int fspvec_lincomb3_wprod_sqsum(double *b, double *v, double c0, const double *x0, double c1, const double *x1, double c2,
                                const double *x2, const double *w, double *out, long n, void *s) {
  return launch_reduce<1, RED_SUM>(Lc3WprodSqF{b, v, c0, x0, c1, x1, c2, x2, w}, n, out, s);
}
int fspvec_scale_div(double *v, double a, double *t, const double *w, long n, void *s) {
  return launch_map(ScaleDivF{v, a, t, w}, n, s);
}
int fspvec_scale_mul(double *v, double a, double *t, const double *winv, long n, void *s) {
  return launch_map(ScaleMulF{v, a, t, winv}, n, s);
}
int fspvec_newton_update_mul(const double *x, const double *winv, double *acor, const double *zn0, double *ycur,
                             double *out, long n, void *s) {
  return launch_reduce<1, RED_SUM>(NewtonUpdateMulF{x, winv, acor, zn0, ycur}, n, out, s);
}
int fspvec_newton_update(const double *x, const double *dw, const double *ewt, double *acor, const double *zn0,
                         double *ycur, double *out, long n, void *s) {
  return launch_reduce<1, RED_SUM>(NewtonUpdateF{x, dw, ewt, acor, zn0, ycur}, n, out, s);
}

int fspvec_nordsieck(double *const *Z, int L, const double *scale, int pascal, long n, void *s) {
  if (L < 1 || L > kMaxNord) { set_error("fspvec_nordsieck: L=%d out of range (1..%d)", L, kMaxNord); return -1; }
  if (n <= 0 || (!scale && pascal == 0)) return 0;
  NordArgs a;
  for (int j = 0; j < L; ++j) { a.Z[j] = Z[j]; a.f[j] = scale ? scale[j] : 1.0; }
  const int    hs = scale ? 1 : 0;
  const int    grid = grid_for(n, 1);
  cudaStream_t st = resolve_stream(s);
  switch (L) {
#define FSP_NORD(N) case N: nordsieck_kernel<N><<<grid, kThreads, 0, st>>>(a, hs, pascal, n); break;
    FSP_NORD(1) FSP_NORD(2) FSP_NORD(3) FSP_NORD(4) FSP_NORD(5) FSP_NORD(6) FSP_NORD(7) FSP_NORD(8)
#undef FSP_NORD
  }
  FSP_LAUNCH_CHECK();
  return 0;
}

int fspvec_multi_axpy(double *const *Z, int L, const double *coef, const double *x, long n, void *s) {
  if (L < 1 || L > kMaxNord) { set_error("fspvec_multi_axpy: L=%d out of range (1..%d)", L, kMaxNord); return -1; }
  if (n <= 0) return 0;
  NordArgs a;
  for (int j = 0; j < L; ++j) { a.Z[j] = Z[j]; a.f[j] = coef[j]; }
  const int    grid = grid_for(n, 1);
  cudaStream_t st = resolve_stream(s);
  switch (L) {
#define FSP_MAX(N) case N: multi_axpy_kernel<N><<<grid, kThreads, 0, st>>>(a, x, n); break;
    FSP_MAX(1) FSP_MAX(2) FSP_MAX(3) FSP_MAX(4) FSP_MAX(5) FSP_MAX(6) FSP_MAX(7) FSP_MAX(8)
#undef FSP_MAX
  }
  FSP_LAUNCH_CHECK();
  return 0;
}

int fspvec_dot_h(double *out, const double *x, const double *y, long n, void *s) {
  double *tmp;
  if (tmp_scalar(&tmp)) return -1;
  if (fspvec_dot(tmp, x, y, n, s)) return -1;
  return host_result(out, tmp, s);
}
int fspvec_norm2_h(double *out, const double *x, long n, void *s) {
  double *tmp;
  if (tmp_scalar(&tmp)) return -1;
  if (fspvec_norm2sq(tmp, x, n, s)) return -1;
  if (host_result(out, tmp, s)) return -1;
  *out = sqrt(*out);
  return 0;
}
int fspvec_sum_h(double *out, const double *x, long n, void *s) {
  double *tmp;
  if (tmp_scalar(&tmp)) return -1;
  if (fspvec_sum(tmp, x, n, s)) return -1;
  return host_result(out, tmp, s);
}
int fspvec_norm1_h(double *out, const double *x, long n, void *s) {
  double *tmp;
  if (tmp_scalar(&tmp)) return -1;
  if (fspvec_norm1(tmp, x, n, s)) return -1;
  return host_result(out, tmp, s);
}

int fspvec_scatter(double *pn, long n_new, const double *po, const int *idx, long n_old, void *s) {
  if (fspvec_set(pn, 0.0, n_new, s)) return -1;
  return launch_map(ScatterF{pn, po, idx}, n_old, s);
}
int fspvec_scatter_range(double *pn, long n_new, const double *v, const int *g, long n, long start, void *s) {
  return launch_map(ScatterRangeF{pn, n_new, v, g, start}, n, s);
}
int fspvec_gather(double *o, const double *x, const int *idx, long n, void *s) {
  return launch_map(GatherF{o, x, idx}, n, s);
}

// Multi-GPU ExpandVec (src/Sys/PetscWrap.cpp:10-45, VecScatter to the new layout): sort the entries (idx[i], val[i]) by the
// rank that owns global index idx[i] under `starts` (n_ranks + 1 offsets), so that each peer's share is one contiguous
// segment to send; entries outside [0, starts[n_ranks]) are dropped.  counts_host[r] = entries for rank r.
int fspvec_route_by_owner(const int *idx, const double *val, long n, const long *starts_host, int n_ranks, int *idx_sorted,
                          double *val_sorted, long *counts_host, void *s) {
  for (int r = 0; r < n_ranks; ++r) counts_host[r] = 0;
  if (n <= 0) return 0;
  if (n_ranks > FSP_P2P_MAX_RANKS || n >= 0x7FFFFFF0L) { set_error("fspvec_route_by_owner: too many ranks or entries"); return -1; }
  cudaStream_t st = resolve_stream(s);
  OwnerStarts  os;
  os.n = n_ranks;
  for (int r = 0; r <= n_ranks; ++r) os.s[r] = starts_host[r];
  unsigned char *key = nullptr, *key2 = nullptr;
  int           *pos = nullptr, *pos2 = nullptr, *bounds = nullptr;
  void          *tmp = nullptr;
  int            rc = -1;
  do {
    if (pmalloc(&key, (size_t) n) != cudaSuccess || pmalloc(&key2, (size_t) n) != cudaSuccess ||
        pmalloc(&pos, sizeof(int) * (size_t) n) != cudaSuccess || pmalloc(&pos2, sizeof(int) * (size_t) n) != cudaSuccess ||
        pmalloc(&bounds, sizeof(int) * (FSP_P2P_MAX_RANKS + 2)) != cudaSuccess) { set_error("fspvec_route_by_owner: allocation failed"); break; }
    if (launch_map(OwnerKeyF{idx, os, key, pos}, n, s)) break;
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, key, key2, pos, pos2, (int) n, 0, 5, st);
    if (pmalloc(&tmp, need) != cudaSuccess) { set_error("fspvec_route_by_owner: allocation failed"); break; }
    if (cub::DeviceRadixSort::SortPairs(tmp, need, key, key2, pos, pos2, (int) n, 0, 5, st) != cudaSuccess) { set_error("fspvec_route_by_owner: sort failed"); break; }
    count_launch();
    owner_bounds_kernel<<<1, 32, 0, st>>>(key2, (int) n, n_ranks, bounds);
    if (cudaGetLastError() != cudaSuccess) { set_error("fspvec_route_by_owner: launch failed"); break; }
    count_launch();
    if (launch_map(PackPairF{idx_sorted, val_sorted, idx, val, pos2}, n, s)) break;
    int hb[FSP_P2P_MAX_RANKS + 2];
    if (cudaMemcpyAsync(hb, bounds, sizeof(int) * (n_ranks + 1), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
    if (cudaStreamSynchronize(st) != cudaSuccess) break;
    for (int r = 0; r < n_ranks; ++r) counts_host[r] = hb[r + 1] - hb[r];
    rc = 0;
  } while (0);
  if (rc) cudaGetLastError();
  pfree(key); pfree(key2); pfree(pos); pfree(pos2); pfree(bounds); pfree(tmp);
  return rc;
}

}  // extern "C"

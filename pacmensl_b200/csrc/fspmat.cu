// fspmat.cu -- the FSP operator y = A(t) x on the device (include/fsp_b200.h "The FSP operator").
//
// Replaces FspMatrixBase::Action / FspMatrixConstrained::Action of the reference
// (src/Matrix/FspMatrixBase.cpp:36-62, src/Matrix/FspMatrixConstrained.cpp:31-64): there the action is
// (R_tv + 1) PETSc MatMult passes into a work vector, each followed by a VecAXPY, then K tiny sink
// MatMults and a VecScatter ADD.  Here it is ONE kernel launch that
//   * reads every matrix byte exactly once with streaming (evict-first) 128-bit loads,
//   * scales by the time-varying coefficients c_r(t) (passed by value as kernel arguments),
//   * fuses the diagonal term, the gathered off-diagonal terms and the accumulation,
//   * and computes the K sink rows in trailing CTAs with warp-shuffle reductions and a fixed-order
//     (deterministic) final sum.
//
// HBM layout ("reaction-plane ELL"): P = n_tv + n_ti planes of leading dimension ld (multiple of 32
// elements so every plane starts 128/256-byte aligned):
//   col  int32 [P][ld]   local column of x_i - nu_r ; -1 = none ; <= -2 = ghost slot -(col+2)
//   off  fp64  [P][ld]   d_r(x_i - nu_r)  (0 where col == -1)
//   diag fp64  [ND][ld]  ND = n_tv + (n_ti > 0): one plane per TV reaction (+d_r(x_i)) and ONE merged
//                        plane sum_{r in TI} d_r(x_i)
// Algorithmic bytes per row = 8 (x) + 8 (y) + 12 P + 8 ND  (SURVEY.md section 8d).
// Roofline: HBM bandwidth; arithmetic intensity ~0.15 flop/byte, so no tensor cores.
#include <cub/cub.cuh>

#include <string.h>

#include <algorithm>
#include <vector>

#include "fsp_common.cuh"

using namespace fspb;

namespace {

constexpr int kThreads = 256;
constexpr int kMaxPlanes = 32;
constexpr int kSinkChunk = 4096;  // sink entries per sink CTA

struct Coefs {
  double c[kMaxPlanes];   // per off-diagonal plane (TI planes: 1.0)
  double cd[kMaxPlanes];  // per diagonal plane (merged TI plane: 1.0)
};

struct MatView {
  int           n;       // rows this launch processes (local states, or the boundary-list length)
  int           n_rows_main;  // local states: the sink rows of y start here
  int           P, ND;
  long          ld;
  const int    *col;
  const double *off;
  const double *diag;
  // sinks
  int           K, G;          // G = groups = ND
  int           main_blocks;   // CTAs [0, main_blocks) do rows, the rest do sink chunks
  int           sink_blocks;
  const int    *sb_seg;        // [sink_blocks] segment id g*K + k
  const long   *sb_begin;      // [sink_blocks]
  const long   *sb_end;        // [sink_blocks]
  const int    *sink_idx;
  const double *sink_val;
  double       *sink_partials; // [sink_blocks]
  unsigned     *sink_counter;
  int           owns_sinks;
  // split (multi-GPU overlap) phases
  int           ghost_zero;    // 1: ghost entries contribute 0 (interior pass, halo still in flight)
  const int    *row_list;      // non-null: process only these rows (boundary pass); n = list length
  int           write_y_sinks; // 0: sink role writes only sink_out (the owner copies the reduced values later)
  int           row0;          // row-range variant (GHOST == 3): first row of this launch; n = one past the last
  const int    *cta_order;     // single-kernel peer-memory action: CTAs without ghost rows first, the others last
  int           n_interior_ctas;
  // peer-memory mode: sink_out is the sink owner's slot row of this rank; publish this flag after writing it
  unsigned long long *sink_flag_remote;
  unsigned long long  sink_epoch;
};

// Fused epilogue of the single-GPU action (fspmat_action_fused): y = scale .* (beta x + alpha A x) and up to two
// inner products of y, accumulated through per-CTA partials and summed in a fixed order by the last CTA.
struct Epi {
  double        alpha, beta;
  const double *scale;
  int           n_dots;
  const double *vec[2];
  double       *out;
  double       *partials;  // [2][pstride]: one slot per main CTA + one slot (index G) for the sink rows
  long          pstride;
  int           G;         // number of main CTAs
};

struct P2PWait {
  const unsigned long long *halo_flags;
  const unsigned long long *sink_flags;  // null: this rank does not finish the sink rows
  const double             *sink_slots;
  unsigned long long        epoch;
  int                       n_ranks, rank;
  unsigned int             *err;
};

__device__ __forceinline__ double fetch_x(const double *__restrict__ x, const double *__restrict__ ghost, int c) {
  // c >= 0: local entry; c == -1: absent neighbour (contributes 0); c <= -2: ghost slot
  // (ghost == nullptr during the interior pass of the split multi-GPU action: contributes 0, the row is redone later)
  if (c >= 0) return __ldg(x + c);
  if (c == -1 || ghost == nullptr) return 0.0;
  return __ldg(ghost + (-(c + 2)));
}

// ---- sink rows: trailing CTAs --------------------------------------------------------------------
__device__ __forceinline__ void sink_role(const MatView &m, const Coefs &cf, const double *__restrict__ x,
                                          double *__restrict__ y, double *__restrict__ sink_out) {
  __shared__ double smem[32];
  __shared__ bool   is_last;
  const int sb = blockIdx.x - m.main_blocks;
  double    acc = 0.0;
  if (m.sink_blocks > 0 && sb < m.sink_blocks) {
    const long b = m.sb_begin[sb], e = m.sb_end[sb];
    for (long q = b + threadIdx.x; q < e; q += blockDim.x) {
      acc = fma(ld_stream(m.sink_val + q), __ldg(x + ld_stream(m.sink_idx + q)), acc);
    }
  }
  double r = block_sum(acc, smem);
  if (threadIdx.x == 0) m.sink_partials[sb] = r;
  __threadfence();
  if (threadIdx.x == 0) {
    unsigned prev = atomicAdd(m.sink_counter, 1u);
    is_last = (prev == (unsigned) m.sink_blocks - 1u);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // fixed-order final sum: warp w handles constraints k = w, w + nwarps, ...
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = warp; k < m.K; k += nw) {
    double s = 0.0;
    for (int b = lane; b < m.sink_blocks; b += 32) {
      int seg = m.sb_seg[b];
      if (seg >= 0 && seg % m.K == k) s = fma(cf.cd[seg / m.K], __ldcg(m.sink_partials + b), s);
    }
    s = warp_sum(s);
    if (lane == 0) {
      if (m.owns_sinks && m.write_y_sinks) y[m.n_rows_main + k] = s;
      if (sink_out) sink_out[k] = s;
    }
  }
  if (threadIdx.x == 0) *m.sink_counter = 0u;
  if (m.sink_flag_remote) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(m.sink_flag_remote, m.sink_epoch);
  }
}

// ---- main rows: V rows per thread, P planes unrolled at compile time ------------------------------
template <int P>
__device__ __forceinline__ double row_scalar(const MatView &m, const Coefs &cf, const double *__restrict__ x,
                                             const double *__restrict__ ghost, long i) {
  int    c[P];
  double o[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    c[p] = ld_stream(m.col + p * m.ld + i);
    o[p] = ld_stream(m.off + p * m.ld + i);
  }
  const double xi = __ldg(x + i);
  double       d = 0.0;
  for (int g = 0; g < m.ND; ++g) d = fma(cf.cd[g], ld_stream(m.diag + g * m.ld + i), d);
  double acc = 0.0;
#pragma unroll
  for (int p = 0; p < P; ++p) acc = fma(cf.c[p] * o[p], fetch_x(x, ghost, c[p]), acc);
  return fma(-d, xi, acc);
}

template <int P>
__global__ void __launch_bounds__(kThreads) fsp_action_rows2(MatView m, Coefs cf, const double *__restrict__ x,
                                                             const double *__restrict__ ghost,
                                                             double *__restrict__ y, double *__restrict__ sink_out) {
  if ((int) blockIdx.x >= m.main_blocks) {
    sink_role(m, cf, x, y, sink_out);
    return;
  }
  const long i0 = 2 * ((long) blockIdx.x * kThreads + threadIdx.x);
  if (i0 + 1 < m.n) {
    int2    c[P];
    double2 o[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      c[p] = ld_stream(reinterpret_cast<const int2 *>(m.col + p * m.ld + i0));
      o[p] = ld_stream(reinterpret_cast<const double2 *>(m.off + p * m.ld + i0));
    }
    const double2 xi = __ldg(reinterpret_cast<const double2 *>(x + i0));
    double2       d = make_double2(0.0, 0.0);
    for (int g = 0; g < m.ND; ++g) {
      double2 dg = ld_stream(reinterpret_cast<const double2 *>(m.diag + g * m.ld + i0));
      d.x = fma(cf.cd[g], dg.x, d.x);
      d.y = fma(cf.cd[g], dg.y, d.y);
    }
    double2 acc = make_double2(0.0, 0.0);
#pragma unroll
    for (int p = 0; p < P; ++p) {
      acc.x = fma(cf.c[p] * o[p].x, fetch_x(x, ghost, c[p].x), acc.x);
      acc.y = fma(cf.c[p] * o[p].y, fetch_x(x, ghost, c[p].y), acc.y);
    }
    acc.x = fma(-d.x, xi.x, acc.x);
    acc.y = fma(-d.y, xi.y, acc.y);
    st_stream(reinterpret_cast<double2 *>(y + i0), acc);
  } else if (i0 < m.n) {
    y[i0] = row_scalar<P>(m, cf, x, ghost, i0);
  }
}

// 4 rows per thread: int4 column loads, 2 x double2 value loads per plane
template <int P>
__global__ void __launch_bounds__(kThreads) fsp_action_rows4(MatView m, Coefs cf, const double *__restrict__ x,
                                                             const double *__restrict__ ghost,
                                                             double *__restrict__ y, double *__restrict__ sink_out) {
  if ((int) blockIdx.x >= m.main_blocks) {
    sink_role(m, cf, x, y, sink_out);
    return;
  }
  const long i0 = 4 * ((long) blockIdx.x * kThreads + threadIdx.x);
  if (i0 + 3 < m.n) {
    int4    c[P];
    double2 oa[P], ob[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      c[p] = ld_stream(reinterpret_cast<const int4 *>(m.col + p * m.ld + i0));
      oa[p] = ld_stream(reinterpret_cast<const double2 *>(m.off + p * m.ld + i0));
      ob[p] = ld_stream(reinterpret_cast<const double2 *>(m.off + p * m.ld + i0 + 2));
    }
    const double2 xa = __ldg(reinterpret_cast<const double2 *>(x + i0));
    const double2 xb = __ldg(reinterpret_cast<const double2 *>(x + i0 + 2));
    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
    for (int g = 0; g < m.ND; ++g) {
      double2 da = ld_stream(reinterpret_cast<const double2 *>(m.diag + g * m.ld + i0));
      double2 db = ld_stream(reinterpret_cast<const double2 *>(m.diag + g * m.ld + i0 + 2));
      d0 = fma(cf.cd[g], da.x, d0);
      d1 = fma(cf.cd[g], da.y, d1);
      d2 = fma(cf.cd[g], db.x, d2);
      d3 = fma(cf.cd[g], db.y, d3);
    }
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      a0 = fma(cf.c[p] * oa[p].x, fetch_x(x, ghost, c[p].x), a0);
      a1 = fma(cf.c[p] * oa[p].y, fetch_x(x, ghost, c[p].y), a1);
      a2 = fma(cf.c[p] * ob[p].x, fetch_x(x, ghost, c[p].z), a2);
      a3 = fma(cf.c[p] * ob[p].y, fetch_x(x, ghost, c[p].w), a3);
    }
    st_stream(reinterpret_cast<double2 *>(y + i0), make_double2(fma(-d0, xa.x, a0), fma(-d1, xa.y, a1)));
    st_stream(reinterpret_cast<double2 *>(y + i0 + 2), make_double2(fma(-d2, xb.x, a2), fma(-d3, xb.y, a3)));
  } else {
    for (long i = i0; i < m.n && i < i0 + 4; ++i) y[i] = row_scalar<P>(m, cf, x, ghost, i);
  }
}

// 1 row per thread (also the path for unaligned x / y)
template <int P, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) fsp_action_rows1(MatView m, Coefs cf, const double *__restrict__ x,
                                                             const double *__restrict__ ghost,
                                                             double *__restrict__ y, double *__restrict__ sink_out) {
  if ((int) blockIdx.x >= m.main_blocks) {
    sink_role(m, cf, x, y, sink_out);
    return;
  }
  const long q = (long) blockIdx.x * kThreads + threadIdx.x;
  if (q < m.n) {
    const long i = m.row_list ? (long) m.row_list[q] : q;
    y[i] = row_scalar<P>(m, cf, x, ghost, i);
  }
}

// generic number of planes (P > 16): runtime loop
__global__ void __launch_bounds__(kThreads) fsp_action_generic(MatView m, Coefs cf, const double *__restrict__ x,
                                                                  const double *__restrict__ ghost,
                                                                  double *__restrict__ y,
                                                                  double *__restrict__ sink_out) {
  if ((int) blockIdx.x >= m.main_blocks) {
    sink_role(m, cf, x, y, sink_out);
    return;
  }
  const long q = (long) blockIdx.x * kThreads + threadIdx.x;
  if (q >= m.n) return;
  const long   i = m.row_list ? (long) m.row_list[q] : q;
  const double xi = __ldg(x + i);
  double       d = 0.0;
  for (int g = 0; g < m.ND; ++g) d = fma(cf.cd[g], ld_stream(m.diag + g * m.ld + i), d);
  double acc = 0.0;
  for (int p = 0; p < m.P; ++p) {
    int c = ld_stream(m.col + p * m.ld + i);
    acc = fma(cf.c[p] * ld_stream(m.off + p * m.ld + i), fetch_x(x, ghost, c), acc);
  }
  y[i] = fma(-d, xi, acc);
}


// Lean hot kernel: 1 row per thread, 32 registers -> 8 CTAs of 256 threads per SM (full occupancy), which matters
// because every thread does TWO dependent memory round trips (column index, then the gathered x entry).
// GHOST: 0 = no ghost columns exist (single GPU), 1 = ghost buffer valid, 2 = interior pass (ghost entries count 0),
//        3 = no ghost columns, rows [m.row0, m.n) only (chunks of the host-vector pipeline, fspmat_action_rows).
// One row of y = A(t) x, the body shared by the single-GPU kernel and the row CTAs of the fused multi-GPU kernel.
// GHOST: 0 = no ghost columns can occur; 1 = ghost entries are read from `ghost` at L2 (ld.cg: other GPUs stored them
// while this kernel was running) plus `poison` (0, or NaN after a timed-out wait); 2 = ghost entries count 0.
template <int P, int GHOST>
__device__ __forceinline__ double lean_row(const MatView &m, const Coefs &cf, const double *__restrict__ x,
                                           const double *ghost, double poison, int i) {
  const int    *cp = m.col + i;
  const double *op = m.off + i;
  double acc = 0.0;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int    c = ld_stream(cp + (size_t) p * m.ld);
    const double o = ld_stream(op + (size_t) p * m.ld);
    double xs = c >= 0 ? __ldg(x + c) : 0.0;
    if (GHOST == 1) xs += c <= -2 ? __ldcg(ghost + (-(c + 2))) + poison : 0.0;
    acc = fma(cf.c[p] * o, xs, acc);
  }
  double d = 0.0;
  const double *dp = m.diag + i;
  for (int g = 0; g < m.ND; ++g) d = fma(cf.cd[g], ld_stream(dp + (size_t) g * m.ld), d);
  return fma(-d, __ldg(x + i), acc);
}

template <int P, int GHOST>
__global__ void __launch_bounds__(kThreads, 8) fsp_action_lean(MatView m, Coefs cf, const double *__restrict__ x,
                                                                const double *__restrict__ ghost,
                                                                double *__restrict__ y, double *__restrict__ sink_out) {
  if ((int) blockIdx.x >= m.main_blocks) {
    sink_role(m, cf, x, y, sink_out);
    return;
  }
  const int i = (GHOST == 3 ? m.row0 : 0) + (int) blockIdx.x * kThreads + (int) threadIdx.x;
  if (i >= m.n) return;
  y[i] = lean_row<P, (GHOST == 1 ? 1 : 0)>(m, cf, x, ghost, 0.0, i);
}

// Action with fused epilogue (solver hot loops).  Same shape as the lean kernel (1 row per thread, 32 registers, 8
// CTAs/SM -- a grid-stride variant with few CTAs measured 1.6x slower, profiles/r01_solve_launches_summary.md); every
// CTA leaves one partial per inner product, the last-arriving sink CTA finishes the K sink rows with the same
// epilogue and leaves the partial of those rows in the extra slot; a fixed-shape reduction over the partials follows
// (fspvec_sum), so the results are deterministic.
template <int P>
__global__ void __launch_bounds__(kThreads, 8) fsp_action_epi(MatView m, Coefs cf, Epi e, const double *__restrict__ x,
                                                              double *__restrict__ y) {
  __shared__ double red[32];
  if ((int) blockIdx.x >= m.main_blocks) {
    // ---- sink rows: partial sums per chunk, then the last sink CTA finishes rows n..n+K-1 with the epilogue ----
    __shared__ bool is_last;
    const int       sb = blockIdx.x - m.main_blocks;
    double          acc = 0.0;
    {
      const long b = m.sb_begin[sb], en = m.sb_end[sb];
      for (long q = b + threadIdx.x; q < en; q += blockDim.x)
        acc = fma(ld_stream(m.sink_val + q), __ldg(x + ld_stream(m.sink_idx + q)), acc);
    }
    const double r = block_sum(acc, red);
    if (threadIdx.x == 0) m.sink_partials[sb] = r;
    __threadfence();
    if (threadIdx.x == 0) is_last = (atomicAdd(m.sink_counter, 1u) == (unsigned) m.sink_blocks - 1u);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = warp; k < m.K; k += nw) {
      double s = 0.0;
      for (int b = lane; b < m.sink_blocks; b += 32) {
        const int seg = m.sb_seg[b];
        if (seg >= 0 && seg % m.K == k) s = fma(cf.cd[seg / m.K], __ldcg(m.sink_partials + b), s);
      }
      s = warp_sum(s);
      if (lane == 0) {
        const long i = (long) m.n_rows_main + k;
        double     v = fma(e.alpha, s, e.beta * __ldg(x + i));
        if (e.scale) v *= __ldg(e.scale + i);
        y[i] = v;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      *m.sink_counter = 0u;
      double d0 = 0.0, d1 = 0.0;
      for (int k = 0; k < m.K; ++k) {  // fixed order
        const long   i = (long) m.n_rows_main + k;
        const double v = y[i];
        if (e.n_dots > 0) d0 = fma(v, e.vec[0] ? __ldg(e.vec[0] + i) : v, d0);
        if (e.n_dots > 1) d1 = fma(v, e.vec[1] ? __ldg(e.vec[1] + i) : v, d1);
      }
      if (e.n_dots > 0) e.partials[e.G] = d0;
      if (e.n_dots > 1) e.partials[(size_t) e.pstride + e.G] = d1;
    }
    return;
  }
  const int i = (int) blockIdx.x * kThreads + (int) threadIdx.x;
  double    d0 = 0.0, d1 = 0.0;
  if (i < m.n) {
    // the epilogue operands are requested FIRST, together with the matrix planes: issued after the gathers they would
    // add a third dependent memory round trip per row (measured: 212 us instead of ~172 us on 9.9 M rows)
    const double sc = e.scale ? __ldg(e.scale + i) : 1.0;
    const double w0 = (e.n_dots > 0 && e.vec[0]) ? __ldg(e.vec[0] + i) : 0.0;
    const double w1 = (e.n_dots > 1 && e.vec[1]) ? __ldg(e.vec[1] + i) : 0.0;
    const double xi = __ldg(x + i);
    const int    *cp = m.col + i;
    const double *op = m.off + i;
    double        acc = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const int    c = ld_stream(cp + (size_t) p * m.ld);
      const double o = ld_stream(op + (size_t) p * m.ld);
      const double xs = c >= 0 ? __ldg(x + c) : 0.0;
      acc = fma(cf.c[p] * o, xs, acc);
    }
    double        d = 0.0;
    const double *dp = m.diag + i;
    for (int g = 0; g < m.ND; ++g) d = fma(cf.cd[g], ld_stream(dp + (size_t) g * m.ld), d);
    const double v = fma(e.alpha, fma(-d, xi, acc), e.beta * xi) * sc;
    y[i] = v;
    if (e.n_dots > 0) d0 = v * (e.vec[0] ? w0 : v);
    if (e.n_dots > 1) d1 = v * (e.vec[1] ? w1 : v);
  }
  if (e.n_dots > 0) {
    const double r0 = block_sum(d0, red);
    if (threadIdx.x == 0) e.partials[blockIdx.x] = r0;
  }
  if (e.n_dots > 1) {
    const double r1 = block_sum(d1, red);
    if (threadIdx.x == 0) e.partials[(size_t) e.pstride + blockIdx.x] = r1;
  }
}

typedef void (*epi_fn)(MatView, Coefs, Epi, const double *, double *);
epi_fn pick_epi(int P) {
  switch (P) {
#define FSP_CASE(N) case N: return fsp_action_epi<N>;
    FSP_CASE(1) FSP_CASE(2) FSP_CASE(3) FSP_CASE(4) FSP_CASE(5) FSP_CASE(6) FSP_CASE(7) FSP_CASE(8)
    FSP_CASE(9) FSP_CASE(10) FSP_CASE(11) FSP_CASE(12) FSP_CASE(13) FSP_CASE(14) FSP_CASE(15) FSP_CASE(16)
#undef FSP_CASE
    default: return nullptr;
  }
}

// boundary rows of the split multi-GPU action (short list; runtime plane loop)
__global__ void __launch_bounds__(kThreads) fsp_action_rowlist(MatView m, Coefs cf, const double *__restrict__ x,
                                                               const double *__restrict__ ghost,
                                                               double *__restrict__ y, double *__restrict__ sink_out) {
  const long q = (long) blockIdx.x * kThreads + threadIdx.x;
  if (q >= m.n) return;
  const long   i = (long) m.row_list[q];
  const double xi = __ldg(x + i);
  double       d = 0.0;
  for (int g = 0; g < m.ND; ++g) d = fma(cf.cd[g], ld_stream(m.diag + g * m.ld + i), d);
  double acc = 0.0;
  for (int p = 0; p < m.P; ++p) {
    int c = ld_stream(m.col + p * m.ld + i);
    acc = fma(cf.c[p] * ld_stream(m.off + p * m.ld + i), fetch_x(x, ghost, c), acc);
  }
  y[i] = fma(-d, xi, acc);
}

// Boundary pass of the peer-memory multi-GPU action, fused with the arrival wait: every CTA first waits (device code,
// acquire at system scope) until every peer's push kernel has published this epoch, then recomputes its rows with the
// ghost entries the peers stored into this GPU's window.  One extra CTA on the sink owner adds the K x n_ranks partial
// sink sums in rank order (deterministic) into y[n..n+K).
__global__ void __launch_bounds__(kThreads) fsp_action_boundary_p2p_kernel(MatView m, Coefs cf, P2PWait w,
                                                                           const double *__restrict__ x,
                                                                           const double *ghost, double *__restrict__ y) {
  if ((int) blockIdx.x >= m.main_blocks) {
    bool ok = true;
    if ((int) threadIdx.x < w.n_ranks) ok = wait_flag(w.sink_flags + threadIdx.x, w.epoch, w.err);
    ok = __syncthreads_and(ok);
    if ((int) threadIdx.x < m.K) {
      double s = 0.0;
      for (int p = 0; p < w.n_ranks; ++p) s += __ldcg(w.sink_slots + (size_t) p * FSP_P2P_MAX_SINKS + threadIdx.x);
      y[m.n_rows_main + threadIdx.x] = ok ? s : __longlong_as_double(0x7ff8000000000000ll);  // time-out: poison
    }
    return;
  }
  bool ok = true;
  if ((int) threadIdx.x < w.n_ranks && (int) threadIdx.x != w.rank) ok = wait_flag(w.halo_flags + threadIdx.x, w.epoch, w.err);
  ok = __syncthreads_and(ok);
  const long q = (long) blockIdx.x * kThreads + threadIdx.x;
  if (q >= m.n) return;
  const long   i = (long) m.row_list[q];
  if (!ok) { y[i] = __longlong_as_double(0x7ff8000000000000ll); return; }  // time-out: poison the rows that need the halo
  const double xi = __ldg(x + i);
  double       d = 0.0;
  for (int g = 0; g < m.ND; ++g) d = fma(cf.cd[g], ld_stream(m.diag + g * m.ld + i), d);
  double acc = 0.0;
  for (int p = 0; p < m.P; ++p) {
    const int c = ld_stream(m.col + p * m.ld + i);
    // ghost entries were written by other GPUs during this kernel's lifetime: read them at L2 (ld.cg), never L1
    const double xs = c >= 0 ? __ldg(x + c) : (c == -1 ? 0.0 : __ldcg(ghost + (-(c + 2))));
    acc = fma(cf.c[p] * ld_stream(m.off + p * m.ld + i), xs, acc);
  }
  y[i] = fma(-d, xi, acc);
}

// Whole multi-GPU action in ONE kernel (peer-memory mode).  CTAs are issued in the order cta_order[]: first the CTAs
// whose 256 rows reference no ghost entry (they overlap the arrival of the halo), then the CTAs with ghost rows, which
// wait in device code for the peers' epoch flags before they start; one trailing CTA on the sink owner adds the
// partial sink sums of all ranks in rank order.  No second pass over the boundary rows, no kernel boundary between
// "interior" and "boundary" work, nothing for the host to do between them.
template <int P>
__global__ void __launch_bounds__(kThreads, 8) fsp_action_p2p_kernel(MatView m, Coefs cf, P2PWait w,
                                                                     const double *__restrict__ x, const double *ghost,
                                                                     double *__restrict__ y) {
  if ((int) blockIdx.x >= m.main_blocks) {
    bool ok = true;
    if ((int) threadIdx.x < w.n_ranks) ok = wait_flag(w.sink_flags + threadIdx.x, w.epoch, w.err);
    ok = __syncthreads_and(ok);
    if ((int) threadIdx.x < m.K) {
      double s = 0.0;
      for (int p = 0; p < w.n_ranks; ++p) s += __ldcg(w.sink_slots + (size_t) p * FSP_P2P_MAX_SINKS + threadIdx.x);
      y[m.n_rows_main + threadIdx.x] = ok ? s : __longlong_as_double(0x7ff8000000000000ll);  // time-out: poison
    }
    return;
  }
  const int cta = m.cta_order[blockIdx.x];
  if ((int) blockIdx.x >= m.n_interior_ctas) {
    bool ok = true;
    if ((int) threadIdx.x < w.n_ranks && (int) threadIdx.x != w.rank) ok = wait_flag(w.halo_flags + threadIdx.x, w.epoch, w.err);
    if (!__syncthreads_and(ok)) {  // time-out: poison the rows that need the halo
      const int j = cta * kThreads + (int) threadIdx.x;
      if (j < m.n) y[j] = __longlong_as_double(0x7ff8000000000000ll);
      return;
    }
  }
  const int i = cta * kThreads + (int) threadIdx.x;
  if (i >= m.n) return;
  const int    *cp = m.col + i;
  const double *op = m.off + i;
  double acc = 0.0;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int    c = ld_stream(cp + (size_t) p * m.ld);
    const double o = ld_stream(op + (size_t) p * m.ld);
    // ghost entries were stored by other GPUs while this kernel was running: read them at L2 (ld.cg), never L1
    const double xs = c >= 0 ? __ldg(x + c) : (c == -1 ? 0.0 : __ldcg(ghost + (-(c + 2))));
    acc = fma(cf.c[p] * o, xs, acc);
  }
  double d = 0.0;
  const double *dp = m.diag + i;
  for (int g = 0; g < m.ND; ++g) d = fma(cf.cd[g], ld_stream(dp + (size_t) g * m.ld), d);
  y[i] = fma(-d, __ldg(x + i), acc);
}
// ---- the whole multi-GPU Action in ONE launch (peer-memory mode, default) ----------------------------------------
// Block roles in blockIdx order: [push CTAs][sink-partial CTAs][row CTAs, rotated][1 finishing CTA].
//   push     pack + store to the peers' ghost windows + epoch flag (push_role): first in the grid, so the halo is in
//            flight for the whole duration of the row pass
//   sinks    K partial sums of this rank -> the owner's slot row + flag
//   rows     the lean row code (1 row per thread, <= 32 registers, 8 CTAs/SM); CTA b handles rows of CTA (b + rot) mod
//            n, where rot was chosen at generate time so that the longest circular run of ghost-free CTAs comes first
//            (lattice blocks: both boundary planes end up at the tail).  No lookup table: a dependent load at the
//            start of every CTA is a third memory round trip per row and costs 30 % (measured with cta_order[] above).
//            The first n_fast CTAs of that order hold no ghost column and run exactly the single-GPU row code; the
//            others first wait for the peers' flags.  Every row is computed exactly once.
//   finish   waits for every peer's halo flag (this is what paces the reuse of the two ghost buffers: a rank can only
//            start epoch e+2 after all peers published e+1, i.e. finished reading e) and, on the sink owner, adds the
//            partial sums in rank order (deterministic) into y[n..n+K).
// Deadlock freedom: push and sink CTAs never wait; waiting CTAs only wait for PEERS' push CTAs, which run as soon as
// the peer's kernel starts, whatever the order in which the hardware issues CTAs.
struct HaloView {
  PushView                  push;
  const unsigned long long *halo_flags;
  const unsigned long long *sink_flags;
  const double             *sink_slots;
  const double             *ghost;
  double                   *sink_slot_remote;
  unsigned long long       *sink_flag_remote;
  unsigned int             *err;
  int                       finish_sinks;   // this rank owns y[n..n+K)
  int                       n_push;         // leading push CTAs of this launch (0: the push is not part of it)
  const double             *push_src;       // null: the push CTAs gather x[send_idx[q]]; else they read this packed buffer
  int                       rot;            // row CTA b handles the rows of CTA (b + rot) mod n
  int                       n_fast;         // the first n_fast row CTAs (in that order) hold no ghost column
};

template <int P>
__global__ void __launch_bounds__(kThreads, 8) fsp_action_halo_kernel(MatView m, Coefs cf, HaloView hv,
                                                                      const double *__restrict__ x,
                                                                      double *__restrict__ y) {
  int b = (int) blockIdx.x;
  if (b < hv.n_push) { push_role(hv.push, hv.push_src ? hv.push_src : x, b); return; }
  b -= hv.n_push;
  if (b < m.sink_blocks) {
    // partial sink sums of this rank; the last-arriving CTA stores the K sums into the owner's slots and signals
    __shared__ double smem[32];
    __shared__ bool   is_last;
    double            acc = 0.0;
    for (long q = m.sb_begin[b] + threadIdx.x; q < m.sb_end[b]; q += blockDim.x)
      acc = fma(ld_stream(m.sink_val + q), __ldg(x + ld_stream(m.sink_idx + q)), acc);
    const double r = block_sum(acc, smem);
    if (threadIdx.x == 0) m.sink_partials[b] = r;
    __threadfence();
    if (threadIdx.x == 0) is_last = (atomicAdd(m.sink_counter, 1u) == (unsigned) m.sink_blocks - 1u);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = warp; k < m.K; k += nw) {
      double s = 0.0;
      for (int q = lane; q < m.sink_blocks; q += 32) {
        const int seg = m.sb_seg[q];
        if (seg >= 0 && seg % m.K == k) s = fma(cf.cd[seg / m.K], __ldcg(m.sink_partials + q), s);
      }
      s = warp_sum(s);
      if (lane == 0) hv.sink_slot_remote[k] = s;
    }
    if (threadIdx.x == 0) *m.sink_counter = 0u;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(hv.sink_flag_remote, hv.push.epoch);
    return;
  }
  b -= m.sink_blocks;
  if (b >= m.main_blocks) {
    // finishing CTA
    bool ok = true;
    if ((int) threadIdx.x < hv.push.size && (int) threadIdx.x != hv.push.rank)
      ok = wait_flag(hv.halo_flags + threadIdx.x, hv.push.epoch, hv.err);
    if (hv.finish_sinks && (int) threadIdx.x < hv.push.size)
      ok = wait_flag(hv.sink_flags + threadIdx.x, hv.push.epoch, hv.err) && ok;
    ok = __syncthreads_and(ok);
    if (hv.finish_sinks && (int) threadIdx.x < m.K) {
      double s = 0.0;
      for (int p = 0; p < hv.push.size; ++p) s += __ldcg(hv.sink_slots + (size_t) p * FSP_P2P_MAX_SINKS + threadIdx.x);
      y[m.n_rows_main + threadIdx.x] = ok ? s : __longlong_as_double(0x7ff8000000000000ll);  // time-out: poison
    }
    return;
  }
  int cta = b + hv.rot;
  if (cta >= m.main_blocks) cta -= m.main_blocks;
  const int i = m.row0 + cta * kThreads + (int) threadIdx.x;
  if (b < hv.n_fast) {
    // ghost-free CTAs (by construction of rot / n_fast at generate time): exactly the single-GPU row code
    if (i < m.n) y[i] = lean_row<P, 0>(m, cf, x, nullptr, 0.0, i);
    return;
  }
  // CTAs that may hold ghost columns: wait for the peers' flags, then the same row code with the ghost window
  bool ok = true;
  if ((int) threadIdx.x < hv.push.size && (int) threadIdx.x != hv.push.rank)
    ok = wait_flag(hv.halo_flags + threadIdx.x, hv.push.epoch, hv.err);
  ok = __syncthreads_and(ok);
  const double poison = ok ? 0.0 : __longlong_as_double(0x7ff8000000000000ll);  // time-out: NaN into the rows that need the halo
  if (i < m.n) y[i] = lean_row<P, 1>(m, cf, x, hv.ghost, poison, i);
}
typedef void (*halo_fn)(MatView, Coefs, HaloView, const double *, double *);
halo_fn pick_halo(int P) {
  switch (P) {
#define FSP_CASE(N) case N: return fsp_action_halo_kernel<N>;
    FSP_CASE(1) FSP_CASE(2) FSP_CASE(3) FSP_CASE(4) FSP_CASE(5) FSP_CASE(6) FSP_CASE(7) FSP_CASE(8)
    FSP_CASE(9) FSP_CASE(10) FSP_CASE(11) FSP_CASE(12) FSP_CASE(13) FSP_CASE(14) FSP_CASE(15) FSP_CASE(16)
#undef FSP_CASE
    default: return nullptr;
  }
}

typedef void (*p2p_fn)(MatView, Coefs, P2PWait, const double *, const double *, double *);
p2p_fn pick_p2p(int P) {
  switch (P) {
#define FSP_CASE(N) case N: return fsp_action_p2p_kernel<N>;
    FSP_CASE(1) FSP_CASE(2) FSP_CASE(3) FSP_CASE(4) FSP_CASE(5) FSP_CASE(6) FSP_CASE(7) FSP_CASE(8)
    FSP_CASE(9) FSP_CASE(10) FSP_CASE(11) FSP_CASE(12) FSP_CASE(13) FSP_CASE(14) FSP_CASE(15) FSP_CASE(16)
#undef FSP_CASE
    default: return nullptr;
  }
}

typedef void (*action_fn)(MatView, Coefs, const double *, const double *, double *, double *);

template <int P>
action_fn pick_variant(int rows_per_thread) {
  switch (rows_per_thread) {
    case 10: return fsp_action_lean<P, 0>;
    case 11: return fsp_action_lean<P, 1>;
    case 12: return fsp_action_lean<P, 2>;
    case 13: return fsp_action_lean<P, 3>;
    case 1: return fsp_action_rows1<P, 1>;
    case 3: return fsp_action_rows1<P, 8>;  // register-capped (32 regs, full occupancy, may spill for P >= 5)
    case 4: return fsp_action_rows4<P>;
    default: return fsp_action_rows2<P>;
  }
}

action_fn pick_kernel(int P, int rows_per_thread) {
  switch (P) {
#define FSP_CASE(N) case N: return pick_variant<N>(rows_per_thread);
    FSP_CASE(1) FSP_CASE(2) FSP_CASE(3) FSP_CASE(4) FSP_CASE(5) FSP_CASE(6) FSP_CASE(7) FSP_CASE(8)
    FSP_CASE(9) FSP_CASE(10) FSP_CASE(11) FSP_CASE(12) FSP_CASE(13) FSP_CASE(14) FSP_CASE(15) FSP_CASE(16)
#undef FSP_CASE
    default: return nullptr;
  }
}

// ---- generate-time packing kernels -----------------------------------------------------------------
__global__ void pack_planes_kernel(int n, int P, long ld_in, long ld, const int *__restrict__ col_in,
                                   const double *__restrict__ off_in, int *__restrict__ col,
                                   double *__restrict__ off) {
  long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  int  p = blockIdx.y;
  if (i >= ld) return;
  int    c = -1;
  double o = 0.0;
  if (i < n) {
    c = col_in[p * ld_in + i];
    o = c == -1 ? 0.0 : off_in[p * ld_in + i];  // dropped entries must not inject NaN/Inf
  }
  col[p * ld + i] = c;
  off[p * ld + i] = o;
}

__global__ void pack_diag_kernel(int n, int n_tv, int n_ti, long ld_in, long ld, const double *__restrict__ diag_in,
                                 double *__restrict__ diag) {
  long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ld) return;
  for (int g = 0; g < n_tv; ++g) diag[g * ld + i] = i < n ? diag_in[g * ld_in + i] : 0.0;
  if (n_ti > 0) {
    // merged TI diagonal, summed in ti_reactions_ order like the reference's ADD_VALUES
    // (src/Matrix/FspMatrixBase.cpp:229-243): ((-d0) + (-d1)) + ... == -((d0 + d1) + ...)
    double s = 0.0;
    if (i < n)
      for (int q = 0; q < n_ti; ++q) s += diag_in[(n_tv + q) * ld_in + i];
    diag[n_tv * ld + i] = s;
  }
}

// nnz bookkeeping for GetLocalMVFlops (src/Matrix/FspMatrixBase.cpp:429-444): counts[0..n_tv) = stored
// off-diagonal entries per TV matrix, counts[n_tv] = distinct off-diagonal columns per row of the merged
// TI matrix (ADD_VALUES merges entries landing on the same (i, j)).
__global__ void count_nnz_kernel(int n, int n_tv, int n_ti, long ld, const int *__restrict__ col,
                                 unsigned long long *__restrict__ counts) {
  long               i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long local[kMaxPlanes + 1];
  for (int g = 0; g <= n_tv; ++g) local[g] = 0;
  if (i < n) {
    for (int g = 0; g < n_tv; ++g) {
      int c = col[g * ld + i];
      local[g] += (c != -1 && c != (int) i) ? 1 : 0;
    }
    for (int q = 0; q < n_ti; ++q) {
      int c = col[(n_tv + q) * ld + i];
      if (c == -1 || c == (int) i) continue;
      bool dup = false;
      for (int q2 = 0; q2 < q; ++q2) dup |= (col[(n_tv + q2) * ld + i] == c);
      local[n_tv] += dup ? 0 : 1;
    }
  }
  for (int g = 0; g <= n_tv; ++g) {
    unsigned long long v = local[g];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&counts[g], v);
  }
}


struct OutOfRange {
  int lo, hi;
  __host__ __device__ bool operator()(const int &c) const { return c >= 0 && (c < lo || c >= hi); }
};
__global__ void remap_cols_kernel(int *col, long n, int lo, int hi, const int *ghost, long n_ghost) {
  long q = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  int c = col[q];
  if (c < 0) return;
  if (c >= lo && c < hi) { col[q] = c - lo; return; }
  long a = 0, b = n_ghost;  // lower_bound in the sorted ghost list
  while (a < b) {
    long mid = (a + b) >> 1;
    if (ghost[mid] < c) a = mid + 1; else b = mid;
  }
  col[q] = -((int) a + 2);
}
__global__ void mark_cta_kernel(const int *__restrict__ rows, long n_rows, int *__restrict__ cta_flag) {
  long q = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n_rows) cta_flag[rows[q] / kThreads] = 1;
}
struct CtaIsInterior {
  const int *flag;
  __host__ __device__ bool operator()(const int &c) const { return flag[c] == 0; }
};
// per row-chunk maximum of the referenced column indices (what part of x a chunk of rows needs): one atomicMax per warp
__global__ void chunk_max_col_kernel(int n, int P, long ld, const int *__restrict__ col, int chunk_rows, int *__restrict__ chunk_max) {
  const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  int        mx = -1;
  if (i < n) {
    mx = (int) i;  // the diagonal term reads x_i
    for (int p = 0; p < P; ++p) mx = max(mx, col[p * ld + i]);
  }
  const int c0 = (int) (min((long) n - 1, (long) blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / chunk_rows);
  const int c1 = (int) (min((long) n - 1, (long) blockIdx.x * blockDim.x + (threadIdx.x | 31)) / chunk_rows);
  if (c0 == c1) {
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx >= 0) atomicMax(chunk_max + c0, mx);
  } else if (i < n) {
    atomicMax(chunk_max + (int) (i / chunk_rows), mx);
  }
}
__global__ void mark_chunk_kernel(const int *__restrict__ rows, long n_rows, long chunk_rows, int *__restrict__ flag) {
  long q = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n_rows) flag[rows[q] / chunk_rows] = 1;
}
// ---- assembled Jacobian in CSR form (CreateRHSJacobian / ComputeRHSJacobian) ---------------------------------------
// state row i: slot 0 = diagonal, slot 1 + p = plane p (an absent neighbour keeps column i with value 0); P + 1 slots
__global__ void csr_state_rows_kernel(MatView m, Coefs cf, int structure, int *__restrict__ row_ptr, int *__restrict__ col,
                                      double *__restrict__ val) {
  const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.n) return;
  const long base = i * (m.P + 1);
  double     d = 0.0;
  for (int g = 0; g < m.ND; ++g) d = fma(cf.cd[g], m.diag[(size_t) g * m.ld + i], d);
  if (structure) { row_ptr[i] = (int) base; col[base] = (int) i; }
  val[base] = -d;
  for (int p = 0; p < m.P; ++p) {
    const int c = m.col[(size_t) p * m.ld + i];
    if (structure) col[base + 1 + p] = c >= 0 ? c : (int) i;
    val[base + 1 + p] = c >= 0 ? cf.c[p] * m.off[(size_t) p * m.ld + i] : 0.0;
  }
}
__global__ void csr_sink_segment_kernel(const int *__restrict__ idx, const double *__restrict__ v, long count, double coef,
                                        int structure, int *__restrict__ col, double *__restrict__ val) {
  const long q = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= count) return;
  if (structure) col[q] = idx[q];
  val[q] = coef * v[q];
}
__global__ void csr_spmv_kernel(int n_rows, const int *__restrict__ row_ptr, const int *__restrict__ col,
                                const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  double acc = 0.0;
  for (int q = row_ptr[i]; q < row_ptr[i + 1]; ++q) acc = fma(val[q], __ldg(x + col[q]), acc);
  y[i] = acc;
}
struct RowHasGhost {
  const int *col; long ld; int P;
  __host__ __device__ bool operator()(const int &i) const {
    for (int p = 0; p < P; ++p)
      if (col[p * ld + i] <= -2) return true;
    return false;
  }
};
__global__ void shift_idx_kernel(int *idx, long n, int delta) {
  long q = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) idx[q] += delta;
}

}  // namespace

// -----------------------------------------------------------------------------------------------------
struct fspmat_s {
  bool has_values = false;
  int  n = 0, n_rows = 0, R = 0, n_tv = 0, n_ti = 0, P = 0, ND = 0, K = 0, owns_sinks = 0;
  long n_ghost = 0, ld = 0;
  std::vector<int> tv, ti;
  int     *d_col = nullptr;
  double  *d_off = nullptr;
  double  *d_diag = nullptr;
  // sinks (merged into G = ND groups)
  long     sink_nnz = 0;
  int     *d_sink_idx = nullptr;
  double  *d_sink_val = nullptr;
  int      sink_blocks = 0;
  int     *d_sb_seg = nullptr;
  long    *d_sb_begin = nullptr, *d_sb_end = nullptr;
  double  *d_sink_partials = nullptr;
  unsigned *d_sink_counter = nullptr;
  std::vector<long> seg_ptr;  // [G*K + 1] host copy of merged segments
  long     flops = 0;
  double   bytes = 0.0;
  int      variant = 0;
  double   ti_coef = 1.0;              // coefficient of the time-invariant reactions (1; 0 while a time derivative of A(t) is applied)
  int     *d_cta_order = nullptr;      // CTA issue order of the single-kernel peer-memory action
  int      n_ctas = 0, n_interior_ctas = 0;
  int      rot = 0, n_fast = 0;        // fused halo action: CTA rotation (longest ghost-free run first) and its length
  double  *d_epi_partials = nullptr;   // fused-epilogue inner-product partials (allocated on first use)
  int     *d_boundary_rows = nullptr;  // rows referencing ghost entries (multi-GPU)
  long     n_boundary = 0;
};

static int free_values(fspmat_s *h) {
  pfree(h->d_col); pfree(h->d_off); pfree(h->d_diag);
  pfree(h->d_sink_idx); pfree(h->d_sink_val);
  pfree(h->d_sb_seg); pfree(h->d_sb_begin); pfree(h->d_sb_end);
  pfree(h->d_sink_partials); pfree(h->d_sink_counter); pfree(h->d_boundary_rows);
  pfree(h->d_epi_partials); pfree(h->d_cta_order);
  int variant = h->variant;
  *h = fspmat_s();
  h->variant = variant;
  return 0;
}

extern "C" {

int fspmat_create(fspmat_t *out) {
  *out = new fspmat_s();
  return 0;
}

int fspmat_destroy(fspmat_t h) {
  if (!h) return 0;
  free_values(h);
  delete h;
  return 0;
}

int fspmat_clear(fspmat_t h) { return free_values(h); }

int fspmat_set_ti_coef(fspmat_t h, double c) {
  h->ti_coef = c;
  return 0;
}

int fspmat_set_variant(fspmat_t h, int variant) {
  h->variant = variant;
  return 0;
}

int fspmat_generate(fspmat_t h, const fspmat_desc *d) {
  free_values(h);
  const int P = d->n_tv + d->n_ti;
  if (P > kMaxPlanes) { set_error("fspmat_generate: %d reactions exceed the supported %d", P, kMaxPlanes); return -1; }
  if (d->n_states < 0 || d->n_rows < d->n_states) { set_error("fspmat_generate: bad sizes"); return -1; }
  h->n = d->n_states; h->n_rows = d->n_rows; h->R = d->n_reactions;
  h->n_tv = d->n_tv; h->n_ti = d->n_ti; h->P = P; h->ND = d->n_tv + (d->n_ti > 0 ? 1 : 0);
  h->K = d->n_constr; h->owns_sinks = d->owns_sinks; h->n_ghost = d->n_ghost;
  h->tv.assign(d->tv_reactions, d->tv_reactions + d->n_tv);
  h->ti.assign(d->ti_reactions, d->ti_reactions + d->n_ti);
  const long n = h->n;
  h->ld = ((n + 31) / 32) * 32;
  if (h->ld == 0) h->ld = 32;
  const long ld = h->ld;
  cudaStream_t st = 0;

  if (P > 0) {
    FSP_CUDA_CHECK(pmalloc(&h->d_col, sizeof(int) * P * ld));
    FSP_CUDA_CHECK(pmalloc(&h->d_off, sizeof(double) * P * ld));
    FSP_CUDA_CHECK(pmalloc(&h->d_diag, sizeof(double) * std::max(h->ND, 1) * ld));
    const int *col_in = d->col; const double *off_in = d->off, *diag_in = d->diag;
    int *t_col = nullptr; double *t_off = nullptr, *t_diag = nullptr;
    long ld_in = d->ld;
    if (!d->arrays_on_device) {
      // stage through the device with a dense leading dimension
      FSP_CUDA_CHECK(pmalloc(&t_col, sizeof(int) * P * std::max(n, 1L)));
      FSP_CUDA_CHECK(pmalloc(&t_off, sizeof(double) * P * std::max(n, 1L)));
      FSP_CUDA_CHECK(pmalloc(&t_diag, sizeof(double) * P * std::max(n, 1L)));
      if (n > 0) {
        FSP_CUDA_CHECK(cudaMemcpy2D(t_col, sizeof(int) * n, d->col, sizeof(int) * d->ld, sizeof(int) * n, P, cudaMemcpyHostToDevice));
        FSP_CUDA_CHECK(cudaMemcpy2D(t_off, sizeof(double) * n, d->off, sizeof(double) * d->ld, sizeof(double) * n, P, cudaMemcpyHostToDevice));
        FSP_CUDA_CHECK(cudaMemcpy2D(t_diag, sizeof(double) * n, d->diag, sizeof(double) * d->ld, sizeof(double) * n, P, cudaMemcpyHostToDevice));
      }
      col_in = t_col; off_in = t_off; diag_in = t_diag; ld_in = n;
    }
    dim3 grid((unsigned) ((ld + 255) / 256), P);
    pack_planes_kernel<<<grid, 256, 0, st>>>((int) n, P, ld_in, ld, col_in, off_in, h->d_col, h->d_off);
    FSP_LAUNCH_CHECK();
    pack_diag_kernel<<<(unsigned) ((ld + 255) / 256), 256, 0, st>>>((int) n, h->n_tv, h->n_ti, ld_in, ld, diag_in, h->d_diag);
    FSP_LAUNCH_CHECK();
    FSP_CUDA_CHECK(cudaStreamSynchronize(st));
    pfree(t_col); pfree(t_off); pfree(t_diag);
  }

  // ---- multi-GPU: list of rows that reference ghost entries (redone after the halo exchange) ----------
  if (h->n_ghost > 0 && P > 0 && n > 0) {
    int   *d_num = nullptr;
    void  *d_tmp = nullptr;
    size_t need = 0;
    FSP_CUDA_CHECK(pmalloc(&d_num, sizeof(int)));
    FSP_CUDA_CHECK(pmalloc(&h->d_boundary_rows, sizeof(int) * n));
    cub::CountingInputIterator<int> iota(0);
    RowHasGhost pred{h->d_col, ld, P};
    cub::DeviceSelect::If(nullptr, need, iota, h->d_boundary_rows, d_num, (int) n, pred);
    FSP_CUDA_CHECK(pmalloc(&d_tmp, need));
    FSP_CUDA_CHECK(cub::DeviceSelect::If(d_tmp, need, iota, h->d_boundary_rows, d_num, (int) n, pred));
    count_launch();
    int nb = 0;
    FSP_CUDA_CHECK(cudaMemcpy(&nb, d_num, sizeof(int), cudaMemcpyDeviceToHost));
    h->n_boundary = nb;
    pfree(d_num); pfree(d_tmp);
  }
  if (h->n_ghost > 0 && P > 0 && n > 0) {
    // CTA issue order of the single-kernel peer-memory action: CTAs free of ghost rows first
    const int n_ctas = (int) ((n + kThreads - 1) / kThreads);
    int      *d_flag = nullptr, *d_num = nullptr;
    void     *d_tmp = nullptr;
    size_t    need = 0;
    FSP_CUDA_CHECK(pmalloc(&d_flag, sizeof(int) * n_ctas));
    FSP_CUDA_CHECK(pmalloc(&d_num, sizeof(int)));
    FSP_CUDA_CHECK(pmalloc(&h->d_cta_order, sizeof(int) * n_ctas));
    FSP_CUDA_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(int) * n_ctas, st));
    if (h->n_boundary > 0) {
      mark_cta_kernel<<<(unsigned) ((h->n_boundary + 255) / 256), 256, 0, st>>>(h->d_boundary_rows, h->n_boundary, d_flag);
      FSP_LAUNCH_CHECK();
    }
    {
      // rotation of the fused halo action: start right after the ghost CTA that precedes the longest circular run of
      // ghost-free CTAs, so that this run comes first and the CTAs that wait for the peers come last
      std::vector<int> flag((size_t) n_ctas);
      FSP_CUDA_CHECK(cudaMemcpyAsync(flag.data(), d_flag, sizeof(int) * n_ctas, cudaMemcpyDeviceToHost, st));
      FSP_CUDA_CHECK(cudaStreamSynchronize(st));
      int best_len = -1, best_start = 0, run = 0;
      for (int q = 0; q < 2 * n_ctas; ++q) {
        const int cta = q % n_ctas;
        if (flag[(size_t) cta]) { run = 0; continue; }
        ++run;
        if (run > n_ctas) run = n_ctas;
        if (run > best_len) { best_len = run; best_start = (q - run + 1 + n_ctas) % n_ctas; }
      }
      h->rot = best_len > 0 ? best_start : 0;
      h->n_fast = best_len > 0 ? best_len : 0;
    }
    cub::CountingInputIterator<int> iota(0);
    CtaIsInterior pred{d_flag};
    cub::DevicePartition::If(nullptr, need, iota, h->d_cta_order, d_num, n_ctas, pred, st);
    FSP_CUDA_CHECK(pmalloc(&d_tmp, need));
    FSP_CUDA_CHECK(cub::DevicePartition::If(d_tmp, need, iota, h->d_cta_order, d_num, n_ctas, pred, st));
    count_launch();
    int ni = 0;
    FSP_CUDA_CHECK(cudaMemcpy(&ni, d_num, sizeof(int), cudaMemcpyDeviceToHost));
    h->n_ctas = n_ctas;
    h->n_interior_ctas = ni;
    pfree(d_flag); pfree(d_num); pfree(d_tmp);
  }

  // ---- flops: 2 nnz per matrix (+ rows per TV axpy); FspMatrixBase.cpp:429-444 -------------------
  std::vector<unsigned long long> counts(h->n_tv + 1, 0ull);
  if (P > 0 && n > 0) {
    unsigned long long *d_counts;
    FSP_CUDA_CHECK(pmalloc(&d_counts, sizeof(unsigned long long) * (h->n_tv + 1)));
    FSP_CUDA_CHECK(cudaMemset(d_counts, 0, sizeof(unsigned long long) * (h->n_tv + 1)));
    count_nnz_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>((int) n, h->n_tv, h->n_ti, ld, h->d_col, d_counts);
    FSP_LAUNCH_CHECK();
    FSP_CUDA_CHECK(cudaMemcpy(counts.data(), d_counts, sizeof(unsigned long long) * (h->n_tv + 1), cudaMemcpyDeviceToHost));
    pfree(d_counts);
  }
  long flops = 0;
  if (h->n_ti > 0) flops += 2 * ((long) counts[h->n_tv] + n);
  for (int g = 0; g < h->n_tv; ++g) flops += 2 * ((long) counts[g] + n) + h->n_rows;

  // ---- sinks: merge TI planes' segments into one group per constraint ----------------------------
  const int K = h->K, G = h->ND;
  long total_sink = 0;
  if (K > 0) {
    const long *sp = d->sink_ptr;
    long nnz_in = sp ? sp[(long) P * K] : 0;
    std::vector<int>    idx_in(nnz_in);
    std::vector<double> val_in(nnz_in);
    if (nnz_in > 0) {
      cudaMemcpyKind kind = d->arrays_on_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost;
      FSP_CUDA_CHECK(cudaMemcpy(idx_in.data(), d->sink_idx, sizeof(int) * nnz_in, kind));
      FSP_CUDA_CHECK(cudaMemcpy(val_in.data(), d->sink_val, sizeof(double) * nnz_in, kind));
    }
    std::vector<int>    idx; idx.reserve(nnz_in);
    std::vector<double> val; val.reserve(nnz_in);
    h->seg_ptr.assign((size_t) G * K + 1, 0);
    long ti_sink_distinct = 0;
    std::vector<long> tv_sink_nnz(h->n_tv, 0);
    for (int g = 0; g < G; ++g) {
      for (int k = 0; k < K; ++k) {
        h->seg_ptr[(size_t) g * K + k] = (long) idx.size();
        if (g < h->n_tv) {
          for (long q = sp[(long) g * K + k]; q < sp[(long) g * K + k + 1]; ++q) { idx.push_back(idx_in[q]); val.push_back(val_in[q]); }
          tv_sink_nnz[g] += sp[(long) g * K + k + 1] - sp[(long) g * K + k];
        } else {
          size_t start = idx.size();
          for (int q2 = 0; q2 < h->n_ti; ++q2) {
            int p = h->n_tv + q2;
            for (long q = sp[(long) p * K + k]; q < sp[(long) p * K + k + 1]; ++q) { idx.push_back(idx_in[q]); val.push_back(val_in[q]); }
          }
          // distinct (k, i) pairs, as PETSc's ADD_VALUES would store them (FspMatrixConstrained.cpp:220-240)
          std::vector<int> tmp(idx.begin() + start, idx.end());
          std::sort(tmp.begin(), tmp.end());
          ti_sink_distinct += (long) (std::unique(tmp.begin(), tmp.end()) - tmp.begin());
        }
      }
    }
    h->seg_ptr[(size_t) G * K] = (long) idx.size();
    total_sink = (long) idx.size();
    h->sink_nnz = total_sink;
    // FspMatrixConstrained.cpp:447-465
    if (h->n_ti > 0) flops += 2 * ti_sink_distinct;
    for (int g = 0; g < h->n_tv; ++g) flops += 2 * tv_sink_nnz[g] + K;

    // chunk the segments into sink CTAs
    std::vector<int>  sb_seg;
    std::vector<long> sb_b, sb_e;
    for (int s = 0; s < G * K; ++s) {
      for (long b = h->seg_ptr[s]; b < h->seg_ptr[s + 1]; b += kSinkChunk) {
        sb_seg.push_back(s); sb_b.push_back(b); sb_e.push_back(std::min(b + kSinkChunk, h->seg_ptr[s + 1]));
      }
    }
    if (sb_seg.empty()) { sb_seg.push_back(-1); sb_b.push_back(0); sb_e.push_back(0); }
    h->sink_blocks = (int) sb_seg.size();
    FSP_CUDA_CHECK(pmalloc(&h->d_sink_idx, sizeof(int) * std::max(total_sink, 1L)));
    FSP_CUDA_CHECK(pmalloc(&h->d_sink_val, sizeof(double) * std::max(total_sink, 1L)));
    FSP_CUDA_CHECK(pmalloc(&h->d_sb_seg, sizeof(int) * h->sink_blocks));
    FSP_CUDA_CHECK(pmalloc(&h->d_sb_begin, sizeof(long) * h->sink_blocks));
    FSP_CUDA_CHECK(pmalloc(&h->d_sb_end, sizeof(long) * h->sink_blocks));
    FSP_CUDA_CHECK(pmalloc(&h->d_sink_partials, sizeof(double) * h->sink_blocks));
    FSP_CUDA_CHECK(pmalloc(&h->d_sink_counter, sizeof(unsigned)));
    FSP_CUDA_CHECK(cudaMemset(h->d_sink_counter, 0, sizeof(unsigned)));
    if (total_sink > 0) {
      FSP_CUDA_CHECK(cudaMemcpy(h->d_sink_idx, idx.data(), sizeof(int) * total_sink, cudaMemcpyHostToDevice));
      FSP_CUDA_CHECK(cudaMemcpy(h->d_sink_val, val.data(), sizeof(double) * total_sink, cudaMemcpyHostToDevice));
    }
    FSP_CUDA_CHECK(cudaMemcpy(h->d_sb_seg, sb_seg.data(), sizeof(int) * h->sink_blocks, cudaMemcpyHostToDevice));
    FSP_CUDA_CHECK(cudaMemcpy(h->d_sb_begin, sb_b.data(), sizeof(long) * h->sink_blocks, cudaMemcpyHostToDevice));
    FSP_CUDA_CHECK(cudaMemcpy(h->d_sb_end, sb_e.data(), sizeof(long) * h->sink_blocks, cudaMemcpyHostToDevice));
  }
  h->flops = flops;
  h->bytes = (double) n * (16.0 + 12.0 * P + 8.0 * h->ND) + 12.0 * (double) total_sink + 8.0 * K;
  h->has_values = true;
  return 0;
}

// phase: 0 = everything in one launch; 1 = interior pass (all rows, ghost entries count as 0, no sink rows);
//        2 = boundary rows only (needs the ghost buffer); 3 = sink partial sums only (written to sink_out)
static int launch_action(fspmat_t h, const double *coef_host, const double *x, const double *ghost, double *y,
                         double *sink_out, int phase, cudaStream_t st) {
  if (!h->has_values) return 0;  // FspMatrixBase.cpp:41 -- an operator without values acts as zero
  Coefs cf;
  for (int g = 0; g < h->n_tv; ++g) { cf.c[g] = coef_host[h->tv[g]]; cf.cd[g] = cf.c[g]; }
  for (int q = 0; q < h->n_ti; ++q) cf.c[h->n_tv + q] = h->ti_coef;
  if (h->n_ti > 0) cf.cd[h->n_tv] = h->ti_coef;

  MatView m;
  m.n = h->n; m.n_rows_main = h->n; m.P = h->P; m.ND = h->ND; m.ld = h->ld;
  m.col = h->d_col; m.off = h->d_off; m.diag = h->d_diag;
  m.K = h->K; m.G = h->ND;
  m.sink_blocks = (h->K > 0 && (h->owns_sinks || sink_out)) ? h->sink_blocks : 0;
  m.sb_seg = h->d_sb_seg; m.sb_begin = h->d_sb_begin; m.sb_end = h->d_sb_end;
  m.sink_idx = h->d_sink_idx; m.sink_val = h->d_sink_val;
  m.sink_partials = h->d_sink_partials; m.sink_counter = h->d_sink_counter;
  m.owns_sinks = h->owns_sinks;
  m.ghost_zero = 0; m.row_list = nullptr; m.write_y_sinks = 1;
  m.sink_flag_remote = nullptr; m.sink_epoch = 0;
  m.cta_order = nullptr; m.n_interior_ctas = 0; m.row0 = 0;

  // Kernel selection.  variant 0 (default) = lean kernel: 1 row per thread, 32 registers, 8 CTAs/SM.  Measured on
  // one B200 (465^3 lattice, same GPU, profiles/r01_variants.md): lean 6.77 TB/s, 2 rows/thread (64 regs) 6.15 TB/s,
  // 1 row/thread capped at 32 regs with spills 5.59 TB/s, 4 rows/thread 5.8 TB/s.  Occupancy wins because each row
  // needs two dependent memory round trips (column index -> gathered x).
  // codes: 10/11/12 = lean (no ghosts / ghost buffer / interior pass), 1 = rows1, 2 = rows2, 3 = rows1 capped, 4 = rows4
  int rows_per_thread;
  switch (h->variant) {
    case 1: rows_per_thread = 1; break;
    case 2: rows_per_thread = 2; break;
    case 3: rows_per_thread = 3; break;
    case 4: rows_per_thread = 4; break;
    default: rows_per_thread = h->n_ghost > 0 ? 11 : 10;
  }
  // the 128-bit vector paths need 16-byte aligned x and y
  if ((rows_per_thread == 2 || rows_per_thread == 4) && (((uintptr_t) x & 15u) || ((uintptr_t) y & 15u)))
    rows_per_thread = h->n_ghost > 0 ? 11 : 10;
  if (phase == 1) { ghost = nullptr; m.ghost_zero = 1; m.sink_blocks = 0; if (rows_per_thread >= 10) rows_per_thread = 12; }
  if (phase == 2) { m.row_list = h->d_boundary_rows; m.n = (int) h->n_boundary; m.sink_blocks = 0; rows_per_thread = 1; }
  if (phase == 3) { m.n = 0; m.write_y_sinks = 0; rows_per_thread = 1; }
  action_fn fn = pick_kernel(h->P, rows_per_thread);
  if (!fn) { fn = fsp_action_generic; rows_per_thread = 1; }
  if (phase == 2) fn = fsp_action_rowlist;
  long per_block = (long) kThreads * ((rows_per_thread == 3 || rows_per_thread >= 10) ? 1 : rows_per_thread);
  m.main_blocks = (int) ((m.n + per_block - 1) / per_block);
  int grid = m.main_blocks + m.sink_blocks;
  if (grid == 0) return 0;
  fn<<<grid, kThreads, 0, st>>>(m, cf, x, ghost, y, sink_out);
  FSP_LAUNCH_CHECK();
  return 0;
}

int fspmat_action(fspmat_t h, const double *coef_host, const double *x, const double *ghost, double *y,
                  double *sink_out, void *stream) {
  return launch_action(h, coef_host, x, ghost, y, sink_out, 0, resolve_stream(stream));
}

int fspmat_action_phase(fspmat_t h, const double *coef_host, const double *x, const double *ghost, double *y,
                        double *sink_out, int phase, void *stream) {
  if (phase < 0 || phase > 3) { set_error("fspmat_action_phase: bad phase %d", phase); return -1; }
  return launch_action(h, coef_host, x, ghost, y, sink_out, phase, resolve_stream(stream));
}

int fspmat_num_boundary_rows(fspmat_t h, long *n) { *n = h->n_boundary; return 0; }

static void fill_coefs_view(fspmat_t h, const double *coef_host, Coefs &cf, MatView &m) {
  for (int g = 0; g < h->n_tv; ++g) { cf.c[g] = coef_host[h->tv[g]]; cf.cd[g] = cf.c[g]; }
  for (int q = 0; q < h->n_ti; ++q) cf.c[h->n_tv + q] = h->ti_coef;
  if (h->n_ti > 0) cf.cd[h->n_tv] = h->ti_coef;
  m.n = h->n; m.n_rows_main = h->n; m.P = h->P; m.ND = h->ND; m.ld = h->ld;
  m.col = h->d_col; m.off = h->d_off; m.diag = h->d_diag;
  m.K = h->K; m.G = h->ND;
  m.sink_blocks = h->K > 0 ? h->sink_blocks : 0;
  m.sb_seg = h->d_sb_seg; m.sb_begin = h->d_sb_begin; m.sb_end = h->d_sb_end;
  m.sink_idx = h->d_sink_idx; m.sink_val = h->d_sink_val;
  m.sink_partials = h->d_sink_partials; m.sink_counter = h->d_sink_counter;
  m.owns_sinks = h->owns_sinks;
  m.ghost_zero = 0; m.row_list = nullptr; m.write_y_sinks = 0;
  m.sink_flag_remote = nullptr; m.sink_epoch = 0;
  m.cta_order = nullptr; m.n_interior_ctas = 0; m.row0 = 0;
  m.main_blocks = 0;
}

// Rows [row_begin, row_end) of y = A(t) x (no ghost columns); with_sinks != 0 also computes the K sink rows (needs all
// of x).  Building block of the host-vector pipeline: chunks of rows run as soon as the part of x they reference has
// been uploaded, and their part of y goes back while later chunks compute.
int fspmat_action_rows(fspmat_t h, const double *coef_host, const double *x, double *y, long row_begin, long row_end,
                       int with_sinks, void *stream) {
  if (!h->has_values) return 0;
  if (h->n_ghost != 0) { set_error("fspmat_action_rows: operator has ghost columns"); return -1; }
  if (row_begin < 0 || row_end > h->n || row_begin > row_end) { set_error("fspmat_action_rows: bad row range"); return -1; }
  action_fn fn = pick_kernel(h->P, 13);
  if (!fn) { set_error("fspmat_action_rows: supports 1..16 reactions (got %d)", h->P); return -1; }
  Coefs cf; MatView m;
  fill_coefs_view(h, coef_host, cf, m);
  m.row0 = (int) row_begin;
  m.n = (int) row_end;
  m.write_y_sinks = 1;
  m.sink_blocks = (with_sinks && h->K > 0 && h->owns_sinks) ? h->sink_blocks : 0;
  m.main_blocks = (int) ((row_end - row_begin + kThreads - 1) / kThreads);
  const int grid = m.main_blocks + m.sink_blocks;
  if (grid == 0) return 0;
  fn<<<grid, kThreads, 0, resolve_stream(stream)>>>(m, cf, x, nullptr, y, nullptr);
  FSP_LAUNCH_CHECK();
  return 0;
}

// chunk_max_host[c] = largest index of x referenced by rows [c*chunk_rows, (c+1)*chunk_rows) (their own index included)
int fspmat_chunk_max_columns(fspmat_t h, long chunk_rows, int n_chunks, int *chunk_max_host) {
  if (!h->has_values || h->n <= 0) { for (int c = 0; c < n_chunks; ++c) chunk_max_host[c] = -1; return 0; }
  if (chunk_rows <= 0 || (long) n_chunks * chunk_rows < h->n) { set_error("fspmat_chunk_max_columns: chunks do not cover the rows"); return -1; }
  int *d = nullptr;
  FSP_CUDA_CHECK(pmalloc(&d, sizeof(int) * n_chunks));
  FSP_CUDA_CHECK(cudaMemsetAsync(d, 0xff, sizeof(int) * n_chunks, (cudaStream_t) 0));  // -1
  chunk_max_col_kernel<<<(unsigned) ((h->n + 255) / 256), 256, 0, (cudaStream_t) 0>>>(h->n, h->P, h->ld, h->d_col, (int) chunk_rows, d);
  FSP_LAUNCH_CHECK();
  FSP_CUDA_CHECK(cudaMemcpy(chunk_max_host, d, sizeof(int) * n_chunks, cudaMemcpyDeviceToHost));
  pfree(d);
  return 0;
}

// chunk_flag_host[c] = 1 when rows [c*chunk_rows, (c+1)*chunk_rows) contain a row that references a ghost entry
int fspmat_chunk_has_ghost(fspmat_t h, long chunk_rows, int n_chunks, int *chunk_flag_host) {
  for (int c = 0; c < n_chunks; ++c) chunk_flag_host[c] = 0;
  if (!h->has_values || h->n_boundary <= 0) return 0;
  if (chunk_rows <= 0 || (long) n_chunks * chunk_rows < h->n) { set_error("fspmat_chunk_has_ghost: chunks do not cover the rows"); return -1; }
  int *d = nullptr;
  FSP_CUDA_CHECK(pmalloc(&d, sizeof(int) * n_chunks));
  FSP_CUDA_CHECK(cudaMemsetAsync(d, 0, sizeof(int) * n_chunks, (cudaStream_t) 0));
  mark_chunk_kernel<<<(unsigned) ((h->n_boundary + 255) / 256), 256, 0, (cudaStream_t) 0>>>(h->d_boundary_rows, h->n_boundary, chunk_rows, d);
  FSP_LAUNCH_CHECK();
  FSP_CUDA_CHECK(cudaMemcpy(chunk_flag_host, d, sizeof(int) * n_chunks, cudaMemcpyDeviceToHost));
  pfree(d);
  return 0;
}

int fspmat_fused_supported(fspmat_t h) { return (h->has_values && h->n_ghost == 0 && h->P >= 1 && h->P <= 16) ? 1 : 0; }

int fspmat_action_fused(fspmat_t h, const double *coef_host, const double *x, double *y, const fspmat_epilogue *ep,
                        void *stream) {
  if (!fspmat_fused_supported(h)) { set_error("fspmat_action_fused: not available for this operator (ghost columns, no values or > 16 reactions)"); return -1; }
  if (ep->n_dots < 0 || ep->n_dots > 2) { set_error("fspmat_action_fused: n_dots must be 0, 1 or 2"); return -1; }
  const int  n_ctas = (int) std::max<long>(1, ((long) h->n + kThreads - 1) / kThreads);
  const long pstride = ((long) n_ctas + 1 + 31) / 32 * 32;
  if (!h->d_epi_partials) FSP_CUDA_CHECK(pmalloc(&h->d_epi_partials, sizeof(double) * 2 * (size_t) pstride));
  Coefs cf; MatView m;
  fill_coefs_view(h, coef_host, cf, m);
  m.sink_blocks = (h->K > 0 && h->owns_sinks) ? h->sink_blocks : 0;
  m.main_blocks = n_ctas;
  Epi e;
  e.alpha = ep->alpha; e.beta = ep->beta; e.scale = ep->scale_dev; e.n_dots = ep->n_dots;
  e.vec[0] = ep->dot_vec_dev[0]; e.vec[1] = ep->dot_vec_dev[1];
  e.out = ep->dot_out_dev; e.partials = h->d_epi_partials; e.pstride = pstride; e.G = n_ctas;
  epi_fn fn = pick_epi(h->P);
  fn<<<m.main_blocks + m.sink_blocks, kThreads, 0, resolve_stream(stream)>>>(m, cf, e, x, y);
  FSP_LAUNCH_CHECK();
  // fixed-shape reduction of the per-CTA partials (+ the sink-row slot)
  const long slots = (long) n_ctas + (m.sink_blocks > 0 ? 1 : 0);
  for (int k = 0; k < ep->n_dots; ++k)
    if (fspvec_sum(ep->dot_out_dev + k, h->d_epi_partials + (size_t) k * pstride, slots, stream)) return -1;
  return 0;
}

int fspmat_action_sinks_p2p(fspmat_t h, const double *coef_host, const double *x, const fsphalo_epoch *e, void *stream) {
  if (!h->has_values || h->K <= 0) return 0;
  Coefs cf; MatView m;
  fill_coefs_view(h, coef_host, cf, m);
  m.n = 0;
  m.sink_flag_remote = e->sink_flag_remote;
  m.sink_epoch = e->epoch;
  action_fn fn = pick_kernel(h->P, 1);
  if (!fn) fn = fsp_action_generic;
  fn<<<m.sink_blocks, kThreads, 0, resolve_stream(stream)>>>(m, cf, x, nullptr, nullptr, e->sink_slot_remote);
  FSP_LAUNCH_CHECK();
  return 0;
}

int fspmat_action_p2p(fspmat_t h, const double *coef_host, const double *x, double *y, const fsphalo_epoch *e, void *stream) {
  if (!h->has_values) return 0;
  p2p_fn fn = pick_p2p(h->P);
  if (!fn) { set_error("fspmat_action_p2p: supports 1..16 reactions (got %d)", h->P); return -1; }
  if (!h->d_cta_order) {
    // this rank references no ghost entry (or owns no state at all): identity order; the last CTA still waits for the
    // peers' flags -- every rank must, the flag wait is what paces the reuse of the two ghost buffers
    const int n_ctas = (int) std::max<long>(1, ((long) h->n + kThreads - 1) / kThreads);
    std::vector<int> order((size_t) n_ctas);
    for (int q = 0; q < n_ctas; ++q) order[(size_t) q] = q;
    FSP_CUDA_CHECK(pmalloc(&h->d_cta_order, sizeof(int) * n_ctas));
    FSP_CUDA_CHECK(cudaMemcpy(h->d_cta_order, order.data(), sizeof(int) * n_ctas, cudaMemcpyHostToDevice));
    h->n_ctas = n_ctas;
    h->n_interior_ctas = n_ctas - 1;
  }
  Coefs cf; MatView m;
  fill_coefs_view(h, coef_host, cf, m);
  m.cta_order = h->d_cta_order;
  m.n_interior_ctas = h->n_interior_ctas;
  m.main_blocks = h->n_ctas;
  P2PWait w;
  w.halo_flags = e->halo_flags;
  const bool finish_sinks = h->K > 0 && h->owns_sinks;
  w.sink_flags = finish_sinks ? e->sink_flags : nullptr;
  w.sink_slots = e->sink_slots;
  w.epoch = e->epoch; w.n_ranks = e->n_ranks; w.err = e->error_flag; w.rank = e->self_rank;
  fn<<<m.main_blocks + (finish_sinks ? 1 : 0), kThreads, 0, resolve_stream(stream)>>>(m, cf, w, x, e->ghost, y);
  FSP_LAUNCH_CHECK();
  return 0;
}

int fspmat_halo_fused_supported(fspmat_t h) { return (h->has_values && h->P >= 1 && h->P <= 16) ? 1 : 0; }

static int launch_halo(fspmat_t h, const double *coef_host, const double *x, double *y, const fsphalo_epoch *e,
                       const fsphalo_push *push, int parts, long row_begin, long row_end, int rows_have_ghosts,
                       const double *packed_send, cudaStream_t st) {
  if (!h->has_values) return 0;
  halo_fn fn = pick_halo(h->P);
  if (!fn) { set_error("fspmat_action_halo: supports 1..16 reactions (got %d)", h->P); return -1; }
  Coefs cf; MatView m;
  fill_coefs_view(h, coef_host, cf, m);
  const bool whole = row_begin == 0 && row_end == h->n;
  m.row0 = (int) row_begin;
  m.n = (int) row_end;
  m.main_blocks = (int) ((row_end - row_begin + kThreads - 1) / kThreads);
  if (!(parts & 2)) m.sink_blocks = 0;
  HaloView hv;
  memcpy(&hv.push, push, sizeof(PushView));
  hv.n_push = (parts & 1) ? hv.push.n_ctas : 0;
  hv.push_src = packed_send;
  if (packed_send) hv.push.send_idx = nullptr;
  hv.halo_flags = e->halo_flags; hv.sink_flags = e->sink_flags; hv.sink_slots = e->sink_slots;
  hv.ghost = e->ghost; hv.sink_slot_remote = e->sink_slot_remote; hv.sink_flag_remote = e->sink_flag_remote;
  hv.err = e->error_flag;
  // bit 3 of parts: no rank publishes sink sums in this exchange (halo-only diagnostics): the finishing CTA must not wait for them
  hv.finish_sinks = (h->K > 0 && h->owns_sinks && !(parts & 8)) ? 1 : 0;
  if (whole) {
    hv.rot = (h->rot < m.main_blocks) ? h->rot : 0;
    // operators without ghost columns on this rank: every CTA is ghost-free
    hv.n_fast = h->n_ghost > 0 ? h->n_fast : m.main_blocks;
  } else {
    hv.rot = 0;
    hv.n_fast = (rows_have_ghosts && h->n_ghost > 0) ? 0 : m.main_blocks;
  }
  const int grid = hv.n_push + m.sink_blocks + m.main_blocks + ((parts & 4) ? 1 : 0);
  if (grid == 0) return 0;
  fn<<<grid, kThreads, 0, st>>>(m, cf, hv, x, y);
  FSP_LAUNCH_CHECK();
  return 0;
}

int fspmat_action_halo(fspmat_t h, const double *coef_host, const double *x, double *y, const fsphalo_epoch *e,
                       const fsphalo_push *push, void *stream) {
  return launch_halo(h, coef_host, x, y, e, push, 7, 0, h->n, 1, nullptr, resolve_stream(stream));
}

int fspmat_action_halo_part(fspmat_t h, const double *coef_host, const double *x, double *y, const fsphalo_epoch *e,
                            const fsphalo_push *push, int parts, long row_begin, long row_end, int rows_have_ghosts,
                            const double *packed_send_dev, void *stream) {
  if (row_begin < 0 || row_end > h->n || row_begin > row_end) { set_error("fspmat_action_halo_part: bad row range"); return -1; }
  return launch_halo(h, coef_host, x, y, e, push, parts, row_begin, row_end, rows_have_ghosts, packed_send_dev,
                     resolve_stream(stream));
}

int fspmat_p2p_cta_counts(fspmat_t h, long *n_interior, long *n_total) {
  *n_interior = h->n_interior_ctas;
  *n_total = h->n_ctas;
  return 0;
}

int fspmat_action_boundary_p2p(fspmat_t h, const double *coef_host, const double *x, double *y, const fsphalo_epoch *e,
                               void *stream) {
  if (!h->has_values) return 0;
  Coefs cf; MatView m;
  fill_coefs_view(h, coef_host, cf, m);
  m.row_list = h->d_boundary_rows;
  m.n = (int) h->n_boundary;
  m.main_blocks = std::max(1, (int) ((h->n_boundary + kThreads - 1) / kThreads));  // >= 1: every rank waits (pacing)
  P2PWait w;
  w.halo_flags = e->halo_flags;
  const bool finish_sinks = h->K > 0 && h->owns_sinks;
  w.sink_flags = finish_sinks ? e->sink_flags : nullptr;
  w.sink_slots = e->sink_slots;
  w.epoch = e->epoch; w.n_ranks = e->n_ranks; w.err = e->error_flag;
  w.rank = e->self_rank;
  const int grid = m.main_blocks + (finish_sinks ? 1 : 0);
  fsp_action_boundary_p2p_kernel<<<grid, kThreads, 0, resolve_stream(stream)>>>(m, cf, w, x, e->ghost, y);
  FSP_LAUNCH_CHECK();
  return 0;
}

int fspmat_build_ghosts(int *col, long n, int lo, int hi, int **ghost_out, long *n_ghost) {
  *ghost_out = nullptr;
  *n_ghost = 0;
  if (n <= 0) return 0;
  int *d_sel = nullptr, *d_sorted = nullptr, *d_uniq = nullptr, *d_num = nullptr;
  void *d_tmp = nullptr;
  size_t need = 0, cap = 0;
  FSP_CUDA_CHECK(pmalloc(&d_num, sizeof(int)));
  FSP_CUDA_CHECK(pmalloc(&d_sel, sizeof(int) * n));
  OutOfRange pred{lo, hi};
  cub::DeviceSelect::If(nullptr, need, col, d_sel, d_num, (int) n, pred);
  FSP_CUDA_CHECK(pmalloc(&d_tmp, need)); cap = need;
  FSP_CUDA_CHECK(cub::DeviceSelect::If(d_tmp, need, col, d_sel, d_num, (int) n, pred));
  count_launch();
  int n_sel = 0;
  FSP_CUDA_CHECK(cudaMemcpy(&n_sel, d_num, sizeof(int), cudaMemcpyDeviceToHost));
  int n_u = 0;
  if (n_sel > 0) {
    FSP_CUDA_CHECK(pmalloc(&d_sorted, sizeof(int) * n_sel));
    FSP_CUDA_CHECK(pmalloc(&d_uniq, sizeof(int) * n_sel));
    cub::DeviceRadixSort::SortKeys(nullptr, need, d_sel, d_sorted, n_sel);
    if (need > cap) { pfree(d_tmp); FSP_CUDA_CHECK(pmalloc(&d_tmp, need)); cap = need; }
    FSP_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(d_tmp, need, d_sel, d_sorted, n_sel));
    count_launch();
    cub::DeviceSelect::Unique(nullptr, need, d_sorted, d_uniq, d_num, n_sel);
    if (need > cap) { pfree(d_tmp); FSP_CUDA_CHECK(pmalloc(&d_tmp, need)); cap = need; }
    FSP_CUDA_CHECK(cub::DeviceSelect::Unique(d_tmp, need, d_sorted, d_uniq, d_num, n_sel));
    count_launch();
    FSP_CUDA_CHECK(cudaMemcpy(&n_u, d_num, sizeof(int), cudaMemcpyDeviceToHost));
  }
  remap_cols_kernel<<<(unsigned) ((n + 255) / 256), 256>>>(col, n, lo, hi, d_uniq, n_u);
  FSP_LAUNCH_CHECK();
  FSP_CUDA_CHECK(cudaDeviceSynchronize());
  pfree(d_sel); pfree(d_sorted); pfree(d_tmp); pfree(d_num);
  *ghost_out = d_uniq;
  *n_ghost = n_u;
  return 0;
}

int fspmat_shift_indices(int *idx, long n, int delta) {
  if (n <= 0) return 0;
  shift_idx_kernel<<<(unsigned) ((n + 255) / 256), 256>>>(idx, n, delta);
  FSP_LAUNCH_CHECK();
  return 0;
}

// Assembled A(t) as CSR on the device (src/Matrix/FspMatrixBase.cpp:308-427, FspMatrixConstrained.cpp:304-445).
// Layout: every state row has P + 1 slots (diagonal first, then one per reaction plane); the K sink rows list the
// boundary entries of their constraint group by group.  structure != 0 also writes row_ptr / col (CreateRHSJacobian);
// structure == 0 refreshes the values for new coefficients only (ComputeRHSJacobian).
int fspmat_csr_size(fspmat_t h, long *nnz, int *n_rows) {
  long z = (long) h->n * (h->P + 1);
  if (h->K > 0 && h->owns_sinks) z += h->sink_nnz;
  *nnz = h->has_values ? z : 0;
  *n_rows = h->n_rows;
  return 0;
}
int fspmat_csr_export(fspmat_t h, const double *coef_host, int structure, int *row_ptr, int *col, double *val, void *stream) {
  cudaStream_t st = resolve_stream(stream);
  if (h->n_ghost > 0) { set_error("fspmat_csr_export: operators with ghost columns have no assembled form"); return -1; }
  if (!h->has_values) {
    if (structure) FSP_CUDA_CHECK(cudaMemsetAsync(row_ptr, 0, sizeof(int) * ((size_t) h->n_rows + 1), st));
    return 0;
  }
  Coefs cf; MatView m;
  fill_coefs_view(h, coef_host, cf, m);
  if (h->n > 0) {
    csr_state_rows_kernel<<<(unsigned) ((h->n + 255) / 256), 256, 0, st>>>(m, cf, structure, row_ptr, col, val);
    FSP_LAUNCH_CHECK();
  }
  const long base = (long) h->n * (h->P + 1);
  std::vector<int> tail;  // row_ptr[n .. n_rows]
  long at = base;
  if (h->K > 0 && h->owns_sinks) {
    for (int k = 0; k < h->K; ++k) {
      tail.push_back((int) at);
      for (int g = 0; g < h->ND; ++g) {
        const long b = h->seg_ptr[(size_t) g * h->K + k], e = h->seg_ptr[(size_t) g * h->K + k + 1];
        if (e > b) {
          csr_sink_segment_kernel<<<(unsigned) ((e - b + 255) / 256), 256, 0, st>>>(h->d_sink_idx + b, h->d_sink_val + b, e - b, cf.cd[g],
                                                                                   structure, col + at, val + at);
          FSP_LAUNCH_CHECK();
          at += e - b;
        }
      }
    }
  } else {
    for (int r = h->n; r < h->n_rows; ++r) tail.push_back((int) at);
  }
  tail.push_back((int) at);
  if (structure)
    FSP_CUDA_CHECK(cudaMemcpyAsync(row_ptr + h->n, tail.data(), sizeof(int) * tail.size(), cudaMemcpyHostToDevice, st));
  FSP_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}
int fspmat_csr_spmv(int n_rows, const int *row_ptr, const int *col, const double *val, const double *x, double *y, void *stream) {
  if (n_rows <= 0) return 0;
  csr_spmv_kernel<<<(unsigned) ((n_rows + 255) / 256), 256, 0, resolve_stream(stream)>>>(n_rows, row_ptr, col, val, x, y);
  FSP_LAUNCH_CHECK();
  return 0;
}

int fspmat_flops(fspmat_t h, long *nflops) { *nflops = h->flops; return 0; }
int fspmat_num_rows(fspmat_t h, int *n_rows) { *n_rows = h->n_rows; return 0; }
int fspmat_action_bytes(fspmat_t h, double *bytes) { *bytes = h->bytes; return 0; }

int fspmat_dense(fspmat_t h, const double *coef_host, double *out) {
  const long nr = h->n_rows, n = h->n, ld = h->ld;
  for (long q = 0; q < nr * nr; ++q) out[q] = 0.0;
  if (!h->has_values) return 0;
  std::vector<int>    col((size_t) h->P * ld);
  std::vector<double> off((size_t) h->P * ld), diag((size_t) std::max(h->ND, 1) * ld);
  if (h->P > 0) {
    FSP_CUDA_CHECK(cudaMemcpy(col.data(), h->d_col, sizeof(int) * h->P * ld, cudaMemcpyDeviceToHost));
    FSP_CUDA_CHECK(cudaMemcpy(off.data(), h->d_off, sizeof(double) * h->P * ld, cudaMemcpyDeviceToHost));
    FSP_CUDA_CHECK(cudaMemcpy(diag.data(), h->d_diag, sizeof(double) * h->ND * ld, cudaMemcpyDeviceToHost));
  }
  for (int p = 0; p < h->P; ++p) {
    double c = p < h->n_tv ? coef_host[h->tv[p]] : 1.0;
    for (long i = 0; i < n; ++i) {
      int j = col[(size_t) p * ld + i];
      if (j >= 0) out[(size_t) j * nr + i] += c * off[(size_t) p * ld + i];
    }
  }
  for (int g = 0; g < h->ND; ++g) {
    double c = g < h->n_tv ? coef_host[h->tv[g]] : 1.0;
    for (long i = 0; i < n; ++i) out[(size_t) i * nr + i] -= c * diag[(size_t) g * ld + i];
  }
  if (h->K > 0 && h->owns_sinks && h->sink_nnz > 0) {
    std::vector<int>    idx(h->sink_nnz);
    std::vector<double> val(h->sink_nnz);
    FSP_CUDA_CHECK(cudaMemcpy(idx.data(), h->d_sink_idx, sizeof(int) * h->sink_nnz, cudaMemcpyDeviceToHost));
    FSP_CUDA_CHECK(cudaMemcpy(val.data(), h->d_sink_val, sizeof(double) * h->sink_nnz, cudaMemcpyDeviceToHost));
    for (int g = 0; g < h->ND; ++g) {
      double c = g < h->n_tv ? coef_host[h->tv[g]] : 1.0;
      for (int k = 0; k < h->K; ++k)
        for (long q = h->seg_ptr[(size_t) g * h->K + k]; q < h->seg_ptr[(size_t) g * h->K + k + 1]; ++q)
          out[(size_t) idx[q] * nr + (n + k)] += c * val[q];
    }
  }
  return 0;
}

}  // extern "C"

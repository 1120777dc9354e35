// fspset.cu -- device-resident state set: state list + hash directory + BFS expansion
// (include/fsp_b200.h "State set on the device").
//
// Replaces, for the FSP hot path, the Zoltan distributed directory and Armadillo bookkeeping of the
// reference:
//   StateSetBase::AddStates / State2Index      src/StateSet/StateSetBase.cpp:188-258, 309-423
//   StateSetConstrained::Expand                src/StateSet/StateSetConstrained.cpp:132-221
//   CheckValidityStates / CheckConstraints     src/StateSet/StateSetConstrained.cpp:33-82
//   unique_columns                             src/Sys/pacmenMath.h:204-213
//
// Data layout in HBM: states int32 [n][S] (state i at states[i*S..], = the reference's column-major
// arma::Mat<int>), status int8 [n], and an open-addressing hash table of uint32 slots holding STATE
// INDICES (keys live only in the state list; 4 bytes per slot, load factor <= 0.5).
//
// Determinism / bit-exact index map: candidates of one batch carry provisional ids n_old + c (c = their
// position in the reaction-major child list).  Insertion claims a slot with atomicCAS and resolves
// duplicate keys with atomicMin on the id, so the FIRST occurrence wins regardless of thread timing;
// winners are then compacted in order (prefix sum) -> the appended order is first-discovery order,
// identical to the CPU oracle (oracle/fsp_oracle.c: orc_set_expand).
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "fsp_common.cuh"

using namespace fspb;

namespace {

constexpr int      kMaxS = 16;
constexpr int      kMaxK = 16;
constexpr unsigned kEmpty = 0xFFFFFFFFu;
constexpr long     kBatch = 1L << 25;  // candidates per insertion batch

struct SmallVec {
  int v[kMaxS];
};
struct Bounds {
  int b[kMaxK];
};

__device__ __forceinline__ unsigned long long hash_state(const int *x, int S) {
  unsigned long long h = 0x9E3779B97F4A7C15ull;
  for (int s = 0; s < S; ++s) {
    h ^= (unsigned long long) (unsigned) x[s] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 32;
  }
  return h;
}

__device__ __forceinline__ const int *key_of(unsigned id, const int *states, long n_old, const int *cand, int S) {
  return id < (unsigned) n_old ? states + (size_t) id * S : cand + (size_t) (id - (unsigned) n_old) * S;
}

__device__ __forceinline__ bool key_equal(const int *a, const int *b, int S) {
  for (int s = 0; s < S; ++s)
    if (a[s] != b[s]) return false;
  return true;
}

// Insert candidate c (provisional id n_old + c) -- first occurrence wins via atomicMin.
__global__ void insert_kernel(unsigned *table, unsigned long long mask, const int *states, long n_old,
                              const int *cand, long m, const signed char *valid, int S) {
  long c = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  if (valid && !valid[c]) return;
  const int         *key = cand + (size_t) c * S;
  const unsigned     id = (unsigned) (n_old + c);
  unsigned long long h = hash_state(key, S) & mask;
  while (true) {
    unsigned cur = __ldcg(&table[h]);  // L2 read: other CTAs are claiming slots concurrently
    if (cur == kEmpty) {
      unsigned prev = atomicCAS(&table[h], kEmpty, id);
      if (prev == kEmpty) return;
      cur = prev;
    }
    if (key_equal(key_of(cur, states, n_old, cand, S), key, S)) {
      atomicMin(&table[h], id);
      return;
    }
    h = (h + 1) & mask;
  }
}

// flag[c] = 1 iff candidate c is the winner for its key (i.e. new and first occurrence); slot[c] = table pos
__global__ void winner_kernel(const unsigned *table, unsigned long long mask, const int *states, long n_old,
                              const int *cand, long m, const signed char *valid, int S, int *flag,
                              unsigned long long *slot) {
  long c = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  int f = 0;
  if (!valid || valid[c]) {
    const int         *key = cand + (size_t) c * S;
    unsigned long long h = hash_state(key, S) & mask;
    while (true) {
      unsigned cur = table[h];
      if (cur == kEmpty) break;  // cannot happen for inserted keys
      if (key_equal(key_of(cur, states, n_old, cand, S), key, S)) {
        if (cur == (unsigned) (n_old + c)) { f = 1; slot[c] = h; }
        break;
      }
      h = (h + 1) & mask;
    }
  }
  flag[c] = f;
}

__global__ void append_kernel(unsigned *table, int *states, signed char *status, long n_old, const int *cand, long m,
                              int S, const int *flag, const int *pos, const unsigned long long *slot) {
  long c = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m || !flag[c]) return;
  long dst = n_old + pos[c];
  for (int s = 0; s < S; ++s) states[(size_t) dst * S + s] = cand[(size_t) c * S + s];
  status[dst] = 1;
  table[slot[c]] = (unsigned) dst;
}

__global__ void rehash_kernel(unsigned *table, unsigned long long mask, const int *states, long n, int S) {
  long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long h = hash_state(states + (size_t) i * S, S) & mask;
  while (true) {
    unsigned prev = atomicCAS(&table[h], kEmpty, (unsigned) i);
    if (prev == kEmpty) return;
    h = (h + 1) & mask;
  }
}

// State2Index (StateSetBase.cpp:309-343): -1 for negative coordinates or absent states
__device__ __forceinline__ int lookup_state(const unsigned *table, unsigned long long mask, const int *states,
                                            const int *key, int S) {
  for (int s = 0; s < S; ++s)
    if (key[s] < 0) return -1;
  unsigned long long h = hash_state(key, S) & mask;
  while (true) {
    unsigned cur = table[h];
    if (cur == kEmpty) return -1;
    if (key_equal(states + (size_t) cur * S, key, S)) return (int) cur;
    h = (h + 1) & mask;
  }
}

__global__ void lookup_kernel(const unsigned *table, unsigned long long mask, const int *states, const int *X,
                              long m, int S, int *idx) {
  long j = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  idx[j] = lookup_state(table, mask, states, X + (size_t) j * S, S);
}

__global__ void lookup_shifted_kernel(const unsigned *table, unsigned long long mask, const int *states, long first,
                                      long count, int S, SmallVec nu, int sign, int *idx) {
  long j = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  int key[kMaxS];
  for (int s = 0; s < S; ++s) key[s] = states[(size_t) (first + j) * S + s] + sign * nu.v[s];
  idx[j] = lookup_state(table, mask, states, key, S);
}

// CheckConstraints with the default identity lhs (StateSetConstrained.cpp:63-82,92-99)
__global__ void check_shifted_default_kernel(const int *states, long first, long count, int S, int K, SmallVec nu,
                                             Bounds bd, int *satisfied) {
  long j = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  int  key[kMaxS];
  bool neg = false;
  for (int s = 0; s < S; ++s) {
    key[s] = states[(size_t) (first + j) * S + s] + nu.v[s];
    neg |= key[s] < 0;
  }
  for (int k = 0; k < K; ++k) satisfied[(size_t) k * count + j] = (neg || key[k] <= bd.b[k]) ? 1 : 0;
}

__global__ void shift_states_kernel(const int *states, long first, long count, int S, SmallVec nu, int sign,
                                    int *out) {
  long j = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  for (int s = 0; s < S; ++s) out[(size_t) j * S + s] = states[(size_t) (first + j) * S + s] + sign * nu.v[s];
}

// children of a frontier chunk under one reaction + validity with default constraints
// (CheckValidityStates, StateSetConstrained.cpp:33-56)
__global__ void children_kernel(const int *states, const int *frontier, long i0, long m, int S, int K, SmallVec nu,
                                Bounds bd, int use_default, int *cand, signed char *valid, signed char *fstatus) {
  long c = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  const int src = frontier[i0 + c];
  bool      ok = true;
  for (int s = 0; s < S; ++s) {
    int v = states[(size_t) src * S + s] + nu.v[s];
    cand[(size_t) c * S + s] = v;
    if (v < 0) ok = false;
    if (use_default && s < K && v > bd.b[s]) ok = false;
  }
  if (use_default) {
    valid[c] = ok ? 1 : 0;
    if (!ok) fstatus[i0 + c] = -1;
  }
}

__global__ void apply_valid_kernel(const signed char *valid, long i0, long m, signed char *fstatus) {
  long c = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  if (!valid[c]) fstatus[i0 + c] = -1;
}

__global__ void status_flag_kernel(const signed char *status, long n, signed char want, int *flag) {
  long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = status[i] == want ? 1 : 0;
}
__global__ void reactivate_kernel(signed char *status, long n) {
  long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && status[i] == -1) status[i] = 1;
}
__global__ void frontier_fill_kernel(const int *flag, const int *pos, long n, int *frontier) {
  long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) frontier[pos[i]] = (int) i;
}
__global__ void set_frontier_status_kernel(signed char *status, const int *frontier, const signed char *fstatus,
                                           long nF) {
  long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nF) status[frontier[i]] = fstatus[i];
}

__global__ void lattice_kernel(long first, long m, int S, SmallVec dims, int *out) {
  long c = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  long k = first + c;
  for (int s = 0; s < S; ++s) {  // ind2sub, species 0 fastest (Sys/pacmenMath.h:109-118)
    out[(size_t) c * S + s] = (int) (k % dims.v[s]);
    k /= dims.v[s];
  }
}

// separable propensity: rate * prod_s ff(x_s, order_s) * table_s[min(x_s, len_s - 1)]   (len_s == 0: no table factor)
__global__ void mass_action_kernel(const int *states, long first, long count, int S, SmallVec nu, int sign,
                                   SmallVec order, double rate, const double *__restrict__ tabs, SmallVec tab_off,
                                   SmallVec tab_len, double *out) {
  long j = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  double v = rate;
  for (int s = 0; s < S; ++s) {
    int x = states[(size_t) (first + j) * S + s] + sign * nu.v[s];
    int o = order.v[s];
    if (o == 1) v *= (double) x;
    else if (o == 2) v *= 0.5 * (double) x * (double) (x - 1);
    else if (o == 3) v *= (double) x * (double) (x - 1) * (double) (x - 2) / 6.0;
    const int len = tab_len.v[s];
    if (len > 0) v *= x < 0 ? 0.0 : __ldg(tabs + tab_off.v[s] + min(x, len - 1));
  }
  out[j] = v;
}


struct StatusNonZero {
  const signed char *status; int use;
  __host__ __device__ bool operator()(const int &i) const { return !use || status[i] != 0; }
};
__global__ void status_nonzero_flag_kernel(const signed char *status, long n, int *flag) {
  long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = status[i] != 0 ? 1 : 0;
}
__global__ void check_list_default_kernel(const int *states, long first, const int *list, long m, int S, int K,
                                          SmallVec nu, Bounds bd, int *satisfied) {
  long j = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const long i = first + list[j];
  int  key[kMaxS];
  bool neg = false;
  for (int s = 0; s < S; ++s) {
    key[s] = states[(size_t) i * S + s] + nu.v[s];
    neg |= key[s] < 0;
  }
  for (int k = 0; k < K; ++k) satisfied[(size_t) k * m + j] = (neg || key[k] <= bd.b[k]) ? 1 : 0;
}
__global__ void shift_list_kernel(const int *states, long first, const int *list, long m, int S, SmallVec nu, int *out) {
  long j = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const long i = first + list[j];
  for (int s = 0; s < S; ++s) out[(size_t) j * S + s] = states[(size_t) i * S + s] + nu.v[s];
}
struct NotFlag {
  const int *sat;
  __host__ __device__ int operator()(int i) const { return sat[i] == 0 ? 1 : 0; }
};

inline unsigned blocks_for(long m) { return (unsigned) ((m + 255) / 256); }

}  // namespace

struct fspset_s {
  int              S = 0, R = 0, K = 0;
  std::vector<int> SM;      // S x R column major
  std::vector<int> bounds;  // K
  fspset_constr_fn lhs = nullptr;
  void            *lhs_args = nullptr;
  long             n = 0, cap = 0;
  int             *d_states = nullptr;
  signed char     *d_status = nullptr;
  unsigned        *d_table = nullptr;
  unsigned long long tsize = 0;
  // scratch (grown on demand)
  int         *d_cand = nullptr;       long cand_cap = 0;
  signed char *d_valid = nullptr;
  int         *d_flag = nullptr, *d_pos = nullptr; long flag_cap = 0;
  unsigned long long *d_slot = nullptr;
  void        *d_cub = nullptr;        size_t cub_bytes = 0;
  bool             expanded = false;   // Expand() has run: status 0 <=> all children inside the set
  std::vector<int> expanded_bounds;    // bounds at the last Expand()
};

namespace {

int ensure_table(fspset_s *h, long need) {
  unsigned long long want = h->tsize ? h->tsize : 1024;
  while ((unsigned long long) need * 2 > want) want *= 2;
  if (want == h->tsize) return 0;
  pfree(h->d_table);
  FSP_CUDA_CHECK(pmalloc(&h->d_table, sizeof(unsigned) * want));
  FSP_CUDA_CHECK(cudaMemset(h->d_table, 0xFF, sizeof(unsigned) * want));
  h->tsize = want;
  if (h->n > 0) {
    rehash_kernel<<<blocks_for(h->n), 256>>>(h->d_table, want - 1, h->d_states, h->n, h->S);
    FSP_LAUNCH_CHECK();
  }
  return 0;
}

int ensure_states(fspset_s *h, long need) {
  if (need <= h->cap) return 0;
  long cap = h->cap ? h->cap : 1024;
  while (cap < need) cap = cap + cap / 2 + 1024;
  int         *ns;
  signed char *nst;
  FSP_CUDA_CHECK(pmalloc(&ns, sizeof(int) * cap * h->S));
  FSP_CUDA_CHECK(pmalloc(&nst, cap));
  if (h->n > 0) {
    FSP_CUDA_CHECK(cudaMemcpy(ns, h->d_states, sizeof(int) * h->n * h->S, cudaMemcpyDeviceToDevice));
    FSP_CUDA_CHECK(cudaMemcpy(nst, h->d_status, h->n, cudaMemcpyDeviceToDevice));
  }
  pfree(h->d_states); pfree(h->d_status);
  h->d_states = ns; h->d_status = nst; h->cap = cap;
  return 0;
}

int ensure_cand(fspset_s *h, long m) {
  if (m > h->cand_cap) {
    pfree(h->d_cand); pfree(h->d_valid); pfree(h->d_slot);
    long cap = std::max(m, 4096L);
    FSP_CUDA_CHECK(pmalloc(&h->d_cand, sizeof(int) * cap * h->S));
    FSP_CUDA_CHECK(pmalloc(&h->d_valid, cap));
    FSP_CUDA_CHECK(pmalloc(&h->d_slot, sizeof(unsigned long long) * cap));
    h->cand_cap = cap;
  }
  return 0;
}

int ensure_flags(fspset_s *h, long m) {
  if (m > h->flag_cap) {
    pfree(h->d_flag); pfree(h->d_pos);
    long cap = std::max(m, 4096L);
    FSP_CUDA_CHECK(pmalloc(&h->d_flag, sizeof(int) * cap));
    FSP_CUDA_CHECK(pmalloc(&h->d_pos, sizeof(int) * cap));
    h->flag_cap = cap;
  }
  return 0;
}

// exclusive prefix sum of d_flag[0..m) into d_pos; returns the total
int scan_flags(fspset_s *h, long m, long *total) {
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, h->d_flag, h->d_pos, (int) m);
  if (need > h->cub_bytes) {
    pfree(h->d_cub);
    FSP_CUDA_CHECK(pmalloc(&h->d_cub, need));
    h->cub_bytes = need;
  }
  FSP_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(h->d_cub, need, h->d_flag, h->d_pos, (int) m));
  count_launch();
  int last_pos = 0, last_flag = 0;
  FSP_CUDA_CHECK(cudaMemcpy(&last_pos, h->d_pos + (m - 1), sizeof(int), cudaMemcpyDeviceToHost));
  FSP_CUDA_CHECK(cudaMemcpy(&last_flag, h->d_flag + (m - 1), sizeof(int), cudaMemcpyDeviceToHost));
  *total = (long) last_pos + last_flag;
  return 0;
}

// Insert the m candidates in h->d_cand (validity in h->d_valid unless all_valid): sheds present states and
// in-batch duplicates, appends the rest in order with status 1.
int insert_batch(fspset_s *h, long m, bool all_valid) {
  if (m <= 0) return 0;
  if (h->n + m >= 0x7FFFFFF0L) { set_error("fspset: more than 2^31 states"); return -1; }
  if (ensure_table(h, h->n + m)) return -1;
  if (ensure_flags(h, m)) return -1;
  const signed char *valid = all_valid ? nullptr : h->d_valid;
  const long         n_old = h->n;
  insert_kernel<<<blocks_for(m), 256>>>(h->d_table, h->tsize - 1, h->d_states, n_old, h->d_cand, m, valid, h->S);
  FSP_LAUNCH_CHECK();
  winner_kernel<<<blocks_for(m), 256>>>(h->d_table, h->tsize - 1, h->d_states, n_old, h->d_cand, m, valid, h->S,
                                        h->d_flag, h->d_slot);
  FSP_LAUNCH_CHECK();
  long total = 0;
  if (scan_flags(h, m, &total)) return -1;
  if (total > 0) {
    if (ensure_states(h, n_old + total)) return -1;
    append_kernel<<<blocks_for(m), 256>>>(h->d_table, h->d_states, h->d_status, n_old, h->d_cand, m, h->S, h->d_flag,
                                          h->d_pos, h->d_slot);
    FSP_LAUNCH_CHECK();
    h->n = n_old + total;
  }
  return 0;
}

SmallVec nu_of(const fspset_s *h, int r, int sign = 1) {
  SmallVec v;
  for (int s = 0; s < kMaxS; ++s) v.v[s] = s < h->S ? sign * h->SM[(size_t) r * h->S + s] : 0;
  return v;
}
Bounds bounds_of(const fspset_s *h) {
  Bounds b;
  for (int k = 0; k < kMaxK; ++k) b.b[k] = k < h->K ? h->bounds[k] : 0x7FFFFFFF;
  return b;
}

// host-side validity through the user's lhs callback (CheckValidityStates, StateSetConstrained.cpp:33-56)
int host_validity(fspset_s *h, long m, std::vector<int> &cand_host, std::vector<int> &fval,
                  std::vector<signed char> &valid_host) {
  cand_host.resize((size_t) m * h->S);
  fval.resize((size_t) m * h->K);
  valid_host.resize(m);
  FSP_CUDA_CHECK(cudaMemcpy(cand_host.data(), h->d_cand, sizeof(int) * m * h->S, cudaMemcpyDeviceToHost));
  int ierr = h->lhs(h->S, h->K, (int) m, cand_host.data(), fval.data(), h->lhs_args);
  if (ierr) { set_error("fspset: constraint callback returned %d", ierr); return ierr; }
  for (long c = 0; c < m; ++c) {
    bool ok = true;
    for (int s = 0; s < h->S; ++s) ok &= cand_host[(size_t) c * h->S + s] >= 0;
    for (int k = 0; k < h->K; ++k) ok &= fval[(size_t) c * h->K + k] <= h->bounds[k];
    valid_host[c] = ok ? 1 : 0;
  }
  FSP_CUDA_CHECK(cudaMemcpy(h->d_valid, valid_host.data(), m, cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace

extern "C" {

int fspset_create(fspset_t *out, int S, int R, const int *SM) {
  if (S <= 0 || S > kMaxS) { set_error("fspset_create: num_species %d out of range (1..%d)", S, kMaxS); return -1; }
  fspset_s *h = new fspset_s();
  h->S = S; h->R = R;
  h->SM.assign(SM, SM + (size_t) S * R);
  *out = h;
  return 0;
}

int fspset_destroy(fspset_t h) {
  if (!h) return 0;
  pfree(h->d_states); pfree(h->d_status); pfree(h->d_table); pfree(h->d_cand); pfree(h->d_valid);
  pfree(h->d_flag); pfree(h->d_pos); pfree(h->d_slot); pfree(h->d_cub);
  delete h;
  return 0;
}

int fspset_set_shape(fspset_t h, int K, fspset_constr_fn lhs, const int *bounds, void *args) {
  if (K < 0 || K > kMaxK) { set_error("fspset_set_shape: %d constraints out of range (max %d)", K, kMaxK); return -1; }
  if (!lhs && K != h->S) {  // StateSetConstrained.cpp:227-233
    set_error("fspset_set_shape: default constraints need num_constr == num_species");
    return -1;
  }
  h->K = K; h->lhs = lhs; h->lhs_args = args;
  h->bounds.assign(bounds, bounds + K);
  return 0;
}

int fspset_set_bounds(fspset_t h, int K, const int *bounds) {
  if (K < 0 || K > kMaxK) { set_error("fspset_set_bounds: %d constraints out of range", K); return -1; }
  h->K = K;
  h->bounds.assign(bounds, bounds + K);
  return 0;
}

int fspset_add_states(fspset_t h, int num_species, long m, const int *X, int on_device) {
  if (num_species != h->S) return -1;  // StateSetBase.cpp:190-192
  for (long b = 0; b < m; b += kBatch) {
    long mb = std::min(kBatch, m - b);
    if (ensure_cand(h, mb)) return -1;
    FSP_CUDA_CHECK(cudaMemcpy(h->d_cand, X + (size_t) b * h->S, sizeof(int) * mb * h->S,
                              on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    if (insert_batch(h, mb, true)) return -1;
  }
  return 0;
}

int fspset_add_box_lattice(fspset_t h, const int *upper) {
  SmallVec dims;
  long     total = 1;
  for (int s = 0; s < kMaxS; ++s) {
    dims.v[s] = s < h->S ? upper[s] + 1 : 1;
    total *= dims.v[s];
    if (total >= 0x7FFFFFF0L) { set_error("fspset_add_box_lattice: lattice exceeds 2^31 states"); return -1; }
  }
  if (ensure_states(h, h->n + total)) return -1;
  for (long b = 0; b < total; b += kBatch) {
    long mb = std::min(kBatch, total - b);
    if (ensure_cand(h, mb)) return -1;
    lattice_kernel<<<blocks_for(mb), 256>>>(b, mb, h->S, dims, h->d_cand);
    FSP_LAUNCH_CHECK();
    if (insert_batch(h, mb, true)) return -1;
  }
  return 0;
}

// the BFS loop; the two scratch buffers belong to the caller, which frees them on every path (error returns included)
static int expand_impl(fspset_t h, int *&d_frontier, signed char *&d_fstatus) {
  const int use_default = h->lhs ? 0 : 1;
  reactivate_kernel<<<blocks_for(h->n), 256>>>(h->d_status, h->n);  // :137-149
  FSP_LAUNCH_CHECK();
  long         fcap = 0;
  std::vector<int>         cand_host, fval;
  std::vector<signed char> valid_host;
  int rc = 0;
  while (true) {
    // frontier = indices with status 1, ascending (:151-152, :199-200)
    const long n = h->n;
    if (ensure_flags(h, n)) { rc = -1; break; }
    status_flag_kernel<<<blocks_for(n), 256>>>(h->d_status, n, 1, h->d_flag);
    FSP_LAUNCH_CHECK();
    long nF = 0;
    if (scan_flags(h, n, &nF)) { rc = -1; break; }
    if (nF == 0) break;
    if (nF > fcap) {
      pfree(d_frontier); pfree(d_fstatus);
      d_frontier = nullptr; d_fstatus = nullptr;
      fcap = nF + nF / 2;
      FSP_CUDA_CHECK(pmalloc(&d_frontier, sizeof(int) * fcap));
      FSP_CUDA_CHECK(pmalloc(&d_fstatus, fcap));
    }
    frontier_fill_kernel<<<blocks_for(n), 256>>>(h->d_flag, h->d_pos, n, d_frontier);
    FSP_LAUNCH_CHECK();
    FSP_CUDA_CHECK(cudaMemset(d_fstatus, 0, nF));
    // children in reaction-major order (:175-179), processed in order so that first-discovery order holds
    const Bounds bd = bounds_of(h);
    for (int j = 0; j < h->R && !rc; ++j) {
      const SmallVec nu = nu_of(h, j);
      for (long i0 = 0; i0 < nF && !rc; i0 += kBatch) {
        long m = std::min(kBatch, nF - i0);
        if (ensure_cand(h, m)) { rc = -1; break; }
        children_kernel<<<blocks_for(m), 256>>>(h->d_states, d_frontier, i0, m, h->S, h->K, nu, bd, use_default,
                                                h->d_cand, h->d_valid, d_fstatus);
        FSP_LAUNCH_CHECK();
        if (!use_default) {
          if ((rc = host_validity(h, m, cand_host, fval, valid_host))) break;
          apply_valid_kernel<<<blocks_for(m), 256>>>(h->d_valid, i0, m, d_fstatus);
          FSP_LAUNCH_CHECK();
        }
        if (insert_batch(h, m, false)) rc = -1;
      }
    }
    if (rc) break;
    set_frontier_status_kernel<<<blocks_for(nF), 256>>>(h->d_status, d_frontier, d_fstatus, nF);  // :198
    FSP_LAUNCH_CHECK();
  }
  return rc;
}

int fspset_expand(fspset_t h) {
  if (h->n == 0) return 0;
  if ((int) h->bounds.size() != h->K || h->K == 0) { set_error("fspset_expand: shape not set"); return -1; }
  int         *d_frontier = nullptr;
  signed char *d_fstatus = nullptr;
  const int    rc = expand_impl(h, d_frontier, d_fstatus);
  pfree(d_frontier); pfree(d_fstatus);
  if (rc == 0) { h->expanded = true; h->expanded_bounds = h->bounds; }
  return rc;
}

int fspset_num_states(fspset_t h, int *n) { *n = (int) h->n; return 0; }

int fspset_state2index(fspset_t h, long m, const int *X, int x_on_device, int *idx, int idx_on_device) {
  if (m <= 0) return 0;
  const int *dX = X;
  int       *tX = nullptr, *dI = idx, *tI = nullptr;
  if (!x_on_device) {
    FSP_CUDA_CHECK(pmalloc(&tX, sizeof(int) * m * h->S));
    FSP_CUDA_CHECK(cudaMemcpy(tX, X, sizeof(int) * m * h->S, cudaMemcpyHostToDevice));
    dX = tX;
  }
  if (!idx_on_device) {
    FSP_CUDA_CHECK(pmalloc(&tI, sizeof(int) * m));
    dI = tI;
  }
  if (h->n == 0 || !h->d_table) {
    FSP_CUDA_CHECK(cudaMemset(dI, 0xFF, sizeof(int) * m));
  } else {
    lookup_kernel<<<blocks_for(m), 256>>>(h->d_table, h->tsize - 1, h->d_states, dX, m, h->S, dI);
    FSP_LAUNCH_CHECK();
  }
  if (!idx_on_device) FSP_CUDA_CHECK(cudaMemcpy(idx, dI, sizeof(int) * m, cudaMemcpyDeviceToHost));
  pfree(tX); pfree(tI);
  return 0;
}

int fspset_lookup_shifted(fspset_t h, const int *nu_host, int sign, long first, long count, int *idx_dev) {
  if (count <= 0) return 0;
  SmallVec nu;
  for (int s = 0; s < kMaxS; ++s) nu.v[s] = s < h->S ? nu_host[s] : 0;
  lookup_shifted_kernel<<<blocks_for(count), 256>>>(h->d_table, h->tsize - 1, h->d_states, first, count, h->S, nu,
                                                    sign, idx_dev);
  FSP_LAUNCH_CHECK();
  return 0;
}

int fspset_check_constraints_shifted(fspset_t h, const int *nu_host, long first, long count, int *satisfied_dev) {
  if (count <= 0) return 0;
  SmallVec nu;
  for (int s = 0; s < kMaxS; ++s) nu.v[s] = s < h->S ? nu_host[s] : 0;
  if (!h->lhs) {
    check_shifted_default_kernel<<<blocks_for(count), 256>>>(h->d_states, first, count, h->S, h->K, nu,
                                                             bounds_of(h), satisfied_dev);
    FSP_LAUNCH_CHECK();
    return 0;
  }
  // custom lhs: evaluate on the host (API contract: user std::function), StateSetConstrained.cpp:63-82
  std::vector<int> X((size_t) count * h->S), fval((size_t) count * h->K), sat((size_t) count * h->K);
  int             *d_tmp;
  FSP_CUDA_CHECK(pmalloc(&d_tmp, sizeof(int) * count * h->S));
  shift_states_kernel<<<blocks_for(count), 256>>>(h->d_states, first, count, h->S, nu, 1, d_tmp);
  FSP_LAUNCH_CHECK();
  FSP_CUDA_CHECK(cudaMemcpy(X.data(), d_tmp, sizeof(int) * count * h->S, cudaMemcpyDeviceToHost));
  pfree(d_tmp);
  int ierr = h->lhs(h->S, h->K, (int) count, X.data(), fval.data(), h->lhs_args);
  if (ierr) { set_error("fspset: constraint callback returned %d", ierr); return ierr; }
  for (int k = 0; k < h->K; ++k)
    for (long i = 0; i < count; ++i) {
      int ok = fval[(size_t) h->K * i + k] <= h->bounds[k] ? 1 : 0;
      for (int s = 0; s < h->S; ++s)
        if (X[(size_t) h->S * i + s] < 0) ok = 1;
      sat[(size_t) k * count + i] = ok;
    }
  FSP_CUDA_CHECK(cudaMemcpy(satisfied_dev, sat.data(), sizeof(int) * count * h->K, cudaMemcpyHostToDevice));
  return 0;
}

int fspset_states_dev(fspset_t h, const int **states_dev) { *states_dev = h->d_states; return 0; }

int fspset_copy_states(fspset_t h, long first, long count, int *out) {
  if (count > 0)
    FSP_CUDA_CHECK(cudaMemcpy(out, h->d_states + (size_t) first * h->S, sizeof(int) * count * h->S, cudaMemcpyDeviceToHost));
  return 0;
}
int fspset_copy_status(fspset_t h, long first, long count, signed char *out) {
  if (count > 0) FSP_CUDA_CHECK(cudaMemcpy(out, h->d_status + first, count, cudaMemcpyDeviceToHost));
  return 0;
}

int fspset_eval_separable(fspset_t h, double rate, const int *order_host, const double *tables_dev, const int *tab_off_host,
                          const int *tab_len_host, const int *nu_host, int sign, long first, long count, double *out_dev) {
  if (count <= 0) return 0;
  SmallVec nu, ord, off, len;
  for (int s = 0; s < kMaxS; ++s) {
    nu.v[s] = s < h->S ? nu_host[s] : 0;
    ord.v[s] = s < h->S ? order_host[s] : 0;
    off.v[s] = (s < h->S && tables_dev && tab_off_host) ? tab_off_host[s] : 0;
    len.v[s] = (s < h->S && tables_dev && tab_len_host) ? tab_len_host[s] : 0;
  }
  mass_action_kernel<<<blocks_for(count), 256>>>(h->d_states, first, count, h->S, nu, sign, ord, rate, tables_dev, off, len, out_dev);
  FSP_LAUNCH_CHECK();
  return 0;
}

int fspset_eval_mass_action(fspset_t h, double rate, const int *order_host, const int *nu_host, int sign, long first,
                            long count, double *out_dev) {
  return fspset_eval_separable(h, rate, order_host, nullptr, nullptr, nullptr, nu_host, sign, first, count, out_dev);
}


// Sink column lists (FspMatrixConstrained.cpp:170-194): for each constraint k the ascending local indices
// i - first of stored states whose destination state_i + nu violates constraint k.  Lists are written
// k after k into idx_out_dev (capacity cap entries); counts_host[k] receives their lengths.
// Only states whose status is not 0 can contribute: status 0 means every child of the state is inside the set
// (StateSetConstrained.cpp:184-198), hence satisfies all constraints -- so the constraint evaluation (a HOST callback
// for custom shapes) runs on the O(surface) subset only.  That shortcut needs the bounds to be no smaller than at
// the last Expand(); otherwise all states are examined.
int fspset_sink_lists(fspset_t h, const int *nu_host, long first, long count, int *idx_out_dev, long cap,
                      long *counts_host) {
  for (int k = 0; k < h->K; ++k) counts_host[k] = 0;
  if (count <= 0) return 0;
  SmallVec nu;
  for (int s = 0; s < kMaxS; ++s) nu.v[s] = s < h->S ? nu_host[s] : 0;
  bool use_status = h->expanded && (int) h->expanded_bounds.size() == h->K;
  for (int k = 0; k < h->K && use_status; ++k) use_status = h->bounds[k] >= h->expanded_bounds[k];
  // candidate list (ascending positions relative to `first`)
  int *d_cand = nullptr, *d_num = nullptr;
  FSP_CUDA_CHECK(pmalloc(&d_cand, sizeof(int) * count));
  FSP_CUDA_CHECK(pmalloc(&d_num, sizeof(int)));
  long m = count;
  {
    cub::CountingInputIterator<int> iota(0);
    StatusNonZero pred{h->d_status + first, use_status ? 1 : 0};
    size_t need = 0;
    cub::DeviceSelect::If(nullptr, need, iota, d_cand, d_num, (int) count, pred);
    if (need > h->cub_bytes) { pfree(h->d_cub); FSP_CUDA_CHECK(pmalloc(&h->d_cub, need)); h->cub_bytes = need; }
    FSP_CUDA_CHECK(cub::DeviceSelect::If(h->d_cub, need, iota, d_cand, d_num, (int) count, pred));
    count_launch();
    int num = 0;
    FSP_CUDA_CHECK(cudaMemcpy(&num, d_num, sizeof(int), cudaMemcpyDeviceToHost));
    m = num;
  }
  int rc = 0;
  if (m > 0) {
    int *d_sat = nullptr, *d_out = nullptr;
    FSP_CUDA_CHECK(pmalloc(&d_sat, sizeof(int) * m * h->K));
    FSP_CUDA_CHECK(pmalloc(&d_out, sizeof(int) * m));
    if (!h->lhs) {
      check_list_default_kernel<<<blocks_for(m), 256>>>(h->d_states, first, d_cand, m, h->S, h->K, nu, bounds_of(h), d_sat);
      FSP_LAUNCH_CHECK();
    } else {
      // custom lhs: the user's host callback on the candidate destinations only
      int *d_tmp = nullptr;
      FSP_CUDA_CHECK(pmalloc(&d_tmp, sizeof(int) * m * h->S));
      shift_list_kernel<<<blocks_for(m), 256>>>(h->d_states, first, d_cand, m, h->S, nu, d_tmp);
      FSP_LAUNCH_CHECK();
      std::vector<int> X((size_t) m * h->S), fval((size_t) m * h->K), sat((size_t) m * h->K);
      FSP_CUDA_CHECK(cudaMemcpy(X.data(), d_tmp, sizeof(int) * m * h->S, cudaMemcpyDeviceToHost));
      pfree(d_tmp);
      int ierr = h->lhs(h->S, h->K, (int) m, X.data(), fval.data(), h->lhs_args);
      if (ierr) { set_error("fspset: constraint callback returned %d", ierr); pfree(d_sat); pfree(d_out); pfree(d_cand); pfree(d_num); return ierr; }
      for (int k = 0; k < h->K; ++k)
        for (long i = 0; i < m; ++i) {
          int ok = fval[(size_t) h->K * i + k] <= h->bounds[k] ? 1 : 0;
          for (int s = 0; s < h->S; ++s)
            if (X[(size_t) h->S * i + s] < 0) ok = 1;
          sat[(size_t) k * m + i] = ok;
        }
      FSP_CUDA_CHECK(cudaMemcpy(d_sat, sat.data(), sizeof(int) * m * h->K, cudaMemcpyHostToDevice));
    }
    long written = 0;
    for (int k = 0; k < h->K && !rc; ++k) {
      NotFlag flags{d_sat + (size_t) k * m};
      cub::TransformInputIterator<int, NotFlag, cub::CountingInputIterator<int>> fl(cub::CountingInputIterator<int>(0), flags);
      size_t need = 0;
      cub::DeviceSelect::Flagged(nullptr, need, d_cand, fl, d_out, d_num, (int) m);
      if (need > h->cub_bytes) { pfree(h->d_cub); if (pmalloc(&h->d_cub, need) != cudaSuccess) { rc = -1; break; } h->cub_bytes = need; }
      if (cub::DeviceSelect::Flagged(h->d_cub, need, d_cand, fl, d_out, d_num, (int) m) != cudaSuccess) {
        set_error("fspset_sink_lists: select failed"); rc = -1; break;
      }
      count_launch();
      int num = 0;
      cudaMemcpy(&num, d_num, sizeof(int), cudaMemcpyDeviceToHost);
      if (written + num > cap) { set_error("fspset_sink_lists: capacity %ld too small", cap); rc = -1; break; }
      if (num > 0) cudaMemcpyAsync(idx_out_dev + written, d_out, sizeof(int) * num, cudaMemcpyDeviceToDevice, 0);
      counts_host[k] = num;
      written += num;
    }
    pfree(d_sat); pfree(d_out);
  }
  pfree(d_cand); pfree(d_num);
  return rc;
}

/* number of stored states in [first, first+count) that can contribute sink entries (status != 0) */
int fspset_num_boundary_states(fspset_t h, long first, long count, long *n) {
  *n = count;
  if (count <= 0 || !h->expanded) return 0;
  if (ensure_flags(h, count)) return -1;
  status_nonzero_flag_kernel<<<blocks_for(count), 256>>>(h->d_status + first, count, h->d_flag);
  FSP_LAUNCH_CHECK();
  long tot = 0;
  if (scan_flags(h, count, &tot)) return -1;
  *n = tot;
  return 0;
}

}  // extern "C"

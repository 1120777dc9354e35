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
//
// Sharded mode (fspset_set_sharded, N GPUs of one node): each rank keeps only its block of states; the directory is a
// table of 64-bit slots striped over the ranks' HBM (shard = hash % N) inside CUDA-IPC peer windows, probed with NVLink
// peer loads and claimed with system-scope atomicCAS / atomicMin.  A slot holds (is_candidate, rank, position): the key
// is read from the owner's state list -- or, while a batch is in flight, from the owner's candidate buffer -- through
// the same peer mapping.  Every rank expands the frontier states it owns; duplicates across ranks are resolved by
// atomicMin (existing states beat candidates, then the lowest rank, then the lowest position: deterministic for a given
// N); winners are appended on the discovering rank.  Three stream-ordered barriers per batch (candidates written /
// claims complete / final values stored) are the only synchronisation.  The set is then re-balanced to the contiguous
// equal-count BLOCK layout by peer-to-peer copies that preserve the rank-concatenated order, and the directory is
// rebuilt with the new positions.
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

// ---- sharded directory (see the file header) -------------------------------------------------------------------
constexpr int                kMaxRanksS = FSP_P2P_MAX_RANKS;
constexpr unsigned long long kEmpty64 = ~0ull;
constexpr unsigned long long kCandBit = 1ull << 62;       // the slot names a candidate of the batch in flight
constexpr unsigned long long kIdxMask = (1ull << 40) - 1;  // position inside the owner's state list / candidate buffer

struct ShardView {
  int                 size, rank, S;
  unsigned long long  mask;  // slots per shard - 1
  unsigned long long *table[kMaxRanksS];
  const int          *states[kMaxRanksS];
  const int          *cand[kMaxRanksS];
  int                 starts[kMaxRanksS + 1];
  int                *err;
};

__device__ __forceinline__ const int *sh_key_of(const ShardView &v, unsigned long long val) {
  const int r = (int) ((val >> 40) & 0xFFFFull);
  return ((val & kCandBit) ? v.cand[r] : v.states[r]) + (size_t) (val & kIdxMask) * v.S;
}
// `there` may live in a peer's HBM and may have been written by a kernel of that peer: read through L2, not L1
__device__ __forceinline__ bool sh_key_equal(const int *there, const int *key, int S) {
  for (int s = 0; s < S; ++s)
    if (__ldcg(there + s) != key[s]) return false;
  return true;
}
__device__ __forceinline__ void sh_home(const ShardView &v, const int *key, int &shard, unsigned long long &slot) {
  const unsigned long long h = hash_state(key, v.S);
  shard = (int) ((h >> 40) % (unsigned long long) v.size);
  slot = h & v.mask;
}
__device__ __forceinline__ unsigned long long sh_load(const unsigned long long *p) {
  return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// Claim: first pass of a batch.  Candidate c of this rank carries the value (candidate, rank, c).
__global__ void sh_insert_kernel(const __grid_constant__ ShardView v, long m, const signed char *valid) {
  const long c = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  if (valid && !valid[c]) return;
  const int               *key = v.cand[v.rank] + (size_t) c * v.S;
  const unsigned long long me = kCandBit | ((unsigned long long) v.rank << 40) | (unsigned long long) c;
  int                      shard;
  unsigned long long       slot;
  sh_home(v, key, shard, slot);
  unsigned long long *table = v.table[shard];
  for (unsigned long long probe = 0; probe <= v.mask; ++probe) {
    unsigned long long cur = sh_load(table + slot);
    if (cur == kEmpty64) {
      const unsigned long long prev = atomicCAS_system(table + slot, kEmpty64, me);
      if (prev == kEmpty64) return;
      cur = prev;
    }
    if (sh_key_equal(sh_key_of(v, cur), key, v.S)) {
      atomicMin_system(table + slot, me);  // a stored state (no candidate bit) always stays
      return;
    }
    slot = (slot + 1) & v.mask;
  }
  atomicExch(v.err, 1);
}

// Second pass (after every rank's claims are complete): flag[c] = 1 iff candidate c holds the slot of its key.
__global__ void sh_winner_kernel(const __grid_constant__ ShardView v, long m, const signed char *valid, int *flag,
                                 unsigned long long *slot_out) {
  const long c = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  int f = 0;
  if (!valid || valid[c]) {
    const int               *key = v.cand[v.rank] + (size_t) c * v.S;
    const unsigned long long me = kCandBit | ((unsigned long long) v.rank << 40) | (unsigned long long) c;
    int                      shard;
    unsigned long long       slot;
    sh_home(v, key, shard, slot);
    const unsigned long long *table = v.table[shard];
    for (unsigned long long probe = 0; probe <= v.mask; ++probe) {
      const unsigned long long cur = sh_load(table + slot);
      if (cur == kEmpty64) break;  // cannot happen for a claimed key
      if (cur == me) { f = 1; slot_out[c] = ((unsigned long long) shard << 56) | slot; break; }
      // another rank may already be replacing its winning candidates by stored states: either names an equal key
      if (sh_key_equal(sh_key_of(v, cur), key, v.S)) break;
      slot = (slot + 1) & v.mask;
    }
  }
  flag[c] = f;
}

// Third pass: winners become states of THIS rank; the slot receives (rank, position) after the key is in place.
__global__ void sh_append_kernel(const __grid_constant__ ShardView v, int *states, signed char *status, long n_old, long m,
                                 const int *flag, const int *pos, const unsigned long long *slot_in) {
  const long c = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m || !flag[c]) return;
  const long dst = n_old + pos[c];
  const int *key = v.cand[v.rank] + (size_t) c * v.S;
  for (int s = 0; s < v.S; ++s) states[(size_t) dst * v.S + s] = key[s];
  status[dst] = 1;
  __threadfence_system();
  const unsigned long long sl = slot_in[c];
  atomicExch_system(v.table[(int) (sl >> 56)] + (sl & ((1ull << 56) - 1)), ((unsigned long long) v.rank << 40) | (unsigned long long) dst);
}

// Rebuild: every stored state of this rank claims a slot (all keys are distinct).
__global__ void sh_rehash_kernel(const __grid_constant__ ShardView v, long n) {
  const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int               *key = v.states[v.rank] + (size_t) i * v.S;
  const unsigned long long me = ((unsigned long long) v.rank << 40) | (unsigned long long) i;
  int                      shard;
  unsigned long long       slot;
  sh_home(v, key, shard, slot);
  unsigned long long *table = v.table[shard];
  for (unsigned long long probe = 0; probe <= v.mask; ++probe) {
    if (atomicCAS_system(table + slot, kEmpty64, me) == kEmpty64) return;
    slot = (slot + 1) & v.mask;
  }
  atomicExch(v.err, 1);
}

// State2Index on the striped directory: global index = starts[owner] + position
__device__ __forceinline__ int sh_lookup(const ShardView &v, const int *key) {
  for (int s = 0; s < v.S; ++s)
    if (key[s] < 0) return -1;
  int                shard;
  unsigned long long slot;
  sh_home(v, key, shard, slot);
  const unsigned long long *table = v.table[shard];
  for (unsigned long long probe = 0; probe <= v.mask; ++probe) {
    const unsigned long long cur = __ldcg(table + slot);
    if (cur == kEmpty64) return -1;
    if (sh_key_equal(sh_key_of(v, cur), key, v.S)) return v.starts[(int) ((cur >> 40) & 0xFFFFull)] + (int) (cur & kIdxMask);
    slot = (slot + 1) & v.mask;
  }
  return -1;
}
__global__ void sh_lookup_kernel(const __grid_constant__ ShardView v, const int *X, long m, int *idx) {
  const long j = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  int key[kMaxS];
  for (int s = 0; s < v.S; ++s) key[s] = X[(size_t) j * v.S + s];
  idx[j] = sh_lookup(v, key);
}
__global__ void sh_lookup_shifted_kernel(const __grid_constant__ ShardView v, long first_local, long count, SmallVec nu,
                                         int sign, int *idx) {
  const long j = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  int        key[kMaxS];
  const int *x = v.states[v.rank] + (size_t) (first_local + j) * v.S;
  for (int s = 0; s < v.S; ++s) key[s] = x[s] + sign * nu.v[s];
  idx[j] = sh_lookup(v, key);
}

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
  long             base = 0;           // global index of local state 0 (0 unless sharded)
  int             *d_rem = nullptr;    // fspset_remember_local
  long             n_rem = 0;
  // ---- sharded mode: n / cap / d_states / d_status / d_cand describe the LOCAL block ----
  bool               sharded = false;
  fspcomm_s         *comm = nullptr;
  int                rank = 0, size = 1;
  long               n_glob = 0;
  std::vector<long>  counts, starts;   // states per rank now / BLOCK layout after the last re-balance (size + 1)
  void              *win_data[FSP_P2P_MAX_RANKS] = {nullptr};   // [states cap*S int | status cap bytes], same cap on all ranks
  void              *win_cand[FSP_P2P_MAX_RANKS] = {nullptr};   // candidate keys of the batch in flight
  void              *win_table[FSP_P2P_MAX_RANKS] = {nullptr};  // tsize 64-bit slots per rank
  void              *win_spare[FSP_P2P_MAX_RANKS] = {nullptr};  // second data window of the same capacity: re-balance target
  bool               have_spare = false;
  size_t             data_bytes = 0, cand_bytes = 0, table_bytes = 0;
  int               *d_err = nullptr;  // probe bound exceeded (a full shard)
  // FSP_SET_TRACE=1: what the collective construction spent its time on
  struct Trace { long waves = 0, batches = 0, data_resizes = 0, table_resizes = 0, cand_resizes = 0, rebalances = 0;
                 double t_resize = 0, t_rebalance = 0, t_batches = 0, t_total = 0; } tr;
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


// ---- sharded mode, host side (every function below is COLLECTIVE: same call sequence and arguments on all ranks) ----
bool sh_trace() { static const bool on = [] { const char *e = getenv("FSP_SET_TRACE"); return e && e[0] == '1'; }(); return on; }
double sh_now() {
  if (!sh_trace()) return 0.0;
  cudaDeviceSynchronize();
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}
ShardView make_view(const fspset_s *h) {
  ShardView v;
  memset(&v, 0, sizeof(v));
  v.size = h->size; v.rank = h->rank; v.S = h->S;
  v.mask = h->tsize ? h->tsize - 1 : 0;
  for (int p = 0; p < h->size; ++p) {
    v.table[p] = (unsigned long long *) h->win_table[p];
    v.states[p] = (const int *) h->win_data[p];
    v.cand[p] = (const int *) h->win_cand[p];
  }
  for (int p = 0; p <= h->size; ++p) v.starts[p] = (int) h->starts[(size_t) p];
  v.err = h->d_err;
  return v;
}

int sh_barrier(fspset_s *h) { return fspcomm_barrier(h->comm, nullptr); }

int sh_check(fspset_s *h, const char *where) {
  int e = 0;
  FSP_CUDA_CHECK(cudaMemcpy(&e, h->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) { set_error("fspset (sharded), %s: a directory shard is full", where); return -1; }
  return fspcomm_check(h->comm);
}

int sh_gather(fspset_s *h, long mine, std::vector<long> &all) {
  all.assign((size_t) h->size, 0);
  return fspcomm_gather_long(h->comm, mine, all.data());
}

size_t sh_data_bytes(const fspset_s *h, long cap) {
  return (((size_t) cap * h->S * sizeof(int) + (size_t) cap) + 255) / 256 * 256;
}

// local block -> a window of capacity new_cap (the old window is released)
int sh_data_resize(fspset_s *h, long new_cap) {
  void *nw[kMaxRanksS] = {nullptr};
  const size_t bytes = sh_data_bytes(h, new_cap);
  if (fspcomm_window_create(h->comm, bytes, nw)) return -1;
  int         *ns = (int *) nw[h->rank];
  signed char *nst = (signed char *) (ns + (size_t) new_cap * h->S);
  if (h->n > 0) {
    FSP_CUDA_CHECK(cudaMemcpy(ns, h->d_states, sizeof(int) * (size_t) h->n * h->S, cudaMemcpyDeviceToDevice));
    FSP_CUDA_CHECK(cudaMemcpy(nst, h->d_status, (size_t) h->n, cudaMemcpyDeviceToDevice));
  }
  FSP_CUDA_CHECK(cudaDeviceSynchronize());
  if (h->win_data[h->rank] && fspcomm_window_destroy(h->comm, h->win_data)) return -1;
  if (h->have_spare) {  // wrong capacity now
    if (fspcomm_window_destroy(h->comm, h->win_spare)) return -1;
    h->have_spare = false;
  }
  for (int p = 0; p < h->size; ++p) h->win_data[p] = nw[p];
  h->d_states = ns; h->d_status = nst; h->cap = new_cap; h->data_bytes = bytes;
  return 0;
}

int sh_rehash_all(fspset_s *h, const char *where) {
  if (sh_barrier(h)) return -1;  // every shard is cleared / every block is in place
  if (h->n > 0) {
    sh_rehash_kernel<<<blocks_for(h->n), 256>>>(make_view(h), h->n);
    FSP_LAUNCH_CHECK();
  }
  if (sh_barrier(h)) return -1;
  return sh_check(h, where);
}

int sh_table_resize(fspset_s *h, unsigned long long want) {
  void *nw[kMaxRanksS] = {nullptr};
  const size_t bytes = sizeof(unsigned long long) * want;
  if (fspcomm_window_create(h->comm, bytes, nw)) return -1;
  FSP_CUDA_CHECK(cudaMemset(nw[h->rank], 0xFF, bytes));
  FSP_CUDA_CHECK(cudaDeviceSynchronize());
  if (h->win_table[h->rank] && fspcomm_window_destroy(h->comm, h->win_table)) return -1;
  for (int p = 0; p < h->size; ++p) h->win_table[p] = nw[p];
  h->tsize = want; h->table_bytes = bytes;
  return sh_rehash_all(h, "table growth");
}

int sh_cand_resize(fspset_s *h, long m) {
  void *nw[kMaxRanksS] = {nullptr};
  const size_t bytes = (sizeof(int) * (size_t) m * h->S + 255) / 256 * 256;
  if (h->win_cand[h->rank] && fspcomm_window_destroy(h->comm, h->win_cand)) return -1;
  if (fspcomm_window_create(h->comm, bytes, nw)) return -1;
  for (int p = 0; p < h->size; ++p) h->win_cand[p] = nw[p];
  h->d_cand = (int *) nw[h->rank];
  h->cand_bytes = bytes;
  pfree(h->d_valid); pfree(h->d_slot);
  h->d_valid = nullptr; h->d_slot = nullptr;
  FSP_CUDA_CHECK(pmalloc(&h->d_valid, (size_t) m));
  FSP_CUDA_CHECK(pmalloc(&h->d_slot, sizeof(unsigned long long) * (size_t) m));
  h->cand_cap = m;
  return 0;
}

// room for `max_loc_after` states on the fullest rank, `glob_after` states in the directory, `max_batch` candidates
int sh_reserve(fspset_s *h, long max_loc_after, long glob_after, long max_batch) {
  if (glob_after >= 0x7FFFFFF0L) { set_error("fspset: more than 2^31 states"); return -1; }
  // every resize re-creates a peer window (allocation, IPC handle exchange, mapping by every peer: milliseconds), so
  // capacities start generous and double
  const double t0 = sh_now();
  if (max_loc_after > h->cap) {
    const long nc = std::max(std::max(max_loc_after, 2 * h->cap), 1L << 16);
    if (sh_data_resize(h, nc)) return -1;
    h->tr.data_resizes += 1;
  }
  unsigned long long want = h->tsize ? h->tsize : (1ull << 16);
  const unsigned long long per_shard = (unsigned long long) ((glob_after + h->size - 1) / h->size);
  while (per_shard * 5 / 2 + 64 > want) want *= 2;  // average load <= 0.4; the shards of a good hash differ by O(sqrt)
  if (want != h->tsize) {
    if (sh_table_resize(h, want)) return -1;
    h->tr.table_resizes += 1;
  }
  if (max_batch > h->cand_cap) {
    if (sh_cand_resize(h, std::max(std::max(max_batch, 2 * h->cand_cap), 1L << 18))) return -1;
    h->tr.cand_resizes += 1;
  }
  h->tr.t_resize += sh_now() - t0;
  return 0;
}

// One batch: the m local candidates in h->d_cand (validity in h->d_valid unless all_valid).  m may be 0 on this rank.
int sh_insert_batch_impl(fspset_s *h, long m, bool all_valid);
int sh_insert_batch(fspset_s *h, long m, bool all_valid) {
  const double t0 = sh_now();
  const int    rc = sh_insert_batch_impl(h, m, all_valid);
  h->tr.t_batches += sh_now() - t0;
  h->tr.batches += 1;
  return rc;
}
int sh_insert_batch_impl(fspset_s *h, long m, bool all_valid) {
  const ShardView     v = make_view(h);
  const signed char *valid = all_valid ? nullptr : h->d_valid;
  // every rank's candidates are written, the previous batch is complete.  With a host constraint callback the ranks can
  // be seconds apart here (the rank that owns the frontier evaluated it): meet through NCCL, which has no time limit,
  // before the flag barrier, whose device-side wait gives up after FSP_SPIN_TIMEOUT_MS
  if (h->lhs && fspcomm_barrier_sync(h->comm)) return -1;
  if (sh_barrier(h)) return -1;
  if (m > 0) {
    sh_insert_kernel<<<blocks_for(m), 256>>>(v, m, valid);
    FSP_LAUNCH_CHECK();
  }
  if (sh_barrier(h)) return -1;  // all claims are in
  if (m > 0) {
    if (ensure_flags(h, m)) return -1;
    sh_winner_kernel<<<blocks_for(m), 256>>>(v, m, valid, h->d_flag, h->d_slot);
    FSP_LAUNCH_CHECK();
    long total = 0;
    if (scan_flags(h, m, &total)) return -1;
    if (total > 0) {
      if (h->n + total > h->cap) { set_error("fspset (sharded): local capacity %ld exceeded", h->cap); return -1; }
      sh_append_kernel<<<blocks_for(m), 256>>>(v, h->d_states, h->d_status, h->n, m, h->d_flag, h->d_pos, h->d_slot);
      FSP_LAUNCH_CHECK();
      h->n += total;
    }
  }
  if (sh_barrier(h)) return -1;  // winners are stored everywhere; the candidate buffers may be overwritten
  return sh_check(h, "insertion");
}

// Re-balance to the BLOCK layout (h->counts must be current).  The rank-concatenated order of the states is kept:
// rank r pulls positions [T_r, T_r+1) of the concatenation out of the peers' windows; then the directory is rebuilt.
int sh_rebalance_impl(fspset_s *h);
int sh_rebalance(fspset_s *h) {
  const double t0 = sh_now();
  const int    rc = sh_rebalance_impl(h);
  h->tr.t_rebalance += sh_now() - t0;
  return rc;
}
int sh_rebalance_impl(fspset_s *h) {
  const int N = h->size;
  std::vector<long> T((size_t) N + 1, 0), seg_src((size_t) N), seg_off((size_t) N), seg_dst((size_t) N), seg_len((size_t) N);
  int               n_seg = 0, same_i = 0;
  if (fspset_rebalance_plan(N, h->counts.data(), h->rank, T.data(), seg_src.data(), seg_off.data(), seg_dst.data(), seg_len.data(),
                            &n_seg, &same_i)) return -1;
  h->n_glob = T[(size_t) N];
  const bool same = same_i != 0;
  if (!same) {
    h->tr.rebalances += 1;
    // target: the spare window (same capacity as the current one, created at the first re-balance after a growth and
    // kept: the two swap roles, so a re-balance costs no allocation / handle exchange / mapping)
    const long   new_cap = h->cap;  // >= the fullest rank's count >= every target count
    void        *nw[kMaxRanksS] = {nullptr};
    const size_t bytes = sh_data_bytes(h, new_cap);
    if (h->have_spare) {
      for (int p = 0; p < N; ++p) nw[p] = h->win_spare[p];
    } else if (fspcomm_window_create(h->comm, bytes, nw)) {
      return -1;
    }
    FSP_CUDA_CHECK(cudaDeviceSynchronize());
    if (sh_barrier(h)) return -1;  // every block is final, every new window exists
    FSP_CUDA_CHECK(cudaDeviceSynchronize());
    int         *ns = (int *) nw[h->rank];
    signed char *nst = (signed char *) (ns + (size_t) new_cap * h->S);
    const long   lo_me = T[(size_t) h->rank], hi_me = T[(size_t) h->rank + 1];
    for (int k = 0; k < n_seg; ++k) {  // pull: seg_len states from position seg_off of rank seg_src to my position seg_dst
      const int         *src = (const int *) h->win_data[seg_src[(size_t) k]];
      const signed char *src_st = (const signed char *) (src + (size_t) h->cap * h->S);
      FSP_CUDA_CHECK(cudaMemcpyAsync(ns + (size_t) seg_dst[(size_t) k] * h->S, src + (size_t) seg_off[(size_t) k] * h->S,
                                     sizeof(int) * (size_t) seg_len[(size_t) k] * h->S, cudaMemcpyDefault, 0));
      FSP_CUDA_CHECK(cudaMemcpyAsync(nst + seg_dst[(size_t) k], src_st + seg_off[(size_t) k], (size_t) seg_len[(size_t) k], cudaMemcpyDefault, 0));
    }
    FSP_CUDA_CHECK(cudaDeviceSynchronize());
    // the old window becomes the spare: nobody writes to it before the next re-balance, whose barrier comes after every
    // rank has finished these pulls
    for (int p = 0; p < N; ++p) { h->win_spare[p] = h->win_data[p]; h->win_data[p] = nw[p]; }
    h->have_spare = true;
    h->d_states = ns; h->d_status = nst; h->cap = new_cap; h->data_bytes = bytes;
    h->n = hi_me - lo_me;
    for (int r = 0; r < N; ++r) h->counts[(size_t) r] = T[(size_t) r + 1] - T[(size_t) r];
  }
  h->base = T[(size_t) h->rank];
  h->starts = T;
  if (!same) {
    FSP_CUDA_CHECK(cudaMemsetAsync(h->win_table[h->rank], 0xFF, h->table_bytes, 0));
    if (sh_rehash_all(h, "re-balance")) return -1;
  }
  return 0;
}

int sh_refresh_counts(fspset_s *h) {
  if (sh_gather(h, h->n, h->counts)) return -1;
  h->n_glob = 0;
  for (long c : h->counts) h->n_glob += c;
  return 0;
}

// AddStates / the lattice: `fill(b, mb)` writes the candidates [b, b + mb) of this rank's list into h->d_cand
template <class Fill>
int sh_add(fspset_s *h, long m, Fill fill) {
  std::vector<long> ms;
  if (sh_gather(h, m, ms)) return -1;
  long maxM = 0, sumM = 0, maxN = 0;
  for (int p = 0; p < h->size; ++p) { maxM = std::max(maxM, ms[(size_t) p]); sumM += ms[(size_t) p]; maxN = std::max(maxN, h->counts[(size_t) p]); }
  if (maxM == 0) return 0;
  if (sh_reserve(h, maxN + maxM, h->n_glob + sumM, std::min(kBatch, maxM))) return -1;
  for (long b = 0; b < maxM; b += kBatch) {
    const long mb = std::max(0L, std::min(kBatch, m - b));
    if (mb > 0 && fill(b, mb)) return -1;
    if (sh_insert_batch(h, mb, true)) return -1;
  }
  if (sh_refresh_counts(h)) return -1;
  return sh_rebalance(h);
}

int sh_expand_impl(fspset_s *h, int *&d_frontier, signed char *&d_fstatus);
int sh_expand(fspset_s *h, int *&d_frontier, signed char *&d_fstatus) {
  h->tr = fspset_s::Trace();
  const double t0 = sh_now();
  const int    rc = sh_expand_impl(h, d_frontier, d_fstatus);
  h->tr.t_total = sh_now() - t0;
  if (sh_trace() && h->rank == 0)
    printf("[set] sharded Expand -> %ld states: %.1f ms | %ld waves, %ld batches %.1f ms | resizes data %ld table %ld cand %ld: %.1f ms | "
           "%ld re-balances (incl. the final one) %.1f ms\n", h->n_glob, 1e3 * h->tr.t_total, h->tr.waves, h->tr.batches,
           1e3 * h->tr.t_batches, h->tr.data_resizes, h->tr.table_resizes, h->tr.cand_resizes, 1e3 * h->tr.t_resize, h->tr.rebalances,
           1e3 * h->tr.t_rebalance);
  return rc;
}
int sh_expand_impl(fspset_s *h, int *&d_frontier, signed char *&d_fstatus) {
  const int use_default = h->lhs ? 0 : 1;
  const int N = h->size;
  if (h->n > 0) {
    reactivate_kernel<<<blocks_for(h->n), 256>>>(h->d_status, h->n);
    FSP_LAUNCH_CHECK();
  }
  long                     fcap = 0;
  std::vector<int>         cand_host, fval;
  std::vector<signed char> valid_host;
  std::vector<long>        nFs;
  while (true) {
    const long n = h->n;
    long       nF = 0;
    if (n > 0) {
      if (ensure_flags(h, n)) return -1;
      status_flag_kernel<<<blocks_for(n), 256>>>(h->d_status, n, 1, h->d_flag);
      FSP_LAUNCH_CHECK();
      if (scan_flags(h, n, &nF)) return -1;
    }
    if (sh_gather(h, nF, nFs)) return -1;
    if (sh_refresh_counts(h)) return -1;
    long sumF = 0, maxF = 0, maxN = 0;
    for (int p = 0; p < N; ++p) { sumF += nFs[(size_t) p]; maxF = std::max(maxF, nFs[(size_t) p]); maxN = std::max(maxN, h->counts[(size_t) p]); }
    if (sumF == 0) break;
    h->tr.waves += 1;
    // one rank holds far more than its share (the frontier of a BFS-ordered BLOCK layout sits on the last ranks):
    // re-balance before growing further, then look at the frontier again (statuses travel with the states)
    if (maxN > (h->n_glob / N + 1) * 5 / 4 + (1L << 16)) {
      if (sh_rebalance(h)) return -1;
      continue;
    }
    if (nF > fcap) {
      pfree(d_frontier); pfree(d_fstatus);
      d_frontier = nullptr; d_fstatus = nullptr;
      fcap = nF + nF / 2;
      FSP_CUDA_CHECK(pmalloc(&d_frontier, sizeof(int) * fcap));
      FSP_CUDA_CHECK(pmalloc(&d_fstatus, fcap));
    }
    if (nF > 0) {
      frontier_fill_kernel<<<blocks_for(n), 256>>>(h->d_flag, h->d_pos, n, d_frontier);
      FSP_LAUNCH_CHECK();
      FSP_CUDA_CHECK(cudaMemset(d_fstatus, 0, nF));
    }
    // One batch = the children of a chunk of the frontier under ALL reactions (reaction-major inside the batch): the
    // order of discovery need not match the single-rank set, and each batch costs three barriers.
    const Bounds bd = bounds_of(h);
    const long   chunkF = std::max(1L, kBatch / std::max(1, h->R));
    for (long i0 = 0; i0 < maxF; i0 += chunkF) {
      const long mF = std::max(0L, std::min(chunkF, nF - i0));
      const long m = mF * h->R;
      // room for this batch: exact counts (one small gather) + the batch sizes, which every rank can derive
      if (sh_refresh_counts(h)) return -1;
      long maxM = 0, sumM = 0, fullest = 0;
      for (int p = 0; p < N; ++p) {
        const long mp = std::max(0L, std::min(chunkF, nFs[(size_t) p] - i0)) * h->R;
        maxM = std::max(maxM, mp); sumM += mp;
        fullest = std::max(fullest, h->counts[(size_t) p] + mp);
      }
      if (sh_reserve(h, fullest, h->n_glob + sumM, maxM)) return -1;
      for (int j = 0; j < h->R && mF > 0; ++j) {
        children_kernel<<<blocks_for(mF), 256>>>(h->d_states, d_frontier, i0, mF, h->S, h->K, nu_of(h, j), bd, use_default,
                                                 h->d_cand + (size_t) j * mF * h->S, h->d_valid + (size_t) j * mF, d_fstatus);
        FSP_LAUNCH_CHECK();
      }
      if (!use_default && m > 0) {
        if (host_validity(h, m, cand_host, fval, valid_host)) return -1;
        for (int j = 0; j < h->R; ++j) {
          apply_valid_kernel<<<blocks_for(mF), 256>>>(h->d_valid + (size_t) j * mF, i0, mF, d_fstatus);
          FSP_LAUNCH_CHECK();
        }
      }
      if (sh_insert_batch(h, m, false)) return -1;
    }
    if (nF > 0) {
      set_frontier_status_kernel<<<blocks_for(nF), 256>>>(h->d_status, d_frontier, d_fstatus, nF);
      FSP_LAUNCH_CHECK();
    }
  }
  return sh_rebalance(h);  // h->counts is current: nothing was added since the last gather
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
  if (h->sharded) {
    // the windows go back to the communicator's pool: destructors are not collective
    cudaDeviceSynchronize();
    fspcomm_window_retire(h->comm, h->win_data, h->data_bytes);
    if (h->have_spare) fspcomm_window_retire(h->comm, h->win_spare, h->data_bytes);
    fspcomm_window_retire(h->comm, h->win_cand, h->cand_bytes);
    fspcomm_window_retire(h->comm, h->win_table, h->table_bytes);
    pfree(h->d_err);
  } else {
    pfree(h->d_states); pfree(h->d_status); pfree(h->d_table); pfree(h->d_cand);
  }
  pfree(h->d_valid); pfree(h->d_flag); pfree(h->d_pos); pfree(h->d_slot); pfree(h->d_cub); pfree(h->d_rem);
  delete h;
  return 0;
}

int fspset_set_sharded(fspset_t h, fspcomm_s *comm) {
  if (h->n > 0 || h->d_table || h->sharded) { set_error("fspset_set_sharded: the set is not empty"); return -1; }
  int rank = 0, size = 1;
  fspcomm_rank(comm, &rank, &size);
  if (size == 1) return 0;
  if (!fspcomm_p2p_enabled(comm)) { set_error("fspset_set_sharded: the communicator has no peer memory"); return -1; }
  if (size > kMaxRanksS) { set_error("fspset_set_sharded: more than %d ranks", kMaxRanksS); return -1; }
  FSP_CUDA_CHECK(pmalloc(&h->d_err, sizeof(int)));
  FSP_CUDA_CHECK(cudaMemset(h->d_err, 0, sizeof(int)));
  h->sharded = true; h->comm = comm; h->rank = rank; h->size = size;
  h->counts.assign((size_t) size, 0);
  h->starts.assign((size_t) size + 1, 0);
  return 0;
}
int fspset_is_sharded(fspset_t h) { return h && h->sharded ? 1 : 0; }

// Host arithmetic of the re-balance (no device access: also exercised on CPU by tests/test_sharded_plan.py).
// counts[r] states on rank r now; the rank-concatenated listing is re-cut into the BLOCK layout starts[0..n_ranks]
// (ranks below n mod n_ranks own one more); rank `rank` pulls segment k = seg_len[k] states from position seg_off[k] of
// rank seg_src[k] to its own position seg_dst[k] (at most n_ranks segments, ascending).
int fspset_rebalance_plan(int n_ranks, const long *counts, int rank, long *starts, long *seg_src, long *seg_off, long *seg_dst,
                          long *seg_len, int *n_seg, int *already_balanced) {
  if (n_ranks <= 0 || rank < 0 || rank >= n_ranks) return -1;
  std::vector<long> C((size_t) n_ranks + 1, 0);
  for (int r = 0; r < n_ranks; ++r) {
    if (counts[r] < 0) return -1;
    C[(size_t) r + 1] = C[(size_t) r] + counts[r];
  }
  const long n = C[(size_t) n_ranks], base = n / n_ranks, rem = n % n_ranks;
  starts[0] = 0;
  for (int r = 0; r < n_ranks; ++r) starts[r + 1] = starts[r] + base + (r < rem ? 1 : 0);
  int same = 1;
  for (int r = 0; r <= n_ranks; ++r) same &= starts[r] == C[(size_t) r] ? 1 : 0;
  *already_balanced = same;
  const long lo_me = starts[rank], hi_me = starts[rank + 1];
  int        k = 0;
  for (int q = 0; q < n_ranks; ++q) {
    const long lo = std::max(lo_me, C[(size_t) q]), hi = std::min(hi_me, C[(size_t) q + 1]);
    if (lo >= hi) continue;
    seg_src[k] = q; seg_off[k] = lo - C[(size_t) q]; seg_dst[k] = lo - lo_me; seg_len[k] = hi - lo;
    ++k;
  }
  *n_seg = k;
  return 0;
}
int fspset_layout(fspset_t h, long *starts_host, long *n_local) {
  if (h->sharded) for (int p = 0; p <= h->size; ++p) starts_host[p] = h->starts[(size_t) p];
  else { starts_host[0] = 0; starts_host[1] = h->n; }
  if (n_local) *n_local = h->n;
  return 0;
}

int fspset_remember_local(fspset_t h) {
  pfree(h->d_rem);
  h->d_rem = nullptr;
  h->n_rem = h->n;
  if (h->n > 0) {
    FSP_CUDA_CHECK(pmalloc(&h->d_rem, sizeof(int) * (size_t) h->n * h->S));
    FSP_CUDA_CHECK(cudaMemcpy(h->d_rem, h->d_states, sizeof(int) * (size_t) h->n * h->S, cudaMemcpyDeviceToDevice));
  }
  return 0;
}
int fspset_remembered_indices(fspset_t h, int *idx_host, long n_expected) {
  if (n_expected != h->n_rem) { set_error("fspset_remembered_indices: %ld states were remembered, not %ld", h->n_rem, n_expected); return -1; }
  int rc = 0;
  if (h->n_rem > 0) rc = fspset_state2index(h, h->n_rem, h->d_rem, 1, idx_host, 0);
  pfree(h->d_rem);
  h->d_rem = nullptr; h->n_rem = 0;
  return rc;
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
  if (h->sharded)  // collective: every rank passes its own list (StateSetBase.cpp:176-178); duplicates across ranks are shed
    return sh_add(h, m, [&](long b, long mb) {
      FSP_CUDA_CHECK(cudaMemcpy(h->d_cand, X + (size_t) b * h->S, sizeof(int) * mb * h->S,
                                on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
      return 0;
    });
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
  if (h->sharded) {  // every rank contributes its BLOCK of the lattice ordinals
    const long lo = total / h->size * h->rank + std::min<long>(h->rank, total % h->size);
    const long cnt = total / h->size + (h->rank < total % h->size ? 1 : 0);
    return sh_add(h, cnt, [&](long b, long mb) {
      lattice_kernel<<<blocks_for(mb), 256>>>(lo + b, mb, h->S, dims, h->d_cand);
      FSP_LAUNCH_CHECK();
      return 0;
    });
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
  if ((h->sharded ? h->n_glob : h->n) == 0) return 0;
  if ((int) h->bounds.size() != h->K || h->K == 0) { set_error("fspset_expand: shape not set"); return -1; }
  int         *d_frontier = nullptr;
  signed char *d_fstatus = nullptr;
  const int    rc = h->sharded ? sh_expand(h, d_frontier, d_fstatus) : expand_impl(h, d_frontier, d_fstatus);
  pfree(d_frontier); pfree(d_fstatus);
  if (rc == 0) { h->expanded = true; h->expanded_bounds = h->bounds; }
  return rc;
}

int fspset_num_states(fspset_t h, int *n) { *n = (int) (h->sharded ? h->n_glob : h->n); return 0; }

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
  if (h->sharded && h->n_glob > 0) {
    sh_lookup_kernel<<<blocks_for(m), 256>>>(make_view(h), dX, m, dI);
    FSP_LAUNCH_CHECK();
  } else if (h->n == 0 || !h->d_table) {
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
  if (h->sharded) {
    if (first < h->base || first + count > h->base + h->n) { set_error("fspset_lookup_shifted: rows outside the local block"); return -1; }
    sh_lookup_shifted_kernel<<<blocks_for(count), 256>>>(make_view(h), first - h->base, count, nu, sign, idx_dev);
    FSP_LAUNCH_CHECK();
    return 0;
  }
  lookup_shifted_kernel<<<blocks_for(count), 256>>>(h->d_table, h->tsize - 1, h->d_states, first, count, h->S, nu,
                                                    sign, idx_dev);
  FSP_LAUNCH_CHECK();
  return 0;
}

int fspset_check_constraints_shifted(fspset_t h, const int *nu_host, long first, long count, int *satisfied_dev) {
  if (count <= 0) return 0;
  first -= h->base;  // sharded: global -> local row
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

// indexed by GLOBAL state index (sharded: only the local block [base, base + n) is backed by memory)
int fspset_states_dev(fspset_t h, const int **states_dev) {
  *states_dev = h->d_states ? h->d_states - (ptrdiff_t) h->base * h->S : nullptr;
  return 0;
}

int fspset_copy_states(fspset_t h, long first, long count, int *out) {
  first -= h->base;
  if (count > 0)
    FSP_CUDA_CHECK(cudaMemcpy(out, h->d_states + (size_t) first * h->S, sizeof(int) * count * h->S, cudaMemcpyDeviceToHost));
  return 0;
}
int fspset_copy_status(fspset_t h, long first, long count, signed char *out) {
  first -= h->base;
  if (count > 0) FSP_CUDA_CHECK(cudaMemcpy(out, h->d_status + first, count, cudaMemcpyDeviceToHost));
  return 0;
}

int fspset_eval_separable(fspset_t h, double rate, const int *order_host, const double *tables_dev, const int *tab_off_host,
                          const int *tab_len_host, const int *nu_host, int sign, long first, long count, double *out_dev) {
  if (count <= 0) return 0;
  first -= h->base;
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
  first -= h->base;
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
  first -= h->base;
  if (ensure_flags(h, count)) return -1;
  status_nonzero_flag_kernel<<<blocks_for(count), 256>>>(h->d_status + first, count, h->d_flag);
  FSP_LAUNCH_CHECK();
  long tot = 0;
  if (scan_flags(h, count, &tot)) return -1;
  *n = tot;
  return 0;
}

}  // extern "C"

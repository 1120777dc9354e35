/*
 * fsp_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's algorithm for the FSP hot path: state-set expansion and
 * index map, matrix generation, and the multi-pass matrix action.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object; the product (pacmensl_b200/) never does.
 *
 * PARITY PIN: the reference cannot be compiled here (PETSc/SUNDIALS/Zoltan/Armadillo/MPI absent and
 * un-vendored), so this restatement is pinned by the reference's own analytic
 * known-answer tests (tests/test_mat.cpp:146,233; tests/test_fss.cpp:108-125;
 * tests/test_sensmat.cpp:168-223; tests/test_fsp_solver.cpp:264-345) -- see tests/test_oracle_kats.py.
 * The third-party arithmetic it restates: PETSc 3.13.6 MatMult/VecAXPY on MATMPISELL (one SpMV into a
 * work vector followed by an AXPY per matrix), Zoltan_DD (state -> (part, local id) hash directory),
 * Armadillo find_unique (first representative per duplicate group, ascending positions).
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../pacmensl_b200/fixtures/fsp_models.h"

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * State set: local state list + hash directory (np = 1 semantics of StateSetBase / Zoltan_DD)
 * ---------------------------------------------------------------------------------------------- */
typedef struct orc_set {
  int           S, R, K;
  int          *SM;     /* S x R column major */
  int           n, cap; /* number of states, capacity */
  int          *states; /* S x n column major: state i at states[i*S .. i*S+S-1]  (StateSetBase.h:83) */
  signed char  *status; /* 1 active frontier, 0 done, -1 blocked (StateSetConstrained.cpp:137-149) */
  int          *bounds; /* K */
  fsp_constr_fn lhs;    /* NULL => identity (StateSetConstrained.cpp:92-99) */
  void         *lhs_args;
  /* open addressing hash: slot -> state index or -1 */
  int     *table;
  uint64_t tsize; /* power of two */
} orc_set;

static uint64_t orc_hash_state(const int *x, int S) {
  uint64_t h = 0x9E3779B97F4A7C15ull;
  for (int s = 0; s < S; ++s) {
    h ^= (uint64_t) (uint32_t) x[s] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 32;
  }
  return h;
}

static void orc_table_insert_index(orc_set *st, int idx) {
  uint64_t mask = st->tsize - 1;
  uint64_t p    = orc_hash_state(st->states + (size_t) idx * st->S, st->S) & mask;
  while (st->table[p] >= 0) p = (p + 1) & mask;
  st->table[p] = idx;
}

static void orc_table_rebuild(orc_set *st, uint64_t new_size) {
  free(st->table);
  st->tsize = new_size;
  st->table = (int *) malloc(sizeof(int) * new_size);
  for (uint64_t i = 0; i < new_size; ++i) st->table[i] = -1;
  for (int i = 0; i < st->n; ++i) orc_table_insert_index(st, i);
}

static int orc_find(const orc_set *st, const int *x) {
  uint64_t mask = st->tsize - 1;
  uint64_t p    = orc_hash_state(x, st->S) & mask;
  while (1) {
    int idx = st->table[p];
    if (idx < 0) return -1;
    if (memcmp(st->states + (size_t) idx * st->S, x, sizeof(int) * st->S) == 0) return idx;
    p = (p + 1) & mask;
  }
}

static void orc_reserve(orc_set *st, int need) {
  if (need > st->cap) {
    int cap = st->cap ? st->cap : 64;
    while (cap < need) cap *= 2;
    st->states = (int *) realloc(st->states, sizeof(int) * (size_t) cap * st->S);
    st->status = (signed char *) realloc(st->status, (size_t) cap);
    st->cap    = cap;
  }
  if ((uint64_t) need * 2 > st->tsize) {
    uint64_t ts = st->tsize;
    while ((uint64_t) need * 2 > ts) ts *= 2;
    orc_table_rebuild(st, ts);
  }
}

ORC_API orc_set *orc_set_create(int S, int R, const int *SM) {
  orc_set *st = (orc_set *) calloc(1, sizeof(orc_set));
  st->S       = S;
  st->R       = R;
  st->SM      = (int *) malloc(sizeof(int) * S * R);
  memcpy(st->SM, SM, sizeof(int) * S * R);
  st->tsize = 1024;
  st->table = NULL;
  orc_table_rebuild(st, 1024);
  return st;
}

ORC_API void orc_set_destroy(orc_set *st) {
  if (!st) return;
  free(st->SM); free(st->states); free(st->status); free(st->bounds); free(st->table);
  free(st);
}

/* SetShape / SetShapeBounds: StateSetConstrained.cpp:101-126 */
ORC_API int orc_set_set_shape(orc_set *st, int K, fsp_constr_fn lhs, const int *bounds, void *args) {
  st->K = K;
  free(st->bounds);
  st->bounds = (int *) malloc(sizeof(int) * K);
  memcpy(st->bounds, bounds, sizeof(int) * K);
  st->lhs      = lhs;
  st->lhs_args = args;
  if (!lhs && K != st->S) return -1; /* StateSetConstrained.cpp:227-233 */
  return 0;
}

ORC_API int orc_set_set_bounds(orc_set *st, int K, const int *bounds) {
  if (K != st->K) {
    st->K = K;
    free(st->bounds);
    st->bounds = (int *) malloc(sizeof(int) * K);
  }
  memcpy(st->bounds, bounds, sizeof(int) * K);
  return 0;
}

static int orc_eval_lhs(const orc_set *st, int m, int *X, int *out) {
  if (st->lhs) return st->lhs(st->S, st->K, m, X, out, st->lhs_args);
  /* default_constr_fun: StateSetConstrained.cpp:92-99 */
  for (int i = 0; i < m * st->S; ++i) out[i] = X[i];
  return 0;
}

/* AddStates (np == 1): StateSetBase.cpp:188-258.  States already present are shed (:207-216); the rest
 * are appended in the given order with status 1 (:236-241).  Deviation, stated: duplicates *within* X
 * are also shed (first occurrence kept); the reference would store a duplicate column, which no
 * caller relies on (Expand de-duplicates before calling AddStates, StateSetConstrained.cpp:193). */
ORC_API int orc_set_add_states(orc_set *st, int num_species, int m, const int *X) {
  if (st->S != 0 && num_species != st->S) return -1; /* StateSetBase.cpp:190-192 (KAT-S2) */
  for (int j = 0; j < m; ++j) {
    const int *x = X + (size_t) j * st->S;
    if (orc_find(st, x) >= 0) continue;
    orc_reserve(st, st->n + 1);
    memcpy(st->states + (size_t) st->n * st->S, x, sizeof(int) * st->S);
    st->status[st->n] = 1;
    st->n += 1;
    orc_table_insert_index(st, st->n - 1);
  }
  return 0;
}

/* State2Index: StateSetBase.cpp:309-343 -- -1 if any coordinate is negative or the state is absent. */
ORC_API void orc_set_state2index(const orc_set *st, int m, const int *X, int *idx) {
  for (int j = 0; j < m; ++j) {
    const int *x   = X + (size_t) j * st->S;
    int        neg = 0;
    for (int s = 0; s < st->S; ++s)
      if (x[s] < 0) { neg = 1; break; }
    idx[j] = neg ? -1 : orc_find(st, x);
  }
}

/* CheckValidityStates: StateSetConstrained.cpp:33-56 (lhs is evaluated on negative candidates too). */
static int orc_check_validity(const orc_set *st, int m, int *Y, int *out) {
  int *fval = (int *) malloc(sizeof(int) * (size_t) (st->K > 0 ? st->K : 1));
  for (int j = 0; j < m; ++j) {
    out[j] = 0;
    for (int s = 0; s < st->S; ++s)
      if (Y[(size_t) j * st->S + s] < 0) out[j] = -1;
    int ierr = orc_eval_lhs(st, 1, Y + (size_t) j * st->S, fval);
    if (ierr) { free(fval); return ierr; }
    for (int k = 0; k < st->K; ++k)
      if (fval[k] > st->bounds[k]) out[j] = -1;
  }
  free(fval);
  return 0;
}

/* CheckConstraints: StateSetConstrained.cpp:63-82.  satisfied is constraint-major: satisfied[m*k + i];
 * a state with any negative coordinate counts as satisfying every constraint. */
ORC_API int orc_set_check_constraints(const orc_set *st, int m, int *X, int *satisfied) {
  int *fval = (int *) malloc(sizeof(int) * (size_t) m * (st->K > 0 ? st->K : 1) + 4);
  int  ierr = orc_eval_lhs(st, m, X, fval);
  if (ierr) { free(fval); return ierr; }
  for (int k = 0; k < st->K; ++k)
    for (int i = 0; i < m; ++i) {
      satisfied[(size_t) m * k + i] = (fval[(size_t) st->K * i + k] <= st->bounds[k]) ? 1 : 0;
      for (int s = 0; s < st->S; ++s)
        if (X[(size_t) st->S * i + s] < 0) satisfied[(size_t) m * k + i] = 1;
    }
  free(fval);
  return 0;
}

/* Expand: StateSetConstrained.cpp:132-221 (BFS closure; np == 1 so no load balancing, :213).
 * unique_columns (Sys/pacmenMath.h:204-213): first representative of each duplicate group, kept in
 * ascending position order == first-discovery order in the reaction-major child list (:175-179). */
ORC_API int orc_set_expand(orc_set *st) {
  const int S = st->S, R = st->R;
  for (int i = 0; i < st->n; ++i)
    if (st->status[i] == -1) st->status[i] = 1; /* :137-149 */

  int *frontier = NULL, *Y = NULL, *ystatus = NULL;
  while (1) {
    int nF = 0;
    for (int i = 0; i < st->n; ++i) nF += (st->status[i] == 1);
    if (nF == 0) break; /* :159-163 */
    frontier = (int *) realloc(frontier, sizeof(int) * nF);
    nF       = 0;
    for (int i = 0; i < st->n; ++i)
      if (st->status[i] == 1) frontier[nF++] = i;

    size_t nY = (size_t) nF * R;
    Y         = (int *) realloc(Y, sizeof(int) * nY * S);
    ystatus   = (int *) realloc(ystatus, sizeof(int) * nY);
    for (int i = 0; i < nF; ++i) /* :175-179 */
      for (int j = 0; j < R; ++j)
        for (int s = 0; s < S; ++s)
          Y[((size_t) j * nF + i) * S + s] = st->states[(size_t) frontier[i] * S + s] + st->SM[j * S + s];
    int ierr = orc_check_validity(st, (int) nY, Y, ystatus); /* :181 */
    if (ierr) { free(frontier); free(Y); free(ystatus); return ierr; }

    signed char *fstatus = (signed char *) calloc((size_t) nF, 1);
    for (int i = 0; i < nF; ++i) /* :184-190 */
      for (int j = 0; j < R; ++j)
        if (ystatus[(size_t) j * nF + i] < 0) fstatus[i] = -1;

    /* valid children in order; AddStates sheds present ones and in-batch duplicates (first kept) */
    for (size_t c = 0; c < nY; ++c)
      if (ystatus[c] == 0) orc_set_add_states(st, S, 1, Y + c * S); /* :192-194 */

    for (int i = 0; i < nF; ++i) st->status[frontier[i]] = fstatus[i]; /* :198 */
    free(fstatus);
  }
  free(frontier); free(Y); free(ystatus);
  return 0;
}

ORC_API int orc_set_num_states(const orc_set *st) { return st->n; }
ORC_API int orc_set_num_species(const orc_set *st) { return st->S; }
ORC_API int orc_set_num_reactions(const orc_set *st) { return st->R; }
ORC_API int orc_set_num_constraints(const orc_set *st) { return st->K; }
ORC_API const int *orc_set_states(const orc_set *st) { return st->states; }
ORC_API void orc_set_copy_states(const orc_set *st, int *out) {
  memcpy(out, st->states, sizeof(int) * (size_t) st->n * st->S);
}
ORC_API void orc_set_copy_status(const orc_set *st, signed char *out) { memcpy(out, st->status, (size_t) st->n); }

/* ------------------------------------------------------------------------------------------------
 * Matrix: generation (FspMatrixBase.cpp:76-251, FspMatrixConstrained.cpp:121-282)
 * ---------------------------------------------------------------------------------------------- */
typedef struct orc_csr {
  int     nrows, ncols;
  int    *ptr;
  int    *col;
  double *val;
} orc_csr;

static void orc_csr_free(orc_csr *m) {
  free(m->ptr); free(m->col); free(m->val);
  memset(m, 0, sizeof(*m));
}

typedef struct orc_mat {
  int constrained;
  int n;      /* number of states */
  int nrows;  /* n (+K when constrained) */
  int R, K;
  int n_en, n_tv, n_ti;
  int *enabled, *tv, *ti;
  /* per-reaction ELL-like arrays as the reference keeps them (FspMatrixBase.h:184-190), R x n */
  int    *col;  /* offdiag_col_idxs_(i, r): index of x_i - nu_r or -1 */
  double *off;  /* offdiag_vals_(i, r) = prop_x(r, x_i - nu_r) */
  double *diag; /* diag_vals_(i, r)    = prop_x(r, x_i) (positive) */
  /* reference-shaped operators: one matrix per TV reaction + one merged TI matrix */
  orc_csr *tv_mats;
  orc_csr  ti_mat;
  int      has_ti;
  /* sink rows: K x n matrices, per TV reaction + merged TI (FspMatrixConstrained.cpp:199-240) */
  orc_csr *tv_sinks;
  orc_csr  ti_sinks;
  /* raw sink lists per (reaction, constraint), for exporting to the device layout */
  int     *sink_nnz;  /* K x R: sink_nnz[r*K + k] */
  int    **sink_inz;  /* [r*K + k] -> state indices */
  double **sink_val;
  fsp_tcoef_fn t_fun;
  void        *t_fun_args;
  double      *coef; /* R */
  double      *work; /* nrows */
  int          has_values;
} orc_mat;

ORC_API orc_mat *orc_mat_create(int constrained) {
  orc_mat *A     = (orc_mat *) calloc(1, sizeof(orc_mat));
  A->constrained = constrained;
  return A;
}

ORC_API void orc_mat_destroy_values(orc_mat *A) {
  if (!A) return;
  free(A->enabled); free(A->tv); free(A->ti); free(A->col); free(A->off); free(A->diag);
  for (int i = 0; i < A->n_tv; ++i) {
    if (A->tv_mats) orc_csr_free(&A->tv_mats[i]);
    if (A->tv_sinks) orc_csr_free(&A->tv_sinks[i]);
  }
  free(A->tv_mats); free(A->tv_sinks);
  orc_csr_free(&A->ti_mat); orc_csr_free(&A->ti_sinks);
  if (A->sink_inz)
    for (int i = 0; i < A->R * A->K; ++i) { free(A->sink_inz[i]); free(A->sink_val[i]); }
  free(A->sink_nnz); free(A->sink_inz); free(A->sink_val); free(A->coef); free(A->work);
  int c = A->constrained;
  memset(A, 0, sizeof(*A));
  A->constrained = c;
}

ORC_API void orc_mat_destroy(orc_mat *A) {
  orc_mat_destroy_values(A);
  free(A);
}

typedef struct { int col; double val; } orc_ent;
static int orc_ent_cmp(const void *a, const void *b) {
  int ca = ((const orc_ent *) a)->col, cb = ((const orc_ent *) b)->col;
  return (ca > cb) - (ca < cb);
}

/* Assemble a CSR from per-row entry lists.  mode 0 = INSERT_VALUES (later entry at the same (i,j)
 * overwrites), 1 = ADD_VALUES (entries at the same (i,j) are summed in insertion order).  Negative
 * columns are ignored, as PETSc's MatSetValue does (FspMatrixBase.cpp:189 relies on this).  Rows are
 * stored with ascending column index like PETSc AIJ/SELL. */
static void orc_csr_assemble(orc_csr *M, int nrows, int ncols, const int *row_cnt, orc_ent **rows, int mode) {
  M->nrows = nrows;
  M->ncols = ncols;
  M->ptr   = (int *) calloc((size_t) nrows + 1, sizeof(int));
  size_t tot = 0;
  for (int i = 0; i < nrows; ++i) tot += row_cnt[i];
  M->col = (int *) malloc(sizeof(int) * (tot + 1));
  M->val = (double *) malloc(sizeof(double) * (tot + 1));
  int nz = 0;
  for (int i = 0; i < nrows; ++i) {
    M->ptr[i] = nz;
    int      start = nz;
    orc_ent *e     = rows[i];
    for (int q = 0; q < row_cnt[i]; ++q) {
      if (e[q].col < 0) continue;
      int found = -1;
      for (int p = start; p < nz; ++p)
        if (M->col[p] == e[q].col) { found = p; break; }
      if (found >= 0) {
        if (mode) M->val[found] += e[q].val; else M->val[found] = e[q].val;
      } else {
        M->col[nz] = e[q].col;
        M->val[nz] = e[q].val;
        nz++;
      }
    }
    /* sort the row by column (stable insertion sort keeps it cheap: rows are short) */
    int len = nz - start;
    if (len > 1) {
      orc_ent  small[64];
      orc_ent *tmp = len <= 64 ? small : (orc_ent *) malloc(sizeof(orc_ent) * len);
      for (int p = 0; p < len; ++p) { tmp[p].col = M->col[start + p]; tmp[p].val = M->val[start + p]; }
      qsort(tmp, len, sizeof(orc_ent), orc_ent_cmp);
      for (int p = 0; p < len; ++p) { M->col[start + p] = tmp[p].col; M->val[start + p] = tmp[p].val; }
      if (tmp != small) free(tmp);
    }
  }
  M->ptr[nrows] = nz;
}

/* GenerateValues.  enable == NULL / n_en == 0 => all reactions (FspMatrixBase.cpp:105-111). */
ORC_API int orc_mat_generate(orc_mat *A, const orc_set *st, int n_tv_in, const int *time_varying,
                             fsp_tcoef_fn prop_t, fsp_prop_fn prop_x, int n_en, const int *enable,
                             void *prop_t_args, void *prop_x_args) {
  const int S = st->S, R = st->R, n = st->n;
  const int K = A->constrained ? st->K : 0;
  orc_mat_destroy_values(A);
  A->n = n; A->R = R; A->K = K;
  A->nrows = n + K; /* DetermineLayout_: FspMatrixBase.cpp:277-300 / FspMatrixConstrained.cpp:284-302 */
  A->t_fun = prop_t; A->t_fun_args = prop_t_args;
  A->coef = (double *) calloc((size_t) R, sizeof(double));
  A->work = (double *) calloc((size_t) A->nrows + 1, sizeof(double));

  A->enabled = (int *) malloc(sizeof(int) * R);
  A->tv      = (int *) malloc(sizeof(int) * R);
  A->ti      = (int *) malloc(sizeof(int) * R);
  if (n_en <= 0 || !enable) {
    A->n_en = R;
    for (int r = 0; r < R; ++r) A->enabled[r] = r;
  } else {
    A->n_en = n_en;
    memcpy(A->enabled, enable, sizeof(int) * n_en);
  }
  for (int e = 0; e < A->n_en; ++e) { /* :112-118 */
    int r = A->enabled[e], is_tv = 0;
    for (int q = 0; q < n_tv_in; ++q)
      if (time_varying[q] == r) is_tv = 1;
    if (is_tv) A->tv[A->n_tv++] = r; else A->ti[A->n_ti++] = r;
  }

  A->col  = (int *) malloc(sizeof(int) * (size_t) R * n + 4);
  A->off  = (double *) calloc((size_t) R * n + 1, sizeof(double));
  A->diag = (double *) calloc((size_t) R * n + 1, sizeof(double));
  for (size_t q = 0; q < (size_t) R * n; ++q) A->col[q] = -1;

  int *shifted = (int *) malloc(sizeof(int) * (size_t) n * S + 4);
  int  ierr    = 0;
  for (int e = 0; e < A->n_en && !ierr; ++e) { /* :132-145 */
    int r = A->enabled[e];
    for (int i = 0; i < n; ++i)
      for (int s = 0; s < S; ++s) shifted[(size_t) i * S + s] = st->states[(size_t) i * S + s] - st->SM[r * S + s];
    orc_set_state2index(st, n, shifted, A->col + (size_t) r * n);
    ierr = prop_x(r, S, n, shifted, A->off + (size_t) r * n, prop_x_args);
    if (ierr) break;
    ierr = prop_x(r, S, n, st->states, A->diag + (size_t) r * n, prop_x_args); /* :180, :232 */
  }
  if (ierr) { free(shifted); return ierr; }

  /* TV matrices: INSERT diag then off-diagonal (:181-191) */
  int       *cnt  = (int *) malloc(sizeof(int) * (size_t) A->nrows + 4);
  orc_ent  **rows = (orc_ent **) malloc(sizeof(orc_ent *) * ((size_t) A->nrows + 1));
  A->tv_mats      = (orc_csr *) calloc((size_t) (A->n_tv ? A->n_tv : 1), sizeof(orc_csr));
  orc_ent *pool   = (orc_ent *) malloc(sizeof(orc_ent) * ((size_t) n * (R + 1) + 1));
  for (int q = 0; q < A->n_tv; ++q) {
    int r = A->tv[q];
    for (int i = 0; i < A->nrows; ++i) { cnt[i] = 0; rows[i] = pool + (size_t) 2 * (i < n ? i : n); }
    for (int i = 0; i < n; ++i) {
      rows[i][0].col = i;                         rows[i][0].val = -1.0 * A->diag[(size_t) r * n + i];
      rows[i][1].col = A->col[(size_t) r * n + i]; rows[i][1].val = A->off[(size_t) r * n + i];
      cnt[i]         = 2;
    }
    orc_csr_assemble(&A->tv_mats[q], A->nrows, A->nrows, cnt, rows, 0);
  }
  /* merged TI matrix: ADD_VALUES, reaction by reaction (:229-243) */
  if (A->n_ti > 0) {
    /* rows >= n (sinks) have no entries; rows < n get 2 entries per TI reaction */
    free(pool);
    pool = (orc_ent *) malloc(sizeof(orc_ent) * ((size_t) n * 2 * A->n_ti + 1));
    for (int i = 0; i < A->nrows; ++i) { cnt[i] = 0; rows[i] = pool + (size_t) 2 * A->n_ti * (i < n ? i : 0); }
    for (int q = 0; q < A->n_ti; ++q) {
      int r = A->ti[q];
      for (int i = 0; i < n; ++i) {
        rows[i][cnt[i]].col = i;                          rows[i][cnt[i]++].val = -1.0 * A->diag[(size_t) r * n + i];
        rows[i][cnt[i]].col = A->col[(size_t) r * n + i]; rows[i][cnt[i]++].val = A->off[(size_t) r * n + i];
      }
    }
    orc_csr_assemble(&A->ti_mat, A->nrows, A->nrows, cnt, rows, 1);
    A->has_ti = 1;
  }
  free(pool); free(cnt); free(rows);

  /* sink rows (FspMatrixConstrained.cpp:170-194) */
  if (A->constrained) {
    A->sink_nnz = (int *) calloc((size_t) R * K + 1, sizeof(int));
    A->sink_inz = (int **) calloc((size_t) R * K + 1, sizeof(int *));
    A->sink_val = (double **) calloc((size_t) R * K + 1, sizeof(double *));
    int *sat    = (int *) malloc(sizeof(int) * ((size_t) n * K + 4));
    for (int e = 0; e < A->n_en && !ierr; ++e) {
      int r = A->enabled[e];
      for (int i = 0; i < n; ++i)
        for (int s = 0; s < S; ++s) shifted[(size_t) i * S + s] = st->states[(size_t) i * S + s] + st->SM[r * S + s];
      ierr = orc_set_check_constraints(st, n, shifted, sat);
      if (ierr) break;
      for (int k = 0; k < K; ++k) {
        int c = 0;
        for (int i = 0; i < n; ++i) c += (sat[(size_t) n * k + i] == 0);
        A->sink_nnz[r * K + k] = c;
        A->sink_inz[r * K + k] = (int *) malloc(sizeof(int) * (c + 1));
        A->sink_val[r * K + k] = (double *) malloc(sizeof(double) * (c + 1));
        c = 0;
        for (int i = 0; i < n; ++i)
          if (sat[(size_t) n * k + i] == 0) {
            A->sink_inz[r * K + k][c] = i;
            /* one callback per boundary entry (:188) */
            ierr = prop_x(r, S, 1, st->states + (size_t) i * S, &A->sink_val[r * K + k][c], prop_x_args);
            c++;
          }
      }
    }
    free(sat);
    if (ierr) { free(shifted); return ierr; }
    /* assemble K x nrows sink matrices with ADD_VALUES (:199-240) */
    A->tv_sinks = (orc_csr *) calloc((size_t) (A->n_tv ? A->n_tv : 1), sizeof(orc_csr));
    int       *kc = (int *) malloc(sizeof(int) * (K + 1));
    orc_ent  **kr = (orc_ent **) malloc(sizeof(orc_ent *) * (K + 1));
    for (int q = 0; q < A->n_tv; ++q) {
      int r = A->tv[q];
      for (int k = 0; k < K; ++k) {
        kc[k] = A->sink_nnz[r * K + k];
        kr[k] = (orc_ent *) malloc(sizeof(orc_ent) * (kc[k] + 1));
        for (int c = 0; c < kc[k]; ++c) { kr[k][c].col = A->sink_inz[r * K + k][c]; kr[k][c].val = A->sink_val[r * K + k][c]; }
      }
      orc_csr_assemble(&A->tv_sinks[q], K, A->nrows, kc, kr, 1);
      for (int k = 0; k < K; ++k) free(kr[k]);
    }
    if (A->n_ti > 0) {
      for (int k = 0; k < K; ++k) {
        int tot = 0;
        for (int q = 0; q < A->n_ti; ++q) tot += A->sink_nnz[A->ti[q] * K + k];
        kr[k] = (orc_ent *) malloc(sizeof(orc_ent) * (tot + 1));
        kc[k] = 0;
        for (int q = 0; q < A->n_ti; ++q) {
          int r = A->ti[q];
          for (int c = 0; c < A->sink_nnz[r * K + k]; ++c) {
            kr[k][kc[k]].col = A->sink_inz[r * K + k][c];
            kr[k][kc[k]++].val = A->sink_val[r * K + k][c];
          }
        }
      }
      orc_csr_assemble(&A->ti_sinks, K, A->nrows, kc, kr, 1);
      for (int k = 0; k < K; ++k) free(kr[k]);
    }
    free(kc); free(kr);
  }
  free(shifted);
  A->has_values = 1;
  return 0;
}

/* Direct generator for the synthetic 3-D birth-death lattice (BASELINE config 4) in the canonical
 * lexicographic ordering idx = x0 + L0 (x1 + L1 x2) (sub2ind_nd convention, src/Sys/pacmenMath.h:33-59): builds the
 * same reference-shaped operators as orc_mat_generate on that state set -- one CSR per TV reaction (diagonal +
 * off-diagonal, INSERT) and one merged TI CSR (ADD_VALUES, ascending columns), plus the K = 3 sink matrices --
 * without the BFS, the hash directory or the R x n staging arrays, so that the full 465^3 problem (1e8 states) can be
 * set up in seconds for the CPU-baseline timing.  tests/test_oracle_kats.py checks it against orc_mat_generate
 * (by state key) on small boxes.  Values come from the same fixture callback (bd3_prop). */
static void orc_lattice_rows(orc_csr *M, int n, int nrows, int L0, int L1, int L2, int n_re, const int *reactions) {
  /* two passes: count, then fill; rows ascending in column; reactions listed in `reactions` are merged (ADD) */
  M->nrows = nrows; M->ncols = nrows;
  M->ptr = (int *) calloc((size_t) nrows + 1, sizeof(int));
  const int  Ls[3] = {L0, L1, L2};
  const long stride[3] = {1, L0, (long) L0 * L1};
  int *cnt = (int *) calloc((size_t) nrows + 1, sizeof(int));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    int x[3] = {i % L0, (i / L0) % L1, i / (L0 * L1)};
    int c = 1; /* diagonal */
    for (int q = 0; q < n_re; ++q) {
      int r = reactions[q], s = r / 2, src = x[s] - ((r & 1) ? -1 : 1);
      if (src >= 0 && src < Ls[s]) c++;
    }
    cnt[i] = c;
  }
  long nz = 0;
  for (int i = 0; i < nrows; ++i) { M->ptr[i] = (int) nz; nz += cnt[i]; }
  M->ptr[nrows] = (int) nz;
  free(cnt);
  M->col = (int *) malloc(sizeof(int) * ((size_t) nz + 1));
  M->val = (double *) malloc(sizeof(double) * ((size_t) nz + 1));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    int     x[3] = {i % L0, (i / L0) % L1, i / (L0 * L1)};
    orc_ent e[16];
    int     m = 0;
    double  dsum = 0.0;
    for (int q = 0; q < n_re; ++q) { /* diagonal: -d_r(x_i), summed in the order of `reactions` (ADD_VALUES) */
      int    r = reactions[q];
      double d = 0.0;
      bd3_prop(r, 3, 1, x, &d, NULL);
      dsum += -1.0 * d;
    }
    e[m].col = i; e[m].val = dsum; m++;
    for (int q = 0; q < n_re; ++q) { /* off-diagonal: d_r(x_i - nu_r) at the column of x_i - nu_r */
      int r = reactions[q], s = r / 2, dir = (r & 1) ? -1 : 1;
      int y[3] = {x[0], x[1], x[2]};
      y[s] -= dir;
      if (y[s] < 0 || y[s] >= Ls[s]) continue;
      double d = 0.0;
      bd3_prop(r, 3, 1, y, &d, NULL);
      e[m].col = (int) (i - dir * stride[s]); e[m].val = d; m++;
    }
    for (int a = 1; a < m; ++a) { /* ascending columns (distinct by construction) */
      orc_ent t = e[a];
      int     b = a - 1;
      while (b >= 0 && e[b].col > t.col) { e[b + 1] = e[b]; b--; }
      e[b + 1] = t;
    }
    int p0 = M->ptr[i];
    for (int a = 0; a < m; ++a) { M->col[p0 + a] = e[a].col; M->val[p0 + a] = e[a].val; }
  }
}
static void orc_lattice_sinks(orc_csr *M, int n, int L0, int L1, int L2, int n_re, const int *reactions) {
  /* K = 3 rows; row k gets (col i, d_r(x_i)) for every state with x_k == L_k - 1 and every listed birth of species k
   * (FspMatrixConstrained.cpp:170-194: x_i + nu_r violates constraint k; a box face violates exactly one) */
  const int Ls[3] = {L0, L1, L2};
  M->nrows = 3; M->ncols = n + 3;
  M->ptr = (int *) calloc(4, sizeof(int));
  long tot = 0;
  int  has[3] = {0, 0, 0};
  for (int q = 0; q < n_re; ++q) if ((reactions[q] & 1) == 0) has[reactions[q] / 2] = 1;
  for (int k = 0; k < 3; ++k) { M->ptr[k] = (int) tot; if (has[k]) tot += (long) n / Ls[k]; }
  M->ptr[3] = (int) tot;
  M->col = (int *) malloc(sizeof(int) * ((size_t) tot + 1));
  M->val = (double *) malloc(sizeof(double) * ((size_t) tot + 1));
  for (int k = 0; k < 3; ++k) {
    if (!has[k]) continue;
    long p = M->ptr[k];
    for (int i = 0; i < n; ++i) {
      int x[3] = {i % L0, (i / L0) % L1, i / (L0 * L1)};
      if (x[k] != Ls[k] - 1) continue;
      double d = 0.0;
      bd3_prop(2 * k, 3, 1, x, &d, NULL);
      M->col[p] = i; M->val[p] = d; p++;
    }
  }
}
ORC_API int orc_mat_generate_lattice(orc_mat *A, int L0, int L1, int L2, int tv) {
  const long nl = (long) L0 * L1 * L2;
  if (L0 < 1 || L1 < 1 || L2 < 1 || nl * 7 > 2147483000L) return -1;
  const int n = (int) nl, R = 6, K = A->constrained ? 3 : 0;
  orc_mat_destroy_values(A);
  A->n = n; A->R = R; A->K = K; A->nrows = n + K;
  A->t_fun = bd3_tfun; A->t_fun_args = NULL;
  A->coef = (double *) calloc((size_t) R, sizeof(double));
  A->work = (double *) calloc((size_t) A->nrows + 1, sizeof(double));
  A->enabled = (int *) malloc(sizeof(int) * R);
  A->tv = (int *) malloc(sizeof(int) * R);
  A->ti = (int *) malloc(sizeof(int) * R);
  A->n_en = R;
  for (int r = 0; r < R; ++r) {
    A->enabled[r] = r;
    if (tv && (r & 1) == 0) A->tv[A->n_tv++] = r; else A->ti[A->n_ti++] = r;
  }
  A->tv_mats = (orc_csr *) calloc((size_t) (A->n_tv ? A->n_tv : 1), sizeof(orc_csr));
  for (int q = 0; q < A->n_tv; ++q) orc_lattice_rows(&A->tv_mats[q], n, A->nrows, L0, L1, L2, 1, &A->tv[q]);
  if (A->n_ti > 0) { orc_lattice_rows(&A->ti_mat, n, A->nrows, L0, L1, L2, A->n_ti, A->ti); A->has_ti = 1; }
  if (A->constrained) {
    A->tv_sinks = (orc_csr *) calloc((size_t) (A->n_tv ? A->n_tv : 1), sizeof(orc_csr));
    for (int q = 0; q < A->n_tv; ++q) orc_lattice_sinks(&A->tv_sinks[q], n, L0, L1, L2, 1, &A->tv[q]);
    if (A->n_ti > 0) orc_lattice_sinks(&A->ti_sinks, n, L0, L1, L2, A->n_ti, A->ti);
  }
  A->has_values = 1;
  return 0;
}
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void) n;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Action (reference-shaped, multi-pass).  FspMatrixBase.cpp:36-62 + FspMatrixConstrained.cpp:31-64
 * ---------------------------------------------------------------------------------------------- */
static void orc_spmv(const orc_csr *M, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M->nrows; ++i) {
    double acc = 0.0;
    for (int p = M->ptr[i]; p < M->ptr[i + 1]; ++p) acc += M->val[p] * x[M->col[p]];
    y[i] = acc;
  }
}
static void orc_axpy(int n, double a, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) y[i] += a * x[i];
}

/* Action with the coefficients supplied directly (used by sensitivity operators as well). */
ORC_API int orc_mat_action_coef(orc_mat *A, const double *coef, const double *x, double *y) {
  const int nr = A->nrows;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nr; ++i) y[i] = 0.0; /* VecSet(y, 0): :39 */
  if (!A->has_values) return 0;            /* :41 */
  for (int q = 0; q < A->n_tv; ++q) {      /* :47-52 */
    orc_spmv(&A->tv_mats[q], x, A->work);
    orc_axpy(nr, coef[A->tv[q]], A->work, y);
  }
  if (A->has_ti) { /* :55-60 */
    orc_spmv(&A->ti_mat, x, A->work);
    orc_axpy(nr, 1.0, A->work, y);
  }
  if (A->constrained) { /* FspMatrixConstrained.cpp:38-60 */
    double sink[FSP_FIXTURE_MAX_CONSTR * 4], tmp[FSP_FIXTURE_MAX_CONSTR * 4];
    for (int k = 0; k < A->K; ++k) sink[k] = 0.0;
    for (int q = 0; q < A->n_tv; ++q) {
      orc_spmv(&A->tv_sinks[q], x, tmp);
      for (int k = 0; k < A->K; ++k) sink[k] += coef[A->tv[q]] * tmp[k];
    }
    if (A->has_ti) {
      orc_spmv(&A->ti_sinks, x, tmp);
      for (int k = 0; k < A->K; ++k) sink[k] += 1.0 * tmp[k];
    }
    for (int k = 0; k < A->K; ++k) y[A->n + k] += sink[k];
  }
  return 0;
}

ORC_API int orc_mat_action(orc_mat *A, double t, const double *x, double *y) {
  if (A->has_values && A->n_tv > 0) { /* :43-45 */
    int ierr = A->t_fun(t, A->R, A->coef, A->t_fun_args);
    if (ierr) {
      for (int i = 0; i < A->nrows; ++i) y[i] = 0.0;
      return ierr;
    }
  }
  return orc_mat_action_coef(A, A->coef, x, y);
}

/* A stronger CPU baseline: single fused pass over the per-reaction arrays (not reference-shaped). */
ORC_API int orc_mat_action_fused(orc_mat *A, double t, const double *x, double *y) {
  if (!A->has_values) { for (int i = 0; i < A->nrows; ++i) y[i] = 0.0; return 0; }
  if (A->n_tv > 0) { int ierr = A->t_fun(t, A->R, A->coef, A->t_fun_args); if (ierr) return ierr; }
  const int n = A->n;
  double c[FSP_FIXTURE_MAX_REACTIONS * 4];
  for (int r = 0; r < A->R; ++r) c[r] = 0.0;
  for (int q = 0; q < A->n_tv; ++q) c[A->tv[q]] = A->coef[A->tv[q]];
  for (int q = 0; q < A->n_ti; ++q) c[A->ti[q]] = 1.0;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    double acc = 0.0, xi = x[i];
    for (int e = 0; e < A->n_en; ++e) {
      int    r  = A->enabled[e];
      int    j  = A->col[(size_t) r * n + i];
      double xs = j >= 0 ? x[j] : 0.0;
      acc += c[r] * (A->off[(size_t) r * n + i] * xs - A->diag[(size_t) r * n + i] * xi);
    }
    y[i] = acc;
  }
  for (int k = 0; k < A->K; ++k) {
    double s = 0.0;
    for (int e = 0; e < A->n_en; ++e) {
      int r = A->enabled[e];
      double p = 0.0;
      for (int q = 0; q < A->sink_nnz[r * A->K + k]; ++q) p += A->sink_val[r * A->K + k][q] * x[A->sink_inz[r * A->K + k][q]];
      s += c[r] * p;
    }
    y[n + k] = s;
  }
  return 0;
}

/* GetLocalMVFlops: FspMatrixBase.cpp:429-444 + FspMatrixConstrained.cpp:447-465 */
ORC_API int orc_mat_flops(const orc_mat *A) {
  long f = 0;
  if (A->has_ti) f += 2L * A->ti_mat.ptr[A->ti_mat.nrows];
  for (int q = 0; q < A->n_tv; ++q) f += 2L * A->tv_mats[q].ptr[A->tv_mats[q].nrows] + A->nrows;
  if (A->constrained) {
    if (A->has_ti) f += 2L * A->ti_sinks.ptr[A->ti_sinks.nrows];
    for (int q = 0; q < A->n_tv; ++q) f += 2L * A->tv_sinks[q].ptr[A->tv_sinks[q].nrows] + A->K;
  }
  return (int) f;
}

ORC_API int orc_mat_num_rows(const orc_mat *A) { return A->nrows; }
ORC_API int orc_mat_num_states(const orc_mat *A) { return A->n; }
ORC_API const int *orc_mat_col(const orc_mat *A) { return A->col; }
ORC_API const double *orc_mat_off(const orc_mat *A) { return A->off; }
ORC_API const double *orc_mat_diag(const orc_mat *A) { return A->diag; }
ORC_API int orc_mat_num_tv(const orc_mat *A) { return A->n_tv; }
ORC_API int orc_mat_num_ti(const orc_mat *A) { return A->n_ti; }
ORC_API const int *orc_mat_tv(const orc_mat *A) { return A->tv; }
ORC_API const int *orc_mat_ti(const orc_mat *A) { return A->ti; }
ORC_API int orc_mat_sink_nnz(const orc_mat *A, int r, int k) { return A->sink_nnz ? A->sink_nnz[r * A->K + k] : 0; }
ORC_API const int *orc_mat_sink_inz(const orc_mat *A, int r, int k) { return A->sink_inz[r * A->K + k]; }
ORC_API const double *orc_mat_sink_val(const orc_mat *A, int r, int k) { return A->sink_val[r * A->K + k]; }

/* Dense A(t) (column major, nrows x nrows) for small problems: the analogue of
 * CreateRHSJacobian + ComputeRHSJacobian (FspMatrixBase.cpp:308-427, FspMatrixConstrained.cpp:304-445). */
ORC_API int orc_mat_dense(orc_mat *A, double t, double *out) {
  const int nr = A->nrows, n = A->n;
  memset(out, 0, sizeof(double) * (size_t) nr * nr);
  if (!A->has_values) return 0;
  if (A->n_tv > 0) { int ierr = A->t_fun(t, A->R, A->coef, A->t_fun_args); if (ierr) return ierr; }
  for (int e = 0; e < A->n_en; ++e) {
    int r = A->enabled[e], is_tv = 0;
    for (int q = 0; q < A->n_tv; ++q) if (A->tv[q] == r) is_tv = 1;
    double c = is_tv ? A->coef[r] : 1.0;
    for (int i = 0; i < n; ++i) {
      int j = A->col[(size_t) r * n + i];
      if (j >= 0) out[(size_t) j * nr + i] += c * A->off[(size_t) r * n + i];
      out[(size_t) i * nr + i] += -1.0 * c * A->diag[(size_t) r * n + i];
    }
    for (int k = 0; k < A->K; ++k)
      for (int q = 0; q < A->sink_nnz[r * A->K + k]; ++q)
        out[(size_t) A->sink_inz[r * A->K + k][q] * nr + (n + k)] += c * A->sink_val[r * A->K + k][q];
  }
  return 0;
}

/* ExpandVec: PetscWrap/PetscWrap.cpp:26-56 -- p_new = 0; p_new[new_idx[i]] = p_old[i]. */
ORC_API void orc_expand_vec(int n_old, const double *p_old, const int *new_idx, int n_new, double *p_new) {
  for (int i = 0; i < n_new; ++i) p_new[i] = 0.0;
  for (int i = 0; i < n_old; ++i) p_new[new_idx[i]] = p_old[i];
}

/* ------------------------------------------------------------------------------------------------
 * Fixture access (so Python can build oracle objects for the named workloads without ctypes callbacks)
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_fixture_get(const char *name, fsp_fixture *out) { return fsp_fixture_get(name, out); }
ORC_API int orc_fixture_sizeof(void) { return (int) sizeof(fsp_fixture); }

ORC_API orc_set *orc_set_from_fixture(const char *name, const int *bounds_override) {
  fsp_fixture f;
  if (fsp_fixture_get(name, &f)) return NULL;
  orc_set *st = orc_set_create(f.num_species, f.num_reactions, f.SM);
  orc_set_set_shape(st, f.num_constr, f.lhs, bounds_override ? bounds_override : f.bounds, NULL);
  orc_set_add_states(st, f.num_species, 1, f.x0);
  return st;
}

ORC_API int orc_mat_generate_fixture(orc_mat *A, const orc_set *st, const char *name) {
  fsp_fixture f;
  if (fsp_fixture_get(name, &f)) return -1;
  return orc_mat_generate(A, st, f.num_tv, f.tv_reactions, f.prop_t, f.prop_x, 0, NULL, NULL, NULL);
}

/* parameters of a named workload (inputs, not algorithm): S, R, K, bounds[K], expansion[K], x0[S], t_final, fsp_tol, rtol, atol */
ORC_API int orc_fixture_info(const char *name, int *dims, int *bounds, double *expansion, int *x0, double *params) {
  fsp_fixture f;
  if (fsp_fixture_get(name, &f)) return -1;
  dims[0] = f.num_species; dims[1] = f.num_reactions; dims[2] = f.num_constr; dims[3] = f.num_tv;
  for (int k = 0; k < f.num_constr; ++k) { bounds[k] = f.bounds[k]; expansion[k] = f.expansion[k]; }
  for (int s = 0; s < f.num_species; ++s) x0[s] = f.x0[s];
  params[0] = f.t_final; params[1] = f.fsp_tol; params[2] = f.rtol; params[3] = f.atol;
  return 0;
}

ORC_API int orc_fixture_tcoef(const char *name, double t, double *out) {
  fsp_fixture f;
  if (fsp_fixture_get(name, &f)) return -1;
  return f.prop_t(t, f.num_reactions, out, NULL);
}

/* ------------------------------------------------------------------------------------------------
 * BLAS-1 kernels of the CPU restatement of KrylovFsp (oracle/krylov_oracle.py): the PETSc calls of
 * src/OdeSolver/KrylovFsp.cpp:280-309 (VecDot, VecAXPY, VecNorm, VecScale) and :244-252 (VecMAXPY),
 * OpenMP over the host cores.
 * ---------------------------------------------------------------------------------------------- */
ORC_API double orc_vec_dot(long n, const double *x, const double *y) {
  double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
  for (long i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}
ORC_API void orc_vec_axpy(long n, double a, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) y[i] += a * x[i];
}
ORC_API void orc_vec_scale(long n, double a, double *y) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) y[i] *= a;
}
ORC_API void orc_vec_copy(long n, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) y[i] = x[i];
}
/* y = sum_k c[k] X[k]  (X: m rows of leading dimension ld) */
ORC_API void orc_vec_maxpy(long n, int m, const double *c, const double *X, long ld, double *y) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    double acc = 0.0;
    for (int k = 0; k < m; ++k) acc += c[k] * X[(size_t) k * ld + i];
    y[i] = acc;
  }
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

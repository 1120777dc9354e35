/*
 * fsp_models.h -- workload definitions (reaction networks) shared by the oracle, the host tests,
 * the examples and bench.py.  These are INPUTS (stoichiometry, propensity factors, time
 * coefficients, shape constraints), not part of the algorithm under test.
 *
 * Callback contracts follow the reference (src/Models/Model.h:44-60,
 * src/StateSet/StateSetConstrained.h:32-33):
 *   prop_x(reaction, S, m, const int* states /+ S x m, column major +/, double* out, void* args) -> int
 *   prop_t(t, R, double* out, void* args) -> int        (fills the R coefficients c_r(t))
 *   lhs(S, K, m, int* states, int* out /+ out[K*j + k] +/, void* args) -> int
 *
 * Parameter values restate the reference fixtures:
 *   random walk     tests/test_mat.cpp:45-69,104-105
 *   toggle          tests/test_ode.cpp:33-89, tests/test_fsp_solver.cpp:31-95
 *   pure birth      tests/test_fsp_solver.cpp:179-215
 *   repressilator   examples/repressilator.cpp:14-84
 *   hog1p (5-d)     examples/hog1p.cpp:14-112
 *   transcr_reg_6d  examples/transcr_reg_6d.cpp:14-85
 *   birth-death 3-d SURVEY.md section 8(d) (synthetic lattice, config 4)
 *
 * Plain C99 so that it can be included from C (oracle), C++ (host) and CUDA.
 */
#ifndef FSP_MODELS_H_
#define FSP_MODELS_H_

#include <math.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int (*fsp_prop_fn)(int reaction, int num_species, int num_states, const int *states, double *out,
                           void *args);
typedef int (*fsp_tcoef_fn)(double t, int num_coefs, double *out, void *args);
typedef int (*fsp_constr_fn)(int num_species, int num_constr, int num_states, int *states, int *out,
                             void *args);

#define FSP_FIXTURE_MAX_SPECIES 8
#define FSP_FIXTURE_MAX_REACTIONS 16
#define FSP_FIXTURE_MAX_CONSTR 8

typedef struct fsp_fixture {
  const char  *name;
  int          num_species;
  int          num_reactions;
  int          SM[FSP_FIXTURE_MAX_SPECIES * FSP_FIXTURE_MAX_REACTIONS]; /* column major: SM[r*S + s] */
  fsp_prop_fn  prop_x;
  fsp_tcoef_fn prop_t;
  int          num_tv;
  int          tv_reactions[FSP_FIXTURE_MAX_REACTIONS];
  int          num_constr;
  fsp_constr_fn lhs;                          /* NULL => default identity constraints (K == S) */
  int          bounds[FSP_FIXTURE_MAX_CONSTR];
  double       expansion[FSP_FIXTURE_MAX_CONSTR];
  int          x0[FSP_FIXTURE_MAX_SPECIES];
  double       t_final;
  double       fsp_tol;
  double       rtol, atol;
} fsp_fixture;

/* ---------------- 1-d random walk (KAT-M1..M5) ---------------- */
static int rw1d_prop(int r, int S, int m, const int *X, double *out, void *a) {
  (void) S; (void) a;
  switch (r) {
    case 0: for (int i = 0; i < m; ++i) out[i] = 2.0; break;
    case 1: for (int i = 0; i < m; ++i) out[i] = 3.0 * (X[i] > 0); break;
    default: return -1;
  }
  return 0;
}
static int rw1d_tfun(double t, int n, double *out, void *a) {
  (void) n; (void) a;
  out[0] = 1.0 + t;
  out[1] = 1.0 + 0.5 * t;
  return 0;
}

/* ---------------- toggle switch (KAT-O*, KAT-F1/2) ---------------- */
static const double tg_ayx = 2.6e-3, tg_axy = 6.1e-3, tg_nyx = 3.0, tg_nxy = 2.1, tg_kx0 = 2.2e-3,
                    tg_kx = 1.7e-2, tg_dx = 3.8e-4, tg_ky0 = 6.8e-5, tg_ky = 1.6e-2, tg_dy = 3.8e-4;
static int toggle_prop(int r, int S, int m, const int *X, double *out, void *a) {
  (void) S; (void) a;
  switch (r) {
    case 0: for (int i = 0; i < m; ++i) out[i] = tg_kx0; break;
    case 1: for (int i = 0; i < m; ++i) out[i] = tg_kx / (1.0 + tg_ayx * pow((double) X[2 * i + 1], tg_nyx)); break;
    case 2: for (int i = 0; i < m; ++i) out[i] = tg_dx * (double) X[2 * i]; break;
    case 3: for (int i = 0; i < m; ++i) out[i] = tg_ky0; break;
    case 4: for (int i = 0; i < m; ++i) out[i] = tg_ky / (1.0 + tg_axy * pow((double) X[2 * i], tg_nxy)); break;
    case 5: for (int i = 0; i < m; ++i) out[i] = tg_dy * (double) X[2 * i + 1]; break;
    default: return -1;
  }
  return 0;
}
static int toggle_tfun(double t, int n, double *out, void *a) {
  (void) t; (void) n; (void) out; (void) a;
  return 0;
}
static int toggle_lhs(int S, int K, int m, int *X, int *out, void *a) {
  (void) a;
  if (S != 2 || K != 3) return -1;
  for (int i = 0; i < m; ++i) {
    out[i * K]     = X[S * i];
    out[i * K + 1] = X[S * i + 1];
    out[i * K + 2] = X[S * i] * X[S * i + 1];
  }
  return 0;
}

/* ---------------- pure birth (Poisson; KAT-F3..5) ---------------- */
static int birth_prop(int r, int S, int m, const int *X, double *out, void *a) {
  (void) r; (void) S; (void) X; (void) a;
  for (int i = 0; i < m; ++i) out[i] = 2.0;
  return 0;
}
static int ones_tfun(double t, int n, double *out, void *a) {
  (void) t; (void) a;
  for (int i = 0; i < n; ++i) out[i] = 1.0;
  return 0;
}

/* ---------------- repressilator (config 1) ---------------- */
static int repress_prop(int r, int S, int m, const int *X, double *out, void *a) {
  (void) S; (void) a;
  const double k1 = 100.0, ka = 20.0, ket = 6.0, kg = 1.0;
  for (int i = 0; i < m; ++i) {
    const int *x = X + 3 * i;
    double     v;
    switch (r) {
      case 0: v = k1 / (1.0 + ka * pow(1.0 * (double) x[1], ket)); break;
      case 1: v = kg * (double) x[0]; break;
      case 2: v = k1 / (1.0 + ka * pow(1.0 * (double) x[2], ket)); break;
      case 3: v = kg * (double) x[1]; break;
      case 4: v = k1 / (1.0 + ka * pow(1.0 * (double) x[0], ket)); break;
      case 5: v = kg * (double) x[2]; break;
      default: v = 0.0;
    }
    out[i] = v;
  }
  return 0;
}
static int repress_lhs(int S, int K, int m, int *X, int *out, void *a) {
  (void) a;
  if (S != 3 || K != 6) return -1;
  for (int i = 0; i < m; ++i) {
    const int *x = X + S * i;
    out[i * K]     = x[0];
    out[i * K + 1] = x[1];
    out[i * K + 2] = x[2];
    out[i * K + 3] = x[0] * x[1];
    out[i * K + 4] = x[2] * x[1];
    out[i * K + 5] = x[0] * x[2];
  }
  return 0;
}

/* ---------------- hog1p 5-d (config 2) ---------------- */
static int hog1p_prop(int r, int S, int m, const int *X, double *out, void *a) {
  (void) S; (void) a;
  const double k12 = 1.29, k23 = 0.0067, k34 = 0.133, k32 = 0.027, k43 = 0.0381, k21 = 1.0, kr21 = 0.005,
               kr31 = 0.45, kr41 = 0.025, kr22 = 0.0116, kr32 = 0.987, kr42 = 0.0538, trans = 0.01,
               gamma1 = 0.001, gamma2 = 0.0049;
  for (int i = 0; i < m; ++i) {
    const int *x = X + 5 * i;
    double     v;
    switch (r) {
      case 0: v = k12 * (double) (x[0] == 0) + k23 * (double) (x[0] == 1) + k34 * (double) (x[0] == 2); break;
      case 1: v = k32 * (double) (x[0] == 2) + k43 * (double) (x[0] == 3); break;
      case 2: v = k21 * (double) (x[0] == 1); break;
      case 3: v = kr21 * (double) (x[0] == 1) + kr31 * (double) (x[0] == 2) + kr41 * (double) (x[0] == 3); break;
      case 4: v = kr22 * (double) (x[0] == 1) + kr32 * (double) (x[0] == 2) + kr42 * (double) (x[0] == 3); break;
      case 5: v = trans * (double) x[1]; break;
      case 6: v = trans * (double) x[2]; break;
      case 7: v = gamma1 * (double) x[3]; break;
      case 8: v = gamma2 * (double) x[4]; break;
      default: v = 0.0;
    }
    out[i] = v;
  }
  return 0;
}
static int hog1p_tfun(double t, int n, double *out, void *a) {
  (void) a;
  if (n != 9) return -1;
  const double r1 = 6.9e-5, r2 = 7.1e-3, eta = 3.1, Ahog = 9.3e09, Mhog = 6.4e-4;
  for (int i = 0; i < 9; ++i) out[i] = 1.0;
  double h1    = (1.0 - exp(-r1 * t)) * exp(-r2 * t);
  double hog1p = pow(h1 / (1.0 + h1 / Mhog), eta) * Ahog;
  double u     = 3200.0 - 7710.0 * hog1p;
  out[2]       = u > 0.0 ? u : 0.0;
  return 0;
}

/* ---------------- transcription regulation 6-d (config 3) ---------------- */
static int transcr_prop(int r, int S, int m, const int *X, double *out, void *a) {
  (void) S; (void) a;
  const double c0 = 0.043, c1 = 0.0007, c2 = 0.078, c3 = 0.0039, c5 = 0.4791, c7 = 0.8765e-11, c9 = 0.5;
  for (int i = 0; i < m; ++i) {
    const int *x = X + 6 * i;
    double     v;
    switch (r) {
      case 0: v = c0 * (double) x[5]; break;
      case 1: v = c1 * (double) x[0]; break;
      case 2: v = c2 * (double) x[3]; break;
      case 3: v = c3 * (double) x[5]; break;
      case 4: v = (double) x[1] * (double) x[2]; break;
      case 5: v = c5 * (double) x[3]; break;
      case 6: v = (double) x[3] * (double) x[1]; break;
      case 7: v = c7 * (double) x[4]; break;
      case 8: v = 0.5 * (double) x[0] * (double) (x[0] - 1); break;
      case 9: v = c9 * (double) x[1]; break;
      default: v = 0.0;
    }
    out[i] = v;
  }
  return 0;
}
static int transcr_tfun(double t, int n, double *out, void *a) {
  (void) a;
  if (n != 10) return -1;
  const double avg_cell_cyc_time = 35 * 60.0;
  for (int i = 0; i < 10; ++i) out[i] = 1.0;
  double AV = 6.022140857 * 1.0e8 * pow(2.0, t / avg_cell_cyc_time);
  out[4]    = 0.012e09 / AV;
  out[6]    = 0.00012e09 / AV;
  out[8]    = 0.05e09 / AV;
  return 0;
}

/* ---------------- synthetic 3-d birth-death lattice (config 4) ---------------- */
static int bd3_prop(int r, int S, int m, const int *X, double *out, void *a) {
  (void) S; (void) a;
  static const double birth[3] = {40.0, 30.0, 20.0};
  static const double death[3] = {1.0, 1.5, 2.0};
  if (r < 0 || r >= 6) return -1;
  int s = r / 2;
  if ((r & 1) == 0) {
    for (int i = 0; i < m; ++i) out[i] = birth[s];
  } else {
    for (int i = 0; i < m; ++i) out[i] = death[s] * (double) X[3 * i + s];
  }
  return 0;
}
static int bd3_tfun(double t, int n, double *out, void *a) {
  (void) a;
  for (int i = 0; i < n; ++i) out[i] = 1.0;
  double c = 1.0 + 0.5 * sin(0.1 * t);
  out[0] = c; out[2] = c; out[4] = c;
  return 0;
}

static int fsp_fixture_fill_SM(fsp_fixture *f, const int *rowmajor) {
  /* rowmajor: S rows of R entries, as written in the reference sources */
  for (int s = 0; s < f->num_species; ++s)
    for (int r = 0; r < f->num_reactions; ++r) f->SM[r * f->num_species + s] = rowmajor[s * f->num_reactions + r];
  return 0;
}

/* Fill *f with the named workload. Returns 0 on success, -1 if the name is unknown. */
static int fsp_fixture_get(const char *name, fsp_fixture *f) {
  memset(f, 0, sizeof(*f));
  f->name = name;
  f->rtol = 1.0e-6;
  f->atol = 1.0e-14;
  if (!strcmp(name, "random_walk_1d") || !strcmp(name, "random_walk_1d_tv")) {
    static const int sm[] = {1, -1};
    f->num_species = 1; f->num_reactions = 2;
    fsp_fixture_fill_SM(f, sm);
    f->prop_x = rw1d_prop; f->prop_t = rw1d_tfun;
    if (!strcmp(name, "random_walk_1d_tv")) { f->num_tv = 2; f->tv_reactions[0] = 0; f->tv_reactions[1] = 1; }
    f->num_constr = 1; f->bounds[0] = 12; f->expansion[0] = 0.2;
    f->t_final = 1.0; f->fsp_tol = 1.0e-6;
    return 0;
  }
  if (!strcmp(name, "toggle") || !strcmp(name, "toggle_custom")) {
    static const int sm[] = {1, 1, -1, 0, 0, 0, 0, 0, 0, 1, 1, -1};
    f->num_species = 2; f->num_reactions = 6;
    fsp_fixture_fill_SM(f, sm);
    f->prop_x = toggle_prop; f->prop_t = toggle_tfun;
    if (!strcmp(name, "toggle")) {
      f->num_constr = 2; f->bounds[0] = 100; f->bounds[1] = 100;
      f->expansion[0] = f->expansion[1] = 0.25;
    } else {
      f->num_constr = 3; f->lhs = toggle_lhs;
      f->bounds[0] = 200; f->bounds[1] = 200; f->bounds[2] = 2000;
      f->expansion[0] = f->expansion[1] = f->expansion[2] = 0.2;
    }
    f->t_final = 100.0; f->fsp_tol = 1.0e-6;
    return 0;
  }
  if (!strcmp(name, "pure_birth")) {
    static const int sm[] = {1};
    f->num_species = 1; f->num_reactions = 1;
    fsp_fixture_fill_SM(f, sm);
    f->prop_x = birth_prop; f->prop_t = ones_tfun;
    f->num_constr = 1; f->bounds[0] = 5; f->expansion[0] = 0.1;
    f->t_final = 10.0; f->fsp_tol = 1.0e-6;
    return 0;
  }
  if (!strcmp(name, "repressilator") || !strcmp(name, "repressilator_custom")) {
    static const int sm[] = {1, -1, 0, 0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 0, 0, 1, -1};
    f->num_species = 3; f->num_reactions = 6;
    fsp_fixture_fill_SM(f, sm);
    f->prop_x = repress_prop; f->prop_t = ones_tfun;
    if (!strcmp(name, "repressilator")) {
      f->num_constr = 3; f->bounds[0] = 22; f->bounds[1] = 2; f->bounds[2] = 2;
      for (int i = 0; i < 3; ++i) f->expansion[i] = 0.2;
    } else {
      static const int b[] = {22, 2, 2, 44, 4, 44};
      f->num_constr = 6; f->lhs = repress_lhs;
      for (int i = 0; i < 6; ++i) { f->bounds[i] = b[i]; f->expansion[i] = 0.2; }
    }
    f->x0[0] = 21;
    f->t_final = 10.0; f->fsp_tol = 1.0e-4; f->rtol = 1.0e-4; f->atol = 1.0e-14;
    return 0;
  }
  if (!strcmp(name, "hog1p")) {
    static const int sm[] = {1, -1, -1, 0, 0, 0, 0, 0, 0,
                             0, 0, 0, 1, 0, -1, 0, 0, 0,
                             0, 0, 0, 0, 1, 0, -1, 0, 0,
                             0, 0, 0, 0, 0, 1, 0, -1, 0,
                             0, 0, 0, 0, 0, 0, 1, 0, -1};
    static const int b[] = {3, 10, 10, 10, 10};
    static const double e[] = {0.0, 0.25, 0.25, 0.25, 0.25};
    f->num_species = 5; f->num_reactions = 9;
    fsp_fixture_fill_SM(f, sm);
    f->prop_x = hog1p_prop; f->prop_t = hog1p_tfun;
    f->num_tv = 1; f->tv_reactions[0] = 2;
    f->num_constr = 5;
    for (int i = 0; i < 5; ++i) { f->bounds[i] = b[i]; f->expansion[i] = e[i]; }
    f->t_final = 180.0; f->fsp_tol = 1.0e-4;
    return 0;
  }
  if (!strcmp(name, "transcr_reg_6d")) {
    static const int sm[] = {1, -1, 0, 0, 0, 0, 0, 0, -2, 2,
                             0, 0, 0, 0, -1, 1, -1, 1, 1, -1,
                             0, 0, 0, 0, -1, 1, 0, 0, 0, 0,
                             0, 0, 0, 0, 1, -1, -1, 1, 0, 0,
                             0, 0, 0, 0, 0, 0, 1, -1, 0, 0,
                             0, 0, 1, -1, 0, 0, 0, 0, 0, 0};
    static const int b[] = {10, 6, 1, 2, 1, 1};
    static const int x0[] = {2, 6, 0, 2, 0, 0};
    f->num_species = 6; f->num_reactions = 10;
    fsp_fixture_fill_SM(f, sm);
    f->prop_x = transcr_prop; f->prop_t = transcr_tfun;
    f->num_tv = 3; f->tv_reactions[0] = 4; f->tv_reactions[1] = 6; f->tv_reactions[2] = 8;
    f->num_constr = 6;
    for (int i = 0; i < 6; ++i) { f->bounds[i] = b[i]; f->expansion[i] = 0.2; f->x0[i] = x0[i]; }
    f->t_final = 300.0; f->fsp_tol = 1.0e-4;
    return 0;
  }
  if (!strcmp(name, "birth_death_3d") || !strcmp(name, "birth_death_3d_tv")) {
    static const int sm[] = {1, -1, 0, 0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 0, 0, 1, -1};
    f->num_species = 3; f->num_reactions = 6;
    fsp_fixture_fill_SM(f, sm);
    f->prop_x = bd3_prop; f->prop_t = bd3_tfun;
    if (!strcmp(name, "birth_death_3d_tv")) { f->num_tv = 3; f->tv_reactions[0] = 0; f->tv_reactions[1] = 2; f->tv_reactions[2] = 4; }
    f->num_constr = 3;
    for (int i = 0; i < 3; ++i) { f->bounds[i] = 99; f->expansion[i] = 0.2; }
    f->t_final = 1.0; f->fsp_tol = -1.0;
    return 0;
  }
  return -1;
}

#ifdef __cplusplus
}
#endif
#endif /* FSP_MODELS_H_ */

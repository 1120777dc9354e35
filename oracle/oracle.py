"""ctypes front-end to the CPU oracle (oracle/fsp_oracle.c) plus small numpy/scipy restatements of
the solver-level reference semantics (FSP driver loop, dense expm).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under pacmensl_b200/ imports this module.

Parity pin: analytic known-answer tests of the reference (tests/test_oracle_kats.py); the reference
binary itself cannot be built in this image.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PROP_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p)
TCOEF_FN = C.CFUNCTYPE(C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double), C.c_void_p)
CONSTR_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libfsp_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        sigs = {
            "orc_set_create": (vp, [ci, ci, ip]),
            "orc_set_destroy": (None, [vp]),
            "orc_set_set_shape": (ci, [vp, ci, vp, ip, vp]),
            "orc_set_set_bounds": (ci, [vp, ci, ip]),
            "orc_set_add_states": (ci, [vp, ci, ci, ip]),
            "orc_set_state2index": (None, [vp, ci, ip, ip]),
            "orc_set_check_constraints": (ci, [vp, ci, ip, ip]),
            "orc_set_expand": (ci, [vp]),
            "orc_set_num_states": (ci, [vp]),
            "orc_set_num_species": (ci, [vp]),
            "orc_set_num_reactions": (ci, [vp]),
            "orc_set_num_constraints": (ci, [vp]),
            "orc_set_copy_states": (None, [vp, ip]),
            "orc_set_copy_status": (None, [vp, C.POINTER(C.c_byte)]),
            "orc_mat_create": (vp, [ci]),
            "orc_mat_destroy": (None, [vp]),
            "orc_mat_destroy_values": (None, [vp]),
            "orc_mat_generate": (ci, [vp, vp, ci, ip, vp, vp, ci, ip, vp, vp]),
            "orc_mat_action": (ci, [vp, cd, dp, dp]),
            "orc_mat_action_coef": (ci, [vp, dp, dp, dp]),
            "orc_mat_action_fused": (ci, [vp, cd, dp, dp]),
            "orc_mat_flops": (ci, [vp]),
            "orc_mat_num_rows": (ci, [vp]),
            "orc_mat_num_states": (ci, [vp]),
            "orc_mat_col": (ip, [vp]),
            "orc_mat_off": (dp, [vp]),
            "orc_mat_diag": (dp, [vp]),
            "orc_mat_num_tv": (ci, [vp]),
            "orc_mat_num_ti": (ci, [vp]),
            "orc_mat_tv": (ip, [vp]),
            "orc_mat_ti": (ip, [vp]),
            "orc_mat_sink_nnz": (ci, [vp, ci, ci]),
            "orc_mat_sink_inz": (ip, [vp, ci, ci]),
            "orc_mat_sink_val": (dp, [vp, ci, ci]),
            "orc_mat_dense": (ci, [vp, cd, dp]),
            "orc_expand_vec": (None, [ci, dp, ip, ci, dp]),
            "orc_set_from_fixture": (vp, [C.c_char_p, ip]),
            "orc_mat_generate_fixture": (ci, [vp, vp, C.c_char_p]),
            "orc_mat_generate_lattice": (ci, [vp, ci, ci, ci, ci]),
            "orc_set_num_threads": (None, [ci]),
            "orc_fixture_tcoef": (ci, [C.c_char_p, cd, dp]),
            "orc_fixture_info": (ci, [C.c_char_p, ip, ip, dp, ip, dp]),
            "orc_num_threads": (ci, []),
            "orc_vec_dot": (cd, [C.c_long, dp, dp]),
            "orc_vec_axpy": (None, [C.c_long, cd, dp, dp]),
            "orc_vec_scale": (None, [C.c_long, cd, dp]),
            "orc_vec_copy": (None, [C.c_long, dp, dp]),
            "orc_vec_maxpy": (None, [C.c_long, ci, dp, dp, C.c_long, dp]),
        }
        for name, (res, args) in sigs.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _as_states(X, S=None):
    """Accept (m, S) row-per-state arrays; memory layout == reference column-major S x m."""
    X = np.ascontiguousarray(np.asarray(X, dtype=np.int32))
    if X.ndim == 1:
        X = X.reshape(1, -1) if S is None or X.size == S else X.reshape(-1, S)
    return X


class StateSet:
    """Oracle mirror of StateSetConstrained (np = 1)."""

    def __init__(self, SM=None, fixture=None, bounds=None):
        L = lib()
        self._keep = []
        if fixture is not None:
            b = None if bounds is None else np.ascontiguousarray(bounds, dtype=np.int32)
            self.h = L.orc_set_from_fixture(fixture.encode(), None if b is None else _ip(b))
            if not self.h:
                raise ValueError("unknown fixture " + fixture)
        else:
            SM = np.asarray(SM, dtype=np.int32)  # S x R as written in the reference
            self.S, self.R = SM.shape
            sm_cm = np.ascontiguousarray(SM.T)  # column major S x R
            self.h = L.orc_set_create(self.S, self.R, _ip(sm_cm))
        self.S = L.orc_set_num_species(self.h)
        self.R = L.orc_set_num_reactions(self.h)

    def __del__(self):
        try:
            lib().orc_set_destroy(self.h)
        except Exception:
            pass

    def set_shape(self, bounds, lhs=None):
        b = np.ascontiguousarray(bounds, dtype=np.int32)
        cb = None
        if lhs is not None:
            def _cb(S, K, m, xs, out, args, _lhs=lhs):
                X = np.ctypeslib.as_array(xs, shape=(m, S))
                O = np.ctypeslib.as_array(out, shape=(m, K))
                return int(_lhs(X, O) or 0)
            cb = CONSTR_FN(_cb)
            self._keep.append(cb)
        return lib().orc_set_set_shape(self.h, len(b), C.cast(cb, C.c_void_p) if cb else None, _ip(b), None)

    def set_bounds(self, bounds):
        b = np.ascontiguousarray(bounds, dtype=np.int32)
        return lib().orc_set_set_bounds(self.h, len(b), _ip(b))

    def add_states(self, X):
        X = _as_states(X)
        return lib().orc_set_add_states(self.h, X.shape[1], X.shape[0], _ip(X))

    def expand(self):
        return lib().orc_set_expand(self.h)

    def state2index(self, X):
        X = _as_states(X, self.S)
        out = np.empty(X.shape[0], dtype=np.int32)
        lib().orc_set_state2index(self.h, X.shape[0], _ip(X), _ip(out))
        return out

    def check_constraints(self, X):
        X = _as_states(X, self.S).copy()
        K = lib().orc_set_num_constraints(self.h)
        out = np.empty((K, X.shape[0]), dtype=np.int32)
        ierr = lib().orc_set_check_constraints(self.h, X.shape[0], _ip(X), _ip(out))
        if ierr:
            raise RuntimeError("lhs callback failed")
        return out

    @property
    def n(self):
        return lib().orc_set_num_states(self.h)

    @property
    def K(self):
        return lib().orc_set_num_constraints(self.h)

    def states(self):
        out = np.empty((self.n, self.S), dtype=np.int32)
        lib().orc_set_copy_states(self.h, _ip(out))
        return out

    def status(self):
        out = np.empty(self.n, dtype=np.int8)
        lib().orc_set_copy_status(self.h, out.ctypes.data_as(C.POINTER(C.c_byte)))
        return out


class FspMatrix:
    """Oracle mirror of FspMatrixBase (constrained=False) / FspMatrixConstrained (constrained=True)."""

    def __init__(self, constrained=True):
        self.h = lib().orc_mat_create(1 if constrained else 0)
        self._keep = []

    def __del__(self):
        try:
            lib().orc_mat_destroy(self.h)
        except Exception:
            pass

    def generate_fixture(self, state_set, name):
        ierr = lib().orc_mat_generate_fixture(self.h, state_set.h, name.encode())
        self.set = state_set
        return ierr

    def generate(self, state_set, prop_x, prop_t=None, tv=(), enable=()):
        """prop_x(reaction, X[m,S]) -> array[m];  prop_t(t, out[R]) -> int."""
        def _px(r, S, m, xs, out, args):
            X = np.ctypeslib.as_array(xs, shape=(m, S))
            O = np.ctypeslib.as_array(out, shape=(m,))
            res = prop_x(r, X)
            if res is None:
                return -1
            O[:] = res
            return 0

        def _pt(t, R, out, args):
            O = np.ctypeslib.as_array(out, shape=(R,))
            return int(prop_t(t, O) or 0) if prop_t else 0

        px, pt = PROP_FN(_px), TCOEF_FN(_pt)
        self._keep += [px, pt]
        tv = np.ascontiguousarray(tv, dtype=np.int32)
        en = np.ascontiguousarray(enable, dtype=np.int32)
        self.set = state_set
        return lib().orc_mat_generate(self.h, state_set.h, len(tv), _ip(tv), C.cast(pt, C.c_void_p),
                                      C.cast(px, C.c_void_p), len(en), _ip(en), None, None)

    def generate_lattice(self, dims, tv=False):
        """Reference-shaped operators of the synthetic birth-death lattice in lexicographic order, built directly
        (no BFS / hash directory): the full 465^3 CPU-baseline workload.  Action only (no per-reaction arrays)."""
        self.set = None
        return lib().orc_mat_generate_lattice(self.h, int(dims[0]), int(dims[1]), int(dims[2]), 1 if tv else 0)

    @property
    def nrows(self):
        return lib().orc_mat_num_rows(self.h)

    @property
    def n(self):
        return lib().orc_mat_num_states(self.h)

    def action(self, t, x, fused=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.nrows, dtype=np.float64)
        f = lib().orc_mat_action_fused if fused else lib().orc_mat_action
        ierr = f(self.h, float(t), _dp(x), _dp(y))
        return ierr, y

    def action_coef(self, coef, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        c = np.ascontiguousarray(coef, dtype=np.float64)
        y = np.empty(self.nrows, dtype=np.float64)
        ierr = lib().orc_mat_action_coef(self.h, _dp(c), _dp(x), _dp(y))
        return ierr, y

    def action_into(self, t, x, y):
        """No-allocation variant for timing (x, y float64 contiguous)."""
        return lib().orc_mat_action(self.h, float(t), _dp(x), _dp(y))

    def flops(self):
        return lib().orc_mat_flops(self.h)

    def dense(self, t):
        nr = self.nrows
        out = np.empty((nr, nr), dtype=np.float64, order="F")
        ierr = lib().orc_mat_dense(self.h, float(t), _dp(out))
        if ierr:
            raise RuntimeError("t_fun failed")
        return out

    def ell(self):
        """(col[R,n], off[R,n], diag[R,n]) views copied out of the oracle."""
        L = lib()
        n = self.n
        col = np.ctypeslib.as_array(L.orc_mat_col(self.h), shape=(self._R(), n)).copy()
        off = np.ctypeslib.as_array(L.orc_mat_off(self.h), shape=(self._R(), n)).copy()
        diag = np.ctypeslib.as_array(L.orc_mat_diag(self.h), shape=(self._R(), n)).copy()
        return col, off, diag

    def _R(self):
        return self.set.R

    def tv_ti(self):
        L = lib()
        ntv, nti = L.orc_mat_num_tv(self.h), L.orc_mat_num_ti(self.h)
        tv = np.ctypeslib.as_array(L.orc_mat_tv(self.h), shape=(max(ntv, 1),))[:ntv].copy()
        ti = np.ctypeslib.as_array(L.orc_mat_ti(self.h), shape=(max(nti, 1),))[:nti].copy()
        return tv, ti

    def sinks(self):
        """dict (r, k) -> (state indices, values)"""
        L = lib()
        out = {}
        K = self.nrows - self.n
        for r in range(self._R()):
            for k in range(K):
                c = L.orc_mat_sink_nnz(self.h, r, k)
                if c > 0:
                    out[(r, k)] = (np.ctypeslib.as_array(L.orc_mat_sink_inz(self.h, r, k), shape=(c,)).copy(),
                                   np.ctypeslib.as_array(L.orc_mat_sink_val(self.h, r, k), shape=(c,)).copy())
        return out


def fixture_tcoef(name, t, R):
    out = np.zeros(R, dtype=np.float64)
    ierr = lib().orc_fixture_tcoef(name.encode(), float(t), _dp(out))
    return ierr, out


def fixture_info(name):
    """Parameters of a named workload of pacmensl_b200/fixtures/fsp_models.h (the reference example's settings)."""
    dims = np.zeros(4, np.int32)
    bounds, x0 = np.zeros(16, np.int32), np.zeros(16, np.int32)
    expansion, params = np.zeros(16), np.zeros(4)
    if lib().orc_fixture_info(name.encode(), _ip(dims), _ip(bounds), _dp(expansion), _ip(x0), _dp(params)):
        raise ValueError("unknown fixture " + name)
    S, R, K, ntv = (int(v) for v in dims)
    return dict(S=S, R=R, K=K, n_tv=ntv, bounds=bounds[:K].copy(), expansion=expansion[:K].copy(), x0=x0[:S].copy(),
                t_final=float(params[0]), fsp_tol=float(params[1]), rtol=float(params[2]), atol=float(params[3]))


def expand_vec(p_old, new_idx, n_new):
    p_old = np.ascontiguousarray(p_old, dtype=np.float64)
    new_idx = np.ascontiguousarray(new_idx, dtype=np.int32)
    out = np.empty(n_new, dtype=np.float64)
    lib().orc_expand_vec(len(p_old), _dp(p_old), _ip(new_idx), n_new, _dp(out))
    return out


def num_threads():
    return lib().orc_num_threads()


def use_all_cores():
    """OpenMP thread count = the cores this process may run on, whatever OMP_NUM_THREADS says (torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would silently turn the CPU baseline into a single-core run)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(n)
    return num_threads()

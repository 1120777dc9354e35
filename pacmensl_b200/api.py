"""Python binding of the host-level C ABI (include/pacmensl_b200_host.h): the objects mirror the reference's
C++ classes (StateSetConstrained, Model, FspMatrixBase/FspMatrixConstrained, FspSolverMultiSinks).

Used by bench.py and the tests.  All compute happens in libpacmensl_b200.so (C++ host + sm_100a kernels);
there is no Python/CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import FspError

vp, ci, cl, cd = C.c_void_p, C.c_int, C.c_long, C.c_double
ip, dp, lp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_long)
vpp = C.POINTER(C.c_void_p)
PROP_FN, TCOEF_FN, CONSTR_FN = _capi.PROP_FN, _capi.TCOEF_FN, _capi.CONSTR_FN

HOST_SIGNATURES = {
    "pfsp_init": (ci, [ci, C.c_char_p, ci, ci]),
    "pfsp_finalize": (ci, []),
    "pfsp_p2p_enabled": (ci, []),
    "pfsp_check": (ci, []),
    "pfsp_last_error": (C.c_char_p, []),
    "pfsp_set_create": (ci, [vpp]),
    "pfsp_set_destroy": (ci, [vp]),
    "pfsp_set_stoichiometry": (ci, [vp, ci, ci, ip]),
    "pfsp_set_shape": (ci, [vp, ci, ip, vp, vp]),
    "pfsp_set_shape_bounds": (ci, [vp, ci, ip]),
    "pfsp_set_add_states": (ci, [vp, ci, ci, ip]),
    "pfsp_set_add_box_lattice": (ci, [vp, ci, ip]),
    "pfsp_set_expand": (ci, [vp]),
    "pfsp_set_sizes": (ci, [vp, ip, ip, ip]),
    "pfsp_set_copy_states": (ci, [vp, ip]),
    "pfsp_set_state2index": (ci, [vp, ci, ip, ip]),
    "pfsp_model_create": (ci, [vpp, ci, ci, ip, vp, vp, vp, vp, ci, ip]),
    "pfsp_model_from_fixture": (ci, [vpp, C.c_char_p, ip, ip, ip, ip, dp, ip, dp, dp, dp, dp, vpp]),
    "pfsp_model_set_mass_action": (ci, [vp, dp, ip]),
    "pfsp_model_set_factor_table": (ci, [vp, ci, ci, ci, dp]),
    "pfsp_model_attach_device_form": (ci, [vp, C.c_char_p]),
    "pfsp_model_get_stoichiometry": (ci, [vp, ip]),
    "pfsp_model_destroy": (ci, [vp]),
    "pfsp_mat_create": (ci, [vpp, ci]),
    "pfsp_mat_destroy": (ci, [vp]),
    "pfsp_mat_generate": (ci, [vp, vp, vp]),
    "pfsp_mat_clear": (ci, [vp]),
    "pfsp_mat_set_variant": (ci, [vp, ci]),
    "pfsp_mat_info": (ci, [vp, ip, lp, dp]),
    "pfsp_mat_action": (ci, [vp, cd, vp, vp]),
    "pfsp_mat_action_host": (ci, [vp, cd, vp, vp]),
    "pfsp_mat_halo_only": (ci, [vp, vp, vp, lp]),
    "pfsp_solver_create": (ci, [vpp, ci]),
    "pfsp_solver_destroy": (ci, [vp]),
    "pfsp_solver_set_model": (ci, [vp, vp]),
    "pfsp_solver_set_initial_bounds": (ci, [vp, ci, ip]),
    "pfsp_solver_set_constraint_function": (ci, [vp, vp, vp]),
    "pfsp_solver_set_expansion_factors": (ci, [vp, ci, dp]),
    "pfsp_solver_set_initial_distribution": (ci, [vp, ci, ci, ip, dp]),
    "pfsp_solver_set_ode_tolerances": (ci, [vp, cd, cd]),
    "pfsp_solver_set_verbosity": (ci, [vp, ci]),
    "pfsp_solver_set_krylov": (ci, [vp, ci, ci, ci]),
    "pfsp_solver_set_warm_restart": (ci, [vp, ci]),
    "pfsp_solver_set_sharded_state_set": (ci, [vp, ci]),
    "pfsp_set_set_sharded": (ci, [vp, ci]),
    "pfsp_set_is_sharded": (ci, [vp]),
    "pfsp_set_remember_local": (ci, [vp]),
    "pfsp_set_remembered_indices": (ci, [vp, ci, ip]),
    "pfsp_solver_num_warm_restarts": (ci, [vp, ip]),
    "pfsp_solver_setup": (ci, [vp]),
    "pfsp_solver_solve": (ci, [vp, cd, cd, cd, ip, ip]),
    "pfsp_solver_copy_result": (ci, [vp, ip, dp]),
    "pfsp_solver_stats": (ci, [vp, ip, ip, lp, ip, ip]),
    "pfsp_solver_clear": (ci, [vp]),
}

_bound = False
KRYLOV, CVODE, PETSC = 0, 1, 2


def lib():
    global _bound
    L = _capi.lib()
    if not _bound:
        for name, (res, args) in HOST_SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _bound = True
    return L


def check(ierr, what):
    if ierr != 0:
        msg = lib().pfsp_last_error()
        raise FspError("%s failed (%d): %s" % (what, ierr, msg.decode() if msg else ""))


def _ip(a):
    return a.ctypes.data_as(ip)


def _dp(a):
    return a.ctypes.data_as(dp)


def init(device=0, dist=None):
    """Select the device and, when torch.distributed is initialised with world size > 1, join the NCCL world
    communicator of the library (the unique id is made on rank 0 and broadcast through torch.distributed)."""
    L = lib()
    rank, size = 0, 1
    idbuf = C.create_string_buffer(128)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        import torch
        rank, size = dist.get_rank(), dist.get_world_size()
        _capi.check(L.fsp_device_set(device), "fsp_device_set")
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            _capi.check(L.fspcomm_unique_id(idbuf), "fspcomm_unique_id")
            t.copy_(torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8))
        dist.broadcast(t, 0)
        idbuf = C.create_string_buffer(bytes(t.cpu().numpy().tobytes()), 128)
    check(L.pfsp_init(device, idbuf, rank, size), "pfsp_init")
    return rank, size


def finalize():
    lib().pfsp_finalize()


def health_check():
    """Device synchronisation + peer-memory health: non-zero once a device-side wait for a peer GPU has timed out."""
    return lib().pfsp_check()


def world_comm():
    """The library's world communicator (fspcomm_t of include/fsp_b200.h) for direct calls of the fspcomm_* functions."""
    f = lib().pfsp_world_comm
    f.restype = vp
    return vp(f())


def p2p_enabled():
    """True when halo exchange / small all-reduces run as fused peer-memory kernels (CUDA IPC over NVLink)."""
    return bool(lib().pfsp_p2p_enabled())


class StateSet:
    def __init__(self, SM, sharded=None):
        SM = np.asarray(SM, dtype=np.int32)  # S x R as written in the reference
        self.S, self.R = SM.shape
        self._sm = np.ascontiguousarray(SM.T)
        h = vp()
        check(lib().pfsp_set_create(C.byref(h)), "pfsp_set_create")
        self.h = h
        if sharded is not None:  # multi-GPU: distributed construction + striped directory (default) or replicated
            check(lib().pfsp_set_set_sharded(self.h, 1 if sharded else 0), "SetSharded")
        check(lib().pfsp_set_stoichiometry(self.h, self.S, self.R, _ip(self._sm)), "SetStoichiometryMatrix")
        self._keep = []

    def __del__(self):
        try:
            lib().pfsp_set_destroy(self.h)
        except Exception:
            pass

    def set_shape(self, bounds, lhs=None, lhs_c=None):
        b = np.ascontiguousarray(bounds, dtype=np.int32)
        fn = None
        if lhs is not None:
            def _cb(S, K, m, xs, out, args, _lhs=lhs):
                X = np.ctypeslib.as_array(xs, shape=(m, S))
                O = np.ctypeslib.as_array(out, shape=(m, K))
                return int(_lhs(X, O) or 0)
            cb = CONSTR_FN(_cb)
            self._keep.append(cb)
            fn = C.cast(cb, vp)
        elif lhs_c:
            fn = lhs_c
        return lib().pfsp_set_shape(self.h, len(b), _ip(b), fn, None)

    def set_bounds(self, bounds):
        b = np.ascontiguousarray(bounds, dtype=np.int32)
        return lib().pfsp_set_shape_bounds(self.h, len(b), _ip(b))

    def add_states(self, X):
        X = np.ascontiguousarray(np.asarray(X, dtype=np.int32))
        if X.ndim == 1:
            X = X.reshape(1, -1)
        return lib().pfsp_set_add_states(self.h, X.shape[1], X.shape[0], _ip(X))

    def add_box_lattice(self, upper):
        u = np.ascontiguousarray(upper, dtype=np.int32)
        check(lib().pfsp_set_add_box_lattice(self.h, len(u), _ip(u)), "AddBoxLattice")

    def expand(self):
        return lib().pfsp_set_expand(self.h)

    def is_sharded(self):
        return bool(lib().pfsp_set_is_sharded(self.h))

    def remember_local(self):
        self._n_rem = self.n_local
        return lib().pfsp_set_remember_local(self.h)

    def remembered_indices(self):
        out = np.empty(self._n_rem, dtype=np.int32)
        check(lib().pfsp_set_remembered_indices(self.h, self._n_rem, _ip(out)), "RememberedIndices")
        return out

    def sizes(self):
        a, b, c = ci(), ci(), ci()
        check(lib().pfsp_set_sizes(self.h, C.byref(a), C.byref(b), C.byref(c)), "sizes")
        return a.value, b.value, c.value

    @property
    def n_local(self):
        return self.sizes()[0]

    @property
    def n_global(self):
        return self.sizes()[1]

    def states(self):
        out = np.empty((self.n_local, self.S), dtype=np.int32)
        check(lib().pfsp_set_copy_states(self.h, _ip(out)), "CopyStatesOnProc")
        return out

    def state2index(self, X):
        X = np.ascontiguousarray(np.asarray(X, dtype=np.int32)).reshape(-1, self.S)
        out = np.empty(X.shape[0], dtype=np.int32)
        check(lib().pfsp_set_state2index(self.h, X.shape[0], _ip(X), _ip(out)), "State2Index")
        return out


class Model:
    def __init__(self, SM=None, prop_x=None, prop_t=None, tv=(), fixture=None, device_form=False):
        """device_form (fixtures only): attach the separable description of the propensities so that matrix generation
        evaluates them on the GPU (hog1p, transcr_reg_6d, birth_death_3d, pure_birth) instead of through prop_x."""
        L = lib()
        self._keep = []
        h = vp()
        self.fixture = None
        self.device_form = False
        if fixture is not None:
            S, R, K = ci(), ci(), ci()
            bounds = np.zeros(16, np.int32)
            expansion = np.zeros(16)
            x0 = np.zeros(16, np.int32)
            tf, ftol, rtol, atol = cd(), cd(), cd(), cd()
            lhs = vp()
            check(L.pfsp_model_from_fixture(C.byref(h), fixture.encode(), C.byref(S), C.byref(R), C.byref(K), _ip(bounds),
                                            _dp(expansion), _ip(x0), C.byref(tf), C.byref(ftol), C.byref(rtol),
                                            C.byref(atol), C.byref(lhs)), "pfsp_model_from_fixture")
            self.S, self.R, self.K = S.value, R.value, K.value
            self.fixture = dict(bounds=bounds[:K.value].copy(), expansion=expansion[:K.value].copy(), x0=x0[:S.value].copy(),
                                t_final=tf.value, fsp_tol=ftol.value, rtol=rtol.value, atol=atol.value,
                                lhs=lhs if lhs.value else None)
        else:
            SM = np.asarray(SM, dtype=np.int32)
            self.S, self.R = SM.shape
            sm = np.ascontiguousarray(SM.T)

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
                return int(prop_t(t, O) or 0)

            px = PROP_FN(_px) if prop_x else None
            pt = TCOEF_FN(_pt) if prop_t else None
            self._keep += [px, pt, sm]
            tvv = np.ascontiguousarray(tv, dtype=np.int32)
            check(L.pfsp_model_create(C.byref(h), self.S, self.R, _ip(sm), C.cast(px, vp) if px else None, None,
                                      C.cast(pt, vp) if pt else None, None, len(tvv), _ip(tvv)), "pfsp_model_create")
        self.h = h
        if fixture is not None and device_form:
            self.device_form = L.pfsp_model_attach_device_form(self.h, fixture.encode()) == 0

    def set_factor_table(self, species, reaction, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        check(lib().pfsp_model_set_factor_table(self.h, int(species), int(reaction), len(v), _dp(v)), "SetFactorTable")

    def stoichiometry(self):
        """S x R matrix as written in the reference (one column per reaction)."""
        sm = np.zeros((self.R, self.S), dtype=np.int32)
        check(lib().pfsp_model_get_stoichiometry(self.h, _ip(sm)), "stoichiometry")
        return np.ascontiguousarray(sm.T)

    def set_mass_action(self, rates, orders):
        r = np.ascontiguousarray(rates, dtype=np.float64)
        o = np.ascontiguousarray(np.asarray(orders, dtype=np.int32).T)  # S x R -> column major
        check(lib().pfsp_model_set_mass_action(self.h, _dp(r), _ip(o)), "SetMassAction")

    def __del__(self):
        try:
            lib().pfsp_model_destroy(self.h)
        except Exception:
            pass


class FspMatrix:
    def __init__(self, constrained=True):
        h = vp()
        check(lib().pfsp_mat_create(C.byref(h), 1 if constrained else 0), "pfsp_mat_create")
        self.h = h

    def __del__(self):
        try:
            lib().pfsp_mat_destroy(self.h)
        except Exception:
            pass

    def generate(self, state_set, model):
        self._refs = (state_set, model)
        return lib().pfsp_mat_generate(self.h, state_set.h, model.h)

    def destroy_values(self):
        return lib().pfsp_mat_clear(self.h)

    def set_variant(self, v):
        lib().pfsp_mat_set_variant(self.h, v)

    def info(self):
        n, f, b = ci(), cl(), cd()
        check(lib().pfsp_mat_info(self.h, C.byref(n), C.byref(f), C.byref(b)), "pfsp_mat_info")
        return n.value, f.value, b.value

    @property
    def n_rows(self):
        return self.info()[0]

    def action(self, t, x, y):
        """x, y: torch float64 CUDA tensors with n_rows entries."""
        return lib().pfsp_mat_action(self.h, float(t), vp(x.data_ptr()), vp(y.data_ptr()))

    def halo_only(self, x, y):
        """Diagnostics (multi-GPU): only the halo exchange of an Action; returns the bytes this rank pushes."""
        b = cl()
        check(lib().pfsp_mat_halo_only(self.h, vp(x.data_ptr()), vp(y.data_ptr()), C.byref(b)), "halo_only")
        return b.value

    def action_host(self, t, x, y):
        """x, y: host buffers (numpy arrays or pinned torch tensors)."""
        xp = x.ctypes.data if isinstance(x, np.ndarray) else x.data_ptr()
        yp = y.ctypes.data if isinstance(y, np.ndarray) else y.data_ptr()
        return lib().pfsp_mat_action_host(self.h, float(t), vp(xp), vp(yp))


class FspSolver:
    def __init__(self, ode_type=CVODE):
        h = vp()
        check(lib().pfsp_solver_create(C.byref(h), ode_type), "pfsp_solver_create")
        self.h = h
        self._keep = []

    def __del__(self):
        try:
            lib().pfsp_solver_destroy(self.h)
        except Exception:
            pass

    def set_model(self, model):
        self._keep.append(model)
        return lib().pfsp_solver_set_model(self.h, model.h)

    def set_initial_bounds(self, b):
        b = np.ascontiguousarray(b, dtype=np.int32)
        return lib().pfsp_solver_set_initial_bounds(self.h, len(b), _ip(b))

    def set_constraint_function_c(self, fn_ptr):
        return lib().pfsp_solver_set_constraint_function(self.h, fn_ptr, None)

    def set_expansion_factors(self, f):
        f = np.ascontiguousarray(f, dtype=np.float64)
        return lib().pfsp_solver_set_expansion_factors(self.h, len(f), _dp(f))

    def set_initial_distribution(self, X, p):
        X = np.ascontiguousarray(np.asarray(X, dtype=np.int32))
        if X.ndim == 1:
            X = X.reshape(1, -1)
        p = np.ascontiguousarray(p, dtype=np.float64)
        return lib().pfsp_solver_set_initial_distribution(self.h, X.shape[1], X.shape[0], _ip(X), _dp(p))

    def set_ode_tolerances(self, rtol, atol):
        return lib().pfsp_solver_set_ode_tolerances(self.h, rtol, atol)

    def set_verbosity(self, v):
        return lib().pfsp_solver_set_verbosity(self.h, v)

    def set_krylov(self, q_iop=2, m_min=25, m_max=60):
        return lib().pfsp_solver_set_krylov(self.h, q_iop, m_min, m_max)

    def set_sharded_state_set(self, on=True):
        return lib().pfsp_solver_set_sharded_state_set(self.h, 1 if on else 0)

    def set_warm_restart(self, on):
        return lib().pfsp_solver_set_warm_restart(self.h, 1 if on else 0)

    def warm_restarts(self):
        n = ci()
        check(lib().pfsp_solver_num_warm_restarts(self.h, C.byref(n)), "num_warm_restarts")
        return n.value

    def setup(self):
        return lib().pfsp_solver_setup(self.h)

    def solve(self, t_final, fsp_tol=-1.0, t_init=0.0):
        n, S = ci(), ci()
        check(lib().pfsp_solver_solve(self.h, t_final, fsp_tol, t_init, C.byref(n), C.byref(S)), "Solve")
        states = np.empty((n.value, S.value), dtype=np.int32)
        p = np.empty(n.value, dtype=np.float64)
        check(lib().pfsp_solver_copy_result(self.h, _ip(states), _dp(p)), "copy_result")
        return states, p

    def stats(self):
        ng, ne, nr, K = ci(), ci(), cl(), ci()
        b = np.zeros(32, np.int32)
        check(lib().pfsp_solver_stats(self.h, C.byref(ng), C.byref(ne), C.byref(nr), C.byref(K), _ip(b)), "stats")
        return dict(n_states=ng.value, expansions=ne.value, rhs_evals=nr.value, bounds=b[:K.value].tolist())

    def clear(self):
        return lib().pfsp_solver_clear(self.h)


def fixture_set_and_matrix(name, bounds=None, model=None, constrained=True, sharded=None):
    """StateSetConstrained (x0 added, Expand()ed within `bounds`) + generated FspMatrix of a named workload."""
    m = model or Model(fixture=name)
    fx = m.fixture
    st = StateSet(m.stoichiometry(), sharded=sharded)
    b = fx["bounds"] if bounds is None else bounds
    ierr = st.set_shape(b, lhs_c=fx["lhs"])
    assert ierr == 0, ierr
    assert st.add_states(fx["x0"].reshape(1, -1)) == 0
    assert st.expand() == 0
    mat = FspMatrix(constrained=constrained)
    ierr = mat.generate(st, m)
    if ierr:
        raise FspError("GenerateValues failed: %d" % ierr)
    mat._model = m
    return st, mat


def fixture_solver(name, ode_type=CVODE, custom_constraints=True):
    """FspSolver configured like the reference example/test of that name."""
    m = Model(fixture=name)
    fx = m.fixture
    s = FspSolver(ode_type)
    s.set_model(m)
    s.set_initial_bounds(fx["bounds"])
    s.set_expansion_factors(fx["expansion"])
    if fx["lhs"] is not None:
        s.set_constraint_function_c(fx["lhs"])
    s.set_initial_distribution(fx["x0"].reshape(1, -1), [1.0])
    s.set_ode_tolerances(fx["rtol"], fx["atol"])
    return s, m

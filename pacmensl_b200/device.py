"""Thin Python handles over the kernel-level C ABI (include/fsp_b200.h).

PyTorch is used only as plumbing here: device buffers (torch tensors), streams and torch.distributed.
All arithmetic happens in libpacmensl_b200.so's sm_100a kernels.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import check, lib


def _torch():
    import torch
    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _stream(stream):
    if stream is None:
        return C.c_void_p(_torch().cuda.current_stream().cuda_stream)
    return C.c_void_p(stream)


class DeviceStateSet:
    """fspset_* : state list + hash directory + Expand on the device (mirrors StateSetConstrained)."""

    def __init__(self, SM):
        SM = np.asarray(SM, dtype=np.int32)  # S x R as written in the reference
        self.S, self.R = SM.shape
        self.SM = np.ascontiguousarray(SM.T)  # column major: SM[r*S + s]
        h = C.c_void_p()
        check(lib().fspset_create(C.byref(h), self.S, self.R, _ip(self.SM)), "fspset_create")
        self.h = h
        self.K = 0
        self._keep = []

    def __del__(self):
        try:
            lib().fspset_destroy(self.h)
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
            cb = _capi.CONSTR_FN(_cb)
            self._keep.append(cb)
        self.K = len(b)
        return lib().fspset_set_shape(self.h, len(b), C.cast(cb, C.c_void_p) if cb else None, _ip(b), None)

    def set_shape_c(self, bounds, lhs_ptr):
        """lhs_ptr: address of a C function with the fsp_constr_multi_fn contract (or None)."""
        b = np.ascontiguousarray(bounds, dtype=np.int32)
        self.K = len(b)
        return lib().fspset_set_shape(self.h, len(b), lhs_ptr, _ip(b), None)

    def set_bounds(self, bounds):
        b = np.ascontiguousarray(bounds, dtype=np.int32)
        self.K = len(b)
        return lib().fspset_set_bounds(self.h, len(b), _ip(b))

    def add_states(self, X):
        X = np.ascontiguousarray(np.asarray(X, dtype=np.int32))
        if X.ndim == 1:
            X = X.reshape(1, -1)
        return lib().fspset_add_states(self.h, X.shape[1], X.shape[0], X.ctypes.data_as(C.c_void_p), 0)

    def add_box_lattice(self, upper):
        u = np.ascontiguousarray(upper, dtype=np.int32)
        check(lib().fspset_add_box_lattice(self.h, _ip(u)), "fspset_add_box_lattice")

    def expand(self):
        return lib().fspset_expand(self.h)

    @property
    def n(self):
        n = C.c_int()
        check(lib().fspset_num_states(self.h, C.byref(n)), "fspset_num_states")
        return n.value

    def states(self, first=0, count=None):
        count = self.n - first if count is None else count
        out = np.empty((count, self.S), dtype=np.int32)
        check(lib().fspset_copy_states(self.h, first, count, _ip(out)), "fspset_copy_states")
        return out

    def status(self):
        out = np.empty(self.n, dtype=np.int8)
        check(lib().fspset_copy_status(self.h, 0, self.n, out.ctypes.data_as(C.POINTER(C.c_byte))), "copy_status")
        return out

    def state2index(self, X):
        X = np.ascontiguousarray(np.asarray(X, dtype=np.int32))
        if X.ndim == 1:
            X = X.reshape(-1, self.S)
        out = np.empty(X.shape[0], dtype=np.int32)
        check(lib().fspset_state2index(self.h, X.shape[0], X.ctypes.data_as(C.c_void_p), 0,
                                       out.ctypes.data_as(C.c_void_p), 0), "fspset_state2index")
        return out

    def lookup_shifted(self, nu, sign, out=None, first=0, count=None):
        torch = _torch()
        count = self.n - first if count is None else count
        if out is None:
            out = torch.empty(count, dtype=torch.int32, device="cuda")
        nu = np.ascontiguousarray(nu, dtype=np.int32)
        check(lib().fspset_lookup_shifted(self.h, _ip(nu), sign, first, count, _ptr(out)), "fspset_lookup_shifted")
        return out

    def check_constraints_shifted(self, nu, first=0, count=None):
        torch = _torch()
        count = self.n - first if count is None else count
        out = torch.empty((self.K, count), dtype=torch.int32, device="cuda")
        nu = np.ascontiguousarray(nu, dtype=np.int32)
        check(lib().fspset_check_constraints_shifted(self.h, _ip(nu), first, count, _ptr(out)), "check_constraints")
        return out

    def eval_mass_action(self, rate, order, nu, sign, out=None, first=0, count=None):
        torch = _torch()
        count = self.n - first if count is None else count
        if out is None:
            out = torch.empty(count, dtype=torch.float64, device="cuda")
        order = np.ascontiguousarray(order, dtype=np.int32)
        nu = np.ascontiguousarray(nu, dtype=np.int32)
        check(lib().fspset_eval_mass_action(self.h, float(rate), _ip(order), _ip(nu), sign, first, count, _ptr(out)),
              "fspset_eval_mass_action")
        return out


class DeviceFspMatrix:
    """fspmat_* : the fused FSP operator on the device."""

    def __init__(self):
        h = C.c_void_p()
        check(lib().fspmat_create(C.byref(h)), "fspmat_create")
        self.h = h
        self.n_rows = 0
        self.R = 0

    def __del__(self):
        try:
            lib().fspmat_destroy(self.h)
        except Exception:
            pass

    def generate(self, n_states, n_reactions, tv, ti, col, off, diag, ld, on_device, K=0, sink_ptr=None,
                 sink_idx=None, sink_val=None, owns_sinks=1, n_ghost=0, n_rows=None):
        """col/off/diag: planes ordered TV first then TI (numpy arrays if on_device == 0, torch CUDA tensors else)."""
        d = _capi.FspMatDesc()
        tv = np.ascontiguousarray(tv, dtype=np.int32)
        ti = np.ascontiguousarray(ti, dtype=np.int32)
        d.n_states = n_states
        d.n_rows = n_rows if n_rows is not None else n_states + (K if owns_sinks else 0)
        d.n_reactions = n_reactions
        d.n_tv, d.n_ti = len(tv), len(ti)
        d.tv_reactions, d.ti_reactions = _ip(tv), _ip(ti)
        if on_device:
            d.col, d.off, d.diag = col.data_ptr(), off.data_ptr(), diag.data_ptr()
        else:
            col = np.ascontiguousarray(col, dtype=np.int32)
            off = np.ascontiguousarray(off, dtype=np.float64)
            diag = np.ascontiguousarray(diag, dtype=np.float64)
            d.col, d.off, d.diag = col.ctypes.data, off.ctypes.data, diag.ctypes.data
        d.ld = ld
        d.arrays_on_device = 1 if on_device else 0
        d.n_constr = K
        keep = [tv, ti, col, off, diag]
        if K > 0:
            sp = np.ascontiguousarray(sink_ptr, dtype=np.int64)
            d.sink_ptr = sp.ctypes.data_as(C.POINTER(C.c_long))
            if on_device:
                d.sink_idx = sink_idx.data_ptr() if sink_idx is not None else None
                d.sink_val = sink_val.data_ptr() if sink_val is not None else None
            else:
                si = np.ascontiguousarray(sink_idx, dtype=np.int32)
                sv = np.ascontiguousarray(sink_val, dtype=np.float64)
                d.sink_idx, d.sink_val = si.ctypes.data, sv.ctypes.data
                keep += [si, sv]
            keep.append(sp)
        d.owns_sinks = owns_sinks
        d.n_ghost = n_ghost
        check(lib().fspmat_generate(self.h, C.byref(d)), "fspmat_generate")
        self.n_rows = d.n_rows
        self.n = n_states
        self.R = n_reactions
        self.K = K

    def set_variant(self, v):
        check(lib().fspmat_set_variant(self.h, v), "fspmat_set_variant")

    def clear(self):
        check(lib().fspmat_clear(self.h), "fspmat_clear")

    def action(self, coef, x, y, ghost=None, sink_out=None, stream=None):
        c = np.ascontiguousarray(coef, dtype=np.float64)
        if len(c) < self.R:
            c = np.concatenate([c, np.ones(self.R - len(c))])
        check(lib().fspmat_action(self.h, _dp(c), _ptr(x), _ptr(ghost), _ptr(y), _ptr(sink_out), _stream(stream)),
              "fspmat_action")

    def action_fused(self, coef, x, y, alpha=1.0, beta=0.0, scale=None, dot_vecs=(), dot_out=None, stream=None):
        """y = scale .* (beta x + alpha A x); dot_out[k] = <y, dot_vecs[k]> (None entry: <y, y>)."""
        c = np.ascontiguousarray(coef, dtype=np.float64)
        if len(c) < self.R:
            c = np.concatenate([c, np.ones(self.R - len(c))])
        ep = _capi.FspMatEpilogue()
        ep.alpha, ep.beta = float(alpha), float(beta)
        ep.scale_dev = scale.data_ptr() if scale is not None else None
        ep.n_dots = len(dot_vecs)
        for k, v in enumerate(dot_vecs):
            ep.dot_vec_dev[k] = v.data_ptr() if v is not None else None
        ep.dot_out_dev = dot_out.data_ptr() if dot_out is not None else None
        check(lib().fspmat_action_fused(self.h, _dp(c), _ptr(x), _ptr(y), C.byref(ep), _stream(stream)), "fspmat_action_fused")

    def flops(self):
        f = C.c_long()
        check(lib().fspmat_flops(self.h, C.byref(f)), "fspmat_flops")
        return f.value

    def action_bytes(self):
        b = C.c_double()
        check(lib().fspmat_action_bytes(self.h, C.byref(b)), "fspmat_action_bytes")
        return b.value

    def dense(self, coef):
        c = np.ascontiguousarray(coef, dtype=np.float64)
        out = np.empty((self.n_rows, self.n_rows), dtype=np.float64, order="F")
        check(lib().fspmat_dense(self.h, _dp(c), _dp(out)), "fspmat_dense")
        return out


# ---- device vector helpers (operate on torch float64 CUDA tensors) ------------------------------------
def vec_call(name, *args):
    check(getattr(lib(), name)(*args), name)


def launch_count():
    return lib().fsp_launch_count()

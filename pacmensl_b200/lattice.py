"""Synthetic 3-D birth-death lattice workload (SURVEY.md section 8d / BASELINE.json config 4), generated
entirely on the device through the C ABI: box lattice -> hash directory -> State2Index of every
neighbour -> mass-action propensities -> sink lists -> fspmat_generate(arrays_on_device=1).

S=3, SM = [+e1,-e1,+e2,-e2,+e3,-e3], births (40,30,20), deaths gamma*x with gamma=(1.0,1.5,2.0),
constraints x_s <= L_s (K=3), lexicographic order with species 0 fastest.
`tv=True`: the three birth reactions are time-varying with c(t) = 1 + 0.5 sin(0.1 t).
"""
import ctypes as C

import numpy as np

from ._capi import check, lib
from .device import DeviceFspMatrix, DeviceStateSet

SM = np.array([[1, -1, 0, 0, 0, 0], [0, 0, 1, -1, 0, 0], [0, 0, 0, 0, 1, -1]], dtype=np.int32)
BIRTH = (40.0, 30.0, 20.0)
DEATH = (1.0, 1.5, 2.0)


def tcoef(t, tv=True):
    c = np.ones(6)
    if tv:
        c[[0, 2, 4]] = 1.0 + 0.5 * np.sin(0.1 * t)
    return c


def mass_action_desc(r):
    s = r // 2
    order = np.zeros(3, dtype=np.int32)
    if r % 2 == 0:
        return BIRTH[s], order
    order[s] = 1
    return DEATH[s], order


def build_birth_death_lattice(upper, tv=False, expand=True, row_range=None, return_set=False):
    """Returns (DeviceFspMatrix, n_states).  row_range=(first, count, owns_sinks) restricts the rows
    (multi-GPU block partition); columns outside the range become ghost slots via remap_ghosts()."""
    import torch
    upper = [int(u) for u in upper]
    st = DeviceStateSet(SM)
    assert st.set_shape(upper) == 0
    st.add_box_lattice(upper)
    if expand:
        assert st.expand() == 0
    N = st.n
    first, count, owns = (0, N, 1) if row_range is None else row_range
    tv_r = [0, 2, 4] if tv else []
    ti_r = [r for r in range(6) if r not in tv_r]
    order = tv_r + ti_r
    P = len(order)
    ld = count
    col = torch.empty((P, ld), dtype=torch.int32, device="cuda")
    off = torch.empty((P, ld), dtype=torch.float64, device="cuda")
    diag = torch.empty((P, ld), dtype=torch.float64, device="cuda")
    zero = np.zeros(3, dtype=np.int32)
    K = 3
    sink_ptr = [0]
    sink_idx_parts, sink_val_parts = [], []
    counts = (C.c_long * K)()
    for p, r in enumerate(order):
        nu = SM[:, r]
        rate, ordv = mass_action_desc(r)
        st.lookup_shifted(nu, -1, out=col[p], first=first, count=count)
        st.eval_mass_action(rate, ordv, nu, -1, out=off[p], first=first, count=count)
        st.eval_mass_action(rate, ordv, zero, 0, out=diag[p], first=first, count=count)
        # sink lists for x + nu (FspMatrixConstrained.cpp:170-194); only surface states qualify
        cap = count
        idx = torch.empty(cap, dtype=torch.int32, device="cuda")
        nuc = np.ascontiguousarray(nu, dtype=np.int32)
        check(lib().fspset_sink_lists(st.h, nuc.ctypes.data_as(C.POINTER(C.c_int)), first, count,
                                      C.c_void_p(idx.data_ptr()), cap, counts), "fspset_sink_lists")
        tot = sum(counts[k] for k in range(K))
        idx = idx[:tot].clone()
        val = torch.empty(tot, dtype=torch.float64, device="cuda")
        if tot:
            check(lib().fspvec_gather(C.c_void_p(val.data_ptr()), C.c_void_p(diag[p].data_ptr()),
                                      C.c_void_p(idx.data_ptr()), tot, None), "fspvec_gather")
        for k in range(K):
            sink_ptr.append(sink_ptr[-1] + counts[k])
        sink_idx_parts.append(idx)
        sink_val_parts.append(val)
    torch.cuda.synchronize()
    sink_idx = torch.cat(sink_idx_parts) if sink_idx_parts else torch.empty(0, dtype=torch.int32, device="cuda")
    sink_val = torch.cat(sink_val_parts) if sink_val_parts else torch.empty(0, dtype=torch.float64, device="cuda")
    n_ghost = 0
    ghost_info = None
    if row_range is not None:
        from .partition import remap_ghosts
        n_ghost, ghost_info = remap_ghosts(col, first, count)
    M = DeviceFspMatrix()
    M.generate(count, 6, tv_r, ti_r, col, off, diag, ld, 1, K=K, sink_ptr=sink_ptr, sink_idx=sink_idx,
               sink_val=sink_val, owns_sinks=owns, n_ghost=n_ghost)
    M.ghost_info = ghost_info
    M.N_global = N
    del col, off, diag
    torch.cuda.empty_cache()
    if return_set:
        return M, N, st
    return M, N

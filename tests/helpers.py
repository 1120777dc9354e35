"""Shared helpers: move an oracle operator into the device layout of include/fsp_b200.h."""
import numpy as np


def planes_from_oracle(A):
    """Return dict with the reaction-plane ELL arrays (TV planes first) and merged sink lists."""
    col, off, diag = A.ell()
    tv, ti = A.tv_ti()
    order = list(tv) + list(ti)
    n = A.n
    K = A.nrows - A.n
    P = len(order)
    pc = np.ascontiguousarray(col[order]) if P else np.zeros((0, n), np.int32)
    po = np.ascontiguousarray(off[order]) if P else np.zeros((0, n))
    pd = np.ascontiguousarray(diag[order]) if P else np.zeros((0, n))
    sinks = A.sinks() if K > 0 else {}
    sp = [0]
    si, sv = [], []
    for p, r in enumerate(order):
        for k in range(K):
            idx, val = sinks.get((int(r), k), (np.zeros(0, np.int32), np.zeros(0)))
            si.append(idx)
            sv.append(val)
            sp.append(sp[-1] + len(idx))
    si = np.concatenate(si) if si else np.zeros(0, np.int32)
    sv = np.concatenate(sv) if sv else np.zeros(0)
    return dict(n=n, K=K, tv=np.array(tv, np.int32), ti=np.array(ti, np.int32), col=pc, off=po, diag=pd,
                sink_ptr=np.array(sp, np.int64), sink_idx=si.astype(np.int32), sink_val=sv.astype(np.float64))


def device_matrix_from_oracle(A, R):
    from pacmensl_b200.device import DeviceFspMatrix
    d = planes_from_oracle(A)
    M = DeviceFspMatrix()
    M.generate(d["n"], R, d["tv"], d["ti"], d["col"], d["off"], d["diag"], d["n"], 0, K=d["K"],
               sink_ptr=d["sink_ptr"], sink_idx=d["sink_idx"], sink_val=d["sink_val"], owns_sinks=1)
    return M


def rel_err(y, yref, scale=None):
    y = np.asarray(y)
    yref = np.asarray(yref)
    s = np.max(np.abs(yref)) if scale is None else scale
    s = s if s > 0 else 1.0
    return float(np.max(np.abs(y - yref)) / s)

"""GPU parity: fused fp64 vector kernels against numpy (fp64 torch/numpy reference; tolerance 1e-13
relative on reductions -- summation order differs, values are O(1))."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def P(t):
    return C.c_void_p(t.data_ptr())


@pytest.fixture()
def L(cuda):
    from pacmensl_b200._capi import lib
    return lib()


@pytest.mark.parametrize("n", [1, 2, 31, 1000, 65537, 1 << 20])
def test_blas1(cuda, L, n):
    torch = cuda
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n); y = rng.standard_normal(n); w = rng.random(n) + 0.5
    xd, yd, wd = (torch.from_numpy(a).cuda() for a in (x, y, w))
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = C.c_double()
    assert L.fspvec_dot_h(C.byref(out), P(xd), P(yd), n, s) == 0
    assert abs(out.value - x @ y) <= 1e-13 * (np.abs(x) @ np.abs(y)) + 1e-300
    assert L.fspvec_norm2_h(C.byref(out), P(xd), n, s) == 0
    assert abs(out.value - np.linalg.norm(x)) <= 1e-13 * np.linalg.norm(x)
    assert L.fspvec_sum_h(C.byref(out), P(wd), n, s) == 0
    assert abs(out.value - w.sum()) <= 1e-13 * w.sum()
    assert L.fspvec_norm1_h(C.byref(out), P(xd), n, s) == 0
    assert abs(out.value - np.abs(x).sum()) <= 1e-13 * np.abs(x).sum()
    # axpy / scale / copy / set / linear_sum
    z = yd.clone()
    assert L.fspvec_axpy(P(z), 0.75, P(xd), n, s) == 0
    assert np.allclose(z.cpu().numpy(), y + 0.75 * x, rtol=1e-15, atol=1e-15)
    assert L.fspvec_scale(P(z), -2.0, n, s) == 0
    assert np.allclose(z.cpu().numpy(), -2.0 * (y + 0.75 * x), rtol=1e-15, atol=1e-15)
    assert L.fspvec_copy(P(z), P(xd), n, s) == 0 and (z == xd).all()
    assert L.fspvec_set(P(z), 3.5, n, s) == 0 and (z == 3.5).all()
    assert L.fspvec_linear_sum(P(z), 2.0, P(xd), -1.0, P(yd), n, s) == 0
    assert np.allclose(z.cpu().numpy(), 2 * x - y, rtol=1e-15, atol=1e-15)
    # weighted square sum (N_VWrmsNorm building block) and error weights
    red = torch.zeros(4, dtype=torch.float64, device="cuda")
    assert L.fspvec_wsqsum(P(red), P(xd), P(wd), n, s) == 0
    assert abs(float(red[0]) - ((x * w) ** 2).sum()) <= 1e-13 * ((x * w) ** 2).sum()
    ew = torch.empty_like(xd)
    assert L.fspvec_ewt(P(ew), P(yd), 1e-4, 1e-9, n, P(red), s) == 0
    assert np.allclose(ew.cpu().numpy(), 1.0 / (1e-4 * np.abs(y) + 1e-9), rtol=1e-15)
    assert np.isclose(float(red[0]), (1e-4 * np.abs(y) + 1e-9).min(), rtol=1e-15)


def test_maxpy_mdot_and_fused_mgs(cuda, L):
    torch = cuda
    n, m = 100003, 61
    rng = np.random.default_rng(0)
    V = rng.standard_normal((m, n))
    a = rng.standard_normal(m)
    Vd = torch.from_numpy(V).cuda()
    ptrs = (C.c_void_p * m)(*[Vd[k].data_ptr() for k in range(m)])
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    y = torch.full((n,), 7.0, dtype=torch.float64, device="cuda")
    ad = np.ascontiguousarray(a)
    assert L.fspvec_maxpy(P(y), 0.0, m, ad.ctypes.data_as(C.POINTER(C.c_double)), ptrs, n, s) == 0
    ref = a @ V
    assert np.abs(y.cpu().numpy() - ref).max() <= 1e-13 * np.abs(a) @ np.abs(V).max(axis=1)
    # mdot: 3 dots in one pass
    out = torch.zeros(8, dtype=torch.float64, device="cuda")
    p3 = (C.c_void_p * 3)(*[Vd[k].data_ptr() for k in (5, 6, 7)])
    assert L.fspvec_mdot(P(out), P(Vd[0]), 3, p3, n, s) == 0
    assert np.allclose(out[:3].cpu().numpy(), V[5:8] @ V[0], rtol=1e-12, atol=1e-10)
    # fused MGS step of the IOP loop (KrylovFsp.cpp:302-309): w -= h*v ; out = <w, u>
    w = Vd[1].clone(); h = torch.tensor([0.37], dtype=torch.float64, device="cuda")
    assert L.fspvec_axpy_dot(P(w), P(h), 1.0, P(Vd[2]), P(Vd[3]), P(out), n, s) == 0
    wref = V[1] - 0.37 * V[2]
    assert np.allclose(w.cpu().numpy(), wref, rtol=1e-15, atol=1e-15)
    assert abs(float(out[0]) - wref @ V[3]) <= 1e-12 * np.abs(wref) @ np.abs(V[3])
    assert L.fspvec_axpy_dot(P(w), P(h), 1.0, P(Vd[2]), None, P(out), n, s) == 0
    wref = wref - 0.37 * V[2]
    assert abs(float(out[0]) - wref @ wref) <= 1e-12 * (wref @ wref)
    assert L.fspvec_scale_rsqrt(P(w), P(out), n, s) == 0
    assert np.allclose(w.cpu().numpy(), wref / np.linalg.norm(wref), rtol=1e-14)


def test_scatter_gather_expandvec(cuda, L, oracle):
    torch = cuda
    rng = np.random.default_rng(2)
    n_old, n_new = 5000, 9000
    idx = rng.permutation(n_new)[:n_old].astype(np.int32)
    p = rng.random(n_old)
    pd, idd = torch.from_numpy(p).cuda(), torch.from_numpy(idx).cuda()
    pn = torch.full((n_new,), -1.0, dtype=torch.float64, device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.fspvec_scatter(P(pn), n_new, P(pd), P(idd), n_old, s) == 0
    assert (pn.cpu().numpy() == oracle.expand_vec(p, idx, n_new)).all()   # PetscWrap.cpp:26-56
    back = torch.empty(n_old, dtype=torch.float64, device="cuda")
    assert L.fspvec_gather(P(back), P(pn), P(idd), n_old, s) == 0
    assert (back.cpu().numpy() == p).all()


@pytest.mark.parametrize("n,n_ranks", [(1, 2), (5000, 2), (70001, 5), (300000, 16)])
def test_route_by_owner(cuda, L, n, n_ranks):
    # multi-GPU ExpandVec routing: entries sorted by the rank that owns their new global index, per-rank counts,
    # out-of-range indices dropped; inside a segment the original order is kept (stable sort)
    torch = cuda
    rng = np.random.default_rng(n)
    N = 4 * n + 7
    cuts = np.sort(rng.integers(0, N + 1, size=n_ranks - 1))
    starts = np.concatenate([[0], cuts, [N]]).astype(np.int64)   # some ranks may own nothing
    idx = rng.integers(-3, N + 3, size=n).astype(np.int32)
    val = rng.random(n)
    idd, vd = torch.from_numpy(idx).cuda(), torch.from_numpy(val).cuda()
    oi = torch.empty(n, dtype=torch.int32, device="cuda")
    ov = torch.empty(n, dtype=torch.float64, device="cuda")
    counts = (C.c_long * n_ranks)()
    st = (C.c_long * (n_ranks + 1))(*[int(v) for v in starts])
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.fspvec_route_by_owner(P(idd), P(vd), n, st, n_ranks, P(oi), P(ov), counts, s) == 0
    oi, ov = oi.cpu().numpy(), ov.cpu().numpy()
    off = 0
    for r in range(n_ranks):
        sel = (idx >= starts[r]) & (idx < starts[r + 1]) & (idx >= 0)
        assert counts[r] == int(sel.sum())
        assert (oi[off: off + counts[r]] == idx[sel]).all() and (ov[off: off + counts[r]] == val[sel]).all()
        off += counts[r]
    assert off == int(((idx >= 0) & (idx < N)).sum())

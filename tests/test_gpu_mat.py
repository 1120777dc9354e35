"""GPU parity: the fused sm_100a Action kernel (through the C ABI) against the CPU oracle.

Tolerance: 1e-12 relative (north_star: "SpMV agrees with the reference's MatMult to 1e-12 relative"),
measured as max|y - y_ref| / (||A||_inf-ish scale = max|y_ref|).
"""
import numpy as np
import pytest

from helpers import device_matrix_from_oracle, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _check(torch, O, A, R, name, times, variants=(0, 1, 2, 4), seed=3):
    M = device_matrix_from_oracle(A, R)
    rng = np.random.default_rng(seed)
    for t in times:
        x = rng.random(A.nrows)
        ierr, yref = A.action(t, x)
        assert ierr == 0
        coef = O.fixture_tcoef(name, t, R)[1] if name else np.ones(R)
        for v in variants:
            M.set_variant(v)
            xd = _dev(torch, x)
            yd = torch.full((A.nrows,), np.nan, dtype=torch.float64, device="cuda")
            M.action(coef, xd, yd)
            torch.cuda.synchronize()
            y = yd.cpu().numpy()
            assert np.isfinite(y).all()
            assert rel_err(y, yref) <= TOL, (name, t, v)
    return M


def test_kat_m1_m2_on_device(cuda, oracle):
    torch, O = cuda, oracle
    st = O.StateSet(fixture="random_walk_1d")
    st.expand()
    for constrained, expect in ((False, -2.0), (True, 0.0)):
        A = O.FspMatrix(constrained=constrained)
        A.generate_fixture(st, "random_walk_1d")
        M = device_matrix_from_oracle(A, 2)
        x = torch.ones(A.nrows, dtype=torch.float64, device="cuda")
        y = torch.empty_like(x)
        M.action(np.ones(2), x, y)
        assert float(y.sum().item()) == expect  # reference tests/test_mat.cpp:146,233 (ASSERT_DOUBLE_EQ)
        assert M.flops() == A.flops()


@pytest.mark.parametrize("name,bounds,times", [
    ("random_walk_1d_tv", None, [0.0, 0.1, 0.2, 1.0, 10.0]),
    ("toggle", [40, 30], [0.0]),
    ("toggle_custom", [30, 30, 200], [0.0]),
    ("repressilator", [22, 6, 6], [0.0]),
    ("repressilator_custom", [22, 4, 4, 60, 12, 60], [0.0]),
    ("hog1p", [3, 6, 6, 5, 5], [0.0, 10.0, 100.0]),
    ("transcr_reg_6d", [10, 6, 1, 2, 1, 1], [0.0, 50.0, 300.0]),
    ("birth_death_3d", [20, 17, 13], [0.0]),
    ("birth_death_3d_tv", [20, 17, 13], [0.0, 3.0]),
])
def test_action_parity_fixtures(cuda, oracle, name, bounds, times):
    torch, O = cuda, oracle
    st = O.StateSet(fixture=name, bounds=bounds)
    assert st.expand() == 0
    A = O.FspMatrix(constrained=True)
    assert A.generate_fixture(st, name) == 0
    M = _check(torch, O, A, st.R, name, times)
    assert M.flops() == A.flops()
    # dense export == oracle dense (KAT-M3/M4 analogue)
    if A.nrows <= 3000:
        coef = O.fixture_tcoef(name, times[-1], st.R)[1]
        assert np.allclose(M.dense(coef), A.dense(times[-1]), rtol=1e-14, atol=0)


def test_base_matrix_without_sinks(cuda, oracle):
    torch, O = cuda, oracle
    st = O.StateSet(fixture="toggle", bounds=[25, 25])
    st.expand()
    A = O.FspMatrix(constrained=False)
    A.generate_fixture(st, "toggle")
    _check(torch, O, A, st.R, "toggle", [0.0])


def test_enabled_subset_and_ragged_sizes(cuda, oracle):
    # enable_reactions subset (FspMatrixBase.cpp:105-117) and sizes that are not multiples of the
    # vector width / block size (tail handling of the 2- and 4-row kernels)
    torch, O = cuda, oracle
    for b in ([1, 1], [2, 1], [6, 4], [16, 15], [31, 32]):
        st = O.StateSet(fixture="toggle", bounds=b)
        st.expand()
        A = O.FspMatrix(constrained=True)
        px = lambda r, X: [np.full(len(X), 0.3), 1.0 / (1.0 + X[:, 1]), 0.1 * X[:, 0], np.full(len(X), 0.2),
                           1.0 / (1.0 + X[:, 0] ** 2), 0.7 * X[:, 1]][r]
        pt = lambda t, out: out.__setitem__(slice(None), 1.0 + t * np.arange(len(out))) or 0
        A.generate(st, px, pt, tv=[1, 4], enable=[0, 1, 4, 5])
        M = device_matrix_from_oracle(A, 6)
        rng = np.random.default_rng(1)
        x = rng.random(A.nrows)
        for t in (0.0, 2.0):
            ierr, yref = A.action(t, x)
            coef = 1.0 + t * np.arange(6)
            for v in (0, 1, 2, 4):
                M.set_variant(v)
                yd = torch.empty(A.nrows, dtype=torch.float64, device="cuda")
                M.action(coef, _dev(torch, x), yd)
                assert rel_err(yd.cpu().numpy(), yref) <= TOL


def test_unaligned_vectors_fall_back_to_scalar_path(cuda, oracle):
    torch, O = cuda, oracle
    st = O.StateSet(fixture="toggle", bounds=[20, 20])
    st.expand()
    A = O.FspMatrix(constrained=True)
    A.generate_fixture(st, "toggle")
    M = device_matrix_from_oracle(A, 6)
    x = np.random.default_rng(5).random(A.nrows)
    buf = torch.zeros(A.nrows + 1, dtype=torch.float64, device="cuda")
    buf[1:] = _dev(torch, x)
    ybuf = torch.zeros(A.nrows + 1, dtype=torch.float64, device="cuda")
    M.action(np.ones(6), buf[1:], ybuf[1:])
    _, yref = A.action(0.0, x)
    assert rel_err(ybuf[1:].cpu().numpy(), yref) <= TOL


def test_linearity_and_mass_conservation_large(cuda, oracle):
    # size-independent properties at a size the oracle is not asked to reproduce, through the host C++ classes
    # (StateSetConstrained::AddBoxLattice -> FspMatrixConstrained::GenerateValues with device propensities):
    # A(ax + by) = aAx + bAy and 1^T A x = 0 for the constrained box-lattice operator
    torch = cuda
    from pacmensl_b200.lattice import build_birth_death_lattice
    M, n = build_birth_death_lattice([63, 62, 61], tv=False)
    assert n == 64 * 63 * 62
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand(M.n_rows, generator=g, dtype=torch.float64, device="cuda")
    z = torch.rand(M.n_rows, generator=g, dtype=torch.float64, device="cuda")
    x[n:] = 0
    z[n:] = 0
    yx, yz, yc = (torch.empty_like(x) for _ in range(3))
    M.action(0.0, x, yx)
    M.action(0.0, z, yz)
    M.action(0.0, 2.0 * x - 0.5 * z, yc)
    scale = float(yx.abs().max())
    assert float((yc - (2.0 * yx - 0.5 * yz)).abs().max()) <= 1e-12 * scale
    assert abs(float(yx.sum())) <= 1e-9 * float(yx.abs().sum())


def test_host_generate_matches_oracle_lattice(cuda, oracle):
    # the device-generated lattice operator (hash lookups + mass-action kernels) == oracle built from host callbacks
    torch, O = cuda, oracle
    from pacmensl_b200.lattice import build_birth_death_lattice
    for tv, name in ((False, "birth_death_3d"), (True, "birth_death_3d_tv")):
        M, n = build_birth_death_lattice([12, 9, 7], tv=tv)
        st = O.StateSet(fixture=name, bounds=[12, 9, 7])
        st.expand()
        A = O.FspMatrix(constrained=True)
        A.generate_fixture(st, name)
        assert M.flops == A.flops()
        # orderings differ (lexicographic lattice vs BFS): compare through the state key
        X_lat = M.set.states()
        perm = st.state2index(X_lat)          # oracle index of each lattice state
        assert (perm >= 0).all()
        rng = np.random.default_rng(4)
        x_lat = rng.random(M.n_rows)
        x_or = np.zeros(A.nrows)
        x_or[perm] = x_lat[:n]
        x_or[st.n:] = x_lat[n:]
        for t in (0.0, 7.0):
            ierr, y_or = A.action(t, x_or)
            yd = torch.empty(M.n_rows, dtype=torch.float64, device="cuda")
            M.action(t, _dev(torch, x_lat), yd)
            y = yd.cpu().numpy()
            assert rel_err(y[:n], y_or[perm], scale=np.abs(y_or).max()) <= TOL
            assert rel_err(y[n:], y_or[st.n:], scale=np.abs(y_or).max()) <= TOL


def test_zero_operator_before_generate(cuda):
    torch = cuda
    from pacmensl_b200.device import DeviceFspMatrix
    M = DeviceFspMatrix()
    M.R = 2
    y = torch.ones(5, dtype=torch.float64, device="cuda")
    M.action(np.ones(2), torch.ones(5, dtype=torch.float64, device="cuda"), y)  # no values: y untouched by kernel
    assert M.flops() == 0


@pytest.mark.parametrize("name,bounds,t", [("random_walk_1d_tv", None, 1.0), ("toggle_custom", [30, 30, 200], 0.0),
                                           ("hog1p", [3, 6, 6, 5, 5], 25.0), ("birth_death_3d_tv", [21, 17, 13], 4.0)])
def test_fused_epilogue_matches_separate_passes(cuda, oracle, name, bounds, t):
    """fspmat_action_fused: y = scale .* (beta x + alpha A x) and <y, v0>, <y, y> in one kernel (the BDF/GMRES and
    Krylov forms), against the oracle Action followed by numpy; sink rows included."""
    torch, O = cuda, oracle
    st = O.StateSet(fixture=name, bounds=bounds) if bounds is not None else O.StateSet(fixture=name)
    st.expand()
    A = O.FspMatrix(constrained=True)
    assert A.generate_fixture(st, name) == 0
    M = device_matrix_from_oracle(A, st.R)
    rng = np.random.default_rng(11)
    x, v0 = rng.random(A.nrows), rng.random(A.nrows)
    scale = 1.0 / (1e-6 * rng.random(A.nrows) + 1e-3)
    ierr, ax = A.action(t, x)
    coef = O.fixture_tcoef(name, t, st.R)[1]
    xd, vd, sd = _dev(torch, x), _dev(torch, v0), _dev(torch, scale)
    for alpha, beta, sc, vecs in ((-0.37, 1.0, sd, (vd, None)), (1.0, 0.0, None, (vd,)), (1.0, 0.0, None, ())):
        yd = torch.full((A.nrows,), np.nan, dtype=torch.float64, device="cuda")
        out = torch.full((2,), np.nan, dtype=torch.float64, device="cuda")
        M.action_fused(coef, xd, yd, alpha=alpha, beta=beta, scale=sc, dot_vecs=vecs, dot_out=out)
        torch.cuda.synchronize()
        yref = (beta * x + alpha * ax) * (scale if sc is not None else 1.0)
        y = yd.cpu().numpy()
        assert rel_err(y, yref) <= TOL
        o = out.cpu().numpy()
        if len(vecs) >= 1:
            assert abs(o[0] - yref @ v0) <= 1e-12 * np.abs(yref * v0).sum()
        if len(vecs) == 2:
            assert abs(o[1] - yref @ yref) <= 1e-12 * (yref @ yref)
    # repeated calls reuse the partial buffers / counter correctly and are deterministic
    outs = []
    for _ in range(3):
        out = torch.zeros(2, dtype=torch.float64, device="cuda")
        M.action_fused(coef, xd, yd, alpha=-0.37, beta=1.0, scale=sd, dot_vecs=(vd, None), dot_out=out)
        outs.append(out.cpu().numpy().copy())
    assert (outs[0] == outs[1]).all() and (outs[1] == outs[2]).all()


@pytest.mark.parametrize("bounds", [[3, 10, 10, 10, 10], [3, 17, 36, 13, 22]])
def test_fused_epilogue_krylov_form_hog1p_sizes(cuda, oracle, bounds):
    """The Krylov form (alpha = 1, beta = 0, one dot) on the hog1p sets of examples/hog1p.cpp: the initial 58 564-state
    set and the 857 808-state set of the last expansion (more rows than 8 CTAs/SM x 256: the grid-stride path), with a
    delta vector (first Krylov step) and a random one, t = 0 (coefficient of the time-varying reaction: 3200)."""
    torch, O = cuda, oracle
    st = O.StateSet(fixture="hog1p", bounds=bounds)
    st.expand()
    A = O.FspMatrix(constrained=True)
    assert A.generate_fixture(st, "hog1p") == 0
    M = device_matrix_from_oracle(A, st.R)
    coef = O.fixture_tcoef("hog1p", 0.0, st.R)[1]
    rng = np.random.default_rng(5)
    delta = np.zeros(A.nrows)
    delta[0] = 1.0
    for x in (delta, rng.random(A.nrows)):
        v0 = rng.random(A.nrows)
        ierr, ax = A.action(0.0, x)
        xd, vd = _dev(torch, x), _dev(torch, v0)
        yd = torch.full((A.nrows,), np.nan, dtype=torch.float64, device="cuda")
        out = torch.full((2,), np.nan, dtype=torch.float64, device="cuda")
        M.action_fused(coef, xd, yd, alpha=1.0, beta=0.0, scale=None, dot_vecs=(vd,), dot_out=out)
        yp = torch.full((A.nrows,), np.nan, dtype=torch.float64, device="cuda")
        M.action(coef, xd, yp)
        torch.cuda.synchronize()
        y = yd.cpu().numpy()
        assert np.isfinite(y).all()
        assert rel_err(y, ax) <= TOL
        assert rel_err(y, yp.cpu().numpy()) <= 1e-15
        o = out.cpu().numpy()
        assert abs(o[0] - ax @ v0) <= 1e-12 * max(np.abs(ax * v0).sum(), 1e-300)


@pytest.mark.parametrize("upper,tv", [([127, 127, 63], False), ([127, 95, 63], True), ([21, 17, 13], False)])
def test_action_host_pipeline_equals_device_action(cuda, upper, tv):
    """pfsp_mat_action_host (host vectors; chunked upload / compute / download pipeline above 524 288 rows, plain
    H2D + Action + D2H below) returns exactly what Action returns on device vectors -- same arithmetic per row."""
    torch = cuda
    from pacmensl_b200 import api
    from pacmensl_b200.lattice import Lattice
    api.init(0)
    lat = Lattice(upper, tv=tv)
    n = lat.n_rows
    rng = np.random.default_rng(21)
    x = rng.random(n)
    for pinned in (True, False):
        xh = torch.from_numpy(x.copy())
        yh = torch.full((n,), float("nan"), dtype=torch.float64)
        if pinned:
            xh, yh = xh.pin_memory(), yh.pin_memory()
        for t in (0.0, 7.0):
            yh.fill_(float("nan"))
            assert lat.mat.action_host(t, xh, yh) == 0
            xd = torch.from_numpy(x).cuda()
            yd = torch.empty_like(xd)
            lat.action(t, xd, yd)
            torch.cuda.synchronize()
            assert torch.equal(yh, yd.cpu()), (upper, tv, pinned, t)


@pytest.mark.gpu
@pytest.mark.parametrize("name,bounds,t", [("hog1p", [3, 6, 6, 5, 5], 25.0), ("transcr_reg_6d", [10, 6, 1, 2, 1, 1], 40.0),
                                            ("pure_birth", [30], 0.0)])
def test_device_propensity_form_is_bit_identical_to_host_callbacks(cuda, name, bounds, t):
    """The separable (rate x falling factorials x per-species factor tables) description of a model's propensities,
    evaluated on the GPU at generate time, must give exactly the operator the host callback prop_x gives."""
    import numpy as np
    from pacmensl_b200 import api
    torch = cuda
    api.init(0)
    m_dev = api.Model(fixture=name, device_form=True)
    assert m_dev.device_form
    st_h, A_h = api.fixture_set_and_matrix(name, bounds=np.asarray(bounds, dtype=np.int32), model=api.Model(fixture=name))
    st_d, A_d = api.fixture_set_and_matrix(name, bounds=np.asarray(bounds, dtype=np.int32), model=m_dev)
    assert st_h.n_global == st_d.n_global and A_h.n_rows == A_d.n_rows and A_h.info()[1] == A_d.info()[1]
    x = torch.rand(A_h.n_rows, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    y_h, y_d = torch.empty_like(x), torch.empty_like(x)
    assert A_h.action(t, x, y_h) == 0 and A_d.action(t, x, y_d) == 0
    torch.cuda.synchronize()
    assert torch.equal(y_h, y_d)


@pytest.mark.gpu
@pytest.mark.parametrize("name,bounds", [("hog1p", [3, 9, 9, 8, 8]), ("birth_death_3d_tv", [30, 20, 25]), ("toggle_custom", [40, 40, 60])])
def test_csr_jacobian_export_equals_action(cuda, oracle, name, bounds):
    """The assembled Jacobian (CreateRHSJacobian / ComputeRHSJacobian, reference src/Matrix/FspMatrixBase.cpp:308-427,
    FspMatrixConstrained.cpp:304-445) as a device CSR matrix: J(t) x == Action(t, x) to rounding (KAT-M5's property) on
    sets well beyond the 20 000-row limit of round 1's dense stand-in, with time-varying coefficients refreshed in place."""
    import ctypes as C
    import numpy as np
    from helpers import device_matrix_from_oracle, rel_err
    from pacmensl_b200 import _capi
    torch = cuda
    O = oracle
    L = _capi.lib()
    so = O.StateSet(fixture=name, bounds=bounds)
    assert so.expand() == 0
    A = O.FspMatrix(constrained=True)
    assert A.generate_fixture(so, name) == 0
    M = device_matrix_from_oracle(A, so.R)
    nnz, nr = C.c_long(), C.c_int()
    assert L.fspmat_csr_size(M.h, C.byref(nnz), C.byref(nr)) == 0
    assert nr.value == A.nrows and nnz.value > 0
    row_ptr = torch.empty(nr.value + 1, dtype=torch.int32, device="cuda")
    col = torch.empty(nnz.value, dtype=torch.int32, device="cuda")
    val = torch.empty(nnz.value, dtype=torch.float64, device="cuda")
    x = torch.rand(nr.value, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(11))
    y_csr, y_act = torch.empty_like(x), torch.empty_like(x)
    vp = C.c_void_p
    for k, t in enumerate((0.0, 0.7, 30.0)):
        coef = np.ascontiguousarray(O.fixture_tcoef(name, t, so.R)[1])
        assert L.fspmat_csr_export(M.h, coef.ctypes.data_as(C.POINTER(C.c_double)), 1 if k == 0 else 0, vp(row_ptr.data_ptr()),
                                   vp(col.data_ptr()), vp(val.data_ptr()), None) == 0
        assert L.fspmat_csr_spmv(nr.value, vp(row_ptr.data_ptr()), vp(col.data_ptr()), vp(val.data_ptr()), vp(x.data_ptr()),
                                 vp(y_csr.data_ptr()), None) == 0
        M.action(coef, x, y_act)
        torch.cuda.synchronize()
        assert int(row_ptr[-1]) == nnz.value and int(col.min()) >= 0 and int(col.max()) < nr.value
        assert rel_err(y_csr.cpu().numpy(), y_act.cpu().numpy()) <= 1e-14
        ierr, y_or = A.action(t, x.cpu().numpy())
        assert rel_err(y_csr.cpu().numpy(), y_or) <= 1e-12

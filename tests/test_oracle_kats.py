"""Pins the CPU oracle against the reference's own analytic known-answer tests (no GPU needed).

KAT ids follow SURVEY.md section 4.
"""
import numpy as np
import pytest


def _random_walk(O, tv=False):
    st = O.StateSet(fixture="random_walk_1d")
    assert st.expand() == 0
    return st


def test_kat_s1_state_count_and_indices(oracle):
    # reference tests/test_fss.cpp:70-125 : 2 species, x0 + x1 <= 3 -> 10 states, all indices >= 0
    O = oracle
    st = O.StateSet(SM=np.array([[1, -1, 0, 0], [0, 0, 1, -1]]))

    def lhs(X, out):
        out[:, 0] = X[:, 0] + X[:, 1]
        return 0

    assert st.set_shape([3], lhs) == 0
    assert st.add_states([[0, 0]]) == 0
    assert st.expand() == 0
    assert st.n == 10
    allst = [(i, j) for i in range(4) for j in range(4 - i)]
    idx = st.state2index(np.array(allst))
    assert (idx >= 0).all() and len(set(idx.tolist())) == 10
    # absent / negative states map to -1 (StateSetBase.cpp:309-343)
    assert st.state2index(np.array([[4, 0], [-1, 0], [2, 2]])).tolist() == [-1, -1, -1]


def test_kat_s2_wrong_species_count(oracle):
    # reference tests/test_fss.cpp:39-65
    st = oracle.StateSet(SM=np.array([[1, -1, 0, 0], [0, 0, 1, -1]]))
    assert st.add_states([[0, 0, 0]]) == -1


def test_kat_m1_base_matrix_sum(oracle):
    # reference tests/test_mat.cpp:110-151 : sum(A * 1) == -rate_right exactly
    O = oracle
    st = _random_walk(O)
    assert st.n == 13
    A = O.FspMatrix(constrained=False)
    assert A.generate_fixture(st, "random_walk_1d") == 0
    ierr, y = A.action(0.0, np.ones(A.nrows))
    assert ierr == 0 and y.sum() == -2.0


def test_kat_m2_constrained_matrix_sum(oracle):
    # reference tests/test_mat.cpp:199-238 : with the sink row, columns sum to zero
    O = oracle
    st = _random_walk(O)
    A = O.FspMatrix(constrained=True)
    assert A.generate_fixture(st, "random_walk_1d") == 0
    assert A.nrows == 14
    ierr, y = A.action(0.0, np.ones(A.nrows))
    assert ierr == 0 and y.sum() == 0.0
    assert y[-1] == 2.0  # the state x=12 leaks with rate 2 into the sink


@pytest.mark.parametrize("t", [0.0, 0.1, 0.2, 1.0, 10.0])
def test_kat_m5_action_equals_assembled_jacobian(oracle, t):
    # reference tests/test_mat.cpp:289-341 (TV reactions {0,1}, c(t) = (1+t, 1+0.5t))
    O = oracle
    st = _random_walk(O)
    A = O.FspMatrix(constrained=True)
    assert A.generate_fixture(st, "random_walk_1d_tv") == 0
    rng = np.random.default_rng(7)
    x = rng.random(A.nrows)
    ierr, y = A.action(t, x)
    J = A.dense(t)
    assert ierr == 0
    assert np.linalg.norm(J @ x - y) <= 1e-14 * max(1.0, np.linalg.norm(y))
    # the fused single-pass CPU variant agrees too
    ierr, yf = A.action(t, x, fused=True)
    assert np.allclose(yf, y, rtol=1e-13, atol=1e-13)


def test_flops_formula(oracle):
    # GetLocalMVFlops: FspMatrixBase.cpp:429-444 + FspMatrixConstrained.cpp:447-465
    O = oracle
    st = _random_walk(O)
    A = O.FspMatrix(constrained=True)
    A.generate_fixture(st, "random_walk_1d")
    # merged TI matrix: 13 diagonal + 12 sub + 12 super entries = 37 nnz; sink matrix 1 nnz
    assert A.flops() == 2 * 37 + 2 * 1
    B = O.FspMatrix(constrained=True)
    B.generate_fixture(st, "random_walk_1d_tv")
    # two TV matrices with 13+12 nnz each (+14 rows each), sinks: reaction 0 has 1 nnz (+K=1), reaction 1: 0 (+1)
    assert B.flops() == 2 * (2 * 25 + 14) + (2 * 1 + 1) + (0 + 1)


def test_column_sums_of_constrained_operator(oracle):
    # probability conservation: with box constraints every leaving transition violates exactly one
    # constraint, so 1^T A = 0
    O = oracle
    st = O.StateSet(fixture="toggle", bounds=[15, 12])
    st.expand()
    assert st.n == 16 * 13
    A = O.FspMatrix(constrained=True)
    A.generate_fixture(st, "toggle")
    J = A.dense(0.0)
    assert np.abs(J.sum(axis=0)[: st.n]).max() < 1e-15
    assert np.abs(J[:, st.n:]).max() == 0.0  # sink columns are empty


def test_expand_is_first_discovery_order(oracle):
    # index map = insertion order; new states of a wave in reaction-major first-occurrence order
    O = oracle
    st = O.StateSet(SM=np.array([[1, -1, 0, 0], [0, 0, 1, -1]]))
    st.set_shape([2, 2])
    st.add_states([[0, 0]])
    st.expand()
    assert st.states().tolist() == [[0, 0], [1, 0], [0, 1], [2, 0], [1, 1], [0, 2], [2, 1], [1, 2], [2, 2]]
    # growing the bounds re-activates blocked states and appends, leaving old indices unchanged
    old = st.states().copy()
    st.set_bounds([3, 2])
    st.expand()
    assert st.n == 12 and (st.states()[:9] == old).all()


def test_t_fun_error_propagates(oracle):
    # FspMatrixBase.cpp:44-45 -> Action returns the callback's code
    O = oracle
    st = _random_walk(O)
    A = O.FspMatrix(constrained=True)
    A.generate(st, lambda r, X: np.full(len(X), 2.0) if r == 0 else 3.0 * (X[:, 0] > 0),
               prop_t=lambda t, out: -1, tv=[0, 1])
    ierr, y = A.action(0.0, np.ones(A.nrows))
    assert ierr == -1


def test_expand_vec(oracle):
    p = oracle.expand_vec([0.1, 0.2, 0.3], [2, 0, 4], 6)
    assert p.tolist() == [0.2, 0.0, 0.1, 0.0, 0.3, 0.0]


@pytest.mark.parametrize("tv", [False, True])
@pytest.mark.parametrize("dims", [[7, 5, 4], [1, 6, 3], [9, 9, 9]])
def test_direct_lattice_generator_matches_generic(oracle, tv, dims):
    # the CPU-baseline operator of the full 465^3 lattice is built directly in lexicographic order
    # (orc_mat_generate_lattice); it must be the operator orc_mat_generate produces on the BFS-ordered set, by state key
    O = oracle
    name = "birth_death_3d_tv" if tv else "birth_death_3d"
    st = O.StateSet(fixture=name, bounds=[d - 1 for d in dims])
    st.expand()
    A = O.FspMatrix(constrained=True)
    assert A.generate_fixture(st, name) == 0
    B = O.FspMatrix(constrained=True)
    assert B.generate_lattice(dims, tv) == 0
    N = st.n
    assert N == int(np.prod(dims)) and B.nrows == A.nrows == N + 3 and A.flops() == B.flops()
    idx = np.arange(N)
    X = np.stack([idx % dims[0], (idx // dims[0]) % dims[1], idx // (dims[0] * dims[1])], axis=1).astype(np.int32)
    perm = st.state2index(X)
    assert (perm >= 0).all()
    xl = np.random.default_rng(5).random(N + 3)
    xo = np.zeros(N + 3)
    xo[perm] = xl[:N]
    xo[N:] = xl[N:]
    for t in (0.0, 2.5):
        _, ya = A.action(t, xo)
        _, yb = B.action(t, xl)
        # same entries; the summation order inside a row follows the column order of each ordering (last-bit effects)
        np.testing.assert_allclose(yb[:N], ya[perm], rtol=1e-13, atol=1e-13 * np.abs(ya).max())
        np.testing.assert_allclose(yb[N:], ya[N:], rtol=1e-13)

"""GPU parity: device state set (hash directory + Expand) against the CPU oracle -- bit-exact index map."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SM_TOGGLE = np.array([[1, 1, -1, 0, 0, 0], [0, 0, 0, 1, 1, -1]])


def _pair(O, SM, bounds, x0, lhs=None):
    from pacmensl_b200.device import DeviceStateSet
    so = O.StateSet(SM=SM)
    sd = DeviceStateSet(SM)
    assert so.set_shape(bounds, lhs) == 0
    assert sd.set_shape(bounds, lhs) == 0
    assert so.add_states(x0) == 0
    assert sd.add_states(x0) == 0
    return so, sd


def test_kat_s1_on_device(cuda, oracle):
    SM = np.array([[1, -1, 0, 0], [0, 0, 1, -1]])

    def lhs(X, out):
        out[:, 0] = X[:, 0] + X[:, 1]
        return 0

    so, sd = _pair(oracle, SM, [3], [[0, 0]], lhs)
    assert so.expand() == 0 and sd.expand() == 0
    assert sd.n == 10
    assert (sd.states() == so.states()).all()
    allst = np.array([(i, j) for i in range(4) for j in range(4 - i)])
    assert (sd.state2index(allst) == so.state2index(allst)).all()
    assert sd.state2index(np.array([[4, 0], [-1, 0], [2, 2]])).tolist() == [-1, -1, -1]


def test_kat_s2_wrong_species(cuda):
    from pacmensl_b200.device import DeviceStateSet
    sd = DeviceStateSet(np.array([[1, -1, 0, 0], [0, 0, 1, -1]]))
    assert sd.add_states([[0, 0, 0]]) == -1


@pytest.mark.parametrize("bounds", [[1, 1], [5, 3], [40, 33], [200, 150]])
def test_expand_box_bit_exact(cuda, oracle, bounds):
    so, sd = _pair(oracle, SM_TOGGLE, bounds, [[0, 0]])
    assert so.expand() == 0 and sd.expand() == 0
    assert sd.n == so.n == (bounds[0] + 1) * (bounds[1] + 1)
    assert (sd.states() == so.states()).all()           # identical insertion order
    assert (sd.status() == so.status()).all()
    # repeated expansion with growing bounds (FspSolverMultiSinks.cpp:116-143)
    for grow in ([bounds[0] + 3, bounds[1]], [bounds[0] + 3, bounds[1] + 7]):
        so.set_bounds(grow)
        sd.set_bounds(grow)
        assert so.expand() == 0 and sd.expand() == 0
        assert (sd.states() == so.states()).all()
        assert (sd.status() == so.status()).all()


def test_expand_custom_constraints_bit_exact(cuda, oracle):
    def lhs(X, out):
        out[:, 0] = X[:, 0]
        out[:, 1] = X[:, 1]
        out[:, 2] = X[:, 0] * X[:, 1]
        return 0

    so, sd = _pair(oracle, SM_TOGGLE, [30, 30, 120], [[0, 0], [3, 2]], lhs)
    assert so.expand() == 0 and sd.expand() == 0
    assert sd.n == so.n
    assert (sd.states() == so.states()).all()
    assert (sd.status() == so.status()).all()


def test_expand_6d_and_lookup(cuda, oracle):
    from pacmensl_b200.device import DeviceStateSet
    st = oracle.StateSet(fixture="transcr_reg_6d")
    st.expand()
    SM = np.array([[1, -1, 0, 0, 0, 0, 0, 0, -2, 2], [0, 0, 0, 0, -1, 1, -1, 1, 1, -1], [0, 0, 0, 0, -1, 1, 0, 0, 0, 0],
                   [0, 0, 0, 0, 1, -1, -1, 1, 0, 0], [0, 0, 0, 0, 0, 0, 1, -1, 0, 0], [0, 0, 1, -1, 0, 0, 0, 0, 0, 0]])
    sd = DeviceStateSet(SM)
    sd.set_shape([10, 6, 1, 2, 1, 1])
    sd.add_states([[2, 6, 0, 2, 0, 0]])
    assert sd.expand() == 0
    assert sd.n == st.n
    assert (sd.states() == st.states()).all()
    # State2Index on shifted states == oracle (what GenerateValues needs, FspMatrixBase.cpp:133-134)
    X = st.states()
    for r in range(10):
        shifted = X - SM[:, r]
        ref = st.state2index(shifted)
        got = sd.lookup_shifted(SM[:, r], -1).cpu().numpy()
        assert (got == ref).all()
        assert (sd.state2index(shifted) == ref).all()


def test_add_states_sheds_present_and_duplicates(cuda, oracle):
    so, sd = _pair(oracle, SM_TOGGLE, [9, 9], [[0, 0]])
    X = [[1, 1], [0, 0], [1, 1], [2, 5], [2, 5], [7, 7]]
    assert so.add_states(X) == 0 and sd.add_states(X) == 0
    assert sd.n == so.n == 4
    assert (sd.states() == so.states()).all()


def test_box_lattice_is_lexicographic_and_closed(cuda, oracle):
    from pacmensl_b200.device import DeviceStateSet
    SM = np.array([[1, -1, 0, 0, 0, 0], [0, 0, 1, -1, 0, 0], [0, 0, 0, 0, 1, -1]])
    sd = DeviceStateSet(SM)
    sd.set_shape([6, 4, 5])
    sd.add_box_lattice([6, 4, 5])
    assert sd.n == 7 * 5 * 6
    X = sd.states()
    idx = X[:, 0] + 7 * (X[:, 1] + 5 * X[:, 2])   # sub2ind_nd convention, Sys/pacmenMath.h:33-59
    assert (idx == np.arange(sd.n)).all()
    assert sd.expand() == 0 and sd.n == 7 * 5 * 6   # already closed under the reactions
    # statuses: interior states done (0), states with an invalid child blocked (-1)
    st = sd.status()
    blocked = (X[:, 0] == 0) | (X[:, 1] == 0) | (X[:, 2] == 0) | (X[:, 0] == 6) | (X[:, 1] == 4) | (X[:, 2] == 5)
    assert ((st == -1) == blocked).all() and ((st == 0) == ~blocked).all()


def test_check_constraints_shifted(cuda, oracle):
    so, sd = _pair(oracle, SM_TOGGLE, [7, 5], [[0, 0]])
    so.expand(); sd.expand()
    X = so.states()
    for r in range(6):
        ref = so.check_constraints(X + SM_TOGGLE[:, r])
        got = sd.check_constraints_shifted(SM_TOGGLE[:, r]).cpu().numpy()
        assert (got == ref).all()

"""Pins the CPU restatement of KrylovFsp (oracle/krylov_oracle.py) with the analytic answers the reference's own solver
tests use: Poisson pmfs (tests/test_fsp_solver.cpp:264-345) and mass conservation (tests/test_ode.cpp:220-259) [CPU];
then the CUDA KrylovFsp (through the FSP driver on a fixed state set) against that restatement [GPU]."""
import math

import numpy as np
import pytest

BOUNDS = [63, 47, 39]          # 64 x 48 x 40 = 122 880 states: truncation error of the box << 1e-9 at t = 1
B, G = (40.0, 30.0, 20.0), (1.0, 1.5, 2.0)


def poisson_product(states, t):
    out = np.ones(len(states))
    for s in range(3):
        lam = B[s] / G[s] * (1.0 - math.exp(-G[s] * t))
        k = states[:, s].astype(np.float64)
        out *= np.exp(-lam + k * math.log(lam) - np.array([math.lgamma(v + 1.0) for v in k]))
    return out


def _oracle_solve(O, t_final=1.0):
    from oracle.krylov_oracle import KrylovOracle
    st = O.StateSet(fixture="birth_death_3d", bounds=BOUNDS)
    st.expand()
    A = O.FspMatrix(constrained=True)
    assert A.generate_fixture(st, "birth_death_3d") == 0
    p0 = np.zeros(A.nrows)
    p0[st.state2index(np.array([[0, 0, 0]], dtype=np.int32))[0]] = 1.0
    kry = KrylovOracle(A)
    p = kry.solve(p0, t_final)
    return st, A, kry, p


def test_krylov_oracle_matches_poisson_product(oracle):
    st, A, kry, p = _oracle_solve(oracle)
    n = st.n
    exact = poisson_product(st.states(), 1.0)
    assert abs(p.sum() - 1.0) <= 1e-12            # mass conservation incl. sinks (KAT-O2 asks 1e-8)
    assert np.abs(p[:n] - exact).sum() <= 1e-9    # KAT-F5 asks 1e-6; the box truncation is ~1e-10
    assert kry.num_rhs > 100 and kry.num_steps >= 2


def test_krylov_oracle_toggle_mass_conservation(oracle):
    from oracle.krylov_oracle import KrylovOracle
    O = oracle
    st = O.StateSet(fixture="toggle", bounds=[100, 100])
    st.expand()
    assert st.n == 10201                          # reference tests/test_ode.cpp: 101 x 101 box
    A = O.FspMatrix(constrained=True)
    A.generate_fixture(st, "toggle")
    p0 = np.zeros(A.nrows)
    p0[0] = 1.0
    p = KrylovOracle(A).solve(p0, 100.0)
    assert abs(p.sum() - 1.0) <= 1e-8 and p.min() > -1e-10


@pytest.mark.gpu
def test_device_krylov_matches_oracle_krylov(cuda, oracle):
    """north_star: the probability vector at t_f within 10 x the solver's absolute tolerance in 1-norm.  KrylovFsp only
    uses atol (1e-14, reference KrylovFsp.cpp:182); on top of 10 x atol the bound allows the roundoff floor of two
    different summation orders over 1e5-dimensional vectors and ~400 operator applications (1e-12)."""
    from pacmensl_b200 import api
    api.init(0)
    st, A, kry, p_or = _oracle_solve(oracle)
    s, m = api.fixture_solver("birth_death_3d", api.KRYLOV)
    s.set_initial_bounds(BOUNDS)
    states, p = s.solve(1.0, -1.0)                # fsp_tol <= 0: fixed state set, no expansion
    assert len(states) == st.n
    idx = st.state2index(states)
    assert (idx >= 0).all()
    diff = np.abs(p - p_or[idx]).sum()
    exact = poisson_product(states, 1.0)
    print("||p_gpu - p_oracle||_1 = %.3e, ||p_gpu - exact||_1 = %.3e" % (diff, np.abs(p - exact).sum()))
    assert diff <= 10 * 1e-14 + 1e-12
    assert np.abs(p - exact).sum() <= 1e-9
    s.clear()

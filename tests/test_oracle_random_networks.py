"""CPU: the oracle (oracle/fsp_oracle.c) against the independent pure-Python restatement of the reference semantics
(tests/golden/make_golden.py: build / action) on RANDOM reaction networks -- random stoichiometry with repeated and
opposite reactions, polynomial propensities, linear custom constraints that overlap (a destination may violate several:
each violated constraint gets the entry, FspMatrixConstrained.cpp:177-193), time-varying subsets, enabled subsets are
exercised elsewhere.  Vectors are compared by STATE, so the comparison is independent of the index order."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


def random_network(seed):
    rng = np.random.default_rng(seed)
    S = int(rng.integers(1, 4))
    R = int(rng.integers(2, 7))
    SM = rng.integers(-2, 3, size=(R, S))
    for r in range(R):
        if not SM[r].any():
            SM[r, rng.integers(0, S)] = 1
    if R >= 3:
        SM[R - 1] = -SM[0]          # an exactly opposite pair
    if R >= 5:
        SM[R - 2] = SM[1]           # two reactions with the same stoichiometry simply add
    rate = rng.random(R) * 3.0 + 0.1
    order = rng.integers(0, 3, size=(R, S))
    K = int(rng.integers(1, 4))
    W = rng.integers(0, 3, size=(K, S))
    for k in range(K):
        if not W[k].any():
            W[k, rng.integers(0, S)] = 1
    bounds = [int(v) for v in rng.integers(4, 9, size=K) * W.sum(axis=1).clip(1)]
    # every species must be bounded by some constraint, otherwise the BFS closure is infinite
    for s in range(S):
        if not W[:, s].any():
            W[int(rng.integers(0, K)), s] = 1
    tv = sorted(set(int(v) for v in rng.choice(R, size=int(rng.integers(0, R)), replace=False)))
    amp = rng.random(R)

    def prop(r, x):
        v = rate[r]
        for s in range(S):
            v = v * float(x[s]) ** int(order[r, s])
        return v

    def tfun(t):
        return [1.0 + amp[r] * np.sin(0.3 * t + r) for r in range(R)]

    def lhs(x):
        return [int(sum(W[k, s] * x[s] for s in range(S))) for k in range(K)]

    w = dict(SM=[list(map(int, SM[r])) for r in range(R)], prop=prop, tfun=tfun, tv=tv, lhs=lhs, bounds=bounds,
             x0=[0] * S, times=[0.0, 2.5])
    return w, S, R, K, W, rate, order, amp


@pytest.mark.parametrize("seed", range(12))
def test_oracle_equals_independent_restatement_on_random_network(oracle, seed):
    O = oracle
    w, S, R, K, W, rate, order, amp = random_network(seed)
    states, K_, lhs, bounds = mg.build(w)
    assert 1 <= len(states) <= 20000
    st = O.StateSet(SM=np.array(w["SM"]).T)

    def lhs_cb(X, out):
        out[:, :] = X @ W.T
        return 0

    assert st.set_shape(bounds, lhs_cb) == 0
    assert st.add_states([w["x0"]]) == 0
    assert st.expand() == 0
    assert st.n == len(states)
    idx = st.state2index(np.array(states, dtype=np.int32))
    assert (idx >= 0).all() and len(set(idx.tolist())) == st.n

    def prop_x(r, X):
        v = np.full(len(X), rate[r])
        for s in range(S):
            v = v * X[:, s].astype(np.float64) ** int(order[r, s])
        return v

    def prop_t(t, out):
        out[:] = [1.0 + amp[r] * np.sin(0.3 * t + r) for r in range(R)]
        return 0

    A = O.FspMatrix(constrained=True)
    assert A.generate(st, prop_x, prop_t=prop_t, tv=w["tv"]) == 0
    rng = np.random.default_rng(100 + seed)
    xs = {s: float(rng.random()) for s in states}
    x = np.zeros(st.n + K)
    x[idx] = [xs[s] for s in states]
    x[st.n:] = rng.random(K)
    for t in w["times"]:
        y_ref, ysink_ref = mg.action(w, states, K, lhs, bounds, t, xs, list(x[st.n:]))
        ierr, y = A.action(t, x)
        assert ierr == 0
        ref = np.array([y_ref[s] for s in states])
        scale = max(np.abs(ref).max(), np.abs(ysink_ref).max() if K else 0.0, 1e-300)
        assert np.abs(y[idx] - ref).max() <= 1e-12 * scale
        assert np.abs(y[st.n:] - np.array(ysink_ref)).max() <= 1e-12 * scale
        # the fused single-pass CPU variant agrees too
        ierr, yf = A.action(t, x, fused=True)
        assert np.abs(yf - y).max() <= 1e-12 * scale


def reference_wave_order(w):
    """Index map of the reference at np = 1: per wave, children Y[:, j*nF + i] = frontier_i + nu_j
    (reaction-major, StateSetConstrained.cpp:175-179) -> keep the valid ones -> first occurrence of every duplicate ->
    drop those already present -> append in that order with status 1 (StateSetBase.cpp:207-241)."""
    lhs, bounds = w["lhs"], w["bounds"]

    def valid(x):
        return all(v >= 0 for v in x) and all(l <= b for l, b in zip(lhs(x), bounds))

    states = [tuple(w["x0"])]
    present = {states[0]}
    frontier = list(states)
    while frontier:
        children = [tuple(a + b for a, b in zip(x, nu)) for nu in w["SM"] for x in frontier]
        new = []
        for y in children:
            if valid(y) and y not in present:
                present.add(y)
                new.append(y)
        states += new
        frontier = new
    return states


@pytest.mark.parametrize("seed", range(12))
def test_oracle_index_map_is_the_reaction_major_first_discovery_order(oracle, seed):
    w, S, R, K, W, rate, order, amp = random_network(seed)
    st = oracle.StateSet(SM=np.array(w["SM"]).T)

    def lhs_cb(X, out):
        out[:, :] = X @ W.T
        return 0

    assert st.set_shape(w["bounds"], lhs_cb) == 0
    assert st.add_states([w["x0"]]) == 0
    assert st.expand() == 0
    expect = reference_wave_order(w)
    assert st.states().tolist() == [list(s) for s in expect]
    # growing a bound re-activates the blocked states and appends; old indices never change (np = 1)
    bigger = [b + 2 for b in w["bounds"]]
    assert st.set_bounds(bigger) == 0 and st.expand() == 0
    assert st.states()[: len(expect)].tolist() == [list(s) for s in expect]

#!/usr/bin/env python
"""Generates tests/golden/*.json: small known-answer vectors for the FSP operator.

The reference (voduchuy/pacmensl) stores NO golden vectors and cannot be built or imported in this image (C++ on
PETSc/SUNDIALS/Zoltan/Armadillo/MPI), so these vectors come from an INDEPENDENT third restatement of the reference
semantics, written directly from the reference sources in dictionary-based pure Python (no code shared with
oracle/fsp_oracle.c or the CUDA path):
    state validity / BFS closure      src/StateSet/StateSetConstrained.cpp:33-56,132-221
    matrix entries                    src/Matrix/FspMatrixBase.cpp:132-145,180-191,231-240
    sink entries (every violated k)   src/Matrix/FspMatrixConstrained.cpp:170-194, StateSetConstrained.cpp:63-82
    Action = sum_r c_r(t) A_r x       src/Matrix/FspMatrixBase.cpp:36-62, FspMatrixConstrained.cpp:31-64
Vectors are keyed by STATE (not by index) so that they are valid for any index order.
It also records the reference's own analytic known answers with their file:line (tests/golden/reference_kats.json).

    python tests/golden/make_golden.py        # rewrites the json files next to this script
"""
import json
import math
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- workload definitions (parameters restate the reference fixtures; see pacmensl_b200/fixtures/fsp_models.h) ----
def rw_prop(r, x):
    return 2.0 if r == 0 else 3.0 * (x[0] > 0)


def rw_tfun(t):
    return [1.0 + t, 1.0 + 0.5 * t]


def toggle_prop(r, x):
    ayx, axy, nyx, nxy = 2.6e-3, 6.1e-3, 3.0, 2.1
    kx0, kx, dx, ky0, ky, dy = 2.2e-3, 1.7e-2, 3.8e-4, 6.8e-5, 1.6e-2, 3.8e-4
    return [kx0, kx / (1.0 + ayx * float(x[1]) ** nyx), dx * x[0], ky0, ky / (1.0 + axy * float(x[0]) ** nxy), dy * x[1]][r]


def hog1p_prop(r, x):
    k12, k23, k34, k32, k43, k21 = 1.29, 0.0067, 0.133, 0.027, 0.0381, 1.0
    kr21, kr31, kr41, kr22, kr32, kr42 = 0.005, 0.45, 0.025, 0.0116, 0.987, 0.0538
    trans, g1, g2 = 0.01, 0.001, 0.0049
    g = x[0]
    return [k12 * (g == 0) + k23 * (g == 1) + k34 * (g == 2), k32 * (g == 2) + k43 * (g == 3), k21 * (g == 1),
            kr21 * (g == 1) + kr31 * (g == 2) + kr41 * (g == 3), kr22 * (g == 1) + kr32 * (g == 2) + kr42 * (g == 3),
            trans * x[1], trans * x[2], g1 * x[3], g2 * x[4]][r]


def hog1p_tfun(t):
    r1, r2, eta, Ahog, Mhog = 6.9e-5, 7.1e-3, 3.1, 9.3e09, 6.4e-4
    c = [1.0] * 9
    h1 = (1.0 - math.exp(-r1 * t)) * math.exp(-r2 * t)
    hog = (h1 / (1.0 + h1 / Mhog)) ** eta * Ahog
    c[2] = max(0.0, 3200.0 - 7710.0 * hog)
    return c


WORKLOADS = {
    "random_walk_1d_tv": dict(SM=[[1], [-1]], prop=rw_prop, tfun=rw_tfun, tv=[0, 1], lhs=None, bounds=[12], x0=[0],
                              times=[0.0, 0.1, 1.0, 10.0]),
    "toggle_custom": dict(SM=[[1, 0], [1, 0], [-1, 0], [0, 1], [0, 1], [0, -1]], prop=toggle_prop, tfun=None, tv=[],
                          lhs=lambda x: [x[0], x[1], x[0] * x[1]], bounds=[9, 7, 20], x0=[0, 0], times=[0.0]),
    "hog1p": dict(SM=[[1, 0, 0, 0, 0], [-1, 0, 0, 0, 0], [-1, 0, 0, 0, 0], [0, 1, 0, 0, 0], [0, 0, 1, 0, 0],
                      [0, -1, 0, 1, 0], [0, 0, -1, 0, 1], [0, 0, 0, -1, 0], [0, 0, 0, 0, -1]], prop=hog1p_prop,
                  tfun=hog1p_tfun, tv=[2], lhs=None, bounds=[3, 3, 3, 2, 2], x0=[0, 0, 0, 0, 0], times=[0.0, 25.0, 120.0]),
}


def build(w):
    lhs = w["lhs"] or (lambda x: list(x))
    bounds = w["bounds"]
    K = len(bounds)

    def valid(x):
        return all(v >= 0 for v in x) and all(l <= b for l, b in zip(lhs(x), bounds))

    # BFS closure (any order; vectors are keyed by state)
    states = [tuple(w["x0"])]
    seen = {states[0]}
    frontier = list(states)
    while frontier:
        nxt = []
        for x in frontier:
            for nu in w["SM"]:
                y = tuple(a + b for a, b in zip(x, nu))
                if valid(y) and y not in seen:
                    seen.add(y)
                    states.append(y)
                    nxt.append(y)
        frontier = nxt
    return states, K, lhs, bounds


def action(w, states, K, lhs, bounds, t, xs, xsink):
    """y = A(t) [x; xsink] keyed by state, plus the K sink rows."""
    R = len(w["SM"])
    c = [1.0] * R
    if w["tv"]:
        ct = w["tfun"](t)
        for r in w["tv"]:
            c[r] = ct[r]
    y = {s: 0.0 for s in states}
    ysink = [0.0] * K
    for s in states:
        for r, nu in enumerate(w["SM"]):
            d = w["prop"](r, s)
            y[s] -= c[r] * d * xs[s]
            src = tuple(a - b for a, b in zip(s, nu))
            if src in xs:
                y[s] += c[r] * w["prop"](r, src) * xs[src]
            dst = tuple(a + b for a, b in zip(s, nu))
            if all(v >= 0 for v in dst):
                for k, (l, b) in enumerate(zip(lhs(dst), bounds)):
                    if l > b:
                        ysink[k] += c[r] * d * xs[s]
    return y, ysink  # sink columns are empty: xsink never feeds back


def main():
    rnd = random.Random(20240607)
    for name, w in WORKLOADS.items():
        states, K, lhs, bounds = build(w)
        xs = {s: rnd.random() for s in states}
        xsink = [rnd.random() for _ in range(K)]
        cases = []
        for t in w["times"]:
            y, ysink = action(w, states, K, lhs, bounds, t, xs, xsink)
            cases.append({"t": t, "y": [y[s] for s in states], "y_sink": ysink})
        out = {"workload": name, "bounds": bounds, "num_states": len(states), "states": [list(s) for s in states],
               "x": [xs[s] for s in states], "x_sink": xsink, "cases": cases,
               "source": "tests/golden/make_golden.py (independent pure-Python restatement; the reference stores no vectors)"}
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f)
        print(name, len(states), "states")
    # The reference's own analytic known answers for this path (file:line in /root/reference)
    lam = 20.0
    kats = {
        "KAT-M1": {"ref": "tests/test_mat.cpp:110-151", "what": "sum(FspMatrixBase * ones), 13-state random walk", "value": -2.0, "tol": 0.0},
        "KAT-M2": {"ref": "tests/test_mat.cpp:199-238", "what": "sum(FspMatrixConstrained * ones)", "value": 0.0, "tol": 0.0},
        "KAT-M5": {"ref": "tests/test_mat.cpp:289-341", "what": "||J(t) x - Action(t, x)|| at t in {0,.1,.2,1,10}", "value": 0.0, "tol": 1e-14},
        "KAT-S1": {"ref": "tests/test_fss.cpp:70-125", "what": "states of {x0 + x1 <= 3}", "value": 10},
        "KAT-S2": {"ref": "tests/test_fss.cpp:39-65", "what": "AddStates with a wrong species count returns", "value": -1},
        "KAT-O1/O2": {"ref": "tests/test_ode.cpp:123-153,220-259", "what": "|sum(p) - 1| after toggle solve to t=100", "tol": 1e-8},
        "KAT-F4/F5": {"ref": "tests/test_fsp_solver.cpp:179-300", "what": "pure birth rate 2, t_f = 10: L1 error vs Poisson(20)", "tol": 1e-6,
                      "poisson_pmf_first_40": [math.exp(-lam + n * math.log(lam) - math.lgamma(n + 1.0)) for n in range(40)]},
        "KAT-SF2": {"ref": "tests/test_sensfsp_solver.cpp", "what": "Poisson pmf (1e-7... L1) and its lambda-derivative (1e-6)", "tol_p": 1e-7, "tol_s": 1e-6},
    }
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(kats, f, indent=1)


if __name__ == "__main__":
    main()

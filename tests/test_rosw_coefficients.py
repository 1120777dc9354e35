"""CPU: the Rosenbrock-W coefficients of pacmensl_b200/host/TsFsp.cpp (RA34PW2, the scheme behind PETSc's TSROSW default
"ra34pw2", which the reference's TsFsp uses: src/OdeSolver/TsFsp.cpp:31-79) are read from the C++ source and checked by
what defines them: third order for the main weights and second order for the embedded ones, with the exact Jacobian
AND with a perturbed one (the W property), on a small nonlinear system."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _coefficients():
    src = open(os.path.join(ROOT, "pacmensl_b200", "host", "TsFsp.cpp")).read()
    gamma = float(re.search(r"kGamma\s*=\s*([-+0-9.eE]+)", src).group(1))

    def arr(name):
        body = re.search(name + r"(?:\[4\])+\s*=\s*\{(.*?)\};", src, re.S).group(1).replace("kGamma", repr(gamma))
        return np.array([float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", body)])
    return arr("kA").reshape(4, 4), arr("kG").reshape(4, 4), arr("kB"), arr("kB2")


def test_ra34pw2_orders():
    A, G, b, b2 = _coefficients()
    assert abs(b.sum() - 1.0) < 1e-14 and abs(b2.sum() - 1.0) < 1e-14

    def f(y):
        return np.array([y[1], -y[0] - 0.5 * y[0] ** 3 + 0.1 * y[1] ** 2])

    def jac(y):
        return np.array([[0.0, 1.0], [-1.0 - 1.5 * y[0] ** 2, 0.2 * y[1]]])

    def run(h, w, inexact):
        y = np.array([1.0, 0.3])
        for _ in range(int(round(1.0 / h))):
            Jm = jac(y) + (np.array([[0.3, -0.2], [0.1, 0.4]]) if inexact else 0.0)
            k = np.zeros((4, 2))
            for i in range(4):
                yi = y + sum(A[i, j] * k[j] for j in range(i))
                rhs = h * f(yi) + h * Jm @ sum((G[i, j] * k[j] for j in range(i)), np.zeros(2))
                k[i] = np.linalg.solve(np.eye(2) - h * G[i, i] * Jm, rhs)
            y = y + sum(w[i] * k[i] for i in range(4))
        return y

    ref = run(1.0 / 4096, b, False)
    for w, order in ((b, 3.0), (b2, 2.0)):
        for inexact in (False, True):
            errs = [np.abs(run(h, w, inexact) - ref).max() for h in (1 / 32, 1 / 64, 1 / 128)]
            observed = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
            assert all(abs(o - order) < 0.1 for o in observed), (order, inexact, observed)

"""CPU ORACLE (test infrastructure, NOT product code): restatement of the adaptive FSP driver
FspSolverMultiSinks::{Solve, Advance_, CheckFspTolerance_} (src/Fsp/FspSolverMultiSinks.cpp:62-224, 576-643) on top
of the oracle state set / operator / KrylovFsp restatement, plus a tight-tolerance reference integration of the same
truncated problem for the solvers whose step sequences the reference does not pin (CVODE).

Only tests/ may import this module.  PARITY PIN: analytic answers (Poisson pmf, tests/test_fsp_solver.cpp:264-345) --
tests/test_examples_small.py.
"""
import math

import numpy as np

from . import oracle as O
from .krylov_oracle import KrylovOracle


class FspDriverOracle:
    def __init__(self, fixture, bounds=None, expansion=None, fsp_tol=None, t_final=None, abs_tol=1.0e-14):
        fx = O.fixture_info(fixture)
        self.name = fixture
        self.bounds = np.array(fx["bounds"] if bounds is None else bounds, dtype=np.int32)
        self.expansion = np.array(fx["expansion"] if expansion is None else expansion, dtype=np.float64)
        self.fsp_tol = fx["fsp_tol"] if fsp_tol is None else fsp_tol
        self.t_final = fx["t_final"] if t_final is None else t_final
        self.x0 = np.array(fx["x0"], dtype=np.int32).reshape(1, -1)
        self.abs_tol = abs_tol
        self.expansions = 0
        self.num_rhs = 0
        self.history = []  # (t, bounds after the expansion, n_states)
        self.steps = []    # Krylov steps (t_now, t_step, m, rejects, err_loc)

    # FspSolverMultiSinks.cpp:576-611
    def _check(self, t, p):
        K = self.K
        sinks = p[-K:]
        excess = 0.0
        for k in range(K):
            if sinks[k] / self.fsp_tol >= (1.0 / K) * (t / self.t_final):
                self.to_expand[k] = 1
                excess = max(excess, sinks[k] * K - self.fsp_tol * (t / self.t_final))
        return excess

    def _build(self):
        A = O.FspMatrix(constrained=True)
        assert A.generate_fixture(self.set, self.name) == 0
        return A

    # FspSolverMultiSinks.cpp:619-643 + :62-224 with ODESolverType KRYLOV
    def solve(self):
        self.set = O.StateSet(fixture=self.name, bounds=list(self.bounds))  # x0 added, SetShapeBounds
        assert self.set.expand() == 0
        self.K = self.set.K
        A = self._build()
        n = self.set.n
        p = np.zeros(n + self.K)
        p[self.set.state2index(self.x0)[0]] = 1.0
        t_now = 0.0
        self.to_expand = np.zeros(self.K, dtype=np.int32)
        m_prev = 30  # KrylovFsp.h:54; the solver object outlives the restarts, so m_ carries over (only m_next_ is reset)
        while True:
            kry = KrylovOracle(A, abs_tol=self.abs_tol)  # SetUp(): fresh workspace, first step re-initialised (:92-101)
            kry.m = m_prev
            self.to_expand[:] = 0
            check = (lambda t, v: self._check(t, v)) if self.fsp_tol > 0.0 else None
            stat, t_now = kry.solve_with_stop(p, t_now, self.t_final, check)
            self.num_rhs += kry.num_rhs
            m_prev = kry.m
            self.steps.extend(kry.trace)
            if stat == 0:
                break
            for k in range(self.K):  # :116-123
                if self.to_expand[k] == 1:
                    # std::round rounds halves away from zero (Python's round() rounds them to even)
                    self.bounds[k] = int(math.floor(float(self.bounds[k]) * (self.expansion[k] + 1.0) + 0.5 + 0.5))
            states_old = self.set.states().copy()
            self.set.set_bounds(list(self.bounds))
            assert self.set.expand() == 0
            A = self._build()
            n_new = self.set.n
            loc = self.set.state2index(states_old)  # :177
            assert (loc >= 0).all()
            sinks_loc = np.arange(n_new, n_new + self.K)  # :183-193
            p = O.expand_vec(p, np.concatenate([loc, sinks_loc]).astype(np.int32), n_new + self.K)  # PetscWrap.cpp:26-56
            self.expansions += 1
            self.history.append((t_now, self.bounds.copy(), n_new))
        self.A = A
        return self.set.states(), p[: self.set.n].copy(), p[self.set.n:].copy()


def tight_reference(A, p0, t_final, t_init=0.0, rtol=1e-12, atol=1e-16):
    """Tight-tolerance integration of dp/dt = A(t) p on a FIXED state set with the oracle operator (scipy LSODA/BDF are
    not needed: the problems used in the tests are small enough for an explicit high-order method with error control)."""
    from scipy.integrate import solve_ivp
    y = np.empty(A.nrows)

    def f(t, x):
        ierr, out = A.action(t, x)
        if ierr:
            raise RuntimeError("rhs failed")
        return out

    sol = solve_ivp(f, (t_init, t_final), np.asarray(p0, dtype=np.float64), method="DOP853", rtol=rtol, atol=atol)
    if not sol.success:
        raise RuntimeError(sol.message)
    return sol.y[:, -1], sol.nfev

"""CPU ORACLE (test infrastructure, NOT product code): restatement of the reference's Krylov exponential integrator
KrylovFsp (src/OdeSolver/KrylovFsp.cpp:29-485, constants src/OdeSolver/KrylovFsp.h:50-72) on top of the oracle Action.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.  The dense 62 x 62 matrix
exponential is scipy.linalg.expm (Pade-13 scaling and squaring, the published algorithm behind arma::expmat, Armadillo
9.880.1); vector operations are the OpenMP kernels orc_vec_* of fsp_oracle.c.

PARITY PIN: the reference binary cannot run here; this restatement is pinned by the analytic answers the reference's
own solver tests use (Poisson pmf, tests/test_fsp_solver.cpp:264-345; mass conservation, tests/test_ode.cpp:220-259)
-- tests/test_oracle_krylov.py.

One deliberate deviation, shared with the product (pacmensl_b200/host/KrylovFsp.cpp): a non-finite error estimate
(overflow of exp(tau H)) shrinks the step by the controller's own lower bound 0.2 instead of feeding inf to pow/log,
which in the reference formulas retries the same step until max_reject_ (`robust=False` restores the formulas as written).
"""
import ctypes as C
import math

import numpy as np
import scipy.linalg

from . import oracle as O


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class KrylovOracle:
    """rhs is evaluated at t = 0 for every basis vector, as the reference does (KrylovFsp.cpp:137,152,296)."""

    def __init__(self, A, abs_tol=1.0e-14, m_min=25, m_max=60, q_iop=2, robust=True):
        self.A = A
        self.L = O.lib()
        self.n = A.nrows
        self.abs_tol = abs_tol
        self.m_min, self.m_max, self.m_next = m_min, m_max, m_min   # KrylovFsp.h:50-60, SetKrylovDimRange
        self.m = 30                                                  # KrylovFsp.h:54
        self.q_iop = q_iop
        self.delta, self.gamma, self.btol, self.max_reject = 1.2, 0.9, 1.0e-14, 10000
        self.robust = robust
        self.Vm = np.zeros((0, self.n))
        self.av = np.empty(self.n)
        self.Hm = np.zeros((m_max + 2, m_max + 2))
        self.first_step_initialized = False
        self.rhs_cost = A.flops()
        self.num_rhs = 0
        self.num_steps = 0
        self.t_step_next = 0.0
        self.trace = []   # (t_now, t_step, m, rejects, err_loc) per accepted step

    # -- helpers ------------------------------------------------------------------------------------------------
    def _ensure(self, count):
        if self.Vm.shape[0] < count:
            V = np.empty((count, self.n))
            V[: self.Vm.shape[0]] = self.Vm
            self.Vm = V

    def _rhs(self, x, y):
        self.num_rhs += 1
        ierr = self.A.action_into(0.0, x, y)
        if ierr:
            raise RuntimeError("rhs failed")

    def _norm(self, x):
        return math.sqrt(self.L.orc_vec_dot(self.n, _dp(x), _dp(x)))

    # -- KrylovFsp.cpp:264-322 ---------------------------------------------------------------------------------------
    def generate_basis(self, v, m_start):
        if m_start >= self.m:
            return
        L, n, Vm, Hm = self.L, self.n, self.Vm, self.Hm
        self.k1, self.mb = 2, self.m
        self.beta = self._norm(v)
        L.orc_vec_copy(n, _dp(v), _dp(Vm[0]))
        L.orc_vec_scale(n, 1.0 / self.beta, _dp(Vm[0]))
        istart = 0
        if m_start == 0:
            Hm[:] = 0.0
        for j in range(m_start, self.m):
            self._rhs(Vm[j], Vm[j + 1])
            if self.q_iop > 0:
                istart = j - self.q_iop + 1 if j - self.q_iop + 1 >= 0 else 0
            for i in range(istart, j + 1):
                Hm[i, j] = L.orc_vec_dot(n, _dp(Vm[j + 1]), _dp(Vm[i]))
                L.orc_vec_axpy(n, -Hm[i, j], _dp(Vm[i]), _dp(Vm[j + 1]))
            s = self._norm(Vm[j + 1])
            L.orc_vec_scale(n, 1.0 / s if s > 0 else 0.0, _dp(Vm[j + 1]))
            Hm[j + 1, j] = s
            if s < self.btol:
                self.k1, self.mb = 0, j + 1
                break

    # -- KrylovFsp.cpp:457-478 ---------------------------------------------------------------------------------------
    def estimate_cost(self, tau_new, m_new):
        hnorm = np.abs(self.Hm).sum(axis=1).max()
        ns = math.ceil(hnorm * tau_new)
        q = self.q_iop
        if q > 0:
            return (m_new + 1) * self.rhs_cost + (4 * q * m_new + 5 * m_new + 2 * q - 2 * q * q + 7) * self.n + \
                2.0 * math.ceil(25.0 / 3.0 + ns) * float((m_new + 2) ** 3)
        return (m_new + 1) * self.rhs_cost + (4 * m_new * m_new + 5 * m_new + 2 * m_new - 2 * m_new * m_new + 7) * self.n + \
            2.0 * math.ceil(25.0 / 3.0 + ns) * float((m_new + 2) ** 3)

    # -- KrylovFsp.cpp:101-262 ---------------------------------------------------------------------------------------
    def advance_one_step(self, v, t_now, t_final):
        ireject, m_start, kappa = 0, 0, 2.0
        m_old, t_step_old, omega, omega_old = 0, 0.0, 0.0, 0.0
        bsize_changed = False
        # KrylovFsp.cpp:119 precedes :122: `order` is seeded with the dimension of the PREVIOUS step (the header's initial
        # m_ = 30, KrylovFsp.h:54, on the very first one), not with the dimension this step is about to use
        order = self.m / 4.0
        err_loc = 0.0
        while True:
            self.m = min(self.m_max, max(self.m_min, self.m_next))
            m = self.m
            self._ensure(m + 1)
            self.generate_basis(v, m_start)
            if not self.first_step_initialized:
                xm = 1.0 / m
                self._rhs(v, self.av)
                avnorm = self._norm(self.av)
                anorm = avnorm / self.beta
                fact = ((m + 1) / math.e) ** (m + 1) * math.sqrt(2 * 3.1416 * (m + 1))
                self.t_step_next = (1.0 / anorm) * ((fact * self.abs_tol) / (4.0 * self.beta * anorm)) ** xm
                self.first_step_initialized = True
            t_step = min(t_final - t_now, self.t_step_next)
            if self.k1 != 0:
                self.Hm[m + 1, m] = 1.0
                self._rhs(self.Vm[m], self.av)
                self.avnorm = self._norm(self.av)
            with np.errstate(over="ignore", invalid="ignore"):
                F = scipy.linalg.expm(t_step * self.Hm)
            if self.k1 == 0:
                err_loc = self.btol
                break
            phi1 = abs(self.beta * F[m, 0])
            phi2 = abs(self.beta * F[m + 1, 0] * self.avnorm)
            with np.errstate(over="ignore", invalid="ignore"):
                if phi1 > phi2 * 10.0:
                    err_loc = phi2
                elif phi1 > phi2:
                    err_loc = (phi1 * phi2) / (phi1 - phi2)
                else:
                    err_loc = phi1
                omega_new = err_loc / (self.abs_tol * t_step)
            if self.robust and not (math.isfinite(err_loc) and math.isfinite(omega_new)):
                if ireject == self.max_reject:
                    raise RuntimeError("KrylovFsp: maximum number of failed steps reached")
                ireject += 1
                t_step_old, m_old, m_start = t_step, m, m
                self.m_next, self.t_step_next, bsize_changed = m, 0.2 * t_step, False
                continue
            omega_old, omega = omega, omega_new
            if bsize_changed and ireject > 0:
                kappa = max(1.1, (omega / omega_old) ** (1.0 / (m_old - m)))
            elif ireject > 0:
                order = max(1.0, math.log(omega / omega_old) / math.log(t_step / t_step_old))
            t_suggest = self.gamma * t_step * omega ** (-1.0 / order)
            s = 10.0 ** (math.floor(math.log10(t_suggest)) - 1)
            t_suggest = math.ceil(t_suggest / s) * s
            t_suggest = min(5.0 * t_step, max(0.2 * t_step, t_suggest))
            t_suggest = min(t_final - t_now, t_suggest)
            m_suggest = m + int(math.ceil(math.log(omega / self.gamma) / math.log(kappa)))
            m_suggest = max(3 * m // 4, min(4 * m // 3 + 1, m_suggest))
            m_suggest = max(self.m_min, min(self.m_max, m_suggest))
            cost_t = self.estimate_cost(t_suggest, m)
            cost_m = self.estimate_cost(t_step, m_suggest)
            if math.ceil((t_final - t_now) / t_suggest) * cost_t <= math.ceil((t_final - t_now) / t_step) * cost_m or m_suggest == m:
                self.t_step_next, self.m_next, bsize_changed = t_suggest, m, False
            else:
                self.t_step_next, self.m_next, bsize_changed = t_step, m_suggest, True
            if omega <= self.delta:
                break
            if bsize_changed:
                self.Hm[m + 1, m] = 0.0
            if ireject == self.max_reject:
                raise RuntimeError("KrylovFsp: maximum number of failed steps reached")
            ireject += 1
            t_step_old, m_old, m_start = t_step, m, m
        mx = self.mb + max(0, self.k1 - 1)
        F0 = np.ascontiguousarray(self.beta * F[:mx, 0])
        self.L.orc_vec_maxpy(self.n, mx, _dp(F0), _dp(self.Vm), self.Vm.shape[1], _dp(v))
        self.num_steps += 1
        self.trace.append((t_now, t_step, m, ireject, err_loc))
        return t_now + t_step

    # -- KrylovFsp.cpp:364-400 (deg = 0): the solution at t inside the last step, from the stored basis ------------------
    def get_dky(self, t, t_now, v):
        with np.errstate(over="ignore", invalid="ignore"):
            F = scipy.linalg.expm((t - t_now) * self.Hm)
        mx = self.mb + max(0, self.k1 - 1)
        F0 = np.ascontiguousarray(self.beta * F[:mx, 0])
        self.L.orc_vec_maxpy(self.n, mx, _dp(F0), _dp(self.Vm), self.Vm.shape[1], _dp(v))

    # -- KrylovFsp.cpp:29-99 without a stop condition ---------------------------------------------------------------------
    def solve(self, p0, t_final, t_init=0.0):
        v = np.ascontiguousarray(p0, dtype=np.float64).copy()
        t = t_init
        while t < t_final:
            t = self.advance_one_step(v, t, t_final)
        return v

    # -- KrylovFsp.cpp:29-99 with the stop condition of the FSP driver ----------------------------------------------------
    def solve_with_stop(self, v, t_init, t_final, stop_check):
        """Integrates v in place from t_init; stop_check(t, v) -> error_excess.  Returns (stop, t_now): stop = 1 when
        the check reported an excess after a step -- the reference then halves the step ten times WITHOUT re-evaluating
        the excess (it is never updated inside that loop, :59-74), so it always ends at t_step_tmp = 0: the solution is
        rolled back to t_now through GetDky(t_now) and t_now is not advanced."""
        t_now = t_init
        while t_now < t_final:
            t_tmp = self.advance_one_step(v, t_now, t_final)
            if stop_check is not None and stop_check(t_tmp, v) > 0.0:
                self.get_dky(t_now, t_now, v)
                return 1, t_now
            t_now = t_tmp
        return 0, t_now

"""Synthetic 3-D birth-death lattice workload (SURVEY.md section 8d / BASELINE.json config 4) through the host
C++ classes: StateSetConstrained::AddBoxLattice (+ Expand), Model with a mass-action description (device-side
propensity evaluation), FspMatrixConstrained::GenerateValues, Action.

S=3, SM = [+e1,-e1,+e2,-e2,+e3,-e3], births (40,30,20), deaths gamma*x with gamma=(1.0,1.5,2.0), constraints
x_s <= L_s (K=3 sinks), lexicographic order with species 0 fastest.  tv=True: the three birth reactions are
time-varying with c(t) = 1 + 0.5 sin(0.1 t).
"""
import numpy as np

from . import api

SM = np.array([[1, -1, 0, 0, 0, 0], [0, 0, 1, -1, 0, 0], [0, 0, 0, 0, 1, -1]], dtype=np.int32)
RATES = [40.0, 1.0, 30.0, 1.5, 20.0, 2.0]
ORDERS = np.array([[0, 1, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 0, 1]], dtype=np.int32)  # S x R


class Lattice:
    def __init__(self, upper, tv=False, expand=True, sharded=None):
        self.upper = [int(u) for u in upper]
        self.set = api.StateSet(SM, sharded=sharded)
        assert self.set.set_shape(self.upper) == 0
        self.set.add_box_lattice(self.upper)
        if expand:  # closure check + status bookkeeping, as the reference's AddStates -> Expand sequence
            assert self.set.expand() == 0
        self.model = api.Model(fixture="birth_death_3d_tv" if tv else "birth_death_3d")
        self.model.set_mass_action(RATES, ORDERS)
        self.mat = api.FspMatrix(constrained=True)
        ierr = self.mat.generate(self.set, self.model)
        if ierr:
            raise api.FspError("GenerateValues failed: %d" % ierr)
        self.n_local, self.n_global, self.start = self.set.sizes()
        self.n_rows, self.flops, self.bytes = self.mat.info()
        self.n = self.n_local

    def action(self, t, x, y):
        ierr = self.mat.action(t, x, y)
        if ierr:
            raise api.FspError("Action failed: %d" % ierr)


def build_birth_death_lattice(upper, tv=False, expand=True):
    lat = Lattice(upper, tv=tv, expand=expand)
    return lat, lat.n_local

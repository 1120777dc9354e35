"""GPU side of the solve-level comparison: run the adaptive FSP solves of the example workloads at reduced t_f with
verbosity 2 (every expansion and every Krylov step is printed by the host classes) and save the final (states, p) so
the traces can be compared offline with oracle/fsp_driver_oracle.py.   python tools/trace_solves.py [out_dir]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = [("pure_birth", None), ("repressilator", 0.5), ("transcr_reg_6d", 10.0), ("hog1p", 5.0)]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
    from pacmensl_b200 import api
    api.init(0)
    for ode, label in ((api.KRYLOV, "krylov"), (api.CVODE, "cvode")):
        for name, tf in CASES:
            s, m = api.fixture_solver(name, ode)
            fx = m.fixture
            s.set_verbosity(2 if ode == api.KRYLOV else 1)
            print("==== %s %s t_final=%g fsp_tol=%g" % (label, name, tf or fx["t_final"], fx["fsp_tol"]), flush=True)
            states, p = s.solve(tf or fx["t_final"], fx["fsp_tol"])
            st = s.stats()
            print("==== done %s %s: %d states, %d expansions, %d rhs, bounds %s, sum %.15f" % (
                label, name, st["n_states"], st["expansions"], st["rhs_evals"], st["bounds"], p.sum()), flush=True)
            np.savez(os.path.join(out, "trace_%s_%s.npz" % (label, name)), states=states, p=p)
            s.clear()
    api.finalize()


if __name__ == "__main__":
    main()

"""Runs the C++ host-layer test programs (tests/cpp/*.cpp, shaped after the reference's gtest programs) on the GPU.

They exercise the pacmensl:: classes (StateSetConstrained, FspMatrixBase/Constrained, KrylovFsp, CvodeFsp,
FspSolverMultiSinks) which call the CUDA library through the C ABI.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(name, timeout=600):
    exe = os.path.join(ROOT, "build", "tests", name)
    assert os.path.exists(exe), "%s not built (run make)" % exe
    r = subprocess.run([exe], capture_output=True, text=True, timeout=timeout)
    print(r.stdout[-6000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0, "%s failed:\n%s" % (name, r.stdout[-4000:])
    assert "0 failed" in r.stdout
    return r.stdout


# (the example workloads at reduced t_f are compared with the oracle in tests/test_examples_small.py)
@pytest.mark.parametrize("name", ["test_mat", "test_fss", "test_ode", "test_fsp_solver", "test_sensmat",
                                  "test_sensfsp_solver"])
def test_cpp_program(cuda, name):
    _run(name)

"""CPU-only: the C-ABI library loads and exports every symbol include/fsp_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fsp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"FSP_API\s+[\w\s\*]+?\b(fsp\w*)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_are_exported_and_bound():
    from pacmensl_b200 import _capi
    L = _capi.lib()  # raises if the library is missing: there is no CPU fallback
    names = declared_symbols()
    assert len(names) > 70
    for n in names:
        assert hasattr(L, n), "symbol %s declared in include/fsp_b200.h is not exported" % n
        assert n in _capi.SIGNATURES, "symbol %s has no ctypes signature in pacmensl_b200/_capi.py" % n
    for n in _capi.SIGNATURES:
        assert n in names, "%s bound in _capi.py but not declared in the header" % n


def test_host_layer_symbols_exported():
    L = ctypes.CDLL(os.path.join(ROOT, "pacmensl_b200", "lib", "libpacmensl_b200.so"))
    for n in ["VecCreate", "VecDestroy", "VecNorm", "MatMult", "pacmensl_comm_world", "MPI_Comm_rank"]:
        assert hasattr(L, n)


def test_cpp_test_programs_are_built():
    for n in ["test_mat", "test_fss", "test_ode", "test_fsp_solver"]:
        assert os.path.exists(os.path.join(ROOT, "build", "tests", n))


def test_launch_counter_and_error_string_without_gpu():
    from pacmensl_b200 import _capi
    L = _capi.lib()
    assert L.fsp_launch_count() >= 0
    assert isinstance(L.fsp_last_error(), bytes)

"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): launches tests/multirank_check.py under torchrun
and the C++ host tests as N ranks (the counterpart of the reference's `mpirun -np N <gtest program>`)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu(cuda):
    return cuda.cuda.device_count()


@pytest.mark.parametrize("nranks", [2, 4])
def test_partitioned_action_and_solves(cuda, nranks):
    if _ngpu(cuda) < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "multirank_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-4000:], r.stderr[-3000:])
    assert r.returncode == 0 and "MULTIRANK OK" in r.stdout


@pytest.mark.parametrize("prog", ["test_mat", "test_fss", "test_ode", "test_fsp_solver", "test_sensmat", "test_sensfsp_solver"])
def test_cpp_programs_two_ranks(cuda, prog):
    if _ngpu(cuda) < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([os.path.join(ROOT, "tools", "launch_ranks.sh"), "2", os.path.join(ROOT, "build", "tests", prog)],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-4000:], r.stderr[-2000:])
    assert r.returncode == 0 and "0 failed" in r.stdout


@pytest.mark.parametrize("nranks", [2, 4])
def test_sharded_state_set(cuda, nranks):
    # distributed construction (src/StateSet/StateSetBase.cpp:134-154,188-258): set == oracle, BLOCK layout,
    # striped directory consistent on every rank, Action parity by state key, solves == the replicated directory's
    if _ngpu(cuda) < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr",
           "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tests", "multirank_sharded_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-6000:], r.stderr[-3000:])
    assert r.returncode == 0 and "SHARDED OK" in r.stdout


@pytest.mark.parametrize("prog", ["test_fss", "test_mat", "test_fsp_solver"])
def test_cpp_programs_two_ranks_replicated_sets(cuda, prog):
    # the state sets of test_cpp_programs_two_ranks are sharded (the multi-GPU default); here the same programs with the
    # replicated directory (FSP_SHARDED_SET=0: every rank holds and expands the whole set)
    if _ngpu(cuda) < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([os.path.join(ROOT, "tools", "launch_ranks.sh"), "2", os.path.join(ROOT, "build", "tests", prog)],
                       capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, FSP_SHARDED_SET="0"))
    print(r.stdout[-4000:], r.stderr[-2000:])
    assert r.returncode == 0 and "0 failed" in r.stdout


def test_missing_peer_is_reported_not_silently_wrong(cuda):
    # a device-side flag wait that times out must poison the outputs and end in a non-zero return code
    if _ngpu(cuda) < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29561", os.path.join(ROOT, "tests", "multirank_timeout_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, FSP_SPIN_TIMEOUT_MS="300"))
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "TIMEOUT CHECK OK" in r.stdout

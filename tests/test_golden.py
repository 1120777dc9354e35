"""Golden vectors (tests/golden/*.json, made by tests/golden/make_golden.py: an independent pure-Python restatement of
the reference semantics, keyed by state) against (1) the CPU oracle [CPU] and (2) the CUDA path through the host C ABI
(StateSetConstrained::Expand -> FspMatrixConstrained::GenerateValues -> Action) [GPU].  Tolerance 1e-12 relative
(north_star: "SpMV agrees with the reference's MatMult to 1e-12 relative")."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["random_walk_1d_tv", "toggle_custom", "hog1p"]
TOL = 1e-12


def load(name):
    return json.load(open(os.path.join(HERE, "golden", name + ".json")))


def test_reference_kats_file_lists_the_reference_pins():
    k = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
    assert k["KAT-M1"]["value"] == -2.0 and k["KAT-M2"]["value"] == 0.0 and k["KAT-S1"]["value"] == 10
    assert abs(sum(k["KAT-F4/F5"]["poisson_pmf_first_40"]) - 1.0) < 1e-4


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_golden(oracle, name):
    g = load(name)
    st = oracle.StateSet(fixture=name, bounds=g["bounds"])
    assert st.expand() == 0
    assert st.n == g["num_states"]
    idx = st.state2index(np.array(g["states"], dtype=np.int32))
    assert (idx >= 0).all() and len(set(idx.tolist())) == st.n
    A = oracle.FspMatrix(constrained=True)
    assert A.generate_fixture(st, name) == 0
    K = len(g["x_sink"])
    x = np.zeros(st.n + K)
    x[idx] = g["x"]
    x[st.n:] = g["x_sink"]
    for c in g["cases"]:
        ierr, y = A.action(c["t"], x)
        assert ierr == 0
        scale = max(np.abs(c["y"]).max(), 1e-300)
        assert np.abs(y[idx] - np.array(c["y"])).max() <= TOL * scale
        assert np.abs(y[st.n:] - np.array(c["y_sink"])).max() <= TOL * scale


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_host_path_matches_golden(cuda, name):
    torch = cuda
    from pacmensl_b200 import api
    g = load(name)
    api.init(0)
    model = api.Model(fixture=name)
    st, mat = api.fixture_set_and_matrix(name, g["bounds"], model=model)
    n = st.n_local
    assert n == g["num_states"]
    idx = st.state2index(np.array(g["states"], dtype=np.int32))
    assert (idx >= 0).all() and len(set(idx.tolist())) == n
    K = len(g["x_sink"])
    assert mat.n_rows == n + K
    x = np.zeros(n + K)
    x[idx] = g["x"]
    x[n:] = g["x_sink"]
    xd = torch.from_numpy(x).cuda()
    for c in g["cases"]:
        yd = torch.full((n + K,), float("nan"), dtype=torch.float64, device="cuda")
        assert mat.action(c["t"], xd, yd) == 0
        y = yd.cpu().numpy()
        scale = max(np.abs(c["y"]).max(), 1e-300)
        assert np.abs(y[idx] - np.array(c["y"])).max() <= TOL * scale
        assert np.abs(y[n:] - np.array(c["y_sink"])).max() <= TOL * scale

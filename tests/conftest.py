import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def unit_vectors():
    return load_json("ref_unit_vectors.json")


@pytest.fixture(scope="session")
def acmove_random():
    return load_npz("acmove_random.npz")


@pytest.fixture(scope="session")
def env_traces():
    return load_npz("env_traces.npz")


@pytest.fixture(scope="session")
def search_cases():
    return load_json("search_cases.json")


@pytest.fixture(scope="session")
def search_visited():
    return load_npz("search_visited.npz")


@pytest.fixture(scope="session")
def miller_schupp():
    return load_npz("miller_schupp.npz")


def ms_row(ms, k):
    """Row k of all_presentations.txt at its own max_relator_length."""
    m = int(ms["mrl"][k])
    p = ms["presentations36"][k]
    return np.concatenate([p[:m], p[36 : 36 + m]]).astype(np.int8)


def ms_path(ms, k):
    """Stored greedy path k in ACMove ids: the data file holds (action+1, length)."""
    o = ms["greedy_path_offsets"]
    seg = ms["greedy_path_flat"][o[k] : o[k + 1]]
    return [(int(a) - 1, int(l)) for a, l in seg]

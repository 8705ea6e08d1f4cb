import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return graft.load_package()


@pytest.fixture(scope="session")
def oracle():
    return graft.load_oracle()


@pytest.fixture(scope="session")
def golden8():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "position8.json")) as f:
        g = json.load(f)
    g["position8"] = np.asarray(g["position8"], np.float32)
    return g


@pytest.fixture(scope="session")
def golden_digests():
    """tests/golden/pair_digests.json (made by tests/golden/make_pair_digests.py): committed digests of exact pair sets."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "pair_digests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def big_handle(pkg):
    """One handle for the GPU tests (1.1M atoms)."""
    h = pkg.Handle(1_100_000)
    yield h
    h.close()


def uniform_positions(n, seed):
    """uniform [0,1)^3: Float64 draw -> Float32, as generate_positions (MDInput.jl:179-187)."""
    return np.random.default_rng(seed).random((n, 3)).astype(np.float32)

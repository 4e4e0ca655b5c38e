"""pytest configuration: the `gpu` marker, import paths and shared fixtures."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_addoption(parser):
    parser.addoption("--run-slow", action="store_true", default=False, help="run deep perft cases")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: long CPU test (deep perft)")


def load_golden(name):
    """-> (dict of arrays, flat weight blob rebuilt from the recorded seed)."""
    from oracle import weights as W
    g = dict(np.load(os.path.join(GOLDEN, "net_%s.npz" % name)))
    blocks, hidden = int(g["blocks"]), int(g["hidden"])
    blob = W.flatten(W.make_weights(blocks, hidden, int(g["weight_seed"]), str(g["variant"])), blocks, hidden)
    # the fixture was computed with exactly these numbers
    assert abs(float(np.abs(blob.astype(np.float64)).sum()) - float(g["weights_checksum"])) < 1e-6
    return g, blob


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get


@pytest.fixture(scope="session")
def cb():
    """The product binding; skips cleanly when there is no GPU, fails loudly when the
    library is missing on a GPU box."""
    import caissa_b200
    caissa_b200.load_library()
    return caissa_b200

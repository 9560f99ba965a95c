import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def state_dicts():
    from ml_conformer_generator_b200.weights import random_state_dicts
    return random_state_dicts(0)


_ENGINES = {}


@pytest.fixture(scope="session")
def engines(state_dicts):
    """Lazily built engines per precision (weights loaded once)."""
    from ml_conformer_generator_b200.engine import Engine

    def get(precision):
        if precision not in _ENGINES:
            e = Engine(torch.device("cuda:0"), precision)
            e.load_edm_state_dict(state_dicts[0])
            e.load_seer_state_dict(state_dicts[1])
            _ENGINES[precision] = e
        return _ENGINES[precision]

    yield get
    for e in _ENGINES.values():
        e.close()
    _ENGINES.clear()


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def tables():
    from mpinets_b200 import franka
    return franka.default_tables()


@pytest.fixture(scope="session")
def state_dict(oracle):
    return oracle.reference_state_dict(0)


@pytest.fixture(scope="session")
def engine(tables):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mpinets_b200.engine import Engine
    return Engine(device=0, tables=tables, seed=0x4D50694E)


@pytest.fixture(scope="session")
def engine_w(engine, state_dict):
    engine.load_state_dict(state_dict)
    return engine


def to_dev(d, keys=None):
    import torch
    keys = keys or d.keys()
    return {k: torch.from_numpy(np.ascontiguousarray(d[k])).cuda() for k in keys if isinstance(d[k], np.ndarray)}

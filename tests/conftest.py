import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    def load(name):
        return np.load(GOLDEN / f'{name}.npz')
    return load


@pytest.fixture(scope='session')
def lib():
    """The built C-ABI library; building is part of the CPU suite's job (nvcc cross-compiles without a GPU)."""
    from raider_b200 import _lib, build
    build.build()
    return _lib.load()

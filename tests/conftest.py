import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    import json
    with open(os.path.join(ROOT, 'tests', 'golden', 'golden.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def fixture_cube():
    import numpy as np
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'anom_test.npz'))
    return d['anom'], d['latitude'], d['longitude']


@pytest.fixture(scope='session')
def reference_run():
    """Results of the unmodified reference source (tests/golden/make_reference_golden.py)."""
    import json
    with open(os.path.join(ROOT, 'tests', 'golden', 'reference_run.json')) as f:
        return json.load(f)

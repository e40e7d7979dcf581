import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def golden():
    import json
    import numpy as np
    gdir = os.path.join(ROOT, 'tests', 'golden')
    arrays = dict(np.load(os.path.join(gdir, 'golden.npz')))
    with open(os.path.join(gdir, 'golden.json')) as f:
        meta = json.load(f)
    return arrays, meta

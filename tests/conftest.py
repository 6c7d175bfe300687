import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def frags():
    from slv_b200.slvpar import Frag
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Frag(name)
        return cache[name]
    return get


@pytest.fixture(scope='session')
def golden():
    import json
    with open(os.path.join(ROOT, 'tests', 'golden', 'small_cluster.json')) as fh:
        return json.load(fh)

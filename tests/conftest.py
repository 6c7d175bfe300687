import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    # the in-tree CUDA library, the host build of the device math and the transpiled reference are git-ignored build
    # products: (re)build whatever is missing or stale before the tests import them (nvcc cross-compiles without a GPU)
    try:
        import __graft_entry__
        __graft_entry__.build()
    except Exception as exc:          # tests that need the products will fail loudly on import
        print('conftest: build() failed: %r' % (exc,))


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

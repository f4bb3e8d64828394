import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')
    config.addinivalue_line('markers', 'slow: long CPU test, enabled by CMU_SLOW_TESTS=1')


def pytest_collection_modifyitems(config, items):
    if os.environ.get('CMU_SLOW_TESTS') == '1':
        return
    skip = pytest.mark.skip(reason='set CMU_SLOW_TESTS=1')
    for it in items:
        if 'slow' in it.keywords:
            it.add_marker(skip)

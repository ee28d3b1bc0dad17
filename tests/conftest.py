import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')
    config.addinivalue_line('markers', 'slow: takes more than a few seconds')


@pytest.fixture(scope='session')
def lib_built():
    """Builds (if stale) and returns the path of libcpsd_b200.so."""
    from cross_patient_speech_decoding_b200 import build
    return build.build()

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def bundles():
    from vajax_b200.models.fixtures import get_bundle

    return {k: get_bundle(k) for k in ["resistor", "capacitor", "diode", "psp103", "opcov", "psp103w"]}

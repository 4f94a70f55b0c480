import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with `-m gpu`")
    # a wedged GPU wait must fail the test instead of hanging the run (pytest-timeout is in the image)
    if config.pluginmanager.hasplugin("timeout") and not config.getoption("timeout", None):
        config.option.timeout = 900


@pytest.fixture(scope="session")
def golden():
    import torch

    def load(name):
        return torch.load(os.path.join(ROOT, "tests", "golden", name), map_location="cpu", weights_only=False)
    return load

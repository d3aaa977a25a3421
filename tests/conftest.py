import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def example_scene():
    from raygun_b200 import scene as S
    return S.load_example_scene()[0]


@pytest.fixture(scope="session")
def oracle_example(example_scene):
    from oracle import oracle as O
    return O.OracleScene(example_scene)

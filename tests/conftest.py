"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI export checks (CPU only).
`-m gpu`      : parity tests proper -- the CUDA path through the C-ABI vs the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def clo():
    import cl_ops_b200
    cl_ops_b200.lib()
    return cl_ops_b200


@pytest.fixture(scope="session")
def ctx(clo):
    c = clo.Context()
    yield c
    c.destroy()


@pytest.fixture(scope="session")
def queue(clo, ctx):
    q = clo.Queue(ctx)
    yield q
    q.destroy()

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def gpu_solver():
    """One handle for the whole GPU session; fails loudly when the extension or GPU is missing."""
    from opm_simulators_legacy_b200.solver import GpuLinearSolver
    s = GpuLinearSolver(0)
    yield s
    s.close()

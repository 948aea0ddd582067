import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import __graft_entry__ as ge  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def pkg():
    return ge.load_package()


@pytest.fixture(scope="session")
def orc():
    return ge.load_oracle()


@pytest.fixture(scope="session")
def wl(pkg):
    return pkg.workloads


def hard_config(wl, N=10, mu=0.3, dt=0.03, wf=1e-2, disc_mode=0):
    """Tracking-heavy weights + low friction: friction-pyramid rows go active (workloads.hard_config)."""
    return wl.hard_config(N, mu, dt, wf, disc_mode)


def to_step_major(forces, N, L):
    """forces [B, L*N*3] (per-leg order) -> U [B, 3LN] (step-major)."""
    B = forces.shape[0]
    return forces.reshape(B, L, N, 3).transpose(0, 2, 1, 3).reshape(B, -1)

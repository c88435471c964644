"""Shared fixtures.  `gpu` marks tests that need a real B200 (run with `-m gpu` via gpurun)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run under gpurun")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "singleroom_golden.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): oracle/oracle.py over oracle/librcw_oracle.so."""
    from oracle import oracle as orc
    orc.build()
    return orc


GOLDEN_CONFIGS = {
    "A": dict(),
    "B": dict(H=64, W=64, N=256, R=128, P=96),
    "C": dict(H=5, W=7, N=36, R=45, P=51, radius=np.float32(0.2), incr=np.float32(0.3),
              sfov=np.float32(0.5), cam_h=np.float32(0.8)),
    "D": dict(tie_le=1, dist_post=1),
}

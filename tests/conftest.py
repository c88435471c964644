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


def pytest_sessionstart(session):
    """The CUDA library is a build artefact (git-ignored): compile it with nvcc if it is missing, so
    that a fresh checkout can run the suite.  No GPU is needed to build."""
    lib = os.path.join(ROOT, "raycastworlds.jl_b200", "lib", "librcw_b200.so")
    if not os.path.exists(lib):
        import importlib.util as u

        spec = u.spec_from_file_location("_rcw_build", os.path.join(ROOT, "raycastworlds.jl_b200", "build.py"))
        mod = u.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build(force=True)


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
    "B": dict(H=64, W=64, N=256, R=128, P=96, pu_per_tu=4),
    "C": dict(H=5, W=7, N=36, R=45, P=51, radius=np.float32(0.2), incr=np.float32(0.3),
              sfov=np.float32(0.5), cam_h=np.float32(0.8), pu_per_tu=7),
    "D": dict(tie_le=1, dist_post=1),
    # exact ties between the side distances (tile centres / corners, 8 directions, odd ray count): T with the
    # default tie rule, U with RCW_DDA_TIE_LE — the states that let a Julia dump resolve decision D1
    "T": dict(H=7, W=7, N=8, R=33, P=40, pu_per_tu=4),
    "U": dict(H=7, W=7, N=8, R=33, P=40, pu_per_tu=4, tie_le=1),
}


# Case L of the golden fixture: NUM_OBJECTS = 5 — an interior wall, blocking pillars (object 3), a
# terminal layer with reward -1 (object 4) and one with reward 0.5 (object 5) on a 9 x 12 map.
LAYERED_CONFIG = dict(H=9, W=12, N=32, R=64, P=48, radius=np.float32(0.15), incr=np.float32(0.2), pu_per_tu=4,
                      num_layers=5, layer_kind=[0, 1, 1, 0], layer_reward=[0.0, -1.0, 0.5, 0.0],
                      layer_palette=[0x00205080, 0x003070A0, 0x00A04000, 0x00C06000, 0x0000A040, 0x0000C060, 0, 0],
                      layer_top_color=[0x000000FF, 0x00FF8000, 0x0000FF00, 0])


def layered_oracle_world(oracle, golden):
    """An oracle World configured and furnished like case L."""
    w = oracle.World(oracle.default_config(**LAYERED_CONFIG))
    w.set_layer(1, golden["L_wall"])
    for k in range(3):
        w.set_layer(3 + k, golden["L_extra"][k])
    return w

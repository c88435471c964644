"""Pins the oracle (and with -m gpu the CUDA path) against the REAL Julia reference, when somebody has
run tools/dump_reference.jl off-box and committed tests/golden/julia_reference.npz.  Without that file
the test is skipped with the reason spelled out: parity versus Julia is unpinned."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_CONFIGS

PIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "julia_reference.npz")


@pytest.mark.skipif(not os.path.exists(PIN), reason="PARITY UNPINNED vs Julia: tests/golden/julia_reference.npz "
                    "is absent (tools/dump_reference.jl needs Julia + RayCaster.jl 0.1, not in this image)")
@pytest.mark.parametrize("case", ["A", "B", "C"])
def test_oracle_matches_julia(oracle, golden, case):
    ref = dict(np.load(PIN))
    w = oracle.World(oracle.default_config(**GOLDEN_CONFIGS[case]))
    states, au, goal = golden[f"{case}_states"], golden[f"{case}_au"], golden[f"{case}_goal"]
    for k in range(len(states)):
        w.set_state(states[k, 0], states[k, 1], au[k], goal[k, 0], goal[k, 1])
        w.cast_rays()
        w.update_camera_view()
        np.testing.assert_array_equal(w.ray_stop, ref[f"{case}_hit"][k])
        np.testing.assert_array_equal(w.ray_dim, ref[f"{case}_dim"][k])
        np.testing.assert_allclose(w.ray_dist, ref[f"{case}_dist"][k], rtol=1e-5)
        np.testing.assert_array_equal(w.camera_view, ref[f"{case}_image"][k])
        if f"{case}_top" in ref:                      # pins the restated SimpleDraw shapes (line, circle) as well
            w.update_top_view()
            np.testing.assert_array_equal(w.top_view, ref[f"{case}_top"][k])

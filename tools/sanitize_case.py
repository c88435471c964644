#!/usr/bin/env python
"""Small end-to-end cases for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
covers the fused step kernel (mirror-pair and pitched renderers, both formats), per-env wall
layers, masked reset with host layouts, the ray dump, and the bulk / split variants."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raycastworlds_jl_b200 as rcw  # noqa: E402


def run(n, **kw):
    env = rcw.BatchedSingleRoom(n, seed=3, **kw)
    env.step_random(6)
    env.act(np.random.default_rng(0).integers(1, 5, n).astype(np.uint8))
    env.reset(mask=(np.arange(n) % 2).astype(np.uint8))
    env.reset(goal_ij=np.full((n, 2), 3), player_ij=np.full((n, 2), 2), dir_au=np.arange(n) % 8)
    env.get_rays()
    env.copy_obs()
    env.get_state()
    env.episode_stats()
    env.close()


run(5)
run(3, obs_format="xrgb32")
run(7, num_rays=45, height_camera_view_pu=51)
run(7, num_rays=33, height_camera_view_pu=84, obs_format="xrgb32")
run(4, height_tile_map_tu=64, width_tile_map_tu=64, num_directions=256, num_rays=96, height_camera_view_pu=64)
e = rcw.BatchedSingleRoom(9, seed=1, height_tile_map_tu=12, width_tile_map_tu=20, num_rays=32, height_camera_view_pu=32)
w = np.zeros((9, 12, 20), bool)
w[:, 0, :] = w[:, -1, :] = w[:, :, 0] = w[:, :, -1] = True
w[:, 5, 5:9] = True
e.set_wall_maps(w)
e.reset()
e.step_random(5)
e.copy_obs()
e.close()
print("sanitize cases done")

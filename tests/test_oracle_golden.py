"""CPU oracle (C) versus the golden fixtures made by the independent Python restatement
(tests/golden/make_golden.py).  Neither is the Julia reference: parity is unpinned."""
import zlib

import numpy as np
import pytest

from conftest import GOLDEN_CONFIGS, layered_oracle_world


@pytest.mark.parametrize("case", ["A", "B", "C", "D", "T", "U"])
def test_cast_and_render_match_golden(oracle, golden, case):
    cfg = oracle.default_config(**GOLDEN_CONFIGS[case])
    w = oracle.World(cfg)
    states, au, goal = golden[f"{case}_states"], golden[f"{case}_au"], golden[f"{case}_goal"]
    for k in range(len(states)):
        w.set_state(states[k, 0], states[k, 1], au[k], goal[k, 0], goal[k, 1])
        w.cast_rays()
        w.update_camera_view()
        np.testing.assert_array_equal(w.ray_stop, golden[f"{case}_hit"][k])
        np.testing.assert_array_equal(w.ray_dim, golden[f"{case}_dim"][k])
        # bit-exact: same binary32 operation order
        np.testing.assert_array_equal(w.ray_dist.view(np.uint32), golden[f"{case}_dist"][k].view(np.uint32))
        np.testing.assert_array_equal(w.ray_dir.view(np.uint32), golden[f"{case}_ray_dir"][k].view(np.uint32))
        np.testing.assert_array_equal(w.wall_heights(), golden[f"{case}_height"][k])
        img = w.camera_view
        assert zlib.crc32(img.tobytes()) == int(golden[f"{case}_crc"][k])
        if f"{case}_image" in golden and k < len(golden[f"{case}_image"]):
            np.testing.assert_array_equal(img, golden[f"{case}_image"][k])
        # update_top_view! (single_room.jl:446-483)
        w.update_top_view()
        top = w.top_view
        assert zlib.crc32(top.tobytes()) == int(golden[f"{case}_top_crc"][k])
        if f"{case}_top_image" in golden and k < len(golden[f"{case}_top_image"]):
            np.testing.assert_array_equal(top, golden[f"{case}_top_image"][k])


@pytest.mark.parametrize("case", ["A", "B", "C"])
def test_act_trajectories_match_golden(oracle, golden, case):
    cfg = oracle.default_config(**GOLDEN_CONFIGS[case])
    w = oracle.World(cfg)
    init, actions = golden[f"{case}_act_init"], golden[f"{case}_act_actions"]
    n_done = 0
    for ep in range(len(init)):
        gi, gj, pi, pj, a0 = (int(v) for v in init[ep])
        w.reset_to(gi, gj, pi, pj, a0)
        for t, a in enumerate(actions[ep]):
            assert w.act(int(a)) == 0
            s = w.state()
            assert s["pos"].view(np.uint32).tolist() == golden[f"{case}_act_pos"][ep, t].view(np.uint32).tolist()
            assert s["au"] == golden[f"{case}_act_au"][ep, t]
            assert s["reward"] == golden[f"{case}_act_reward"][ep, t]
            assert s["done"] == bool(golden[f"{case}_act_done"][ep, t])
            n_done += s["done"]
    if case == "A":
        assert n_done > 0, "fixture should contain goal hits"


@pytest.mark.parametrize("case", ["A", "C"])
def test_byte_formats_written_directly_equal_the_converted_uint32_view(oracle, golden, case):
    """bench.py's CPU arm writes RGB8 / GRAY8 frames directly (orc_update_camera_view_bytes) so that both arms
    store the same bytes per frame; the bytes must be those of the reference's UInt32 picture."""
    kw = dict(GOLDEN_CONFIGS[case])
    if case == "C":
        kw["palette"] = [0x112233, 0x445566, 0x778899, 0xAABBCC, 0xDD1122, 0x3344EE]
    w = oracle.World(oracle.default_config(**kw))
    states, au, goal = golden[f"{case}_states"], golden[f"{case}_au"], golden[f"{case}_goal"]
    for k in range(0, len(states), 3):
        w.set_state(states[k, 0], states[k, 1], au[k], goal[k, 0], goal[k, 1])
        w.cast_rays()
        w.update_camera_view()
        np.testing.assert_array_equal(w.frame_bytes("rgb8"), w.obs_rgb8())
        c = w.camera_view
        luma = ((77 * ((c >> 16) & 255) + 150 * ((c >> 8) & 255) + 29 * (c & 255) + 128) >> 8).astype(np.uint8)
        np.testing.assert_array_equal(w.frame_bytes("gray8"), luma)


def test_object_layers_match_golden(oracle, golden):
    """NUM_OBJECTS = 5 (case L): rays stop at any object, columns take the first object's colours, the top view shows
    findfirst, terminal layers end the episode with their own reward, blocking layers refuse the move."""
    w = layered_oracle_world(oracle, golden)
    states, au, goal = golden["L_states"], golden["L_au"], golden["L_goal"]
    for k in range(len(states)):
        w.set_state(states[k, 0], states[k, 1], au[k], goal[k, 0], goal[k, 1])
        w.cast_rays()
        w.update_camera_view()
        np.testing.assert_array_equal(w.ray_stop, golden["L_hit"][k])
        np.testing.assert_array_equal(w.ray_dim, golden["L_dim"][k])
        np.testing.assert_array_equal(w.ray_dist.view(np.uint32), golden["L_dist"][k].view(np.uint32))
        np.testing.assert_array_equal(w.camera_view, golden["L_image"][k])
        w.update_top_view()
        np.testing.assert_array_equal(w.top_view, golden["L_top_image"][k])
        cols = w.camera_columns()                       # column words carry the object's colour id 2 * object + (dim != 1)
        assert set(np.unique(cols >> 16)) <= set(range(2, 12))
    colours = set(np.unique(golden["L_image"]))
    assert {0x00205080, 0x003070A0} & colours and {0x00A04000, 0x00C06000} & colours, "the fixture should show the extra objects"
    init, actions = golden["L_act_init"], golden["L_act_actions"]
    for ep in range(len(init)):
        gi, gj, pi, pj, a0 = (int(v) for v in init[ep])
        w.reset_to(gi, gj, pi, pj, a0)
        for t, a in enumerate(actions[ep]):
            assert w.act(int(a)) == 0
            s = w.state()
            assert s["pos"].view(np.uint32).tolist() == golden["L_act_pos"][ep, t].view(np.uint32).tolist()
            assert s["au"] == golden["L_act_au"][ep, t]
            assert s["reward"] == golden["L_act_reward"][ep, t]
            assert s["done"] == bool(golden["L_act_done"][ep, t])
    assert set(np.unique(golden["L_act_reward"])) == {-1.0, 0.0, 0.5, 1.0}

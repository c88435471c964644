"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol that
include/rcw_b200.h declares (and nothing else under the rcw_ prefix), the ctypes mirror of
rcw_config matches the C layout, argument validation answers without a GPU, and creation fails
loudly (RCW_ECUDA) instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import pytest

import raycastworlds_jl_b200 as rcw
from raycastworlds_jl_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rcw_b200.h")


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rcw_[a-z_0-9]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_capi.LIB_PATH), "run `python __graft_entry__.py build`"
    assert os.path.realpath(_capi.LIB_PATH).startswith(os.path.realpath(ROOT))


def test_exports_match_header():
    declared = header_functions()
    assert declared == sorted(_capi.SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (rcw_[a-z_0-9]+)$", out, flags=re.M)))
    assert exported == declared
    lib = _capi.load()
    for name in declared:
        assert hasattr(lib, name)


def test_library_has_sm100a_code_and_no_oracle():
    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    needed = subprocess.run(["readelf", "-d", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in needed.lower()


def test_config_struct_layout_and_defaults():
    cfg = _capi.default_config()
    assert cfg.struct_size == C.sizeof(_capi.RcwConfig) == 248
    assert (cfg.height_tile_map_tu, cfg.width_tile_map_tu) == (8, 16)        # single_room.jl:44-45
    assert (cfg.num_directions, cfg.num_rays, cfg.height_camera_view_pu) == (128, 512, 256)
    assert cfg.player_radius_wu == 0.125 and cfg.position_increment_wu == 0.125
    assert cfg.semi_field_of_view_wu == pytest.approx(2 / 3, rel=1e-7)
    assert list(cfg.palette) == [0xFFFFFF, 0x404040, 0x808080, 0xC0C0C0, 0x800000, 0xC00000]
    assert (cfg.obs_window_envs, cfg.top_view, cfg.pu_per_tu, cfg.frame_stack, cfg.result_ring) == (0, 0, 32, 0, 0)  # :269; the top view is opt-in for a batch
    assert list(cfg.top_palette) == [0xFFFFFF, 0xFF0000, 0x000000, 0xCCCCCC, 0x808080, 0xC0C0C0]  # :288-290, 364-367
    assert cfg.num_object_layers == 2 and not any(cfg.layer_kind) and not any(cfg.layer_top_color)      # NUM_OBJECTS = 2, :16
    assert _capi.load().rcw_version() == _capi.ABI_VERSION


def test_validation_errors_without_gpu():
    lib = _capi.load()
    h = C.c_void_p()
    cfg = _capi.default_config()
    cfg.struct_size = 8
    assert lib.rcw_create(C.byref(cfg), None, C.byref(h)) == _capi.RCW_ESIZE
    assert b"struct_size" in lib.rcw_last_error()
    for field, value in [("num_envs", 0), ("height_tile_map_tu", 2), ("player_radius_wu", 0.6),
                         ("obs_format", 7), ("num_rays", 0), ("dda_flags", 8), ("obs_window_envs", -1), ("pu_per_tu", 0),
                         ("top_view", 2), ("frame_stack", 65), ("result_ring", 65), ("result_ring", -1),
                         ("num_object_layers", 1), ("num_object_layers", 7)]:
        cfg = _capi.default_config()
        setattr(cfg, field, value)
        assert lib.rcw_create(C.byref(cfg), None, C.byref(h)) == _capi.RCW_EINVAL, field
        assert not h.value
    assert lib.rcw_step(None, None) == _capi.RCW_EINVAL
    assert lib.rcw_step_tape(None, None, 3) == _capi.RCW_EINVAL and lib.rcw_set_layer(None, 3, None) == _capi.RCW_EINVAL
    assert lib.rcw_step_sharded(None, 2, None) == _capi.RCW_EINVAL and lib.rcw_sync_sharded(None, 0) == _capi.RCW_EINVAL
    for field, value in [("layer_kind", 2)]:                 # per-layer settings are validated for the layers in use
        cfg = _capi.default_config()
        cfg.num_object_layers = 3
        cfg.layer_kind[0] = value
        assert lib.rcw_create(C.byref(cfg), None, C.byref(h)) == _capi.RCW_EINVAL, field
    cfg = _capi.default_config()
    cfg.obs_format, cfg.num_rays = _capi.RCW_OBS_GRAY8_HALF, 33       # the 2 x 2 box filter needs even sizes
    assert lib.rcw_create(C.byref(cfg), None, C.byref(h)) == _capi.RCW_EINVAL
    assert lib.rcw_sync(None) == _capi.RCW_EINVAL
    assert lib.rcw_destroy(None) == _capi.RCW_OK


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rcw.RcwError) as e:
        rcw.BatchedSingleRoom(4)
    assert e.value.code == _capi.RCW_ECUDA
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "raycastworlds.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".jl")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "rcw_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_action_validation_in_host_mirror():
    assert rcw.NUM_ACTIONS == 4
    assert rcw.ACTION_NAMES == ("MOVE_FORWARD", "MOVE_BACKWARD", "TURN_LEFT", "TURN_RIGHT")
    assert issubclass(rcw.InvalidActionError, AssertionError)


def test_julia_binding_covers_the_header():
    """The Julia file cannot run here (no julia binary); what can be checked is that it binds every entry point the
    header declares (rcw_config_init excepted: the binding fills the struct itself) and that its mirror of
    rcw_config lists the C fields in the C order."""
    jl = open(os.path.join(ROOT, "raycastworlds.jl_b200", "julia", "BatchedRayCastWorlds.jl")).read()
    bound = set(re.findall(r"\(:(rcw_[a-z_0-9]+), LIB\)", jl))
    assert bound == set(header_functions()) - {"rcw_config_init"}
    struct = jl[jl.index("struct RcwConfig"):]
    struct = struct[:struct.index("\nend")]
    julia_fields = re.findall(r"^\s+([a-z_0-9]+)::", struct, flags=re.M)
    assert julia_fields == [name for name, _ in _capi.RcwConfig._fields_]
    assert f"const RCW_ABI_VERSION = Int32({_capi.ABI_VERSION})" in jl

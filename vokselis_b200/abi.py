"""ctypes mirrors of the frozen ABI structs in include/vokselis_rt.h (no compute, importable on CPU).

CameraUniform  <- src/camera.rs:5-11            (144 B)
Uniform        <- src/context/global_ubo.rs:52-65 (48 B)
Offset         <- examples/xor/main.rs:20-25      (8 B)
"""
from __future__ import annotations

import ctypes as C

MODE_M0, MODE_M1 = 0, 1
DTYPE_U8, DTYPE_F16, DTYPE_F32 = 0, 1, 2
LAYOUT_LINEAR, LAYOUT_BRICKED, LAYOUT_TEXTURE, LAYOUT_GATHER, LAYOUT_QUAD = 0, 1, 2, 3, 4

OK, ERR_INVALID, ERR_CUDA, ERR_NO_VOLUME, ERR_UNSUPPORTED = 0, -1, -2, -3, -4


class CameraUniform(C.Structure):
    _fields_ = [("view_position", C.c_float * 4), ("proj_view", C.c_float * 16), ("inv_proj", C.c_float * 16)]


class Uniform(C.Structure):
    _fields_ = [
        ("pos", C.c_float * 3), ("frame", C.c_uint32), ("resolution", C.c_float * 2), ("mouse", C.c_float * 2),
        ("mouse_pressed", C.c_uint32), ("time", C.c_float), ("time_delta", C.c_float), ("_padding", C.c_float),
    ]

    @classmethod
    def default(cls) -> "Uniform":
        """`impl Default for Uniform` (src/context/global_ubo.rs:67-81)."""
        u = cls()
        u.resolution[0], u.resolution[1] = 1920.0, 780.0
        u.time_delta = 1.0 / 60.0
        return u


class Offset(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("mode", C.c_int32), ("dt_scale", C.c_float), ("dt_floor", C.c_float),
        ("alpha_threshold", C.c_float), ("initial_alpha", C.c_float), ("clear_color", C.c_float * 4),
        ("tile_size", C.c_int32), ("layout", C.c_int32), ("skip_empty", C.c_int32), ("count_samples", C.c_int32),
        ("m1_srgb", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("rays_hit", C.c_uint64), ("samples_reference", C.c_uint64), ("samples_fetched", C.c_uint64),
        ("last_render_ms", C.c_float), ("_pad", C.c_float),
    ]


assert C.sizeof(CameraUniform) == 144
assert C.sizeof(Uniform) == 48
assert C.sizeof(Offset) == 8
assert C.sizeof(Params) == 88
assert C.sizeof(Stats) == 32


def default_params(mode: int = MODE_M0) -> Params:
    """Pure-Python twin of vkrt_default_params (include/vokselis_rt.h); tests check they agree."""
    p = Params()
    p.struct_size = C.sizeof(Params)
    p.mode = mode
    p.dt_scale = 1.0
    p.alpha_threshold = 0.95
    p.tile_size = 256
    p.layout = LAYOUT_LINEAR
    p.skip_empty = 0
    p.count_samples = 0
    if mode == MODE_M0:
        p.dt_floor = 0.01
        p.initial_alpha = 0.1
        p.clear_color[:] = [0.023, 0.02, 0.02, 0.0]
        p.m1_srgb = 0
    else:
        p.dt_floor = 0.0
        p.initial_alpha = 0.0
        p.clear_color[:] = [0.0, 0.0, 0.0, 0.0]
        p.m1_srgb = 0
    return p

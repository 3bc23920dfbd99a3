"""ctypes binding of oracle/_ref/libwgsl_ref.so — the reference's own WGSL, machine-translated to
C++ by oracle/wgsl2cpp.py (TEST INFRASTRUCTURE). Built only where /root/reference exists
(`make -C oracle ref`); the built .so travels to the GPU box, the sources never enter the repo."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from vokselis_b200.abi import CameraUniform, Uniform

_DIR = Path(__file__).resolve().parent
SO = _DIR / "_ref" / "libwgsl_ref.so"
_lib = None


def available() -> bool:
    return SO.exists()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(str(SO))
        vp, ci = C.c_void_p, C.c_int
        L.wref_raycast_compute.restype = ci
        L.wref_raycast_compute.argtypes = [ci, C.POINTER(CameraUniform), C.POINTER(Uniform), vp, ci, ci, vp, vp, ci, ci, ci, ci,
                                           ci, vp, ci]
        L.wref_xor_generate.restype = ci
        L.wref_xor_generate.argtypes = [ci, C.POINTER(Uniform), vp, vp, ci]
        L.wref_present.restype = ci
        L.wref_present.argtypes = [vp, ci, ci, ci, ci, vp]
        L.wref_naive_fs.restype = ci
        L.wref_naive_fs.argtypes = [vp, ci, ci, ci, ci, vp, vp, vp, ci]
        L.wref_num_threads.restype = ci
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def raycast_compute(cam: CameraUniform, color, normal, W, H, *, entry="single", offsets=None, tile_size=256, uniform=None,
                    frame=None, nthreads=0):
    color = np.ascontiguousarray(color).view(np.uint16)
    normal = np.ascontiguousarray(normal).view(np.uint16)
    nz, ny, nx = color.shape[:3]
    un = uniform if uniform is not None else Uniform.default()
    if frame is None:
        frame = np.zeros((H, W, 4), np.uint16)
    offs, n = None, 0
    if entry == "tile":
        offs = np.ascontiguousarray(np.asarray(offsets, np.float32).reshape(-1, 2))
        n = offs.shape[0]
    rc = lib().wref_raycast_compute(0 if entry == "single" else 1, C.byref(cam), C.byref(un), _p(offs), n, tile_size, _p(color),
                                    _p(normal), nx, ny, nz, W, H, _p(frame), nthreads)
    assert rc == 0
    return frame


def xor_generate(n, time=0.0, nthreads=0):
    un = Uniform.default()
    un.time = time
    color = np.zeros((n, n, n, 4), np.uint16)
    normal = np.zeros((n, n, n, 4), np.uint16)
    rc = lib().wref_xor_generate(n, C.byref(un), _p(color), _p(normal), nthreads)
    assert rc == 0
    return color, normal


def present(frame, out_w=None, out_h=None):
    frame = np.ascontiguousarray(frame).view(np.uint16)
    H, W = frame.shape[:2]
    out_w, out_h = out_w or W, out_h or H
    out = np.zeros((out_h, out_w, 4), np.uint8)
    rc = lib().wref_present(_p(frame), W, H, out_w, out_h, _p(out))
    assert rc == 0
    return out


def naive_fs(vol_u8, eyes, dirs, nthreads=0):
    vol_u8 = np.ascontiguousarray(vol_u8, np.uint8)
    nz, ny, nx = vol_u8.shape
    eyes = np.ascontiguousarray(eyes, np.float32).reshape(-1, 3)
    dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    out = np.zeros((eyes.shape[0], 4), np.float32)
    rc = lib().wref_naive_fs(_p(vol_u8), nx, ny, nz, eyes.shape[0], _p(eyes), _p(dirs), _p(out), nthreads)
    assert rc == 0
    return out


def num_threads():
    return int(lib().wref_num_threads())

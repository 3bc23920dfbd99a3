"""ctypes binding of the CPU ORACLE (oracle/_build/libvk_oracle.so). TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module. The product package (vokselis_b200) must never import it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from vokselis_b200.abi import CameraUniform, Offset, Params, Stats  # ABI structs only (no compute)

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_build" / "libvk_oracle.so"


class Volume(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("dtype", C.c_int32),
        ("color", C.c_void_p), ("normal", C.c_void_p), ("scalar", C.c_void_p),
    ]


class Brick(C.Structure):
    _fields_ = [("lo", C.c_float * 3), ("hi", C.c_float * 3)]


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    if force or not _SO.exists() or _SO.stat().st_mtime < (_DIR / "vk_oracle.cpp").stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR), "all"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        L.vko_render.restype = C.c_int
        L.vko_render.argtypes = [C.POINTER(Volume), C.POINTER(Params), C.POINTER(CameraUniform), C.c_int, C.c_int,
                                 C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Stats), C.c_int]
        L.vko_render_partial.restype = C.c_int
        L.vko_render_partial.argtypes = [C.POINTER(Volume), C.POINTER(Params), C.POINTER(CameraUniform), C.c_int,
                                         C.c_int, C.POINTER(Brick), C.c_void_p, C.c_void_p, C.c_int]
        L.vko_generate_xor.restype = C.c_int
        L.vko_generate_xor.argtypes = [C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.vko_scalar_to_rgba16f.restype = C.c_int
        L.vko_scalar_to_rgba16f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.vko_present.restype = C.c_int
        L.vko_present.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.vko_camera_uniform.restype = C.c_int
        L.vko_camera_uniform.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float * 3), C.c_float,
                                         C.POINTER(CameraUniform)]
        L.vko_rays.restype = C.c_int
        L.vko_rays.argtypes = [C.POINTER(CameraUniform), C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.vko_naive_fs.restype = C.c_int
        L.vko_naive_fs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.vko_f32_to_f16.restype = C.c_uint16
        L.vko_f32_to_f16.argtypes = [C.c_float]
        L.vko_f16_to_f32.restype = C.c_float
        L.vko_f16_to_f32.argtypes = [C.c_uint16]
        L.vko_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_volume(color=None, normal=None, scalar=None) -> tuple[Volume, list]:
    """M0: color/normal uint16 arrays [nz,ny,nx,4]. M1: scalar uint8/float16/float32 [nz,ny,nx]."""
    keep = []
    v = Volume()
    if scalar is None:
        color = np.ascontiguousarray(color).view(np.uint16)
        normal = np.ascontiguousarray(normal).view(np.uint16)
        assert color.shape == normal.shape and color.shape[-1] == 4
        v.nz, v.ny, v.nx = color.shape[:3]
        v.dtype = -1
        v.color, v.normal = color.ctypes.data, normal.ctypes.data
        keep += [color, normal]
    else:
        scalar = np.ascontiguousarray(scalar)
        v.nz, v.ny, v.nx = scalar.shape
        v.dtype = {np.dtype(np.uint8): 0, np.dtype(np.float16): 1, np.dtype(np.float32): 2}[scalar.dtype]
        v.scalar = scalar.ctypes.data
        keep.append(scalar)
    return v, keep


def camera_uniform(zoom, pitch, yaw, target=(0.0, 0.0, 0.0), aspect=16 / 9) -> CameraUniform:
    out = CameraUniform()
    t = (C.c_float * 3)(*target)
    rc = lib().vko_camera_uniform(zoom, pitch, yaw, C.byref(t), aspect, C.byref(out))
    assert rc == 0
    return out


def generate_xor(n: int, time: float = 0.0, which: int = 0, nthreads: int = 0):
    color = np.empty((n, n, n, 4), np.uint16)
    normal = np.empty((n, n, n, 4), np.uint16)
    rc = lib().vko_generate_xor(n, time, which, _ptr(color), _ptr(normal), nthreads)
    assert rc == 0
    return color, normal


def render(params: Params, cam: CameraUniform, W: int, H: int, *, color=None, normal=None, scalar=None,
           offsets=None, frame=None, want_aux=True, nthreads: int = 0):
    """Returns (frame uint16 [H,W,4], aux uint32 [H,W] or None, Stats)."""
    vol, keep = make_volume(color, normal, scalar)
    if frame is None:
        frame = np.zeros((H, W, 4), np.uint16)
    aux = np.zeros((H, W), np.uint32) if want_aux else None
    st = Stats()
    offs, n = None, 0
    if offsets is not None:
        offs = np.ascontiguousarray(np.asarray(offsets, np.float32).reshape(-1, 2))
        n = offs.shape[0]
    rc = lib().vko_render(C.byref(vol), C.byref(params), C.byref(cam), W, H, _ptr(offs), n, _ptr(frame), _ptr(aux),
                          C.byref(st), nthreads)
    if rc != 0:
        raise RuntimeError(f"vko_render failed: {rc}")
    del keep
    return frame, aux, st


def render_partial(params: Params, cam: CameraUniform, W: int, H: int, lo, hi, *, color=None, normal=None,
                   scalar=None, a_in=None, nthreads: int = 0):
    vol, keep = make_volume(color, normal, scalar)
    b = Brick()
    b.lo[:] = [float(x) for x in lo]
    b.hi[:] = [float(x) for x in hi]
    out = np.zeros((H, W, 4), np.float32)
    if a_in is not None:
        a_in = np.ascontiguousarray(a_in, np.float32)
    rc = lib().vko_render_partial(C.byref(vol), C.byref(params), C.byref(cam), W, H, C.byref(b), _ptr(a_in),
                                  _ptr(out), nthreads)
    if rc != 0:
        raise RuntimeError(f"vko_render_partial failed: {rc}")
    del keep
    return out


def scalar_to_rgba16f(scalar: np.ndarray):
    scalar = np.ascontiguousarray(scalar)
    nz, ny, nx = scalar.shape
    dt = {np.dtype(np.uint8): 0, np.dtype(np.float16): 1, np.dtype(np.float32): 2}[scalar.dtype]
    color = np.empty((nz, ny, nx, 4), np.uint16)
    normal = np.empty((nz, ny, nx, 4), np.uint16)
    rc = lib().vko_scalar_to_rgba16f(_ptr(scalar), dt, nx, ny, nz, _ptr(color), _ptr(normal))
    assert rc == 0
    return color, normal


def present(frame: np.ndarray, out_w: int | None = None, out_h: int | None = None) -> np.ndarray:
    H, W = frame.shape[:2]
    frame = np.ascontiguousarray(frame).view(np.uint16)
    if out_w is None and out_h is None:
        out = np.empty((H, W, 4), np.uint8)
        rc = lib().vko_present(_ptr(frame), W, H, _ptr(out))
    else:
        out_w, out_h = out_w or W, out_h or H
        out = np.empty((out_h, out_w, 4), np.uint8)
        L = lib()
        L.vko_present_scaled.restype = C.c_int
        L.vko_present_scaled.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        rc = L.vko_present_scaled(_ptr(frame), W, H, out_w, out_h, _ptr(out))
    assert rc == 0
    return out


def rays(cam: CameraUniform, W: int, H: int, offx: float = 0.0, offy: float = 0.0) -> np.ndarray:
    out = np.empty((H, W, 8), np.float32)
    rc = lib().vko_rays(C.byref(cam), W, H, offx, offy, _ptr(out))
    assert rc == 0
    return out


def naive_fs(vol_u8, eyes, dirs, nthreads: int = 0) -> np.ndarray:
    vol_u8 = np.ascontiguousarray(vol_u8, np.uint8)
    nz, ny, nx = vol_u8.shape
    eyes = np.ascontiguousarray(eyes, np.float32).reshape(-1, 3)
    dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    out = np.zeros((eyes.shape[0], 4), np.float32)
    rc = lib().vko_naive_fs(_ptr(vol_u8), nx, ny, nz, eyes.shape[0], _ptr(eyes), _ptr(dirs), _ptr(out), nthreads)
    assert rc == 0
    return out


def num_threads() -> int:
    return int(lib().vko_num_threads())

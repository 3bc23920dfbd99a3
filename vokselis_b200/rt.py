"""ctypes binding of libvokselis_rt.so — the product's C ABI (include/vokselis_rt.h).

There is no fallback: if the shared library is missing, importing a symbol from here raises, and if
no CUDA device is usable `Context(...)` raises `VokselisError` (VKRT_ERR_CUDA). Names follow the
reference's host interface (src/lib.rs:13-18): Camera, Context, RaycastPipeline ("single"/"tile"),
XorCompute, VolumeTexture, dispatch_optimal.
"""
from __future__ import annotations

import ctypes as C
import math
from pathlib import Path

import numpy as np

from . import abi
from .abi import CameraUniform, Offset, Params, Stats, Uniform

import os

# VKRT_LIB overrides the library path (A/B runs of two builds on the same box).
_SO = Path(os.environ.get("VKRT_LIB") or (Path(__file__).resolve().parent / "libvokselis_rt.so"))

# Every symbol include/vokselis_rt.h declares (tests check the .so exports all of them).
EXPORTS = [
    "vkrt_create", "vkrt_destroy", "vkrt_resize", "vkrt_last_error", "vkrt_default_params", "vkrt_set_params",
    "vkrt_get_params", "vkrt_upload_rgba16f", "vkrt_upload_scalar", "vkrt_generate_xor", "vkrt_download_rgba16f",
    "vkrt_render", "vkrt_render_tiles", "vkrt_tile_table", "vkrt_box_screen_bounds", "vkrt_box_screen_hull",
    "vkrt_render_batch", "vkrt_batch_frame_device_ptr", "vkrt_readback_batch", "vkrt_frames_host", "vkrt_present", "vkrt_present_scaled", "vkrt_readback", "vkrt_readback_rgba8", "vkrt_readback_rgba8_async",
    "vkrt_readback_aux", "vkrt_sync", "vkrt_frame_host", "vkrt_frame_host_async", "vkrt_frame_host_wait",
    "vkrt_frame_host_slot_ptr", "vkrt_frame_device_ptr", "vkrt_frame_rgba8_device_ptr", "vkrt_stream", "vkrt_stats",
    "vkrt_reset_stats", "vkrt_timing_enable", "vkrt_timing_read", "vkrt_flush_l2", "vkrt_set_occupancy_brick", "vkrt_volume_info", "vkrt_camera_uniform", "vkrt_dispatch_optimal",
    "vkrt_sortfirst_create_root", "vkrt_sortfirst_join", "vkrt_sortfirst_leave", "vkrt_sortfirst_partition", "vkrt_sortfirst_render", "vkrt_sortfirst_render_batch", "vkrt_sortfirst_render_tiles_batch",
    "vkrt_sortfirst_consume", "vkrt_sortfirst_timeouts", "vkrt_sortfirst_wait", "vkrt_mark", "vkrt_mark_elapsed",
    "vkrt_alloc_host", "vkrt_free_host", "vkrt_host_register", "vkrt_host_unregister", "vkrt_generate_synthetic", "vkrt_download_scalar", "vkrt_scalar_to_rgba16f",
    "vkrt_upload_window", "vkrt_generate_synthetic_window", "vkrt_window_info", "vkrt_partial_alpha", "vkrt_partial_ain",
    "vkrt_partial_color", "vkrt_partial_finalize", "vkrt_partial_relative", "vkrt_partial_resolve",
    "vkrt_exchange_create", "vkrt_exchange_open", "vkrt_exchange_push", "vkrt_exchange_wait", "vkrt_exchange_table", "vkrt_exchange_done",
    "vkrt_exchange_timeouts", "vkrt_exchange_close",
]


class VokselisError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; raises loudly when it has not been built (no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _SO.exists():
        raise ImportError(f"{_SO} is missing: build it with `make` (or __graft_entry__.build()). "
                          "vokselis_b200 has no CPU fallback.")
    L = C.CDLL(str(_SO))
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    sig = {
        "vkrt_create": (ci, [ci, ci, ci, C.POINTER(vp)]),
        "vkrt_destroy": (ci, [vp]),
        "vkrt_resize": (ci, [vp, ci, ci]),
        "vkrt_last_error": (C.c_char_p, []),
        "vkrt_default_params": (None, [ci, C.POINTER(Params)]),
        "vkrt_set_params": (ci, [vp, C.POINTER(Params)]),
        "vkrt_get_params": (ci, [vp, C.POINTER(Params)]),
        "vkrt_upload_rgba16f": (ci, [vp, vp, vp, ci, ci, ci]),
        "vkrt_upload_scalar": (ci, [vp, vp, ci, ci, ci, ci]),
        "vkrt_generate_xor": (ci, [vp, C.POINTER(Uniform), ci, ci]),
        "vkrt_download_rgba16f": (ci, [vp, vp, vp]),
        "vkrt_render": (ci, [vp, C.POINTER(CameraUniform), C.POINTER(Uniform), C.POINTER(Offset)]),
        "vkrt_render_tiles": (ci, [vp, C.POINTER(CameraUniform), C.POINTER(Uniform), vp, ci]),
        "vkrt_tile_table": (ci, [ci, ci, ci, vp, ci]),
        "vkrt_box_screen_bounds": (ci, [C.POINTER(CameraUniform), ci, ci, C.POINTER(C.c_float * 4), C.POINTER(ci)]),
        "vkrt_box_screen_hull": (ci, [C.POINTER(CameraUniform), ci, ci, vp]),
        "vkrt_render_batch": (ci, [vp, vp, ci, C.POINTER(Uniform)]),
        "vkrt_batch_frame_device_ptr": (vp, [vp, ci]),
        "vkrt_readback_batch": (ci, [vp, ci, vp]),
        "vkrt_frames_host": (ci, [vp, vp, ci, C.POINTER(Uniform), vp, ci]),
        "vkrt_present": (ci, [vp]),
        "vkrt_present_scaled": (ci, [vp, ci, ci, vp]),
        "vkrt_readback": (ci, [vp, vp]),
        "vkrt_readback_rgba8": (ci, [vp, vp]),
        "vkrt_readback_rgba8_async": (ci, [vp, vp]),
        "vkrt_readback_aux": (ci, [vp, vp]),
        "vkrt_sync": (ci, [vp]),
        "vkrt_frame_host": (ci, [vp, C.POINTER(CameraUniform), C.POINTER(Uniform), vp]),
        "vkrt_frame_host_async": (ci, [vp, C.POINTER(CameraUniform), C.POINTER(Uniform), ci]),
        "vkrt_frame_host_wait": (ci, [vp, ci, vp]),
        "vkrt_frame_host_slot_ptr": (vp, [vp, ci]),
        "vkrt_frame_device_ptr": (vp, [vp]),
        "vkrt_frame_rgba8_device_ptr": (vp, [vp]),
        "vkrt_stream": (vp, [vp]),
        "vkrt_stats": (ci, [vp, C.POINTER(Stats)]),
        "vkrt_reset_stats": (ci, [vp]),
        "vkrt_timing_enable": (ci, [vp, ci]),
        "vkrt_timing_read": (ci, [vp, vp, ci]),
        "vkrt_flush_l2": (ci, [vp]),
        "vkrt_set_occupancy_brick": (ci, [vp, ci]),
        "vkrt_volume_info": (ci, [vp, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci * 3), C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_uint64)]),
        "vkrt_camera_uniform": (ci, [cf, cf, cf, C.POINTER(cf * 3), cf, C.POINTER(CameraUniform)]),
        "vkrt_dispatch_optimal": (C.c_uint32, [C.c_uint32, C.c_uint32]),
        "vkrt_sortfirst_create_root": (ci, [vp, ci, ci, vp]),
        "vkrt_sortfirst_wait": (ci, [vp, C.c_uint64, C.c_uint64]),
        "vkrt_mark": (ci, [vp, ci]),
        "vkrt_alloc_host": (ci, [C.c_size_t, C.POINTER(vp)]),
        "vkrt_generate_synthetic": (ci, [vp, ci, ci, ci, ci, ci, C.c_uint32]),
        "vkrt_download_scalar": (ci, [vp, vp]),
        "vkrt_scalar_to_rgba16f": (ci, [vp]),
        "vkrt_upload_window": (ci, [vp, vp, vp, ci, C.POINTER(ci * 3), C.POINTER(ci * 3), C.POINTER(ci * 3)]),
        "vkrt_generate_synthetic_window": (ci, [vp, ci, ci, C.POINTER(ci * 3), C.POINTER(ci * 3), C.POINTER(ci * 3), C.c_uint32]),
        "vkrt_window_info": (ci, [vp, C.POINTER(ci * 3), C.POINTER(ci * 3)]),
        "vkrt_partial_alpha": (ci, [vp, C.POINTER(CameraUniform), C.POINTER(Uniform), vp]),
        "vkrt_partial_ain": (ci, [vp, vp, vp, ci, vp]),
        "vkrt_partial_color": (ci, [vp, C.POINTER(CameraUniform), C.POINTER(Uniform), vp, vp]),
        "vkrt_partial_finalize": (ci, [vp, C.POINTER(CameraUniform), C.POINTER(Uniform), vp]),
        "vkrt_partial_relative": (ci, [vp, C.POINTER(CameraUniform), C.POINTER(Uniform), vp, vp]),
        "vkrt_partial_resolve": (ci, [vp, vp, vp, ci, vp, vp]),
        "vkrt_free_host": (ci, [vp]),
        "vkrt_host_register": (ci, [vp, C.c_size_t]),
        "vkrt_host_unregister": (ci, [vp]),
        "vkrt_mark_elapsed": (ci, [vp, ci, ci, C.POINTER(cf)]),
        "vkrt_sortfirst_join": (ci, [vp, ci, vp]),
        "vkrt_sortfirst_leave": (ci, [vp]),
        "vkrt_sortfirst_partition": (ci, [ci, ci, ci, ci, ci, vp, ci]),
        "vkrt_sortfirst_render": (ci, [vp, C.POINTER(CameraUniform), C.POINTER(Uniform), vp, ci, C.c_uint64]),
        "vkrt_sortfirst_render_batch": (ci, [vp, vp, ci, C.POINTER(Uniform), C.c_uint64, ci]),
        "vkrt_sortfirst_render_tiles_batch": (ci, [vp, vp, ci, C.POINTER(Uniform), vp, ci, C.c_uint64]),
        "vkrt_sortfirst_consume": (ci, [vp, C.c_uint64, ci]),
        "vkrt_sortfirst_timeouts": (ci, [vp, C.POINTER(C.c_uint64)]),
        "vkrt_exchange_create": (ci, [vp, ci, ci, vp]),
        "vkrt_exchange_open": (ci, [vp, vp]),
        "vkrt_exchange_push": (ci, [vp, vp, vp, ci, C.c_uint64]),
        "vkrt_exchange_wait": (ci, [vp, C.c_uint64, ci]),
        "vkrt_exchange_table": (vp, [vp, C.c_uint64]),
        "vkrt_exchange_done": (ci, [vp, C.c_uint64]),
        "vkrt_exchange_timeouts": (ci, [vp, C.POINTER(C.c_uint64)]),
        "vkrt_exchange_close": (ci, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise VokselisError(rc, (lib().vkrt_last_error() or b"").decode(errors="replace"))


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dispatch_optimal(length: int, subgroup_size: int) -> int:
    """src/utils/mod.rs:15-18"""
    return int(lib().vkrt_dispatch_optimal(length, subgroup_size))


MAX_BATCH = 32  # VKRT_MAX_BATCH


def default_params(mode: int = abi.MODE_M0) -> Params:
    p = Params()
    lib().vkrt_default_params(mode, C.byref(p))
    return p


def tile_table(width: int, height: int, tile_size: int = 256) -> np.ndarray:
    """The reference's offset table (examples/xor/main.rs:80-95) as float32 [n, 2]."""
    n = lib().vkrt_tile_table(width, height, tile_size, None, 0)
    if n < 0:
        _check(n)
    out = np.zeros((n, 2), np.float32)
    lib().vkrt_tile_table(width, height, tile_size, _vp(out), n)
    return out


def box_screen_bounds(cam: CameraUniform, width: int, height: int):
    """(x0, y0, x1, y1), centre_row: the pixel rectangle outside which no ray can hit the box (host-only)."""
    rect = (C.c_float * 4)()
    row = C.c_int(-1)
    _check(lib().vkrt_box_screen_bounds(C.byref(cam), width, height, C.byref(rect), C.byref(row)))
    return tuple(float(v) for v in rect), int(row.value)


class Camera:
    """Orbit camera, src/camera.rs:75-171 (math runs in the library's C++ host layer)."""

    ZFAR, ZNEAR, FOVY = 100.0, 0.1, math.pi / 2.0

    def __init__(self, zoom: float, pitch: float, yaw: float, target=(0.0, 0.0, 0.0), aspect: float = 16 / 9):
        self.zoom, self.pitch, self.yaw, self.target, self.aspect = zoom, pitch, yaw, tuple(target), aspect
        self.updated = False

    def set_zoom(self, zoom: float):
        self.zoom = min(max(zoom, 0.3), self.ZFAR / 2.0)
        self.updated = True

    def add_zoom(self, delta: float):
        self.set_zoom(self.zoom + delta)

    def set_pitch(self, pitch: float):
        eps = float(np.finfo(np.float32).eps)
        self.pitch = min(max(pitch, -math.pi / 2.0 + eps), math.pi / 2.0 - eps)
        self.updated = True

    def add_pitch(self, delta: float):
        self.set_pitch(self.pitch + delta)

    def set_yaw(self, yaw: float):
        self.yaw = yaw
        self.updated = True

    def add_yaw(self, delta: float):
        self.set_yaw(self.yaw + delta)

    def set_aspect(self, width: int, height: int):
        self.aspect = width / height
        self.updated = True

    def get_proj_view_matrix(self) -> CameraUniform:
        out = CameraUniform()
        t = (C.c_float * 3)(*self.target)
        _check(lib().vkrt_camera_uniform(self.zoom, self.pitch, self.yaw, C.byref(t), self.aspect, C.byref(out)))
        return out


class OrbitInput:
    """The event -> camera mapping of the reference's `run` (src/lib.rs:64-66,150-176), float32 like the Rust code."""

    ROTATE_SPEED = np.float32(0.0025)  # src/lib.rs:65
    ZOOM_SPEED = np.float32(0.002)     # src/lib.rs:66

    def __init__(self):
        self.mouse_dragged = False

    def button(self, pressed: bool):
        self.mouse_dragged = bool(pressed)

    def mouse_wheel_lines(self, cam: "Camera", scroll: float):
        cam.add_zoom(float(-(np.float32(scroll) * np.float32(1.0)) * self.ZOOM_SPEED))

    def mouse_wheel_pixels(self, cam: "Camera", scroll_y: float):
        cam.add_zoom(float(-np.float32(scroll_y) * self.ZOOM_SPEED))

    def mouse_motion(self, cam: "Camera", dx: float, dy: float):
        if not self.mouse_dragged:
            return
        cam.add_yaw(float(-np.float32(dx) * self.ROTATE_SPEED))
        cam.add_pitch(float(np.float32(dy) * self.ROTATE_SPEED))


class Context:
    """Device + stream + rgba16f frame: the part of src/context.rs the raycast path needs."""

    def __init__(self, device: int = 0, width: int = 1280, height: int = 720):
        self._h = C.c_void_p()
        _check(lib().vkrt_create(device, width, height, C.byref(self._h)))
        self.width, self.height, self.device = width, height, device
        self.global_uniform = Uniform.default()

    def close(self):
        if getattr(self, "_h", None):
            lib().vkrt_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- params ---------------------------------------------------------------------------
    def set_params(self, p: Params):
        _check(lib().vkrt_set_params(self._h, C.byref(p)))

    def get_params(self) -> Params:
        p = Params()
        _check(lib().vkrt_get_params(self._h, C.byref(p)))
        return p

    def resize(self, width: int, height: int):
        _check(lib().vkrt_resize(self._h, width, height))
        self.width, self.height = width, height

    # -- volumes --------------------------------------------------------------------------
    def upload_rgba16f(self, color: np.ndarray, normal: np.ndarray):
        color = np.ascontiguousarray(color).view(np.uint16)
        normal = np.ascontiguousarray(normal).view(np.uint16)
        if color.shape != normal.shape or color.ndim != 4 or color.shape[3] != 4:
            raise ValueError("color/normal must be [nz, ny, nx, 4] half arrays")
        nz, ny, nx = color.shape[:3]
        _check(lib().vkrt_upload_rgba16f(self._h, _vp(color), _vp(normal), nx, ny, nz))

    def upload_scalar(self, vol: np.ndarray):
        vol = np.ascontiguousarray(vol)
        dt = {np.dtype(np.uint8): abi.DTYPE_U8, np.dtype(np.float16): abi.DTYPE_F16,
              np.dtype(np.float32): abi.DTYPE_F32}.get(vol.dtype)
        if dt is None or vol.ndim != 3:
            raise ValueError("scalar volume must be [nz, ny, nx] uint8/float16/float32")
        nz, ny, nx = vol.shape
        _check(lib().vkrt_upload_scalar(self._h, _vp(vol), dt, nx, ny, nz))

    def generate_xor(self, n: int = 256, which: int = 0, uniform: Uniform | None = None):
        un = uniform if uniform is not None else self.global_uniform
        _check(lib().vkrt_generate_xor(self._h, C.byref(un), n, which))

    def generate_synthetic(self, kind: int, dtype, nx: int, ny: int | None = None, nz: int | None = None, seed: int = 1):
        """kind 0 noise / 1 sparse blobs / 2 smooth lattice; dtype numpy uint8/float16/float32."""
        dt = {np.dtype(np.uint8): abi.DTYPE_U8, np.dtype(np.float16): abi.DTYPE_F16, np.dtype(np.float32): abi.DTYPE_F32}[np.dtype(dtype)]
        _check(lib().vkrt_generate_synthetic(self._h, kind, dt, nx, ny or nx, nz or nx, seed))

    def scalar_to_rgba16f(self):
        """N3: the resident scalar volume becomes the rgba16f colour/normal pair that mode M0 renders."""
        _check(lib().vkrt_scalar_to_rgba16f(self._h))

    def download_scalar(self) -> np.ndarray:
        info = self.volume_info()
        nx, ny, nz = info["dims"]
        dt = {abi.DTYPE_U8: np.uint8, abi.DTYPE_F16: np.float16, abi.DTYPE_F32: np.float32}[info["dtype"]]
        out = np.empty((nz, ny, nx), dt)
        _check(lib().vkrt_download_scalar(self._h, _vp(out)))
        return out

    def download_rgba16f(self):
        kind, dims = C.c_int(), (C.c_int * 3)()
        _check(lib().vkrt_volume_info(self._h, C.byref(kind), None, C.byref(dims), None, None))
        nx, ny, nz = dims[0], dims[1], dims[2]
        color = np.empty((nz, ny, nx, 4), np.uint16)
        normal = np.empty((nz, ny, nx, 4), np.uint16)
        _check(lib().vkrt_download_rgba16f(self._h, _vp(color), _vp(normal)))
        return color, normal

    def set_occupancy_brick(self, edge: int = 0):
        """Brick edge (voxels) of the occupancy grid for volumes uploaded after this call; 0 = automatic. Frames do not depend on it."""
        _check(lib().vkrt_set_occupancy_brick(self._h, int(edge)))

    def volume_info(self) -> dict:
        kind, dtype, dims = C.c_int(), C.c_int(), (C.c_int * 3)()
        tot, occ = C.c_uint64(), C.c_uint64()
        _check(lib().vkrt_volume_info(self._h, C.byref(kind), C.byref(dtype), C.byref(dims), C.byref(tot), C.byref(occ)))
        return {"kind": kind.value, "dtype": dtype.value, "dims": tuple(dims), "bricks_total": tot.value,
                "bricks_occupied": occ.value}

    # -- the hot path ---------------------------------------------------------------------
    def render(self, cam: CameraUniform, offset=None, uniform: Uniform | None = None):
        un = uniform if uniform is not None else self.global_uniform
        off = None
        if offset is not None:
            off = Offset(float(offset[0]), float(offset[1]))
        _check(lib().vkrt_render(self._h, C.byref(cam), C.byref(un), C.byref(off) if off is not None else None))

    def render_tiles(self, cam: CameraUniform, offsets, uniform: Uniform | None = None):
        un = uniform if uniform is not None else self.global_uniform
        offs = np.ascontiguousarray(np.asarray(offsets, np.float32).reshape(-1, 2))
        _check(lib().vkrt_render_tiles(self._h, C.byref(cam), C.byref(un), _vp(offs), offs.shape[0]))

    def present(self):
        _check(lib().vkrt_present(self._h))

    def present_scaled(self, out_w: int, out_h: int) -> np.ndarray:
        """vkrt_present_scaled: the present pass stretched onto an out_w x out_h target (window != backbuffer)."""
        out = np.empty((out_h, out_w, 4), np.uint8)
        _check(lib().vkrt_present_scaled(self._h, out_w, out_h, _vp(out)))
        return out

    def sync(self):
        _check(lib().vkrt_sync(self._h))

    def readback(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), np.uint16)
        _check(lib().vkrt_readback(self._h, _vp(out)))
        return out

    def readback_rgba8(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width, 4), np.uint8)
        _check(lib().vkrt_readback_rgba8(self._h, _vp(out)))
        return out

    def readback_aux(self) -> np.ndarray:
        out = np.empty((self.height, self.width), np.uint32)
        _check(lib().vkrt_readback_aux(self._h, _vp(out)))
        return out

    def readback_rgba8_async(self, out: np.ndarray):
        """Enqueue the D2H of the presented frame into page-locked `out`; valid after sync()."""
        _check(lib().vkrt_readback_rgba8_async(self._h, _vp(out)))

    def frame_host(self, cam: CameraUniform, out: np.ndarray | None = None, uniform: Uniform | None = None) -> np.ndarray:
        un = uniform if uniform is not None else self.global_uniform
        if out is None:
            out = np.empty((self.height, self.width, 4), np.uint8)
        _check(lib().vkrt_frame_host(self._h, C.byref(cam), C.byref(un), _vp(out)))
        return out

    # -- batches: several cameras of a sweep in ONE launch (grid.z = frame) ---------------------------
    @staticmethod
    def _cam_array(cams):
        arr = (CameraUniform * len(cams))()
        for i, cam in enumerate(cams):
            C.memmove(C.byref(arr, i * C.sizeof(CameraUniform)), C.byref(cam), C.sizeof(CameraUniform))
        return arr

    def render_batch(self, cams, uniform: Uniform | None = None):
        """vkrt_render_batch: len(cams) <= MAX_BATCH frames into the context's batch buffers, one launch."""
        un = uniform if uniform is not None else self.global_uniform
        arr = self._cam_array(cams)
        _check(lib().vkrt_render_batch(self._h, C.cast(arr, C.c_void_p), len(cams), C.byref(un)))

    def readback_batch(self, i: int) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), np.uint16)
        _check(lib().vkrt_readback_batch(self._h, i, _vp(out)))
        return out

    def frames_host(self, cams, out: np.ndarray | None = None, group: int = 0, uniform: Uniform | None = None) -> np.ndarray:
        """vkrt_frames_host: len(cams) presented RGBA8 frames into host memory [n, H, W, 4], pipelined in groups."""
        un = uniform if uniform is not None else self.global_uniform
        if out is None:
            out = np.empty((len(cams), self.height, self.width, 4), np.uint8)
        arr = self._cam_array(cams)
        _check(lib().vkrt_frames_host(self._h, C.cast(arr, C.c_void_p), len(cams), C.byref(un), _vp(out), group))
        return out

    def frame_host_async(self, cam: CameraUniform, slot: int, uniform: Uniform | None = None):
        un = uniform if uniform is not None else self.global_uniform
        _check(lib().vkrt_frame_host_async(self._h, C.byref(cam), C.byref(un), slot))

    def frame_host_wait(self, slot: int, out: np.ndarray | None = None):
        _check(lib().vkrt_frame_host_wait(self._h, slot, _vp(out) if out is not None else None))

    def stats(self) -> Stats:
        st = Stats()
        _check(lib().vkrt_stats(self._h, C.byref(st)))
        return st

    def timing_enable(self, capacity: int):
        _check(lib().vkrt_timing_enable(self._h, capacity))

    def timing_read(self, n: int) -> np.ndarray:
        out = np.empty(n, np.float32)
        _check(lib().vkrt_timing_read(self._h, _vp(out), n))
        return out

    def flush_l2(self):
        _check(lib().vkrt_flush_l2(self._h))

    # -- sort-first group (one process per GPU) ---------------------------------------------
    def sortfirst_create_root(self, world: int, slots: int = 2) -> bytes:
        h = (C.c_ubyte * 80)()
        _check(lib().vkrt_sortfirst_create_root(self._h, world, slots, C.byref(h)))
        return bytes(h)

    def sortfirst_wait(self, frame_index: int, arrivals_target: int):
        _check(lib().vkrt_sortfirst_wait(self._h, frame_index, arrivals_target))

    def mark(self, idx: int):
        _check(lib().vkrt_mark(self._h, idx))

    def mark_elapsed(self, a: int, b: int) -> float:
        ms = C.c_float()
        _check(lib().vkrt_mark_elapsed(self._h, a, b, C.byref(ms)))
        return float(ms.value)

    def sortfirst_join(self, rank: int, handle: bytes):
        h = (C.c_ubyte * 80).from_buffer_copy(handle)
        _check(lib().vkrt_sortfirst_join(self._h, rank, C.byref(h)))

    def sortfirst_leave(self):
        _check(lib().vkrt_sortfirst_leave(self._h))

    def sortfirst_render(self, cam: CameraUniform, offsets: np.ndarray, frame_index: int, uniform: Uniform | None = None):
        un = uniform if uniform is not None else self.global_uniform
        n = 0 if offsets is None else offsets.shape[0]
        _check(lib().vkrt_sortfirst_render(self._h, C.byref(cam), C.byref(un), _vp(offsets), n, frame_index))

    def sortfirst_render_batch(self, cams, first_frame: int, uniform: Uniform | None = None, flush_l2: bool = False):
        un = uniform if uniform is not None else self.global_uniform
        arr = self._cam_array(cams)
        _check(lib().vkrt_sortfirst_render_batch(self._h, C.cast(arr, C.c_void_p), len(cams), C.byref(un), first_frame, 1 if flush_l2 else 0))

    def sortfirst_render_tiles_batch(self, cams, offsets: np.ndarray, first_frame: int, uniform: Uniform | None = None):
        un = uniform if uniform is not None else self.global_uniform
        arr = self._cam_array(cams)
        _check(lib().vkrt_sortfirst_render_tiles_batch(self._h, C.cast(arr, C.c_void_p), len(cams), C.byref(un), _vp(offsets), offsets.shape[0], first_frame))

    def sortfirst_consume(self, frame_index: int, present: bool = False):
        _check(lib().vkrt_sortfirst_consume(self._h, frame_index, 1 if present else 0))

    def sortfirst_timeouts(self) -> int:
        v = C.c_uint64()
        _check(lib().vkrt_sortfirst_timeouts(self._h, C.byref(v)))
        return int(v.value)

    # -- sort-last: this context holds one brick of a partitioned grid ------------------------
    def upload_window(self, gn, own_lo, own_hi, *, scalar=None, color=None, normal=None):
        """The window array = the global grid restricted to [own_lo-1, own_hi+1) (clamped): pass that slice."""
        i3 = C.c_int * 3
        if scalar is not None:
            a = np.ascontiguousarray(scalar)
            dt = {np.dtype(np.uint8): abi.DTYPE_U8, np.dtype(np.float16): abi.DTYPE_F16, np.dtype(np.float32): abi.DTYPE_F32}[a.dtype]
            b = None
        else:
            a = np.ascontiguousarray(color).view(np.uint16)
            b = np.ascontiguousarray(normal).view(np.uint16)
            dt = -1
        _check(lib().vkrt_upload_window(self._h, _vp(a), _vp(b), dt, C.byref(i3(*gn)), C.byref(i3(*own_lo)), C.byref(i3(*own_hi))))

    def generate_synthetic_window(self, kind: int, dtype, gn, own_lo, own_hi, seed: int = 1):
        i3 = C.c_int * 3
        dt = {np.dtype(np.uint8): abi.DTYPE_U8, np.dtype(np.float16): abi.DTYPE_F16, np.dtype(np.float32): abi.DTYPE_F32}[np.dtype(dtype)]
        _check(lib().vkrt_generate_synthetic_window(self._h, kind, dt, C.byref(i3(*gn)), C.byref(i3(*own_lo)), C.byref(i3(*own_hi)), seed))

    def partial_alpha(self, cam: CameraUniform, d_T: int, uniform: Uniform | None = None):
        un = uniform if uniform is not None else self.global_uniform
        _check(lib().vkrt_partial_alpha(self._h, C.byref(cam), C.byref(un), d_T))

    def partial_ain(self, d_T_all: int, ranks_before, d_ain: int):
        rb = np.ascontiguousarray(np.asarray(list(ranks_before), np.int32))
        _check(lib().vkrt_partial_ain(self._h, d_T_all, _vp(rb) if rb.size else None, int(rb.size), d_ain))

    def partial_color(self, cam: CameraUniform, d_ain: int, d_rgba: int, uniform: Uniform | None = None):
        un = uniform if uniform is not None else self.global_uniform
        _check(lib().vkrt_partial_color(self._h, C.byref(cam), C.byref(un), d_ain, d_rgba))

    def partial_relative(self, cam: CameraUniform, d_rgba: int, d_T: int, uniform: Uniform | None = None):
        un = uniform if uniform is not None else self.global_uniform
        _check(lib().vkrt_partial_relative(self._h, C.byref(cam), C.byref(un), d_rgba, d_T))

    def partial_resolve(self, d_T_all: int, ranks_before, d_rgba: int, d_ain: int):
        rb = np.ascontiguousarray(np.asarray(list(ranks_before), np.int32))
        _check(lib().vkrt_partial_resolve(self._h, d_T_all, _vp(rb) if rb.size else None, int(rb.size), d_rgba, d_ain))

    # -- sort-last direct-send exchange of the transmittance images (one process per GPU) -------------------
    def exchange_create(self, rank: int, world: int) -> bytes:
        h = (C.c_ubyte * 80)()
        _check(lib().vkrt_exchange_create(self._h, rank, world, C.byref(h)))
        return bytes(h)

    def exchange_open(self, handles):
        """handles: the 80-byte blobs of exchange_create of ALL ranks, indexed by rank."""
        blob = (C.c_ubyte * (80 * len(handles))).from_buffer_copy(b"".join(handles))
        _check(lib().vkrt_exchange_open(self._h, C.byref(blob)))

    def exchange_push(self, d_T: int, ranks_behind, frame: int):
        rb = np.ascontiguousarray(np.asarray(list(ranks_behind), np.int32))
        _check(lib().vkrt_exchange_push(self._h, d_T, _vp(rb) if rb.size else None, int(rb.size), frame))

    def exchange_wait(self, frame: int, arrivals: int):
        _check(lib().vkrt_exchange_wait(self._h, frame, int(arrivals)))

    def exchange_table(self, frame: int) -> int:
        return int(lib().vkrt_exchange_table(self._h, frame) or 0)

    def exchange_done(self, frame: int):
        _check(lib().vkrt_exchange_done(self._h, frame))

    def exchange_timeouts(self) -> int:
        v = C.c_uint64(0)
        _check(lib().vkrt_exchange_timeouts(self._h, C.byref(v)))
        return int(v.value)

    def exchange_close(self):
        _check(lib().vkrt_exchange_close(self._h))

    def partial_finalize(self, cam: CameraUniform, d_sum: int, uniform: Uniform | None = None):
        un = uniform if uniform is not None else self.global_uniform
        _check(lib().vkrt_partial_finalize(self._h, C.byref(cam), C.byref(un), d_sum))

    def reset_stats(self):
        _check(lib().vkrt_reset_stats(self._h))

    @property
    def frame_device_ptr(self) -> int:
        return int(lib().vkrt_frame_device_ptr(self._h) or 0)

    @property
    def stream(self) -> int:
        return int(lib().vkrt_stream(self._h) or 0)


class PinnedArray:
    """A numpy view of page-locked host memory from vkrt_alloc_host (freed on close / garbage collection)."""

    def __init__(self, shape, dtype=np.uint8):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = C.c_void_p()
        _check(lib().vkrt_alloc_host(self.nbytes, C.byref(self._p)))
        buf = (C.c_ubyte * self.nbytes).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def close(self):
        if getattr(self, "_p", None) and self._p.value:
            self.array = None
            lib().vkrt_free_host(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SharedHostFrames:
    """n frames of RGBA8 [n, H, W, 4] in ONE host segment that every rank's process maps (POSIX shared memory created by
    rank 0) and page-locks (vkrt_host_register), so that each GPU delivers its frames into the consumer's memory over
    its own PCIe link: the gather to rank 0 of a sort-first sweep whose consumer lives on the host. `dist` is only the
    plumbing that ships the segment's name. register=False skips the page-locking (CPU tests of the plumbing)."""

    def __init__(self, rank: int, world: int, dist, n: int, height: int, width: int, register: bool = True):
        from multiprocessing import shared_memory

        self.rank, self.world, self.dist = rank, world, dist
        self.nbytes = n * height * width * 4
        box = [None]
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=self.nbytes)
            box[0] = self.shm.name
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=box[0])
            try:  # this process does not own the segment: keep Python's resource tracker from unlinking it at exit
                from multiprocessing import resource_tracker

                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.array = np.ndarray((n, height, width, 4), np.uint8, buffer=self.shm.buf)
        self._addr = self.array.ctypes.data
        self._registered = False
        if register:
            _check(lib().vkrt_host_register(C.c_void_p(self._addr), self.nbytes))
            self._registered = True
        if world > 1:
            dist.barrier()

    def close(self):
        if getattr(self, "shm", None) is None:
            return
        if self._registered:
            lib().vkrt_host_unregister(C.c_void_p(self._addr))
            self._registered = False
        self.array = None
        if self.world > 1:
            self.dist.barrier()
        self.shm.close()
        if self.rank == 0:
            self.shm.unlink()
        self.shm = None


def sortfirst_partition(width: int, height: int, tile_size: int, rank: int, world: int) -> np.ndarray:
    """This rank's tiles (float32 [n, 2] origins): the tiles intersecting the frame, dealt round-robin."""
    n = lib().vkrt_sortfirst_partition(width, height, tile_size, rank, world, None, 0)
    if n < 0:
        _check(n)
    out = np.zeros((n, 2), np.float32)
    lib().vkrt_sortfirst_partition(width, height, tile_size, rank, world, _vp(out), n)
    return out


def box_screen_hull(cam: CameraUniform, width: int, height: int) -> np.ndarray:
    """Six inward half-planes [a, b, c] (a*cx + b*cy + c >= 0 inside) of the box's silhouette, 2 px of margin (host-only)."""
    planes = np.zeros((6, 3), np.float32)
    _check(lib().vkrt_box_screen_hull(C.byref(cam), width, height, _vp(planes)))
    return planes


class RaycastPipeline:
    """examples/xor/raycast.rs `RaycastPipeline` with its entry point ("single" | "tile")."""

    def __init__(self, entry_point: str):
        if entry_point not in ("single", "tile"):
            raise VokselisError(abi.ERR_INVALID, f"unknown entry point: {entry_point}")
        self.entry_point = entry_point

    def record(self, ctx: Context, cam: CameraUniform):
        """What `Xor::render` records for this pipeline (examples/xor/main.rs:223-254)."""
        if self.entry_point == "single":
            ctx.render(cam)
        else:
            ctx.render_tiles(cam, tile_table(ctx.width, ctx.height, ctx.get_params().tile_size))

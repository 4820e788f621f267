"""ctypes binding of include/cloudsky.h.

The binding is backend-agnostic: the product loads ``csrc/libcloudsky_b200.so`` (hand-written
sm_100a CUDA); tests additionally load the CPU oracle through the very same binding so every
parity test calls both sides through the same C-ABI.  This module never falls back to anything:
if the CUDA library is missing, ``load_product()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(PKG_DIR, "csrc", "libcloudsky_b200.so")

CS_OK = 0
MODE_FAST = 0
MODE_STRICT = 1
MODE_EARLY_OUT = 2  # flag for MODE_FAST: stop a ray once T < 2^-12 (not reference behaviour)
MODE_TEX = 4  # flag for MODE_FAST: the texture unit filters the noise volumes / weather map (8-bit filter weights)
MODE_HALF = 8  # flag for MODE_FAST: the in-kernel filter runs in packed fp16 on the exact-integer records (11-bit weights)
TRANSMITTANCE_W, TRANSMITTANCE_H = 256, 64  # transmittance_lut.gd:6
SKY_LUT_W, SKY_LUT_H = 200, 100  # sky_lut.gd:4
REF_PRIMARY_STEPS, REF_CONE_SAMPLES = 128, 6  # clouds.glsl:228, :186


class CloudParams(C.Structure):
    """The 112-byte push-constant block (clouds.glsl:18-40, cloud_sky.gd:251-289)."""

    _fields_ = [
        ("texture_size", C.c_float * 2),
        ("update_position", C.c_float * 2),
        ("cloud_pos", C.c_float * 2),
        ("detailed_pos", C.c_float * 2),
        ("weather_pos", C.c_float * 2),
        ("pad1", C.c_float * 2),
        ("ground_color", C.c_float * 4),
        ("light_direction", C.c_float * 3),
        ("light_energy", C.c_float),
        ("light_color", C.c_float * 3),
        ("time", C.c_float),
        ("pad2", C.c_float),
        ("density", C.c_float),
        ("cloud_coverage", C.c_float),
        ("time_offset", C.c_float),
    ]

    def copy(self) -> "CloudParams":
        o = CloudParams()
        C.memmove(C.byref(o), C.byref(self), C.sizeof(self))
        return o

    def as_floats(self) -> np.ndarray:
        return np.frombuffer(bytes(self), dtype=np.float32).copy()

    @classmethod
    def from_floats(cls, arr: Sequence[float]) -> "CloudParams":
        a = np.asarray(arr, dtype=np.float32)
        assert a.size == 28
        o = cls()
        C.memmove(C.byref(o), a.ctypes.data, 112)
        return o


assert C.sizeof(CloudParams) == 112


class SkySettings(C.Structure):
    """Exported properties of cloud_sky.gd:4-50."""

    _fields_ = [
        ("wind_direction", C.c_float),
        ("wind_speed", C.c_float),
        ("density", C.c_float),
        ("cloud_coverage", C.c_float),
        ("time_offset", C.c_float),
        ("sun_disk_scale", C.c_float),
        ("ground_color", C.c_float * 4),
        ("frames_to_update", C.c_int32),
        ("texture_size", C.c_int32),
    ]


class FrameState(C.Structure):
    """FrameData's derived state (cloud_sky.gd:56-79)."""

    _fields_ = [
        ("time", C.c_float),
        ("cloud_pos", C.c_float * 2),
        ("detailed_pos", C.c_float * 2),
        ("weather_pos", C.c_float * 2),
        ("light_direction", C.c_float * 3),
        ("light_energy", C.c_float),
        ("light_color", C.c_float * 3),
    ]


class Counters(C.Structure):
    _fields_ = [
        ("marched_pixels", C.c_uint64),
        ("primary_steps", C.c_uint64),
        ("lit_steps", C.c_uint64),
        ("density_evals", C.c_uint64),
        ("large_fetches", C.c_uint64),
        ("small_fetches", C.c_uint64),
    ]

    def as_dict(self) -> dict:
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class NoiseParams(C.Structure):
    """cs_noise_params: knobs of the noise generator (README.md:30 TODO)."""
    _fields_ = [
        ("seed", C.c_uint32), ("worley_frequency", C.c_int32), ("worley_scale", C.c_float), ("perlin_frequency", C.c_int32), ("perlin_octaves", C.c_int32),
        ("perlin_scale", C.c_float), ("remap_lo", C.c_float), ("remap_hi", C.c_float), ("type_lo", C.c_float), ("type_hi", C.c_float),
    ]


NOISE_LARGE, NOISE_SMALL, NOISE_WEATHER = 0, 1, 2
TLUT_LINEAR, TLUT_BRUNETON2017 = 0, 1  # cs_set_transmittance_parametrisation


class View(C.Structure):
    """cs_view: which directions the presentation composite shades (clouds.gdshader EYEDIR)."""
    _fields_ = [
        ("projection", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("basis_columns", C.c_float * 9), ("fov_y_degrees", C.c_float),
        ("sun_direction", C.c_float * 3), ("sun_disk_scale", C.c_float), ("blend_amount", C.c_float),
    ]

    @classmethod
    def equirect(cls, width, height, sun_direction, sun_disk_scale=1.0, blend_amount=0.0):
        v = cls()
        v.projection, v.width, v.height = 0, width, height
        v.basis_columns[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1]
        v.fov_y_degrees = 75.0
        v.sun_direction[:] = [float(x) for x in sun_direction]
        v.sun_disk_scale, v.blend_amount = sun_disk_scale, blend_amount
        return v

    @classmethod
    def perspective(cls, width, height, basis_columns, fov_y_degrees, sun_direction, sun_disk_scale=1.0, blend_amount=0.0):
        v = cls.equirect(width, height, sun_direction, sun_disk_scale, blend_amount)
        v.projection = 1
        v.basis_columns[:] = [float(x) for x in basis_columns]
        v.fov_y_degrees = fov_y_degrees
        return v


class SkyFrame(C.Structure):
    """cs_sky_frame: the observable state of the Sky resource (cloud_sky.gd:82-94, sky_lut.gd:15-18)."""
    _fields_ = [
        ("frame", C.c_int32), ("frames_to_update", C.c_int32), ("texture_size", C.c_int32),
        ("update_position", C.c_int32 * 2), ("update_region_size", C.c_int32), ("num_workgroups", C.c_int32),
        ("texture_to_update", C.c_int32), ("texture_to_blend_from", C.c_int32), ("texture_to_blend_to", C.c_int32),
        ("blend_amount", C.c_float),
        ("sky_current_texture", C.c_int32), ("sky_blend_from", C.c_int32), ("sky_blend_to", C.c_int32), ("sky_updates", C.c_int32),
        ("cloud_textures", C.c_void_p * 3), ("sky_luts", C.c_void_p * 3),
        ("frame_data", FrameState),
    ]


# name -> (restype, argtypes); every symbol include/cloudsky.h declares.
_P = C.c_void_p
_PROTOTYPES = {
    "cs_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "cs_destroy": (None, [_P]),
    "cs_last_error": (C.c_char_p, [_P]),
    "cs_backend_name": (C.c_char_p, []),
    "cs_set_stream": (C.c_int, [_P, _P]),
    "cs_sync": (C.c_int, [_P]),
    "cs_set_threads": (C.c_int, [_P, C.c_int]),
    "cs_upload_textures": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int]),
    "cs_load_texture_files": (C.c_int, [_P, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p]),
    "cs_decode_image_file": (C.c_int, [C.c_char_p, C.POINTER(_P), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cs_free": (None, [_P]),
    "cs_read_volume_level": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_size_t]),
    "cs_noise_params_default": (None, [C.c_int, C.POINTER(NoiseParams)]),
    "cs_generate_noise": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(NoiseParams), _P, C.c_size_t]),
    "cs_set_transmittance_parametrisation": (C.c_int, [_P, C.c_int]),
    "cs_build_transmittance_lut": (C.c_int, [_P]),
    "cs_build_sky_lut": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "cs_read_transmittance_lut": (C.c_int, [_P, _P, C.c_size_t]),
    "cs_read_sky_lut": (C.c_int, [_P, _P, C.c_size_t]),
    "cs_write_transmittance_lut": (C.c_int, [_P, _P, C.c_size_t]),
    "cs_write_sky_lut": (C.c_int, [_P, _P, C.c_size_t]),
    "cs_resize": (C.c_int, [_P, C.c_int, C.c_int]),
    "cs_set_march_config": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "cs_set_step_budget": (C.c_int, [_P, C.c_float, C.c_int]),
    "cs_set_counters_enabled": (C.c_int, [_P, C.c_int]),
    "cs_get_counters": (C.c_int, [_P, C.POINTER(Counters)]),
    "cs_dispatch_clouds": (C.c_int, [_P, C.POINTER(CloudParams), C.c_int, C.c_int]),
    "cs_render_frame": (C.c_int, [_P, C.POINTER(CloudParams)]),
    "cs_render_rows_to": (C.c_int, [_P, C.POINTER(CloudParams), C.c_int, C.c_int, _P]),
    "cs_render_row_bands_to": (C.c_int, [_P, C.POINTER(CloudParams), C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "cs_image_device_ptr": (_P, [_P]),
    "cs_read_image": (C.c_int, [_P, _P, C.c_size_t]),
    "cs_render_frame_host": (C.c_int, [_P, C.POINTER(CloudParams), _P, C.c_size_t]),
    "cs_render_frame_host_async": (C.c_int, [_P, C.POINTER(CloudParams), _P, C.c_size_t]),
    "cs_wait_host": (C.c_int, [_P]),
    "cs_render_sun_batch_to": (C.c_int, [_P, C.POINTER(CloudParams), C.POINTER(C.c_float), C.c_int, _P]),
    "cs_time_render_frame": (C.c_int, [_P, C.POINTER(CloudParams), C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "cs_peer_alloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P), C.c_char_p]),
    "cs_peer_open": (C.c_int, [_P, C.c_char_p, C.POINTER(_P)]),
    "cs_peer_close": (C.c_int, [_P, _P]),
    "cs_peer_free": (C.c_int, [_P, _P]),
    "cs_set_output_mirrors": (C.c_int, [_P, _P, C.c_size_t, C.c_int, C.POINTER(_P)]),
    "cs_peer_barrier": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P), C.c_uint32]),
    "cs_peer_check": (C.c_int, [_P]),
    "cs_sky_create": (C.c_int, [_P, C.POINTER(SkySettings), C.POINTER(_P)]),
    "cs_sky_destroy": (None, [_P]),
    "cs_sky_set_settings": (C.c_int, [_P, C.POINTER(SkySettings)]),
    "cs_sky_set_sun": (C.c_int, [_P, C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float)]),
    "cs_sky_update": (C.c_int, [_P, C.c_float]),
    "cs_sky_get_frame": (C.c_int, [_P, C.POINTER(SkyFrame)]),
    "cs_sky_read_texture": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "cs_composite": (C.c_int, [_P, C.POINTER(View), _P, _P, C.c_int, C.c_int, _P, _P, _P]),
    "cs_sky_composite_host": (C.c_int, [_P, C.POINTER(View), _P, C.c_size_t]),
    "cs_set_kernel_timing": (C.c_int, [_P, C.c_int]),
    "cs_read_kernel_timings": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "cs_settings_default": (None, [C.POINTER(SkySettings)]),
    "cs_settings_demo": (None, [C.POINTER(SkySettings)]),
    "cs_frame_state_init": (None, [C.POINTER(FrameState)]),
    "cs_frame_state_set_light": (None, [C.POINTER(FrameState), C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float)]),
    "cs_frame_advance": (None, [C.POINTER(FrameState), C.POINTER(SkySettings), C.c_float]),
    "cs_fill_cloud_params": (None, [C.POINTER(CloudParams), C.POINTER(SkySettings), C.POINTER(FrameState), C.c_int, C.c_int, C.c_int, C.c_int]),
    "cs_update_performance": (None, [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cs_next_update_position": (None, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


class CloudSkyError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cloudsky error {code}: {msg}")
        self.code = code


class Library:
    """One loaded implementation of include/cloudsky.h."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found — build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                "There is no CPU fallback for the product path.")
        self.path = path
        self.dll = C.CDLL(path, mode=C.RTLD_LOCAL)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(self.dll, name)  # raises AttributeError when a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
        self.backend = self.dll.cs_backend_name().decode()

    def context(self, device: int = 0) -> "Context":
        return Context(self, device)

    # host-side parameter logic -------------------------------------------------------------
    def settings_default(self) -> SkySettings:
        s = SkySettings()
        self.dll.cs_settings_default(C.byref(s))
        return s

    def settings_demo(self) -> SkySettings:
        s = SkySettings()
        self.dll.cs_settings_demo(C.byref(s))
        return s

    def noise_params_default(self, kind: int) -> NoiseParams:
        p = NoiseParams()
        self.dll.cs_noise_params_default(kind, C.byref(p))
        return p

    def frame_state_init(self) -> FrameState:
        st = FrameState()
        self.dll.cs_frame_state_init(C.byref(st))
        return st

    def frame_state_set_light(self, st: FrameState, basis_columns, energy: float, color_srgb) -> None:
        b = (C.c_float * 9)(*[float(v) for v in basis_columns])
        c = (C.c_float * 3)(*[float(v) for v in color_srgb])
        self.dll.cs_frame_state_set_light(C.byref(st), b, float(energy), c)

    def frame_advance(self, st: FrameState, s: SkySettings, abs_time: float) -> None:
        self.dll.cs_frame_advance(C.byref(st), C.byref(s), float(abs_time))

    def fill_cloud_params(self, s: SkySettings, st: FrameState, w: int, h: int, ux: int = 0, uy: int = 0) -> CloudParams:
        p = CloudParams()
        self.dll.cs_fill_cloud_params(C.byref(p), C.byref(s), C.byref(st), w, h, ux, uy)
        return p

    def update_performance(self, texture_size: int, frames_to_update: int):
        ts, region, groups = C.c_int(texture_size), C.c_int(0), C.c_int(0)
        self.dll.cs_update_performance(C.byref(ts), frames_to_update, C.byref(region), C.byref(groups))
        return ts.value, region.value, groups.value

    def next_update_position(self, x: int, y: int, region: int, texture_size: int):
        cx, cy = C.c_int(x), C.c_int(y)
        self.dll.cs_next_update_position(C.byref(cx), C.byref(cy), region, texture_size)
        return cx.value, cy.value

    def decode_image_file(self, path: str) -> np.ndarray:
        ptr, w, h, ch = _P(), C.c_int(), C.c_int(), C.c_int()
        r = self.dll.cs_decode_image_file(path.encode(), C.byref(ptr), C.byref(w), C.byref(h), C.byref(ch))
        if r != CS_OK:
            raise CloudSkyError(r, f"cs_decode_image_file({path})")
        n = w.value * h.value * ch.value
        out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,)).copy().reshape(h.value, w.value, ch.value)
        self.dll.cs_free(ptr)
        return out


def _u8(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint8, a.dtype
    return a


class Context:
    """One cs_context (one device + one stream)."""

    def __init__(self, lib: Library, device: int = 0):
        self.lib = lib
        self._h = _P()
        r = lib.dll.cs_create(device, C.byref(self._h))
        if r != CS_OK or not self._h:
            raise CloudSkyError(r, f"cs_create(device={device}) failed on backend {lib.backend}")
        self.width = self.height = 0
        self._skies = []  # Sky objects created on this context

    def close(self) -> None:
        if getattr(self, "_h", None):
            for sky in list(getattr(self, "_skies", [])):
                sky.close()
            self.lib.dll.cs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, r: int) -> None:
        if r != CS_OK:
            msg = self.lib.dll.cs_last_error(self._h)
            raise CloudSkyError(r, msg.decode() if msg else "")

    # plumbing -----------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int) -> None:
        self._ck(self.lib.dll.cs_set_stream(self._h, _P(cuda_stream)))

    def sync(self) -> None:
        self._ck(self.lib.dll.cs_sync(self._h))

    def set_threads(self, n: int) -> None:
        self._ck(self.lib.dll.cs_set_threads(self._h, n))

    # textures -----------------------------------------------------------------------------
    def upload_textures(self, large: np.ndarray, small: np.ndarray, weather: np.ndarray) -> None:
        """large [z][y][x][3|4] u8, small [z][y][x][3|4] u8, weather [y][x][3|4] u8."""
        large, small, weather = _u8(large), _u8(small), _u8(weather)
        assert large.ndim == 4 and small.ndim == 4 and weather.ndim == 3
        self._ck(self.lib.dll.cs_upload_textures(
            self._h, large.ctypes.data, large.shape[0], large.shape[3], small.ctypes.data, small.shape[0], small.shape[3],
            weather.ctypes.data, weather.shape[1], weather.shape[0], weather.shape[2]))

    def load_texture_files(self, large_path: str, large_slices: int, small_path: str, small_slices: int, weather_path: str) -> None:
        self._ck(self.lib.dll.cs_load_texture_files(self._h, large_path.encode(), large_slices, small_path.encode(),
                                                    small_slices, weather_path.encode()))

    def read_volume_level(self, which: int, n0: int, level: int) -> np.ndarray:
        n = n0 >> level
        out = np.empty((n, n, n, 4), np.uint8)
        self._ck(self.lib.dll.cs_read_volume_level(self._h, which, level, out.ctypes.data, out.nbytes))
        return out

    def generate_noise(self, kind: int, n: int, params: "NoiseParams | None" = None) -> np.ndarray:
        """cs_generate_noise: RGBA8 [n,n,n,4] (z,y,x order, as upload_textures takes it) or [n,n,4] for the weather map."""
        p = params if params is not None else self.lib.noise_params_default(kind)
        out = np.empty((n, n, 4) if kind == NOISE_WEATHER else (n, n, n, 4), np.uint8)
        self._ck(self.lib.dll.cs_generate_noise(self._h, kind, n, C.byref(p), out.ctypes.data, out.nbytes))
        return out

    # LUTs ---------------------------------------------------------------------------------
    def set_transmittance_parametrisation(self, which: int) -> None:
        self._ck(self.lib.dll.cs_set_transmittance_parametrisation(self._h, which))

    def build_transmittance_lut(self) -> None:
        self._ck(self.lib.dll.cs_build_transmittance_lut(self._h))

    def build_sky_lut(self, sun_direction) -> None:
        s = (C.c_float * 3)(*[float(v) for v in sun_direction])
        self._ck(self.lib.dll.cs_build_sky_lut(self._h, s))

    def read_transmittance_lut(self) -> np.ndarray:
        out = np.empty((TRANSMITTANCE_H, TRANSMITTANCE_W, 4), np.float16)
        self._ck(self.lib.dll.cs_read_transmittance_lut(self._h, out.ctypes.data, out.nbytes))
        return out

    def read_sky_lut(self) -> np.ndarray:
        out = np.empty((SKY_LUT_H, SKY_LUT_W, 4), np.float16)
        self._ck(self.lib.dll.cs_read_sky_lut(self._h, out.ctypes.data, out.nbytes))
        return out

    def write_transmittance_lut(self, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a, dtype=np.float16)
        self._ck(self.lib.dll.cs_write_transmittance_lut(self._h, a.ctypes.data, a.nbytes))

    def write_sky_lut(self, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a, dtype=np.float16)
        self._ck(self.lib.dll.cs_write_sky_lut(self._h, a.ctypes.data, a.nbytes))

    # clouds -------------------------------------------------------------------------------
    def resize(self, width: int, height: int) -> None:
        self._ck(self.lib.dll.cs_resize(self._h, width, height))
        self.width, self.height = width, height

    def set_march_config(self, primary_steps: int = REF_PRIMARY_STEPS, cone_samples: int = REF_CONE_SAMPLES, mode: int = MODE_FAST) -> None:
        self._ck(self.lib.dll.cs_set_march_config(self._h, primary_steps, cone_samples, mode))

    def set_step_budget(self, min_step_length_m: float = 0.0, min_steps: int = 1) -> None:
        """Adaptive per-direction primary step count (0 = the reference's fixed count); see include/cloudsky.h."""
        self._ck(self.lib.dll.cs_set_step_budget(self._h, float(min_step_length_m), int(min_steps)))

    def set_counters_enabled(self, on: bool) -> None:
        self._ck(self.lib.dll.cs_set_counters_enabled(self._h, int(on)))

    def get_counters(self) -> Counters:
        k = Counters()
        self._ck(self.lib.dll.cs_get_counters(self._h, C.byref(k)))
        return k

    def dispatch_clouds(self, params: CloudParams, groups_x: int, groups_y: int) -> None:
        self._ck(self.lib.dll.cs_dispatch_clouds(self._h, C.byref(params), groups_x, groups_y))

    def render_frame(self, params: CloudParams) -> None:
        self._ck(self.lib.dll.cs_render_frame(self._h, C.byref(params)))

    def render_rows_to(self, params: CloudParams, row_begin: int, row_end: int, device_ptr: int) -> None:
        self._ck(self.lib.dll.cs_render_rows_to(self._h, C.byref(params), row_begin, row_end, _P(device_ptr)))

    def render_row_bands_to(self, params: CloudParams, first_row: int, band_rows: int, band_pitch_rows: int, n_bands: int, device_ptr: int) -> None:
        self._ck(self.lib.dll.cs_render_row_bands_to(self._h, C.byref(params), first_row, band_rows, band_pitch_rows, n_bands, _P(device_ptr)))

    def image_device_ptr(self) -> int:
        return int(self.lib.dll.cs_image_device_ptr(self._h) or 0)

    def read_image(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), np.float16)
        self._ck(self.lib.dll.cs_read_image(self._h, out.ctypes.data, out.nbytes))
        return out

    def render_frame_host(self, params: CloudParams, out: Optional[np.ndarray] = None, out_ptr: Optional[int] = None) -> Optional[np.ndarray]:
        nbytes = self.width * self.height * 8
        if out_ptr is not None:
            self._ck(self.lib.dll.cs_render_frame_host(self._h, C.byref(params), _P(out_ptr), nbytes))
            return None
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float16)
        assert out.nbytes == nbytes and out.flags.c_contiguous
        self._ck(self.lib.dll.cs_render_frame_host(self._h, C.byref(params), out.ctypes.data, nbytes))
        return out

    def render_frame_host_async(self, params: CloudParams, out_ptr: int) -> None:
        """Queue sky LUT + march + D2H into host memory at out_ptr (valid after wait_host())."""
        self._ck(self.lib.dll.cs_render_frame_host_async(self._h, C.byref(params), _P(out_ptr), self.width * self.height * 8))

    def wait_host(self) -> None:
        self._ck(self.lib.dll.cs_wait_host(self._h))

    def render_sun_batch_to(self, params: CloudParams, sun_dirs: np.ndarray, device_ptr: int) -> None:
        s = np.ascontiguousarray(sun_dirs, dtype=np.float32).reshape(-1, 3)
        self._ck(self.lib.dll.cs_render_sun_batch_to(self._h, C.byref(params), s.ctypes.data_as(C.POINTER(C.c_float)), s.shape[0], _P(device_ptr)))

    # ---- multi-GPU: peer-mapped output replicas (fused all-gather) ----
    def peer_alloc(self, nbytes: int):
        """-> (device pointer, 64-byte IPC handle) of a zero-filled exportable buffer."""
        ptr = _P()
        handle = C.create_string_buffer(64)
        self._ck(self.lib.dll.cs_peer_alloc(self._h, nbytes, C.byref(ptr), handle))
        return ptr.value, handle.raw

    def peer_open(self, handle: bytes) -> int:
        ptr = _P()
        self._ck(self.lib.dll.cs_peer_open(self._h, handle, C.byref(ptr)))
        return ptr.value

    def peer_close(self, ptr: int) -> None:
        self._ck(self.lib.dll.cs_peer_close(self._h, _P(ptr)))

    def peer_free(self, ptr: int) -> None:
        self._ck(self.lib.dll.cs_peer_free(self._h, _P(ptr)))

    def set_output_mirrors(self, base: int, nbytes: int, mirror_bases) -> None:
        arr = (_P * max(1, len(mirror_bases)))(*mirror_bases)
        self._ck(self.lib.dll.cs_set_output_mirrors(self._h, _P(base), nbytes, len(mirror_bases), arr))

    def peer_barrier(self, rank: int, world: int, flag_arrays, epoch: int) -> None:
        arr = (_P * world)(*flag_arrays)
        self._ck(self.lib.dll.cs_peer_barrier(self._h, rank, world, arr, epoch & 0xFFFFFFFF))

    def peer_check(self) -> None:
        self._ck(self.lib.dll.cs_peer_check(self._h))

    def set_kernel_timing(self, on: bool) -> None:
        self._ck(self.lib.dll.cs_set_kernel_timing(self._h, int(on)))

    def read_kernel_timings(self) -> dict:
        m, s = C.c_float(), C.c_float()
        nm, ns = C.c_int(), C.c_int()
        self._ck(self.lib.dll.cs_read_kernel_timings(self._h, C.byref(m), C.byref(nm), C.byref(s), C.byref(ns)))
        return {"march_ms": float(m.value), "march_launches": nm.value, "sky_ms": float(s.value), "sky_launches": ns.value}

    def time_render_frame(self, params: CloudParams, warmup: int, iters: int) -> float:
        ms = C.c_float()
        self._ck(self.lib.dll.cs_time_render_frame(self._h, C.byref(params), warmup, iters, C.byref(ms)))
        return float(ms.value)


class Sky:
    """cs_sky: the time-sliced Sky resource (cloud_sky.gd's update_sky state machine inside the library)."""

    def __init__(self, ctx: Context, settings: SkySettings):
        self.ctx = ctx
        self._h = _P()
        ctx._ck(ctx.lib.dll.cs_sky_create(ctx._h, C.byref(settings), C.byref(self._h)))
        ctx._skies.append(self)  # Context.close() destroys its skies first (a cs_sky keeps a pointer to its context)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.ctx.lib.dll.cs_sky_destroy(self._h)
            self._h = None
            if self in self.ctx._skies:
                self.ctx._skies.remove(self)

    def set_settings(self, settings: SkySettings) -> None:
        self.ctx._ck(self.ctx.lib.dll.cs_sky_set_settings(self._h, C.byref(settings)))

    def set_sun(self, basis_columns, energy: float, color_srgb) -> None:
        b = (C.c_float * 9)(*[float(v) for v in basis_columns])
        c = (C.c_float * 3)(*[float(v) for v in color_srgb])
        self.ctx._ck(self.ctx.lib.dll.cs_sky_set_sun(self._h, b, float(energy), c))

    def update(self, now: float) -> None:
        self.ctx._ck(self.ctx.lib.dll.cs_sky_update(self._h, float(now)))

    def frame(self) -> SkyFrame:
        f = SkyFrame()
        self.ctx._ck(self.ctx.lib.dll.cs_sky_get_frame(self._h, C.byref(f)))
        return f

    def composite(self, view: "View") -> np.ndarray:
        """clouds.gdshader for every pixel of `view` with this sky's blend textures; float32 [H, W, 4] linear radiance."""
        out = np.empty((view.height, view.width, 4), np.float32)
        self.ctx._ck(self.ctx.lib.dll.cs_sky_composite_host(self._h, C.byref(view), out.ctypes.data, out.nbytes))
        return out

    def read_texture(self, index: int) -> np.ndarray:
        n = self.frame().texture_size
        out = np.empty((n, n, 4), np.float16)
        self.ctx._ck(self.ctx.lib.dll.cs_sky_read_texture(self._h, index, out.ctypes.data, out.nbytes))
        return out


_product: Optional[Library] = None


def load_product() -> Library:
    """The CUDA library.  Raises if it has not been built — no fallback."""
    global _product
    if _product is None:
        _product = Library(PRODUCT_LIB)
        if _product.backend != "cuda-sm100a":
            raise RuntimeError(f"{PRODUCT_LIB} reports backend {_product.backend!r}, expected 'cuda-sm100a'")
    return _product

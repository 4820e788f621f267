"""Host-side mirror of the reference's GDScript operator interface for the hot path
(cloud_sky/cloud_sky.gd, sky_lut.gd, transmittance_lut.gd, sun.gd): same property names, the same
defaults and side effects, the same call sequence — with Godot's RenderingDevice dispatches replaced by
calls through the C-ABI (include/cloudsky.h).  Time is passed in (the reference reads
Time.get_ticks_msec(), cloud_sky.gd:174) so that every frame is reproducible.

    sky = CloudSky(lib, device=0)               # load("clouds_sky.tres")
    sky.load_textures(large, small, weather)    # preload(...) of the three bitmaps
    sky.sun = DirectionalLight(basis=..., light_energy=1.0, light_color=(1, 1, 1))   # sun.gd:11-13
    sky.update_sky(now_seconds)                 # frame_pre_draw -> update_sky (cloud_sky.gd:107,129)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import capi


@dataclass
class DirectionalLight:
    """What sun.gd hands to the sky: transform basis (three column vectors), energy, sRGB colour."""
    basis: Sequence[float] = (1, 0, 0, 0, 0, -1, 0, 1, 0)  # columns x, y, z; z column = direction toward the sun
    light_energy: float = 1.0
    light_color: Sequence[float] = (1.0, 1.0, 1.0)

    @classmethod
    def looking_from(cls, direction: Sequence[float], **kw) -> "DirectionalLight":
        """A light whose basis z column is `direction` (toward the sun)."""
        d = np.asarray(direction, np.float64)
        d = d / np.linalg.norm(d)
        up = np.array([0.0, 1.0, 0.0]) if abs(d[1]) < 0.99 else np.array([1.0, 0.0, 0.0])
        x = np.cross(up, d); x /= np.linalg.norm(x)
        y = np.cross(d, x)
        return cls(basis=tuple(x) + tuple(y) + tuple(d), **kw)


class TransmittanceLUT:
    """transmittance_lut.gd: a 256x64 RGBA16F texture generated once when loaded."""
    texture_size = (capi.TRANSMITTANCE_W, capi.TRANSMITTANCE_H)  # transmittance_lut.gd:6

    def __init__(self, ctx: capi.Context, parametrisation: int = capi.TLUT_LINEAR):
        self.ctx = ctx
        self.parametrisation = parametrisation  # TLUT_BRUNETON2017: the mapping the reference's README lists as a TODO (README.md:29)
        self._initialize_compute_code()

    def _initialize_compute_code(self) -> None:  # transmittance_lut.gd:51-78: one dispatch of 32x8 groups
        self.ctx.set_transmittance_parametrisation(self.parametrisation)
        self.ctx.build_transmittance_lut()

    def read(self) -> np.ndarray:
        return self.ctx.read_transmittance_lut()


class SkyLUT:
    """sky_lut.gd: a 200x100 RGBA16F sky-view LUT re-rendered for every sun direction."""
    texture_size = (capi.SKY_LUT_W, capi.SKY_LUT_H)  # sky_lut.gd:4

    def __init__(self, ctx: capi.Context, transmittance: Optional[TransmittanceLUT]):
        self.ctx = ctx
        self.light_direction = (0.0, -1.0, 0.0)  # sky_lut.gd:5
        self.needs_update = True
        self.needs_full_update = True  # sky_lut.gd:10
        self.initialized = transmittance is not None  # sky_lut.gd:7,120
        self.current_texture = 0
        self.updates = 0

    def request_update(self) -> None:  # sky_lut.gd:39-40
        self.needs_update = True

    def update_lut(self, sun_direction: Sequence[float]) -> None:  # sky_lut.gd:43-52
        self.light_direction = tuple(float(v) for v in sun_direction)
        if not self.initialized:
            print("Attempting to update uninitialized sky lut")
            return
        self.render_lut()
        if self.needs_full_update:  # fill all three copies the first time (sky_lut.gd:49-52)
            self.render_lut()
            self.render_lut()
            self.needs_full_update = False

    def render_lut(self) -> None:  # sky_lut.gd:122-148
        self.ctx.build_sky_lut(self.light_direction)
        self.current_texture = (self.current_texture + 1) % 3
        self.updates += 1
        self.needs_update = False

    def read(self) -> np.ndarray:
        return self.ctx.read_sky_lut()


class CloudSky:
    """cloud_sky.gd: exported cloud / sky / performance settings, FrameData, the tile walk of update_sky
    and the push-constant packing.  `frames_to_update = 1` renders the whole texture in one dispatch."""

    def __init__(self, lib: capi.Library, device: int = 0, demo_values: bool = True):
        self.lib = lib
        self.ctx = lib.context(device)
        s = lib.settings_demo() if demo_values else lib.settings_default()  # clouds_sky.tres:11-18 / cloud_sky.gd:4-50
        self._s = s
        self.frame_data = lib.frame_state_init()  # FrameData (cloud_sky.gd:56-79)
        self.sun: Optional[DirectionalLight] = None
        self.update_position = [0, 0]
        self.update_region_size = 96
        self.num_workgroups = 12
        self.texture_to_update, self.texture_to_blend_from, self.texture_to_blend_to = 0, 1, 2
        self.textures = [None, None, None]  # finished hemisphere textures (host copies), rotated like cloud_sky.gd:137-141
        self.frame = 0
        self.blend_amount = 0.0
        self.can_run = False
        self.needs_full_sky_init = True
        self._params = None
        self.transmittance_tex = TransmittanceLUT(self.ctx)      # cloud_sky.gd:92
        self.sky_lut = SkyLUT(self.ctx, self.transmittance_tex)  # cloud_sky.gd:91
        self._have_textures = False
        self.update_performance()

    # ---- exported properties (cloud_sky.gd:4-50) ------------------------------------------------
    wind_direction = property(lambda self: self._s.wind_direction, lambda self, v: setattr(self._s, "wind_direction", float(v)))
    wind_speed = property(lambda self: self._s.wind_speed, lambda self, v: setattr(self._s, "wind_speed", float(v)))
    density = property(lambda self: self._s.density, lambda self, v: setattr(self._s, "density", float(v)))
    cloud_coverage = property(lambda self: self._s.cloud_coverage, lambda self, v: setattr(self._s, "cloud_coverage", float(v)))
    time_offset = property(lambda self: self._s.time_offset, lambda self, v: setattr(self._s, "time_offset", float(v)))
    sun_disk_scale = property(lambda self: self._s.sun_disk_scale, lambda self, v: setattr(self._s, "sun_disk_scale", float(v)))

    @property
    def ground_color(self):
        return tuple(self._s.ground_color)

    @ground_color.setter
    def ground_color(self, c):
        c = tuple(c) + (1.0,) * (4 - len(tuple(c)))
        self._s.ground_color[:] = [float(v) for v in c]

    @property
    def frames_to_update(self) -> int:
        return self._s.frames_to_update

    @frames_to_update.setter
    def frames_to_update(self, v: int) -> None:  # cloud_sky.gd:37-42: cleanup + update_performance + full init
        self._s.frames_to_update = int(v)
        self.cleanup()
        self.update_performance()
        self.request_full_sky_init()

    @property
    def texture_size(self) -> int:
        return self._s.texture_size

    @texture_size.setter
    def texture_size(self, v: int) -> None:  # cloud_sky.gd:45-50
        self._s.texture_size = int(v)
        self.cleanup()
        self.update_performance()
        self.request_full_sky_init()

    # ---- setup ------------------------------------------------------------------------------------
    def load_textures(self, large: np.ndarray, small: np.ndarray, weather: np.ndarray) -> None:
        self.ctx.upload_textures(large, small, weather)  # _create_noise_uniform_set (cloud_sky.gd:298-341)
        self._have_textures = True
        self.can_run = True

    def generate_textures(self, seed: int = 1, large_n: int = 128, small_n: int = 32, weather_n: int = 512) -> None:
        """Synthesise the three inputs on the device instead of loading the bitmaps (the reference's TODO, README.md:30)."""
        tex = []
        for kind, n in ((capi.NOISE_LARGE, large_n), (capi.NOISE_SMALL, small_n), (capi.NOISE_WEATHER, weather_n)):
            p = self.lib.noise_params_default(kind)
            p.seed = seed
            tex.append(self.ctx.generate_noise(kind, n, p))
        self.load_textures(*tex)

    def load_texture_files(self, directory: str) -> None:  # preload("perlworlnoise.tga") ... (cloud_sky.gd:311,321,331)
        import os
        self.ctx.load_texture_files(os.path.join(directory, "perlworlnoise.tga"), 128, os.path.join(directory, "worlnoise.bmp"), 32,
                                    os.path.join(directory, "weather.bmp"))
        self._have_textures = True
        self.can_run = True

    def update_performance(self) -> None:  # cloud_sky.gd:109-118
        ts, region, groups = self.lib.update_performance(self._s.texture_size, self._s.frames_to_update)
        if ts != self._s.texture_size:
            print("texture_size is not a multiple of sqrt(frames_to_update), changing to: ", ts)
            self._s.texture_size = ts
        self.update_region_size, self.num_workgroups = region, groups
        self.ctx.resize(ts, ts)  # _initialize_compute_code (cloud_sky.gd:355-408): the output texture
        self.can_run = self._have_textures

    def request_full_sky_init(self) -> None:  # cloud_sky.gd:120-121
        self.needs_full_sky_init = True

    def cleanup(self) -> None:  # cloud_sky.gd:197-212
        self.can_run = False
        self.frame = 0
        self.texture_to_update, self.texture_to_blend_from, self.texture_to_blend_to = 0, 1, 2
        self.update_position = [0, 0]
        self.textures = [None, None, None]

    # ---- per-frame path ---------------------------------------------------------------------------
    def initialize_sky(self, now: float) -> None:  # cloud_sky.gd:124-127
        self._update_per_frame_data(now)
        for _ in range(self.frames_to_update * 2):
            self.update_sky(now)

    def update_sky(self, now: float) -> None:  # cloud_sky.gd:129-163
        if not self.can_run:
            return
        if self.needs_full_sky_init:
            self.needs_full_sky_init = False
            self.initialize_sky(now)
        if self.frame >= self.frames_to_update:
            self.textures[self.texture_to_update] = self.ctx.read_image()  # the texture just completed
            self.texture_to_update = (self.texture_to_update + 1) % 3
            self.texture_to_blend_from = (self.texture_to_blend_from + 1) % 3
            self.texture_to_blend_to = (self.texture_to_blend_to + 1) % 3
            self._update_per_frame_data(now)  # only once per full texture, otherwise tiles get out of sync
            self.frame = 0
        self.blend_amount = float(self.frame) / float(self.frames_to_update)
        self._render_process(self.texture_to_update)
        self.update_position[0], self.update_position[1] = self.lib.next_update_position(
            self.update_position[0], self.update_position[1], self.update_region_size, self._s.texture_size)
        self.frame += 1

    def _update_per_frame_data(self, now: float) -> None:  # cloud_sky.gd:165-187
        if self.sun is not None:
            self.lib.frame_state_set_light(self.frame_data, self.sun.basis, self.sun.light_energy, self.sun.light_color)
        self.lib.frame_advance(self.frame_data, self._s, now)
        self.sky_lut.update_lut(tuple(self.frame_data.light_direction))

    def _fill_push_constant(self) -> capi.CloudParams:  # cloud_sky.gd:251-289
        ts = self._s.texture_size
        return self.lib.fill_cloud_params(self._s, self.frame_data, ts, ts, self.update_position[0], self.update_position[1])

    def _render_process(self, texture_to_update: int) -> None:  # cloud_sky.gd:234-248
        self._params = self._fill_push_constant()
        self.ctx.dispatch_clouds(self._params, self.num_workgroups, self.num_workgroups)

    # ---- convenience beyond the reference ----------------------------------------------------------
    def render_full(self, now: float, width: Optional[int] = None, height: Optional[int] = None) -> np.ndarray:
        """north_star's single dispatch: advance the frame data to `now`, refresh the sky LUT and render the
        whole (optionally non-square) hemisphere texture in one launch; returns the RGBA16F image."""
        w = width or self._s.texture_size
        h = height or self._s.texture_size
        if (self.ctx.width, self.ctx.height) != (w, h):
            self.ctx.resize(w, h)
        self._update_per_frame_data(now)
        p = self.lib.fill_cloud_params(self._s, self.frame_data, w, h, 0, 0)
        self.ctx.render_frame(p)
        return self.ctx.read_image()

    def close(self) -> None:
        self.ctx.close()

"""B200-native volumetric-cloud sky renderer — the hot path of
clayjohn/godot-volumetric-cloud-demo-v2 (clouds.glsl + sky-lut.glsl + transmittance-lut.glsl and the
cloud_sky.gd parameter surface) as hand-written sm_100a CUDA behind the C-ABI of include/cloudsky.h.

Import name: ``cloudsky_b200`` (see cloudsky_b200.py at the repo root; the directory name
``godot-volumetric-cloud-demo-v2_b200`` is not a valid Python identifier).
"""
from . import capi  # noqa: F401
from .capi import (CloudParams, CloudSkyError, Context, Counters, FrameState, Library, NoiseParams, Sky, SkyFrame, SkySettings, View,  # noqa: F401
                   NOISE_LARGE, NOISE_SMALL, NOISE_WEATHER, TLUT_BRUNETON2017, TLUT_LINEAR,
                   MODE_EARLY_OUT, MODE_FAST, MODE_HALF, MODE_STRICT, MODE_TEX, load_product)

from .sky import CloudSky, DirectionalLight, SkyLUT, TransmittanceLUT  # noqa: F401,E402

__all__ = ["CloudSky", "DirectionalLight", "SkyLUT", "TransmittanceLUT", "capi", "CloudParams", "CloudSkyError", "Context", "Counters", "FrameState", "Library", "NoiseParams", "NOISE_LARGE", "NOISE_SMALL", "NOISE_WEATHER", "TLUT_LINEAR", "TLUT_BRUNETON2017", "Sky", "SkyFrame", "SkySettings", "View",
           "MODE_EARLY_OUT", "MODE_FAST", "MODE_HALF", "MODE_STRICT", "MODE_TEX", "load_product"]

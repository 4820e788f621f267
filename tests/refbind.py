"""ctypes binding of oracle/_ref/libcloudsky_ref.so — the REFERENCE's own GLSL compiled by g++ (oracle/build_ref.sh).

TEST INFRASTRUCTURE.  Imported only by tests/, tests/golden/make_ref_golden.py and bench.py's CPU-baseline /
``--impl reference`` legs.  The product never loads it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_LIB = os.path.join(REF_DIR, "libcloudsky_ref.so")
REFERENCE_SHADERS = "/root/reference/cloud_sky"

TRANSMITTANCE_W, TRANSMITTANCE_H = 256, 64  # transmittance_lut.gd:6
SKY_LUT_W, SKY_LUT_H = 200, 100  # sky_lut.gd:4


def available() -> bool:
    return os.path.exists(REF_LIB)


def build(force: bool = False) -> bool:
    """(Re)build oracle/_ref when the reference tree is mounted (build container only).  Returns available()."""
    if os.path.isdir(REFERENCE_SHADERS):
        srcs = [os.path.join(ROOT, "oracle", f) for f in ("glsl_compat.h", "ref_glue.inc", "build_ref.sh")]
        srcs += [os.path.join(REFERENCE_SHADERS, f) for f in ("clouds.glsl", "sky-lut.glsl", "transmittance-lut.glsl")]
        if force or not available() or os.path.getmtime(REF_LIB) < max(os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")], stdout=subprocess.DEVNULL)
    return available()


def box_mips(level0: np.ndarray) -> list:
    """[n,n,n,4] uint8 -> mip chain, 2x2x2 box average re-quantised (sum + 4) >> 3.  Godot's generator
    (perlworlnoise.tga.import:24 "mipmaps/generate=true") is outside the reference tree; this is the repo's stated
    definition of it, written here independently of oracle/cloudsky_oracle.cpp::build_mips."""
    chain = [np.ascontiguousarray(level0, dtype=np.uint8)]
    while chain[-1].shape[0] > 1:
        a = chain[-1].astype(np.uint32)
        m = a.shape[0] // 2
        s = a.reshape(m, 2, m, 2, m, 2, 4).sum(axis=(1, 3, 5))
        chain.append(np.ascontiguousarray(((s + 4) >> 3).astype(np.uint8)))
    return chain


def _rgba(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a, dtype=np.uint8)
    if a.shape[-1] == 3:  # RGB8 -> RGBA8, alpha 255 (what Godot's RGB8 import presents to the sampler)
        a = np.concatenate([a, np.full(a.shape[:-1] + (1,), 255, np.uint8)], -1)
    return np.ascontiguousarray(a)


class Reference:
    """One 'device': holds the textures the way cloud_sky.gd binds them and dispatches the compiled shaders."""

    def __init__(self, threads: int | None = None):
        if not available():
            raise RuntimeError(f"{REF_LIB} missing: run oracle/build_ref.sh in the build container (needs /root/reference)")
        self.lib = C.CDLL(REF_LIB)
        self.threads = threads or max(1, os.cpu_count() or 1)
        self.lib.ref_transmittance_lut.argtypes = [C.c_void_p] + [C.c_int] * 5
        self.lib.ref_sky_lut.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p] + [C.c_int] * 5
        self.lib.ref_clouds.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                        C.c_void_p, C.c_int, C.c_int, C.c_void_p] + [C.c_int] * 5
        self.large = self.small = self.weather = None
        self.tlut = None
        self.sky = None

    # cloud_sky.gd:311-339
    def upload_textures(self, large, small, weather):
        self.large = box_mips(_rgba(large))
        self.small = box_mips(_rgba(small))
        self.weather = _rgba(weather)

    # transmittance_lut.gd:72-78: 32x8 groups of 8x8
    def build_transmittance_lut(self) -> np.ndarray:
        out = np.zeros((TRANSMITTANCE_H, TRANSMITTANCE_W, 4), np.float16)
        r = self.lib.ref_transmittance_lut(out.ctypes.data, TRANSMITTANCE_W, TRANSMITTANCE_H, 32 * 8, 8 * 8, self.threads)
        assert r == 0
        self.tlut = out
        return out

    # sky_lut.gd:122-148: 25x13 groups of 8x8 (200 x 104 invocations over a 200 x 100 image)
    def build_sky_lut(self, sun_direction, tlut: np.ndarray | None = None) -> np.ndarray:
        t = np.ascontiguousarray(self.tlut if tlut is None else tlut).view(np.uint16)
        sun = np.asarray(sun_direction, np.float32)
        out = np.zeros((SKY_LUT_H, SKY_LUT_W, 4), np.float16)
        r = self.lib.ref_sky_lut(sun.ctypes.data, t.ctypes.data, TRANSMITTANCE_W, TRANSMITTANCE_H, out.ctypes.data, SKY_LUT_W, SKY_LUT_H,
                                 25 * 8, 13 * 8, self.threads)
        assert r == 0
        self.sky = out
        return out

    def write_sky_lut(self, a: np.ndarray):
        self.sky = np.ascontiguousarray(a).view(np.float16).reshape(SKY_LUT_H, SKY_LUT_W, 4).copy()

    # cloud_sky.gd:234-248.  `params` = the 112-byte push-constant block (anything with the buffer protocol / ctypes struct).
    # rows = None renders the whole W x H image with one dispatch; otherwise only the listed rows (one W x 1 dispatch each,
    # update_position = (0, row): the reference's own tile addressing, cloud_sky.gd:156-161).
    def render(self, params, width: int, height: int, rows=None, out: np.ndarray | None = None) -> np.ndarray:
        p = np.frombuffer(bytes(params), dtype=np.float32).copy()
        assert p.size == 28
        img = np.zeros((height, width, 4), np.float16) if out is None else out
        lp = (C.c_void_p * len(self.large))(*[a.ctypes.data for a in self.large])
        sp = (C.c_void_p * len(self.small))(*[a.ctypes.data for a in self.small])
        sky = np.ascontiguousarray(self.sky).view(np.uint16)

        def go(inv_w, inv_h):
            r = self.lib.ref_clouds(p.ctypes.data, lp, self.large[0].shape[0], len(self.large), sp, self.small[0].shape[0], len(self.small),
                                    self.weather.ctypes.data, self.weather.shape[1], self.weather.shape[0], sky.ctypes.data, SKY_LUT_W,
                                    SKY_LUT_H, img.ctypes.data, width, height, inv_w, inv_h, self.threads)
            assert r == 0

        if rows is None:
            go(width, height)
        else:
            ux, uy = p[2], p[3]
            rows = sorted(int(r_) for r_ in rows)
            i = 0
            while i < len(rows):  # runs of consecutive rows become one W x n dispatch at update_position = (0, first row)
                j = i
                while j + 1 < len(rows) and rows[j + 1] == rows[j] + 1:
                    j += 1
                p[2], p[3] = 0.0, float(rows[i])
                go(width, j - i + 1)
                i = j + 1
            p[2], p[3] = ux, uy
        return img

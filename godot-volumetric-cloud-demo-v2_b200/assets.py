"""Input textures for the cloud march: fixture loader and a procedural stand-in.

The reference's three input bitmaps (cloud_sky/perlworlnoise.tga, worlnoise.bmp, weather.bmp;
SURVEY §8(a) T1-T3) are decoded once in the build container into xz-compressed planar arrays
(tests/golden/make_asset_fixture.py) because /root/reference does not exist on the GPU box.
``load_fixture`` reads those; ``synthetic_textures`` generates tileable noise of the same shapes
for runs without the fixture (bench ``data: synthetic``).
"""
from __future__ import annotations

import lzma
import os
from typing import Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE_DIR = os.path.join(ROOT, "tests", "golden", "assets")


def _planar(path: str, shape) -> np.ndarray:
    raw = lzma.decompress(open(path, "rb").read())
    a = np.frombuffer(raw, dtype=np.uint8).reshape(shape)
    return np.ascontiguousarray(np.moveaxis(a, 0, -1))


def fixture_available(d: str = FIXTURE_DIR) -> bool:
    return all(os.path.exists(os.path.join(d, f)) for f in ("large_128_rgba8.xz", "small_32_rgb8.xz", "weather_512_rgb8.xz"))


def load_fixture(d: str = FIXTURE_DIR) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns (large [128,128,128,4], small [32,32,32,3], weather [512,512,3]) uint8, index [z][y][x][c]."""
    large = _planar(os.path.join(d, "large_128_rgba8.xz"), (4, 128, 128, 128))
    small = _planar(os.path.join(d, "small_32_rgb8.xz"), (3, 32, 32, 32))
    weather = _planar(os.path.join(d, "weather_512_rgb8.xz"), (3, 512, 512))
    return large, small, weather


def _tile_worley(n: int, cells: int, rng: np.random.Generator, dims: int) -> np.ndarray:
    """Tileable inverted Worley noise in [0,1] on an n^dims grid with cells^dims feature points."""
    pts = rng.random((cells,) * dims + (dims,), dtype=np.float32)
    grid = np.stack(np.meshgrid(*[(np.arange(n, dtype=np.float32) + 0.5) * cells / n] * dims, indexing="ij"), -1)
    base = np.floor(grid).astype(np.int32)
    best = np.full((n,) * dims, 1e9, np.float32)
    for off in np.ndindex(*(3,) * dims):
        o = np.array(off, np.int32) - 1
        c = base + o
        p = pts[tuple((c[..., k] % cells) for k in range(dims))] + c
        d = np.sqrt(((p - grid) ** 2).sum(-1))
        best = np.minimum(best, d)
    return np.clip(1.0 - best, 0.0, 1.0)


def synthetic_textures(seed: int = 0, large_n: int = 128, small_n: int = 32, weather_n: int = 512):
    """Procedural tileable stand-ins with the reference textures' shapes and rough statistics
    (large channel means ~0.85/0.71/0.71/0.71, weather R in [0.59,0.91], B in [0.07,1])."""
    rng = np.random.default_rng(seed)

    def vol(n, cells):
        return _tile_worley(n, cells, rng, 3)

    w = [vol(large_n, c) for c in (4, 8, 16, 32)]
    perlin_like = 0.5 * vol(large_n, 2) + 0.5 * vol(large_n, 3)
    r = np.clip(0.6 + 0.4 * (perlin_like * 0.6 + 0.4 * w[0]), 0, 1)
    large = np.stack([r, w[1] * 0.5 + 0.45, w[2] * 0.5 + 0.45, w[3] * 0.5 + 0.45], -1)
    small = np.stack([vol(small_n, c) * 0.5 + 0.45 for c in (2, 4, 8)], -1)
    w2 = _tile_worley(weather_n, 6, rng, 2)
    w3 = _tile_worley(weather_n, 3, rng, 2)
    weather = np.stack([0.59 + 0.32 * w3, np.zeros_like(w2), np.clip(0.07 + 1.2 * (w2 - 0.25), 0.07, 1.0)], -1)
    q = lambda a: np.ascontiguousarray(np.clip(np.rint(a * 255.0), 0, 255).astype(np.uint8))
    return q(large), q(small), q(weather)


def load_default_textures():
    """(large, small, weather, description): the decoded reference assets when the fixture is
    present, else the procedural stand-ins."""
    if fixture_available():
        return (*load_fixture(), "reference textures (decoded fixture tests/golden/assets)")
    return (*synthetic_textures(0), "synthetic tileable worley noise of the reference shapes")

#!/usr/bin/env python
"""Decode the three reference input textures into committed fixtures.

Runs ONLY in the build container (needs /root/reference and PIL).  The GPU box has no
/root/reference, so the decoded texels travel as xz-compressed planar arrays:

  large_128_rgba8.xz   128^3 RGBA8, planar [c][z][y][x]  <- cloud_sky/perlworlnoise.tga
                       (16384x128 strip, 128 horizontal slices: perlworlnoise.tga.import:26)
  small_32_rgb8.xz     32^3  RGB8,  planar [c][z][y][x]  <- cloud_sky/worlnoise.bmp
                       (1024x32 strip, 32 horizontal slices: worlnoise.bmp.import:26)
  weather_512_rgb8.xz  512^2 RGB8,  planar [c][y][x]     <- cloud_sky/weather.bmp

Texel (x, y, z) of a volume = strip column z*N + x, row y, row 0 = top of the decoded image
(SURVEY §8(a) T1/T2).  These are input DATA (not reference source code); manifest.json records
the sha256 of the source files and of the decoded arrays so tests can pin the decode.
"""
import hashlib, json, lzma, os, sys
import numpy as np
from PIL import Image

REF = "/root/reference/cloud_sky"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")
Image.MAX_IMAGE_PIXELS = None


def sha(b):
    return hashlib.sha256(b).hexdigest()


def strip_to_volume(a, n):
    h, w, c = a.shape
    assert h == n and w == n * n, a.shape
    return np.ascontiguousarray(a.reshape(n, n, n, c).transpose(1, 0, 2, 3))  # [z][y][x][c]


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    specs = [("perlworlnoise.tga", "large_128_rgba8.xz", 128, "RGBA"),
             ("worlnoise.bmp", "small_32_rgb8.xz", 32, "RGB"),
             ("weather.bmp", "weather_512_rgb8.xz", 0, "RGB")]
    for src, dst, n, mode in specs:
        path = os.path.join(REF, src)
        raw = open(path, "rb").read()
        im = Image.open(path)
        assert im.mode == mode, (src, im.mode)
        a = np.asarray(im)
        decoded_sha = sha(a.tobytes())
        vol = strip_to_volume(a, n) if n else a  # [z][y][x][c] or [y][x][c]
        planar = np.ascontiguousarray(np.moveaxis(vol, -1, 0))
        blob = lzma.compress(planar.tobytes(), preset=9 | lzma.PRESET_EXTREME)
        open(os.path.join(OUT, dst), "wb").write(blob)
        manifest[dst] = {"source": "cloud_sky/" + src, "source_sha256": sha(raw), "decoded_strip_sha256": decoded_sha,
                         "shape_interleaved": list(vol.shape), "interleaved_sha256": sha(vol.tobytes()),
                         "xz_bytes": len(blob)}
        print(dst, vol.shape, len(blob))
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    sys.exit(main())

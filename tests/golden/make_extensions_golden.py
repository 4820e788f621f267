#!/usr/bin/env python
"""Writes tests/golden/extensions_golden.npz from the CPU oracle: frozen outputs of the two extensions that have no reference
implementation to follow (the reference lists them as TODOs, README.md:29-30) — the noise generator's bytes and the
Bruneton-2017-mapped transmittance LUT.  Run from the repo root: python tests/golden/make_extensions_golden.py"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import cloudsky_b200 as cs

CASES = [("large", cs.NOISE_LARGE, 32), ("small", cs.NOISE_SMALL, 32), ("weather", cs.NOISE_WEATHER, 128)]


def main():
    ora = cs.Library(os.path.join(ROOT, "oracle", "libcloudsky_oracle.so"))
    ctx = ora.context(0)
    ctx.set_threads(os.cpu_count() or 1)
    out = {}
    for name, kind, n in CASES:
        for tag, seed in (("default", None), ("seed77", 77)):
            p = ora.noise_params_default(kind)
            if seed is not None:
                p.seed = seed
                p.worley_frequency = 3
            a = ctx.generate_noise(kind, n, p)
            out[f"noise_{name}_{tag}_sha256"] = np.frombuffer(hashlib.sha256(a.tobytes()).digest(), np.uint8)
            out[f"noise_{name}_{tag}_slice"] = a[0, :8, :8].copy() if kind != cs.NOISE_WEATHER else a[:8, :8].copy()
    ctx.set_transmittance_parametrisation(cs.TLUT_BRUNETON2017)
    ctx.build_transmittance_lut()
    out["transmittance_bruneton"] = ctx.read_transmittance_lut()[::2, ::4].copy()  # every 2nd row, 4th column: 32 x 64 texels
    ctx.build_sky_lut((0.3, 0.5, -0.81))
    out["sky_bruneton"] = ctx.read_sky_lut()[::4, ::4].copy()
    ctx.close()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "extensions_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

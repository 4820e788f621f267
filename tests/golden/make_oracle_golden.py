#!/usr/bin/env python
"""Generate tests/golden/oracle_golden.npz: outputs of the CPU oracle on the decoded reference textures.

The reference cannot be executed anywhere we have access to (no Godot / Vulkan / GLSL toolchain), so these
vectors do NOT come from the reference itself; they freeze the oracle's output so that (a) the oracle cannot
drift silently and (b) the GPU path can be checked against committed numbers on a box without the oracle.
Run from the repo root:  python tests/golden/make_oracle_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {  # name -> make_params kwargs (tests/conftest.py)
    "noon": dict(sun=(0.0, 1.0, 0.0)),
    "sunset_wind": dict(sun=(-0.998773, 0.0495291, 2.69869e-07), time=37.5, wind_direction=0.7, wind_speed=3.0),
    "overcast": dict(sun=(0.2, 0.9, -0.3), coverage=1.0, density=0.1, energy=2.0, color=(1.0, 0.8, 0.6)),
}
W, H = 64, 32


def main():
    import cloudsky_b200 as cs
    from cloudsky_b200 import assets
    from conftest import make_params, prepared_context
    lib = cs.Library(os.path.join(ROOT, "oracle", "libcloudsky_oracle.so"))
    tex = assets.load_fixture()
    ctx = prepared_context(lib, tex, W, H, threads=os.cpu_count())
    out = {"transmittance": ctx.read_transmittance_lut()}
    for name, kw in CASES.items():
        p = make_params(lib, W, H, **kw)
        ctx.build_sky_lut(tuple(p.light_direction))
        out[f"sky_{name}"] = ctx.read_sky_lut()
        ctx.set_march_config(128, 6)
        ctx.render_frame(p)
        out[f"clouds_{name}"] = ctx.read_image()
        out[f"params_{name}"] = p.as_floats()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

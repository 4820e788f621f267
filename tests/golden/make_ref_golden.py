#!/usr/bin/env python
"""Generate tests/golden/ref_golden.npz: outputs of the REFERENCE ITSELF — its three GLSL compute shaders compiled
unmodified by g++ behind oracle/glsl_compat.h (oracle/build_ref.sh -> oracle/_ref/libcloudsky_ref.so) — on the decoded
reference textures.  These vectors are what pins the hand-written oracle and the CUDA kernels to the reference:

  transmittance            256x64  RGBA16F   transmittance-lut.glsl, the one dispatch of transmittance_lut.gd:72-78
  sky_<case>               200x100 RGBA16F   sky-lut.glsl for the case's sun direction (sky_lut.gd:122-148)
  clouds_<case>            128x64  RGBA16F   clouds.glsl, whole image, the reference's fixed 128 primary / 6+1 light steps
  c3_rows, c3_frame<k>     16 rows of the 2048x1024 bench frame k (animated wind, noon sun; bench.py's own parameters)
  params_*                 the 112-byte push-constant blocks that were pushed (so a box without the host logic can replay them)

Needs /root/reference (build container).  Run from the repo root:  python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {  # name -> make_params kwargs (tests/conftest.py)
    "noon": dict(sun=(0.0, 1.0, 0.0)),
    "sunset": dict(sun=(-0.998773, 0.0495291, 2.69869e-07)),  # cloud-demo.tscn:21
    "sunset_wind": dict(sun=(-0.998773, 0.0495291, 2.69869e-07), time=37.5, wind_direction=0.7, wind_speed=3.0),
    "oblique_wind": dict(sun=(0.5, 0.6, -0.3), time=12.25, wind_direction=2.1, wind_speed=2.0),
    "overcast": dict(sun=(0.2, 0.9, -0.3), coverage=1.0, density=0.1, energy=2.0, color=(1.0, 0.8, 0.6)),
}
W, H = 128, 64
C3_W, C3_H = 2048, 1024
C3_ROWS = list(range(5, C3_H, 64))  # 16 rows
C3_FRAMES = (0, 15)


def main():
    import cloudsky_b200 as cs
    from cloudsky_b200 import assets
    from conftest import make_params, ORACLE_LIB, _build_oracle
    import refbind
    import bench

    assert refbind.build(force=True), "oracle/_ref could not be built (is /root/reference mounted?)"
    _build_oracle()
    host = cs.Library(ORACLE_LIB)  # only its HOST logic (settings -> push constants) is used here
    tex = assets.load_fixture()
    ref = refbind.Reference()
    ref.upload_textures(*tex)
    out = {"transmittance": ref.build_transmittance_lut().copy()}
    for name, kw in CASES.items():
        p = make_params(host, W, H, **kw)
        out[f"sky_{name}"] = ref.build_sky_lut(tuple(p.light_direction)).copy()
        out[f"clouds_{name}"] = ref.render(p, W, H)
        out[f"params_{name}"] = p.as_floats()
    out["c3_rows"] = np.asarray(C3_ROWS, np.int32)
    sun = (0.0, 1.0, 0.0)
    out["c3_sky"] = ref.build_sky_lut(sun).copy()
    for k in C3_FRAMES:
        p = bench.frame_params(host, k, sun)
        img = ref.render(p, C3_W, C3_H, rows=C3_ROWS)
        out[f"c3_frame{k}"] = np.ascontiguousarray(img[C3_ROWS])
        out[f"c3_params{k}"] = p.as_floats()
    with open(os.path.join(refbind.REF_DIR, "SOURCES.sha256")) as f:
        out["reference_sha256"] = np.asarray(f.read())
    path = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
    np.savez_compressed(path, **out)
    print({k: getattr(v, "shape", None) for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

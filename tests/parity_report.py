#!/usr/bin/env python
"""Parity margins of the CUDA path against the CPU oracle (test infrastructure) on a 256x128 frame: pass
fraction at the stated tolerances and error percentiles, for the DESIGN.md table."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cloudsky_b200 as cs
from cloudsky_b200 import assets
from conftest import make_params, prepared_context

CASES = {"noon": dict(sun=(0, 1, 0)), "sunset": dict(sun=(-0.998773, 0.0495291, 2.69869e-07)),
         "wind": dict(sun=(0.5, 0.5, 0.70710678), time=37.5, wind_direction=0.7, wind_speed=3.0),
         "overcast": dict(sun=(0, 1, 0), coverage=1.0, density=0.1)}

def main():
    W, H = 256, 128
    lib = cs.load_product(); ora = cs.Library(os.path.join(ROOT, "oracle", "libcloudsky_oracle.so"))
    tex = assets.load_fixture()
    o = prepared_context(ora, tex, W, H, threads=os.cpu_count()); g = prepared_context(lib, tex, W, H)
    rows = []
    for name, kw in CASES.items():
        p = make_params(lib, W, H, **kw)
        o.build_sky_lut(tuple(p.light_direction)); o.render_frame(p); ref = o.read_image().astype(np.float32)[1:, 1:]
        g.write_sky_lut(o.read_sky_lut())
        for mode, mname in ((cs.MODE_STRICT, "strict"), (cs.MODE_FAST, "fast"), (cs.MODE_FAST | cs.MODE_TEX, "fast+tex"), (cs.MODE_FAST | cs.MODE_HALF, "fast+half")):
            g.set_march_config(128, 6, mode); g.render_frame(p)
            d = np.abs(g.read_image().astype(np.float32)[1:, 1:] - ref)
            r = dict(case=name, mode=mname,
                     pass_1e3_2e3=round(float((d <= 1e-3 + 2e-3 * np.abs(ref)).all(-1).mean()), 5),
                     pass_2e3_1e2=round(float((d <= 2e-3 + 1e-2 * np.abs(ref)).all(-1).mean()), 5),
                     bit_identical=round(float((d == 0).all(-1).mean()), 4), max_abs=round(float(d.max()), 5),
                     p999_abs=round(float(np.percentile(d, 99.9)), 6), mean_abs=float(f"{d.mean():.3g}"))
            print(json.dumps(r), flush=True); rows.append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/parity_report.json", "w"), indent=1)

if __name__ == "__main__":
    main()

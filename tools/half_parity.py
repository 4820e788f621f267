#!/usr/bin/env python
"""Development aid: parity margins and speed of CS_MODE_HALF (optionally of a variant library built by tools/build_variants.sh)
against the CPU oracle (test infrastructure): 256x128 at 128/6, and BASELINE config 2's 1024x512 at 64/5 where few steps make
every lit/unlit flip count.  usage: half_parity.py [--lib path]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cloudsky_b200 as cs
from cloudsky_b200 import assets
from conftest import make_params, prepared_context, ORACLE_LIB

lib = cs.Library(sys.argv[sys.argv.index("--lib") + 1]) if "--lib" in sys.argv else cs.load_product()
ora = cs.Library(ORACLE_LIB)
tex = assets.load_fixture()
out = {"lib": os.path.basename(lib.path)}
for tag, (W, H, P, cone, kw) in {"c1_noon": (256, 128, 128, 6, dict(sun=(0, 1, 0))), "c1_sunset": (256, 128, 128, 6, dict(sun=(-0.998773, 0.0495291, 2.69869e-07))),
                                 "c2": (1024, 512, 64, 5, dict(sun=(0, 1, 0)))}.items():
    o = prepared_context(ora, tex, W, H, threads=os.cpu_count()); g = prepared_context(lib, tex, W, H)
    p = make_params(lib, W, H, **kw)
    o.build_sky_lut(tuple(p.light_direction)); o.set_march_config(P, cone); o.render_frame(p); ref = o.read_image().astype(np.float32)[1:, 1:]
    g.write_sky_lut(o.read_sky_lut())
    for mode, name in ((cs.MODE_FAST, "fast"), (cs.MODE_FAST | cs.MODE_TEX, "tex"), (cs.MODE_FAST | cs.MODE_HALF, "half")):
        g.set_march_config(P, cone, mode); g.render_frame(p)
        d = np.abs(g.read_image().astype(np.float32)[1:, 1:] - ref)
        out[f"{tag}_{name}"] = [round(float((d <= 2e-3 + 1e-2 * np.abs(ref)).all(-1).mean()), 5), float(f"{d.mean():.3g}")]
    o.close(); g.close()
g = prepared_context(lib, tex, 2048, 1024)
p = make_params(lib, 2048, 1024, time=1.0)
for mode, name in ((cs.MODE_FAST, "fast"), (cs.MODE_FAST | cs.MODE_HALF, "half")):
    g.set_march_config(128, 7, mode)
    out[f"c3_ms_{name}"] = round(min(g.time_render_frame(p, 2, 5) for _ in range(3)), 4)
print(json.dumps(out), flush=True)

#!/usr/bin/env python
"""Development aid: parity margins (vs the CPU oracle, test infrastructure) and C3 speed of one library build — the default or a
variant from tools/build_variants.sh (--lib path).  Cases: the parity-report four + the hardest fuzz case found in round 2 (low
sun, high density)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cloudsky_b200 as cs
from cloudsky_b200 import assets
from conftest import make_params, prepared_context, ORACLE_LIB

lib = cs.Library(sys.argv[sys.argv.index("--lib") + 1]) if "--lib" in sys.argv else cs.load_product()
modes = {"fast": cs.MODE_FAST, "tex": cs.MODE_FAST | cs.MODE_TEX, "half": cs.MODE_FAST | cs.MODE_HALF}
ora = cs.Library(ORACLE_LIB)
tex = assets.load_fixture()
CASES = {"noon": dict(sun=(0, 1, 0)), "sunset": dict(sun=(-0.998773, 0.0495291, 2.69869e-07)),
         "wind": dict(sun=(0.5, 0.5, 0.70710678), time=37.5, wind_direction=0.7, wind_speed=3.0),
         "overcast": dict(sun=(0, 1, 0), coverage=1.0, density=0.1),
         "fuzz0": dict(sun=(-0.7406591555684899, 0.028887514395757697, 0.6712596567533359), coverage=0.6400634729041019, density=0.14526252276707127,
                       time=58.848768727964455, wind_direction=2.1166201293218947, wind_speed=3.4108498920879664, energy=1.7568986792654124,
                       color=(0.815233898290137, 0.601573160741395, 0.3402688374655125))}
W, H = 256, 128
o = prepared_context(ora, tex, W, H, threads=os.cpu_count()); g = prepared_context(lib, tex, W, H)
for name, kw in CASES.items():
    p = make_params(lib, W, H, **kw)
    o.build_sky_lut(tuple(p.light_direction)); o.set_march_config(128, 6); o.render_frame(p); ref = o.read_image().astype(np.float32)[1:, 1:]
    g.write_sky_lut(o.read_sky_lut())
    row = {"lib": os.path.basename(lib.path), "case": name}
    for mname, mode in modes.items():
        g.set_march_config(128, 6, mode); g.render_frame(p)
        d = np.abs(g.read_image().astype(np.float32)[1:, 1:] - ref)
        row[mname] = dict(tight=round(float((d <= 1e-3 + 2e-3 * np.abs(ref)).all(-1).mean()), 5), fast_tol=round(float((d <= 2e-3 + 1e-2 * np.abs(ref)).all(-1).mean()), 5),
                          identical=round(float((d == 0).all(-1).mean()), 4), max_abs=round(float(d.max()), 5), mean_abs=float(f"{d.mean():.3g}"))
    print(json.dumps(row), flush=True)
o.close(); g.close()
g = prepared_context(lib, tex, 2048, 1024)
t = {"lib": os.path.basename(lib.path)}
for cov in (0.2, 1.0):
    p = make_params(lib, 2048, 1024, time=1.0, coverage=cov)
    for mname, mode in modes.items():
        g.set_march_config(128, 7, mode)
        t[f"c3_ms_{mname}_cov{cov}"] = round(min(g.time_render_frame(p, 2, 5) for _ in range(3)), 4)
print(json.dumps(t), flush=True)

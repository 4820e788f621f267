#!/usr/bin/env python
"""Development aid: which push-constant makes a fuzz case leave the FAST gate?  Re-creates case `it` of
tests/test_gpu_parity.py::test_random_parameter_sets_all_sampler_modes and resets one parameter at a time."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cloudsky_b200 as cs
from cloudsky_b200 import assets
from conftest import make_params, prepared_context, compare_images, ORACLE_LIB

W, H = 256, 128
lib = cs.load_product(); ora = cs.Library(ORACLE_LIB)
tex = assets.load_fixture()
o = prepared_context(ora, tex, W, H, threads=os.cpu_count()); g = prepared_context(lib, tex, W, H)
rng = np.random.default_rng(777)
cases = []
for it in range(6):
    el = rng.uniform(0.05, 1.0) if it % 3 else rng.uniform(-0.02, 0.06)
    az = rng.uniform(0, 2 * np.pi)
    c = np.sqrt(max(0.0, 1 - el * el))
    cases.append(dict(sun=(float(np.cos(az) * c), float(el), float(np.sin(az) * c)), coverage=float(rng.uniform(0.1, 1.0)), density=float(rng.uniform(0.02, 0.15)),
                      time=float(rng.uniform(0.0, 300.0)), wind_direction=float(rng.uniform(0, 6.28)), wind_speed=float(rng.uniform(0, 6)),
                      energy=float(rng.uniform(0.3, 3.0)), color=tuple(float(v) for v in rng.uniform(0.3, 1.0, 3))))

def run(kw, modes=(("strict", cs.MODE_STRICT), ("fast", cs.MODE_FAST))):
    po = make_params(ora, W, H, **kw)
    o.set_march_config(128, 6); o.build_sky_lut(tuple(po.light_direction)); o.render_frame(po); ref = o.read_image()
    g.write_sky_lut(o.read_sky_lut())
    out = {}
    for name, mode in modes:
        g.set_march_config(128, 6, mode); g.render_frame(make_params(lib, W, H, **kw))
        img = g.read_image()
        frac, mx = compare_images(img, ref, 2e-3, 1e-2)
        d = np.abs(img.astype(np.float32) - ref.astype(np.float32))[1:, 1:]
        bad = ~(d <= 2e-3 + 1e-2 * np.abs(ref.astype(np.float32)[1:, 1:])).all(-1)
        rows = np.where(bad.any(1))[0]
        out[name] = (round(frac, 5), round(mx, 4), int(bad.sum()), (int(rows.min()), int(rows.max())) if rows.size else None, [round(float(v), 4) for v in d.reshape(-1, 4).max(0)])
    return out

it = int(sys.argv[1]) if len(sys.argv) > 1 else 0
kw = cases[it]
print("case", it, json.dumps(kw))
print("full      ", run(kw))
defaults = dict(sun=(0.0, 1.0, 0.0), coverage=0.2, density=0.05, time=0.0, wind_direction=0.0, wind_speed=1.0, energy=1.0, color=(1.0, 1.0, 1.0))
for k in kw:
    kk = dict(kw); kk[k] = defaults[k]
    print(f"{k:14s}->default", run(kk))
for t in (10.0, 30.0, 58.0, 100.0, 200.0):
    kk = dict(kw); kk["time"] = t
    print(f"time={t}", run(kk))

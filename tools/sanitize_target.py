#!/usr/bin/env python
"""Small workload touching every kernel variant, for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudsky_b200 as cs
from cloudsky_b200 import assets

lib = cs.load_product()
tex = assets.synthetic_textures(seed=3, large_n=16, small_n=8, weather_n=32)
W, H = 72, 40
ctx = lib.context(0)
ctx.upload_textures(*tex); ctx.build_transmittance_lut(); ctx.build_sky_lut((0.3, 0.8, 0.1)); ctx.resize(W, H)
s = lib.settings_demo(); s.cloud_coverage = 0.6
st = lib.frame_state_init(); st.light_direction[:] = [0.3, 0.8, 0.1]
lib.frame_advance(st, s, 2.0)
p = lib.fill_cloud_params(s, st, W, H)
for mode in (cs.MODE_FAST, cs.MODE_FAST | cs.MODE_TEX, cs.MODE_FAST | cs.MODE_HALF, cs.MODE_FAST | cs.MODE_EARLY_OUT, cs.MODE_STRICT):
    ctx.set_march_config(32, 6, mode)
    ctx.render_frame(p)
    ctx.set_counters_enabled(True); ctx.render_frame(p); ctx.get_counters(); ctx.set_counters_enabled(False)
ctx.set_march_config(32, 6, cs.MODE_FAST)
ctx.set_step_budget(60.0, 8); ctx.render_frame(p); ctx.set_step_budget(0.0, 1)
own, _ = ctx.peer_alloc(4 * W * H * 8); mirror, _ = ctx.peer_alloc(4 * W * H * 8); flags, _ = ctx.peer_alloc(256)
ctx.set_output_mirrors(own, 4 * W * H * 8, [mirror])
ctx.render_row_bands_to(p, 0, 8, 16, 3, own)
ctx.render_row_bands_to(p, 8, 8, 16, 2, own)
suns = np.array([[0.0, 1.0, 0.0], [0.6, 0.8, 0.0], [-0.5, 0.5, 0.3]], np.float32)
ctx.render_sun_batch_to(p, suns, own + W * H * 8)
ctx.peer_barrier(0, 1, [flags], 1)
ctx.sync(); ctx.peer_check()
ctx.set_output_mirrors(0, 0, [])
for q in (own, mirror, flags):
    ctx.peer_free(q)
img = ctx.render_frame_host(p)
assert np.isfinite(img.astype(np.float32)).all()
ctx.close()
print("sanitize target ok")

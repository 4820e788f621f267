#!/usr/bin/env python
"""Development aid: A/B two builds of the library over several dispatch sizes."""
import glob, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cloudsky_b200 as cs
from cloudsky_b200 import assets
large, small, weather, _ = assets.load_default_textures()
libs = [("current", cs.capi.PRODUCT_LIB)] + [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob("build/variants/lib_*.so"))]
cases = [(2048, 1024, 128, 7, 0.2), (2048, 1024, 128, 7, 1.0), (1024, 512, 64, 5, 0.2), (768, 768, 128, 6, 0.2), (256, 128, 128, 6, 0.2), (8192, 4096, 256, 11, 1.0)]
for name, path in libs:
    lib = cs.Library(path); ctx = lib.context(0)
    ctx.upload_textures(large, small, weather); ctx.build_transmittance_lut(); ctx.build_sky_lut((0, 1, 0))
    out = {"lib": name}
    for (W, H, P, cone, cov) in cases:
        ctx.resize(W, H)
        s = lib.settings_demo(); s.cloud_coverage = cov
        st = lib.frame_state_init(); st.light_direction[:] = [0, 1, 0]; lib.frame_advance(st, s, 1.0)
        p = lib.fill_cloud_params(s, st, W, H)
        ctx.set_march_config(P, cone, cs.MODE_FAST)
        out[f"{W}x{H}/{P}/{cone}/cov{cov}"] = round(min(ctx.time_render_frame(p, 2, 4) for _ in range(3)), 4)
    print(json.dumps(out), flush=True); ctx.close()

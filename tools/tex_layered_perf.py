#!/usr/bin/env python
"""Development aid: CS_MODE_TEX with the volumes fetched from the two-slice layered textures (CLOUDSKY_TEX_LAYERED = 0..3: bit 0
small volume, bit 1 large volume) — device time on the C3 shape and agreement with the plain 3-D texture path."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudsky_b200 as cs
from cloudsky_b200 import assets

libpath = sys.argv[sys.argv.index("--lib") + 1] if "--lib" in sys.argv else cs.capi.PRODUCT_LIB
lib = cs.Library(libpath)
large, small, weather, _ = assets.load_default_textures()
W, H, P, cone = 2048, 1024, 128, 7
ref = {}
for layered in (0, 1, 2, 3):
    os.environ["CLOUDSKY_TEX_LAYERED"] = str(layered)
    ctx = lib.context(0)
    ctx.upload_textures(large, small, weather); ctx.build_transmittance_lut(); ctx.build_sky_lut((0, 1, 0)); ctx.resize(W, H)
    out = {"lib": os.path.basename(libpath), "tex_layered": layered}
    for cov in (0.2, 1.0):
        s = lib.settings_demo(); s.cloud_coverage = cov
        st = lib.frame_state_init(); st.light_direction[:] = [0, 1, 0]
        lib.frame_advance(st, s, 1.0)
        p = lib.fill_cloud_params(s, st, W, H)
        ctx.set_march_config(P, cone, cs.MODE_FAST | cs.MODE_TEX)
        out[f"ms_cov{cov}"] = round(min(ctx.time_render_frame(p, 2, 5) for _ in range(3)), 4)
        ctx.render_frame(p)
        img = ctx.read_image().astype(np.float32)[1:, 1:]
        if layered == 0:
            ref[cov] = img
        else:
            d = np.abs(img - ref[cov])
            out[f"in_fast_tol_of_3d_cov{cov}"] = round(float((d <= 2e-3 + 1e-2 * np.abs(ref[cov])).all(-1).mean()), 6)
            out[f"max_abs_cov{cov}"] = float(d.max())
    print(json.dumps(out), flush=True)
    ctx.close()

#!/usr/bin/env python
"""Development aid: time launch-shape variants of the fast kernel built into build/variants/lib_*.so
(see the CS_* macros at the top of clouds_fast.cu)."""
import glob, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cloudsky_b200 as cs
from cloudsky_b200 import assets

large, small, weather, _ = assets.load_default_textures()
libs = [("default", cs.capi.PRODUCT_LIB)] + [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob("build/variants/lib_*.so"))]
W, H, P, cone = 2048, 1024, 128, 7
flags = int(sys.argv[sys.argv.index("--flags") + 1]) if "--flags" in sys.argv else 0  # MODE_EARLY_OUT (2) | MODE_TEX (4)
for name, path in libs:
    lib = cs.Library(path)
    ctx = lib.context(0)
    ctx.upload_textures(large, small, weather); ctx.build_transmittance_lut(); ctx.build_sky_lut((0, 1, 0)); ctx.resize(W, H)
    out = {"variant": name}
    for cov in (0.2, 1.0):
        s = lib.settings_demo(); s.cloud_coverage = cov
        st = lib.frame_state_init(); st.light_direction[:] = [0, 1, 0]
        lib.frame_advance(st, s, 1.0)
        p = lib.fill_cloud_params(s, st, W, H)
        ctx.set_march_config(P, cone, cs.MODE_FAST | flags)
        out[f"ms_cov{cov}"] = round(min(ctx.time_render_frame(p, 2, 5) for _ in range(3)), 4)
    print(json.dumps(out), flush=True)
    ctx.close()

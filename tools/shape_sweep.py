#!/usr/bin/env python
"""Development aid: time compile-time variants of the fast kernel built into build/variants/lib_*.so (tools/build_variants.sh;
see the CS_* macros at the top of clouds_fast.cu) against the default library, and compare their images with the default's
(bit-identical? inside the FAST parity tolerance?).  usage: shape_sweep.py [--flags N] [--only name,name]"""
import glob, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudsky_b200 as cs
from cloudsky_b200 import assets

large, small, weather, _ = assets.load_default_textures()
libs = [("default", cs.capi.PRODUCT_LIB)] + [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob("build/variants/lib_*.so"))]
if "--only" in sys.argv:
    keep = set(sys.argv[sys.argv.index("--only") + 1].split(",")) | {"default"}
    libs = [l for l in libs if l[0] in keep]
W, H, P, cone = 2048, 1024, 128, 7
flags = int(sys.argv[sys.argv.index("--flags") + 1]) if "--flags" in sys.argv else 0  # MODE_EARLY_OUT (2) | MODE_TEX (4)
ref = {}
for name, path in libs:
    lib = cs.Library(path)
    ctx = lib.context(0)
    ctx.upload_textures(large, small, weather); ctx.build_transmittance_lut(); ctx.build_sky_lut((0, 1, 0)); ctx.resize(W, H)
    out = {"variant": name, "flags": flags}
    for cov in (0.2, 1.0):
        s = lib.settings_demo(); s.cloud_coverage = cov
        st = lib.frame_state_init(); st.light_direction[:] = [0, 1, 0]
        lib.frame_advance(st, s, 1.0)
        p = lib.fill_cloud_params(s, st, W, H)
        ctx.set_march_config(P, cone, cs.MODE_FAST | flags)
        out[f"ms_cov{cov}"] = round(min(ctx.time_render_frame(p, 2, 5) for _ in range(3)), 4)
        ctx.render_frame(p)
        img = ctx.read_image()
        if name == "default":
            ref[cov] = img
        else:
            a, b = img.astype(np.float32)[1:, 1:], ref[cov].astype(np.float32)[1:, 1:]
            d = np.abs(a - b)
            out[f"same_bits_cov{cov}"] = bool((img.view(np.uint16) == ref[cov].view(np.uint16)).all())
            out[f"in_fast_tol_cov{cov}"] = round(float((d <= 2e-3 + 1e-2 * np.abs(b)).all(-1).mean()), 6)
            out[f"max_abs_cov{cov}"] = float(d.max())
    print(json.dumps(out), flush=True)
    ctx.close()

#!/usr/bin/env python
"""Device time of the fast kernel on C3 (coverage 0.2) and the high-coverage config (development aid; A/B runs are
done by setting the library's environment knobs, e.g. CLOUDSKY_FP32_RECORDS=1 or CLOUDSKY_RECORDS=1|3|7)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudsky_b200 as cs
from cloudsky_b200 import assets

def main():
    variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,1,2,3".split(","))]
    lib = cs.load_product()
    large, small, weather, _ = assets.load_default_textures()
    W, H, P, cone = 2048, 1024, 128, 7
    ctx = lib.context(0)
    ctx.upload_textures(large, small, weather)
    ctx.build_transmittance_lut(); ctx.build_sky_lut((0, 1, 0)); ctx.resize(W, H)
    for cov in (0.2, 1.0):
        s = lib.settings_demo(); s.cloud_coverage = cov
        st = lib.frame_state_init(); st.light_direction[:] = [0, 1, 0]
        lib.frame_advance(st, s, 1.0)
        p = lib.fill_cloud_params(s, st, W, H)
        ref = None
        for v in variants:
            ctx.set_march_config(P, cone, cs.MODE_FAST | v)  # variant = mode flags: 2 early-out, 4 texture-unit filtering
            ms = min(ctx.time_render_frame(p, 2, 5) for _ in range(3))
            ctx.render_frame(p)
            img = ctx.read_image()
            if ref is None:
                ref = img
            same = bool((img.view(np.uint16) == ref.view(np.uint16)).all())
            d = np.abs(img.astype(np.float32) - ref.astype(np.float32))[1:, 1:]
            ok = float((d <= 2e-3 + 1e-2 * np.abs(ref.astype(np.float32)[1:, 1:])).all(-1).mean())
            print(json.dumps(dict(coverage=cov, variant=v, ms=round(ms, 4), mray_steps_s=round((W * H - W - H + 1) * P / ms / 1e3, 1), identical_to_first=same,
                                  within_fast_tol_of_first=round(ok, 6), max_abs=float(d.max()))), flush=True)
    ctx.close()

if __name__ == "__main__":
    main()

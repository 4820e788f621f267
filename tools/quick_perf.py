#!/usr/bin/env python
"""Quick device-side timing of the cloud kernels (development aid; bench.py is the contract)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudsky_b200 as cs
from cloudsky_b200 import assets

def main():
    lib = cs.Library(sys.argv[sys.argv.index("--lib") + 1]) if "--lib" in sys.argv else cs.load_product()  # --lib: a variant built by tools/build_variants.sh
    large, small, weather, desc = assets.load_default_textures()
    res = []
    cfgs = [(2048, 1024, 128, 6, 0.2), (2048, 1024, 128, 7, 0.2), (2048, 1024, 128, 7, 1.0), (1024, 512, 64, 5, 0.2)]
    if "--only" in sys.argv:
        cfgs = [cfgs[int(sys.argv[sys.argv.index("--only") + 1])]]
    flags = int(sys.argv[sys.argv.index("--flags") + 1]) if "--flags" in sys.argv else 0  # MODE_EARLY_OUT (2) | MODE_TEX (4)
    for (W, H, P, cone, cov) in cfgs:
        ctx = lib.context(0)
        ctx.upload_textures(large, small, weather)
        ctx.build_transmittance_lut(); ctx.build_sky_lut((0, 1, 0)); ctx.resize(W, H)
        s = lib.settings_demo(); s.cloud_coverage = cov
        st = lib.frame_state_init(); st.light_direction[:] = [0, 1, 0]
        lib.frame_advance(st, s, 1.0)
        p = lib.fill_cloud_params(s, st, W, H)
        for mode, name in ((cs.MODE_FAST | flags, "fast" + (f"+{flags}" if flags else "")), (cs.MODE_STRICT, "strict")):
            if name == "strict" and "--strict" not in sys.argv and (cone != 6):
                continue
            ctx.set_march_config(P, cone, mode)
            ctx.set_counters_enabled(True); ctx.render_frame(p); k = ctx.get_counters().as_dict(); ctx.set_counters_enabled(False)
            ms = ctx.time_render_frame(p, 2, 5)
            steps = k["marched_pixels"] * P
            r = dict(W=W, H=H, P=P, cone=cone, coverage=cov, mode=name, ms=round(ms, 4), mray_steps_s=round(steps / ms / 1e3, 1),
                     gevals_s=round(k["density_evals"] / ms / 1e6, 2), lit_frac=round(k["lit_steps"] / k["primary_steps"], 4),
                     alg_GBps=round((80 * k["density_evals"] + 8 * W * H) / ms / 1e6, 1), **k)
            print(json.dumps(r), flush=True)
            res.append(r)
        ctx.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/quick_perf.json", "w"), indent=1)

if __name__ == "__main__":
    main()

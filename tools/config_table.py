#!/usr/bin/env python
"""Device time of BASELINE.json's configurations 2-5 on ONE B200 (config 1 is the CPU-only plumbing case; configs 4 and 5 are
multi-GPU configurations in BASELINE — here: what one GPU's share costs).  CUDA events on the context's stream.
Writes gpurun_out/config_table.jsonl (copied to profiles/ by hand)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cloudsky_b200 as cs
from cloudsky_b200 import assets

lib = cs.load_product()
large, small, weather, _ = assets.load_default_textures()
stream = torch.cuda.Stream()
rows = []


def timed(fn, warm=1, iters=3):
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def ctx_for(W, H, cov, P, cone, mode):
    ctx = lib.context(0)
    ctx.set_stream(stream.cuda_stream)
    ctx.upload_textures(large, small, weather); ctx.build_transmittance_lut(); ctx.build_sky_lut((0.0, 1.0, 0.0)); ctx.resize(W, H)
    s = lib.settings_demo(); s.cloud_coverage = cov
    st = lib.frame_state_init(); st.light_direction[:] = [0.0, 1.0, 0.0]
    lib.frame_advance(st, s, 1.0)
    ctx.set_march_config(P, cone, mode)
    return ctx, lib.fill_cloud_params(s, st, W, H)


def emit(**kw):
    rows.append(kw); print(json.dumps(kw), flush=True)


# C2: 1024x512, 64 primary / 6 light (5 cone + 1 distant), transmittance + sky LUT precompute in the timed region
ctx, p = ctx_for(1024, 512, 0.2, 64, 5, cs.MODE_FAST)
px = 1024 * 512 - 1024 - 512 + 1
t_lut = timed(lambda: ctx.build_transmittance_lut(), 2, 5); s_lut = timed(lambda: ctx.build_sky_lut((0.0, 1.0, 0.0)), 2, 5); m = timed(lambda: ctx.render_frame(p), 2, 5)
allin = timed(lambda: (ctx.build_transmittance_lut(), ctx.build_sky_lut((0.0, 1.0, 0.0)), ctx.render_frame(p)), 2, 5)
emit(config="C2 1024x512 64/6", transmittance_lut_ms=round(t_lut, 4), sky_lut_ms=round(s_lut, 4), march_ms=round(m, 4), all_three_ms=round(allin, 4), mray_steps_s=round(px * 64 / allin / 1e3, 1))
ctx.close()
# C3: the headline (bench.py)
ctx, p = ctx_for(2048, 1024, 0.2, 128, 7, cs.MODE_FAST)
px = 2048 * 1024 - 2048 - 1024 + 1
m = timed(lambda: ctx.render_frame(p), 2, 5)
emit(config="C3 2048x1024 128/8", march_ms=round(m, 4), mray_steps_s=round(px * 128 / m / 1e3, 1))
# C4: one GPU's share of the 64-sun sweep over 8 GPUs = 8 suns of the C3 frame, sky LUTs included
th = np.pi * (np.arange(8) + 0.5) / 64
suns = np.stack([np.cos(th), np.sin(th), np.zeros(8)], 1).astype(np.float32)
out = torch.empty((8, 1024, 2048, 4), dtype=torch.float16, device="cuda")
m = timed(lambda: ctx.render_sun_batch_to(p, suns, out.data_ptr()), 1, 3)
emit(config="C4 share: 8 suns x C3 (first 8 of the 64-sun sweep: low suns)", total_ms=round(m, 3), ms_per_frame=round(m / 8, 4), mray_steps_s=round(px * 128 * 8 / m / 1e3, 1))
del out
ctx.close()
# C5: 8192x4096, 256 primary / 12 light (11 cone + 1 distant), coverage 1.0: fixed steps, the opt-in early out, and the adaptive
# per-direction step budget (cs_set_step_budget: never step finer than `len` metres) compared with the fixed-step render
px = 8192 * 4096 - 8192 - 4096 + 1
fixed_img = None
for mode, budget, name in ((cs.MODE_FAST, 0.0, "fixed steps"), (cs.MODE_FAST | cs.MODE_EARLY_OUT, 0.0, "early out"),
                           (cs.MODE_FAST, 19.53125, "step budget 19.53 m"), (cs.MODE_FAST, 30.0, "step budget 30 m"), (cs.MODE_FAST, 40.0, "step budget 40 m"),
                           (cs.MODE_FAST | cs.MODE_EARLY_OUT, 19.53125, "step budget 19.53 m + early out"),
                           (cs.MODE_FAST | cs.MODE_EARLY_OUT, 40.0, "step budget 40 m + early out"),
                           (cs.MODE_FAST | cs.MODE_TEX | cs.MODE_EARLY_OUT, 0.0, "texture unit + early out"),
                           (cs.MODE_FAST | cs.MODE_TEX | cs.MODE_EARLY_OUT, 19.53125, "texture unit + step budget 19.53 m + early out")):
    ctx, p = ctx_for(8192, 4096, 1.0, 256, 11, mode)
    ctx.set_step_budget(budget, 64)
    m = timed(lambda: ctx.render_frame(p), 1, 2)
    ctx.set_counters_enabled(True); ctx.render_frame(p); k = ctx.get_counters().as_dict(); ctx.set_counters_enabled(False)
    img = ctx.read_image().astype(np.float32)
    if fixed_img is None:
        fixed_img = img
    d = np.abs(img - fixed_img)[1:, 1:]
    ok = float((d <= 2e-3 + 1e-2 * np.abs(fixed_img[1:, 1:])).all(-1).mean())
    emit(config="C5 8192x4096 256/12 coverage 1.0", mode=name, march_ms=round(m, 2), mray_steps_s_nominal=round(px * 256 / m / 1e3, 1),
         mray_steps_s_executed=round(k["primary_steps"] / m / 1e3, 1), executed_step_fraction=round(k["primary_steps"] / (px * 256), 4),
         pixels_in_fast_tolerance_of_fixed=round(ok, 6), max_abs_vs_fixed=round(float(d.max()), 5), one_eighth_ms=round(m / 8, 2))
    ctx.close()
# the same budget on the thin-cloud headline sky, for the record (it does NOT hold parity there)
ctx, p = ctx_for(2048, 1024, 0.2, 256, 11, cs.MODE_FAST)
ctx.render_frame(p); ref = ctx.read_image().astype(np.float32)
ctx.set_step_budget(19.53125, 64); ctx.render_frame(p); img = ctx.read_image().astype(np.float32)
d = np.abs(img - ref)[1:, 1:]
emit(config="2048x1024 256/12 coverage 0.2 (thin cloud)", mode="step budget 19.53 m vs fixed", pixels_in_fast_tolerance_of_fixed=round(float((d <= 2e-3 + 1e-2 * np.abs(ref[1:, 1:])).all(-1).mean()), 6), max_abs_vs_fixed=round(float(d.max()), 5))
ctx.close()
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/config_table.jsonl", "w").write("\n".join(json.dumps(r) for r in rows) + "\n")

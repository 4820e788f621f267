#!/usr/bin/env python
"""Development aid: BASELINE config 4's shape on one GPU — 8 sun angles of the C3 frame through cs_render_sun_batch_to, with the
sun-batch kernel (4 suns per launch, shared primary march) and with one launch per sun (CLOUDSKY_SUN_BATCH=0)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cloudsky_b200 as cs
from cloudsky_b200 import assets

lib = cs.Library(sys.argv[sys.argv.index("--lib") + 1]) if "--lib" in sys.argv else cs.load_product()
large, small, weather, _ = assets.load_default_textures()
W, H, P, cone, N = 2048, 1024, 128, 7, 8
th = np.pi * (np.arange(N) + 0.5) / N
suns = np.stack([np.cos(th), np.sin(th), np.zeros(N)], 1).astype(np.float32)  # dir_k = (cos, sin, 0)
out = torch.zeros((N, H, W, 4), dtype=torch.float16, device="cuda")
stream = torch.cuda.Stream()  # not the default stream: cs_set_stream(NULL) would mean "the context's own stream"
ref = None
for cov in (0.2, 1.0):
    for batched in ("1", "0"):
        os.environ["CLOUDSKY_SUN_BATCH"] = batched
        ctx = lib.context(0)
        ctx.upload_textures(large, small, weather); ctx.build_transmittance_lut(); ctx.resize(W, H)
        s = lib.settings_demo(); s.cloud_coverage = cov
        st = lib.frame_state_init(); lib.frame_advance(st, s, 1.0)
        p = lib.fill_cloud_params(s, st, W, H)
        ctx.set_march_config(P, cone, cs.MODE_FAST)
        ctx.set_stream(stream.cuda_stream)
        best = 1e9
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ctx.render_sun_batch_to(p, suns, out.data_ptr()); e1.record(stream); torch.cuda.synchronize()
            if it: best = min(best, e0.elapsed_time(e1))
        img = out.cpu().numpy()
        same = None if batched == "1" else bool((img.view(np.uint16) == ref.view(np.uint16)).all())
        ref = img
        print(json.dumps({"coverage": cov, "sun_batch_kernel": batched == "1", "suns": N, "ms_total": round(best, 3), "ms_per_frame": round(best / N, 4),
                          "mray_steps_s": round((W * H - W - H + 1) * P * N / best / 1e3, 1), "same_bits_as_batched": same}), flush=True)
        ctx.set_stream(0); ctx.close()

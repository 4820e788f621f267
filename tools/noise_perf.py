#!/usr/bin/env python
"""Development aid: wall time of cs_generate_noise (kernel + device->host copy) at the reference sizes and larger."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cloudsky_b200 as cs

lib = cs.load_product()
ctx = lib.context(0)
ctx.generate_noise(cs.NOISE_SMALL, 16)  # warm-up (module load)
for kind, name, n in ((cs.NOISE_LARGE, "large", 128), (cs.NOISE_LARGE, "large", 256), (cs.NOISE_SMALL, "small", 32), (cs.NOISE_SMALL, "small", 128),
                      (cs.NOISE_WEATHER, "weather", 512), (cs.NOISE_WEATHER, "weather", 4096)):
    best = 1e9
    for _ in range(3):
        t = time.perf_counter(); a = ctx.generate_noise(kind, n); best = min(best, time.perf_counter() - t)
    texels = a.size // 4
    print(json.dumps({"kind": name, "n": n, "ms": round(best * 1e3, 3), "Mtexels_per_s": round(texels / best / 1e6, 1), "mean": [round(float(v), 4) for v in a.reshape(-1, 4).mean(0) / 255.0]}), flush=True)
ctx.close()

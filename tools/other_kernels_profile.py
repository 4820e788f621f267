#!/usr/bin/env python
"""Launches every kernel of the library that bench.py's headline does not (profiling target:
ncu --set full -k regex:'sunbatch|strict|composite|noise|peer_barrier|prologue')."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cloudsky_b200 as cs
from cloudsky_b200 import assets
import bench

lib = cs.load_product()
large, small, weather, _ = assets.load_default_textures()
ctx = lib.context(0)
ctx.upload_textures(large, small, weather); ctx.build_transmittance_lut(); ctx.build_sky_lut((0, 1, 0)); ctx.resize(bench.W, bench.H)
p = bench.frame_params(lib, 0, (0.0, 1.0, 0.0))
ctx.set_march_config(bench.PRIMARY, bench.CONE, cs.MODE_FAST)
from cloudsky_b200 import sharding
suns = sharding.sun_sweep(64)[:4]
out = torch.empty((4, bench.H, bench.W, 4), dtype=torch.float16, device="cuda")
for _ in range(2):
    ctx.render_sun_batch_to(p, suns, out.data_ptr())                     # clouds_fast_sunbatch_kernel (4 suns of the C3 frame)
flags, _ = ctx.peer_alloc(256)
for e in (1, 2):
    ctx.peer_barrier(0, 1, [flags], e)                                  # peer_barrier_kernel
ctx.resize(512, 256)
q = bench.frame_params(lib, 0, (0.0, 1.0, 0.0), 512, 256)
ctx.set_march_config(128, 6, cs.MODE_STRICT)
for _ in range(2):
    ctx.render_frame(q)                                                 # clouds_strict_kernel (512x256)
ctx.set_march_config(128, 6, cs.MODE_FAST)
sky = cs.Sky(ctx, lib.settings_demo())
sky.set_sun(cs.DirectionalLight.looking_from((0.3, 0.6, 0.2)).basis, 1.0, (1.0, 1.0, 1.0))
sky.update(1.0)
view = cs.View.equirect(1024, 512, (0.3, 0.6, 0.2))
sky.composite(view); sky.composite(view)                                # composite_kernel
for kind, n in ((cs.NOISE_LARGE, 128), (cs.NOISE_SMALL, 32), (cs.NOISE_WEATHER, 512)):
    ctx.generate_noise(kind, n)                                         # noise_kernel<kind>
ctx.sync()
print("ok")

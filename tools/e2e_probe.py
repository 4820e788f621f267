#!/usr/bin/env python
"""Development probe: per-step wall time of the synchronous vs streaming host-readback paths."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cloudsky_b200 as cs
from cloudsky_b200 import assets
import bench

lib = cs.load_product()
large, small, weather, _ = assets.load_default_textures()
ctx = lib.context(0)
ctx.upload_textures(large, small, weather); ctx.build_transmittance_lut(); ctx.resize(bench.W, bench.H)
ctx.set_march_config(bench.PRIMARY, bench.CONE, cs.MODE_FAST)
params = [bench.frame_params(lib, k, (0.0, 1.0, 0.0)) for k in range(16)]
hosts = [torch.empty((bench.H, bench.W, 4), dtype=torch.float16).pin_memory() for _ in range(2)]
N = 40
def timeit(name, fn, fin=lambda: None):
    for k in range(5): fn(k)
    fin(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(N): fn(k)
    fin(); torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / N * 1e3:.3f} ms/step", flush=True)
for use_torch_stream in (False, True):
    if use_torch_stream:
        s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
    tag = "torch-stream" if use_torch_stream else "own-stream"
    timeit(f"{tag} kernels only (render_frame)", lambda k: (ctx.build_sky_lut((0, 1, 0)), ctx.render_frame(params[k % 16])), ctx.sync)
    timeit(f"{tag} sync host", lambda k: ctx.render_frame_host(params[k % 16], out_ptr=hosts[0].data_ptr()))
    timeit(f"{tag} async host, 2 buffers", lambda k: ctx.render_frame_host_async(params[k % 16], hosts[k & 1].data_ptr()), ctx.wait_host)
    timeit(f"{tag} async host, 1 buffer", lambda k: ctx.render_frame_host_async(params[k % 16], hosts[0].data_ptr()), ctx.wait_host)

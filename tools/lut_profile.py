#!/usr/bin/env python
"""Launches the two LUT kernels a few times (profiling target: ncu -k regex:lut_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cloudsky_b200 as cs

lib = cs.load_product()
ctx = lib.context(0)
for i in range(4):
    ctx.build_transmittance_lut()
    ctx.build_sky_lut((0.3 * i, 1.0, 0.1))
ctx.sync()
ctx.close()

#!/usr/bin/env python
"""Study, CPU only (it executes the oracle, so it lives under tests/): what would hierarchical ray marching — the first
TODO of the reference's README (README.md:28) and the "hierarchical/adaptive step" of BASELINE config 5 — buy on this
workload, and what would it cost in accuracy?

Scheme evaluated (oracle hook cso_set_hierarchical): primary steps grouped into blocks of `stride`; one probe per block
(density() of clouds.glsl:109-126 without the detail erosion, at the centre of the block's sample positions, from the
large-volume mip level whose texel matches the block length) decides whether the block is marched (probe > -margin) or
skipped.  Sample positions and everything evaluated at a marched sample are the fixed-step march's own.

Prints, per case and setting: fraction of pixels inside the FAST parity tolerance of the fixed-step render, max abs error,
and the executed primary steps / lit steps / density evaluations relative to the fixed-step march.
usage: python tests/hierarchical_study.py [W H [P Lc]]      (results quoted in DESIGN.md section 8)"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import cloudsky_b200 as cs
from cloudsky_b200 import assets
from conftest import ORACLE_LIB, _build_oracle, compare_images, make_params, prepared_context


def main():
    W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 128)
    P, cone = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (128, 6)
    _build_oracle()
    ora = cs.Library(ORACLE_LIB)
    ora.dll.cso_set_hierarchical.restype = C.c_int
    ora.dll.cso_set_hierarchical.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int]
    tex = assets.load_fixture()
    cases = [("noon, coverage 0.2", (0, 1, 0), 0.2), ("low sun, coverage 0.2", (0.9, 0.12, 0.4), 0.2), ("noon, coverage 1.0", (0, 1, 0), 1.0)]
    settings = [(2, 0.05), (4, 0.0), (4, 0.05), (4, 0.1), (4, 0.2), (8, 0.1)]
    rows = []
    for name, sun, cov in cases:
        ctx = prepared_context(ora, tex, W, H, sun=sun, threads=os.cpu_count() or 1)
        p = make_params(ora, W, H, sun=sun, coverage=cov, time=2.5)
        ctx.set_march_config(P, cone, cs.MODE_FAST)
        ctx.render_frame(p)
        ref = ctx.read_image().astype(np.float32)
        k0 = ctx.get_counters().as_dict()
        for stride, margin in settings:
            assert ora.dll.cso_set_hierarchical(ctx._h, stride, margin, 0) == 0
            ctx.render_frame(p)
            img = ctx.read_image().astype(np.float32)
            k = ctx.get_counters().as_dict()
            ok, mx = compare_images(img, ref, 2e-3, 1e-2)
            rows.append(dict(case=name, W=W, H=H, P=P, cone=cone, stride=stride, margin=margin, pass_fast_tol=round(ok, 5), max_abs=round(mx, 4),
                             primary_steps=round(k["primary_steps"] / k0["primary_steps"], 3), lit_steps=round(k["lit_steps"] / k0["lit_steps"], 4),
                             density_evals=round(k["density_evals"] / k0["density_evals"], 3)))
            print(json.dumps(rows[-1]), flush=True)
        assert ora.dll.cso_set_hierarchical(ctx._h, 0, 0.0, 0) == 0
        ctx.close()
    return rows


if __name__ == "__main__":
    main()

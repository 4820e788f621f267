"""The hierarchical-march study hook of the oracle (cso_set_hierarchical): off by default and
identical to the fixed-step march when off; when on, the numbers DESIGN.md quotes hold in kind (few pixels change, about half
of the primary steps are skipped, only about a tenth of the density evaluations)."""
import ctypes as C

import numpy as np


def test_hierarchical_hook(cs, oracle_lib, textures, helpers):
    d = oracle_lib.dll
    d.cso_set_hierarchical.restype = C.c_int
    d.cso_set_hierarchical.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int]
    W, H = 128, 64
    ctx = helpers.prepared_context(oracle_lib, textures, W, H, threads=helpers.cpu_threads)
    p = helpers.make_params(oracle_lib, W, H, time=2.5)
    ctx.set_march_config(128, 6, cs.MODE_FAST)
    ctx.render_frame(p)
    ref, k0 = ctx.read_image().copy(), ctx.get_counters().as_dict()
    assert d.cso_set_hierarchical(ctx._h, 1, 0.1, 0) != 0 and d.cso_set_hierarchical(ctx._h, 4, -1.0, 0) != 0  # bad arguments
    assert d.cso_set_hierarchical(ctx._h, 4, 0.1, 0) == 0
    ctx.render_frame(p)
    img, k = ctx.read_image().copy(), ctx.get_counters().as_dict()
    ok, mx = helpers.compare_images(img, ref, 2e-3, 1e-2)
    assert 0.99 < ok < 1.0 and mx < 0.5                               # inexact, by a little
    assert 0.4 < k["primary_steps"] / k0["primary_steps"] < 0.7       # about half of the primary steps are skipped ...
    assert 0.8 < k["density_evals"] / k0["density_evals"] < 0.95      # ... but only about a tenth of the work
    assert k["lit_steps"] <= k0["lit_steps"] and k["lit_steps"] > 0.995 * k0["lit_steps"]
    assert d.cso_set_hierarchical(ctx._h, 0, 0.0, 0) == 0              # off again: bit-identical to the fixed-step march
    ctx.render_frame(p)
    assert (ctx.read_image().view(np.uint16) == ref.view(np.uint16)).all()
    ctx.close()

"""Committed golden vectors (tests/golden/oracle_golden.npz, made by tests/golden/make_oracle_golden.py).
CPU: the oracle still reproduces them.  GPU: the CUDA path matches them without running the oracle."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "oracle_golden.npz")
CASES = ["noon", "sunset_wind", "overcast"]
W, H = 64, 32


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def half_ulps(a, b):
    """Distance in fp16 representable steps (both non-negative here)."""
    return np.abs(a.view(np.int16).astype(np.int32) - b.view(np.int16).astype(np.int32))


def test_oracle_reproduces_golden(cs, oracle_lib, textures, helpers, gold):
    ctx = helpers.prepared_context(oracle_lib, textures, W, H, threads=helpers.cpu_threads)
    # libm's expf/powf may be dispatched to different (FMA / non-FMA) variants on other CPUs: allow 1 fp16 ulp on a few texels
    d = half_ulps(ctx.read_transmittance_lut(), gold["transmittance"])
    assert d.max() <= 1 and (d > 0).mean() < 0.01
    for name in CASES:
        p = cs.CloudParams.from_floats(gold[f"params_{name}"])
        ctx.build_sky_lut(tuple(p.light_direction))
        d = half_ulps(ctx.read_sky_lut(), gold[f"sky_{name}"])
        assert d.max() <= 2 and (d > 0).mean() < 0.02, name
        ctx.write_sky_lut(gold[f"sky_{name}"])
        ctx.set_march_config(128, 6)
        ctx.render_frame(p)
        img = ctx.read_image()
        frac, mx = helpers.compare_images(img, gold[f"clouds_{name}"], 2e-4, 2e-3, skip_edges=False)
        assert frac >= 0.999 and mx < 5e-3, (name, frac, mx)
    ctx.close()


@pytest.mark.gpu
def test_gpu_matches_committed_golden(cs, product_lib, textures, helpers, gold):
    ctx = helpers.prepared_context(product_lib, textures, W, H)
    a = ctx.read_transmittance_lut().astype(np.float32)
    g = gold["transmittance"].astype(np.float32)
    assert (np.abs(a - g) <= 1e-3 + 2e-3 * np.abs(g)).all()
    for name in CASES:
        p = cs.CloudParams.from_floats(gold[f"params_{name}"])
        ctx.build_sky_lut(tuple(p.light_direction))
        a = ctx.read_sky_lut().astype(np.float32)
        g = gold[f"sky_{name}"].astype(np.float32)
        assert (np.abs(a - g) <= 1e-3 + 2e-3 * np.abs(g)).all(), name
        ctx.write_sky_lut(gold[f"sky_{name}"])
        for mode, atol, rtol, need in ((cs.MODE_STRICT, 1e-3, 2e-3, 0.999), (cs.MODE_FAST, 2e-3, 1e-2, 0.999)):
            ctx.set_march_config(128, 6, mode)
            ctx.render_frame(p)
            frac, mx = helpers.compare_images(ctx.read_image(), gold[f"clouds_{name}"], atol, rtol)
            assert frac >= need, (name, mode, frac, mx)
    ctx.close()

"""Committed golden vectors of the two extensions without a reference implementation (tests/golden/extensions_golden.npz,
made by tests/golden/make_extensions_golden.py): the noise generator's bytes (SHA-256 of whole textures + a readable slice)
and the Bruneton-mapped LUTs.  CPU: the oracle still reproduces them.  GPU: the CUDA path matches them without the oracle."""
import hashlib
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "extensions_golden.npz")
SUN = (0.3, 0.5, -0.81)


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def _noise_cases(cs):
    return [("large", cs.NOISE_LARGE, 32), ("small", cs.NOISE_SMALL, 32), ("weather", cs.NOISE_WEATHER, 128)]


def _check_noise(cs, lib, ctx, gold):
    for name, kind, n in _noise_cases(cs):
        for tag, seed in (("default", None), ("seed77", 77)):
            p = lib.noise_params_default(kind)
            if seed is not None:
                p.seed = seed
                p.worley_frequency = 3
            a = ctx.generate_noise(kind, n, p)
            sl = a[0, :8, :8] if kind != cs.NOISE_WEATHER else a[:8, :8]
            assert (sl == gold[f"noise_{name}_{tag}_slice"]).all(), (name, tag)
            assert hashlib.sha256(a.tobytes()).digest() == gold[f"noise_{name}_{tag}_sha256"].tobytes(), (name, tag)


def _half_ulps(a, b):
    return np.abs(a.view(np.int16).astype(np.int32) - b.view(np.int16).astype(np.int32))


def test_oracle_reproduces_extension_golden(cs, oracle_lib, helpers, gold):
    ctx = oracle_lib.context(0)
    ctx.set_threads(helpers.cpu_threads)
    _check_noise(cs, oracle_lib, ctx, gold)  # integer hash + exactly rounded fp32 only: byte-exact on any CPU
    ctx.set_transmittance_parametrisation(cs.TLUT_BRUNETON2017)
    ctx.build_transmittance_lut()
    d = _half_ulps(ctx.read_transmittance_lut()[::2, ::4], gold["transmittance_bruneton"])
    assert d.max() <= 1 and (d > 0).mean() < 0.01  # libm expf/powf variants, as in test_oracle_golden.py
    ctx.build_sky_lut(SUN)
    d = _half_ulps(ctx.read_sky_lut()[::4, ::4], gold["sky_bruneton"])
    assert d.max() <= 2 and (d > 0).mean() < 0.02
    ctx.close()


@pytest.mark.gpu
def test_gpu_matches_extension_golden(cs, product_lib, gold):
    ctx = product_lib.context(0)
    _check_noise(cs, product_lib, ctx, gold)
    ctx.set_transmittance_parametrisation(cs.TLUT_BRUNETON2017)
    ctx.build_transmittance_lut()
    a, g = ctx.read_transmittance_lut()[::2, ::4].astype(np.float32), gold["transmittance_bruneton"].astype(np.float32)
    assert (np.abs(a - g) <= 1e-3 + 2e-3 * np.abs(g)).all()
    ctx.build_sky_lut(SUN)
    a, g = ctx.read_sky_lut()[::4, ::4].astype(np.float32), gold["sky_bruneton"].astype(np.float32)
    assert (np.abs(a - g) <= 1e-3 + 2e-3 * np.abs(g)).all()
    ctx.close()

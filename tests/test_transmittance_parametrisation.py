"""CS_TLUT_BRUNETON2017 (cs_set_transmittance_parametrisation; the reference README's TODO 2).

CPU: the mapping is a bijection texel -> ray -> texel centre; the coordinate arithmetic the CUDA kernels execute
(csrc/tlut_param.h, compiled with g++ by tests/tlut_host_check.cpp) equals the oracle's; lookups through the oracle's
Bruneton LUT are several times closer to a float64 4000-step integral of the same extinction model than lookups through
the reference's linear LUT, most of all near the horizon; the default mapping is untouched.  GPU: both LUT kernels and the
composite match the oracle in this mode at the usual LUT tolerance."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS_SRC = os.path.join(ROOT, "tests", "tlut_host_check.cpp")
HARNESS_LIB = os.path.join(ROOT, "build", "libtlut_host_check.so")
HDR = os.path.join(ROOT, "godot-volumetric-cloud-demo-v2_b200", "csrc", "tlut_param.h")
RG, THICK = 6371.0, 100.0
RT = RG + THICK
W, H = 256, 64


def _hooks(oracle_lib):
    d = oracle_lib.dll
    f, fp = C.c_float, C.POINTER(C.c_float)
    d.cso_transmittance_lookup.restype = C.c_int; d.cso_transmittance_lookup.argtypes = [C.c_void_p, f, f, fp]
    d.cso_bruneton_texel_ray.restype = None; d.cso_bruneton_texel_ray.argtypes = [C.c_int, C.c_int, fp]
    d.cso_bruneton_lookup_coords.restype = None; d.cso_bruneton_lookup_coords.argtypes = [f, f, fp]
    return d


@pytest.fixture(scope="module")
def harness():
    os.makedirs(os.path.dirname(HARNESS_LIB), exist_ok=True)
    if not os.path.exists(HARNESS_LIB) or os.path.getmtime(HARNESS_LIB) < max(os.path.getmtime(HARNESS_SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-std=c++17", "-shared", "-o", HARNESS_LIB, HARNESS_SRC])
    dll = C.CDLL(HARNESS_LIB)
    fp = C.POINTER(C.c_float)
    dll.tlh_ray_from_texel.restype = None; dll.tlh_ray_from_texel.argtypes = [C.c_int, C.c_int, fp]
    dll.tlh_uv.restype = None; dll.tlh_uv.argtypes = [C.c_float, C.c_float, fp]
    return dll


def test_mapping_round_trip_and_kernel_arithmetic(oracle_lib, harness):
    d = _hooks(oracle_lib)
    a4, b4, a3, b3 = (C.c_float * 4)(), (C.c_float * 4)(), (C.c_float * 3)(), (C.c_float * 3)()
    for py in (0, 1, 7, 31, 62, 63):
        for px in (0, 1, 50, 128, 254, 255):
            d.cso_bruneton_texel_ray(px, py, a4)
            harness.tlh_ray_from_texel(px, py, b4)
            assert list(a4) == list(b4)
            alt, r, mu, dist = a4
            # analytic facts about the stored ray: it starts between the two radii, misses the ground, ends on the top boundary
            assert -1e-3 <= alt <= THICK + 1e-3 and abs(r - (RG + alt)) < 2e-3
            end = np.hypot(r + dist * mu, dist * np.sqrt(max(1.0 - mu * mu, 0.0)))
            assert abs(end - RT) < 0.05
            mu_h = -np.sqrt(max(1.0 - (RG / r) ** 2, 0.0))
            assert mu >= mu_h - 2e-3
            # texel -> ray -> texture coordinate lands on the texel centre
            d.cso_bruneton_lookup_coords(alt / THICK, mu, a3)
            harness.tlh_uv(alt / THICK, mu, b3)
            assert np.allclose(list(a3), list(b3), rtol=0, atol=1e-7)
            assert abs(a3[0] * W - (px + 0.5)) < 0.15 and abs(a3[1] * H - (py + 0.5)) < 0.02, (px, py, list(a3))
    # the planet hides the sun below the horizon, shows it above, half of it on the horizon
    for alt in (0.0, 0.002, 0.3):
        r = RG + alt * THICK
        mu_h = -np.sqrt(max(1.0 - (RG / r) ** 2, 0.0))
        for dm, want in ((-0.02, 0.0), (0.02, 1.0), (0.0, 0.5)):
            d.cso_bruneton_lookup_coords(alt, mu_h + dm, a3)
            assert abs(a3[2] - want) < 0.02


def _extinction(h):
    """float64 restatement of get_atmosphere_collision_coefficients (transmittance-lut.glsl:100-145), 4 wavelengths x samples."""
    h = np.maximum(h, 0.0)
    aerosol = 1.3681e20 * (np.exp(-h / 0.73) + 2e6 / 1.3681e20)
    aer = (np.array([2.8722e-24, 4.6168e-24, 7.9706e-24, 1.3578e-23]) + np.array([1.5908e-22, 1.7711e-22, 2.0942e-22, 2.4033e-22]))[:, None] * aerosol
    ho = h + 1e-4
    t = np.log(ho) - 3.22261
    ozone = (np.array([3.472e-21, 3.914e-21, 1.349e-21, 11.03e-23]) * 1e-4 * 350.0)[:, None] * (3.78547397e20 / ho * np.exp(-t * t * 5.55555555))
    rayleigh = np.array([6.605e-3, 1.067e-2, 1.842e-2, 3.156e-2])[:, None] * np.exp(-0.07771971 * h ** 1.16364243)
    return aer + ozone + rayleigh


def _truth(alt_km, mu, n=4000):
    r = RG + alt_km
    dist = -r * mu + np.sqrt(r * r * (mu * mu - 1.0) + RT * RT)
    t = (np.arange(n) + 0.5) * dist / n
    rr = np.sqrt(r * r + t * t + 2.0 * r * mu * t)
    return np.exp(-(_extinction(rr - RG) * (dist / n)).sum(1))


def test_bruneton_lut_is_closer_to_the_integral(cs, oracle_lib, helpers):
    d = _hooks(oracle_lib)
    rng = np.random.default_rng(1)
    samples = []
    for k in range(600):
        alt = rng.uniform(0, 1) ** 3 * 60.0
        r = RG + alt
        mu_h = -np.sqrt(max(1.0 - (RG / r) ** 2, 0.0))
        mu = min(mu_h + 0.006 + (rng.uniform(0, 0.15) if k % 2 else rng.uniform(0, 1.0 - mu_h - 0.006)), 1.0)
        samples.append((alt, mu, mu - mu_h, _truth(alt, mu)))
    err = {}
    out = (C.c_float * 4)()
    for name, mode in (("linear", cs.TLUT_LINEAR), ("bruneton", cs.TLUT_BRUNETON2017)):
        ctx = oracle_lib.context(0)
        ctx.set_threads(helpers.cpu_threads)
        ctx.set_transmittance_parametrisation(mode)
        ctx.build_transmittance_lut()
        e = []
        for alt, mu, _, want in samples:
            assert d.cso_transmittance_lookup(ctx._h, mu, alt / THICK, out) == 0
            e.append(np.abs(np.array(out[:]) - want).max())
        err[name] = np.array(e)
        ctx.close()
    near = np.array([s[2] < 0.05 for s in samples])
    # measured: linear max 0.18 / mean 0.03, Bruneton max 0.015 / mean 0.002
    assert err["bruneton"].max() < 0.03 and err["bruneton"].mean() < 0.005
    assert err["bruneton"].max() * 5 < err["linear"].max() and err["bruneton"].mean() * 5 < err["linear"].mean()
    assert err["bruneton"][near].mean() * 10 < err["linear"][near].mean()


def test_switching_invalidates_and_default_is_untouched(cs, oracle_lib, helpers):
    ctx = oracle_lib.context(0)
    ctx.set_threads(helpers.cpu_threads)
    ctx.build_transmittance_lut()
    ref = ctx.read_transmittance_lut().copy()
    ctx.set_transmittance_parametrisation(cs.TLUT_LINEAR)  # no change: LUT stays valid
    assert (ctx.read_transmittance_lut() == ref).all()
    ctx.set_transmittance_parametrisation(cs.TLUT_BRUNETON2017)
    with pytest.raises(cs.CloudSkyError):
        ctx.read_transmittance_lut()
    with pytest.raises(cs.CloudSkyError):
        ctx.build_sky_lut((0.0, 1.0, 0.0))
    ctx.build_transmittance_lut()
    b = ctx.read_transmittance_lut().astype(np.float32)
    assert np.isfinite(b).all() and b.min() >= 0.0 and b.max() <= 1.0
    assert (np.diff(b[:, :, 2], axis=1) <= 1e-3).all()   # along x_mu the path to the top boundary grows: transmittance falls
    ctx.build_sky_lut((0.3, 0.6, 0.2))
    sky_b = ctx.read_sky_lut().astype(np.float32)
    ctx.set_transmittance_parametrisation(cs.TLUT_LINEAR)
    ctx.build_transmittance_lut()
    assert (ctx.read_transmittance_lut() == ref).all()
    ctx.build_sky_lut((0.3, 0.6, 0.2))
    sky_l = ctx.read_sky_lut().astype(np.float32)
    # same physics, better-resolved transmittance: the sky-view LUT moves by a few per cent, it does not change character
    rel = np.abs(sky_b[..., :3] - sky_l[..., :3]) / (np.abs(sky_l[..., :3]) + 0.05)
    assert np.isfinite(sky_b).all() and np.median(rel) < 0.1
    with pytest.raises(cs.CloudSkyError):
        ctx.set_transmittance_parametrisation(7)
    ctx.close()


def _sky_with_mapping(cs, lib, textures, sun, threads=None, mapping=None):
    ctx = lib.context(0)
    if threads:
        ctx.set_threads(threads)
    ctx.upload_textures(*textures)
    ctx.set_transmittance_parametrisation(cs.TLUT_BRUNETON2017 if mapping is None else mapping)
    ctx.build_transmittance_lut()
    ctx.set_march_config(32, 4, cs.MODE_FAST)
    s = lib.settings_demo()
    s.texture_size, s.frames_to_update, s.cloud_coverage, s.sun_disk_scale = 64, 4, 0.3, 2.0
    sky = cs.Sky(ctx, s)
    sky.set_sun(cs.DirectionalLight.looking_from(sun).basis, 1.0, (1.0, 1.0, 1.0))
    for k in range(3):
        sky.update(1.0 + k)
    return ctx, sky


def test_composite_uses_the_mapping(cs, oracle_lib, small_textures, helpers):
    """Inside the sun's disc the composite adds transmittance(viewPos, LIGHT0_DIRECTION) to the sky (clouds.gdshader:96-99).
    The sky term does not depend on the view's sun uniform, so shading the same camera with the disc elsewhere isolates it."""
    sun = (0.0, 0.05, -0.99875)
    cam = cs.DirectionalLight.looking_from(tuple(-c for c in sun)).basis
    d = _hooks(oracle_lib)
    out = (C.c_float * 4)()
    trans = {}
    for mapping in (cs.TLUT_LINEAR, cs.TLUT_BRUNETON2017):
        ctx, sky = _sky_with_mapping(cs, oracle_lib, small_textures, sun, threads=helpers.cpu_threads, mapping=mapping)
        with_disc = sky.composite(cs.View.perspective(33, 33, cam, 4.0, sun, 2.0))[16, 16, :3].astype(np.float64)
        without = sky.composite(cs.View.perspective(33, 33, cam, 4.0, (0.0, 0.05, 0.99875), 2.0))[16, 16, :3].astype(np.float64)
        assert d.cso_transmittance_lookup(ctx._h, sun[1], 0.002, out) == 0
        t = trans[mapping] = np.array(out[:3], np.float64)
        added = with_disc - without
        # horizon fade (clouds.gdshader:115) is 0.96 at this elevation, so at least 96 % of the term survives any cloud alpha
        assert np.all(added <= t * 1.001 + 1e-4) and np.all(added >= t * 0.95 - 1e-4), (mapping, added, t)
        sky.close(); ctx.close()
    tb, tl_ = trans[cs.TLUT_BRUNETON2017], trans[cs.TLUT_LINEAR]
    assert 0.0 < tb[2] < tb[0] < 1.0
    assert np.abs(tb - tl_).max() > 0.01  # 2.9 degrees above the horizon the two mappings disagree visibly


@pytest.mark.gpu
def test_gpu_bruneton_luts_and_composite_match_oracle(cs, oracle_lib, product_lib, textures, helpers):
    sun = (0.55, 0.25, 0.3)
    n = float(np.sqrt(sum(v * v for v in sun)))
    sun = tuple(v / n for v in sun)
    res = {}
    for name, lib in (("gpu", product_lib), ("oracle", oracle_lib)):
        ctx, sky = _sky_with_mapping(cs, lib, textures, sun, threads=helpers.cpu_threads if lib is oracle_lib else None)
        ctx.build_sky_lut(sun)
        cam = cs.DirectionalLight.looking_from(tuple(-c for c in sun)).basis
        comp = sky.composite(cs.View.perspective(128, 96, cam, 40.0, sun, 2.0))
        res[name] = (ctx.read_transmittance_lut().astype(np.float32), ctx.read_sky_lut().astype(np.float32), comp)
        sky.close(); ctx.close()
    for i, what in enumerate(("transmittance LUT", "sky LUT")):
        g, o = res["gpu"][i], res["oracle"][i]
        assert np.isfinite(g).all()
        assert (np.abs(g - o) <= 1e-3 + 2e-3 * np.abs(o)).all(), (what, float(np.abs(g - o).max()))
    g, o = res["gpu"][2], res["oracle"][2]
    ok = (np.abs(g - o) <= 2e-3 + 1e-2 * np.abs(o)).all(-1).mean()
    assert ok >= 0.998, ok
    # the default mapping is still the reference's, bit for bit, after switching there and back
    ctx = product_lib.context(0)
    ctx.build_transmittance_lut()
    ref = ctx.read_transmittance_lut().copy()
    ctx.set_transmittance_parametrisation(cs.TLUT_BRUNETON2017); ctx.build_transmittance_lut()
    assert (ctx.read_transmittance_lut() != ref).any()
    ctx.set_transmittance_parametrisation(cs.TLUT_LINEAR); ctx.build_transmittance_lut()
    assert (ctx.read_transmittance_lut() == ref).all()
    ctx.close()

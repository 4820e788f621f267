"""Pins the CPU oracle (the parity checker) with analytic known-answer tests derived from the shader
source, because the reference ships no tests, golden vectors or fixtures and cannot
be executed here: PARITY UNPINNED against the reference itself."""
import ctypes as C
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def probe(oracle_lib):
    d = oracle_lib.dll
    f, fp, u16 = C.c_float, C.POINTER(C.c_float), C.c_uint16
    d.cso_intersect_sphere.restype = f; d.cso_intersect_sphere.argtypes = [fp, f]
    d.cso_hash.restype = f; d.cso_hash.argtypes = [fp]
    d.cso_henyey_greenstein.restype = f; d.cso_henyey_greenstein.argtypes = [f, f]
    d.cso_remap.restype = f; d.cso_remap.argtypes = [f] * 5
    d.cso_height_fraction.restype = f; d.cso_height_fraction.argtypes = [f]
    d.cso_density_height_gradient.restype = f; d.cso_density_height_gradient.argtypes = [f, f]
    d.cso_oct_to_dir.restype = None; d.cso_oct_to_dir.argtypes = [f, f, fp]
    d.cso_f32_to_f16.restype = u16; d.cso_f32_to_f16.argtypes = [f]
    d.cso_f16_to_f32.restype = f; d.cso_f16_to_f32.argtypes = [u16]
    d.cso_sample_volume.restype = C.c_int; d.cso_sample_volume.argtypes = [C.c_void_p, C.c_int, fp, f, fp]
    d.cso_sample_weather.restype = C.c_int; d.cso_sample_weather.argtypes = [C.c_void_p, f, f, fp]
    d.cso_sample_lut.restype = C.c_int; d.cso_sample_lut.argtypes = [C.c_void_p, C.c_int, f, f, fp]
    d.cso_density.restype = f; d.cso_density.argtypes = [C.c_void_p, C.c_void_p, fp, fp, f]
    d.cso_ray_setup.restype = None; d.cso_ray_setup.argtypes = [fp, C.c_int, fp]
    return d


def vec(*v):
    return (C.c_float * len(v))(*v)


def ray_setup(probe, d, steps=128):
    out = (C.c_float * 6)()
    probe.cso_ray_setup(vec(*d), steps, out)
    return list(out)


# ---- geometry --------------------------------------------------------------------------------------
def test_slab_geometry(probe):
    """intersectSphere + sky(): zenith start 1500 m, end 4000 m, step 2500/128 (clouds.glsl:97-105,218-237)."""
    t0, t1, shell, step, _, r0 = ray_setup(probe, (0.0, 1.0, 0.0))
    assert t0 == 1500.0 and t1 == 4000.0 and shell == 2500.0
    assert abs(step - 19.53125) < 1e-4 and r0 == 6001500.0
    e = math.radians(30.0)
    _, _, shell, _, _, _ = ray_setup(probe, (math.cos(e), math.sin(e), 0.0))
    R, b, t = 6000000.0, 6001500.0, 6004000.0
    exact = lambda s, r: -R * s + math.sqrt(R * R * s * s + r * r - R * R)
    assert abs(shell - (exact(math.sin(e), t) - exact(math.sin(e), b))) < 3.0  # fp32 cancellation: ~2 m (SURVEY key fact 3)
    assert 4985 < shell < 5001
    e = math.radians(0.1)
    _, _, shell, _, _, _ = ray_setup(probe, (math.cos(e), math.sin(e), 0.0))
    assert abs(shell - (exact(math.sin(e), t) - exact(math.sin(e), b))) < 8.0
    assert 84000 < shell < 85600


def test_jitter_hash_is_identically_zero_in_fp32(probe):
    """SURVEY key fact 2: pos.y*10*0.3183099 > 2^24, so fract() of that component is 0 and the product is 0."""
    rng = np.random.default_rng(0)
    for _ in range(2000):
        az, el = rng.uniform(0, 2 * math.pi), rng.uniform(0.001, math.pi / 2)
        d = (math.cos(el) * math.cos(az), math.sin(el), math.cos(el) * math.sin(az))
        assert ray_setup(probe, d)[4] == 0.0
    assert probe.cso_hash(vec(0.3, 0.7, 0.2)) != 0.0  # the hash itself is not degenerate for small arguments


def vec3_to_oct(e):
    """clouds.gdshader:22-32 (the presentation shader's inverse mapping) in float64."""
    e = np.asarray(e, np.float64) / np.abs(e).sum()
    ny = e[1] * 0.5 + 0.5
    return np.array([e[0] * 0.5 + ny, e[0] * -0.5 + ny])


def test_oct_mapping_round_trip(probe):
    rng = np.random.default_rng(1)
    out = (C.c_float * 3)()
    for _ in range(500):
        u, v = rng.uniform(0.02, 0.98, 2)
        probe.cso_oct_to_dir(u, v, out)
        d = np.array(list(out), np.float64)          # .xzy applied: (n.x, n.z, n.y)
        assert abs(np.linalg.norm(d) - 1.0) < 1e-6
        n_z = 1.0 - abs(u - v) - abs(u + v - 1.0)
        assert (d[1] > 0) == (n_z > 0)
        if n_z > 0.01:                               # open hemisphere: composition is the identity
            uv = vec3_to_oct([d[0], d[2], d[1]])     # norm.xz = vec3_to_oct(norm.xzy) (clouds.gdshader:109)
            np.testing.assert_allclose(uv, [u, v], atol=2e-6)
    probe.cso_oct_to_dir(0.5, 0.5, out)
    np.testing.assert_allclose(list(out), [0.0, 1.0, 0.0], atol=1e-7)  # texture centre = zenith


# ---- scalar helpers ----------------------------------------------------------------------------------
def test_henyey_greenstein(probe):
    for g in (0.6, -0.2, 0.3, 0.0):
        mu = np.linspace(-1, 1, 20001)
        vals = np.array([probe.cso_henyey_greenstein(float(m), g) for m in mu[::50]])
        integral = 2 * math.pi * np.trapezoid(vals, mu[::50])
        assert abs(integral - 1.0) < 2e-3, (g, integral)   # normalised over the sphere
        c = 0.3
        closed = (1 - g * g) / (4 * math.pi * (1 + g * g - 2 * g * c) ** 1.5)
        assert abs(probe.cso_henyey_greenstein(c, g) - closed) < 1e-6
    assert probe.cso_henyey_greenstein(0.5, -1.0) == 0.0       # noon sun: g2 = 0.4 - 1.4 = -1


def test_remap_height_fraction_gradient(probe):
    assert probe.cso_remap(0.5, 0.0, 1.0, 0.0, 1.0) == 0.5
    assert probe.cso_remap(0.25, 0.0, 0.5, 10.0, 20.0) == 15.0
    assert probe.cso_height_fraction(6001500.0) == 0.0 and probe.cso_height_fraction(6004000.0) == 1.0
    assert probe.cso_height_fraction(6002750.0) == 0.5 and probe.cso_height_fraction(5.9e6) == 0.0 and probe.cso_height_fraction(7e6) == 1.0
    # cumulus (type 1): smoothstep(0.01, 0.0625, h) - smoothstep(0.78, 1, h) (clouds.glsl:83-94)
    sm = lambda a, b, x: (lambda t: t * t * (3 - 2 * t))(min(max((x - a) / (b - a), 0.0), 1.0))
    for h in (0.0, 0.03, 0.3, 0.9, 1.0):
        assert abs(probe.cso_density_height_gradient(h, 1.0) - (sm(0.01, 0.0625, h) - sm(0.78, 1.0, h))) < 1e-6
    for h in (0.03, 0.3, 0.5, 0.6):  # stratocumulus (type 0.5)
        assert abs(probe.cso_density_height_gradient(h, 0.5) - (sm(0.02, 0.2, h) - sm(0.48, 0.625, h))) < 1e-6
    assert probe.cso_density_height_gradient(0.0, 0.8) == 0.0 and abs(probe.cso_density_height_gradient(1.0, 0.8)) < 1e-6


def test_fp16_conversion_matches_ieee(probe):
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.normal(0, 1, 2000), rng.uniform(-70000, 70000, 500), 10.0 ** rng.uniform(-9, -3, 500),
                         [0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e-8, 5.96e-8, 2.98e-8, 2.9802322e-08, 6.1e-5, np.inf, -np.inf]]).astype(np.float32)
    for x in xs:
        with np.errstate(over="ignore"):
            want = np.float32(x).astype(np.float16).view(np.uint16)
        assert probe.cso_f32_to_f16(float(x)) == want, x
    for h in list(range(0, 0x7c01, 7)) + [0x8000, 0x8001, 0xfbff, 0x0001, 0x03ff, 0x0400]:
        assert probe.cso_f16_to_f32(h) == np.uint16(h).view(np.float16).astype(np.float32)


# ---- samplers ----------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx_small(oracle_lib, small_textures):
    c = oracle_lib.context(0)
    c.upload_textures(*small_textures)
    c.build_transmittance_lut()
    c.build_sky_lut((0.0, 1.0, 0.0))
    yield c
    c.close()


def test_volume_sampler(probe, ctx_small, small_textures):
    large, small, weather = small_textures
    n = large.shape[0]
    out = (C.c_float * 4)()
    rng = np.random.default_rng(4)
    for _ in range(50):  # texel centres return the texel, REPEAT wraps by whole periods
        x, y, z = rng.integers(0, n, 3)
        for shift in (0, 3, -2):
            s = vec((x + 0.5) / n + shift, (y + 0.5) / n - shift, (z + 0.5) / n)
            assert probe.cso_sample_volume(ctx_small._h, 0, s, 0.0, out) == 0
            np.testing.assert_allclose(list(out), large[z, y, x] / 255.0, atol=2e-6)
    # halfway between two texels along x, across the wrap seam: the average of texel n-1 and texel 0
    s = vec(0.0, 0.5 / n, 0.5 / n)
    probe.cso_sample_volume(ctx_small._h, 0, s, 0.0, out)
    np.testing.assert_allclose(list(out), (large[0, 0, n - 1].astype(np.float64) + large[0, 0, 0]) / 510.0, atol=2e-6)
    # negative LOD clamps to level 0 (clouds.glsl:117 "mip - 2.0"); the last level is 1x1x1
    a, b = (C.c_float * 4)(), (C.c_float * 4)()
    s = vec(0.37, 0.11, 0.93)
    probe.cso_sample_volume(ctx_small._h, 0, s, -2.0, a); probe.cso_sample_volume(ctx_small._h, 0, s, 0.0, b)
    assert list(a) == list(b)
    probe.cso_sample_volume(ctx_small._h, 1, s, 5.0, a); probe.cso_sample_volume(ctx_small._h, 1, vec(0.9, 0.2, 0.4), 9.0, b)
    assert list(a) == list(b)


def test_mip_chain_is_rounded_box_filter(ctx_small, small_textures):
    large = small_textures[0].astype(np.int64)
    n = large.shape[0]
    prev = large
    for level in range(1, 5):
        m = prev.shape[0] // 2
        box = prev.reshape(m, 2, m, 2, m, 2, 4).sum((1, 3, 5))
        want = (box + 4) >> 3
        got = ctx_small.read_volume_level(0, n, level)
        np.testing.assert_array_equal(got, want.astype(np.uint8))
        prev = want


def test_weather_and_lut_samplers(probe, ctx_small, small_textures):
    weather = small_textures[2]
    h, w, _ = weather.shape
    out3, out4 = (C.c_float * 3)(), (C.c_float * 4)()
    probe.cso_sample_weather(ctx_small._h, (5 + 0.5) / w + 2.0, (7 + 0.5) / h - 1.0, out3)
    np.testing.assert_allclose(list(out3), weather[7, 5] / 255.0, atol=2e-6)
    T = ctx_small.read_transmittance_lut().astype(np.float32)
    probe.cso_sample_lut(ctx_small._h, 0, (100 + 0.5) / 256, (10 + 0.5) / 64, out4)
    np.testing.assert_allclose(list(out4), T[10, 100], rtol=1e-6)
    probe.cso_sample_lut(ctx_small._h, 0, -3.0, 7.0, out4)   # CLAMP_TO_EDGE
    np.testing.assert_allclose(list(out4), T[63, 0], rtol=1e-6)
    probe.cso_sample_lut(ctx_small._h, 0, 1.0, 0.0, out4)
    np.testing.assert_allclose(list(out4), T[0, 255], rtol=1e-6)


# ---- LUTs --------------------------------------------------------------------------------------------
def test_transmittance_lut_properties(ctx_small):
    T = ctx_small.read_transmittance_lut().astype(np.float32)
    assert T.shape == (64, 256, 4) and (T >= 0).all() and (T <= 1).all()
    np.testing.assert_allclose(T[0, 255], [0.903, 0.867, 0.831, 0.750], atol=2e-3)        # SURVEY §4 spot values
    np.testing.assert_allclose(T[0, 128], [1.4e-2, 3.1e-3, 2.7e-4, 3.9e-6], rtol=0.03)
    assert (T[63, 255] == 1.0).all()
    assert (np.diff(T[:, 128:, :], axis=0) >= -2e-3).all()       # u >= 0.5: non-decreasing with altitude
    assert (T[:, 255] >= T[:, 128]).all() and (T[:32, 255] > T[:32, 128]).all()  # sun overhead > sun on the horizon
    assert T[0, 255, 0] > T[0, 255, 1] > T[0, 255, 2] > T[0, 255, 3]  # 630 > 560 > 490 > 430 nm at sea level


def test_sky_lut_properties(oracle_lib, ctx_small):
    K = ctx_small.read_sky_lut().astype(np.float32)
    assert K.shape == (100, 200, 4) and (K[..., :3] >= 0).all() and (K[..., 3] == 1).all()
    assert 0.1 < K[..., :3].min() < 0.25 and 15 < K[..., :3].max() < 20     # noon 0.17 .. 17.9
    # noon sun: no azimuth dependence
    assert np.abs(K[:, 1:, :3] - K[:, :1, :3]).max() < 2e-2 * K[..., :3].max()
    # low sun toward Godot -x (the demo's sunset direction): the ray toward the sun has cos_theta = dot(-ray, sun_dir) = -1
    sun = np.array([-0.998773, 0.0495291, 0.0])
    ctx_small.build_sky_lut(tuple(sun))
    K = ctx_small.read_sky_lut().astype(np.float32)
    lum = K[..., :3].sum(-1)
    y, x = np.unravel_index(lum.argmax(), lum.shape)
    az = 2 * math.pi * x / 200                 # sky-lut.glsl:286
    l = 2 * y / 100 - 1
    elev = l * l * np.sign(l) * math.pi / 2    # :290-291
    ray = np.array([math.cos(elev) * math.cos(az), math.cos(elev) * math.sin(az), math.sin(elev)])
    sun_dir = np.array([-sun[0], -sun[2], sun[1]])  # :221-223
    assert np.dot(-ray, sun_dir) < -0.97        # brightest texel sits at the HG forward peak (+2g cos convention, :122-126)
    assert abs(elev) < math.radians(8)          # near the horizon, where the low sun is
    # mirror symmetry in azimuth about the sun azimuth (az = 0 here): column x <-> column 200 - x
    np.testing.assert_allclose(K[:, 1:100, :3], K[:, 199:100:-1, :3], rtol=2e-2, atol=2e-3)
    ctx_small.build_sky_lut((0.0, 1.0, 0.0))


# ---- density and the march ---------------------------------------------------------------------------
def test_density_zero_outside_slab_and_nan_safe(cs, oracle_lib, probe, ctx_small, helpers):
    p = helpers.make_params(oracle_lib, 64, 32)
    w = vec(0.8, 0.0, 1.0)
    for r in (6001000.0, 6001500.0, 6004000.0, 6004500.0):
        assert probe.cso_density(ctx_small._h, C.byref(p), vec(0.0, r, 0.0), w, 0.0) == 0.0
    inside = [probe.cso_density(ctx_small._h, C.byref(p), vec(x, 6002500.0, 300.0 * k), vec(0.8, 0.0, 1.0), 0.0) for k, x in enumerate(np.linspace(0, 9000, 40))]
    assert all(0.0 <= v <= 1.0 for v in inside)
    q = helpers.make_params(oracle_lib, 64, 32, coverage=0.0)   # remap divides by zero (clouds.glsl:124): must not poison the image
    v = probe.cso_density(ctx_small._h, C.byref(q), vec(10.0, 6002500.0, 20.0), w, 0.0)
    assert v == 0.0 or math.isnan(v)


def test_march_image_properties(cs, oracle_lib, textures, helpers):
    W, H = 96, 48
    ctx = helpers.prepared_context(oracle_lib, textures, W, H, threads=helpers.cpu_threads)
    p = helpers.make_params(oracle_lib, W, H)
    ctx.render_frame(p)
    img = ctx.read_image().astype(np.float32)
    k = ctx.get_counters().as_dict()
    assert np.isfinite(img).all() and (img[..., 3] >= 0).all() and (img[..., 3] <= 1).all() and (img[..., :3] >= 0).all()
    assert k["primary_steps"] == k["marched_pixels"] * 128
    assert k["density_evals"] == k["primary_steps"] + 7 * k["lit_steps"]           # 6 cone + 1 distant (clouds.glsl:186-199)
    assert 0.08 < k["lit_steps"] / k["primary_steps"] < 0.18                      # ~0.12 at coverage 0.2
    assert W * H - (W + H) <= k["marched_pixels"] <= W * H                          # all but the row-0 / column-0 edge
    # coverage 0: every density is 0/NaN -> unlit -> an all-zero image (no NaN leaks out)
    ctx.render_frame(helpers.make_params(oracle_lib, W, H, coverage=0.0))
    z = ctx.read_image().astype(np.float32)
    assert (z == 0).all()
    # zero light energy: only the ambient terms remain, image stays finite and darker
    ctx.render_frame(helpers.make_params(oracle_lib, W, H, energy=0.0))
    dark = ctx.read_image().astype(np.float32)
    assert (dark[..., :3] <= img[..., :3] + 1e-3).all() and (dark[..., 3] == img[..., 3]).all()
    # tile invariance: 8x8-group dispatches over 4 tiles == one full-frame dispatch
    ctx.render_frame(p)
    full = ctx.read_image().copy()
    ctx.resize(W, H)
    for ty in range(2):
        for tx in range(2):
            q = p.copy(); q.update_position[0] = tx * W // 2; q.update_position[1] = ty * H // 2
            ctx.dispatch_clouds(q, W // 16, H // 16)
    assert (ctx.read_image().view(np.uint16) == full.view(np.uint16)).all()
    ctx.close()


def test_oracle_error_behaviour(cs, oracle_lib, small_textures, helpers):
    ctx = oracle_lib.context(0)
    with pytest.raises(cs.CloudSkyError) as e:
        ctx.build_sky_lut((0, 1, 0))
    assert e.value.code == 5
    ctx.build_transmittance_lut(); ctx.build_sky_lut((0, 1, 0)); ctx.resize(32, 16)
    with pytest.raises(cs.CloudSkyError):
        ctx.render_frame(helpers.make_params(oracle_lib, 32, 16))
    ctx.upload_textures(*small_textures)
    ctx.render_frame(helpers.make_params(oracle_lib, 32, 16))
    with pytest.raises(cs.CloudSkyError):
        ctx.render_frame(helpers.make_params(oracle_lib, 64, 16))
    with pytest.raises(cs.CloudSkyError):
        ctx.set_stream(1)
    ctx.close()

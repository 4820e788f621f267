"""Parity pinned to the REFERENCE ITSELF.

oracle/_ref/libcloudsky_ref.so is the reference's three GLSL compute shaders compiled unmodified by g++ behind
oracle/glsl_compat.h (recipe: oracle/build_ref.sh); tests/golden/ref_golden.npz holds its outputs
(tests/golden/make_ref_golden.py).  Chain of evidence:

  reference GLSL (compiled) == committed golden   (CPU, whenever _ref is present — build container and GPU box)
  hand-written oracle       == reference GLSL     (CPU, bit for bit: both LUTs every texel, clouds every pixel incl. the edges)
  CUDA kernels              ~  reference golden   (GPU, the tolerances of DESIGN.md §5, without running any CPU code)
"""
import hashlib
import os

import numpy as np
import pytest

import refbind

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
CASES = ["noon", "sunset", "sunset_wind", "oblique_wind", "overcast"]
W, H = 128, 64
C3_W, C3_H = 2048, 1024


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


@pytest.fixture(scope="module")
def ref(textures):
    if not refbind.build():
        pytest.skip("oracle/_ref not built and /root/reference not mounted")
    r = refbind.Reference()
    r.upload_textures(*textures)
    r.build_transmittance_lut()
    return r


def u16(a):
    return np.ascontiguousarray(a).view(np.uint16)


def half_ulps(a, b):
    return np.abs(u16(a).astype(np.int32) - u16(b).astype(np.int32))  # all values here are non-negative


def assert_same_up_to_libm(a, b, what):
    """Bit-identical on the machine that made the fixture; another CPU may pick a different expf/powf variant in glibc
    (FMA / non-FMA ifuncs), worth at most one fp16 step on a few values (cf. tests/test_oracle_golden.py)."""
    d = half_ulps(a, b)
    assert d.max() <= 2 and (d > 0).mean() < 0.02, (what, int(d.max()), float((d > 0).mean()))


def test_reference_sources_are_the_ones_the_fixture_was_made_from(gold):
    if not os.path.isdir(refbind.REFERENCE_SHADERS):
        pytest.skip("/root/reference not mounted (GPU box)")
    want = dict(line.split()[::-1] for line in str(gold["reference_sha256"]).strip().splitlines())
    for name in ("clouds.glsl", "sky-lut.glsl", "transmittance-lut.glsl"):
        with open(os.path.join(refbind.REFERENCE_SHADERS, name), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == want[name], name


def test_build_recipe_edits_nothing_but_what_it_states():
    """oracle/build_ref.sh may only drop `#[compute]` / `#version 450` and rewrite `out T x` parameters."""
    if not os.path.isdir(refbind.REFERENCE_SHADERS):
        pytest.skip("/root/reference not mounted (GPU box)")
    import re
    import subprocess
    script = open(os.path.join(ROOT, "oracle", "build_ref.sh")).read()
    sed = re.search(r"^\s*(sed -e .*) \"\$2\"$", script, re.M).group(1)
    for name in ("clouds.glsl", "sky-lut.glsl", "transmittance-lut.glsl"):
        path = os.path.join(refbind.REFERENCE_SHADERS, name)
        src = open(path).read().splitlines()
        got = subprocess.check_output(["bash", "-c", f"{sed} {path}"], text=True).splitlines()
        assert src[0] == "#[compute]" and src[1] == "#version 450"
        body = src[2:]
        assert len(got) == len(body)
        changed = [(a, b) for a, b in zip(body, got) if a != b]
        for a, b in changed:
            assert re.sub(r"\bout (vec[234]|float) ", r"\1& ", a) == b
        if name == "clouds.glsl":
            assert not changed  # the hot kernel's shader is compiled exactly as it is written


def test_compiled_reference_reproduces_committed_golden(cs, ref, gold):
    assert_same_up_to_libm(ref.tlut, gold["transmittance"], "transmittance")
    for name in CASES:
        p = cs.CloudParams.from_floats(gold[f"params_{name}"])
        sky = ref.build_sky_lut(tuple(p.light_direction), tlut=gold["transmittance"])
        assert_same_up_to_libm(sky, gold[f"sky_{name}"], f"sky_{name}")
        ref.write_sky_lut(gold[f"sky_{name}"])
        assert_same_up_to_libm(ref.render(p, W, H), gold[f"clouds_{name}"], f"clouds_{name}")


def test_oracle_is_bit_identical_to_reference_golden(cs, oracle_lib, textures, helpers, gold):
    """The hand-written restatement against the vectors the compiled reference produced (no _ref needed)."""
    ctx = helpers.prepared_context(oracle_lib, textures, W, H, threads=helpers.cpu_threads)
    assert_same_up_to_libm(ctx.read_transmittance_lut(), gold["transmittance"], "transmittance")
    ctx.write_transmittance_lut(gold["transmittance"])
    for name in CASES:
        p = cs.CloudParams.from_floats(gold[f"params_{name}"])
        ctx.build_sky_lut(tuple(p.light_direction))
        assert_same_up_to_libm(ctx.read_sky_lut(), gold[f"sky_{name}"], f"sky_{name}")
        ctx.write_sky_lut(gold[f"sky_{name}"])
        ctx.set_march_config(128, 6)
        ctx.render_frame(p)
        assert_same_up_to_libm(ctx.read_image(), gold[f"clouds_{name}"], f"clouds_{name}")
    # bench-sized frame (C3, 2048x1024), frames 0 and 15 of the wind animation, 16 rows
    ctx.resize(C3_W, C3_H)
    ctx.write_sky_lut(gold["c3_sky"])
    buf = np.zeros((C3_H, C3_W, 4), np.float16)
    for k in (0, 15):
        p = cs.CloudParams.from_floats(gold[f"c3_params{k}"])
        for r in gold["c3_rows"]:
            ctx.render_rows_to(p, int(r), int(r) + 1, buf.ctypes.data)
        assert_same_up_to_libm(buf[gold["c3_rows"]], gold[f"c3_frame{k}"], f"c3_frame{k}")
    ctx.close()


def test_oracle_is_bit_identical_to_live_reference_c1(cs, oracle_lib, ref, textures, helpers):
    """same machine, same libm -> EXACT equality, every texel of both LUTs and every pixel
    (row 0 / column 0 included) of the C1-sized 256x128 frame at the reference's step counts, 5 parameter sets."""
    from golden.make_ref_golden import CASES as KW
    w, h = 256, 128
    ctx = helpers.prepared_context(oracle_lib, textures, w, h, threads=helpers.cpu_threads)
    assert np.array_equal(u16(ctx.read_transmittance_lut()), u16(ref.tlut))
    for name, kw in KW.items():
        p = helpers.make_params(oracle_lib, w, h, **kw)
        sun = tuple(p.light_direction)
        ctx.build_sky_lut(sun)
        sky = ref.build_sky_lut(sun)
        assert np.array_equal(u16(ctx.read_sky_lut()), u16(sky)), name
        ctx.set_march_config(128, 6)
        ctx.render_frame(p)
        assert np.array_equal(u16(ctx.read_image()), u16(ref.render(p, w, h))), name
    ctx.close()


def test_reference_tile_addressing_matches_full_dispatch(cs, ref, oracle_lib, helpers, gold):
    """update_position tiles (cloud_sky.gd:156-161) of the compiled shader assemble to the single full dispatch."""
    p = cs.CloudParams.from_floats(gold["params_oblique_wind"])
    ref.write_sky_lut(gold["sky_oblique_wind"])
    full = ref.render(p, W, H)
    tiled = np.zeros_like(full)
    ref.render(p, W, H, rows=range(H), out=tiled)
    assert np.array_equal(u16(full), u16(tiled))


# ---------------------------------------------------------------------------------------------------------
# GPU: the CUDA kernels against the reference-compiled vectors (no CPU code involved)
# ---------------------------------------------------------------------------------------------------------
LUT_TOL = (1e-3, 2e-3)
STRICT_TOL = (1e-3, 2e-3, 0.999)
FAST_TOL = (2e-3, 1e-2, 0.999)


@pytest.mark.gpu
def test_gpu_lut_kernels_match_reference_golden(cs, product_lib, textures, helpers, gold):
    ctx = helpers.prepared_context(product_lib, textures, W, H)
    a, g = ctx.read_transmittance_lut().astype(np.float32), gold["transmittance"].astype(np.float32)
    assert (np.abs(a - g) <= LUT_TOL[0] + LUT_TOL[1] * np.abs(g)).all()
    assert (half_ulps(ctx.read_transmittance_lut(), gold["transmittance"]) == 0).mean() > 0.98
    for name in CASES + ["c3"]:
        sun = (0.0, 1.0, 0.0) if name == "c3" else tuple(cs.CloudParams.from_floats(gold[f"params_{name}"]).light_direction)
        ctx.write_transmittance_lut(gold["transmittance"])
        ctx.build_sky_lut(sun)
        key = "c3_sky" if name == "c3" else f"sky_{name}"
        a, g = ctx.read_sky_lut().astype(np.float32), gold[key].astype(np.float32)
        assert (np.abs(a - g) <= LUT_TOL[0] + LUT_TOL[1] * np.abs(g)).all(), name
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["strict", "fast", "tex", "half"])
def test_gpu_march_matches_reference_golden(cs, product_lib, textures, helpers, gold, mode):
    m, tol = {"strict": (cs.MODE_STRICT, STRICT_TOL), "fast": (cs.MODE_FAST, FAST_TOL), "tex": (cs.MODE_FAST | cs.MODE_TEX, FAST_TOL),
              "half": (cs.MODE_FAST | cs.MODE_HALF, FAST_TOL)}[mode]
    ctx = helpers.prepared_context(product_lib, textures, W, H)
    ctx.set_march_config(128, 6, m)
    for name in CASES:
        p = cs.CloudParams.from_floats(gold[f"params_{name}"])
        ctx.write_sky_lut(gold[f"sky_{name}"])
        ctx.render_frame(p)
        frac, mx = helpers.compare_images(ctx.read_image(), gold[f"clouds_{name}"], tol[0], tol[1])
        assert frac >= tol[2] and mx < 0.1, (mode, name, frac, mx)
        if mode == "fast":  # the product kernel also holds the strict kernel's tolerance against the compiled reference
            frac, mx = helpers.compare_images(ctx.read_image(), gold[f"clouds_{name}"], STRICT_TOL[0], STRICT_TOL[1])
            assert frac >= 0.9995 and mx < 5e-3, (mode, name, frac, mx)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fast", "tex", "half"])
def test_gpu_bench_frames_match_reference_golden_rows(cs, product_lib, textures, helpers, gold, mode):
    """The 2048x1024 frames bench.py renders (k = 0 and k = 15 of the wind animation), at the reference's 6+1 light
    samples, against 16 rows of the compiled reference."""
    ctx = helpers.prepared_context(product_lib, textures, C3_W, C3_H)
    ctx.set_march_config(128, 6, cs.MODE_FAST | {"fast": 0, "tex": cs.MODE_TEX, "half": cs.MODE_HALF}[mode])
    ctx.write_sky_lut(gold["c3_sky"])
    rows = gold["c3_rows"]
    for k in (0, 15):
        p = cs.CloudParams.from_floats(gold[f"c3_params{k}"])
        ctx.render_frame(p)
        img = ctx.read_image()[rows]
        frac, mx = helpers.compare_images(img, gold[f"c3_frame{k}"], FAST_TOL[0], FAST_TOL[1], skip_edges=False)
        # column 0 of every row is the dir.y > 0 coin flip: compare without it
        frac, mx = helpers.compare_images(img[:, 1:], gold[f"c3_frame{k}"][:, 1:], FAST_TOL[0], FAST_TOL[1], skip_edges=False)
        assert frac >= FAST_TOL[2] and mx < 0.1, (mode, k, frac, mx)
    ctx.close()


@pytest.mark.gpu
def test_gpu_march_matches_live_reference_when_present(cs, product_lib, textures, helpers):
    """On a box that carries oracle/_ref (it travels with the snapshot): a fresh parameter set that is in no fixture."""
    if not refbind.available():
        pytest.skip("oracle/_ref not present")
    r = refbind.Reference()
    r.upload_textures(*textures)
    r.build_transmittance_lut()
    w, h = 320, 160
    p = helpers.make_params(product_lib, w, h, sun=(-0.3, 0.35, 0.6), time=91.0, wind_direction=4.0, wind_speed=2.5, coverage=0.35)
    sky = r.build_sky_lut(tuple(p.light_direction))
    want = r.render(p, w, h)
    ctx = helpers.prepared_context(product_lib, textures, w, h)
    ctx.write_sky_lut(sky)
    for m, tol in ((cs.MODE_STRICT, STRICT_TOL), (cs.MODE_FAST, FAST_TOL), (cs.MODE_FAST | cs.MODE_TEX, FAST_TOL)):
        ctx.set_march_config(128, 6, m)
        ctx.render_frame(p)
        frac, mx = helpers.compare_images(ctx.read_image(), want, tol[0], tol[1])
        assert frac >= tol[2] and mx < 0.1, (m, frac, mx)
    ctx.close()


def test_oracle_equals_live_reference_on_random_parameter_sets(cs, oracle_lib, ref, textures, helpers):
    """Seeded fuzz over the whole push-constant surface (sun direction incl. below-horizon and near-horizon suns, energy, colour,
    ground colour, coverage, density, wind direction / speed / elapsed time, time_offset, non-square sizes): the hand-written
    oracle and the compiled reference shader must agree bit for bit on both the sky LUT and the cloud image."""
    rng = np.random.default_rng(20261017)
    ctx = helpers.prepared_context(oracle_lib, textures, 64, 32, threads=helpers.cpu_threads)
    ctx.set_march_config(128, 6)
    for it in range(12):
        w, h = [(64, 32), (48, 48), (40, 24), (96, 16)][it % 4]
        ctx.resize(w, h)
        el = rng.uniform(-0.2, 1.0) if it % 3 else rng.uniform(-0.02, 0.06)  # every third case: sun at the horizon
        az = rng.uniform(0, 2 * np.pi)
        sun = (np.cos(az) * np.sqrt(max(0.0, 1 - el * el)), el, np.sin(az) * np.sqrt(max(0.0, 1 - el * el)))
        p = helpers.make_params(oracle_lib, w, h, sun=sun, coverage=float(rng.uniform(0.05, 1.0)), density=float(rng.uniform(0.01, 0.2)),
                                time=float(rng.uniform(0.0, 500.0)), wind_direction=float(rng.uniform(0, 6.28)), wind_speed=float(rng.uniform(0, 8)),
                                energy=float(rng.uniform(0.2, 4.0)), color=tuple(rng.uniform(0.2, 1.0, 3)), time_offset=float(rng.uniform(0, 50)))
        p.ground_color[:] = [float(v) for v in rng.uniform(0, 1, 4)]
        sun_n = tuple(p.light_direction)
        ctx.build_sky_lut(sun_n)
        sky = ref.build_sky_lut(sun_n)
        assert np.array_equal(u16(ctx.read_sky_lut()), u16(sky)), (it, sun_n)
        ctx.render_frame(p)
        assert np.array_equal(u16(ctx.read_image()), u16(ref.render(p, w, h))), (it, sun_n)
    ctx.close()

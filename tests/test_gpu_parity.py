"""GPU parity tests proper: the CUDA path vs the CPU oracle, both called through the C-ABI.

Tolerances (stated here, justified in DESIGN.md §Parity):
  * LUTs (accurate kernels, --fmad=false):   |gpu - oracle| <= 1e-3 + 2e-3*|oracle|  on every texel
  * clouds STRICT (oracle operation order):  <= 1e-3 + 2e-3*|oracle| on >= 99.9 % of pixels
  * clouds FAST (FMA + MUFU intrinsics):     <= 2e-3 + 1e-2*|oracle| on >= 99.9 % of pixels
  * clouds FAST | TEX (texture-unit filter): the same tolerance and pixel fraction as FAST (the texture unit's 8-bit
    filter weights move a texel fetch by <= 1/512 of the local texel difference)
    (SURVEY 8(c)'s recommended criterion; an FMA-contracted build of the oracle itself only reaches
     99.91 % at this tolerance — the reference's fp32 arithmetic at 6e6 m is that ill-conditioned
     near the horizon; measured margins are in DESIGN.md section 5)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SUNS = [(0.0, 1.0, 0.0), (-0.998773, 0.0495291, 2.69869e-07), (0.5, 0.5, 0.70710678), (0.3, -0.2, 0.9)]


@pytest.fixture(scope="module")
def pair(cs, oracle_lib, product_lib, textures, helpers):
    W, H = 256, 128
    o = helpers.prepared_context(oracle_lib, textures, W, H, threads=helpers.cpu_threads)
    g = helpers.prepared_context(product_lib, textures, W, H)
    yield o, g, W, H
    o.close()
    g.close()


def test_transmittance_lut_matches_oracle(pair):
    o, g, _, _ = pair
    a = o.read_transmittance_lut().astype(np.float32)
    b = g.read_transmittance_lut().astype(np.float32)
    d = np.abs(a - b)
    assert (d <= 1e-3 + 2e-3 * np.abs(a)).all(), d.max()
    assert (a.view(np.uint32) == b.view(np.uint32)).mean() > 0.98  # nearly every texel is bit-identical


@pytest.mark.parametrize("sun", SUNS)
def test_sky_lut_matches_oracle(pair, sun):
    o, g, _, _ = pair
    n = np.linalg.norm(sun)
    sun = tuple(float(v / n) for v in sun)
    o.build_sky_lut(sun)
    g.build_sky_lut(sun)
    a = o.read_sky_lut().astype(np.float32)
    b = g.read_sky_lut().astype(np.float32)
    d = np.abs(a - b)
    assert (d <= 1e-3 + 2e-3 * np.abs(a)).all(), d.max()


CASES = [
    dict(sun=(0.0, 1.0, 0.0)),                                                    # BASELINE config noon sun
    dict(sun=(-0.998773, 0.0495291, 2.69869e-07)),                                # demo scene sunset (cloud-demo.tscn:21)
    dict(sun=(0.5, 0.5, 0.70710678), time=37.5, wind_direction=0.7, wind_speed=3.0),  # animated wind
    dict(sun=(0.0, 1.0, 0.0), coverage=1.0, density=0.1),                         # high coverage (config 5)
    dict(sun=(0.2, 0.9, -0.3), coverage=0.5, energy=2.0, color=(1.0, 0.8, 0.6), demo=False, time_offset=3.0),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mode", ["strict", "fast", "tex", "half"])
def test_clouds_match_oracle(cs, pair, helpers, oracle_lib, product_lib, case, mode):
    o, g, W, H = pair
    po = helpers.make_params(oracle_lib, W, H, **case)
    pg = helpers.make_params(product_lib, W, H, **case)
    assert bytes(po) == bytes(pg)  # both host-logic implementations produce the same push constants
    sun = tuple(po.light_direction)
    o.set_march_config(128, 6)
    o.build_sky_lut(sun)
    o.render_frame(po)
    ref = o.read_image()
    g.set_march_config(128, 6, {"strict": cs.MODE_STRICT, "fast": cs.MODE_FAST, "tex": cs.MODE_FAST | cs.MODE_TEX, "half": cs.MODE_FAST | cs.MODE_HALF}[mode])
    g.write_sky_lut(o.read_sky_lut())  # identical LUT texels on both sides isolates the march
    g.render_frame(pg)
    out = g.read_image()
    assert np.isfinite(out.astype(np.float32)).all()
    if mode == "strict":
        frac, mx = helpers.compare_images(out, ref, 1e-3, 2e-3)
        assert frac >= 0.999, (frac, mx)
        same = (out[1:, 1:].view(np.uint16) == ref[1:, 1:].view(np.uint16)).all(-1).mean()
        assert same >= 0.98, same  # the strict kernel is the oracle's arithmetic: measured 99.5-99.9 % bit-identical pixels
    else:
        frac, mx = helpers.compare_images(out, ref, 2e-3, 1e-2)
        assert frac >= 0.999, (frac, mx)
        if mode == "fast":
            # Since round 2 the fast kernel follows the reference's rounding trajectory (ray setup, light-sample offsets and the
            # quantised fp32 length() in the shader's own roundings): it holds the STRICT tolerance too, and nearly all pixels are
            # bit-identical to the oracle (measured 97.5-99.9 %, max abs error 4.9e-4).
            tight, mx_t = helpers.compare_images(out, ref, 1e-3, 2e-3)
            same = (out[1:, 1:].view(np.uint16) == ref[1:, 1:].view(np.uint16)).all(-1).mean()
            assert tight >= 0.9995 and mx_t < 5e-3 and same >= 0.93, (tight, mx_t, same)
    assert mx < 0.1


def test_all_cloud_types_and_small_volumes(cs, helpers, oracle_lib, product_lib):
    """Weather map whose cloud type spans 0..1 (stratus / stratocumulus / cumulus mixes: the general
    mixGradients path, clouds.glsl:82-90) on small non-reference volume sizes (16^3 / 8^3 / 32^2)."""
    from cloudsky_b200 import assets
    large, small, weather = assets.synthetic_textures(seed=11, large_n=16, small_n=8, weather_n=32)
    rng = np.random.default_rng(5)
    weather = weather.copy()
    weather[..., 0] = rng.integers(0, 256, weather.shape[:2], dtype=np.uint8)   # type on both sides of 0.5
    weather[..., 2] = rng.integers(60, 256, weather.shape[:2], dtype=np.uint8)
    W, H = 128, 64
    o = helpers.prepared_context(oracle_lib, (large, small, weather), W, H, threads=helpers.cpu_threads)
    g = helpers.prepared_context(product_lib, (large, small, weather), W, H)
    for lvl in range(5):  # the library's mip chain is the oracle's, bit for bit
        assert (g.read_volume_level(0, 16, lvl) == o.read_volume_level(0, 16, lvl)).all()
    for lvl in range(4):
        assert (g.read_volume_level(1, 8, lvl) == o.read_volume_level(1, 8, lvl)).all()
    g.write_sky_lut(o.read_sky_lut())
    for kw in (dict(coverage=0.6), dict(coverage=1.0, time=5.0, wind_direction=2.0), dict(coverage=0.0)):
        p = helpers.make_params(product_lib, W, H, **kw)
        o.render_frame(p)
        ref = o.read_image()
        for mode, atol, rtol, need in ((cs.MODE_STRICT, 1e-3, 2e-3, 0.999), (cs.MODE_FAST, 2e-3, 1e-2, 0.999),
                                       (cs.MODE_FAST | cs.MODE_TEX, 2e-3, 1e-2, 0.999)):
            g.set_march_config(128, 6, mode)
            g.render_frame(p)
            out = g.read_image()
            assert np.isfinite(out.astype(np.float32)).all()
            frac, mx = helpers.compare_images(out, ref, atol, rtol)
            assert frac >= need, (kw, mode, frac, mx)
    assert (ref == 0).all()  # coverage 0 -> empty image (remap by zero must not leak NaN, clouds.glsl:124)
    o.close(); g.close()


def test_record_formats(cs, helpers, oracle_lib, product_lib, textures, monkeypatch):
    """The fast kernel reads exact-integer fp16 records when every interpolation coefficient is representable
    (the reference textures are) and fp32 records otherwise; both must match the oracle, and rough random
    texels (coefficients beyond fp16's exact range) must take the fp32 path automatically."""
    W, H = 128, 64
    o = helpers.prepared_context(oracle_lib, textures, W, H, threads=helpers.cpu_threads)
    p = helpers.make_params(product_lib, W, H, time=4.0)
    o.render_frame(p)
    ref = o.read_image()
    imgs = []
    for force in (False, True):
        if force:
            monkeypatch.setenv("CLOUDSKY_FP32_RECORDS", "1")
        g = helpers.prepared_context(product_lib, textures, W, H)
        g.write_sky_lut(o.read_sky_lut())
        g.set_march_config(128, 6, cs.MODE_FAST)
        g.render_frame(p)
        imgs.append(g.read_image())
        frac, mx = helpers.compare_images(imgs[-1], ref, 2e-3, 1e-2)
        assert frac >= 0.999, (force, frac, mx)
        g.close()
    monkeypatch.delenv("CLOUDSKY_FP32_RECORDS")
    frac, mx = helpers.compare_images(imgs[0], imgs[1], 1e-3, 2e-3)
    assert frac >= 0.995, (frac, mx)  # same maths, different rounding of the coefficients
    # rough random texels: second differences of K = 5G+2B+A exceed 2048 and are odd -> not exact in fp16 -> fp32 records
    rng = np.random.default_rng(9)
    large = rng.integers(0, 256, (16, 16, 16, 4), dtype=np.uint8)
    small = rng.integers(0, 256, (8, 8, 8, 3), dtype=np.uint8)
    weather = rng.integers(100, 256, (32, 32, 3), dtype=np.uint8)
    o2 = helpers.prepared_context(oracle_lib, (large, small, weather), W, H, threads=helpers.cpu_threads)
    g2 = helpers.prepared_context(product_lib, (large, small, weather), W, H)
    g2.write_sky_lut(o2.read_sky_lut())
    q = helpers.make_params(product_lib, W, H, coverage=0.7)
    o2.render_frame(q); g2.set_march_config(128, 6, cs.MODE_FAST); g2.render_frame(q)
    frac, mx = helpers.compare_images(g2.read_image(), o2.read_image(), 2e-3, 1e-2)
    assert frac >= 0.998, (frac, mx)
    o.close(); o2.close(); g2.close()


def test_odd_sizes_and_partial_tiles(cs, helpers, oracle_lib, product_lib, textures):
    """Image sizes that are not multiples of the 8x8 group / 16x8 CTA tile, and dispatches from unaligned origins."""
    for (W, H) in ((100, 37), (33, 129), (7, 5)):
        o = helpers.prepared_context(oracle_lib, textures, W, H, threads=helpers.cpu_threads)
        g = helpers.prepared_context(product_lib, textures, W, H)
        g.write_sky_lut(o.read_sky_lut())
        p = helpers.make_params(product_lib, W, H, coverage=0.5)
        o.render_frame(p)
        ref = o.read_image()
        for mode, atol, rtol, need in ((cs.MODE_STRICT, 1e-3, 2e-3, 0.995), (cs.MODE_FAST, 2e-3, 1e-2, 0.99)):
            g.set_march_config(128, 6, mode)
            g.render_frame(p)
            full = g.read_image()
            frac, mx = helpers.compare_images(full, ref, atol, rtol)
            assert frac >= need, (W, H, mode, frac, mx)
            if W < 32 or H < 32:
                continue
            # the same image from four dispatches with origins that are not multiples of 8
            g.resize(W, H)
            for (x0, y0) in ((0, 0), (13, 0), (0, 11), (13, 11)):
                q = p.copy(); q.update_position[0] = x0; q.update_position[1] = y0
                g.dispatch_clouds(q, 2 if x0 == 0 else (W - 13 + 7) // 8, 2 if y0 == 0 else (H - 11 + 7) // 8)
            q = p.copy(); q.update_position[0] = 0; q.update_position[1] = 0
            tiled = g.read_image()
            # pixels covered by the four rectangles ([0,16)x[0,16) U [13,W)x[0,16) U ...) are all pixels with x>=0,y>=0 except none: compare where rendered
            cover = np.zeros((H, W), bool)
            cover[:16, :16] = True; cover[:16, 13:] = True; cover[11:, :16] = True; cover[11:, 13:] = True
            if mode == cs.MODE_STRICT:  # one thread per pixel: bit-identical whatever the dispatch origin
                assert (tiled.view(np.uint16)[cover] == full.view(np.uint16)[cover]).all(), (W, H)
            else:  # warp composition differs with the origin only through rounding-free paths: still bit-identical by construction
                assert (tiled.view(np.uint16)[cover] == full.view(np.uint16)[cover]).all(), (W, H)
        o.close(); g.close()


def test_early_out_mode(cs, pair, helpers, oracle_lib, product_lib):
    """CS_MODE_FAST | CS_MODE_EARLY_OUT (opt-in, not reference behaviour): rays stop once T < 2^-12.  alpha is unchanged
    (it already rounds to 1.0 in fp16), RGB moves by at most a couple of fp16 ulps, the oracle tolerance still holds,
    and overcast skies execute far fewer primary steps."""
    o, g, W, H = pair
    for cov, min_saving in ((1.0, 0.2), (0.2, 0.0)):
        p = helpers.make_params(product_lib, W, H, coverage=cov, density=0.1 if cov == 1.0 else None)
        o.set_march_config(128, 6); o.build_sky_lut(tuple(p.light_direction)); o.render_frame(p)
        ref = o.read_image()
        g.write_sky_lut(o.read_sky_lut())
        g.set_counters_enabled(True)
        g.set_march_config(128, 6, cs.MODE_FAST); g.render_frame(p)
        full, k_full = g.read_image(), g.get_counters().as_dict()
        g.set_march_config(128, 6, cs.MODE_FAST | cs.MODE_EARLY_OUT); g.render_frame(p)
        early, k_early = g.read_image(), g.get_counters().as_dict()
        g.set_counters_enabled(False)
        assert (early[..., 3].view(np.uint16) == full[..., 3].view(np.uint16)).all()              # alpha bit-identical
        d = np.abs(early[..., :3].view(np.int16).astype(np.int32) - full[..., :3].view(np.int16).astype(np.int32))
        assert d.max() <= 2, d.max()                                                              # <= 2 fp16 ulps in RGB
        frac, mx = helpers.compare_images(early, ref, 2e-3, 1e-2)
        assert frac >= 0.999, (cov, frac, mx)
        assert k_early["primary_steps"] <= k_full["primary_steps"] * (1.0 - min_saving), (cov, k_early["primary_steps"], k_full["primary_steps"])
        assert k_full["primary_steps"] == k_full["marched_pixels"] * 128                          # nominal steps without the flag
    g.set_march_config(128, 6, cs.MODE_FAST)


def test_counters_match_oracle(cs, pair, helpers, oracle_lib, product_lib):
    o, g, W, H = pair
    p = helpers.make_params(product_lib, W, H)
    o.set_march_config(128, 6)
    o.build_sky_lut((0, 1, 0)); g.build_sky_lut((0, 1, 0))
    o.render_frame(p)
    ko = o.get_counters().as_dict()
    g.set_march_config(128, 6, cs.MODE_STRICT)
    g.set_counters_enabled(True)
    g.render_frame(p)
    kg = g.get_counters().as_dict()
    g.set_counters_enabled(False)
    assert kg["marched_pixels"] == ko["marched_pixels"]
    assert kg["primary_steps"] == ko["primary_steps"]
    assert abs(kg["lit_steps"] - ko["lit_steps"]) <= 1e-3 * ko["lit_steps"]
    assert abs(kg["density_evals"] - ko["density_evals"]) <= 1e-3 * ko["density_evals"]


@pytest.mark.parametrize("mode", ["strict", "fast", "tex", "half"])
def test_tile_invariance(cs, pair, helpers, product_lib, mode):
    """1, 4 and 64 tiles give bit-identical textures (cloud_sky.gd:156-161 tile walk)."""
    _, g, W, H = pair
    g.set_march_config(128, 6, {"strict": cs.MODE_STRICT, "fast": cs.MODE_FAST, "tex": cs.MODE_FAST | cs.MODE_TEX, "half": cs.MODE_FAST | cs.MODE_HALF}[mode])
    g.build_sky_lut((0, 1, 0))
    p = helpers.make_params(product_lib, W, H, time=12.0)
    g.render_frame(p)
    full = g.read_image().copy()
    for tiles in (2, 8):
        g.resize(W, H)  # clears the image
        tw, th = W // tiles, H // tiles
        for ty in range(tiles):
            for tx in range(tiles):
                q = p.copy()
                q.update_position[0] = tx * tw
                q.update_position[1] = ty * th
                g.dispatch_clouds(q, (tw + 7) // 8, (th + 7) // 8)
        tiled = g.read_image()
        assert (tiled.view(np.uint16) == full.view(np.uint16)).all()
    g.set_march_config(128, 6, cs.MODE_FAST)


def test_texture_unit_mode(cs, pair, helpers, oracle_lib, product_lib):
    """CS_MODE_FAST | CS_MODE_TEX (opt-in): the texture unit filters the noise volumes and the weather map, as the
    reference's sampler bindings do (cloud_sky.gd:389-398), with 8-bit fixed-point filter weights instead of the oracle's
    fp32 ones.  It stays close to the in-kernel-filter FAST image, composes with the early-out flag, handles non-reference
    march shapes (sequential fallback beyond 16 light samples) and is rejected together with STRICT."""
    o, g, W, H = pair
    p = helpers.make_params(product_lib, W, H, time=3.0)
    o.set_march_config(128, 6); o.build_sky_lut(tuple(p.light_direction)); o.render_frame(p)
    ref = o.read_image()
    g.write_sky_lut(o.read_sky_lut())
    g.set_march_config(128, 6, cs.MODE_FAST); g.render_frame(p)
    fast = g.read_image()
    g.set_march_config(128, 6, cs.MODE_FAST | cs.MODE_TEX); g.render_frame(p)
    tex = g.read_image()
    assert not (tex.view(np.uint16) == fast.view(np.uint16)).all()      # it really is a different sampler
    frac, mx = helpers.compare_images(tex, fast, 2e-3, 1e-2)
    assert frac >= 0.9995 and mx < 0.02, (frac, mx)
    g.set_march_config(128, 6, cs.MODE_FAST | cs.MODE_TEX | cs.MODE_EARLY_OUT); g.render_frame(p)
    early = g.read_image()
    assert (early[..., 3].view(np.uint16) == tex[..., 3].view(np.uint16)).all()
    d = np.abs(early[..., :3].view(np.int16).astype(np.int32) - tex[..., :3].view(np.int16).astype(np.int32))
    assert d.max() <= 2, d.max()
    for P, cone in ((64, 3), (128, 20)):  # 20 cone samples: more than the cooperative tables hold -> sequential light walk
        o.set_march_config(P, cone); o.render_frame(p)
        g.set_march_config(P, cone, cs.MODE_FAST | cs.MODE_TEX); g.render_frame(p)
        frac, mx = helpers.compare_images(g.read_image(), o.read_image(), 2e-3, 1e-2)
        assert frac >= 0.998, (P, cone, frac, mx)
    with pytest.raises(cs.CloudSkyError):
        g.set_march_config(128, 6, cs.MODE_STRICT | cs.MODE_TEX)
    o.set_march_config(128, 6)
    g.set_march_config(128, 6, cs.MODE_FAST)


def test_generalised_step_counts(cs, pair, helpers, oracle_lib, product_lib):
    """Extension rule of SURVEY §8(d): P primary steps, Lc cone samples (j%6 vectors, LOD clamp)."""
    o, g, W, H = pair
    p = helpers.make_params(product_lib, W, H)
    o.build_sky_lut((0, 1, 0)); g.write_sky_lut(o.read_sky_lut())
    for P, Lc in ((32, 3), (64, 5), (128, 7), (48, 11), (16, 20), (24, 0)):  # 20 cone samples: beyond the cooperative tables; 0: distant sample only
        o.set_march_config(P, Lc)
        o.render_frame(p)
        ref = o.read_image()
        for mode, tol in ((cs.MODE_STRICT, (1e-3, 2e-3, 0.999)), (cs.MODE_FAST, (2e-3, 1e-2, 0.998))):
            g.set_march_config(P, Lc, mode)
            g.render_frame(p)
            frac, mx = helpers.compare_images(g.read_image(), ref, tol[0], tol[1])
            # 16-24 primary steps: one lit/unlit flip of a single sample is 1/16 of the pixel, so the fast kernel's
            # rounding noise crosses the tolerance on more pixels (measured 99.45 %, independent of the cone count;
            # the strict kernel stays at 100 %)
            need = tol[2] if (P >= 32 or mode == cs.MODE_STRICT) else 0.99
            assert frac >= need, (P, Lc, mode, frac, mx)
    o.set_march_config(128, 6)
    g.set_march_config(128, 6, cs.MODE_FAST)


def test_full_size_properties_and_row_subsample(cs, helpers, oracle_lib, product_lib, textures):
    """BASELINE config 3 size (2048x1024): size-independent properties + oracle parity on 6 rows."""
    W, H = 2048, 1024
    g = helpers.prepared_context(product_lib, textures, W, H)
    o = helpers.prepared_context(oracle_lib, textures, W, H, threads=helpers.cpu_threads)
    p = helpers.make_params(product_lib, W, H, time=1.0)
    g.write_sky_lut(o.read_sky_lut())
    g.set_march_config(128, 6, cs.MODE_FAST)
    g.render_frame(p)
    img = g.read_image().astype(np.float32)
    assert np.isfinite(img).all()
    assert (img[..., 3] >= 0).all() and (img[..., 3] <= 1).all()
    assert (img[..., :3] >= 0).all()
    assert img[..., 3].mean() > 0.05  # there are clouds
    rows = [1, 100, 333, 512, 800, 1023]
    buf = np.zeros((H, W, 4), np.float16)
    for r in rows:
        o.render_rows_to(p, r, r + 1, buf.ctypes.data)
    ok_total, n_total = 0, 0
    for r in rows:
        d = np.abs(img[r, 1:] - buf[r, 1:].astype(np.float32))
        ok = (d <= 2e-3 + 1e-2 * np.abs(buf[r, 1:].astype(np.float32))).all(-1)
        ok_total += ok.sum(); n_total += ok.size
    assert ok_total / n_total >= 0.999, ok_total / n_total
    # tile/row-band invariance at full size: two half-frames into caller-owned device memory, against the single dispatch
    import torch
    halves = torch.zeros((H, W, 4), dtype=torch.float16, device="cuda")
    g.render_rows_to(p, 0, H // 2, halves.data_ptr())
    g.render_rows_to(p, H // 2, H, halves.data_ptr())
    g.sync()
    assert (halves.cpu().numpy().astype(np.float32) == img).all()
    g.close(); o.close()


def test_config5_shape_rows(cs, helpers, oracle_lib, product_lib, textures):
    """BASELINE config 5 shape: 8192x4096, 256 primary / 12 light (11 cone + 1 distant) steps, coverage 1.0.
    (Its adaptive stepping is not in the reference; the fixed-step render is compared on three rows.)"""
    W, H = 8192, 4096
    g = helpers.prepared_context(product_lib, textures, W, H)
    o = helpers.prepared_context(oracle_lib, textures, W, H, threads=helpers.cpu_threads)
    p = helpers.make_params(product_lib, W, H, coverage=1.0, time=2.0)
    g.write_sky_lut(o.read_sky_lut())
    g.set_march_config(256, 11, cs.MODE_FAST)
    o.set_march_config(256, 11)
    g.render_frame(p)
    img = g.read_image()
    f = img.astype(np.float32)
    assert np.isfinite(f).all() and (f[..., 3] >= 0).all() and (f[..., 3] <= 1).all() and f[..., 3].mean() > 0.3
    rows = [7, 2048, 3900]
    buf = np.zeros((H, W, 4), np.float16)
    ok = n = 0
    for r in rows:
        o.render_rows_to(p, r, r + 1, buf.ctypes.data)
        d = np.abs(f[r, 1:] - buf[r, 1:].astype(np.float32))
        good = (d <= 2e-3 + 1e-2 * np.abs(buf[r, 1:].astype(np.float32))).all(-1)
        ok += good.sum(); n += good.size
    assert ok / n >= 0.999, ok / n
    g.close(); o.close()


def test_render_frame_host_and_sun_batch(cs, pair, helpers, product_lib):
    _, g, W, H = pair
    g.set_march_config(128, 6, cs.MODE_FAST)
    p = helpers.make_params(product_lib, W, H, sun=(0.3, 0.6, 0.2))
    host = g.render_frame_host(p)
    g.build_sky_lut(tuple(p.light_direction))
    g.render_frame(p)
    assert (g.read_image().view(np.uint16) == host.view(np.uint16)).all()
    import torch
    # streaming readback: three frames in flight through two slots give the same bits as the synchronous call
    pinned = [torch.empty((H, W, 4), dtype=torch.float16).pin_memory() for _ in range(3)]
    ps = [helpers.make_params(product_lib, W, H, sun=(0.3, 0.6, 0.2), time=float(k)) for k in range(3)]
    for k in range(3):
        g.render_frame_host_async(ps[k], pinned[k].data_ptr())
    g.wait_host()
    for k in range(3):
        assert (g.render_frame_host(ps[k]).view(np.uint16) == pinned[k].numpy().view(np.uint16)).all()
    suns = np.array([[0.0, 1.0, 0.0], [0.6, 0.8, 0.0], [-0.998773, 0.0495291, 0.0]], np.float32)
    out = torch.zeros((3, H, W, 4), dtype=torch.float16, device="cuda")
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    g.render_sun_batch_to(p, suns, out.data_ptr())
    g.sync()
    batch = out.cpu().numpy()
    g.set_stream(0)
    for i in range(3):
        q = p.copy()
        q.light_direction[:] = suns[i].tolist()
        single = g.render_frame_host(q)
        assert (single.view(np.uint16) == batch[i].view(np.uint16)).all()


@pytest.mark.parametrize("which", ["reference_textures", "fp32_records"])
def test_sun_batch_kernel_matches_single_launches(cs, product_lib, textures, small_textures, helpers, which, monkeypatch):
    """cs_render_sun_batch_to marches up to 4 suns per launch (the primary loop is sun-independent); every image must be the
    one a single-sun dispatch produces, bit for bit — chunks of 4 + 3, both lit-lane paths (coverage 0.2 and 1.0), an image
    whose size is not a multiple of the CTA tile, and both record formats."""
    import torch
    tex = textures if which == "reference_textures" else small_textures
    W, H = (200, 100) if which == "reference_textures" else (72, 40)
    th = np.linspace(0.15, 2.9, 7)
    suns = np.stack([np.cos(th), np.sin(th), 0.2 * np.cos(3 * th)], 1)
    suns = (suns / np.linalg.norm(suns, axis=1, keepdims=True)).astype(np.float32)
    for cov in (0.2, 1.0):
        p = helpers.make_params(product_lib, W, H, coverage=cov, time=4.0, wind_direction=0.4)
        g = helpers.prepared_context(product_lib, tex, W, H)
        g.set_march_config(64, 6, cs.MODE_FAST)
        out = torch.zeros((7, H, W, 4), dtype=torch.float16, device="cuda")
        g.set_stream(torch.cuda.current_stream().cuda_stream)
        g.render_sun_batch_to(p, suns, out.data_ptr())
        g.sync()
        batch = out.cpu().numpy()
        g.set_stream(0)
        assert np.isfinite(batch.astype(np.float32)).all()
        for i in range(7):
            q = p.copy()
            q.light_direction[:] = suns[i].tolist()
            single = g.render_frame_host(q)
            assert (single.view(np.uint16) == batch[i].view(np.uint16)).all(), (which, cov, i)
        assert (batch[0].view(np.uint16) != batch[3].view(np.uint16)).any()
        g.close()
    # the per-sun fallback (CLOUDSKY_SUN_BATCH=0, read at cs_create) gives the same bits
    monkeypatch.setenv("CLOUDSKY_SUN_BATCH", "0")
    g = helpers.prepared_context(product_lib, tex, W, H)
    g.set_march_config(64, 6, cs.MODE_FAST)
    out2 = torch.zeros((7, H, W, 4), dtype=torch.float16, device="cuda")
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    g.render_sun_batch_to(p, suns, out2.data_ptr())
    g.sync()
    g.set_stream(0)
    assert (out2.cpu().numpy().view(np.uint16) == batch.view(np.uint16)).all()
    g.close()


def test_error_behaviour(cs, product_lib, small_textures, helpers):
    ctx = product_lib.context(0)
    p = helpers.make_params(product_lib, 64, 32)
    with pytest.raises(cs.CloudSkyError) as e:
        ctx.build_sky_lut((0, 1, 0))  # sky_lut.gd:45-47 "Attempting to update uninitialized sky lut"
    assert e.value.code == 5
    ctx.build_transmittance_lut()
    ctx.build_sky_lut((0, 1, 0))
    ctx.resize(64, 32)
    with pytest.raises(cs.CloudSkyError) as e:
        ctx.render_frame(p)  # no textures: can_run == false
    assert e.value.code == 5
    ctx.upload_textures(*small_textures)
    ctx.render_frame(p)
    bad = p.copy(); bad.texture_size[0] = 128
    with pytest.raises(cs.CloudSkyError):
        ctx.render_frame(bad)
    # dispatch that overhangs the image is clipped, not a fault (the reference does not bounds-check)
    q = p.copy(); q.update_position[0] = 56; q.update_position[1] = 24
    ctx.dispatch_clouds(q, 4, 4)
    ctx.sync()
    ctx.close()


def test_random_parameter_sets_all_sampler_modes(cs, pair, helpers, oracle_lib, product_lib):
    """Seeded fuzz over the push-constant surface (the same generator family as tests/test_reference_pin.py's CPU fuzz, where the
    oracle is shown bit-identical to the compiled reference): every sampler mode stays inside its gate on every case."""
    o, g, W, H = pair
    rng = np.random.default_rng(777)
    worst = {}
    for it in range(6):
        el = rng.uniform(0.05, 1.0) if it % 3 else rng.uniform(-0.02, 0.06)
        az = rng.uniform(0, 2 * np.pi)
        c = np.sqrt(max(0.0, 1 - el * el))
        kw = dict(sun=(np.cos(az) * c, el, np.sin(az) * c), coverage=float(rng.uniform(0.1, 1.0)), density=float(rng.uniform(0.02, 0.15)),
                  time=float(rng.uniform(0.0, 300.0)), wind_direction=float(rng.uniform(0, 6.28)), wind_speed=float(rng.uniform(0, 6)),
                  energy=float(rng.uniform(0.3, 3.0)), color=tuple(float(v) for v in rng.uniform(0.3, 1.0, 3)))
        po = helpers.make_params(oracle_lib, W, H, **kw)
        pg = helpers.make_params(product_lib, W, H, **kw)
        assert bytes(po) == bytes(pg)
        o.set_march_config(128, 6)
        o.build_sky_lut(tuple(po.light_direction))
        o.render_frame(po)
        ref = o.read_image()
        g.write_sky_lut(o.read_sky_lut())
        for name, mode, tol in (("strict", cs.MODE_STRICT, (1e-3, 2e-3)), ("fast", cs.MODE_FAST, (2e-3, 1e-2)), ("tex", cs.MODE_FAST | cs.MODE_TEX, (2e-3, 1e-2)),
                                ("half", cs.MODE_FAST | cs.MODE_HALF, (2e-3, 1e-2))):
            g.set_march_config(128, 6, mode)
            g.render_frame(pg)
            frac, mx = helpers.compare_images(g.read_image(), ref, tol[0], tol[1])
            worst[name] = min(worst.get(name, 1.0), frac)
            assert frac >= 0.999 and mx < 0.1, (it, name, frac, mx, kw)
    print("worst pass fractions:", worst)
    g.set_march_config(128, 6, cs.MODE_FAST)

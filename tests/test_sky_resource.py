"""The Sky resource (cs_sky_*): cloud_sky.gd's time-sliced update + temporal-blend state machine inside the
library.  CPU: the oracle's restatement against the independent Python mirror (sky.py), tile by
tile.  GPU: the product against the oracle, and tiles against one full dispatch."""
import numpy as np
import pytest

SUN_BASIS = (0.0, 0.6, 0.8, 1.0, 0.0, 0.0, 0.3, 0.8, 0.52)  # columns x, y, z (z is not unit: the library normalises it)


def make_settings(lib, size=64, frames=4, coverage=0.6):
    s = lib.settings_demo()
    s.texture_size, s.frames_to_update, s.cloud_coverage, s.wind_speed = size, frames, coverage, 4.0
    return s


def drive(sky, times):
    states = []
    for t in times:
        sky.update(t)
        f = sky.frame()
        states.append((f.frame, tuple(f.update_position), f.texture_to_update, f.texture_to_blend_from, f.texture_to_blend_to,
                       round(f.blend_amount, 6), f.sky_current_texture, f.sky_blend_from, f.sky_blend_to, f.sky_updates, bytes(f.frame_data)))
    return states


def test_oracle_sky_resource_matches_python_mirror(cs, oracle_lib, small_textures):
    times = [2.0 + 0.5 * k for k in range(11)]
    ctx = oracle_lib.context(0); ctx.set_threads(4)
    ctx.upload_textures(*small_textures); ctx.build_transmittance_lut(); ctx.set_march_config(16, 2)
    sky = cs.Sky(ctx, make_settings(oracle_lib))
    sky.set_sun(SUN_BASIS, 1.3, (1.0, 0.9, 0.7))
    states = drive(sky, times)
    f = sky.frame()
    assert (f.texture_size, f.update_region_size, f.num_workgroups, f.frames_to_update) == (64, 32, 4, 4)

    m = cs.CloudSky(oracle_lib)  # the independent restatement of the same GDScript
    m.ctx.set_threads(4); m.ctx.set_march_config(16, 2)
    m._s.texture_size = 64; m._s.cloud_coverage = 0.6; m._s.wind_speed = 4.0
    m.frames_to_update = 4
    m.load_textures(*small_textures)
    m.sun = cs.DirectionalLight(basis=SUN_BASIS, light_energy=1.3, light_color=(1.0, 0.9, 0.7))
    for t, st in zip(times, states):
        m.update_sky(t)
        assert (m.frame, tuple(m.update_position), m.texture_to_update, m.texture_to_blend_from, m.texture_to_blend_to,
                round(m.blend_amount, 6)) == st[:6]
        assert (m.sky_lut.current_texture, m.sky_lut.updates) == (st[6], st[9])
        assert bytes(m.frame_data) == st[10]
    # first call = initialize_sky: 2*4 tiles + 1, three LUT renders on the first update, then one per texture
    assert states[0][0] == 1 and states[0][9] == 3 + 2
    assert [s[5] for s in states[:5]] == [0.0, 0.25, 0.5, 0.75, 0.0]
    # completed textures are identical to the mirror's snapshots
    for i in range(3):
        if m.textures[i] is not None and i != m.texture_to_update:
            assert (sky.read_texture(i).view(np.uint16) == m.textures[i].view(np.uint16)).all()
    # property setter semantics: changing frames_to_update re-initialises (cloud_sky.gd:37-42)
    s2 = make_settings(oracle_lib, size=64, frames=16)
    sky.set_settings(s2)
    f = sky.frame()
    assert (f.frame, tuple(f.update_position), f.update_region_size, f.num_workgroups) == (0, (0, 0), 16, 2)
    sky.update(20.0)
    assert sky.frame().frame == 1 and sky.frame().sky_updates > states[-1][9]
    sky.close(); ctx.close(); m.close()


def test_sky_resource_errors(cs, oracle_lib, small_textures):
    ctx = oracle_lib.context(0)
    with pytest.raises(cs.CloudSkyError) as e:
        cs.Sky(ctx, make_settings(oracle_lib))
    assert e.value.code == 5
    ctx.upload_textures(*small_textures); ctx.build_transmittance_lut()
    bad = make_settings(oracle_lib); bad.frames_to_update = 0
    with pytest.raises(cs.CloudSkyError):
        cs.Sky(ctx, bad)
    ctx.close()


@pytest.mark.gpu
def test_gpu_sky_resource_matches_oracle(cs, oracle_lib, product_lib, small_textures, helpers):
    times = [1.0 + 0.7 * k for k in range(10)]
    out = {}
    for name, lib in (("oracle", oracle_lib), ("gpu", product_lib)):
        ctx = lib.context(0)
        if name == "oracle":
            ctx.set_threads(helpers.cpu_threads)
        ctx.upload_textures(*small_textures); ctx.build_transmittance_lut(); ctx.set_march_config(64, 6, cs.MODE_FAST)
        sky = cs.Sky(ctx, make_settings(lib, size=96, frames=4))
        sky.set_sun(SUN_BASIS, 1.3, (1.0, 0.9, 0.7))
        st = drive(sky, times)
        f = sky.frame()
        out[name] = (st, [sky.read_texture(i) for i in range(3)], (f.texture_to_update, f.texture_to_blend_from, f.texture_to_blend_to))
        if name == "gpu":  # a completed texture == one full-frame dispatch with its frame data and the LUT it was rendered with
            done = f.texture_to_blend_to
            fd_snap = None
        sky.close(); ctx.close()
    assert [s[:10] for s in out["gpu"][0]] == [s[:10] for s in out["oracle"][0]]
    assert [s[10] for s in out["gpu"][0]] == [s[10] for s in out["oracle"][0]]  # FrameData bit-identical
    assert out["gpu"][2] == out["oracle"][2]
    for i in range(3):
        if i == out["gpu"][2][0]:
            continue  # the texture being updated is a mix of two cycles on both sides; compare the finished ones
        frac, mx = helpers.compare_images(out["gpu"][1][i], out["oracle"][1][i], 2e-3, 1e-2)
        assert frac >= 0.998, (i, frac, mx)


@pytest.mark.gpu
def test_gpu_sky_tiles_equal_single_dispatch(cs, product_lib, small_textures, helpers):
    ctx = product_lib.context(0)
    ctx.upload_textures(*small_textures); ctx.build_transmittance_lut(); ctx.set_march_config(64, 6, cs.MODE_FAST)
    s = make_settings(product_lib, size=128, frames=16)
    sky = cs.Sky(ctx, s)
    sky.set_sun(SUN_BASIS, 1.0, (1.0, 1.0, 1.0))
    sky.update(3.0)  # initialize_sky: 32 tiles + 1
    f = sky.frame()
    finished = f.texture_to_blend_to  # rendered completely with the frame data of t = 3.0
    tiled = sky.read_texture(finished)
    ctx.build_sky_lut(tuple(f.frame_data.light_direction))
    p = product_lib.fill_cloud_params(s, f.frame_data, 128, 128, 0, 0)
    ctx.resize(128, 128)  # the sky renders into its own textures and leaves the context's image alone
    ctx.render_frame(p)
    assert (ctx.read_image().view(np.uint16) == tiled.view(np.uint16)).all()
    sky.close(); ctx.close()

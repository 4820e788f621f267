"""The Python mirror of cloud_sky.gd / sky_lut.gd / transmittance_lut.gd / sun.gd (cloudsky_b200.sky),
driven on the CPU through the oracle library, reads like the reference's own usage."""
import numpy as np
import pytest


@pytest.fixture()
def sky(cs, oracle_lib, small_textures):
    s = cs.CloudSky(oracle_lib)
    s.ctx.set_threads(4)
    s.ctx.set_march_config(16, 2)
    s._s.texture_size = 64
    s.frames_to_update = 4          # setter: cleanup + update_performance + request_full_sky_init
    s.load_textures(*small_textures)
    yield s
    s.close()


def test_defaults_and_setters(cs, oracle_lib):
    s = cs.CloudSky(oracle_lib)
    assert (s.wind_direction, s.wind_speed, s.time_offset, s.frames_to_update, s.texture_size) == (0.0, 1.0, 0.0, 64, 768)
    assert abs(s.cloud_coverage - 0.2) < 1e-7 and abs(s.density - 0.05) < 1e-7       # clouds_sky.tres:13-14
    assert (s.update_region_size, s.num_workgroups) == (96, 12)                      # cloud_sky.gd:83-84
    assert not s.can_run                                                             # no textures yet
    s.update_sky(1.0)                                                                # silently returns (cloud_sky.gd:130-131)
    assert s.frame == 0
    s.texture_size = 1000
    s.frames_to_update = 64
    assert (s.texture_size, s.update_region_size, s.num_workgroups) == (1000, 125, 16)
    s.texture_size = 1001                                                            # coerced (cloud_sky.gd:112-114)
    assert s.texture_size == 1000
    s2 = cs.CloudSky(oracle_lib, demo_values=False)
    assert abs(s2.cloud_coverage - 0.25) < 1e-7 and s2.ground_color == (1.0, 1.0, 1.0, 1.0)
    s.close(); s2.close()


def test_update_sky_tile_walk_equals_single_dispatch(cs, oracle_lib, sky):
    sky.sun = cs.DirectionalLight.looking_from((0.3, 0.8, 0.1), light_energy=1.2, light_color=(1.0, 0.9, 0.8))
    sky.update_sky(2.0)   # first call: initialize_sky renders 2 * frames_to_update tiles (cloud_sky.gd:124-127)
    assert sky.frame == 1 and sky.sky_lut.updates >= 2
    np.testing.assert_allclose(np.linalg.norm(list(sky.frame_data.light_direction)), 1.0, atol=1e-6)
    assert sky.textures[0] is not None and sky.textures[1] is not None
    # keep ticking: blend_amount ramps 0, 1/4, 2/4, 3/4 and the tile origin walks in raster order
    seen = []
    for k in range(4):
        seen.append((tuple(sky.update_position), sky.blend_amount))
        sky.update_sky(3.0 + k)
    assert [b for _, b in seen] == [0.0, 0.25, 0.5, 0.75]
    assert [p for p, _ in seen] == [(32, 0), (0, 32), (32, 32), (0, 0)]
    # a completed texture equals one full-frame dispatch with the same frame data
    done = sky.textures[(sky.texture_to_update - 1) % 3]
    p = sky._fill_push_constant(); p.update_position[0] = 0; p.update_position[1] = 0
    # the completed texture was rendered with the frame data snapshotted at its cycle start (time 2.0 data)
    ref = cs.CloudSky(oracle_lib)
    ref.ctx.set_threads(4); ref.ctx.set_march_config(16, 2)
    ref._s.texture_size = 64; ref.frames_to_update = 1
    ref.load_textures(*[sky_tex for sky_tex in sky_textures(sky)])
    ref.sun = sky.sun
    img = ref.render_full(2.0)
    assert done.shape == img.shape == (64, 64, 4)
    assert (done.view(np.uint16) == img.view(np.uint16)).all()
    ref.close()


def sky_textures(sky):
    """Read the level-0 texels back out of the library so the second sky gets identical inputs."""
    from cloudsky_b200 import assets
    return assets.synthetic_textures(seed=7, large_n=16, small_n=8, weather_n=32)


def test_luts_and_uninitialised_sky_lut(cs, oracle_lib, sky, capsys):
    assert sky.transmittance_tex.read().shape == (64, 256, 4)
    sky.sky_lut.update_lut((0.0, 1.0, 0.0))
    assert sky.sky_lut.read().shape == (100, 200, 4) and not sky.sky_lut.needs_update
    sky.sky_lut.request_update()
    assert sky.sky_lut.needs_update
    orphan = cs.SkyLUT(sky.ctx, None)
    orphan.update_lut((0.0, 1.0, 0.0))
    assert "uninitialized sky lut" in capsys.readouterr().out   # sky_lut.gd:45-47

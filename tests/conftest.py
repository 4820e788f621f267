import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libcloudsky_oracle.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _build_oracle():
    if not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(os.path.join(ORACLE_DIR, "cloudsky_oracle.cpp")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


@pytest.fixture(scope="session")
def cs():
    import cloudsky_b200
    return cloudsky_b200


@pytest.fixture(scope="session")
def oracle_lib(cs):
    """The CPU oracle behind the same C-ABI — the checker, never the thing under test."""
    _build_oracle()
    return cs.Library(ORACLE_LIB)


@pytest.fixture(scope="session")
def product_lib(cs):
    """libcloudsky_b200.so (fails loudly when it has not been built)."""
    return cs.load_product()


@pytest.fixture(scope="session")
def textures(cs):
    from cloudsky_b200 import assets
    assert assets.fixture_available(), "tests/golden/assets fixture missing"
    return assets.load_fixture()


@pytest.fixture(scope="session")
def small_textures(cs):
    """Small synthetic volumes (16^3 / 8^3 / 32^2) for edge-case tests."""
    from cloudsky_b200 import assets
    return assets.synthetic_textures(seed=7, large_n=16, small_n=8, weather_n=32)


def make_params(lib, width, height, sun=(0.0, 1.0, 0.0), coverage=None, density=None, time=0.0, wind_direction=0.0,
                wind_speed=1.0, energy=1.0, color=(1.0, 1.0, 1.0), demo=True, time_offset=0.0):
    """Settings -> FrameData -> push constants, through the library's own host logic."""
    s = lib.settings_demo() if demo else lib.settings_default()
    if coverage is not None:
        s.cloud_coverage = coverage
    if density is not None:
        s.density = density
    s.wind_direction = wind_direction
    s.wind_speed = wind_speed
    s.time_offset = time_offset
    st = lib.frame_state_init()
    n = float(np.sqrt(sum(v * v for v in sun)))
    st.light_direction[:] = [v / n for v in sun]
    st.light_energy = energy
    st.light_color[:] = list(color)
    if time:
        lib.frame_advance(st, s, time)
    return lib.fill_cloud_params(s, st, width, height)


def prepared_context(lib, textures, width, height, sun=(0.0, 1.0, 0.0), threads=None, device=0):
    ctx = lib.context(device)
    if threads:
        ctx.set_threads(threads)
    ctx.upload_textures(*textures)
    ctx.build_transmittance_lut()
    ctx.build_sky_lut(sun)
    ctx.resize(width, height)
    return ctx


def compare_images(test, ref, atol, rtol, min_pass=0.999, skip_edges=True):
    """Per-channel |test - ref| <= atol + rtol*|ref|; returns (pass fraction over pixels, max abs err)."""
    t = np.asarray(test, np.float32)
    r = np.asarray(ref, np.float32)
    if skip_edges:  # row 0 / column 0: dir.y > 0 is a rounding coin flip there
        t, r = t[1:, 1:], r[1:, 1:]
    d = np.abs(t - r)
    ok = (d <= atol + rtol * np.abs(r)).all(-1)
    return float(ok.mean()), float(d.max())


@pytest.fixture(scope="session")
def helpers():
    class H:
        pass
    H.make_params = staticmethod(make_params)
    H.prepared_context = staticmethod(prepared_context)
    H.compare_images = staticmethod(compare_images)
    H.cpu_threads = max(1, min(os.cpu_count() or 1, 16))
    return H

"""The thin C++ driver (tools/cloudsky_cli.cpp) over the C-ABI: decodes TGA/BMP strips written here from the
fixture texels (cs_load_texture_files: RLE TGA, bottom-up BMP, strip slicing), renders, and must reproduce what
the Python path renders from the decoded arrays, bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "tools", "cloudsky_cli")


def test_cli_builds_and_prints_usage():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s"])
    out = subprocess.run([CLI, "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "usage: cloudsky_cli" in out.stdout


def test_cli_procedural_inputs_through_any_backend(cs, oracle_lib, helpers, tmp_path):
    """--procedural: cs_generate_noise -> cs_upload_textures -> render, here through the oracle's implementation of the ABI
    (the driver only sees include/cloudsky.h); must equal the same calls made from Python."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s"])
    out_f16 = tmp_path / "out.f16"
    W, H = 48, 24
    r = subprocess.run([CLI, "--lib", oracle_lib.path, "--procedural", "5", "32", "16", "64", "--bruneton", "--size", str(W), str(H), "--steps", "24", "3",
                        "--sun", "0", "1", "0", "--coverage", "0.5", "--out", str(out_f16)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "backend oracle-cpu" in r.stdout
    got = np.fromfile(out_f16, dtype=np.float16).reshape(H, W, 4)
    ctx = oracle_lib.context(0)
    ctx.set_threads(helpers.cpu_threads)
    tex = []
    for kind, n in ((cs.NOISE_LARGE, 32), (cs.NOISE_SMALL, 16), (cs.NOISE_WEATHER, 64)):
        p = oracle_lib.noise_params_default(kind); p.seed = 5
        tex.append(ctx.generate_noise(kind, n, p))
    ctx.upload_textures(*tex)
    ctx.set_transmittance_parametrisation(cs.TLUT_BRUNETON2017)
    ctx.build_transmittance_lut(); ctx.resize(W, H); ctx.set_march_config(24, 3, cs.MODE_FAST)
    want = ctx.render_frame_host(helpers.make_params(oracle_lib, W, H, coverage=0.5))
    ctx.close()
    assert (got.view(np.uint16) == want.view(np.uint16)).all()
    assert 0.02 < float(got[..., 3].astype(np.float32).mean()) < 0.98


@pytest.mark.gpu
def test_cli_matches_python_path(cs, product_lib, textures, helpers, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_assets import write_bmp, write_tga
    large, small, weather = textures
    # volume [z][y][x][c] -> horizontal strip [y][z*n + x][c] (perlworlnoise.tga.import:26)
    strip = lambda v: np.ascontiguousarray(v.transpose(1, 0, 2, 3).reshape(v.shape[1], v.shape[0] * v.shape[2], v.shape[3]))
    write_tga(str(tmp_path / "perlworlnoise.tga"), strip(large), rle=True, top_origin=False)
    write_bmp(str(tmp_path / "worlnoise.bmp"), strip(small), 24, False)
    write_bmp(str(tmp_path / "weather.bmp"), weather, 24, False)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s"])
    out_f16 = tmp_path / "out.f16"
    W, H = 192, 96
    r = subprocess.run([CLI, "--lib", product_lib.path, "--assets", str(tmp_path), "--size", str(W), str(H), "--steps", "128", "6",
                        "--sun", "0", "1", "0", "--time", "7.5", "--out", str(out_f16), "--ppm", str(tmp_path / "out.ppm")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "backend cuda-sm100a" in r.stdout and "Mray-steps/s" in r.stdout
    got = np.fromfile(out_f16, dtype=np.float16).reshape(H, W, 4)
    ctx = helpers.prepared_context(product_lib, textures, W, H)
    p = helpers.make_params(product_lib, W, H, sun=(0.0, 1.0, 0.0), time=7.5)  # an exactly representable unit vector: both hosts normalise it identically
    ctx.set_march_config(128, 6, cs.MODE_FAST)
    want = ctx.render_frame_host(p)
    assert (got.view(np.uint16) == want.view(np.uint16)).all()
    assert os.path.getsize(tmp_path / "out.ppm") > W * H * 3
    ctx.close()

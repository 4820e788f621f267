"""Asset pipeline: the product's TGA/BMP decoders against synthetic files written here, against the real
reference bitmaps when /root/reference is mounted (build container only), and the committed fixture
against its manifest."""
import hashlib
import json
import os
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/cloud_sky"


def write_tga(path, img, rle, top_origin):
    """img: [h][w][3|4] RGB(A) uint8, row 0 = top.  Writes type 2 / 10 true-colour TGA (BGR(A) on disk)."""
    h, w, ch = img.shape
    rows = img if top_origin else img[::-1]
    px = rows[..., [2, 1, 0, 3][:ch] if ch == 4 else [2, 1, 0]].reshape(-1, ch)
    hdr = struct.pack("<BBBHHBHHHHBB", 0, 0, 10 if rle else 2, 0, 0, 0, 0, 0, w, h, ch * 8, (0x20 if top_origin else 0) | (8 if ch == 4 else 0))
    body = bytearray()
    if not rle:
        body += px.tobytes()
    else:
        i, n = 0, len(px)
        while i < n:
            run = 1
            while i + run < n and run < 128 and (px[i + run] == px[i]).all():
                run += 1
            if run > 1:
                body.append(0x80 | (run - 1)); body += px[i].tobytes(); i += run
            else:
                lit = 1
                while i + lit < n and lit < 128 and not (i + lit + 1 < n and (px[i + lit] == px[i + lit + 1]).all()):
                    lit += 1
                body.append(lit - 1); body += px[i:i + lit].tobytes(); i += lit
    open(path, "wb").write(hdr + bytes(body))


def write_bmp(path, img, bpp=24, top_down=False):
    h, w, _ = img.shape
    bytes_pp = bpp // 8
    stride = (w * bytes_pp + 3) & ~3
    rows = img if top_down else img[::-1]
    data = bytearray()
    for r in rows:
        line = bytearray()
        for p in r:
            line += bytes([p[2], p[1], p[0]] + ([0] if bytes_pp == 4 else []))
        line += b"\0" * (stride - len(line))
        data += line
    hdr = b"BM" + struct.pack("<IHHI", 54 + len(data), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, w, -h if top_down else h, 1, bpp, 0, len(data), 2835, 2835, 0, 0)
    open(path, "wb").write(hdr + bytes(data))


@pytest.mark.parametrize("rle", [False, True])
@pytest.mark.parametrize("top", [False, True])
@pytest.mark.parametrize("ch", [3, 4])
def test_tga_decode(product_lib, tmp_path, rle, top, ch):
    rng = np.random.default_rng(ch * 4 + rle * 2 + top)
    img = rng.integers(0, 256, (9, 37, ch), dtype=np.uint8)
    img[2, 3:30] = img[2, 3]  # long runs, also crossing the 128-pixel packet limit below
    img[5:8] = 77
    p = str(tmp_path / "t.tga")
    write_tga(p, img, rle, top)
    out = product_lib.decode_image_file(p)
    assert out.shape == img.shape
    np.testing.assert_array_equal(out, img)


@pytest.mark.parametrize("bpp,top_down", [(24, False), (24, True), (32, False)])
def test_bmp_decode(product_lib, tmp_path, bpp, top_down):
    rng = np.random.default_rng(bpp + top_down)
    img = rng.integers(0, 256, (11, 13, 3), dtype=np.uint8)  # 13*3 = 39 bytes -> rows padded to 40
    p = str(tmp_path / "t.bmp")
    write_bmp(p, img, bpp, top_down)
    out = product_lib.decode_image_file(p)
    np.testing.assert_array_equal(out, img)


def test_decode_errors(cs, product_lib, tmp_path):
    with pytest.raises(cs.CloudSkyError) as e:
        product_lib.decode_image_file(str(tmp_path / "missing.tga"))
    assert e.value.code == 3
    bad = tmp_path / "bad.bmp"
    bad.write_bytes(b"BM" + b"\0" * 10)
    with pytest.raises(cs.CloudSkyError):
        product_lib.decode_image_file(str(bad))
    trunc = tmp_path / "trunc.tga"
    img = np.zeros((4, 4, 3), np.uint8)
    write_tga(str(trunc), img, False, True)
    trunc.write_bytes(trunc.read_bytes()[:-5])
    with pytest.raises(cs.CloudSkyError):
        product_lib.decode_image_file(str(trunc))


def test_fixture_matches_manifest(textures):
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "assets", "manifest.json")))
    large, small, weather = textures
    assert large.shape == (128, 128, 128, 4) and small.shape == (32, 32, 32, 3) and weather.shape == (512, 512, 3)
    for arr, key in ((large, "large_128_rgba8.xz"), (small, "small_32_rgb8.xz"), (weather, "weather_512_rgb8.xz")):
        assert hashlib.sha256(arr.tobytes()).hexdigest() == man[key]["interleaved_sha256"]
    # decoded-strip sha256 prefixes measured from the reference files
    assert man["large_128_rgba8.xz"]["decoded_strip_sha256"].startswith("6702c6f5c780c0cb")
    assert man["small_32_rgb8.xz"]["decoded_strip_sha256"].startswith("b6f44679543d510d")
    assert man["weather_512_rgb8.xz"]["decoded_strip_sha256"].startswith("2c15fb3c19a0e5fc")
    # channel statistics of the reference textures
    np.testing.assert_allclose(large.reshape(-1, 4).mean(0) / 255.0, [0.85, 0.71, 0.71, 0.71], atol=0.01)
    assert 0.58 < weather[..., 0].min() / 255 < 0.60 and 0.90 < weather[..., 0].max() / 255 < 0.92
    assert weather[..., 2].min() == 17 and weather[..., 2].max() == 255


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference only exists in the build container")
def test_real_reference_assets_decode_to_the_fixture(product_lib, textures):
    """The product's own TGA/BMP decoders reproduce what PIL decoded into the fixture, bit for bit."""
    large, small, weather = textures
    strip = product_lib.decode_image_file(os.path.join(REF, "perlworlnoise.tga"))
    assert strip.shape == (128, 16384, 4)
    assert hashlib.sha256(strip.tobytes()).hexdigest().startswith("6702c6f5c780c0cb")
    vol = strip.reshape(128, 128, 128, 4).transpose(1, 0, 2, 3)  # texel (x,y,z) = column z*128+x, row y
    np.testing.assert_array_equal(vol, large)
    s = product_lib.decode_image_file(os.path.join(REF, "worlnoise.bmp"))
    assert s.shape == (32, 1024, 3)
    np.testing.assert_array_equal(s.reshape(32, 32, 32, 3).transpose(1, 0, 2, 3), small)
    w = product_lib.decode_image_file(os.path.join(REF, "weather.bmp"))
    np.testing.assert_array_equal(w, weather)


def test_synthetic_textures_have_reference_shapes(cs):
    from cloudsky_b200 import assets
    l, s, w = assets.synthetic_textures(seed=1, large_n=16, small_n=8, weather_n=32)
    assert l.shape == (16, 16, 16, 4) and s.shape == (8, 8, 8, 3) and w.shape == (32, 32, 3)
    assert l.dtype == np.uint8 and w[..., 2].min() >= 17
    l2, _, _ = assets.synthetic_textures(seed=1, large_n=16, small_n=8, weather_n=32)
    np.testing.assert_array_equal(l, l2)  # deterministic

"""Pins every numeric literal of the three reference shaders to the restatements (build container only: needs
/root/reference).  The reference ships no golden vectors, but its constants are checkable text: each float literal in
clouds.glsl / sky-lut.glsl / transmittance-lut.glsl must occur, with the same fp32 value, in the oracle and in the
reference-order CUDA sources — a digit typo in a physical constant would otherwise go unnoticed by self-consistent tests."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/cloud_sky"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference only exists in the build container")

LIT = re.compile(r"(?<![\w.])(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)f?(?![\w.])")


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def literals(path):
    vals = set()
    for m in LIT.finditer(strip_comments(open(path).read())):
        vals.add(np.float32(float(m.group(1))))
    return vals


def f32set(paths):
    out = set()
    for p in paths:
        out |= literals(p)
    return out


CSRC = os.path.join(ROOT, "godot-volumetric-cloud-demo-v2_b200", "csrc")
@pytest.mark.parametrize("shader,targets", [
    ("clouds.glsl", ["oracle/cloudsky_oracle.cpp"]),
    ("clouds.glsl", [os.path.join(CSRC, "clouds_generic.cuh"), os.path.join(CSRC, "cs_device.cuh"), os.path.join(CSRC, "clouds_strict.cu")]),
    ("sky-lut.glsl", ["oracle/cloudsky_oracle.cpp"]),
    ("sky-lut.glsl", [os.path.join(CSRC, "lut_kernels.cu"), os.path.join(CSRC, "cs_device.cuh")]),
    ("transmittance-lut.glsl", ["oracle/cloudsky_oracle.cpp"]),
    ("transmittance-lut.glsl", [os.path.join(CSRC, "lut_kernels.cu"), os.path.join(CSRC, "cs_device.cuh")]),
    ("clouds.gdshader", ["oracle/cloudsky_oracle.cpp"]),
    ("clouds.gdshader", [os.path.join(CSRC, "composite.cu"), os.path.join(CSRC, "cs_device.cuh")]),
])
def test_every_shader_literal_is_restated(shader, targets):
    want = literals(os.path.join(REF, shader))
    have = f32set([t if os.path.isabs(t) else os.path.join(ROOT, t) for t in targets])
    # trivial literals are everywhere; compare the distinctive ones
    trivial = {np.float32(v) for v in (0.0, 1.0, 2.0, 0.5, 3.0, 4.0, 5.0, 8.0)}
    # the unused spectral matrix M of transmittance-lut.glsl:150-155 (declared, never read) need not be restated there
    unused = set()
    if shader == "transmittance-lut.glsl":
        unused = {np.float32(v) for v in (137.672389239975, -8.632904716299537, 8.632904716299537, 1.7181567391931372, 32.549094028629234,
                                          91.29801417199785, 12.005406444382531, 38.91428392614275, 34.31665471469816, 29.89044807197628,
                                          8.572844237945445, 11.103384660054624, 117.47585277566478)}
    missing = sorted(float(v) for v in want - have - trivial - unused)
    allowed_missing = {
        # clouds.glsl:228 "float steps = 128.0" is the primary_steps parameter here (CS_REF_PRIMARY_STEPS, checked below)
        "clouds.glsl": {128.0},
        "sky-lut.glsl": set(),
        "transmittance-lut.glsl": set(),
        # uniform defaults / hints of the material (blend_amount hint_range step 0.01) are not shader arithmetic
        "clouds.gdshader": {0.01},
    }[shader]
    really_missing = [v for v in missing if np.float32(v) not in {np.float32(a) for a in allowed_missing}]
    assert not really_missing, f"{shader}: literals not found in {targets}: {really_missing}"


def test_reference_step_counts_are_the_defaults():
    hdr = open(os.path.join(ROOT, "include", "cloudsky.h")).read()
    assert re.search(r"#define CS_REF_PRIMARY_STEPS 128\b", hdr) and re.search(r"#define CS_REF_CONE_SAMPLES 6\b", hdr)
    src = strip_comments(open(os.path.join(REF, "clouds.glsl")).read())
    assert "float steps = 128.0;" in src and "for (int j = 0; j < 6; j++)" in src  # clouds.glsl:228,186


def test_host_script_constants_are_restated():
    """cloud_sky.gd's integration constants (delta * 0.001 + 0.005 * time_offset, cloud_sky.gd:176) and defaults."""
    gd = open(os.path.join(REF, "cloud_sky.gd")).read()
    assert "delta * 0.001 + 0.005 * frame_data.time_offset" in gd
    for target in ("oracle/cloudsky_oracle.cpp", os.path.join(CSRC, "host_logic.cpp")):
        have = literals(target if os.path.isabs(target) else os.path.join(ROOT, target))
        for v in (0.001, 0.005, 0.05, 0.25, 12.92, 0.04045, 0.055, 2.4, 0.270588, 0.188235, 0.027451):
            assert np.float32(v) in have, (target, v)

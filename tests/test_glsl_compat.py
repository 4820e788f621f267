"""oracle/glsl_compat.h gives GLSL's vocabulary a C++ meaning so that the reference shaders compile unmodified (oracle/_ref).
These checks hold it to the GLSL 4.50 / Vulkan rules directly — swizzle write-through, built-in formulas, column-major mat4x3,
texel-centre / REPEAT / CLAMP_TO_EDGE sampling, fp16 round-to-nearest-even stores, discarded out-of-range image stores —
independently of any shader and of the hand-written oracle."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_glsl_compat_semantics(tmp_path):
    exe = str(tmp_path / "glsl_compat_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fsingle-precision-constant", "-ffp-contract=off", "-fno-fast-math", "-Wno-unused",
                           "-I", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "cxx", "glsl_compat_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "glsl_compat ok" in out.stdout, out.stdout + out.stderr

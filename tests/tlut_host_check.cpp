// TEST HARNESS, CPU only — not part of the product and not linked into libcloudsky_b200.so.
// Compiles the CS_TLUT_BRUNETON2017 coordinate mappings the CUDA kernels execute (csrc/tlut_param.h) with g++ so that
// tests/test_transmittance_parametrisation.py can compare them with the oracle's statement when no GPU is present.
#include "../godot-volumetric-cloud-demo-v2_b200/csrc/tlut_param.h"

extern "C" void tlh_ray_from_texel(int px, int py, float out[4]) { tl::bruneton_ray_from_texel(px, py, out[0], out[1], out[2], out[3]); }
extern "C" void tlh_uv(float normalized_altitude, float mu, float out[3]) { tl::bruneton_uv(normalized_altitude, mu, out[0], out[1], out[2]); }

// TEST HARNESS, CPU only — not part of the product and not linked into libcloudsky_b200.so.
// Compiles the texel arithmetic the CUDA generator kernel executes (csrc/noise_core.h) with g++ so that
// tests/test_noise_generator.py can compare it with the oracle's independent statement when no GPU is present
// (-ffp-contract=off here == --fmad=false in noise_gen.cu).  The GPU test then compares the kernel's bytes themselves.
#include <cstddef>
#include <cstdint>

#include "../godot-volumetric-cloud-demo-v2_b200/csrc/noise_core.h"

extern "C" int nzh_generate(int kind, int n, const cs_noise_params* P, uint8_t* out) {
    if (nz::check_request(kind, n, P)) return 1;
    const int depth = kind == CS_NOISE_WEATHER ? 1 : n;
    for (int z = 0; z < depth; z++)
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) {
                uint8_t* t = out + (((size_t)z * n + y) * n + x) * 4;
                if (kind == CS_NOISE_LARGE) nz::large_texel(*P, n, x, y, z, t);
                else if (kind == CS_NOISE_SMALL) nz::small_texel(*P, n, x, y, z, t);
                else nz::weather_texel(*P, n, x, y, t);
            }
    return 0;
}

// Noise synthesis kernel (cs_generate_noise): one thread per texel of the requested product, the texel arithmetic of
// noise_core.h.  A one-off, compute-only generator (no reads, 4 bytes written per texel; a 128^3 Perlin-Worley volume is
// ~2 M texels x ~17 k integer/fp32 operations), so the only launch rule that matters is filling the machine: 256-thread
// CTAs over a flat texel index, x fastest so the uchar4 stores of a warp are one 128-byte line.
// Compiled in the accurate configuration (--fmad=false, IEEE sqrt/div) so the bytes equal the oracle's.
#include <cuda_runtime.h>

#include "cs_internal.h"
#include "noise_core.h"

namespace {

template <int KIND>
__global__ void __launch_bounds__(256) noise_kernel(const __grid_constant__ cs_noise_params P, int n, int log2n, uchar4* __restrict__ out) {
    const size_t total = KIND == CS_NOISE_WEATHER ? (size_t)n * n : (size_t)n * n * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i & (size_t)(n - 1)), y = (int)((i >> log2n) & (size_t)(n - 1)), z = (int)(i >> (2 * log2n));
        uint8_t t[4];
        if constexpr (KIND == CS_NOISE_LARGE) nz::large_texel(P, n, x, y, z, t);
        else if constexpr (KIND == CS_NOISE_SMALL) nz::small_texel(P, n, x, y, z, t);
        else nz::weather_texel(P, n, x, y, t);
        out[i] = make_uchar4(t[0], t[1], t[2], t[3]);
    }
}

}  // namespace

namespace cs {

void launch_noise(int kind, int n, const cs_noise_params& P, uint32_t* out_rgba8, void* stream) {
    int log2n = 0;
    while ((1 << log2n) < n) log2n++;
    const size_t total = kind == CS_NOISE_WEATHER ? (size_t)n * n : (size_t)n * n * n;
    const unsigned blocks = (unsigned)((total + 255) / 256 < 148u * 64u ? (total + 255) / 256 : 148u * 64u);  // grid-stride beyond 64 CTAs per SM
    cudaStream_t st = (cudaStream_t)stream;
    uchar4* out = reinterpret_cast<uchar4*>(out_rgba8);
    if (kind == CS_NOISE_LARGE) noise_kernel<CS_NOISE_LARGE><<<blocks, 256, 0, st>>>(P, n, log2n, out);
    else if (kind == CS_NOISE_SMALL) noise_kernel<CS_NOISE_SMALL><<<blocks, 256, 0, st>>>(P, n, log2n, out);
    else noise_kernel<CS_NOISE_WEATHER><<<blocks, 256, 0, st>>>(P, n, log2n, out);
}

}  // namespace cs

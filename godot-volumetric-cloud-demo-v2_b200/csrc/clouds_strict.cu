// Strict-parity cloud march (CS_MODE_STRICT) and the per-dispatch prologue kernel.
// This translation unit is compiled with --fmad=false -prec-div=true -prec-sqrt=true so that
// every fp32 operation rounds exactly like the CPU oracle's (-ffp-contract=off); the only
// remaining differences are the last-ulp behaviour of expf/powf/atan2f/asinf.
#include "clouds_generic.cuh"

using namespace csd;

namespace {

template <bool STRICT>
__global__ void clouds_prologue_kernel(const __grid_constant__ cs::CloudLaunch L, cs::FrameConsts* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        cs::FrameConsts fc;
        compute_frame_consts<STRICT>(L, fc);
        *out = fc;
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(64) clouds_strict_kernel(const __grid_constant__ cs::CloudLaunch L) {
    const int by = blockIdx.y, band = L.band_ctas ? by / L.band_ctas : 0;
    const int row0 = L.band_ctas ? band * L.band_pitch_rows + (by - band * L.band_ctas) * 8 : by * 8;  // interleaved bands (cs_render_row_bands_to)
    int px = L.x0 + blockIdx.x * 8 + threadIdx.x, py = L.y0 + row0 + threadIdx.y;  // 8x8 groups (clouds.glsl:5)
    if (px >= L.x1 || py >= L.y1) return;  // the reference does not bounds-check; we do
    const cs::FrameConsts fc = *reinterpret_cast<const cs::FrameConsts*>(L.frame_consts);
    V3 dir = pixel_direction<true>(px, py, L.P.texture_size[0], L.P.texture_size[1]);
    V4 col = {0.0f, 0.0f, 0.0f, 0.0f};
    Tally tl = {0u, 0u, 0u};
    bool marched = dir.y > 0.0f;  // clouds.glsl:221
    if (marched) col = sky_pixel_ref<true, COUNT>(L, fc, dir, tl);
    ushort4 o = {f2h(col.x), f2h(col.y), f2h(col.z), f2h(col.w)};
    const size_t at = (size_t)py * L.out_pitch_px + px;
    reinterpret_cast<ushort4*>(L.out)[at] = o;
    for (int m = 0; m < L.n_mirrors; m++) reinterpret_cast<ushort4*>(L.mirror[m])[at] = o;
    if constexpr (COUNT) {
        atomicAdd(L.counters + 0, marched ? 1ull : 0ull);
        atomicAdd(L.counters + 1, (unsigned long long)tl.steps);
        atomicAdd(L.counters + 2, (unsigned long long)tl.lit);
        atomicAdd(L.counters + 3, (unsigned long long)tl.evals);
        atomicAdd(L.counters + 4, (unsigned long long)tl.evals);
        atomicAdd(L.counters + 5, (unsigned long long)tl.evals);
    }
}

}  // namespace

namespace cs {

void launch_clouds_prologue(const CloudLaunch& L, bool strict, void* stream) {
    FrameConsts* out = reinterpret_cast<FrameConsts*>(const_cast<float*>(L.frame_consts));
    (void)strict;  // one accurate prologue serves both modes: it runs once per dispatch
    clouds_prologue_kernel<true><<<1, 32, 0, (cudaStream_t)stream>>>(L, out);
}

void launch_clouds_strict(const CloudLaunch& L, void* stream) {
    dim3 block(8, 8), grid((L.x1 - L.x0 + 7) / 8, L.grid_y > 0 ? L.grid_y : (L.y1 - L.y0 + 7) / 8);
    if (grid.x == 0 || grid.y == 0) return;
    if (L.counters) clouds_strict_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(L);
    else clouds_strict_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(L);
}

}  // namespace cs

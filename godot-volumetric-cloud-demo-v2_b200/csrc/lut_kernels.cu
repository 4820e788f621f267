// Atmosphere LUT precomputes for sm_100a.
//
//   transmittance_lut_kernel : transmittance-lut.glsl:157-196  (256x64 RGBA16F, 40-step extinction integral)
//   sky_lut_kernel           : sky-lut.glsl:219-315            (200x100 RGBA16F, 30-step in-scattering integral;
//                              persistent, transmittance LUT staged in shared memory by TMA bulk copies, warp per texel)
//
// Both are launch-latency-scale (16 K and 20 K threads); they are compiled in the accurate
// configuration (--fmad=false, IEEE div/sqrt, libm-grade exp/pow/log) so that they track the
// CPU oracle to the last fp16 ulp.  The transmittance kernel is one thread per texel in 8x8 blocks exactly
// like the reference dispatch (transmittance_lut.gd:77 -> 32x8 groups).
#include "cs_device.cuh"
#include "cs_internal.h"
#include "tlut_param.h"

using namespace csd;

namespace {

constexpr bool S = true;  // accurate math in this translation unit

// Shared atmosphere model: transmittance-lut.glsl:45-145 == sky-lut.glsl:56-202.
constexpr float EARTH_RADIUS = 6371.0f;
constexpr float ATMOSPHERE_THICKNESS = 100.0f;
constexpr float ATMOSPHERE_RADIUS = EARTH_RADIUS + ATMOSPHERE_THICKNESS;

struct Coeffs { V4 aerosol_scattering, molecular_scattering, extinction; };

__device__ __forceinline__ float ray_sphere_intersection(V3 ro, V3 rd, float radius) {
    float b = dot3(ro, rd);
    float c = dot3(ro, ro) - radius * radius;
    if (c > 0.0f && b > 0.0f) return -1.0f;
    float d = b * b - c;
    if (d < 0.0f) return -1.0f;
    if (d > b * b) return (-b + sqrtf(d));
    return (-b - sqrtf(d));
}

__device__ __forceinline__ Coeffs atmosphere_coefficients(float h) {
    const V4 mol_base = {6.605e-3f, 1.067e-2f, 1.842e-2f, 3.156e-2f};
    const V4 ozone_xs = {3.472e-21f * 1e-4f, 3.914e-21f * 1e-4f, 1.349e-21f * 1e-4f, 11.03e-23f * 1e-4f};
    const V4 aer_abs_xs = {2.8722e-24f, 4.6168e-24f, 7.9706e-24f, 1.3578e-23f};
    const V4 aer_sca_xs = {1.5908e-22f, 1.7711e-22f, 2.0942e-22f, 2.4033e-22f};
    const float aerosol_base_density = 1.3681e20f;
    const float bg_over_base = 2e6f / 1.3681e20f;
    h = fmaxf(h, 0.0f);
    float aerosol_density = aerosol_base_density * (expf(-h / 0.73f) + bg_over_base);
    V4 aer_abs = aer_abs_xs * aerosol_density;
    V4 aer_sca = aer_sca_xs * aerosol_density;
    float ho = h + 1e-4f;
    float t = logf(ho) - 3.22261f;
    float ozone_density = 3.78547397e20f * (1.0f / ho) * expf(-t * t * 5.55555555f);
    V4 mol_abs = (ozone_xs * 350.0f) * ozone_density;
    V4 mol_sca = mol_base * expf(-0.07771971f * powf(h, 1.16364243f));
    Coeffs c;
    c.aerosol_scattering = aer_sca;
    c.molecular_scattering = mol_sca;
    c.extinction = aer_abs + aer_sca + mol_abs + mol_sca;
    return c;
}

__global__ void __launch_bounds__(64) transmittance_lut_kernel(uint16_t* __restrict__ out, int param) {
    int px = blockIdx.x * 8 + threadIdx.x, py = blockIdx.y * 8 + threadIdx.y;
    if (px >= CS_TRANSMITTANCE_W || py >= CS_TRANSMITTANCE_H) return;
    float u = (float)px / (float)CS_TRANSMITTANCE_W, v = (float)py / (float)CS_TRANSMITTANCE_H;
    float sun_cos_theta = u * 2.0f - 1.0f;
    float distance_to_earth_center = mixf(EARTH_RADIUS, ATMOSPHERE_RADIUS, v);
    float t_d;
    if (param == CS_TLUT_BRUNETON2017) {  // the texel stores the ray (r, mu) of Bruneton's mapping; its length to the top boundary is d
        float h;
        tl::bruneton_ray_from_texel(px, py, h, distance_to_earth_center, sun_cos_theta, t_d);
    }
    V3 sun_dir = {-sqrtf(1.0f - sun_cos_theta * sun_cos_theta), 0.0f, sun_cos_theta};
    V3 ray_origin = {0.0f, 0.0f, distance_to_earth_center};
    if (param != CS_TLUT_BRUNETON2017) t_d = ray_sphere_intersection(ray_origin, sun_dir, ATMOSPHERE_RADIUS);
    float dt = t_d / 40.0f;
    V4 result = splat4(0.0f);
    for (int i = 0; i < 40; ++i) {
        float t = ((float)i + 0.5f) * dt;
        V3 x_t = ray_origin + sun_dir * t;
        float altitude = length3<S>(x_t) - EARTH_RADIUS;
        Coeffs c = atmosphere_coefficients(altitude);
        result = result + c.extinction * dt;
    }
    ushort4 o = {f2h(expf(-result.x)), f2h(expf(-result.y)), f2h(expf(-result.z)), f2h(expf(-result.w))};
    reinterpret_cast<ushort4*>(out)[py * CS_TRANSMITTANCE_W + px] = o;
}

// Bilinear CLAMP_TO_EDGE fetch of the transmittance LUT from a shared-memory copy (same arithmetic as
// sample_lut_half4; plain loads instead of __ldg).
template <bool BRUNETON>
__device__ __forceinline__ V4 transmittance_from_smem(const uint2* __restrict__ lut, float cos_theta, float normalized_altitude) {
    const int w = CS_TRANSMITTANCE_W, h = CS_TRANSMITTANCE_H;
    float su = clampf(cos_theta * 0.5f + 0.5f, 0.0f, 1.0f);
    float sv = clampf(normalized_altitude, 0.0f, 1.0f);
    float visible = 1.0f;
    if constexpr (BRUNETON) tl::bruneton_uv(normalized_altitude, cos_theta, su, sv, visible);
    float ux = su * (float)w - 0.5f, uy = sv * (float)h - 0.5f;
    float fx0 = floorf(ux), fy0 = floorf(uy);
    float fx = ux - fx0, fy = uy - fy0;
    int x0 = (int)fx0, y0 = (int)fy0;
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = min(max(x0, 0), w - 1); x1 = min(max(x1, 0), w - 1);
    y0 = min(max(y0, 0), h - 1); y1 = min(max(y1, 0), h - 1);
    uint2 t00 = lut[y0 * w + x0], t10 = lut[y0 * w + x1], t01 = lut[y1 * w + x0], t11 = lut[y1 * w + x1];
    auto ch = [](uint2 t, int c) -> float {
        uint32_t wv = c < 2 ? t.x : t.y;
        return h2f((uint16_t)((c & 1) ? (wv >> 16) : (wv & 0xffffu)));
    };
    float o[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        float a = lerpf(ch(t00, c), ch(t10, c), fx);
        float b = lerpf(ch(t01, c), ch(t11, c), fx);
        o[c] = lerpf(a, b, fy);
        if constexpr (BRUNETON) o[c] = o[c] * visible;
    }
    return {o[0], o[1], o[2], o[3]};
}

// ---- sky-view LUT (sky-lut.glsl:219-315) -----------------------------------------------------------------------
// Persistent kernel, one CTA of 32 warps per SM.  The 128 KiB transmittance LUT is staged into shared memory once per
// CTA with bulk async copies (TMA, cp.async.bulk + mbarrier).  One WARP per texel: lane i evaluates integration step i
// of compute_inscattering (the extinction model, three LUT fetches, the source term and exp(-dt * extinction), which
// do not depend on the other steps); the running transmittance product and radiance sum are then accumulated in the
// shader's sequential order from shuffled values, so the result is bit-identical to the one-thread-per-texel form.
constexpr int kSkyWarps = 32;
constexpr int kSkyLutBytes = CS_TRANSMITTANCE_W * CS_TRANSMITTANCE_H * 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool BRUNETON>
__global__ void __launch_bounds__(32 * kSkyWarps, 1) sky_lut_kernel(const uint16_t* __restrict__ tlut, float sx, float sy, float sz,
                                                                    uint16_t* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint2* lut = reinterpret_cast<uint2*>(smem_raw);
    __shared__ __align__(8) unsigned long long mbar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(kSkyLutBytes) : "memory");
        constexpr int kChunk = 16384;
        for (int off = 0; off < kSkyLutBytes; off += kChunk)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_raw + off)),
                         "l"(reinterpret_cast<const unsigned char*>(tlut) + off), "r"(kChunk), "r"(smem_u32(&mbar))
                         : "memory");
    }
    {  // every thread waits for the LUT to land (phase 0)
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    }

    const float PI = 3.14159265358979323846f, INV_PI = 0.31830988618379067154f;
    const float INV_4PI = 0.25f * INV_PI, RAYLEIGH_PHASE_SCALE = (3.0f / 16.0f) * INV_PI;
    const float g = 0.8f, gg = g * g;
    const V3 sun_dir = {-sx, -sz, sy};  // params.sun_direction.xzy with x and y negated (:221-223)
    const V4 irradiance = {1.679f, 1.828f, 1.986f, 1.307f};
    const float albedo_over_pi = 0.3f / PI;
    const V4 T_1_0 = transmittance_from_smem<BRUNETON>(lut, 1.0f, 0.0f);  // loop-invariant fetch of :153
    const int n_texels = CS_SKY_LUT_W * CS_SKY_LUT_H;

    for (int texel = blockIdx.x * kSkyWarps + warp; texel < n_texels; texel += gridDim.x * kSkyWarps) {
        const int px = texel % CS_SKY_LUT_W, py = texel / CS_SKY_LUT_W;
        float u = (float)px / (float)CS_SKY_LUT_W, v = (float)py / (float)CS_SKY_LUT_H;
        float azimuth = 2.0f * PI * u;
        float l = v * 2.0f - 1.0f;
        float elev = l * l * signf(l) * PI * 0.5f;
        float ce = cosf(elev);
        V3 ray_dir = {ce * cosf(azimuth), ce * sinf(azimuth), sinf(elev)};
        V3 ray_origin = {0.0f, 0.0f, EARTH_RADIUS + 0.5f};
        float atmos_dist = ray_sphere_intersection(ray_origin, ray_dir, ATMOSPHERE_RADIUS);
        float ground_dist = ray_sphere_intersection(ray_origin, ray_dir, EARTH_RADIUS);
        float t_d = ground_dist < 0.0f ? atmos_dist : ground_dist;
        V3 neg_ray = {-ray_dir.x, -ray_dir.y, -ray_dir.z};
        float cos_theta = dot3(neg_ray, sun_dir);
        float molecular_phase = RAYLEIGH_PHASE_SCALE * (1.0f + cos_theta * cos_theta);
        float den = 1.0f + gg + 2.0f * g * cos_theta;
        float aerosol_phase = INV_4PI * (1.0f - gg) / (den * sqrtf(den));
        float dt = t_d / 30.0f;

        // step `lane` of the 30-step integral (lanes 30, 31 idle)
        V4 S_int = splat4(0.0f), step_T = splat4(1.0f);
        if (lane < 30) {
            float t = ((float)lane + 0.5f) * dt;
            V3 x_t = ray_origin + ray_dir * t;
            float d = length3<S>(x_t);
            V3 zenith_dir = {x_t.x / d, x_t.y / d, x_t.z / d};
            float altitude = d - EARTH_RADIUS;
            float normalized_altitude = altitude / ATMOSPHERE_THICKNESS;
            float sample_cos_theta = dot3(zenith_dir, sun_dir);
            Coeffs c = atmosphere_coefficients(altitude);
            V4 T_sun = transmittance_from_smem<BRUNETON>(lut, sample_cos_theta, normalized_altitude);
            // get_multiple_scattering (:144-164)
            float omega = 2.0f * PI * (1.0f - sqrtf(d * d - EARTH_RADIUS * EARTH_RADIUS) / d);
            V4 T_to_ground = transmittance_from_smem<BRUNETON>(lut, sample_cos_theta, 0.0f);
            V4 T_1_h = transmittance_from_smem<BRUNETON>(lut, 1.0f, normalized_altitude);
            V4 T_g2s = {T_1_0.x / T_1_h.x, T_1_0.y / T_1_h.y, T_1_0.z / T_1_h.z, T_1_0.w / T_1_h.w};
            V4 L_ground = (((splat4(albedo_over_pi) * (INV_4PI * omega)) * T_to_ground) * T_g2s) * sample_cos_theta;
            const V4 fit = {0.217f, 0.347f, 0.594f, 1.0f};
            V4 L_ms = (fit * 0.02f) * (1.0f / (1.0f + 5.0f * expf(-17.92f * sample_cos_theta)));
            V4 ms = L_ms + L_ground;
            V4 Ssrc = irradiance * (c.molecular_scattering * (T_sun * molecular_phase + ms) + c.aerosol_scattering * (T_sun * aerosol_phase + ms));
            V4 e = c.extinction;
            step_T = {expf(e.x * -dt), expf(e.y * -dt), expf(e.z * -dt), expf(e.w * -dt)};
            V4 num = Ssrc - Ssrc * step_T;
            S_int = {num.x / fmaxf(e.x, 1e-7f), num.y / fmaxf(e.y, 1e-7f), num.z / fmaxf(e.z, 1e-7f), num.w / fmaxf(e.w, 1e-7f)};
        }
        // sequential accumulation in the shader's order (:270-272)
        V4 L_in = splat4(0.0f), transmittance = splat4(1.0f);
#pragma unroll 6
        for (int i = 0; i < 30; ++i) {
            V4 si = {__shfl_sync(0xffffffffu, S_int.x, i), __shfl_sync(0xffffffffu, S_int.y, i), __shfl_sync(0xffffffffu, S_int.z, i), __shfl_sync(0xffffffffu, S_int.w, i)};
            V4 ti = {__shfl_sync(0xffffffffu, step_T.x, i), __shfl_sync(0xffffffffu, step_T.y, i), __shfl_sync(0xffffffffu, step_T.z, i), __shfl_sync(0xffffffffu, step_T.w, i)};
            L_in = L_in + transmittance * si;
            transmittance = transmittance * ti;
        }
        if (lane == 0) {
            // linear_srgb_from_spectral_samples (:207-217), mat4x3 column-major
            float r = 137.672389239975f * L_in.x + 32.549094028629234f * L_in.y + -38.91428392614275f * L_in.z + 8.572844237945445f * L_in.w;
            float gch = -8.632904716299537f * L_in.x + 91.29801417199785f * L_in.y + 34.31665471469816f * L_in.z + -11.103384660054624f * L_in.w;
            float b = -1.7181567391931372f * L_in.x + -12.005406444382531f * L_in.y + 29.89044807197628f * L_in.z + 117.47585277566478f * L_in.w;
            ushort4 o = {f2h(r), f2h(gch), f2h(b), f2h(1.0f)};
            reinterpret_cast<ushort4*>(out)[py * CS_SKY_LUT_W + px] = o;
        }
    }
}

}  // namespace

namespace cs {

void launch_transmittance_lut(uint16_t* out, int param, void* stream) {
    dim3 grid(CS_TRANSMITTANCE_W / 8, CS_TRANSMITTANCE_H / 8), block(8, 8);  // 32 x 8 groups (transmittance_lut.gd:77)
    transmittance_lut_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(out, param);
}

void launch_sky_lut(const uint16_t* tlut, int param, const float sun[3], uint16_t* out, void* stream) {
    // The reference dispatches 25 x 13 groups of 8 x 8 (sky_lut.gd:140); here: one persistent CTA per SM, warp per texel.
    static int sm_count[64] = {};  // per device: SM count, and "opt-in shared memory size configured"
    int dev = 0;
    cudaGetDevice(&dev);
    dev = dev < 0 || dev >= 64 ? 0 : dev;
    if (sm_count[dev] == 0) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(sky_lut_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkyLutBytes);
        cudaFuncSetAttribute(sky_lut_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkyLutBytes);
        sm_count[dev] = sms > 0 ? sms : 148;
    }
    if (param == CS_TLUT_BRUNETON2017) sky_lut_kernel<true><<<sm_count[dev], 32 * kSkyWarps, kSkyLutBytes, (cudaStream_t)stream>>>(tlut, sun[0], sun[1], sun[2], out);
    else sky_lut_kernel<false><<<sm_count[dev], 32 * kSkyWarps, kSkyLutBytes, (cudaStream_t)stream>>>(tlut, sun[0], sun[1], sun[2], out);
}

}  // namespace cs

// Reference-order cloud march: one thread per pixel, RGBA8 texels gathered from global memory
// (L1/L2-resident) and filtered in software with fp32 weights.  Function by function this
// follows clouds.glsl (citations inline); it is instantiated
//   * in clouds_strict.cu  (STRICT = true, compiled with --fmad=false): the bit-for-bit
//     operation order of the CPU oracle — the parity anchor, and
//   * by the prologue kernel shared with the fast path.
// The throughput kernel lives in clouds_fast.cu.
#pragma once
#include "cs_device.cuh"
#include "cs_internal.h"

namespace csd {

constexpr float g_radius = 6000000.0f;      // clouds.glsl:43
constexpr float sky_b_radius = 6001500.0f;  // clouds.glsl:44
constexpr float sky_t_radius = 6004000.0f;  // clouds.glsl:45
constexpr float CLOUDS_PI = 3.141592f;      // clouds.glsl:47

__device__ __constant__ float kRandomVectors[6][3] = {  // clouds.glsl:140
    {0.38051305f, 0.92453449f, -0.02111345f}, {-0.50625799f, -0.03590792f, -0.86163418f},
    {-0.32509218f, -0.94557439f, 0.01428793f}, {0.09026238f, -0.27376545f, 0.95755165f},
    {0.28128598f, 0.42443639f, -0.86065785f}, {-0.16852403f, 0.14748697f, 0.97460106f}};

__device__ __forceinline__ int wrapi(int i, int n) { return i & (n - 1); }  // n is a power of two

__device__ __forceinline__ float unorm8(uint32_t texel, int c) { return (float)((texel >> (8 * c)) & 0xffu) / 255.0f; }

// textureLod(sampler3D, s, integer lod), REPEAT, fp32 weights (same arithmetic as the oracle's sample_volume).
__device__ __forceinline__ V4 sample_volume_rgba8(const uint32_t* __restrict__ t, int n, V3 s) {
    float fn = (float)n;
    float ux = s.x * fn - 0.5f, uy = s.y * fn - 0.5f, uz = s.z * fn - 0.5f;
    float fx0 = floorf(ux), fy0 = floorf(uy), fz0 = floorf(uz);
    float fx = ux - fx0, fy = uy - fy0, fz = uz - fz0;
    int x0 = wrapi((int)fx0, n), y0 = wrapi((int)fy0, n), z0 = wrapi((int)fz0, n);
    int x1 = wrapi(x0 + 1, n), y1 = wrapi(y0 + 1, n), z1 = wrapi(z0 + 1, n);
    uint32_t t000 = __ldg(t + (z0 * n + y0) * n + x0), t100 = __ldg(t + (z0 * n + y0) * n + x1);
    uint32_t t010 = __ldg(t + (z0 * n + y1) * n + x0), t110 = __ldg(t + (z0 * n + y1) * n + x1);
    uint32_t t001 = __ldg(t + (z1 * n + y0) * n + x0), t101 = __ldg(t + (z1 * n + y0) * n + x1);
    uint32_t t011 = __ldg(t + (z1 * n + y1) * n + x0), t111 = __ldg(t + (z1 * n + y1) * n + x1);
    float o[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        float c00 = lerpf(unorm8(t000, c), unorm8(t100, c), fx);
        float c10 = lerpf(unorm8(t010, c), unorm8(t110, c), fx);
        float c01 = lerpf(unorm8(t001, c), unorm8(t101, c), fx);
        float c11 = lerpf(unorm8(t011, c), unorm8(t111, c), fx);
        o[c] = lerpf(lerpf(c00, c10, fy), lerpf(c01, c11, fy), fz);
    }
    return {o[0], o[1], o[2], o[3]};
}

// texture(weather_noise, uv): REPEAT bilinear at LOD 0 (power-of-two sizes only).
__device__ __forceinline__ V3 sample_weather_rgba8(const uint32_t* __restrict__ t, int w, int h, float su, float sv) {
    float ux = su * (float)w - 0.5f, uy = sv * (float)h - 0.5f;
    float fx0 = floorf(ux), fy0 = floorf(uy);
    float fx = ux - fx0, fy = uy - fy0;
    int x0 = wrapi((int)fx0, w), y0 = wrapi((int)fy0, h);
    int x1 = wrapi(x0 + 1, w), y1 = wrapi(y0 + 1, h);
    uint32_t t00 = __ldg(t + y0 * w + x0), t10 = __ldg(t + y0 * w + x1);
    uint32_t t01 = __ldg(t + y1 * w + x0), t11 = __ldg(t + y1 * w + x1);
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; c++)
        o[c] = lerpf(lerpf(unorm8(t00, c), unorm8(t10, c), fx), lerpf(unorm8(t01, c), unorm8(t11, c), fx), fy);
    return {o[0], o[1], o[2]};
}

// clouds.glsl:49-57
template <bool STRICT>
__device__ __forceinline__ V3 getValFromSkyLUT(const uint16_t* __restrict__ lut, V3 rayDir) {
    float phi = atan2f(rayDir.z, rayDir.x);
    float theta = asinf(rayDir.y);
    float u = (phi / CLOUDS_PI * 0.5f + 0.5f);
    float v = sqrtf(fabsf(theta) / (CLOUDS_PI * 0.5f)) * signf(theta) * 0.5f + 0.5f;
    V4 t = sample_lut_half4(lut, CS_SKY_LUT_W, CS_SKY_LUT_H, u, v);
    return {t.x, t.y, t.z};
}
// clouds.glsl:67-69
template <bool STRICT>
__device__ __forceinline__ float remap(float v, float omin, float omax, float nmin, float nmax) {
    return nmin + (fdiv<STRICT>(v - omin, omax - omin) * (nmax - nmin));
}
// clouds.glsl:72-75
template <bool STRICT>
__device__ __forceinline__ float henyey_greenstein(float cos_theta, float g) {
    const float k = 0.0795774715459f;
    return fdiv<STRICT>(k * (1.0f - g * g), fpow<STRICT>(1.0f + g * g - 2.0f * g * cos_theta, 1.5f));
}
// clouds.glsl:77-80
template <bool STRICT>
__device__ __forceinline__ float GetHeightFractionForPoint(float inPosition) {
    float hf = fdiv<STRICT>(inPosition - sky_b_radius, sky_t_radius - sky_b_radius);
    return clampf(hf, 0.0f, 1.0f);
}
// clouds.glsl:82-95
template <bool STRICT>
__device__ __forceinline__ float densityHeightGradient(float heightFrac, float cloudType) {
    float stratus = 1.0f - clampf(cloudType * 2.0f, 0.0f, 1.0f);
    float stratocumulus = 1.0f - fabsf(cloudType - 0.5f) * 2.0f;
    float cumulus = clampf(cloudType - 0.5f, 0.0f, 1.0f) * 2.0f;
    float gx = 0.02f * stratus + 0.02f * stratocumulus + 0.01f * cumulus;
    float gy = 0.05f * stratus + 0.2f * stratocumulus + 0.0625f * cumulus;
    float gz = 0.09f * stratus + 0.48f * stratocumulus + 0.78f * cumulus;
    float gw = 0.11f * stratus + 0.625f * stratocumulus + 1.0f * cumulus;
    return smoothstepf<STRICT>(gx, gy, heightFrac) - smoothstepf<STRICT>(gz, gw, heightFrac);
}
// clouds.glsl:97-105
template <bool STRICT>
__device__ __forceinline__ float intersectSphere(V3 pos, V3 dir, float r) {
    float a = dot3(dir, dir);
    float b = 2.0f * dot3(dir, pos);
    float c = dot3(pos, pos) - (r * r);
    float d = fsqrt<STRICT>((b * b) - 4.0f * a * c);
    float p = -b - d;
    float p2 = -b + d;
    return fdiv<STRICT>(fmaxf(p, p2), 2.0f * a);
}
// clouds.glsl:239-256 followed by the .xzy swizzle of :262
template <bool STRICT>
__device__ __forceinline__ V3 pixel_direction(int px, int py, float tw, float th) {
    float ex = fdiv<STRICT>((float)px, tw), ey = fdiv<STRICT>((float)py, th);
    V3 n;
    n.x = (ex - ey);
    n.y = (ex + ey) - 1.0f;
    n.z = 1.0f - fabsf(n.x) - fabsf(n.y);
    if (!(n.z >= 0.0f)) {
        float sx = n.x >= 0.0f ? 1.0f : -1.0f, sy = n.y >= 0.0f ? 1.0f : -1.0f;
        float wx = (1.0f - fabsf(n.y)) * sx, wy = (1.0f - fabsf(n.x)) * sy;
        n.x = wx; n.y = wy;
    }
    n = normalize3<STRICT>(n);
    return {n.x, n.z, n.y};
}

// The pixel-independent part of march()'s prologue (clouds.glsl:149-167).
template <bool STRICT>
__device__ __forceinline__ void compute_frame_consts(const cs::CloudLaunch& L, cs::FrameConsts& fc) {
    const cs_cloud_params& P = L.P;
    V3 LD = {P.light_direction[0], P.light_direction[1], P.light_direction[2]};
    V3 ldir = normalize3<STRICT>(LD);
    V3 sun = ((getValFromSkyLUT<STRICT>(L.sky_lut, LD) * 0.1f) * P.light_energy) * v3(P.light_color[0], P.light_color[1], P.light_color[2]);
    V3 amb = getValFromSkyLUT<STRICT>(L.sky_lut, normalize3<STRICT>(v3(1.0f, 1.0f, 0.0f))) * 0.05f;
    float la = length3<STRICT>(amb);
    amb = mix3(amb, v3(la, la, la), 0.5f);
    V3 gnd = (getValFromSkyLUT<STRICT>(L.sky_lut, normalize3<STRICT>(v3(1.0f, -1.0f, 0.0f))) * 5.0f) * 0.05f;
    float lg = length3<STRICT>(gnd);
    gnd = mix3(gnd, v3(P.ground_color[0] * lg, P.ground_color[1] * lg, P.ground_color[2] * lg), 0.5f);
    fc.ldir[0] = ldir.x; fc.ldir[1] = ldir.y; fc.ldir[2] = ldir.z;
    fc.atmosphere_sun[0] = sun.x; fc.atmosphere_sun[1] = sun.y; fc.atmosphere_sun[2] = sun.z;
    fc.atmosphere_ambient[0] = amb.x; fc.atmosphere_ambient[1] = amb.y; fc.atmosphere_ambient[2] = amb.z;
    fc.atmosphere_ground[0] = gnd.x; fc.atmosphere_ground[1] = gnd.y; fc.atmosphere_ground[2] = gnd.z;
    fc.hg_g2 = (0.4f - 1.4f * ldir.y);
    fc.pad[0] = fc.pad[1] = fc.pad[2] = 0.0f;
}

struct Tally { unsigned int steps, lit, evals; };

// clouds.glsl:109-137
template <bool STRICT, bool COUNT>
__device__ __forceinline__ float density_ref(const cs::CloudLaunch& L, V3 p, V3 weather, int mip, Tally& tl) {
    if constexpr (COUNT) tl.evals++;
    const cs_cloud_params& P = L.P;
    float height_fraction = GetHeightFractionForPoint<STRICT>(length3<STRICT>(p));
    p.x += 20.0f * P.cloud_pos[0] * 0.6f;
    p.z += 20.0f * P.cloud_pos[1] * 0.6f;
    int ll = min(max(mip - 2, 0), L.large_levels - 1);
    V4 n = sample_volume_rgba8(L.large[ll], L.large_n >> ll, v3(p.x * 0.00008f, p.y * 0.00008f, p.z * 0.00008f));
    float fbm = n.y * 0.625f + n.z * 0.25f + n.w * 0.125f;
    float g = densityHeightGradient<STRICT>(height_fraction, weather.x);
    float base_cloud = remap<STRICT>(n.x, -(1.0f - fbm), 1.0f, 0.0f, 1.0f);
    float weather_coverage = P.cloud_coverage * weather.z;
    base_cloud = remap<STRICT>(base_cloud * g, 1.0f - (weather_coverage), 1.0f, 0.0f, 1.0f);
    base_cloud *= weather_coverage;
    p.x -= P.detailed_pos[0] * 40.0f;
    p.z -= P.detailed_pos[1] * 40.0f;
    p.y -= P.time * 40.0f;
    int sl = min(max(mip, 0), L.small_levels - 1);
    V4 hn = sample_volume_rgba8(L.small[sl], L.small_n >> sl, v3(p.x * 0.001f, p.y * 0.001f, p.z * 0.001f));
    float hfbm = hn.x * 0.625f + hn.y * 0.25f + hn.z * 0.125f;
    hfbm = mixf(hfbm, 1.0f - hfbm, clampf(height_fraction * 4.0f, 0.0f, 1.0f));
    base_cloud = remap<STRICT>(base_cloud, hfbm * 0.4f * height_fraction, 1.0f, 0.0f, 1.0f);
    return fpow<STRICT>(clampf(base_cloud, 0.0f, 1.0f), (1.0f - height_fraction) * 0.8f + 0.5f);
}

// clouds.glsl:139-237 for one pixel: sky() + march()
template <bool STRICT, bool COUNT>
__device__ __forceinline__ V4 sky_pixel_ref(const cs::CloudLaunch& L, const cs::FrameConsts& fc, V3 dir, Tally& tl) {
    const cs_cloud_params& P = L.P;
    V3 camPos = {0.0f, g_radius, 0.0f};
    V3 start = camPos + dir * intersectSphere<STRICT>(camPos, dir, sky_b_radius);
    V3 end = camPos + dir * intersectSphere<STRICT>(camPos, dir, sky_t_radius);
    float shelldist = length3<STRICT>(end - start);
    int n_steps = L.primary_steps;
    if (L.budget_len > 0.0f) n_steps = min(L.primary_steps, max(L.budget_min, (int)ceilf(fdiv<STRICT>(shelldist, L.budget_len))));  // cs_set_step_budget
    float steps = (float)n_steps;
    V3 ds = dir * shelldist;
    V3 raystep = {fdiv<STRICT>(ds.x, steps), fdiv<STRICT>(ds.y, steps), fdiv<STRICT>(ds.z, steps)};

    // march() (:139-215)
    float ss = length3<STRICT>(raystep);
    V3 d = normalize3<STRICT>(raystep);
    V3 h = {start.x * 10.0f * 0.3183099f + 0.1f, start.y * 10.0f * 0.3183099f + 0.1f, start.z * 10.0f * 0.3183099f + 0.1f};
    h = {h.x - floorf(h.x), h.y - floorf(h.y), h.z - floorf(h.z)};
    h = h * 17.0f;
    float hh = h.x * h.y * h.z * (h.x + h.y + h.z);
    hh = hh - floorf(hh);  // hash(pos * 10.0) (:60-64); identically 0 in fp32 here
    V3 p = start + (d * hh) * ss;

    const float lss = (sky_t_radius - sky_b_radius) / 64.0f;
    V3 ldir = {fc.ldir[0], fc.ldir[1], fc.ldir[2]};
    float T = 1.0f, alpha = 0.0f;
    V3 Lacc = {0.0f, 0.0f, 0.0f};
    float costheta = dot3(ldir, d);
    float phase = fmaxf(fmaxf(henyey_greenstein<STRICT>(costheta, 0.6f), henyey_greenstein<STRICT>(costheta, fc.hg_g2)),
                        henyey_greenstein<STRICT>(costheta, -0.2f));
    V3 atmosphere_sun = {fc.atmosphere_sun[0], fc.atmosphere_sun[1], fc.atmosphere_sun[2]};
    V3 atmosphere_ambient = {fc.atmosphere_ambient[0], fc.atmosphere_ambient[1], fc.atmosphere_ambient[2]};
    V3 atmosphere_ground = {fc.atmosphere_ground[0], fc.atmosphere_ground[1], fc.atmosphere_ground[2]};
    const float weather_scale = 0.00006f;
    const float wpx = P.weather_pos[0], wpy = P.weather_pos[1];

    for (int i = 0; i < n_steps; i++) {
        if constexpr (COUNT) tl.steps++;
        p = p + d * ss;
        V3 weather_sample = sample_weather_rgba8(L.weather, L.weather_w, L.weather_h, p.x * weather_scale + 0.5f + wpx, p.z * weather_scale + 0.5f + wpy);
        float height_fraction = GetHeightFractionForPoint<STRICT>(length3<STRICT>(p));
        float t = density_ref<STRICT, COUNT>(L, p, weather_sample, 0, tl);
        float dt = fexp<STRICT>(-P.density * t * ss);
        if (t > 0.0f) {
            if constexpr (COUNT) tl.lit++;
            V3 lp = p;
            float cd = 0.0f;
            for (int j = 0; j < L.cone_samples; j++) {
                int r = j % 6;
                V3 rv = {kRandomVectors[r][0], kRandomVectors[r][1], kRandomVectors[r][2]};
                lp = lp + (ldir + rv * (float)j) * lss;
                V3 lweather = sample_weather_rgba8(L.weather, L.weather_w, L.weather_h, lp.x * weather_scale + 0.5f + wpx, lp.z * weather_scale + 0.5f + wpy);
                cd += density_ref<STRICT, COUNT>(L, lp, lweather, j, tl);
            }
            lp = p + (ldir * 18.0f) * lss;
            float lheight_fraction = GetHeightFractionForPoint<STRICT>(length3<STRICT>(lp));
            V3 lweather = sample_weather_rgba8(L.weather, L.weather_w, L.weather_h, lp.x * weather_scale + 0.5f, lp.z * weather_scale + 0.5f);
            cd += fpow<STRICT>(density_ref<STRICT, COUNT>(L, lp, lweather, 5, tl), (1.0f - lheight_fraction) * 0.8f + 0.5f);

            float beers = fexp<STRICT>(-P.density * cd * lss * 3.0f);
            float powder = 1.0f - fexp<STRICT>(-P.density * cd * lss * 3.0f * 2.0f);
            float beers_total = 2.0f * beers * powder;
            V3 ambient = mix3(atmosphere_ground, atmosphere_ambient, smoothstepf<STRICT>(0.0f, 1.0f, height_fraction));
            alpha += (1.0f - dt) * (1.0f - alpha);
            V3 radiance = (ambient + (atmosphere_sun * beers_total) * phase) * t;
            V3 num = (radiance - radiance * dt) * T;
            float den = fmaxf(0.0000001f, t);
            Lacc = Lacc + v3(fdiv<STRICT>(num.x, den), fdiv<STRICT>(num.y, den), fdiv<STRICT>(num.z, den));
            T *= dt;
        }
    }
    alpha = clampf(alpha, 0.0f, 1.0f);
    return {Lacc.x, Lacc.y, Lacc.z, alpha};
}

}  // namespace csd

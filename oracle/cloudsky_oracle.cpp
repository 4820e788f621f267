/*
 * cloudsky_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar fp32 CPU restatement of the reference's three compute shaders and of the host-side
 * parameter logic, used ONLY as the parity checker (tests/, __graft_entry__.smoke(),
 * bench.py's cpu_baseline / --impl reference legs).  Nothing in the product path may link,
 * import or call this file.
 *
 * PARITY PINNED (round 2): the reference ships no tests, golden vectors or fixtures, and nothing here can run Godot /
 * Vulkan, but its three GLSL compute shaders compile UNMODIFIED as C++ behind oracle/glsl_compat.h (oracle/build_ref.sh
 * -> oracle/_ref/libcloudsky_ref.so).  This restatement is bit-identical to that compiled reference: both LUTs on every
 * texel and the cloud image on every pixel, 5 parameter sets (tests/test_reference_pin.py), and to the vectors the
 * compiled reference produced (tests/golden/ref_golden.npz).  What stays unpinned is third-party arithmetic outside the
 * reference tree (Godot Engine "4.2 or later", un-vendored): BC7 compression of the inputs, Godot's mip generator, a GPU
 * driver's pow/exp/atan/asin and the texture unit's 8-bit filter weights; the oracle (like _ref) samples the raw 8-bit
 * texels with fp32 weights and 2x2x2 box-filter mips re-quantised to 8 bits.
 *
 * Rules of the restatement: fp32 throughout, GLSL operation order, built with
 * -O2 -ffp-contract=off (no FMA contraction), round-to-nearest-even fp16 at every RGBA16F
 * store (both LUTs and the output image).
 *
 * Citations are file:line relative to the reference root.
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../include/cloudsky.h"

namespace {

// ---------------------------------------------------------------------------------------------
// small vector helpers (GLSL semantics, left-to-right evaluation)
// ---------------------------------------------------------------------------------------------
struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 operator-(V4 a, V4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline V4 operator*(V4 a, V4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline V4 operator/(V4 a, V4 b) { return {a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w}; }
inline V4 operator*(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline V4 operator+(V4 a, float s) { return {a.x + s, a.y + s, a.z + s, a.w + s}; }
inline V4 splat4(float s) { return {s, s, s, s}; }
inline float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length3(V3 a) { return sqrtf(dot3(a, a)); }
inline V3 normalize3(V3 a) { return a / length3(a); }
inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }  // GLSL mix
inline V3 mix3(V3 a, V3 b, float t) { return {mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)}; }
inline float fractf(float x) { return x - floorf(x); }
inline float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float smoothstepf(float e0, float e1, float x) {
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline V4 exp4(V4 a) { return {expf(a.x), expf(a.y), expf(a.z), expf(a.w)}; }

// ---------------------------------------------------------------------------------------------
// fp16 conversion (RGBA16F stores: clouds.glsl:8, sky-lut.glsl:8, transmittance-lut.glsl:8)
// ---------------------------------------------------------------------------------------------
uint16_t f32_to_f16(float f) {  // round-to-nearest-even, IEEE binary16
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t mant = x & 0x007fffffu;
    int32_t exp = (int32_t)((x >> 23) & 0xff);
    if (exp == 0xff) return (uint16_t)(sign | 0x7c00u | (mant ? 0x200u | (mant >> 13) : 0u));
    int32_t e = exp - 127 + 15;
    if (e >= 0x1f) return (uint16_t)(sign | 0x7c00u);  // overflow -> inf
    if (e <= 0) {                                      // subnormal half or zero
        if (e < -10) return (uint16_t)sign;
        mant |= 0x00800000u;
        uint32_t shift = (uint32_t)(14 - e);
        uint32_t half = mant >> shift;
        uint32_t rem = mant & ((1u << shift) - 1u);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half & 1u))) half++;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)e << 10) | (mant >> 13);
    uint32_t rem = mant & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) half++;  // may carry into exponent (ok)
    return (uint16_t)(sign | half);
}
float f16_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f;
    uint32_t mant = h & 0x3ffu;
    uint32_t x;
    if (exp == 0) {
        if (mant == 0) x = sign;
        else {
            int e = -1;
            do { mant <<= 1; e++; } while (!(mant & 0x400u));
            mant &= 0x3ffu;
            x = sign | ((uint32_t)(127 - 15 - e) << 23) | (mant << 13);
        }
    } else if (exp == 0x1f) x = sign | 0x7f800000u | (mant << 13);
    else x = sign | ((exp - 15 + 127) << 23) | (mant << 13);
    float f;
    memcpy(&f, &x, 4);
    return f;
}

// ---------------------------------------------------------------------------------------------
// software samplers
//   noise sampler: REPEAT x3, linear min/mag/mip (cloud_sky.gd:301-307)
//   LUT samplers : CLAMP_TO_EDGE, linear (cloud_sky.gd:383-388, sky_lut.gd:62-68)
// texel centres at (i+0.5)/N; weights in fp32; lerp(a,b,f) = a + (b-a)*f, x then y then z.
// ---------------------------------------------------------------------------------------------
inline float lerpf(float a, float b, float f) { return a + (b - a) * f; }
inline int wrapi(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

struct Volume {          // RGBA8 mip chain of an n^3 volume, x fastest
    int n = 0;           // level-0 edge
    int levels = 0;
    std::vector<std::vector<uint8_t>> mip;  // each (n>>l)^3 * 4
};

// 2x2x2 box filter re-quantised to 8 bits with round-half-up ((sum + 4) >> 3).  This is the
// oracle's DEFINITION of "mipmaps/generate=true" (perlworlnoise.tga.import:24); Godot's real
// generator is outside the reference tree (see header).
void build_mips(Volume& v) {
    int n = v.n;
    while (n > 1) {
        int m = n / 2;
        const std::vector<uint8_t>& src = v.mip.back();
        std::vector<uint8_t> dst((size_t)m * m * m * 4);
        for (int z = 0; z < m; z++)
            for (int y = 0; y < m; y++)
                for (int x = 0; x < m; x++)
                    for (int c = 0; c < 4; c++) {
                        int s = 0;
                        for (int dz = 0; dz < 2; dz++)
                            for (int dy = 0; dy < 2; dy++)
                                for (int dx = 0; dx < 2; dx++)
                                    s += src[((((size_t)(2 * z + dz) * n) + (2 * y + dy)) * n + (2 * x + dx)) * 4 + c];
                        dst[(((size_t)z * m + y) * m + x) * 4 + c] = (uint8_t)((s + 4) >> 3);
                    }
        v.mip.push_back(std::move(dst));
        n = m;
    }
    v.levels = (int)v.mip.size();
}

// textureLod(sampler3D, s, lod) with an integer-valued lod (all LODs in clouds.glsl are
// integers after the clamp at 0: clouds.glsl:117 "mip - 2.0", :132 "mip").
V4 sample_volume(const Volume& v, V3 s, float lod) {
    int l = (int)fmaxf(lod, 0.0f);
    if (l > v.levels - 1) l = v.levels - 1;
    int n = v.n >> l;
    const uint8_t* t = v.mip[l].data();
    float ux = s.x * (float)n - 0.5f, uy = s.y * (float)n - 0.5f, uz = s.z * (float)n - 0.5f;
    float fx0 = floorf(ux), fy0 = floorf(uy), fz0 = floorf(uz);
    float fx = ux - fx0, fy = uy - fy0, fz = uz - fz0;
    int x0 = wrapi((int)fx0, n), y0 = wrapi((int)fy0, n), z0 = wrapi((int)fz0, n);
    int x1 = x0 + 1 == n ? 0 : x0 + 1, y1 = y0 + 1 == n ? 0 : y0 + 1, z1 = z0 + 1 == n ? 0 : z0 + 1;
    float out[4];
    for (int c = 0; c < 4; c++) {
        auto T = [&](int x, int y, int z) {
            return (float)t[((((size_t)z * n) + y) * n + x) * 4 + c] / 255.0f;  // UNORM8
        };
        float c00 = lerpf(T(x0, y0, z0), T(x1, y0, z0), fx);
        float c10 = lerpf(T(x0, y1, z0), T(x1, y1, z0), fx);
        float c01 = lerpf(T(x0, y0, z1), T(x1, y0, z1), fx);
        float c11 = lerpf(T(x0, y1, z1), T(x1, y1, z1), fx);
        float c0 = lerpf(c00, c10, fy);
        float c1 = lerpf(c01, c11, fy);
        out[c] = lerpf(c0, c1, fz);
    }
    return {out[0], out[1], out[2], out[3]};
}

struct Image8 { int w = 0, h = 0; std::vector<uint8_t> px; };  // RGBA8

// texture(sampler2D weather_noise, uv) in a compute shader = LOD 0; REPEAT; (clouds.glsl:174)
V3 sample_weather(const Image8& im, V2 s) {
    float ux = s.x * (float)im.w - 0.5f, uy = s.y * (float)im.h - 0.5f;
    float fx0 = floorf(ux), fy0 = floorf(uy);
    float fx = ux - fx0, fy = uy - fy0;
    int x0 = wrapi((int)fx0, im.w), y0 = wrapi((int)fy0, im.h);
    int x1 = x0 + 1 == im.w ? 0 : x0 + 1, y1 = y0 + 1 == im.h ? 0 : y0 + 1;
    float out[3];
    for (int c = 0; c < 3; c++) {
        auto T = [&](int x, int y) { return (float)im.px[((size_t)y * im.w + x) * 4 + c] / 255.0f; };
        float a = lerpf(T(x0, y0), T(x1, y0), fx);
        float b = lerpf(T(x0, y1), T(x1, y1), fx);
        out[c] = lerpf(a, b, fy);
    }
    return {out[0], out[1], out[2]};
}

// RGBA16F LUT, CLAMP_TO_EDGE bilinear on normalised coordinates.
V4 sample_lut(const uint16_t* lut, int w, int h, float su, float sv) {
    float ux = su * (float)w - 0.5f, uy = sv * (float)h - 0.5f;
    float fx0 = floorf(ux), fy0 = floorf(uy);
    float fx = ux - fx0, fy = uy - fy0;
    int x0 = (int)fx0, y0 = (int)fy0;
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = std::min(std::max(x0, 0), w - 1); x1 = std::min(std::max(x1, 0), w - 1);
    y0 = std::min(std::max(y0, 0), h - 1); y1 = std::min(std::max(y1, 0), h - 1);
    float out[4];
    for (int c = 0; c < 4; c++) {
        auto T = [&](int x, int y) { return f16_to_f32(lut[((size_t)y * w + x) * 4 + c]); };
        float a = lerpf(T(x0, y0), T(x1, y0), fx);
        float b = lerpf(T(x0, y1), T(x1, y1), fx);
        out[c] = lerpf(a, b, fy);
    }
    return {out[0], out[1], out[2], out[3]};
}

// ---------------------------------------------------------------------------------------------
// atmosphere model shared by both LUT shaders
//   (transmittance-lut.glsl:45-145 == sky-lut.glsl:56-202)
// ---------------------------------------------------------------------------------------------
const float EARTH_RADIUS = 6371.0f;                                        // :50
const float ATMOSPHERE_THICKNESS = 100.0f;                                 // :51
const float ATMOSPHERE_RADIUS = EARTH_RADIUS + ATMOSPHERE_THICKNESS;       // :52
const V4 sun_spectral_irradiance = {1.679f, 1.828f, 1.986f, 1.307f};       // :56
const V4 molecular_scattering_coefficient_base = {6.605e-3f, 1.067e-2f, 1.842e-2f, 3.156e-2f};  // :60
const V4 ozone_absorption_cross_section = {3.472e-21f * 1e-4f, 3.914e-21f * 1e-4f, 1.349e-21f * 1e-4f, 11.03e-23f * 1e-4f};  // :64
const float ozone_mean_monthly_dobson = 350.0f;                            // :67
const V4 aerosol_absorption_cross_section = {2.8722e-24f, 4.6168e-24f, 7.9706e-24f, 1.3578e-23f};  // :74
const V4 aerosol_scattering_cross_section = {1.5908e-22f, 1.7711e-22f, 2.0942e-22f, 2.4033e-22f};  // :75
const float aerosol_base_density = 1.3681e20f;                             // :76
const float aerosol_background_density = 2e6f;                             // :77
const float aerosol_height_scale = 0.73f;                                  // :78
const float aerosol_background_divided_by_base_density = aerosol_background_density / aerosol_base_density;  // :80

// transmittance-lut.glsl:89-98 / sky-lut.glsl:100-109
float ray_sphere_intersection(V3 ro, V3 rd, float radius) {
    float b = dot3(ro, rd);
    float c = dot3(ro, ro) - radius * radius;
    if (c > 0.0f && b > 0.0f) return -1.0f;
    float d = b * b - c;
    if (d < 0.0f) return -1.0f;
    if (d > b * b) return (-b + sqrtf(d));
    return (-b - sqrtf(d));
}
// transmittance-lut.glsl:104-107
V4 get_molecular_scattering_coefficient(float h) {
    return molecular_scattering_coefficient_base * expf(-0.07771971f * powf(h, 1.16364243f));
}
// transmittance-lut.glsl:113-119
V4 get_molecular_absorption_coefficient(float h) {
    h += 1e-4f;
    float t = logf(h) - 3.22261f;
    float density = 3.78547397e20f * (1.0f / h) * expf(-t * t * 5.55555555f);
    return ozone_absorption_cross_section * ozone_mean_monthly_dobson * density;
}
// transmittance-lut.glsl:121-125
float get_aerosol_density(float h) {
    return aerosol_base_density * (expf(-h / aerosol_height_scale) + aerosol_background_divided_by_base_density);
}
// transmittance-lut.glsl:131-145
void get_atmosphere_collision_coefficients(float h, V4& aerosol_absorption, V4& aerosol_scattering,
                                           V4& molecular_absorption, V4& molecular_scattering, V4& extinction) {
    h = fmaxf(h, 0.0f);
    float aerosol_density = get_aerosol_density(h);
    aerosol_absorption = aerosol_absorption_cross_section * aerosol_density;
    aerosol_scattering = aerosol_scattering_cross_section * aerosol_density;
    molecular_absorption = get_molecular_absorption_coefficient(h);
    molecular_scattering = get_molecular_scattering_coefficient(h);
    extinction = aerosol_absorption + aerosol_scattering + molecular_absorption + molecular_scattering;
}

// CS_TLUT_BRUNETON2017 (extension: README.md:29 TODO; include/cloudsky.h): E. Bruneton 2017, "Precomputed Atmospheric
// Scattering: a New Implementation" — transmittance texture coordinates from (r, mu) through the distance to the top
// boundary, texel centres on the ends of the unit range, written with cancellation-free differences of squares.
const float BRUNETON_H2 = ATMOSPHERE_THICKNESS * (ATMOSPHERE_RADIUS + EARTH_RADIUS);  // Rt^2 - Rg^2
const float SUN_ANGULAR_RADIUS = 0.53f * 3.14159265358979f / 180.0f * 0.5f;           // half of clouds.gdshader:49's disc
struct BrunetonRay { float altitude, r, mu, d; };
BrunetonRay bruneton_texel_ray(int px, int py) {
    float H = sqrtf(BRUNETON_H2);
    float x_mu = (float)px / (float)(CS_TRANSMITTANCE_W - 1);
    float x_r = (float)py / (float)(CS_TRANSMITTANCE_H - 1);
    float rho = H * x_r;
    BrunetonRay o;
    o.r = sqrtf(rho * rho + EARTH_RADIUS * EARTH_RADIUS);
    o.altitude = (rho * rho) / (o.r + EARTH_RADIUS);
    float d_min = ATMOSPHERE_THICKNESS - o.altitude, d_max = rho + H;
    o.d = d_min + x_mu * (d_max - d_min);
    o.mu = o.d == 0.0f ? 1.0f : (BRUNETON_H2 - rho * rho - o.d * o.d) / (2.0f * o.r * o.d);
    o.mu = clampf(o.mu, -1.0f, 1.0f);
    return o;
}
// -> (u, v, visible fraction of the sun's disc)
V3 bruneton_lookup_coords(float normalized_altitude, float mu) {
    float H = sqrtf(BRUNETON_H2);
    float h = clampf(normalized_altitude, 0.0f, 1.0f) * ATMOSPHERE_THICKNESS;
    float r = EARTH_RADIUS + h;
    float rho = sqrtf(h * (2.0f * EARTH_RADIUS + h));
    float d_min = ATMOSPHERE_THICKNESS - h, d_max = rho + H;
    float rmu = r * mu;
    float discriminant = rmu * rmu + d_min * (ATMOSPHERE_RADIUS + r);  // r^2 (mu^2 - 1) + Rt^2
    float d = fmaxf(sqrtf(fmaxf(discriminant, 0.0f)) - rmu, 0.0f);
    float x_mu = clampf((d - d_min) / (d_max - d_min), 0.0f, 1.0f), x_r = rho / H;
    float u = 0.5f / (float)CS_TRANSMITTANCE_W + x_mu * (1.0f - 1.0f / (float)CS_TRANSMITTANCE_W);
    float v = 0.5f / (float)CS_TRANSMITTANCE_H + x_r * (1.0f - 1.0f / (float)CS_TRANSMITTANCE_H);
    float sin_horizon = EARTH_RADIUS / r, cos_horizon = -(rho / r);
    float visible = smoothstepf(-sin_horizon * SUN_ANGULAR_RADIUS, sin_horizon * SUN_ANGULAR_RADIUS, mu - cos_horizon);
    return {u, v, visible};
}

// transmittance-lut.glsl:157-196, one texel
void transmittance_texel(int px, int py, uint16_t* out4, int param) {
    const int TRANSMITTANCE_STEPS = 40;  // :45
    float u = (float)px / (float)CS_TRANSMITTANCE_W, v = (float)py / (float)CS_TRANSMITTANCE_H;  // :162
    float sun_cos_theta = u * 2.0f - 1.0f;
    float distance_to_earth_center = mixf(EARTH_RADIUS, ATMOSPHERE_RADIUS, v);
    BrunetonRay br{};
    if (param == CS_TLUT_BRUNETON2017) { br = bruneton_texel_ray(px, py); sun_cos_theta = br.mu; distance_to_earth_center = br.r; }
    V3 sun_dir = {-sqrtf(1.0f - sun_cos_theta * sun_cos_theta), 0.0f, sun_cos_theta};
    V3 ray_origin = {0.0f, 0.0f, distance_to_earth_center};
    float t_d = param == CS_TLUT_BRUNETON2017 ? br.d : ray_sphere_intersection(ray_origin, sun_dir, ATMOSPHERE_RADIUS);
    float dt = t_d / (float)TRANSMITTANCE_STEPS;
    V4 result = splat4(0.0f);
    for (int i = 0; i < TRANSMITTANCE_STEPS; ++i) {
        float t = ((float)i + 0.5f) * dt;
        V3 x_t = ray_origin + sun_dir * t;
        float altitude = length3(x_t) - EARTH_RADIUS;
        V4 aa, as, ma, ms, ext;
        get_atmosphere_collision_coefficients(altitude, aa, as, ma, ms, ext);
        result = result + ext * dt;
    }
    V4 tr = exp4(result * -1.0f);
    out4[0] = f32_to_f16(tr.x); out4[1] = f32_to_f16(tr.y); out4[2] = f32_to_f16(tr.z); out4[3] = f32_to_f16(tr.w);
}

// ---- sky-lut.glsl ---------------------------------------------------------------------------
const float SKY_PI = 3.14159265358979323846f;   // sky-lut.glsl:44
const float INV_PI = 0.31830988618379067154f;   // :45
const float INV_4PI = 0.25f * INV_PI;           // :46
const float PHASE_ISOTROPIC = INV_4PI;          // :47
const float RAYLEIGH_PHASE_SCALE = (3.0f / 16.0f) * INV_PI;  // :48
const float sky_g = 0.8f;                       // :49
const float sky_gg = sky_g * sky_g;             // :50
const float EYE_ALTITUDE = 0.5f;                // :61
const float EYE_DISTANCE_TO_EARTH_CENTER = EARTH_RADIUS + EYE_ALTITUDE;  // :62
const V4 GROUND_ALBEDO = {0.3f, 0.3f, 0.3f, 0.3f};  // :63

float molecular_phase_function(float c) { return RAYLEIGH_PHASE_SCALE * (1.0f + c * c); }  // :114-117
float aerosol_phase_function(float c) {                                                    // :122-126
    float den = 1.0f + sky_gg + 2.0f * sky_g * c;
    return INV_4PI * (1.0f - sky_gg) / (den * sqrtf(den));
}
V4 transmittance_from_lut(const uint16_t* lut, float cos_theta, float normalized_altitude, int param) {  // :137-142
    if (param == CS_TLUT_BRUNETON2017) {
        V3 c = bruneton_lookup_coords(normalized_altitude, cos_theta);
        return sample_lut(lut, CS_TRANSMITTANCE_W, CS_TRANSMITTANCE_H, c.x, c.y) * c.z;
    }
    float u = clampf(cos_theta * 0.5f + 0.5f, 0.0f, 1.0f);
    float v = clampf(normalized_altitude, 0.0f, 1.0f);
    return sample_lut(lut, CS_TRANSMITTANCE_W, CS_TRANSMITTANCE_H, u, v);
}
// GROUND_ALBEDO / PI (sky-lut.glsl:157) is a true fp32 division in GLSL.
V4 ground_albedo_over_pi() { return {0.3f / SKY_PI, 0.3f / SKY_PI, 0.3f / SKY_PI, 0.3f / SKY_PI}; }

// sky-lut.glsl:219-276
V4 compute_inscattering(const uint16_t* tlut, int param, V3 sun_direction_param, V3 ray_origin, V3 ray_dir, float t_d) {
    const int IN_SCATTERING_STEPS = 30;  // :53
    // :221-223  sun_dir = params.sun_direction.xzy; x = -x; y = -y
    V3 sun_dir = {-sun_direction_param.x, -sun_direction_param.z, sun_direction_param.y};
    V3 neg_ray = {-ray_dir.x, -ray_dir.y, -ray_dir.z};
    float cos_theta = dot3(neg_ray, sun_dir);
    float molecular_phase = molecular_phase_function(cos_theta);
    float aerosol_phase = aerosol_phase_function(cos_theta);
    float dt = t_d / (float)IN_SCATTERING_STEPS;
    V4 L_inscattering = splat4(0.0f);
    V4 transmittance = splat4(1.0f);
    for (int i = 0; i < IN_SCATTERING_STEPS; ++i) {
        float t = ((float)i + 0.5f) * dt;
        V3 x_t = ray_origin + ray_dir * t;
        float distance_to_earth_center = length3(x_t);
        V3 zenith_dir = x_t / distance_to_earth_center;
        float altitude = distance_to_earth_center - EARTH_RADIUS;
        float normalized_altitude = altitude / ATMOSPHERE_THICKNESS;
        float sample_cos_theta = dot3(zenith_dir, sun_dir);
        V4 aa, as, ma, msc, ext;
        get_atmosphere_collision_coefficients(altitude, aa, as, ma, msc, ext);
        V4 transmittance_to_sun = transmittance_from_lut(tlut, sample_cos_theta, normalized_altitude, param);
        // get_multiple_scattering (:144-164) inlined with GLSL's left-to-right order
        V4 ms;
        {
            float d = distance_to_earth_center;
            float omega = 2.0f * SKY_PI * (1.0f - sqrtf(d * d - EARTH_RADIUS * EARTH_RADIUS) / d);
            V4 T_to_ground = transmittance_from_lut(tlut, sample_cos_theta, 0.0f, param);
            V4 T_ground_to_sample = transmittance_from_lut(tlut, 1.0f, 0.0f, param) / transmittance_from_lut(tlut, 1.0f, normalized_altitude, param);
            V4 L_ground = (((ground_albedo_over_pi() * (PHASE_ISOTROPIC * omega)) * T_to_ground) * T_ground_to_sample) * sample_cos_theta;
            V4 fit = {0.217f, 0.347f, 0.594f, 1.0f};
            V4 L_ms = (fit * 0.02f) * (1.0f / (1.0f + 5.0f * expf(-17.92f * sample_cos_theta)));
            ms = L_ms + L_ground;
        }
        V4 S = sun_spectral_irradiance *
               (msc * (transmittance_to_sun * molecular_phase + ms) + as * (transmittance_to_sun * aerosol_phase + ms));
        V4 step_transmittance = exp4(ext * -dt);
        V4 ext_c = {fmaxf(ext.x, 1e-7f), fmaxf(ext.y, 1e-7f), fmaxf(ext.z, 1e-7f), fmaxf(ext.w, 1e-7f)};
        V4 S_int = (S - S * step_transmittance) / ext_c;
        L_inscattering = L_inscattering + transmittance * S_int;
        transmittance = transmittance * step_transmittance;
    }
    return L_inscattering;
}

// sky-lut.glsl:278-315, one texel
void sky_texel(const uint16_t* tlut, int param, V3 sun_direction, int px, int py, uint16_t* out4) {
    float u = (float)px / (float)CS_SKY_LUT_W, v = (float)py / (float)CS_SKY_LUT_H;  // :284
    float azimuth = 2.0f * SKY_PI * u;
    float l = v * 2.0f - 1.0f;
    float elev = l * l * signf(l) * SKY_PI * 0.5f;
    V3 ray_dir = {cosf(elev) * cosf(azimuth), cosf(elev) * sinf(azimuth), sinf(elev)};
    V3 ray_origin = {0.0f, 0.0f, EYE_DISTANCE_TO_EARTH_CENTER};
    float atmos_dist = ray_sphere_intersection(ray_origin, ray_dir, ATMOSPHERE_RADIUS);
    float ground_dist = ray_sphere_intersection(ray_origin, ray_dir, EARTH_RADIUS);
    float t_d = ground_dist < 0.0f ? atmos_dist : ground_dist;
    V4 L = compute_inscattering(tlut, param, sun_direction, ray_origin, ray_dir, t_d);
    // linear_srgb_from_spectral_samples: mat4x3 M * L, column-major (:207-217)
    const float M[4][3] = {{137.672389239975f, -8.632904716299537f, -1.7181567391931372f},
                           {32.549094028629234f, 91.29801417199785f, -12.005406444382531f},
                           {-38.91428392614275f, 34.31665471469816f, 29.89044807197628f},
                           {8.572844237945445f, -11.103384660054624f, 117.47585277566478f}};
    float rgb[3];
    for (int r = 0; r < 3; r++) rgb[r] = M[0][r] * L.x + M[1][r] * L.y + M[2][r] * L.z + M[3][r] * L.w;
    out4[0] = f32_to_f16(rgb[0]); out4[1] = f32_to_f16(rgb[1]); out4[2] = f32_to_f16(rgb[2]); out4[3] = f32_to_f16(1.0f);
}

// ---------------------------------------------------------------------------------------------
// clouds.glsl
// ---------------------------------------------------------------------------------------------
const float g_radius = 6000000.0f;      // clouds.glsl:43
const float sky_b_radius = 6001500.0f;  // :44
const float sky_t_radius = 6004000.0f;  // :45
const float CLOUDS_PI = 3.141592f;      // :47 (deliberately truncated)

struct CloudCtx {
    const Volume* large; const Volume* small; const Image8* weather; const uint16_t* sky_lut;
    cs_cloud_params P;
    int primary_steps, cone_samples;
    int hier_stride = 0;        // > 1: hierarchical-march STUDY (README.md:28 TODO, not reference behaviour; cso_set_hierarchical)
    float hier_margin = 0.0f;
    int hier_lod_bias = 0;
    // per-direction step budget STUDY (clouds.glsl:227 "Take fewer steps towards horizon", never implemented there; cso_set_step_budget):
    // steps(dir) = clamp(ceil(shell length / budget_len), budget_min, primary_steps); 0 = the reference's fixed count
    float budget_len = 0.0f;
    int budget_min = 1;
    int study_variant = 0;  // cso_set_study_variant: bit 0 = light-sample positions as p + (cumulative offset) in ONE add, the way the fast CUDA
                            // kernel's per-CTA tables do, instead of the shader's sequential lp += step (clouds.glsl:187) — a noise-source study
};
struct Tally { uint64_t px = 0, steps = 0, lit = 0, evals = 0; };

// clouds.glsl:49-57
V3 getValFromSkyLUT(const CloudCtx& c, V3 rayDir) {
    float phi = atan2f(rayDir.z, rayDir.x);
    float theta = asinf(rayDir.y);
    float u = (phi / CLOUDS_PI * 0.5f + 0.5f);
    float v = sqrtf(fabsf(theta) / (CLOUDS_PI * 0.5f)) * signf(theta) * 0.5f + 0.5f;
    V4 t = sample_lut(c.sky_lut, CS_SKY_LUT_W, CS_SKY_LUT_H, u, v);
    return {t.x, t.y, t.z};
}
// clouds.glsl:60-64
float hash3(V3 p) {
    p = {fractf(p.x * 0.3183099f + 0.1f), fractf(p.y * 0.3183099f + 0.1f), fractf(p.z * 0.3183099f + 0.1f)};
    p = p * 17.0f;
    return fractf(p.x * p.y * p.z * (p.x + p.y + p.z));
}
// clouds.glsl:67-69
float remap(float v, float omin, float omax, float nmin, float nmax) {
    return nmin + (((v - omin) / (omax - omin)) * (nmax - nmin));
}
// clouds.glsl:72-75
float henyey_greenstein(float cos_theta, float g) {
    const float k = 0.0795774715459f;
    return k * (1.0f - g * g) / (powf(1.0f + g * g - 2.0f * g * cos_theta, 1.5f));
}
// clouds.glsl:77-80
float GetHeightFractionForPoint(float inPosition) {
    float height_fraction = (inPosition - sky_b_radius) / (sky_t_radius - sky_b_radius);
    return clampf(height_fraction, 0.0f, 1.0f);
}
// clouds.glsl:82-90
V4 mixGradients(float cloudType) {
    const V4 STRATUS_GRADIENT = {0.02f, 0.05f, 0.09f, 0.11f};
    const V4 STRATOCUMULUS_GRADIENT = {0.02f, 0.2f, 0.48f, 0.625f};
    const V4 CUMULUS_GRADIENT = {0.01f, 0.0625f, 0.78f, 1.0f};
    float stratus = 1.0f - clampf(cloudType * 2.0f, 0.0f, 1.0f);
    float stratocumulus = 1.0f - fabsf(cloudType - 0.5f) * 2.0f;
    float cumulus = clampf(cloudType - 0.5f, 0.0f, 1.0f) * 2.0f;
    return STRATUS_GRADIENT * stratus + STRATOCUMULUS_GRADIENT * stratocumulus + CUMULUS_GRADIENT * cumulus;
}
// clouds.glsl:92-95
float densityHeightGradient(float heightFrac, float cloudType) {
    V4 g = mixGradients(cloudType);
    return smoothstepf(g.x, g.y, heightFrac) - smoothstepf(g.z, g.w, heightFrac);
}
// clouds.glsl:97-105
float intersectSphere(V3 pos, V3 dir, float r) {
    float a = dot3(dir, dir);
    float b = 2.0f * dot3(dir, pos);
    float c = dot3(pos, pos) - (r * r);
    float d = sqrtf((b * b) - 4.0f * a * c);
    float p = -b - d;
    float p2 = -b + d;
    return fmaxf(p, p2) / (2.0f * a);
}
// clouds.glsl:109-137
float density(const CloudCtx& c, V3 pip, V3 weather, float mip, Tally& tl) {
    tl.evals++;
    V3 p = pip;
    float height_fraction = GetHeightFractionForPoint(length3(p));
    p.x += 20.0f * c.P.cloud_pos[0] * 0.6f;  // p.xz += 20.0 * cloud_pos * 0.6
    p.z += 20.0f * c.P.cloud_pos[1] * 0.6f;
    V4 n = sample_volume(*c.large, {p.x * 0.00008f, p.y * 0.00008f, p.z * 0.00008f}, mip - 2.0f);
    float fbm = n.y * 0.625f + n.z * 0.25f + n.w * 0.125f;
    float g = densityHeightGradient(height_fraction, weather.x);
    float base_cloud = remap(n.x, -(1.0f - fbm), 1.0f, 0.0f, 1.0f);
    float weather_coverage = c.P.cloud_coverage * weather.z;
    base_cloud = remap(base_cloud * g, 1.0f - (weather_coverage), 1.0f, 0.0f, 1.0f);
    base_cloud *= weather_coverage;
    p.x -= c.P.detailed_pos[0] * 40.0f;
    p.z -= c.P.detailed_pos[1] * 40.0f;
    p.y -= c.P.time * 40.0f;
    V4 hn = sample_volume(*c.small, {p.x * 0.001f, p.y * 0.001f, p.z * 0.001f}, mip);
    float hfbm = hn.x * 0.625f + hn.y * 0.25f + hn.z * 0.125f;
    hfbm = mixf(hfbm, 1.0f - hfbm, clampf(height_fraction * 4.0f, 0.0f, 1.0f));
    base_cloud = remap(base_cloud, hfbm * 0.4f * height_fraction, 1.0f, 0.0f, 1.0f);
    return powf(clampf(base_cloud, 0.0f, 1.0f), (1.0f - height_fraction) * 0.8f + 0.5f);
}

// Probe of the hierarchical-march study (cso_set_hierarchical): density() of clouds.glsl:109-126 up to the coverage
// remap, i.e. WITHOUT the detail erosion of :127-136 (which only ever lowers the value), before the `* weather_coverage`
// (positive factor) — the sign of `base_cloud * g - (1 - weather_coverage)` is the sign of that remap.
float density_probe(const CloudCtx& c, V3 pip, V3 weather, float lod, Tally& tl) {
    tl.evals++;
    V3 p = pip;
    float height_fraction = GetHeightFractionForPoint(length3(p));
    p.x += 20.0f * c.P.cloud_pos[0] * 0.6f;
    p.z += 20.0f * c.P.cloud_pos[1] * 0.6f;
    V4 n = sample_volume(*c.large, {p.x * 0.00008f, p.y * 0.00008f, p.z * 0.00008f}, lod);
    float fbm = n.y * 0.625f + n.z * 0.25f + n.w * 0.125f;
    float g = densityHeightGradient(height_fraction, weather.x);
    float base_cloud = remap(n.x, -(1.0f - fbm), 1.0f, 0.0f, 1.0f);
    float weather_coverage = c.P.cloud_coverage * weather.z;
    return base_cloud * g - (1.0f - weather_coverage);
}

const V3 RANDOM_VECTORS[6] = {  // clouds.glsl:140
    {0.38051305f, 0.92453449f, -0.02111345f}, {-0.50625799f, -0.03590792f, -0.86163418f},
    {-0.32509218f, -0.94557439f, 0.01428793f}, {0.09026238f, -0.27376545f, 0.95755165f},
    {0.28128598f, 0.42443639f, -0.86065785f}, {-0.16852403f, 0.14748697f, 0.97460106f}};

// clouds.glsl:139-215
V4 march(const CloudCtx& c, V3 pos, V3 /*end*/, V3 dir, int depth, Tally& tl) {
    const cs_cloud_params& P = c.P;
    float ss = length3(dir);
    dir = normalize3(dir);
    V3 pos10 = pos * 10.0f;
    V3 p = pos + (dir * hash3(pos10)) * ss;  // :145 (hash == 0 in fp32 at these magnitudes)

    const float t_dist = sky_t_radius - sky_b_radius;
    float lss = (t_dist / 64.0f);
    V3 LD = {P.light_direction[0], P.light_direction[1], P.light_direction[2]};
    V3 ldir = normalize3(LD);

    float t = 1.0f, T = 1.0f, alpha = 0.0f;
    V3 L = {0.0f, 0.0f, 0.0f};

    float costheta = dot3(ldir, dir);
    float phase = fmaxf(fmaxf(henyey_greenstein(costheta, 0.6f), henyey_greenstein(costheta, (0.4f - 1.4f * ldir.y))),
                        henyey_greenstein(costheta, -0.2f));  // :160

    V3 LC = {P.light_color[0], P.light_color[1], P.light_color[2]};
    V3 atmosphere_sun = ((getValFromSkyLUT(c, LD) * 0.1f) * P.light_energy) * LC;  // :163
    V3 atmosphere_ambient = getValFromSkyLUT(c, normalize3({1.0f, 1.0f, 0.0f})) * 0.05f;
    float la = length3(atmosphere_ambient);
    atmosphere_ambient = mix3(atmosphere_ambient, {la, la, la}, 0.5f);
    V3 atmosphere_ground = (getValFromSkyLUT(c, normalize3({1.0f, -1.0f, 0.0f})) * 5.0f) * 0.05f;
    float lg = length3(atmosphere_ground);
    V3 gc = {P.ground_color[0] * lg, P.ground_color[1] * lg, P.ground_color[2] * lg};
    atmosphere_ground = mix3(atmosphere_ground, gc, 0.5f);

    const float weather_scale = 0.00006f;
    V2 weather_pos = {P.weather_pos[0], P.weather_pos[1]};
    const int max_small_mip = 5;

    // Hierarchical-march study (off unless cso_set_hierarchical was called): the primary steps are grouped into blocks of hier_stride; one probe at the centre of a block's
    // sample positions (large volume only, at the mip level whose texel matches the block length) decides whether the
    // block is marched at all.  The ray positions are the reference's own (p advances by the same fp32 adds).
    const int S = c.hier_stride > 1 ? c.hier_stride : 0;
    float probe_lod = 0.0f;
    if (S) {
        float texel = 1.0f / ((float)c.large->n * 0.00008f);
        int l = (int)floorf(log2f(fmaxf((float)S * ss / texel, 1.0f))) + c.hier_lod_bias;
        probe_lod = (float)std::min(std::max(l, 0), c.large->levels - 1);
    }
    int skip = 0;
    for (int i = 0; i < depth; i++) {
        if (S && i % S == 0) {
            int nb = std::min(S, depth - i);
            V3 q = p + dir * (ss * (0.5f * (float)(nb + 1)));
            V3 wq = sample_weather(*c.weather, {q.x * weather_scale + 0.5f + weather_pos.x, q.z * weather_scale + 0.5f + weather_pos.y});
            skip = density_probe(c, q, wq, probe_lod, tl) > -c.hier_margin ? 0 : nb;
        }
        if (skip > 0) { skip--; p = p + dir * ss; continue; }
        tl.steps++;
        p = p + dir * ss;
        V3 weather_sample = sample_weather(*c.weather, {p.x * weather_scale + 0.5f + weather_pos.x, p.z * weather_scale + 0.5f + weather_pos.y});
        float height_fraction = GetHeightFractionForPoint(length3(p));
        t = density(c, p, weather_sample, 0.0f, tl);
        float dt = expf(-P.density * t * ss);
        V3 lp = p;
        float lt = 1.0f, cd = 0.0f;
        if (t > 0.0f) {
            tl.lit++;
            float lheight_fraction = 0.0f;
            V3 cum = {0.0f, 0.0f, 0.0f};
            for (int j = 0; j < c.cone_samples; j++) {  // 6 in the reference (:186)
                V3 step = (ldir + RANDOM_VECTORS[j % 6] * (float)j) * lss;
                if (c.study_variant & 1) { cum = cum + step; lp = p + cum; }
                else lp = lp + step;
                lheight_fraction = GetHeightFractionForPoint(length3(lp));
                V3 lweather = sample_weather(*c.weather, {lp.x * weather_scale + 0.5f + weather_pos.x, lp.z * weather_scale + 0.5f + weather_pos.y});
                lt = density(c, lp, lweather, (float)j, tl);
                cd += lt;
            }
            lp = p + (ldir * 18.0f) * lss;  // :195
            lheight_fraction = GetHeightFractionForPoint(length3(lp));
            V3 lweather = sample_weather(*c.weather, {lp.x * weather_scale + 0.5f, lp.z * weather_scale + 0.5f});  // :197 no weather_pos
            lt = powf(density(c, lp, lweather, (float)max_small_mip, tl), (1.0f - lheight_fraction) * 0.8f + 0.5f);
            cd += lt;

            float beers = expf(-P.density * cd * lss * 3.0f);
            float powder_sugar_effect = 1.0f - expf(-P.density * cd * lss * 3.0f * 2.0f);
            float beers_total = 2.0f * beers * powder_sugar_effect;

            V3 ambient = mix3(atmosphere_ground, atmosphere_ambient, smoothstepf(0.0f, 1.0f, height_fraction));
            alpha += (1.0f - dt) * (1.0f - alpha);
            V3 radiance = (ambient + (atmosphere_sun * beers_total) * phase) * t;
            L = L + ((radiance - radiance * dt) * T) / fmaxf(0.0000001f, t);
            T *= dt;
        }
    }
    alpha = clampf(alpha, 0.0f, 1.0f);
    return {L.x, L.y, L.z, alpha};
}

// clouds.glsl:218-237
V4 sky(const CloudCtx& c, V3 dir, Tally& tl) {
    V4 col = {0, 0, 0, 0};
    if (dir.y > 0.0f) {
        tl.px++;
        V3 camPos = {0.0f, g_radius, 0.0f};
        V3 start = camPos + dir * intersectSphere(camPos, dir, sky_b_radius);
        V3 end = camPos + dir * intersectSphere(camPos, dir, sky_t_radius);
        float shelldist = length3(end - start);
        int n_steps = c.primary_steps;
        if (c.budget_len > 0.0f) n_steps = std::min(c.primary_steps, std::max(c.budget_min, (int)ceilf(shelldist / c.budget_len)));
        float steps = (float)n_steps;  // 128.0 in the reference (:228)
        V3 raystep = (dir * shelldist) / steps;
        col = march(c, start, end, raystep, n_steps, tl);
    }
    return col;
}
// clouds.glsl:239-256
V3 oct_to_vec3(V2 e) {
    V3 n;
    n.x = (e.x - e.y);
    n.y = (e.x + e.y) - 1.0f;
    n.z = 1.0f - fabsf(n.x) - fabsf(n.y);
    if (!(n.z >= 0.0f)) {  // oct_wrap (:239-244)
        float sx = n.x >= 0.0f ? 1.0f : -1.0f, sy = n.y >= 0.0f ? 1.0f : -1.0f;
        float wx = (1.0f - fabsf(n.y)) * sx, wy = (1.0f - fabsf(n.x)) * sy;
        n.x = wx; n.y = wy;
    }
    return normalize3(n);
}
// clouds.glsl:258-266, one pixel
V4 cloud_pixel(const CloudCtx& c, int px, int py, Tally& tl) {
    V2 uv = {(float)px / c.P.texture_size[0], (float)py / c.P.texture_size[1]};
    V3 n = oct_to_vec3(uv);
    V3 dir = {n.x, n.z, n.y};  // .xzy
    return sky(c, dir, tl);
}

template <class F>
void parallel_rows(int n_threads, int rows, F f) {
    if (n_threads <= 1 || rows <= 1) { for (int r = 0; r < rows; r++) f(r, 0); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
        th.emplace_back([&, t]() { for (;;) { int r = next.fetch_add(1); if (r >= rows) break; f(r, t); } });
    for (auto& x : th) x.join();
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI (include/cloudsky.h) — oracle backend
// ---------------------------------------------------------------------------------------------
struct cs_context {
    std::string err;
    int threads = 1;
    Volume large, small;
    Image8 weather;
    bool have_tex = false, have_tlut = false, have_sky = false;
    int tlut_param = CS_TLUT_LINEAR;
    std::vector<uint16_t> tlut = std::vector<uint16_t>((size_t)CS_TRANSMITTANCE_W * CS_TRANSMITTANCE_H * 4);
    std::vector<uint16_t> skylut = std::vector<uint16_t>((size_t)CS_SKY_LUT_W * CS_SKY_LUT_H * 4);
    int W = 0, H = 0;
    std::vector<uint16_t> image;
    int primary_steps = CS_REF_PRIMARY_STEPS, cone_samples = CS_REF_CONE_SAMPLES;
    int hier_stride = 0, hier_lod_bias = 0;
    float hier_margin = 0.0f;
    float budget_len = 0.0f;
    int budget_min = 1;
    int study_variant = 0;
    cs_counters counters{};
};

static int fail(cs_context* c, int code, const char* msg) { if (c) c->err = msg; return code; }

extern "C" {

int cs_create(int, cs_context** out) {
    if (!out) return CS_ERR_INVALID;
    *out = new cs_context();
    return CS_OK;
}
void cs_destroy(cs_context* c) { delete c; }
const char* cs_last_error(const cs_context* c) { return c ? c->err.c_str() : "null context"; }
const char* cs_backend_name(void) { return "oracle-cpu"; }
int cs_set_stream(cs_context* c, void*) { return fail(c, CS_ERR_UNSUPPORTED, "oracle backend has no CUDA stream"); }
int cs_sync(cs_context*) { return CS_OK; }
int cs_set_threads(cs_context* c, int n) { if (!c || n < 1) return CS_ERR_INVALID; c->threads = n; return CS_OK; }

static void to_rgba(const uint8_t* src, size_t texels, int ch, std::vector<uint8_t>& dst) {
    dst.resize(texels * 4);
    for (size_t i = 0; i < texels; i++) {
        for (int k = 0; k < 4; k++) dst[i * 4 + k] = k < ch ? src[i * ch + k] : 255;
    }
}

int cs_upload_textures(cs_context* c, const uint8_t* large, int ln, int lch, const uint8_t* small, int sn, int sch,
                       const uint8_t* weather, int ww, int wh, int wch) {
    if (!c) return CS_ERR_INVALID;
    if (!large || !small || !weather || ln < 1 || sn < 1 || ww < 1 || wh < 1 || lch < 3 || lch > 4 || sch < 3 || sch > 4 ||
        wch < 3 || wch > 4 || (ln & (ln - 1)) || (sn & (sn - 1)))
        return fail(c, CS_ERR_INVALID, "cs_upload_textures: bad dimensions (volumes must be power-of-two cubes, 3 or 4 channels)");
    c->large = Volume(); c->small = Volume();
    c->large.n = ln; c->large.mip.emplace_back(); to_rgba(large, (size_t)ln * ln * ln, lch, c->large.mip[0]); build_mips(c->large);
    c->small.n = sn; c->small.mip.emplace_back(); to_rgba(small, (size_t)sn * sn * sn, sch, c->small.mip[0]); build_mips(c->small);
    c->weather.w = ww; c->weather.h = wh; to_rgba(weather, (size_t)ww * wh, wch, c->weather.px);
    c->have_tex = true;
    return CS_OK;
}
int cs_load_texture_files(cs_context* c, const char*, int, const char*, int, const char*) {
    return fail(c, CS_ERR_UNSUPPORTED, "oracle backend takes decoded texels only (cs_upload_textures)");
}
int cs_decode_image_file(const char*, uint8_t**, int*, int*, int*) { return CS_ERR_UNSUPPORTED; }
void cs_free(void* p) { free(p); }
int cs_read_volume_level(cs_context* c, int which, int level, uint8_t* out, size_t bytes) {
    if (!c || !c->have_tex) return fail(c, CS_ERR_NOT_READY, "no textures");
    const Volume& v = which == 0 ? c->large : c->small;
    if (level < 0 || level >= v.levels || bytes != v.mip[level].size()) return fail(c, CS_ERR_INVALID, "bad level/size");
    memcpy(out, v.mip[level].data(), bytes);
    return CS_OK;
}

int cs_build_transmittance_lut(cs_context* c) {
    if (!c) return CS_ERR_INVALID;
    parallel_rows(c->threads, CS_TRANSMITTANCE_H, [&](int y, int) {
        for (int x = 0; x < CS_TRANSMITTANCE_W; x++) transmittance_texel(x, y, &c->tlut[((size_t)y * CS_TRANSMITTANCE_W + x) * 4], c->tlut_param);
    });
    c->have_tlut = true;
    return CS_OK;
}
int cs_set_transmittance_parametrisation(cs_context* c, int which) {
    if (!c || (which != CS_TLUT_LINEAR && which != CS_TLUT_BRUNETON2017)) return fail(c, CS_ERR_INVALID, "cs_set_transmittance_parametrisation: CS_TLUT_LINEAR or CS_TLUT_BRUNETON2017");
    if (which != c->tlut_param) { c->tlut_param = which; c->have_tlut = false; c->have_sky = false; }  // both LUTs depend on the mapping
    return CS_OK;
}
int cs_build_sky_lut(cs_context* c, const float sun[3]) {
    if (!c || !sun) return CS_ERR_INVALID;
    if (!c->have_tlut) return fail(c, CS_ERR_NOT_READY, "Attempting to update uninitialized sky lut (no transmittance LUT)");
    V3 s = {sun[0], sun[1], sun[2]};
    parallel_rows(c->threads, CS_SKY_LUT_H, [&](int y, int) {
        for (int x = 0; x < CS_SKY_LUT_W; x++) sky_texel(c->tlut.data(), c->tlut_param, s, x, y, &c->skylut[((size_t)y * CS_SKY_LUT_W + x) * 4]);
    });
    c->have_sky = true;
    return CS_OK;
}
int cs_read_transmittance_lut(cs_context* c, uint16_t* out, size_t bytes) {
    if (!c || !c->have_tlut) return fail(c, CS_ERR_NOT_READY, "no transmittance LUT");
    if (bytes != c->tlut.size() * 2) return fail(c, CS_ERR_INVALID, "size");
    memcpy(out, c->tlut.data(), bytes); return CS_OK;
}
int cs_read_sky_lut(cs_context* c, uint16_t* out, size_t bytes) {
    if (!c || !c->have_sky) return fail(c, CS_ERR_NOT_READY, "no sky LUT");
    if (bytes != c->skylut.size() * 2) return fail(c, CS_ERR_INVALID, "size");
    memcpy(out, c->skylut.data(), bytes); return CS_OK;
}
int cs_write_transmittance_lut(cs_context* c, const uint16_t* in, size_t bytes) {
    if (!c || bytes != c->tlut.size() * 2) return fail(c, CS_ERR_INVALID, "size");
    memcpy(c->tlut.data(), in, bytes); c->have_tlut = true; return CS_OK;
}
int cs_write_sky_lut(cs_context* c, const uint16_t* in, size_t bytes) {
    if (!c || bytes != c->skylut.size() * 2) return fail(c, CS_ERR_INVALID, "size");
    memcpy(c->skylut.data(), in, bytes); c->have_sky = true; return CS_OK;
}

int cs_resize(cs_context* c, int w, int h) {
    if (!c || w < 1 || h < 1) return fail(c, CS_ERR_INVALID, "bad size");
    c->W = w; c->H = h; c->image.assign((size_t)w * h * 4, 0);
    return CS_OK;
}
int cs_set_march_config(cs_context* c, int p, int cone, int) {
    if (!c || p < 1 || p > 4096 || cone < 0 || cone > 64) return fail(c, CS_ERR_INVALID, "bad march config");
    c->primary_steps = p; c->cone_samples = cone; return CS_OK;
}
int cs_set_step_budget(cs_context* c, float step_len_m, int min_steps) {  // include/cloudsky.h: per-direction primary-step budget (0 = fixed count)
    if (!c) return CS_ERR_INVALID;
    if (!(step_len_m >= 0.0f) || min_steps < 1) return fail(c, CS_ERR_INVALID, "cs_set_step_budget: min_step_length_m >= 0, min_steps >= 1");
    c->budget_len = step_len_m; c->budget_min = min_steps; return CS_OK;
}
int cs_set_counters_enabled(cs_context*, int) { return CS_OK; }
int cs_get_counters(cs_context* c, cs_counters* out) { if (!c || !out) return CS_ERR_INVALID; *out = c->counters; return CS_OK; }

static int render_region(cs_context* c, const cs_cloud_params* P, int x0, int y0, int x1, int y1, uint16_t* dst) {
    if (!c || !P) return CS_ERR_INVALID;
    if (!c->have_tex || !c->have_sky) return fail(c, CS_ERR_NOT_READY, "textures or sky LUT missing (can_run == false)");
    if (c->W < 1) return fail(c, CS_ERR_NOT_READY, "cs_resize not called");
    if ((int)P->texture_size[0] != c->W || (int)P->texture_size[1] != c->H) return fail(c, CS_ERR_INVALID, "params.texture_size != image size");
    x0 = std::max(x0, 0); y0 = std::max(y0, 0); x1 = std::min(x1, c->W); y1 = std::min(y1, c->H);
    CloudCtx cc{&c->large, &c->small, &c->weather, c->skylut.data(), *P, c->primary_steps, c->cone_samples};
    cc.hier_stride = c->hier_stride; cc.hier_margin = c->hier_margin; cc.hier_lod_bias = c->hier_lod_bias;
    cc.budget_len = c->budget_len; cc.budget_min = c->budget_min; cc.study_variant = c->study_variant;
    std::vector<Tally> tl((size_t)std::max(c->threads, 1));
    parallel_rows(c->threads, std::max(y1 - y0, 0), [&](int r, int t) {
        int y = y0 + r;
        for (int x = x0; x < x1; x++) {
            V4 col = cloud_pixel(cc, x, y, tl[t]);
            uint16_t* o = &dst[((size_t)y * c->W + x) * 4];
            o[0] = f32_to_f16(col.x); o[1] = f32_to_f16(col.y); o[2] = f32_to_f16(col.z); o[3] = f32_to_f16(col.w);
        }
    });
    cs_counters k{};
    for (auto& t : tl) { k.marched_pixels += t.px; k.primary_steps += t.steps; k.lit_steps += t.lit; k.density_evals += t.evals; }
    k.large_fetches = k.density_evals; k.small_fetches = k.density_evals;
    c->counters = k;
    return CS_OK;
}
int cs_dispatch_clouds(cs_context* c, const cs_cloud_params* P, int gx, int gy) {
    if (!c || !P || gx < 1 || gy < 1) return fail(c, CS_ERR_INVALID, "bad dispatch");
    int x0 = (int)P->update_position[0], y0 = (int)P->update_position[1];
    return render_region(c, P, x0, y0, x0 + 8 * gx, y0 + 8 * gy, c->image.data());
}
int cs_render_frame(cs_context* c, const cs_cloud_params* P) {
    if (!c || !P) return CS_ERR_INVALID;
    return render_region(c, P, 0, 0, c->W, c->H, c->image.data());
}
int cs_render_rows_to(cs_context* c, const cs_cloud_params* P, int r0, int r1, void* out) {
    if (!c || !P || !out) return CS_ERR_INVALID;
    return render_region(c, P, 0, r0, c->W, r1, (uint16_t*)out);  // "device" memory is host memory here
}
int cs_render_row_bands_to(cs_context* c, const cs_cloud_params* P, int first_row, int band_rows, int band_pitch_rows, int n_bands, void* out) {
    if (!c || !P || !out) return CS_ERR_INVALID;
    if (first_row < 0 || band_rows < 8 || band_rows % 8 != 0 || band_pitch_rows < band_rows || n_bands < 1)
        return fail(c, CS_ERR_INVALID, "cs_render_row_bands_to: band_rows must be a positive multiple of 8, band_pitch_rows >= band_rows, n_bands >= 1");
    for (int b = 0; b < n_bands; b++) {  // independent tiles (cloud_sky.gd:156-161): the oracle simply renders them one after the other
        const int r0 = first_row + b * band_pitch_rows;
        if (r0 >= c->H) break;
        int r = render_region(c, P, 0, r0, c->W, std::min(r0 + band_rows, c->H), (uint16_t*)out);
        if (r) return r;
    }
    return CS_OK;
}
void* cs_image_device_ptr(cs_context* c) { return c ? c->image.data() : nullptr; }
int cs_read_image(cs_context* c, uint16_t* out, size_t bytes) {
    if (!c || !out || bytes != c->image.size() * 2) return fail(c, CS_ERR_INVALID, "size");
    memcpy(out, c->image.data(), bytes); return CS_OK;
}
int cs_render_frame_host(cs_context* c, const cs_cloud_params* P, uint16_t* out, size_t bytes) {
    if (!c || !P || !out || bytes != c->image.size() * 2) return fail(c, CS_ERR_INVALID, "size");
    int r = cs_build_sky_lut(c, P->light_direction); if (r) return r;
    r = cs_render_frame(c, P); if (r) return r;
    memcpy(out, c->image.data(), bytes); return CS_OK;
}
int cs_render_frame_host_async(cs_context* c, const cs_cloud_params* P, uint16_t* out, size_t bytes) { return cs_render_frame_host(c, P, out, bytes); }
int cs_wait_host(cs_context*) { return CS_OK; }
int cs_render_sun_batch_to(cs_context* c, const cs_cloud_params* P, const float* suns, int n, void* out) {
    if (!c || !P || !suns || !out || n < 1) return CS_ERR_INVALID;
    for (int i = 0; i < n; i++) {
        cs_cloud_params q = *P;
        memcpy(q.light_direction, suns + 3 * i, 12);
        int r = cs_build_sky_lut(c, q.light_direction); if (r) return r;
        r = render_region(c, &q, 0, 0, c->W, c->H, (uint16_t*)out + (size_t)i * c->W * c->H * 4); if (r) return r;
    }
    return CS_OK;
}
int cs_time_render_frame(cs_context* c, const cs_cloud_params*, int, int, float*) {
    return fail(c, CS_ERR_UNSUPPORTED, "device timing is CUDA-only");
}

int cs_set_kernel_timing(cs_context* c, int) { return fail(c, CS_ERR_UNSUPPORTED, "device timing is CUDA-only"); }
// peer-mapped output replicas are a CUDA/NVLink mechanism; the oracle renders into host memory (gloo tests gather with torch)
int cs_peer_alloc(cs_context* c, size_t, void**, uint8_t*) { return fail(c, CS_ERR_UNSUPPORTED, "peer buffers are CUDA-only"); }
int cs_peer_open(cs_context* c, const uint8_t*, void**) { return fail(c, CS_ERR_UNSUPPORTED, "peer buffers are CUDA-only"); }
int cs_peer_close(cs_context* c, void*) { return fail(c, CS_ERR_UNSUPPORTED, "peer buffers are CUDA-only"); }
int cs_peer_free(cs_context* c, void*) { return fail(c, CS_ERR_UNSUPPORTED, "peer buffers are CUDA-only"); }
int cs_set_output_mirrors(cs_context* c, void*, size_t, int, void* const*) { return fail(c, CS_ERR_UNSUPPORTED, "peer buffers are CUDA-only"); }
int cs_peer_barrier(cs_context* c, int, int, void* const*, uint32_t) { return fail(c, CS_ERR_UNSUPPORTED, "peer buffers are CUDA-only"); }
int cs_peer_check(cs_context* c) { return fail(c, CS_ERR_UNSUPPORTED, "peer buffers are CUDA-only"); }
int cs_read_kernel_timings(cs_context* c, float*, int*, float*, int*) { return fail(c, CS_ERR_UNSUPPORTED, "device timing is CUDA-only"); }

// ---- noise synthesis (cs_generate_noise; README.md:30 TODO) ----------------------------------------
// CPU statement of the generator defined in include/cloudsky.h: integer lattice hash, inverted Worley F1 with one
// feature point per cell, hash-gradient Perlin noise with quintic fade, fBm 1/2,1/4,..., the shader's own 0.625/0.25/0.125
// channel weights (clouds.glsl:118,133).  The reference has no generator to follow (its noise ships as bitmaps), so this
// is the definition the CUDA kernel is checked against, byte for byte.
} // extern "C"
namespace noisegen {
uint32_t finalise(uint32_t h) { h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16; return h; }
uint32_t hash_cell(int x, int y, int z, uint32_t seed) {
    uint32_t h = finalise(seed);
    h = finalise((uint32_t)z + h);
    h = finalise((uint32_t)y + h);
    return finalise((uint32_t)x + h);
}
float to_unit(uint32_t h) { return (float)(h >> 8) / 16777216.0f; }
int modp(int i, int f) { int m = i % f; return m < 0 ? m + f : m; }

float worley_inv(const float p[3], int freq, float scale, uint32_t seed) {
    float q[3], fr[3]; int cell[3];
    for (int a = 0; a < 3; a++) { q[a] = p[a] * (float)freq; float fl = floorf(q[a]); fr[a] = q[a] - fl; cell[a] = (int)fl; }
    float nearest2 = 1.0e30f;
    for (int k = 0; k < 27; k++) {  // dx fastest, then dy, then dz
        int d[3] = {k % 3 - 1, (k / 3) % 3 - 1, k / 9 - 1};
        uint32_t hx = hash_cell(modp(cell[0] + d[0], freq), modp(cell[1] + d[1], freq), modp(cell[2] + d[2], freq), seed);
        uint32_t hy = finalise(hx + 0x9e3779b9u);
        uint32_t hz = finalise(hy + 0x9e3779b9u);
        uint32_t hh[3] = {hx, hy, hz};
        float r[3];
        for (int a = 0; a < 3; a++) r[a] = ((float)d[a] + to_unit(hh[a])) - fr[a];
        float d2 = (r[0] * r[0] + r[1] * r[1]) + r[2] * r[2];
        if (d2 < nearest2) nearest2 = d2;
    }
    float dist = sqrtf(nearest2) * scale;
    return 1.0f - (dist < 1.0f ? dist : 1.0f);
}
float corner_gradient(uint32_t h, float x, float y, float z) {
    static const int gx[16] = {1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0, 1, 0, -1, 0};
    static const int gy[16] = {1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1};
    static const int gz[16] = {0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1, 0, 1, 0, -1};
    // the 12 edge directions (+4 repeats), written as the sum of the two non-zero signed components in (first, second) order
    int i = (int)(h & 15u);
    float first, second;
    if (gx[i] != 0 && gy[i] != 0) { first = gx[i] > 0 ? x : -x; second = gy[i] > 0 ? y : -y; }
    else if (gx[i] != 0) { first = gx[i] > 0 ? x : -x; second = gz[i] > 0 ? z : -z; }
    else { first = gy[i] > 0 ? y : -y; second = gz[i] > 0 ? z : -z; }
    return first + second;
}
float quintic(float t) { return t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f); }
float gradient_noise(const float p[3], int freq, uint32_t seed) {
    float t[3]; int lo[3], hi[3];
    for (int a = 0; a < 3; a++) { float q = p[a] * (float)freq; float fl = floorf(q); t[a] = q - fl; lo[a] = (int)fl; hi[a] = modp(lo[a] + 1, freq); }
    float corner[2][2][2];
    for (int cz = 0; cz < 2; cz++)
        for (int cy = 0; cy < 2; cy++)
            for (int cx = 0; cx < 2; cx++)
                corner[cz][cy][cx] = corner_gradient(hash_cell(cx ? hi[0] : lo[0], cy ? hi[1] : lo[1], cz ? hi[2] : lo[2], seed),
                                                     cx ? t[0] - 1.0f : t[0], cy ? t[1] - 1.0f : t[1], cz ? t[2] - 1.0f : t[2]);
    float u = quintic(t[0]), v = quintic(t[1]), w = quintic(t[2]);
    float plane[2];
    for (int cz = 0; cz < 2; cz++) {
        float e0 = lerpf(corner[cz][0][0], corner[cz][0][1], u), e1 = lerpf(corner[cz][1][0], corner[cz][1][1], u);
        plane[cz] = lerpf(e0, e1, v);
    }
    return lerpf(plane[0], plane[1], w);
}
float gradient_fbm(const float p[3], int freq, int octaves, uint32_t seed) {
    float total = 0.0f, amplitude = 0.5f;
    for (int o = 0; o < octaves; o++) { total = total + amplitude * gradient_noise(p, freq << o, seed + (uint32_t)o); amplitude = amplitude * 0.5f; }
    return total;
}
float sat01(float x) { return clampf(x, 0.0f, 1.0f); }
uint8_t quantise(float v) { return (uint8_t)(int)(sat01(v) * 255.0f + 0.5f); }
float weights3(float a, float b, float c) { return (a * 0.625f + b * 0.25f) + c * 0.125f; }

void texel(int kind, const cs_noise_params& P, int n, int x, int y, int z, uint8_t* out) {
    float p[3] = {((float)x + 0.5f) * (1.0f / (float)n), ((float)y + 0.5f) * (1.0f / (float)n), kind == CS_NOISE_WEATHER ? 0.0f : ((float)z + 0.5f) * (1.0f / (float)n)};
    if (kind == CS_NOISE_WEATHER) {
        float wf = weights3(worley_inv(p, P.worley_frequency, P.worley_scale, P.seed), worley_inv(p, P.worley_frequency * 2, P.worley_scale, P.seed + 101u),
                            worley_inv(p, P.worley_frequency * 4, P.worley_scale, P.seed + 202u));
        float p01 = sat01(0.5f + P.perlin_scale * gradient_fbm(p, P.perlin_frequency, P.perlin_octaves, P.seed ^ 0x5bd1e995u));
        float pw = wf + p01 * (1.0f - wf);
        float t01 = sat01(0.5f + P.perlin_scale * gradient_fbm(p, 2, 3, P.seed ^ 0x2545f491u));
        out[0] = quantise(P.type_lo + (P.type_hi - P.type_lo) * t01);
        out[1] = 0;
        out[2] = quantise(sat01((pw - P.remap_lo) / (P.remap_hi - P.remap_lo)));
        out[3] = 255;
        return;
    }
    float oct[5];
    for (int k = 0; k < 5; k++) oct[k] = worley_inv(p, P.worley_frequency << k, P.worley_scale, P.seed + 101u * (uint32_t)k);
    float f0 = weights3(oct[0], oct[1], oct[2]), f1 = weights3(oct[1], oct[2], oct[3]), f2 = weights3(oct[2], oct[3], oct[4]);
    if (kind == CS_NOISE_SMALL) { out[0] = quantise(f0); out[1] = quantise(f1); out[2] = quantise(f2); out[3] = 255; return; }
    float p01 = sat01(0.5f + P.perlin_scale * gradient_fbm(p, P.perlin_frequency, P.perlin_octaves, P.seed ^ 0x5bd1e995u));
    out[0] = quantise(f0 + p01 * (1.0f - f0));
    out[1] = quantise(f0); out[2] = quantise(f1); out[3] = quantise(f2);
}
}  // namespace noisegen
extern "C" {
void cs_noise_params_default(int kind, cs_noise_params* p) {
    if (!p) return;
    p->seed = 1u;
    p->worley_frequency = kind == CS_NOISE_SMALL ? 2 : 4;
    p->worley_scale = kind == CS_NOISE_WEATHER ? 1.0f : 0.56f;
    p->perlin_frequency = 4;
    p->perlin_octaves = kind == CS_NOISE_WEATHER ? 4 : 5;
    p->perlin_scale = 1.0f;
    p->remap_lo = 0.55f; p->remap_hi = 0.95f;
    p->type_lo = 0.59f; p->type_hi = 0.91f;
}
int cs_generate_noise(cs_context* c, int kind, int n, const cs_noise_params* P, uint8_t* out, size_t bytes) {
    if (!c) return CS_ERR_INVALID;
    if (kind < CS_NOISE_LARGE || kind > CS_NOISE_WEATHER || !P) return fail(c, CS_ERR_INVALID, "cs_generate_noise: bad kind / params");
    if (n < 1 || n > (kind == CS_NOISE_WEATHER ? 8192 : 512) || (n & (n - 1))) return fail(c, CS_ERR_INVALID, "cs_generate_noise: n must be a power of two");
    if (P->worley_frequency < 1 || P->worley_frequency > 256 || P->perlin_frequency < 1 || P->perlin_frequency > 256 || P->perlin_octaves < 1 || P->perlin_octaves > 8 ||
        !(P->remap_hi > P->remap_lo) || !(P->type_hi >= P->type_lo) || !(P->perlin_scale >= 0.0f) || !(P->worley_scale > 0.0f))
        return fail(c, CS_ERR_INVALID, "cs_generate_noise: bad parameters");
    const int depth = kind == CS_NOISE_WEATHER ? 1 : n;
    if (!out || bytes != (size_t)n * n * depth * 4) return fail(c, CS_ERR_INVALID, "cs_generate_noise: out_bytes must be texels * 4");
    parallel_rows(c->threads, n * depth, [&](int row, int) {
        int y = row % n, z = row / n;
        for (int x = 0; x < n; x++) noisegen::texel(kind, *P, n, x, y, z, out + (((size_t)z * n + y) * n + x) * 4);
    });
    return CS_OK;
}

// ---- host-side parameter logic (cloud_sky.gd) -------------------------------------------------
void cs_settings_default(cs_sky_settings* s) {  // cloud_sky.gd:4-50
    s->wind_direction = 0.0f; s->wind_speed = 1.0f; s->density = 0.05f; s->cloud_coverage = 0.25f; s->time_offset = 0.0f;
    s->sun_disk_scale = 1.0f; s->ground_color[0] = s->ground_color[1] = s->ground_color[2] = s->ground_color[3] = 1.0f;
    s->frames_to_update = 64; s->texture_size = 768;
}
void cs_settings_demo(cs_sky_settings* s) {  // clouds_sky.tres:11-18
    cs_settings_default(s);
    s->cloud_coverage = 0.2f; s->sun_disk_scale = 2.0f;
    s->ground_color[0] = 0.270588f; s->ground_color[1] = 0.188235f; s->ground_color[2] = 0.027451f; s->ground_color[3] = 1.0f;
}
void cs_frame_state_init(cs_frame_state* st) {  // cloud_sky.gd:66-74
    memset(st, 0, sizeof(*st));
    st->light_direction[1] = -1.0f; st->light_energy = 1.0f;
    st->light_color[0] = st->light_color[1] = st->light_color[2] = 1.0f;
}
static float srgb_to_linear(float c) {  // Godot Color::srgb_to_linear
    return c < 0.04045f ? c * (1.0f / 12.92f) : powf((c + 0.055f) * (float)(1.0 / (1.0 + 0.055)), 2.4f);
}
void cs_frame_state_set_light(cs_frame_state* st, const float b[9], float energy, const float srgb[3]) {  // cloud_sky.gd:76-79
    V3 d = {b[6], b[7], b[8]};  // basis * (0,0,1) = third column
    d = normalize3(d);
    st->light_direction[0] = d.x; st->light_direction[1] = d.y; st->light_direction[2] = d.z;
    st->light_energy = energy;
    for (int i = 0; i < 3; i++) st->light_color[i] = srgb_to_linear(srgb[i]);
}
void cs_frame_advance(cs_frame_state* st, const cs_sky_settings* s, float time) {  // cloud_sky.gd:165-187
    float dx = cosf(s->wind_direction), dy = sinf(s->wind_direction);  // Vector2.from_angle
    float delta = time - st->time;
    float delta2 = delta * 0.001f + 0.005f * s->time_offset;
    float len = sqrtf(dx * dx + dy * dy);  // .normalized()
    float nx = dx / len, ny = dy / len;
    st->time = time;
    st->detailed_pos[0] += delta * nx; st->detailed_pos[1] += delta * ny;
    st->cloud_pos[0] += delta * nx * s->wind_speed; st->cloud_pos[1] += delta * ny * s->wind_speed;
    st->weather_pos[0] += delta2 * nx * s->wind_speed; st->weather_pos[1] += delta2 * ny * s->wind_speed;
}
void cs_fill_cloud_params(cs_cloud_params* o, const cs_sky_settings* s, const cs_frame_state* st, int w, int h, int ux, int uy) {  // cloud_sky.gd:251-289
    memset(o, 0, sizeof(*o));
    o->texture_size[0] = (float)w; o->texture_size[1] = (float)h;
    o->update_position[0] = (float)ux; o->update_position[1] = (float)uy;
    memcpy(o->cloud_pos, st->cloud_pos, 8); memcpy(o->detailed_pos, st->detailed_pos, 8); memcpy(o->weather_pos, st->weather_pos, 8);
    memcpy(o->ground_color, s->ground_color, 16);
    memcpy(o->light_direction, st->light_direction, 12); o->light_energy = st->light_energy;
    memcpy(o->light_color, st->light_color, 12); o->time = st->time;
    o->density = s->density; o->cloud_coverage = s->cloud_coverage; o->time_offset = s->time_offset;
}
void cs_update_performance(int* ts, int frames, int* region, int* groups) {  // cloud_sky.gd:109-118
    int fs = (int)sqrt((double)frames);
    int r = *ts / fs;
    if (*ts % fs != 0) *ts = r * fs;
    *region = r; *groups = (r + 7) / 8;
}
void cs_next_update_position(int* x, int* y, int region, int ts) {  // cloud_sky.gd:156-161
    *x += region;
    if (*x >= ts) { *x = 0; *y += region; }
    if (*y >= ts) { *x = 0; *y = 0; }
}


// ---- the Sky resource: restatement of cloud_sky.gd's update_sky state machine (cloud_sky.gd:109-163) ----------
}  // extern "C" (reopened below)

struct cs_sky {
    cs_context* c;
    cs_sky_settings s;
    cs_frame_state fd;
    bool sun_attached = false;
    float basis[9], energy, color[3];
    std::vector<uint16_t> textures[3], luts[3];
    int size = 0, region = 0, groups = 0, pos[2] = {0, 0};
    int tex_update = 0, tex_from = 1, tex_to = 2, frame = 0;
    float blend = 0.0f;
    bool can_run = false, full_init = true, lut_full = true;
    int lut_current = 0, lut_updates = 0;
};

static int sky_tick(cs_sky* k, float now);

static void sky_alloc(cs_sky* k) {  // update_performance + texture creation (cloud_sky.gd:109-118,368-402)
    int ts = k->s.texture_size;
    cs_update_performance(&ts, k->s.frames_to_update, &k->region, &k->groups);
    k->s.texture_size = k->size = ts;
    for (auto& t : k->textures) t.assign((size_t)ts * ts * 4, 0);
    if (k->c->W != ts || k->c->H != ts) cs_resize(k->c, ts, ts);
    k->can_run = true;
}
static int sky_render_lut(cs_sky* k) {  // sky_lut.gd:122-148
    int r = cs_build_sky_lut(k->c, k->fd.light_direction);
    if (r) return r;
    k->luts[k->lut_current] = k->c->skylut;
    k->lut_current = (k->lut_current + 1) % 3;
    k->lut_updates++;
    return CS_OK;
}
static int sky_frame_data(cs_sky* k, float now) {  // cloud_sky.gd:165-187 + sky_lut.gd:43-52
    if (k->sun_attached) cs_frame_state_set_light(&k->fd, k->basis, k->energy, k->color);
    cs_frame_advance(&k->fd, &k->s, now);
    int r = sky_render_lut(k);
    if (r == CS_OK && k->lut_full) {
        r = sky_render_lut(k);
        if (r == CS_OK) r = sky_render_lut(k);
        k->lut_full = false;
    }
    return r;
}
static int sky_tick(cs_sky* k, float now) {  // update_sky (cloud_sky.gd:129-163)
    if (!k->can_run) return CS_OK;
    int r;
    if (k->full_init) {
        k->full_init = false;
        if ((r = sky_frame_data(k, now)) != CS_OK) return r;                      // initialize_sky (cloud_sky.gd:124-127)
        for (int i = 0; i < k->s.frames_to_update * 2; i++) if ((r = sky_tick(k, now)) != CS_OK) return r;
    }
    if (k->frame >= k->s.frames_to_update) {
        k->tex_update = (k->tex_update + 1) % 3; k->tex_from = (k->tex_from + 1) % 3; k->tex_to = (k->tex_to + 1) % 3;
        if ((r = sky_frame_data(k, now)) != CS_OK) return r;
        k->frame = 0;
    }
    k->blend = (float)k->frame / (float)k->s.frames_to_update;
    cs_cloud_params P;
    cs_fill_cloud_params(&P, &k->s, &k->fd, k->size, k->size, k->pos[0], k->pos[1]);
    k->c->skylut = k->luts[(k->lut_current + 2) % 3];  // the LUT rendered last (cloud_sky.gd:242)
    k->c->have_sky = true;
    r = render_region(k->c, &P, k->pos[0], k->pos[1], k->pos[0] + 8 * k->groups, k->pos[1] + 8 * k->groups, k->textures[k->tex_update].data());
    if (r) return r;
    cs_next_update_position(&k->pos[0], &k->pos[1], k->region, k->size);
    k->frame++;
    return CS_OK;
}

extern "C" {

int cs_sky_create(cs_context* c, const cs_sky_settings* s, cs_sky** out) {
    if (!c || !s || !out) return CS_ERR_INVALID;
    *out = nullptr;
    if (!c->have_tex || !c->have_tlut) return fail(c, CS_ERR_NOT_READY, "cs_sky_create: textures and transmittance LUT first");
    if (s->frames_to_update < 1) return fail(c, CS_ERR_INVALID, "frames_to_update must be >= 1");
    cs_sky* k = new cs_sky();
    k->c = c; k->s = *s;
    cs_frame_state_init(&k->fd);
    for (auto& l : k->luts) l.assign((size_t)CS_SKY_LUT_W * CS_SKY_LUT_H * 4, 0);
    sky_alloc(k);
    *out = k;
    return CS_OK;
}
void cs_sky_destroy(cs_sky* k) { delete k; }
int cs_sky_set_settings(cs_sky* k, const cs_sky_settings* s) {
    if (!k || !s) return CS_ERR_INVALID;
    if (s->frames_to_update < 1) return fail(k->c, CS_ERR_INVALID, "frames_to_update must be >= 1");
    bool rebuild = s->texture_size != k->s.texture_size || s->frames_to_update != k->s.frames_to_update;
    k->s = *s;
    if (rebuild) {  // cloud_sky.gd:37-50,197-212
        k->frame = 0; k->tex_update = 0; k->tex_from = 1; k->tex_to = 2; k->pos[0] = k->pos[1] = 0;
        sky_alloc(k);
        k->full_init = true;
    }
    return CS_OK;
}
int cs_sky_set_sun(cs_sky* k, const float b[9], float e, const float col[3]) {
    if (!k || !b || !col) return CS_ERR_INVALID;
    memcpy(k->basis, b, 36); k->energy = e; memcpy(k->color, col, 12);
    if (!k->sun_attached) k->full_init = true;  // sun.gd:11-13
    k->sun_attached = true;
    return CS_OK;
}
int cs_sky_update(cs_sky* k, float now) { return k ? sky_tick(k, now) : CS_ERR_INVALID; }
int cs_sky_get_frame(cs_sky* k, cs_sky_frame* o) {
    if (!k || !o) return CS_ERR_INVALID;
    memset(o, 0, sizeof(*o));
    o->frame = k->frame; o->frames_to_update = k->s.frames_to_update; o->texture_size = k->size;
    o->update_position[0] = k->pos[0]; o->update_position[1] = k->pos[1];
    o->update_region_size = k->region; o->num_workgroups = k->groups;
    o->texture_to_update = k->tex_update; o->texture_to_blend_from = k->tex_from; o->texture_to_blend_to = k->tex_to;
    o->blend_amount = k->blend;
    o->sky_current_texture = k->lut_current; o->sky_blend_from = k->lut_current; o->sky_blend_to = (k->lut_current + 1) % 3;
    o->sky_updates = k->lut_updates;
    for (int i = 0; i < 3; i++) { o->cloud_textures[i] = k->textures[i].data(); o->sky_luts[i] = k->luts[i].data(); }
    o->frame_data = k->fd;
    return CS_OK;
}
int cs_sky_read_texture(cs_sky* k, int i, uint16_t* out, size_t bytes) {
    if (!k || !out || i < 0 || i > 2 || bytes != k->textures[i].size() * 2) return CS_ERR_INVALID;
    memcpy(out, k->textures[i].data(), bytes);
    return CS_OK;
}

// ---- presentation composite: clouds.gdshader -----------------------------------------------------------------
}  // extern "C" (reopened below)

namespace {
const float GD_PI = 3.14159265358979323846f;  // Godot shading language PI

// bilinear CLAMP_TO_EDGE fetch of an RGBA16F texture (filter_linear, repeat_disable: clouds.gdshader:4-10)
V4 sample_half4_clamp(const uint16_t* t, int w, int h, float u, float v) { return sample_lut(t, w, h, u, v); }

// clouds.gdshader:15-32
V2 gd_vec3_to_oct(V3 e) {
    float s = fabsf(e.x) + fabsf(e.y) + fabsf(e.z);
    e = e / s;
    if (!(e.z >= 0.0f)) {
        float sx = e.x >= 0.0f ? 1.0f : -1.0f, sy = e.y >= 0.0f ? 1.0f : -1.0f;
        float wx = (1.0f - fabsf(e.y)) * sx, wy = (1.0f - fabsf(e.x)) * sy;
        e.x = wx; e.y = wy;
    }
    V2 n;
    n.y = e.y * 0.5f + 0.5f;
    n.x = e.x * 0.5f + n.y;
    n.y = e.x * -0.5f + n.y;
    return n;
}
// clouds.gdshader:34-45
V3 gd_sky_lut(const uint16_t* from, const uint16_t* to, float blend, V3 d) {
    float phi = atan2f(d.z, d.x), theta = asinf(d.y);
    float u = (phi / GD_PI * 0.5f + 0.5f);
    float v = sqrtf(fabsf(theta) / (GD_PI * 0.5f)) * signf(theta) * 0.5f + 0.5f;
    V4 a = sample_half4_clamp(from, CS_SKY_LUT_W, CS_SKY_LUT_H, u, v), b = sample_half4_clamp(to, CS_SKY_LUT_W, CS_SKY_LUT_H, u, v);
    V3 m = mix3({a.x, a.y, a.z}, {b.x, b.y, b.z}, blend);
    return m / 50.0f;
}
// clouds.gdshader:48-59
float gd_sun_with_bloom(V3 ray, V3 sun, float disk_scale) {
    float sunSolidAngle = disk_scale * 0.53f * GD_PI / 180.0f;
    float minSunCosTheta = cosf(sunSolidAngle);
    float cosTheta = dot3(ray, sun);
    if (cosTheta >= minSunCosTheta) return 1.0f;
    float offset = minSunCosTheta - cosTheta;
    float gaussianBloom = expf(-offset * 50000.0f) * 0.5f;
    float invBloom = 1.0f / (0.02f + offset * 300.0f) * 0.01f;
    return gaussianBloom + invBloom;
}
// clouds.gdshader:61-71
float gd_ray_intersect_sphere(V3 ro, V3 rd, float rad) {
    float b = dot3(ro, rd);
    float c = dot3(ro, ro) - rad * rad;
    if (c > 0.0f && b > 0.0f) return -1.0f;
    float discr = b * b - c;
    if (discr < 0.0f) return -1.0f;
    if (discr > b * b) return (-b + sqrtf(discr));
    return -b - sqrtf(discr);
}
// clouds.gdshader:87-102
V3 gd_get_atmo(const uint16_t* sky_from, const uint16_t* sky_to, const uint16_t* tlut, int tlut_param, float blend, V3 dir, V3 sun, float disk_scale) {
    const float groundRadiusMM = 6.360f, atmosphereRadiusMM = 6.460f;
    const V3 viewPos = {0.0f, groundRadiusMM + 0.0002f, 0.0f};
    V3 col = gd_sky_lut(sky_from, sky_to, blend, dir);
    float sl = smoothstepf(0.002f, 1.0f, gd_sun_with_bloom(dir, sun, disk_scale));
    V3 sunLum = {sl, sl, sl};
    if (length3(sunLum) > 0.0f) {
        if (gd_ray_intersect_sphere(viewPos, dir, groundRadiusMM) >= 0.0f) {
            sunLum = sunLum * 0.0f;
        } else {  // getValFromTLUT (:77-85)
            float height = length3(viewPos);
            V3 up = viewPos / height;
            float c = dot3(up, sun);
            float u = 256.0f * clampf(0.5f + 0.5f * c, 0.0f, 1.0f) / 256.0f;
            float v = 64.0f * fmaxf(0.0f, fminf(1.0f, (height - groundRadiusMM) / (atmosphereRadiusMM - groundRadiusMM))) / 64.0f;
            V4 t = sample_half4_clamp(tlut, CS_TRANSMITTANCE_W, CS_TRANSMITTANCE_H, u, v);
            if (tlut_param == CS_TLUT_BRUNETON2017) t = transmittance_from_lut(tlut, c, v, tlut_param);  // same (mu, normalised altitude), other mapping
            sunLum = sunLum * V3{t.x, t.y, t.z};
        }
    }
    return col + sunLum;
}
V3 view_direction(const cs_view& vw, int x, int y) {
    if (vw.projection == CS_VIEW_EQUIRECT) {
        float a = ((float)x + 0.5f) / (float)vw.width * (2.0f * GD_PI) - GD_PI;
        float e = GD_PI * 0.5f - ((float)y + 0.5f) / (float)vw.height * GD_PI;
        return {sinf(a) * cosf(e), sinf(e), -cosf(a) * cosf(e)};
    }
    float th = tanf(vw.fov_y_degrees * (GD_PI / 180.0f) * 0.5f);
    float nx = (((float)x + 0.5f) / (float)vw.width * 2.0f - 1.0f) * th * ((float)vw.width / (float)vw.height);
    float ny = (1.0f - ((float)y + 0.5f) / (float)vw.height * 2.0f) * th;
    const float* b = vw.basis_columns;
    V3 d = {b[0] * nx + b[3] * ny - b[6], b[1] * nx + b[4] * ny - b[7], b[2] * nx + b[5] * ny - b[8]};
    return normalize3(d);
}
// sky() (clouds.gdshader:104-116)
V3 gd_sky_pixel(const cs_view& vw, const uint16_t* cf, const uint16_t* ct, int tw, int th, const uint16_t* sf, const uint16_t* st,
                const uint16_t* tlut, int tlut_param, V3 eyedir) {
    V3 norm = eyedir;
    norm.y = fmaxf(0.0f, norm.y);
    norm = normalize3(norm);
    V2 uv = gd_vec3_to_oct({norm.x, norm.z, norm.y});
    V4 a = sample_half4_clamp(cf, tw, th, uv.x, uv.y), b = sample_half4_clamp(ct, tw, th, uv.x, uv.y);
    float k = vw.blend_amount;
    V4 clouds = {mixf(a.x, b.x, k), mixf(a.y, b.y, k), mixf(a.z, b.z, k), mixf(a.w, b.w, k)};
    V3 sun = {vw.sun_direction[0], vw.sun_direction[1], vw.sun_direction[2]};
    V3 background = gd_get_atmo(sf, st, tlut, tlut_param, k, eyedir, sun, vw.sun_disk_scale);
    V3 color = background * (1.0f - clouds.w) + V3{clouds.x, clouds.y, clouds.z};
    float f = smoothstepf(0.6f, 1.0f, 1.0f - eyedir.y);
    auto cl = [](float v) { return clampf(v, 0.0f, 100.0f); };
    return mix3({cl(color.x), cl(color.y), cl(color.z)}, {cl(background.x), cl(background.y), cl(background.z)}, f);
}
}  // namespace

extern "C" {

int cs_composite(cs_context* c, const cs_view* vw, const void* cf, const void* ct, int tw, int th, const void* sf, const void* st, float* out) {
    if (!c || !vw || !cf || !ct || !sf || !st || !out) return CS_ERR_INVALID;
    if (!c->have_tlut) return fail(c, CS_ERR_NOT_READY, "cs_composite: build the transmittance LUT first");
    if (vw->width < 1 || vw->height < 1 || tw < 1 || th < 1 || (vw->projection != CS_VIEW_EQUIRECT && vw->projection != CS_VIEW_PERSPECTIVE))
        return fail(c, CS_ERR_INVALID, "cs_composite: bad view");
    parallel_rows(c->threads, vw->height, [&](int y, int) {
        for (int x = 0; x < vw->width; x++) {
            V3 col = gd_sky_pixel(*vw, (const uint16_t*)cf, (const uint16_t*)ct, tw, th, (const uint16_t*)sf, (const uint16_t*)st, c->tlut.data(), c->tlut_param, view_direction(*vw, x, y));
            float* o = out + ((size_t)y * vw->width + x) * 4;
            o[0] = col.x; o[1] = col.y; o[2] = col.z; o[3] = 1.0f;
        }
    });
    return CS_OK;
}
int cs_sky_composite_host(cs_sky* k, const cs_view* vw, float* out, size_t bytes) {
    if (!k || !vw || !out) return CS_ERR_INVALID;
    if (bytes != (size_t)vw->width * vw->height * 16) return fail(k->c, CS_ERR_INVALID, "cs_sky_composite_host: bad buffer size");
    cs_view v = *vw;
    v.blend_amount = k->blend;
    return cs_composite(k->c, &v, k->textures[k->tex_from].data(), k->textures[k->tex_to].data(), k->size, k->size,
                        k->luts[k->lut_current].data(), k->luts[(k->lut_current + 1) % 3].data(), out);
}

// ---- oracle-only probes for the known-answer tests (tests/test_oracle_known_answers.py) --------
// Not part of include/cloudsky.h; they expose the internal functions of the restatement so that
// each can be pinned against an analytic answer derived from the shader source.
// transmittance_from_lut (sky-lut.glsl:137-142) through the context's LUT and parametrisation.
int cso_transmittance_lookup(cs_context* c, float cos_theta, float normalized_altitude, float out[4]) {
    if (!c || !c->have_tlut) return CS_ERR_NOT_READY;
    V4 t = transmittance_from_lut(c->tlut.data(), cos_theta, normalized_altitude, c->tlut_param);
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
    return CS_OK;
}
void cso_bruneton_texel_ray(int px, int py, float out[4]) { BrunetonRay b = bruneton_texel_ray(px, py); out[0] = b.altitude; out[1] = b.r; out[2] = b.mu; out[3] = b.d; }
void cso_bruneton_lookup_coords(float normalized_altitude, float mu, float out[3]) { V3 c = bruneton_lookup_coords(normalized_altitude, mu); out[0] = c.x; out[1] = c.y; out[2] = c.z; }
// Study hook, oracle only (tests/hierarchical_study.py): stride 0 = the reference's fixed-step march.
int cso_set_hierarchical(cs_context* c, int stride, float margin, int lod_bias) {
    if (!c || stride < 0 || stride == 1 || stride > 32 || !(margin >= 0.0f) || margin > 1.0f || lod_bias < -8 || lod_bias > 8) return CS_ERR_INVALID;
    c->hier_stride = stride; c->hier_margin = margin; c->hier_lod_bias = lod_bias; return CS_OK;
}
int cso_set_study_variant(cs_context* c, int flags) { if (!c) return CS_ERR_INVALID; c->study_variant = flags; return CS_OK; }
float cso_intersect_sphere(const float dir[3], float r) { return intersectSphere({0.0f, g_radius, 0.0f}, {dir[0], dir[1], dir[2]}, r); }
float cso_hash(const float p[3]) { return hash3({p[0], p[1], p[2]}); }
float cso_henyey_greenstein(float c, float g) { return henyey_greenstein(c, g); }
float cso_remap(float v, float a, float b, float c, float d) { return remap(v, a, b, c, d); }
float cso_height_fraction(float r) { return GetHeightFractionForPoint(r); }
float cso_density_height_gradient(float hf, float type) { return densityHeightGradient(hf, type); }
void cso_oct_to_dir(float u, float v, float out[3]) { V3 n = oct_to_vec3({u, v}); out[0] = n.x; out[1] = n.z; out[2] = n.y; }
uint16_t cso_f32_to_f16(float f) { return f32_to_f16(f); }
float cso_f16_to_f32(uint16_t h) { return f16_to_f32(h); }
int cso_sample_volume(cs_context* c, int which, const float s[3], float lod, float out[4]) {
    if (!c || !c->have_tex) return CS_ERR_NOT_READY;
    V4 v = sample_volume(which == 0 ? c->large : c->small, {s[0], s[1], s[2]}, lod);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    return CS_OK;
}
int cso_sample_weather(cs_context* c, float u, float v, float out[3]) {
    if (!c || !c->have_tex) return CS_ERR_NOT_READY;
    V3 w = sample_weather(c->weather, {u, v});
    out[0] = w.x; out[1] = w.y; out[2] = w.z;
    return CS_OK;
}
int cso_sample_lut(cs_context* c, int which, float u, float v, float out[4]) {
    if (!c) return CS_ERR_INVALID;
    V4 t = which == 0 ? sample_lut(c->tlut.data(), CS_TRANSMITTANCE_W, CS_TRANSMITTANCE_H, u, v) : sample_lut(c->skylut.data(), CS_SKY_LUT_W, CS_SKY_LUT_H, u, v);
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
    return CS_OK;
}
float cso_density(cs_context* c, const cs_cloud_params* P, const float p[3], const float weather[3], float mip) {
    if (!c || !c->have_tex) return -1.0f;
    CloudCtx cc{&c->large, &c->small, &c->weather, c->skylut.data(), *P, c->primary_steps, c->cone_samples};
    Tally tl;
    return density(cc, {p[0], p[1], p[2]}, {weather[0], weather[1], weather[2]}, mip, tl);
}
// start distance, end distance, step length and first marched position for a view direction (sky(), clouds.glsl:218-237)
void cso_ray_setup(const float dir[3], int steps, float out[6]) {
    V3 d = {dir[0], dir[1], dir[2]}, cam = {0.0f, g_radius, 0.0f};
    float t0 = intersectSphere(cam, d, sky_b_radius), t1 = intersectSphere(cam, d, sky_t_radius);
    V3 start = cam + d * t0, end = cam + d * t1;
    float shell = length3(end - start);
    V3 rs = (d * shell) / (float)steps;
    out[0] = t0; out[1] = t1; out[2] = shell; out[3] = length3(rs); out[4] = hash3(start * 10.0f); out[5] = length3(start);
}

}  // extern "C"

// Transmittance-LUT parametrisations.  CS_TLUT_LINEAR is the reference's (transmittance-lut.glsl:161-168: u -> cosine of the
// sun zenith angle, v -> altitude, both linear).  CS_TLUT_BRUNETON2017 is the second TODO of the reference's README
// (README.md:29 "Use the transmittance LUT parametrization from Bruneton (2017)"): E. Bruneton, "Precomputed Atmospheric
// Scattering: a New Implementation" (2017) maps (r, mu) to
//     x_r  = rho / H,                     rho = sqrt(r^2 - Rg^2),  H = sqrt(Rt^2 - Rg^2)
//     x_mu = (d - d_min) / (d_max - d_min), d = distance to the top boundary, d_min = Rt - r, d_max = rho + H
// with texel centres at the ends of the unit range, which spends the texels near the horizon where the transmittance
// varies fastest; only rays that miss the ground are stored, the planet's shadow is applied at lookup time with the
// paper's smoothstep over the sun's angular radius.
// The formulas below use the cancellation-free forms (r^2 - Rg^2 = h (2 Rg + h) with h the altitude, and so on) because
// everything is fp32 at |r| ~ 6.4e3.
//
// Shared by lut_kernels.cu / composite.cu and — CPU test harness only — tests/tlut_host_check.cpp.
#pragma once
#include <math.h>

#include "../../include/cloudsky.h"

#if defined(__CUDACC__)
#define TL_HD __host__ __device__ __forceinline__
#else
#define TL_HD inline
#endif

namespace tl {

constexpr float kRg = 6371.0f, kThickness = 100.0f, kRt = kRg + kThickness;  // transmittance-lut.glsl:50-52
constexpr float kH2 = kThickness * (kRt + kRg);                               // Rt^2 - Rg^2 = 1284200, exact
constexpr float kSunAngularRadius = 0.53f * 3.14159265358979f / 180.0f * 0.5f;  // half of the sun disc of clouds.gdshader:49

// Texel (px, py) of a W x H LUT -> the ray it stores: start radius r (as altitude h = r - Rg and r), cosine mu of the
// zenith angle, and the distance d to the top atmosphere boundary.
TL_HD void bruneton_ray_from_texel(int px, int py, float& h, float& r, float& mu, float& d) {
    const float H = sqrtf(kH2);
    float x_mu = (float)px / (float)(CS_TRANSMITTANCE_W - 1), x_r = (float)py / (float)(CS_TRANSMITTANCE_H - 1);
    float rho = H * x_r;
    r = sqrtf(rho * rho + kRg * kRg);
    h = (rho * rho) / (r + kRg);
    float d_min = kThickness - h, d_max = rho + H;
    d = d_min + x_mu * (d_max - d_min);
    mu = d == 0.0f ? 1.0f : (kH2 - rho * rho - d * d) / (2.0f * r * d);
    mu = fminf(fmaxf(mu, -1.0f), 1.0f);
}

// (normalised altitude in [0,1], mu) -> normalised texture coordinates of the stored ray, and the fraction of the sun's
// disc above the horizon (0 when the planet hides it).
TL_HD void bruneton_uv(float normalized_altitude, float mu, float& u, float& v, float& visible) {
    const float H = sqrtf(kH2);
    float h = fminf(fmaxf(normalized_altitude, 0.0f), 1.0f) * kThickness;
    float r = kRg + h;
    float rho = sqrtf(h * (2.0f * kRg + h));
    float d_min = kThickness - h, d_max = rho + H;
    float rmu = r * mu;
    float disc = rmu * rmu + d_min * (kRt + r);
    float d = fmaxf(sqrtf(fmaxf(disc, 0.0f)) - rmu, 0.0f);
    float x_mu = fminf(fmaxf((d - d_min) / (d_max - d_min), 0.0f), 1.0f), x_r = rho / H;
    u = 0.5f / (float)CS_TRANSMITTANCE_W + x_mu * (1.0f - 1.0f / (float)CS_TRANSMITTANCE_W);
    v = 0.5f / (float)CS_TRANSMITTANCE_H + x_r * (1.0f - 1.0f / (float)CS_TRANSMITTANCE_H);
    float sin_h = kRg / r, cos_h = -(rho / r);  // horizon: sin = Rg / r, cos = -sqrt(1 - sin^2) = -rho / r
    float e0 = -sin_h * kSunAngularRadius, e1 = sin_h * kSunAngularRadius;
    float t = fminf(fmaxf(((mu - cos_h) - e0) / (e1 - e0), 0.0f), 1.0f);
    visible = t * t * (3.0f - 2.0f * t);
}

}  // namespace tl

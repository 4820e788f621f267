// Tileable cloud-noise synthesis: the arithmetic of one texel, shared by the CUDA generator kernel (noise_gen.cu) and by
// the CPU-side harness tests/noise_host_check.cpp, which compiles this header with g++ to check the arithmetic against
// the oracle without a GPU.  The product only ever runs it inside noise_kernel.
//
// The reference ships its noise as bitmaps and lists the generator as a TODO (README.md:30 "Implement a noise generator
// so custom noise can be created and tweaked"); SURVEY 8(f)-3.  The recipe is the published one the reference's textures
// follow (A. Schneider, "Real-time volumetric cloudscapes", GPU Pro 7: a 128^3 RGBA volume with Perlin-Worley in R and
// Worley fBm at three rising frequencies in G, B, A; a 32^3 RGB volume of Worley fBm; a 2-D weather map with cloud type
// in R and coverage in B — perlworlnoise.tga, worlnoise.bmp, weather.bmp and their uses at clouds.glsl:117-133).
//
// Everything is built from an integer lattice hash and fp32 + - * / sqrt floor only, evaluated in a fixed order with no
// FMA contraction (the CUDA TU is compiled --fmad=false, the CPU sides -ffp-contract=off), so every backend produces the
// same bytes.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/cloudsky.h"

#if defined(__CUDACC__)
#define NZ_HD __host__ __device__ __forceinline__
#else
#define NZ_HD inline
#endif

namespace nz {

// 32-bit finaliser ("lowbias32") chained over the wrapped lattice coordinates and the seed.
NZ_HD uint32_t mix(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352du;
    h ^= h >> 15; h *= 0x846ca68bu;
    h ^= h >> 16;
    return h;
}
NZ_HD uint32_t lattice_hash(int x, int y, int z, uint32_t seed) { return mix((uint32_t)x + mix((uint32_t)y + mix((uint32_t)z + mix(seed)))); }
NZ_HD float unit24(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }  // [0,1), exact
NZ_HD int wrap1(int i, int f) { return i < 0 ? i + f : (i >= f ? i - f : i); }     // i in [-1, f]

// Inverted Worley F1: one feature point per cell of an f^3 lattice that repeats with period 1; 1 - min(scale * distance, 1)
// with the distance in cell units.  p in [0,1)^3.
NZ_HD float worley(float px, float py, float pz, int f, float scale, uint32_t seed) {
    float qx = px * (float)f, qy = py * (float)f, qz = pz * (float)f;
    float cx = floorf(qx), cy = floorf(qy), cz = floorf(qz);
    float fx = qx - cx, fy = qy - cy, fz = qz - cz;
    int ix = (int)cx, iy = (int)cy, iz = (int)cz;
    float best = 1.0e30f;
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                uint32_t h0 = lattice_hash(wrap1(ix + dx, f), wrap1(iy + dy, f), wrap1(iz + dz, f), seed);
                uint32_t h1 = mix(h0 + 0x9e3779b9u), h2 = mix(h1 + 0x9e3779b9u);
                float rx = ((float)dx + unit24(h0)) - fx, ry = ((float)dy + unit24(h1)) - fy, rz = ((float)dz + unit24(h2)) - fz;
                float d2 = (rx * rx + ry * ry) + rz * rz;
                best = fminf(best, d2);
            }
    return 1.0f - fminf(sqrtf(best) * scale, 1.0f);
}

// Gradient noise on an f^3 lattice with period 1 (12 edge gradients picked by the hash, quintic fade), roughly [-1,1].
NZ_HD float grad_dot(uint32_t h, float x, float y, float z) {
    h &= 15u;
    float u = h < 8u ? x : y;
    float v = h < 4u ? y : ((h == 12u || h == 14u) ? x : z);
    return ((h & 1u) ? -u : u) + ((h & 2u) ? -v : v);
}
NZ_HD float fade(float t) { return t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f); }
NZ_HD float lerp(float a, float b, float w) { return a + (b - a) * w; }
NZ_HD float perlin(float px, float py, float pz, int f, uint32_t seed) {
    float qx = px * (float)f, qy = py * (float)f, qz = pz * (float)f;
    float cx = floorf(qx), cy = floorf(qy), cz = floorf(qz);
    float tx = qx - cx, ty = qy - cy, tz = qz - cz;
    int x0 = (int)cx, y0 = (int)cy, z0 = (int)cz;
    int x1 = wrap1(x0 + 1, f), y1 = wrap1(y0 + 1, f), z1 = wrap1(z0 + 1, f);
    float g000 = grad_dot(lattice_hash(x0, y0, z0, seed), tx, ty, tz);
    float g100 = grad_dot(lattice_hash(x1, y0, z0, seed), tx - 1.0f, ty, tz);
    float g010 = grad_dot(lattice_hash(x0, y1, z0, seed), tx, ty - 1.0f, tz);
    float g110 = grad_dot(lattice_hash(x1, y1, z0, seed), tx - 1.0f, ty - 1.0f, tz);
    float g001 = grad_dot(lattice_hash(x0, y0, z1, seed), tx, ty, tz - 1.0f);
    float g101 = grad_dot(lattice_hash(x1, y0, z1, seed), tx - 1.0f, ty, tz - 1.0f);
    float g011 = grad_dot(lattice_hash(x0, y1, z1, seed), tx, ty - 1.0f, tz - 1.0f);
    float g111 = grad_dot(lattice_hash(x1, y1, z1, seed), tx - 1.0f, ty - 1.0f, tz - 1.0f);
    float u = fade(tx), v = fade(ty), w = fade(tz);
    float a = lerp(lerp(g000, g100, u), lerp(g010, g110, u), v);
    float b = lerp(lerp(g001, g101, u), lerp(g011, g111, u), v);
    return lerp(a, b, w);
}
// fBm: amplitudes 1/2, 1/4, ... over lattice frequencies f, 2f, 4f, ...; octave o is seeded with seed + o.
NZ_HD float perlin_fbm(float px, float py, float pz, int f, int octaves, uint32_t seed) {
    float sum = 0.0f, amp = 0.5f;
    for (int o = 0; o < octaves; o++) {
        sum = sum + amp * perlin(px, py, pz, f << o, seed + (uint32_t)o);
        amp = amp * 0.5f;
    }
    return sum;
}
NZ_HD float sat(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
NZ_HD uint8_t unorm8(float v) { return (uint8_t)(int)(sat(v) * 255.0f + 0.5f); }
// The shader's own fBm weights (clouds.glsl:118,133).
NZ_HD float fbm3(float a, float b, float c) { return (a * 0.625f + b * 0.25f) + c * 0.125f; }

constexpr uint32_t kPerlinSeedXor = 0x5bd1e995u, kTypeSeedXor = 0x2545f491u;

// One texel of each product.  (x, y, z) are texel indices, n the edge: p = (index + 0.5) / n.
// Worley octave k (lattice worley_frequency << k) is seeded with seed + 101 k, so the channels that share an octave share
// its feature points, as in the published recipe.
NZ_HD void large_texel(const cs_noise_params& P, int n, int x, int y, int z, uint8_t out[4]) {
    float inv = 1.0f / (float)n;
    float px = ((float)x + 0.5f) * inv, py = ((float)y + 0.5f) * inv, pz = ((float)z + 0.5f) * inv;
    float w[5];
    for (int k = 0; k < 5; k++) w[k] = worley(px, py, pz, P.worley_frequency << k, P.worley_scale, P.seed + 101u * (uint32_t)k);
    float g = fbm3(w[0], w[1], w[2]), b = fbm3(w[1], w[2], w[3]), a = fbm3(w[2], w[3], w[4]);
    float p01 = sat(0.5f + P.perlin_scale * perlin_fbm(px, py, pz, P.perlin_frequency, P.perlin_octaves, P.seed ^ kPerlinSeedXor));
    float r = g + p01 * (1.0f - g);  // remap(perlin, 0, 1, low-frequency Worley fBm, 1): "Perlin-Worley"
    out[0] = unorm8(r); out[1] = unorm8(g); out[2] = unorm8(b); out[3] = unorm8(a);
}
NZ_HD void small_texel(const cs_noise_params& P, int n, int x, int y, int z, uint8_t out[4]) {
    float inv = 1.0f / (float)n;
    float px = ((float)x + 0.5f) * inv, py = ((float)y + 0.5f) * inv, pz = ((float)z + 0.5f) * inv;
    float w[5];
    for (int k = 0; k < 5; k++) w[k] = worley(px, py, pz, P.worley_frequency << k, P.worley_scale, P.seed + 101u * (uint32_t)k);
    out[0] = unorm8(fbm3(w[0], w[1], w[2])); out[1] = unorm8(fbm3(w[1], w[2], w[3])); out[2] = unorm8(fbm3(w[2], w[3], w[4])); out[3] = 255;
}
// Weather map (2-D: the z = 0 slice of the 3-D functions).  B = coverage: Perlin-Worley pushed through
// (v - remap_lo) / (remap_hi - remap_lo); R = cloud type in [type_lo, type_hi] from a low-frequency Perlin fBm; G = 0
// (clouds.glsl reads .x and .z only: clouds.glsl:121,123).
NZ_HD void weather_texel(const cs_noise_params& P, int n, int x, int y, uint8_t out[4]) {
    float inv = 1.0f / (float)n;
    float px = ((float)x + 0.5f) * inv, py = ((float)y + 0.5f) * inv;
    float w0 = worley(px, py, 0.0f, P.worley_frequency, P.worley_scale, P.seed), w1 = worley(px, py, 0.0f, P.worley_frequency << 1, P.worley_scale, P.seed + 101u),
          w2 = worley(px, py, 0.0f, P.worley_frequency << 2, P.worley_scale, P.seed + 202u);
    float wf = fbm3(w0, w1, w2);
    float p01 = sat(0.5f + P.perlin_scale * perlin_fbm(px, py, 0.0f, P.perlin_frequency, P.perlin_octaves, P.seed ^ kPerlinSeedXor));
    float pw = wf + p01 * (1.0f - wf);
    float coverage = sat((pw - P.remap_lo) / (P.remap_hi - P.remap_lo));
    float t01 = sat(0.5f + P.perlin_scale * perlin_fbm(px, py, 0.0f, 2, 3, P.seed ^ kTypeSeedXor));
    float type = P.type_lo + (P.type_hi - P.type_lo) * t01;
    out[0] = unorm8(type); out[1] = 0; out[2] = unorm8(coverage); out[3] = 255;
}

// Shared validation (both backends reject the same inputs).  Returns nullptr when the request is fine.
inline const char* check_request(int kind, int n, const cs_noise_params* P) {
    if (kind < CS_NOISE_LARGE || kind > CS_NOISE_WEATHER) return "cs_generate_noise: kind must be CS_NOISE_LARGE, _SMALL or _WEATHER";
    if (!P) return "cs_generate_noise: params missing";
    if (n < 1 || n > (kind == CS_NOISE_WEATHER ? 8192 : 512) || (n & (n - 1))) return "cs_generate_noise: n must be a power of two (<= 512 for volumes, <= 8192 for the weather map)";
    if (P->worley_frequency < 1 || P->worley_frequency > 256 || P->perlin_frequency < 1 || P->perlin_frequency > 256 || P->perlin_octaves < 1 || P->perlin_octaves > 8)
        return "cs_generate_noise: frequencies in [1,256], perlin_octaves in [1,8]";
    if (!(P->remap_hi > P->remap_lo) || !(P->type_hi >= P->type_lo) || !(P->perlin_scale >= 0.0f) || !(P->worley_scale > 0.0f))
        return "cs_generate_noise: need remap_hi > remap_lo, type_hi >= type_lo, perlin_scale >= 0, worley_scale > 0";
    return nullptr;
}

}  // namespace nz

// libcloudsky_b200.so — context management and the C-ABI of include/cloudsky.h on top of the
// sm_100a kernels (lut_kernels.cu, clouds_strict.cu, clouds_fast.cu).
//
// HBM layout per context (everything stays resident; the part of it a frame touches, ~40 MiB, lives in
// L2 after the first frame):
//   large volume : RGBA8 mip chain 128^3 .. 1^3 (9.14 MiB, strict kernel) + coefficient records (fast kernel): fp16 32 B/texel
//                  = 73 MiB when exactly representable, else fp32 64 B/texel = 146 MiB
//   small volume : RGBA8 mip chain 32^3 .. 1^3 (146 KiB)                  + records 16 / 32 B per texel (0.57 / 1.14 MiB)
//   weather map  : RGBA8 512^2 (1 MiB)                                    + records 16 / 32 B per texel (4 / 8 MiB)
//   transmittance LUT 256x64 half4, sky LUT 200x100 half4, FrameConsts (64 B)
//   output image : W*H half4, tightly packed, row 0 = uv.y 0
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <new>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cs_context.h"
#include "noise_core.h"

using namespace cs;

namespace {

int fail(cs_context* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}
int cuda_fail(cs_context* c, cudaError_t e, const char* what) {
    return fail(c, CS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                                     \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return cuda_fail(c, e__, #call);     \
    } while (0)

int bind(cs_context* c) {
    CU(cudaSetDevice(c->device));
    return CS_OK;
}

void free_textures(cs_context* c) {
    for (auto& p : c->d_large) { if (p) cudaFree(p); p = nullptr; }
    for (auto& p : c->d_small) { if (p) cudaFree(p); p = nullptr; }
    for (auto& p : c->d_large_f) { if (p) cudaFree(p); p = nullptr; }
    for (auto& p : c->d_small_f) { if (p) cudaFree(p); p = nullptr; }
    for (auto& p : c->d_large_h2) { if (p) cudaFree(p); p = nullptr; }
    for (auto& p : c->d_small_h2) { if (p) cudaFree(p); p = nullptr; }
    if (c->d_weather) cudaFree(c->d_weather);
    if (c->d_weather_f) cudaFree(c->d_weather_f);
    if (c->d_weather_h2) cudaFree(c->d_weather_h2);
    c->d_weather = nullptr; c->d_weather_f = nullptr; c->d_weather_h2 = nullptr;
    c->have_h2 = false;
    if (c->t_large) cudaDestroyTextureObject(c->t_large);
    if (c->t_small) cudaDestroyTextureObject(c->t_small);
    if (c->t_weather) cudaDestroyTextureObject(c->t_weather);
    if (c->a_large) cudaFreeMipmappedArray(c->a_large);
    if (c->a_small) cudaFreeMipmappedArray(c->a_small);
    if (c->a_weather) cudaFreeArray(c->a_weather);
    c->t_large = c->t_small = c->t_weather = 0;
    c->a_large = c->a_small = nullptr; c->a_weather = nullptr;
    c->have_tex = false;
}

// CS_MODE_TEX: the RGBA8 mip chain of a volume as a CUDA mipmapped 3D array behind a texture object with the reference's
// sampler state (REPEAT, linear filter inside a level; the kernel names the level itself, like textureLod()).
cudaError_t make_volume_texture(const std::vector<std::vector<uint8_t>>& levels, int n, cudaMipmappedArray_t* arr, cudaTextureObject_t* tex) {
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<uchar4>();
    cudaError_t e = cudaMallocMipmappedArray(arr, &fd, make_cudaExtent(n, n, n), (unsigned)levels.size(), 0);
    if (e != cudaSuccess) return e;
    for (size_t l = 0; l < levels.size(); l++) {
        int nl = n >> l;
        cudaArray_t lvl = nullptr;
        if ((e = cudaGetMipmappedArrayLevel(&lvl, *arr, (unsigned)l)) != cudaSuccess) return e;
        cudaMemcpy3DParms cp = {};
        cp.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(levels[l].data()), (size_t)nl * 4, nl, nl);
        cp.dstArray = lvl;
        cp.extent = make_cudaExtent(nl, nl, nl);
        cp.kind = cudaMemcpyHostToDevice;
        if ((e = cudaMemcpy3D(&cp)) != cudaSuccess) return e;
    }
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeMipmappedArray;
    rd.res.mipmap.mipmap = *arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear;
    td.mipmapFilterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    td.minMipmapLevelClamp = 0.0f;
    td.maxMipmapLevelClamp = (float)(levels.size() - 1);
    return cudaCreateTextureObject(tex, &rd, &td, nullptr);
}
cudaError_t make_weather_texture(const std::vector<uint8_t>& rgba, int w, int h, cudaArray_t* arr, cudaTextureObject_t* tex) {
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<uchar4>();
    cudaError_t e = cudaMallocArray(arr, &fd, w, h, 0);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemcpy2DToArray(*arr, 0, 0, rgba.data(), (size_t)w * 4, (size_t)w * 4, h, cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = *arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    return cudaCreateTextureObject(tex, &rd, &td, nullptr);
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// Interpolation-coefficient records for the fast kernel.  Only the channel combinations clouds.glsl reads
// are kept (fbm = .625G+.25B+.125A, clouds.glsl:118; hfbm = .625R+.25G+.125B, clouds.glsl:133; weather R and
// B, clouds.glsl:121,123).  Each texel (x,y,z) stores the 8 coefficients of the trilinear polynomial over the
// cell [x,x+1]x[y,y+1]x[z,z+1] (REPEAT wrap baked in):
//     v(fx,fy,fz) = c0 + fx c1 + fy (c2 + fx c3) + fz (c4 + fx c5 + fy (c6 + fx c7))
// so one filtered fetch reads ONE aligned record (one 128-byte line), needs a single address, no unpacking
// and 7 FFMAs per channel.  large: 64 B per texel (R coefficients, then fbm), small: 32 B, weather (bilinear,
// 4 coefficients per channel): 32 B.
inline float un8(uint8_t v) { return (float)v / 255.0f; }

template <class F>
void trilinear_coeffs(F v, int n, int x, int y, int z, float* c) {
    int x1 = (x + 1) % n, y1 = (y + 1) % n, z1 = (z + 1) % n;
    float v000 = v(x, y, z), v100 = v(x1, y, z), v010 = v(x, y1, z), v110 = v(x1, y1, z);
    float v001 = v(x, y, z1), v101 = v(x1, y, z1), v011 = v(x, y1, z1), v111 = v(x1, y1, z1);
    c[0] = v000;
    c[1] = v100 - v000;
    c[2] = v010 - v000;
    c[3] = (v110 - v010) - c[1];
    c[4] = v001 - v000;
    c[5] = (v101 - v001) - c[1];
    c[6] = (v011 - v001) - c[2];
    c[7] = ((v111 - v011) - (v101 - v001)) - c[3];
}
void pack_large_f(const std::vector<uint8_t>& rgba, int n, std::vector<float>& out) {
    std::vector<float> R((size_t)n * n * n), K(R.size());
    for (size_t i = 0; i < R.size(); i++) {
        R[i] = un8(rgba[i * 4]);
        K[i] = un8(rgba[i * 4 + 1]) * 0.625f + un8(rgba[i * 4 + 2]) * 0.25f + un8(rgba[i * 4 + 3]) * 0.125f;
    }
    out.resize(R.size() * 16);
    auto fr = [&](int x, int y, int z) { return R[((size_t)z * n + y) * n + x]; };
    auto fk = [&](int x, int y, int z) { return K[((size_t)z * n + y) * n + x]; };
    for (int z = 0; z < n; z++)
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) {
                float* o = &out[(((size_t)z * n + y) * n + x) * 16];
                trilinear_coeffs(fr, n, x, y, z, o);
                trilinear_coeffs(fk, n, x, y, z, o + 8);
            }
}
void pack_small_f(const std::vector<uint8_t>& rgba, int n, std::vector<float>& out) {
    std::vector<float> Hh((size_t)n * n * n);
    for (size_t i = 0; i < Hh.size(); i++) Hh[i] = un8(rgba[i * 4]) * 0.625f + un8(rgba[i * 4 + 1]) * 0.25f + un8(rgba[i * 4 + 2]) * 0.125f;
    out.resize(Hh.size() * 8);
    auto fh = [&](int x, int y, int z) { return Hh[((size_t)z * n + y) * n + x]; };
    for (int z = 0; z < n; z++)
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) trilinear_coeffs(fh, n, x, y, z, &out[(((size_t)z * n + y) * n + x) * 8]);
}
void pack_weather_f(const std::vector<uint8_t>& rgba, int w, int h, std::vector<float>& out) {
    out.resize((size_t)w * h * 8);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int x1 = (x + 1) % w, y1 = (y + 1) % h;
            float* o = &out[((size_t)y * w + x) * 8];
            for (int ch = 0; ch < 2; ch++) {  // ch 0: cloud type (R), ch 1: coverage (B)
                int k = ch == 0 ? 0 : 2;
                float v00 = un8(rgba[((size_t)y * w + x) * 4 + k]), v10 = un8(rgba[((size_t)y * w + x1) * 4 + k]);
                float v01 = un8(rgba[((size_t)y1 * w + x) * 4 + k]), v11 = un8(rgba[((size_t)y1 * w + x1) * 4 + k]);
                o[ch * 4 + 0] = v00; o[ch * 4 + 1] = v10 - v00; o[ch * 4 + 2] = v01 - v00; o[ch * 4 + 3] = (v11 - v01) - (v10 - v00);
            }
        }
}
// ---- compact records: the same coefficients as exact INTEGERS in fp16 ---------------------------------------
// All inputs are 8-bit texels, so R, K = 5G+2B+A, h = 5R+2G+B and the weather channels are integers and so are
// all their interpolation coefficients (differences of differences).  fp16 holds every integer of magnitude
// <= 2048 exactly (and even / multiple-of-4 ones beyond), which covers smooth noise textures such as the
// reference's.  When EVERY coefficient of EVERY level is exactly representable the fast kernel reads these
// half-size records (large 32 B, small 16 B, weather 16 B: half the L1 wavefronts per fetch) and scales the
// interpolated integer by 1/255 or 1/2040 afterwards; otherwise it falls back to the fp32 records above.
inline bool half_exact(int v) {
    int a = v < 0 ? -v : v;
    if (a <= 2048) return true;
    if (a <= 4096) return (a & 1) == 0;
    if (a <= 8192) return (a & 3) == 0;
    if (a <= 16384) return (a & 7) == 0;
    return false;
}
inline uint16_t half_bits_exact(int v) {  // v must satisfy half_exact()
    if (v == 0) return 0;
    uint16_t sign = v < 0 ? 0x8000u : 0u;
    unsigned a = (unsigned)(v < 0 ? -v : v);
    int e = 31 - __builtin_clz(a);
    unsigned mant = e <= 10 ? (a << (10 - e)) : (a >> (e - 10));
    return (uint16_t)(sign | ((unsigned)(e + 15) << 10) | (mant & 0x3ffu));
}
template <class F>
bool trilinear_coeffs_i(F v, int n, int x, int y, int z, uint16_t* c) {
    int x1 = (x + 1) % n, y1 = (y + 1) % n, z1 = (z + 1) % n;
    int v000 = v(x, y, z), v100 = v(x1, y, z), v010 = v(x, y1, z), v110 = v(x1, y1, z);
    int v001 = v(x, y, z1), v101 = v(x1, y, z1), v011 = v(x, y1, z1), v111 = v(x1, y1, z1);
    int k[8];
    k[0] = v000; k[1] = v100 - v000; k[2] = v010 - v000; k[3] = (v110 - v010) - k[1];
    k[4] = v001 - v000; k[5] = (v101 - v001) - k[1]; k[6] = (v011 - v001) - k[2];
    k[7] = ((v111 - v011) - (v101 - v001)) - k[3];
    bool ok = true;
    for (int i = 0; i < 8; i++) { ok = ok && half_exact(k[i]); c[i] = ok ? half_bits_exact(k[i]) : 0; }
    return ok;
}
bool pack_large_h(const std::vector<uint8_t>& rgba, int n, std::vector<uint16_t>& out) {
    out.resize((size_t)n * n * n * 16);
    auto fr = [&](int x, int y, int z) { return (int)rgba[(((size_t)z * n + y) * n + x) * 4]; };
    auto fk = [&](int x, int y, int z) { const uint8_t* t = &rgba[(((size_t)z * n + y) * n + x) * 4]; return 5 * t[1] + 2 * t[2] + t[3]; };
    for (int z = 0; z < n; z++)
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) {
                uint16_t* o = &out[(((size_t)z * n + y) * n + x) * 16];
                if (!trilinear_coeffs_i(fr, n, x, y, z, o) || !trilinear_coeffs_i(fk, n, x, y, z, o + 8)) return false;
            }
    return true;
}
bool pack_small_h(const std::vector<uint8_t>& rgba, int n, std::vector<uint16_t>& out) {
    out.resize((size_t)n * n * n * 8);
    auto fh = [&](int x, int y, int z) { const uint8_t* t = &rgba[(((size_t)z * n + y) * n + x) * 4]; return 5 * t[0] + 2 * t[1] + t[2]; };
    for (int z = 0; z < n; z++)
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++)
                if (!trilinear_coeffs_i(fh, n, x, y, z, &out[(((size_t)z * n + y) * n + x) * 8])) return false;
    return true;
}
bool pack_weather_h(const std::vector<uint8_t>& rgba, int w, int h, std::vector<uint16_t>& out) {
    out.resize((size_t)w * h * 8);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int x1 = (x + 1) % w, y1 = (y + 1) % h;
            uint16_t* o = &out[((size_t)y * w + x) * 8];
            for (int ch = 0; ch < 2; ch++) {
                int k = ch == 0 ? 0 : 2;
                int v00 = rgba[((size_t)y * w + x) * 4 + k], v10 = rgba[((size_t)y * w + x1) * 4 + k];
                int v01 = rgba[((size_t)y1 * w + x) * 4 + k], v11 = rgba[((size_t)y1 * w + x1) * 4 + k];
                int c[4] = {v00, v10 - v00, v01 - v00, (v11 - v01) - (v10 - v00)};  // |c| <= 510: always exact in fp16
                for (int i = 0; i < 4; i++) o[ch * 4 + i] = half_bits_exact(c[i]);
            }
        }
    return true;
}

// ---- CS_MODE_HALF records: the exact-integer coefficients again, arranged for packed-fp16 evaluation ----------------------
// large   : 8 half2 (R_ci, K_ci), i = 0..7, with the constant terms centred (R_c0 - 128, K_c0 - 1020) so that the fp16 roundings
//           of the running sums act on values half as large; the kernel adds the centres back in fp32 for free (FFMA).
// small   : 4 half2 (c0 - 1020, c4), (c1, c5), (c2, c6), (c3, c7): the two z-halves of the polynomial side by side.
// weather : 4 half2 (type_ci, coverage_ci), i = 0..3, constant terms centred by 128.
template <class F>
void trilinear_coeffs_int(F v, int n, int x, int y, int z, int* k) {
    int x1 = (x + 1) % n, y1 = (y + 1) % n, z1 = (z + 1) % n;
    int v000 = v(x, y, z), v100 = v(x1, y, z), v010 = v(x, y1, z), v110 = v(x1, y1, z);
    int v001 = v(x, y, z1), v101 = v(x1, y, z1), v011 = v(x, y1, z1), v111 = v(x1, y1, z1);
    k[0] = v000; k[1] = v100 - v000; k[2] = v010 - v000; k[3] = (v110 - v010) - k[1];
    k[4] = v001 - v000; k[5] = (v101 - v001) - k[1]; k[6] = (v011 - v001) - k[2];
    k[7] = ((v111 - v011) - (v101 - v001)) - k[3];
}
void pack_large_h2(const std::vector<uint8_t>& rgba, int n, std::vector<uint16_t>& out) {
    out.resize((size_t)n * n * n * 16);
    auto fr = [&](int x, int y, int z) { return (int)rgba[(((size_t)z * n + y) * n + x) * 4]; };
    auto fk = [&](int x, int y, int z) { const uint8_t* t = &rgba[(((size_t)z * n + y) * n + x) * 4]; return 5 * t[1] + 2 * t[2] + t[3]; };
    for (int z = 0; z < n; z++)
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) {
                int r[8], k[8];
                trilinear_coeffs_int(fr, n, x, y, z, r);
                trilinear_coeffs_int(fk, n, x, y, z, k);
                r[0] -= 128; k[0] -= 1020;
                uint16_t* o = &out[(((size_t)z * n + y) * n + x) * 16];
                for (int i = 0; i < 8; i++) { o[2 * i] = half_bits_exact(r[i]); o[2 * i + 1] = half_bits_exact(k[i]); }
            }
}
void pack_small_h2(const std::vector<uint8_t>& rgba, int n, std::vector<uint16_t>& out) {
    out.resize((size_t)n * n * n * 8);
    auto fh = [&](int x, int y, int z) { const uint8_t* t = &rgba[(((size_t)z * n + y) * n + x) * 4]; return 5 * t[0] + 2 * t[1] + t[2]; };
    for (int z = 0; z < n; z++)
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) {
                int k[8];
                trilinear_coeffs_int(fh, n, x, y, z, k);
                k[0] -= 1020;
                uint16_t* o = &out[(((size_t)z * n + y) * n + x) * 8];
                for (int i = 0; i < 4; i++) { o[2 * i] = half_bits_exact(k[i]); o[2 * i + 1] = half_bits_exact(k[i + 4]); }
            }
}
void pack_weather_h2(const std::vector<uint8_t>& rgba, int w, int h, std::vector<uint16_t>& out) {
    out.resize((size_t)w * h * 8);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int x1 = (x + 1) % w, y1 = (y + 1) % h;
            int c[2][4];
            for (int ch = 0; ch < 2; ch++) {
                int k = ch == 0 ? 0 : 2;
                int v00 = rgba[((size_t)y * w + x) * 4 + k], v10 = rgba[((size_t)y * w + x1) * 4 + k];
                int v01 = rgba[((size_t)y1 * w + x) * 4 + k], v11 = rgba[((size_t)y1 * w + x1) * 4 + k];
                c[ch][0] = v00 - 128; c[ch][1] = v10 - v00; c[ch][2] = v01 - v00; c[ch][3] = (v11 - v01) - (v10 - v00);
            }
            uint16_t* o = &out[((size_t)y * w + x) * 8];
            for (int i = 0; i < 4; i++) { o[2 * i] = half_bits_exact(c[0][i]); o[2 * i + 1] = half_bits_exact(c[1][i]); }
        }
}

int ilog2(int v) { int s = 0; while ((1 << s) < v) s++; return s; }

int upload_levels(cs_context* c, const std::vector<uint8_t>& large0, int ln, const std::vector<uint8_t>& small0, int sn,
                  const std::vector<uint8_t>& weather, int ww, int wh) {
    const bool force_fp32_records = getenv("CLOUDSKY_FP32_RECORDS") != nullptr;  // development / test knob
    if (!is_pow2(ln) || !is_pow2(sn) || !is_pow2(ww) || !is_pow2(wh))
        return fail(c, CS_ERR_INVALID, "texture dimensions must be powers of two (REPEAT addressing uses masks)");
    // Sizes: the round-down-add floor of the kernel needs |coordinate * edge| < 2^22 at the slab's 6e6 m world coordinates
    // (clouds.glsl:43-45 with the texture scales of :117,132,174): large <= 512, small <= 256, weather <= 8192.
    if (ln > (1 << (kMaxLargeLevels - 1)) || sn > (1 << (kMaxSmallLevels - 1)) || ww > 8192 || wh > 8192)
        return fail(c, CS_ERR_INVALID, "texture too large (large <= 512^3, small <= 256^3, weather <= 8192^2)");
    int r = bind(c);
    if (r) return r;
    CU(cudaStreamSynchronize(c->stream));
    free_textures(c);
    c->h_large.assign(1, large0);
    c->h_small.assign(1, small0);
    build_volume_mips(c->h_large, ln);
    build_volume_mips(c->h_small, sn);
    c->large_n = ln; c->large_levels = (int)c->h_large.size();
    c->small_n = sn; c->small_levels = (int)c->h_small.size();
    c->weather_w = ww; c->weather_h = wh;
    c->h_weather = weather;
    c->weather_type_hi = 1;
    for (size_t i = 0; i < weather.size(); i += 4) if (weather[i] < 128) { c->weather_type_hi = 0; break; }
    for (int l = 0; l < c->large_levels; l++) {
        CU(cudaMalloc(&c->d_large[l], c->h_large[l].size()));
        CU(cudaMemcpy(c->d_large[l], c->h_large[l].data(), c->h_large[l].size(), cudaMemcpyHostToDevice));
    }
    for (int l = 0; l < c->small_levels; l++) {
        CU(cudaMalloc(&c->d_small[l], c->h_small[l].size()));
        CU(cudaMemcpy(c->d_small[l], c->h_small[l].data(), c->h_small[l].size(), cudaMemcpyHostToDevice));
    }
    CU(cudaMalloc(&c->d_weather, weather.size()));
    CU(cudaMemcpy(c->d_weather, weather.data(), weather.size(), cudaMemcpyHostToDevice));

    // interpolation records for the fast kernel: exact-integer fp16 records when representable, else fp32
    std::vector<std::vector<uint16_t>> lh(c->large_levels), sh(c->small_levels);
    std::vector<uint16_t> wh16;
    bool half_ok = force_fp32_records ? false : true;
    for (int l = 0; l < c->large_levels && half_ok; l++) half_ok = pack_large_h(c->h_large[l], ln >> l, lh[l]);
    for (int l = 0; l < c->small_levels && half_ok; l++) half_ok = pack_small_h(c->h_small[l], sn >> l, sh[l]);
    if (half_ok) half_ok = pack_weather_h(weather, ww, wh, wh16);
    // format mask: bit 0 large, bit 1 small, bit 2 weather.  All fp16 when exact, else all fp32 (mixed layouts were
    // measured within 0.2 % of all-fp16, DESIGN.md, and are not instantiated).
    int fmt = half_ok ? 7 : 0;
    c->records_half = fmt;
    auto put = [&](float** dst, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes);
        return e != cudaSuccess ? e : cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    std::vector<float> pk;
    for (int l = 0; l < c->large_levels; l++) {
        if (fmt & 1) { CU(put(&c->d_large_f[l], lh[l].data(), lh[l].size() * 2)); }
        else { pack_large_f(c->h_large[l], ln >> l, pk); CU(put(&c->d_large_f[l], pk.data(), pk.size() * 4)); }
    }
    for (int l = 0; l < c->small_levels; l++) {
        if (fmt & 2) { CU(put(&c->d_small_f[l], sh[l].data(), sh[l].size() * 2)); }
        else { pack_small_f(c->h_small[l], sn >> l, pk); CU(put(&c->d_small_f[l], pk.data(), pk.size() * 4)); }
    }
    if (fmt & 4) { CU(put(&c->d_weather_f, wh16.data(), wh16.size() * 2)); }
    else { pack_weather_f(weather, ww, wh, pk); CU(put(&c->d_weather_f, pk.data(), pk.size() * 4)); }
    CU(make_volume_texture(c->h_large, ln, &c->a_large, &c->t_large));
    CU(make_volume_texture(c->h_small, sn, &c->a_small, &c->t_small));
    CU(make_weather_texture(weather, ww, wh, &c->a_weather, &c->t_weather));
    c->have_tex = true;
    return CS_OK;
}

// CS_MODE_HALF: build the half2 records on first use (another 8 x the texel bytes; only contexts that ask for the mode pay for it).
int ensure_half2_records(cs_context* c) {
    if (c->have_h2) return CS_OK;
    if (!c->have_tex) return fail(c, CS_ERR_NOT_READY, "input textures not uploaded");
    if (c->records_half != 7) return fail(c, CS_ERR_UNSUPPORTED, "CS_MODE_HALF needs textures whose interpolation coefficients are exact in fp16");
    int r = bind(c);
    if (r) return r;
    auto put = [&](float** dst, const std::vector<uint16_t>& v) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, v.size() * 2);
        return e != cudaSuccess ? e : cudaMemcpy(*dst, v.data(), v.size() * 2, cudaMemcpyHostToDevice);
    };
    std::vector<uint16_t> pk;
    for (int l = 0; l < c->large_levels; l++) { pack_large_h2(c->h_large[l], c->large_n >> l, pk); CU(put(&c->d_large_h2[l], pk)); }
    for (int l = 0; l < c->small_levels; l++) { pack_small_h2(c->h_small[l], c->small_n >> l, pk); CU(put(&c->d_small_h2[l], pk)); }
    pack_weather_h2(c->h_weather, c->weather_w, c->weather_h, pk);
    CU(put(&c->d_weather_h2, pk));
    c->have_h2 = true;
    return CS_OK;
}

// img_w/img_h > 0: `out` is a caller-owned image of that size (a cs_sky's own textures); otherwise the context's image size.
int make_launch(cs_context* c, const cs_cloud_params* P, int x0, int y0, int x1, int y1, uint16_t* out, const uint16_t* sky_lut, CloudLaunch& L,
                int img_w = 0, int img_h = 0) {
    const int IW = img_w > 0 ? img_w : c->W, IH = img_h > 0 ? img_h : c->H;
    if (!c->have_tex) return fail(c, CS_ERR_NOT_READY, "input textures not uploaded (can_run == false)");
    if (!sky_lut && !c->have_sky) return fail(c, CS_ERR_NOT_READY, "sky LUT not built (Attempting to render with an uninitialized sky lut)");
    if (IW < 1 || !out) return fail(c, CS_ERR_NOT_READY, "cs_resize not called");
    if ((int)P->texture_size[0] != IW || (int)P->texture_size[1] != IH)
        return fail(c, CS_ERR_INVALID, "params.texture_size does not match the image size set with cs_resize");
    std::memset(&L, 0, sizeof(L));
    L.P = *P;
    L.width = IW; L.height = IH;
    L.x0 = x0 < 0 ? 0 : x0; L.y0 = y0 < 0 ? 0 : y0;
    L.x1 = x1 > IW ? IW : x1; L.y1 = y1 > IH ? IH : y1;
    L.out_pitch_px = IW;
    L.primary_steps = c->primary_steps; L.cone_samples = c->cone_samples;
    L.budget_len = c->budget_len; L.budget_min = c->budget_min;
    // below 2^-12 the remaining radiance is < 1 fp16 ulp and alpha = 1 - T already rounds to 1.0 in fp16
    L.early_out_T = (c->mode & CS_MODE_EARLY_OUT) ? 0.000244140625f : 0.0f;
    L.large_n = c->large_n; L.large_levels = c->large_levels;
    L.small_n = c->small_n; L.small_levels = c->small_levels;
    L.weather_w = c->weather_w; L.weather_h = c->weather_h;
    const bool half2 = (c->mode != CS_MODE_STRICT) && (c->mode & CS_MODE_HALF);
    if (half2 && !c->have_h2) {  // textures were (re)uploaded after the mode was set
        int r = ensure_half2_records(c);
        if (r) return r;
    }
    for (int l = 0; l < kMaxLargeLevels; l++) { L.large[l] = c->d_large[l]; L.large_f[l] = half2 ? c->d_large_h2[l] : c->d_large_f[l]; }
    for (int l = 0; l < kMaxSmallLevels; l++) { L.small[l] = c->d_small[l]; L.small_f[l] = half2 ? c->d_small_h2[l] : c->d_small_f[l]; }
    L.weather = c->d_weather; L.weather_f = half2 ? c->d_weather_h2 : c->d_weather_f;
    L.large_shift = ilog2(c->large_n); L.small_shift = ilog2(c->small_n);
    L.weather_shx = ilog2(c->weather_w); L.weather_shy = ilog2(c->weather_h);
    L.weather_type_hi = c->weather_type_hi;
    L.records_half = half2 ? 16 : c->records_half;
    // texels per metre at level 0 (exact: power-of-two edge times the shader's texture scale, clouds.glsl:117,132)
    L.large_fn0 = (float)c->large_n * 0.00008f; L.small_fn0 = (float)c->small_n * 0.001f;
    L.weather_fw = (float)c->weather_w; L.weather_fh = (float)c->weather_h;
    L.large_mask0 = c->large_n - 1; L.small_mask0 = c->small_n - 1; L.weather_maskx = c->weather_w - 1; L.weather_masky = c->weather_h - 1;
    // The 1^3 tail of the small volume's mip chain: any filtered fetch of it returns that texel's hfbm (clouds.glsl:133).
    L.small_tail_level = -1;
    if ((c->small_n >> (c->small_levels - 1)) == 1) {
        const uint8_t* t = c->h_small[c->small_levels - 1].data();
        L.small_tail_level = c->small_levels - 1;
        if (c->mode & CS_MODE_TEX) L.small_tail_value = fmaf(un8(t[0]), 0.625f, fmaf(un8(t[1]), 0.25f, un8(t[2]) * 0.125f));
        else if (half2) L.small_tail_value = fmaf((float)(5 * t[0] + 2 * t[1] + t[2] - 1020), 1.0f / 2040.0f, 0.5f);
        else if (c->records_half & 2) L.small_tail_value = (float)(5 * t[0] + 2 * t[1] + t[2]) * (1.0f / 2040.0f);
        else L.small_tail_value = un8(t[0]) * 0.625f + un8(t[1]) * 0.25f + un8(t[2]) * 0.125f;
    }
    L.tex_large = c->t_large; L.tex_small = c->t_small; L.tex_weather = c->t_weather;
    L.hw_filter = (c->mode & CS_MODE_TEX) ? 1 : 0;
    L.sky_lut = sky_lut ? sky_lut : c->d_sky;
    L.frame_consts = c->d_frame_consts;
    L.out = out;
    L.counters = c->counters_on ? c->d_counters : nullptr;
    L.n_suns = 1; L.sun_stride_px = 0;
    L.n_mirrors = 0;
    if (c->n_mirrors > 0) {
        const uint8_t* o = reinterpret_cast<const uint8_t*>(out);
        if (o >= c->mirror_base && o < c->mirror_base + c->mirror_bytes) {  // `out` lies in the registered range: replicate at the same offset
            if ((size_t)(o - c->mirror_base) + (size_t)IW * IH * 8 > c->mirror_bytes)
                return fail(c, CS_ERR_INVALID, "output image starts inside the range registered with cs_set_output_mirrors but does not fit in it");
            L.n_mirrors = c->n_mirrors;
            for (int m = 0; m < c->n_mirrors; m++) L.mirror[m] = reinterpret_cast<uint16_t*>(c->mirror_peer[m] + (o - c->mirror_base));
        }
    }
    return CS_OK;
}

// Record one side of an event pair (pairs are created on demand and reused after each read).
void timing_mark(cs_context* c, std::vector<cudaEvent_t>& evs, size_t pair, int side) {
    while (evs.size() < 2 * (pair + 1)) {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        evs.push_back(e);
    }
    cudaEventRecord(evs[2 * pair + side], c->stream);
}

// prologue + march for one rectangle, asynchronous on c->stream
int dispatch(cs_context* c, const cs_cloud_params* P, int x0, int y0, int x1, int y1, uint16_t* out, const uint16_t* sky_lut = nullptr,
             int img_w = 0, int img_h = 0, int band_ctas = 0, int band_pitch_rows = 0, int n_bands = 0) {
    if (!c || !P) return CS_ERR_INVALID;
    int r = bind(c);
    if (r) return r;
    CloudLaunch L;
    r = make_launch(c, P, x0, y0, x1, y1, out, sky_lut, L, img_w, img_h);
    if (r) return r;
    L.band_ctas = band_ctas; L.band_pitch_rows = band_pitch_rows; L.grid_y = band_ctas * n_bands;
    if (L.x1 <= L.x0 || L.y1 <= L.y0) return CS_OK;
    if (L.counters) CU(cudaMemsetAsync(c->d_counters, 0, 6 * sizeof(unsigned long long), c->stream));
    launch_clouds_prologue(L, c->mode == CS_MODE_STRICT, c->stream);
    if (c->timing_on) timing_mark(c, c->ev_march, c->n_march, 0);
    if (c->mode == CS_MODE_STRICT) launch_clouds_strict(L, c->stream);
    else launch_clouds_fast(L, c->stream);
    if (c->timing_on) timing_mark(c, c->ev_march, c->n_march++, 1);
    CU(cudaGetLastError());
    return CS_OK;
}

}  // namespace

namespace cs {
int ctx_dispatch(cs_context* c, const cs_cloud_params* P, int x0, int y0, int x1, int y1, uint16_t* out, const uint16_t* sky_lut, int img_w, int img_h) {
    return dispatch(c, P, x0, y0, x1, y1, out, sky_lut, img_w, img_h);
}
int ctx_build_sky_lut_into(cs_context* c, const float sun[3], uint16_t* dst) {
    if (!c || !sun || !dst) return CS_ERR_INVALID;
    if (!c->have_tlut) return fail(c, CS_ERR_NOT_READY, "Attempting to update uninitialized sky lut (build the transmittance LUT first)");
    int r = bind(c);
    if (r) return r;
    if (c->timing_on) timing_mark(c, c->ev_sky, c->n_sky, 0);
    launch_sky_lut(c->d_tlut, c->tlut_param, sun, dst, c->stream);
    if (c->timing_on) timing_mark(c, c->ev_sky, c->n_sky++, 1);
    CU(cudaGetLastError());
    return CS_OK;
}
int ctx_fail(cs_context* c, int code, const std::string& msg) { return fail(c, code, msg); }
}  // namespace cs

extern "C" {

int cs_create(int device, cs_context** out) {
    if (!out) return CS_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        fprintf(stderr, "cloudsky_b200: cs_create(device=%d) failed: %s (devices: %d). There is no CPU fallback.\n", device,
                e != cudaSuccess ? cudaGetErrorString(e) : "no such device", count);
        return CS_ERR_CUDA;
    }
    cs_context* c = new cs_context();
    c->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&c->d_tlut, (size_t)CS_TRANSMITTANCE_W * CS_TRANSMITTANCE_H * 8) == cudaSuccess &&
              cudaMalloc(&c->d_sky, (size_t)CS_SKY_LUT_W * CS_SKY_LUT_H * 8) == cudaSuccess &&
              cudaMalloc(&c->d_frame_consts, sizeof(FrameConsts) * kMaxSunBatch) == cudaSuccess &&
              cudaMalloc(&c->d_sky_batch, (size_t)CS_SKY_LUT_W * CS_SKY_LUT_H * 8 * kMaxSunBatch) == cudaSuccess &&
              cudaMalloc(&c->d_counters, 6 * sizeof(unsigned long long)) == cudaSuccess;
    if (!ok) {
        fprintf(stderr, "cloudsky_b200: cs_create: %s\n", cudaGetErrorString(cudaGetLastError()));
        cs_destroy(c);
        return CS_ERR_CUDA;
    }
    c->stream = c->own_stream;
    if (const char* e = getenv("CLOUDSKY_SUN_BATCH")) c->sun_batching = e[0] != '0';  // development knob: 0 = one launch per sun
    *out = c;
    return CS_OK;
}

void cs_destroy(cs_context* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);  // an in-flight device->host copy must not outlive its source image
    free_textures(c);
    if (c->d_tlut) cudaFree(c->d_tlut);
    if (c->d_sky) cudaFree(c->d_sky);
    if (c->d_frame_consts) cudaFree(c->d_frame_consts);
    if (c->d_counters) cudaFree(c->d_counters);
    if (c->d_sky_batch) cudaFree(c->d_sky_batch);
    if (c->d_image) cudaFree(c->d_image);
    if (c->d_image2) cudaFree(c->d_image2);
    if (c->d_peer_err) cudaFree(c->d_peer_err);
    for (int i = 0; i < 2; i++) {
        if (c->ev_rendered[i]) cudaEventDestroy(c->ev_rendered[i]);
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (auto e : c->ev_march) cudaEventDestroy(e);
    for (auto e : c->ev_sky) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char* cs_last_error(const cs_context* c) { return c ? c->err.c_str() : "null context"; }
const char* cs_backend_name(void) { return "cuda-sm100a"; }

int cs_set_stream(cs_context* c, void* s) {
    if (!c) return CS_ERR_INVALID;
    int r = bind(c);
    if (r) return r;
    CU(cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return CS_OK;
}
int cs_sync(cs_context* c) {
    if (!c) return CS_ERR_INVALID;
    int r = bind(c);
    if (r) return r;
    CU(cudaStreamSynchronize(c->stream));
    return CS_OK;
}
int cs_set_threads(cs_context* c, int) { return c ? CS_OK : CS_ERR_INVALID; }

// The C ABI never lets a C++ exception escape: allocation failures of the host-side staging buffers become error codes.
#define CS_GUARD_BEGIN try {
#define CS_GUARD_END(ctx)                                                                    \
    }                                                                                        \
    catch (const std::bad_alloc&) { return fail(ctx, CS_ERR_INVALID, "out of host memory"); } \
    catch (...) { return fail(ctx, CS_ERR_INVALID, "unexpected C++ exception"); }

int cs_upload_textures(cs_context* c, const uint8_t* large, int ln, int lch, const uint8_t* small, int sn, int sch,
                       const uint8_t* weather, int ww, int wh, int wch) {
    if (!c) return CS_ERR_INVALID;
    CS_GUARD_BEGIN
    if (!large || !small || !weather || ln < 1 || sn < 1 || ww < 1 || wh < 1 || lch < 3 || lch > 4 || sch < 3 || sch > 4 || wch < 3 || wch > 4)
        return fail(c, CS_ERR_INVALID, "cs_upload_textures: null pointer or bad dimensions / channel count");
    std::vector<uint8_t> l0, s0, w0;
    expand_rgba(large, (size_t)ln * ln * ln, lch, l0);
    expand_rgba(small, (size_t)sn * sn * sn, sch, s0);
    expand_rgba(weather, (size_t)ww * wh, wch, w0);
    return upload_levels(c, l0, ln, s0, sn, w0, ww, wh);
    CS_GUARD_END(c)
}

int cs_load_texture_files(cs_context* c, const char* large_path, int large_slices, const char* small_path, int small_slices,
                          const char* weather_path) {
    if (!c) return CS_ERR_INVALID;
    CS_GUARD_BEGIN
    HostImage li, si, wi;
    std::string e = decode_image_file(large_path, li);
    if (!e.empty()) return fail(c, CS_ERR_IO, e);
    e = decode_image_file(small_path, si);
    if (!e.empty()) return fail(c, CS_ERR_IO, e);
    e = decode_image_file(weather_path, wi);
    if (!e.empty()) return fail(c, CS_ERR_IO, e);
    std::vector<uint8_t> l0, s0, w0;
    int ln = 0, sn = 0;
    e = strip_to_volume_rgba(li, large_slices, l0, ln);
    if (!e.empty()) return fail(c, CS_ERR_IO, std::string(large_path) + ": " + e);
    e = strip_to_volume_rgba(si, small_slices, s0, sn);
    if (!e.empty()) return fail(c, CS_ERR_IO, std::string(small_path) + ": " + e);
    expand_rgba(wi.px.data(), (size_t)wi.w * wi.h, wi.ch, w0);
    return upload_levels(c, l0, ln, s0, sn, w0, wi.w, wi.h);
    CS_GUARD_END(c)
}

int cs_decode_image_file(const char* path, uint8_t** out_pixels, int* w, int* h, int* ch) {
    if (!path || !out_pixels || !w || !h || !ch) return CS_ERR_INVALID;
    try {
        HostImage im;
        std::string e = decode_image_file(path, im);
        if (!e.empty()) return CS_ERR_IO;
        uint8_t* p = (uint8_t*)malloc(im.px.size());
        if (!p) return CS_ERR_IO;
        memcpy(p, im.px.data(), im.px.size());
        *out_pixels = p; *w = im.w; *h = im.h; *ch = im.ch;
        return CS_OK;
    } catch (...) {
        return CS_ERR_IO;
    }
}
void cs_free(void* p) { free(p); }

int cs_read_volume_level(cs_context* c, int which, int level, uint8_t* out, size_t bytes) {
    if (!c || !out) return CS_ERR_INVALID;
    if (!c->have_tex) return fail(c, CS_ERR_NOT_READY, "no textures");
    int levels = which == 0 ? c->large_levels : c->small_levels;
    if (which < 0 || which > 1 || level < 0 || level >= levels) return fail(c, CS_ERR_INVALID, "bad volume / level");
    const uint32_t* d = which == 0 ? c->d_large[level] : c->d_small[level];
    int n = (which == 0 ? c->large_n : c->small_n) >> level;
    if (bytes != (size_t)n * n * n * 4) return fail(c, CS_ERR_INVALID, "bad output size");
    int r = bind(c);
    if (r) return r;
    CU(cudaMemcpy(out, d, bytes, cudaMemcpyDeviceToHost));  // read back what the kernels actually sample
    return CS_OK;
}

int cs_generate_noise(cs_context* c, int kind, int n, const cs_noise_params* P, uint8_t* out, size_t bytes) {
    if (!c) return CS_ERR_INVALID;
    if (const char* why = nz::check_request(kind, n, P)) return fail(c, CS_ERR_INVALID, why);
    const size_t texels = kind == CS_NOISE_WEATHER ? (size_t)n * n : (size_t)n * n * n;
    if (!out || bytes != texels * 4) return fail(c, CS_ERR_INVALID, "cs_generate_noise: out_bytes must be texels * 4");
    int r = bind(c);
    if (r) return r;
    uint32_t* d = nullptr;
    CU(cudaMalloc(&d, bytes));
    launch_noise(kind, n, *P, d, c->stream);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(c, e, "cs_generate_noise");
    return CS_OK;
}

int cs_set_transmittance_parametrisation(cs_context* c, int which) {
    if (!c) return CS_ERR_INVALID;
    if (which != CS_TLUT_LINEAR && which != CS_TLUT_BRUNETON2017) return fail(c, CS_ERR_INVALID, "cs_set_transmittance_parametrisation: CS_TLUT_LINEAR or CS_TLUT_BRUNETON2017");
    if (which != c->tlut_param) { c->tlut_param = which; c->have_tlut = false; c->have_sky = false; }  // both LUTs depend on the mapping
    return CS_OK;
}

int cs_build_transmittance_lut(cs_context* c) {
    if (!c) return CS_ERR_INVALID;
    int r = bind(c);
    if (r) return r;
    launch_transmittance_lut(c->d_tlut, c->tlut_param, c->stream);
    CU(cudaGetLastError());
    c->have_tlut = true;
    return CS_OK;
}
int cs_build_sky_lut(cs_context* c, const float sun[3]) {
    if (!c || !sun) return CS_ERR_INVALID;
    if (!c->have_tlut) return fail(c, CS_ERR_NOT_READY, "Attempting to update uninitialized sky lut (build the transmittance LUT first)");
    int r = bind(c);
    if (r) return r;
    if (c->timing_on) timing_mark(c, c->ev_sky, c->n_sky, 0);
    launch_sky_lut(c->d_tlut, c->tlut_param, sun, c->d_sky, c->stream);
    if (c->timing_on) timing_mark(c, c->ev_sky, c->n_sky++, 1);
    CU(cudaGetLastError());
    c->have_sky = true;
    return CS_OK;
}

static int read_dev(cs_context* c, const void* d, void* out, size_t bytes, size_t expect, bool ready, const char* what) {
    if (!c || !out) return CS_ERR_INVALID;
    if (!ready) return fail(c, CS_ERR_NOT_READY, std::string(what) + " not built");
    if (bytes != expect) return fail(c, CS_ERR_INVALID, std::string(what) + ": bad buffer size");
    int r = bind(c);
    if (r) return r;
    CU(cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return CS_OK;
}
static int write_dev(cs_context* c, void* d, const void* in, size_t bytes, size_t expect, const char* what) {
    if (!c || !in) return CS_ERR_INVALID;
    if (bytes != expect) return fail(c, CS_ERR_INVALID, std::string(what) + ": bad buffer size");
    int r = bind(c);
    if (r) return r;
    CU(cudaMemcpyAsync(d, in, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return CS_OK;
}
int cs_read_transmittance_lut(cs_context* c, uint16_t* out, size_t bytes) {
    return read_dev(c, c ? c->d_tlut : nullptr, out, bytes, (size_t)CS_TRANSMITTANCE_W * CS_TRANSMITTANCE_H * 8, c && c->have_tlut, "transmittance LUT");
}
int cs_read_sky_lut(cs_context* c, uint16_t* out, size_t bytes) {
    return read_dev(c, c ? c->d_sky : nullptr, out, bytes, (size_t)CS_SKY_LUT_W * CS_SKY_LUT_H * 8, c && c->have_sky, "sky LUT");
}
int cs_write_transmittance_lut(cs_context* c, const uint16_t* in, size_t bytes) {
    int r = write_dev(c, c ? c->d_tlut : nullptr, in, bytes, (size_t)CS_TRANSMITTANCE_W * CS_TRANSMITTANCE_H * 8, "transmittance LUT");
    if (r == CS_OK) c->have_tlut = true;
    return r;
}
int cs_write_sky_lut(cs_context* c, const uint16_t* in, size_t bytes) {
    int r = write_dev(c, c ? c->d_sky : nullptr, in, bytes, (size_t)CS_SKY_LUT_W * CS_SKY_LUT_H * 8, "sky LUT");
    if (r == CS_OK) c->have_sky = true;
    return r;
}

int cs_resize(cs_context* c, int w, int h) {
    if (!c) return CS_ERR_INVALID;
    if (w < 1 || h < 1 || w > 16384 || h > 16384) return fail(c, CS_ERR_INVALID, "cs_resize: size must be in [1, 16384]");
    int r = bind(c);
    if (r) return r;
    CU(cudaStreamSynchronize(c->stream));
    if (c->copy_stream) CU(cudaStreamSynchronize(c->copy_stream));
    if (c->d_image) { cudaFree(c->d_image); c->d_image = nullptr; }
    if (c->d_image2) { cudaFree(c->d_image2); c->d_image2 = nullptr; }
    c->slot_busy[0] = c->slot_busy[1] = false;
    c->W = c->H = 0;
    CU(cudaMalloc(&c->d_image, (size_t)w * h * 8));
    CU(cudaMemsetAsync(c->d_image, 0, (size_t)w * h * 8, c->stream));
    c->W = w; c->H = h;
    return CS_OK;
}
int cs_set_march_config(cs_context* c, int p, int cone, int mode) {
    if (!c) return CS_ERR_INVALID;
    if (p < 1 || p > 4096 || cone < 0 || cone > 64 || (mode != CS_MODE_STRICT && (mode & ~(CS_MODE_EARLY_OUT | CS_MODE_TEX | CS_MODE_HALF)) != CS_MODE_FAST) ||
        (mode != CS_MODE_STRICT && (mode & CS_MODE_TEX) && (mode & CS_MODE_HALF)))
        return fail(c, CS_ERR_INVALID, "cs_set_march_config: primary_steps in [1,4096], cone_samples in [0,64], mode STRICT, or FAST optionally | EARLY_OUT | one of TEX, HALF");
    if (mode != CS_MODE_STRICT && (mode & CS_MODE_HALF)) {
        int r = ensure_half2_records(c);
        if (r) return r;
    }
    c->primary_steps = p; c->cone_samples = cone; c->mode = mode;
    return CS_OK;
}
int cs_set_step_budget(cs_context* c, float len, int min_steps) {
    if (!c) return CS_ERR_INVALID;
    if (!(len >= 0.0f) || min_steps < 1) return fail(c, CS_ERR_INVALID, "cs_set_step_budget: min_step_length_m >= 0, min_steps >= 1");
    c->budget_len = len; c->budget_min = min_steps;
    return CS_OK;
}
int cs_set_counters_enabled(cs_context* c, int on) {
    if (!c) return CS_ERR_INVALID;
    c->counters_on = on != 0;
    return CS_OK;
}
int cs_get_counters(cs_context* c, cs_counters* out) {
    if (!c || !out) return CS_ERR_INVALID;
    int r = bind(c);
    if (r) return r;
    unsigned long long h[6];
    CU(cudaMemcpyAsync(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    out->marched_pixels = h[0]; out->primary_steps = h[1]; out->lit_steps = h[2];
    out->density_evals = h[3]; out->large_fetches = h[4]; out->small_fetches = h[5];
    return CS_OK;
}

int cs_set_kernel_timing(cs_context* c, int on) {
    if (!c) return CS_ERR_INVALID;
    int r = bind(c);
    if (r) return r;
    CU(cudaStreamSynchronize(c->stream));
    c->timing_on = on != 0;
    c->n_march = c->n_sky = 0;
    return CS_OK;
}
int cs_read_kernel_timings(cs_context* c, float* march_ms, int* n_march, float* sky_ms, int* n_sky) {
    if (!c || !march_ms || !n_march || !sky_ms || !n_sky) return CS_ERR_INVALID;
    int r = bind(c);
    if (r) return r;
    CU(cudaStreamSynchronize(c->stream));
    float sm = 0.0f, ss = 0.0f, t = 0.0f;
    for (size_t i = 0; i < c->n_march; i++) { CU(cudaEventElapsedTime(&t, c->ev_march[2 * i], c->ev_march[2 * i + 1])); sm += t; }
    for (size_t i = 0; i < c->n_sky; i++) { CU(cudaEventElapsedTime(&t, c->ev_sky[2 * i], c->ev_sky[2 * i + 1])); ss += t; }
    *march_ms = sm; *n_march = (int)c->n_march; *sky_ms = ss; *n_sky = (int)c->n_sky;
    c->n_march = c->n_sky = 0;
    return CS_OK;
}

int cs_dispatch_clouds(cs_context* c, const cs_cloud_params* P, int gx, int gy) {
    if (!c || !P) return CS_ERR_INVALID;
    if (gx < 1 || gy < 1) return fail(c, CS_ERR_INVALID, "cs_dispatch_clouds: group counts must be >= 1");
    int x0 = (int)P->update_position[0], y0 = (int)P->update_position[1];  // ivec2(params.update_position), clouds.glsl:260
    return dispatch(c, P, x0, y0, x0 + 8 * gx, y0 + 8 * gy, c->d_image);
}
int cs_render_frame(cs_context* c, const cs_cloud_params* P) {
    if (!c || !P) return CS_ERR_INVALID;
    return dispatch(c, P, 0, 0, c->W, c->H, c->d_image);
}
int cs_render_rows_to(cs_context* c, const cs_cloud_params* P, int r0, int r1, void* out) {
    if (!c || !P || !out) return CS_ERR_INVALID;
    return dispatch(c, P, 0, r0, c->W, r1, (uint16_t*)out);
}
int cs_render_row_bands_to(cs_context* c, const cs_cloud_params* P, int first_row, int band_rows, int band_pitch_rows, int n_bands, void* out) {
    if (!c || !P || !out) return CS_ERR_INVALID;
    if (first_row < 0 || band_rows < 8 || band_rows % 8 != 0 || band_pitch_rows < band_rows || n_bands < 1 || n_bands > 65535 / (band_rows / 8))
        return fail(c, CS_ERR_INVALID, "cs_render_row_bands_to: band_rows must be a positive multiple of 8, band_pitch_rows >= band_rows, n_bands >= 1");
    return dispatch(c, P, 0, first_row, c->W, c->H, (uint16_t*)out, nullptr, 0, 0, band_rows / 8, band_pitch_rows, n_bands);
}
void* cs_image_device_ptr(cs_context* c) { return c ? c->d_image : nullptr; }
int cs_read_image(cs_context* c, uint16_t* out, size_t bytes) {
    if (!c) return CS_ERR_INVALID;
    return read_dev(c, c->d_image, out, bytes, (size_t)c->W * c->H * 8, c->d_image != nullptr, "output image");
}
int cs_render_frame_host(cs_context* c, const cs_cloud_params* P, uint16_t* out, size_t bytes) {
    if (!c || !P || !out) return CS_ERR_INVALID;
    if (bytes != (size_t)c->W * c->H * 8) return fail(c, CS_ERR_INVALID, "cs_render_frame_host: bad buffer size");
    int r = cs_build_sky_lut(c, P->light_direction);
    if (r) return r;
    r = dispatch(c, P, 0, 0, c->W, c->H, c->d_image);
    if (r) return r;
    CU(cudaMemcpyAsync(out, c->d_image, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return CS_OK;
}
int cs_render_frame_host_async(cs_context* c, const cs_cloud_params* P, uint16_t* out, size_t bytes) {
    if (!c || !P || !out) return CS_ERR_INVALID;
    if (bytes != (size_t)c->W * c->H * 8) return fail(c, CS_ERR_INVALID, "cs_render_frame_host_async: bad buffer size");
    int r = bind(c);
    if (r) return r;
    if (!c->copy_stream) {
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CU(cudaEventCreateWithFlags(&c->ev_rendered[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
        }
    }
    if (!c->d_image2) CU(cudaMalloc(&c->d_image2, bytes));
    const int slot = (int)(c->async_frame++ & 1u);
    uint16_t* img = slot ? c->d_image2 : c->d_image;
    if (c->slot_busy[slot]) CU(cudaStreamWaitEvent(c->stream, c->ev_copied[slot], 0));  // the image is free once its last copy finished
    r = cs_build_sky_lut(c, P->light_direction);
    if (r) return r;
    r = dispatch(c, P, 0, 0, c->W, c->H, img);
    if (r) return r;
    CU(cudaEventRecord(c->ev_rendered[slot], c->stream));
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_rendered[slot], 0));
    CU(cudaMemcpyAsync(out, img, bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    CU(cudaEventRecord(c->ev_copied[slot], c->copy_stream));
    c->slot_busy[slot] = true;
    return CS_OK;
}
int cs_wait_host(cs_context* c) {
    if (!c) return CS_ERR_INVALID;
    int r = bind(c);
    if (r) return r;
    CU(cudaStreamSynchronize(c->stream));
    if (c->copy_stream) CU(cudaStreamSynchronize(c->copy_stream));
    c->slot_busy[0] = c->slot_busy[1] = false;
    return CS_OK;
}
int cs_render_sun_batch_to(cs_context* c, const cs_cloud_params* P, const float* suns, int n, void* out) {
    if (!c || !P || !suns || !out || n < 1) return CS_ERR_INVALID;
    const size_t image_px = (size_t)c->W * c->H;
    // CS_MODE_FAST with the record sampler: up to kMaxSunBatch suns share one march (clouds_fast_sunbatch_kernel) — the primary
    // loop is sun-independent.  Other modes, instrumented runs and step counts beyond the tables: one launch per sun.
    const bool batched = c->mode == CS_MODE_FAST /* no flags: the batch kernel exists for the fp32-filter record formats only */ && !c->counters_on && !c->timing_on && c->cone_samples + 1 <= 16 && n > 1 && c->have_tlut && c->sun_batching;
    int i = 0;
    while (i < n) {
        const int k = batched ? std::min(n - i, (int)kMaxSunBatch) : 1;
        if (k == 1) {
            cs_cloud_params q = *P;
            memcpy(q.light_direction, suns + 3 * i, 12);
            int r = cs_build_sky_lut(c, q.light_direction);
            if (r) return r;
            r = dispatch(c, &q, 0, 0, c->W, c->H, (uint16_t*)out + (size_t)i * image_px * 4);
            if (r) return r;
        } else {
            int r = bind(c);
            if (r) return r;
            CloudLaunch L;
            for (int s = 0; s < k; s++) {  // one sky LUT and one prologue (FrameConsts) per sun, exactly as for a single frame
                cs_cloud_params q = *P;
                memcpy(q.light_direction, suns + 3 * (i + s), 12);
                uint16_t* lut = c->d_sky_batch + (size_t)s * CS_SKY_LUT_W * CS_SKY_LUT_H * 4;
                r = ctx_build_sky_lut_into(c, q.light_direction, lut);
                if (r) return r;
                r = make_launch(c, &q, 0, 0, c->W, c->H, (uint16_t*)out + (size_t)i * image_px * 4, lut, L);
                if (r) return r;
                L.frame_consts = c->d_frame_consts + (size_t)s * (sizeof(FrameConsts) / sizeof(float));
                launch_clouds_prologue(L, false, c->stream);
            }
            L.frame_consts = c->d_frame_consts;
            L.n_suns = k; L.sun_stride_px = image_px;
            if (L.n_mirrors > 0 && (size_t)(reinterpret_cast<const uint8_t*>(L.out) - c->mirror_base) + (size_t)k * image_px * 8 > c->mirror_bytes)
                return fail(c, CS_ERR_INVALID, "cs_render_sun_batch_to: the batch does not fit in the range registered with cs_set_output_mirrors");
            if (!launch_clouds_fast_sunbatch(L, c->stream)) return fail(c, CS_ERR_INVALID, "cs_render_sun_batch_to: no batch kernel for this configuration");
            CU(cudaGetLastError());
        }
        i += k;
    }
    return CS_OK;
}
int cs_time_render_frame(cs_context* c, const cs_cloud_params* P, int warmup, int iters, float* ms) {
    if (!c || !P || !ms || iters < 1 || warmup < 0) return CS_ERR_INVALID;
    int r = bind(c);
    if (r) return r;
    for (int i = 0; i < warmup; i++) { r = dispatch(c, P, 0, 0, c->W, c->H, c->d_image); if (r) return r; }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t ce = cudaEventCreate(&e0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e1);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
    if (ce == cudaSuccess) ce = cudaEventRecord(e0, c->stream);
    if (ce != cudaSuccess) {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        return cuda_fail(c, ce, "cs_time_render_frame");
    }
    for (int i = 0; i < iters; i++) { r = dispatch(c, P, 0, 0, c->W, c->H, c->d_image); if (r) break; }
    cudaEventRecord(e1, c->stream);
    cudaError_t e = cudaEventSynchronize(e1);
    float t = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (r) return r;
    if (e != cudaSuccess) return cuda_fail(c, e, "cs_time_render_frame");
    *ms = t / (float)iters;
    return CS_OK;
}

}  // extern "C"

// Internal types shared by the host-side C++ and the CUDA translation units of
// libcloudsky_b200.so.  Not part of the public ABI (include/cloudsky.h is).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/cloudsky.h"

namespace cs {

constexpr int kMaxLargeLevels = 10;  // up to 512^3 -> 1 (the reference asset is 128^3: 8 levels)
constexpr int kMaxSmallLevels = 9;   // up to 256^3 -> 1 (the reference asset is 32^3: 6 levels)
constexpr int kMaxSunBatch = 4;     // suns marched together by the sun-batch kernel (cs_render_sun_batch_to)
constexpr int kMaxMirrors = 7;      // peer replicas of the output a march kernel also stores to (cs_set_output_mirrors): 8 GPUs - 1

// ---- host-side assets (assets.cpp) ----------------------------------------------------------
struct HostImage {
    int w = 0, h = 0, ch = 0;  // top-row-first, RGB or RGBA
    std::vector<uint8_t> px;
};
// Decode an RLE / raw true-colour TGA or a BI_RGB BMP.  Returns empty string on success.
std::string decode_image_file(const char* path, HostImage& out);
// Slice a (n*slices) x n strip into an n^3 RGBA8 volume, x fastest (texel (x,y,z) = column z*n+x, row y).
std::string strip_to_volume_rgba(const HostImage& strip, int slices, std::vector<uint8_t>& out, int& n);
// Expand 3/4-channel interleaved texels to RGBA8 (alpha 255 when absent).
void expand_rgba(const uint8_t* src, size_t texels, int ch, std::vector<uint8_t>& dst);
// 2x2x2 box filter, (sum + 4) >> 3, all levels down to 1^3.  levels[0] must hold level 0.
void build_volume_mips(std::vector<std::vector<uint8_t>>& levels, int n);

// ---- device-side parameter blocks -------------------------------------------------------------
// Everything the cloud kernels need, passed by value as a __grid_constant__ argument.
struct CloudLaunch {
    cs_cloud_params P;        // the push-constant block, verbatim
    int width, height;        // image size (== P.texture_size)
    int x0, y0, x1, y1;       // pixel rectangle to render (already clipped)
    int out_pitch_px;         // row pitch of out, in pixels
    // Interleaved row bands in ONE launch (cs_render_row_bands_to; multi-GPU strong scaling of a single frame): CTA row `by` renders image
    // rows y0 + (by / band_ctas) * band_pitch_rows + (by % band_ctas) * 8 ...  band_ctas == 0: contiguous rows from y0 (the default).
    int band_ctas, band_pitch_rows, grid_y;
    float early_out_T;        // 0: run every primary step like the reference; > 0: CS_MODE_EARLY_OUT transmittance threshold
    float budget_len;         // > 0: cs_set_step_budget — steps(dir) = clamp(ceil(shell length / budget_len), budget_min, primary_steps)
    int budget_min;
    int primary_steps;        // 128 in the reference
    int cone_samples;         // 6 in the reference
    int large_n, large_levels;
    int small_n, small_levels;
    int weather_w, weather_h;
    const uint32_t* large[kMaxLargeLevels];   // RGBA8 texels, x fastest
    const uint32_t* small[kMaxSmallLevels];   // RGBA8 texels (alpha unused)
    const uint32_t* weather;                  // RGBA8 texels
    // fp32 interpolation-coefficient records for the fast kernel (see context.cu / clouds_fast.cu)
    int large_shift, small_shift, weather_shx, weather_shy;  // log2 of the level-0 edges
    float large_fn0, small_fn0, weather_fw, weather_fh;   // level-0 texels per metre (edge * texture scale) and weather edges, used straight from the constant bank
    int large_mask0, small_mask0, weather_maskx, weather_masky;
    int records_half;     // format mask: bit 0 large_f, bit 1 small_f, bit 2 weather_f; set = fp16 records (32/16/16 B), clear = fp32 (64/32/32 B)
    int small_tail_level;    // index of the 1^3 level of the small volume (-1: none)
    float small_tail_value;  // its hfbm, computed the way the active sampler format would (context.cu)
    int weather_type_hi;  // 1 when every weather texel has R >= 128 (cloud type >= 0.5): affine height-gradient fast path
    const float* large_f[kMaxLargeLevels];  // 64 B per texel: 8 trilinear coefficients of R, then 8 of fbm
    const float* small_f[kMaxSmallLevels];  // 32 B per texel: 8 trilinear coefficients of hfbm
    const float* weather_f;                 // 32 B per texel: 4 bilinear coefficients of type, then 4 of coverage
    // CS_MODE_TEX: texture objects over the same RGBA8 mip chains (REPEAT, linear, explicit level), filtered by the texture unit
    unsigned long long tex_large, tex_small, tex_weather;  // cudaTextureObject_t
    int hw_filter;                                          // 1: the fast kernel samples through the texture objects
    const uint16_t* sky_lut;                    // half4 200x100
    const float* frame_consts;                  // FrameConsts written by the prologue kernel
    uint16_t* out;                              // half4 image
    unsigned long long* counters;               // 6 x u64 or nullptr
    int n_suns;                                 // sun-batch kernel: frame_consts holds this many FrameConsts, out this many images
    size_t sun_stride_px;                       // pixels between consecutive images of a sun batch
    // Fused all-gather (cs_set_output_mirrors): every finished pixel is also stored, at the same offset, into these
    // peer-mapped replicas of the output buffer (other GPUs' memory, reached over NVLink by plain st.global).
    int n_mirrors;
    uint16_t* mirror[kMaxMirrors];
};

// Pixel-independent values of march()'s prologue (clouds.glsl:149-167), computed once per
// dispatch by clouds_prologue_kernel and read by every thread.
struct FrameConsts {
    float ldir[3];
    float atmosphere_sun[3];
    float atmosphere_ambient[3];
    float atmosphere_ground[3];
    float hg_g2;  // 0.4 - 1.4 * ldir.y
    float pad[3];
};

// kernel launchers (defined in the .cu files) — all asynchronous on `stream`.
void launch_transmittance_lut(uint16_t* out_half4, int parametrisation, void* stream);
void launch_sky_lut(const uint16_t* transmittance_half4, int parametrisation, const float sun_dir[3], uint16_t* out_half4, void* stream);
void launch_clouds_prologue(const CloudLaunch& L, bool strict, void* stream);
void launch_clouds_strict(const CloudLaunch& L, void* stream);
void launch_clouds_fast(const CloudLaunch& L, void* stream);
bool launch_clouds_fast_sunbatch(const CloudLaunch& L, void* stream);  // false: this configuration has no batch kernel
void launch_noise(int kind, int n, const cs_noise_params& P, uint32_t* out_rgba8, void* stream);  // noise_gen.cu

}  // namespace cs

/*
 * cloudsky.h — C-ABI of the B200-native cloud-sky hot path.
 *
 * The reference (clayjohn/godot-volumetric-cloud-demo-v2) has no FFI / plugin API.
 * Its operator boundary for this path is the Godot RenderingDevice compute dispatch:
 * a push-constant block + descriptor sets + compute_list_dispatch
 * (cloud_sky/cloud_sky.gd:234-248, sky_lut.gd:122-148, transmittance_lut.gd:51-78).
 * Every entry point below names the reference interface it replaces (file:line,
 * paths relative to the reference root).
 *
 * Plain C: opaque context pointer, plain pointers and sizes, int return codes
 * (0 = CS_OK).  Never aborts; on error the code is returned and
 * cs_last_error(ctx) holds a message (the reference only has boolean flags + prints:
 * cloud_sky.gd:96,130-131,362-364; sky_lut.gd:45-47).
 *
 * Two shared libraries implement this same header:
 *   libcloudsky_b200.so   — the product: hand-written sm_100a CUDA (this repo's csrc/)
 *   libcloudsky_oracle.so — TEST INFRASTRUCTURE ONLY: scalar fp32 CPU restatement (oracle/)
 *
 * Threading: one context = one CUDA device + one stream.  Calls on one context are
 * not thread-safe; different contexts are independent (the reference marshals all GPU
 * work to a single render thread: cloud_sky.gd:118,154).
 */
#ifndef CLOUDSKY_H
#define CLOUDSKY_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CS_OK 0
#define CS_ERR_INVALID 1     /* bad argument / bad state                              */
#define CS_ERR_CUDA 2        /* CUDA runtime failure (message in cs_last_error)       */
#define CS_ERR_IO 3          /* asset file could not be read / decoded                */
#define CS_ERR_UNSUPPORTED 4 /* e.g. CUDA call on the oracle backend                  */
#define CS_ERR_NOT_READY 5   /* textures / LUTs missing ("can_run == false")          */

/* Fixed sizes of the two atmosphere LUTs. */
#define CS_TRANSMITTANCE_W 256 /* transmittance_lut.gd:6 */
#define CS_TRANSMITTANCE_H 64
#define CS_SKY_LUT_W 200 /* sky_lut.gd:4 */
#define CS_SKY_LUT_H 100

/* Reference march constants (clouds.glsl:228 and :186). */
#define CS_REF_PRIMARY_STEPS 128
#define CS_REF_CONE_SAMPLES 6

typedef struct cs_context cs_context;

/*
 * The 112-byte std430 push-constant block of clouds.glsl (clouds.glsl:18-40), in the
 * exact order cloud_sky.gd:_fill_push_constant packs it (cloud_sky.gd:251-289), so a
 * Godot-side caller can memcpy its PackedFloat32Array into this struct.
 */
typedef struct cs_cloud_params {
    float texture_size[2];    /* @0   W, H of the hemisphere texture (pixels)            */
    float update_position[2]; /* @8   pixel origin of the tile being dispatched          */
    float cloud_pos[2];       /* @16  base wind offset                                   */
    float detailed_pos[2];    /* @24  detail wind offset                                 */
    float weather_pos[2];     /* @32  weather-map offset                                 */
    float pad1[2];            /* @40                                                     */
    float ground_color[4];    /* @48                                                     */
    float light_direction[3]; /* @64  unit vector TOWARD the sun (cloud_sky.gd:76-79)    */
    float light_energy;       /* @76                                                     */
    float light_color[3];     /* @80  linear RGB                                         */
    float time;               /* @92  absolute seconds (cloud_sky.gd:282)                */
    float pad2;               /* @96                                                     */
    float density;            /* @100                                                    */
    float cloud_coverage;     /* @104                                                    */
    float time_offset;        /* @108 pushed but unused by the shader                    */
} cs_cloud_params;

/* User-level parameter surface: the @export properties of cloud_sky.gd:4-50. */
typedef struct cs_sky_settings {
    float wind_direction;  /* radians (cloud_sky.gd:9-10)                              */
    float wind_speed;      /* m/s     (cloud_sky.gd:13-14)                             */
    float density;         /* cloud_sky.gd:19-20                                       */
    float cloud_coverage;  /* cloud_sky.gd:21-22                                       */
    float time_offset;     /* cloud_sky.gd:23-24                                       */
    float sun_disk_scale;  /* cloud_sky.gd:27-31 (presentation only)                   */
    float ground_color[4]; /* cloud_sky.gd:32-33                                       */
    int32_t frames_to_update; /* 4 / 16 / 64 / 256 (cloud_sky.gd:36-42); 1 = one dispatch */
    int32_t texture_size;     /* cloud_sky.gd:44-50                                    */
} cs_sky_settings;

/* FrameData's derived state (cloud_sky.gd:56-79): wind accumulators + light snapshot. */
typedef struct cs_frame_state {
    float time;            /* _time          cloud_sky.gd:66 */
    float cloud_pos[2];    /* _cloud_pos     cloud_sky.gd:67 */
    float detailed_pos[2]; /* _detailed_pos  cloud_sky.gd:68 */
    float weather_pos[2];  /* _weather_pos   cloud_sky.gd:69 */
    float light_direction[3]; /* LIGHT_DIRECTION cloud_sky.gd:72 */
    float light_energy;       /* LIGHT_ENERGY    cloud_sky.gd:73 */
    float light_color[3];     /* LIGHT_COLOR (linear) cloud_sky.gd:74 */
} cs_frame_state;

/* Work counters of the last instrumented clouds dispatch. */
typedef struct cs_counters {
    uint64_t marched_pixels;    /* pixels with dir.y > 0                                   */
    uint64_t primary_steps;     /* executed primary-loop iterations                        */
    uint64_t lit_steps;         /* primary steps with t > 0 (light march executed)         */
    uint64_t density_evals;     /* calls of density() (primary + cone + distant)           */
    uint64_t large_fetches;     /* trilinear fetches of the large volume actually issued   */
    uint64_t small_fetches;     /* trilinear fetches of the small volume actually issued   */
} cs_counters;

/* march mode flags for cs_set_march_config */
#define CS_MODE_FAST 0   /* product kernel: fast intrinsics and exact-zero skips inside density(); sample positions and heights
                          * in the shader's own fp32 roundings (the reference's rounding trajectory)                      */
#define CS_MODE_STRICT 1 /* same operation order as the oracle, --fmad=false, IEEE div/sqrt  */
/* Optional flag for CS_MODE_FAST (mode = CS_MODE_FAST | CS_MODE_EARLY_OUT), NOT reference behaviour (the reference runs all
 * primary steps even after T ~ 0, clouds.glsl:172): stop marching a ray once its transmittance T < 2^-12.  alpha = 1 - T
 * then already rounds to 1.0 in fp16 and the skipped radiance is below one fp16 ulp; executed (not nominal) primary steps
 * are reported by the counters.  Pays off for overcast skies (BASELINE config 5). */
#define CS_MODE_EARLY_OUT 2
/* Optional flag for CS_MODE_FAST (mode = CS_MODE_FAST | CS_MODE_TEX [| CS_MODE_EARLY_OUT]): sample the two noise volumes and
 * the weather map through the GPU's texture unit (REPEAT, linear, explicit mip level) — what the reference's own sampler2D /
 * sampler3D bindings do (cloud_sky.gd:389-398) — instead of the in-kernel fp32 trilinear filter.  The texture unit
 * interpolates with 8-bit fixed-point weights, so results differ from the fp32-filtered oracle by up to ~3e-3 absolute
 * (inside the FAST parity tolerance, tests/test_gpu_parity.py); texel values and mip chains are identical. */
#define CS_MODE_TEX 4
/* Optional flag for CS_MODE_FAST (mode = CS_MODE_FAST | CS_MODE_HALF [| CS_MODE_EARLY_OUT]; not together with CS_MODE_TEX): the
 * in-kernel trilinear / bilinear filter is evaluated in packed fp16 (HFMA2) directly on the exact-integer coefficient records —
 * no half->float conversions, two channels per instruction.  Filter weights and the three lerp levels round to 11 bits
 * (finer than the texture unit's 8-bit weights, coarser than the fp32 filter of plain CS_MODE_FAST); texel values and mip
 * chains are identical.  Same parity tolerance as CS_MODE_FAST (tests/test_gpu_parity.py).  Needs textures whose
 * interpolation coefficients are exact in fp16 (true for the reference's; CS_ERR_UNSUPPORTED otherwise). */
#define CS_MODE_HALF 8

/* ---- lifetime -------------------------------------------------------------------------- */

/* Replaces _initialize_compute_code (cloud_sky.gd:355-408): device/pipeline setup.
 * device = CUDA ordinal (ignored by the oracle backend). */
int cs_create(int device, cs_context** out_ctx);
/* Replaces cleanup() / NOTIFICATION_PREDELETE (cloud_sky.gd:193-212, sky_lut.gd:29-37,
 * transmittance_lut.gd:20-25). */
void cs_destroy(cs_context* ctx);
/* Replaces the print()-based error reporting (sky_lut.gd:45-47,105-107). */
const char* cs_last_error(const cs_context* ctx);
/* "cuda-sm100a" or "oracle-cpu". */
const char* cs_backend_name(void);
/* Use an existing cudaStream_t (e.g. torch's current stream) for all work of this context.
 * NULL = the context's own stream.  (Godot: the render-thread queue, cloud_sky.gd:154.) */
int cs_set_stream(cs_context* ctx, void* cuda_stream);
/* Wait for all work queued on the context's stream. */
int cs_sync(cs_context* ctx);
/* Worker threads for the oracle backend (no-op for CUDA). */
int cs_set_threads(cs_context* ctx, int n_threads);

/* ---- input textures (set 1 bindings 0/1/2; cloud_sky.gd:298-341, clouds.glsl:10-12) ------ */

/* large: n^3 texels, `large_ch` (4) bytes per texel, x fastest then y then z (texel (x,y,z) of
 *        the sliced strip: perlworlnoise.tga.import:24-27);
 * small: n^3 texels, 3 or 4 bytes per texel (worlnoise.bmp.import:24-27);
 * weather: w*h texels, 3 or 4 bytes per texel, row 0 = v 0 (weather.bmp.import:25, no mips).
 * The library copies the data and builds the box-filter mip chains of the two volumes.
 * Sizes (CUDA backend): every edge a power of two (REPEAT addressing by mask); large <= 512^3, small <= 256^3, weather <=
 * 8192^2 (the reference assets are 128^3 / 32^3 / 512^2).  Resident device footprint of the volumes: 8 x the RGBA8 texels
 * (one 32-byte / 16-byte interpolation record per texel and mip level) + the texels themselves + a mipmapped CUDA array. */
int cs_upload_textures(cs_context* ctx,
                       const uint8_t* large, int large_n, int large_ch,
                       const uint8_t* small, int small_n, int small_ch,
                       const uint8_t* weather, int weather_w, int weather_h, int weather_ch);
/* Replaces preload("perlworlnoise.tga") / preload("worlnoise.bmp") / preload("weather.bmp")
 * (cloud_sky.gd:311,321,331): decodes RLE TGA / BI_RGB BMP strips and slices them
 * horizontally (slices/horizontal, *.import:26). */
int cs_load_texture_files(cs_context* ctx, const char* large_path, int large_slices,
                          const char* small_path, int small_slices, const char* weather_path);
/* Stand-alone decoder used by the call above: returns malloc'ed top-row-first RGB(A) bytes
 * (free with cs_free). */
int cs_decode_image_file(const char* path, uint8_t** out_pixels, int* out_w, int* out_h,
                         int* out_channels);
void cs_free(void* p);
/* Read back one mip level of the large (which=0) or small (which=1) volume as RGBA8
 * (level 0 = the upload).  out must hold (n>>level)^3*4 bytes. */
int cs_read_volume_level(cs_context* ctx, int which, int level, uint8_t* out, size_t out_bytes);

/* ---- noise synthesis (README.md:30 TODO "Implement a noise generator so custom noise can be created and tweaked"; ----
 * ---- SURVEY 8(f)-3): tileable stand-ins for the three bitmap inputs, at any power-of-two resolution --------------- */

#define CS_NOISE_LARGE 0   /* n^3 RGBA8: R Perlin-Worley, G/B/A Worley fBm at rising frequencies (perlworlnoise.tga; clouds.glsl:117-119) */
#define CS_NOISE_SMALL 1   /* n^3 RGBA8: R/G/B Worley fBm at rising frequencies, A = 255 (worlnoise.bmp; clouds.glsl:132-133)            */
#define CS_NOISE_WEATHER 2 /* n^2 RGBA8: R cloud type, G = 0, B coverage, A = 255 (weather.bmp; clouds.glsl:121,123)                      */
typedef struct cs_noise_params {
    uint32_t seed;
    int32_t worley_frequency; /* feature cells across the tile of the lowest Worley octave (octave k uses frequency << k) */
    float worley_scale;       /* Worley value = 1 - min(worley_scale * distance to the nearest feature point in cells, 1) */
    int32_t perlin_frequency; /* lattice cells across the tile of the first Perlin octave                               */
    int32_t perlin_octaves;   /* 1..8                                                                                   */
    float perlin_scale;       /* perlin01 = clamp(0.5 + perlin_scale * fBm)                                             */
    float remap_lo, remap_hi; /* weather coverage = clamp((perlin-worley - remap_lo) / (remap_hi - remap_lo))           */
    float type_lo, type_hi;   /* weather cloud type range (the reference map spans 0.59..0.91)                          */
} cs_noise_params;
/* Defaults per kind (large: Worley 4 / Perlin 4 x 5 octaves; small: Worley 2; weather: Worley 4 / Perlin 4 x 4 octaves,
 * coverage remap 0.55..0.95, type 0.59..0.91; worley_scale 0.56 puts the Worley channels' mean near the 0.71 of the
 * reference bitmaps). */
void cs_noise_params_default(int kind, cs_noise_params* out);
/* Generate one texture on the device and copy it to out_rgba8 (host; n^3*4 bytes for the volumes, x fastest then y then z
 * as cs_upload_textures expects with 4 channels; n*n*4 for the weather map).  n: power of two.  Every backend produces the
 * same bytes (integer lattice hash, fp32 + - * / sqrt only, fixed evaluation order). */
int cs_generate_noise(cs_context* ctx, int kind, int n, const cs_noise_params* params, uint8_t* out_rgba8, size_t out_bytes);

/* ---- atmosphere LUTs -------------------------------------------------------------------- */

/* How the transmittance LUT maps (altitude, cosine of the sun zenith angle) to texels.
 * CS_TLUT_LINEAR (default) is the reference's mapping: u = 0.5 + 0.5 cos, v = altitude / thickness, the integral running
 * to the top boundary even through the planet (transmittance-lut.glsl:161-171, sky-lut.glsl:137-142, clouds.gdshader:77-85).
 * CS_TLUT_BRUNETON2017 is the second TODO of the reference's README (README.md:29 "Use the transmittance LUT
 * parametrization from Bruneton (2017)"): (r, mu) -> (x_r, x_mu) through the distance to the top boundary, texel centres on
 * the ends of the unit range, only rays that miss the ground stored, the planet's shadow applied at lookup with a
 * smoothstep over the sun's angular radius.  Same 256x64 RGBA16F texture, same 40-step integral, every consumer
 * (sky-LUT build, presentation composite) switches with it.  Changing it invalidates both LUTs (rebuild them). */
#define CS_TLUT_LINEAR 0
#define CS_TLUT_BRUNETON2017 1
int cs_set_transmittance_parametrisation(cs_context* ctx, int which);
/* Replaces transmittance_lut.gd:_initialize_compute_code's one dispatch (32x8 groups,
 * transmittance_lut.gd:66-78) of transmittance-lut.glsl:157-196.  256x64 RGBA16F. */
int cs_build_transmittance_lut(cs_context* ctx);
/* Replaces SkyLUT.update_lut(sun_direction) -> render_lut (sky_lut.gd:43-52,122-148):
 * one dispatch of sky-lut.glsl:278-315.  200x100 RGBA16F.  Needs the transmittance LUT. */
int cs_build_sky_lut(cs_context* ctx, const float sun_direction[3]);
/* Readback / injection of the two LUTs as tightly packed half4 texels (row 0 = v 0). */
int cs_read_transmittance_lut(cs_context* ctx, uint16_t* out_half4, size_t out_bytes);
int cs_read_sky_lut(cs_context* ctx, uint16_t* out_half4, size_t out_bytes);
int cs_write_transmittance_lut(cs_context* ctx, const uint16_t* half4, size_t bytes);
int cs_write_sky_lut(cs_context* ctx, const uint16_t* half4, size_t bytes);

/* ---- the cloud march (set 0 binding 0 output; cloud_sky.gd:234-248) ----------------------- */

/* (Re)allocate the RGBA16F output image (cloud_sky.gd:368-376,399).  W != H is allowed
 * (clouds.glsl:19 texture_size is a vec2). */
int cs_resize(cs_context* ctx, int width, int height);
/* Step-count parametrisation (extension; the reference is fixed at 128 primary steps,
 * clouds.glsl:228, and 6 cone + 1 distant light samples, clouds.glsl:186-199).
 * mode = CS_MODE_STRICT, or CS_MODE_FAST optionally OR-ed with CS_MODE_EARLY_OUT and with one of CS_MODE_TEX / CS_MODE_HALF.
 * SURVEY §8(d) rule: cone sample j uses RANDOM_VECTORS[j % 6] * j and mip j, LODs clamp to the last mip. */
int cs_set_march_config(cs_context* ctx, int primary_steps, int cone_samples, int mode);
/* Enable/disable the device work counters (slower instrumented kernel when enabled). */
/* Adaptive primary step count per direction (extension; the reference's own unimplemented hint "Take fewer steps towards
 * horizon", clouds.glsl:227, and BASELINE configs[4]'s "hierarchical/adaptive step"): never step finer than
 * min_step_length_m along a ray —
 *     steps(dir) = clamp(ceil(shell_length(dir) / min_step_length_m), min_steps, primary_steps)
 * so directions whose slab crossing is short (towards the zenith: 2.5 km vs 84.8 km at the horizon) take fewer than
 * primary_steps steps.  min_step_length_m = 0 (default) restores the reference's fixed count.  NOT reference behaviour: it
 * changes the sample positions.  Measured against the fixed-count render: inside the parity tolerance
 * for optically thick skies (coverage 1.0: 100 % of pixels at 19.53 m, the reference's own zenith step), outside it for thin
 * cloud (coverage 0.2).  cs_counters reports the executed steps. */
int cs_set_step_budget(cs_context* ctx, float min_step_length_m, int min_steps);
int cs_set_counters_enabled(cs_context* ctx, int enabled);
int cs_get_counters(cs_context* ctx, cs_counters* out);

/* Replaces compute_list_dispatch(num_workgroups, num_workgroups, 1) of clouds.glsl with
 * 8x8 groups (cloud_sky.gd:247, clouds.glsl:5,258-266): renders pixels
 * [update_position, update_position + 8*groups) clipped to the image (the reference does
 * not bounds-check; this does).  Asynchronous on the context's stream. */
int cs_dispatch_clouds(cs_context* ctx, const cs_cloud_params* params, int groups_x, int groups_y);
/* Whole image in one dispatch (north_star's "single dispatch"); params->update_position is
 * ignored (treated as 0,0). */
int cs_render_frame(cs_context* ctx, const cs_cloud_params* params);
/* Same, writing into caller-owned DEVICE memory (tightly packed half4[W*H]); used to render
 * straight into a shard of a gathered buffer.  rows [row_begin,row_end) only. */
int cs_render_rows_to(cs_context* ctx, const cs_cloud_params* params, int row_begin, int row_end,
                      void* device_out_half4);
/* Device pointer of the context-owned output image (Texture2DRD handle, cloud_sky.gd:235). */
/* Interleaved row bands of one frame in ONE launch: bands b = 0 .. n_bands-1 cover rows
 * [first_row + b * band_pitch_rows, + band_rows) (clipped to the image); band_rows is a multiple of 8 (the 8x8 workgroup of
 * clouds.glsl:5 / one CTA row).  This is the multi-GPU strong-scaling shard of a single frame: rank r of N
 * calls it with first_row = r * band_rows, band_pitch_rows = N * band_rows, so the lit fraction — which varies with the
 * row — is spread evenly over the ranks; tiles are independent like the reference's update_position tiles
 * (cloud_sky.gd:156-161), so the assembled image is bit-identical to one full dispatch. */
int cs_render_row_bands_to(cs_context* ctx, const cs_cloud_params* params, int first_row, int band_rows, int band_pitch_rows,
                           int n_bands, void* device_out_half4);
void* cs_image_device_ptr(cs_context* ctx);
/* Copy the context-owned image to host: tightly packed half4[W*H], row 0 = uv.y 0. */
int cs_read_image(cs_context* ctx, uint16_t* out_half4, size_t out_bytes);
/* End-to-end convenience with HOST buffers: sky LUT for params->light_direction, full-frame
 * render, device->host copy into (pinned or pageable) out_half4, synchronised on return.
 * This is the _update_per_frame_data + _render_process pair (cloud_sky.gd:165-187,234-248). */
int cs_render_frame_host(cs_context* ctx, const cs_cloud_params* params, uint16_t* out_half4,
                         size_t out_bytes);
/* Streaming variant of cs_render_frame_host for back-to-back frames: the kernels of frame k+1 overlap the
 * device->host copy of frame k (two device images, a copy stream, events).  The call returns as soon as the work
 * is queued; out_half4 (pinned memory recommended) is valid after cs_wait_host(ctx).  At most two frames are in
 * flight; a third call waits for the oldest copy.  This is update_sky's steady state (cloud_sky.gd:129-163) for
 * a caller that consumes the texture on the host. */
int cs_render_frame_host_async(cs_context* ctx, const cs_cloud_params* params, uint16_t* out_half4, size_t out_bytes);
int cs_wait_host(cs_context* ctx);
/* Sun-angle batch (BASELINE config 4): for each of n suns build its sky LUT and render one full
 * frame into device_out_half4 + i*W*H*4 halfs.  Other params are shared.  In CS_MODE_FAST (no flags, counters and kernel
 * timing off) up to 4 suns are marched per launch: everything a primary step computes before its light march is
 * sun-independent and is computed once; every image is bit-identical to a single-sun cs_render_frame for that sun. */
int cs_render_sun_batch_to(cs_context* ctx, const cs_cloud_params* params, const float* sun_dirs_xyz,
                           int n_suns, void* device_out_half4);
/* Device-side timing of `iters` back-to-back full-frame dispatches with CUDA events on the
 * context's stream (after `warmup` untimed ones).  Returns average milliseconds per dispatch. */
int cs_time_render_frame(cs_context* ctx, const cs_cloud_params* params, int warmup, int iters,
                         float* out_ms_avg);

/* ---- multi-GPU: fused all-gather through peer-mapped output replicas -------------------------
 * The reference renders its texture as independent tiles addressed by update_position (cloud_sky.gd:156-161) and every
 * pixel depends on nothing but the push constants and the textures (clouds.glsl:258-266), so ranks (one process per GPU)
 * render disjoint tiles / sun angles with no exchange while rendering.  The gather of the finished texture is fused into
 * the march kernel: each rank holds a full copy of the gathered buffer, maps its peers' copies (CUDA IPC over
 * NVLink/NVSwitch peer access) and the kernel stores every finished pixel into all copies.  cs_peer_barrier is what is
 * left of the collective: "my stores have landed everywhere, and so have everyone else's".
 * Host protocol per rank: cs_peer_alloc (buffer + 64-byte handle) -> exchange handles out of band (torch.distributed,
 * MPI, a file ...) -> cs_peer_open each peer's handle -> cs_set_output_mirrors -> render with the ordinary calls
 * (cs_render_rows_to / cs_render_sun_batch_to into the own copy) -> cs_peer_barrier.  CUDA-only. */
#define CS_IPC_HANDLE_BYTES 64
/* cudaMalloc'ed, zero-filled, exportable device buffer + its IPC handle. */
int cs_peer_alloc(cs_context* ctx, size_t bytes, void** out_device_ptr, uint8_t out_handle[CS_IPC_HANDLE_BYTES]);
/* Map a peer's buffer (the handle came from cs_peer_alloc in another process on the same node). */
int cs_peer_open(cs_context* ctx, const uint8_t handle[CS_IPC_HANDLE_BYTES], void** out_device_ptr);
int cs_peer_close(cs_context* ctx, void* opened_device_ptr);
int cs_peer_free(cs_context* ctx, void* allocated_device_ptr);
/* From now on every march dispatch whose output pointer lies inside [base, base + bytes) also stores each finished pixel
 * at the same offset into mirror_bases[0 .. n_mirrors) (peer-mapped copies; n_mirrors <= 7).  n_mirrors = 0 turns it off. */
int cs_set_output_mirrors(cs_context* ctx, void* base, size_t bytes, int n_mirrors, void* const* mirror_bases);
/* Stream-ordered completion barrier of the fused gather: flag_arrays[k] is rank k's flag array (>= world uint32, from
 * cs_peer_alloc; the own one at [rank], peers' opened with cs_peer_open).  Queues one tiny kernel that publishes `epoch`
 * (any value increasing by < 2^31 per call) to every rank and waits until every rank published it.  After it, on this
 * context's stream, the own copy holds every rank's pixels.  A rank that never arrives trips a ~20 s watchdog, reported by
 * cs_peer_check (which synchronises the stream). */
int cs_peer_barrier(cs_context* ctx, int rank, int world, void* const* flag_arrays, uint32_t epoch);
int cs_peer_check(cs_context* ctx);

/* ---- time-sliced update + temporal blend: the Sky resource of cloud_sky.gd ---------- */

/* One instance of cloud_sky.gd on a context: three hemisphere textures (render target / blend from / blend to,
 * cloud_sky.gd:213-218), three sky LUTs (sky_lut.gd:15-18), the FrameData accumulators and the tile walk. */
typedef struct cs_sky cs_sky;
typedef struct cs_sky_frame {
    int32_t frame;                 /* tiles rendered so far into the texture being updated (cloud_sky.gd:94)      */
    int32_t frames_to_update;
    int32_t texture_size;
    int32_t update_position[2];    /* origin of the NEXT tile (cloud_sky.gd:82)                                   */
    int32_t update_region_size, num_workgroups;                       /* cloud_sky.gd:83-84                       */
    int32_t texture_to_update, texture_to_blend_from, texture_to_blend_to;   /* cloud_sky.gd:87-89                */
    float blend_amount;            /* frame / frames_to_update as set by the last update (cloud_sky.gd:152)       */
    int32_t sky_current_texture;   /* sky_lut.gd:18                                                               */
    int32_t sky_blend_from, sky_blend_to;  /* indices of back_texture[0], back_texture[1] (sky_lut.gd:143-146)    */
    int32_t sky_updates;           /* number of sky-LUT renders so far                                            */
    void* cloud_textures[3];       /* device pointers: half4[texture_size * texture_size]                         */
    void* sky_luts[3];             /* device pointers: half4[200 * 100]                                           */
    cs_frame_state frame_data;     /* FrameData snapshot used by the texture being updated                        */
} cs_sky_frame;
/* load("clouds_sky.tres") + delayed_init (cloud_sky.gd:99-107): allocate the textures for settings->texture_size.
 * The context must already hold the input textures and the transmittance LUT.  A sky owns its three textures and
 * three sky LUTs and renders into them with its own texture size: the context's image and cs_resize() state are
 * never changed by a sky, so several skies of different sizes may share one context.  Ownership: a cs_sky keeps a
 * pointer to its context — destroy every sky BEFORE cs_destroy(ctx) (cleanup() order of cloud_sky.gd:197-212). */
int cs_sky_create(cs_context* ctx, const cs_sky_settings* settings, cs_sky** out_sky);
void cs_sky_destroy(cs_sky* sky);
/* Property setters (cloud_sky.gd:4-50).  A change of texture_size or frames_to_update runs cleanup() +
 * update_performance() + request_full_sky_init() like the GDScript setters (cloud_sky.gd:37-50). */
int cs_sky_set_settings(cs_sky* sky, const cs_sky_settings* settings);
/* sun.gd:11-13 + FrameData.update_light_data (cloud_sky.gd:76-79): attach / update the sun. */
int cs_sky_set_sun(cs_sky* sky, const float basis_columns[9], float energy, const float color_srgb[3]);
/* update_sky() (cloud_sky.gd:129-163) with the clock passed in: renders ONE tile (2 * frames_to_update + 1 tiles
 * on the first call after a full-init request), rotates the textures when a texture is complete, refreshes the
 * sky LUT once per texture. */
int cs_sky_update(cs_sky* sky, float now_seconds);
int cs_sky_get_frame(cs_sky* sky, cs_sky_frame* out);
/* Host copy of one of the three hemisphere textures (index 0..2). */
int cs_sky_read_texture(cs_sky* sky, int index, uint16_t* out_half4, size_t out_bytes);

/* ---- presentation composite: the sky material shader (clouds.gdshader) ------------------- */

/* Which directions to shade.  Equirect: pixel (x, y) looks toward azimuth a = (x+.5)/W*2pi - pi, elevation
 * e = pi/2 - (y+.5)/H*pi, EYEDIR = (sin a cos e, sin e, -cos a cos e) (Godot axes: +Y up, -Z forward).
 * Perspective: Godot camera convention, EYEDIR = normalize(basis * (ndc.x * tan(fov/2) * aspect, ndc.y * tan(fov/2), -1)). */
#define CS_VIEW_EQUIRECT 0
#define CS_VIEW_PERSPECTIVE 1
typedef struct cs_view {
    int32_t projection;      /* CS_VIEW_EQUIRECT or CS_VIEW_PERSPECTIVE                                       */
    int32_t width, height;   /* output size in pixels                                                       */
    float basis_columns[9];  /* camera basis (x right, y up, z back), perspective only                      */
    float fov_y_degrees;     /* perspective only                                                            */
    float sun_direction[3];  /* LIGHT0_DIRECTION: unit vector toward the sun                                */
    float sun_disk_scale;    /* uniform sun_disk_scale (clouds.gdshader:13, cloud_sky.gd:27-31)             */
    float blend_amount;      /* uniform blend_amount (clouds.gdshader:12, cloud_sky.gd:152)                 */
} cs_view;
/* sky() of clouds.gdshader:104-116 for every pixel of the view: EYEDIR -> hemi-octahedral lookup of the two cloud
 * textures (vec3_to_oct, :22-32) blended by blend_amount, get_atmo (:87-102: two sky LUTs / 50, sun disk with bloom
 * :48-59 attenuated by the transmittance LUT :77-85, hidden below the horizon :62-71), composite and horizon fade
 * (:114-115).  clouds_from/to: device half4[tex_w*tex_h]; sky_from/to: device half4[200*100]; out: device
 * float4[width*height] (linear radiance, alpha 1).  Needs the transmittance LUT. */
int cs_composite(cs_context* ctx, const cs_view* view, const void* clouds_from, const void* clouds_to, int tex_w, int tex_h,
                 const void* sky_from, const void* sky_to, float* out_rgba32f_device);
/* The same for a Sky resource's current blend textures / sky-LUT back buffers / blend_amount (view->blend_amount is
 * ignored), copied to host memory: what the viewport shows for that camera (before tonemapping). */
int cs_sky_composite_host(cs_sky* sky, const cs_view* view, float* out_rgba32f_host, size_t out_bytes);

/* Per-kernel device timing: when enabled, every sky-LUT build and every cloud-march launch is
 * bracketed by CUDA events on the context's stream.  cs_read_kernel_timings synchronises, returns
 * the summed milliseconds and launch counts since the last read, and resets them. */
int cs_set_kernel_timing(cs_context* ctx, int enabled);
int cs_read_kernel_timings(cs_context* ctx, float* march_ms_sum, int* march_launches, float* sky_ms_sum,
                           int* sky_launches);

/* ---- host-side parameter logic (pure CPU, no context) ------------------------------------ */

/* Script defaults of cloud_sky.gd:4-50 (coverage 0.25, white ground, 768, 64 frames). */
void cs_settings_default(cs_sky_settings* s);
/* Demo resource values of clouds_sky.tres:11-18 (coverage 0.2, brown ground, sun_disk 2). */
void cs_settings_demo(cs_sky_settings* s);
/* FrameData initial values (cloud_sky.gd:66-74): zero offsets, light (0,-1,0), energy 1, white. */
void cs_frame_state_init(cs_frame_state* st);
/* FrameData.update_light_data (cloud_sky.gd:76-79): direction = normalize(basis * (0,0,1)) with
 * basis given as 3 column vectors (Godot Basis x,y,z axes); colour sRGB -> linear. */
void cs_frame_state_set_light(cs_frame_state* st, const float basis_columns[9], float energy,
                              const float color_srgb[3]);
/* _update_per_frame_data (cloud_sky.gd:165-187) with the clock passed in: integrates the three
 * wind offsets from st->time to abs_time_seconds and stores the new time. */
void cs_frame_advance(cs_frame_state* st, const cs_sky_settings* s, float abs_time_seconds);
/* _fill_push_constant (cloud_sky.gd:251-289). */
void cs_fill_cloud_params(cs_cloud_params* out, const cs_sky_settings* s, const cs_frame_state* st,
                          int width, int height, int update_x, int update_y);
/* update_performance (cloud_sky.gd:109-118): region = texture_size / isqrt(frames_to_update),
 * texture_size coerced to a multiple, groups = ceil(region / 8). */
void cs_update_performance(int* texture_size_inout, int frames_to_update, int* out_region,
                           int* out_groups);
/* The raster-order tile walk of update_sky (cloud_sky.gd:156-161). */
void cs_next_update_position(int* x_inout, int* y_inout, int region, int texture_size);

#ifdef __cplusplus
}
#endif
#endif /* CLOUDSKY_H */

// Internal: the context object shared by context.cu and sky_resource.cu.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "cs_internal.h"

struct cs_context {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;

    // textures
    bool have_tex = false;
    int large_n = 0, large_levels = 0, small_n = 0, small_levels = 0, weather_w = 0, weather_h = 0;
    int weather_type_hi = 0;
    int records_half = 0;  // 1: d_*_f hold exact-integer fp16 records, 0: fp32 records
    uint32_t* d_large[cs::kMaxLargeLevels] = {};
    uint32_t* d_small[cs::kMaxSmallLevels] = {};
    uint32_t* d_weather = nullptr;
    float* d_large_f[cs::kMaxLargeLevels] = {};
    float* d_small_f[cs::kMaxSmallLevels] = {};
    float* d_weather_f = nullptr;
    // CS_MODE_HALF: the same exact-integer fp16 coefficients, centred and interleaved as half2 pairs (built on first use)
    float* d_large_h2[cs::kMaxLargeLevels] = {};
    float* d_small_h2[cs::kMaxSmallLevels] = {};
    float* d_weather_h2 = nullptr;
    bool have_h2 = false;
    std::vector<uint8_t> h_weather;  // host copy of the RGBA8 weather map (repack)
    // CS_MODE_TEX: the same mip chains as CUDA mipmapped arrays behind texture objects
    cudaMipmappedArray_t a_large = nullptr, a_small = nullptr;
    cudaArray_t a_weather = nullptr;
    cudaTextureObject_t t_large = 0, t_small = 0, t_weather = 0;
    std::vector<std::vector<uint8_t>> h_large, h_small;  // host copies of the mip chains (readback / repack)

    // LUTs
    uint16_t* d_tlut = nullptr;
    uint16_t* d_sky = nullptr;
    bool have_tlut = false, have_sky = false;
    int tlut_param = CS_TLUT_LINEAR;  // cs_set_transmittance_parametrisation
    float* d_frame_consts = nullptr;     // kMaxSunBatch x FrameConsts (a single frame uses the first)
    uint16_t* d_sky_batch = nullptr;     // kMaxSunBatch sky LUTs for cs_render_sun_batch_to
    bool sun_batching = true;            // CLOUDSKY_SUN_BATCH=0 in the environment: one launch per sun (A/B runs)

    // output
    int W = 0, H = 0;
    uint16_t* d_image = nullptr;
    // streaming host readback (cs_render_frame_host_async): second image, copy stream, per-slot events
    uint16_t* d_image2 = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_rendered[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    bool slot_busy[2] = {false, false};
    unsigned async_frame = 0;

    // fused all-gather: peer replicas of a registered output range (cs_set_output_mirrors)
    uint8_t* mirror_base = nullptr;
    size_t mirror_bytes = 0;
    int n_mirrors = 0;
    uint8_t* mirror_peer[cs::kMaxMirrors] = {};
    unsigned* d_peer_err = nullptr;  // set by the peer barrier kernel when a peer never arrives
    long long peer_watchdog_cycles = 0;

    // march config
    int primary_steps = CS_REF_PRIMARY_STEPS, cone_samples = CS_REF_CONE_SAMPLES, mode = CS_MODE_FAST;
    float budget_len = 0.0f;  // cs_set_step_budget
    int budget_min = 1;
    bool counters_on = false;
    unsigned long long* d_counters = nullptr;

    // optional per-kernel event timing (cs_set_kernel_timing)
    bool timing_on = false;
    std::vector<cudaEvent_t> ev_march, ev_sky;  // begin/end pairs, recycled
    size_t n_march = 0, n_sky = 0;              // pairs recorded since the last read
};


namespace cs {
// prologue + march of one pixel rectangle into `out`, sampling `sky_lut` (nullptr = the context's own LUT).  `out` is an image of
// img_w x img_h pixels owned by the caller (params.texture_size is validated against THAT size; the context's own image and size,
// which other users of the context may rely on, are not touched).
int ctx_dispatch(cs_context* c, const cs_cloud_params* P, int x0, int y0, int x1, int y1, uint16_t* out, const uint16_t* sky_lut, int img_w, int img_h);
// sky-LUT kernel into `dst` (device half4[200*100]); requires the transmittance LUT.
int ctx_build_sky_lut_into(cs_context* c, const float sun[3], uint16_t* dst);
int ctx_fail(cs_context* c, int code, const std::string& msg);
}  // namespace cs

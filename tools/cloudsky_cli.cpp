// cloudsky_cli — thin C++ host driver over the C-ABI (include/cloudsky.h).
//
// Loads libcloudsky_b200.so with dlopen (or any other implementation of the header given with
// --lib), decodes the three input bitmaps, builds both atmosphere LUTs, renders the hemisphere
// texture for the given sun / wind settings and writes it as raw half4 (.f16) and/or a tonemapped
// PPM preview.  It plays the role of cloud_sky.gd's _update_per_frame_data + _render_process
// (cloud_sky.gd:165-187,234-248) for a headless caller.  --procedural SEED LARGE_N SMALL_N WEATHER_N replaces the bitmaps by
// noise synthesised on the device (cs_generate_noise, the reference README's TODO 3); --bruneton selects the Bruneton 2017
// transmittance-LUT mapping (README TODO 2); --flags N ORs march-mode flags (2 early out, 4 texture unit, 8 packed-fp16 filter) into the
// mode; --budget LEN MIN sets the adaptive per-direction step count (cs_set_step_budget; clouds.glsl:227's "fewer steps" hint).
#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/cloudsky.h"

#define SYM(name) decltype(&::name) name = (decltype(&::name))dlsym(h, #name); if (!name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

static float half_to_float(uint16_t v) {
    uint32_t s = (v & 0x8000u) << 16, e = (v >> 10) & 31, m = v & 1023, x;
    if (e == 0) { if (!m) x = s; else { int k = -1; do { m <<= 1; k++; } while (!(m & 1024)); x = s | ((uint32_t)(112 - k) << 23) | ((m & 1023) << 13); } }
    else if (e == 31) x = s | 0x7f800000u | (m << 13);
    else x = s | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &x, 4); return f;
}

int main(int argc, char** argv) {
    std::string lib = "godot-volumetric-cloud-demo-v2_b200/csrc/libcloudsky_b200.so", dir = "cloud_sky", out_f16, out_ppm;
    int W = 768, H = 768, P = CS_REF_PRIMARY_STEPS, cone = CS_REF_CONE_SAMPLES, mode = CS_MODE_FAST, device = 0, iters = 1;
    float sun[3] = {0.0f, 1.0f, 0.0f}, time_s = 0.0f, coverage = -1.0f, density = -1.0f, wind_dir = 0.0f, wind_speed = 1.0f;
    int procedural = 0, gen_n[3] = {128, 32, 512}, tlut_mapping = CS_TLUT_LINEAR, flags = 0;
    unsigned seed = 1;
    float budget_len = 0.0f;
    int budget_min = 1;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--lib") lib = next();
        else if (a == "--assets") dir = next();
        else if (a == "--size") { W = atoi(next()); H = atoi(next()); }
        else if (a == "--steps") { P = atoi(next()); cone = atoi(next()); }
        else if (a == "--strict") mode = CS_MODE_STRICT;
        else if (a == "--device") device = atoi(next());
        else if (a == "--sun") { sun[0] = (float)atof(next()); sun[1] = (float)atof(next()); sun[2] = (float)atof(next()); }
        else if (a == "--time") time_s = (float)atof(next());
        else if (a == "--coverage") coverage = (float)atof(next());
        else if (a == "--density") density = (float)atof(next());
        else if (a == "--wind") { wind_dir = (float)atof(next()); wind_speed = (float)atof(next()); }
        else if (a == "--iters") iters = atoi(next());
        else if (a == "--procedural") { procedural = 1; seed = (unsigned)strtoul(next(), nullptr, 0); for (int k = 0; k < 3; k++) gen_n[k] = atoi(next()); }
        else if (a == "--bruneton") tlut_mapping = CS_TLUT_BRUNETON2017;
        else if (a == "--flags") flags = atoi(next());
        else if (a == "--budget") { budget_len = (float)atof(next()); budget_min = atoi(next()); }
        else if (a == "--out") out_f16 = next();
        else if (a == "--ppm") out_ppm = next();
        else {
            printf("usage: cloudsky_cli [--lib so] [--assets dir] [--size W H] [--steps P cone] [--strict] [--device n]\n"
                   "       [--sun x y z] [--time s] [--coverage c] [--density d] [--wind dir_rad speed] [--iters n] [--out img.f16] [--ppm img.ppm]\n"
                   "       [--procedural seed large_n small_n weather_n] [--bruneton] [--flags march_mode_flags] [--budget min_step_m min_steps]\n");
            return a == "--help" ? 0 : 1;
        }
    }
    void* h = dlopen(lib.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "dlopen(%s): %s\n", lib.c_str(), dlerror()); return 2; }
    SYM(cs_create) SYM(cs_destroy) SYM(cs_last_error) SYM(cs_backend_name) SYM(cs_load_texture_files) SYM(cs_build_transmittance_lut)
    SYM(cs_build_sky_lut) SYM(cs_resize) SYM(cs_set_march_config) SYM(cs_render_frame_host) SYM(cs_settings_demo) SYM(cs_frame_state_init)
    SYM(cs_frame_advance) SYM(cs_fill_cloud_params) SYM(cs_generate_noise) SYM(cs_noise_params_default) SYM(cs_upload_textures)
    SYM(cs_set_transmittance_parametrisation) SYM(cs_set_step_budget)

    cs_context* ctx = nullptr;
    if (cs_create(device, &ctx) != CS_OK) { fprintf(stderr, "cs_create failed (backend %s)\n", cs_backend_name()); return 3; }
#define CK(call) do { int r__ = (call); if (r__ != CS_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, r__, cs_last_error(ctx)); cs_destroy(ctx); return 3; } } while (0)
    if (procedural) {  // synthesise the three inputs instead of preload()-ing the bitmaps
        std::vector<uint8_t> tex[3];
        for (int k = 0; k < 3; k++) {
            cs_noise_params np; cs_noise_params_default(k, &np); np.seed = seed;
            tex[k].resize((size_t)gen_n[k] * gen_n[k] * (k == CS_NOISE_WEATHER ? 1 : gen_n[k]) * 4);
            CK(cs_generate_noise(ctx, k, gen_n[k], &np, tex[k].data(), tex[k].size()));
        }
        CK(cs_upload_textures(ctx, tex[0].data(), gen_n[0], 4, tex[1].data(), gen_n[1], 4, tex[2].data(), gen_n[2], gen_n[2], 4));
    } else {
        CK(cs_load_texture_files(ctx, (dir + "/perlworlnoise.tga").c_str(), 128, (dir + "/worlnoise.bmp").c_str(), 32, (dir + "/weather.bmp").c_str()));
    }
    CK(cs_set_transmittance_parametrisation(ctx, tlut_mapping));
    CK(cs_build_transmittance_lut(ctx));
    CK(cs_resize(ctx, W, H));
    CK(cs_set_march_config(ctx, P, cone, mode == CS_MODE_STRICT ? mode : (mode | flags)));
    CK(cs_set_step_budget(ctx, budget_len, budget_min));
    cs_sky_settings s; cs_settings_demo(&s);
    if (coverage >= 0) s.cloud_coverage = coverage;
    if (density >= 0) s.density = density;
    s.wind_direction = wind_dir; s.wind_speed = wind_speed;
    cs_frame_state st; cs_frame_state_init(&st);
    float n = std::sqrt(sun[0] * sun[0] + sun[1] * sun[1] + sun[2] * sun[2]);
    for (int i = 0; i < 3; i++) st.light_direction[i] = sun[i] / n;
    cs_frame_advance(&st, &s, time_s);
    cs_cloud_params p; cs_fill_cloud_params(&p, &s, &st, W, H, 0, 0);
    std::vector<uint16_t> img((size_t)W * H * 4);
    CK(cs_render_frame_host(ctx, &p, img.data(), img.size() * 2));  // warm-up
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; i++) CK(cs_render_frame_host(ctx, &p, img.data(), img.size() * 2));
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / iters;
    printf("backend %s: %dx%d, %d primary / %d+1 light steps: %.3f ms per frame end-to-end (sky LUT + march + D2H), %.1f Mray-steps/s\n",
           cs_backend_name(), W, H, P, cone, sec * 1e3, (double)W * H * P / sec / 1e6);
    if (!out_f16.empty()) { FILE* f = fopen(out_f16.c_str(), "wb"); if (f) { fwrite(img.data(), 2, img.size(), f); fclose(f); } }
    if (!out_ppm.empty()) {
        FILE* f = fopen(out_ppm.c_str(), "wb");
        if (f) {
            fprintf(f, "P6\n%d %d\n255\n", W, H);
            for (size_t i = 0; i < (size_t)W * H; i++) {
                float a = half_to_float(img[i * 4 + 3]);
                const float sky[3] = {0.10f, 0.20f, 0.50f};
                for (int c = 0; c < 3; c++) {
                    float v = (sky[c] * (1 - a) + half_to_float(img[i * 4 + c])) * 2.5f;
                    v = std::pow(std::fmin(std::fmax(v, 0.0f), 1.0f), 1.0f / 2.2f);
                    fputc((int)(v * 255.0f + 0.5f), f);
                }
            }
            fclose(f);
        }
    }
    cs_destroy(ctx);
    dlclose(h);
    return 0;
}

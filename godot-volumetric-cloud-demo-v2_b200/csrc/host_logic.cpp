// Host-side parameter logic of cloud_sky.gd, with the clock passed in instead of read
// (Time.get_ticks_msec(), cloud_sky.gd:174) so that frames are reproducible.
#include <cmath>
#include <cstring>

#include "../../include/cloudsky.h"

extern "C" {

// Generator defaults (cs_generate_noise; README.md:30 TODO).  Chosen so that the products have the rough statistics of the
// reference bitmaps (perlworlnoise R mean ~0.85, weather type 0.59..0.91, coverage spanning 0..1).
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

// Exported property defaults (cloud_sky.gd:9-50).
void cs_settings_default(cs_sky_settings* s) {
    if (!s) return;
    s->wind_direction = 0.0f;   // :10
    s->wind_speed = 1.0f;       // :14
    s->density = 0.05f;         // :20
    s->cloud_coverage = 0.25f;  // :22
    s->time_offset = 0.0f;      // :24
    s->sun_disk_scale = 1.0f;   // :28
    for (int i = 0; i < 4; i++) s->ground_color[i] = 1.0f;  // :33
    s->frames_to_update = 64;   // :37
    s->texture_size = 768;      // :45
}

// The demo scene's overrides (clouds_sky.tres:11-18).
void cs_settings_demo(cs_sky_settings* s) {
    if (!s) return;
    cs_settings_default(s);
    s->cloud_coverage = 0.2f;
    s->sun_disk_scale = 2.0f;
    s->ground_color[0] = 0.270588f;
    s->ground_color[1] = 0.188235f;
    s->ground_color[2] = 0.027451f;
    s->ground_color[3] = 1.0f;
}

// FrameData initialisers (cloud_sky.gd:66-74).
void cs_frame_state_init(cs_frame_state* st) {
    if (!st) return;
    std::memset(st, 0, sizeof(*st));
    st->light_direction[1] = -1.0f;
    st->light_energy = 1.0f;
    st->light_color[0] = st->light_color[1] = st->light_color[2] = 1.0f;
}

static float srgb_channel_to_linear(float c) {  // Godot's Color::srgb_to_linear()
    return c < 0.04045f ? c * (1.0f / 12.92f) : std::pow((c + 0.055f) * (float)(1.0 / (1.0 + 0.055)), 2.4f);
}

// FrameData.update_light_data (cloud_sky.gd:76-79).
void cs_frame_state_set_light(cs_frame_state* st, const float basis_columns[9], float energy, const float color_srgb[3]) {
    if (!st || !basis_columns || !color_srgb) return;
    const float* z = basis_columns + 6;  // basis * Vector3(0, 0, 1) selects the third column
    float len = std::sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
    for (int i = 0; i < 3; i++) st->light_direction[i] = z[i] / len;
    st->light_energy = energy;
    for (int i = 0; i < 3; i++) st->light_color[i] = srgb_channel_to_linear(color_srgb[i]);
}

// _update_per_frame_data (cloud_sky.gd:165-187) minus the sky-LUT refresh, which the caller
// issues with cs_build_sky_lut(ctx, st->light_direction).
void cs_frame_advance(cs_frame_state* st, const cs_sky_settings* s, float now) {
    if (!st || !s) return;
    float wx = std::cos(s->wind_direction), wy = std::sin(s->wind_direction);  // Vector2.from_angle (:168)
    float delta = now - st->time;                                              // :175
    float delta2 = delta * 0.001f + 0.005f * s->time_offset;                   // :176
    float wl = std::sqrt(wx * wx + wy * wy);                                   // .normalized() (:177)
    wx /= wl; wy /= wl;
    st->time = now;                                                            // :180
    st->detailed_pos[0] += delta * wx;                                         // :181
    st->detailed_pos[1] += delta * wy;
    st->cloud_pos[0] += delta * wx * s->wind_speed;                            // :182
    st->cloud_pos[1] += delta * wy * s->wind_speed;
    st->weather_pos[0] += delta2 * wx * s->wind_speed;                         // :183
    st->weather_pos[1] += delta2 * wy * s->wind_speed;
}

// _fill_push_constant (cloud_sky.gd:251-289): same order, same padding.
void cs_fill_cloud_params(cs_cloud_params* o, const cs_sky_settings* s, const cs_frame_state* st, int width, int height,
                          int update_x, int update_y) {
    if (!o || !s || !st) return;
    float* f = reinterpret_cast<float*>(o);
    int k = 0;
    f[k++] = (float)width; f[k++] = (float)height;
    f[k++] = (float)update_x; f[k++] = (float)update_y;
    f[k++] = st->cloud_pos[0]; f[k++] = st->cloud_pos[1];
    f[k++] = st->detailed_pos[0]; f[k++] = st->detailed_pos[1];
    f[k++] = st->weather_pos[0]; f[k++] = st->weather_pos[1];
    f[k++] = 0.0f; f[k++] = 0.0f;
    for (int i = 0; i < 4; i++) f[k++] = s->ground_color[i];
    for (int i = 0; i < 3; i++) f[k++] = st->light_direction[i];
    f[k++] = st->light_energy;
    for (int i = 0; i < 3; i++) f[k++] = st->light_color[i];
    f[k++] = st->time;
    f[k++] = 0.0f;
    f[k++] = s->density;
    f[k++] = s->cloud_coverage;
    f[k++] = s->time_offset;
    static_assert(sizeof(cs_cloud_params) == 28 * sizeof(float), "push constant block must be 112 bytes");
}

// update_performance (cloud_sky.gd:109-118).
void cs_update_performance(int* texture_size, int frames_to_update, int* region, int* groups) {
    if (!texture_size || !region || !groups || frames_to_update < 1) return;
    int side = (int)std::sqrt((double)frames_to_update);
    int r = *texture_size / side;
    if (*texture_size % side != 0) *texture_size = r * side;
    *region = r;
    *groups = (r + 7) / 8;
}

// The tile walk at the end of update_sky (cloud_sky.gd:156-161).
void cs_next_update_position(int* x, int* y, int region, int texture_size) {
    if (!x || !y) return;
    *x += region;
    if (*x >= texture_size) { *x = 0; *y += region; }
    if (*y >= texture_size) { *x = 0; *y = 0; }
}

}  // extern "C"

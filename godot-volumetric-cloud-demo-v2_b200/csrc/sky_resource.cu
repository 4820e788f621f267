// The Sky resource: cloud_sky.gd's time-sliced update and temporal-blend state machine on top of the context
// (SURVEY 8(f)-2).  Three hemisphere textures are rotated exactly like texture_to_update / _blend_from /
// _blend_to (cloud_sky.gd:87-89,137-141), three sky LUTs like sky_lut.gd:15-18,143-146; one call of
// cs_sky_update is one call of update_sky (cloud_sky.gd:129-163) with the clock passed in.
#include <cstring>

#include "cs_context.h"

using namespace cs;

struct cs_sky {
    cs_context* c = nullptr;
    cs_sky_settings s{};
    cs_frame_state fd{};                  // FrameData (cloud_sky.gd:56-79)
    bool have_sun = false;                // `if sun:` (cloud_sky.gd:166)
    float sun_basis[9] = {}, sun_energy = 1.0f, sun_color[3] = {1.0f, 1.0f, 1.0f};
    uint16_t* tex[3] = {nullptr, nullptr, nullptr};
    uint16_t* lut[3] = {nullptr, nullptr, nullptr};
    int texture_size = 0, region = 0, groups = 0;
    int update_position[2] = {0, 0};
    int to_update = 0, blend_from = 1, blend_to = 2;
    int frame = 0;
    float blend_amount = 0.0f;
    bool can_run = false, needs_full_init = true;
    // SkyLUT (sky_lut.gd)
    int sky_current = 0, sky_updates = 0;
    bool sky_needs_full_update = true;
};

namespace {

#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) return ctx_fail(k->c, CS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

void free_textures(cs_sky* k) {
    for (auto& p : k->tex) { if (p) cudaFree(p); p = nullptr; }
}

// cleanup() (cloud_sky.gd:197-212)
void cleanup(cs_sky* k) {
    k->can_run = false;
    k->frame = 0;
    k->to_update = 0; k->blend_from = 1; k->blend_to = 2;
    k->update_position[0] = k->update_position[1] = 0;
    free_textures(k);
}

// update_performance() + _initialize_compute_code (cloud_sky.gd:109-118,355-408)
int update_performance(cs_sky* k) {
    int ts = k->s.texture_size;
    cs_update_performance(&ts, k->s.frames_to_update, &k->region, &k->groups);
    k->s.texture_size = ts;
    k->texture_size = ts;
    if (ts < 8 || ts > 16384) return ctx_fail(k->c, CS_ERR_INVALID, "cs_sky: texture_size out of range");
    cudaSetDevice(k->c->device);
    CU(cudaStreamSynchronize(k->c->stream));
    const size_t bytes = (size_t)ts * ts * 8;
    for (int i = 0; i < 3; i++) {
        CU(cudaMalloc(&k->tex[i], bytes));
        CU(cudaMemsetAsync(k->tex[i], 0, bytes, k->c->stream));  // the reference clears to debug colours (cloud_sky.gd:402)
    }
    k->can_run = true;
    return CS_OK;
}

// SkyLUT.render_lut (sky_lut.gd:122-148)
int render_lut(cs_sky* k) {
    int r = ctx_build_sky_lut_into(k->c, k->fd.light_direction, k->lut[k->sky_current]);
    if (r) return r;
    k->sky_current = (k->sky_current + 1) % 3;
    k->sky_updates++;
    return CS_OK;
}
// SkyLUT.update_lut (sky_lut.gd:43-52)
int update_lut(cs_sky* k) {
    int r = render_lut(k);
    if (r) return r;
    if (k->sky_needs_full_update) {
        if ((r = render_lut(k)) != CS_OK) return r;
        if ((r = render_lut(k)) != CS_OK) return r;
        k->sky_needs_full_update = false;
    }
    return CS_OK;
}

// _update_per_frame_data (cloud_sky.gd:165-187)
int update_per_frame_data(cs_sky* k, float now) {
    if (k->have_sun) cs_frame_state_set_light(&k->fd, k->sun_basis, k->sun_energy, k->sun_color);
    cs_frame_advance(&k->fd, &k->s, now);
    return update_lut(k);
}

// _render_process (cloud_sky.gd:234-248): one tile of texture `to_update` with the LUT rendered last.
int render_process(cs_sky* k) {
    cs_cloud_params p;
    cs_fill_cloud_params(&p, &k->s, &k->fd, k->texture_size, k->texture_size, k->update_position[0], k->update_position[1]);
    const int x0 = k->update_position[0], y0 = k->update_position[1];
    const uint16_t* lut = k->lut[(k->sky_current + 2) % 3];  // sky_uniform_set[(sky_lut.current_texture + 2) % 3] (cloud_sky.gd:242)
    // the sky's own texture size travels with the dispatch: the shared context is never resized under its other users
    return ctx_dispatch(k->c, &p, x0, y0, x0 + 8 * k->groups, y0 + 8 * k->groups, k->tex[k->to_update], lut, k->texture_size, k->texture_size);
}

int update_sky(cs_sky* k, float now);

// initialize_sky (cloud_sky.gd:124-127)
int initialize_sky(cs_sky* k, float now) {
    int r = update_per_frame_data(k, now);
    for (int i = 0; i < k->s.frames_to_update * 2 && r == CS_OK; i++) r = update_sky(k, now);
    return r;
}

// update_sky (cloud_sky.gd:129-163)
int update_sky(cs_sky* k, float now) {
    if (!k->can_run) return CS_OK;  // silently returns, like the reference (cloud_sky.gd:130-131)
    int r;
    if (k->needs_full_init) {
        k->needs_full_init = false;
        if ((r = initialize_sky(k, now)) != CS_OK) return r;
    }
    if (k->frame >= k->s.frames_to_update) {
        k->to_update = (k->to_update + 1) % 3;
        k->blend_from = (k->blend_from + 1) % 3;
        k->blend_to = (k->blend_to + 1) % 3;
        if ((r = update_per_frame_data(k, now)) != CS_OK) return r;  // only once per texture, otherwise tiles get out of sync
        k->frame = 0;
    }
    k->blend_amount = (float)k->frame / (float)k->s.frames_to_update;
    if ((r = render_process(k)) != CS_OK) return r;
    cs_next_update_position(&k->update_position[0], &k->update_position[1], k->region, k->texture_size);
    k->frame += 1;
    return CS_OK;
}

}  // namespace

extern "C" {

int cs_sky_create(cs_context* c, const cs_sky_settings* s, cs_sky** out) {
    if (!c || !s || !out) return CS_ERR_INVALID;
    *out = nullptr;
    if (!c->have_tex) return ctx_fail(c, CS_ERR_NOT_READY, "cs_sky_create: upload the input textures first");
    if (!c->have_tlut) return ctx_fail(c, CS_ERR_NOT_READY, "cs_sky_create: build the transmittance LUT first");
    if (s->frames_to_update < 1) return ctx_fail(c, CS_ERR_INVALID, "cs_sky_create: frames_to_update must be >= 1");
    cs_sky* k = new cs_sky();
    k->c = c;
    k->s = *s;
    cs_frame_state_init(&k->fd);
    cudaSetDevice(c->device);
    for (int i = 0; i < 3; i++) {
        if (cudaMalloc(&k->lut[i], (size_t)CS_SKY_LUT_W * CS_SKY_LUT_H * 8) != cudaSuccess) {
            cs_sky_destroy(k);
            return ctx_fail(c, CS_ERR_CUDA, "cs_sky_create: out of device memory");
        }
        cudaMemsetAsync(k->lut[i], 0, (size_t)CS_SKY_LUT_W * CS_SKY_LUT_H * 8, c->stream);
    }
    int r = update_performance(k);
    if (r) { cs_sky_destroy(k); return r; }
    *out = k;
    return CS_OK;
}

void cs_sky_destroy(cs_sky* k) {
    if (!k) return;
    cudaSetDevice(k->c->device);
    cudaStreamSynchronize(k->c->stream);
    free_textures(k);
    for (auto& p : k->lut) if (p) cudaFree(p);
    delete k;
}

int cs_sky_set_settings(cs_sky* k, const cs_sky_settings* s) {
    if (!k || !s) return CS_ERR_INVALID;
    if (s->frames_to_update < 1) return ctx_fail(k->c, CS_ERR_INVALID, "frames_to_update must be >= 1");
    const bool rebuild = s->texture_size != k->s.texture_size || s->frames_to_update != k->s.frames_to_update;
    k->s = *s;
    if (rebuild) {  // the texture_size / frames_to_update setters (cloud_sky.gd:37-50)
        cleanup(k);
        int r = update_performance(k);
        if (r) return r;
        k->needs_full_init = true;
    }
    return CS_OK;
}

int cs_sky_set_sun(cs_sky* k, const float basis[9], float energy, const float color[3]) {
    if (!k || !basis || !color) return CS_ERR_INVALID;
    memcpy(k->sun_basis, basis, sizeof(k->sun_basis));
    k->sun_energy = energy;
    memcpy(k->sun_color, color, sizeof(k->sun_color));
    if (!k->have_sun) k->needs_full_init = true;  // sun.gd:11-13: cloud_sky.sun = self; request_full_sky_init()
    k->have_sun = true;
    return CS_OK;
}

int cs_sky_update(cs_sky* k, float now) {
    if (!k) return CS_ERR_INVALID;
    return update_sky(k, now);
}

int cs_sky_get_frame(cs_sky* k, cs_sky_frame* o) {
    if (!k || !o) return CS_ERR_INVALID;
    memset(o, 0, sizeof(*o));
    o->frame = k->frame; o->frames_to_update = k->s.frames_to_update; o->texture_size = k->texture_size;
    o->update_position[0] = k->update_position[0]; o->update_position[1] = k->update_position[1];
    o->update_region_size = k->region; o->num_workgroups = k->groups;
    o->texture_to_update = k->to_update; o->texture_to_blend_from = k->blend_from; o->texture_to_blend_to = k->blend_to;
    o->blend_amount = k->blend_amount;
    o->sky_current_texture = k->sky_current;
    o->sky_blend_from = k->sky_current;            // back_texture[0] = texture_rd[current_texture] (sky_lut.gd:145)
    o->sky_blend_to = (k->sky_current + 1) % 3;    // back_texture[1] (sky_lut.gd:146)
    o->sky_updates = k->sky_updates;
    for (int i = 0; i < 3; i++) { o->cloud_textures[i] = k->tex[i]; o->sky_luts[i] = k->lut[i]; }
    o->frame_data = k->fd;
    return CS_OK;
}

int cs_sky_composite_host(cs_sky* k, const cs_view* vw, float* out, size_t bytes) {
    if (!k || !vw || !out) return CS_ERR_INVALID;
    if (vw->width < 1 || vw->height < 1 || bytes != (size_t)vw->width * vw->height * 16) return ctx_fail(k->c, CS_ERR_INVALID, "cs_sky_composite_host: bad buffer size");
    if (!k->tex[0]) return ctx_fail(k->c, CS_ERR_NOT_READY, "cs_sky_composite_host: textures released (cleanup)");
    cudaSetDevice(k->c->device);
    float* d_out = nullptr;
    CU(cudaMalloc(&d_out, bytes));
    cs_view v = *vw;
    v.blend_amount = k->blend_amount;  // sky_material.set_shader_parameter("blend_amount", ...) (cloud_sky.gd:152)
    // blend_from / blend_to textures (cloud_sky.gd:144-145) and the sky LUT back buffers (cloud_sky.gd:147-148, sky_lut.gd:145-146)
    int r = cs_composite(k->c, &v, k->tex[k->blend_from], k->tex[k->blend_to], k->texture_size, k->texture_size, k->lut[k->sky_current],
                         k->lut[(k->sky_current + 1) % 3], d_out);
    cudaError_t e = cudaSuccess;
    if (r == CS_OK) {
        e = cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, k->c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(k->c->stream);
    }
    cudaFree(d_out);
    if (r) return r;
    if (e != cudaSuccess) return ctx_fail(k->c, CS_ERR_CUDA, std::string("cs_sky_composite_host: ") + cudaGetErrorString(e));
    return CS_OK;
}

int cs_sky_read_texture(cs_sky* k, int index, uint16_t* out, size_t bytes) {
    if (!k || !out || index < 0 || index > 2) return CS_ERR_INVALID;
    if (!k->tex[index]) return ctx_fail(k->c, CS_ERR_NOT_READY, "cs_sky_read_texture: textures released (cleanup)");
    if (bytes != (size_t)k->texture_size * k->texture_size * 8) return ctx_fail(k->c, CS_ERR_INVALID, "cs_sky_read_texture: bad buffer size");
    cudaSetDevice(k->c->device);
    CU(cudaMemcpyAsync(out, k->tex[index], bytes, cudaMemcpyDeviceToHost, k->c->stream));
    CU(cudaStreamSynchronize(k->c->stream));
    return CS_OK;
}

}  // extern "C"

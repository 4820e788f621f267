// Presentation composite for sm_100a: the sky material shader clouds.gdshader evaluated for every
// pixel of an equirectangular or perspective view.  One thread per output pixel; all inputs (two hemisphere
// textures, two sky LUTs, the transmittance LUT) are small and cache resident, so this is a latency-scale kernel;
// it is compiled in the accurate configuration (--fmad=false, IEEE div/sqrt) to track the CPU oracle closely.
#include "cs_context.h"
#include "cs_device.cuh"
#include "tlut_param.h"

using namespace csd;

namespace {

constexpr float GD_PI = 3.14159265358979323846f;

struct CompositeArgs {
    cs_view view;
    const uint16_t* clouds_from; const uint16_t* clouds_to; int tex_w, tex_h;
    const uint16_t* sky_from; const uint16_t* sky_to; const uint16_t* tlut;
    int tlut_param;
    float4* out;
};

// clouds.gdshader:22-32
__device__ __forceinline__ void vec3_to_oct(V3 e, float& u, float& v) {
    float s = fabsf(e.x) + fabsf(e.y) + fabsf(e.z);
    e = {e.x / s, e.y / s, e.z / s};
    if (!(e.z >= 0.0f)) {
        float sx = e.x >= 0.0f ? 1.0f : -1.0f, sy = e.y >= 0.0f ? 1.0f : -1.0f;
        float wx = (1.0f - fabsf(e.y)) * sx, wy = (1.0f - fabsf(e.x)) * sy;
        e.x = wx; e.y = wy;
    }
    float ny = e.y * 0.5f + 0.5f;
    u = e.x * 0.5f + ny;
    v = e.x * -0.5f + ny;
}

__device__ __forceinline__ V3 view_direction(const cs_view& vw, int x, int y) {
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
    return normalize3<true>(d);
}

__global__ void __launch_bounds__(128) composite_kernel(const __grid_constant__ CompositeArgs A) {
    const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 8 + (threadIdx.x >> 4);
    const cs_view& vw = A.view;
    if (x >= vw.width || y >= vw.height) return;
    const V3 eyedir = view_direction(vw, x, y);
    // sky() (clouds.gdshader:104-116)
    V3 norm = {eyedir.x, fmaxf(0.0f, eyedir.y), eyedir.z};
    norm = normalize3<true>(norm);
    float u, v;
    vec3_to_oct({norm.x, norm.z, norm.y}, u, v);
    V4 a = sample_lut_half4(A.clouds_from, A.tex_w, A.tex_h, u, v), b = sample_lut_half4(A.clouds_to, A.tex_w, A.tex_h, u, v);
    const float k = vw.blend_amount;
    V4 clouds = {mixf(a.x, b.x, k), mixf(a.y, b.y, k), mixf(a.z, b.z, k), mixf(a.w, b.w, k)};
    // get_atmo (:87-102) -> getValFromSkyLUT (:34-45)
    float phi = atan2f(eyedir.z, eyedir.x), theta = asinf(eyedir.y);
    float su = (phi / GD_PI * 0.5f + 0.5f);
    float sv = sqrtf(fabsf(theta) / (GD_PI * 0.5f)) * signf(theta) * 0.5f + 0.5f;
    V4 sa = sample_lut_half4(A.sky_from, CS_SKY_LUT_W, CS_SKY_LUT_H, su, sv), sb = sample_lut_half4(A.sky_to, CS_SKY_LUT_W, CS_SKY_LUT_H, su, sv);
    V3 col = {mixf(sa.x, sb.x, k) / 50.0f, mixf(sa.y, sb.y, k) / 50.0f, mixf(sa.z, sb.z, k) / 50.0f};
    // sunWithBloom (:48-59)
    V3 sun = {vw.sun_direction[0], vw.sun_direction[1], vw.sun_direction[2]};
    float minSunCosTheta = cosf(vw.sun_disk_scale * 0.53f * GD_PI / 180.0f);
    float cosTheta = dot3(eyedir, sun);
    float lum = 1.0f;
    if (!(cosTheta >= minSunCosTheta)) {
        float offset = minSunCosTheta - cosTheta;
        lum = expf(-offset * 50000.0f) * 0.5f + 1.0f / (0.02f + offset * 300.0f) * 0.01f;
    }
    float sl = smoothstepf<true>(0.002f, 1.0f, lum);
    V3 sunLum = {sl, sl, sl};
    if (length3<true>(sunLum) > 0.0f) {
        const float groundRadiusMM = 6.360f, atmosphereRadiusMM = 6.460f;
        const V3 viewPos = {0.0f, groundRadiusMM + 0.0002f, 0.0f};
        // rayIntersectSphere(viewPos, dir, groundRadiusMM) >= 0 (:61-71)
        float bq = dot3(viewPos, eyedir);
        float cq = dot3(viewPos, viewPos) - groundRadiusMM * groundRadiusMM;
        float hit = -1.0f;
        if (!(cq > 0.0f && bq > 0.0f)) {
            float discr = bq * bq - cq;
            if (!(discr < 0.0f)) hit = discr > bq * bq ? (-bq + sqrtf(discr)) : (-bq - sqrtf(discr));
        }
        if (hit >= 0.0f) {
            sunLum = {0.0f, 0.0f, 0.0f};
        } else {  // getValFromTLUT (:77-85)
            float height = length3<true>(viewPos);
            float c = dot3(V3{viewPos.x / height, viewPos.y / height, viewPos.z / height}, sun);
            float tu = 256.0f * clampf(0.5f + 0.5f * c, 0.0f, 1.0f) / 256.0f;
            float tv = 64.0f * fmaxf(0.0f, fminf(1.0f, (height - groundRadiusMM) / (atmosphereRadiusMM - groundRadiusMM))) / 64.0f;
            float visible = 1.0f;
            if (A.tlut_param == CS_TLUT_BRUNETON2017) tl::bruneton_uv(tv, c, tu, tv, visible);  // same (mu, normalised altitude), other mapping
            V4 t = sample_lut_half4(A.tlut, CS_TRANSMITTANCE_W, CS_TRANSMITTANCE_H, tu, tv) * visible;
            sunLum = {sunLum.x * t.x, sunLum.y * t.y, sunLum.z * t.z};
        }
    }
    V3 background = col + sunLum;
    V3 color = background * (1.0f - clouds.w) + V3{clouds.x, clouds.y, clouds.z};
    float f = smoothstepf<true>(0.6f, 1.0f, 1.0f - eyedir.y);
    V3 o = mix3({clampf(color.x, 0.0f, 100.0f), clampf(color.y, 0.0f, 100.0f), clampf(color.z, 0.0f, 100.0f)},
                {clampf(background.x, 0.0f, 100.0f), clampf(background.y, 0.0f, 100.0f), clampf(background.z, 0.0f, 100.0f)}, f);
    A.out[(size_t)y * vw.width + x] = make_float4(o.x, o.y, o.z, 1.0f);
}

}  // namespace

extern "C" int cs_composite(cs_context* c, const cs_view* vw, const void* cf, const void* ct, int tw, int th, const void* sf, const void* st,
                            float* out) {
    if (!c || !vw || !cf || !ct || !sf || !st || !out) return CS_ERR_INVALID;
    if (!c->have_tlut) return cs::ctx_fail(c, CS_ERR_NOT_READY, "cs_composite: build the transmittance LUT first");
    if (vw->width < 1 || vw->height < 1 || tw < 1 || th < 1 || (vw->projection != CS_VIEW_EQUIRECT && vw->projection != CS_VIEW_PERSPECTIVE))
        return cs::ctx_fail(c, CS_ERR_INVALID, "cs_composite: bad view");
    if (cudaSetDevice(c->device) != cudaSuccess) return cs::ctx_fail(c, CS_ERR_CUDA, "cudaSetDevice");
    CompositeArgs A{*vw, (const uint16_t*)cf, (const uint16_t*)ct, tw, th, (const uint16_t*)sf, (const uint16_t*)st, c->d_tlut, c->tlut_param, (float4*)out};
    dim3 block(128), grid((vw->width + 15) / 16, (vw->height + 7) / 8);
    composite_kernel<<<grid, block, 0, c->stream>>>(A);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cs::ctx_fail(c, CS_ERR_CUDA, std::string("composite_kernel: ") + cudaGetErrorString(e));
    return CS_OK;
}

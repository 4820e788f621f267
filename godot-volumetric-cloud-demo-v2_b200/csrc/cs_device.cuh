// Device-side helpers shared by the LUT kernels and the cloud kernels.
// STRICT = true selects IEEE division / sqrt / libm-grade transcendentals and is meant for
// translation units compiled with --fmad=false (bit-for-bit the oracle's operation order);
// STRICT = false selects the fast intrinsics (MUFU paths) and lets nvcc contract FMAs.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace csd {

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

__device__ __forceinline__ V3 v3(float x, float y, float z) { return {x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
__device__ __forceinline__ V4 operator-(V4 a, V4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
__device__ __forceinline__ V4 operator*(V4 a, V4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
__device__ __forceinline__ V4 operator*(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
__device__ __forceinline__ V4 splat4(float s) { return {s, s, s, s}; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

template <bool STRICT> __device__ __forceinline__ float fdiv(float a, float b) {
    if constexpr (STRICT) return a / b; else return __fdividef(a, b);
}
template <bool STRICT> __device__ __forceinline__ float fsqrt(float a) {
    if constexpr (STRICT) return sqrtf(a); else return __fsqrt_rn(a);
}
template <bool STRICT> __device__ __forceinline__ float fexp(float a) {
    if constexpr (STRICT) return expf(a); else return __expf(a);
}
template <bool STRICT> __device__ __forceinline__ float fpow(float a, float b) {
    if constexpr (STRICT) return powf(a, b); else return exp2f(b * __log2f(a));  // a >= 0 here; log2(0) = -inf -> 0
}
template <bool STRICT> __device__ __forceinline__ float flog(float a) {
    if constexpr (STRICT) return logf(a); else return __logf(a);
}
template <bool STRICT> __device__ __forceinline__ float length3(V3 a) { return fsqrt<STRICT>(dot3(a, a)); }
template <bool STRICT> __device__ __forceinline__ V3 normalize3(V3 a) {
    float l = length3<STRICT>(a);
    return {fdiv<STRICT>(a.x, l), fdiv<STRICT>(a.y, l), fdiv<STRICT>(a.z, l)};
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ V3 mix3(V3 a, V3 b, float t) { return {mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)}; }
__device__ __forceinline__ float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
__device__ __forceinline__ float lerpf(float a, float b, float f) { return a + (b - a) * f; }
template <bool STRICT> __device__ __forceinline__ float smoothstepf(float e0, float e1, float x) {
    float t = clampf(fdiv<STRICT>(x - e0, e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

__device__ __forceinline__ uint16_t f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }

// Bilinear CLAMP_TO_EDGE fetch from a tightly packed half4 LUT (normalised coordinates).
__device__ __forceinline__ V4 sample_lut_half4(const uint16_t* __restrict__ lut, int w, int h, float su, float sv) {
    float ux = su * (float)w - 0.5f, uy = sv * (float)h - 0.5f;
    float fx0 = floorf(ux), fy0 = floorf(uy);
    float fx = ux - fx0, fy = uy - fy0;
    int x0 = (int)fx0, y0 = (int)fy0;
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = min(max(x0, 0), w - 1); x1 = min(max(x1, 0), w - 1);
    y0 = min(max(y0, 0), h - 1); y1 = min(max(y1, 0), h - 1);
    const uint2* l2 = reinterpret_cast<const uint2*>(lut);
    uint2 t00 = __ldg(l2 + y0 * w + x0), t10 = __ldg(l2 + y0 * w + x1);
    uint2 t01 = __ldg(l2 + y1 * w + x0), t11 = __ldg(l2 + y1 * w + x1);
    auto ch = [](uint2 t, int c) -> float {
        uint32_t wv = c < 2 ? t.x : t.y;
        return h2f((uint16_t)((c & 1) ? (wv >> 16) : (wv & 0xffffu)));
    };
    float o[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        float a = lerpf(ch(t00, c), ch(t10, c), fx);
        float b = lerpf(ch(t01, c), ch(t11, c), fx);
        o[c] = lerpf(a, b, fy);
    }
    return {o[0], o[1], o[2], o[3]};
}

}  // namespace csd

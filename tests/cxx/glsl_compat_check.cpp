// Unit checks of oracle/glsl_compat.h against the GLSL 4.50 specification, independent of any shader.
// Built and run by tests/test_glsl_compat.py (g++ -fsingle-precision-constant -ffp-contract=off, like oracle/build_ref.sh).
#include <cstdio>
#include <vector>

#include "glsl_compat.h"
#undef layout
#undef uniform
#undef restrict
#undef writeonly
#undef in
#undef main

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

namespace check {  // the shader translation units do the same: a namespace of their own with using-declarations of the built-ins
GLSL_USING_BUILTINS
using glsl::f16_to_f32; using glsl::f32_to_f16_rne; using glsl::GLSL_ADDR_CLAMP_TO_EDGE; using glsl::GLSL_ADDR_REPEAT; using glsl::GLSL_FMT_RGBA8_UNORM;

int run() {
    // literals are single precision under -fsingle-precision-constant (GLSL: an unsuffixed literal is a float)
    CHECK(sizeof(0.1) == 4);
    CHECK(0.1 + 0.2 == 0.1f + 0.2f);

    // swizzles: reads, write-through, compound assignment, rvalue swizzle, ternary with a swizzle operand (clouds.glsl:253)
    vec3 p(1.0, 2.0, 3.0);
    vec2 xz = p.xz;
    CHECK(xz.x == 1.0 && xz.y == 3.0);
    p.xz += vec2(10.0, 20.0);
    CHECK(p.x == 11.0 && p.y == 2.0 && p.z == 23.0);
    p.xz -= vec2(1.0, 1.0) * 2.0;
    CHECK(p.x == 9.0 && p.z == 21.0);
    vec3 s = p.xzy;
    CHECK(s.x == 9.0 && s.y == 21.0 && s.z == 2.0);
    vec4 q(1.0, 2.0, 3.0, 4.0);
    vec3 rgb = q.rgb;
    CHECK(rgb.r == 1.0 && rgb.g == 2.0 && rgb.b == 3.0 && q.a == 4.0);
    vec3 n(0.25, -0.5, -1.0);
    n.xy = n.z >= 0.0 ? n.xy : vec2(7.0, 8.0);
    CHECK(n.x == 7.0 && n.y == 8.0 && n.z == -1.0);
    vec2 v(3.0, -4.0);
    vec2 a = abs(v.yx);
    CHECK(a.x == 4.0 && a.y == 3.0);
    CHECK(vec4(vec3(1.0, 2.0, 3.0), 9.0).w == 9.0 && vec3(5.0).y == 5.0);

    // operators are component-wise, scalar on either side
    vec3 w = 2.0 * vec3(1.0, 2.0, 3.0) * vec3(1.0, 0.5, 2.0) / 2.0 + 1.0 - vec3(1.0);
    CHECK(w.x == 1.0 && w.y == 1.0 && w.z == 6.0);
    vec3 m1 = -vec3(1.0, -2.0, 0.0);
    CHECK(m1.x == -1.0 && m1.y == 2.0);

    // built-ins (GLSL 4.50 section 8)
    CHECK(mix(2.0, 4.0, 0.25) == 2.5 && mix(vec3(0.0), vec3(8.0), 0.5).z == 4.0);
    CHECK(clamp(1.5, 0.0, 1.0) == 1.0 && clamp(-1.0, 0.0, 1.0) == 0.0);
    CHECK(smoothstep(0.0, 1.0, 0.5) == 0.5 && smoothstep(0.0, 2.0, 3.0) == 1.0 && smoothstep(1.0, 2.0, 0.0) == 0.0);
    CHECK(smoothstep(0.0, 1.0, 0.25) == 0.25f * 0.25f * (3.0f - 2.0f * 0.25f));
    CHECK(fract(1.75) == 0.75 && fract(-0.25) == 0.75);  // x - floor(x)
    CHECK(sign(-3.0) == -1.0 && sign(0.0) == 0.0 && sign(2.0) == 1.0);
    CHECK(atan(1.0, 0.0) > 1.5707 && atan(1.0, 0.0) < 1.5709 && atan(0.0, -1.0) > 3.1415);  // atan(y, x)
    CHECK(dot(vec3(1.0, 2.0, 3.0), vec3(4.0, 5.0, 6.0)) == 32.0 && length(vec3(3.0, 4.0, 0.0)) == 5.0);
    CHECK(normalize(vec3(0.0, 0.0, 2.0)).z == 1.0);
    CHECK(max(vec4(1.0, -1.0, 3.0, 0.0), 0.5).y == 0.5 && exp(vec4(0.0)).w == 1.0);
    CHECK(pow(2.0, 3.0) == 8.0 && int(3.99) == 3);

    // mat4x3: 4 columns of 3 rows, column-major constructor; M * v = sum of column_i * v_i
    const mat4x3 M(1.0, 2.0, 3.0, 10.0, 20.0, 30.0, 100.0, 200.0, 300.0, 1000.0, 2000.0, 3000.0);
    vec3 mv = M * vec4(1.0, 1.0, 1.0, 1.0);
    CHECK(mv.x == 1111.0 && mv.y == 2222.0 && mv.z == 3333.0);
    vec3 mc = M * vec4(0.0, 0.0, 1.0, 0.0);
    CHECK(mc.x == 100.0 && mc.y == 200.0 && mc.z == 300.0);

    // ivec2 conversions truncate toward zero; vec2(ivec2)
    CHECK(ivec2(vec2(3.9, -2.9)).x == 3 && ivec2(vec2(3.9, -2.9)).y == -2 && vec2(ivec2(5, 7)).y == 7.0);

    // fp16: round to nearest even at ties, exact decode
    CHECK(f32_to_f16_rne(1.0f) == 0x3c00 && f16_to_f32(0x3c00) == 1.0f);
    CHECK(f32_to_f16_rne(1.0f + 1.0f / 2048.0f) == 0x3c00);           // tie -> even (mantissa 0)
    CHECK(f32_to_f16_rne(1.0f + 3.0f / 2048.0f) == 0x3c02);           // tie -> even (mantissa 2)
    CHECK(f32_to_f16_rne(65520.0f) == 0x7c00 && f32_to_f16_rne(-0.0f) == 0x8000);
    CHECK(f16_to_f32(f32_to_f16_rne(6.1035156e-5f)) == 6.1035156e-5f && f32_to_f16_rne(2.98e-8f) == 0x0000 && f32_to_f16_rne(5.97e-8f) == 0x0001);

    // texture(): texel centres at (i + 0.5) / N, REPEAT wraps (also negative coordinates), CLAMP_TO_EDGE clamps, fp32 lerp weights
    std::vector<uint8_t> t(4 * 2 * 4, 0);  // 4 x 2 RGBA8, R channel = 0, 85, 170, 255 in row 0 and 255 in row 1
    for (int x = 0; x < 4; x++) { t[x * 4] = (uint8_t)(85 * x); t[(4 + x) * 4] = 255; }
    sampler2D rep; rep.texels = t.data(); rep.w = 4; rep.h = 2; rep.format = GLSL_FMT_RGBA8_UNORM; rep.address = GLSL_ADDR_REPEAT;
    CHECK(texture(rep, vec2(0.125, 0.25)).r == 0.0);                          // centre of texel (0, 0)
    CHECK(texture(rep, vec2(0.375, 0.25)).r == 85.0f / 255.0f);              // centre of texel (1, 0)
    CHECK(texture(rep, vec2(0.25, 0.25)).r == 0.0f + (85.0f / 255.0f - 0.0f) * 0.5f);  // halfway between texels 0 and 1
    CHECK(texture(rep, vec2(0.0, 0.25)).r == 1.0f + (0.0f - 1.0f) * 0.5f);   // REPEAT: halfway between texel 3 (wrapped) and texel 0
    CHECK(texture(rep, vec2(-0.875, 0.25)).r == texture(rep, vec2(0.125, 0.25)).r && texture(rep, vec2(1.375, 2.25)).r == texture(rep, vec2(0.375, 0.25)).r);
    sampler2D cl = rep; cl.address = GLSL_ADDR_CLAMP_TO_EDGE;
    CHECK(texture(cl, vec2(0.0, 0.25)).r == 0.0 && texture(cl, vec2(1.0, 0.25)).r == 1.0 && texture(cl, vec2(-3.0, 0.25)).r == 0.0);
    // textureLod(sampler3D): integer LOD picks one level, clamps to the last level; REPEAT in all three axes
    std::vector<uint8_t> l0(2 * 2 * 2 * 4, 0), l1(4, 200);
    l0[0] = 255;  // texel (0,0,0).r
    sampler3D vol; vol.n = 2; vol.levels = 2; vol.level[0] = l0.data(); vol.level[1] = l1.data();
    CHECK(textureLod(vol, vec3(0.25, 0.25, 0.25), 0.0).r == 1.0 && textureLod(vol, vec3(0.75, 0.25, 0.25), 0.0).r == 0.0);
    CHECK(textureLod(vol, vec3(0.5, 0.25, 0.25), 0.0).r == 0.5 && textureLod(vol, vec3(0.0, 0.25, 0.25), 0.0).r == 0.5);  // wraps to texel 1
    CHECK(textureLod(vol, vec3(0.3, 0.9, 0.1), 1.0).r == 200.0f / 255.0f && textureLod(vol, vec3(0.3, 0.9, 0.1), 7.0).r == 200.0f / 255.0f);
    CHECK(textureLod(vol, vec3(0.25, 0.25, 0.25), -2.0).r == 1.0);  // clouds.glsl:117 passes mip - 2: negative LODs sample level 0

    // imageStore: rgba16f, out-of-range stores are discarded (sky-lut.glsl:281 dispatches one row too many)
    std::vector<uint16_t> img(2 * 2 * 4, 0x1234);
    image2D im; im.texels = img.data(); im.w = 2; im.h = 2;
    imageStore(im, ivec2(1, 1), vec4(1.0, 0.5, -2.0, 65504.0));
    CHECK(img[12] == 0x3c00 && img[13] == 0x3800 && img[14] == 0xc000 && img[15] == 0x7bff);
    imageStore(im, ivec2(2, 0), vec4(9.0)); imageStore(im, ivec2(0, 2), vec4(9.0)); imageStore(im, ivec2(-1, 0), vec4(9.0));
    CHECK(img[0] == 0x1234 && img[4] == 0x1234 && img[8] == 0x1234);

    printf(failures ? "%d failures\n" : "glsl_compat ok\n", failures);
    return failures ? 1 : 0;
}
}  // namespace check

int main() { return check::run(); }

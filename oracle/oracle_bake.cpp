// oracle_bake.cpp — CPU restatement of the reference's load-time environment bakes. TEST INFRASTRUCTURE:
// the product computes these on the GPU (swraster-viewer_b200/csrc/swr_bake.cuh, include/swr.h swr_bake_*); only tests/
// call this file, as the checker. Built into liboracle*.so next to oracle.cpp.
//   hammersley / GGX importance sampling  src/texture.rs:135-165
//   integrate_brdf, generate_brdf_lut     src/texture.rs:167-235   (texels only; the mip chain is texture.rs:45-128)
//   cubemap direction <-> face uv         src/texture.rs:237-272
//   sample_cubemap_direction_linear       src/texture.rs:274-287   (bilinear, texture.rs:730-790, then sRGB -> linear)
//   compute_irradiance_sh4                src/texture.rs:289-328
//   generate_prefiltered_specular_cubemap src/texture.rs:330-420   (every mip at full face resolution, one roughness per mip)
// f32 in the reference's order (glam Vec3A: dot = (x*x' + y*y') + z*z', normalize = v * (1 / length)), -ffp-contract=off;
// sinf / cosf / powf / sqrtf are the C library's, which is what Rust's f32 methods call on Linux; f32::powi(5) is
// compiler-rt's square-and-multiply x * ((x*x) * (x*x)). Parity unpinned against the Rust binary (no toolchain here); pinned
// by the float64 restatements and analytic cases in tests/test_bakes.py.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct V3 {
    float x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3 normalize(V3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }  // glam: self * length_recip()
inline float fmax_rs(float a, float b) { return std::fmax(a, b); }       // f32::max: the non-NaN operand
inline float clamp_rs(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }  // f32::clamp (NaN stays NaN)
inline uint32_t as_u32(float v) {  // Rust `as u32`: NaN -> 0, saturating
    if (!(v == v) || v <= 0.0f) return 0u;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}
inline uint32_t pack(float r, float g, float b, float a) {  // util.rs:91-96
    return (as_u32(r * 255.0f) << 24) | (as_u32(g * 255.0f) << 16) | (as_u32(b * 255.0f) << 8) | as_u32(a * 255.0f);
}
inline float srgb_to_linear_scalar(float s) { return s <= 0.04045f ? s / 12.92f : std::pow((s + 0.055f) / 1.055f, 2.4f); }  // util.rs:50-56

inline float radical_inverse_vdc(uint32_t bits) {  // texture.rs:135-142
    bits = (bits << 16) | (bits >> 16);
    bits = ((bits & 0x55555555u) << 1) | ((bits & 0xAAAAAAAAu) >> 1);
    bits = ((bits & 0x33333333u) << 2) | ((bits & 0xCCCCCCCCu) >> 2);
    bits = ((bits & 0x0F0F0F0Fu) << 4) | ((bits & 0xF0F0F0F0u) >> 4);
    bits = ((bits & 0x00FF00FFu) << 8) | ((bits & 0xFF00FF00u) >> 8);
    return (float)bits * 2.3283064e-10f;
}

V3 importance_sample_ggx(float xi_x, float xi_y, V3 n, float roughness) {  // texture.rs:148-165
    const float a = roughness * roughness;
    const float phi = 2.0f * 3.14159274f * xi_x;
    const float cos_theta = std::sqrt((1.0f - xi_y) / (1.0f + (a * a - 1.0f) * xi_y));
    const float sin_theta = std::sqrt(fmax_rs(1.0f - cos_theta * cos_theta, 0.0f));
    const V3 h{std::cos(phi) * sin_theta, std::sin(phi) * sin_theta, cos_theta};
    const V3 up = std::fabs(n.z) < 0.999f ? V3{0.0f, 0.0f, 1.0f} : V3{1.0f, 0.0f, 0.0f};
    const V3 tangent = normalize(cross(n, up));
    const V3 bitangent = cross(n, tangent);
    return normalize((tangent * h.x + bitangent * h.y) + n * h.z);
}

void integrate_brdf(float ndotv, float roughness, float &out_a, float &out_b) {  // texture.rs:167-197
    const V3 v{std::sqrt(fmax_rs(1.0f - ndotv * ndotv, 0.0f)), 0.0f, ndotv};
    const V3 n{0.0f, 0.0f, 1.0f};
    const uint32_t sample_count = 128;
    float a = 0.0f, b = 0.0f;
    for (uint32_t i = 0; i < sample_count; i++) {
        const float xx = (float)i / (float)sample_count, xy = radical_inverse_vdc(i);
        const V3 h = importance_sample_ggx(xx, xy, n, roughness);
        const V3 l = normalize(h * (2.0f * dot(v, h)) - v);
        const float ndotl = fmax_rs(l.z, 0.0f), ndoth = fmax_rs(h.z, 0.0f), vdoth = fmax_rs(dot(v, h), 0.0f);
        if (ndotl > 0.0f) {
            const float alpha = roughness * roughness;
            const float k = (alpha + 1.0f) * (alpha + 1.0f) * 0.125f;
            const float g_v = ndotv / (ndotv * (1.0f - k) + k);
            const float g_l = ndotl / (ndotl * (1.0f - k) + k);
            const float g_vis = fmax_rs(g_v * g_l * vdoth / (ndoth * fmax_rs(ndotv, 1.0e-5f)), 0.0f);
            const float om = 1.0f - vdoth, om2 = om * om;
            const float fc = om * (om2 * om2);  // powi(5)
            a += (1.0f - fc) * g_vis;
            b += fc * g_vis;
        }
    }
    out_a = a / (float)sample_count;
    out_b = b / (float)sample_count;
}

V3 face_uv_to_direction(uint32_t face, float u, float v) {  // texture.rs:237-247
    V3 d;
    switch (face) {
        case 0: d = V3{1.0f, -v, -u}; break;
        case 1: d = V3{-1.0f, -v, u}; break;
        case 2: d = V3{u, 1.0f, v}; break;
        case 3: d = V3{u, -1.0f, -v}; break;
        case 4: d = V3{u, -v, 1.0f}; break;
        default: d = V3{-u, -v, -1.0f}; break;
    }
    return normalize(d);
}
void direction_to_face_uv(V3 n, uint32_t &face, float &u, float &v) {  // texture.rs:249-272
    const float ax = std::fabs(n.x), ay = std::fabs(n.y), az = std::fabs(n.z);
    if (ax >= ay && ax >= az) {
        if (n.x >= 0.0f)
            face = 0, u = (-n.z / ax) * 0.5f + 0.5f, v = (-n.y / ax) * 0.5f + 0.5f;
        else
            face = 1, u = (n.z / ax) * 0.5f + 0.5f, v = (-n.y / ax) * 0.5f + 0.5f;
    } else if (ay > ax && ay >= az) {
        if (n.y >= 0.0f)
            face = 2, u = (n.x / ay) * 0.5f + 0.5f, v = (n.z / ay) * 0.5f + 0.5f;
        else
            face = 3, u = (n.x / ay) * 0.5f + 0.5f, v = (-n.z / ay) * 0.5f + 0.5f;
    } else if (n.z >= 0.0f) {
        face = 4, u = (n.x / az) * 0.5f + 0.5f, v = (-n.y / az) * 0.5f + 0.5f;
    } else {
        face = 5, u = (-n.x / az) * 0.5f + 0.5f, v = (-n.y / az) * 0.5f + 0.5f;
    }
}

struct Cube {  // mip 0 of the sky cubemap: six faces of w x h RGBA8 texels (R in bits 31..24), face-major
    const uint32_t *texels;
    uint32_t w, h;
};

// sample_bilinear_rgb at mip 0 of one face with ClampToEdge (texture.rs:730-790, :578-589), then sRGB -> linear (:274-287)
V3 sample_direction_linear(const Cube &c, V3 dir) {
    uint32_t face;
    float u, v;
    direction_to_face_uv(dir, face, u, v);
    u = clamp_rs(u, 0.0f, 1.0f), v = clamp_rs(v, 0.0f, 1.0f);
    const float wf = (float)c.w, hf = (float)c.h;
    const uint32_t off = face * c.w * c.h;
    const float xf = u * wf - 0.5f, yf = v * hf - 0.5f;
    const float x0 = std::floor(xf), y0 = std::floor(yf), x1 = x0 + 1.0f, y1 = y0 + 1.0f;
    const float fx = xf - x0, fy = yf - y0, ofx = 1.0f - fx, ofy = 1.0f - fy;
    auto clampi = [](float texel, float dim) { return as_u32(std::fmin(texel, dim - 1.0f)); };  // _mm_min_ps + as_uvec4
    const uint32_t x0i = clampi(x0, wf), y0i = clampi(y0, hf), x1i = clampi(x1, wf), y1i = clampi(y1, hf);
    const uint32_t t[4] = {c.texels[off + y0i * c.w + x0i], c.texels[off + y0i * c.w + x1i], c.texels[off + y1i * c.w + x0i], c.texels[off + y1i * c.w + x1i]};
    const float wgt[4] = {ofx * ofy, fx * ofy, ofx * fy, fx * fy};
    float rgb[3];
    for (int ch = 0; ch < 3; ch++) {
        const int sh = 24 - 8 * ch;
        const float p00 = (float)((t[0] >> sh) & 0xFF) / 255.0f, p10 = (float)((t[1] >> sh) & 0xFF) / 255.0f;
        const float p01 = (float)((t[2] >> sh) & 0xFF) / 255.0f, p11 = (float)((t[3] >> sh) & 0xFF) / 255.0f;
        rgb[ch] = ((p00 * wgt[0] + p10 * wgt[1]) + p01 * wgt[2]) + p11 * wgt[3];
    }
    return V3{srgb_to_linear_scalar(rgb[0]), srgb_to_linear_scalar(rgb[1]), srgb_to_linear_scalar(rgb[2])};
}

}  // namespace

extern "C" {

// one texel's worth of integrate_brdf (texture.rs:167-197): out = (scale, bias)
int orc_integrate_brdf(float ndotv, float roughness, float *out) {
    integrate_brdf(ndotv, roughness, out[0], out[1]);
    return 0;
}

// generate_brdf_lut(size) texels (texture.rs:199-235), row-major size x size
int orc_bake_brdf_lut(uint32_t size, uint32_t *out) {
    const float size_f = (float)size;
    for (uint32_t y = 0; y < size; y++) {
        const float roughness = clamp_rs(((float)y + 0.5f) / size_f, 0.0f, 1.0f);
        for (uint32_t x = 0; x < size; x++) {
            const float ndotv = clamp_rs(((float)x + 0.5f) / size_f, 0.0f, 1.0f);
            float a, b;
            integrate_brdf(fmax_rs(ndotv, 1.0e-4f), fmax_rs(roughness, 1.0e-4f), a, b);
            out[(size_t)y * size + x] = pack(clamp_rs(a, 0.0f, 1.0f), clamp_rs(b, 0.0f, 1.0f), 0.0f, 1.0f);
        }
    }
    return 0;
}

// compute_irradiance_sh4 (texture.rs:289-328): out = 4 coefficients x rgb
int orc_bake_irradiance_sh4(const uint32_t *faces, uint32_t w, uint32_t h, float *out) {
    const Cube cube{faces, w, h};
    V3 sh[4] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const float width = (float)w, height = (float)h;
    const float texel_omega = (2.0f / width) * (2.0f / height);
    for (uint32_t face = 0; face < 6; face++)
        for (uint32_t y = 0; y < h; y++) {
            const float v = (((float)y + 0.5f) / height) * 2.0f - 1.0f;
            for (uint32_t x = 0; x < w; x++) {
                const float u = (((float)x + 0.5f) / width) * 2.0f - 1.0f;
                const V3 dir = face_uv_to_direction(face, u, v);
                const float weight = texel_omega / std::pow((1.0f + u * u) + v * v, 1.5f);
                const V3 color = sample_direction_linear(cube, dir);
                const float basis[4] = {0.282095f, 0.488603f * dir.y, 0.488603f * dir.z, 0.488603f * dir.x};
                for (int i = 0; i < 4; i++) sh[i] = sh[i] + color * (basis[i] * weight);
            }
        }
    sh[0] = sh[0] * 3.14159274f;
    for (int i = 1; i < 4; i++) sh[i] = sh[i] * (2.0f * 3.14159274f / 3.0f);
    sh[0] = sh[0] * 0.282095f;
    for (int i = 1; i < 4; i++) sh[i] = sh[i] * 0.488603f;
    for (int i = 0; i < 4; i++) out[3 * i] = sh[i].x, out[3 * i + 1] = sh[i].y, out[3 * i + 2] = sh[i].z;
    return 0;
}

// generate_prefiltered_specular_cubemap (texture.rs:330-420): out = num_mips x 6 x h x w texels, num_mips = floor(log2(max(w, h))) + 1
int orc_bake_prefilter_specular(const uint32_t *faces, uint32_t w, uint32_t h, uint32_t sample_count, uint32_t *out) {
    const Cube cube{faces, w, h};
    uint32_t num_mips = 1;
    for (uint32_t m = std::max(w, h); m >>= 1;) num_mips++;
    const uint32_t max_mip = num_mips - 1;
    for (uint32_t mip = 0; mip < num_mips; mip++) {
        const float roughness = max_mip > 0 ? (float)mip / (float)max_mip : 0.0f;
        const size_t mip_offset = (size_t)mip * w * h * 6;
#pragma omp parallel for schedule(dynamic, 4)
        for (int64_t row = 0; row < (int64_t)6 * h; row++) {  // texels are independent: the thread count does not change them
            const uint32_t face = (uint32_t)(row / h), y = (uint32_t)(row % h);
            const float v = (((float)y + 0.5f) / (float)h) * 2.0f - 1.0f;
            for (uint32_t x = 0; x < w; x++) {
                const float u = (((float)x + 0.5f) / (float)w) * 2.0f - 1.0f;
                const V3 r = face_uv_to_direction(face, u, v);
                V3 color;
                if (mip == 0) {
                    color = sample_direction_linear(cube, r);
                } else {
                    V3 accum{0, 0, 0};
                    float total = 0.0f;
                    for (uint32_t i = 0; i < sample_count; i++) {
                        const float xx = (float)i / (float)sample_count, xy = radical_inverse_vdc(i);
                        const V3 hv = importance_sample_ggx(xx, xy, r, fmax_rs(roughness, 0.045f));
                        const V3 l = normalize(hv * (2.0f * dot(r, hv)) - r);
                        const float ndotl = fmax_rs(dot(r, l), 0.0f);
                        if (ndotl > 0.0f) {
                            accum = accum + sample_direction_linear(cube, l) * ndotl;
                            total += ndotl;
                        }
                    }
                    color = total > 0.0f ? accum / total : sample_direction_linear(cube, r);
                }
                out[mip_offset + (size_t)face * w * h + (size_t)y * w + x] = pack(color.x, color.y, color.z, 1.0f);
            }
        }
    }
    return (int)num_mips;
}

}  // extern "C"

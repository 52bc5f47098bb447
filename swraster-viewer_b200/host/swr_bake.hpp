// swr_bake.hpp — load-time environment bakes on the host (host side). The reference computes these on the
// CPU when a scene is loaded (Scene::from_gltf, scene.rs:151-231); so does this mirror. They are inputs of the shading
// kernel, not part of the per-frame path. GPU versions of the two heavy ones (GGX prefilter, SH projection) are future
// work; nothing here is a fallback for a device code path.
//   cubemap cross -> 6 faces              src/texture.rs:922-959   (layout: +Y on top, -X +Z +X -Z in the middle row, -Y below)
//   hammersley / GGX importance sampling  src/texture.rs:135-165
//   integrate_brdf, generate_brdf_lut     src/texture.rs:167-235
//   cubemap direction <-> face uv         src/texture.rs:237-272
//   sample_cubemap_direction_linear       src/texture.rs:274-287   (bilinear, texture.rs:730-790, then sRGB -> linear)
//   compute_irradiance_sh4                src/texture.rs:289-328
//   generate_prefiltered_specular_cubemap src/texture.rs:330-420   (every mip at full face resolution, one roughness per mip)
//   initialize_voxel_gi_from_scene        src/gi.rs:123-149        (main.rs:280: scale 0.25, sky visibility 1.0)
// Arithmetic is f32 in the reference's order (glam Vec3A: dot = (x*x' + y*y') + z*z', normalize = v * (1 / length), unfused;
// translation unit built with -ffp-contract=off). sinf/cosf/powf/sqrtf are the C library's, which is what Rust's f32 methods
// call on Linux. f32::powi(5) is compiler-rt's square-and-multiply: x * ((x*x) * (x*x)).
#pragma once
#include "swr_gltf.hpp"

namespace swr {
namespace bake {

using gltf::V3;
using gltf::TextureData;

inline float radical_inverse_vdc(uint32_t bits) {  // texture.rs:135-142
    bits = (bits << 16) | (bits >> 16);
    bits = ((bits & 0x55555555u) << 1) | ((bits & 0xAAAAAAAAu) >> 1);
    bits = ((bits & 0x33333333u) << 2) | ((bits & 0xCCCCCCCCu) >> 2);
    bits = ((bits & 0x0F0F0F0Fu) << 4) | ((bits & 0xF0F0F0F0u) >> 4);
    bits = ((bits & 0x00FF00FFu) << 8) | ((bits & 0xFF00FF00u) >> 8);
    return (float)bits * 2.3283064e-10f;
}
inline void hammersley(uint32_t i, uint32_t n, float &x, float &y) {  // texture.rs:144-146
    x = (float)i / (float)n;
    y = radical_inverse_vdc(i);
}
inline float fmax_rs(float a, float b) { return std::fmax(a, b); }  // f32::max: the non-NaN operand

inline V3 importance_sample_ggx(float xi_x, float xi_y, V3 n, float roughness) {  // texture.rs:148-165
    const float a = roughness * roughness;
    const float phi = 2.0f * 3.14159274f * xi_x;
    const float cos_theta = std::sqrt((1.0f - xi_y) / (1.0f + (a * a - 1.0f) * xi_y));
    const float sin_theta = std::sqrt(fmax_rs(1.0f - cos_theta * cos_theta, 0.0f));
    const V3 h{std::cos(phi) * sin_theta, std::sin(phi) * sin_theta, cos_theta};
    const V3 up = std::fabs(n.z) < 0.999f ? V3{0.0f, 0.0f, 1.0f} : V3{1.0f, 0.0f, 0.0f};
    const V3 tangent = gltf::normalize(gltf::cross(n, up));
    const V3 bitangent = gltf::cross(n, tangent);
    return gltf::normalize((tangent * h.x + bitangent * h.y) + n * h.z);
}

inline void integrate_brdf(float ndotv, float roughness, float &out_a, float &out_b) {  // texture.rs:167-197
    const V3 v{std::sqrt(fmax_rs(1.0f - ndotv * ndotv, 0.0f)), 0.0f, ndotv};
    const V3 n{0.0f, 0.0f, 1.0f};
    const uint32_t sample_count = 128;
    float a = 0.0f, b = 0.0f;
    for (uint32_t i = 0; i < sample_count; i++) {
        float xx, xy;
        hammersley(i, sample_count, xx, xy);
        const V3 h = importance_sample_ggx(xx, xy, n, roughness);
        const V3 l = gltf::normalize(h * (2.0f * gltf::dot(v, h)) - v);
        const float ndotl = fmax_rs(l.z, 0.0f), ndoth = fmax_rs(h.z, 0.0f), vdoth = fmax_rs(gltf::dot(v, h), 0.0f);
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

inline float clamp_rs(float v, float lo, float hi) {  // f32::clamp (NaN stays NaN)
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}

// generate_brdf_lut(size) (texture.rs:199-235): Linear texture, mip chain generated
inline TextureData generate_brdf_lut(uint32_t size) {
    TextureData t;
    t.width = t.height = size;
    t.type = SWR_TEX_LINEAR;
    t.data.reserve((size_t)size * size);
    const float size_f = (float)size;
    for (uint32_t y = 0; y < size; y++) {
        const float roughness = clamp_rs(((float)y + 0.5f) / size_f, 0.0f, 1.0f);
        for (uint32_t x = 0; x < size; x++) {
            const float ndotv = clamp_rs(((float)x + 0.5f) / size_f, 0.0f, 1.0f);
            float a, b;
            integrate_brdf(fmax_rs(ndotv, 1.0e-4f), fmax_rs(roughness, 1.0e-4f), a, b);
            const float texel[4] = {clamp_rs(a, 0.0f, 1.0f), clamp_rs(b, 0.0f, 1.0f), 0.0f, 1.0f};
            t.data.push_back(gltf::rgba8_pack_vec4(texel));
        }
    }
    t.mip_offsets = {0}, t.mip_widths = {size}, t.mip_heights = {size}, t.array_stride = {0};
    t.generate_mipmaps();
    return t;
}

// TextureCache::load_texture, cubemap branch (texture.rs:922-959) + generate_mipmaps
inline TextureData cubemap_from_cross(const gltf::Image &img) {
    if (img.width < 4 || img.height < 3) throw std::runtime_error("Invalid data: cubemap cross image is smaller than 4x3");
    TextureData t;
    const uint32_t fw = img.width / 4, fh = img.height / 3;
    t.width = fw, t.height = fh, t.type = SWR_TEX_CUBEMAP;
    static const uint32_t face_pos[6][2] = {{2, 1}, {0, 1}, {1, 0}, {1, 2}, {1, 1}, {3, 1}};  // +X -X +Y -Y +Z -Z
    t.data.reserve((size_t)fw * fh * 6);
    for (int f = 0; f < 6; f++)
        for (uint32_t y = 0; y < fh; y++)
            for (uint32_t x = 0; x < fw; x++) {
                const size_t s = ((size_t)(face_pos[f][1] * fh + y) * img.width + face_pos[f][0] * fw + x) * 4;
                t.data.push_back(((uint32_t)img.rgba[s] << 24) | ((uint32_t)img.rgba[s + 1] << 16) | ((uint32_t)img.rgba[s + 2] << 8) | img.rgba[s + 3]);
            }
    t.mip_offsets = {0}, t.mip_widths = {fw}, t.mip_heights = {fh}, t.array_stride = {fw * fh};
    t.generate_mipmaps();
    return t;
}

inline V3 cubemap_face_uv_to_direction(uint32_t face, float u, float v) {  // texture.rs:237-247
    V3 d;
    switch (face) {
        case 0: d = V3{1.0f, -v, -u}; break;
        case 1: d = V3{-1.0f, -v, u}; break;
        case 2: d = V3{u, 1.0f, v}; break;
        case 3: d = V3{u, -1.0f, -v}; break;
        case 4: d = V3{u, -v, 1.0f}; break;
        default: d = V3{-u, -v, -1.0f}; break;
    }
    return gltf::normalize(d);
}
inline void cubemap_direction_to_face_uv(V3 n, uint32_t &face, float &u, float &v) {  // texture.rs:249-272
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

// sample_bilinear_rgb at mip 0 of one face with ClampToEdge (texture.rs:730-790, :578-589); one lane of the reference's four
inline void sample_bilinear_rgb_face(const TextureData &t, float u, float v, uint32_t face, float rgb[3]) {
    const float wf = (float)t.mip_widths[0], hf = (float)t.mip_heights[0];
    const uint32_t wi = t.mip_widths[0], off = t.mip_offsets[0] + face * t.array_stride[0];
    const float xf = u * wf - 0.5f, yf = v * hf - 0.5f;
    const float x0 = std::floor(xf), y0 = std::floor(yf), x1 = x0 + 1.0f, y1 = y0 + 1.0f;
    const float fx = xf - x0, fy = yf - y0, ofx = 1.0f - fx, ofy = 1.0f - fy;
    auto clampi = [](float texel, float dim) { return gltf::f32_as_u32_saturating(std::fmin(texel, dim - 1.0f)); };  // _mm_min_ps + as_uvec4
    const uint32_t x0i = clampi(x0, wf), y0i = clampi(y0, hf), x1i = clampi(x1, wf), y1i = clampi(y1, hf);
    float p00[4], p10[4], p01[4], p11[4];
    gltf::rgba8_unpack_vec4(t.data[off + y0i * wi + x0i], p00);
    gltf::rgba8_unpack_vec4(t.data[off + y0i * wi + x1i], p10);
    gltf::rgba8_unpack_vec4(t.data[off + y1i * wi + x0i], p01);
    gltf::rgba8_unpack_vec4(t.data[off + y1i * wi + x1i], p11);
    const float w00 = ofx * ofy, w10 = fx * ofy, w01 = ofx * fy, w11 = fx * fy;
    for (int c = 0; c < 3; c++) rgb[c] = ((p00[c] * w00 + p10[c] * w10) + p01[c] * w01) + p11[c] * w11;
}

inline V3 sample_cubemap_direction_linear(const TextureData &cubemap, V3 dir) {  // texture.rs:274-287
    uint32_t face;
    float u, v, rgb[3];
    cubemap_direction_to_face_uv(dir, face, u, v);
    sample_bilinear_rgb_face(cubemap, clamp_rs(u, 0.0f, 1.0f), clamp_rs(v, 0.0f, 1.0f), face, rgb);
    return V3{gltf::srgb_to_linear_scalar(rgb[0]), gltf::srgb_to_linear_scalar(rgb[1]), gltf::srgb_to_linear_scalar(rgb[2])};
}

// compute_irradiance_sh4 (texture.rs:289-328): out = 4 coefficients x rgb
inline void compute_irradiance_sh4(const TextureData &cubemap, float out[12]) {
    V3 sh[4] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const float width = (float)cubemap.width, height = (float)cubemap.height;
    const float texel_omega = (2.0f / width) * (2.0f / height);
    for (uint32_t face = 0; face < 6; face++)
        for (uint32_t y = 0; y < cubemap.height; y++) {
            const float v = (((float)y + 0.5f) / height) * 2.0f - 1.0f;
            for (uint32_t x = 0; x < cubemap.width; x++) {
                const float u = (((float)x + 0.5f) / width) * 2.0f - 1.0f;
                const V3 dir = cubemap_face_uv_to_direction(face, u, v);
                const float weight = texel_omega / std::pow((1.0f + u * u) + v * v, 1.5f);
                const V3 color = sample_cubemap_direction_linear(cubemap, dir);
                const float basis[4] = {0.282095f, 0.488603f * dir.y, 0.488603f * dir.z, 0.488603f * dir.x};
                for (int i = 0; i < 4; i++) sh[i] = sh[i] + color * (basis[i] * weight);
            }
        }
    sh[0] = sh[0] * 3.14159274f;
    for (int i = 1; i < 4; i++) sh[i] = sh[i] * (2.0f * 3.14159274f / 3.0f);
    sh[0] = sh[0] * 0.282095f;
    for (int i = 1; i < 4; i++) sh[i] = sh[i] * 0.488603f;
    for (int i = 0; i < 4; i++) out[3 * i] = sh[i].x, out[3 * i + 1] = sh[i].y, out[3 * i + 2] = sh[i].z;
}

// generate_prefiltered_specular_cubemap (texture.rs:330-420)
inline TextureData generate_prefiltered_specular_cubemap(const TextureData &cubemap, uint32_t sample_count) {
    TextureData t;
    const uint32_t bw = cubemap.width, bh = cubemap.height;
    uint32_t num_mips = 1;
    for (uint32_t m = std::max(bw, bh); m >>= 1;) num_mips++;
    const uint32_t max_mip = num_mips - 1;
    t.width = bw, t.height = bh, t.type = SWR_TEX_LINEAR;
    t.data.assign((size_t)bw * bh * 6 * num_mips, 0u);
    for (uint32_t mip = 0; mip < num_mips; mip++) {
        const float roughness = max_mip > 0 ? (float)mip / (float)max_mip : 0.0f;
        const uint32_t mip_offset = mip * bw * bh * 6;
        t.mip_offsets.push_back(mip_offset);
        t.mip_widths.push_back(bw);
        t.mip_heights.push_back(bh);
        t.array_stride.push_back(bw * bh);
        // texels are independent: rows are spread over the host threads (the reference's loop is serial and cached on disk);
        // every texel is computed exactly as in the serial order, so the result does not depend on the thread count
#pragma omp parallel for schedule(dynamic, 4)
        for (int64_t row = 0; row < (int64_t)6 * bh; row++) {
            const uint32_t face = (uint32_t)(row / bh), y = (uint32_t)(row % bh);
            const float v = (((float)y + 0.5f) / (float)bh) * 2.0f - 1.0f;
            for (uint32_t x = 0; x < bw; x++) {
                const float u = (((float)x + 0.5f) / (float)bw) * 2.0f - 1.0f;
                const V3 r = cubemap_face_uv_to_direction(face, u, v);
                V3 color;
                if (mip == 0) {
                    color = sample_cubemap_direction_linear(cubemap, r);
                } else {
                    V3 accum{0, 0, 0};
                    float total = 0.0f;
                    for (uint32_t i = 0; i < sample_count; i++) {
                        float xx, xy;
                        hammersley(i, sample_count, xx, xy);
                        const V3 h = importance_sample_ggx(xx, xy, r, fmax_rs(roughness, 0.045f));
                        const V3 l = gltf::normalize(h * (2.0f * gltf::dot(r, h)) - r);
                        const float ndotl = fmax_rs(gltf::dot(r, l), 0.0f);
                        if (ndotl > 0.0f) {
                            accum = accum + sample_cubemap_direction_linear(cubemap, l) * ndotl;
                            total += ndotl;
                        }
                    }
                    color = total > 0.0f ? accum / total : sample_cubemap_direction_linear(cubemap, r);
                }
                const float c4[4] = {color.x, color.y, color.z, 1.0f};
                t.data[(size_t)mip_offset + (size_t)face * bw * bh + (size_t)y * bw + x] = gltf::rgba8_pack_vec4(c4);
            }
        }
    }
    return t;
}

// initialize_voxel_gi_from_scene (gi.rs:123-149): every voxel gets the scene SH x scale; coefficient 0's w = the voxel's
// sun visibility (light_intensity, computed by ray casting in the reference — N4; a constant here), coefficient 1's w = sky
// visibility.
inline std::vector<float> initialize_voxel_gi(const float sh[12], uint32_t w, uint32_t h, uint32_t d, float irradiance_scale, float sky_visibility,
                                              float light_intensity) {
    std::vector<float> out((size_t)w * h * d * 16, 0.0f);
    for (size_t vtx = 0; vtx < (size_t)w * h * d; vtx++) {
        float *o = &out[vtx * 16];
        for (int i = 0; i < 4; i++) {
            o[4 * i] = sh[3 * i] * irradiance_scale;
            o[4 * i + 1] = sh[3 * i + 1] * irradiance_scale;
            o[4 * i + 2] = sh[3 * i + 2] * irradiance_scale;
            o[4 * i + 3] = 0.0f;
        }
        o[3] = light_intensity;
        o[7] = sky_visibility;
    }
    return out;
}

// Everything Scene::from_gltf derives from the sky image (scene.rs:151-231) plus the default voxel grid of main.rs:228-281.
struct EnvironmentBake {
    TextureData cubemap, cubemap_specular, brdf_lut;
    float irradiance_sh[12];
    std::vector<float> voxels;
    uint32_t voxel_dims[3];
    // flat views
    swr_texture_desc descs[3];

    static std::unique_ptr<EnvironmentBake> from_cross(const gltf::Image &cross, uint32_t lut_size, uint32_t specular_samples, uint32_t voxel_dim,
                                                       float irradiance_scale, float sky_visibility, float light_intensity) {
        std::unique_ptr<EnvironmentBake> e(new EnvironmentBake());
        e->cubemap = cubemap_from_cross(cross);
        e->cubemap_specular = generate_prefiltered_specular_cubemap(e->cubemap, specular_samples);
        compute_irradiance_sh4(e->cubemap, e->irradiance_sh);
        e->brdf_lut = generate_brdf_lut(lut_size);
        e->voxel_dims[0] = e->voxel_dims[1] = e->voxel_dims[2] = voxel_dim;
        e->voxels = initialize_voxel_gi(e->irradiance_sh, voxel_dim, voxel_dim, voxel_dim, irradiance_scale, sky_visibility, light_intensity);
        const TextureData *ts[3] = {&e->cubemap, &e->cubemap_specular, &e->brdf_lut};
        for (int k = 0; k < 3; k++) {
            const TextureData &t = *ts[k];
            swr_texture_desc &d = e->descs[k];
            std::memset(&d, 0, sizeof(d));
            d.data = t.data.data(), d.ntexels = (uint32_t)t.data.size(), d.width = t.width, d.height = t.height, d.texture_type = t.type;
            d.max_mip_level = t.max_mip_level();
            d.mip_offsets = t.mip_offsets.data(), d.mip_widths = t.mip_widths.data(), d.mip_heights = t.mip_heights.data(), d.array_stride = t.array_stride.data();
            d.wrap_s = d.wrap_t = SWR_WRAP_CLAMP_TO_EDGE;  // scene.rs:153-158, :206-211, :222-227
        }
        return e;
    }
};

}  // namespace bake
}  // namespace swr

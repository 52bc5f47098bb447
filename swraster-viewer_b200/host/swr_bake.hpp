// swr_bake.hpp — load-time environment bakes. The reference computes these on the CPU when a scene is loaded
// (Scene::from_gltf, scene.rs:151-231). Here the three integrals — BRDF LUT, irradiance SH4, GGX-prefiltered cubemap — run on
// the GPU (include/swr.h swr_bake_*, csrc/swr_bake.cuh); what stays on the host is glue: cutting the cross image into faces,
// the mip chains, filling the voxel grid from the SH. There is no CPU fallback for the integrals (the CPU restatement is test infrastructure outside this package: the tests' checker).
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
#include "swr_cache.hpp"

namespace swr {
namespace bake {

using gltf::V3;
using gltf::TextureData;

// generate_brdf_lut(size) (texture.rs:199-235): texels integrated on the device (swr_bake_brdf_lut), Linear texture, mip chain here
inline TextureData generate_brdf_lut(uint32_t size) {
    TextureData t;
    t.width = t.height = size;
    t.type = SWR_TEX_LINEAR;
    t.data.assign((size_t)size * size, 0u);
    if (swr_bake_brdf_lut(-1, size, t.data.data()) != SWR_OK) throw std::runtime_error(std::string("swr_bake_brdf_lut: ") + swr_bake_last_error());
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

// compute_irradiance_sh4 (texture.rs:289-328) on the device: out = 4 coefficients x rgb
inline void compute_irradiance_sh4(const TextureData &cubemap, float out[12]) {
    if (swr_bake_irradiance_sh4(-1, cubemap.data.data() + cubemap.mip_offsets[0], cubemap.width, cubemap.height, out) != SWR_OK)
        throw std::runtime_error(std::string("swr_bake_irradiance_sh4: ") + swr_bake_last_error());
}

// generate_prefiltered_specular_cubemap (texture.rs:330-420) on the device: every mip at full face resolution
inline TextureData generate_prefiltered_specular_cubemap(const TextureData &cubemap, uint32_t sample_count) {
    TextureData t;
    const uint32_t bw = cubemap.width, bh = cubemap.height;
    uint32_t num_mips = 1;
    for (uint32_t m = std::max(bw, bh); m >>= 1;) num_mips++;
    t.width = bw, t.height = bh, t.type = SWR_TEX_LINEAR;
    t.data.assign((size_t)bw * bh * 6 * num_mips, 0u);
    for (uint32_t mip = 0; mip < num_mips; mip++) {
        t.mip_offsets.push_back(mip * bw * bh * 6);
        t.mip_widths.push_back(bw);
        t.mip_heights.push_back(bh);
        t.array_stride.push_back(bw * bh);
    }
    // the six faces of mip 0 are the first 6 * w * h texels of the cubemap (face stride = w * h at mip 0)
    const int got = swr_bake_prefilter_specular(-1, cubemap.data.data() + cubemap.mip_offsets[0], bw, bh, sample_count, t.data.data());
    if (got != (int)num_mips) throw std::runtime_error(std::string("swr_bake_prefilter_specular: ") + swr_bake_last_error());
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
    bool specular_from_cache = false;

    static std::unique_ptr<EnvironmentBake> from_cross(const gltf::Image &cross, uint32_t lut_size, uint32_t specular_samples, uint32_t voxel_dim,
                                                       float irradiance_scale, float sky_visibility, float light_intensity,
                                                       const char *ggx_cache_path = nullptr) {
        std::unique_ptr<EnvironmentBake> e(new EnvironmentBake());
        e->cubemap = cubemap_from_cross(cross);
        // scene.rs:164-206: the prefiltered cubemap comes from `assets/cubemap.ggx` when that file matches the sky's face size,
        // otherwise it is baked (here: on the device) and the cache is written; a cache that cannot be written is not an error
        std::vector<uint32_t> cached;
        if (ggx_cache_path && cache::ggx_load(ggx_cache_path, e->cubemap.width, e->cubemap.height, cached)) {
            TextureData &t = e->cubemap_specular;
            const uint32_t bw = e->cubemap.width, bh = e->cubemap.height, mips = cache::ggx_mips(bw, bh);
            t.width = bw, t.height = bh, t.type = SWR_TEX_LINEAR;
            t.data = std::move(cached);
            for (uint32_t mip = 0; mip < mips; mip++) {
                t.mip_offsets.push_back(mip * bw * bh * 6);
                t.mip_widths.push_back(bw);
                t.mip_heights.push_back(bh);
                t.array_stride.push_back(bw * bh);
            }
            e->specular_from_cache = true;
        } else {
            e->cubemap_specular = generate_prefiltered_specular_cubemap(e->cubemap, specular_samples);
            if (ggx_cache_path) {
                try {
                    const TextureData &t = e->cubemap_specular;
                    cache::ggx_save(ggx_cache_path, t.width, t.height, (uint32_t)t.mip_offsets.size(), t.data.data(), t.data.size());
                } catch (const std::exception &) {
                }
            }
        }
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

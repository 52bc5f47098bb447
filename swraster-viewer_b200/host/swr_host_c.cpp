// C exports of the host mirror (swr_host.hpp) so Python tests/bench can drive the
// same Renderer / RenderCamera code a C++ application would. Links libswr_b200.so.
#include "swr_host.hpp"
#include "swr_gltf.hpp"
#include "swr_bake.hpp"
#include "swr_sunvis.hpp"
#include "../../include/swr_gltf.h"
#include "../../include/swr_host.h"

static thread_local std::string g_err;

#define SWRH_TRY(stmt)                \
    try {                             \
        stmt;                         \
        return 0;                     \
    } catch (const std::exception &e) { \
        g_err = e.what();             \
        return -1;                    \
    }

// a NULL handle / pointer from the caller is an error, not a crash
static swr::Renderer &need(void *r) {
    if (!r) throw std::runtime_error("null renderer handle");
    return *(swr::Renderer *)r;
}
template <typename T>
static T &need_ptr(T *p, const char *what) {
    if (!p) throw std::runtime_error(std::string("null ") + what);
    return *p;
}

extern "C" {

const char *swrh_last_error(void) { return g_err.c_str(); }

// RenderCamera::new + update_matrices (rendercamera.rs:28-86)
int swrh_camera_build(const float pos[3], const float look_at[3], float fov, float width, float height, float far_plane,
                      swr_camera *out) {
    SWRH_TRY(need_ptr(pos, "position"); need_ptr(look_at, "look_at"); need_ptr(out, "camera") = swr::RenderCamera(pos, look_at, fov, width, height, far_plane).to_abi());
}

// RenderCamera::new towards a level target, then rotate_mouse(dx, dy) and update_matrices (main.rs:478-500 path).
int swrh_camera_build_rotated(const float pos[3], const float look_at_level[3], float mouse_dx, float mouse_dy, float fov, float width,
                              float height, float far_plane, swr_camera *out) {
    SWRH_TRY(need_ptr(pos, "position"); need_ptr(look_at_level, "look_at"); swr::RenderCamera c(pos, look_at_level, fov, width, height, far_plane);
             c.rotate_mouse(mouse_dx, mouse_dy); c.update_matrices(); need_ptr(out, "camera") = c.to_abi());
}

void *swrh_renderer_new(int width, int height, int device) {
    try {
        return new swr::Renderer(width, height, device);
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void *swrh_renderer_new_lanes(int width, int height, int device, int lanes) {
    try {
        return new swr::Renderer(width, height, device, lanes);
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
int swrh_renderer_lanes(void *r) { return r ? ((swr::Renderer *)r)->lanes() : -1; }
swr_ctx *swrh_renderer_lane_ctx(void *r, int lane) { return r ? ((swr::Renderer *)r)->lane_ctx(lane) : nullptr; }
void *swrh_renderer_new_multi(int width, int height, const int *devices, int ndev) {
    try {
        if (!devices || ndev < 1) throw std::runtime_error("swrh_renderer_new_multi: device list is empty");
        return new swr::Renderer(width, height, std::vector<int>(devices, devices + ndev));
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
int swrh_renderer_tile_rows(void *r, int device_index, int *row_begin, int *row_end) {
    SWRH_TRY(swr::Renderer &R = need(r); if (!R.multi()) throw std::runtime_error("not a multi-device renderer");
             if (swr_multi_tile_rows(R.multi(), device_index, &need_ptr(row_begin, "row_begin"), &need_ptr(row_end, "row_end")) != 0) throw std::runtime_error("device index out of range"));
}
void swrh_renderer_free(void *r) { delete (swr::Renderer *)r; }
swr_ctx *swrh_renderer_ctx(void *r) { return r ? ((swr::Renderer *)r)->ctx() : nullptr; }

int swrh_set_reference_rsqrt(void *r, int on) { SWRH_TRY(need(r).set_reference_rsqrt(on != 0)); }
int swrh_reference_rsqrt_bits(void *r) { return r ? ((swr::Renderer *)r)->reference_rsqrt_bits() : -1; }
int swrh_set_tile_rows(void *r, int r0, int r1) { SWRH_TRY(need(r).set_tile_rows(r0, r1)); }

int swrh_render_scene(void *r, const swr_scene_desc *scene, const swr_camera *cam, int shade, int shard, int nshards) {
    SWRH_TRY(need(r).render_scene(swr::Scene(&need_ptr(scene, "scene")), need_ptr(cam, "camera"), shade != 0, shard, nshards));
}

int swrh_build_draws_band(const swr_scene_desc *scene, const swr_camera *cam, swr_draw *out, int max_draws, int y0, int y1, int height) {
    try {
        std::vector<swr_draw> draws;
        swr::validate_scene_ranges(need_ptr(scene, "scene"));
        swr::build_draw_list(*scene, need_ptr(cam, "camera"), draws, 0, 1, y0, y1, height);
        for (size_t i = 0; out && i < draws.size() && (int)i < max_draws; i++) out[i] = draws[i];
        return (int)draws.size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
int swrh_invalidate_scene(void *r) { SWRH_TRY(need(r).invalidate_scene()); }
int swrh_num_draws(void *r) { return r ? (int)((swr::Renderer *)r)->draws().size() : -1; }
int swrh_auto_exposure_step(float state[3], const float *tile_luminance, int ntiles, float delta_time) {
    SWRH_TRY(if (ntiles < 0) throw std::runtime_error("negative tile count");
             swr::auto_exposure_step(&need_ptr(state, "state"), ntiles ? &need_ptr(tile_luminance, "tile luminance") : nullptr, (size_t)ntiles, delta_time));
}
int swrh_update_auto_exposure(void *r, float dt) { SWRH_TRY(need(r).update_auto_exposure(dt)); }
int swrh_frame_exposure(void *r, float exposure) { SWRH_TRY(need(r).frame_exposure(exposure)); }
int swrh_frame_hdr(void *r) { SWRH_TRY(need(r).frame_hdr()); }
float swrh_auto_exposure(void *r) { return r ? ((swr::Renderer *)r)->auto_exposure() : 0.0f; }

int swrh_blit_to_buffer(void *r, uint32_t *pixels, size_t width, size_t height) {
    SWRH_TRY(swr::RenderBuffer buf(width, height, &need_ptr(pixels, "pixel buffer")); need(r).blit_to_buffer(buf));
}

int swrh_blit_to_buffer_async(void *r, uint32_t *pixels, size_t width, size_t height, int *ticket) {
    SWRH_TRY(swr::RenderBuffer buf(width, height, &need_ptr(pixels, "pixel buffer")); need_ptr(ticket, "ticket") = need(r).blit_to_buffer_async(buf));
}
int swrh_wait_blit(void *r, int ticket) { SWRH_TRY(need(r).wait_blit(ticket)); }

// Host draw list only (no device work): for cross-checks against the oracle's own R1/R2.
int swrh_build_draws(const swr_scene_desc *scene, const swr_camera *cam, swr_draw *out, int max_draws, int shard, int nshards) {
    try {
        std::vector<swr_draw> draws;
        swr::validate_scene_ranges(need_ptr(scene, "scene"));
        swr::build_draw_list(*scene, need_ptr(cam, "camera"), draws, shard, nshards);
        for (size_t i = 0; out && i < draws.size() && (int)i < max_draws; i++) out[i] = draws[i];
        return (int)draws.size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// ---- glTF loader (include/swr_gltf.h) ---------------------------------------------------------------------------------
void *swrh_gltf_load(const char *path, const swrh_gltf_env *env) {
    try {
        if (!path) throw std::runtime_error("Invalid data: null path");
        swr::gltf::Environment e;
        if (env) {
            e.cubemap = env->cubemap;
            e.cubemap_specular = env->cubemap_specular;
            e.brdf_lut = env->brdf_lut;
            e.voxel_grid = env->voxel_grid;
            std::memcpy(e.light_direction, env->light_direction, 12);
            std::memcpy(e.light_color, env->light_color, 12);
        }
        return swr::gltf::Document::load(path, e).release();
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return nullptr;
    }
}
void swrh_gltf_free(void *doc) { delete (swr::gltf::Document *)doc; }
const swr_scene_desc *swrh_gltf_scene(void *doc) { return doc ? &((swr::gltf::Document *)doc)->desc : nullptr; }
int swrh_gltf_get_info(void *doc, swrh_gltf_info *out) {
    if (!doc || !out) return -1;
    const swr::gltf::Document &d = *(swr::gltf::Document *)doc;
    std::memcpy(out->bounds_min, d.bounds_min, 12);
    std::memcpy(out->bounds_max, d.bounds_max, 12);
    std::memcpy(out->bounds_center, d.bounds_center, 12);
    out->bounds_diagonal = d.bounds_diagonal;
    out->ncameras = (uint32_t)d.cameras.size();
    uint32_t nfile = 0;
    for (const std::string &u : d.texture_uri) nfile += u.empty() || u[0] != '<' ? 1u : 0u;
    out->nfile_textures = nfile;
    return 0;
}
int swrh_gltf_get_camera(void *doc, uint32_t index, swrh_gltf_camera *out) {
    if (!doc || !out) return -1;
    const swr::gltf::Document &d = *(swr::gltf::Document *)doc;
    if (index >= d.cameras.size()) {
        g_err = "camera index out of range";
        return -1;
    }
    const swr::gltf::CameraData &c = d.cameras[index];
    out->perspective = c.perspective ? 1 : 0;
    out->yfov_or_xmag = c.fov_or_xmag;
    out->aspect_or_ymag = c.aspect_or_ymag;
    out->znear = c.znear;
    out->zfar = c.zfar;
    std::memcpy(out->transform, c.transform, 64);
    return 0;
}
const char *swrh_gltf_texture_uri(void *doc, uint32_t slot) {
    if (!doc) return nullptr;
    const swr::gltf::Document &d = *(swr::gltf::Document *)doc;
    return slot < d.texture_uri.size() ? d.texture_uri[slot].c_str() : nullptr;
}
int swrh_gltf_register_image(const char *uri, const uint8_t *rgba, uint32_t width, uint32_t height) {
    if (!uri) return -1;
    std::lock_guard<std::mutex> lock(swr::gltf::Document::registered_images_mutex());
    auto &m = swr::gltf::Document::registered_images();
    if (!rgba) {
        m.erase(uri);
        return 0;
    }
    swr::gltf::Image img;
    img.width = width;
    img.height = height;
    img.rgba.assign(rgba, rgba + (size_t)width * height * 4);
    m[uri] = std::move(img);
    return 0;
}
int swrh_compute_smooth_normals(const float *positions4, uint32_t nverts, const uint32_t *indices, uint32_t nindices, float *normals4_out) {
    try {
        std::vector<float> pos(positions4, positions4 + (size_t)nverts * 4);
        std::vector<uint32_t> idx(indices, indices + nindices);
        for (uint32_t i : idx)
            if (i >= nverts) throw std::runtime_error("Invalid data: vertex index out of range");
        std::vector<float> n = swr::gltf::compute_smooth_normals(pos, idx);
        std::memcpy(normals4_out, n.data(), n.size() * 4);
        return 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return -1;
    }
}
int swrh_compute_tangents(const float *positions4, const float *texcoords2, const float *normals4, uint32_t nverts, const uint32_t *indices,
                          uint32_t nindices, float *tangents4_out) {
    try {
        std::vector<float> pos(positions4, positions4 + (size_t)nverts * 4);
        std::vector<float> uv(texcoords2, texcoords2 + (size_t)nverts * 2);
        std::vector<float> nrm(normals4, normals4 + (size_t)nverts * 4);
        std::vector<uint32_t> idx(indices, indices + nindices);
        for (uint32_t i : idx)
            if (i >= nverts) throw std::runtime_error("Invalid data: vertex index out of range");
        std::vector<float> t = swr::gltf::compute_tangents(pos, uv, nrm, idx);
        std::memcpy(tangents4_out, t.data(), t.size() * 4);
        return 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return -1;
    }
}
int swrh_build_mip_chain(const uint32_t *base_texels, uint32_t width, uint32_t height, uint32_t texture_type, uint32_t *data_out, uint32_t *ntexels_out,
                         uint32_t *nmips_out, uint32_t *mip_table_out) {
    try {
        if (!base_texels || !width || !height) throw std::runtime_error("Invalid data: empty image");
        swr::gltf::TextureData t;
        const uint32_t slices = texture_type == SWR_TEX_CUBEMAP ? 6u : 1u;
        t.width = width, t.height = height, t.type = texture_type;
        t.data.assign(base_texels, base_texels + (size_t)width * height * slices);
        t.mip_offsets = {0}, t.mip_widths = {width}, t.mip_heights = {height};
        t.array_stride = {slices > 1 ? width * height : 0u};
        t.generate_mipmaps();
        const uint32_t nm = t.max_mip_level() + 1;
        if (ntexels_out) *ntexels_out = (uint32_t)t.data.size();
        if (nmips_out) *nmips_out = nm;
        if (data_out) std::memcpy(data_out, t.data.data(), t.data.size() * 4);
        if (mip_table_out)
            for (uint32_t i = 0; i < nm; i++) {
                mip_table_out[i] = t.mip_offsets[i];
                mip_table_out[nm + i] = t.mip_widths[i];
                mip_table_out[2 * nm + i] = t.mip_heights[i];
                mip_table_out[3 * nm + i] = t.array_stride[i];
            }
        return 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return -1;
    }
}
int swrh_decode_png(const uint8_t *file, size_t nbytes, uint8_t *rgba_out, uint32_t *width_out, uint32_t *height_out) {
    try {
        if (!file) throw std::runtime_error("Invalid data: null image");
        swr::gltf::Image img = swr::gltf::decode_image(std::vector<uint8_t>(file, file + nbytes), "<memory>");
        if (width_out) *width_out = img.width;
        if (height_out) *height_out = img.height;
        if (rgba_out) std::memcpy(rgba_out, img.rgba.data(), img.rgba.size());
        return 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return -1;
    }
}

// ---- environment bakes (include/swr_gltf.h, host/swr_bake.hpp) ----------------------------------------------------------
void *swrh_env_bake(const uint8_t *cross_rgba, uint32_t width, uint32_t height, uint32_t lut_size, uint32_t specular_samples, uint32_t voxel_dim,
                    float irradiance_scale, float sky_visibility, float light_intensity) {
    try {
        if (!cross_rgba || !lut_size || !specular_samples || !voxel_dim) throw std::runtime_error("Invalid data: null image or zero size");
        if (lut_size > 4096 || voxel_dim > 1024 || width > 32768 || height > 32768) throw std::runtime_error("Invalid data: bake size out of range");
        swr::gltf::Image img;
        img.width = width;
        img.height = height;
        img.rgba.assign(cross_rgba, cross_rgba + (size_t)width * height * 4);
        return swr::bake::EnvironmentBake::from_cross(img, lut_size, specular_samples, voxel_dim, irradiance_scale, sky_visibility, light_intensity).release();
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return nullptr;
    }
}
void *swrh_env_bake_cached(const uint8_t *cross_rgba, uint32_t width, uint32_t height, uint32_t lut_size, uint32_t specular_samples, uint32_t voxel_dim,
                           float irradiance_scale, float sky_visibility, float light_intensity, const char *ggx_cache_path) {
    try {
        if (!cross_rgba || !lut_size || !specular_samples || !voxel_dim) throw std::runtime_error("Invalid data: null image or zero size");
        if (lut_size > 4096 || voxel_dim > 1024 || width > 32768 || height > 32768) throw std::runtime_error("Invalid data: bake size out of range");
        swr::gltf::Image img;
        img.width = width;
        img.height = height;
        img.rgba.assign(cross_rgba, cross_rgba + (size_t)width * height * 4);
        return swr::bake::EnvironmentBake::from_cross(img, lut_size, specular_samples, voxel_dim, irradiance_scale, sky_visibility, light_intensity, ggx_cache_path)
            .release();
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return nullptr;
    }
}
int swrh_env_specular_from_cache(void *env) { return env ? (((swr::bake::EnvironmentBake *)env)->specular_from_cache ? 1 : 0) : -1; }

// ---- bake caches (include/swr_gltf.h, host/swr_cache.hpp) -----------------------------------------------------------------
int swrh_ggx_cache_load(const char *path, uint32_t width, uint32_t height, uint32_t *out_texels) {
    try {
        if (!path || !out_texels) throw std::runtime_error("Invalid data: null argument");
        std::vector<uint32_t> t;
        if (!swr::cache::ggx_load(path, width, height, t)) return 0;
        std::memcpy(out_texels, t.data(), t.size() * sizeof(uint32_t));
        return 1;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return -1;
    }
}
int swrh_ggx_cache_save(const char *path, uint32_t width, uint32_t height, uint32_t mips, const uint32_t *texels) {
    try {
        if (!path || !texels) throw std::runtime_error("Invalid data: null argument");
        swr::cache::ggx_save(path, width, height, mips, texels, (size_t)width * height * 6 * mips);
        return 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return -1;
    }
}
int swrh_gi_cache_load(const char *path, uint32_t w, uint32_t h, uint32_t d, float *gi_sh4_out) {
    try {
        if (!path || !gi_sh4_out) throw std::runtime_error("Invalid data: null argument");
        return swr::cache::gi_load(path, w, h, d, gi_sh4_out) ? 1 : 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return -1;
    }
}
int swrh_gi_cache_save(const char *path, uint32_t w, uint32_t h, uint32_t d, const float *gi_sh4) {
    try {
        if (!path || !gi_sh4) throw std::runtime_error("Invalid data: null argument");
        swr::cache::gi_save(path, w, h, d, gi_sh4);
        return 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return -1;
    }
}

int swrh_env_get(void *env, swrh_gltf_env *out, float irradiance_sh_out[12]) {
    if (!env || !out) return -1;
    swr::bake::EnvironmentBake &e = *(swr::bake::EnvironmentBake *)env;
    out->cubemap = &e.descs[0];
    out->cubemap_specular = &e.descs[1];
    out->brdf_lut = &e.descs[2];
    for (int c = 0; c < 3; c++) out->voxel_grid.dims[c] = e.voxel_dims[c];
    out->voxel_grid.gi_sh4 = e.voxels.data();
    if (irradiance_sh_out) std::memcpy(irradiance_sh_out, e.irradiance_sh, sizeof(e.irradiance_sh));
    return 0;
}
void swrh_env_free(void *env) { delete (swr::bake::EnvironmentBake *)env; }

// ---- voxel sun visibility (include/swr_gltf.h, host/swr_sunvis.hpp) ------------------------------------------------------
int swrh_compute_sun_visibility(const swr_scene_desc *scene, float *out_per_voxel) {
    try {
        if (!scene || !out_per_voxel) throw std::runtime_error("Invalid data: null argument");
        swr::validate_scene_ranges(*scene);
        std::vector<float> v = swr::sunvis::compute_sun_visibility(*scene, scene->voxel_grid, scene->light_direction);
        std::memcpy(out_per_voxel, v.data(), v.size() * sizeof(float));
        return 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return -1;
    }
}
int swrh_gltf_bake_sun_visibility(void *doc) {
    try {
        if (!doc) throw std::runtime_error("Invalid data: null document");
        swr::gltf::Document &d = *(swr::gltf::Document *)doc;
        if (d.voxels.empty()) throw std::runtime_error("Invalid data: the document has no voxel grid (load it with an environment)");
        std::vector<float> v = swr::sunvis::compute_sun_visibility(d.desc, d.desc.voxel_grid, d.desc.light_direction);
        for (size_t i = 0; i < v.size(); i++) d.voxels[i * 16 + 3] = v[i];
        return 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return -1;
    }
}
}

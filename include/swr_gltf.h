/* swr_gltf.h — glTF 2.0 / GLB loader in front of the hot path (C ABI of libswr_host.so).
 *
 * Produces the flat swr_scene_desc that swr_upload_scene consumes, following the reference's own loader for every
 * field the renderer reads (src/scene.rs:145-354 Scene::from_gltf, :383-419 nodes, :440-502 primitives with automatic
 * normals/tangents :520-646, :648-787 textures/samplers/materials, :789-818 cameras; src/texture.rs:45-128 mip chains,
 * :897-1010 texel packing). A Rust host keeps using its own `Scene`; this entry point is for hosts
 * that have no loader of their own (C, C++, Python).
 *
 * Not in a glTF file and therefore supplied by the caller: the sky cubemap, the prefiltered specular cubemap, the BRDF
 * LUT and the GI voxel grid (the reference bakes them from assets/cubemap.jpg at load time).
 * Images: PNG and JPEG files / data URIs are decoded here (host/swr_gltf.hpp, host/swr_jpeg.hpp); other formats can be
 * registered decoded beforehand.
 * All functions return 0 / a handle on success, -1 / NULL on error with the message in swrh_last_error() — the wording
 * follows the reference's SceneError ("Missing data: No positions in primitive", ...). */
#ifndef SWR_GLTF_H
#define SWR_GLTF_H
#include "swr.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct swrh_gltf_env {
    const swr_texture_desc *cubemap;          /* type SWR_TEX_CUBEMAP, may be NULL (then the scene cannot be rendered) */
    const swr_texture_desc *cubemap_specular;
    const swr_texture_desc *brdf_lut;
    swr_voxel_grid_desc voxel_grid;           /* copied; world_min == world_max means: span the scene bounds (main.rs:228-235) */
    float light_direction[3];                 /* scene.rs:236-239 */
    float light_color[3];
} swrh_gltf_env;

typedef struct swrh_gltf_info {
    float bounds_min[3], bounds_max[3], bounds_center[3]; /* scene.rs:331-351 */
    float bounds_diagonal;
    uint32_t ncameras;
    uint32_t nfile_textures; /* texture slots that came from the file; the environment follows them */
} swrh_gltf_info;

/* scene.rs:123-143 SceneCamera: projection parameters + the world transform of the node that carries the camera */
typedef struct swrh_gltf_camera {
    int32_t perspective;   /* 1: yfov / aspect, 0: xmag / ymag */
    float yfov_or_xmag, aspect_or_ymag, znear, zfar;
    float transform[16];
} swrh_gltf_camera;

const char *swrh_last_error(void);
void *swrh_gltf_load(const char *path, const swrh_gltf_env *env);
void swrh_gltf_free(void *doc);
const swr_scene_desc *swrh_gltf_scene(void *doc); /* valid until swrh_gltf_free */
int swrh_gltf_get_info(void *doc, swrh_gltf_info *out);
int swrh_gltf_get_camera(void *doc, uint32_t index, swrh_gltf_camera *out);
const char *swrh_gltf_texture_uri(void *doc, uint32_t slot);
/* Hand in an image the loader cannot decode itself (width*height*4 bytes, R G B A per texel, as image::to_rgba8);
 * looked up by the exact `uri` string of the glTF image. rgba == NULL removes the entry. */
int swrh_gltf_register_image(const char *uri, const uint8_t *rgba, uint32_t width, uint32_t height);

/* Load-time environment bakes (scene.rs:151-231 + main.rs:228-281). The three integrals (BRDF LUT, irradiance SH4, GGX
 * prefilter) run on the CURRENT CUDA device (include/swr.h swr_bake_*; there is no CPU fallback), the glue (faces, mip chains,
 * voxel fill) on the host. From a sky image in
 * the reference's cross layout (+Y on top; -X +Z +X -Z in the middle row; -Y below; face size = width/4 x height/3) build
 * the sky cubemap with mips (texture.rs:922-959, :45-128), the GGX-prefiltered specular cubemap (texture.rs:330-420, every
 * mip at full face resolution, `specular_samples` = 64 in the reference), the irradiance SH4 (texture.rs:289-328), the
 * BRDF LUT (texture.rs:199-235, `lut_size` = 128) and a voxel grid initialised from the SH (gi.rs:123-149 with main.rs's
 * irradiance_scale = 0.25, sky_visibility = 1.0; `light_intensity` stands in for the ray-cast sun visibility, N4).
 * swrh_env_get fills the three texture pointers and voxel_grid.dims / gi_sh4 of `out` (world bounds and light stay the
 * caller's); they stay valid until swrh_env_free. */
void *swrh_env_bake(const uint8_t *cross_rgba, uint32_t width, uint32_t height, uint32_t lut_size, uint32_t specular_samples, uint32_t voxel_dim,
                    float irradiance_scale, float sky_visibility, float light_intensity);
/* The same with the reference's `.ggx` cache (scene.rs:164-206, texture.rs:426-552): when `ggx_cache_path` names a cache of the
 * sky's face size the prefiltered cubemap is read from it; otherwise it is baked on the device and the cache is written.
 * swrh_env_specular_from_cache says which happened. */
void *swrh_env_bake_cached(const uint8_t *cross_rgba, uint32_t width, uint32_t height, uint32_t lut_size, uint32_t specular_samples, uint32_t voxel_dim,
                           float irradiance_scale, float sky_visibility, float light_intensity, const char *ggx_cache_path);
int swrh_env_specular_from_cache(void *env);
int swrh_env_get(void *env, swrh_gltf_env *out, float irradiance_sh_out[12]);
void swrh_env_free(void *env);
/* Voxel sun visibility as the reference's default load path computes it (main.rs:237-246; gi.rs:151-314, raytracer.rs,
 * voxelgrid.rs:371-419); triangles, hierarchy and active-voxel mask are built on the host, the rays and the blur run on the
 * current CUDA device (swr_bake_sun_visibility, no CPU fallback): voxels near geometry cast one ray towards scene->light_direction, opaque hits give 0, translucent
 * hits multiply their transmission, the grid is blurred (3x3x3 mean, squared). out_per_voxel receives dims[0]*dims[1]*dims[2]
 * floats (index z*w*h + y*w + x); the value belongs into gi_sh4[voxel][0].w. swrh_gltf_bake_sun_visibility does that for
 * a loaded document in place (call it before the scene is first rendered: uploads are once per scene). */
int swrh_compute_sun_visibility(const swr_scene_desc *scene, float *out_per_voxel);
int swrh_gltf_bake_sun_visibility(void *doc);

/* The reference's bake caches on their own. `.ggx` (texture.rs:12-17, 426-552): "GGX0", version 1, width, height, mips, then the
 * brotli stream of width*height*6*mips RGBA8 texels (every mip at full face resolution, mips = 1 + ilog2(max(w, h))). `.gi`
 * (gi.rs:17-22, 30-122): "VGI0", version 8, w, h, d, then the brotli stream of w*h*d x 4 coefficients x (r, g, b, w) f32 — the
 * layout of swr_voxel_grid.gi_sh4. Loads return 1 = read, 0 = no usable cache (missing file, other magic / version / size, payload
 * of another length: the reference then bakes and saves), -1 = error (swrh_last_error). Brotli is the system's libbrotlidec /
 * libbrotlienc, bound at run time; without them the calls fail. */
int swrh_ggx_cache_load(const char *path, uint32_t width, uint32_t height, uint32_t *out_texels);
int swrh_ggx_cache_save(const char *path, uint32_t width, uint32_t height, uint32_t mips, const uint32_t *texels);
int swrh_gi_cache_load(const char *path, uint32_t w, uint32_t h, uint32_t d, float *gi_sh4_out);
int swrh_gi_cache_save(const char *path, uint32_t w, uint32_t h, uint32_t d, const float *gi_sh4);

/* The pieces of the loader that are useful on their own (and are what the tests pin): */
int swrh_compute_smooth_normals(const float *positions4, uint32_t nverts, const uint32_t *indices, uint32_t nindices, float *normals4_out);
int swrh_compute_tangents(const float *positions4, const float *texcoords2, const float *normals4, uint32_t nverts, const uint32_t *indices,
                          uint32_t nindices, float *tangents4_out);
/* Type-aware mip chain of one RGBA8 image (texture.rs:45-128). Call with data_out == NULL to get the texel count and the
 * number of mips; mip_table_out receives 4 x nmips u32: offsets, widths, heights, array strides. */
int swrh_build_mip_chain(const uint32_t *base_texels, uint32_t width, uint32_t height, uint32_t texture_type, uint32_t *data_out,
                         uint32_t *ntexels_out, uint32_t *nmips_out, uint32_t *mip_table_out);
/* PNG or JPEG (by signature) -> RGBA8; call with rgba_out == NULL for the size. (The name predates the JPEG decoder.) */
int swrh_decode_png(const uint8_t *file, size_t nbytes, uint8_t *rgba_out, uint32_t *width_out, uint32_t *height_out);

#ifdef __cplusplus
}
#endif
#endif

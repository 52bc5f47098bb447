/*
 * swr.h — C ABI of the B200-native rasterisation hot path (libswr_b200.so).
 *
 * This is the drop-in boundary under the reference's `Renderer`
 * (reference: src/renderer.rs:165 `Renderer::new`, :201 `render_scene`,
 * :258 `update_auto_exposure`, :293 `blit_to_buffer`).  The reference has no
 * FFI of its own; the entry points below are what a Rust shim behind
 * `renderer.rs` binds (see INTEGRATION.md for the `extern "C"` block).
 *
 * Conventions
 *  - plain C99, pointers + sizes only, no C++/torch types;
 *  - matrices are column-major 16 floats, exactly the bytes of glam `Mat4`;
 *  - every call returns 0 on success or a negative swr_status; the message is
 *    available from swr_last_error();
 *  - one context per GPU, callable from any host thread, one call at a time
 *    per context (mirrors `&mut self` on the reference's Renderer);
 *  - there is NO CPU fallback: if no CUDA device is usable swr_create fails.
 */
#ifndef SWR_H_
#define SWR_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWR_ABI_VERSION 1

/* Numerical contract constants (reference: renderer.rs:14, tilerasterizer.rs:10-22). */
#define SWR_TILE_SIZE 64
#define SWR_SUBPIXEL_SHIFT 4
#define SWR_SUBPIXEL_SCALE 16
#define SWR_COARSE_BLOCK_PIXELS 16
#define SWR_DEFAULT_EXPOSURE 2.0f

typedef enum swr_status {
    SWR_OK = 0,
    SWR_ERR_INVALID = -1,   /* bad argument / inconsistent scene description */
    SWR_ERR_CUDA = -2,      /* CUDA runtime error (message has the detail) */
    SWR_ERR_NO_DEVICE = -3, /* no usable sm_100 device: there is no CPU fallback */
    SWR_ERR_NO_SCENE = -4,  /* render before upload */
    SWR_ERR_OOM = -5
} swr_status;

/* texture.rs:19-26 */
typedef enum swr_texture_type {
    SWR_TEX_SRGB = 0,
    SWR_TEX_NORMAL = 1,
    SWR_TEX_METALLIC_ROUGHNESS = 2,
    SWR_TEX_CUBEMAP = 3,
    SWR_TEX_LINEAR = 4
} swr_texture_type;

/* texture.rs:554-558 */
typedef enum swr_wrap_mode {
    SWR_WRAP_REPEAT = 0,
    SWR_WRAP_MIRRORED_REPEAT = 1,
    SWR_WRAP_CLAMP_TO_EDGE = 2
} swr_wrap_mode;

/* scene.rs:94-102 `Primitive`.  All attribute arrays have nverts entries. */
typedef struct swr_primitive_desc {
    const float *positions;  /* nverts x 4 (glam Vec4, w = 1.0)            */
    const float *normals;    /* nverts x 4 (glam Vec3A: 16-byte stride)    */
    const float *tangents;   /* nverts x 4 (xyz + handedness)              */
    const float *texcoords;  /* nverts x 2                                  */
    const uint32_t *indices; /* nindices, triangle list                     */
    uint32_t nverts;
    uint32_t nindices;
    uint32_t material_index;
    float bounding_sphere[4]; /* centre xyz, radius (scene.rs:31-34)        */
} swr_primitive_desc;

/* scene.rs:87-92 `Mesh`: a contiguous range of the scene's primitive array.
 * primitives_opaque / primitives_translucent are derived from the material's
 * translucent flag in primitive order (scene.rs:276-284). */
typedef struct swr_mesh_desc {
    uint32_t first_primitive;
    uint32_t num_primitives;
} swr_mesh_desc;

/* scene.rs:79-85 `Node` (hierarchy already flattened). */
typedef struct swr_node_desc {
    float transform[16];            /* local -> world */
    int32_t mesh_index;             /* -1 = none */
    float bounding_sphere_world[4]; /* centre xyz, radius */
} swr_node_desc;

/* texture.rs:28-42 `Texture` + :565-570 `Sampler`. RGBA8 with R in bits 31..24. */
typedef struct swr_texture_desc {
    const uint32_t *data; /* all mips (and 6 faces for cubemaps) concatenated */
    uint32_t ntexels;
    uint32_t width, height;
    uint32_t texture_type;  /* swr_texture_type */
    uint32_t max_mip_level; /* number of mips - 1 */
    const uint32_t *mip_offsets;  /* max_mip_level + 1 entries each */
    const uint32_t *mip_widths;
    const uint32_t *mip_heights;
    const uint32_t *array_stride; /* texels between faces per mip */
    uint32_t wrap_s, wrap_t;      /* swr_wrap_mode */
} swr_texture_desc;

#define SWR_MAT_ALPHA_TESTED 1u
#define SWR_MAT_TRANSLUCENT 2u

/* scene.rs:104-121 `Material`. Texture slots index swr_scene_desc.textures; -1 = None. */
typedef struct swr_material_desc {
    float base_color_factor[4];
    float metallic_factor;
    float roughness_factor;
    float emissive_factor[3];
    float occlusion_strength;
    float transmission;
    float alpha_cutoff;
    uint32_t flags;
    int32_t base_color_texture;
    int32_t metallic_roughness_texture;
    int32_t normal_texture;
    int32_t emissive_texture;
    int32_t occlusion_texture;
    int32_t transmission_texture;
} swr_material_desc;

/* voxelgrid.rs:6-18. gi_sh4 index = z*w*h + y*w + x, 4 coeffs x Vec4 per voxel. */
typedef struct swr_voxel_grid_desc {
    uint32_t dims[3];
    float world_min[3];
    float world_max[3];
    const float *gi_sh4; /* nvoxels x 16 floats */
} swr_voxel_grid_desc;

/* scene.rs:65-77 `Scene` (immutable after load; uploaded once). */
typedef struct swr_scene_desc {
    const swr_primitive_desc *primitives;
    uint32_t nprimitives;
    const swr_mesh_desc *meshes;
    uint32_t nmeshes;
    const swr_node_desc *nodes;
    uint32_t nnodes;
    const swr_material_desc *materials;
    uint32_t nmaterials;
    const swr_texture_desc *textures;
    uint32_t ntextures;
    swr_voxel_grid_desc voxel_grid;
    int32_t cubemap;          /* texture index (type CUBEMAP) */
    int32_t cubemap_specular; /* texture index (prefiltered, all mips full res) */
    int32_t brdf_lut;         /* texture index */
    float light_direction[3]; /* scene.rs:236-239 */
    float light_color[3];
} swr_scene_desc;

/* rendercamera.rs:5-25: the cached matrices are INPUTS of the path. */
typedef struct swr_camera {
    float position[4];
    float view_matrix[16];
    float view_project_matrix[16];
    float skybox_matrix_transposed[16];
    float view_clip_planes[6][4];
    float one_over_width;
    float one_over_height;
    float reserved[2];
} swr_camera;

#define SWR_DRAW_CLIP 1u        /* primitive sphere intersects the frustum: renderer.rs:453-462 */
#define SWR_DRAW_TRANSLUCENT 2u /* primitive of mesh.primitives_translucent (renderer.rs:407-419): forward-shaded after the opaque
                                 * pass, back to front per tile; first_triangle counts within the translucent draws only */

/* One (node, opaque primitive) pair that survived the sphere/frustum test, in the
 * reference's serial submission order (renderer.rs:207 -> :396 -> :490).
 * `first_triangle` is the running sum of triangle counts of the draws before
 * it; the fragment tie-break id is seq = (first_triangle + tri) * 8 + fan. */
typedef struct swr_draw {
    float model[16];
    float mvp[16];
    uint32_t primitive; /* index into swr_scene_desc.primitives */
    uint32_t flags;
    uint32_t first_triangle;
    uint32_t reserved;
} swr_draw;

typedef struct swr_frame_stats {
    uint64_t triangles_submitted; /* T: input triangles over all draws       */
    uint64_t vertices_submitted;  /* V: sum over draws of primitive nverts   */
    uint64_t triangles_binned;    /* after cull/clip (fan triangles counted) */
    uint64_t triangles_clipped;   /* polygons that went through the clipper  */
    uint64_t tile_refs;           /* R: (triangle, tile) references          */
    uint32_t tiles;
    uint32_t clusters_culled;     /* 128-triangle clusters rejected before set-up (frustum / row band) */
    float ms_setup_bin;           /* CUDA-event ms: set-up + count + scan + scatter */
    float ms_raster;              /* tile rasteriser */
    float ms_shade;               /* vis-buffer shading */
    float ms_resolve;             /* last swr_resolve */
} swr_frame_stats;

typedef struct swr_ctx swr_ctx;

int swr_abi_version(void);
/* sizeof() of the ABI structs as compiled into the library, for binding self-checks:
 * 0 primitive, 1 mesh, 2 node, 3 texture, 4 material, 5 voxel grid, 6 scene, 7 camera, 8 draw, 9 stats. */
size_t swr_sizeof(int which);
const char *swr_last_error(const swr_ctx *ctx); /* ctx may be NULL: last create error */

/* Renderer::new (renderer.rs:165). device = CUDA ordinal. */
swr_ctx *swr_create(int width, int height, int device);
void swr_destroy(swr_ctx *ctx);

/* Sort-first partition: this context owns tile rows [row_begin,row_end) only
 * (default: all rows).  Triangles are binned, rasterised and shaded for owned
 * tiles; other pixels are left untouched. */
int swr_set_tile_rows(swr_ctx *ctx, int row_begin, int row_end);

/* normalize() compatibility (math.rs:34-39, :101-108). On x86-64 the reference normalises with the hardware
 * estimate _mm_rsqrt_ps, whose value is a pure function of the exponent parity and the top `mantissa_bits`
 * mantissa bits of its input (a table that differs between CPU vendors). Passing that table — 2^(bits+1)
 * entries: the result bits for inputs (127+p)<<23 | k<<(23-bits), index p<<bits | k — makes the CUDA shading
 * normalise bit-identically to the reference running on that host. table == NULL (the default) selects
 * rsqrtf(), which is within the estimate's error envelope but not bit-identical to it. */
int swr_set_rsqrt_table(swr_ctx *ctx, const uint32_t *table, int mantissa_bits);

/* Copies the scene to device memory; the descriptor may be freed afterwards. */
int swr_upload_scene(swr_ctx *ctx, const swr_scene_desc *scene);

/* A second context on the SAME device renders the scene `owner` uploaded, without a second copy in device memory (the
 * scene is immutable; `owner` must outlive `ctx` and must not upload again while `ctx` uses the scene). This is how a
 * pipelined host alternates frames between two contexts — two streams, two sets of per-frame buffers — so that frame
 * N+1's geometry pass fills the SMs frame N's raster tail and single-block phases leave idle (host mirror: lanes). */
int swr_share_scene(swr_ctx *ctx, const swr_ctx *owner);

/* Device half of render_scene (renderer.rs:201-220): set-up, clip, bin, raster,
 * shade.  Asynchronous on the context's stream. With shade=0 only the
 * visibility buffer is produced (used by the sort-last composite). */
int swr_render(swr_ctx *ctx, const swr_camera *camera, const swr_draw *draws, int ndraws, int shade);

/* Fixed-exposure frames. The reference only changes its exposure in update_auto_exposure (renderer.rs:258-290); a host that
 * knows the exposure BEFORE the frame is shaded (auto exposure idle, dt = 0, a fixed-exposure capture) says so here, and
 * shading then applies exposure, tonemap and the RGBA8 pack of blit_to_buffer (renderer.rs:293-355) itself: the 16-byte HDR
 * colour per pixel is neither written nor read back, and the tile metering values (tilerasterizer.rs:103-106) come from the
 * same pass. exposure > 0 switches it on for frames rendered from now on, 0 switches it off (default). Plain opaque frames
 * only: a frame with translucent draws or a sort-last composite is shaded to HDR as before. Such a frame can be resolved
 * with exactly that exposure; any other exposure (and swr_read_color) fails with SWR_ERR_INVALID — call
 * swr_set_fixed_exposure(ctx, 0) and swr_shade() to shade the same visibility buffer to HDR (the host mirror does that by
 * itself when update_auto_exposure moved the exposure between render_scene and blit_to_buffer). */
int swr_set_fixed_exposure(swr_ctx *ctx, float exposure);

/* Shade the current visibility-key buffer again (same frame, e.g. a different camera block is NOT supported:
 * the records belong to the last swr_render). */
int swr_shade(swr_ctx *ctx, const swr_camera *camera);

/* Sort-last composite (one context per rank, each rendered with shade = 0 from its own subset of the global
 * draw list; first_triangle stays global):
 *   1. swr_keys_to_global    low word of every key becomes ~seq, so keys are comparable across ranks;
 *   2. (caller) unsigned 64-bit MIN all-reduce over swr_device_keys();
 *   3. swr_keys_localize     winners owned by this rank get their local id back, others are marked foreign;
 *                            the winner's barycentrics are written to swr_device_bary() (0 where not owned);
 *   4. (caller) float SUM all-reduce over swr_device_bary()  — the quad-coupled mip needs every lane's barycentrics;
 *   5. swr_shade_composited  shades the pixels this rank owns, sky only in tile rows [sky_row_begin, sky_row_end);
 *   6. swr_resolve + (caller) integer SUM over swr_device_pixels(): unowned pixels resolve to 0. */
int swr_keys_to_global(swr_ctx *ctx);
int swr_keys_localize(swr_ctx *ctx);
int swr_shade_composited(swr_ctx *ctx, const swr_camera *camera, int sky_row_begin, int sky_row_end);
void *swr_device_bary(swr_ctx *ctx); /* W*H float2, row-major */

/* blit_to_buffer (renderer.rs:293-355): exposure, tonemap, RGBA8 pack into a
 * row-major W*H u32 image on the device; if out_pixels != NULL it is copied to
 * that HOST buffer and the call synchronises. */
int swr_resolve(swr_ctx *ctx, float exposure, uint32_t *out_pixels);

/* Pipelined form of swr_resolve for callers that overlap the read-back of frame N with the rendering of frame N+1 (what
 * the reference's App does with present(N-1) || render(N), main.rs:526-597): resolves into one of two device pixel
 * buffers and copies it to `out_pixels` (pinned host memory) on a separate stream, without blocking. *ticket identifies
 * the copy; swr_wait_pixels(ticket) blocks until `out_pixels` is complete. At most two copies may be outstanding. */
int swr_resolve_async(swr_ctx *ctx, float exposure, uint32_t *out_pixels, int *ticket);
int swr_wait_pixels(swr_ctx *ctx, int ticket);

/* Per-tile cost signals of the last frame (row-major tiles; 0 for tiles this context does not own), for balancing
 * sort-first row bands: triangle references binned into the tile and SM cycles its rasterisation took. Either may be NULL. */
int swr_read_tile_costs(swr_ctx *ctx, uint32_t *refs_per_tile, uint32_t *raster_cycles_per_tile);

/* Per-tile metering luminance (tilerasterizer.rs:103-106), row-major tiles. */
int swr_read_tile_luminance(swr_ctx *ctx, float *out_per_tile);

/* Parity read-back: per pixel depth bits, seq id and the two barycentrics, row-major
 * W*H each (any pointer may be NULL). Uncovered = (+INF bits, 0xFFFFFFFF, 0, 0). */
int swr_read_visbuffer(swr_ctx *ctx, uint32_t *depth_bits, uint32_t *seq, float *bary1, float *bary2);

/* Linear HDR colour before exposure/tonemap, row-major W*H*3 floats (parity only). */
int swr_read_color(swr_ctx *ctx, float *rgb);

int swr_synchronize(swr_ctx *ctx);
int swr_get_stats(swr_ctx *ctx, swr_frame_stats *out);
/* Number of CUDA kernels this context has launched since swr_create (monotonic; what bench.py reports as gpu_launches). */
uint64_t swr_launch_count(swr_ctx *ctx);

/* Sort-first frame assembly over NVLink peer memory: instead of resolving locally and gathering strips with a
 * collective, every contributing rank's resolve kernel stores its rows straight into the assembling rank's pixel buffer
 * and signals completion there (one kernel: tonemap + pack + peer stores + release). Replaces the strip gather a host
 * would otherwise issue after Renderer::blit_to_buffer (renderer.rs:293-355) per rank.
 *   assembling rank:   swr_peer_export -> 64-byte handle (send it to the others with any transport);
 *                      per frame f = 1,2,...: swr_render, swr_resolve(NULL) (its own rows), swr_peer_collect(f, n_contributors),
 *                      consume swr_device_pixels() on the context's stream, swr_peer_release(f);
 *   contributing rank: swr_peer_open(handle) once (or swr_peer_attach with a device pointer that is already mapped
 *                      in this process: several contexts driven by one process), per frame swr_render,
 *                      swr_resolve_peer(exposure, f).
 * Everything is enqueued on the contexts' streams; the waits are device-side and give up after ~2 s
 * (SWR_ERR_CUDA at the next synchronising call) so a protocol error cannot wedge the GPU. Frame numbers must
 * increase by one per frame on every rank. Do not mix with swr_resolve_async on the assembling rank. */
#define SWR_PEER_HANDLE_BYTES 64
int swr_peer_export(swr_ctx *ctx, void *handle_out);
int swr_peer_open(swr_ctx *ctx, const void *handle);
int swr_peer_attach(swr_ctx *ctx, void *assembler_device_pixels);
int swr_resolve_peer(swr_ctx *ctx, float exposure, uint32_t frame);
int swr_peer_collect(swr_ctx *ctx, uint32_t frame, int contributors);
int swr_peer_release(swr_ctx *ctx, uint32_t frame);

/* ---- several devices of one process behind one handle ------------------------------------------------------------
 * What a single `Renderer` (renderer.rs:145-355; SURVEY 8b: "swr_create(w, h, devices, ndev, mode)") needs to drive all
 * GPUs of a box: the scene is replicated, device i owns a contiguous range of tile rows whose measured cost is 1/ndev of
 * the frame's (probed on the first frame after an upload), every device culls the same draw list against its band, and
 * the frame is assembled in devices[0]'s pixel buffer by the other devices' resolve kernels storing over NVLink peer
 * memory (the swr_peer_* protocol with direct peer pointers: no IPC, no collective). One persistent host thread per
 * device does the enqueueing. Same call order as the single-device API:
 *   swr_multi_upload_scene; per frame swr_multi_render, [swr_multi_read_tile_luminance -> update_auto_exposure on the
 *   host], swr_multi_resolve(exposure, host pixels or NULL). ndev = 1 is the single-device path through the same calls.
 * Sort-last (shard by primitive + depth composite) keeps one context per rank: swr_keys_to_global / swr_keys_localize. */
#define SWR_MULTI_SORT_FIRST 1
typedef struct swr_multi swr_multi;
swr_multi *swr_multi_create(int width, int height, const int *devices, int ndev, int mode);
void swr_multi_destroy(swr_multi *m);
const char *swr_multi_last_error(const swr_multi *m); /* m may be NULL: last create error */
int swr_multi_device_count(const swr_multi *m);
swr_ctx *swr_multi_context(swr_multi *m, int i); /* the per-device context (parity read-back, statistics) */
int swr_multi_tile_rows(const swr_multi *m, int i, int *row_begin, int *row_end);
int swr_multi_set_rsqrt_table(swr_multi *m, const uint32_t *table, int mantissa_bits);
int swr_multi_upload_scene(swr_multi *m, const swr_scene_desc *scene);
int swr_multi_render(swr_multi *m, const swr_camera *camera, const swr_draw *draws, int ndraws);
int swr_multi_resolve(swr_multi *m, float exposure, uint32_t *out_pixels);
int swr_multi_read_tile_luminance(swr_multi *m, float *out_per_tile);
int swr_multi_get_stats(swr_multi *m, swr_frame_stats *out); /* counters summed over the bands, phase times = slowest device */
int swr_multi_synchronize(swr_multi *m);

/* ---- load-time bakes on the device ---------------------------------------------------------------
 * The inputs of the shading kernel the reference derives when a scene is loaded, computed on the GPU. Stateless calls
 * (device < 0: the current device); host pointers in and out; no CPU fallback. The cheap glue around them (cross -> six
 * faces, mip chains, hierarchy build, voxel marking) stays on the host (swr_gltf.h, like R1/R2 of the frame path).
 *   swr_bake_brdf_lut            generate_brdf_lut texels, size x size                          texture.rs:167-235
 *   swr_bake_irradiance_sh4      compute_irradiance_sh4: 4 coefficients x rgb                   texture.rs:289-328
 *   swr_bake_prefilter_specular  generate_prefiltered_specular_cubemap: num_mips x 6 x h x w texels, every mip at full
 *                                resolution; returns num_mips = floor(log2(max(w, h))) + 1        texture.rs:330-420
 *   swr_bake_sun_visibility      one ray per active voxel towards the light through the caller's hierarchy, opaque hit
 *                                -> 0, translucent hits multiply, then the 3x3x3 blur, squared   gi.rs:267-314,
 *                                raytracer.rs:177-259, voxelgrid.rs:371-419
 * cubemap_faces: mip 0 of the sky, six faces of w x h RGBA8 texels (R in bits 31..24), face-major (+X -X +Y -Y +Z -Z). */
typedef struct swr_bvh_node {
    float lo[3], hi[3];
    uint32_t first, count; /* leaf: order[first .. first + count); inner (count = 0): children first and first + 1 */
} swr_bvh_node;
typedef struct swr_sun_triangle {
    float p0[3], p1[3], p2[3]; /* world space */
    float transmission;        /* material transmission of a translucent surface; < 0 = opaque */
} swr_sun_triangle;
typedef struct swr_sunvis_desc {
    const swr_bvh_node *nodes; /* node 0 = root; nnodes = 0: no geometry */
    uint32_t nnodes;
    const uint32_t *order; /* triangle indices referenced by the leaves */
    uint32_t norder;
    const swr_sun_triangle *triangles;
    uint32_t ntriangles;
    const uint8_t *active; /* per voxel (z*W*H + y*W + x): cast a ray from this voxel */
    uint32_t dims[3];
    float world_min[3], world_max[3];
    float light_direction[3];
} swr_sunvis_desc;
int swr_bake_brdf_lut(int device, uint32_t size, uint32_t *out_texels);
int swr_bake_irradiance_sh4(int device, const uint32_t *cubemap_faces, uint32_t w, uint32_t h, float *out12);
int swr_bake_prefilter_specular(int device, const uint32_t *cubemap_faces, uint32_t w, uint32_t h, uint32_t sample_count, uint32_t *out_texels);
int swr_bake_sun_visibility(int device, const swr_sunvis_desc *desc, float *out_per_voxel);
const char *swr_bake_last_error(void);

/* Device pointers for zero-copy interop (NCCL gather / composite from the host
 * language): RGBA8 image (W*H u32, row-major) and the 64-bit visibility keys
 * (tile-major: tile (ty*tiles_x+tx) owns 4096 consecutive keys; inside the tile pixel
 * (x, y) is entry y*64 + (x ^ (((y >> 1) & 7) << 1)) — the rasteriser's bank-conflict-free shared-memory order, kept so
 * that a tile leaves the SM as one bulk copy; key = orderable(depth)<<32 | ~record id, empty = all ones). Valid until the
 * next swr_render that grows buffers, or swr_destroy. */
void *swr_device_pixels(swr_ctx *ctx);
void *swr_device_keys(swr_ctx *ctx);
size_t swr_device_keys_bytes(swr_ctx *ctx);
void *swr_cuda_stream(swr_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SWR_H_ */

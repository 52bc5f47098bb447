/* swr_host.h — C entry points of the host-side mirror of the reference's Renderer API (libswr_host.so,
 * swraster-viewer_b200/host/swr_host.hpp). They sit ABOVE the device C ABI of swr.h and keep on the host what the
 * reference keeps on the host: RenderCamera matrices (rendercamera.rs:28-150), the node sort and per-primitive
 * sphere/frustum classification (renderer.rs:357-468), auto-exposure metering (renderer.rs:258-290).
 * All int functions return 0 on success, -1 on error (message: swrh_last_error(), per thread). */
#ifndef SWR_HOST_H
#define SWR_HOST_H
#include "swr.h"
#ifdef __cplusplus
extern "C" {
#endif

const char *swrh_last_error(void);

/* RenderCamera::new(position, look_at, fov, width, height, far_plane) + update_matrices (rendercamera.rs:28-86) */
int swrh_camera_build(const float pos[3], const float look_at[3], float fov, float width, float height, float far_plane, swr_camera *out);
/* ... towards a level target, then rotate_mouse(dx, dy) (rendercamera.rs:88-101) */
int swrh_camera_build_rotated(const float pos[3], const float look_at_level[3], float mouse_dx, float mouse_dy, float fov, float width, float height,
                              float far_plane, swr_camera *out);

/* Renderer::new(width, height) (renderer.rs:165) on CUDA device `device`; NULL on error (no CPU fallback) */
void *swrh_renderer_new(int width, int height, int device);
/* the same Renderer with `lanes` (1..4) contexts on one device that frames alternate between (own stream and per-frame
 * buffers each, one shared scene: swr_share_scene): for callers that pipeline with swrh_blit_to_buffer_async, frame N+1's
 * geometry pass then overlaps frame N's raster tail and shading. swrh_renderer_ctx returns the lane of the last frame. */
void *swrh_renderer_new_lanes(int width, int height, int device, int lanes);
int swrh_renderer_lanes(void *renderer);                  /* 0 for a multi-device renderer */
swr_ctx *swrh_renderer_lane_ctx(void *renderer, int lane); /* NULL when out of range */
/* the same Renderer over several CUDA devices of this process (sort-first; include/swr.h swr_multi_*): one frame, one
 * blit, every method below works on it except swrh_set_tile_rows (it assigns its own cost-balanced bands, readable
 * with swrh_renderer_tile_rows) and shard / nshards of swrh_render_scene */
void *swrh_renderer_new_multi(int width, int height, const int *devices, int ndev);
int swrh_renderer_tile_rows(void *renderer, int device_index, int *row_begin, int *row_end);
void swrh_renderer_free(void *renderer);
swr_ctx *swrh_renderer_ctx(void *renderer); /* the device context underneath, for the swr_* calls of swr.h */
/* Renderer::render_scene(&scene, &camera) (renderer.rs:201): uploads the scene on first sight, builds the draw list on
 * the host and enqueues the frame. shade = 0 stops after the visibility buffer; shard / nshards select every nshards-th
 * draw (sort-last), 0 / 1 for everything. */
int swrh_render_scene(void *renderer, const swr_scene_desc *scene, const swr_camera *camera, int shade, int shard, int nshards);
/* The scene upload is cached by descriptor address (the reference's Scene is immutable, scene.rs:65-77). Call this when
 * the same address now describes a different scene: the next swrh_render_scene uploads again. */
int swrh_invalidate_scene(void *renderer);
/* Renderer::update_auto_exposure(delta_time) (renderer.rs:258) and the exposure blit_to_buffer applies */
int swrh_update_auto_exposure(void *renderer, float delta_time);
float swrh_auto_exposure(void *renderer);
/* Fixed-exposure shading (include/swr.h, swr_set_fixed_exposure): while update_auto_exposure leaves the exposure where it
 * is, swrh_render_scene shades straight to RGBA8 with it. Callers that go to the context directly (swrh_renderer_ctx) say
 * what they are about to ask of the frame rendered last: swrh_frame_exposure = "resolve it with this exposure",
 * swrh_frame_hdr = "read its HDR colour". Either shades the frame again to HDR if it has to; swrh_blit_to_buffer(_async)
 * do this by themselves. */
int swrh_frame_exposure(void *renderer, float exposure);
int swrh_frame_hdr(void *renderer);
/* the metering maths of update_auto_exposure on its own (no device): state = {auto_exposure, auto_exposure_target,
 * auto_exposure_ev}, updated in place from `ntiles` per-tile center_luminance values (tilerasterizer.rs:103-106) */
int swrh_auto_exposure_step(float state[3], const float *tile_luminance, int ntiles, float delta_time);
/* Renderer::blit_to_buffer(&mut RenderBuffer) (renderer.rs:293): W*H u32, row-major, (R<<24)|(G<<16)|(B<<8)|A */
int swrh_blit_to_buffer(void *renderer, uint32_t *pixels, size_t width, size_t height);
/* pipelined form: resolve + read-back into pinned `pixels` in the background; swrh_wait_blit(ticket) completes it */
int swrh_blit_to_buffer_async(void *renderer, uint32_t *pixels, size_t width, size_t height, int *ticket);
int swrh_wait_blit(void *renderer, int ticket);
/* sort-first: this renderer owns tile rows [row_begin, row_end) */
int swrh_set_tile_rows(void *renderer, int row_begin, int row_end);
/* normalize() as the reference computes it on this host (_mm_rsqrt_ps table, default on) or rsqrtf() */
int swrh_set_reference_rsqrt(void *renderer, int on);
int swrh_reference_rsqrt_bits(void *renderer);
int swrh_num_draws(void *renderer);

/* The host draw list on its own (renderer.rs:357-468), no device involved. Returns the number of draws (call with
 * out = NULL, max_draws = 0 to size the array), -1 on error. */
int swrh_build_draws(const swr_scene_desc *scene, const swr_camera *camera, swr_draw *out, int max_draws, int shard, int nshards);
int swrh_build_draws_band(const swr_scene_desc *scene, const swr_camera *camera, swr_draw *out, int max_draws, int y0, int y1, int height);

#ifdef __cplusplus
}
#endif
#endif

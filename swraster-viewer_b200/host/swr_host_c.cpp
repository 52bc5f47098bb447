// C exports of the host mirror (swr_host.hpp) so Python tests/bench can drive the
// same Renderer / RenderCamera code a C++ application would. Links libswr_b200.so.
#include "swr_host.hpp"

static thread_local std::string g_err;

#define SWRH_TRY(stmt)                \
    try {                             \
        stmt;                         \
        return 0;                     \
    } catch (const std::exception &e) { \
        g_err = e.what();             \
        return -1;                    \
    }

extern "C" {

const char *swrh_last_error(void) { return g_err.c_str(); }

// RenderCamera::new + update_matrices (rendercamera.rs:28-86)
int swrh_camera_build(const float pos[3], const float look_at[3], float fov, float width, float height, float far_plane,
                      swr_camera *out) {
    SWRH_TRY(*out = swr::RenderCamera(pos, look_at, fov, width, height, far_plane).to_abi());
}

// RenderCamera::new towards a level target, then rotate_mouse(dx, dy) and update_matrices (main.rs:478-500 path).
int swrh_camera_build_rotated(const float pos[3], const float look_at_level[3], float mouse_dx, float mouse_dy, float fov, float width,
                              float height, float far_plane, swr_camera *out) {
    SWRH_TRY(swr::RenderCamera c(pos, look_at_level, fov, width, height, far_plane); c.rotate_mouse(mouse_dx, mouse_dy);
             c.update_matrices(); *out = c.to_abi());
}

void *swrh_renderer_new(int width, int height, int device) {
    try {
        return new swr::Renderer(width, height, device);
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void swrh_renderer_free(void *r) { delete (swr::Renderer *)r; }
void *swrh_renderer_ctx(void *r) { return ((swr::Renderer *)r)->ctx(); }

int swrh_set_reference_rsqrt(void *r, int on) { SWRH_TRY(((swr::Renderer *)r)->set_reference_rsqrt(on != 0)); }
int swrh_reference_rsqrt_bits(void *r) { return ((swr::Renderer *)r)->reference_rsqrt_bits(); }
int swrh_set_tile_rows(void *r, int r0, int r1) { SWRH_TRY(((swr::Renderer *)r)->set_tile_rows(r0, r1)); }

int swrh_render_scene(void *r, const swr_scene_desc *scene, const swr_camera *cam, int shade, int shard, int nshards) {
    SWRH_TRY(((swr::Renderer *)r)->render_scene(swr::Scene(scene), *cam, shade != 0, shard, nshards));
}

int swrh_build_draws_band(const swr_scene_desc *scene, const swr_camera *cam, swr_draw *out, int max_draws, int y0, int y1, int height) {
    try {
        std::vector<swr_draw> draws;
        swr::build_draw_list(*scene, *cam, draws, 0, 1, y0, y1, height);
        for (size_t i = 0; i < draws.size() && (int)i < max_draws; i++) out[i] = draws[i];
        return (int)draws.size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
int swrh_num_draws(void *r) { return (int)((swr::Renderer *)r)->draws().size(); }
int swrh_update_auto_exposure(void *r, float dt) { SWRH_TRY(((swr::Renderer *)r)->update_auto_exposure(dt)); }
float swrh_auto_exposure(void *r) { return ((swr::Renderer *)r)->auto_exposure(); }

int swrh_blit_to_buffer(void *r, uint32_t *pixels, size_t width, size_t height) {
    SWRH_TRY(swr::RenderBuffer buf(width, height, pixels); ((swr::Renderer *)r)->blit_to_buffer(buf));
}

int swrh_blit_to_buffer_async(void *r, uint32_t *pixels, size_t width, size_t height, int *ticket) {
    SWRH_TRY(swr::RenderBuffer buf(width, height, pixels); *ticket = ((swr::Renderer *)r)->blit_to_buffer_async(buf));
}
int swrh_wait_blit(void *r, int ticket) { SWRH_TRY(((swr::Renderer *)r)->wait_blit(ticket)); }

// Host draw list only (no device work): for cross-checks against the oracle's own R1/R2.
int swrh_build_draws(const swr_scene_desc *scene, const swr_camera *cam, swr_draw *out, int max_draws, int shard, int nshards) {
    try {
        std::vector<swr_draw> draws;
        swr::build_draw_list(*scene, *cam, draws, shard, nshards);
        for (size_t i = 0; i < draws.size() && (int)i < max_draws; i++) out[i] = draws[i];
        return (int)draws.size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
}

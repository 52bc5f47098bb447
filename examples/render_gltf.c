/* render_gltf.c — the whole boundary from plain C99: load a glTF / GLB file, render one frame on a B200, write a PPM.
 *
 *   cc -std=c99 -I include examples/render_gltf.c -L swraster-viewer_b200/lib -lswr_host -lswr_b200 \
 *      -Wl,-rpath,'$ORIGIN/../swraster-viewer_b200/lib' -lm -o render_gltf
 *   ./render_gltf scene.glb out.ppm [width height [sky_cross.png]]
 *
 * What a glTF file does not carry (sky, prefiltered sky, BRDF LUT, GI voxels) is filled with small procedural
 * stand-ins here; a real host bakes them the way the reference does (texture.rs:135-552, gi.rs). There is no CPU
 * fallback: without a usable sm_100 device swrh_renderer_new fails and the program says why. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "swr_gltf.h"
#include "swr_host.h"

static uint32_t pack(float r, float g, float b) {
    uint32_t R = (uint32_t)(r * 255.0f), G = (uint32_t)(g * 255.0f), B = (uint32_t)(b * 255.0f);
    return (R << 24) | (G << 16) | (B << 8) | 0xFFu;
}

/* single-mip texture descriptor over caller-owned tables */
typedef struct simple_tex {
    uint32_t *texels;
    uint32_t mip_offset, mip_width, mip_height, stride;
    swr_texture_desc desc;
} simple_tex;

static void make_tex(simple_tex *t, uint32_t w, uint32_t h, uint32_t faces, uint32_t type) {
    t->texels = (uint32_t *)malloc(sizeof(uint32_t) * w * h * faces);
    t->mip_offset = 0;
    t->mip_width = w;
    t->mip_height = h;
    t->stride = faces > 1 ? w * h : 0;
    memset(&t->desc, 0, sizeof(t->desc));
    t->desc.data = t->texels;
    t->desc.ntexels = w * h * faces;
    t->desc.width = w;
    t->desc.height = h;
    t->desc.texture_type = type;
    t->desc.max_mip_level = 0;
    t->desc.mip_offsets = &t->mip_offset;
    t->desc.mip_widths = &t->mip_width;
    t->desc.mip_heights = &t->mip_height;
    t->desc.array_stride = &t->stride;
    t->desc.wrap_s = t->desc.wrap_t = SWR_WRAP_CLAMP_TO_EDGE;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s scene.gltf|scene.glb out.ppm [width height [sky_cross.png]]\n", argv[0]);
        return 2;
    }
    const int W = argc > 4 ? atoi(argv[3]) : 1280, H = argc > 4 ? atoi(argv[4]) : 720;

    /* Environment. With a sky image in the viewer's cross layout (argv[5], PNG) everything is baked on the host the way the
     * reference does at load time (swrh_env_bake); otherwise small procedural stand-ins: a sky gradient on all six faces, a
     * flat BRDF LUT, one voxel of dim ambient light. */
    simple_tex sky, spec, lut;
    float voxel[16];
    void *baked = NULL;
    swrh_gltf_env env;
    memset(&env, 0, sizeof(env));
    memset(&sky, 0, sizeof(sky)), memset(&spec, 0, sizeof(spec)), memset(&lut, 0, sizeof(lut));
    if (argc > 5) {
        FILE *pf = fopen(argv[5], "rb");
        if (!pf) {
            perror(argv[5]);
            return 1;
        }
        fseek(pf, 0, SEEK_END);
        long n = ftell(pf);
        fseek(pf, 0, SEEK_SET);
        uint8_t *bytes = (uint8_t *)malloc((size_t)n);
        if (fread(bytes, 1, (size_t)n, pf) != (size_t)n) return 1;
        fclose(pf);
        uint32_t cw = 0, ch = 0;
        if (swrh_decode_png(bytes, (size_t)n, NULL, &cw, &ch)) {
            fprintf(stderr, "sky image: %s\n", swrh_last_error());
            return 1;
        }
        uint8_t *rgba = (uint8_t *)malloc((size_t)cw * ch * 4);
        swrh_decode_png(bytes, (size_t)n, rgba, &cw, &ch);
        baked = swrh_env_bake(rgba, cw, ch, 128, 64, 16, 0.25f, 1.0f, 1.0f); /* texture.rs:216,225 and main.rs:54,280 */
        free(rgba), free(bytes);
        if (!baked || swrh_env_get(baked, &env, NULL)) {
            fprintf(stderr, "environment bake: %s\n", swrh_last_error());
            return 1;
        }
    } else {
        make_tex(&sky, 16, 16, 6, SWR_TEX_CUBEMAP);
        make_tex(&spec, 16, 16, 6, SWR_TEX_CUBEMAP);
        make_tex(&lut, 16, 16, 1, SWR_TEX_LINEAR);
        for (uint32_t f = 0; f < 6; f++)
            for (uint32_t y = 0; y < 16; y++)
                for (uint32_t x = 0; x < 16; x++) {
                    float t = (float)y / 15.0f;
                    uint32_t c = f == 2 ? pack(0.35f, 0.55f, 0.9f) : f == 3 ? pack(0.25f, 0.22f, 0.2f) : pack(0.35f + 0.3f * t, 0.55f + 0.15f * t, 0.9f - 0.2f * t);
                    sky.texels[(f * 16 + y) * 16 + x] = c;
                    spec.texels[(f * 16 + y) * 16 + x] = c;
                }
        for (uint32_t i = 0; i < 256; i++) lut.texels[i] = pack(0.5f, 0.1f, 0.0f);
        memset(voxel, 0, sizeof(voxel));
        voxel[0] = voxel[1] = voxel[2] = 0.4f; /* SH band 0, rgb */
        voxel[3] = 1.0f;                       /* sun visibility */
        env.cubemap = &sky.desc;
        env.cubemap_specular = &spec.desc;
        env.brdf_lut = &lut.desc;
        env.voxel_grid.dims[0] = env.voxel_grid.dims[1] = env.voxel_grid.dims[2] = 1;
        env.voxel_grid.gi_sh4 = voxel;
    }
    { /* scene.rs:236-239 */
        const float d[3] = {-0.2f, 1.0f, 0.5f}, n = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        for (int c = 0; c < 3; c++) env.light_direction[c] = d[c] * n;
        env.light_color[0] = 5.0f, env.light_color[1] = 5.0f, env.light_color[2] = 4.75f;
    }

    void *doc = swrh_gltf_load(argv[1], &env);
    if (!doc) {
        fprintf(stderr, "load failed: %s\n", swrh_last_error());
        return 1;
    }
    swrh_gltf_info info;
    swrh_gltf_get_info(doc, &info);
    if (baked && swrh_gltf_bake_sun_visibility(doc)) { /* main.rs:237-246: one ray per voxel near geometry towards the sun */
        fprintf(stderr, "sun visibility: %s\n", swrh_last_error());
        return 1;
    }
    /* env.voxel_grid carried no extent, so the loader made the grid span the scene bounds (main.rs:228-235) */
    swr_scene_desc scene = *swrh_gltf_scene(doc);
    printf("%s: %u primitives, %u nodes, %u materials, %u textures, bounds diagonal %.3f\n", argv[1], scene.nprimitives, scene.nnodes, scene.nmaterials,
           scene.ntextures, info.bounds_diagonal);

    /* the reference's default camera (main.rs:210-224): eye = centre + (0, 0, diagonal), fov pi/4, far = 2 diagonal */
    float eye[3] = {info.bounds_center[0], info.bounds_center[1], info.bounds_center[2] + info.bounds_diagonal};
    swr_camera cam;
    if (swrh_camera_build(eye, info.bounds_center, 0.78539816f, (float)W, (float)H, 2.0f * info.bounds_diagonal, &cam)) {
        fprintf(stderr, "camera: %s\n", swrh_last_error());
        return 1;
    }

    void *r = swrh_renderer_new(W, H, 0);
    if (!r) {
        fprintf(stderr, "Renderer::new failed (there is no CPU fallback): %s\n", swrh_last_error());
        return 3;
    }
    uint32_t *pixels = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)W * H);
    if (swrh_render_scene(r, &scene, &cam, 1, 0, 1) || swrh_update_auto_exposure(r, 0.0f) || swrh_blit_to_buffer(r, pixels, (size_t)W, (size_t)H)) {
        fprintf(stderr, "render failed: %s\n", swrh_last_error());
        return 1;
    }
    swr_frame_stats st;
    swr_get_stats(swrh_renderer_ctx(r), &st);
    printf("%llu triangles submitted, %llu binned, %llu tile refs; set-up+bin %.3f ms, raster %.3f ms, shade %.3f ms\n", (unsigned long long)st.triangles_submitted,
           (unsigned long long)st.triangles_binned, (unsigned long long)st.tile_refs, st.ms_setup_bin, st.ms_raster, st.ms_shade);

    FILE *f = fopen(argv[2], "wb");
    if (!f) {
        perror(argv[2]);
        return 1;
    }
    fprintf(f, "P6\n%d %d\n255\n", W, H);
    for (size_t i = 0; i < (size_t)W * H; i++) {
        unsigned char rgb[3] = {(unsigned char)(pixels[i] >> 24), (unsigned char)(pixels[i] >> 16), (unsigned char)(pixels[i] >> 8)};
        fwrite(rgb, 1, 3, f);
    }
    fclose(f);
    free(pixels);
    swrh_renderer_free(r);
    swrh_gltf_free(doc);
    free(sky.texels), free(spec.texels), free(lut.texels);
    swrh_env_free(baked);
    return 0;
}

// swr_api.cu — implementation of the C ABI in include/swr.h on CUDA (sm_100a).
// Frame = k_setup -> k_scan_tiles -> k_scatter -> k_raster_tiles -> k_shade -> k_luminance, all on one
// stream, no host synchronisation inside swr_render. Buffer growth (tile refs, clip vertices) is detected
// from device counters at the next synchronising call and the frame is replayed once.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include "../../include/swr.h"
#include "swr_shade.cuh"
#include "swr_bake.cuh"

static thread_local std::string g_create_error;

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                          \
            return SWR_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct Staging {  // pinned host mirror of the per-frame draw table, rotated so a pending copy is never overwritten
    void *host = nullptr;
    size_t bytes = 0;
    cudaEvent_t done = nullptr;
};

// Everything one geometry pass (set-up, clip, bin) owns. Two instances: the opaque primitives (visibility buffer) and the
// translucent ones (forward-shaded afterwards, tilerasterizer.rs:92-101).
struct GeomSet {
    DevBuf<DevDraw> draws;
    DevBuf<uint32_t> tri_prefix;  // ndraws + 1 triangle prefix, then ndraws + 1 cluster prefix
    // second copy of both tables: frame N+1's tables are uploaded (on the upload stream) while frame N's kernels still read theirs
    DevBuf<DevDraw> draws_alt;
    DevBuf<uint32_t> tri_prefix_alt;
    DevBuf<uint32_t> cull;  // k_cull output: keep mask, per-block counts and offsets, draw of every cluster
    DevBuf<uint4> work;     // k_compact output: surviving clusters in submission order (draw, cluster, dense base, triangles)
    DevBuf<TriRecord> records;
    DevBuf<uint32_t> rects;
    DevBuf<uint8_t> zb;  // depth bucket per record (tile lists are bucket-major)
    DevBuf<float> avgz;  // translucent set only: packet.avg_z per record (renderer.rs:765-775)
    DevBuf<ClipVertex> clip_verts;
    DevBuf<uint2> clip_queue;  // (dense triangle, draw) pairs waiting for k_clip
    DevBuf<uint32_t> clip_ext, clip_list;
    size_t ext_cap = 0;
    DevBuf<uint32_t> tile_count, tile_offset, tile_cursor;
    DevBuf<uint32_t> refs;
    DevBuf<FrameCounters> counters;
    FrameCounters *h_counters = nullptr;  // where this frame's counters are copied: the current FrameSlot's pinned block (not owned)
    Staging staging[4];
    int staging_next = 0;
    std::vector<swr_draw> last_draws;
    uint32_t ndraws = 0, total_tris = 0, clusters = 0;
    unsigned geom_grid = 1;
    uint32_t work_hint = 0;  // clusters that survived k_cull in the last finished frame: sizes the next set-up / scatter grids
    uint64_t total_verts = 0;
    uint64_t refs_emitted = 0;
    bool rendered_once = false;
};

// Two frames may be in flight on the stream: the host enqueues frame N+1 while frame N still runs, and a frame's device
// counters (buffer overflow flags, statistics) are only looked at when somebody needs that frame — the next-but-one
// swr_render, swr_wait_pixels, or any synchronising call. A slot keeps what a replay needs if a buffer turned out too small.
#define SWR_FRAMES_IN_FLIGHT 2
struct FrameSlot {
    bool pending = false;
    std::vector<swr_draw> op_draws, tr_draws;
    swr_camera cam{};
    int shade = 0;
    float fixed_exposure = 0.0f;  // > 0: shade straight to RGBA8 with this exposure when the frame allows it (swr_set_fixed_exposure)
    bool tr_ran = false;      // the translucent geometry pass was launched (its counters are valid)
    int resolve_kind = 0;     // queued behind the frame: 0 nothing, 1 device-only resolve, 2 resolve + read-back (swr_resolve_async),
                              // 3 peer resolve (swr_resolve_peer, frame number in resolve_frame)
    uint32_t resolve_frame = 0;
    int resolve_idx = 0;
    float resolve_exposure = 0.0f;
    uint32_t *resolve_host = nullptr;
    FrameCounters *h_op = nullptr, *h_tr = nullptr;  // pinned
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // phase timing: begin, geometry done, raster done, shade done
    cudaEvent_t done = nullptr;                                // both counter blocks are on the host
    uint32_t total_tris = 0, clusters = 0;
    uint64_t total_verts = 0;
};

struct swr_ctx {
    int device = 0;
    int num_sms = 148;
    int W = 0, H = 0, tiles_x = 0, tiles_y = 0, ntiles = 0;
    int row_begin = 0, row_end = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t upload_stream = nullptr;  // per-frame draw tables: issued at enqueue time, ahead of the frame that needs them
    FrameSlot slots[SWR_FRAMES_IN_FLIGHT];
    int slot_head = 0, slots_pending = 0;  // ring: oldest pending slot, number pending
    FrameSlot *cur = nullptr;              // slot of the frame being (or last) enqueued
    cudaEvent_t ev_res[2] = {nullptr, nullptr};
    std::string err;
    uint64_t launches = 0;  // kernels this context has launched (swr_launch_count)

    // scene
    bool have_scene = false;
    bool scene_borrowed = false;  // swr_share_scene: the device scene belongs to another context of the same device
    std::vector<void *> scene_allocs;
    std::vector<cudaTextureObject_t> tex_objs;
    DevScene scene{};
    std::vector<uint32_t> prim_ntris, prim_nverts;

    // frame: geometry sets (opaque pass, translucent pass) + shared per-frame buffers
    GeomSet op, tr;
    DevBuf<uint32_t> unit_list, tile_cycles, tile_cycles_prev, tile_count_prev, tile_unit;
    bool have_history = false;
    DevBuf<unsigned long long> keys;
    DevBuf<float4> color;
    DevBuf<uint32_t> pixels;
    // asynchronous read-back (swr_resolve_async): second pixel buffer, copy stream, per-buffer events
    DevBuf<uint32_t> pixels_alt;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_resolved[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    bool copy_pending[2] = {false, false};
    int pix_cur = 0;
    DevBuf<float> lum;
    // fixed-exposure frames: k_shade<true> packs RGBA8 itself into `fused_px`; the resolve entry points then only move it
    float fixed_exposure = 0.0f;   // swr_set_fixed_exposure (applies to frames rendered from now on)
    float units_floor = 1.0f;      // least raster units per CTA slot (development knob: SWR_UNITS_FLOOR)
    DevBuf<uint32_t> fused_px;
    bool last_fused = false;       // the frame shaded last holds RGBA8 in fused_px (exposure last_fused_exposure) and NO HDR colour
    float last_fused_exposure = 0.0f;
    DevBuf<float2> bary;
    bool composited = false;
    int sky_r0 = 0, sky_r1 = 0;
    DevBuf<unsigned long long> dbg_tiles;
    DevBuf<uint32_t> rsqrt_tab;
    int rsqrt_bits = 0;
    bool rsqrt_on = false;
    swr_camera last_cam{};
    // peer frame assembly (swr_peer_*): mapping of the assembling rank's pixel buffer + local bookkeeping
    uint32_t *peer_pixels = nullptr;   // assembler's pixels as seen from this context (own buffer on the assembler)
    void *peer_ipc_base = nullptr;     // non-null when opened through an IPC handle (closed on destroy)
    DevBuf<uint32_t> peer_local;       // [0] CTAs done (k_resolve_peer), [1] timeout flag
    bool peer_exported = false;
    bool frame_valid = false;
    DevCamera dcam{};
    swr_frame_stats stats{};
};

template <typename T>
static int upload(swr_ctx *ctx, const T *src, size_t n, T **out) {
    *out = nullptr;
    if (n == 0) return SWR_OK;
    void *d = nullptr;
    CK(cudaMalloc(&d, n * sizeof(T)));
    ctx->scene_allocs.push_back(d);
    CK(cudaMemcpyAsync(d, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *out = (T *)d;
    return SWR_OK;
}

static size_t raster_smem_bytes() { return SWR_TILE_PIXELS * 8 + sizeof(TileBatch) + RASTER_WARPS * sizeof(FragQueue); }

extern "C" {

int swr_abi_version(void) { return SWR_ABI_VERSION; }

size_t swr_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(swr_primitive_desc);
        case 1: return sizeof(swr_mesh_desc);
        case 2: return sizeof(swr_node_desc);
        case 3: return sizeof(swr_texture_desc);
        case 4: return sizeof(swr_material_desc);
        case 5: return sizeof(swr_voxel_grid_desc);
        case 6: return sizeof(swr_scene_desc);
        case 7: return sizeof(swr_camera);
        case 8: return sizeof(swr_draw);
        case 9: return sizeof(swr_frame_stats);
        default: return 0;
    }
}

const char *swr_last_error(const swr_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

swr_ctx *swr_create(int width, int height, int device) {
    if (width <= 0 || height <= 0 || (width & 1) || (height & 1)) {
        g_create_error = "width and height must be positive and even (2x2 quads; renderer.rs:296 processes row pairs)";
        return nullptr;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        g_create_error = std::string("no usable CUDA device (there is no CPU fallback): ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range");
        return nullptr;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
        g_create_error = "device is not sm_100 class: this library carries sm_100a code only";
        return nullptr;
    }
    swr_ctx *ctx = new swr_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if (const char *e = getenv("SWR_UNITS_FLOOR")) ctx->units_floor = (float)atof(e) > 0.0f ? (float)atof(e) : 1.0f;  // development knob
    ctx->W = width;
    ctx->H = height;
    ctx->tiles_x = (width + SWR_TILE - 1) / SWR_TILE;
    ctx->tiles_y = (height + SWR_TILE - 1) / SWR_TILE;
    ctx->ntiles = ctx->tiles_x * ctx->tiles_y;
    ctx->row_begin = 0;
    ctx->row_end = ctx->tiles_y;
    if (ctx->tiles_x > 255 || ctx->tiles_y > 255) {
        g_create_error = "resolution above 16320 pixels per axis is not supported (tile rectangles are packed in 8 bits)";
        delete ctx;
        return nullptr;
    }
    bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (FrameSlot &f : ctx->slots) {
        for (int i = 0; ok && i < 4; i++) ok = cudaEventCreate(&f.ev[i]) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&f.done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaMallocHost(&f.h_op, sizeof(FrameCounters)) == cudaSuccess && cudaMallocHost(&f.h_tr, sizeof(FrameCounters)) == cudaSuccess;
    }
    ctx->cur = &ctx->slots[0];
    for (int i = 0; ok && i < 2; i++) ok = cudaEventCreate(&ctx->ev_res[i]) == cudaSuccess;
    for (int i = 0; ok && i < 4; i++) ok = cudaEventCreateWithFlags(&ctx->op.staging[i].done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && ctx->op.tile_count.reserve((size_t)(ctx->ntiles + 1) * SWR_ZBUCKETS) == cudaSuccess && ctx->op.tile_offset.reserve(ctx->ntiles + 1) == cudaSuccess &&
         ctx->op.tile_cursor.reserve((size_t)(ctx->ntiles + 1) * SWR_ZBUCKETS) == cudaSuccess && ctx->tile_cycles.reserve(ctx->ntiles + 1) == cudaSuccess && ctx->tile_cycles_prev.reserve(ctx->ntiles + 1) == cudaSuccess &&
         ctx->tile_count_prev.reserve(ctx->ntiles + 1) == cudaSuccess && ctx->tile_unit.reserve(ctx->ntiles + 1) == cudaSuccess && ctx->unit_list.reserve(ctx->ntiles + 1) == cudaSuccess && ctx->keys.reserve((size_t)ctx->ntiles * SWR_TILE_PIXELS) == cudaSuccess &&
         ctx->color.reserve((size_t)ctx->ntiles * SWR_TILE_PIXELS) == cudaSuccess && ctx->pixels.reserve((size_t)width * height + SWR_PEER_WORDS) == cudaSuccess && ctx->peer_local.reserve(16) == cudaSuccess &&
         ctx->lum.reserve(ctx->ntiles) == cudaSuccess && ctx->op.counters.reserve(1) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_raster_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)raster_smem_bytes()) == cudaSuccess;
    if (ok) {
        ok = cudaMemsetAsync(ctx->keys.p, 0xFF, (size_t)ctx->ntiles * SWR_TILE_PIXELS * 8, ctx->stream) == cudaSuccess &&
             cudaMemsetAsync(ctx->color.p, 0, (size_t)ctx->ntiles * SWR_TILE_PIXELS * sizeof(float4), ctx->stream) == cudaSuccess &&
             cudaMemsetAsync(ctx->pixels.p, 0, ((size_t)width * height + SWR_PEER_WORDS) * 4, ctx->stream) == cudaSuccess &&
             cudaMemsetAsync(ctx->peer_local.p, 0, 16 * 4, ctx->stream) == cudaSuccess &&
             cudaMemsetAsync(ctx->lum.p, 0, ctx->ntiles * sizeof(float), ctx->stream) == cudaSuccess &&
             cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    }
    if (!ok) {
        g_create_error = std::string("CUDA initialisation failed: ") + cudaGetErrorString(cudaGetLastError());
        swr_destroy(ctx);
        return nullptr;
    }
    return ctx;
}

static void free_scene(swr_ctx *ctx) {
    if (ctx->scene_borrowed) {  // nothing here is ours
        ctx->tex_objs.clear();
        ctx->scene_allocs.clear();
        ctx->scene_borrowed = false;
        ctx->have_scene = false;
        return;
    }
    for (cudaTextureObject_t t : ctx->tex_objs) cudaDestroyTextureObject(t);
    ctx->tex_objs.clear();
    for (void *p : ctx->scene_allocs) cudaFree(p);
    ctx->scene_allocs.clear();
    ctx->have_scene = false;
}

void swr_destroy(swr_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_scene(ctx);
    for (GeomSet *g : {&ctx->op, &ctx->tr}) {
        g->draws.release();
        g->tri_prefix.release();
        g->draws_alt.release();
        g->tri_prefix_alt.release();
        g->cull.release();
        g->work.release();
        g->records.release();
        g->rects.release();
        g->zb.release();
        g->avgz.release();
        g->clip_verts.release();
        g->clip_queue.release();
        g->clip_ext.release();
        g->clip_list.release();
        g->tile_count.release();
        g->tile_offset.release();
        g->tile_cursor.release();
        g->refs.release();
        g->counters.release();
        for (auto &st : g->staging) {
            if (st.host) cudaFreeHost(st.host);
            if (st.done) cudaEventDestroy(st.done);
        }
    }
    if (ctx->peer_ipc_base) cudaIpcCloseMemHandle(ctx->peer_ipc_base);
    ctx->peer_local.release();
    ctx->unit_list.release();
    ctx->tile_cycles.release();
    ctx->tile_cycles_prev.release();
    ctx->tile_count_prev.release();
    ctx->tile_unit.release();
    ctx->keys.release();
    ctx->color.release();
    ctx->pixels.release();
    ctx->pixels_alt.release();
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_resolved[i]) cudaEventDestroy(ctx->ev_resolved[i]);
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    ctx->lum.release();
    ctx->bary.release();
    ctx->rsqrt_tab.release();
    ctx->dbg_tiles.release();
    ctx->fused_px.release();
    for (FrameSlot &f : ctx->slots) {
        for (auto &e : f.ev)
            if (e) cudaEventDestroy(e);
        if (f.done) cudaEventDestroy(f.done);
        if (f.h_op) cudaFreeHost(f.h_op);
        if (f.h_tr) cudaFreeHost(f.h_tr);
    }
    for (auto &e : ctx->ev_res)
        if (e) cudaEventDestroy(e);
    if (ctx->upload_stream) cudaStreamDestroy(ctx->upload_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int swr_set_rsqrt_table(swr_ctx *ctx, const uint32_t *table, int mantissa_bits) {
    if (!ctx) return SWR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int rc = SWR_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    if (!table) {
        ctx->rsqrt_on = false;
        return rc;
    }
    if (mantissa_bits < 1 || mantissa_bits > 16) {
        ctx->err = "rsqrt table: mantissa_bits must be in 1..16";
        return SWR_ERR_INVALID;
    }
    size_t n = (size_t)2 << mantissa_bits;
    if (ctx->rsqrt_tab.reserve(n) != cudaSuccess) return SWR_ERR_OOM;
    CK(cudaMemcpy(ctx->rsqrt_tab.p, table, n * 4, cudaMemcpyHostToDevice));
    ctx->rsqrt_bits = mantissa_bits;
    ctx->rsqrt_on = true;
    return rc;
}

int swr_set_tile_rows(swr_ctx *ctx, int row_begin, int row_end) {
    if (!ctx) return SWR_ERR_INVALID;
    if (row_begin < 0 || row_end > ctx->tiles_y || row_begin > row_end) {
        ctx->err = "tile row range out of bounds";
        return SWR_ERR_INVALID;
    }
    ctx->row_begin = row_begin;
    ctx->row_end = row_end;
    ctx->have_history = false;
    return SWR_OK;
}

int swr_upload_scene(swr_ctx *ctx, const swr_scene_desc *s) {
    if (!ctx || !s) return SWR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    free_scene(ctx);
    ctx->frame_valid = false;
    ctx->have_history = false;
    if (s->nmaterials == 0 || s->cubemap < 0 || s->cubemap_specular < 0 || s->brdf_lut < 0 || (uint32_t)s->cubemap >= s->ntextures ||
        (uint32_t)s->cubemap_specular >= s->ntextures || (uint32_t)s->brdf_lut >= s->ntextures) {
        ctx->err = "scene needs at least one material (scene.rs:493) and cubemap / cubemap_specular / brdf_lut textures";
        return SWR_ERR_INVALID;
    }
    int rc;
    std::vector<DevPrim> prims(s->nprimitives);
    ctx->prim_ntris.assign(s->nprimitives, 0);
    ctx->prim_nverts.assign(s->nprimitives, 0);
    for (uint32_t i = 0; i < s->nprimitives; i++) {
        const swr_primitive_desc &p = s->primitives[i];
        if (p.material_index >= s->nmaterials) {
            ctx->err = "primitive material_index out of range";
            return SWR_ERR_INVALID;
        }
        for (uint32_t k = 0; k < p.nindices; k++)
            if (p.indices[k] >= p.nverts) {
                ctx->err = "primitive index out of range";  // the reference would panic on the slice access
                return SWR_ERR_INVALID;
            }
        DevPrim d{};
        float4 *pos, *nrm, *tan;
        float2 *uv;
        uint32_t *idx;
        if ((rc = upload(ctx, (const float4 *)p.positions, p.nverts, &pos))) return rc;
        if ((rc = upload(ctx, (const float4 *)p.normals, p.nverts, &nrm))) return rc;
        if ((rc = upload(ctx, (const float4 *)p.tangents, p.nverts, &tan))) return rc;
        if ((rc = upload(ctx, (const float2 *)p.texcoords, p.nverts, &uv))) return rc;
        if ((rc = upload(ctx, p.indices, p.nindices, &idx))) return rc;
        d.pos = pos;
        d.nrm = nrm;
        d.tan = tan;
        d.uv = uv;
        d.idx = idx;
        d.nverts = p.nverts;
        d.ntris = p.nindices / 3;
        d.material = p.material_index;
        d.ncl = (d.ntris + SWR_CLUSTER_TRIS - 1) / SWR_CLUSTER_TRIS;
        if (d.ncl) {
            float4 *sph = nullptr;
            CK(cudaMalloc(&sph, (size_t)d.ncl * sizeof(float4)));
            ctx->scene_allocs.push_back(sph);
            ctx->launches++;
            k_cluster_bounds<<<(d.ncl + 7) / 8, 256, 0, ctx->stream>>>(pos, idx, d.ntris, sph, d.ncl);
            d.cl_sphere = sph;
        }
        prims[i] = d;
        ctx->prim_ntris[i] = d.ntris;
        ctx->prim_nverts[i] = p.nverts;
    }
    std::vector<DevMat> mats(s->nmaterials);
    for (uint32_t i = 0; i < s->nmaterials; i++) {
        const swr_material_desc &m = s->materials[i];
        DevMat d{};
        memcpy(d.base, m.base_color_factor, 16);
        d.metallic = m.metallic_factor;
        d.roughness = m.roughness_factor;
        memcpy(d.emissive, m.emissive_factor, 12);
        d.occlusion_strength = m.occlusion_strength;
        d.transmission = m.transmission;
        d.alpha_cutoff = m.alpha_cutoff;
        d.flags = m.flags;
        const int32_t t[6] = {m.base_color_texture, m.metallic_roughness_texture, m.normal_texture, m.emissive_texture, m.occlusion_texture, m.transmission_texture};
        for (int k = 0; k < 6; k++)
            if (t[k] >= (int32_t)s->ntextures) {
                ctx->err = "material texture index out of range";
                return SWR_ERR_INVALID;
            }
        d.tex_base = t[0];
        d.tex_mr = t[1];
        d.tex_normal = t[2];
        d.tex_emissive = t[3];
        d.tex_occlusion = t[4];
        d.tex_transmission = t[5];
        mats[i] = d;
    }
    std::vector<DevTex> texs(s->ntextures);
    for (uint32_t i = 0; i < s->ntextures; i++) {
        const swr_texture_desc &t = s->textures[i];
        if (t.max_mip_level + 1 > SWR_MAX_MIPS) {
            ctx->err = "texture has more than 16 mip levels";
            return SWR_ERR_INVALID;
        }
        if (!t.data || !t.mip_offsets || !t.mip_widths || !t.mip_heights || !t.array_stride || t.width == 0 || t.height == 0 ||
            t.texture_type > SWR_TEX_LINEAR || t.wrap_s > SWR_WRAP_CLAMP_TO_EDGE || t.wrap_t > SWR_WRAP_CLAMP_TO_EDGE) {
            ctx->err = "texture " + std::to_string(i) + ": null table, empty image or unknown type / wrap mode";
            return SWR_ERR_INVALID;
        }
        for (uint32_t k = 0; k <= t.max_mip_level; k++) {  // every mip (all six faces of a cubemap) must lie inside the texel array
            // the sky and the prefiltered sky are sampled as six faces whatever their declared type
            const uint64_t faces = (t.texture_type == SWR_TEX_CUBEMAP || (int32_t)i == s->cubemap || (int32_t)i == s->cubemap_specular) ? 6 : 1;
            const uint64_t wh = (uint64_t)t.mip_widths[k] * t.mip_heights[k];
            const uint64_t last = (uint64_t)t.mip_offsets[k] + (faces - 1) * t.array_stride[k] + wh;
            if (wh == 0 || last > t.ntexels || (faces > 1 && t.array_stride[k] < wh)) {
                ctx->err = "texture " + std::to_string(i) + ": mip " + std::to_string(k) + " reaches outside the texel array";
                return SWR_ERR_INVALID;
            }
        }
        DevTex d{};
        uint32_t *data;
        if ((rc = upload(ctx, t.data, t.ntexels, &data))) return rc;
        d.data = data;
        d.width = t.width;
        d.height = t.height;
        d.type = t.texture_type;
        d.max_mip = t.max_mip_level;
        d.wrap_s = t.wrap_s;
        d.wrap_t = t.wrap_t;
        for (uint32_t k = 0; k <= t.max_mip_level; k++) {
            d.mip_off[k] = t.mip_offsets[k];
            d.mip_w[k] = t.mip_widths[k];
            d.mip_h[k] = t.mip_heights[k];
            d.stride[k] = t.array_stride[k];
        }
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = data;
        rd.res.linear.desc = cudaCreateChannelDesc<unsigned int>();
        rd.res.linear.sizeInBytes = (size_t)t.ntexels * 4;
        cudaTextureDesc td{};
        td.readMode = cudaReadModeElementType;
        td.filterMode = cudaFilterModePoint;
        td.addressMode[0] = cudaAddressModeClamp;
        cudaTextureObject_t obj = 0;
        CK(cudaCreateTextureObject(&obj, &rd, &td, nullptr));
        ctx->tex_objs.push_back(obj);
        d.obj = obj;
        texs[i] = d;
    }
    DevPrim *dprims;
    DevMat *dmats;
    DevTex *dtexs;
    float4 *gi;
    if ((rc = upload(ctx, prims.data(), prims.size(), &dprims))) return rc;
    if ((rc = upload(ctx, mats.data(), mats.size(), &dmats))) return rc;
    if ((rc = upload(ctx, texs.data(), texs.size(), &dtexs))) return rc;
    const swr_voxel_grid_desc &g = s->voxel_grid;
    size_t nvox = (size_t)g.dims[0] * g.dims[1] * g.dims[2];
    if (nvox == 0 || !g.gi_sh4) {
        ctx->err = "voxel grid is empty";
        return SWR_ERR_INVALID;
    }
    if ((rc = upload(ctx, (const float4 *)g.gi_sh4, nvox * 4, &gi))) return rc;
    DevScene &sc = ctx->scene;
    sc.prims = dprims;
    sc.mats = dmats;
    sc.texs = dtexs;
    sc.gi = gi;
    for (int k = 0; k < 3; k++) {
        sc.gdim[k] = g.dims[k];
        sc.gmin[k] = g.world_min[k];
        sc.gmax[k] = g.world_max[k];
        sc.gvs[k] = (g.world_max[k] - g.world_min[k]) / (float)g.dims[k];  // one IEEE f32 divide, the same on host and device
        sc.light_dir[k] = s->light_direction[k];
        sc.light_color[k] = s->light_color[k];
    }
    sc.cubemap = s->cubemap;
    sc.cubemap_specular = s->cubemap_specular;
    sc.brdf_lut = s->brdf_lut;
    CK(cudaStreamSynchronize(ctx->stream));  // host staging vectors go out of scope
    ctx->have_scene = true;
    return SWR_OK;
}

int swr_share_scene(swr_ctx *ctx, const swr_ctx *owner) {
    if (!ctx || !owner || ctx == owner) return SWR_ERR_INVALID;
    if (!owner->have_scene || owner->scene_borrowed || owner->device != ctx->device) {
        ctx->err = "swr_share_scene: the owner must hold an uploaded scene of its own on the same device";
        return SWR_ERR_INVALID;
    }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(owner->stream));  // the owner's upload (its own stream) is complete before this context reads it
    free_scene(ctx);
    ctx->scene = owner->scene;
    ctx->prim_ntris = owner->prim_ntris;
    ctx->prim_nverts = owner->prim_nverts;
    ctx->scene_borrowed = true;
    ctx->have_scene = true;
    ctx->frame_valid = false;
    ctx->have_history = false;
    return SWR_OK;
}

// Geometry half of a pass for one set of draws: host tables -> k_setup -> k_clip -> scan -> scatter. No host synchronisation.
static int launch_geometry(swr_ctx *ctx, GeomSet &g, bool translucent) {
    const std::vector<swr_draw> &draws = g.last_draws;
    const uint32_t nd = (uint32_t)draws.size();
    // the previous frame may still be reading its draw table: this frame gets the other copy (the frame before that has
    // been settled, shading included, before this one is enqueued)
    std::swap(g.draws, g.draws_alt);
    std::swap(g.tri_prefix, g.tri_prefix_alt);
    size_t bytes = (size_t)nd * sizeof(DevDraw) + 2 * (size_t)(nd + 1) * sizeof(uint32_t);
    Staging &st = g.staging[g.staging_next];
    g.staging_next = (g.staging_next + 1) & 3;
    if (!st.done) CK(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
    CK(cudaEventSynchronize(st.done));
    if (st.bytes < bytes) {
        if (st.host) cudaFreeHost(st.host);
        st.host = nullptr;
        st.bytes = 0;
        CK(cudaMallocHost(&st.host, bytes + 4096));
        st.bytes = bytes + 4096;
    }
    DevDraw *hd = (DevDraw *)st.host;
    uint32_t *hp = (uint32_t *)((char *)st.host + (size_t)nd * sizeof(DevDraw));
    uint32_t *hc = hp + (nd + 1);
    uint64_t tris = 0, verts = 0, clip_tris = 0, clusters = 0;
    for (uint32_t i = 0; i < nd; i++) {
        const swr_draw &d = draws[i];
        if (d.primitive >= ctx->prim_ntris.size()) {
            ctx->err = "draw references a primitive that is not in the uploaded scene";
            return SWR_ERR_INVALID;
        }
        memcpy(hd[i].model, d.model, 64);
        memcpy(hd[i].mvp, d.mvp, 64);
        hd[i].prim = d.primitive;
        hd[i].flags = d.flags;
        hd[i].first_tri = d.first_triangle;
        hd[i].reserved = 0;
        hp[i] = (uint32_t)tris;
        hc[i] = (uint32_t)clusters;
        uint32_t nt = ctx->prim_ntris[d.primitive];
        tris += nt;
        clusters += (nt + SWR_CLUSTER_TRIS - 1) / SWR_CLUSTER_TRIS;
        verts += ctx->prim_nverts[d.primitive];
        if (d.flags & SWR_DRAW_CLIP) clip_tris += nt;
        if ((uint64_t)d.first_triangle + nt > 0x1FFFFFFFull) {
            ctx->err = "more than 2^29 triangles in one frame: the seq id (tri*8+fan) would overflow";
            return SWR_ERR_INVALID;
        }
    }
    hp[nd] = (uint32_t)tris;
    hc[nd] = (uint32_t)clusters;
    if (tris >= 0x1FFFFFFFull) {
        ctx->err = "frame too large for 32-bit ids (triangle * 8 + fan)";
        return SWR_ERR_INVALID;
    }
    // records: [0, tris) = fan 0 of every triangle (dense), then the extension area for fans >= 1 of clipped polygons
    size_t want_ext = (size_t)clip_tris / 16 + 4096;
    if (g.ext_cap < want_ext && !g.rendered_once) g.ext_cap = want_ext;
    if (g.ext_cap < 4096) g.ext_cap = 4096;
    const size_t slots = tris + g.ext_cap;
    g.ndraws = nd;
    g.total_tris = (uint32_t)tris;
    g.clusters = (uint32_t)clusters;
    g.total_verts = verts;
    bool ok = g.draws.reserve(nd + 1) == cudaSuccess && g.tri_prefix.reserve(2 * (size_t)nd + 4) == cudaSuccess && g.cull.reserve(((clusters + 255) / 256) * 10 + clusters + 16) == cudaSuccess && g.work.reserve(clusters + 1) == cudaSuccess && g.records.reserve(slots + 1) == cudaSuccess &&
              g.rects.reserve(slots + 1) == cudaSuccess && g.zb.reserve(slots + 16) == cudaSuccess && g.tile_count.reserve((size_t)(ctx->ntiles + 1) * SWR_ZBUCKETS) == cudaSuccess &&
              g.tile_offset.reserve(ctx->ntiles + 1) == cudaSuccess && g.tile_cursor.reserve((size_t)(ctx->ntiles + 1) * SWR_ZBUCKETS) == cudaSuccess &&
              g.counters.reserve(1) == cudaSuccess && (!translucent || g.avgz.reserve(slots + 1) == cudaSuccess);
    if (!ok) {
        ctx->err = "out of device memory for per-frame triangle records";
        return SWR_ERR_OOM;
    }
    size_t want_refs = (size_t)(tris + tris / 2) + (1u << 16);
    if (g.refs.cap < want_refs && !g.rendered_once) {
        if (g.refs.reserve(want_refs) != cudaSuccess) {
            ctx->err = "out of device memory for tile lists";
            return SWR_ERR_OOM;
        }
    }
    if (g.refs.cap == 0 && g.refs.reserve(1u << 16) != cudaSuccess) return SWR_ERR_OOM;
    size_t want_clip = (size_t)clip_tris / 4 + 4096;
    if (g.clip_verts.cap < want_clip && !g.rendered_once) {
        if (g.clip_verts.reserve(want_clip) != cudaSuccess) return SWR_ERR_OOM;
    }
    if (g.clip_verts.cap == 0 && g.clip_verts.reserve(4096) != cudaSuccess) return SWR_ERR_OOM;
    if (g.clip_queue.reserve(clip_tris + 1) != cudaSuccess || g.clip_ext.reserve(tris + 1) != cudaSuccess || g.clip_list.reserve(g.ext_cap + 1) != cudaSuccess)
        return SWR_ERR_OOM;

    cudaStream_t s = ctx->stream;
    // Upload on its own stream, now: the frame in front is still running and the PCIe link is quiet. Issued in stream
    // order behind that frame it would share the link with the frame's 33 MB read-back and take ~0.1 ms instead of ~10 us.
    if (nd) CK(cudaMemcpyAsync(g.draws.p, hd, (size_t)nd * sizeof(DevDraw), cudaMemcpyHostToDevice, ctx->upload_stream));
    CK(cudaMemcpyAsync(g.tri_prefix.p, hp, 2 * (size_t)(nd + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->upload_stream));
    CK(cudaEventRecord(st.done, ctx->upload_stream));
    CK(cudaStreamWaitEvent(s, st.done, 0));
    CK(cudaMemsetAsync(g.tile_count.p, 0, (size_t)(ctx->ntiles + 1) * SWR_ZBUCKETS * sizeof(uint32_t), s));
    CK(cudaMemsetAsync(g.counters.p, 0, sizeof(FrameCounters), s));

    const int rb = ctx->row_begin, re = ctx->row_end;
    SetupParams sp{};
    if (tris > 0) {
        sp.draws = g.draws.p;
        sp.tri_prefix = g.tri_prefix.p;
        sp.ndraws = nd;
        sp.total_tris = (uint32_t)tris;
        sp.cl_prefix = g.tri_prefix.p + (nd + 1);
        sp.total_clusters = (uint32_t)clusters;
        sp.ncull_blocks = (uint32_t)((clusters + 255) / 256);
        sp.cl_mask = g.cull.p;
        sp.cl_blk = sp.cl_mask + (size_t)sp.ncull_blocks * 8;
        sp.cl_blk_off = sp.cl_blk + sp.ncull_blocks;
        sp.cl_draw = sp.cl_blk_off + sp.ncull_blocks + 1;
        sp.work = g.work.p;
        // sort-first: NDC y range of the rows this rank owns, widened by 2 px against the snapping of renderer.rs:834-846
        sp.use_band = (rb > 0 || re < ctx->tiles_y) ? 1 : 0;
        sp.band_hi = 1.0f - 2.0f * ((float)(rb * 64) - 2.0f) / (float)ctx->H;
        sp.band_lo = 1.0f - 2.0f * ((float)(re * 64) + 2.0f) / (float)ctx->H;
        sp.prims = ctx->scene.prims;
        sp.mats = ctx->scene.mats;
        sp.records = g.records.p;
        sp.rects = g.rects.p;
        sp.zb = g.zb.p;
        sp.avgz = translucent ? g.avgz.p : nullptr;
        sp.clip_verts = g.clip_verts.p;
        sp.clip_capacity = (uint32_t)g.clip_verts.cap;
        sp.clip_queue = g.clip_queue.p;
        sp.clip_ext = g.clip_ext.p;
        sp.clip_list = g.clip_list.p;
        sp.ext_capacity = (uint32_t)g.ext_cap;
        sp.tile_count = g.tile_count.p;
        sp.counters = g.counters.p;
        sp.W = ctx->W;
        sp.H = ctx->H;
        sp.tiles_x = ctx->tiles_x;
        sp.tiles_y = ctx->tiles_y;
        sp.row_begin = rb;
        sp.row_end = re;
        // One block per surviving cluster, dispatched in list order (the hardware retires blocks roughly in order, which keeps
        // the tile lists close to submission order: front to back, as the reference sorts its nodes — early-Z lives on that).
        // The survivor count is only known on the device, so the grid is sized from the previous frame's count; blocks
        // stride over whatever is left when the guess is short and exit at once when it is long.
        const uint64_t guess = g.work_hint ? (uint64_t)g.work_hint + g.work_hint / 8 + 64 : clusters;
        const unsigned geom_grid = (unsigned)std::min<uint64_t>(clusters, std::max<uint64_t>(guess, (uint64_t)ctx->num_sms * 16u));
        g.geom_grid = geom_grid;
        ctx->launches++;
        k_cull<<<sp.ncull_blocks, 256, 0, s>>>(sp);
        ctx->launches++;
        k_compact<<<sp.ncull_blocks, 256, 0, s>>>(sp);
        ctx->launches++;
        k_setup<<<geom_grid, SETUP_THREADS, 0, s>>>(sp);
        if (clip_tris > 0) {
            uint64_t want = (clip_tris + CLIP_GROUPS - 1) / CLIP_GROUPS;
            // many short CTAs: an iteration lasts as long as its slowest polygon (block barriers in the accounting), so fine grains win
            const uint64_t cap = (uint64_t)ctx->num_sms * 8u;
            ctx->launches++;
            k_clip<<<(unsigned)(want < cap ? want : cap), CLIP_THREADS, 0, s>>>(sp);
        }
    }
    if (translucent) {
        ctx->launches++;
        k_scan_simple<<<1, 1024, 0, s>>>(g.tile_count.p, g.tile_offset.p, g.tile_cursor.p, ctx->ntiles, g.counters.p, (uint32_t)(g.refs.cap - 16));
    } else {
        const size_t unit_cap = (size_t)ctx->ntiles + g.refs.cap / RASTER_UNIT_MIN + 1;
        const uint32_t cta_slots = (uint32_t)ctx->num_sms * RASTER_MINB;  // k_raster_tiles: resident CTAs per SM
        // a band's tiles are cut no finer than a full frame's would be (floor: one unit per slot)
        float units_per_slot = SWR_UNITS_PER_SLOT * (float)(re - rb) / (float)(ctx->tiles_y > 0 ? ctx->tiles_y : 1);
        if (units_per_slot < ctx->units_floor) units_per_slot = ctx->units_floor;
        if (ctx->unit_list.reserve(unit_cap) != cudaSuccess) return SWR_ERR_OOM;
        ctx->launches++;
        k_scan_tiles<<<1, 1024, 0, s>>>(g.tile_count.p, g.tile_offset.p, g.tile_cursor.p, ctx->ntiles, g.counters.p, (uint32_t)(g.refs.cap - 16), ctx->unit_list.p,
                                        (uint32_t)unit_cap, rb * ctx->tiles_x, re * ctx->tiles_x, cta_slots, units_per_slot, ctx->tile_unit.p, ctx->tile_count_prev.p,
                                        ctx->tile_cycles_prev.p, ctx->have_history ? 1 : 0);
        ctx->have_history = true;
    }
    {
        // cluster blocks (none when nothing was submitted) + the auxiliary blocks: key initialisation of split tiles, clipped fans
        ScatterAux aux{};
        if (!translucent) {
            aux.keys = ctx->keys.p;
            aux.tile_unit = ctx->tile_unit.p;
            aux.tile_offset = g.tile_offset.p;
            aux.tile_begin = rb * ctx->tiles_x;
            aux.tile_end = re * ctx->tiles_x;
        }
        if (tris > 0) {  // nothing submitted: no tile is split, no list
            ctx->launches++;
            const unsigned sc_blocks = (g.geom_grid + SCATTER_BATCH - 1) / SCATTER_BATCH;
            k_scatter<<<sc_blocks + SCATTER_AUX_BLOCKS, SWR_CLUSTER_TRIS, 0, s>>>(sp, g.tile_cursor.p, g.refs.p, sc_blocks, aux);
        }
    }
    CK(cudaGetLastError());
    return SWR_OK;
}

// Enqueue the opaque pass of one frame from ctx->op.last_draws: geometry + tile rasteriser (no host synchronisation).
static int launch_frame(swr_ctx *ctx) {
    cudaStream_t s = ctx->stream;
    GeomSet &g = ctx->op;
    CK(cudaEventRecord(ctx->cur->ev[0], s));
    // raster cycles of the previous frame become the history that sizes this frame's work units
    std::swap(ctx->tile_cycles.p, ctx->tile_cycles_prev.p);
    CK(cudaMemsetAsync(ctx->tile_cycles.p, 0, (ctx->ntiles + 1) * sizeof(uint32_t), s));
    int rc = launch_geometry(ctx, g, false);
    if (rc) return rc;
    const int rb = ctx->row_begin, re = ctx->row_end;
    const uint32_t cta_slots = (uint32_t)ctx->num_sms * RASTER_MINB;
    CK(cudaEventRecord(ctx->cur->ev[1], s));
    if (re > rb) {
        RasterParams rp{};
        rp.records = g.records.p;
        rp.refs = g.refs.p;
        rp.tile_offset = g.tile_offset.p;
        rp.unit_list = ctx->unit_list.p;
        rp.clip_ext = g.clip_ext.p;
        rp.draws = g.draws.p;
        rp.prims = ctx->scene.prims;
        rp.mats = ctx->scene.mats;
        rp.texs = ctx->scene.texs;
        rp.clip_verts = g.clip_verts.p;
        rp.keys = ctx->keys.p;
        rp.counters = g.counters.p;
        rp.tile_cycles = ctx->tile_cycles.p;
        rp.tile_unit = ctx->tile_unit.p;
#ifdef SWR_PROFILE_COUNTERS
        if (ctx->dbg_tiles.reserve((size_t)8192 * 4) == cudaSuccess) rp.dbg_tiles = ctx->dbg_tiles.p;
#endif
        rp.W = ctx->W;
        rp.H = ctx->H;
        rp.tiles_x = ctx->tiles_x;
        rp.tiles_y = ctx->tiles_y;
        rp.row_begin = rb;
        rp.row_end = re;
        // (the global keys of split tiles were set to EMPTY by k_scatter's auxiliary blocks; whole tiles are simply overwritten)
        ctx->launches++;
        k_raster_tiles<<<cta_slots, RASTER_THREADS, raster_smem_bytes(), s>>>(rp);  // persistent: one CTA per resident slot
    }
    CK(cudaEventRecord(ctx->cur->ev[2], s));
    CK(cudaMemcpyAsync(g.h_counters, g.counters.p, sizeof(FrameCounters), cudaMemcpyDeviceToHost, s));
    CK(cudaGetLastError());
    return SWR_OK;
}

static void fill_shade_params(swr_ctx *ctx, const GeomSet &g, ShadeParams &sp) {
    sp.keys = ctx->keys.p;
    sp.records = g.records.p;
    sp.clip_ext = g.clip_ext.p;
    sp.draws = g.draws.p;
    sp.clip_verts = g.clip_verts.p;
    sp.scene = ctx->scene;
    sp.cam = ctx->dcam;
    sp.color = ctx->color.p;
    sp.W = ctx->W;
    sp.H = ctx->H;
    sp.tiles_x = ctx->tiles_x;
    sp.Wp = ctx->tiles_x * SWR_TILE;
    sp.Hp = ctx->tiles_y * SWR_TILE;
    sp.row_begin = ctx->row_begin;
    sp.row_end = ctx->row_end;
    sp.rsqrt_tab = ctx->rsqrt_on ? ctx->rsqrt_tab.p : nullptr;
    sp.rsqrt_bits = ctx->rsqrt_bits;
    sp.ext_bary = ctx->composited ? ctx->bary.p : nullptr;
    sp.sky_row_begin = ctx->sky_r0;
    sp.sky_row_end = ctx->sky_r1;
    sp.counters = ctx->op.counters.p;
}

static int launch_shade(swr_ctx *ctx) {
    cudaStream_t s = ctx->stream;
    const int rb = ctx->row_begin, re = ctx->row_end;
    ctx->last_fused = false;  // also for an empty band: the state always describes the frame shaded last
    if (re > rb) {
        ShadeParams sp{};
        fill_shade_params(ctx, ctx->op, sp);
        dim3 grid(sp.Wp / 16, (re - rb) * SWR_TILE / SHADE_ROWS);
        // Fixed exposure: the frame leaves shading as packed RGBA8 (+ the metering values); only plain opaque frames qualify
        // (the translucent pass blends over the HDR colour, the sort-last composite sums HDR-resolved strips).
        const float fx = ctx->cur ? ctx->cur->fixed_exposure : 0.0f;
        const bool fused = fx > 0.0f && ctx->tr.last_draws.empty() && !ctx->composited;
        if (fused) {
            if (ctx->fused_px.reserve((size_t)ctx->W * ctx->H) != cudaSuccess) return SWR_ERR_OOM;
            // a read-back of the previous fixed-exposure frame may still be copying out of fused_px
            for (int i = 0; i < 2; i++)
                if (ctx->copy_pending[i]) CK(cudaStreamWaitEvent(s, ctx->ev_copied[i], 0));
            sp.rgba = ctx->fused_px.p;
            sp.lum = ctx->lum.p;
            sp.exposure = fx;
            ctx->launches++;
            k_shade<true><<<grid, SHADE_BLOCK, 0, s>>>(sp);
            ctx->last_fused = true;
            ctx->last_fused_exposure = fx;
            CK(cudaEventRecord(ctx->cur->ev[3], s));
            CK(cudaGetLastError());
            return SWR_OK;
        }
        ctx->launches++;
        k_shade<false><<<grid, SHADE_BLOCK, 0, s>>>(sp);
        if (!ctx->tr.last_draws.empty() && !ctx->composited) {
            // translucent pass (tilerasterizer.rs:92-101): own geometry set, per-tile back-to-front sort, forward shading
            int rc = launch_geometry(ctx, ctx->tr, true);
            if (rc) return rc;
            GeomSet &t = ctx->tr;
            ctx->launches++;
            k_sort_translucent<<<ctx->ntiles, 256, 0, s>>>(t.refs.p, t.tile_offset.p, t.avgz.p, t.clip_ext.p, t.counters.p);
            ForwardParams fp{};
            fill_shade_params(ctx, t, fp.sh);
            fp.refs = t.refs.p;
            fp.offset = t.tile_offset.p;
            fp.counters = t.counters.p;
            ctx->launches++;
            k_forward_translucent<<<grid, SHADE_BLOCK, 0, s>>>(fp);
            CK(cudaMemcpyAsync(t.h_counters, t.counters.p, sizeof(FrameCounters), cudaMemcpyDeviceToHost, s));
            ctx->cur->tr_ran = true;
        }
        const int t0 = rb * ctx->tiles_x, t1 = re * ctx->tiles_x;
        ctx->launches++;
        k_luminance<<<(t1 - t0 + 127) / 128, 128, 0, s>>>(ctx->color.p, ctx->lum.p, sp.Wp, sp.Hp, ctx->tiles_x, ctx->ntiles, t0, t1);
    }
    CK(cudaEventRecord(ctx->cur->ev[3], s));
    CK(cudaGetLastError());
    return SWR_OK;
}

static void set_camera(swr_ctx *ctx, const swr_camera *cam) {
    ctx->last_cam = *cam;
    memcpy(ctx->dcam.position, cam->position, 12);
    memcpy(ctx->dcam.skybox_T, cam->skybox_matrix_transposed, 64);
    ctx->dcam.one_over_width = cam->one_over_width;
    ctx->dcam.one_over_height = cam->one_over_height;
}

static int issue_resolve_copy(swr_ctx *ctx, int idx, float exposure, uint32_t *host);

static int resolve_into(swr_ctx *ctx, uint32_t *dst, float exposure);
static int launch_peer_resolve(swr_ctx *ctx, float exposure, uint32_t frame, bool tr_ran);

// Enqueue (or re-enqueue) everything a slot describes: opaque pass, shading, and the resolve that was queued behind it.
static int enqueue_slot(swr_ctx *ctx, FrameSlot &f) {
    ctx->cur = &f;
    ctx->op.h_counters = f.h_op;
    ctx->tr.h_counters = f.h_tr;
    ctx->op.last_draws = f.op_draws;
    ctx->tr.last_draws = f.tr_draws;
    set_camera(ctx, &f.cam);
    f.tr_ran = false;
    int rc;
    if ((rc = launch_frame(ctx))) return rc;
    if (!f.shade) ctx->last_fused = false;  // a visibility-only frame holds neither HDR colour nor packed pixels
    if (f.shade && (rc = launch_shade(ctx))) return rc;
    f.total_tris = ctx->op.total_tris;
    f.total_verts = ctx->op.total_verts;
    f.clusters = ctx->op.clusters;
    CK(cudaEventRecord(f.done, ctx->stream));
    if (f.resolve_kind == 1) return resolve_into(ctx, f.resolve_idx ? ctx->pixels_alt.p : ctx->pixels.p, f.resolve_exposure);
    if (f.resolve_kind == 2) return issue_resolve_copy(ctx, f.resolve_idx, f.resolve_exposure, f.resolve_host);
    if (f.resolve_kind == 3) return launch_peer_resolve(ctx, f.resolve_exposure, f.resolve_frame, f.tr_ran);
    return SWR_OK;
}

static bool overflowed(const FrameSlot &f) {
    const FrameCounters &c = *f.h_op;
    if (c.overflow_refs || c.overflow_clip || c.overflow_ext) return true;
    if (f.tr_ran) {
        const FrameCounters &t = *f.h_tr;
        if (t.overflow_refs || t.overflow_clip || t.overflow_ext) return true;
    }
    return false;
}

// Wait for the OLDEST pending frame's counters. Fine -> publish its statistics. A buffer was too small -> drain the stream,
// grow the buffers and re-enqueue every pending frame in order (their resolves included), then look again.
static int settle_oldest(swr_ctx *ctx) {
    for (int attempt = 0; attempt < 4; attempt++) {
        if (ctx->slots_pending == 0) return SWR_OK;
        FrameSlot &f = ctx->slots[ctx->slot_head];
        CK(cudaEventSynchronize(f.done));
        const FrameCounters c = *f.h_op;
        FrameCounters ct{};
        if (f.tr_ran) ct = *f.h_tr;
        if (ct.overflow_sort) {
            ctx->err = "a tile holds more than 4096 translucent triangles: the in-kernel back-to-front sort does not support that";
            f.pending = false;
            ctx->slot_head = (ctx->slot_head + 1) % SWR_FRAMES_IN_FLIGHT;
            ctx->slots_pending--;
            return SWR_ERR_INVALID;
        }
        if (!overflowed(f)) {
            f.pending = false;
            ctx->slot_head = (ctx->slot_head + 1) % SWR_FRAMES_IN_FLIGHT;
            ctx->slots_pending--;
            ctx->tr.rendered_once = ctx->tr.rendered_once || f.tr_ran;
            ctx->frame_valid = true;
            ctx->op.rendered_once = true;
            swr_frame_stats &st = ctx->stats;
            st.triangles_submitted = f.total_tris;
            st.vertices_submitted = f.total_verts;
            uint64_t binned = 0, uncovered = 0;
            for (int k = 0; k < 32; k++) {
                binned += c.tris_binned[k];
                uncovered += c.refs_uncovered[k];
            }
            st.triangles_binned = binned;
            st.triangles_clipped = c.tris_clipped;
            st.tile_refs = c.tile_refs + uncovered;  // the reference's R: every (triangle, tile) packet it would push
            ctx->op.refs_emitted = c.tile_refs;
            st.tiles = (uint32_t)ctx->ntiles;
            st.clusters_culled = f.clusters - c.work_n;
            ctx->op.work_hint = c.work_n;
            if (f.tr_ran) ctx->tr.work_hint = ct.work_n;
#ifdef SWR_PROFILE_COUNTERS
            if (ctx->dbg_tiles.p) {
                const int nu = c.raster_units < 8192 ? (int)c.raster_units : 8192;
                std::vector<unsigned long long> h((size_t)8192 * 4);
                cudaMemcpy(h.data(), ctx->dbg_tiles.p, h.size() * 8, cudaMemcpyDeviceToHost);
                unsigned long long sum = 0, tmin = ~0ull, tmax = 0;
                std::vector<int> idx(nu);
                for (int t = 0; t < nu; t++) {
                    idx[t] = t;
                    sum += h[t * 4];
                    unsigned long long te = h[t * 4 + 3] >> 20, ts = te - h[t * 4];
                    if (ts < tmin) tmin = ts;
                    if (te > tmax) tmax = te;
                }
                std::sort(idx.begin(), idx.end(), [&](int a, int b) { return h[a * 4] > h[b * 4]; });
                fprintf(stderr, "[swr dbg] units %u (unit refs default %u): sum cycles %llu, /592 = %llu, span %llu; slowest:\n", c.raster_units, c.raster_unit_refs, sum, sum / 592, tmax - tmin);
                for (int k = 0; k < 6 && k < nu; k++)
                    fprintf(stderr, "   unit %d tile %llu cycles %llu refs %llu items %llu\n", idx[k], h[idx[k] * 4 + 3] & 0xFFFFF, h[idx[k] * 4], h[idx[k] * 4 + 1], h[idx[k] * 4 + 2]);
                int med = idx[nu / 2];
                fprintf(stderr, "   median unit cycles %llu refs %llu items %llu\n", h[med * 4], h[med * 4 + 1], h[med * 4 + 2]);
            }
            fprintf(stderr, "[swr dbg] items %llu batches %llu quad-steps %llu fragments %llu warp-iters %llu\n", c.dbg[0], c.dbg[1], c.dbg[2], c.dbg[3], c.dbg[4]);
#endif
            cudaEventElapsedTime(&st.ms_setup_bin, f.ev[0], f.ev[1]);
            cudaEventElapsedTime(&st.ms_raster, f.ev[1], f.ev[2]);
            st.ms_shade = 0.0f;
            if (f.shade) cudaEventElapsedTime(&st.ms_shade, f.ev[2], f.ev[3]);
            return SWR_OK;
        }
        // growth: nothing may be in flight while buffers are reallocated
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
        for (int k = 0; k < ctx->slots_pending; k++) {
            const FrameSlot &g = ctx->slots[(ctx->slot_head + k) % SWR_FRAMES_IN_FLIGHT];
            struct Grow {
                GeomSet *g;
                const FrameCounters *c;
            } grow[2] = {{&ctx->op, g.h_op}, {&ctx->tr, g.tr_ran ? g.h_tr : nullptr}};
            for (const Grow &gw : grow) {
                if (!gw.c) continue;
                if (gw.c->overflow_refs) {
                    size_t want = (size_t)gw.c->tile_refs + (size_t)gw.c->tile_refs / 4 + 4096;
                    if (gw.g->refs.reserve(want) != cudaSuccess) {
                        ctx->err = "out of device memory growing tile lists";
                        return SWR_ERR_OOM;
                    }
                }
                if (gw.c->overflow_ext) gw.g->ext_cap = std::max(gw.g->ext_cap, (size_t)gw.c->ext_records + (size_t)gw.c->ext_records / 4 + 4096);
                if (gw.c->overflow_clip) {
                    size_t want = (size_t)gw.c->clip_verts + (size_t)gw.c->clip_verts / 4 + 4096;
                    if (gw.g->clip_verts.reserve(want) != cudaSuccess) {
                        ctx->err = "out of device memory growing clip vertex buffer";
                        return SWR_ERR_OOM;
                    }
                }
            }
        }
        for (int k = 0; k < ctx->slots_pending; k++) {
            int rc = enqueue_slot(ctx, ctx->slots[(ctx->slot_head + k) % SWR_FRAMES_IN_FLIGHT]);
            if (rc) return rc;
        }
    }
    ctx->err = "frame did not fit after growing buffers";
    return SWR_ERR_OOM;
}

// Settle every pending frame and leave the stream idle.
static int finish_frame(swr_ctx *ctx) {
    CK(cudaSetDevice(ctx->device));
    int rc;
    while (ctx->slots_pending > 0)
        if ((rc = settle_oldest(ctx))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return SWR_OK;
}

int swr_render(swr_ctx *ctx, const swr_camera *camera, const swr_draw *draws, int ndraws, int shade) {
    if (!ctx || !camera || ndraws < 0 || (ndraws > 0 && !draws)) return SWR_ERR_INVALID;
    if (!ctx->have_scene) {
        ctx->err = "swr_render called before swr_upload_scene";
        return SWR_ERR_NO_SCENE;
    }
    CK(cudaSetDevice(ctx->device));
    int rc;
    // at most SWR_FRAMES_IN_FLIGHT frames are unsettled; until the first frame has sized the buffers, one at a time
    while (ctx->slots_pending >= (ctx->op.rendered_once ? SWR_FRAMES_IN_FLIGHT : 1))
        if ((rc = settle_oldest(ctx))) return rc;
    FrameSlot &f = ctx->slots[(ctx->slot_head + ctx->slots_pending) % SWR_FRAMES_IN_FLIGHT];
    f.op_draws.clear();
    f.tr_draws.clear();
    for (int i = 0; i < ndraws; i++) ((draws[i].flags & SWR_DRAW_TRANSLUCENT) ? f.tr_draws : f.op_draws).push_back(draws[i]);
    f.cam = *camera;
    f.shade = shade;
    f.fixed_exposure = ctx->fixed_exposure;
    f.resolve_kind = 0;
    f.pending = true;
    ctx->slots_pending++;
    return enqueue_slot(ctx, f);
}

int swr_set_fixed_exposure(swr_ctx *ctx, float exposure) {
    if (!ctx || !(exposure >= 0.0f) || exposure > 3.0e38f) return SWR_ERR_INVALID;
    ctx->fixed_exposure = exposure;
    return SWR_OK;
}

int swr_shade(swr_ctx *ctx, const swr_camera *camera) {
    if (!ctx || !camera) return SWR_ERR_INVALID;
    if (!ctx->have_scene) return SWR_ERR_NO_SCENE;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = finish_frame(ctx))) return rc;
    set_camera(ctx, camera);
    ctx->cur->shade = 1;  // not pending any more: only its phase events and counter block are reused
    ctx->cur->fixed_exposure = ctx->fixed_exposure;
    ctx->tr.h_counters = ctx->cur->h_tr;
    CK(cudaEventRecord(ctx->cur->ev[2], ctx->stream));
    return launch_shade(ctx);
}

int swr_keys_to_global(swr_ctx *ctx) {
    if (!ctx) return SWR_ERR_INVALID;
    int rc;
    if ((rc = finish_frame(ctx))) return rc;
    const size_t n = (size_t)ctx->ntiles * SWR_TILE_PIXELS;
    ctx->launches++;
    k_keys_to_global<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->keys.p, n, ctx->op.records.p, ctx->op.clip_ext.p);
    CK(cudaGetLastError());
    return SWR_OK;
}

int swr_keys_localize(swr_ctx *ctx) {
    if (!ctx) return SWR_ERR_INVALID;
    int rc;
    if ((rc = finish_frame(ctx))) return rc;
    if (ctx->bary.reserve((size_t)ctx->W * ctx->H) != cudaSuccess) return SWR_ERR_OOM;
    dim3 blk(32, 8), grid((ctx->W + 31) / 32, (ctx->H + 7) / 8);
    ctx->launches++;
    k_keys_localize<<<grid, blk, 0, ctx->stream>>>(ctx->keys.p, ctx->tiles_x, ctx->W, ctx->H, ctx->op.records.p, ctx->op.clip_ext.p, ctx->op.draws.p,
                                                   ctx->op.tri_prefix.p, ctx->op.ndraws, ctx->scene.prims, ctx->bary.p);
    CK(cudaGetLastError());
    return SWR_OK;
}

int swr_shade_composited(swr_ctx *ctx, const swr_camera *camera, int sky_row_begin, int sky_row_end) {
    if (!ctx || !camera) return SWR_ERR_INVALID;
    if (!ctx->bary.p) {
        ctx->err = "swr_shade_composited needs swr_keys_localize first";
        return SWR_ERR_INVALID;
    }
    ctx->composited = true;
    ctx->sky_r0 = sky_row_begin;
    ctx->sky_r1 = sky_row_end;
    int rc = swr_shade(ctx, camera);
    ctx->composited = false;
    return rc;
}

void *swr_device_bary(swr_ctx *ctx) {
    if (!ctx) return nullptr;
    if (ctx->bary.reserve((size_t)ctx->W * ctx->H) != cudaSuccess) return nullptr;
    return ctx->bary.p;
}

// k_resolve of the owned rows into `dst` on the main stream
static int resolve_into(swr_ctx *ctx, uint32_t *dst, float exposure) {
    cudaStream_t s = ctx->stream;
    const size_t W = ctx->W;
    size_t y0 = (size_t)ctx->row_begin * SWR_TILE, y1 = (size_t)ctx->row_end * SWR_TILE;
    if (y1 > (size_t)ctx->H) y1 = ctx->H;
    CK(cudaEventRecord(ctx->ev_res[0], s));
    if (y1 > y0) {
        const int idx = dst == ctx->pixels_alt.p ? 1 : 0;
        dim3 grid((unsigned)((W + 255) / 256), (unsigned)((y1 - y0 + 3) / 4));
        if (ctx->copy_pending[idx]) CK(cudaStreamWaitEvent(s, ctx->ev_copied[idx], 0));
        ctx->launches++;
        k_resolve<<<grid, 256, 0, s>>>(ctx->color.p, ctx->tiles_x * SWR_TILE, dst, ctx->W, (int)y0, (int)y1, exposure, ctx->last_fused ? ctx->fused_px.p : nullptr);
    }
    CK(cudaEventRecord(ctx->ev_res[1], s));
    CK(cudaGetLastError());
    return SWR_OK;
}

// A frame shaded with a fixed exposure holds packed RGBA8 only: it can be handed out with that exposure and no other.
static int check_fused_exposure(swr_ctx *ctx, float exposure) {
    if (ctx->last_fused && exposure != ctx->last_fused_exposure) {
        ctx->err = "the frame was shaded with a fixed exposure (swr_set_fixed_exposure) and holds no HDR colour: resolve it with that exposure, or "
                   "call swr_set_fixed_exposure(ctx, 0) and swr_shade() to shade it again";
        return SWR_ERR_INVALID;
    }
    return SWR_OK;
}

static FrameSlot *newest_pending(swr_ctx *ctx) {
    return ctx->slots_pending ? &ctx->slots[(ctx->slot_head + ctx->slots_pending - 1) % SWR_FRAMES_IN_FLIGHT] : nullptr;
}

int swr_resolve(swr_ctx *ctx, float exposure, uint32_t *out_pixels) {
    if (!ctx) return SWR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int rc;
    // The host form settles the frame first (a buffer-growth replay must happen before pixels are handed out). The
    // device-only form stays asynchronous: it is noted in the frame's slot so that a replay re-issues it.
    if ((rc = check_fused_exposure(ctx, exposure))) return rc;
    if (out_pixels && (rc = finish_frame(ctx))) return rc;
    uint32_t *dst = (uint32_t *)swr_device_pixels(ctx);
    if (!out_pixels) {
        if (FrameSlot *f = newest_pending(ctx)) {
            f->resolve_kind = 1;
            f->resolve_idx = ctx->pix_cur;
            f->resolve_exposure = exposure;
        }
    }
    if ((rc = resolve_into(ctx, dst, exposure))) return rc;
    if (out_pixels) {
        cudaStream_t s = ctx->stream;
        const size_t W = ctx->W;
        size_t y0 = (size_t)ctx->row_begin * SWR_TILE, y1 = (size_t)ctx->row_end * SWR_TILE;
        if (y1 > (size_t)ctx->H) y1 = ctx->H;
        if (y1 > y0) CK(cudaMemcpyAsync(out_pixels + y0 * W, dst + y0 * W, (y1 - y0) * W * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        cudaEventElapsedTime(&ctx->stats.ms_resolve, ctx->ev_res[0], ctx->ev_res[1]);
    }
    return SWR_OK;
}

// ---- peer frame assembly (include/swr.h) ---------------------------------------------------------------------------
static int peer_check_timeout(swr_ctx *ctx) {  // called from synchronising entry points
    uint32_t flag = 0;
    CK(cudaMemcpy(&flag, ctx->peer_local.p + 1, 4, cudaMemcpyDeviceToHost));
    if (flag) {
        cudaMemset(ctx->peer_local.p + 1, 0, 4);
        ctx->err = "peer frame assembly: a device-side wait timed out (frame numbers out of step, or a rank stopped)";
        return SWR_ERR_CUDA;
    }
    return SWR_OK;
}

int swr_peer_export(swr_ctx *ctx, void *handle_out) {
    if (!ctx || !handle_out) return SWR_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) <= SWR_PEER_HANDLE_BYTES, "handle size");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const uint32_t init[2] = {0u, 1u};  // nothing received; frame 1 may be written
    CK(cudaMemset(ctx->pixels.p + (size_t)ctx->W * ctx->H, 0, SWR_PEER_WORDS * 4));
    CK(cudaMemcpy(ctx->pixels.p + (size_t)ctx->W * ctx->H + SWR_PEER_FREE, &init[1], 4, cudaMemcpyHostToDevice));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->pixels.p));
    memset(handle_out, 0, SWR_PEER_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof(h));
    ctx->peer_pixels = ctx->pixels.p;
    ctx->peer_exported = true;
    ctx->pix_cur = 0;
    return SWR_OK;
}

int swr_peer_attach(swr_ctx *ctx, void *assembler_device_pixels) {
    if (!ctx || !assembler_device_pixels) return SWR_ERR_INVALID;
    ctx->peer_pixels = (uint32_t *)assembler_device_pixels;
    return SWR_OK;
}

int swr_peer_open(swr_ctx *ctx, const void *handle) {
    if (!ctx || !handle) return SWR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->peer_ipc_base) {
        cudaIpcCloseMemHandle(ctx->peer_ipc_base);
        ctx->peer_ipc_base = nullptr;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_ipc_base = p;
    ctx->peer_pixels = (uint32_t *)p;
    return SWR_OK;
}

static int launch_peer_resolve(swr_ctx *ctx, float exposure, uint32_t frame, bool tr_ran) {
    cudaStream_t s = ctx->stream;
    const size_t W = ctx->W;
    size_t y0 = (size_t)ctx->row_begin * SWR_TILE, y1 = (size_t)ctx->row_end * SWR_TILE;
    if (y1 > (size_t)ctx->H) y1 = ctx->H;
    uint32_t *ctrl = ctx->peer_pixels + (size_t)ctx->W * ctx->H;
    CK(cudaEventRecord(ctx->ev_res[0], s));
    ctx->launches++;
    k_peer_wait<<<1, 1, 0, s>>>(ctrl + SWR_PEER_FREE, frame, ctx->peer_local.p + 1);
    // an empty band still contributes (one CTA that only signals) so the assembler's count stays in step
    dim3 grid((unsigned)((W + 255) / 256), (unsigned)(y1 > y0 ? (y1 - y0 + 3) / 4 : 1));
    ctx->launches++;
    k_resolve_peer<<<grid, 256, 0, s>>>(ctx->color.p, ctx->tiles_x * SWR_TILE, ctx->peer_pixels, ctrl, ctx->W, (int)y0, (int)(y1 > y0 ? y1 : y0), exposure,
                                        ctx->peer_local.p, ctx->peer_local.p + 1, ctx->op.counters.p, tr_ran ? ctx->tr.counters.p : nullptr,
                                        ctx->last_fused ? ctx->fused_px.p : nullptr);
    CK(cudaEventRecord(ctx->ev_res[1], s));
    CK(cudaGetLastError());
    return SWR_OK;
}

int swr_resolve_peer(swr_ctx *ctx, float exposure, uint32_t frame) {
    if (!ctx) return SWR_ERR_INVALID;
    if (!ctx->peer_pixels || ctx->peer_exported) {
        ctx->err = "swr_resolve_peer needs swr_peer_open / swr_peer_attach first (and is for contributing ranks)";
        return SWR_ERR_INVALID;
    }
    CK(cudaSetDevice(ctx->device));
    // A contribution cannot be taken back once it is signalled. The frame in front of this resolve is guarded on the device
    // (k_resolve_peer neither stores nor signals when the frame's buffers overflowed; the replay re-issues it), so the host
    // does not have to wait for it here. Older frames are settled first: at most one unsettled frame carries a peer
    // resolve, which is what keeps a replay from queueing behind a device-side wait for its own contribution.
    int rc;
    if ((rc = check_fused_exposure(ctx, exposure))) return rc;
    while (ctx->slots_pending > 1)
        if ((rc = settle_oldest(ctx))) return rc;
    bool tr_ran = false;
    if (FrameSlot *f = newest_pending(ctx)) {
        f->resolve_kind = 3;
        f->resolve_exposure = exposure;
        f->resolve_frame = frame;
        tr_ran = f->tr_ran;
    }
    return launch_peer_resolve(ctx, exposure, frame, tr_ran);
}

int swr_peer_collect(swr_ctx *ctx, uint32_t frame, int contributors) {
    if (!ctx || contributors < 0) return SWR_ERR_INVALID;
    if (!ctx->peer_exported) {
        ctx->err = "swr_peer_collect is for the rank that called swr_peer_export";
        return SWR_ERR_INVALID;
    }
    CK(cudaSetDevice(ctx->device));
    uint32_t *ctrl = ctx->pixels.p + (size_t)ctx->W * ctx->H;
    if (contributors > 0) {
        ctx->launches++;
        k_peer_wait<<<1, 1, 0, ctx->stream>>>(ctrl + SWR_PEER_DONE, frame * (uint32_t)contributors, ctx->peer_local.p + 1);
    }
    CK(cudaGetLastError());
    return SWR_OK;
}

int swr_peer_release(swr_ctx *ctx, uint32_t frame) {
    if (!ctx) return SWR_ERR_INVALID;
    if (!ctx->peer_exported) {
        ctx->err = "swr_peer_release is for the rank that called swr_peer_export";
        return SWR_ERR_INVALID;
    }
    CK(cudaSetDevice(ctx->device));
    ctx->launches++;
    k_peer_release<<<1, 1, 0, ctx->stream>>>(ctx->pixels.p + (size_t)ctx->W * ctx->H, frame + 1u);
    CK(cudaGetLastError());
    return SWR_OK;
}

// resolve into pixel buffer `idx` on the main stream, then copy it to `host` on the copy stream
static int issue_resolve_copy(swr_ctx *ctx, int idx, float exposure, uint32_t *host) {
    cudaStream_t s = ctx->stream;
    const size_t W = ctx->W;
    size_t y0 = (size_t)ctx->row_begin * SWR_TILE, y1 = (size_t)ctx->row_end * SWR_TILE;
    if (y1 > (size_t)ctx->H) y1 = ctx->H;
    uint32_t *dst = idx ? ctx->pixels_alt.p : ctx->pixels.p;
    if (ctx->copy_pending[idx]) CK(cudaStreamWaitEvent(s, ctx->ev_copied[idx], 0));  // the previous copy out of this buffer
    if (ctx->last_fused) {
        dst = ctx->fused_px.p;  // already packed: read back straight from where k_shade wrote it (the next fixed-exposure frame waits for this copy)
    } else if (y1 > y0) {
        dim3 grid((unsigned)((W + 255) / 256), (unsigned)((y1 - y0 + 3) / 4));
        ctx->launches++;
        k_resolve<<<grid, 256, 0, s>>>(ctx->color.p, ctx->tiles_x * SWR_TILE, dst, ctx->W, (int)y0, (int)y1, exposure, nullptr);
    }
    CK(cudaEventRecord(ctx->ev_resolved[idx], s));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_resolved[idx], 0));
    if (y1 > y0) CK(cudaMemcpyAsync(host + y0 * W, dst + y0 * W, (y1 - y0) * W * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CK(cudaEventRecord(ctx->ev_copied[idx], ctx->copy_stream));
    ctx->copy_pending[idx] = true;
    CK(cudaGetLastError());
    return SWR_OK;
}

int swr_resolve_async(swr_ctx *ctx, float exposure, uint32_t *out_pixels, int *ticket) {
    if (!ctx || !out_pixels || !ticket) return SWR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CK(cudaEventCreateWithFlags(&ctx->ev_resolved[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
        }
        if (ctx->pixels_alt.reserve((size_t)ctx->W * ctx->H) != cudaSuccess) return SWR_ERR_OOM;
    }
    const int idx = ctx->pix_cur ^ 1;
    int rc = check_fused_exposure(ctx, exposure);
    if (rc) return rc;
    if ((rc = issue_resolve_copy(ctx, idx, exposure, out_pixels))) return rc;
    ctx->pix_cur = idx;
    if (FrameSlot *f = newest_pending(ctx)) {  // re-issued if the frame has to be replayed
        f->resolve_kind = 2;
        f->resolve_idx = idx;
        f->resolve_exposure = exposure;
        f->resolve_host = out_pixels;
    }
    *ticket = idx;
    return SWR_OK;
}

int swr_wait_pixels(swr_ctx *ctx, int ticket) {
    if (!ctx || ticket < 0 || ticket > 1) return SWR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int rc;
    // the frame behind this ticket may still need a buffer-growth replay (which re-issues its resolve + copy): settle the
    // pending frames up to and including it, oldest first
    for (;;) {
        bool mine = false;
        for (int k = 0; k < ctx->slots_pending; k++) {
            const FrameSlot &f = ctx->slots[(ctx->slot_head + k) % SWR_FRAMES_IN_FLIGHT];
            if (f.resolve_kind == 2 && f.resolve_idx == ticket) mine = true;
        }
        if (!mine) break;
        if ((rc = settle_oldest(ctx))) return rc;
    }
    if (ctx->copy_pending[ticket]) CK(cudaEventSynchronize(ctx->ev_copied[ticket]));
    return SWR_OK;
}

int swr_synchronize(swr_ctx *ctx) {
    if (!ctx) return SWR_ERR_INVALID;
    int rc = finish_frame(ctx);
    if (rc == SWR_OK && ctx->peer_pixels) rc = peer_check_timeout(ctx);
    return rc;
}

int swr_read_tile_costs(swr_ctx *ctx, uint32_t *refs, uint32_t *cycles) {
    if (!ctx) return SWR_ERR_INVALID;
    int rc;
    if ((rc = finish_frame(ctx))) return rc;
    if (refs) {  // refs per tile = difference of the scanned offsets (the counters themselves are per depth bucket)
        std::vector<uint32_t> off((size_t)ctx->ntiles + 1);
        CK(cudaMemcpy(off.data(), ctx->op.tile_offset.p, off.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        for (int i = 0; i < ctx->ntiles; i++) refs[i] = off[(size_t)i + 1] - off[(size_t)i];
    }
    if (cycles) CK(cudaMemcpy(cycles, ctx->tile_cycles.p, ctx->ntiles * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return SWR_OK;
}

int swr_read_tile_luminance(swr_ctx *ctx, float *out) {
    if (!ctx || !out) return SWR_ERR_INVALID;
    int rc;
    if ((rc = finish_frame(ctx))) return rc;
    CK(cudaMemcpy(out, ctx->lum.p, ctx->ntiles * sizeof(float), cudaMemcpyDeviceToHost));
    return SWR_OK;
}

int swr_read_visbuffer(swr_ctx *ctx, uint32_t *depth_bits, uint32_t *seq, float *bary1, float *bary2) {
    if (!ctx) return SWR_ERR_INVALID;
    int rc;
    if ((rc = finish_frame(ctx))) return rc;
    if (!ctx->frame_valid) {
        ctx->err = "no frame has been rendered";
        return SWR_ERR_INVALID;
    }
    const size_t n = (size_t)ctx->W * ctx->H;
    uint32_t *d = nullptr;
    CK(cudaMalloc(&d, n * 16));
    VisParams vp{};
    vp.keys = ctx->keys.p;
    vp.records = ctx->op.records.p;
    vp.clip_ext = ctx->op.clip_ext.p;
    vp.draws = ctx->op.draws.p;
    vp.ndraws = ctx->op.ndraws;
    vp.W = ctx->W;
    vp.H = ctx->H;
    vp.tiles_x = ctx->tiles_x;
    dim3 blk(32, 8), grid((ctx->W + 31) / 32, (ctx->H + 7) / 8);
    ctx->launches++;
    k_read_vis<<<grid, blk, 0, ctx->stream>>>(vp, d, d + n, (float *)(d + 2 * n), (float *)(d + 3 * n));
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && depth_bits) e = cudaMemcpy(depth_bits, d, n * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && seq) e = cudaMemcpy(seq, d + n, n * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && bary1) e = cudaMemcpy(bary1, d + 2 * n, n * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && bary2) e = cudaMemcpy(bary2, d + 3 * n, n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    CK(e);
    return SWR_OK;
}

int swr_read_color(swr_ctx *ctx, float *rgb) {
    if (!ctx || !rgb) return SWR_ERR_INVALID;
    int rc;
    if ((rc = finish_frame(ctx))) return rc;
    if (ctx->last_fused) {
        ctx->err = "swr_read_color: the frame was shaded with a fixed exposure and holds no HDR colour (swr_set_fixed_exposure(ctx, 0) + swr_shade())";
        return SWR_ERR_INVALID;
    }
    const size_t Wp = (size_t)ctx->tiles_x * SWR_TILE, n = Wp * ctx->tiles_y * SWR_TILE;
    std::vector<float4> tmp(n);
    CK(cudaMemcpy(tmp.data(), ctx->color.p, n * sizeof(float4), cudaMemcpyDeviceToHost));
    for (size_t y = 0; y < (size_t)ctx->H; y++)
        for (size_t x = 0; x < (size_t)ctx->W; x++) {
            const float4 &c = tmp[y * Wp + x];
            float *o = rgb + (y * ctx->W + x) * 3;
            o[0] = c.x;
            o[1] = c.y;
            o[2] = c.z;
        }
    return SWR_OK;
}

int swr_get_stats(swr_ctx *ctx, swr_frame_stats *out) {
    if (!ctx || !out) return SWR_ERR_INVALID;
    int rc;
    if ((rc = finish_frame(ctx))) return rc;
    ctx->stats.tiles = (uint32_t)ctx->ntiles;
    *out = ctx->stats;
    return SWR_OK;
}

uint64_t swr_launch_count(swr_ctx *ctx) { return ctx ? ctx->launches : 0; }

void *swr_device_pixels(swr_ctx *ctx) { return ctx ? (ctx->pix_cur ? ctx->pixels_alt.p : ctx->pixels.p) : nullptr; }
void *swr_device_keys(swr_ctx *ctx) { return ctx ? ctx->keys.p : nullptr; }
size_t swr_device_keys_bytes(swr_ctx *ctx) { return ctx ? (size_t)ctx->ntiles * SWR_TILE_PIXELS * 8 : 0; }
void *swr_cuda_stream(swr_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

#include "swr_multi.inl"
#include "swr_bake_api.inl"

}  // extern "C"

// swr_raster.cuh — geometry and visibility kernels:
//   k_setup        K1: per-triangle projection, snap, cull, bbox, record write, tile counting; triangles of
//                  clip-flagged draws are classified exactly (all-in / provably empty / needs clipping)
//                  (reference: renderer.rs:470-574, 668-760)
//   k_clip         K2: cooperative Sutherland-Hodgman clipper, 16 lanes = 16 polygon vertices, one group per
//                  straddling triangle (reference: renderer.rs:579-665)
//   k_scan_tiles   exclusive scan of per-tile counts + heaviest-first tile order (replaces bumpqueue.rs)
//   k_scatter      K3 pass 2: warp-aggregated scatter of triangle refs into per-tile lists
//   k_raster_tiles K4: one CTA per 64x64 tile, 64-bit depth|~id keys in shared memory, atomicMin
//                  (reference: tilerasterizer.rs:72-81, 114-383, 511-523; shader.rs:32-63)
//   k_read_vis     parity read-back of (depth bits, seq, bary1, bary2)
#pragma once
#include "swr_device.cuh"
#include "swr_texture.cuh"

struct SetupParams {
    const DevDraw *draws;
    const uint32_t *tri_prefix;  // ndraws + 1, local triangle prefix
    const uint32_t *cl_prefix;   // ndraws + 1, cluster prefix (clusters of SWR_CLUSTER_TRIS consecutive triangles)
    uint32_t total_clusters;
    // k_cull output, in submission order (no atomically appended list: tile lists must stay close to the reference's
    // front-to-back node order, which is what makes early-Z effective): one keep bit per global cluster, the draw of
    // every cluster, and per k_cull block (256 clusters = 8 mask words) the exclusive prefix of survivors.
    uint32_t *cl_mask, *cl_blk, *cl_blk_off, *cl_draw;
    uint32_t ncull_blocks;
    // k_compact: the surviving clusters in submission order, counters->work_n entries of (draw, cluster in draw, dense index of
    // the cluster's first triangle, triangles in the cluster): everything k_setup / k_scatter need to address their
    // triangle without walking draws -> prims first
    uint4 *work;
    float band_lo, band_hi;      // NDC y range of this rank's rows (widened by 2 px); used when use_band != 0 (sort-first)
    int use_band;
    uint32_t ndraws;
    uint32_t total_tris;
    const DevPrim *prims;
    const DevMat *mats;
    TriRecord *records;
    uint32_t *rects;
    uint8_t *zb;  // depth bucket of every record with a rectangle (tile lists are bucket-major: near buckets first)
    float *avgz;  // translucent set: packet.avg_z per record (renderer.rs:765-775), else NULL
    ClipVertex *clip_verts;
    uint32_t clip_capacity;
    uint2 *clip_queue;     // (dense triangle id, draw) of the triangles that need the clipper
    uint32_t *clip_ext;    // per dense triangle: first extension record of its fans 1..n-3 (written by k_clip)
    uint32_t *clip_list;   // ids (tri*8+fan, fan >= 1) of clipped fan triangles that survived, for the list blocks of k_scatter
    uint32_t ext_capacity; // extension records available after the total_tris dense ones
    uint32_t *tile_count;
    FrameCounters *counters;
    int W, H, tiles_x, tiles_y;
    int row_begin, row_end;  // owned tile rows (sort-first)
};

#define SETUP_THREADS SWR_CLUSTER_TRIS  // one block iteration = one cluster

__device__ __forceinline__ uint32_t find_draw(const uint32_t *prefix, uint32_t n, uint32_t g) {
    uint32_t lo = 0, hi = n;  // prefix[lo] <= g < prefix[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (prefix[mid] <= g)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint32_t record_of_id(uint32_t id, const uint32_t *clip_ext) {
    const uint32_t fan = id & 7u, t = id >> 3;
    return fan == 0 ? t : __ldg(clip_ext + t) + fan - 1u;
}

// Depth-bucketed tile lists. The per-tile counters and cursors are SWR_ZBUCKETS wide: a triangle is counted / scattered under
// (tile, bucket of its nearest vertex), the scan lays a tile's buckets out one after the other, so the raster kernel — which
// just walks the tile's ref range — meets near geometry first and its Hi-Z / early-Z tests reject most of what lies behind
// (C3: fragments that reach the depth test 19.8 M -> 11.4 M of 8.3 M visible; raster 0.54 -> 0.46 ms). The order of a list
// never changes a result (the 64-bit min is order-independent), only the work. Buckets are octaves of 1 - z/w, which is
// proportional to 1/w away from the far plane: bucket = exponent distance of (1 - z_ndc) from 1.0.
#ifndef SWR_ZBUCKET_SHIFT
#define SWR_ZBUCKET_SHIFT 0  // buckets per octave = 1 << shift
#endif
#define SWR_ZBUCKETS (8 << SWR_ZBUCKET_SHIFT)
__device__ __forceinline__ uint32_t depth_bucket(float zmin_ndc) {
    const float t = fminf(fmaxf(1.0f - zmin_ndc, 1.0e-30f), 1.0f);  // NaN -> far bucket
    return min((uint32_t)SWR_ZBUCKETS - 1u, (0x3F800000u - __float_as_uint(t)) >> (23 - SWR_ZBUCKET_SHIFT));
}

// Count one triangle's tile rectangle into tile_count. Single-tile rectangles (the common case) are
// aggregated across the warp with match.any so each distinct tile costs one atomic. Must be called by
// all 32 lanes of the warp.
__device__ __forceinline__ void count_tiles(uint32_t rect, uint32_t bucket, uint32_t *tile_count, int tiles_x) {
    int tx0 = rect & 0xFF, ty0 = (rect >> 8) & 0xFF, tx1 = (rect >> 16) & 0xFF, ty1 = rect >> 24;
    bool valid = rect != 0;
    bool single = valid && (tx1 - tx0 == 1) && (ty1 - ty0 == 1);
    unsigned sm = __ballot_sync(0xFFFFFFFFu, single);
    if (single) {
        int tile = (ty0 * tiles_x + tx0) * SWR_ZBUCKETS + (int)bucket;
        unsigned peers = __match_any_sync(sm, tile);
        if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&tile_count[tile], (uint32_t)__popc(peers));
    } else if (valid) {
        for (int ty = ty0; ty < ty1; ty++)
            for (int tx = tx0; tx < tx1; tx++) atomicAdd(&tile_count[(ty * tiles_x + tx) * SWR_ZBUCKETS + (int)bucket], 1u);
    }
}

// renderer.rs:668-760 minus attribute set-up (deferred to shading). Returns the packed tile rectangle
// (0 = culled / off-screen / not owned) and writes the record when it survives.
__device__ __forceinline__ uint32_t emit_triangle(const SetupParams &P, float4 c0, float4 c1, float4 c2, uint32_t slot,
                                                  uint32_t draw, uint32_t seq, uint32_t clipref, bool &nocover, uint32_t &bucket) {
    nocover = false;
    bucket = 0;
    float Wf = (float)P.W, Hf = (float)P.H;
    TriRecord r;
    snap_vertex(c0, Wf, Hf, r.X0, r.Y0);
    snap_vertex(c1, Wf, Hf, r.X1, r.Y1);
    snap_vertex(c2, Wf, Hf, r.X2, r.Y2);
    int area = wsub(wmul(wsub(r.X1, r.X0), wsub(r.Y2, r.Y0)), wmul(wsub(r.X2, r.X0), wsub(r.Y1, r.Y0)));
    if (area > 0) return 0;  // renderer.rs:680 backface cull (zero area is kept)
    int bminx, bminy, bmaxx, bmaxy;
    tri_bbox_pixels(r, P.W, P.H, bminx, bminy, bmaxx, bmaxy);
    if (bmaxx - bminx < 1 || bmaxy - bminy < 1) return 0;
    // renderer.rs:757-760; i32 `/` truncates toward zero, operands are >= 0 here
    int tx0 = bminx / SWR_TILE, ty0 = bminy / SWR_TILE;
    int tx1 = min((bmaxx + SWR_TILE - 1) / SWR_TILE, P.tiles_x);  // column == tiles_x wraps to a duplicate or is skipped
    int ty1 = min((bmaxy + SWR_TILE - 1) / SWR_TILE, P.tiles_y);
    ty0 = max(ty0, P.row_begin);
    ty1 = min(ty1, P.row_end);
    if (tx1 <= tx0 || ty1 <= ty0) return 0;
    const uint32_t rect = (uint32_t)tx0 | ((uint32_t)ty0 << 8) | ((uint32_t)tx1 << 16) | ((uint32_t)ty1 << 24);
    {
        // The reference bins every surviving triangle, including sub-pixel ones that touch no pixel centre
        // (SURVEY 8a R6). When the edge arithmetic is provably exact everywhere on screen (|a|*x+|b|*y+|c| < 2^24), coverage
        // implies a pixel centre inside the triangle's bounding box; if there is none the packet cannot produce a
        // fragment, so it is only COUNTED (stats equal the reference's) and neither record nor refs are written.
        const int mnX = min(min(r.X0, r.X1), r.X2), mxX = max(max(r.X0, r.X1), r.X2);
        const int mnY = min(min(r.Y0, r.Y1), r.Y2), mxY = max(max(r.Y0, r.Y1), r.Y2);
        const bool no_centre = (((mxX - 8) >> 4) < ((mnX - 8 + 15) >> 4)) || (((mxY - 8) >> 4) < ((mnY - 8 + 15) >> 4));
        if (no_centre) {
            const long long xhi = (long long)P.W * 16 + 64, yhi = (long long)P.H * 16 + 64;
            const int ax[3] = {wsub(r.Y2, r.Y1), wsub(r.Y0, r.Y2), wsub(r.Y1, r.Y0)};
            const int bx[3] = {wsub(r.X1, r.X2), wsub(r.X2, r.X0), wsub(r.X0, r.X1)};
            const int cx[3] = {wsub(wmul(r.X2, r.Y1), wmul(r.X1, r.Y2)), wsub(wmul(r.X0, r.Y2), wmul(r.X2, r.Y0)), wsub(wmul(r.X1, r.Y0), wmul(r.X0, r.Y1))};
            bool ex = true;
#pragma unroll
            for (int e = 0; e < 3; e++)
                ex = ex && ((long long)abs((long long)ax[e]) * xhi + (long long)abs((long long)bx[e]) * yhi + abs((long long)cx[e]) + 1 < (1ll << 24));
            if (ex) {
                nocover = true;
                return rect;
            }
        }
    }
    int aabs = area < 0 ? (int)(0u - (unsigned)area) : area;
    r.ooa = fdiv(1.0f, i2f(aabs));
    r.iw0 = fdiv(1.0f, c0.w);
    r.iw1 = fdiv(1.0f, c1.w);
    r.iw2 = fdiv(1.0f, c2.w);
    r.zw0 = fmul(c0.z, r.iw0);
    r.zw1 = fmul(c1.z, r.iw1);
    r.zw2 = fmul(c2.z, r.iw2);
    r.draw = draw;
    r.seq = seq;
    r.clip = clipref;
    if (P.avgz) P.avgz[slot] = fdiv(fadd(fadd(c0.z, c1.z), c2.z), 3.0f);
    else bucket = depth_bucket(fminf(r.zw0, fminf(r.zw1, r.zw2)));  // translucent lists are sorted per tile anyway: one bucket
    uint4 *dst = reinterpret_cast<uint4 *>(&P.records[slot]);
    const uint4 *src = reinterpret_cast<const uint4 *>(&r);
    dst[0] = src[0];
    dst[1] = src[1];
    dst[2] = src[2];
    dst[3] = src[3];
    return rect;
}

// Book-keeping for triangles emit_triangle classified as "cannot cover": they count as binned and their tile
// references count towards R, exactly like the reference's packets, but nothing is written. Block-collective:
// one pair of atomics per block, spread over 32 counter slots. Returns the rectangle to bin (0 if uncovered).
__device__ __forceinline__ uint32_t account_block(uint32_t rect, bool nocover, FrameCounters *counters, uint32_t *s_unc) {
    uint32_t n = 0;
    if (nocover && rect) n = (((rect >> 16) & 0xFF) - (rect & 0xFF)) * ((rect >> 24) - ((rect >> 8) & 0xFF));
    if (threadIdx.x == 0) *s_unc = 0;
    const int binned = __syncthreads_count(rect != 0);
    const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, n);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(s_unc, tot);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (binned) atomicAdd(&counters->tris_binned[blockIdx.x & 31], (unsigned long long)binned);
        if (*s_unc) atomicAdd(&counters->refs_uncovered[blockIdx.x & 31], (unsigned long long)*s_unc);
    }
    return nocover ? 0u : rect;
}

// The same book-keeping per warp, without block barriers (k_clip: its polygon groups finish at very different times).
__device__ __forceinline__ uint32_t account_warp(uint32_t rect, bool nocover, FrameCounters *counters) {
    uint32_t n = 0;
    if (nocover && rect) n = (((rect >> 16) & 0xFF) - (rect & 0xFF)) * ((rect >> 24) - ((rect >> 8) & 0xFF));
    const unsigned bm = __ballot_sync(0xFFFFFFFFu, rect != 0);
    const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, n);
    if ((threadIdx.x & 31) == 0) {
        const uint32_t slot = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & 31u;
        if (bm) atomicAdd(&counters->tris_binned[slot], (unsigned long long)__popc(bm));
        if (tot) atomicAdd(&counters->refs_uncovered[slot], (unsigned long long)tot);
    }
    return nocover ? 0u : rect;
}

// glam dot4 with the clip-plane constants of renderer.rs:581-588: (px*x + pz*z) + (py*y + pw*w)
__device__ __forceinline__ float plane_dist(int pl, float4 v) {
    const float px = (pl == 2) ? 1.0f : (pl == 3 ? -1.0f : 0.0f);
    const float py = (pl == 4) ? 1.0f : (pl == 5 ? -1.0f : 0.0f);
    const float pz = (pl == 0) ? 1.0f : (pl == 1 ? -1.0f : 0.0f);
    return fadd(fadd(fmul(px, v.x), fmul(pz, v.z)), fadd(fmul(py, v.y), fmul(1.0f, v.w)));
}

// Upload-time: bounding sphere (object space) of every cluster of SWR_CLUSTER_TRIS consecutive triangles. One warp per cluster.
__global__ void k_cluster_bounds(const float4 *pos, const uint32_t *idx, uint32_t ntris, float4 *spheres, uint32_t ncl) {
    const uint32_t cl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (cl >= ncl) return;
    const uint32_t i0 = cl * SWR_CLUSTER_TRIS * 3, i1 = min(i0 + SWR_CLUSTER_TRIS * 3, ntris * 3);
    float3 mn = make_float3(3.0e38f, 3.0e38f, 3.0e38f), mx = make_float3(-3.0e38f, -3.0e38f, -3.0e38f);
    for (uint32_t i = i0 + lane; i < i1; i += 32) {
        const float4 p = pos[idx[i]];
        mn = make_float3(fminf(mn.x, p.x), fminf(mn.y, p.y), fminf(mn.z, p.z));
        mx = make_float3(fmaxf(mx.x, p.x), fmaxf(mx.y, p.y), fmaxf(mx.z, p.z));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn.x = fminf(mn.x, __shfl_xor_sync(0xFFFFFFFFu, mn.x, o));
        mn.y = fminf(mn.y, __shfl_xor_sync(0xFFFFFFFFu, mn.y, o));
        mn.z = fminf(mn.z, __shfl_xor_sync(0xFFFFFFFFu, mn.z, o));
        mx.x = fmaxf(mx.x, __shfl_xor_sync(0xFFFFFFFFu, mx.x, o));
        mx.y = fmaxf(mx.y, __shfl_xor_sync(0xFFFFFFFFu, mx.y, o));
        mx.z = fmaxf(mx.z, __shfl_xor_sync(0xFFFFFFFFu, mx.z, o));
    }
    const float3 c = make_float3(0.5f * (mn.x + mx.x), 0.5f * (mn.y + mx.y), 0.5f * (mn.z + mx.z));
    float r2 = 0.0f;
    for (uint32_t i = i0 + lane; i < i1; i += 32) {
        const float4 p = pos[idx[i]];
        const float dx = p.x - c.x, dy = p.y - c.y, dz = p.z - c.z;
        r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xFFFFFFFFu, r2, o));
    if (lane == 0) spheres[cl] = make_float4(c.x, c.y, c.z, sqrtf(r2) * 1.0001f + 1.0e-30f);
}

// Per frame: keep the (draw, cluster) pairs whose bounding sphere can reach the clip-space frustum (and, for sort-first,
// this rank's row band). The planes are the clipper's own (renderer.rs:581-588) pulled back to object space through the
// draw's mvp, so the test needs no scale extraction; every comparison is pushed out by a margin two orders of magnitude
// above the f32 rounding of k_setup's own mvp product. A dropped cluster therefore only holds triangles the reference's
// clipper reduces to nothing, or triangles whose pixels all lie in rows this rank does not own.
// Draws the reference classified Inside are never clipped, whatever their real position (the node sphere quirk), so
// they are kept unless they are provably in front of the eye (inside both z planes => w > 0) and outside the band.
// One thread per (draw, cluster).
__device__ __forceinline__ float cull_plane(const float *m, float kx, float ky, float kz, float kw, float4 sp, float &reach) {
    // plane = kx*row0 + ky*row1 + kz*row2 + kw*row3 of the column-major mvp
    const float a = kx * m[0] + ky * m[1] + kz * m[2] + kw * m[3], b = kx * m[4] + ky * m[5] + kz * m[6] + kw * m[7];
    const float c = kx * m[8] + ky * m[9] + kz * m[10] + kw * m[11], d = kx * m[12] + ky * m[13] + kz * m[14] + kw * m[15];
    const float mag = fabsf(kx) + fabsf(ky) + fabsf(kz) + fabsf(kw);
    const float terms = fabsf(m[0] * sp.x) + fabsf(m[1] * sp.x) + fabsf(m[2] * sp.x) + fabsf(m[3] * sp.x) + fabsf(m[4] * sp.y) + fabsf(m[5] * sp.y) +
                        fabsf(m[6] * sp.y) + fabsf(m[7] * sp.y) + fabsf(m[8] * sp.z) + fabsf(m[9] * sp.z) + fabsf(m[10] * sp.z) + fabsf(m[11] * sp.z) +
                        fabsf(m[12]) + fabsf(m[13]) + fabsf(m[14]) + fabsf(m[15]);
    const float rn = sp.w * sqrtf(a * a + b * b + c * c);
    reach = rn * 1.001f + 1.0e-4f * mag * (terms + rn) + 1.0e-30f;
    return a * sp.x + b * sp.y + c * sp.z + d;
}

__global__ void __launch_bounds__(256) k_cull(SetupParams P) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    bool keep = false;
    uint32_t d = 0, cl = 0;
    if (i < P.total_clusters) {
        d = find_draw(P.cl_prefix, P.ndraws, i);
        cl = i - P.cl_prefix[d];
        const DevDraw &dr = P.draws[d];
        const float4 sp = __ldg(P.prims[dr.prim].cl_sphere + cl);
        const float *m = dr.mvp;
        const bool clip = (dr.flags & 1u) != 0;
        float reach;
        bool out = false;      // provably outside one plane
        bool front = true;     // provably inside both z planes (every point has w > 0)
        float dist = cull_plane(m, 0.0f, 0.0f, 1.0f, 1.0f, sp, reach);
        out |= dist < -reach, front &= dist > reach;
        dist = cull_plane(m, 0.0f, 0.0f, -1.0f, 1.0f, sp, reach);
        out |= dist < -reach, front &= dist > reach;
        dist = cull_plane(m, 1.0f, 0.0f, 0.0f, 1.0f, sp, reach);
        out |= dist < -reach;
        dist = cull_plane(m, -1.0f, 0.0f, 0.0f, 1.0f, sp, reach);
        out |= dist < -reach;
        dist = cull_plane(m, 0.0f, 1.0f, 0.0f, 1.0f, sp, reach);
        out |= dist < -reach;
        dist = cull_plane(m, 0.0f, -1.0f, 0.0f, 1.0f, sp, reach);
        out |= dist < -reach;
        if (!clip) out = false;
        if (P.use_band && (clip || front)) {
            dist = cull_plane(m, 0.0f, 1.0f, 0.0f, -P.band_lo, sp, reach);  // y >= lo * w
            out |= dist < -reach;
            dist = cull_plane(m, 0.0f, -1.0f, 0.0f, P.band_hi, sp, reach);  // y <= hi * w
            out |= dist < -reach;
        }
        keep = !out;  // NaNs compare false everywhere: never culled
    }
    const unsigned bm = __ballot_sync(0xFFFFFFFFu, keep);
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (i < P.total_clusters) P.cl_draw[i] = d;
    if (lane == 0) P.cl_mask[i >> 5] = bm;  // the buffer holds 8 words per launched block
    __shared__ uint32_t s_last, s_warp[8], s_carry;
    const int cnt = __syncthreads_count(keep);
    if (threadIdx.x == 0) {
        P.cl_blk[blockIdx.x] = (uint32_t)cnt;
        __threadfence();
        s_last = atomicAdd(&P.counters->cull_done, 1u) == gridDim.x - 1u ? 1u : 0u;
        s_carry = 0;
    }
    __syncthreads();
    if (!s_last) return;
    // the block that finishes last turns the per-block counts into exclusive offsets (<= a few thousand values)
    __threadfence();
    for (uint32_t base = 0; base < gridDim.x; base += 256) {
        const uint32_t j = base + threadIdx.x;
        const uint32_t v = j < gridDim.x ? __ldcg(P.cl_blk + j) : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) wbase += (w < (int)wid) ? s_warp[w] : 0u;
        const uint32_t carry = s_carry;
        if (j < gridDim.x) P.cl_blk_off[j] = carry + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == 255) s_carry = carry + wbase + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        P.cl_blk_off[gridDim.x] = s_carry;
        P.counters->work_n = s_carry;
    }
}

// Ordered compaction of the survivors: block b of k_cull owns mask words 8b..8b+7 and the offset cl_blk_off[b].
__global__ void __launch_bounds__(256) k_compact(SetupParams P) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t m = __ldg(P.cl_mask + (i >> 5));
    if (!((m >> lane) & 1u)) return;
    uint32_t pos = __ldg(P.cl_blk_off + blockIdx.x) + (uint32_t)__popc(m & ((1u << lane) - 1u));
    for (uint32_t w = 0; w < wid; w++) pos += (uint32_t)__popc(__ldg(P.cl_mask + blockIdx.x * 8u + w));
    const uint32_t d = __ldg(P.cl_draw + i);
    const uint32_t cl = i - __ldg(P.cl_prefix + d);
    const uint32_t ntris = P.prims[P.draws[d].prim].ntris;
    P.work[pos] = make_uint4(d, cl, __ldg(P.tri_prefix + d) + cl * SWR_CLUSTER_TRIS, min((uint32_t)SWR_CLUSTER_TRIS, ntris - cl * SWR_CLUSTER_TRIS));
}

// K1: persistent over the surviving clusters; one block iteration = one cluster, one thread = one triangle.
__global__ void __launch_bounds__(SETUP_THREADS) k_setup(SetupParams P) {
    __shared__ uint32_t s_unc;
    const uint32_t tid = threadIdx.x;
    const unsigned lane = tid & 31;
    const uint32_t nwork = P.counters->work_n;
    for (uint32_t wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
        const uint4 w = P.work[wi];
        const uint32_t d = w.x;
        const DevDraw &dr = P.draws[d];
        const DevPrim &pr = P.prims[dr.prim];
        const uint32_t tri = w.y * SWR_CLUSTER_TRIS + tid;
        const uint32_t g = w.z + tid;
        uint32_t rect = 0, bucket = 0;
        bool queued = false, nocover = false;
        if (tid < w.w) {
            const bool clip = (dr.flags & 1u) != 0;
            const uint32_t dflag = d | ((P.mats[pr.material].flags & 1u) ? SWR_REC_ALPHA : 0u);
            const uint32_t slot = g;  // record of fan 0 = dense triangle index; id = g * 8 + fan
            uint32_t i0 = __ldg(pr.idx + 3 * tri), i1 = __ldg(pr.idx + 3 * tri + 1), i2 = __ldg(pr.idx + 3 * tri + 2);
            float4 c0 = mul_vec4(dr.mvp, __ldg(pr.pos + i0));
            float4 c1 = mul_vec4(dr.mvp, __ldg(pr.pos + i1));
            float4 c2 = mul_vec4(dr.mvp, __ldg(pr.pos + i2));
            const uint32_t seq = (dr.first_tri + tri) * 8u;
            if (clip) {
                // Exact classification against renderer.rs:621-648: planes are visited in order; while all three
                // vertices are inside the polygon is passed through unchanged, so the first plane with any vertex
                // outside decides: all three outside -> the polygon becomes empty (nothing is emitted);
                // mixed -> the real clipper is needed; no such plane -> the triangle goes through untouched.
                int state = 0;  // 0 all-in, 1 empty, 2 needs clipping
#pragma unroll
                for (int pl = 0; pl < 6; pl++) {
                    if (state == 0) {
                        int in = (plane_dist(pl, c0) >= 0.0f ? 1 : 0) + (plane_dist(pl, c1) >= 0.0f ? 1 : 0) + (plane_dist(pl, c2) >= 0.0f ? 1 : 0);
                        if (in == 0)
                            state = 1;
                        else if (in != 3)
                            state = 2;
                    }
                }
                if (state == 0) rect = emit_triangle(P, c0, c1, c2, slot, dflag, seq, SWR_NO_CLIP, nocover, bucket);
                queued = state == 2;
            } else {
                rect = emit_triangle(P, c0, c1, c2, slot, dflag, seq, SWR_NO_CLIP, nocover, bucket);
            }
            P.rects[slot] = nocover ? 0u : rect;  // k_clip overwrites it when fan 0 of a clipped polygon survives
            P.zb[slot] = (uint8_t)bucket;         // written for every submitted triangle: k_scatter loads it beside the rectangle
        }
        rect = account_block(rect, nocover, P.counters, &s_unc);
        count_tiles(rect, bucket, P.tile_count, P.tiles_x);
        // warp-aggregated append to the clip queue
        unsigned qm = __ballot_sync(0xFFFFFFFFu, queued);
        if (qm) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&P.counters->clip_queue_n, (uint32_t)__popc(qm));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (queued) P.clip_queue[base + __popc(qm & ((1u << lane) - 1u))] = make_uint2(g, d);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2: cooperative clipper over the queue written by k_setup
// ---------------------------------------------------------------------------------------------
#define CLIP_THREADS 256
#define CLIP_GROUPS (CLIP_THREADS / 16)

__global__ void __launch_bounds__(CLIP_THREADS) k_clip(SetupParams P) {
    __shared__ float s_poly[CLIP_GROUPS][2][16][17];  // renderer.rs:31-37 as 16 floats per vertex (+1 pad)
    const uint32_t qn = P.counters->clip_queue_n;
    const uint32_t tid = threadIdx.x;
    const uint32_t grp = tid >> 4, lane = tid & 15;
    const unsigned gmask = 0xFFFFu << (tid & 16);  // the 16 lanes of my group inside the warp
    for (uint32_t base = blockIdx.x * CLIP_GROUPS; base < qn; base += gridDim.x * CLIP_GROUPS) {
        const uint32_t e = base + grp;
        uint32_t rect_out = 0, bucket_out = 0;
        bool nocover_out = false;
        if (e < qn) {
            const uint2 qe = P.clip_queue[e];
            const uint32_t gg = qe.x, dd = qe.y;
            const uint32_t ttri = gg - P.tri_prefix[dd];
            const DevDraw &dr = P.draws[dd];
            const DevPrim &pr = P.prims[dr.prim];
            float(*poly)[16][17] = s_poly[grp];
            if (lane < 3) {  // renderer.rs:494-560: build the three input vertices
                uint32_t iv = __ldg(pr.idx + 3 * ttri + lane);
                float4 lp = __ldg(pr.pos + iv);
                float4 wp = mul_vec4(dr.model, lp);
                float4 cp = mul_vec4(dr.mvp, lp);
                float4 n4 = __ldg(pr.nrm + iv), t4 = __ldg(pr.tan + iv);
                float3 nw = mul_mat3(dr.model, make_float3(n4.x, n4.y, n4.z));
                float3 tw = mul_mat3(dr.model, make_float3(t4.x, t4.y, t4.z));
                float2 uv = __ldg(pr.uv + iv);
                float *v = poly[0][lane];
                v[0] = cp.x; v[1] = cp.y; v[2] = cp.z; v[3] = cp.w;
                v[4] = wp.x; v[5] = wp.y; v[6] = wp.z;
                v[7] = nw.x; v[8] = nw.y; v[9] = nw.z;
                v[10] = tw.x; v[11] = tw.y; v[12] = tw.z; v[13] = t4.w;
                v[14] = uv.x; v[15] = uv.y;
            }
            __syncwarp(gmask);
            int n = 3, cur = 0;
            for (int pl = 0; pl < 6 && n > 0; pl++) {  // renderer.rs:621-648
                bool cin = false, pin = false;
                float dc = 0.0f, dp = 0.0f;
                const float *c = poly[cur][lane], *p = poly[cur][(int)lane < n ? (lane + n - 1) % n : 0];
                if ((int)lane < n) {
                    dc = plane_dist(pl, make_float4(c[0], c[1], c[2], c[3]));
                    dp = plane_dist(pl, make_float4(p[0], p[1], p[2], p[3]));
                    cin = dc >= 0.0f;
                    pin = dp >= 0.0f;
                }
                const unsigned allin = __ballot_sync(gmask, (int)lane >= n || cin);
                if (allin == gmask) continue;  // polygon unchanged by this plane
                int cnt = cin ? (pin ? 1 : 2) : (pin ? 1 : 0);
                int incl = cnt;  // inclusive scan over the 16-lane group
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    int t = __shfl_up_sync(gmask, incl, o, 16);
                    if ((int)lane >= o) incl += t;
                }
                const int total = __shfl_sync(gmask, incl, 15, 16);
                const int off = incl - cnt;
                if (cnt > 0 && off + cnt <= 16) {
                    float(*np)[17] = poly[cur ^ 1];
                    if (cin != pin) {  // intersect(prev, curr): t = d0 / (d0 - d1), v0 + (v1 - v0) * t (renderer.rs:598-609)
                        const float t = fdiv(dp, fsub(dp, dc));
#pragma unroll
                        for (int k = 0; k < 16; k++) np[off][k] = fadd(p[k], fmul(fsub(c[k], p[k]), t));
                        if (cin) {
#pragma unroll
                            for (int k = 0; k < 16; k++) np[off + 1][k] = c[k];
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; k++) np[off][k] = c[k];
                    }
                }
                __syncwarp(gmask);
                n = min(total, 16);
                cur ^= 1;
            }
            if (n >= 3) {  // renderer.rs:650-664 fan (0, i, i+1)
                const int nfan = min(n - 2, 8);  // seq carries 3 fan bits; a triangle cut by 6 planes has at most 7 fans
                uint32_t vbase = 0, ext = 0;
                if (lane == 0) {  // clip vertices and extension records in one round trip (FrameCounters: paired allocators)
                    atomicAdd(&P.counters->tris_clipped, 1ull);
                    const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(&P.counters->clip_verts),
                                                             (unsigned long long)(uint32_t)n | ((unsigned long long)(uint32_t)(nfan > 1 ? nfan - 1 : 0) << 32));
                    vbase = (uint32_t)old;
                    ext = (uint32_t)(old >> 32);
                }
                vbase = __shfl_sync(gmask, vbase, 0, 16);
                const bool room = vbase + (uint32_t)n <= P.clip_capacity;
                if (!room && lane == 0) P.counters->overflow_clip = 1;
                if (room && (int)lane < n) {
                    const float *v = poly[cur][lane];
                    ClipVertex cv;
                    cv.wx = v[4]; cv.wy = v[5]; cv.wz = v[6];
                    cv.nx = v[7]; cv.ny = v[8]; cv.nz = v[9];
                    cv.tx = v[10]; cv.ty = v[11]; cv.tz = v[12]; cv.tw = v[13];
                    cv.u = v[14]; cv.v = v[15];
                    P.clip_verts[vbase + lane] = cv;
                }
                if (lane == 0 && nfan > 1) {
                    if (ext + (uint32_t)(nfan - 1) > P.ext_capacity) P.counters->overflow_ext = 1;
                    P.clip_ext[gg] = P.total_tris + ext;
                }
                ext = __shfl_sync(gmask, ext, 0, 16);
                const bool room2 = room && (nfan <= 1 || ext + (uint32_t)(nfan - 1) <= P.ext_capacity);
                if (room2 && lane >= 1 && (int)lane <= nfan) {
                    const float *v0 = poly[cur][0], *v1 = poly[cur][lane], *v2 = poly[cur][lane + 1];
                    const uint32_t fan = lane - 1;
                    const uint32_t sl = fan == 0 ? gg : P.total_tris + ext + fan - 1;
                    rect_out = emit_triangle(P, make_float4(v0[0], v0[1], v0[2], v0[3]), make_float4(v1[0], v1[1], v1[2], v1[3]),
                                             make_float4(v2[0], v2[1], v2[2], v2[3]), sl, dd | ((P.mats[pr.material].flags & 1u) ? SWR_REC_ALPHA : 0u),
                                             (dr.first_tri + ttri) * 8u + fan, vbase, nocover_out, bucket_out);
                    P.rects[sl] = nocover_out ? 0u : rect_out;
                    P.zb[sl] = (uint8_t)bucket_out;
                    if (fan > 0 && rect_out != 0 && !nocover_out) P.clip_list[atomicAdd(&P.counters->clip_list_n, 1u)] = gg * 8u + fan;
                }
            }
        }
        rect_out = account_warp(rect_out, nocover_out, P.counters);  // every lane of the warp gets here (warp-uniform trip count)
        count_tiles(rect_out, bucket_out, P.tile_count, P.tiles_x);
    }
}

// ---------------------------------------------------------------------------------------------
// exclusive scan over tiles (<= 65,025) + the raster work list: one CTA.
// A raster unit is (tile, chunk of <= RASTER_UNIT_REFS refs); tiles with more refs are split over several CTAs (their
// keys are merged with a global 64-bit atomicMin), so one hot tile cannot become the long pole of the frame.
// Units are ordered heaviest-first (bit-length buckets of their ref count).
// ---------------------------------------------------------------------------------------------
#ifndef SWR_UNITS_PER_SLOT
#define SWR_UNITS_PER_SLOT 3.0f  // raster units per resident CTA slot for a FULL frame (measured on C3: 1-3 the same, 4 +4 %, 6 +12 %: every extra split
                                  // costs a key init, a global min-merge and the occlusion the other chunks would have provided); a sort-first band
                                  // scales it by its share of the screen (swr_api.cu) so that its tiles are not cut finer than a full frame's
#endif
#define RASTER_UNIT_MAX 2048u
#define RASTER_UNIT_MIN 32u

// Unit size of one tile. With history (previous frame's measured raster cycles and ref count of this tile) the chunk is
// sized so that one unit costs about `target_cycles`; without history a ref-count based default is used.
__device__ __forceinline__ uint32_t tile_unit_refs(uint32_t refs, uint32_t prev_refs, uint32_t prev_cycles, float target_cycles, uint32_t default_unit) {
    uint32_t u = default_unit;
    if (prev_refs >= 16u && prev_cycles > 0u && target_cycles > 0.0f) {
        const float cycles_per_ref = (float)prev_cycles / (float)prev_refs;
        const float want = target_cycles / cycles_per_ref;
        u = want >= (float)RASTER_UNIT_MAX ? RASTER_UNIT_MAX : (uint32_t)want;
    }
    u = ((u + 31u) / 32u) * 32u;
    return min(max(u, RASTER_UNIT_MIN), RASTER_UNIT_MAX);
}

// One tile's SWR_ZBUCKETS counters (32 bytes) -> total; and its cursors: bucket b starts at base + (counts of buckets < b).
__device__ __forceinline__ uint32_t load_tile_counts(const uint32_t *tile_count, int i, uint32_t c[SWR_ZBUCKETS]) {
    uint32_t sum = 0;
#pragma unroll
    for (int q = 0; q < SWR_ZBUCKETS / 4; q++) {
        const uint4 a = reinterpret_cast<const uint4 *>(tile_count)[(SWR_ZBUCKETS / 4) * i + q];
        c[4 * q] = a.x; c[4 * q + 1] = a.y; c[4 * q + 2] = a.z; c[4 * q + 3] = a.w;
        sum += (a.x + a.y) + (a.z + a.w);
    }
    return sum;
}
__device__ __forceinline__ void store_tile_cursors(uint32_t *tile_cursor, int i, uint32_t base, const uint32_t c[SWR_ZBUCKETS]) {
#pragma unroll
    for (int q = 0; q < SWR_ZBUCKETS / 4; q++) {
        uint4 a;
        a.x = base;
        a.y = a.x + c[4 * q];
        a.z = a.y + c[4 * q + 1];
        a.w = a.z + c[4 * q + 2];
        base = a.w + c[4 * q + 3];
        reinterpret_cast<uint4 *>(tile_cursor)[(SWR_ZBUCKETS / 4) * i + q] = a;
    }
}

__global__ void __launch_bounds__(1024) k_scan_tiles(const uint32_t *tile_count, uint32_t *tile_offset, uint32_t *tile_cursor, int ntiles,
                                                     FrameCounters *counters, uint32_t ref_capacity, uint32_t *unit_list, uint32_t unit_capacity,
                                                     int tile_begin, int tile_end, uint32_t cta_slots, float units_per_slot, uint32_t *tile_unit,
                                                     uint32_t *prev_count, const uint32_t *prev_cycles, int have_history) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry, s_unit;
    __shared__ float s_target;
    __shared__ float s_cyc;
    __shared__ uint32_t s_hist[34], s_cur[34];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float my_cyc = 0.0f;
    if (tid == 0) {
        s_carry = 0;
        s_cyc = 0.0f;
    }
    if (tid < 34) s_hist[tid] = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        int i = base + tid;
        uint32_t cb[SWR_ZBUCKETS];
        uint32_t v = i < ntiles ? load_tile_counts(tile_count, i, cb) : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = s_warp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
                if (lane >= o) wi += t;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        uint32_t excl = s_carry + s_warp[wid] + incl - v;
        if (i < ntiles) {
            tile_offset[i] = excl;
            store_tile_cursors(tile_cursor, i, excl, cb);
            if (have_history && i >= tile_begin && i < tile_end && prev_count[i] >= 16u) {
                // expected cycles of this tile now = cycles per ref last frame x refs now
                my_cyc += (float)prev_cycles[i] / (float)prev_count[i] * (float)v;
            }
        }
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_cyc += __shfl_xor_sync(0xFFFFFFFFu, my_cyc, o);
    if (lane == 0 && my_cyc > 0.0f) atomicAdd(&s_cyc, my_cyc);
    __syncthreads();
    if (tid == 0) {
        tile_offset[ntiles] = s_carry;
        counters->tile_refs = s_carry;
        if (s_carry > ref_capacity) counters->overflow_refs = 1;
        // default unit: ~3 units per resident CTA slot, so the tail stays short even when this context owns only a band
        uint32_t u = (uint32_t)((float)s_carry / (units_per_slot * (float)cta_slots));
        u = ((u + 255u) / 256u) * 256u;
        s_unit = min(max(u, 256u), RASTER_UNIT_MAX);
        s_target = have_history ? s_cyc / (units_per_slot * (float)cta_slots) : 0.0f;
    }
    __syncthreads();
    if (counters->overflow_refs) return;
    const uint32_t U = s_unit;
    const float target = s_target;
    for (int i = tile_begin + tid; i < tile_end; i += 1024) {
        const uint32_t v = tile_offset[i + 1] - tile_offset[i];
        const uint32_t ut = tile_unit_refs(v, have_history ? prev_count[i] : 0u, have_history ? prev_cycles[i] : 0u, target, U);
        tile_unit[i] = ut;
        const uint32_t nfull = v / ut, rem = v % ut;
        if (nfull) atomicAdd(&s_hist[32 - __clz(ut)], nfull);
        if (rem || !nfull) atomicAdd(&s_hist[rem ? 32 - __clz(rem) : 0], 1u);  // an empty tile still needs its keys written
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (int b = 33; b >= 0; b--) {  // biggest units first (with history all full units cost about the same)
            s_cur[b] = acc;
            acc += s_hist[b];
        }
        counters->raster_units = acc <= unit_capacity ? acc : 0;
        counters->raster_unit_refs = U;
        counters->raster_next = 0;
        if (acc > unit_capacity) counters->overflow_refs = 1;  // cannot happen: capacity covers ntiles + ref_capacity / RASTER_UNIT_MIN
    }
    __syncthreads();
    if (counters->overflow_refs) return;
    // one thread per tile; a tile split into many units (a hot tile has hundreds) is left to a whole warp afterwards
    for (int i = tile_begin + tid; i < tile_end; i += 1024) {
        const uint32_t v = tile_offset[i + 1] - tile_offset[i];
        const uint32_t ut = tile_unit[i];
        const uint32_t nfull = v / ut, rem = v % ut;
        prev_count[i] = v;  // history for the next frame
        if (rem || !nfull) unit_list[atomicAdd(&s_cur[rem ? 32 - __clz(rem) : 0], 1u)] = ((uint32_t)i << 14) | nfull;
        if (nfull && nfull <= 8u) {
            const uint32_t base = atomicAdd(&s_cur[32 - __clz(ut)], nfull);
            for (uint32_t k = 0; k < nfull; k++) unit_list[base + k] = ((uint32_t)i << 14) | k;
        }
    }
    for (int i0 = tile_begin + wid * 32; i0 < tile_end; i0 += 1024) {  // warp-uniform trip count
        const int i = i0 + lane;
        uint32_t nfull = 0, ut = 1;
        if (i < tile_end) {
            ut = tile_unit[i];
            nfull = (tile_offset[i + 1] - tile_offset[i]) / ut;
        }
        unsigned big = __ballot_sync(0xFFFFFFFFu, nfull > 8u);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const uint32_t n = __shfl_sync(0xFFFFFFFFu, nfull, src), u = __shfl_sync(0xFFFFFFFFu, ut, src);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&s_cur[32 - __clz(u)], n);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            for (uint32_t k = lane; k < n; k += 32) unit_list[base + k] = ((uint32_t)(i0 + src) << 14) | k;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3 pass 2: scatter refs (ids = dense triangle * 8 + fan). k_scatter: one thread per dense triangle (fan 0);
// the grid's last blocks: the surviving fans >= 1 of clipped polygons.
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ void scatter_rect(uint32_t rect, uint32_t bucket, uint32_t id, uint32_t *tile_cursor, uint32_t *refs, int tiles_x) {
    int tx0 = rect & 0xFF, ty0 = (rect >> 8) & 0xFF, tx1 = (rect >> 16) & 0xFF, ty1 = rect >> 24;
    bool valid = rect != 0;
    bool single = valid && (tx1 - tx0 == 1) && (ty1 - ty0 == 1);
    unsigned sm = __ballot_sync(0xFFFFFFFFu, single);
    if (single) {
        int tile = (ty0 * tiles_x + tx0) * SWR_ZBUCKETS + (int)bucket;
        unsigned peers = __match_any_sync(sm, tile);
        int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(&tile_cursor[tile], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        uint32_t rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
        refs[base + rank] = id;
    } else if (valid) {
        for (int ty = ty0; ty < ty1; ty++)
            for (int tx = tx0; tx < tx1; tx++) refs[atomicAdd(&tile_cursor[(ty * tiles_x + tx) * SWR_ZBUCKETS + (int)bucket], 1u)] = id;
    }
}

// fan-0 records of the surviving clusters (same ordered work list as k_setup): persistent, one cluster per block iteration.
// The last SCATTER_AUX_BLOCKS blocks of the grid do two short, latency-bound side jobs beside the cluster pass instead of
// behind it: (1) tiles that k_scan_tiles split over several raster units get their global keys set to EMPTY (their CTAs
// merge into them with atomicMin; every other tile is written whole by its one unit, so no frame-wide 67 MB clear is
// needed), (2) the surviving fans >= 1 of clipped polygons (k_clip's list) are scattered.
#define SCATTER_AUX_BLOCKS 148u
#ifndef SCATTER_BATCH
#define SCATTER_BATCH 4  // clusters per block iteration: their loads are in flight together (the pass is latency-bound); with depth-bucketed
                         // lists the coarser arrival order no longer costs raster time (before: -0.02 ms here, +0.06 ms there)
#endif
struct ScatterAux {
    unsigned long long *keys;   // NULL: no key initialisation (translucent set)
    const uint32_t *tile_unit;  // refs per raster unit of each tile (k_scan_tiles)
    const uint32_t *tile_offset;
    int tile_begin, tile_end;   // owned tiles
};
__global__ void __launch_bounds__(SWR_CLUSTER_TRIS) k_scatter(SetupParams P, uint32_t *tile_cursor, uint32_t *refs, uint32_t cluster_blocks, ScatterAux A) {
    if (P.counters->overflow_refs) return;  // lists would not fit: the host grows the buffer and replays the frame
    if (blockIdx.x >= cluster_blocks) {
        const uint32_t nb = gridDim.x - cluster_blocks, b = blockIdx.x - cluster_blocks;
        if (A.keys) {
            for (int t = A.tile_begin + (int)b; t < A.tile_end; t += (int)nb) {
                if (__ldg(A.tile_offset + t + 1) - __ldg(A.tile_offset + t) <= __ldg(A.tile_unit + t)) continue;  // one unit: written whole by k_raster_tiles
                uint4 *dst = reinterpret_cast<uint4 *>(A.keys + (size_t)t * SWR_TILE_PIXELS);
                for (int i = threadIdx.x; i < SWR_TILE_PIXELS / 2; i += SWR_CLUSTER_TRIS) dst[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            }
        }
        if (P.counters->overflow_ext) return;
        const uint32_t n = P.counters->clip_list_n;
        for (uint32_t base = b * SWR_CLUSTER_TRIS; base < n; base += nb * SWR_CLUSTER_TRIS) {  // uniform trip count per block
            const uint32_t i = base + threadIdx.x;
            uint32_t id = 0, rect = 0, bucket = 0;
            if (i < n) {
                id = P.clip_list[i];
                const uint32_t slot = record_of_id(id, P.clip_ext);
                rect = __ldg(P.rects + slot);
                bucket = __ldg(P.zb + slot);
            }
            scatter_rect(rect, bucket, id, tile_cursor, refs, P.tiles_x);
        }
        return;
    }
    const uint32_t nwork = P.counters->work_n;
    for (uint32_t w0 = blockIdx.x * SCATTER_BATCH; w0 < nwork; w0 += cluster_blocks * SCATTER_BATCH) {
        uint32_t t[SCATTER_BATCH], rect[SCATTER_BATCH], bucket[SCATTER_BATCH];
#pragma unroll
        for (int k = 0; k < SCATTER_BATCH; k++) {
            uint4 w = make_uint4(0, 0, 0, 0);
            if (w0 + k < nwork) w = P.work[w0 + k];
            t[k] = w.z + threadIdx.x;
            rect[k] = threadIdx.x < w.w ? __ldg(P.rects + t[k]) : 0u;
            bucket[k] = threadIdx.x < w.w ? __ldg(P.zb + t[k]) : 0u;
        }
#pragma unroll
        for (int k = 0; k < SCATTER_BATCH; k++) scatter_rect(rect[k], rect[k] ? bucket[k] : 0u, t[k] * 8u, tile_cursor, refs, P.tiles_x);
    }
}

// ---------------------------------------------------------------------------------------------
// K4: tile rasteriser
// ---------------------------------------------------------------------------------------------
#define RASTER_THREADS 256
#ifndef RASTER_MINB
#define RASTER_MINB 3  // resident CTAs per SM: 80 registers (no spills in the row loop) beat 4 x 64 (C3: 0.61 -> 0.57 ms)
#endif
#define RASTER_WARPS (RASTER_THREADS / 32)

struct RasterParams {
    const TriRecord *records;
    const uint32_t *refs;
    const uint32_t *tile_offset;
    const uint32_t *unit_list;  // tile << 14 | chunk, heaviest first; counters->raster_units entries
    const uint32_t *clip_ext;
    const DevDraw *draws;  // the next five are only touched by alpha-tested fragments
    const DevPrim *prims;
    const DevMat *mats;
    const DevTex *texs;
    const ClipVertex *clip_verts;
    unsigned long long *keys;  // tile-major: tile * 4096 + y * 64 + x; low word = ~id
    FrameCounters *counters;
    unsigned long long *dbg_tiles;  // SWR_PROFILE_COUNTERS: per tile {cycles, refs, items, launch slot}
    const uint32_t *tile_unit;      // refs per unit of each tile (k_scan_tiles)
    uint32_t *tile_cycles;          // per tile: SM cycles spent rasterising it (summed over its units): load-balancing signal
    int W, H, tiles_x, tiles_y;
    int row_begin, row_end;
};

// One batch of up to 256 packets of the tile, structure of arrays in shared memory. Work is then re-distributed
// over the CTA at the granularity of one quad row inside one 16x16 block ("item"), so a triangle that fills the
// tile (128 items of 8 quads) is shared by all warps instead of serialising one of them.
#define REF_STAGES 3
#define REF_STAGE_WORDS (RASTER_THREADS + 4)  // one batch of refs + up to 3 leading words (16-byte alignment of the bulk copy)
struct TileBatch {
    // edge 0 is not stored: a12 + a20 + a01 == 0 and b12 + b20 + b01 == 0 in wrapping i32 arithmetic, so a[0] = -(a[1] + a[2])
    int a[2][RASTER_THREADS], b[2][RASTER_THREADS], c[3][RASTER_THREADS];
    // region origin in quads relative to the tile (5 + 5 bits) | nqx << 10 | nqy << 16 | coarse << 22 | exact << 23 | alpha-tested << 24
    uint32_t geom[RASTER_THREADS];
    // The unit's ref list is contiguous in global memory: each batch's 1 KiB chunk is brought in by ONE bulk async copy
    // (cp.async.bulk, the TMA engine: no LSU instructions, no registers) two batches ahead, completion on an mbarrier.
    // refs[s][lead + k] is the record id of packet k of the batch staged in s (it doubles as the fragment's id).
    __align__(16) uint32_t refs[REF_STAGES][REF_STAGE_WORDS];
    __align__(8) unsigned long long mbar[REF_STAGES];
    float ooa[RASTER_THREADS], iw0[RASTER_THREADS], iwda[RASTER_THREADS], iwdb[RASTER_THREADS];
    float zw0[RASTER_THREADS], zwda[RASTER_THREADS], zwdb[RASTER_THREADS];
    uint32_t zmin_hi[RASTER_THREADS];  // orderable lower bound of every fragment depth of the packet (0 = unknown)
    uint16_t prefix[RASTER_THREADS + 2];  // item prefix (<= 256 * 165 items per batch)
    uint32_t wsum[RASTER_WARPS];
    // Hierarchical Z: per 8x8-pixel block of the tile, an upper bound of the orderable depth (key high word) of its 64
    // pixels. Keys only ever decrease, so a stale value stays an upper bound; refreshed once per batch.
    uint32_t zmax[64];
};

// Shared-memory key layout: pixel (x, y) of the tile lives at y * 64 + (x ^ swz(y)), swz(y) = ((y >> 1) & 7) << 1.
// Lanes of a warp walk different quad rows at the same x; without the swizzle their key addresses differ by multiples
// of 1 KiB and all hit one bank. The XOR is even, so the two pixels of a quad row stay an aligned 16-byte pair.
__device__ __forceinline__ int key_swz(int y) { return ((y >> 1) & 7) << 1; }
__device__ __forceinline__ int key_index(int x, int y) { return y * SWR_TILE + (x ^ key_swz(y)); }
__device__ __forceinline__ uint4 lds_volatile_v4(const void *p) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}

// ---- bulk async copies (TMA engine) + mbarrier, sm_90+ PTX ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion counted in bytes on `bar`; src, dst and bytes are multiples of 16
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes),
                 "r"(smem_addr(bar))
                 : "memory");
}
// shared -> global; the caller orders its generic-proxy writes first (fence.proxy.async) and waits before reusing the source
__device__ __forceinline__ void bulk_store(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("fence.proxy.async.shared::cta;\n\tcp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;" ::"l"(dst_gmem),
                 "r"(smem_addr(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Per-warp queue of covered pixels waiting for the depth computation: coverage is found by lanes walking different
// quad rows (divergent by nature); the expensive part — perspective depth, 64-bit min — then runs up to 32 wide.
#define FRAGQ_CAP 42  // a drain takes min(qn, 32) entries, leaving <= 10: the next plane of <= 32 fragments always fits
struct FragQueue {
    uint32_t pkpix[FRAGQ_CAP];  // packet index in the batch << 16 | pixel index in the tile
    float w1[FRAGQ_CAP], w2[FRAGQ_CAP];
};

// Texture coordinates of a record's three vertices: the primitive's uv stream, or the clipped polygon's stored vertices.
__device__ __forceinline__ void fetch_tri_uv(const DevDraw *draws, const DevPrim *prims, const ClipVertex *clip_verts, const TriRecord &r,
                                             float uu[3], float vv[3]) {
    if (r.clip == SWR_NO_CLIP) {
        const DevDraw &dr = draws[r.draw & SWR_REC_DRAW_MASK];
        const DevPrim &pr = prims[dr.prim];
        const uint32_t tri = (r.seq >> 3) - dr.first_tri;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float2 uv = __ldg(pr.uv + __ldg(pr.idx + 3 * tri + k));
            uu[k] = uv.x;
            vv[k] = uv.y;
        }
    } else {
        const uint32_t fan = r.seq & 7u;
        const uint32_t vi[3] = {r.clip, r.clip + fan + 1u, r.clip + fan + 2u};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            uu[k] = __ldg(&clip_verts[vi[k]].u);
            vv[k] = __ldg(&clip_verts[vi[k]].v);
        }
    }
}

// renderer.rs:738-754: packet.du_dv from the snapped positions and uv/w
__device__ __forceinline__ void packet_du_dv(const TriRecord &r, const float uw[3], const float vw[3], float du_dv[4]) {
    float dx1 = i2f(wsub(r.X1, r.X0)), dx2 = i2f(wsub(r.X2, r.X0)), dy1 = i2f(wsub(r.Y1, r.Y0)), dy2 = i2f(wsub(r.Y2, r.Y0));
    float du1 = uw[1] - uw[0], du2 = uw[2] - uw[0], dv1 = vw[1] - vw[0], dv2 = vw[2] - vw[0];
    float s = r.ooa * 16.0f;
    du_dv[0] = (du1 * dy2 - du2 * dy1) * s;
    du_dv[1] = (du2 * dx1 - du1 * dx2) * s;
    du_dv[2] = (dv1 * dy2 - dv2 * dy1) * s;
    du_dv[3] = (dv2 * dx1 - dv1 * dx2) * s;
}

struct RasterParams;
// shader.rs:311-329 get_alpha_test_mask for one fragment: base-colour alpha at the fragment's uv, mip from the packet's
// UNSCALED du_dv, compared with the material's cutoff. Slow path (re-reads the record and the uv stream): alpha-tested
// materials are the exception.
__device__ __noinline__ bool alpha_test_fragment(const TriRecord *records, const uint32_t *clip_ext, const DevDraw *draws, const DevPrim *prims,
                                                const DevMat *mats, const DevTex *texs, const ClipVertex *clip_verts, uint32_t id, float b1,
                                                float b2, float w) {
    const TriRecord r = records[record_of_id(id, clip_ext)];
    const DevMat &mat = mats[prims[draws[r.draw & SWR_REC_DRAW_MASK].prim].material];
    if (mat.tex_base < 0) return true;  // out_mask = mask
    float uu[3], vv[3], uw[3], vw[3], du_dv[4];
    fetch_tri_uv(draws, prims, clip_verts, r, uu, vv);
    const float iw[3] = {r.iw0, r.iw1, r.iw2};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uw[k] = uu[k] * iw[k];
        vw[k] = vv[k] * iw[k];
    }
    packet_du_dv(r, uw, vw, du_dv);
    const float u = (uw[0] + b1 * (uw[1] - uw[0]) + b2 * (uw[2] - uw[0])) * w;
    const float v = (vw[0] + b1 * (vw[1] - vw[0]) + b2 * (vw[2] - vw[0])) * w;
    const float alpha = sample4(texs[mat.tex_base], u, v, du_dv).w;
    return alpha >= mat.alpha_cutoff;
}

// tilerasterizer.rs:337-342 + depth_test :511-523 (+ the alpha test of shader.rs:40-43) folded into one 64-bit atomicMin.
// The alpha test does not depend on the depth state, so dropping alpha-failed fragments before the min is equivalent to
// the reference's "depth test, then mask, then conditional depth write" for any packet order.
template <typename RP>
__device__ __forceinline__ void shade_fragment(const RP &P, unsigned long long *skeys, const TileBatch &tb, const uint32_t *ids, uint32_t pkpix, float w1, float w2) {
    const int pk = pkpix >> 16, pix = pkpix & 0xFFFF;
    const float ooa = tb.ooa[pk];
    float b1 = fmul(w1, ooa), b2 = fmul(w2, ooa);
    float qq = fadd(fadd(tb.iw0[pk], fmul(b1, tb.iwda[pk])), fmul(b2, tb.iwdb[pk]));
    float wpix = frcp(qq);
    float zz = fadd(fadd(tb.zw0[pk], fmul(b1, tb.zwda[pk])), fmul(b2, tb.zwdb[pk]));
    float z = fmul(zz, wpix);
    if (z == z) {  // NaN never passes `z <= current` (tilerasterizer.rs:516)
        const uint32_t id = ids[pk];
        unsigned long long key = ((unsigned long long)depth_orderable(z) << 32) | (0xFFFFFFFFu - id);
        // keys only ever decrease, so a (possibly stale) read that is already <= key proves the atomic would be a no-op;
        // the shared-memory 64-bit min is a CAS loop (ATOMS.CAST.SPIN.64), worth skipping for occluded fragments
        if (key < *reinterpret_cast<volatile unsigned long long *>(&skeys[pix])) {
            if ((tb.geom[pk] >> 24) & 1u) {
                if (!alpha_test_fragment(P.records, P.clip_ext, P.draws, P.prims, P.mats, P.texs, P.clip_verts, id, b1, b2, wpix)) return;
            }
            atomicMin(&skeys[pix], key);
        }
    }
}

// Drain up to 32 fragments from the tail of the warp's queue; returns the new count.
template <typename RP>
__device__ __forceinline__ int drain_queue(const RP &P, unsigned long long *skeys, const TileBatch &tb, const uint32_t *ids, const FragQueue &fq, int qn, int lane) {
    const int n = min(qn, 32);
    __syncwarp();
    if (lane < n) shade_fragment(P, skeys, tb, ids, fq.pkpix[qn - n + lane], fq.w1[qn - n + lane], fq.w2[qn - n + lane]);
    __syncwarp();
    return qn - n;
}

// One lane's quad row: the four pixels of the current quad as three f32 edge values each. Exact packets use the same
// f32 stepping (their chain is exact by construction), only their origin is computed with integers.
struct RowState {
    float v[4][3];
    float step[3];
    int len;   // quads in the row
    int ybase; // (tile-local y of the quad's upper row) * 64
    int x;     // tile-local x of lane 0 of the current quad (even)
    int swz;   // key_swz(y): the same for both rows of a quad
    int pk;
    uint32_t zmin_hi;
};

__device__ __forceinline__ void row_setup(const TileBatch &tb, int pk, uint32_t local, int tile_x0, int tile_y0, RowState &st) {
    const uint32_t geom = tb.geom[pk];
    const int nqx = (geom >> 10) & 0x3F;
    const bool coarse = (geom >> 22) & 1u, exact = (geom >> 23) & 1u;
    st.pk = pk;
    st.zmin_hi = tb.zmin_hi[pk];
    int qy, bi;
    if (coarse) {
        const uint32_t nbx = (uint32_t)(nqx + 7) >> 3;
        qy = (int)(local / nbx);
        bi = (int)(local - (uint32_t)qy * nbx);
    } else {
        qy = (int)local;
        bi = 0;
    }
    const int qx0 = bi * 8, qx1 = coarse ? min(qx0 + 8, nqx) : nqx;
    st.len = qx1 - qx0;
    const int xq = geom & 31, yq = (geom >> 5) & 31;
    const int xs = tile_x0 * 16 + xq * 32, ys = tile_y0 * 16 + yq * 32;
    int a[3], b[3], c[3];
#pragma unroll
    for (int e = 0; e < 3; e++) {
        if (e > 0) {
            a[e] = tb.a[e - 1][pk];
            b[e] = tb.b[e - 1][pk];
        }
        c[e] = tb.c[e][pk];
    }
    a[0] = wsub(0, wadd(a[1], a[2]));  // the three edge vectors of a triangle sum to zero (wrapping i32, like the reference's)
    b[0] = wsub(0, wadd(b[1], b[2]));
#pragma unroll
    for (int e = 0; e < 3; e++) st.step[e] = i2f(wmul(a[e], 32));
    st.ybase = (yq + qy) * 2 * SWR_TILE;
    st.x = (xq + qx0) * 2;
    st.swz = key_swz((yq + qy) * 2);
    {
        // hierarchical Z: the row lies in one band of 8x8 blocks and touches at most three of them
        const int by = ((yq + qy) * 2) >> 3, bx0 = st.x >> 3, bx1 = (st.x + st.len * 2 - 1) >> 3;
        const volatile uint32_t *zm = tb.zmax + by * 8;
        uint32_t zmx = zm[bx0];
        if (bx1 > bx0) zmx = max(zmx, zm[bx0 + 1]);
        if (bx1 > bx0 + 1) zmx = max(zmx, zm[bx1]);
        if (st.zmin_hi > zmx) {  // every pixel of the row already holds something nearer than the packet can produce
            st.len = 0;
            return;
        }
    }
    if (exact) {
        // integer edge functions at the pixel centres of the first quad (== the f32 chain values, all exact)
        const int sx = xs + qx0 * 32 + 8, sy = ys + qy * 32 + 8;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int e0 = a[k] * sx + b[k] * sy + c[k];
            st.v[0][k] = i2f(e0);
            st.v[1][k] = i2f(e0 + a[k] * 16);
            st.v[2][k] = i2f(e0 + b[k] * 16);
            st.v[3][k] = i2f(e0 + b[k] * 16 + a[k] * 16);
        }
        return;
    }
    // f32 chain replay (tilerasterizer.rs:239-269, 307-329, 371-381)
    const int bx = coarse ? xs + bi * 256 : xs;
    const int by = coarse ? ys + (qy >> 3) * 256 : ys;
    const int j = coarse ? (qy & 7) : qy;
    float A[3], B[3], C[3];
#pragma unroll
    for (int e = 0; e < 3; e++) {
        A[e] = i2f(a[e]);
        B[e] = i2f(b[e]);
        C[e] = i2f(c[e]);
    }
    if (coarse) {
        const float cx0 = i2f(bx + 8), cx1 = i2f(bx + 248), cy0 = i2f(by + 8), cy1 = i2f(by + 248);
        bool reject = false;
#pragma unroll
        for (int e = 0; e < 3; e++) {
            const float ax0 = fmul(A[e], cx0), ax1 = fmul(A[e], cx1), by0 = fmul(B[e], cy0), by1 = fmul(B[e], cy1);
            const float e00 = fadd(fadd(ax0, by0), C[e]), e10 = fadd(fadd(ax1, by0), C[e]);
            const float e01 = fadd(fadd(ax0, by1), C[e]), e11 = fadd(fadd(ax1, by1), C[e]);
            reject = reject || (fmaxf(fmaxf(e00, e10), fmaxf(e01, e11)) < 0.0f);  // block fully outside this edge
        }
        if (reject) {
            st.len = 0;
            return;
        }
    }
    const float x0 = i2f(bx + 8), x1 = i2f(bx + 24), y0 = i2f(by + 8), y1 = i2f(by + 24);
#pragma unroll
    for (int e = 0; e < 3; e++) {
        const float ax0 = fmul(A[e], x0), ax1 = fmul(A[e], x1), by0 = fmul(B[e], y0), by1 = fmul(B[e], y1);
        float v0 = fadd(fadd(ax0, by0), C[e]), v1 = fadd(fadd(ax1, by0), C[e]);
        float v2 = fadd(fadd(ax0, by1), C[e]), v3 = fadd(fadd(ax1, by1), C[e]);
        const float syf = i2f(wmul(b[e], 32));
        for (int k = 0; k < j; k++) {
            v0 = fadd(v0, syf);
            v1 = fadd(v1, syf);
            v2 = fadd(v2, syf);
            v3 = fadd(v3, syf);
        }
        st.v[0][e] = v0;
        st.v[1][e] = v1;
        st.v[2][e] = v2;
        st.v[3][e] = v3;
    }
}

__global__ void __launch_bounds__(RASTER_THREADS, RASTER_MINB) k_raster_tiles(RasterParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *skeys = reinterpret_cast<unsigned long long *>(smem_raw);
    TileBatch &tb = *reinterpret_cast<TileBatch *>(smem_raw + SWR_TILE_PIXELS * 8);
    FragQueue *fqs = reinterpret_cast<FragQueue *>(smem_raw + SWR_TILE_PIXELS * 8 + sizeof(TileBatch));

    if (P.counters->overflow_refs || P.counters->overflow_ext) return;  // lists incomplete; the host replays the frame
    __shared__ uint32_t s_unit_index;
    const uint32_t nunits = P.counters->raster_units;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < REF_STAGES; s++) mbar_init(&tb.mbar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t mphase = 0;  // bit s: parity of the next completion of stage s (every thread waits on every stage, so all agree)
  // persistent CTA: fetch work units (heaviest first) until the list is drained
  for (;;) {
    __syncthreads();  // everybody is done with the previous unit's shared memory
    if (threadIdx.x == 0) s_unit_index = atomicAdd(&P.counters->raster_next, 1u);
    __syncthreads();
    if (s_unit_index >= nunits) break;
    const uint32_t unit = P.unit_list[s_unit_index];
    const int tile = (int)(unit >> 14);
    const uint32_t chunk = unit & 0x3FFFu;
    const int tx = tile % P.tiles_x, ty = tile / P.tiles_x;
    const int tile_x0 = tx * SWR_TILE, tile_y0 = ty * SWR_TILE;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long unit_t0 = clock64();
    FragQueue &fq = fqs[wid];
    uint32_t fq_addr;  // shared-window address of fq.pkpix[0]; w1 / w2 follow at FRAGQ_CAP * 4 / * 8 bytes
    asm volatile("mov.u32 %0, %1;" : "=r"(fq_addr) : "r"((uint32_t)__cvta_generic_to_shared(&fq.pkpix[0])));
    unsigned lt_mask;  // one S2R when the compiler rematerialises it (it does, at 64 registers)
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));

    const uint32_t tile_beg = P.tile_offset[tile], tile_end = P.tile_offset[tile + 1];
    const uint32_t unit_refs = P.tile_unit[tile];
    const bool split = tile_end - tile_beg > unit_refs;  // several CTAs share this tile: merge with atomics at the end
    for (int i = tid; i < SWR_TILE_PIXELS; i += RASTER_THREADS) skeys[i] = SWR_KEY_EMPTY;
    if (tid < 64) tb.zmax[tid] = 0xFFFFFFFFu;
#ifdef SWR_PROFILE_COUNTERS
    const long long dbg_t0 = clock64();
    unsigned long long dbg_items = 0;
#endif

    const uint32_t beg = tile_beg + chunk * unit_refs, end = min(beg + unit_refs, tile_end);
    // Ref staging: batch k of the unit (refs [beg + 256k, ...)) is copied by one bulk async copy into stage k % 3, two
    // batches ahead of its use. The copy starts at the 16-byte boundary below `beg` (`lead` extra words in front) and is
    // rounded up to 16 bytes (the ref buffer carries that much slack behind its last entry).
    const uint32_t lead = beg & 3u;
    const uint32_t nb = (end - beg + RASTER_THREADS - 1) / RASTER_THREADS;
    auto issue = [&](uint32_t k) {
        const uint32_t count = min((uint32_t)RASTER_THREADS, end - beg - k * RASTER_THREADS) + lead;
        const uint32_t bytes = (count * 4u + 15u) & ~15u;
        unsigned long long *bar = &tb.mbar[k % REF_STAGES];
        mbar_expect_tx(bar, bytes);
        bulk_load(tb.refs[k % REF_STAGES], P.refs + (beg - lead + k * RASTER_THREADS), bytes, bar);
    };
    if (tid == 0) {  // the __syncthreads at the top of the unit loop ordered every read of the previous unit's stages before this
        if (nb > 0) issue(0);
        if (nb > 1) issue(1);
    }
    // software pipeline over batches: the record of batch n+1 is in flight (registers) while batch n is rasterised
    uint32_t slot_next = 0;
    uint4 rq0 = make_uint4(0, 0, 0, 0), rq1 = rq0, rq2 = rq0, rq3 = rq0;
    if (nb > 0) {
        mbar_wait(&tb.mbar[0], mphase & 1u);
        mphase ^= 1u;
        if (beg + tid < end) {
            slot_next = tb.refs[0][lead + tid];
            const uint4 *src = reinterpret_cast<const uint4 *>(P.records + record_of_id(slot_next, P.clip_ext));
            rq0 = __ldg(src);
            rq1 = __ldg(src + 1);
            rq2 = __ldg(src + 2);
            rq3 = __ldg(src + 3);
        }
    }
    uint32_t bn = 0;  // batch number inside the unit
    for (uint32_t base = beg; base < end; base += RASTER_THREADS, bn++) {
        __syncthreads();  // previous batch fully consumed (and key init done)
        if (tid == 0 && bn + 2 < nb) issue(bn + 2);  // into the stage batch bn - 1 has just released
        if (base != beg) {
            // Hi-Z refresh: warp w takes blocks 8w..8w+7 (one band), two pixels per lane. Nobody writes keys until the item
            // phase, and readers of zmax in the packet phase below may see the old or the new value: both are upper bounds.
#pragma unroll
            for (int bb = 0; bb < 8; bb++) {
                const int bx = bb, by = wid;
                const int x = bx * 8 + (lane & 7), y0 = by * 8 + (lane >> 3);
                const uint32_t h0 = reinterpret_cast<const uint32_t *>(skeys)[2 * key_index(x, y0) + 1];
                const uint32_t h1 = reinterpret_cast<const uint32_t *>(skeys)[2 * key_index(x, y0 + 4) + 1];
                const uint32_t m = __reduce_max_sync(0xFFFFFFFFu, max(h0, h1));
                if (lane == 0) tb.zmax[by * 8 + bx] = m;
            }
        }
        const uint32_t ri = base + tid;
        uint32_t nitems = 0;
        const uint32_t *ids = &tb.refs[bn % REF_STAGES][lead];  // ids[k] = record id of packet k of this batch
        TriRecord r;
        {
            uint4 *dst = reinterpret_cast<uint4 *>(&r);
            dst[0] = rq0;
            dst[1] = rq1;
            dst[2] = rq2;
            dst[3] = rq3;
        }
        if (bn + 1 < nb) {
            const uint32_t s1 = (bn + 1) % REF_STAGES;
            mbar_wait(&tb.mbar[s1], (mphase >> s1) & 1u);
            mphase ^= 1u << s1;
            if (ri + RASTER_THREADS < end) {
                slot_next = tb.refs[s1][lead + tid];
                const uint4 *src = reinterpret_cast<const uint4 *>(P.records + record_of_id(slot_next, P.clip_ext));
                rq0 = __ldg(src);
                rq1 = __ldg(src + 1);
                rq2 = __ldg(src + 2);
                rq3 = __ldg(src + 3);
            }
        }
        if (ri < end) {
            PacketSetup ps;
            packet_setup(r, P.W, P.H, tile_x0, tile_y0, ps);
            if (!ps.empty) {
                nitems = (uint32_t)((ps.coarse ? ((ps.nqx + 7) >> 3) : 1) * ps.nqy);
#pragma unroll
                for (int e = 0; e < 3; e++) {
                    if (e > 0) {
                        tb.a[e - 1][tid] = ps.a[e];
                        tb.b[e - 1][tid] = ps.b[e];
                    }
                    tb.c[e][tid] = ps.c[e];
                }
                tb.geom[tid] = (uint32_t)((ps.xs >> 5) - tile_x0 / 2) | ((uint32_t)((ps.ys >> 5) - tile_y0 / 2) << 5) | ((uint32_t)ps.nqx << 10) |
                               ((uint32_t)ps.nqy << 16) | (ps.coarse ? 1u << 22 : 0u) | (ps.exact ? 1u << 23 : 0u) | ((r.draw & SWR_REC_ALPHA) ? 1u << 24 : 0u);
                tb.ooa[tid] = r.ooa;
                tb.iw0[tid] = r.iw0;
                tb.iwda[tid] = fsub(r.iw1, r.iw0);
                tb.iwdb[tid] = fsub(r.iw2, r.iw0);
                tb.zw0[tid] = r.zw0;
                tb.zwda[tid] = fsub(r.zw1, r.zw0);
                tb.zwdb[tid] = fsub(r.zw2, r.zw0);
                // Early-Z bound. Fragment depth = interp(z/w) / interp(1/w) is the perspective-correct blend of the vertex
                // clip-space z_k = zw_k / iw_k, so it lies in [min z_k, max z_k] up to a few ulps (and a hair of
                // extrapolation when the f32 chain covers a centre just outside the exact triangle). The bound is pushed down
                // by 1e-3 of the range + 1e-5 relative; when 1/w is not strictly positive and finite no bound is used.
                uint32_t zhi = 0u;
                if (r.iw0 > 0.0f && r.iw1 > 0.0f && r.iw2 > 0.0f && r.iw0 < 3.0e38f && r.iw1 < 3.0e38f && r.iw2 < 3.0e38f) {
                    const float z0 = r.zw0 / r.iw0, z1 = r.zw1 / r.iw1, z2 = r.zw2 / r.iw2;
                    const float zmn = fminf(z0, fminf(z1, z2)), zmx = fmaxf(z0, fmaxf(z1, z2));
                    const float lo = zmn - (zmx - zmn) * 1.0e-3f - fabsf(zmn) * 1.0e-5f - 1.0e-30f;
                    if (lo == lo && fabsf(lo) < 3.0e38f) zhi = depth_orderable(lo);
                }
                tb.zmin_hi[tid] = zhi;
                // Hi-Z at packet level: a packet whose region touches at most 2x2 blocks and whose depth lower bound is behind
                // all of them produces no fragment that could win: it contributes no items at all.
                const int rx0 = (ps.xs >> 4) - tile_x0, ry0 = (ps.ys >> 4) - tile_y0;
                const int bx0 = rx0 >> 3, bx1 = (rx0 + ps.nqx * 2 - 1) >> 3, by0 = ry0 >> 3, by1 = (ry0 + ps.nqy * 2 - 1) >> 3;
                if (zhi != 0u && bx1 - bx0 <= 1 && by1 - by0 <= 1) {
                    const volatile uint32_t *zm = tb.zmax;
                    const uint32_t zmx = max(max(zm[by0 * 8 + bx0], zm[by0 * 8 + bx1]), max(zm[by1 * 8 + bx0], zm[by1 * 8 + bx1]));
                    if (zhi > zmx) nitems = 0;
                }
            }
        }
        // CTA-wide exclusive prefix of item counts
        uint32_t incl = nitems;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) tb.wsum[wid] = incl;
        __syncthreads();
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < RASTER_WARPS; w++) wbase += (w < wid) ? tb.wsum[w] : 0u;
        tb.prefix[tid + 1] = (uint16_t)(wbase + incl);
        if (tid == 0) tb.prefix[0] = 0;
        __syncthreads();
        const uint32_t total = tb.prefix[RASTER_THREADS];
#ifdef SWR_PROFILE_COUNTERS
        dbg_items += total;
        if (tid == 0) {
            atomicAdd((unsigned long long *)&P.counters->dbg[0], (unsigned long long)total);  // items (quad rows)
            atomicAdd((unsigned long long *)&P.counters->dbg[1], 1ull);                       // batches
        }
        unsigned long long dbg_steps = 0, dbg_frags = 0, dbg_maxsteps = 0;
#endif
        int qn = 0;  // fragments queued by this warp (warp-uniform)
        for (uint32_t it0 = (uint32_t)wid * 32u; it0 < total; it0 += RASTER_THREADS) {  // warp-uniform trip count
            const uint32_t it = it0 + lane;
            RowState st;
            st.len = 0;
            st.ybase = 0;
            st.x = 0;
            st.swz = 0;
            st.pk = 0;
            st.zmin_hi = 0;
            if (it < total) {
                int lo = 0, hi = RASTER_THREADS;
#pragma unroll
                for (int s = 0; s < 8; s++) {
                    const int mid = (lo + hi) >> 1;
                    if (tb.prefix[mid] <= it)
                        lo = mid;
                    else
                        hi = mid;
                }
                row_setup(tb, lo, it - tb.prefix[lo], tile_x0, tile_y0, st);
            }
            const int maxlen = __reduce_max_sync(0xFFFFFFFFu, st.len);
#ifdef SWR_PROFILE_COUNTERS
            dbg_steps += st.len;
            if (lane == 0) dbg_maxsteps += maxlen;
#endif
            for (int s = 0; s < maxlen; s++) {
                // coverage of the quad's four pixels: all three edge values >= 0. The chain values are sums and products of
                // finite integers-as-floats with positive coordinates: never NaN, never -0, so ">= 0" is "sign bit clear".
                uint32_t m4 = 0;
                if (s < st.len) {
#pragma unroll
                    for (int l = 0; l < 4; l++) {
                        const uint32_t sg = __float_as_uint(st.v[l][0]) | __float_as_uint(st.v[l][1]) | __float_as_uint(st.v[l][2]);
                        m4 |= ((sg >> 31) ^ 1u) << l;
                    }
                }
                const int phys = st.ybase + (st.x ^ st.swz);  // key index of lane 0; lane 1 = +1, lanes 2,3 = +64
                const uint32_t pkbase = (uint32_t)st.pk << 16;
                if (m4) {
                    // early-Z: the packet's depth lower bound is already behind what the pixel holds -> it cannot win
                    const uint4 k0 = lds_volatile_v4(skeys + phys), k1 = lds_volatile_v4(skeys + phys + SWR_TILE);
                    if (st.zmin_hi > k0.y) m4 &= ~1u;
                    if (st.zmin_hi > k0.w) m4 &= ~2u;
                    if (st.zmin_hi > k1.y) m4 &= ~4u;
                    if (st.zmin_hi > k1.w) m4 &= ~8u;
                }
                if (__any_sync(0xFFFFFFFFu, m4 != 0u)) {
#pragma unroll
                    for (int l = 0; l < 4; l++) {
                        const bool cov = (m4 >> l) & 1u;
                        const unsigned m = __ballot_sync(0xFFFFFFFFu, cov);
                        if (m) {
                            // lazy drain: only when this plane's fragments would not fit, so drains run (nearly) 32 wide
                            if (qn + __popc(m) > FRAGQ_CAP) qn = drain_queue(P, skeys, tb, ids, fq, qn, lane);
                            if (cov) {
                                // explicit shared-window address kept in a register: the compiler would otherwise rebuild the
                                // queue's base (warp id, dynamic-smem window) at every one of the four push sites
                                const uint32_t a = fq_addr + 4u * (uint32_t)(qn + __popc(m & lt_mask));
                                const uint32_t pp = pkbase + (uint32_t)(phys + (l & 1) + (l >> 1) * SWR_TILE);
                                asm volatile("st.shared.u32 [%0], %1;\n\tst.shared.f32 [%0+%4], %2;\n\tst.shared.f32 [%0+%5], %3;" ::"r"(a), "r"(pp), "f"(st.v[l][1]),
                                             "f"(st.v[l][2]), "n"(FRAGQ_CAP * 4), "n"(FRAGQ_CAP * 8)
                                             : "memory");
                            }
                            qn += __popc(m);
#ifdef SWR_PROFILE_COUNTERS
                            if (lane == 0) dbg_frags += __popc(m);
#endif
                        }
                    }
                }
#pragma unroll
                for (int l = 0; l < 4; l++)
#pragma unroll
                    for (int e = 0; e < 3; e++) st.v[l][e] = fadd(st.v[l][e], st.step[e]);  // one quad to the right
                st.x += 2;
            }
        }
        while (qn > 0) qn = drain_queue(P, skeys, tb, ids, fq, qn, lane);  // fragments reference this batch's packets
#ifdef SWR_PROFILE_COUNTERS
        atomicAdd((unsigned long long *)&P.counters->dbg[2], dbg_steps);      // quads stepped (useful lanes)
        if (lane == 0) {
            atomicAdd((unsigned long long *)&P.counters->dbg[3], dbg_frags);     // covered pixels
            atomicAdd((unsigned long long *)&P.counters->dbg[4], dbg_maxsteps);  // warp row-loop iterations
        }
#endif
    }
    __syncthreads();
    unsigned long long *out = P.keys + (size_t)tile * SWR_TILE_PIXELS;
    // the global key buffer uses the same swizzled in-tile order as shared memory (key_index), so a finished tile leaves
    // as ONE 32 KiB bulk async store issued by one thread
    if (!split) {
        if (tid == 0) {
            bulk_store(out, skeys, SWR_TILE_PIXELS * 8);
            bulk_store_wait_read();  // the source may be overwritten once it has been read (next unit's key init)
        }
    } else {  // the key buffer was reset to EMPTY before the launch
        for (int i = tid; i < SWR_TILE_PIXELS; i += RASTER_THREADS) {
            const unsigned long long k = skeys[i];
            if (k != SWR_KEY_EMPTY) atomicMin(&out[i], k);
        }
    }
    if (tid == 0) atomicAdd(&P.tile_cycles[tile], (uint32_t)(clock64() - unit_t0));
#ifdef SWR_PROFILE_COUNTERS
    if (tid == 0 && P.dbg_tiles && s_unit_index < 8192) {
        P.dbg_tiles[s_unit_index * 4 + 0] = (unsigned long long)(clock64() - dbg_t0);
        P.dbg_tiles[s_unit_index * 4 + 1] = end - beg;
        P.dbg_tiles[s_unit_index * 4 + 2] = dbg_items;
        P.dbg_tiles[s_unit_index * 4 + 3] = (unsigned long long)tile | ((unsigned long long)(clock64() & 0xFFFFFFFFFFull) << 20);
    }
#endif
  }
}

// ---------------------------------------------------------------------------------------------
// key -> (record, pixel values)
// ---------------------------------------------------------------------------------------------
struct VisParams {
    const unsigned long long *keys;
    const TriRecord *records;
    const uint32_t *clip_ext;
    const DevDraw *draws;
    uint32_t ndraws;
    int W, H, tiles_x;
};

__device__ __forceinline__ unsigned long long load_key(const unsigned long long *keys, int tiles_x, int px, int py) {
    int tile = (py >> 6) * tiles_x + (px >> 6);
    return keys[(size_t)tile * SWR_TILE_PIXELS + key_index(px & 63, py & 63)];
}

__global__ void k_read_vis(VisParams P, uint32_t *depth_bits, uint32_t *seq, float *bary1, float *bary2) {
    int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= P.W || py >= P.H) return;
    unsigned long long key = load_key(P.keys, P.tiles_x, px, py);
    uint32_t db = SWR_INF_BITS, sq = 0xFFFFFFFFu;
    float b1 = 0.0f, b2 = 0.0f;
    if (key != SWR_KEY_EMPTY && 0xFFFFFFFFu - (uint32_t)key == SWR_ID_FOREIGN) {
        // sort-last: the winner lives on another rank; only its depth is known here (inverse of depth_orderable)
        const uint32_t u = (uint32_t)(key >> 32);
        db = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
        sq = SWR_ID_FOREIGN;
    } else if (key != SWR_KEY_EMPTY) {
        uint32_t slot = 0xFFFFFFFFu - (uint32_t)key;
        TriRecord r = P.records[record_of_id(slot, P.clip_ext)];
        float z;
        // the owner must cover the pixel AND the depth it re-derives must be the depth the rasteriser ranked it by: a
        // raster-vs-read-back disagreement shows up as a poisoned depth word, not as a silently different winner
        if (resolve_pixel(r, P.W, P.H, px, py, b1, b2, z) && depth_orderable(z) == (uint32_t)(key >> 32)) {
            db = __float_as_uint(z);
            sq = r.seq;
        } else {
            db = 0xDEADBEEFu;  // must never happen: the owner of a key covers its pixel
            sq = r.seq;
        }
    }
    size_t o = (size_t)py * P.W + px;
    if (depth_bits) depth_bits[o] = db;
    if (seq) seq[o] = sq;
    if (bary1) bary1[o] = b1;
    if (bary2) bary2[o] = b2;
}

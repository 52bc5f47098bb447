// swr_raster.cuh — geometry and visibility kernels:
//   k_setup        K1 + K2: per-triangle projection, snap, cull, bbox, record write, tile counting, and a
//                  cooperative Sutherland-Hodgman clipper (16 lanes = 16 polygon vertices) for straddlers
//                  (reference: renderer.rs:470-574, 579-665, 668-760)
//   k_scan_tiles   exclusive scan of per-tile counts (replaces bumpqueue.rs block lists)
//   k_scatter      K3 pass 2: warp-aggregated scatter of triangle refs into per-tile lists
//   k_raster_tiles K4: one CTA per 64x64 tile, 64-bit depth|~seq keys in shared memory, atomicMin
//                  (reference: tilerasterizer.rs:72-81, 114-383, 511-523; shader.rs:32-63)
//   k_read_vis     parity read-back of (depth bits, seq, bary1, bary2)
#pragma once
#include "swr_device.cuh"

struct SetupParams {
    const DevDraw *draws;
    const uint32_t *tri_prefix;  // ndraws + 1, local triangle prefix
    uint32_t ndraws;
    uint32_t total_tris;
    const DevPrim *prims;
    TriRecord *records;
    uint32_t *rects;
    ClipVertex *clip_verts;
    uint32_t clip_capacity;
    uint32_t *tile_count;
    FrameCounters *counters;
    int W, H, tiles_x, tiles_y;
    int row_begin, row_end;  // owned tile rows (sort-first)
};

#define SETUP_THREADS 256

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

// Count one triangle's tile rectangle into tile_count. Single-tile rectangles (the common case) are
// aggregated across the warp with match.any so each distinct tile costs one atomic.
__device__ __forceinline__ void count_tiles(uint32_t rect, uint32_t *tile_count, int tiles_x) {
    int tx0 = rect & 0xFF, ty0 = (rect >> 8) & 0xFF, tx1 = (rect >> 16) & 0xFF, ty1 = rect >> 24;
    bool valid = rect != 0;
    bool single = valid && (tx1 - tx0 == 1) && (ty1 - ty0 == 1);
    unsigned sm = __ballot_sync(0xFFFFFFFFu, single);
    if (single) {
        int tile = ty0 * tiles_x + tx0;
        unsigned peers = __match_any_sync(sm, tile);
        if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&tile_count[tile], (uint32_t)__popc(peers));
    } else if (valid) {
        for (int ty = ty0; ty < ty1; ty++)
            for (int tx = tx0; tx < tx1; tx++) atomicAdd(&tile_count[ty * tiles_x + tx], 1u);
    }
}

// renderer.rs:668-760 minus attribute set-up (deferred to shading). Returns the packed tile rectangle
// (0 = culled / off-screen / not owned) and writes the record when it survives.
__device__ __forceinline__ uint32_t emit_triangle(const SetupParams &P, float4 c0, float4 c1, float4 c2, uint32_t slot,
                                                  uint32_t draw, uint32_t seq, uint32_t clipref) {
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
    uint4 *dst = reinterpret_cast<uint4 *>(&P.records[slot]);
    const uint4 *src = reinterpret_cast<const uint4 *>(&r);
    dst[0] = src[0];
    dst[1] = src[1];
    dst[2] = src[2];
    dst[3] = src[3];
    return (uint32_t)tx0 | ((uint32_t)ty0 << 8) | ((uint32_t)tx1 << 16) | ((uint32_t)ty1 << 24);
}

// glam dot4 with the clip-plane constants of renderer.rs:581-588: (px*x + pz*z) + (py*y + pw*w)
__device__ __forceinline__ float plane_dist(int pl, float4 v) {
    const float px = (pl == 2) ? 1.0f : (pl == 3 ? -1.0f : 0.0f);
    const float py = (pl == 4) ? 1.0f : (pl == 5 ? -1.0f : 0.0f);
    const float pz = (pl == 0) ? 1.0f : (pl == 1 ? -1.0f : 0.0f);
    return fadd(fadd(fmul(px, v.x), fmul(pz, v.z)), fadd(fmul(py, v.y), fmul(1.0f, v.w)));
}

struct ClipVtx {  // renderer.rs:31-37 as 16 floats
    float f[16];  // 0-3 pos_clip, 4-6 pos_world, 7-9 normal, 10-13 tangent, 14-15 uv
};

#define CLIP_GROUPS (SETUP_THREADS / 16)

__global__ void __launch_bounds__(SETUP_THREADS) k_setup(SetupParams P) {
    __shared__ uint32_t s_queue[SETUP_THREADS];  // local thread ids of triangles that need the clipper
    __shared__ uint32_t s_qn;
    __shared__ uint32_t s_draw0;
    __shared__ float s_poly[CLIP_GROUPS][2][16][17];  // +1 pad: lanes read rows

    const uint32_t tid = threadIdx.x;
    const uint32_t g0 = blockIdx.x * SETUP_THREADS;
    if (tid == 0) {
        s_qn = 0;
        s_draw0 = find_draw(P.tri_prefix, P.ndraws, g0);
    }
    __syncthreads();
    const uint32_t g = g0 + tid;
    bool active = g < P.total_tris;
    uint32_t d = s_draw0;
    uint32_t tri = 0, slot = 0, rect = 0;
    bool survived = false;
    bool queued = false;
    if (active) {
        while (g >= P.tri_prefix[d + 1]) d++;
        tri = g - P.tri_prefix[d];
        const DevDraw &dr = P.draws[d];
        const DevPrim &pr = P.prims[dr.prim];
        const bool clip = (dr.flags & 1u) != 0;
        slot = dr.slot_base + tri * (clip ? 7u : 1u);
        uint32_t i0 = __ldg(pr.idx + 3 * tri), i1 = __ldg(pr.idx + 3 * tri + 1), i2 = __ldg(pr.idx + 3 * tri + 2);
        float4 c0 = mul_vec4(dr.mvp, __ldg(pr.pos + i0));
        float4 c1 = mul_vec4(dr.mvp, __ldg(pr.pos + i1));
        float4 c2 = mul_vec4(dr.mvp, __ldg(pr.pos + i2));
        uint32_t seq = (dr.first_tri + tri) * 8u;
        if (clip) {
            bool all_in = true;
#pragma unroll
            for (int pl = 0; pl < 6; pl++)
                all_in = all_in && (plane_dist(pl, c0) >= 0.0f) && (plane_dist(pl, c1) >= 0.0f) && (plane_dist(pl, c2) >= 0.0f);
            if (all_in) {
                rect = emit_triangle(P, c0, c1, c2, slot, d, seq, SWR_NO_CLIP);  // polygon passes S-H unchanged
            } else {
                queued = true;
                s_queue[atomicAdd(&s_qn, 1u)] = tid;
            }
            P.rects[slot] = rect;
#pragma unroll
            for (int k = 1; k < 7; k++) P.rects[slot + k] = 0;
        } else {
            rect = emit_triangle(P, c0, c1, c2, slot, d, seq, SWR_NO_CLIP);
            P.rects[slot] = rect;
        }
        survived = rect != 0;
    }
    count_tiles(rect, P.tile_count, P.tiles_x);
    unsigned surv = __ballot_sync(0xFFFFFFFFu, survived);
    if ((tid & 31) == 0 && surv) atomicAdd(&P.counters->tris_binned, (unsigned long long)__popc(surv));
    __syncthreads();

    // ---- K2: cooperative clipper, one 16-lane group per straddling triangle ------------------------------
    const uint32_t qn = s_qn;
    if (qn == 0) return;
    const uint32_t grp = tid >> 4, lane = tid & 15;
    const unsigned gmask = 0xFFFFu << ((tid & 16));  // the 16 lanes of my group inside the warp
    for (uint32_t base = 0; base < qn; base += CLIP_GROUPS) {
        const uint32_t e = base + grp;
        const bool have = e < qn;  // warp-uniform per half only; all syncs below use gmask
        uint32_t rect_out = 0;
        bool surv_out = false;
        if (have) {
            const uint32_t ltid = s_queue[e];
            const uint32_t gg = g0 + ltid;
            uint32_t dd = s_draw0;
            while (gg >= P.tri_prefix[dd + 1]) dd++;
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
                ClipVtx vc, vp;
                bool cin = false, pin = false;
                float dc = 0.0f, dp = 0.0f;
                if ((int)lane < n) {
                    const float *c = poly[cur][lane];
                    const float *p = poly[cur][(lane + n - 1) % n];
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        vc.f[k] = c[k];
                        vp.f[k] = p[k];
                    }
                    dc = plane_dist(pl, make_float4(vc.f[0], vc.f[1], vc.f[2], vc.f[3]));
                    dp = plane_dist(pl, make_float4(vp.f[0], vp.f[1], vp.f[2], vp.f[3]));
                    cin = dc >= 0.0f;
                    pin = dp >= 0.0f;
                }
                int cnt = cin ? (pin ? 1 : 2) : (pin ? 1 : 0);
                int incl = cnt;  // inclusive scan over the 16-lane group
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    int t = __shfl_up_sync(gmask, incl, o, 16);
                    if ((int)lane >= o) incl += t;
                }
                int total = __shfl_sync(gmask, incl, 15, 16);
                int off = incl - cnt;
                __syncwarp(gmask);
                if (cnt > 0 && off + cnt <= 16) {
                    float(*np)[17] = poly[cur ^ 1];
                    if (cin != pin) {  // intersect(prev, curr): t = d0 / (d0 - d1), v0 + (v1 - v0) * t (renderer.rs:598-609)
                        float t = fdiv(dp, fsub(dp, dc));
#pragma unroll
                        for (int k = 0; k < 16; k++) np[off][k] = fadd(vp.f[k], fmul(fsub(vc.f[k], vp.f[k]), t));
                        if (cin) {
#pragma unroll
                            for (int k = 0; k < 16; k++) np[off + 1][k] = vc.f[k];
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; k++) np[off][k] = vc.f[k];
                    }
                }
                __syncwarp(gmask);
                n = min(total, 16);
                cur ^= 1;
            }
            if (n >= 3) {  // renderer.rs:650-664 fan (0, i, i+1)
                if (lane == 0) atomicAdd(&P.counters->tris_clipped, 1ull);
                uint32_t vbase = 0;
                if (lane == 0) vbase = atomicAdd(&P.counters->clip_verts, (uint32_t)n);
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
                if (room && lane >= 1 && (int)lane <= n - 2) {
                    const float *v0 = poly[cur][0], *v1 = poly[cur][lane], *v2 = poly[cur][lane + 1];
                    const uint32_t fan = lane - 1;
                    const uint32_t sl = dr.slot_base + ttri * 7u + fan;
                    if (fan < 7) {
                        rect_out = emit_triangle(P, make_float4(v0[0], v0[1], v0[2], v0[3]), make_float4(v1[0], v1[1], v1[2], v1[3]),
                                                 make_float4(v2[0], v2[1], v2[2], v2[3]), sl, dd, (dr.first_tri + ttri) * 8u + fan, vbase);
                        P.rects[sl] = rect_out;
                        surv_out = rect_out != 0;
                    }
                }
            }
        }
        // all 32 lanes reconverge here for the warp-aggregated counting
        __syncwarp();
        count_tiles(rect_out, P.tile_count, P.tiles_x);
        unsigned sv = __ballot_sync(0xFFFFFFFFu, surv_out);
        if ((tid & 31) == 0 && sv) atomicAdd(&P.counters->tris_binned, (unsigned long long)__popc(sv));
    }
}

// ---------------------------------------------------------------------------------------------
// exclusive scan over tiles (<= 65,025): one CTA
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_scan_tiles(const uint32_t *tile_count, uint32_t *tile_offset, uint32_t *tile_cursor, int ntiles,
                                                     FrameCounters *counters, uint32_t ref_capacity) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        int i = base + tid;
        uint32_t v = i < ntiles ? tile_count[i] : 0u;
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
            tile_cursor[i] = excl;
        }
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        tile_offset[ntiles] = s_carry;
        counters->tile_refs = s_carry;
        if (s_carry > ref_capacity) counters->overflow_refs = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// K3 pass 2: scatter refs. One thread per record slot.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter(const uint32_t *rects, uint32_t nslots, uint32_t *tile_cursor, uint32_t *refs,
                                                 uint32_t ref_capacity, const FrameCounters *counters, int tiles_x) {
    if (counters->overflow_refs) return;  // lists would not fit: the host grows the buffer and replays the frame
    uint32_t slot = blockIdx.x * 256 + threadIdx.x;
    uint32_t rect = slot < nslots ? __ldg(rects + slot) : 0u;
    int tx0 = rect & 0xFF, ty0 = (rect >> 8) & 0xFF, tx1 = (rect >> 16) & 0xFF, ty1 = rect >> 24;
    bool valid = rect != 0;
    bool single = valid && (tx1 - tx0 == 1) && (ty1 - ty0 == 1);
    unsigned sm = __ballot_sync(0xFFFFFFFFu, single);
    if (single) {
        int tile = ty0 * tiles_x + tx0;
        unsigned peers = __match_any_sync(sm, tile);
        int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(&tile_cursor[tile], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        uint32_t rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
        refs[base + rank] = slot;
    } else if (valid) {
        for (int ty = ty0; ty < ty1; ty++)
            for (int tx = tx0; tx < tx1; tx++) refs[atomicAdd(&tile_cursor[ty * tiles_x + tx], 1u)] = slot;
    }
}

// ---------------------------------------------------------------------------------------------
// K4: tile rasteriser
// ---------------------------------------------------------------------------------------------
#define RASTER_THREADS 256
#define RASTER_WARPS (RASTER_THREADS / 32)

struct RasterParams {
    const TriRecord *records;
    const uint32_t *refs;
    const uint32_t *tile_offset;
    unsigned long long *keys;  // tile-major: tile * 4096 + y * 64 + x
    const FrameCounters *counters;
    int W, H, tiles_x, tiles_y;
    int row_begin, row_end;
};

// Per-warp staging of 32 packets (structure of arrays: conflict-free when lanes read different packets).
struct WarpPackets {
    int a[3][32], b[3][32], c[3][32];
    int xs[32], ys[32], nqx[32];
    uint32_t flags[32];  // bit0 coarse, bit1 exact
    uint32_t slot[32];
    float ooa[32], iw0[32], iwda[32], iwdb[32], zw0[32], zwda[32], zwdb[32];
    uint32_t prefix[33];
};

__device__ __forceinline__ void raster_lane(unsigned long long *skeys, const WarpPackets &wp, int pk, int q, int tile_x0, int tile_y0) {
    const int nqx = wp.nqx[pk];
    const int qy = q / nqx, qx = q - qy * nqx;
    PacketSetup p;
#pragma unroll
    for (int e = 0; e < 3; e++) {
        p.a[e] = wp.a[e][pk];
        p.b[e] = wp.b[e][pk];
        p.c[e] = wp.c[e][pk];
    }
    p.xs = wp.xs[pk];
    p.ys = wp.ys[pk];
    const uint32_t fl = wp.flags[pk];
    p.coarse = fl & 1u;
    p.exact = fl & 2u;
    const float ooa = wp.ooa[pk], iw0 = wp.iw0[pk], iwda = wp.iwda[pk], iwdb = wp.iwdb[pk];
    const float zw0 = wp.zw0[pk], zwda = wp.zwda[pk], zwdb = wp.zwdb[pk];
    const uint32_t idlow = 0xFFFFFFFFu - wp.slot[pk];
    const int px0 = (p.xs >> 4) + 2 * qx, py0 = (p.ys >> 4) + 2 * qy;
#pragma unroll
    for (int l = 0; l < 4; l++) {
        const int lx = l & 1, ly = l >> 1;
        float w1, w2;
        bool cov;
        if (p.exact) {
            int e[3];
            eval_exact(p, (px0 + lx) * 16 + 8, (py0 + ly) * 16 + 8, e);
            cov = (e[0] | e[1] | e[2]) >= 0;
            w1 = i2f(e[1]);
            w2 = i2f(e[2]);
        } else {
            float r[3];
            cov = eval_chain(p, qx, qy, lx, ly, r);
            cov = cov && (r[0] >= 0.0f && r[1] >= 0.0f && r[2] >= 0.0f);
            w1 = r[1];
            w2 = r[2];
        }
        if (cov) {
            float b1 = fmul(w1, ooa), b2 = fmul(w2, ooa);
            float qq = fadd(fadd(iw0, fmul(b1, iwda)), fmul(b2, iwdb));
            float wpix = fdiv(1.0f, qq);
            float zz = fadd(fadd(zw0, fmul(b1, zwda)), fmul(b2, zwdb));
            float z = fmul(zz, wpix);
            if (z == z) {  // NaN never passes `z <= current` (tilerasterizer.rs:516)
                unsigned long long key = ((unsigned long long)depth_orderable(z) << 32) | idlow;
                atomicMin(&skeys[(py0 + ly - tile_y0) * SWR_TILE + (px0 + lx - tile_x0)], key);
            }
        }
    }
}

__global__ void __launch_bounds__(RASTER_THREADS) k_raster_tiles(RasterParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *skeys = reinterpret_cast<unsigned long long *>(smem_raw);
    WarpPackets *wps = reinterpret_cast<WarpPackets *>(smem_raw + SWR_TILE_PIXELS * 8);

    if (P.counters->overflow_refs) return;  // tile lists were not written; the host replays the frame
    const int tile = blockIdx.x + P.row_begin * P.tiles_x;
    const int tx = tile % P.tiles_x, ty = tile / P.tiles_x;
    const int tile_x0 = tx * SWR_TILE, tile_y0 = ty * SWR_TILE;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (int i = tid; i < SWR_TILE_PIXELS; i += RASTER_THREADS) skeys[i] = SWR_KEY_EMPTY;
    __syncthreads();

    const uint32_t beg = P.tile_offset[tile], end = P.tile_offset[tile + 1];
    WarpPackets &wp = wps[wid];
    for (uint32_t base = beg + wid * 32; base < end; base += RASTER_WARPS * 32) {
        const uint32_t ri = base + lane;
        uint32_t nq = 0;
        if (ri < end) {
            const uint32_t slot = __ldg(P.refs + ri);
            TriRecord r;
            const uint4 *src = reinterpret_cast<const uint4 *>(P.records + slot);
            uint4 *dst = reinterpret_cast<uint4 *>(&r);
            dst[0] = __ldg(src);
            dst[1] = __ldg(src + 1);
            dst[2] = __ldg(src + 2);
            dst[3] = __ldg(src + 3);
            PacketSetup ps;
            packet_setup(r, P.W, P.H, tile_x0, tile_y0, ps);
            if (!ps.empty) {
                nq = (uint32_t)(ps.nqx * ps.nqy);
#pragma unroll
                for (int e = 0; e < 3; e++) {
                    wp.a[e][lane] = ps.a[e];
                    wp.b[e][lane] = ps.b[e];
                    wp.c[e][lane] = ps.c[e];
                }
                wp.xs[lane] = ps.xs;
                wp.ys[lane] = ps.ys;
                wp.nqx[lane] = ps.nqx;
                wp.flags[lane] = (ps.coarse ? 1u : 0u) | (ps.exact ? 2u : 0u);
                wp.slot[lane] = slot;
                wp.ooa[lane] = r.ooa;
                wp.iw0[lane] = r.iw0;
                wp.iwda[lane] = fsub(r.iw1, r.iw0);
                wp.iwdb[lane] = fsub(r.iw2, r.iw0);
                wp.zw0[lane] = r.zw0;
                wp.zwda[lane] = fsub(r.zw1, r.zw0);
                wp.zwdb[lane] = fsub(r.zw2, r.zw0);
            }
        }
        // exclusive prefix of quad counts across the warp
        uint32_t incl = nq;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        wp.prefix[lane + 1] = incl;
        if (lane == 0) wp.prefix[0] = 0;
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        __syncwarp();
        // load-balanced expansion: lane handles quad item `it`; find its packet by binary search
        for (uint32_t it = lane; it < total; it += 32) {
            int lo = 0, hi = 32;
#pragma unroll
            for (int s = 0; s < 5; s++) {
                int mid = (lo + hi) >> 1;
                if (wp.prefix[mid] <= it)
                    lo = mid;
                else
                    hi = mid;
            }
            raster_lane(skeys, wp, lo, (int)(it - wp.prefix[lo]), tile_x0, tile_y0);
        }
        __syncwarp();
    }
    __syncthreads();
    unsigned long long *out = P.keys + (size_t)tile * SWR_TILE_PIXELS;
    for (int i = tid; i < SWR_TILE_PIXELS; i += RASTER_THREADS) out[i] = skeys[i];
}

// ---------------------------------------------------------------------------------------------
// key -> (record, pixel values)
// ---------------------------------------------------------------------------------------------
struct VisParams {
    const unsigned long long *keys;
    const TriRecord *records;
    const DevDraw *draws;
    uint32_t ndraws;
    int W, H, tiles_x;
};

__device__ __forceinline__ unsigned long long load_key(const unsigned long long *keys, int tiles_x, int px, int py) {
    int tile = (py >> 6) * tiles_x + (px >> 6);
    return keys[(size_t)tile * SWR_TILE_PIXELS + (py & 63) * SWR_TILE + (px & 63)];
}

__global__ void k_read_vis(VisParams P, uint32_t *depth_bits, uint32_t *seq, float *bary1, float *bary2) {
    int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= P.W || py >= P.H) return;
    unsigned long long key = load_key(P.keys, P.tiles_x, px, py);
    uint32_t db = SWR_INF_BITS, sq = 0xFFFFFFFFu;
    float b1 = 0.0f, b2 = 0.0f;
    if (key != SWR_KEY_EMPTY) {
        uint32_t slot = 0xFFFFFFFFu - (uint32_t)key;
        TriRecord r = P.records[slot];
        float z;
        if (resolve_pixel(r, P.W, P.H, px, py, b1, b2, z)) {
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

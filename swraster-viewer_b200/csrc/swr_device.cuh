// swr_device.cuh — device-side data layout and the bit-exact raster arithmetic shared by the
// set-up, tile-raster, shading and read-back kernels.
//
// Everything on the visibility path uses explicit round-to-nearest intrinsics
// (__fmul_rn/__fadd_rn/__fdiv_rn: never contracted into FMA) so that the result is the
// reference's f32 arithmetic bit for bit. The translation unit is also
// compiled with -fmad=false as a second guard.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>
#include <stdint.h>

#define SWR_TILE 64
#define SWR_TILE_PIXELS 4096
#define SWR_MAX_MIPS 16
#define SWR_CLUSTER_TRIS 128  // culling granularity: consecutive triangles of a primitive
#define SWR_NO_CLIP 0xFFFFFFFFu
#define SWR_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
#define SWR_INF_BITS 0x7F800000u
#define SWR_REC_ALPHA 0x80000000u  // TriRecord.draw bit 31: the triangle's material is alpha-tested (shader.rs:40-43)
#define SWR_REC_DRAW_MASK 0x7FFFFFFFu
#define SWR_ID_FOREIGN 0xFFFFFFFEu  // sort-last: the pixel's winner belongs to another rank

// One surviving (post cull/clip) triangle: 64 bytes, 4 x 128-bit.
//   q0 = X0 Y0 X1 Y1      28.4 fixed-point screen positions (renderer.rs:834-846)
//   q1 = X2 Y2 iw0 iw1    iw_k = 1 / clip.w_k             (renderer.rs:701-705)
//   q2 = iw2 zw0 zw1 zw2  zw_k = clip.z_k * iw_k          (renderer.rs:706-710)
//   q3 = ooa draw seq clip   ooa = 1/|area| (renderer.rs:699); draw = index into this frame's
//        draw array; seq = (first_triangle + tri) * 8 + fan; clip = first vertex of the clipped
//        polygon in the clip-vertex buffer, or SWR_NO_CLIP.
struct __align__(16) TriRecord {
    int X0, Y0, X1, Y1;
    int X2, Y2;
    float iw0, iw1;
    float iw2, zw0, zw1, zw2;
    float ooa;
    uint32_t draw, seq, clip;
};
static_assert(sizeof(TriRecord) == 64, "TriRecord must be 64 bytes");

// Post-clip vertex attributes kept for shading (renderer.rs:602-608 minus pos_clip): 48 bytes.
struct __align__(16) ClipVertex {
    float wx, wy, wz, nx;  // pos_world.xyz, normal.x
    float ny, nz, tx, ty;  // normal.yz, tangent.xy
    float tz, tw, u, v;    // tangent.zw, uv
};

struct DevDraw {
    float model[16];
    float mvp[16];
    uint32_t prim;
    uint32_t flags;
    uint32_t first_tri;  // global triangle base (seq)
    uint32_t reserved;
};

struct DevPrim {
    const float4 *pos;
    const float4 *nrm;
    const float4 *tan;
    const float2 *uv;
    const uint32_t *idx;
    const float4 *cl_sphere;  // object-space bounding sphere per cluster (k_cluster_bounds at upload)
    uint32_t nverts, ntris, material, ncl;
};

struct DevTex {
    cudaTextureObject_t obj;  // linear-memory texture object over the RGBA8 texels (point fetch)
    const uint32_t *data;
    uint32_t width, height, type, max_mip;
    uint32_t wrap_s, wrap_t;
    uint32_t mip_off[SWR_MAX_MIPS], mip_w[SWR_MAX_MIPS], mip_h[SWR_MAX_MIPS], stride[SWR_MAX_MIPS];
};

struct DevMat {
    float base[4];
    float metallic, roughness;
    float emissive[3];
    float occlusion_strength, transmission, alpha_cutoff;
    uint32_t flags;
    int tex_base, tex_mr, tex_normal, tex_emissive, tex_occlusion, tex_transmission;
};

struct DevScene {
    const DevPrim *prims;
    const DevMat *mats;
    const DevTex *texs;
    const float4 *gi;  // voxels x 4 coefficients
    uint32_t gdim[3];
    float gmin[3], gmax[3];
    float gvs[3];  // voxel size (gmax - gmin) / gdim in f32, as VoxelGrid::new computes it (voxelgrid.rs:154-161)
    int cubemap, cubemap_specular, brdf_lut;
    float light_dir[3], light_color[3];
};

struct DevCamera {
    float position[3];
    float skybox_T[16];
    float one_over_width, one_over_height;
};

struct FrameCounters {
    unsigned long long tris_clipped, tile_refs;
    // statistics spread over 32 slots (block index & 31) so that the per-block atomics do not serialise on one address
    unsigned long long tris_binned[32];
    unsigned long long refs_uncovered[32];  // refs of triangles that provably cover no pixel: counted, not emitted
    // the two bump allocators of k_clip sit in one aligned 64-bit word so that a polygon takes both with ONE atomic
    uint32_t clip_verts;      // bump allocator for ClipVertex (low word)
    uint32_t ext_records;     // bump allocator for the records of fans >= 1 (high word)
    uint32_t overflow_refs;   // tile_refs exceeded the ref buffer
    uint32_t overflow_clip;   // clip vertex buffer exhausted
    uint32_t clip_queue_n;    // triangles queued for k_clip
    uint32_t overflow_ext;    // extension records exhausted
    uint32_t clip_list_n;     // surviving fans >= 1
    uint32_t raster_units;    // entries of the raster work list
    uint32_t raster_unit_refs; // refs per unit chosen for this frame
    uint32_t raster_next;     // work-list cursor of the persistent raster CTAs
    uint32_t work_n;          // surviving clusters (k_cull)
    uint32_t cull_done;       // k_cull blocks finished (the last one scans the block counts)
    uint32_t overflow_sort;   // a tile holds more translucent packets than the in-kernel sort supports
    unsigned long long dbg[8];  // SWR_PROFILE_COUNTERS builds only
};
static_assert(offsetof(FrameCounters, clip_verts) % 8 == 0 && offsetof(FrameCounters, ext_records) == offsetof(FrameCounters, clip_verts) + 4, "k_clip's paired allocator");

// ---------------------------------------------------------------------------------------------
// exact float helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float frcp(float a) { return __frcp_rn(a); }  // correctly rounded 1/a == IEEE 1.0f / a
__device__ __forceinline__ float i2f(int a) { return __int2float_rn(a); }

// Packed fp32 pairs (sm_100+: add/mul/fma.rn.f32x2 -> FADD2 / FMUL2 / FFMA2). Each half is an ordinary IEEE round-to-nearest
// operation, so a packed operation is bit-identical to its two scalar ones; it just costs one issue slot instead of two.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long pack2(float v) { return pack2(v, v); }
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ ulonglong2 ldg_pairs(const float4 *p) { return __ldg(reinterpret_cast<const ulonglong2 *>(p)); }  // (xy, zw)

__device__ __forceinline__ int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }

// glam sse2 Mat4::mul_vec4: ((c0*x + c1*y) + c2*z) + c3*w
__device__ __forceinline__ float4 mul_vec4(const float *m, float4 v) {
    float4 r;
    r.x = fadd(fadd(fadd(fmul(m[0], v.x), fmul(m[4], v.y)), fmul(m[8], v.z)), fmul(m[12], v.w));
    r.y = fadd(fadd(fadd(fmul(m[1], v.x), fmul(m[5], v.y)), fmul(m[9], v.z)), fmul(m[13], v.w));
    r.z = fadd(fadd(fadd(fmul(m[2], v.x), fmul(m[6], v.y)), fmul(m[10], v.z)), fmul(m[14], v.w));
    r.w = fadd(fadd(fadd(fmul(m[3], v.x), fmul(m[7], v.y)), fmul(m[11], v.z)), fmul(m[15], v.w));
    return r;
}
// glam Mat3A::from_mat4(m) * Vec3A
__device__ __forceinline__ float3 mul_mat3(const float *m, float3 v) {
    float3 r;
    r.x = fadd(fadd(fmul(m[0], v.x), fmul(m[4], v.y)), fmul(m[8], v.z));
    r.y = fadd(fadd(fmul(m[1], v.x), fmul(m[5], v.y)), fmul(m[9], v.z));
    r.z = fadd(fadd(fmul(m[2], v.x), fmul(m[6], v.y)), fmul(m[10], v.z));
    return r;
}

// renderer.rs:834-846 clip_to_screen_subpixels. roundf = half away from zero (f32::round);
// __float2int_rz saturates and maps NaN to 0 exactly like Rust's `as i32`.
__device__ __forceinline__ void snap_vertex(float4 c, float Wf, float Hf, int &X, int &Y) {
    float nx = fdiv(c.x, c.w), ny = fdiv(c.y, c.w);
    // "/ 2.0" (renderer.rs:838-839) as "* 0.5": identical bits for every input (a power-of-two scale is exact in binary
    // floating point, subnormal results included), without the IEEE division sequence
    float sx = fmul(fmul(fadd(nx, 1.0f), Wf), 0.5f);
    float sy = fmul(fmul(fsub(1.0f, ny), Hf), 0.5f);
    X = __float2int_rz(roundf(fmul(sx, 16.0f)));
    Y = __float2int_rz(roundf(fmul(sy, 16.0f)));
}

// Order-preserving map f32 -> u32 for finite/inf values; -0 and +0 share one code so that the
// reference's `<=` tie (later packet wins) is kept for mixed zeros.
__device__ __forceinline__ uint32_t depth_orderable(float z) {
    uint32_t b = __float_as_uint(z);
    if ((b << 1) == 0u) b = 0u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// ---------------------------------------------------------------------------------------------
// per-(triangle, tile) packet set-up: tilerasterizer.rs:114-218 on top of renderer.rs:685-695,785-786
// ---------------------------------------------------------------------------------------------
struct PacketSetup {
    int a[3], b[3], c[3];  // index 0: edge v1->v2 (w0), 1: edge v2->v0 (w1 -> bary1), 2: edge v0->v1 (w2 -> bary2)
    int xs, ys, xe, ye;    // evaluated region in sub-pixels: start even-aligned, end exclusive
    int nqx, nqy;          // quads in the region
    bool coarse;           // clipped bbox wider/taller than 16 px: 16x16 block reject first
    bool exact;            // every f32 edge value in the region is an exactly representable integer
    bool empty;
};

__device__ __forceinline__ int top_left_bias(int a, int b) { return (a < 0 || (a == 0 && b > 0)) ? 0 : -1; }

__device__ __forceinline__ void tri_bbox_pixels(const TriRecord &r, int W, int H, int &bminx, int &bminy, int &bmaxx, int &bmaxy) {
    int mnx = max(min(min(r.X0, r.X1), r.X2), 0), mny = max(min(min(r.Y0, r.Y1), r.Y2), 0);
    int mxx = min(max(max(r.X0, r.X1), r.X2), W * 16), mxy = min(max(max(r.Y0, r.Y1), r.Y2), H * 16);
    bminx = mnx >> 4;
    bminy = mny >> 4;
    bmaxx = wadd(mxx, 16) >> 4;
    bmaxy = wadd(mxy, 16) >> 4;
}

// |a| * xhi + |b| * yhi + |c| < 2^24. Fast path in 32 bits: xhi, yhi < 2^18 + 256 (screens up to 255 tiles of 1024 sub-pixel
// units), so with |a|, |b| < 2^12 and |c| < 2^24 the sum stays below 2^32; anything larger takes the 64-bit form (same value).
__device__ __forceinline__ bool edge_bound_ok(int a, int b, int c, uint32_t xhi, uint32_t yhi) {
    const uint32_t ua = (uint32_t)(a < 0 ? -(long long)a : a), ub = (uint32_t)(b < 0 ? -(long long)b : b), uc = (uint32_t)(c < 0 ? -(long long)c : c);
    if ((ua | ub) >= 4096u || uc >= (1u << 24)) {
        return (long long)ua * xhi + (long long)ub * yhi + (long long)uc < (1ll << 24);
    }
    return ua * xhi + ub * yhi + uc < (1u << 24);
}

__device__ __forceinline__ void packet_setup(const TriRecord &r, int W, int H, int tile_x0, int tile_y0, PacketSetup &p) {
    int bminx, bminy, bmaxx, bmaxy;
    tri_bbox_pixels(r, W, H, bminx, bminy, bmaxx, bmaxy);
    int cminx = max(bminx, tile_x0), cminy = max(bminy, tile_y0);
    int cmaxx = min(bmaxx, tile_x0 + SWR_TILE), cmaxy = min(bmaxy, tile_y0 + SWR_TILE);
    int wpx = cmaxx - cminx, hpx = cmaxy - cminy;
    p.empty = (wpx < 1 || hpx < 1);
    p.coarse = (wpx > 16 || hpx > 16);
    p.xs = (cminx & ~1) * 16;
    p.ys = (cminy & ~1) * 16;
    p.xe = cmaxx * 16;
    p.ye = cmaxy * 16;
    p.nqx = (p.xe - p.xs + 31) >> 5;
    p.nqy = (p.ye - p.ys + 31) >> 5;
    // tilerasterizer.rs:143-156
    int a01 = wsub(r.Y1, r.Y0), b01 = wsub(r.X0, r.X1);
    int c01 = wadd(wsub(wmul(r.X1, r.Y0), wmul(r.X0, r.Y1)), top_left_bias(a01, b01));
    int a12 = wsub(r.Y2, r.Y1), b12 = wsub(r.X1, r.X2);
    int c12 = wadd(wsub(wmul(r.X2, r.Y1), wmul(r.X1, r.Y2)), top_left_bias(a12, b12));
    int a20 = wsub(r.Y0, r.Y2), b20 = wsub(r.X2, r.X0);
    int c20 = wadd(wsub(wmul(r.X0, r.Y2), wmul(r.X2, r.Y0)), top_left_bias(a20, b20));
    p.a[0] = a12; p.b[0] = b12; p.c[0] = c12;
    p.a[1] = a20; p.b[1] = b20; p.c[1] = c20;
    p.a[2] = a01; p.b[2] = b01; p.c[2] = c01;
    // Exactness bound: all evaluated coordinates are positive and <= (xhi, yhi); every partial sum of the f32
    // chain is then bounded by |a|*xhi + |b|*yhi + |c|. Below 2^24 all of them are exact integers, so the chain
    // equals the integer edge function and the coarse reject can never drop a covered pixel.
    int xhi = p.coarse ? p.xs + (((p.xe - p.xs) + 255) >> 8) * 256 : p.xs + p.nqx * 32;
    int yhi = p.coarse ? p.ys + (((p.ye - p.ys) + 255) >> 8) * 256 : p.ys + p.nqy * 32;
    bool ex = true;
#pragma unroll
    for (int e = 0; e < 3; e++) ex = ex && edge_bound_ok(p.a[e], p.b[e], p.c[e], (uint32_t)xhi, (uint32_t)yhi);
    p.exact = ex;
}

// Edge value of one pixel of a NON-exact packet, replaying the reference's f32 chain:
// origin value at the fine region's first quad, `j` row steps, then `i` column steps
// (tilerasterizer.rs:321-323, 371-373, 378-380). (qx,qy) = quad inside the packet region, (lx,ly) = lane.
// Returns false when the quad's 16x16 block is rejected by the coarse test (tilerasterizer.rs:239-269).
__device__ __forceinline__ bool eval_chain(const PacketSetup &p, int qx, int qy, int lx, int ly, float r[3]) {
    int bx = p.xs, by = p.ys, i = qx, j = qy;
    float A[3], B[3], C[3];
#pragma unroll
    for (int e = 0; e < 3; e++) {
        A[e] = i2f(p.a[e]);
        B[e] = i2f(p.b[e]);
        C[e] = i2f(p.c[e]);
    }
    if (p.coarse) {
        bx = p.xs + (qx >> 3) * 256;
        by = p.ys + (qy >> 3) * 256;
        i = qx & 7;
        j = qy & 7;
        float cx0 = i2f(bx + 8), cx1 = i2f(bx + 248), cy0 = i2f(by + 8), cy1 = i2f(by + 248);
#pragma unroll
        for (int e = 0; e < 3; e++) {
            float ax0 = fmul(A[e], cx0), ax1 = fmul(A[e], cx1), by0 = fmul(B[e], cy0), by1 = fmul(B[e], cy1);
            float e00 = fadd(fadd(ax0, by0), C[e]), e10 = fadd(fadd(ax1, by0), C[e]);
            float e01 = fadd(fadd(ax0, by1), C[e]), e11 = fadd(fadd(ax1, by1), C[e]);
            // max_element via _mm_max_ps: comparisons only matter through `< 0`; NaN cannot occur (finite ints)
            float m = fmaxf(fmaxf(e00, e10), fmaxf(e01, e11));
            if (m < 0.0f) return false;
        }
    }
    float x = i2f(bx + 8 + 16 * lx), y = i2f(by + 8 + 16 * ly);
#pragma unroll
    for (int e = 0; e < 3; e++) {
        float v = fadd(fadd(fmul(A[e], x), fmul(B[e], y)), C[e]);
        float sy = i2f(wmul(p.b[e], 32)), sx = i2f(wmul(p.a[e], 32));
        for (int k = 0; k < j; k++) v = fadd(v, sy);
        for (int k = 0; k < i; k++) v = fadd(v, sx);
        r[e] = v;
    }
    return true;
}

// Exact packets: the integer edge function at the pixel centre (16*px + 8, 16*py + 8).
__device__ __forceinline__ void eval_exact(const PacketSetup &p, int sx, int sy, int e[3]) {
#pragma unroll
    for (int k = 0; k < 3; k++) e[k] = p.a[k] * sx + p.b[k] * sy + p.c[k];
}

// tilerasterizer.rs:337-342 with util.rs:169-171: barycentrics, perspective w and depth.
__device__ __forceinline__ float frag_depth(const TriRecord &r, float w1, float w2, float &b1, float &b2) {
    b1 = fmul(w1, r.ooa);
    b2 = fmul(w2, r.ooa);
    float iwda = fsub(r.iw1, r.iw0), iwdb = fsub(r.iw2, r.iw0);
    float zwda = fsub(r.zw1, r.zw0), zwdb = fsub(r.zw2, r.zw0);
    float q = fadd(fadd(r.iw0, fmul(b1, iwda)), fmul(b2, iwdb));
    float wp = frcp(q);
    float zz = fadd(fadd(r.zw0, fmul(b1, zwda)), fmul(b2, zwdb));
    return fmul(zz, wp);
}

// Edge values w1, w2 of a pixel its packet is KNOWN to cover (the key's owner): same chain as eval_chain, but neither the
// coarse reject nor edge 0 (both only decide coverage) is evaluated.
__device__ __forceinline__ void eval_chain_owner(const PacketSetup &p, int qx, int qy, int lx, int ly, float &w1, float &w2) {
    int bx = p.xs, by = p.ys, i = qx, j = qy;
    if (p.coarse) {
        bx = p.xs + (qx >> 3) * 256;
        by = p.ys + (qy >> 3) * 256;
        i = qx & 7;
        j = qy & 7;
    }
    const float x = i2f(bx + 8 + 16 * lx), y = i2f(by + 8 + 16 * ly);
    // both edges walk the same j row steps and i column steps: one loop each (the per-edge order of additions is the reference's)
    float t1 = fadd(fadd(fmul(i2f(p.a[1]), x), fmul(i2f(p.b[1]), y)), i2f(p.c[1]));
    float t2 = fadd(fadd(fmul(i2f(p.a[2]), x), fmul(i2f(p.b[2]), y)), i2f(p.c[2]));
    const float sy1 = i2f(wmul(p.b[1], 32)), sx1 = i2f(wmul(p.a[1], 32));
    const float sy2 = i2f(wmul(p.b[2], 32)), sx2 = i2f(wmul(p.a[2], 32));
    for (int k = 0; k < j; k++) {
        t1 = fadd(t1, sy1);
        t2 = fadd(t2, sy2);
    }
    for (int k = 0; k < i; k++) {
        t1 = fadd(t1, sx1);
        t2 = fadd(t2, sx2);
    }
    w1 = t1;
    w2 = t2;
}

// Barycentrics and depth of a pixel for the record that owns its key (shading): no coverage re-test.
__device__ __forceinline__ float resolve_owner(const TriRecord &r, int W, int H, int px, int py, float &b1, float &b2) {
    PacketSetup p;
    packet_setup(r, W, H, px & ~(SWR_TILE - 1), py & ~(SWR_TILE - 1), p);
    float w1, w2;
    if (p.exact) {
        const int sx = px * 16 + 8, sy = py * 16 + 8;
        w1 = i2f(p.a[1] * sx + p.b[1] * sy + p.c[1]);
        w2 = i2f(p.a[2] * sx + p.b[2] * sy + p.c[2]);
    } else {
        eval_chain_owner(p, ((px & ~1) * 16 - p.xs) >> 5, ((py & ~1) * 16 - p.ys) >> 5, px & 1, py & 1, w1, w2);
    }
    return frag_depth(r, w1, w2, b1, b2);
}

// Full re-evaluation of one pixel against one record (used after the key buffer is final: shading and
// read-back). Returns false if the pixel is not covered by this packet (cannot happen for a key's owner).
__device__ __forceinline__ bool resolve_pixel(const TriRecord &r, int W, int H, int px, int py, float &b1, float &b2, float &z) {
    PacketSetup p;
    packet_setup(r, W, H, px & ~(SWR_TILE - 1), py & ~(SWR_TILE - 1), p);
    float w1, w2;
    if (p.exact) {
        int e[3];
        eval_exact(p, px * 16 + 8, py * 16 + 8, e);
        if ((e[0] | e[1] | e[2]) < 0) return false;
        w1 = i2f(e[1]);
        w2 = i2f(e[2]);
    } else {
        int lx = px & 1, ly = py & 1;
        int qx = ((px & ~1) * 16 - p.xs) >> 5, qy = ((py & ~1) * 16 - p.ys) >> 5;
        float rr[3];
        if (!eval_chain(p, qx, qy, lx, ly, rr)) return false;
        if (!(rr[0] >= 0.0f && rr[1] >= 0.0f && rr[2] >= 0.0f)) return false;
        w1 = rr[1];
        w2 = rr[2];
    }
    z = frag_depth(r, w1, w2, b1, b2);
    return true;
}

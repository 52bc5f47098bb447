// oracle.cpp — CPU restatement of swraster-viewer's per-frame rasterisation path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (swraster-viewer_b200/)
// may include, link or call this file; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it, as the checker or as
// the timed CPU baseline.
//
// PARITY UNPINNED: the reference ships no tests, golden images or known-answer
// vectors, and it cannot be compiled here (no cargo/rustc, crates
// not vendored).  This file therefore restates the Rust source operation for
// operation; every function cites the file:line it follows.  The arithmetic of the
// un-vendored dependency glam 0.30.10 (SSE2 backend) is restated from its
// published source: mul_vec4 = ((c0*x + c1*y) + c2*z) + c3*w, dot4 =
// (x*x'+z*z') + (y*y'+w*w'), dot3 = (x*x'+y*y') + z*z', min/max = _mm_min/max_ps,
// Vec4::round = round-half-even, casts saturate.  Build with -ffp-contract=off.
//
// Two schedules:
//   * serial (nthreads = 1): the reference's RAYON_NUM_THREADS=1 execution; packets
//     reach each tile queue in submission order. This is the parity oracle.
//   * parallel (nthreads > 1): same algorithm with the reference's parallel
//     structure (triangles -> shared per-tile queues of full packets, one task per
//     tile, row-pair resolve). This is the timed CPU baseline.
#include <immintrin.h>
#include <sys/mman.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/swr.h"

namespace {

// ---------------------------------------------------------------------------------
// 4-lane vector = glam Vec4 on SSE2
// ---------------------------------------------------------------------------------
struct V4 {
    __m128 v;
    V4() : v(_mm_setzero_ps()) {}
    V4(__m128 x) : v(x) {}
    V4(float a, float b, float c, float d) : v(_mm_setr_ps(a, b, c, d)) {}
    float operator[](int i) const {
        alignas(16) float t[4];
        _mm_store_ps(t, v);
        return t[i];
    }
};
static inline V4 splat(float x) { return V4(_mm_set1_ps(x)); }
static inline V4 operator+(V4 a, V4 b) { return _mm_add_ps(a.v, b.v); }
static inline V4 operator-(V4 a, V4 b) { return _mm_sub_ps(a.v, b.v); }
static inline V4 operator*(V4 a, V4 b) { return _mm_mul_ps(a.v, b.v); }
static inline V4 operator/(V4 a, V4 b) { return _mm_div_ps(a.v, b.v); }
static inline V4 operator+(V4 a, float b) { return a + splat(b); }
static inline V4 operator-(V4 a, float b) { return a - splat(b); }
static inline V4 operator*(V4 a, float b) { return a * splat(b); }
static inline V4 operator/(V4 a, float b) { return a / splat(b); }
static inline V4 operator+(float a, V4 b) { return splat(a) + b; }
static inline V4 operator-(float a, V4 b) { return splat(a) - b; }
static inline V4 operator*(float a, V4 b) { return splat(a) * b; }
static inline V4 operator/(float a, V4 b) { return splat(a) / b; }
static inline V4 operator-(V4 a) { return _mm_xor_ps(a.v, _mm_set1_ps(-0.0f)); }  // glam Neg
static inline V4 vmin(V4 a, V4 b) { return _mm_min_ps(a.v, b.v); }
static inline V4 vmax(V4 a, V4 b) { return _mm_max_ps(a.v, b.v); }
static inline V4 vclamp(V4 a, V4 lo, V4 hi) { return vmin(vmax(a, lo), hi); }  // glam clamp = max then min
static inline V4 vabs(V4 a) { return _mm_andnot_ps(_mm_set1_ps(-0.0f), a.v); }
static inline V4 vfloor(V4 a) { return _mm_floor_ps(a.v); }
// glam sse2 Vec4::round: magic-number round under RNE = half to even.
static inline V4 vround(V4 a) { return _mm_round_ps(a.v, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC); }
typedef __m128 M4;  // BVec4A
static inline M4 cmpge(V4 a, V4 b) { return _mm_cmpge_ps(a.v, b.v); }
static inline M4 cmpgt(V4 a, V4 b) { return _mm_cmpgt_ps(a.v, b.v); }
static inline M4 cmple(V4 a, V4 b) { return _mm_cmple_ps(a.v, b.v); }
static inline M4 cmpne(V4 a, V4 b) { return _mm_cmpneq_ps(a.v, b.v); }
static inline M4 mand(M4 a, M4 b) { return _mm_and_ps(a, b); }
static inline M4 mandnot(M4 nota, M4 b) { return _mm_andnot_ps(nota, b); }
static inline int bitmask(M4 m) { return _mm_movemask_ps(m); }
static inline bool many(M4 m) { return bitmask(m) != 0; }
static inline bool mall(M4 m) { return bitmask(m) == 0xF; }
static inline V4 select(M4 m, V4 a, V4 b) { return _mm_or_ps(_mm_and_ps(m, a.v), _mm_andnot_ps(m, b.v)); }
static inline float max_element(V4 a) {
    __m128 t = _mm_max_ps(a.v, _mm_shuffle_ps(a.v, a.v, _MM_SHUFFLE(0, 0, 3, 2)));
    t = _mm_max_ps(t, _mm_shuffle_ps(t, t, _MM_SHUFFLE(0, 0, 0, 1)));
    return _mm_cvtss_f32(t);
}

static bool g_exact_rsqrt = false;
// math.rs:34-39 (x86_64): _mm_rsqrt_ps. The exact variant exists only so tests can
// separate "approximation envelope" from logic differences.
static inline V4 rsqrt_vec(V4 a) {
    if (g_exact_rsqrt) return _mm_div_ps(_mm_set1_ps(1.0f), _mm_sqrt_ps(a.v));
    return _mm_rsqrt_ps(a.v);
}

struct U4 {
    uint32_t v[4];
    uint32_t operator[](int i) const { return v[i]; }
};
// Rust `as u32` (saturating, NaN -> 0) per lane: glam as_uvec4.
static inline uint32_t f2u(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;
}
static inline int32_t f2i(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}
static inline U4 as_uvec4(V4 a) {
    alignas(16) float t[4];
    _mm_store_ps(t, a.v);
    U4 r;
    for (int i = 0; i < 4; i++) r.v[i] = f2u(t[i]);
    return r;
}

// math.rs:52-171 Vec3x4
struct V3x4 {
    V4 x, y, z;
};
static inline V3x4 v3splat(float s) { return {splat(s), splat(s), splat(s)}; }
static inline V3x4 operator+(V3x4 a, V3x4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3x4 operator-(V3x4 a, V3x4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3x4 operator*(V3x4 a, V3x4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3x4 operator*(V3x4 a, V4 b) { return {a.x * b, a.y * b, a.z * b}; }
static inline V3x4 operator*(V3x4 a, float b) { return {a.x * b, a.y * b, a.z * b}; }
static inline V3x4 operator/(V3x4 a, V3x4 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
static inline V3x4 operator+(V3x4 a, float b) { return {a.x + b, a.y + b, a.z + b}; }
static inline V3x4 operator+(V3x4 a, V4 b) { return {a.x + b, a.y + b, a.z + b}; }
static inline V4 dot(V3x4 a, V3x4 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // math.rs:88-90
static inline V3x4 normalize(V3x4 a) {                                             // math.rs:101-108
    V4 r = rsqrt_vec(dot(a, a));
    return {a.x * r, a.y * r, a.z * r};
}
static inline V3x4 cross(V3x4 a, V3x4 b) {  // math.rs:110-116
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline V3x4 reflect(V3x4 a, V3x4 n) { return a - n * dot(a, n) * 2.0f; }  // math.rs:118-120
static inline V3x4 v3max(V3x4 a, V4 m) { return {vmax(a.x, m), vmax(a.y, m), vmax(a.z, m)}; }
static inline V3x4 v3select(M4 m, V3x4 a, V3x4 b) {
    return {select(m, a.x, b.x), select(m, a.y, b.y), select(m, a.z, b.z)};
}
static inline V3x4 v3lerp(V3x4 a, V3x4 b, V4 t) {  // math.rs:160-166
    return {a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t};
}

// ---------------------------------------------------------------------------------
// scalar glam pieces used per triangle
// ---------------------------------------------------------------------------------
struct F4 {
    float x, y, z, w;
};
struct F3 {
    float x, y, z;
};
struct F2 {
    float x, y;
};
// glam sse2 Mat4::mul_vec4
static inline F4 mul_vec4(const float *m, F4 v) {
    F4 r;
    r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
    r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
    r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
    r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
    return r;
}
// glam Mat4 * Mat4: each column of rhs through mul_vec4
static inline void mul_mat4(const float *a, const float *b, float *out) {
    for (int c = 0; c < 4; c++) {
        F4 r = mul_vec4(a, F4{b[c * 4 + 0], b[c * 4 + 1], b[c * 4 + 2], b[c * 4 + 3]});
        out[c * 4 + 0] = r.x;
        out[c * 4 + 1] = r.y;
        out[c * 4 + 2] = r.z;
        out[c * 4 + 3] = r.w;
    }
}
// glam Mat3A::from_mat4(m) * Vec3A
static inline F3 mul_mat3(const float *m, F3 v) {
    F3 r;
    r.x = (m[0] * v.x + m[4] * v.y) + m[8] * v.z;
    r.y = (m[1] * v.x + m[5] * v.y) + m[9] * v.z;
    r.z = (m[2] * v.x + m[6] * v.y) + m[10] * v.z;
    return r;
}
static inline float dot4(F4 a, F4 b) { return (a.x * b.x + a.z * b.z) + (a.y * b.y + a.w * b.w); }
static inline float dot3(F3 a, F3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

// util.rs:149-194
struct ScalarInterp {
    float a, da, db;
    void set(float v0, float v1, float v2) {
        a = v0;
        da = v1 - v0;
        db = v2 - v0;
    }
    V4 interpolate(V4 b1, V4 b2) const { return a + b1 * da + b2 * db; }
};
struct Vec3Interp {
    F3 a, da, db;
    void set(F3 v0, F3 v1, F3 v2) {
        a = v0;
        da = F3{v1.x - v0.x, v1.y - v0.y, v1.z - v0.z};
        db = F3{v2.x - v0.x, v2.y - v0.y, v2.z - v0.z};
    }
    V3x4 interpolate(V4 b1, V4 b2) const {
        return {a.x + b1 * da.x + b2 * db.x, a.y + b1 * da.y + b2 * db.y, a.z + b1 * da.z + b2 * db.z};
    }
};

// renderer.rs:31-37
struct Vertex {
    F4 pos_clip;
    F3 pos_world;
    F3 normal;
    F4 tangent;
    F2 uv;
};

// renderer.rs:59-76 (plus the global ids the oracle reports)
struct RasterPacket {
    int32_t min_x, min_y, max_x, max_y;  // screen_min_pixels / screen_max_pixels
    int32_t px[3], py[3];                // pos_screen_subpixels
    ScalarInterp z_over_w, one_over_w;
    Vec3Interp normals, tangents;
    ScalarInterp tangent_sign, u_over_w, v_over_w;
    Vec3Interp pos_world_over_w;
    float one_over_area;
    uint32_t mesh_index, primitive_index;  // primitive_index = global primitive id here
    float avg_z;
    float du_dv[4];
    uint32_t seq;  // oracle-only: (first_triangle + tri) * 8 + fan
    uint32_t pad[3];
};

// bumpqueue.rs:73-233 semantics: concurrent append (fetch_add slot claim, blocks of
// 1024 from a shared pool), indexed get, reset. Blocks are kept across frames.
// Block storage for the packet queues: one process-wide bump arena of 2 MiB-aligned, huge-page-advised chunks — the
// stand-in for the reference's shared BumpPool (bumpqueue.rs:38-63). Blocks are recycled per queue across frames, so
// after the first frame no allocation happens; huge pages keep the ~1 GB/frame packet traffic off the dTLB.
struct BlockArena {
    std::mutex m;
    char *cur = nullptr, *end = nullptr;
    size_t nallocs = 0;
    void *alloc(size_t bytes) {
        std::lock_guard<std::mutex> g(m);
        bytes = (bytes + 63) & ~size_t(63);
        if (cur == nullptr || cur + bytes + 4096 > end) {
            size_t chunk = std::max<size_t>(bytes, size_t(256) << 20);
            chunk = (chunk + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);
            void *p = aligned_alloc(size_t(2) << 20, chunk);
            if (!p) abort();
#ifdef MADV_HUGEPAGE
            madvise(p, chunk, MADV_HUGEPAGE);
#endif
            cur = (char *)p;
            end = cur + chunk;
        }
        void *r = cur;
        // skew consecutive blocks by an odd number of cache lines: 256 KiB blocks at a power-of-two stride would put every
        // tile's write front into the same L2 sets (measured: clip/bin 2x slower), which ordinary heap allocations avoid
        cur += bytes + 64 * (size_t)(1 + (nallocs++ * 37) % 61);
        return r;
    }
};
static BlockArena g_arena;

// bumpqueue.rs:73-233 semantics: concurrent append (fetch_add slot claim, blocks of 1024), indexed get, reset.
struct PacketQueue {
    static const uint32_t BLOCK = 1024;  // bumpqueue.rs:8
    static const uint32_t MAX_BLOCKS = 2048;
    // read-mostly fields and the contended slot counter live on separate cache lines (no false sharing between the
    // fetch_add traffic and the block-table reads), which is the best case for the reference's push path
    alignas(64) RasterPacket **blocks;  // fixed table (boxcar::Vec stand-in): readers never see a reallocation
    std::atomic<uint32_t> blocks_ready{0};
    alignas(64) std::atomic<uint32_t> count{0};
    alignas(64) std::mutex grow;
    PacketQueue() {
        blocks = (RasterPacket **)g_arena.alloc(MAX_BLOCKS * sizeof(RasterPacket *));
        memset(blocks, 0, MAX_BLOCKS * sizeof(RasterPacket *));
    }
    void push(const RasterPacket &p) {  // bumpqueue.rs:96-112
        uint32_t slot = count.fetch_add(1, std::memory_order_relaxed);
        uint32_t bi = slot / BLOCK;
        if (bi >= MAX_BLOCKS) abort();
        if (bi >= blocks_ready.load(std::memory_order_acquire)) {
            std::lock_guard<std::mutex> g(grow);
            uint32_t have = blocks_ready.load(std::memory_order_relaxed);
            while (have <= bi) blocks[have++] = (RasterPacket *)g_arena.alloc(sizeof(RasterPacket) * BLOCK);
            blocks_ready.store(have, std::memory_order_release);
        }
        blocks[bi][slot % BLOCK] = p;
    }
    RasterPacket get(uint32_t i) const { return blocks[i / BLOCK][i % BLOCK]; }  // :114-120 copies the packet out
    const RasterPacket &ref(uint32_t i) const { return blocks[i / BLOCK][i % BLOCK]; }
    uint32_t len() const { return count.load(std::memory_order_relaxed); }
    void reset() { count.store(0); }  // blocks are recycled across frames (bumpqueue.rs:124-148)
};

struct Tile {  // tilerasterizer.rs:25-38
    int32_t min_x, min_y, max_x, max_y;
    PacketQueue packets_opaque;
    PacketQueue packets_translucent;
    std::vector<V3x4> color;
    std::vector<V4> depth;
    std::vector<U4> packet_index;
    std::vector<V4> bary1, bary2;
    std::vector<U4> written;  // oracle bookkeeping only: lane received an id/bary write this frame
    float center_luminance = 1.0f;
};

struct Draw {
    float model[16], mvp[16];
    uint32_t primitive, mesh, flags, first_triangle;
};

// Frame counters. The reference keeps none on this path (renderer.rs:668-846 pushes packets and nothing else), so they
// must not cost the timed parallel bin phase anything: one cache-line-sized slot per worker thread, folded after the
// phase (a shared atomic here would serialise the workers on one line and handicap the CPU baseline).
struct alignas(64) ThreadStats {
    uint64_t tris_binned = 0, tris_clipped = 0, packets = 0, packets_dup = 0;
    uint64_t pad[4];
};
struct Stats {
    uint64_t tris_binned = 0, tris_clipped = 0, packets = 0, packets_dup = 0;
};
static const int MAX_THREADS = 1024;

struct Oracle {
    int W, H, tiles_x, tiles_y;
    const swr_scene_desc *scene;
    const swr_camera *cam;
    std::vector<Tile *> tiles;
    std::vector<Draw> draws;
    Stats stats;
    std::vector<ThreadStats> tstats;  // indexed by omp_get_thread_num()
    bool track_written = true;        // per-lane "written" bookkeeping is only needed when the caller reads seq back
    int nthreads;
    uint64_t tris_submitted = 0, verts_submitted = 0;
    double ms_clipbin = 0, ms_raster = 0;
    ~Oracle() {
        for (auto *t : tiles) delete t;
    }
};

// ---------------------------------------------------------------------------------
// util.rs / texture.rs sampling
// ---------------------------------------------------------------------------------
static inline F4 rgba8_unpack(uint32_t p) {  // util.rs:83-89
    return F4{(float)((p >> 24) & 0xFF) / 255.0f, (float)((p >> 16) & 0xFF) / 255.0f, (float)((p >> 8) & 0xFF) / 255.0f,
              (float)(p & 0xFF) / 255.0f};
}
static inline V3x4 srgb_to_linear_fast(V3x4 x) {  // util.rs:111-113
    return x * (x * (x * 0.305306011f + 0.682171111f) + 0.012522878f);
}
static inline V3x4 gather_rgb(const swr_texture_desc &t, U4 idx) {  // texture.rs:716-727
    F4 a = rgba8_unpack(t.data[idx[0]]), b = rgba8_unpack(t.data[idx[1]]), c = rgba8_unpack(t.data[idx[2]]),
       d = rgba8_unpack(t.data[idx[3]]);
    return {V4(a.x, b.x, c.x, d.x), V4(a.y, b.y, c.y, d.y), V4(a.z, b.z, c.z, d.z)};
}
static inline V4 gather_alpha(const swr_texture_desc &t, U4 idx) {  // texture.rs:792-799
    return V4(rgba8_unpack(t.data[idx[0]]).w, rgba8_unpack(t.data[idx[1]]).w, rgba8_unpack(t.data[idx[2]]).w,
              rgba8_unpack(t.data[idx[3]]).w);
}
static inline V4 apply_wrap_mode(V4 texel, V4 dim, uint32_t mode) {  // texture.rs:578-589
    V4 t2 = texel;
    if (mode == SWR_WRAP_REPEAT) {
        t2 = texel - vfloor(texel / dim) * dim;
    } else if (mode == SWR_WRAP_MIRRORED_REPEAT) {
        V4 two = dim * splat(2.0f);
        V4 t = texel - vfloor(texel / two) * two;
        t2 = vmin(t, two - t);
    }
    return vmin(t2, dim - splat(1.0f));
}
static inline uint32_t ilog2_u32(uint32_t v) { return 31u - (uint32_t)__builtin_clz(v); }
static inline uint32_t compute_mip_level(const swr_texture_desc &t, const float du_dv[4]) {  // texture.rs:851-863
    float wh[4] = {(float)t.width, (float)t.width, (float)t.height, (float)t.height};
    float d0 = du_dv[0] * wh[0], d1 = du_dv[1] * wh[1], d2 = du_dv[2] * wh[2], d3 = du_dv[3] * wh[3];
    float dx2 = d0 * d0 + d2 * d2;
    float dy2 = d1 * d1 + d3 * d3;
    float footprint = (dx2 + dy2) * 0.5f;
    float fm = (footprint > 1.0f) ? footprint : 1.0f;  // f32::max(1.0): NaN -> 1.0
    uint32_t mip = ilog2_u32(f2u(fm)) >> 1;
    return mip < t.max_mip_level ? mip : t.max_mip_level;
}
static inline U4 sample4_index(const swr_texture_desc &t, V4 u, V4 v, const float du_dv[4]) {  // texture.rs:680-694
    uint32_t mip = compute_mip_level(t, du_dv);
    uint32_t wi = t.mip_widths[mip];
    V4 wf = splat((float)t.mip_widths[mip]), hf = splat((float)t.mip_heights[mip]);
    uint32_t off = t.mip_offsets[mip];
    U4 x = as_uvec4(apply_wrap_mode(vfloor(u * wf), wf, t.wrap_s));
    U4 y = as_uvec4(apply_wrap_mode(vfloor(v * hf), hf, t.wrap_t));
    U4 idx;
    for (int i = 0; i < 4; i++) idx.v[i] = off + y[i] * wi + x[i];
    return idx;
}
static inline V3x4 sample4_rgb(const swr_texture_desc &t, V4 u, V4 v, const float du_dv[4]) {
    return gather_rgb(t, sample4_index(t, u, v, du_dv));
}
static inline V4 sample4_alpha(const swr_texture_desc &t, V4 u, V4 v, const float du_dv[4]) {  // texture.rs:698-714
    return gather_alpha(t, sample4_index(t, u, v, du_dv));
}
// texture.rs:593-644
static inline void cubemap_uv_from_normal(V3x4 n, V4 &uf, V4 &vf, U4 &slice) {
    V4 ax = vabs(n.x), ay = vabs(n.y), az = vabs(n.z);
    M4 mx = mand(cmpge(ax, ay), cmpge(ax, az));
    M4 my = mand(cmpgt(ay, ax), cmpge(ay, az));
    V4 zero = splat(0.0f), one = splat(1.0f), neg = splat(-1.0f);
    V4 sx = select(cmpge(n.x, zero), one, neg);
    V4 sy = select(cmpge(n.y, zero), one, neg);
    V4 sz = select(cmpge(n.z, zero), one, neg);
    V4 u_x = -n.z * sx, v_x = -n.y, d_x = ax;
    V4 u_y = n.x, v_y = n.z * sy, d_y = ay;
    V4 u_z = n.x * sz, v_z = -n.y, d_z = az;
    V4 u = select(mx, u_x, select(my, u_y, u_z));
    V4 v = select(mx, v_x, select(my, v_y, v_z));
    V4 denom = select(mx, d_x, select(my, d_y, d_z));
    denom = vmax(denom, splat(1.0e-19f));
    uf = 0.5f * (u / denom + 1.0f);
    vf = 0.5f * (v / denom + 1.0f);
    V4 idx_x = select(cmpge(n.x, zero), zero, one);
    V4 idx_y = select(cmpge(n.y, zero), splat(2.0f), splat(3.0f));
    V4 idx_z = select(cmpge(n.z, zero), splat(4.0f), splat(5.0f));
    slice = as_uvec4(select(mx, idx_x, select(my, idx_y, idx_z)));
}
static inline V3x4 sample_cubemap_rgb(const swr_texture_desc &t, V3x4 normal, U4 mip) {  // texture.rs:646-663
    V4 u, v;
    U4 slice;
    cubemap_uv_from_normal(normal, u, v, slice);
    float wf_[4], hf_[4];
    uint32_t wi[4], off[4];
    for (int i = 0; i < 4; i++) {
        uint32_t m = mip[i] < t.max_mip_level ? mip[i] : t.max_mip_level;
        wi[i] = t.mip_widths[m];
        wf_[i] = (float)t.mip_widths[m];
        hf_[i] = (float)t.mip_heights[m];
        off[i] = t.mip_offsets[m] + slice[i] * t.array_stride[m];
    }
    V4 wf(wf_[0], wf_[1], wf_[2], wf_[3]), hf(hf_[0], hf_[1], hf_[2], hf_[3]);
    V4 one = splat(1.0f), zero = splat(0.0f);
    V4 x = vround(u * (wf - one));
    V4 y = vround(v * (hf - one));
    U4 xi = as_uvec4(vclamp(x, zero, wf - one));
    U4 yi = as_uvec4(vclamp(y, zero, hf - one));
    U4 idx;
    for (int i = 0; i < 4; i++) idx.v[i] = off[i] + yi[i] * wi[i] + xi[i];
    return gather_rgb(t, idx);
}
static inline V3x4 sample_cubemap_trilinear_rgb(const swr_texture_desc &t, V3x4 normal, V4 mip_level) {  // :665-678
    V4 maxv = splat((float)t.max_mip_level);
    V4 mip = vclamp(mip_level, splat(0.0f), maxv);
    V4 mip0f = vfloor(mip);
    V4 mip1f = vmin(mip0f + splat(1.0f), maxv);
    V4 tt = mip - mip0f;
    V3x4 c0 = sample_cubemap_rgb(t, normal, as_uvec4(mip0f));
    V3x4 c1 = sample_cubemap_rgb(t, normal, as_uvec4(mip1f));
    return {c0.x + (c1.x - c0.x) * tt, c0.y + (c1.y - c0.y) * tt, c0.z + (c1.z - c0.z) * tt};
}
// texture.rs:730-790 with mip_level = 0, array_slice = 0 (the only call site on the path: shader.rs:250-255)
static inline V3x4 sample_bilinear_rgb0(const swr_texture_desc &t, V4 u, V4 v) {
    uint32_t wi = t.mip_widths[0];
    V4 wf = splat((float)t.mip_widths[0]), hf = splat((float)t.mip_heights[0]);
    uint32_t off = t.mip_offsets[0];
    V4 x_f = u * wf - splat(0.5f), y_f = v * hf - splat(0.5f);
    V4 x0 = vfloor(x_f), y0 = vfloor(y_f);
    V4 x1 = x0 + splat(1.0f), y1 = y0 + splat(1.0f);
    V4 fx = x_f - x0, fy = y_f - y0;
    V4 omfx = splat(1.0f) - fx, omfy = splat(1.0f) - fy;
    x0 = apply_wrap_mode(x0, wf, t.wrap_s);
    y0 = apply_wrap_mode(y0, hf, t.wrap_t);
    x1 = apply_wrap_mode(x1, wf, t.wrap_s);
    y1 = apply_wrap_mode(y1, hf, t.wrap_t);
    U4 x0i = as_uvec4(x0), y0i = as_uvec4(y0), x1i = as_uvec4(x1), y1i = as_uvec4(y1);
    U4 i00, i10, i01, i11;
    for (int i = 0; i < 4; i++) {
        i00.v[i] = off + y0i[i] * wi + x0i[i];
        i10.v[i] = off + y0i[i] * wi + x1i[i];
        i01.v[i] = off + y1i[i] * wi + x0i[i];
        i11.v[i] = off + y1i[i] * wi + x1i[i];
    }
    V3x4 p00 = gather_rgb(t, i00), p10 = gather_rgb(t, i10), p01 = gather_rgb(t, i01), p11 = gather_rgb(t, i11);
    V4 w00 = omfx * omfy, w10 = fx * omfy, w01 = omfx * fy, w11 = fx * fy;
    return {p00.x * w00 + p10.x * w10 + p01.x * w01 + p11.x * w11, p00.y * w00 + p10.y * w10 + p01.y * w01 + p11.y * w11,
            p00.z * w00 + p10.z * w10 + p01.z * w01 + p11.z * w11};
}

// voxelgrid.rs:264-368
static inline size_t f2usize(float f) {  // Rust `as usize`
    if (!(f > 0.0f)) return 0;
    if (f >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)f;
}
static void get_filtered_gi_sh4(const swr_voxel_grid_desc &g, V3x4 pos, V3x4 out_rgb[4], V4 out_w[4]) {
    const size_t W = g.dims[0], Hh = g.dims[1], D = g.dims[2];
    // voxelgrid.rs:154-161 voxel_size
    float vsx = (g.world_max[0] - g.world_min[0]) / (float)W;
    float vsy = (g.world_max[1] - g.world_min[1]) / (float)Hh;
    float vsz = (g.world_max[2] - g.world_min[2]) / (float)D;
    V4 vx = (pos.x - g.world_min[0]) / vsx;
    V4 vy = (pos.y - g.world_min[1]) / vsy;
    V4 vz = (pos.z - g.world_min[2]) / vsz;
    V4 x0f = vfloor(vx), y0f = vfloor(vy), z0f = vfloor(vz);
    size_t x0[4], y0[4], z0[4], x1[4], y1[4], z1[4];
    for (int i = 0; i < 4; i++) {
        x0[i] = std::min(f2usize(x0f[i]), W - 1);
        y0[i] = std::min(f2usize(y0f[i]), Hh - 1);
        z0[i] = std::min(f2usize(z0f[i]), D - 1);
        x1[i] = std::min(x0[i] + 1, W - 1);
        y1[i] = std::min(y0[i] + 1, Hh - 1);
        z1[i] = std::min(z0[i] + 1, D - 1);
    }
    V4 fx = vx - x0f, fy = vy - y0f, fz = vz - z0f;
    V4 one = splat(1.0f);
    auto fetch = [&](int c, const size_t *x, const size_t *y, const size_t *z, V3x4 &rgb, V4 &w) {
        const float *p[4];
        for (int i = 0; i < 4; i++) p[i] = g.gi_sh4 + ((z[i] * W * Hh + y[i] * W + x[i]) * 16 + (size_t)c * 4);
        rgb.x = V4(p[0][0], p[1][0], p[2][0], p[3][0]);
        rgb.y = V4(p[0][1], p[1][1], p[2][1], p[3][1]);
        rgb.z = V4(p[0][2], p[1][2], p[2][2], p[3][2]);
        w = V4(p[0][3], p[1][3], p[2][3], p[3][3]);
    };
    for (int c = 0; c < 4; c++) {
        V3x4 v000, v001, v010, v011, v100, v101, v110, v111;
        V4 w000, w001, w010, w011, w100, w101, w110, w111;
        fetch(c, x0, y0, z0, v000, w000);
        fetch(c, x0, y0, z1, v001, w001);
        fetch(c, x0, y1, z0, v010, w010);
        fetch(c, x0, y1, z1, v011, w011);
        fetch(c, x1, y0, z0, v100, w100);
        fetch(c, x1, y0, z1, v101, w101);
        fetch(c, x1, y1, z0, v110, w110);
        fetch(c, x1, y1, z1, v111, w111);
        V3x4 v00 = v000 * (one - fx) + v100 * fx;
        V3x4 v01 = v001 * (one - fx) + v101 * fx;
        V3x4 v10 = v010 * (one - fx) + v110 * fx;
        V3x4 v11 = v011 * (one - fx) + v111 * fx;
        V3x4 v0 = v00 * (one - fy) + v10 * fy;
        V3x4 v1 = v01 * (one - fy) + v11 * fy;
        out_rgb[c] = v0 * (one - fz) + v1 * fz;
        V4 w00 = w000 * (one - fx) + w100 * fx;
        V4 w01 = w001 * (one - fx) + w101 * fx;
        V4 w10 = w010 * (one - fx) + w110 * fx;
        V4 w11 = w011 * (one - fx) + w111 * fx;
        V4 w0 = w00 * (one - fy) + w10 * fy;
        V4 w1 = w01 * (one - fy) + w11 * fy;
        out_w[c] = w0 * (one - fz) + w1 * fz;
    }
}

// ---------------------------------------------------------------------------------
// shader.rs:110-309 pbr_shader::<false>
// ---------------------------------------------------------------------------------
// TRANSLUCENT = true adds the KHR_materials_transmission blend over `current_color` (shader.rs:265-277)
static V3x4 pbr_shader(const Oracle &o, const RasterPacket &packet, const swr_material_desc &mat, V4 bary1, V4 bary2, bool translucent = false,
                       V3x4 current_color = V3x4()) {
    const swr_scene_desc &sc = *o.scene;
    const V4 EPS = splat(1e-6f), PI = splat(3.14159265358979323846f), ONE = splat(1.0f), ZERO = splat(0.0f);
    V4 w = 1.0f / packet.one_over_w.interpolate(bary1, bary2);
    V3x4 input_normal = normalize(packet.normals.interpolate(bary1, bary2));
    V3x4 input_tangent = normalize(packet.tangents.interpolate(bary1, bary2));
    V4 tangent_sign = packet.tangent_sign.interpolate(bary1, bary2);
    V3x4 pos_world = packet.pos_world_over_w.interpolate(bary1, bary2) * w;
    V4 uv_x = packet.u_over_w.interpolate(bary1, bary2) * w;
    V4 uv_y = packet.v_over_w.interpolate(bary1, bary2) * w;
    // shader.rs:130 — lane-wise Vec4 * Vec4: component k of du_dv is scaled by lane k's w.
    V4 du_dv_v = V4(packet.du_dv[0], packet.du_dv[1], packet.du_dv[2], packet.du_dv[3]) * w;
    float du_dv[4] = {du_dv_v[0], du_dv_v[1], du_dv_v[2], du_dv_v[3]};

    V3x4 tangent_world = normalize(input_tangent - input_normal * dot(input_normal, input_tangent));
    V4 handed = select(cmpge(tangent_sign, ZERO), ONE, splat(-1.0f));
    V3x4 bitangent_world = cross(input_normal, tangent_world) * handed;
    V3x4 normal_world = normalize(input_normal);
    if (mat.normal_texture >= 0) {
        V3x4 tsn = sample4_rgb(sc.textures[mat.normal_texture], uv_x, uv_y, du_dv) * 2.0f + (-1.0f);
        normal_world = tangent_world * tsn.x + bitangent_world * tsn.y + input_normal * tsn.z;
        normal_world = normalize(normal_world);
    }
    V3x4 light_dir = {splat(sc.light_direction[0]), splat(sc.light_direction[1]), splat(sc.light_direction[2])};
    V3x4 light_color = {splat(sc.light_color[0]), splat(sc.light_color[1]), splat(sc.light_color[2])};
    V3x4 campos = {splat(o.cam->position[0]), splat(o.cam->position[1]), splat(o.cam->position[2])};
    V3x4 view_dir = campos - pos_world;
    V3x4 view_normal = normalize(view_dir);

    V4 n_dot_l = vmax(dot(normal_world, light_dir), ZERO);
    V3x4 half_vector = normalize(light_dir + view_normal);
    V4 n_dot_h = vmax(dot(normal_world, half_vector), ZERO);
    V4 n_dot_v = vmax(dot(normal_world, view_normal), splat(1.0e-4f));
    V4 v_dot_h = vmax(dot(view_normal, half_vector), ZERO);

    V3x4 gi_rgb[4];
    V4 gi_w[4];
    get_filtered_gi_sh4(sc.voxel_grid, pos_world, gi_rgb, gi_w);
    V4 voxel_light_intensity = vclamp(gi_w[0], ZERO, ONE);
    V4 sky_visibility = vclamp(gi_w[1], ZERO, ONE);

    V3x4 base = {splat(mat.base_color_factor[0]), splat(mat.base_color_factor[1]), splat(mat.base_color_factor[2])};
    if (mat.base_color_texture >= 0)
        base = base * srgb_to_linear_fast(sample4_rgb(sc.textures[mat.base_color_texture], uv_x, uv_y, du_dv));
    V4 roughness = splat(mat.roughness_factor), metallic = splat(mat.metallic_factor);
    if (mat.metallic_roughness_texture >= 0) {
        V3x4 mr = sample4_rgb(sc.textures[mat.metallic_roughness_texture], uv_x, uv_y, du_dv);
        roughness = roughness * mr.y;
        metallic = metallic * mr.z;
    }
    roughness = vclamp(roughness, splat(0.045f), ONE);
    metallic = vclamp(metallic, ZERO, ONE);
    V4 ao = ONE;
    if (mat.occlusion_texture >= 0) {
        ao = sample4_rgb(sc.textures[mat.occlusion_texture], uv_x, uv_y, du_dv).x;
        ao = ONE + (ao - ONE) * splat(mat.occlusion_strength);
    }
    V3x4 f0 = v3lerp(v3splat(0.04f), base, metallic);
    V4 omvh = ONE - v_dot_h;
    V4 omvh2 = omvh * omvh;
    V4 omvh4 = omvh2 * omvh2;
    V4 omvh5 = omvh4 * omvh;
    V3x4 brdf_f_direct = f0 + (v3splat(1.0f) - f0) * omvh5;

    V4 alpha = roughness * roughness;
    V4 alpha_2 = alpha * alpha;
    V4 ndh_2 = n_dot_h * n_dot_h;
    V4 denom_d = ndh_2 * (alpha_2 - ONE) + ONE;
    V4 brdf_d = alpha_2 / (PI * (denom_d * denom_d) + EPS);

    V4 k = roughness + ONE;
    k = (k * k) * splat(0.125f);
    V4 gv_denom = n_dot_v * (ONE - k) + k;
    V4 gv = n_dot_v / (gv_denom + EPS);
    V4 gl_denom = n_dot_l * (ONE - k) + k;
    V4 gl = n_dot_l / (gl_denom + EPS);
    V4 brdf_g = gv * gl;

    V4 specular_dg = (brdf_d * brdf_g) / (splat(4.0f) * n_dot_l * n_dot_v + EPS);
    V3x4 k_d_direct = (v3splat(1.0f) - brdf_f_direct) * (ONE - metallic);
    V3x4 lambert = base * splat(1.0f / 3.14159265358979323846f);
    V3x4 color_direct_diffuse = light_color * k_d_direct * lambert * n_dot_l * voxel_light_intensity;
    V3x4 color_direct_specular = light_color * brdf_f_direct * specular_dg * n_dot_l * voxel_light_intensity;

    V3x4 k_d_indirect = (v3splat(1.0f) - f0) * (ONE - metallic);
    // shader.rs:102-108
    V3x4 irradiance = v3max(gi_rgb[0] + gi_rgb[1] * normal_world.y + gi_rgb[2] * normal_world.z + gi_rgb[3] * normal_world.x, ZERO);
    V3x4 color_indirect_diffuse = irradiance * base * k_d_indirect * splat(1.0f / 3.14159265358979323846f);

    V3x4 reflect_dir = reflect(view_normal * -1.0f, normal_world);
    const swr_texture_desc &spec = sc.textures[sc.cubemap_specular];
    V4 spec_mip = roughness * splat((float)spec.max_mip_level);
    V3x4 prefiltered_env = sample_cubemap_trilinear_rgb(spec, reflect_dir, spec_mip);
    V3x4 brdf_lut = sample_bilinear_rgb0(sc.textures[sc.brdf_lut], vclamp(n_dot_v, ZERO, ONE), vclamp(roughness, ZERO, ONE));
    V3x4 brdf_spec_factor = f0 * brdf_lut.x;
    brdf_spec_factor = brdf_spec_factor + brdf_lut.y;
    V3x4 color_indirect_specular = prefiltered_env * brdf_spec_factor;
    V4 ao_spec = ONE + (ao - ONE) * splat(0.5f);
    color_indirect_specular = color_indirect_specular * (ao_spec * sky_visibility);

    V3x4 color;
    if (translucent) {
        V4 transmission = splat(mat.transmission);
        if (mat.transmission_texture >= 0) transmission = transmission * sample4_rgb(sc.textures[mat.transmission_texture], uv_x, uv_y, du_dv).x;
        transmission = vclamp(transmission, ZERO, ONE);
        V4 inv_transmission = ONE - transmission;
        color = (current_color * base * transmission) + ((color_direct_diffuse + color_indirect_diffuse) * inv_transmission) + color_direct_specular +
                color_indirect_specular;
    } else {
        color = color_direct_diffuse + color_direct_specular + color_indirect_diffuse + color_indirect_specular;
    }
    V3x4 emissive_mat = v3splat(1.0f);
    if (mat.emissive_texture >= 0)
        emissive_mat = srgb_to_linear_fast(sample4_rgb(sc.textures[mat.emissive_texture], uv_x, uv_y, du_dv));
    // math.rs Mul<Vec3A>: component-wise by the emissive factor
    color = color + V3x4{emissive_mat.x * mat.emissive_factor[0], emissive_mat.y * mat.emissive_factor[1],
                         emissive_mat.z * mat.emissive_factor[2]};
    return color;
}

// ---------------------------------------------------------------------------------
// tilerasterizer.rs
// ---------------------------------------------------------------------------------
static const int SUBPIXEL_SCALE = 16, SUBPIXEL_SHIFT = 4, STEP = 32, HALF_PIXEL = 8, ONE_HALF_PIXEL = 24,
                 COARSE_PX = 16, COARSE_SUB = 256, TILE = 64;

struct RasterParams {
    bool translucent = false;  // TranslucentForwardShader instead of VBufferOpaqueShader
    const RasterPacket *packet;
    uint32_t packet_index;
    const swr_material_desc *material;
    V4 a01, b01, c01, a12, b12, c12, a20, b20, c20;
    V4 sx01, sx12, sx20, sy01, sy12, sy20;
    V4 ooa;
};

static inline size_t index_from_xy(const Tile &t, int px, int py) {  // tilerasterizer.rs:525-543
    int lx = px - t.min_x, ly = py - t.min_y;
    int qx = lx / 2, qy = ly / 2;
    int wq = (t.max_x - t.min_x) / 2;
    return (size_t)(qy * wq + qx);
}

// shader.rs:311-329
static M4 get_alpha_test_mask(const Oracle &o, const RasterParams &rp, V4 bary1, V4 bary2, V4 w, M4 mask) {
    const RasterPacket &p = *rp.packet;
    V4 u = p.u_over_w.interpolate(bary1, bary2) * w;
    V4 v = p.v_over_w.interpolate(bary1, bary2) * w;
    M4 out = mask;
    if (rp.material->base_color_texture >= 0) {
        V4 alpha = sample4_alpha(o.scene->textures[rp.material->base_color_texture], u, v, p.du_dv);
        out = mand(out, cmpge(alpha, splat(rp.material->alpha_cutoff)));
    }
    return out;
}

static void fine_raster(const Oracle &o, Tile &tile, const RasterParams &rp, int xs, int ys, int xe, int ye) {  // :293-383
    V4 x((float)(xs + HALF_PIXEL), (float)(xs + ONE_HALF_PIXEL), (float)(xs + HALF_PIXEL), (float)(xs + ONE_HALF_PIXEL));
    V4 y((float)(ys + HALF_PIXEL), (float)(ys + HALF_PIXEL), (float)(ys + ONE_HALF_PIXEL), (float)(ys + ONE_HALF_PIXEL));
    V4 w0_row = rp.a12 * x + rp.b12 * y + rp.c12;
    V4 w1_row = rp.a20 * x + rp.b20 * y + rp.c20;
    V4 w2_row = rp.a01 * x + rp.b01 * y + rp.c01;
    const V4 zero = splat(0.0f);
    for (int py = ys; py < ye; py += STEP) {
        V4 w0 = w0_row, w1 = w1_row, w2 = w2_row;
        for (int px = xs; px < xe; px += STEP) {
            M4 mask = mand(mand(cmpge(w0, zero), cmpge(w1, zero)), cmpge(w2, zero));
            if (many(mask)) {
                V4 bary1 = w1 * rp.ooa;
                V4 bary2 = w2 * rp.ooa;
                V4 w = 1.0f / rp.packet->one_over_w.interpolate(bary1, bary2);
                V4 z = rp.packet->z_over_w.interpolate(bary1, bary2) * w;
                int pix_x = px >> SUBPIXEL_SHIFT, pix_y = py >> SUBPIXEL_SHIFT;
                size_t index = index_from_xy(tile, pix_x, pix_y);
                // depth_test :511-523
                V4 cur = tile.depth[index];
                M4 fmask = mand(mask, cmple(z, cur));
                V4 fdepth = select(fmask, z, cur);
                if (many(fmask) && rp.translucent) {
                    // TranslucentForwardShader::shade shader.rs:66-76 (depth is tested but never written)
                    V3x4 cur_col = tile.color[index];
                    V3x4 col = pbr_shader(o, *rp.packet, *rp.material, bary1, bary2, true, cur_col);
                    tile.color[index] = v3select(fmask, col, cur_col);
                } else if (many(fmask)) {
                    // VBufferOpaqueShader::shade shader.rs:32-63
                    M4 m = fmask;
                    bool alpha_tested = (rp.material->flags & SWR_MAT_ALPHA_TESTED) != 0;
                    if (alpha_tested) m = mand(m, get_alpha_test_mask(o, rp, bary1, bary2, w, fmask));
                    int bm = bitmask(m);
                    for (int l = 0; l < 4; l++)
                        if (bm & (1 << l)) {
                            tile.packet_index[index].v[l] = rp.packet_index;
                            if (o.track_written) tile.written[index].v[l] = 1;
                        }
                    tile.bary1[index] = select(m, bary1, tile.bary1[index]);
                    tile.bary2[index] = select(m, bary2, tile.bary2[index]);
                    if (!alpha_tested)
                        tile.depth[index] = fdepth;
                    else
                        tile.depth[index] = select(m, z, tile.depth[index]);
                }
            }
            w0 = w0 + rp.sx12;
            w1 = w1 + rp.sx20;
            w2 = w2 + rp.sx01;
        }
        w0_row = w0_row + rp.sy12;
        w1_row = w1_row + rp.sy20;
        w2_row = w2_row + rp.sy01;
    }
}

static void coarse_raster(const Oracle &o, Tile &tile, const RasterParams &rp) {  // :220-291
    const RasterPacket &p = *rp.packet;
    int xs = (p.min_x & ~1) * SUBPIXEL_SCALE, ys = (p.min_y & ~1) * SUBPIXEL_SCALE;
    int xe = p.max_x * SUBPIXEL_SCALE, ye = p.max_y * SUBPIXEL_SCALE;
    for (int by = ys; by < ye; by += COARSE_SUB) {
        for (int bx = xs; bx < xe; bx += COARSE_SUB) {
            float x0 = (float)(bx + HALF_PIXEL), x1 = (float)(bx + COARSE_SUB - HALF_PIXEL);
            float y0 = (float)(by + HALF_PIXEL), y1 = (float)(by + COARSE_SUB - HALF_PIXEL);
            V4 cx(x0, x1, x0, x1), cy(y0, y0, y1, y1);
            V4 w0c = rp.a12 * cx + rp.b12 * cy + rp.c12;
            V4 w1c = rp.a20 * cx + rp.b20 * cy + rp.c20;
            V4 w2c = rp.a01 * cx + rp.b01 * cy + rp.c01;
            bool outside = (max_element(w0c) < 0.0f) || (max_element(w1c) < 0.0f) || (max_element(w2c) < 0.0f);
            if (outside) continue;
            fine_raster(o, tile, rp, bx, by, std::min(bx + COARSE_SUB, xe), std::min(by + COARSE_SUB, ye));
        }
    }
}

static inline int32_t top_left_bias(int32_t a, int32_t b) { return (a < 0 || (a == 0 && b > 0)) ? 0 : -1; }  // :139-141
static inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }           // release-mode wrap
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

static void rasterize_packet(const Oracle &o, Tile &tile, uint32_t packet_index, const RasterPacket &p, bool translucent = false) {  // :114-218
    RasterParams rp;
    rp.translucent = translucent;
    rp.packet = &p;
    rp.packet_index = packet_index;
    const swr_primitive_desc &prim = o.scene->primitives[p.primitive_index];
    rp.material = &o.scene->materials[prim.material_index];
    int32_t x0 = p.px[0], y0 = p.py[0], x1 = p.px[1], y1 = p.py[1], x2 = p.px[2], y2 = p.py[2];
    int32_t a01 = wsub(y1, y0), b01 = wsub(x0, x1);
    int32_t c01 = wadd(wsub(wmul(x1, y0), wmul(x0, y1)), top_left_bias(a01, b01));
    int32_t a12 = wsub(y2, y1), b12 = wsub(x1, x2);
    int32_t c12 = wadd(wsub(wmul(x2, y1), wmul(x1, y2)), top_left_bias(a12, b12));
    int32_t a20 = wsub(y0, y2), b20 = wsub(x2, x0);
    int32_t c20 = wadd(wsub(wmul(x0, y2), wmul(x2, y0)), top_left_bias(a20, b20));
    rp.sx01 = splat((float)wmul(a01, STEP));
    rp.sx12 = splat((float)wmul(a12, STEP));
    rp.sx20 = splat((float)wmul(a20, STEP));
    rp.sy01 = splat((float)wmul(b01, STEP));
    rp.sy12 = splat((float)wmul(b12, STEP));
    rp.sy20 = splat((float)wmul(b20, STEP));
    rp.a01 = splat((float)a01);
    rp.b01 = splat((float)b01);
    rp.c01 = splat((float)c01);
    rp.a12 = splat((float)a12);
    rp.b12 = splat((float)b12);
    rp.c12 = splat((float)c12);
    rp.a20 = splat((float)a20);
    rp.b20 = splat((float)b20);
    rp.c20 = splat((float)c20);
    rp.ooa = splat(p.one_over_area);
    int wpx = p.max_x - p.min_x, hpx = p.max_y - p.min_y;
    if (wpx > COARSE_PX || hpx > COARSE_PX) {
        coarse_raster(o, tile, rp);
    } else {
        int xs = (p.min_x & ~1) * SUBPIXEL_SCALE, ys = (p.min_y & ~1) * SUBPIXEL_SCALE;
        fine_raster(o, tile, rp, xs, ys, p.max_x * SUBPIXEL_SCALE, p.max_y * SUBPIXEL_SCALE);
    }
}

static V3x4 compute_skybox(const Oracle &o, int px, int py) {  // :478-508
    const swr_camera &c = *o.cam;
    float fx = (float)px, fy = (float)py;
    V4 pixel_x(fx + 0.5f, fx + 1.5f, fx + 0.5f, fx + 1.5f);
    V4 pixel_y(fy + 0.5f, fy + 0.5f, fy + 1.5f, fy + 1.5f);
    V4 one = splat(1.0f);
    V4 ndc_x = pixel_x * c.one_over_width * 2.0f - one;
    V4 ndc_y = (one - pixel_y * c.one_over_height) * 2.0f - one;
    V3x4 v = {ndc_x, ndc_y, one};
    const float *m = c.skybox_matrix_transposed;  // math.rs:135-149: rows = columns of the transposed matrix
    V3x4 d;
    d.x = v.x * m[0] + v.y * m[1] + v.z * m[2] + m[3];
    d.y = v.x * m[4] + v.y * m[5] + v.z * m[6] + m[7];
    d.z = v.x * m[8] + v.y * m[9] + v.z * m[10] + m[11];
    V3x4 n = normalize(d);
    U4 zero = {{0, 0, 0, 0}};
    return srgb_to_linear_fast(sample_cubemap_rgb(o.scene->textures[o.scene->cubemap], n, zero));
}

static void shade_vbuffer(const Oracle &o, Tile &tile, bool skybox_only) {  // :386-476
    int tw = tile.max_x - tile.min_x, th = tile.max_y - tile.min_y;
    int wq = tw / 2;
    const V4 INF = splat(INFINITY);
    for (int y = 0; y < th / 2; y++) {
        int sy = tile.min_y + y * 2;
        for (int qx = 0; qx < wq; qx++) {
            int sx = tile.min_x + qx * 2;
            size_t index = index_from_xy(tile, sx, sy);
            if (skybox_only) {
                tile.color[index] = compute_skybox(o, sx, sy);
                continue;
            }
            U4 pidx = tile.packet_index[index];
            V4 depth = tile.depth[index];
            M4 depth_mask = cmpne(depth, INF);
            V3x4 out = v3splat(0.0f);
            if (many(depth_mask)) {
                int remaining = bitmask(depth_mask);
                while (remaining) {
                    int lane = __builtin_ctz(remaining);
                    uint32_t pi = pidx[lane];
                    int eq = 0;
                    for (int l = 0; l < 4; l++)
                        if (pidx[l] == pi) eq |= 1 << l;
                    alignas(16) uint32_t mbits[4];
                    for (int l = 0; l < 4; l++) mbits[l] = (eq >> l) & 1 ? 0xFFFFFFFFu : 0u;
                    M4 mask = _mm_load_ps((const float *)mbits);
                    RasterPacket packet = tile.packets_opaque.get(pi);  // copy, as bumpqueue.rs:114-120
                    const swr_primitive_desc &prim = o.scene->primitives[packet.primitive_index];
                    const swr_material_desc &mat = o.scene->materials[prim.material_index];
                    V3x4 color = pbr_shader(o, packet, mat, tile.bary1[index], tile.bary2[index]);
                    out = v3select(mask, color, out);
                    remaining &= ~eq;
                }
            }
            if (!mall(depth_mask)) {
                V3x4 sky = compute_skybox(o, sx, sy);
                out = v3select(depth_mask, out, sky);
            }
            tile.color[index] = out;
        }
    }
}

// OrderedFloat total order used by sort_by_key (renderer.rs:361-366): NaN sorts last.
static inline bool of_less(float a, float b) {
    bool an = a != a, bn = b != b;
    if (an || bn) return !an && bn;
    return a < b;
}

static void render_tile(Oracle &o, Tile &tile, bool shade) {  // :72-111
    std::fill(tile.depth.begin(), tile.depth.end(), splat(INFINITY));
    if (o.track_written) std::fill(tile.written.begin(), tile.written.end(), U4{{0, 0, 0, 0}});
    uint32_t n = tile.packets_opaque.len();
    for (uint32_t i = 0; i < n; i++) {
        RasterPacket packet = tile.packets_opaque.get(i);
        rasterize_packet(o, tile, i, packet);
    }
    if (shade) {
        shade_vbuffer(o, tile, n == 0);
        // :92-101 sort translucent packets back to front and forward-shade them. The reference's quicksort
        // (bumpqueue.rs:155-202) is unstable; equal avg_z keeps submission order here (documented in DESIGN.md).
        uint32_t nt = tile.packets_translucent.len();
        if (nt) {
            std::vector<uint32_t> order(nt);
            for (uint32_t i = 0; i < nt; i++) order[i] = i;
            // OrderedFloat descending; with a parallel bin phase the push order is arbitrary, so ties fall back to seq
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
                const RasterPacket &pa = tile.packets_translucent.ref(a), &pb = tile.packets_translucent.ref(b);
                if (pa.avg_z != pb.avg_z) return of_less(pb.avg_z, pa.avg_z);
                return pa.seq < pb.seq;
            });
            for (uint32_t i = 0; i < nt; i++) {
                RasterPacket packet = tile.packets_translucent.get(order[i]);
                rasterize_packet(o, tile, i, packet, true);
            }
        }
        // :103-106
        const V3x4 &cq = tile.color[tile.color.size() / 2];
        V4 lum = cq.x * 0.2126f + cq.y * 0.7152f + cq.z * 0.0722f;
        tile.center_luminance = (lum[0] + lum[1] + lum[2] + lum[3]) * 0.25f;
    }
}

// ---------------------------------------------------------------------------------
// renderer.rs
// ---------------------------------------------------------------------------------
static inline void clip_to_screen_subpixels(const Oracle &o, F4 v, int32_t &X, int32_t &Y) {  // :834-846
    float nx = v.x / v.w, ny = v.y / v.w;
    float sx = (nx + 1.0f) * (float)o.W / 2.0f;
    float sy = (1.0f - ny) * (float)o.H / 2.0f;
    X = f2i(roundf(sx * 16.0f));
    Y = f2i(roundf(sy * 16.0f));
}

static void bin_triangle(Oracle &o, const Vertex tri[3], uint32_t mesh_index, uint32_t primitive_index, uint32_t seq, bool opaque) {  // :668-831
    int32_t X[3], Y[3];
    for (int i = 0; i < 3; i++) clip_to_screen_subpixels(o, tri[i].pos_clip, X[i], Y[i]);
    int32_t area = wsub(wmul(wsub(X[1], X[0]), wsub(Y[2], Y[0])), wmul(wsub(X[2], X[0]), wsub(Y[1], Y[0])));
    if (area > 0) return;
    int32_t mnx = std::max(std::min(std::min(X[0], X[1]), X[2]), 0);
    int32_t mny = std::max(std::min(std::min(Y[0], Y[1]), Y[2]), 0);
    int32_t mxx = std::min(std::max(std::max(X[0], X[1]), X[2]), o.W * SUBPIXEL_SCALE);
    int32_t mxy = std::min(std::max(std::max(Y[0], Y[1]), Y[2]), o.H * SUBPIXEL_SCALE);
    int32_t bminx = mnx >> SUBPIXEL_SHIFT, bminy = mny >> SUBPIXEL_SHIFT;
    int32_t bmaxx = wadd(mxx, SUBPIXEL_SCALE) >> SUBPIXEL_SHIFT, bmaxy = wadd(mxy, SUBPIXEL_SCALE) >> SUBPIXEL_SHIFT;

    RasterPacket pk;
    memset(&pk, 0, sizeof(pk));
    // i32::abs wraps for MIN in release; (float) of that is -2^31
    int32_t aabs = area < 0 ? (int32_t)(0u - (uint32_t)area) : area;
    pk.one_over_area = 1.0f / (float)aabs;
    float iw[3], zw[3];
    F2 uvw[3];
    F3 pww[3];
    for (int i = 0; i < 3; i++) {
        iw[i] = 1.0f / tri[i].pos_clip.w;
        zw[i] = tri[i].pos_clip.z * iw[i];
        uvw[i] = F2{tri[i].uv.x * iw[i], tri[i].uv.y * iw[i]};
        pww[i] = F3{tri[i].pos_world.x * iw[i], tri[i].pos_world.y * iw[i], tri[i].pos_world.z * iw[i]};
    }
    float dx1 = (float)wsub(X[1], X[0]), dx2 = (float)wsub(X[2], X[0]);
    float dy1 = (float)wsub(Y[1], Y[0]), dy2 = (float)wsub(Y[2], Y[0]);
    float du1 = uvw[1].x - uvw[0].x, du2 = uvw[2].x - uvw[0].x;
    float dv1 = uvw[1].y - uvw[0].y, dv2 = uvw[2].y - uvw[0].y;
    float du_dx = du1 * dy2 - du2 * dy1;
    float du_dy = du2 * dx1 - du1 * dx2;
    float dv_dx = dv1 * dy2 - dv2 * dy1;
    float dv_dy = dv2 * dx1 - dv1 * dx2;
    float s = pk.one_over_area * 16.0f;
    pk.du_dv[0] = du_dx * s;
    pk.du_dv[1] = du_dy * s;
    pk.du_dv[2] = dv_dx * s;
    pk.du_dv[3] = dv_dy * s;

    // Rust `/` on i32 truncates toward zero; operands here are >= 0 except possibly bmax (negative when offscreen)
    int32_t min_bin_x = bminx / TILE, min_bin_y = bminy / TILE;
    int32_t max_bin_x = (bmaxx + TILE - 1) / TILE, max_bin_y = (bmaxy + TILE - 1) / TILE;
    int32_t tile_w = o.tiles_x;
    int32_t ntiles = (int32_t)o.tiles.size();

    for (int i = 0; i < 3; i++) {
        pk.px[i] = X[i];
        pk.py[i] = Y[i];
    }
    pk.z_over_w.set(zw[0], zw[1], zw[2]);
    pk.one_over_w.set(iw[0], iw[1], iw[2]);
    pk.normals.set(tri[0].normal, tri[1].normal, tri[2].normal);
    pk.tangents.set(F3{tri[0].tangent.x, tri[0].tangent.y, tri[0].tangent.z}, F3{tri[1].tangent.x, tri[1].tangent.y, tri[1].tangent.z},
                    F3{tri[2].tangent.x, tri[2].tangent.y, tri[2].tangent.z});
    pk.tangent_sign.set(tri[0].tangent.w, tri[1].tangent.w, tri[2].tangent.w);
    pk.u_over_w.set(uvw[0].x, uvw[1].x, uvw[2].x);
    pk.v_over_w.set(uvw[0].y, uvw[1].y, uvw[2].y);
    pk.pos_world_over_w.set(pww[0], pww[1], pww[2]);
    pk.primitive_index = primitive_index;
    pk.mesh_index = mesh_index;
    // renderer.rs:765-775
    pk.avg_z = opaque ? 0.0f : (tri[0].pos_clip.z + tri[1].pos_clip.z + tri[2].pos_clip.z) / 3.0f;
    pk.seq = seq;

    bool any = false;
    uint64_t npk = 0, ndup = 0;
    for (int32_t y = min_bin_y; y < max_bin_y; y++) {
        for (int32_t x = min_bin_x; x < max_bin_x; x++) {
            int32_t bin = y * tile_w + x;
            if (bin >= 0 && bin < ntiles) {
                Tile &t = *o.tiles[bin];
                int32_t cminx = std::max(bminx, t.min_x), cminy = std::max(bminy, t.min_y);
                int32_t cmaxx = std::min(bmaxx, t.max_x), cmaxy = std::min(bmaxy, t.max_y);
                if (cmaxx - cminx < 1 || cmaxy - cminy < 1) continue;
                pk.min_x = cminx;
                pk.min_y = cminy;
                pk.max_x = cmaxx;
                pk.max_y = cmaxy;
                if (opaque)
                    t.packets_opaque.push(pk);
                else
                    t.packets_translucent.push(pk);
                any = true;
                if (x >= tile_w)
                    ndup++;  // column overflow wrapped into the next tile row: identical duplicate
                else
                    npk++;
            }
        }
    }
    if (opaque) {  // the reported counters describe the opaque pass (what the BASELINE configs exercise)
        ThreadStats &ts = o.tstats[omp_get_thread_num()];
        ts.tris_binned += any ? 1 : 0;
        ts.packets += npk;
        ts.packets_dup += ndup;
    }
}

static inline Vertex intersect(const Vertex &v0, const Vertex &v1, F4 plane) {  // :598-609
    float d0 = dot4(plane, v0.pos_clip), d1 = dot4(plane, v1.pos_clip);
    float t = d0 / (d0 - d1);
    Vertex r;
#define LERP(f) r.f = v0.f + (v1.f - v0.f) * t
    LERP(pos_clip.x);
    LERP(pos_clip.y);
    LERP(pos_clip.z);
    LERP(pos_clip.w);
    LERP(pos_world.x);
    LERP(pos_world.y);
    LERP(pos_world.z);
    LERP(normal.x);
    LERP(normal.y);
    LERP(normal.z);
    LERP(tangent.x);
    LERP(tangent.y);
    LERP(tangent.z);
    LERP(tangent.w);
    LERP(uv.x);
    LERP(uv.y);
#undef LERP
    return r;
}

static void clip_against_frustum(Oracle &o, const Vertex tri[3], uint32_t mesh_index, uint32_t primitive_index, uint32_t seq_base, bool opaque) {  // :579-665
    static const F4 PLANES[6] = {{0, 0, 1, 1}, {0, 0, -1, 1}, {1, 0, 0, 1}, {-1, 0, 0, 1}, {0, 1, 0, 1}, {0, -1, 0, 1}};
    Vertex poly[16], np[16];
    int n = 3;
    poly[0] = tri[0];
    poly[1] = tri[1];
    poly[2] = tri[2];
    bool touched = false;
    for (int pl = 0; pl < 6; pl++) {
        if (n == 0) return;
        F4 plane = PLANES[pl];
        int m = 0;
        for (int i = 0; i < n; i++) {
            const Vertex &curr = poly[i];
            const Vertex &prev = poly[(i + n - 1) % n];
            bool cin = dot4(plane, curr.pos_clip) >= 0.0f;
            bool pin = dot4(plane, prev.pos_clip) >= 0.0f;
            if (cin) {
                if (!pin) {
                    np[m++] = intersect(prev, curr, plane);
                    touched = true;
                }
                np[m++] = curr;
            } else {
                touched = true;
                if (pin) np[m++] = intersect(prev, curr, plane);
            }
        }
        for (int i = 0; i < m; i++) poly[i] = np[i];
        n = m;
    }
    if (n < 3) return;
    if (touched && opaque) o.tstats[omp_get_thread_num()].tris_clipped++;
    for (int i = 1; i < n - 1; i++) {
        Vertex t[3] = {poly[0], poly[i], poly[i + 1]};
        bin_triangle(o, t, mesh_index, primitive_index, seq_base + (uint32_t)(i - 1), opaque);
    }
}

static void process_triangle(Oracle &o, const Draw &d, uint32_t tri_idx) {  // renderer.rs:493-573
    const swr_primitive_desc &prim = o.scene->primitives[d.primitive];
    Vertex tri[3];
    for (int k = 0; k < 3; k++) {
        uint32_t i = prim.indices[tri_idx * 3 + k];
        F4 local = {prim.positions[i * 4 + 0], prim.positions[i * 4 + 1], prim.positions[i * 4 + 2], prim.positions[i * 4 + 3]};
        F4 world = mul_vec4(d.model, local);
        tri[k].pos_world = F3{world.x, world.y, world.z};
        tri[k].pos_clip = mul_vec4(d.mvp, local);
        tri[k].normal = mul_mat3(d.model, F3{prim.normals[i * 4 + 0], prim.normals[i * 4 + 1], prim.normals[i * 4 + 2]});
        F3 tw = mul_mat3(d.model, F3{prim.tangents[i * 4 + 0], prim.tangents[i * 4 + 1], prim.tangents[i * 4 + 2]});
        tri[k].tangent = F4{tw.x, tw.y, tw.z, prim.tangents[i * 4 + 3]};
        tri[k].uv = F2{prim.texcoords[i * 2 + 0], prim.texcoords[i * 2 + 1]};
    }
    uint32_t seq = (d.first_triangle + tri_idx) * 8u;
    const bool opaque = !(d.flags & SWR_DRAW_TRANSLUCENT);
    if (d.flags & SWR_DRAW_CLIP)
        clip_against_frustum(o, tri, d.mesh, d.primitive, seq, opaque);
    else
        bin_triangle(o, tri, d.mesh, d.primitive, seq, opaque);
}

// scene.rs:53-63 Mat4 * &BoundingSphere, renderer.rs:130-142 test_sphere_frustum
static int classify_sphere(const swr_camera &cam, const float *model, const float *sphere) {
    float lx = sqrtf(dot4(F4{model[0], model[1], model[2], model[3]}, F4{model[0], model[1], model[2], model[3]}));
    float ly = sqrtf(dot4(F4{model[4], model[5], model[6], model[7]}, F4{model[4], model[5], model[6], model[7]}));
    float lz = sqrtf(dot4(F4{model[8], model[9], model[10], model[11]}, F4{model[8], model[9], model[10], model[11]}));
    float max_scale = (lx + ly + lz) / 3.0f;
    // glam transform_point3a: ((x_axis*x + y_axis*y) + z_axis*z) + w_axis
    F3 c;
    c.x = ((model[0] * sphere[0] + model[4] * sphere[1]) + model[8] * sphere[2]) + model[12];
    c.y = ((model[1] * sphere[0] + model[5] * sphere[1]) + model[9] * sphere[2]) + model[13];
    c.z = ((model[2] * sphere[0] + model[6] * sphere[1]) + model[10] * sphere[2]) + model[14];
    float radius = sphere[3] * max_scale;
    F4 cv = mul_vec4(cam.view_matrix, F4{c.x, c.y, c.z, 1.0f});
    int result = 0;  // 0 inside, 1 intersecting, 2 outside
    for (int p = 0; p < 6; p++) {
        const float *pl = cam.view_clip_planes[p];
        float dist = dot4(F4{pl[0], pl[1], pl[2], pl[3]}, cv);
        if (dist < -radius)
            return 2;
        else if (dist < radius)
            result = 1;
    }
    return result;
}

static void build_draws(Oracle &o) {  // renderer.rs:357-468
    const swr_scene_desc &sc = *o.scene;
    std::vector<uint32_t> order(sc.nnodes);
    std::vector<float> key(sc.nnodes);
    for (uint32_t i = 0; i < sc.nnodes; i++) {
        order[i] = i;
        const float *s = sc.nodes[i].bounding_sphere_world;
        F3 d = {o.cam->position[0] - s[0], o.cam->position[1] - s[1], o.cam->position[2] - s[2]};
        key[i] = dot3(d, d);
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return of_less(key[a], key[b]); });
    uint32_t first_tri = 0, first_tri_t = 0;
    o.draws.clear();
    o.tris_submitted = 0;
    o.verts_submitted = 0;
    for (uint32_t ni : order) {
        const swr_node_desc &node = sc.nodes[ni];
        if (node.mesh_index < 0) continue;
        const swr_mesh_desc &mesh = sc.meshes[node.mesh_index];
        for (int pass = 0; pass < 2; pass++) {  // render_mesh: primitives_opaque, then primitives_translucent (renderer.rs:386-420)
            for (uint32_t pi = mesh.first_primitive; pi < mesh.first_primitive + mesh.num_primitives; pi++) {
                const swr_primitive_desc &prim = sc.primitives[pi];
                const bool tr = (sc.materials[prim.material_index].flags & SWR_MAT_TRANSLUCENT) != 0;
                if (tr != (pass == 1)) continue;
                int cls = classify_sphere(*o.cam, node.transform, prim.bounding_sphere);
                if (cls == 2) continue;
                Draw d;
                memcpy(d.model, node.transform, sizeof(d.model));
                mul_mat4(o.cam->view_project_matrix, node.transform, d.mvp);  // renderer.rs:378
                d.primitive = pi;
                d.mesh = (uint32_t)node.mesh_index;
                d.flags = (cls == 1 ? SWR_DRAW_CLIP : 0) | (tr ? SWR_DRAW_TRANSLUCENT : 0);
                uint32_t nt = prim.nindices / 3;
                if (tr) {  // translucent packets have their own queue, hence their own submission order ids
                    d.first_triangle = first_tri_t;
                    first_tri_t += nt;
                } else {
                    d.first_triangle = first_tri;
                    first_tri += nt;
                    o.tris_submitted += nt;
                    o.verts_submitted += prim.nverts;
                }
                o.draws.push_back(d);
            }
        }
    }
}

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void render_scene(Oracle &o, bool shade) {  // renderer.rs:201-220
    build_draws(o);
    double t0 = now_ms();
    if (o.nthreads <= 1) {
        for (const Draw &d : o.draws) {
            uint32_t nt = o.scene->primitives[d.primitive].nindices / 3;
            for (uint32_t t = 0; t < nt; t++) process_triangle(o, d, t);
        }
    } else {
        // rayon fan-out #1/#2 (renderer.rs:207, :490-493): batches of >= 128 triangles, dynamic scheduling
        struct Batch {
            uint32_t draw, t0, t1;
        };
        std::vector<Batch> batches;
        for (uint32_t di = 0; di < o.draws.size(); di++) {
            uint32_t nt = o.scene->primitives[o.draws[di].primitive].nindices / 3;
            for (uint32_t t = 0; t < nt; t += 128) batches.push_back({di, t, std::min(nt, t + 128)});
        }
// rayon splits a range recursively (large contiguous pieces first, smaller ones when stealing): guided scheduling
#pragma omp parallel for schedule(guided, 1) num_threads(o.nthreads)
        for (long b = 0; b < (long)batches.size(); b++) {
            const Batch &bt = batches[b];
            for (uint32_t t = bt.t0; t < bt.t1; t++) process_triangle(o, o.draws[bt.draw], t);
        }
    }
    for (const ThreadStats &ts : o.tstats) {
        o.stats.tris_binned += ts.tris_binned;
        o.stats.tris_clipped += ts.tris_clipped;
        o.stats.packets += ts.packets;
        o.stats.packets_dup += ts.packets_dup;
    }
    double t1 = now_ms();
    // rayon fan-out #3 (renderer.rs:216): one job per tile
#pragma omp parallel for schedule(dynamic, 1) num_threads(o.nthreads > 1 ? o.nthreads : 1)
    for (long i = 0; i < (long)o.tiles.size(); i++) render_tile(o, *o.tiles[i], shade);
    double t2 = now_ms();
    o.ms_clipbin = t1 - t0;
    o.ms_raster = t2 - t1;
}

static inline float tonemap1(float c) { return c / (c + 0.2f) * (1.0f + 0.2f); }  // util.rs:37-41
static inline uint32_t f2u8(float f) {                                            // Rust `as u8`
    if (!(f > 0.0f)) return 0;
    if (f >= 255.0f) return 255;
    return (uint32_t)f;
}

}  // namespace

// ---------------------------------------------------------------------------------
// C exports (ctypes)
// ---------------------------------------------------------------------------------
extern "C" {

struct orc_stats {
    uint64_t triangles_submitted, vertices_submitted, triangles_binned, triangles_clipped, tile_refs, tile_refs_dup;
    uint32_t ndraws, tiles;
    double ms_clipbin, ms_raster, ms_resolve;
};

void orc_set_exact_rsqrt(int on) { g_exact_rsqrt = on != 0; }

void *orc_create(int width, int height) {  // Renderer::new renderer.rs:165-198
    Oracle *o = new Oracle();
    o->W = width;
    o->H = height;
    o->tiles_x = (width + TILE - 1) / TILE;
    o->tiles_y = (height + TILE - 1) / TILE;
    for (int y = 0; y < o->tiles_y; y++)
        for (int x = 0; x < o->tiles_x; x++) {
            Tile *t = new Tile();
            t->min_x = x * TILE;
            t->min_y = y * TILE;
            t->max_x = (x + 1) * TILE;
            t->max_y = (y + 1) * TILE;
            size_t nq = (TILE / 2) * (TILE / 2);
            t->color.assign(nq, v3splat(0.0f));
            t->depth.assign(nq, splat(INFINITY));
            t->packet_index.assign(nq, U4{{0, 0, 0, 0}});
            t->bary1.assign(nq, splat(0.0f));
            t->bary2.assign(nq, splat(0.0f));
            t->written.assign(nq, U4{{0, 0, 0, 0}});
            o->tiles.push_back(t);
        }
    return o;
}

void orc_destroy(void *h) { delete (Oracle *)h; }

// One frame: render_scene. fresh != 0 zeroes the barycentric/id buffers first so the frame behaves as the first
// frame of a fresh Renderer. Outputs (any may be NULL) are row-major W*H.
int orc_render(void *h, const swr_scene_desc *scene, const swr_camera *cam, int nthreads, int shade, int fresh,
               uint32_t *depth_bits, uint32_t *seq, float *bary1, float *bary2, float *color_rgb, float *tile_luminance,
               orc_stats *st) {
    Oracle &o = *(Oracle *)h;
    o.scene = scene;
    o.cam = cam;
    o.nthreads = nthreads;
    o.stats = Stats();
    o.tstats.assign((size_t)std::min(std::max(std::max(nthreads, omp_get_max_threads()), 1), MAX_THREADS), ThreadStats());
    const bool any_output = depth_bits || seq || bary1 || bary2 || color_rgb;
    o.track_written = seq != nullptr;
    for (Tile *t : o.tiles) {
        t->packets_opaque.reset();
        t->packets_translucent.reset();
        if (fresh) {
            std::fill(t->packet_index.begin(), t->packet_index.end(), U4{{0, 0, 0, 0}});
            std::fill(t->bary1.begin(), t->bary1.end(), splat(0.0f));
            std::fill(t->bary2.begin(), t->bary2.end(), splat(0.0f));
        }
    }
    render_scene(o, shade != 0);
    for (size_t ti = 0; ti < o.tiles.size(); ti++) {
        Tile &t = *o.tiles[ti];
        if (tile_luminance) tile_luminance[ti] = t.center_luminance;
        if (!any_output) continue;  // timed runs pass no output arrays: nothing to copy out
        for (int qy = 0; qy < TILE / 2; qy++)
            for (int qx = 0; qx < TILE / 2; qx++) {
                size_t qi = (size_t)qy * (TILE / 2) + qx;
                for (int l = 0; l < 4; l++) {
                    int px = t.min_x + qx * 2 + (l & 1), py = t.min_y + qy * 2 + (l >> 1);
                    if (px >= o.W || py >= o.H) continue;
                    size_t pi = (size_t)py * o.W + px;
                    float d = t.depth[qi][l];
                    uint32_t db;
                    memcpy(&db, &d, 4);
                    if (depth_bits) depth_bits[pi] = db;
                    // A lane that was never written this frame is "uncovered" (+INF, 0xFFFFFFFF). A lane written by
                    // a fragment whose depth is +INF keeps its id (depth_test passes INF <= INF) — reported as such.
                    if (seq) seq[pi] = t.written[qi][l] ? t.packets_opaque.ref(t.packet_index[qi][l]).seq : 0xFFFFFFFFu;
                    if (bary1) bary1[pi] = t.bary1[qi][l];
                    if (bary2) bary2[pi] = t.bary2[qi][l];
                    if (color_rgb && shade) {
                        color_rgb[pi * 3 + 0] = t.color[qi].x[l];
                        color_rgb[pi * 3 + 1] = t.color[qi].y[l];
                        color_rgb[pi * 3 + 2] = t.color[qi].z[l];
                    }
                }
            }
    }
    if (st) {
        st->triangles_submitted = o.tris_submitted;
        st->vertices_submitted = o.verts_submitted;
        st->triangles_binned = o.stats.tris_binned;
        st->triangles_clipped = o.stats.tris_clipped;
        st->tile_refs = o.stats.packets;
        st->tile_refs_dup = o.stats.packets_dup;
        st->ndraws = (uint32_t)o.draws.size();
        st->tiles = (uint32_t)o.tiles.size();
        st->ms_clipbin = o.ms_clipbin;
        st->ms_raster = o.ms_raster;
    }
    for (Tile *t : o.tiles) {  // tilerasterizer.rs:109-110
        t->packets_opaque.reset();
        t->packets_translucent.reset();
    }
    return 0;
}

// blit_to_buffer (renderer.rs:293-355) from the tiles' colour buffers of the last shaded frame.
int orc_resolve(void *h, float exposure, int nthreads, uint32_t *out_pixels, orc_stats *st) {
    Oracle &o = *(Oracle *)h;
    double t0 = now_ms();
    int W = o.W;
    int nrows2 = o.H / 2;
#pragma omp parallel for schedule(static) num_threads(nthreads > 1 ? nthreads : 1)
    for (int qy = 0; qy < nrows2; qy++) {
        int pixel_y = qy * 2;
        int tile_y = pixel_y / TILE;
        for (int tx = 0; tx < o.tiles_x; tx++) {
            const Tile &t = *o.tiles[tile_y * o.tiles_x + tx];
            int lq = (pixel_y - t.min_y) / 2;
            int wq = (t.max_x - t.min_x) / 2;
            for (int qx = 0; qx < wq; qx++) {
                V3x4 c = t.color[lq * wq + qx];
                c = c * exposure;
                // tonemap: color / (color + k) * (1.0 + k)
                V3x4 tm = (c / (c + 0.2f)) * (1.0f + 0.2f);
                int bx = t.min_x + qx * 2;
                for (int sy = 0; sy < 2; sy++)
                    for (int sx = 0; sx < 2; sx++) {
                        int px = bx + sx;
                        if (px < W) {
                            int l = sy * 2 + sx;
                            uint32_t r = f2u8(tm.x[l] * 255.0f), g = f2u8(tm.y[l] * 255.0f), b = f2u8(tm.z[l] * 255.0f);
                            out_pixels[(size_t)(pixel_y + sy) * W + px] = (r << 24) | (g << 16) | (b << 8) | 0xFFu;
                        }
                    }
            }
        }
    }
    if (st) st->ms_resolve = now_ms() - t0;
    return 0;
}

// Per-tile digests of the tile state left by the last orc_render (serial schedule), in the reference's own storage order
// (tilerasterizer.rs:25-38: 2x2 quads, lane = 2*(y&1) + (x&1)), so that tools/reference_digest.rs — a few lines a
// maintainer adds to the reference crate — can print the same numbers from the real Rust renderer and pin this oracle:
//   out[3*t + 0] = FNV-1a-64 over, per quad and lane: depth bits, and for lanes with depth != +INF also packet_index
//                  (the tile queue's index: equals the reference's when it runs with RAYON_NUM_THREADS=1), bary1, bary2 bits;
//   out[3*t + 1] = number of lanes with depth != +INF;
//   out[3*t + 2] = FNV-1a-64 over the colour quads (x, y, z lanes) — depends on the host's _mm_rsqrt_ps and libm.
static inline void fnv_u32(uint64_t &h, uint32_t w) {
    for (int k = 0; k < 4; k++) {
        h ^= (w >> (8 * k)) & 0xFFu;
        h *= 0x100000001b3ull;
    }
}
int orc_tile_digests(void *hnd, uint64_t *out) {
    Oracle &o = *(Oracle *)hnd;
    for (size_t ti = 0; ti < o.tiles.size(); ti++) {
        const Tile &t = *o.tiles[ti];
        uint64_t hv = 0xcbf29ce484222325ull, hc = 0xcbf29ce484222325ull, covered = 0;
        for (size_t q = 0; q < t.depth.size(); q++) {
            for (int l = 0; l < 4; l++) {
                float d = t.depth[q][l];
                uint32_t db;
                memcpy(&db, &d, 4);
                fnv_u32(hv, db);
                if (db != 0x7F800000u) {
                    covered++;
                    float b1 = t.bary1[q][l], b2 = t.bary2[q][l];
                    uint32_t u1, u2;
                    memcpy(&u1, &b1, 4);
                    memcpy(&u2, &b2, 4);
                    fnv_u32(hv, t.packet_index[q][l]);
                    fnv_u32(hv, u1);
                    fnv_u32(hv, u2);
                }
            }
            for (int c = 0; c < 3; c++)
                for (int l = 0; l < 4; l++) {
                    float v = c == 0 ? t.color[q].x[l] : (c == 1 ? t.color[q].y[l] : t.color[q].z[l]);
                    uint32_t u;
                    memcpy(&u, &v, 4);
                    fnv_u32(hc, u);
                }
        }
        out[3 * ti + 0] = hv;
        out[3 * ti + 1] = covered;
        out[3 * ti + 2] = hc;
    }
    return (int)o.tiles.size();
}

// update_auto_exposure (renderer.rs:258-290) on a caller-supplied list of per-tile center_luminance values.
// state = {auto_exposure, auto_exposure_target, auto_exposure_ev} (renderer.rs:194-196 start: 2.0, 2.0, log2(2.0)).
int orc_update_auto_exposure(float *state, const float *center_luminance, int sample_count, float delta_time) {
    if (sample_count == 0) return 0;  // :260-262
    std::vector<float> tile_log_luminance((size_t)sample_count);
    for (int i = 0; i < sample_count; i++) {  // :264-267  f32::max ignores a NaN operand, like fmaxf
        tile_log_luminance[i] = log2f(fmaxf(center_luminance[i], 1e-4f));
    }
    // :269-270 sort_unstable_by(total_cmp): IEEE totalOrder (-NaN < -inf < ... < -0 < +0 < ... < +inf < +NaN)
    auto total_key = [](float f) {
        int32_t b;
        memcpy(&b, &f, 4);
        return b ^ (int32_t)(((uint32_t)(b >> 31)) >> 1);
    };
    std::sort(tile_log_luminance.begin(), tile_log_luminance.end(), [&](float a, float b) { return total_key(a) < total_key(b); });
    // :272-276
    size_t trim_count = (size_t)floorf((float)sample_count * 0.10f);
    trim_count = std::min(trim_count, (size_t)(sample_count - 1) / 2);
    float sum = 0.0f;  // iter().sum::<f32>() adds left to right starting from 0.0
    for (size_t i = trim_count; i < (size_t)sample_count - trim_count; i++) sum += tile_log_luminance[i];
    const float mean_log_luminance = sum / (float)((size_t)sample_count - 2 * trim_count);
    // :279 tonemap_inverse_scalar(0.45) (util.rs:43-47), TONEMAP_K = 0.2
    const float y = std::min(std::max(0.45f, 0.0f), 1.0f);
    const float denom = fmaxf(1.0f + 0.2f - y, 1e-6f);
    const float meter_key = fmaxf((y * 0.2f) / denom, 1e-4f);
    // :280-283
    float target_ev = log2f(meter_key) - mean_log_luminance;
    float target = powf(2.0f, target_ev);
    target = target < 0.05f ? 0.05f : (target > 32.0f ? 32.0f : target);  // clamp(AUTO_EXPOSURE_MIN, AUTO_EXPOSURE_MAX); NaN passes through
    state[1] = target;
    // :285-289
    target_ev = log2f(state[1]);
    const float tau = fmaxf(1.0f, 1e-4f);
    const float alpha = 1.0f - expf(-(fmaxf(delta_time, 0.0f) / tau));
    state[2] += (target_ev - state[2]) * alpha;
    state[0] = powf(2.0f, state[2]);
    return 0;
}

// Host-side draw list exactly as the oracle builds it (for cross-checking the product's host mirror).
int orc_build_draws(const swr_scene_desc *scene, const swr_camera *cam, swr_draw *out, int max_draws) {
    Oracle o;
    o.W = o.H = 0;
    o.scene = scene;
    o.cam = cam;
    build_draws(o);
    int n = (int)o.draws.size();
    for (int i = 0; i < n && i < max_draws; i++) {
        memcpy(out[i].model, o.draws[i].model, 64);
        memcpy(out[i].mvp, o.draws[i].mvp, 64);
        out[i].primitive = o.draws[i].primitive;
        out[i].flags = o.draws[i].flags;
        out[i].first_triangle = o.draws[i].first_triangle;
        out[i].reserved = 0;
    }
    return n;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
}

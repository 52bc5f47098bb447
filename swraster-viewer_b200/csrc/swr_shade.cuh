// swr_shade.cuh — K5 visibility-buffer shading and K6 resolve.
//   k_shade     one thread per pixel, 2x2 quads inside a warp (lanes 4q..4q+3) so that the reference's
//               quad-coupled mip selection (shader.rs:130: du_dv * w is lane-wise) is reproduced with shuffles.
//               Reference: tilerasterizer.rs:386-508 (shade_vbuffer, compute_skybox), shader.rs:102-309
//               (pbr_shader<false>), texture.rs:577-864 (samplers), voxelgrid.rs:264-368, util.rs, math.rs.
//   k_luminance per-tile metering value (tilerasterizer.rs:103-106)
//   k_resolve   exposure, tonemap, RGBA8 pack with 128-bit stores (renderer.rs:293-355, util.rs:37-41,98-100)
// Arithmetic is restated operation for operation (translation unit built with -fmad=false, IEEE div/sqrt);
// normalize(): the reference uses the hardware estimate _mm_rsqrt_ps (math.rs:34-39); with the host's table
// installed (swr_set_rsqrt_table) it is emulated bit for bit, otherwise rsqrtf() is used.
#pragma once
#include "swr_raster.cuh"

struct ShadeParams {
    const unsigned long long *keys;
    const TriRecord *records;
    const uint32_t *clip_ext;
    const DevDraw *draws;
    const ClipVertex *clip_verts;
    DevScene scene;
    DevCamera cam;
    float4 *color;  // row-major, padded to whole tiles (Wp x Hp): the reference shades every quad of a tile
    int W, H, tiles_x;
    int Wp, Hp;
    int row_begin, row_end;
    const uint32_t *rsqrt_tab;  // host _mm_rsqrt_ps table (swr_set_rsqrt_table) or NULL
    int rsqrt_bits;
    // sort-last (after the cross-rank key composite): barycentrics of every pixel's winner, whoever owns it; pixels whose
    // winner belongs to another rank carry SWR_ID_FOREIGN and are left unshaded (colour.w = 0); sky only in sky rows.
    const float2 *ext_bary;
    int sky_row_begin, sky_row_end;
    // Fixed-exposure frames (swr_set_fixed_exposure): k_shade<true> applies exposure, tonemap and the RGBA8 pack itself
    // (renderer.rs:293-355) and writes `rgba` (W x H) plus the tile's metering value `lum` (tilerasterizer.rs:103-106); the
    // 16-byte HDR value never goes to memory.
    uint32_t *rgba;
    float *lum;
    float exposure;
    // opaque pass counters: a frame whose tile lists overflowed left the keys of the previous frame in place (the host grows
    // the buffers and replays it); such a frame is not shaded
    const FrameCounters *counters;
};

// Value-domain arithmetic. The translation unit is built with -fmad=false and everything that decides an ADDRESS (texel,
// mip level, voxel cell, cubemap face, LUT cell) keeps the reference's unfused operation order, so those decisions are bit
// for bit the reference's. Arithmetic whose result is only ever a colour VALUE (trilinear SH blend, BRDF terms, colour
// sums) may contract a*b+c into one FMA: the result moves by <= 1 ulp per operation, far inside the +-1 LSB RGBA8 budget
// (measured max / mean error: DESIGN.md 5). -DSWR_SHADE_FMA=0 restores the unfused form (then RGBA8 is bit-exact).
#ifndef SWR_SHADE_FMA
#define SWR_SHADE_FMA 1
#endif
__device__ __forceinline__ float vfma(float a, float b, float c) {  // a * b + c in the value domain
#if SWR_SHADE_FMA
    return __fmaf_rn(a, b, c);
#else
    return a * b + c;
#endif
}

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
// Value-domain quotient: reciprocal estimate times numerator (MUFU.RCP + FMUL, <= 2 ulp) instead of the IEEE division
// sequence (~9 instructions). Only for colour values; anything that decides an address keeps `/`.
__device__ __forceinline__ float vdiv(float a, float b) {
#if SWR_SHADE_FMA
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float b) { return V3{a.x * b, a.y * b, a.z * b}; }
__device__ __forceinline__ V3 operator+(V3 a, float b) { return V3{a.x + b, a.y + b, a.z + b}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // math.rs:88-90
__device__ __forceinline__ float vdot(V3 a, V3 b) { return vfma(a.z, b.z, vfma(a.y, b.y, a.x * b.x)); }  // value domain
__device__ __forceinline__ V3 vfma3(V3 a, V3 b, V3 c) { return V3{vfma(a.x, b.x, c.x), vfma(a.y, b.y, c.y), vfma(a.z, b.z, c.z)}; }
__device__ __forceinline__ V3 vfma3(V3 a, float b, V3 c) { return V3{vfma(a.x, b, c.x), vfma(a.y, b, c.y), vfma(a.z, b, c.z)}; }
// math.rs:34-39 rsqrt_vec. With a host table: exact emulation of _mm_rsqrt_ps (value = table[parity][top mantissa
// bits] scaled by the even part of the exponent; zero/denormal -> inf, negative -> NaN, inf -> 0).
struct Rsq {
    const uint32_t *tab;
    int bits;
};
__device__ __forceinline__ float rsqrt_ref(float x, Rsq q) {
    if (q.tab == nullptr) return rsqrtf(x);
    const uint32_t b = __float_as_uint(x);
    const uint32_t e = (b >> 23) & 0xFFu;
    if (b - 0x00800000u >= 0x7F000000u) {  // not a positive normal number: zero / denormal, inf / NaN, negative
        if (e == 0u) return __uint_as_float((b & 0x80000000u) | 0x7F800000u);
        if (e == 255u) return (b & 0x007FFFFFu) ? __uint_as_float(b | 0x00400000u) : ((b >> 31) ? __uint_as_float(0xFFC00000u) : 0.0f);
        return __uint_as_float(0xFFC00000u);
    }
    const int ue = (int)e - 127;
    const int p = ue & 1;
    const int k = (ue - p) >> 1;
    const uint32_t t = __ldg(q.tab + (((uint32_t)p << q.bits) | ((b & 0x007FFFFFu) >> (23 - q.bits))));
    return __uint_as_float(t - ((uint32_t)k << 23));
}
__device__ __forceinline__ V3 normalize(V3 a, Rsq q) {  // math.rs:101-108
    float r = rsqrt_ref(dot(a, a), q);
    return V3{a.x * r, a.y * r, a.z * r};
}
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ float srgb1(float x) { return x * vfma(x, vfma(x, 0.305306011f, 0.682171111f), 0.012522878f); }
__device__ __forceinline__ V3 srgb_to_linear_fast(V3 x) { return V3{srgb1(x.x), srgb1(x.y), srgb1(x.z)}; }  // util.rs:111-113

__device__ __forceinline__ V3 sample_cubemap_rgb(const DevTex &t, V3 n, uint32_t mip) {  // texture.rs:593-663
    float ax = fabsf(n.x), ay = fabsf(n.y), az = fabsf(n.z);
    bool mx = (ax >= ay) && (ax >= az);
    bool my = (ay > ax) && (ay >= az);
    float sx = n.x >= 0.0f ? 1.0f : -1.0f, sy = n.y >= 0.0f ? 1.0f : -1.0f, sz = n.z >= 0.0f ? 1.0f : -1.0f;
    float u = mx ? (-n.z * sx) : (my ? n.x : n.x * sz);
    float v = mx ? (-n.y) : (my ? n.z * sy : -n.y);
    float den = mx ? ax : (my ? ay : az);
    den = sse_max(den, 1.0e-19f);
    float uf = 0.5f * (u / den + 1.0f), vf = 0.5f * (v / den + 1.0f);
    uint32_t slice = mx ? (n.x >= 0.0f ? 0u : 1u) : (my ? (n.y >= 0.0f ? 2u : 3u) : (n.z >= 0.0f ? 4u : 5u));
    mip = min(mip, t.max_mip);
    float wf = (float)t.mip_w[mip], hf = (float)t.mip_h[mip];
    float x = rintf(uf * (wf - 1.0f)), y = rintf(vf * (hf - 1.0f));  // glam Vec4::round: half to even
    uint32_t xi = __float2uint_rz(sse_clamp(x, 0.0f, wf - 1.0f)), yi = __float2uint_rz(sse_clamp(y, 0.0f, hf - 1.0f));
    float4 c = fetch_texel(t, t.mip_off[mip] + slice * t.stride[mip] + yi * t.mip_w[mip] + xi);
    return v3(c.x, c.y, c.z);
}
__device__ __forceinline__ V3 sample_cubemap_trilinear_rgb(const DevTex &t, V3 n, float mip_level) {  // texture.rs:665-678
    float maxm = (float)t.max_mip;
    float mip = sse_clamp(mip_level, 0.0f, maxm);
    float m0 = floorf(mip), m1 = sse_min(m0 + 1.0f, maxm), tt = mip - m0;
    V3 c0 = sample_cubemap_rgb(t, n, __float2uint_rz(m0));
    V3 c1 = sample_cubemap_rgb(t, n, __float2uint_rz(m1));
    return v3(vfma(c1.x - c0.x, tt, c0.x), vfma(c1.y - c0.y, tt, c0.y), vfma(c1.z - c0.z, tt, c0.z));
}
__device__ __forceinline__ V3 sample_bilinear_rgb0(const DevTex &t, float u, float v) {  // texture.rs:730-790, mip 0 slice 0
    float wf = (float)t.mip_w[0], hf = (float)t.mip_h[0];
    uint32_t wi = t.mip_w[0], off = t.mip_off[0];
    float xf = u * wf - 0.5f, yf = v * hf - 0.5f;
    float x0 = floorf(xf), y0 = floorf(yf), x1 = x0 + 1.0f, y1 = y0 + 1.0f;
    float fx = xf - x0, fy = yf - y0, omfx = 1.0f - fx, omfy = 1.0f - fy;
    uint32_t x0i = __float2uint_rz(apply_wrap_mode(x0, wf, t.wrap_s)), y0i = __float2uint_rz(apply_wrap_mode(y0, hf, t.wrap_t));
    uint32_t x1i = __float2uint_rz(apply_wrap_mode(x1, wf, t.wrap_s)), y1i = __float2uint_rz(apply_wrap_mode(y1, hf, t.wrap_t));
    float4 p00 = fetch_texel(t, off + y0i * wi + x0i), p10 = fetch_texel(t, off + y0i * wi + x1i);
    float4 p01 = fetch_texel(t, off + y1i * wi + x0i), p11 = fetch_texel(t, off + y1i * wi + x1i);
    float w00 = omfx * omfy, w10 = fx * omfy, w01 = omfx * fy, w11 = fx * fy;
    return v3(vfma(p11.x, w11, vfma(p01.x, w01, vfma(p10.x, w10, p00.x * w00))), vfma(p11.y, w11, vfma(p01.y, w01, vfma(p10.y, w10, p00.y * w00))),
              vfma(p11.z, w11, vfma(p01.z, w01, vfma(p10.z, w10, p00.z * w00))));
}

// voxelgrid.rs:264-368 for one position.
__device__ __forceinline__ uint32_t f2usize_clamped(float f, uint32_t hi) { return min(__float2uint_rz(f), hi); }
__device__ __forceinline__ void sample_gi(const DevScene &s, V3 pos, V3 rgb[4], float w[4]) {
    const uint32_t W = s.gdim[0], H = s.gdim[1], D = s.gdim[2];
    const float vsx = s.gvs[0], vsy = s.gvs[1], vsz = s.gvs[2];  // (world_max - world_min) / dims, voxelgrid.rs:154-161
    // value-domain quotients although they pick the voxel: the trilinear blend is continuous across cell boundaries (a cell
    // flip by a last-bit difference swaps (cell k, weight ~1) for (cell k+1, weight ~0)), so 2 ulp in the grid coordinate is
    // 2 ulp in the colour, nothing more
    float vx = vdiv(pos.x - s.gmin[0], vsx), vy = vdiv(pos.y - s.gmin[1], vsy), vz = vdiv(pos.z - s.gmin[2], vsz);
    float x0f = floorf(vx), y0f = floorf(vy), z0f = floorf(vz);
    uint32_t x0 = f2usize_clamped(x0f, W - 1), y0 = f2usize_clamped(y0f, H - 1), z0 = f2usize_clamped(z0f, D - 1);
    uint32_t x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1), z1 = min(z0 + 1, D - 1);
    float fx = vx - x0f, fy = vy - y0f, fz = vz - z0f;
    float ox = 1.0f - fx, oy = 1.0f - fy, oz = 1.0f - fz;
    const size_t WH = (size_t)W * H;
    const float4 *g000 = s.gi + ((size_t)z0 * WH + (size_t)y0 * W + x0) * 4, *g001 = s.gi + ((size_t)z1 * WH + (size_t)y0 * W + x0) * 4;
    const float4 *g010 = s.gi + ((size_t)z0 * WH + (size_t)y1 * W + x0) * 4, *g011 = s.gi + ((size_t)z1 * WH + (size_t)y1 * W + x0) * 4;
    const float4 *g100 = s.gi + ((size_t)z0 * WH + (size_t)y0 * W + x1) * 4, *g101 = s.gi + ((size_t)z1 * WH + (size_t)y0 * W + x1) * 4;
    const float4 *g110 = s.gi + ((size_t)z0 * WH + (size_t)y1 * W + x1) * 4, *g111 = s.gi + ((size_t)z1 * WH + (size_t)y1 * W + x1) * 4;
    // Value domain (colour only): the eight corner weights once, then every SH coefficient is a weighted sum of its eight
    // corners. Two channels at a time with Blackwell's packed fp32 pipe (fma.rn.f32x2 -> FFMA2): each 128-bit corner fetch
    // is already two aligned register pairs (xy, zw), so the 4 x 4 channels cost 64 packed operations instead of 196 scalar
    // ones. Differs from the reference's nested lerp (voxelgrid.rs:300-368) by rounding only (a few ulp).
#if SWR_SHADE_FMA
    const float wy0z0 = oy * oz, wy1z0 = fy * oz, wy0z1 = oy * fz, wy1z1 = fy * fz;
    const unsigned long long w000 = pack2(ox * wy0z0), w100 = pack2(fx * wy0z0), w010 = pack2(ox * wy1z0), w110 = pack2(fx * wy1z0);
    const unsigned long long w001 = pack2(ox * wy0z1), w101 = pack2(fx * wy0z1), w011 = pack2(ox * wy1z1), w111 = pack2(fx * wy1z1);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const ulonglong2 v000 = ldg_pairs(g000 + c), v001 = ldg_pairs(g001 + c), v010 = ldg_pairs(g010 + c), v011 = ldg_pairs(g011 + c);
        const ulonglong2 v100 = ldg_pairs(g100 + c), v101 = ldg_pairs(g101 + c), v110 = ldg_pairs(g110 + c), v111 = ldg_pairs(g111 + c);
#define SWR_SUM8(F) fma2(v111.F, w111, fma2(v011.F, w011, fma2(v101.F, w101, fma2(v001.F, w001, fma2(v110.F, w110, fma2(v010.F, w010, fma2(v100.F, w100, mul2(v000.F, w000))))))))
        const unsigned long long xy = SWR_SUM8(x), zw = SWR_SUM8(y);
#undef SWR_SUM8
        float zz, ww;
        unpack2(xy, rgb[c].x, rgb[c].y);
        unpack2(zw, zz, ww);
        rgb[c].z = zz;
        if (c < 2) w[c] = ww;  // only the .w of coefficients 0 and 1 is consumed (shader.rs:172-173)
    }
#else
#pragma unroll
    for (int c = 0; c < 4; c++) {
        float4 v000 = __ldg(g000 + c), v001 = __ldg(g001 + c), v010 = __ldg(g010 + c), v011 = __ldg(g011 + c);
        float4 v100 = __ldg(g100 + c), v101 = __ldg(g101 + c), v110 = __ldg(g110 + c), v111 = __ldg(g111 + c);
#define SWR_TRI(F)                                                                                   \
    ({                                                                                               \
        float v00 = vfma(v100.F, fx, v000.F * ox), v01 = vfma(v101.F, fx, v001.F * ox);              \
        float v10 = vfma(v110.F, fx, v010.F * ox), v11 = vfma(v111.F, fx, v011.F * ox);              \
        float v0 = vfma(v10, fy, v00 * oy), v1 = vfma(v11, fy, v01 * oy);                            \
        vfma(v1, fz, v0 * oz);                                                                       \
    })
        rgb[c] = v3(SWR_TRI(x), SWR_TRI(y), SWR_TRI(z));
        if (c < 2) w[c] = SWR_TRI(w);  // only the .w of coefficients 0 and 1 is consumed (shader.rs:172-173)
#undef SWR_TRI
    }
#endif
}

// Per-packet data the shader interpolates (renderer.rs:697-754, util.rs:149-194).
struct ShadePacket {
    V3 n_a, n_da, n_db;
    V3 t_a, t_da, t_db;
    float ts_a, ts_da, ts_db;
    float u_a, u_da, u_db, v_a, v_da, v_db;
    V3 p_a, p_da, p_db;
    float du_dv[4];
    uint32_t material;
};

__device__ __forceinline__ void build_shade_packet(const ShadeParams &P, const TriRecord &r, ShadePacket &sp) {
    const DevDraw &dr = P.draws[r.draw & SWR_REC_DRAW_MASK];
    const DevPrim &pr = P.scene.prims[dr.prim];
    V3 n[3], t[3], pw[3];
    float tw[3], uu[3], vv[3];
    if (r.clip == SWR_NO_CLIP) {  // renderer.rs:494-560
        const uint32_t tri = (r.seq >> 3) - dr.first_tri;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            uint32_t iv = __ldg(pr.idx + 3 * tri + k);
            float4 wp = mul_vec4(dr.model, __ldg(pr.pos + iv));
            float4 n4 = __ldg(pr.nrm + iv), t4 = __ldg(pr.tan + iv);
            float3 nw = mul_mat3(dr.model, make_float3(n4.x, n4.y, n4.z));
            float3 tv = mul_mat3(dr.model, make_float3(t4.x, t4.y, t4.z));
            float2 uv = __ldg(pr.uv + iv);
            pw[k] = v3(wp.x, wp.y, wp.z);
            n[k] = v3(nw.x, nw.y, nw.z);
            t[k] = v3(tv.x, tv.y, tv.z);
            tw[k] = t4.w;
            uu[k] = uv.x;
            vv[k] = uv.y;
        }
    } else {  // fan (0, j, j+1) of the clipped polygon (renderer.rs:655-664)
        const uint32_t fan = r.seq & 7u;
        const uint32_t vi[3] = {r.clip, r.clip + fan + 1u, r.clip + fan + 2u};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float4 *src = reinterpret_cast<const float4 *>(P.clip_verts + vi[k]);
            float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
            pw[k] = v3(a.x, a.y, a.z);
            n[k] = v3(a.w, b.x, b.y);
            t[k] = v3(b.z, b.w, c.x);
            tw[k] = c.y;
            uu[k] = c.z;
            vv[k] = c.w;
        }
    }
    const float iw[3] = {r.iw0, r.iw1, r.iw2};
    float uw[3], vw[3];
    V3 pww[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uw[k] = uu[k] * iw[k];
        vw[k] = vv[k] * iw[k];
        pww[k] = pw[k] * iw[k];
    }
    sp.n_a = n[0]; sp.n_da = n[1] - n[0]; sp.n_db = n[2] - n[0];
    sp.t_a = t[0]; sp.t_da = t[1] - t[0]; sp.t_db = t[2] - t[0];
    sp.ts_a = tw[0]; sp.ts_da = tw[1] - tw[0]; sp.ts_db = tw[2] - tw[0];
    sp.u_a = uw[0]; sp.u_da = uw[1] - uw[0]; sp.u_db = uw[2] - uw[0];
    sp.v_a = vw[0]; sp.v_da = vw[1] - vw[0]; sp.v_db = vw[2] - vw[0];
    sp.p_a = pww[0]; sp.p_da = pww[1] - pww[0]; sp.p_db = pww[2] - pww[0];
    // renderer.rs:738-754
    float dx1 = i2f(wsub(r.X1, r.X0)), dx2 = i2f(wsub(r.X2, r.X0)), dy1 = i2f(wsub(r.Y1, r.Y0)), dy2 = i2f(wsub(r.Y2, r.Y0));
    float du1 = uw[1] - uw[0], du2 = uw[2] - uw[0], dv1 = vw[1] - vw[0], dv2 = vw[2] - vw[0];
    float s = r.ooa * 16.0f;
    sp.du_dv[0] = (du1 * dy2 - du2 * dy1) * s;
    sp.du_dv[1] = (du2 * dx1 - du1 * dx2) * s;
    sp.du_dv[2] = (dv1 * dy2 - dv2 * dy1) * s;
    sp.du_dv[3] = (dv2 * dx1 - dv1 * dx2) * s;
    sp.material = pr.material;
}

__device__ __forceinline__ float interp1(float a, float da, float db, float b1, float b2) { return a + b1 * da + b2 * db; }
__device__ __forceinline__ V3 interp3(V3 a, V3 da, V3 db, float b1, float b2) {
    return v3(a.x + b1 * da.x + b2 * db.x, a.y + b1 * da.y + b2 * db.y, a.z + b1 * da.z + b2 * db.z);
}

// shader.rs:110-309, TRANSLUCENT = false. `du_dv` is already scaled lane-wise by the quad's w values.
// shader.rs:110-309; `translucent` adds the KHR_materials_transmission blend over `current` (shader.rs:265-277).
__device__ __forceinline__ V3 pbr_shader(const ShadeParams &P, const ShadePacket &sp, float b1, float b2, float w, const float du_dv[4],
                                         bool translucent = false, V3 current = V3{0.0f, 0.0f, 0.0f}) {
    const DevScene &sc = P.scene;
    const DevMat &mat = sc.mats[sp.material];
    const float EPS = 1e-6f, PI = 3.14159265358979323846f;
    const Rsq rq{P.rsqrt_tab, P.rsqrt_bits};
    V3 input_normal = normalize(interp3(sp.n_a, sp.n_da, sp.n_db, b1, b2), rq);
    V3 input_tangent = normalize(interp3(sp.t_a, sp.t_da, sp.t_db, b1, b2), rq);
    float tangent_sign = interp1(sp.ts_a, sp.ts_da, sp.ts_db, b1, b2);
    V3 pos_world = interp3(sp.p_a, sp.p_da, sp.p_db, b1, b2) * w;
    float uv_x = interp1(sp.u_a, sp.u_da, sp.u_db, b1, b2) * w;
    float uv_y = interp1(sp.v_a, sp.v_da, sp.v_db, b1, b2) * w;

    V3 tangent_world = normalize(input_tangent - input_normal * dot(input_normal, input_tangent), rq);
    float handed = tangent_sign >= 0.0f ? 1.0f : -1.0f;
    V3 bitangent_world = cross(input_normal, tangent_world) * handed;
    V3 normal_world = normalize(input_normal, rq);
    if (mat.tex_normal >= 0) {
        float4 s = sample4(sc.texs[mat.tex_normal], uv_x, uv_y, du_dv);
        V3 tsn = v3(s.x, s.y, s.z) * 2.0f + (-1.0f);
        normal_world = tangent_world * tsn.x + bitangent_world * tsn.y + input_normal * tsn.z;
        normal_world = normalize(normal_world, rq);
    }
    V3 light_dir = v3(sc.light_dir[0], sc.light_dir[1], sc.light_dir[2]);
    V3 light_color = v3(sc.light_color[0], sc.light_color[1], sc.light_color[2]);
    V3 view_dir = v3(P.cam.position[0], P.cam.position[1], P.cam.position[2]) - pos_world;
    V3 view_normal = normalize(view_dir, rq);

    float n_dot_l = sse_max(vdot(normal_world, light_dir), 0.0f);
    V3 half_vector = normalize(light_dir + view_normal, rq);
    float n_dot_h = sse_max(vdot(normal_world, half_vector), 0.0f);
    float n_dot_v = sse_max(dot(normal_world, view_normal), 1.0e-4f);  // addresses the BRDF LUT: unfused
    float v_dot_h = sse_max(vdot(view_normal, half_vector), 0.0f);

    V3 gi_rgb[4];
    float gi_w[4];
    sample_gi(sc, pos_world, gi_rgb, gi_w);
    float voxel_light_intensity = sse_clamp(gi_w[0], 0.0f, 1.0f);
    float sky_visibility = sse_clamp(gi_w[1], 0.0f, 1.0f);

    V3 base = v3(mat.base[0], mat.base[1], mat.base[2]);
    if (mat.tex_base >= 0) {
        float4 s = sample4(sc.texs[mat.tex_base], uv_x, uv_y, du_dv);
        base = base * srgb_to_linear_fast(v3(s.x, s.y, s.z));
    }
    float roughness = mat.roughness, metallic = mat.metallic;
    if (mat.tex_mr >= 0) {
        float4 s = sample4(sc.texs[mat.tex_mr], uv_x, uv_y, du_dv);
        roughness = roughness * s.y;
        metallic = metallic * s.z;
    }
    roughness = sse_clamp(roughness, 0.045f, 1.0f);
    metallic = sse_clamp(metallic, 0.0f, 1.0f);
    float ao = 1.0f;
    if (mat.tex_occlusion >= 0) {
        ao = sample4(sc.texs[mat.tex_occlusion], uv_x, uv_y, du_dv).x;
        ao = vfma(ao - 1.0f, mat.occlusion_strength, 1.0f);
    }
    V3 f0 = v3(vfma(base.x - 0.04f, metallic, 0.04f), vfma(base.y - 0.04f, metallic, 0.04f), vfma(base.z - 0.04f, metallic, 0.04f));
    float omvh = 1.0f - v_dot_h;
    float omvh2 = omvh * omvh;
    float omvh4 = omvh2 * omvh2;
    float omvh5 = omvh4 * omvh;
    V3 one3 = v3(1.0f, 1.0f, 1.0f);
    V3 brdf_f_direct = vfma3(one3 - f0, omvh5, f0);

    float alpha = roughness * roughness;
    float alpha_2 = alpha * alpha;
    float ndh_2 = n_dot_h * n_dot_h;
    float denom_d = vfma(ndh_2, alpha_2 - 1.0f, 1.0f);
    float brdf_d = vdiv(alpha_2, vfma(PI, denom_d * denom_d, EPS));
    float k = roughness + 1.0f;
    k = (k * k) * 0.125f;
    float gv = vdiv(n_dot_v, vfma(n_dot_v, 1.0f - k, k) + EPS);
    float gl = vdiv(n_dot_l, vfma(n_dot_l, 1.0f - k, k) + EPS);
    float brdf_g = gv * gl;
    float specular_dg = vdiv(brdf_d * brdf_g, vfma(4.0f * n_dot_l, n_dot_v, EPS));
    V3 k_d_direct = (one3 - brdf_f_direct) * (1.0f - metallic);
    const float INV_PI = 1.0f / 3.14159265358979323846f;
    V3 lambert = base * INV_PI;
    V3 color_direct_diffuse = light_color * k_d_direct * lambert * n_dot_l * voxel_light_intensity;
    V3 color_direct_specular = light_color * brdf_f_direct * specular_dg * n_dot_l * voxel_light_intensity;

    V3 k_d_indirect = (one3 - f0) * (1.0f - metallic);
    V3 irr = vfma3(gi_rgb[3], normal_world.x, vfma3(gi_rgb[2], normal_world.z, vfma3(gi_rgb[1], normal_world.y, gi_rgb[0])));  // shader.rs:102-108
    irr = v3(sse_max(irr.x, 0.0f), sse_max(irr.y, 0.0f), sse_max(irr.z, 0.0f));
    V3 color_indirect_diffuse = irr * base * k_d_indirect * INV_PI;

    V3 vneg = view_normal * -1.0f;
    V3 reflect_dir = vneg - normal_world * dot(vneg, normal_world) * 2.0f;  // math.rs:118-120
    const DevTex &spec = sc.texs[sc.cubemap_specular];
    V3 prefiltered_env = sample_cubemap_trilinear_rgb(spec, reflect_dir, roughness * (float)spec.max_mip);
    V3 lut = sample_bilinear_rgb0(sc.texs[sc.brdf_lut], sse_clamp(n_dot_v, 0.0f, 1.0f), sse_clamp(roughness, 0.0f, 1.0f));
    V3 brdf_spec_factor = v3(vfma(f0.x, lut.x, lut.y), vfma(f0.y, lut.x, lut.y), vfma(f0.z, lut.x, lut.y));
    V3 color_indirect_specular = prefiltered_env * brdf_spec_factor;
    float ao_spec = vfma(ao - 1.0f, 0.5f, 1.0f);
    color_indirect_specular = color_indirect_specular * (ao_spec * sky_visibility);

    V3 color;
    if (translucent) {
        float transmission = mat.transmission;
        if (mat.tex_transmission >= 0) transmission = transmission * sample4(sc.texs[mat.tex_transmission], uv_x, uv_y, du_dv).x;
        transmission = sse_clamp(transmission, 0.0f, 1.0f);
        const float inv_transmission = 1.0f - transmission;
        color = (current * base * transmission) + ((color_direct_diffuse + color_indirect_diffuse) * inv_transmission) + color_direct_specular +
                color_indirect_specular;
    } else {
        color = color_direct_diffuse + color_direct_specular + color_indirect_diffuse + color_indirect_specular;
    }
    V3 emissive_mat = one3;
    if (mat.tex_emissive >= 0) {
        float4 s = sample4(sc.texs[mat.tex_emissive], uv_x, uv_y, du_dv);
        emissive_mat = srgb_to_linear_fast(v3(s.x, s.y, s.z));
    }
    color = vfma3(emissive_mat, v3(mat.emissive[0], mat.emissive[1], mat.emissive[2]), color);
    return color;
}

__device__ __forceinline__ V3 compute_skybox(const ShadeParams &P, int px, int py) {  // tilerasterizer.rs:478-508
    float pixel_x = (float)px + 0.5f, pixel_y = (float)py + 0.5f;
    float ndc_x = pixel_x * P.cam.one_over_width * 2.0f - 1.0f;
    float ndc_y = (1.0f - pixel_y * P.cam.one_over_height) * 2.0f - 1.0f;
    const float *m = P.cam.skybox_T;
    V3 d = v3(ndc_x * m[0] + ndc_y * m[1] + 1.0f * m[2] + m[3], ndc_x * m[4] + ndc_y * m[5] + 1.0f * m[6] + m[7],
              ndc_x * m[8] + ndc_y * m[9] + 1.0f * m[10] + m[11]);
    const Rsq rq{P.rsqrt_tab, P.rsqrt_bits};
    return srgb_to_linear_fast(sample_cubemap_rgb(P.scene.texs[P.scene.cubemap], normalize(d, rq), 0u));
}

__device__ __forceinline__ uint32_t resolve_pixel_rgba(float4 c, float exposure);

#ifndef SWR_SHADE_MINB
#define SWR_SHADE_MINB 8  // 64 registers: measured best (0.75 ms vs 0.95 ms at 4 blocks/SM on C3)
#endif
#define SHADE_BLOCK 128
#define SHADE_ROWS (SHADE_BLOCK / 32 * 2)
template <bool FUSED>
__global__ void __launch_bounds__(SHADE_BLOCK, SWR_SHADE_MINB) k_shade(ShadeParams P) {
    if (P.counters->overflow_refs | P.counters->overflow_ext | P.counters->overflow_clip) return;  // stale keys: the frame is replayed
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int quad = lane >> 2, sub = lane & 3;
    const int px = blockIdx.x * 16 + quad * 2 + (sub & 1);
    const int py = P.row_begin * SWR_TILE + blockIdx.y * SHADE_ROWS + warp * 2 + (sub >> 1);
    const bool inside = px < P.Wp && py < P.Hp;  // off-screen quads of edge tiles are shaded too (metering, tilerasterizer.rs:392-475)
    const unsigned qmask = 0xFu << (lane & ~3);

    unsigned long long key = inside ? load_key(P.keys, P.tiles_x, px, py) : SWR_KEY_EMPTY;
    uint32_t slot = 0xFFFFFFFFu;
    float b1 = 0.0f, b2 = 0.0f;
    bool covered = false;
    TriRecord rec;
    bool foreign = false;
    if (key != SWR_KEY_EMPTY) {
        slot = 0xFFFFFFFFu - (uint32_t)key;
        if (slot == SWR_ID_FOREIGN) {
            foreign = true;
        } else {
            rec = P.records[record_of_id(slot, P.clip_ext)];
            // the key's owner covers the pixel (k_read_vis re-checks that in the parity read-back): barycentrics only
            const float z = resolve_owner(rec, P.W, P.H, px, py, b1, b2);
            covered = __float_as_uint(z) != SWR_INF_BITS;  // depth.cmpne(INF)
        }
    }
    if (P.ext_bary != nullptr && inside && px < P.W && py < P.H) {
        const float2 eb = P.ext_bary[(size_t)py * P.W + px];
        b1 = eb.x;
        b2 = eb.y;
    }
    V3 out = v3(0.0f, 0.0f, 0.0f);
    // tilerasterizer.rs:419-463 evaluates pbr_shader once per distinct packet of the quad, on all four lanes, and keeps
    // the lanes that own the packet. The only cross-lane term is du_dv * w (shader.rs:130: component k is scaled by
    // lane k's w, where w = 1 / one_over_w(P) at lane k's STORED barycentrics, 0,0 for never-written lanes). So each
    // lane shades its own pixel once, with its own packet P evaluated at the four lanes' barycentrics.
    float qb1[4], qb2[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        qb1[j] = __shfl_sync(qmask, b1, (lane & ~3) + j);
        qb2[j] = __shfl_sync(qmask, b2, (lane & ~3) + j);
    }
    if (covered) {
        const float iwda = rec.iw1 - rec.iw0, iwdb = rec.iw2 - rec.iw0;
        ShadePacket sp;
        build_shade_packet(P, rec, sp);
        const float w = 1.0f / interp1(rec.iw0, iwda, iwdb, b1, b2);  // shader.rs:123, my lane
        float dd[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const DevMat &mat = P.scene.mats[sp.material];
        if ((mat.tex_base & mat.tex_mr & mat.tex_normal & mat.tex_emissive & mat.tex_occlusion) >= 0) {
            // du_dv * w is lane-wise over the quad (shader.rs:130) and only texture sampling (mip selection) reads it
#pragma unroll
            for (int j = 0; j < 4; j++) dd[j] = sp.du_dv[j] * (1.0f / interp1(rec.iw0, iwda, iwdb, qb1[j], qb2[j]));
        }
        out = pbr_shader(P, sp, b1, b2, w, dd);
    }
    bool wrote = covered;
    if (!covered && inside && !(foreign && (uint32_t)(key >> 32) != 0xFF800000u)) {  // foreign winners with finite depth are shaded by their owner
        const int trow = py >> 6;
        if (P.ext_bary == nullptr || (trow >= P.sky_row_begin && trow < P.sky_row_end)) {
            out = compute_skybox(P, px, py);
            wrote = true;
        }
    }
    if (!FUSED) {
        if (inside) P.color[(size_t)py * P.Wp + px] = make_float4(out.x, out.y, out.z, wrote ? 1.0f : 0.0f);
    } else {
        if (px < P.W && py < P.H) P.rgba[(size_t)py * P.W + px] = resolve_pixel_rgba(make_float4(out.x, out.y, out.z, wrote ? 1.0f : 0.0f), P.exposure);
        // metering quad of the tile = pixels (0..1, 32..33): k_luminance's sum, in its order, inside the quad's four lanes
        if ((px & (SWR_TILE - 1)) < 2 && ((py & (SWR_TILE - 1)) >> 1) == 16 && inside) {  // quad-uniform (and the whole quad is inside)
            const float l = out.x * 0.2126f + out.y * 0.7152f + out.z * 0.0722f;
            const float l0 = __shfl_sync(qmask, l, (lane & ~3)), l1 = __shfl_sync(qmask, l, (lane & ~3) + 1);
            const float l2 = __shfl_sync(qmask, l, (lane & ~3) + 2), l3 = __shfl_sync(qmask, l, (lane & ~3) + 3);
            if (sub == 0) P.lum[(py >> 6) * P.tiles_x + (px >> 6)] = (l0 + l1 + l2 + l3) * 0.25f;
        }
    }
}

// tilerasterizer.rs:103-106: quad #512 of the tile = pixels (0..1, 32..33)
__global__ void k_luminance(const float4 *color, float *lum, int W, int H, int tiles_x, int ntiles, int tile_begin, int tile_end) {
    int t = tile_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tile_end || t >= ntiles) return;
    int x0 = (t % tiles_x) * SWR_TILE, y0 = (t / tiles_x) * SWR_TILE + 32;
    float l[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int x = x0 + (k & 1), y = y0 + (k >> 1);
        float4 c = (x < W && y < H) ? color[(size_t)y * W + x] : make_float4(0, 0, 0, 0);
        l[k] = c.x * 0.2126f + c.y * 0.7152f + c.z * 0.0722f;
    }
    lum[t] = (l[0] + l[1] + l[2] + l[3]) * 0.25f;
}

// renderer.rs:293-355. One thread = 4 horizontally adjacent pixels -> one 128-bit store.
__device__ __forceinline__ uint32_t f2u8(float f) { return min(__float2uint_rz(f), 255u); }  // Rust `as u8`
__device__ __forceinline__ uint32_t resolve_pixel_rgba(float4 c, float exposure) {
    if (c.w == 0.0f) return 0u;  // not shaded by this rank (sort-last): contributes nothing to the sum-combine
    float r = c.x * exposure, g = c.y * exposure, b = c.z * exposure;
    const float k = 0.2f, opk = 1.0f + 0.2f;  // util.rs:37-41
    r = vdiv(r, r + k) * opk;  // value domain: a 2-ulp quotient moves a byte only when the product sits on an integer
    g = vdiv(g, g + k) * opk;
    b = vdiv(b, b + k) * opk;
    return (f2u8(r * 255.0f) << 24) | (f2u8(g * 255.0f) << 16) | (f2u8(b * 255.0f) << 8) | 0xFFu;
}
// `rgba` != NULL: the frame was shaded with a fixed exposure and is already packed: the kernel only moves it.
__global__ void __launch_bounds__(256) k_resolve(const float4 *color, int Wp, uint32_t *pixels, int W, int y0, int y1, float exposure, const uint32_t *rgba) {
    const int x = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
    const int y = y0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    if (y >= y1 || x >= W) return;
    const float4 *c = color + (size_t)y * Wp + x;
    uint32_t *o = pixels + (size_t)y * W + x;
    if (rgba) {
        const uint32_t *src = rgba + (size_t)y * W + x;
        if (x + 3 < W && (W & 3) == 0) {
            *reinterpret_cast<uint4 *>(o) = *reinterpret_cast<const uint4 *>(src);
        } else {
            for (int k = 0; k < 4 && x + k < W; k++) o[k] = src[k];
        }
        return;
    }
    if (x + 3 < W && (W & 3) == 0) {
        uint4 v;
        v.x = resolve_pixel_rgba(c[0], exposure);
        v.y = resolve_pixel_rgba(c[1], exposure);
        v.z = resolve_pixel_rgba(c[2], exposure);
        v.w = resolve_pixel_rgba(c[3], exposure);
        *reinterpret_cast<uint4 *>(o) = v;
    } else {
        for (int k = 0; k < 4 && x + k < W; k++) o[k] = resolve_pixel_rgba(c[k], exposure);
    }
}

// Peer frame assembly (sort-first): the pixel buffer of the assembling rank carries SWR_PEER_WORDS control words behind
// its W*H pixels: [0] = contributions received (monotonic count), [16] = frame number contributors may write.
#define SWR_PEER_WORDS 64
#define SWR_PEER_DONE 0
#define SWR_PEER_FREE 16
#define SWR_PEER_TIMEOUT_NS 2000000000ull

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spin until *ctr >= target (system scope); false after SWR_PEER_TIMEOUT_NS
__device__ __forceinline__ bool peer_wait(const uint32_t *ctr, uint32_t target) {
    const unsigned long long t0 = global_timer_ns();
    while ((int32_t)(ld_acquire_sys(ctr) - target) < 0) {
        if (global_timer_ns() - t0 > SWR_PEER_TIMEOUT_NS) return false;
        __nanosleep(200);
    }
    return true;
}

// k_resolve into the assembling rank's buffer: store this rank's rows over NVLink and let the last CTA publish the
// contribution (fence + system-scope atomic on the assembler's counter). The buffer-reuse guard (k_peer_wait) runs
// as a one-thread kernel in front so that only one remote poll is in flight instead of one per CTA.
// A frame whose tile lists / clip buffers overflowed holds no valid image: it contributes nothing and does NOT signal — the
// host replays the frame (and this resolve with it) when it next looks at the frame's counters, and the assembler simply
// keeps waiting for that contribution.
__device__ __forceinline__ bool frame_overflowed(const FrameCounters *c) { return c != nullptr && (c->overflow_refs | c->overflow_clip | c->overflow_ext) != 0u; }

__global__ void __launch_bounds__(256) k_resolve_peer(const float4 *color, int Wp, uint32_t *pixels, uint32_t *ctrl, int W, int y0, int y1, float exposure,
                                                      uint32_t *local_done, const uint32_t *timeout_flag, const FrameCounters *op_counters,
                                                      const FrameCounters *tr_counters, const uint32_t *rgba) {
    const bool bad = frame_overflowed(op_counters) || frame_overflowed(tr_counters);
    if (*timeout_flag == 0u && !bad) {  // after a timeout the buffer may still be in use: contribute nothing but still signal
        const int x = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
        const int y = y0 + blockIdx.y * 4 + (threadIdx.x >> 6);
        if (y < y1 && x < W) {
            const float4 *c = color + (size_t)y * Wp + x;
            uint32_t *o = pixels + (size_t)y * W + x;
            if (rgba) {  // fixed-exposure frame: already packed
                const uint32_t *src = rgba + (size_t)y * W + x;
                if (x + 3 < W && (W & 3) == 0) {
                    *reinterpret_cast<uint4 *>(o) = *reinterpret_cast<const uint4 *>(src);
                } else {
                    for (int k = 0; k < 4 && x + k < W; k++) o[k] = src[k];
                }
            } else if (x + 3 < W && (W & 3) == 0) {
                uint4 v;
                v.x = resolve_pixel_rgba(c[0], exposure);
                v.y = resolve_pixel_rgba(c[1], exposure);
                v.z = resolve_pixel_rgba(c[2], exposure);
                v.w = resolve_pixel_rgba(c[3], exposure);
                *reinterpret_cast<uint4 *>(o) = v;
            } else {
                for (int k = 0; k < 4 && x + k < W; k++) o[k] = resolve_pixel_rgba(c[k], exposure);
            }
        }
    }
    __threadfence_system();  // this thread's peer stores are performed before the CTA is counted
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t n = gridDim.x * gridDim.y;
        if (atomicAdd(local_done, 1u) == n - 1u) {
            *local_done = 0u;  // next frame
            __threadfence_system();
            if (!bad) atomicAdd_system(ctrl + SWR_PEER_DONE, 1u);
        }
    }
}

__global__ void k_peer_wait(const uint32_t *ctr, uint32_t target, uint32_t *timeout_flag) {
    if (!peer_wait(ctr, target)) *timeout_flag = 1u;
    __threadfence_system();
}

__global__ void k_peer_release(uint32_t *ctrl, uint32_t next_frame) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ctrl + SWR_PEER_FREE), "r"(next_frame) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Translucent pass (tilerasterizer.rs:92-101, shader.rs:66-76): per tile, packets sorted back to front, forward-shaded
// over the opaque colour, depth-tested against (never written to) the opaque depth.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_scan_simple(const uint32_t *count, uint32_t *offset, uint32_t *cursor, int ntiles, FrameCounters *counters,
                                                     uint32_t ref_capacity) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        int i = base + tid;
        uint32_t cb[SWR_ZBUCKETS];
        uint32_t v = i < ntiles ? load_tile_counts(count, i, cb) : 0u;
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
            offset[i] = excl;
            store_tile_cursors(cursor, i, excl, cb);
        }
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        offset[ntiles] = s_carry;
        counters->tile_refs = s_carry;
        if (s_carry > ref_capacity) counters->overflow_refs = 1;
    }
}

// Sort one tile's translucent refs by (avg_z descending, submission order ascending) — the reference's sort key is avg_z
// only and its quicksort is unstable (bumpqueue.rs:155-202); ties keep submission order here. Bitonic sort in shared memory.
#define TSORT_MAX 4096
__global__ void __launch_bounds__(256) k_sort_translucent(uint32_t *refs, const uint32_t *offset, const float *avgz, const uint32_t *clip_ext,
                                                          FrameCounters *counters) {
    __shared__ unsigned long long s_key[TSORT_MAX];
    if (counters->overflow_refs || counters->overflow_ext) return;
    const uint32_t beg = offset[blockIdx.x], n = offset[blockIdx.x + 1] - beg;
    if (n <= 1) return;
    if (n > TSORT_MAX) {
        if (threadIdx.x == 0) counters->overflow_sort = 1;
        return;
    }
    uint32_t m = 1;
    while (m < n) m <<= 1;
    for (uint32_t i = threadIdx.x; i < m; i += 256) {
        unsigned long long k = ~0ull;
        if (i < n) {
            const uint32_t id = refs[beg + i];
            const float az = avgz[record_of_id(id, clip_ext)];
            uint32_t b = __float_as_uint(az);
            uint32_t ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // OrderedFloat order (NaN above +inf)
            k = ((unsigned long long)(0xFFFFFFFFu - ord) << 32) | id;
        }
        s_key[i] = k;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < m; i += 256) {
                const uint32_t l = i ^ j;
                if (l > i) {
                    const unsigned long long a = s_key[i], b = s_key[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        s_key[i] = b;
                        s_key[l] = a;
                    }
                }
            }
            __syncthreads();
        }
    for (uint32_t i = threadIdx.x; i < n; i += 256) refs[beg + i] = (uint32_t)s_key[i];
}

struct ForwardParams {
    ShadeParams sh;            // records / clip_ext / draws / clip_verts of the TRANSLUCENT set; keys = opaque visibility keys
    const uint32_t *refs;      // translucent refs, sorted per tile
    const uint32_t *offset;
    const FrameCounters *counters;
};

__global__ void __launch_bounds__(SHADE_BLOCK, SWR_SHADE_MINB) k_forward_translucent(ForwardParams F) {
    const ShadeParams &P = F.sh;
    if (F.counters->overflow_refs || F.counters->overflow_ext || F.counters->overflow_sort) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int quad = lane >> 2, sub = lane & 3;
    const int bx0 = blockIdx.x * 16, by0 = P.row_begin * SWR_TILE + blockIdx.y * SHADE_ROWS;
    const int px = bx0 + quad * 2 + (sub & 1);
    const int py = by0 + warp * 2 + (sub >> 1);
    const unsigned qmask = 0xFu << (lane & ~3);
    const int tile = (by0 >> 6) * P.tiles_x + (bx0 >> 6);
    const uint32_t beg = F.offset[tile], end = F.offset[tile + 1];
    if (beg == end) return;
    // opaque depth of my pixel (never modified by this pass): inverse of depth_orderable; empty -> +inf
    const unsigned long long key = load_key(P.keys, P.tiles_x, px, py);
    float zopaque = __uint_as_float(SWR_INF_BITS);
    if (key != SWR_KEY_EMPTY) {
        const uint32_t u = (uint32_t)(key >> 32);
        zopaque = __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
    }
    const size_t ci = (size_t)py * P.Wp + px;
    float4 c4 = P.color[ci];
    V3 colour = v3(c4.x, c4.y, c4.z);
    const int tile_x0 = bx0 & ~(SWR_TILE - 1), tile_y0 = by0 & ~(SWR_TILE - 1);
    for (uint32_t i = beg; i < end; i++) {
        const uint32_t id = __ldg(F.refs + i);
        const TriRecord rec = P.records[record_of_id(id, P.clip_ext)];
        PacketSetup ps;
        packet_setup(rec, P.W, P.H, tile_x0, tile_y0, ps);
        if (ps.empty) continue;
        // block-uniform reject: does the packet's quad region touch this 16 x SHADE_ROWS block?
        const int rx0 = ps.xs >> 4, ry0 = ps.ys >> 4, rx1 = rx0 + ps.nqx * 2, ry1 = ry0 + ps.nqy * 2;
        if (rx1 <= bx0 || rx0 >= bx0 + 16 || ry1 <= by0 || ry0 >= by0 + SHADE_ROWS) continue;
        // my quad inside the region?  (quads are aligned to even pixels, so all four lanes agree)
        const bool inreg = px >= rx0 && px < rx1 && py >= ry0 && py < ry1;
        float w0 = -1.0f, w1 = -1.0f, w2 = -1.0f;
        bool evaluated = false;
        if (inreg) {
            if (ps.exact) {
                int e[3];
                eval_exact(ps, px * 16 + 8, py * 16 + 8, e);
                w0 = i2f(e[0]);
                w1 = i2f(e[1]);
                w2 = i2f(e[2]);
                evaluated = true;
            } else {
                float rr[3];
                evaluated = eval_chain(ps, ((px & ~1) * 16 - ps.xs) >> 5, ((py & ~1) * 16 - ps.ys) >> 5, px & 1, py & 1, rr);
                w0 = rr[0];
                w1 = rr[1];
                w2 = rr[2];
            }
        }
        const bool cov = evaluated && w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f;
        if (!(__ballot_sync(qmask, cov) & qmask)) continue;  // tilerasterizer.rs:334-335 mask.any() per quad
        // tilerasterizer.rs:337-342 for all four lanes of the quad
        float b1, b2;
        const float z = frag_depth(rec, w1, w2, b1, b2);
        const bool pass = cov && z <= zopaque;  // depth_test :511-523
        if (!(__ballot_sync(qmask, pass) & qmask)) continue;
        const float iwda = rec.iw1 - rec.iw0, iwdb = rec.iw2 - rec.iw0;
        const float w = 1.0f / interp1(rec.iw0, iwda, iwdb, b1, b2);  // shader.rs:123
        float wq[4];
#pragma unroll
        for (int j = 0; j < 4; j++) wq[j] = __shfl_sync(qmask, w, (lane & ~3) + j);
        if (pass) {
            ShadePacket sp;
            build_shade_packet(P, rec, sp);
            float dd[4] = {sp.du_dv[0] * wq[0], sp.du_dv[1] * wq[1], sp.du_dv[2] * wq[2], sp.du_dv[3] * wq[3]};
            colour = pbr_shader(P, sp, b1, b2, w, dd, true, colour);
        }
    }
    P.color[ci] = make_float4(colour.x, colour.y, colour.z, c4.w);
}

// ---------------------------------------------------------------------------------------------
// sort-last helpers: local ids <-> global seq in the key's low word
// ---------------------------------------------------------------------------------------------
// Before the cross-rank min: replace ~local_id by ~seq (seq is global and ordered like the serial schedule).
__global__ void k_keys_to_global(unsigned long long *keys, size_t n, const TriRecord *records, const uint32_t *clip_ext) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = keys[i];
    if (k == SWR_KEY_EMPTY) return;
    const uint32_t id = 0xFFFFFFFFu - (uint32_t)k;
    const uint32_t seq = records[record_of_id(id, clip_ext)].seq;
    keys[i] = (k & 0xFFFFFFFF00000000ull) | (unsigned long long)(0xFFFFFFFFu - seq);
}

// After the min: map seq back to a local id when the winner is one of this rank's draws (binary search over the
// draws' first_tri), else mark it foreign; write the winner's barycentrics (0 when not mine) for the bary exchange.
__global__ void k_keys_localize(unsigned long long *keys, int tiles_x, int W, int H, const TriRecord *records, const uint32_t *clip_ext,
                                const DevDraw *draws, const uint32_t *tri_prefix, uint32_t ndraws, const DevPrim *prims, float2 *bary) {
    int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= W || py >= H) return;
    const size_t ki = (size_t)((py >> 6) * tiles_x + (px >> 6)) * SWR_TILE_PIXELS + key_index(px & 63, py & 63);
    unsigned long long k = keys[ki];
    float2 b = make_float2(0.0f, 0.0f);
    if (k != SWR_KEY_EMPTY) {
        const uint32_t seq = 0xFFFFFFFFu - (uint32_t)k;
        const uint32_t G = seq >> 3, fan = seq & 7u;
        uint32_t id = SWR_ID_FOREIGN;
        if (ndraws > 0 && G >= draws[0].first_tri) {
            uint32_t lo = 0, hi = ndraws;  // draws[lo].first_tri <= G
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) >> 1;
                if (draws[mid].first_tri <= G)
                    lo = mid;
                else
                    hi = mid;
            }
            const uint32_t t = G - draws[lo].first_tri;
            if (t < prims[draws[lo].prim].ntris) id = (tri_prefix[lo] + t) * 8u + fan;
        }
        if (id != SWR_ID_FOREIGN) {
            TriRecord r = records[record_of_id(id, clip_ext)];
            float z;
            if (r.seq == seq) resolve_pixel(r, W, H, px, py, b.x, b.y, z);
        }
        keys[ki] = (k & 0xFFFFFFFF00000000ull) | (unsigned long long)(0xFFFFFFFFu - id);
    }
    bary[(size_t)py * W + px] = b;
}

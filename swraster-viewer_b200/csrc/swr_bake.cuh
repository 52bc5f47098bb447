// swr_bake.cuh — load-time bakes on the device: the inputs of the shading kernel the reference derives
// from its sky image and its geometry when a scene is loaded.
//   k_bake_brdf_lut      integrate_brdf / generate_brdf_lut            src/texture.rs:167-235 (128 GGX samples per texel)
//   k_bake_prefilter     generate_prefiltered_specular_cubemap         src/texture.rs:330-420 (every mip at full face
//                        resolution, one roughness per mip, `samples` importance samples x one bilinear sky tap each)
//   k_bake_sh4           compute_irradiance_sh4                        src/texture.rs:289-328 (projection of the whole sky)
//   k_sunvis_trace       compute_sun_visibility's ray cast             src/gi.rs:267-314, raytracer.rs:177-259
//   k_sunvis_blur        VoxelGrid::blur_grid + squaring               src/voxelgrid.rs:371-419
// Arithmetic follows the reference's f32 operation order (this translation unit is built with -fmad=false, IEEE sqrt / div);
// sinf / cosf / powf are CUDA's (<= 2 ulp) where the reference calls the C library's, so texels can differ from the CPU
// checker the tests use by one RGBA8 step here and there — the tests bound that.
// One thread per output texel / voxel; the sky (six faces, a few MB) is read through the read-only cache.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace bake {

struct B3 {
    float x, y, z;
};
__device__ __forceinline__ B3 operator+(B3 a, B3 b) { return B3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ B3 operator-(B3 a, B3 b) { return B3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ B3 operator*(B3 a, float s) { return B3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ B3 operator/(B3 a, float s) { return B3{a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ float dot(B3 a, B3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }  // glam Vec3A::dot
__device__ __forceinline__ B3 cross(B3 a, B3 b) { return B3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ B3 normalize(B3 a) { return a * (1.0f / sqrtf(dot(a, a))); }            // self * length_recip()
__device__ __forceinline__ float fmax_rs(float a, float b) { return fmaxf(a, b); }                 // f32::max
__device__ __forceinline__ float clamp_rs(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ uint32_t pack(float r, float g, float b, float a) {  // util.rs:91-96: `as u32` saturates, NaN -> 0
    return (__float2uint_rz(r * 255.0f) << 24) | (__float2uint_rz(g * 255.0f) << 16) | (__float2uint_rz(b * 255.0f) << 8) | __float2uint_rz(a * 255.0f);
}
__device__ __forceinline__ float srgb_to_linear_scalar(float s) { return s <= 0.04045f ? s / 12.92f : powf((s + 0.055f) / 1.055f, 2.4f); }  // util.rs:50-56

__device__ __forceinline__ float radical_inverse_vdc(uint32_t bits) { return (float)__brev(bits) * 2.3283064e-10f; }  // texture.rs:135-142

__device__ __forceinline__ B3 importance_sample_ggx(float xi_x, float xi_y, B3 n, float roughness) {  // texture.rs:148-165
    const float a = roughness * roughness;
    const float phi = 2.0f * 3.14159274f * xi_x;
    const float cos_theta = sqrtf((1.0f - xi_y) / (1.0f + (a * a - 1.0f) * xi_y));
    const float sin_theta = sqrtf(fmax_rs(1.0f - cos_theta * cos_theta, 0.0f));
    const B3 h{cosf(phi) * sin_theta, sinf(phi) * sin_theta, cos_theta};
    const B3 up = fabsf(n.z) < 0.999f ? B3{0.0f, 0.0f, 1.0f} : B3{1.0f, 0.0f, 0.0f};
    const B3 tangent = normalize(cross(n, up));
    const B3 bitangent = cross(n, tangent);
    return normalize((tangent * h.x + bitangent * h.y) + n * h.z);
}

__device__ __forceinline__ B3 face_uv_to_direction(uint32_t face, float u, float v) {  // texture.rs:237-247
    B3 d;
    switch (face) {
        case 0: d = B3{1.0f, -v, -u}; break;
        case 1: d = B3{-1.0f, -v, u}; break;
        case 2: d = B3{u, 1.0f, v}; break;
        case 3: d = B3{u, -1.0f, -v}; break;
        case 4: d = B3{u, -v, 1.0f}; break;
        default: d = B3{-u, -v, -1.0f}; break;
    }
    return normalize(d);
}

struct Cube {  // mip 0 of the sky: six faces of w x h RGBA8 texels, face-major
    const uint32_t *texels;
    uint32_t w, h;
};

// cubemap_direction_to_face_uv (texture.rs:249-272) + sample_bilinear_rgb at mip 0 with ClampToEdge (:730-790) + sRGB -> linear
__device__ B3 sample_direction_linear(const Cube &c, B3 n) {
    const float ax = fabsf(n.x), ay = fabsf(n.y), az = fabsf(n.z);
    uint32_t face;
    float u, v;
    if (ax >= ay && ax >= az) {
        if (n.x >= 0.0f)
            face = 0, u = (-n.z / ax) * 0.5f + 0.5f, v = (-n.y / ax) * 0.5f + 0.5f;
        else
            face = 1, u = (n.z / ax) * 0.5f + 0.5f, v = (-n.y / ax) * 0.5f + 0.5f;
    } else if (ay > ax && ay >= az) {
        if (n.y >= 0.0f)
            face = 2, u = (n.x / ay) * 0.5f + 0.5f, v = (n.z / ay) * 0.5f + 0.5f;
        else
            face = 3, u = (n.x / ay) * 0.5f + 0.5f, v = (-n.z / ay) * 0.5f + 0.5f;
    } else if (n.z >= 0.0f) {
        face = 4, u = (n.x / az) * 0.5f + 0.5f, v = (-n.y / az) * 0.5f + 0.5f;
    } else {
        face = 5, u = (-n.x / az) * 0.5f + 0.5f, v = (-n.y / az) * 0.5f + 0.5f;
    }
    u = clamp_rs(u, 0.0f, 1.0f), v = clamp_rs(v, 0.0f, 1.0f);
    const float wf = (float)c.w, hf = (float)c.h;
    const uint32_t off = face * c.w * c.h;
    const float xf = u * wf - 0.5f, yf = v * hf - 0.5f;
    const float x0 = floorf(xf), y0 = floorf(yf), x1 = x0 + 1.0f, y1 = y0 + 1.0f;
    const float fx = xf - x0, fy = yf - y0, ofx = 1.0f - fx, ofy = 1.0f - fy;
    // _mm_min_ps(texel, dim - 1) then as_uvec4 (negative -> 0)
    const uint32_t x0i = __float2uint_rz(fminf(x0, wf - 1.0f)), y0i = __float2uint_rz(fminf(y0, hf - 1.0f));
    const uint32_t x1i = __float2uint_rz(fminf(x1, wf - 1.0f)), y1i = __float2uint_rz(fminf(y1, hf - 1.0f));
    const uint32_t t00 = __ldg(c.texels + off + y0i * c.w + x0i), t10 = __ldg(c.texels + off + y0i * c.w + x1i);
    const uint32_t t01 = __ldg(c.texels + off + y1i * c.w + x0i), t11 = __ldg(c.texels + off + y1i * c.w + x1i);
    const float w00 = ofx * ofy, w10 = fx * ofy, w01 = ofx * fy, w11 = fx * fy;
    float rgb[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const int sh = 24 - 8 * ch;
        const float p00 = (float)((t00 >> sh) & 0xFF) / 255.0f, p10 = (float)((t10 >> sh) & 0xFF) / 255.0f;
        const float p01 = (float)((t01 >> sh) & 0xFF) / 255.0f, p11 = (float)((t11 >> sh) & 0xFF) / 255.0f;
        rgb[ch] = ((p00 * w00 + p10 * w10) + p01 * w01) + p11 * w11;
    }
    return B3{srgb_to_linear_scalar(rgb[0]), srgb_to_linear_scalar(rgb[1]), srgb_to_linear_scalar(rgb[2])};
}

// ---- BRDF LUT: one thread per texel -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_bake_brdf_lut(uint32_t size, uint32_t *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size * size) return;
    const uint32_t x = i % size, y = i / size;
    const float size_f = (float)size;
    const float roughness = fmax_rs(clamp_rs(((float)y + 0.5f) / size_f, 0.0f, 1.0f), 1.0e-4f);
    const float ndotv = fmax_rs(clamp_rs(((float)x + 0.5f) / size_f, 0.0f, 1.0f), 1.0e-4f);
    // integrate_brdf (texture.rs:167-197)
    const B3 v{sqrtf(fmax_rs(1.0f - ndotv * ndotv, 0.0f)), 0.0f, ndotv};
    const B3 n{0.0f, 0.0f, 1.0f};
    float a = 0.0f, b = 0.0f;
    for (uint32_t s = 0; s < 128u; s++) {
        const B3 h = importance_sample_ggx((float)s / 128.0f, radical_inverse_vdc(s), n, roughness);
        const B3 l = normalize(h * (2.0f * dot(v, h)) - v);
        const float ndotl = fmax_rs(l.z, 0.0f), ndoth = fmax_rs(h.z, 0.0f), vdoth = fmax_rs(dot(v, h), 0.0f);
        if (ndotl > 0.0f) {
            const float alpha = roughness * roughness;
            const float k = (alpha + 1.0f) * (alpha + 1.0f) * 0.125f;
            const float g_v = ndotv / (ndotv * (1.0f - k) + k);
            const float g_l = ndotl / (ndotl * (1.0f - k) + k);
            const float g_vis = fmax_rs(g_v * g_l * vdoth / (ndoth * fmax_rs(ndotv, 1.0e-5f)), 0.0f);
            const float om = 1.0f - vdoth, om2 = om * om;
            const float fc = om * (om2 * om2);  // f32::powi(5)
            a += (1.0f - fc) * g_vis;
            b += fc * g_vis;
        }
    }
    out[i] = pack(clamp_rs(a / 128.0f, 0.0f, 1.0f), clamp_rs(b / 128.0f, 0.0f, 1.0f), 0.0f, 1.0f);
}

// ---- GGX prefilter: one thread per texel of one mip (face, y, x) ------------------------------------------------------------
__global__ void __launch_bounds__(128) k_bake_prefilter(Cube cube, uint32_t mip, uint32_t max_mip, uint32_t sample_count, uint32_t *out_mip) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = cube.w, h = cube.h;
    if (i >= 6u * w * h) return;
    const uint32_t face = i / (w * h), y = (i / w) % h, x = i % w;
    const float u = (((float)x + 0.5f) / (float)w) * 2.0f - 1.0f;
    const float v = (((float)y + 0.5f) / (float)h) * 2.0f - 1.0f;
    const B3 r = face_uv_to_direction(face, u, v);
    B3 color;
    if (mip == 0) {
        color = sample_direction_linear(cube, r);
    } else {
        const float roughness = fmax_rs((float)mip / (float)max_mip, 0.045f);
        B3 accum{0.0f, 0.0f, 0.0f};
        float total = 0.0f;
        for (uint32_t s = 0; s < sample_count; s++) {
            const B3 hv = importance_sample_ggx((float)s / (float)sample_count, radical_inverse_vdc(s), r, roughness);
            const B3 l = normalize(hv * (2.0f * dot(r, hv)) - r);
            const float ndotl = fmax_rs(dot(r, l), 0.0f);
            if (ndotl > 0.0f) {
                accum = accum + sample_direction_linear(cube, l) * ndotl;
                total += ndotl;
            }
        }
        color = total > 0.0f ? accum / total : sample_direction_linear(cube, r);
    }
    out_mip[i] = pack(color.x, color.y, color.z, 1.0f);
}

// ---- irradiance SH4: per-block partial sums in a fixed order, then one thread adds the blocks in order (deterministic) -------
#define SH_BLOCK 256
__global__ void __launch_bounds__(SH_BLOCK) k_bake_sh4_partial(Cube cube, float *partial /* gridDim.x x 12 */) {
    __shared__ float s[SH_BLOCK][13];
    const uint32_t w = cube.w, h = cube.h, n = 6u * w * h;
    const uint32_t i = blockIdx.x * SH_BLOCK + threadIdx.x;
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; k++) acc[k] = 0.0f;
    if (i < n) {
        const uint32_t face = i / (w * h), y = (i / w) % h, x = i % w;
        const float width = (float)w, height = (float)h;
        const float texel_omega = (2.0f / width) * (2.0f / height);
        const float v = (((float)y + 0.5f) / height) * 2.0f - 1.0f;
        const float u = (((float)x + 0.5f) / width) * 2.0f - 1.0f;
        const B3 dir = face_uv_to_direction(face, u, v);
        const float weight = texel_omega / powf((1.0f + u * u) + v * v, 1.5f);
        const B3 color = sample_direction_linear(cube, dir);
        const float basis[4] = {0.282095f, 0.488603f * dir.y, 0.488603f * dir.z, 0.488603f * dir.x};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float bw = basis[k] * weight;
            acc[3 * k] = color.x * bw, acc[3 * k + 1] = color.y * bw, acc[3 * k + 2] = color.z * bw;
        }
    }
#pragma unroll
    for (int k = 0; k < 12; k++) s[threadIdx.x][k] = acc[k];
    __syncthreads();
    for (int stride = SH_BLOCK / 2; stride > 0; stride >>= 1) {  // fixed tree
        if ((int)threadIdx.x < stride)
#pragma unroll
            for (int k = 0; k < 12; k++) s[threadIdx.x][k] += s[threadIdx.x + stride][k];
        __syncthreads();
    }
    if (threadIdx.x < 12) partial[blockIdx.x * 12 + threadIdx.x] = s[0][threadIdx.x];
}
__global__ void k_bake_sh4_final(const float *partial, uint32_t nblocks, float *out12) {
    const uint32_t k = threadIdx.x;
    if (k >= 12) return;
    float a = 0.0f;
    for (uint32_t b = 0; b < nblocks; b++) a += partial[b * 12 + k];
    // texture.rs:318-326: cosine-lobe convolution, then the basis constants folded in
    const bool l0 = k < 3;
    a = a * (l0 ? 3.14159274f : (2.0f * 3.14159274f / 3.0f));
    out12[k] = a * (l0 ? 0.282095f : 0.488603f);
}

// ---- voxel sun visibility: one ray per active voxel through a BVH built on the host ------------------------------------------
struct SunNode {  // median-split hierarchy: leaf = [first, first + count) in `order`; inner = children first, first + 1 (count = 0)
    float lo[3], hi[3];
    uint32_t first, count;
};
struct SunTri {
    float p0[3], p1[3], p2[3];
    float transmission;  // < 0: opaque material
};
struct SunParams {
    const SunNode *nodes;
    const uint32_t *order;
    const SunTri *tris;
    uint32_t nnodes;
    const uint8_t *active;
    uint32_t W, H, D;
    float center_min[3], vs[3], L[3], bias;
    float *out;  // per voxel: 1 where inactive, transmittance where active
    uint32_t *overflow;
};
#define SUN_MAX_TRANSLUCENT 24

__device__ __forceinline__ bool ray_triangle(B3 o, B3 d, const SunTri &t, float t_min, float &t_out) {  // raytracer.rs:223-259
    const B3 p0{t.p0[0], t.p0[1], t.p0[2]}, p1{t.p1[0], t.p1[1], t.p1[2]}, p2{t.p2[0], t.p2[1], t.p2[2]};
    const B3 edge1 = p1 - p0, edge2 = p2 - p0;
    const B3 pvec = cross(d, edge2);
    const float det = dot(edge1, pvec);
    if (fabsf(det) <= 1.0e-8f) return false;
    const float inv_det = 1.0f / det;
    const B3 tvec = o - p0;
    const float u = dot(tvec, pvec) * inv_det;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    const B3 qvec = cross(tvec, edge1);
    const float v = dot(d, qvec) * inv_det;
    if (v < 0.0f || (u + v) > 1.0f) return false;
    const float tt = dot(edge2, qvec) * inv_det;
    if (tt < t_min) return false;  // t_max = +inf
    t_out = tt;
    return true;
}

__global__ void __launch_bounds__(128) k_sunvis_trace(SunParams P) {
    const size_t total = (size_t)P.W * P.H * P.D;
    const size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (index >= total) return;
    if (!P.active[index]) {
        P.out[index] = 1.0f;
        return;
    }
    const uint32_t z = (uint32_t)(index / ((size_t)P.W * P.H)), rem = (uint32_t)(index % ((size_t)P.W * P.H)), y = rem / P.W, x = rem % P.W;
    const B3 L{P.L[0], P.L[1], P.L[2]};
    const B3 c = B3{P.center_min[0], P.center_min[1], P.center_min[2]} + B3{(float)x * P.vs[0], (float)y * P.vs[1], (float)z * P.vs[2]};
    const B3 o = c + L * P.bias;
    const float inv[3] = {1.0f / L.x, 1.0f / L.y, 1.0f / L.z}, oo[3] = {o.x, o.y, o.z}, dd[3] = {L.x, L.y, L.z};
    // trace_transmittance (raytracer.rs:177-211): an opaque hit anywhere on the ray gives 0 whatever the order; translucent hits
    // multiply in order of distance (ties by triangle index), kept in a small sorted list
    float ht[SUN_MAX_TRANSLUCENT];
    uint32_t hi_[SUN_MAX_TRANSLUCENT];
    int nh = 0;
    bool opaque = false;
    uint32_t stack[64];
    int sp = 0;
    stack[sp++] = 0;
    while (sp && !opaque) {
        const SunNode nd = P.nodes[stack[--sp]];
        float t0 = 0.0f, t1 = INFINITY;
        bool miss = false;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (miss) continue;
            if (dd[a] == 0.0f) {
                miss = oo[a] < nd.lo[a] || oo[a] > nd.hi[a];
            } else {
                float ta = (nd.lo[a] - oo[a]) * inv[a], tb = (nd.hi[a] - oo[a]) * inv[a];
                if (ta > tb) {
                    const float tmp = ta;
                    ta = tb;
                    tb = tmp;
                }
                ta -= fabsf(ta) * 4.0e-7f;  // widened by a few ulps: rounding can only add candidates
                tb += fabsf(tb) * 4.0e-7f;
                t0 = fmaxf(t0, ta);
                t1 = fminf(t1, tb);
                miss = t0 > t1;
            }
        }
        if (miss) continue;
        if (nd.count) {
            for (uint32_t k = 0; k < nd.count && !opaque; k++) {
                const uint32_t ti = P.order[nd.first + k];
                const SunTri tri = P.tris[ti];
                float t;
                if (!ray_triangle(o, L, tri, 1.0e-4f, t)) continue;
                if (tri.transmission < 0.0f) {
                    opaque = true;
                } else if (nh < SUN_MAX_TRANSLUCENT) {
                    int j = nh++;
                    while (j > 0 && (ht[j - 1] > t || (ht[j - 1] == t && hi_[j - 1] > ti))) {
                        ht[j] = ht[j - 1];
                        hi_[j] = hi_[j - 1];
                        j--;
                    }
                    ht[j] = t;
                    hi_[j] = ti;
                } else {
                    atomicAdd(P.overflow, 1u);
                }
            }
        } else if (sp + 2 <= 64) {
            stack[sp++] = nd.first;
            stack[sp++] = nd.first + 1;
        } else {
            atomicAdd(P.overflow, 1u);
        }
    }
    float tr = 1.0f;
    if (opaque) {
        tr = 0.0f;
    } else {
        for (int j = 0; j < nh; j++) {
            tr *= P.tris[hi_[j]].transmission;
            if (tr <= 0.0001f) {
                tr = 0.0f;
                break;
            }
        }
    }
    P.out[index] = tr;
}

// blur_grid: 3x3x3 mean of what lies inside the grid, squared (voxelgrid.rs:371-419; powf(x, 2) == x * x for every finite x)
__global__ void __launch_bounds__(128) k_sunvis_blur(const float *in, float *out, uint32_t W, uint32_t H, uint32_t D) {
    const size_t total = (size_t)W * H * D;
    const size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (index >= total) return;
    const int z = (int)(index / ((size_t)W * H)), rem = (int)(index % ((size_t)W * H)), y = rem / (int)W, x = rem % (int)W;
    float sum = 0.0f;
    int count = 0;
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                const int nx = x + dx, ny = y + dy, nz = z + dz;
                if (nx < 0 || ny < 0 || nz < 0 || nx >= (int)W || ny >= (int)H || nz >= (int)D) continue;
                sum += in[((size_t)nz * H + (size_t)ny) * W + (size_t)nx];
                count++;
            }
    const float m = sum / (float)count;
    out[index] = m * m;
}

}  // namespace bake

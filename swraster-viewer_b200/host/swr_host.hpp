// swr_host.hpp — host-side mirror of the reference's Renderer API, written above
// the C ABI in include/swr.h (no CUDA here, no torch).
//
// The reference is Rust and no Rust toolchain exists in this image, so the host
// side is C++ and mirrors the reference's names, argument meaning and error
// behaviour 1:1:
//   swr::RenderCamera  <- src/rendercamera.rs:5-150
//   swr::RenderBuffer  <- src/renderer.rs:24-28,78-96
//   swr::Renderer      <- src/renderer.rs:145-355 (new / render_scene /
//                         update_auto_exposure / blit_to_buffer)
//   swr::Scene         <- src/scene.rs:65-77 (borrowed, immutable: a view over
//                         swr_scene_desc)
// What stays on the host (SURVEY §8b "who computes what"): node ordering
// (renderer.rs:357-367), per-primitive sphere/frustum classification and the
// mvp product (renderer.rs:369-468), auto-exposure metering (renderer.rs:258-290).
// Everything per-triangle and per-pixel is behind swr_render / swr_resolve.
// Errors: the reference panics on this path; this mirror throws
// std::runtime_error carrying swr_last_error().
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#endif
#include "../../include/swr.h"

namespace swr {

// ---- small glam-shaped helpers (column-major) -------------------------------------
struct Mat4 {
    float m[16];
    static Mat4 identity() {
        Mat4 r{};
        r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
        return r;
    }
};
inline void mul_vec4(const float *m, const float *v, float *out) {  // glam sse2 Mat4::mul_vec4 association
    for (int r = 0; r < 4; r++) out[r] = ((m[r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2]) + m[12 + r] * v[3];
}
inline Mat4 mul(const Mat4 &a, const Mat4 &b) {  // Mat4 * Mat4: columns of b through mul_vec4
    Mat4 r;
    for (int c = 0; c < 4; c++) mul_vec4(a.m, b.m + 4 * c, r.m + 4 * c);
    return r;
}
inline float dot4(const float *a, const float *b) { return (a[0] * b[0] + a[2] * b[2]) + (a[1] * b[1] + a[3] * b[3]); }
inline Mat4 transpose(const Mat4 &a) {
    Mat4 r;
    for (int c = 0; c < 4; c++)
        for (int k = 0; k < 4; k++) r.m[c * 4 + k] = a.m[k * 4 + c];
    return r;
}
inline Mat4 inverse(const Mat4 &a) {  // general 4x4 inverse (cofactors, double accumulate)
    const float *m = a.m;
    double inv[16];
    inv[0] = (double)m[5] * m[10] * m[15] - (double)m[5] * m[11] * m[14] - (double)m[9] * m[6] * m[15] + (double)m[9] * m[7] * m[14] + (double)m[13] * m[6] * m[11] - (double)m[13] * m[7] * m[10];
    inv[4] = -(double)m[4] * m[10] * m[15] + (double)m[4] * m[11] * m[14] + (double)m[8] * m[6] * m[15] - (double)m[8] * m[7] * m[14] - (double)m[12] * m[6] * m[11] + (double)m[12] * m[7] * m[10];
    inv[8] = (double)m[4] * m[9] * m[15] - (double)m[4] * m[11] * m[13] - (double)m[8] * m[5] * m[15] + (double)m[8] * m[7] * m[13] + (double)m[12] * m[5] * m[11] - (double)m[12] * m[7] * m[9];
    inv[12] = -(double)m[4] * m[9] * m[14] + (double)m[4] * m[10] * m[13] + (double)m[8] * m[5] * m[14] - (double)m[8] * m[6] * m[13] - (double)m[12] * m[5] * m[10] + (double)m[12] * m[6] * m[9];
    inv[1] = -(double)m[1] * m[10] * m[15] + (double)m[1] * m[11] * m[14] + (double)m[9] * m[2] * m[15] - (double)m[9] * m[3] * m[14] - (double)m[13] * m[2] * m[11] + (double)m[13] * m[3] * m[10];
    inv[5] = (double)m[0] * m[10] * m[15] - (double)m[0] * m[11] * m[14] - (double)m[8] * m[2] * m[15] + (double)m[8] * m[3] * m[14] + (double)m[12] * m[2] * m[11] - (double)m[12] * m[3] * m[10];
    inv[9] = -(double)m[0] * m[9] * m[15] + (double)m[0] * m[11] * m[13] + (double)m[8] * m[1] * m[15] - (double)m[8] * m[3] * m[13] - (double)m[12] * m[1] * m[11] + (double)m[12] * m[3] * m[9];
    inv[13] = (double)m[0] * m[9] * m[14] - (double)m[0] * m[10] * m[13] - (double)m[8] * m[1] * m[14] + (double)m[8] * m[2] * m[13] + (double)m[12] * m[1] * m[10] - (double)m[12] * m[2] * m[9];
    inv[2] = (double)m[1] * m[6] * m[15] - (double)m[1] * m[7] * m[14] - (double)m[5] * m[2] * m[15] + (double)m[5] * m[3] * m[14] + (double)m[13] * m[2] * m[7] - (double)m[13] * m[3] * m[6];
    inv[6] = -(double)m[0] * m[6] * m[15] + (double)m[0] * m[7] * m[14] + (double)m[4] * m[2] * m[15] - (double)m[4] * m[3] * m[14] - (double)m[12] * m[2] * m[7] + (double)m[12] * m[3] * m[6];
    inv[10] = (double)m[0] * m[5] * m[15] - (double)m[0] * m[7] * m[13] - (double)m[4] * m[1] * m[15] + (double)m[4] * m[3] * m[13] + (double)m[12] * m[1] * m[7] - (double)m[12] * m[3] * m[5];
    inv[14] = -(double)m[0] * m[5] * m[14] + (double)m[0] * m[6] * m[13] + (double)m[4] * m[1] * m[14] - (double)m[4] * m[2] * m[13] - (double)m[12] * m[1] * m[6] + (double)m[12] * m[2] * m[5];
    inv[3] = -(double)m[1] * m[6] * m[11] + (double)m[1] * m[7] * m[10] + (double)m[5] * m[2] * m[11] - (double)m[5] * m[3] * m[10] - (double)m[9] * m[2] * m[7] + (double)m[9] * m[3] * m[6];
    inv[7] = (double)m[0] * m[6] * m[11] - (double)m[0] * m[7] * m[10] - (double)m[4] * m[2] * m[11] + (double)m[4] * m[3] * m[10] + (double)m[8] * m[2] * m[7] - (double)m[8] * m[3] * m[6];
    inv[11] = -(double)m[0] * m[5] * m[11] + (double)m[0] * m[7] * m[9] + (double)m[4] * m[1] * m[11] - (double)m[4] * m[3] * m[9] - (double)m[8] * m[1] * m[7] + (double)m[8] * m[3] * m[5];
    inv[15] = (double)m[0] * m[5] * m[10] - (double)m[0] * m[6] * m[9] - (double)m[4] * m[1] * m[10] + (double)m[4] * m[2] * m[9] + (double)m[8] * m[1] * m[6] - (double)m[8] * m[2] * m[5];
    double det = (double)m[0] * inv[0] + (double)m[1] * inv[4] + (double)m[2] * inv[8] + (double)m[3] * inv[12];
    Mat4 r;
    double id = 1.0 / det;
    for (int i = 0; i < 16; i++) r.m[i] = (float)(inv[i] * id);
    return r;
}
struct Quat {
    float x, y, z, w;
};
inline Quat qmul(Quat a, Quat b) {
    return Quat{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
                a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline Mat4 mat4_from_quat(Quat q) {
    float x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
    float xx = q.x * x2, xy = q.x * y2, xz = q.x * z2, yy = q.y * y2, yz = q.y * z2, zz = q.z * z2;
    float wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
    Mat4 r = Mat4::identity();
    r.m[0] = 1.0f - (yy + zz);
    r.m[1] = xy + wz;
    r.m[2] = xz - wy;
    r.m[4] = xy - wz;
    r.m[5] = 1.0f - (xx + zz);
    r.m[6] = yz + wx;
    r.m[8] = xz + wy;
    r.m[9] = yz - wx;
    r.m[10] = 1.0f - (xx + yy);
    return r;
}

// ---- rendercamera.rs ----------------------------------------------------------------
struct RenderCamera {
    float position[3];
    float fov, width, height, one_over_width, one_over_height, near_, far_, yaw, pitch, roll;
    Mat4 view_matrix, projection_matrix, view_project_matrix, inverse_view_project_matrix, skybox_matrix_transposed;
    float view_clip_planes[6][4];

    // RenderCamera::new (rendercamera.rs:28-63)
    RenderCamera(const float pos[3], const float look_at[3], float fov_, float w, float h, float far_plane) {
        float f[3] = {look_at[0] - pos[0], look_at[1] - pos[1], look_at[2] - pos[2]};
        float len = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        for (float &c : f) c /= len;
        std::memcpy(position, pos, sizeof(position));
        yaw = std::atan2(f[0], -f[2]);
        pitch = std::asin(f[1]);
        roll = 0.0f;
        fov = fov_;
        width = w;
        height = h;
        one_over_width = 1.0f / w;
        one_over_height = 1.0f / h;
        near_ = far_plane * 0.01f;
        far_ = far_plane;
        update_matrices();
    }
    // rotate_mouse (rendercamera.rs:119-123): the reference's way of changing yaw/pitch after construction.
    // Note RenderCamera::new derives pitch = asin(forward.y) while this path treats positive pitch as looking
    // DOWN, so `new` only aims at look_at when the view is level (true for the reference's own call sites,
    // main.rs:210-224); both behaviours are kept as they are.
    void rotate_mouse(float dx, float dy) {
        const float PI = 3.14159265358979323846f;
        yaw += dx * 0.01f;
        pitch = std::fmin(std::fmax(pitch + dy * 0.01f, -PI / 2.0f + 0.1f), PI / 2.0f - 0.1f);
    }
    Quat get_rotation() const {  // :125-135 roll * pitch * yaw
        Quat qy{0.0f, std::sin(yaw * 0.5f), 0.0f, std::cos(yaw * 0.5f)};
        Quat qp{std::sin(pitch * 0.5f), 0.0f, 0.0f, std::cos(pitch * 0.5f)};
        Quat qr{0.0f, 0.0f, std::sin(roll * 0.5f), std::cos(roll * 0.5f)};
        return qmul(qmul(qr, qp), qy);
    }
    Mat4 compute_projection_matrix() const {  // glam Mat4::perspective_rh (depth 0..1)
        float s = std::sin(0.5f * fov), c = std::cos(0.5f * fov);
        float h = c / s, w = h / (width / height), r = far_ / (near_ - far_);
        Mat4 p{};
        p.m[0] = w;
        p.m[5] = h;
        p.m[10] = r;
        p.m[11] = -1.0f;
        p.m[14] = r * near_;
        return p;
    }
    void update_matrices() {  // :78-86
        Mat4 rot = mat4_from_quat(get_rotation());
        Mat4 tr = Mat4::identity();
        tr.m[12] = -position[0];
        tr.m[13] = -position[1];
        tr.m[14] = -position[2];
        view_matrix = mul(rot, tr);
        projection_matrix = compute_projection_matrix();
        view_project_matrix = mul(projection_matrix, view_matrix);
        inverse_view_project_matrix = inverse(view_project_matrix);
        skybox_matrix_transposed = transpose(inverse(mul(projection_matrix, rot)));  // :102-107
        // :137-150
        float aspect = width / height;
        float tan_y = std::tan(fov * 0.5f), tan_x = tan_y * aspect;
        auto set = [&](int i, float x, float y, float z, float w) {
            view_clip_planes[i][0] = x;
            view_clip_planes[i][1] = y;
            view_clip_planes[i][2] = z;
            view_clip_planes[i][3] = w;
        };
        auto setn = [&](int i, float x, float y, float z) {
            float l = std::sqrt(x * x + y * y + z * z);
            set(i, x / l, y / l, z / l, 0.0f);
        };
        set(0, 0.0f, 0.0f, -1.0f, near_);
        set(1, 0.0f, 0.0f, 1.0f, far_);
        setn(2, 1.0f, 0.0f, -tan_x);
        setn(3, -1.0f, 0.0f, -tan_x);
        setn(4, 0.0f, 1.0f, -tan_y);
        setn(5, 0.0f, -1.0f, -tan_y);
    }
    swr_camera to_abi() const {
        swr_camera c{};
        c.position[0] = position[0];
        c.position[1] = position[1];
        c.position[2] = position[2];
        std::memcpy(c.view_matrix, view_matrix.m, 64);
        std::memcpy(c.view_project_matrix, view_project_matrix.m, 64);
        std::memcpy(c.skybox_matrix_transposed, skybox_matrix_transposed.m, 64);
        std::memcpy(c.view_clip_planes, view_clip_planes, sizeof(view_clip_planes));
        c.one_over_width = one_over_width;
        c.one_over_height = one_over_height;
        return c;
    }
};

// ---- scene.rs:65-77 (borrowed view) ---------------------------------------------------
struct Scene {
    const swr_scene_desc *desc;
    explicit Scene(const swr_scene_desc *d) : desc(d) {}
};

// ---- renderer.rs:24-28,78-96 ----------------------------------------------------------
struct RenderBuffer {
    size_t width, height;
    uint32_t *pixels;
    RenderBuffer(size_t w, size_t h, uint32_t *p) : width(w), height(h), pixels(p) {}
    void clear() { std::fill(pixels, pixels + width * height, 0u); }
    void set_pixel(size_t x, size_t y, uint32_t color) {
        if (x < width && y < height) pixels[y * width + x] = color;
    }
};

enum class FrustumTestResult { Inside, Outside, Intersecting };

// scene.rs:53-63 (Mat4 * &BoundingSphere) + renderer.rs:130-142
inline FrustumTestResult test_sphere_frustum(const float *model, const float *sphere, const swr_camera &cam) {
    float lx = std::sqrt(dot4(model + 0, model + 0)), ly = std::sqrt(dot4(model + 4, model + 4)),
          lz = std::sqrt(dot4(model + 8, model + 8));
    float max_scale = (lx + ly + lz) / 3.0f;
    float c[4];
    for (int r = 0; r < 3; r++) c[r] = ((model[r] * sphere[0] + model[4 + r] * sphere[1]) + model[8 + r] * sphere[2]) + model[12 + r];
    c[3] = 1.0f;
    float radius = sphere[3] * max_scale;
    float cv[4];
    mul_vec4(cam.view_matrix, c, cv);
    FrustumTestResult result = FrustumTestResult::Inside;
    for (int p = 0; p < 6; p++) {
        float d = dot4(cam.view_clip_planes[p], cv);
        if (d < -radius)
            return FrustumTestResult::Outside;
        else if (d < radius)
            result = FrustumTestResult::Intersecting;
    }
    return result;
}

// Host half of render_scene: renderer.rs:357-367 (node order) and :369-468 (classification, mvp).
// `shard`/`nshards` select every draw whose index % nshards == shard (sort-last); ids stay global.
// Sort-first helper (not in the reference): true when the world-space sphere cannot touch the pixel rows [y0, y1) of a
// `height`-row image. Conservative: the band is widened by 2 pixels (snapping, exclusive bbox) and the radius uses the
// LARGEST axis scale of the model matrix. A draw culled here has no triangle whose bounding box reaches the band, so
// dropping it from this rank's draw list changes nothing in the band (ids stay global through first_triangle).
inline bool sphere_outside_row_band(const float *model, const float *sphere, const swr_camera &cam, int y0, int y1, int height) {
    float lx = std::sqrt(dot4(model + 0, model + 0)), ly = std::sqrt(dot4(model + 4, model + 4)), lz = std::sqrt(dot4(model + 8, model + 8));
    float radius = sphere[3] * std::fmax(lx, std::fmax(ly, lz)) * 1.0001f;
    float c[4], cv[4];
    for (int r = 0; r < 3; r++) c[r] = ((model[r] * sphere[0] + model[4 + r] * sphere[1]) + model[8 + r] * sphere[2]) + model[12 + r];
    c[3] = 1.0f;
    mul_vec4(cam.view_matrix, c, cv);
    const float *top = cam.view_clip_planes[4];  // normalize(0, 1, -tan_y): rendercamera.rs:146
    if (!(top[1] > 0.0f)) return false;
    float h = -top[1] / top[2];                  // 1 / tan(fov/2) == projection[1][1]
    if (!(h > 0.0f) || !std::isfinite(h)) return false;
    float nhi = 1.0f - 2.0f * (float)(y0 - 2) / (float)height;  // ndc y of the band's upper edge
    float nlo = 1.0f - 2.0f * (float)(y1 + 2) / (float)height;
    // inside the band (in front of the camera, z < 0):  nlo <= h*y / (-z) <= nhi
    float du = (h * cv[1] + nhi * cv[2]) / std::sqrt(h * h + nhi * nhi);  // > 0: above the upper plane
    float dl = (h * cv[1] + nlo * cv[2]) / std::sqrt(h * h + nlo * nlo);  // < 0: below the lower plane
    return du > radius || dl < -radius;
}

// The part of the scene description only the host reads (node -> mesh -> primitive ranges, material indices): the
// reference would panic on a slice access, this mirror throws before anything is indexed.
inline void validate_scene_ranges(const swr_scene_desc &sc) {
    for (uint32_t i = 0; i < sc.nmeshes; i++)
        if ((uint64_t)sc.meshes[i].first_primitive + sc.meshes[i].num_primitives > sc.nprimitives)
            throw std::runtime_error("scene: mesh " + std::to_string(i) + " names primitives beyond the primitive array");
    for (uint32_t i = 0; i < sc.nnodes; i++)
        if (sc.nodes[i].mesh_index >= (int32_t)sc.nmeshes) throw std::runtime_error("scene: node " + std::to_string(i) + " names a mesh that does not exist");
    for (uint32_t i = 0; i < sc.nprimitives; i++)
        if (sc.primitives[i].material_index >= sc.nmaterials) throw std::runtime_error("scene: primitive " + std::to_string(i) + " names a material that does not exist");
}

inline void build_draw_list(const swr_scene_desc &sc, const swr_camera &cam, std::vector<swr_draw> &draws, int shard = 0,
                            int nshards = 1, int band_y0 = 0, int band_y1 = 0, int band_height = 0) {
    std::vector<uint32_t> nodes_by_distance(sc.nnodes);
    std::vector<float> key(sc.nnodes);
    for (uint32_t i = 0; i < sc.nnodes; i++) {
        nodes_by_distance[i] = i;
        const float *s = sc.nodes[i].bounding_sphere_world;
        float d[3] = {cam.position[0] - s[0], cam.position[1] - s[1], cam.position[2] - s[2]};
        key[i] = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
    }
    std::stable_sort(nodes_by_distance.begin(), nodes_by_distance.end(), [&](uint32_t a, uint32_t b) {
        float x = key[a], y = key[b];  // OrderedFloat: NaN is greatest
        bool xn = x != x, yn = y != y;
        if (xn || yn) return !xn && yn;
        return x < y;
    });
    draws.clear();
    uint32_t first_tri = 0, first_tri_t = 0, di = 0;
    for (uint32_t ni : nodes_by_distance) {
        const swr_node_desc &node = sc.nodes[ni];
        if (node.mesh_index < 0) continue;
        const swr_mesh_desc &mesh = sc.meshes[node.mesh_index];
        Mat4 model, vp, mvp;
        std::memcpy(model.m, node.transform, 64);
        std::memcpy(vp.m, cam.view_project_matrix, 64);
        mvp = mul(vp, model);  // renderer.rs:378
        for (int pass = 0; pass < 2; pass++) {  // render_mesh: primitives_opaque, then primitives_translucent (renderer.rs:386-420)
            for (uint32_t pi = mesh.first_primitive; pi < mesh.first_primitive + mesh.num_primitives; pi++) {
                const swr_primitive_desc &prim = sc.primitives[pi];
                const bool translucent = (sc.materials[prim.material_index].flags & SWR_MAT_TRANSLUCENT) != 0;
                if (translucent != (pass == 1)) continue;
                FrustumTestResult t = test_sphere_frustum(node.transform, prim.bounding_sphere, cam);
                if (t == FrustumTestResult::Outside) continue;
                uint32_t ntris = prim.nindices / 3;
                const bool band_culled = band_y1 > band_y0 && sphere_outside_row_band(node.transform, prim.bounding_sphere, cam, band_y0, band_y1, band_height);
                // translucent primitives are never sharded (sort-last composites the opaque visibility buffer only)
                const bool mine = translucent ? shard == 0 || nshards == 1 : (int)(di % (uint32_t)nshards) == shard;
                if (mine && !band_culled) {
                    swr_draw d{};
                    std::memcpy(d.model, model.m, 64);
                    std::memcpy(d.mvp, mvp.m, 64);
                    d.primitive = pi;
                    d.flags = (t == FrustumTestResult::Intersecting ? SWR_DRAW_CLIP : 0u) | (translucent ? SWR_DRAW_TRANSLUCENT : 0u);
                    d.first_triangle = translucent ? first_tri_t : first_tri;
                    draws.push_back(d);
                }
                if (translucent) {
                    first_tri_t += ntris;  // translucent packets have their own queue, hence their own submission ids
                } else {
                    first_tri += ntris;
                    di++;
                }
            }
        }
    }
}

// math.rs:34-39: on x86-64 the reference's rsqrt_vec IS the hardware estimate _mm_rsqrt_ps. Its value depends only on
// the exponent parity and the top K mantissa bits of the input (K = 10 on Intel parts), i.e. it is a small table that
// belongs to the CPU the host program runs on. Probe it once and hand it to the device (swr_set_rsqrt_table) so that the
// CUDA shading normalises exactly as the reference would on this very host. Returns K, or 0 when the estimate has no such
// structure here (then the device keeps rsqrtf()).
inline int probe_host_rsqrt_table(std::vector<uint32_t> &table) {
#if defined(__x86_64__) || defined(_M_X64)
    auto rs = [](uint32_t b) {
        float f;
        std::memcpy(&f, &b, 4);
        float r = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(f)));
        uint32_t o;
        std::memcpy(&o, &r, 4);
        return o;
    };
    for (int K = 8; K <= 16; K++) {
        const size_t n = (size_t)1 << K;
        table.assign(2 * n, 0);
        bool ok = true;
        for (int p = 0; p < 2 && ok; p++) {
            for (uint32_t k = 0; k < n; k++) table[(size_t)p * n + k] = rs(((127u + p) << 23) | (k << (23 - K)));
            // every mantissa must agree with its bucket (exhaustive for the accepted K, strided pre-check otherwise)
            for (uint32_t m = 0; m < (1u << 23) && ok; m += 61) ok = rs(((127u + p) << 23) | m) == table[(size_t)p * n + (m >> (23 - K))];
        }
        if (!ok) continue;
        for (int p = 0; p < 2 && ok; p++)
            for (uint32_t m = 0; m < (1u << 23) && ok; m++) ok = rs(((127u + p) << 23) | m) == table[(size_t)p * n + (m >> (23 - K))];
        for (uint32_t e = 1; e < 253 && ok; e += 2)  // exponent scaling: +2 in the exponent halves the result
            for (uint32_t m = 0; m < (1u << 23) && ok; m += 4099) ok = rs((e << 23) | m) - rs(((e + 2) << 23) | m) == (1u << 23);
        if (ok) return K;
    }
#endif
    table.clear();
    return 0;
}

// renderer.rs:258-290 on a given set of per-tile metering values (tile.center_luminance): trimmed mean of log2 luminance,
// target exposure from the post-tonemap mid grey, exponential smoothing over delta_time.
// state = {auto_exposure, auto_exposure_target, auto_exposure_ev}, updated in place.
inline void auto_exposure_step(float state[3], const float *tile_luminance, size_t n, float delta_time) {
    if (n == 0) return;
    std::vector<float> lum(tile_luminance, tile_luminance + n);
    for (float &v : lum) v = std::log2(std::fmax(v, 1e-4f));
    std::sort(lum.begin(), lum.end(), [](float a, float b) {  // f32::total_cmp
        int32_t x, y;
        std::memcpy(&x, &a, 4);
        std::memcpy(&y, &b, 4);
        x ^= (int32_t)((uint32_t)(x >> 31) >> 1);
        y ^= (int32_t)((uint32_t)(y >> 31) >> 1);
        return x < y;
    });
    size_t trim = (size_t)std::floor((float)n * 0.10f);
    trim = std::min(trim, (n - 1) / 2);
    float sum = 0.0f;
    for (size_t i = trim; i < n - trim; i++) sum += lum[i];
    const float mean_log = sum / (float)(n - 2 * trim);
    // tonemap_inverse_scalar(0.45) (util.rs:43-47)
    const float y = 0.45f, denom = std::fmax(1.0f + 0.2f - y, 1e-6f);
    const float meter_key = std::fmax((y * 0.2f) / denom, 1e-4f);
    const float target_ev = std::log2(meter_key) - mean_log;
    const float target = std::fmin(std::fmax(std::pow(2.0f, target_ev), 0.05f), 32.0f);
    state[1] = target;
    const float tev = std::log2(target);
    const float alpha = 1.0f - std::exp(-(std::fmax(delta_time, 0.0f) / 1.0f));
    state[2] += (tev - state[2]) * alpha;
    state[0] = std::pow(2.0f, state[2]);
}

// ---- renderer.rs:145-355 --------------------------------------------------------------
class Renderer {
   public:
    // Renderer::new(width, height); device selects the GPU (one Renderer per GPU).
    // lanes > 1: frames alternate between that many contexts of the device (own stream, own per-frame buffers, one shared
    // scene), so a caller that pipelines (render N+1 before waiting for N's pixels) gets frame N+1's geometry pass
    // overlapped with frame N's raster tail and shading. Every per-frame query (stats, luminance, blit) goes to the lane of
    // the frame rendered last. lanes = 1 is the plain single-stream renderer.
    Renderer(int width, int height, int device = 0, int lanes = 1) : width_(width), height_(height) {
        if (lanes < 1 || lanes > 4) throw std::runtime_error("Renderer: lanes must be 1..4");
        for (int l = 0; l < lanes; l++) {
            swr_ctx *c = swr_create(width, height, device);
            if (!c) {
                const std::string msg = std::string("swr_create: ") + swr_last_error(nullptr);
                for (swr_ctx *o : lanes_) swr_destroy(o);
                throw std::runtime_error(msg);
            }
            lanes_.push_back(c);
        }
        ctx_ = lanes_[0];
        auto_exposure_ = auto_exposure_target_ = SWR_DEFAULT_EXPOSURE;  // renderer.rs:194-196
        auto_exposure_ev_ = std::log2(SWR_DEFAULT_EXPOSURE);
        set_reference_rsqrt(true);
    }
    // Renderer::new over several GPUs of this process (sort-first, include/swr.h swr_multi_*): same methods, one frame.
    Renderer(int width, int height, const std::vector<int> &devices) : width_(width), height_(height) {
        multi_ = swr_multi_create(width, height, devices.data(), (int)devices.size(), SWR_MULTI_SORT_FIRST);
        if (!multi_) throw std::runtime_error(std::string("swr_multi_create: ") + swr_multi_last_error(nullptr));
        ctx_ = swr_multi_context(multi_, 0);
        auto_exposure_ = auto_exposure_target_ = SWR_DEFAULT_EXPOSURE;
        auto_exposure_ev_ = std::log2(SWR_DEFAULT_EXPOSURE);
        set_reference_rsqrt(true);
    }
    // Match the reference's normalize() on this host (default), or use the device's own rsqrtf().
    void set_reference_rsqrt(bool on) {
        static std::vector<uint32_t> table;
        static int bits = -1;
        if (bits < 0) bits = probe_host_rsqrt_table(table);
        const uint32_t *t = (on && bits > 0) ? table.data() : nullptr;
        if (multi_)
            check(swr_multi_set_rsqrt_table(multi_, t, t ? bits : 0), "swr_multi_set_rsqrt_table");
        else
            for (swr_ctx *c : lanes_) check(swr_set_rsqrt_table(c, t, t ? bits : 0), "swr_set_rsqrt_table");
        rsqrt_bits_ = on ? bits : 0;
    }
    int reference_rsqrt_bits() const { return rsqrt_bits_; }
    ~Renderer() {
        if (multi_) {
            swr_multi_destroy(multi_);
        } else {
            for (size_t l = lanes_.size(); l-- > 0;) swr_destroy(lanes_[l]);  // lane 0 owns the scene: last
        }
    }
    int lanes() const { return (int)lanes_.size(); }
    swr_ctx *lane_ctx(int i) const { return (i >= 0 && i < (int)lanes_.size()) ? lanes_[i] : nullptr; }
    swr_multi *multi() const { return multi_; }
    Renderer(const Renderer &) = delete;
    Renderer &operator=(const Renderer &) = delete;

    swr_ctx *ctx() const { return ctx_; }
    float auto_exposure() const { return auto_exposure_; }
    const std::vector<swr_draw> &draws() const { return draws_; }
    void set_tile_rows(int r0, int r1) {
        if (multi_) throw std::runtime_error("set_tile_rows: a multi-device renderer assigns its own row bands");
        for (swr_ctx *c : lanes_) check(swr_set_tile_rows(c, r0, r1), "swr_set_tile_rows");
        row0_ = r0;
        row1_ = r1;
    }
    // The upload is cached by descriptor address (the reference's Scene is immutable and outlives the renderer). A caller
    // that reuses an address for different contents says so here; the next render_scene uploads again.
    void invalidate_scene() { uploaded_ = nullptr; }

    // render_scene(&mut self, scene, camera) — renderer.rs:201
    void render_scene(const Scene &scene, const RenderCamera &camera) { render_scene(scene, camera.to_abi()); }
    void render_scene(const Scene &scene, const swr_camera &cam, bool shade = true, int shard = 0, int nshards = 1) {
        if (scene.desc != uploaded_) {  // immutable scene: upload on first sight (SURVEY §8b ownership)
            validate_scene_ranges(*scene.desc);
            if (multi_) {
                check(swr_multi_upload_scene(multi_, scene.desc), "swr_multi_upload_scene");
            } else {
                ctx_ = lanes_[0];
                check(swr_upload_scene(lanes_[0], scene.desc), "swr_upload_scene");
                for (size_t l = 1; l < lanes_.size(); l++) {
                    ctx_ = lanes_[l];
                    check(swr_share_scene(lanes_[l], lanes_[0]), "swr_share_scene");
                }
                cur_ = (int)lanes_.size() - 1;  // the next frame goes to lane 0
            }
            uploaded_ = scene.desc;
        }
        if (!multi_) {
            cur_ = (cur_ + 1) % (int)lanes_.size();
            ctx_ = lanes_[cur_];
        }
        if (multi_) {  // every device culls the full draw list against its own band (k_cull)
            if (!shade || nshards != 1) throw std::runtime_error("multi-device renderer: sort-first only (shade = 1, no shards)");
            build_draw_list(*scene.desc, cam, draws_);
            check(swr_multi_render(multi_, &cam, draws_.data(), (int)draws_.size()), "swr_multi_render");
            return;
        }
        const int tiles_y = (height_ + SWR_TILE_SIZE - 1) / SWR_TILE_SIZE;
        const bool band = row1_ > row0_ && (row0_ > 0 || row1_ < tiles_y);  // sort-first: drop draws that cannot reach my rows
        build_draw_list(*scene.desc, cam, draws_, shard, nshards, band ? row0_ * SWR_TILE_SIZE : 0, band ? row1_ * SWR_TILE_SIZE : 0, height_);
        // The exposure blit_to_buffer will use is the one held now unless update_auto_exposure moves it in between
        // (renderer.rs:258-290 is the only writer). While the meter is idle the frame is shaded straight to RGBA8 with it
        // (swr_set_fixed_exposure); if the exposure does move, frame_exposure() shades the frame again to HDR and later
        // frames take the HDR path until the meter has settled.
        const float fixed = (shade && nshards == 1 && !meter_live_) ? auto_exposure_ : 0.0f;
        check(swr_set_fixed_exposure(ctx_, fixed), "swr_set_fixed_exposure");
        check(swr_render(ctx_, &cam, draws_.data(), (int)draws_.size(), shade ? 1 : 0), "swr_render");
        if (fused_.size() != lanes_.size()) fused_.assign(lanes_.size(), 0.0f);
        fused_[cur_] = fixed;
        last_cam_ = cam;
    }

    // The frame rendered last must be resolvable with `exposure`: a frame that was shaded to RGBA8 with another exposure is
    // shaded again from its visibility buffer, this time to HDR.
    void frame_exposure(float exposure) {
        if (multi_ || fused_.empty() || fused_[cur_] == 0.0f || fused_[cur_] == exposure) return;
        frame_hdr();
    }
    // The frame rendered last must hold HDR colour (read_color, a caller that wants to meter before it picks an exposure).
    void frame_hdr() {
        if (multi_ || fused_.empty() || fused_[cur_] == 0.0f) return;
        check(swr_set_fixed_exposure(ctx_, 0.0f), "swr_set_fixed_exposure");
        check(swr_shade(ctx_, &last_cam_), "swr_shade");
        fused_[cur_] = 0.0f;
    }

    // update_auto_exposure(&mut self, delta_time) — renderer.rs:258-290
    void update_auto_exposure(float delta_time) {
        swr_frame_stats st;
        check(multi_ ? swr_multi_get_stats(multi_, &st) : swr_get_stats(ctx_, &st), "swr_get_stats");
        if (st.tiles == 0) return;
        std::vector<float> lum(st.tiles);
        check(multi_ ? swr_multi_read_tile_luminance(multi_, lum.data()) : swr_read_tile_luminance(ctx_, lum.data()), "swr_read_tile_luminance");
        // sort-first: only the tiles of the rows this renderer owns were shaded; the others hold no metering value
        if (row1_ > row0_) {
            const size_t tiles_x = (size_t)(width_ + SWR_TILE_SIZE - 1) / SWR_TILE_SIZE;
            const size_t t0 = std::min((size_t)row0_ * tiles_x, lum.size()), t1 = std::min((size_t)row1_ * tiles_x, lum.size());
            lum = std::vector<float>(lum.begin() + t0, lum.begin() + t1);
        }
        float state[3] = {auto_exposure_, auto_exposure_target_, auto_exposure_ev_};
        auto_exposure_step(state, lum.data(), lum.size(), delta_time);
        meter_live_ = state[0] != auto_exposure_;  // the exposure is moving: frames are shaded to HDR until it holds still
        auto_exposure_ = state[0];
        auto_exposure_target_ = state[1];
        auto_exposure_ev_ = state[2];
    }

    // blit_to_buffer(&self, buffer) — renderer.rs:293-355
    void blit_to_buffer(RenderBuffer &buffer) {
        if ((int)buffer.width != width_ || (int)buffer.height < height_) throw std::runtime_error("RenderBuffer size mismatch");
        frame_exposure(auto_exposure_);
        if (multi_)
            check(swr_multi_resolve(multi_, auto_exposure_, buffer.pixels), "swr_multi_resolve");
        else
            check(swr_resolve(ctx_, auto_exposure_, buffer.pixels), "swr_resolve");
    }

    // Pipelined blit (the reference's App overlaps present(N-1) with render(N), main.rs:526-597): starts resolve + read-back
    // of the frame just rendered into `buffer` (pinned memory) and returns a ticket; wait_blit(ticket) completes it.
    int blit_to_buffer_async(RenderBuffer &buffer) {
        if ((int)buffer.width != width_ || (int)buffer.height < height_) throw std::runtime_error("RenderBuffer size mismatch");
        int ticket = -1;
        if (multi_) {  // the assembled frame lives on device 0 and is handed back by the peer protocol: synchronous form
            check(swr_multi_resolve(multi_, auto_exposure_, buffer.pixels), "swr_multi_resolve");
            return 0;
        }
        frame_exposure(auto_exposure_);
        check(swr_resolve_async(ctx_, auto_exposure_, buffer.pixels, &ticket), "swr_resolve_async");
        return cur_ * 2 + ticket;  // the lane that holds the frame + its pixel buffer
    }
    void wait_blit(int ticket) {
        if (multi_) return;
        if (ticket < 0 || ticket >= 2 * (int)lanes_.size()) throw std::runtime_error("wait_blit: unknown ticket");
        swr_ctx *c = lanes_[ticket >> 1];
        if (swr_wait_pixels(c, ticket & 1) != 0) throw std::runtime_error(std::string("swr_wait_pixels: ") + swr_last_error(c));
    }

   private:
    void check(int rc, const char *what) {
        if (rc != 0) throw std::runtime_error(std::string(what) + ": " + (multi_ && std::strncmp(what, "swr_multi", 9) == 0 ? swr_multi_last_error(multi_) : swr_last_error(ctx_)));
    }
    int width_, height_;
    std::vector<float> fused_;  // per lane: the fixed exposure its last frame was shaded with (0 = HDR)
    bool meter_live_ = false;   // update_auto_exposure changed the exposure last time it ran
    swr_camera last_cam_{};
    int rsqrt_bits_ = 0;
    int row0_ = 0, row1_ = 0;
    std::vector<swr_ctx *> lanes_;  // single device: the contexts frames alternate between (lane 0 owns the scene)
    int cur_ = 0;                   // lane of the frame rendered last
    swr_ctx *ctx_ = nullptr;        // = lanes_[cur_]; with multi_: the context of devices[0] (owned by multi_)
    swr_multi *multi_ = nullptr;
    const swr_scene_desc *uploaded_ = nullptr;
    std::vector<swr_draw> draws_;
    float auto_exposure_, auto_exposure_target_, auto_exposure_ev_;
};

}  // namespace swr

// oracle_sunvis.cpp — CPU restatement of the reference's voxel sun-visibility bake. TEST INFRASTRUCTURE: the
// product traces on the GPU (swraster-viewer_b200/csrc/swr_bake.cuh k_sunvis_trace / k_sunvis_blur through a hierarchy the
// host mirror builds); only tests/ call this file, as the checker. Deliberately hierarchy-free: every ray is tested against
// every triangle, so nothing of the product's BVH (or its conservative slab test) is shared with the checker.
//   RayTracer::new            src/raytracer.rs:71-132   world-space triangles, degenerate ones skipped
//   ray_triangle_intersect    src/raytracer.rs:223-259  Moeller-Trumbore, |det| <= 1e-8 rejected, u in [0,1], v >= 0, u+v <= 1
//   trace_transmittance       src/raytracer.rs:177-211  every hit in [t_min, t_max] sorted by t; opaque -> 0, translucent multiplies
//   build_active_voxel_mask   src/gi.rs:151-265         barycentric point samples per triangle -> occupied -> dilated by 2
//   compute_sun_visibility    src/gi.rs:267-314         origin = voxel centre + L * 3 |voxel_size|
//   blur_grid / blur_intensity src/voxelgrid.rs:371-419 3x3x3 mean, squared
// Parity unpinned against the Rust binary (no toolchain); pinned by the hand-made cases of tests/test_sunvis.py.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>
#include "../include/swr.h"

namespace {
struct V3 {
    float x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { return a * (1.0f / length(a)); }

struct Tri {
    V3 p0, p1, p2;
    uint32_t material;
};

std::vector<Tri> collect(const swr_scene_desc &sc) {  // raytracer.rs:71-132
    std::vector<Tri> tris;
    for (uint32_t ni = 0; ni < sc.nnodes; ni++) {
        const swr_node_desc &node = sc.nodes[ni];
        if (node.mesh_index < 0) continue;
        const swr_mesh_desc &mesh = sc.meshes[node.mesh_index];
        for (uint32_t pi = mesh.first_primitive; pi < mesh.first_primitive + mesh.num_primitives; pi++) {
            const swr_primitive_desc &p = sc.primitives[pi];
            for (uint32_t t = 0; t + 2 < p.nindices; t += 3) {
                V3 v[3];
                for (int k = 0; k < 3; k++) {
                    const float *q = p.positions + 4 * (size_t)p.indices[t + k];
                    const float *m = node.transform;  // glam Mat4 * Vec4: ((c0*x + c1*y) + c2*z) + c3*w
                    float w[3];
                    for (int r = 0; r < 3; r++) w[r] = ((m[r] * q[0] + m[4 + r] * q[1]) + m[8 + r] * q[2]) + m[12 + r] * q[3];
                    v[k] = V3{w[0], w[1], w[2]};
                }
                const V3 fn = cross(v[1] - v[0], v[2] - v[0]);
                if (dot(fn, fn) <= 1.0e-12f) continue;
                tris.push_back(Tri{v[0], v[1], v[2], p.material_index});
            }
        }
    }
    return tris;
}

bool hit(V3 origin, V3 direction, const Tri &tri, float t_min, float t_max, float &t_out) {  // raytracer.rs:223-259
    const V3 edge1 = tri.p1 - tri.p0, edge2 = tri.p2 - tri.p0;
    const V3 pvec = cross(direction, edge2);
    const float det = dot(edge1, pvec);
    if (std::fabs(det) <= 1.0e-8f) return false;
    const float inv_det = 1.0f / det;
    const V3 tvec = origin - tri.p0;
    const float u = dot(tvec, pvec) * inv_det;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    const V3 qvec = cross(tvec, edge1);
    const float v = dot(direction, qvec) * inv_det;
    if (v < 0.0f || (u + v) > 1.0f) return false;
    const float t = dot(edge2, qvec) * inv_det;
    if (t < t_min || t > t_max) return false;
    t_out = t;
    return true;
}
}  // namespace

extern "C" {

// build_active_voxel_mask (gi.rs:151-265): out[z*W*H + y*W + x] = 1 for voxels within 2 of a voxel a triangle sample falls in
int orc_sunvis_active_mask(const swr_scene_desc *sc, const swr_voxel_grid_desc *g, uint8_t *out) {
    const std::vector<Tri> tris = collect(*sc);
    const size_t W = g->dims[0], H = g->dims[1], D = g->dims[2], total = W * H * D;
    const V3 vs{(g->world_max[0] - g->world_min[0]) / (float)W, (g->world_max[1] - g->world_min[1]) / (float)H, (g->world_max[2] - g->world_min[2]) / (float)D};
    std::vector<uint8_t> occupied(total, 0);
    const float min_edge = std::fmax(std::fmin(vs.x, std::fmin(vs.y, vs.z)), 1.0e-6f);
    const float area_ref = min_edge * min_edge;
    for (const Tri &tri : tris) {
        const float tri_area = 0.5f * length(cross(tri.p1 - tri.p0, tri.p2 - tri.p0));
        const float want = std::ceil((tri_area / area_ref) * 2.0f);
        const size_t target = want >= 4096.0f ? 4096 : (want <= 1.0f || !(want == want) ? 1 : (size_t)want);
        const size_t n = (size_t)std::ceil(std::sqrt((float)target));
        for (size_t iu = 0; iu < n; iu++)
            for (size_t iv = 0; iv < n - iu; iv++) {
                const float u = ((float)iu + 0.5f) / (float)n, v = ((float)iv + 0.5f) / (float)n, w = 1.0f - u - v;
                if (w < 0.0f) continue;
                const V3 p = (tri.p0 * w + tri.p1 * u) + tri.p2 * v;
                const float fx = (p.x - g->world_min[0]) / vs.x, fy = (p.y - g->world_min[1]) / vs.y, fz = (p.z - g->world_min[2]) / vs.z;
                if (fx < 0.0f || fy < 0.0f || fz < 0.0f) continue;
                const float ffx = std::floor(fx), ffy = std::floor(fy), ffz = std::floor(fz);
                if (!(ffx < (float)W && ffy < (float)H && ffz < (float)D)) continue;
                occupied[((size_t)ffz * H + (size_t)ffy) * W + (size_t)ffx] = 1;
            }
    }
    std::copy(occupied.begin(), occupied.end(), out);
    const size_t R = 2;  // GI_ACTIVE_DILATION_RADIUS
    for (size_t z = 0; z < D; z++)
        for (size_t y = 0; y < H; y++)
            for (size_t x = 0; x < W; x++) {
                if (!occupied[(z * H + y) * W + x]) continue;
                for (size_t nz = z >= R ? z - R : 0; nz <= std::min(z + R, D - 1); nz++)
                    for (size_t ny = y >= R ? y - R : 0; ny <= std::min(y + R, H - 1); ny++)
                        for (size_t nx = x >= R ? x - R : 0; nx <= std::min(x + R, W - 1); nx++) out[(nz * H + ny) * W + nx] = 1;
            }
    return 0;
}

// compute_sun_visibility (gi.rs:267-314) + blur_grid (voxelgrid.rs:371-419): blurred, squared light intensity per voxel
int orc_sun_visibility(const swr_scene_desc *sc, const swr_voxel_grid_desc *g, const float *light_direction, float *out) {
    const size_t W = g->dims[0], H = g->dims[1], D = g->dims[2], total = W * H * D;
    if (!total) return -1;
    const std::vector<Tri> tris = collect(*sc);
    std::vector<uint8_t> active(total);
    orc_sunvis_active_mask(sc, g, active.data());
    const V3 L = normalize(V3{light_direction[0], light_direction[1], light_direction[2]});
    const V3 vs{(g->world_max[0] - g->world_min[0]) / (float)W, (g->world_max[1] - g->world_min[1]) / (float)H, (g->world_max[2] - g->world_min[2]) / (float)D};
    const V3 center_min = V3{g->world_min[0], g->world_min[1], g->world_min[2]} + vs * 0.5f;
    const float bias = length(vs) * 3.0f;
    std::vector<float> vis(total, 1.0f);
    size_t nactive = 0;
    for (uint8_t a : active) nactive += a;
    if (!nactive) {
        std::copy(vis.begin(), vis.end(), out);
        return 0;
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t index = 0; index < (int64_t)total; index++) {
        if (!active[(size_t)index]) continue;
        const size_t z = (size_t)index / (W * H), rem = (size_t)index % (W * H), y = rem / W, x = rem % W;
        const V3 origin = (center_min + V3{(float)x * vs.x, (float)y * vs.y, (float)z * vs.z}) + L * bias;
        std::vector<std::pair<float, uint32_t>> hits;  // trace_transmittance (raytracer.rs:177-211), t in [1e-4, inf]
        for (uint32_t ti = 0; ti < tris.size(); ti++) {
            float t;
            if (hit(origin, L, tris[ti], 1.0e-4f, INFINITY, t)) hits.emplace_back(t, ti);
        }
        std::sort(hits.begin(), hits.end());  // by t; ties by triangle index (the reference's tie order depends on its BVH)
        float transmittance = 1.0f;
        for (const auto &h : hits) {
            const swr_material_desc &m = sc->materials[tris[h.second].material];
            if (!(m.flags & SWR_MAT_TRANSLUCENT)) {
                transmittance = 0.0f;
                break;
            }
            transmittance *= m.transmission;
            if (transmittance <= 0.0001f) {
                transmittance = 0.0f;
                break;
            }
        }
        vis[(size_t)index] = transmittance;
    }
    for (size_t z = 0; z < D; z++)
        for (size_t y = 0; y < H; y++)
            for (size_t x = 0; x < W; x++) {
                float sum = 0.0f;
                int count = 0;
                for (int dz = -1; dz <= 1; dz++)
                    for (int dy = -1; dy <= 1; dy++)
                        for (int dx = -1; dx <= 1; dx++) {
                            const int64_t nx = (int64_t)x + dx, ny = (int64_t)y + dy, nz = (int64_t)z + dz;
                            if (nx < 0 || ny < 0 || nz < 0 || nx >= (int64_t)W || ny >= (int64_t)H || nz >= (int64_t)D) continue;
                            sum += vis[((size_t)nz * H + (size_t)ny) * W + (size_t)nx];
                            count++;
                        }
                out[(z * H + y) * W + x] = count ? std::pow(sum / (float)count, 2.0f) : 0.0f;
            }
    return 0;
}

}  // extern "C"

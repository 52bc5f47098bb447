// swr_sunvis.hpp — voxel sun visibility (the part the default load path runs: main.rs:237-246).
// The reference marks the voxels near geometry, casts one ray per such voxel towards the sun through a BVH and blurs the
// result; the value lands in gi_sh4[voxel][0].w and is what the shader uses to shadow the direct light.
// Here the ray cast and the blur run on the GPU (include/swr.h swr_bake_sun_visibility, csrc/swr_bake.cuh); the host builds
// what they consume: world-space triangles, the hierarchy, the active-voxel mask. No CPU fallback for the trace (the CPU
// restatement the tests check against is hierarchy-free and lives outside this package).
//   RayTracer::new            src/raytracer.rs:71-132   world-space triangles, degenerate ones skipped, AABB padded by 1e-5
//   ray_triangle_intersect    src/raytracer.rs:223-259  Moeller-Trumbore, |det| <= 1e-8 rejected, u in [0,1], v >= 0, u+v <= 1
//   trace_transmittance       src/raytracer.rs:177-211  every hit in [t_min, t_max], sorted by t; opaque -> 0, translucent multiplies
//   build_active_voxel_mask   src/gi.rs:151-265         barycentric point samples per triangle -> occupied -> dilated by 2
//   compute_sun_visibility    src/gi.rs:267-314         origin = voxel centre + L * 3 |voxel_size|, then VoxelGrid::blur_grid
//   blur_grid / blur_intensity src/voxelgrid.rs:371-419 3x3x3 mean, squared
// The `bvh` crate (0.11, Cargo.lock) is not vendored; any exact BVH gives the same answer because a voxel's value only depends
// on the set of triangles the ray really hits, which the Moeller-Trumbore test decides — the hierarchy below (median split,
// conservative slab test) merely has to offer a superset of them. The full GI bake (gi.rs:316-408, `--gi`) is not built.
#pragma once
#include "swr_gltf.hpp"

namespace swr {
namespace sunvis {

using gltf::V3;

struct Triangle {
    V3 p0, p1, p2, n;
    uint32_t material_index;
};

// raytracer.rs:71-132
inline std::vector<Triangle> collect_triangles(const swr_scene_desc &sc) {
    std::vector<Triangle> tris;
    for (uint32_t ni = 0; ni < sc.nnodes; ni++) {
        const swr_node_desc &node = sc.nodes[ni];
        if (node.mesh_index < 0) continue;
        const swr_mesh_desc &mesh = sc.meshes[node.mesh_index];
        for (uint32_t pi = mesh.first_primitive; pi < mesh.first_primitive + mesh.num_primitives; pi++) {
            const swr_primitive_desc &p = sc.primitives[pi];
            for (uint32_t t = 0; t + 2 < p.nindices; t += 3) {
                V3 v[3];
                for (int k = 0; k < 3; k++) {
                    float w[4];
                    mul_vec4(node.transform, p.positions + 4 * (size_t)p.indices[t + k], w);
                    v[k] = V3{w[0], w[1], w[2]};
                }
                const V3 fn = gltf::cross(v[1] - v[0], v[2] - v[0]);
                if (gltf::dot(fn, fn) <= 1.0e-12f) continue;
                tris.push_back(Triangle{v[0], v[1], v[2], gltf::normalize(fn), p.material_index});
            }
        }
    }
    return tris;
}

class Bvh {
   public:
    explicit Bvh(const std::vector<Triangle> &tris) : tris_(tris) {
        const size_t n = tris.size();
        order_.resize(n);
        lo_.resize(n), hi_.resize(n), centroid_.resize(n);
        for (size_t i = 0; i < n; i++) {
            order_[i] = (uint32_t)i;
            const Triangle &t = tris[i];
            const float pad = 1.0e-5f;  // raytracer.rs:104-105
            lo_[i] = V3{std::fmin(t.p0.x, std::fmin(t.p1.x, t.p2.x)) - pad, std::fmin(t.p0.y, std::fmin(t.p1.y, t.p2.y)) - pad, std::fmin(t.p0.z, std::fmin(t.p1.z, t.p2.z)) - pad};
            hi_[i] = V3{std::fmax(t.p0.x, std::fmax(t.p1.x, t.p2.x)) + pad, std::fmax(t.p0.y, std::fmax(t.p1.y, t.p2.y)) + pad, std::fmax(t.p0.z, std::fmax(t.p1.z, t.p2.z)) + pad};
            centroid_[i] = (lo_[i] + hi_[i]) * 0.5f;
        }
        if (n) {
            nodes_.push_back(Node{});
            build_into(0, 0, (uint32_t)n);
        }
    }

    // flat tables for the device traversal (swr_bvh_node: leaf = order[first, first + count); inner = children first, first + 1)
    std::vector<swr_bvh_node> flat_nodes() const {
        std::vector<swr_bvh_node> out(nodes_.size());
        for (size_t i = 0; i < nodes_.size(); i++) {
            const Node &n = nodes_[i];
            out[i] = swr_bvh_node{{n.lo.x, n.lo.y, n.lo.z}, {n.hi.x, n.hi.y, n.hi.z}, n.first, n.count};
        }
        return out;
    }
    const std::vector<uint32_t> &order() const { return order_; }

   private:
    struct Node {
        V3 lo, hi;
        uint32_t first, count;  // leaf: range in order_; inner: first child (children are adjacent), count = 0
    };
    const std::vector<Triangle> &tris_;
    std::vector<uint32_t> order_;
    std::vector<V3> lo_, hi_, centroid_;
    std::vector<Node> nodes_;

    // fills nodes_[slot] (already allocated) for the triangles order_[begin, end)
    void build_into(uint32_t slot, uint32_t begin, uint32_t end) {
        V3 lo{INFINITY, INFINITY, INFINITY}, hi{-INFINITY, -INFINITY, -INFINITY}, clo = lo, chi = hi;
        for (uint32_t i = begin; i < end; i++) {
            const uint32_t t = order_[i];
            lo = V3{std::fmin(lo.x, lo_[t].x), std::fmin(lo.y, lo_[t].y), std::fmin(lo.z, lo_[t].z)};
            hi = V3{std::fmax(hi.x, hi_[t].x), std::fmax(hi.y, hi_[t].y), std::fmax(hi.z, hi_[t].z)};
            clo = V3{std::fmin(clo.x, centroid_[t].x), std::fmin(clo.y, centroid_[t].y), std::fmin(clo.z, centroid_[t].z)};
            chi = V3{std::fmax(chi.x, centroid_[t].x), std::fmax(chi.y, centroid_[t].y), std::fmax(chi.z, centroid_[t].z)};
        }
        const float ext[3] = {chi.x - clo.x, chi.y - clo.y, chi.z - clo.z};
        const int axis = ext[0] >= ext[1] && ext[0] >= ext[2] ? 0 : (ext[1] >= ext[2] ? 1 : 2);
        if (end - begin <= 4 || !(ext[axis] > 0.0f)) {
            nodes_[slot] = Node{lo, hi, begin, end - begin};
            return;
        }
        const uint32_t mid = begin + (end - begin) / 2;
        auto key = [&](uint32_t t) { return axis == 0 ? centroid_[t].x : axis == 1 ? centroid_[t].y : centroid_[t].z; };
        std::nth_element(order_.begin() + begin, order_.begin() + mid, order_.begin() + end, [&](uint32_t a, uint32_t b) { return key(a) < key(b); });
        const uint32_t left = (uint32_t)nodes_.size();
        nodes_.push_back(Node{});  // the two children sit next to each other
        nodes_.push_back(Node{});
        nodes_[slot] = Node{lo, hi, left, 0};
        build_into(left, begin, mid);
        build_into(left + 1, mid, end);
    }
};

// gi.rs:151-265 (the mask only; surface normals / albedo feed the full GI bake, which is not built)
inline std::vector<uint8_t> build_active_voxel_mask(const std::vector<Triangle> &tris, const swr_voxel_grid_desc &g) {
    const size_t W = g.dims[0], H = g.dims[1], D = g.dims[2], total = W * H * D;
    const V3 vs{(g.world_max[0] - g.world_min[0]) / (float)W, (g.world_max[1] - g.world_min[1]) / (float)H, (g.world_max[2] - g.world_min[2]) / (float)D};
    std::vector<uint8_t> occupied(total, 0);
    const float min_edge = std::fmax(std::fmin(vs.x, std::fmin(vs.y, vs.z)), 1.0e-6f);
    const float area_ref = min_edge * min_edge;
    for (const Triangle &tri : tris) {
        const V3 e1 = tri.p1 - tri.p0, e2 = tri.p2 - tri.p0;
        const float tri_area = 0.5f * gltf::length(gltf::cross(e1, e2));
        float want = std::ceil((tri_area / area_ref) * 2.0f);
        size_t target = want >= 4096.0f ? 4096 : (want <= 1.0f || !(want == want) ? 1 : (size_t)want);
        const size_t n = (size_t)std::ceil(std::sqrt((float)target));
        for (size_t iu = 0; iu < n; iu++)
            for (size_t iv = 0; iv < n - iu; iv++) {
                const float u = ((float)iu + 0.5f) / (float)n, v = ((float)iv + 0.5f) / (float)n, w = 1.0f - u - v;
                if (w < 0.0f) continue;
                const V3 p = (tri.p0 * w + tri.p1 * u) + tri.p2 * v;
                const float fx = (p.x - g.world_min[0]) / vs.x, fy = (p.y - g.world_min[1]) / vs.y, fz = (p.z - g.world_min[2]) / vs.z;
                if (fx < 0.0f || fy < 0.0f || fz < 0.0f) continue;
                const float ffx = std::floor(fx), ffy = std::floor(fy), ffz = std::floor(fz);
                if (!(ffx < (float)W && ffy < (float)H && ffz < (float)D)) continue;
                occupied[((size_t)ffz * H + (size_t)ffy) * W + (size_t)ffx] = 1;
            }
    }
    std::vector<uint8_t> dilated(occupied);
    const size_t R = 2;  // GI_ACTIVE_DILATION_RADIUS
    for (size_t z = 0; z < D; z++)
        for (size_t y = 0; y < H; y++)
            for (size_t x = 0; x < W; x++) {
                if (!occupied[(z * H + y) * W + x]) continue;
                for (size_t nz = z >= R ? z - R : 0; nz <= std::min(z + R, D - 1); nz++)
                    for (size_t ny = y >= R ? y - R : 0; ny <= std::min(y + R, H - 1); ny++)
                        for (size_t nx = x >= R ? x - R : 0; nx <= std::min(x + R, W - 1); nx++) dilated[(nz * H + ny) * W + nx] = 1;
            }
    return dilated;
}

// gi.rs:267-314 + voxelgrid.rs:371-419. Returns the blurred light intensity per voxel (index = z*W*H + y*W + x): triangles,
// hierarchy and active mask from the host, rays and blur on the device.
inline std::vector<float> compute_sun_visibility(const swr_scene_desc &sc, const swr_voxel_grid_desc &g, const float light_direction[3]) {
    const size_t W = g.dims[0], H = g.dims[1], D = g.dims[2], total = W * H * D;
    if (!total) throw std::runtime_error("Invalid data: empty voxel grid");
    const std::vector<Triangle> tris = collect_triangles(sc);
    const Bvh bvh(tris);
    const std::vector<uint8_t> active = build_active_voxel_mask(tris, g);
    std::vector<swr_sun_triangle> flat(tris.size());
    for (size_t i = 0; i < tris.size(); i++) {
        const Triangle &t = tris[i];
        const swr_material_desc &m = sc.materials[t.material_index];
        flat[i] = swr_sun_triangle{{t.p0.x, t.p0.y, t.p0.z}, {t.p1.x, t.p1.y, t.p1.z}, {t.p2.x, t.p2.y, t.p2.z},
                                   (m.flags & SWR_MAT_TRANSLUCENT) ? std::fmax(m.transmission, 0.0f) : -1.0f};
    }
    const std::vector<swr_bvh_node> nodes = bvh.flat_nodes();
    swr_sunvis_desc d{};
    d.nodes = nodes.data();
    d.nnodes = (uint32_t)nodes.size();
    d.order = bvh.order().data();
    d.norder = (uint32_t)bvh.order().size();
    d.triangles = flat.data();
    d.ntriangles = (uint32_t)flat.size();
    d.active = active.data();
    for (int k = 0; k < 3; k++) {
        d.dims[k] = g.dims[k];
        d.world_min[k] = g.world_min[k];
        d.world_max[k] = g.world_max[k];
        d.light_direction[k] = light_direction[k];
    }
    std::vector<float> out(total, 1.0f);
    if (swr_bake_sun_visibility(-1, &d, out.data()) != SWR_OK) throw std::runtime_error(std::string("swr_bake_sun_visibility: ") + swr_bake_last_error());
    return out;
}

}  // namespace sunvis
}  // namespace swr

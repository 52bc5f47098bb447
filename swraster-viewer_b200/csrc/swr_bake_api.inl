// swr_bake_api.inl — C entry points of the device bakes (include/swr.h, swr_bake_*). Included by swr_api.cu inside extern "C".
// Stateless: every call allocates, launches on the legacy default stream of `device`, copies back and frees. No CPU fallback:
// without a usable sm_100 device the call fails.
static thread_local std::string g_bake_error;
const char *swr_bake_last_error(void) { return g_bake_error.c_str(); }

#define BK(call)                                                                 \
    do {                                                                         \
        cudaError_t e__ = (call);                                                \
        if (e__ != cudaSuccess) {                                                \
            g_bake_error = std::string(#call) + ": " + cudaGetErrorString(e__);  \
            for (void *p__ : allocs) cudaFree(p__);                              \
            return SWR_ERR_CUDA;                                                 \
        }                                                                        \
    } while (0)

static int bake_select_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_bake_error = "no usable CUDA device (the bakes have no CPU fallback)";
        return SWR_ERR_NO_DEVICE;
    }
    if (device < 0) cudaGetDevice(&device);
    if (device >= ndev) {
        g_bake_error = "device ordinal out of range";
        return SWR_ERR_INVALID;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
        g_bake_error = "device is not sm_100 class: this library carries sm_100a code only";
        return SWR_ERR_NO_DEVICE;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        g_bake_error = "cudaSetDevice failed";
        return SWR_ERR_CUDA;
    }
    return SWR_OK;
}

int swr_bake_brdf_lut(int device, uint32_t size, uint32_t *out_texels) {
    if (!out_texels || size == 0 || size > 4096) {
        g_bake_error = "swr_bake_brdf_lut: size must be 1..4096 and out_texels non-NULL";
        return SWR_ERR_INVALID;
    }
    int rc = bake_select_device(device);
    if (rc) return rc;
    std::vector<void *> allocs;
    uint32_t *d = nullptr;
    const size_t n = (size_t)size * size;
    BK(cudaMalloc(&d, n * 4));
    allocs.push_back(d);
    bake::k_bake_brdf_lut<<<(unsigned)((n + 127) / 128), 128>>>(size, d);
    BK(cudaGetLastError());
    BK(cudaMemcpy(out_texels, d, n * 4, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return SWR_OK;
}

static int bake_upload_cube(const uint32_t *faces, uint32_t w, uint32_t h, std::vector<void *> &allocs, bake::Cube &cube) {
    if (!faces || w == 0 || h == 0 || (uint64_t)w * h * 6 > (1u << 28)) {
        g_bake_error = "cubemap faces: need six w x h RGBA8 faces (face-major), w * h * 6 <= 2^28";
        return SWR_ERR_INVALID;
    }
    uint32_t *d = nullptr;
    BK(cudaMalloc(&d, (size_t)w * h * 6 * 4));
    allocs.push_back(d);
    BK(cudaMemcpy(d, faces, (size_t)w * h * 6 * 4, cudaMemcpyHostToDevice));
    cube.texels = d;
    cube.w = w;
    cube.h = h;
    return SWR_OK;
}

int swr_bake_irradiance_sh4(int device, const uint32_t *cubemap_faces, uint32_t w, uint32_t h, float *out12) {
    if (!out12) {
        g_bake_error = "swr_bake_irradiance_sh4: out12 is NULL";
        return SWR_ERR_INVALID;
    }
    int rc = bake_select_device(device);
    if (rc) return rc;
    std::vector<void *> allocs;
    bake::Cube cube{};
    if ((rc = bake_upload_cube(cubemap_faces, w, h, allocs, cube))) return rc;
    const uint32_t n = 6u * w * h, nblocks = (n + SH_BLOCK - 1) / SH_BLOCK;
    float *partial = nullptr, *res = nullptr;
    BK(cudaMalloc(&partial, (size_t)nblocks * 12 * 4));
    allocs.push_back(partial);
    BK(cudaMalloc(&res, 12 * 4));
    allocs.push_back(res);
    bake::k_bake_sh4_partial<<<nblocks, SH_BLOCK>>>(cube, partial);
    bake::k_bake_sh4_final<<<1, 32>>>(partial, nblocks, res);
    BK(cudaGetLastError());
    BK(cudaMemcpy(out12, res, 12 * 4, cudaMemcpyDeviceToHost));
    for (void *p : allocs) cudaFree(p);
    return SWR_OK;
}

int swr_bake_prefilter_specular(int device, const uint32_t *cubemap_faces, uint32_t w, uint32_t h, uint32_t sample_count, uint32_t *out_texels) {
    if (!out_texels || sample_count == 0 || sample_count > 65536) {
        g_bake_error = "swr_bake_prefilter_specular: out_texels is NULL or sample_count outside 1..65536";
        return SWR_ERR_INVALID;
    }
    int rc = bake_select_device(device);
    if (rc) return rc;
    std::vector<void *> allocs;
    bake::Cube cube{};
    if ((rc = bake_upload_cube(cubemap_faces, w, h, allocs, cube))) return rc;
    uint32_t num_mips = 1;
    for (uint32_t m = std::max(w, h); m >>= 1;) num_mips++;
    const size_t per_mip = (size_t)w * h * 6;
    uint32_t *d = nullptr;
    BK(cudaMalloc(&d, per_mip * num_mips * 4));
    allocs.push_back(d);
    for (uint32_t mip = 0; mip < num_mips; mip++)
        bake::k_bake_prefilter<<<(unsigned)((per_mip + 127) / 128), 128>>>(cube, mip, num_mips - 1, sample_count, d + per_mip * mip);
    BK(cudaGetLastError());
    BK(cudaMemcpy(out_texels, d, per_mip * num_mips * 4, cudaMemcpyDeviceToHost));
    for (void *p : allocs) cudaFree(p);
    return (int)num_mips;
}

int swr_bake_sun_visibility(int device, const swr_sunvis_desc *s, float *out_per_voxel) {
    if (!s || !out_per_voxel || !s->active || s->dims[0] == 0 || s->dims[1] == 0 || s->dims[2] == 0) {
        g_bake_error = "swr_bake_sun_visibility: NULL argument or empty voxel grid";
        return SWR_ERR_INVALID;
    }
    if (s->nnodes && (!s->nodes || !s->order || !s->triangles)) {
        g_bake_error = "swr_bake_sun_visibility: hierarchy tables are NULL";
        return SWR_ERR_INVALID;
    }
    for (uint32_t i = 0; i < s->nnodes; i++) {  // the traversal trusts these indices
        const swr_bvh_node &n = s->nodes[i];
        const bool ok = n.count ? ((uint64_t)n.first + n.count <= s->norder) : ((uint64_t)n.first + 1 < s->nnodes);
        if (!ok) {
            g_bake_error = "swr_bake_sun_visibility: hierarchy node " + std::to_string(i) + " points outside its tables";
            return SWR_ERR_INVALID;
        }
    }
    for (uint32_t i = 0; i < s->norder; i++)
        if (s->order[i] >= s->ntriangles) {
            g_bake_error = "swr_bake_sun_visibility: order[] entry outside the triangle table";
            return SWR_ERR_INVALID;
        }
    int rc = bake_select_device(device);
    if (rc) return rc;
    static_assert(sizeof(bake::SunNode) == sizeof(swr_bvh_node) && sizeof(bake::SunTri) == sizeof(swr_sun_triangle), "ABI structs mirror the device structs");
    std::vector<void *> allocs;
    const size_t total = (size_t)s->dims[0] * s->dims[1] * s->dims[2];
    bake::SunParams P{};
    auto up = [&](const void *src, size_t bytes, void **dst) -> cudaError_t {
        *dst = nullptr;
        if (bytes == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(dst, bytes);
        if (e != cudaSuccess) return e;
        allocs.push_back(*dst);
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    void *dn, *dord, *dtri, *dact;
    BK(up(s->nodes, (size_t)s->nnodes * sizeof(swr_bvh_node), &dn));
    BK(up(s->order, (size_t)s->norder * 4, &dord));
    BK(up(s->triangles, (size_t)s->ntriangles * sizeof(swr_sun_triangle), &dtri));
    BK(up(s->active, total, &dact));
    float *d0 = nullptr, *d1 = nullptr;
    uint32_t *dovf = nullptr;
    BK(cudaMalloc(&d0, total * 4));
    allocs.push_back(d0);
    BK(cudaMalloc(&d1, total * 4));
    allocs.push_back(d1);
    BK(cudaMalloc(&dovf, 4));
    allocs.push_back(dovf);
    BK(cudaMemset(dovf, 0, 4));
    P.nodes = (const bake::SunNode *)dn;
    P.order = (const uint32_t *)dord;
    P.tris = (const bake::SunTri *)dtri;
    P.nnodes = s->nnodes;
    P.active = (const uint8_t *)dact;
    P.W = s->dims[0], P.H = s->dims[1], P.D = s->dims[2];
    // gi.rs:267-314: voxel size, first voxel centre, origin bias 3 |voxel|, L normalised — in f32 on the host, like the reference
    float vs[3], len2 = 0.0f;
    for (int k = 0; k < 3; k++) {
        vs[k] = (s->world_max[k] - s->world_min[k]) / (float)s->dims[k];
        P.vs[k] = vs[k];
        P.center_min[k] = s->world_min[k] + vs[k] * 0.5f;
    }
    len2 = (vs[0] * vs[0] + vs[1] * vs[1]) + vs[2] * vs[2];
    P.bias = std::sqrt(len2) * 3.0f;
    const float l2 = (s->light_direction[0] * s->light_direction[0] + s->light_direction[1] * s->light_direction[1]) + s->light_direction[2] * s->light_direction[2];
    const float rl = 1.0f / std::sqrt(l2);
    for (int k = 0; k < 3; k++) P.L[k] = s->light_direction[k] * rl;
    P.out = d0;
    P.overflow = dovf;
    bool any_active = false;
    for (size_t i = 0; i < total && !any_active; i++) any_active = s->active[i] != 0;
    const unsigned grid = (unsigned)((total + 127) / 128);
    if (s->nnodes == 0) {  // no geometry: every ray reaches the sun
        std::vector<float> ones(total, 1.0f);
        BK(cudaMemcpy(d0, ones.data(), total * 4, cudaMemcpyHostToDevice));
    } else {
        bake::k_sunvis_trace<<<grid, 128>>>(P);
    }
    float *result = d0;
    if (any_active) {  // the reference blurs only when something was traced (gi.rs:276-281)
        bake::k_sunvis_blur<<<grid, 128>>>(d0, d1, P.W, P.H, P.D);
        result = d1;
    }
    BK(cudaGetLastError());
    BK(cudaMemcpy(out_per_voxel, result, total * 4, cudaMemcpyDeviceToHost));
    uint32_t ovf = 0;
    BK(cudaMemcpy(&ovf, dovf, 4, cudaMemcpyDeviceToHost));
    for (void *p : allocs) cudaFree(p);
    if (ovf) {
        g_bake_error = "swr_bake_sun_visibility: a ray crossed more than 24 translucent surfaces or the hierarchy is deeper than 64 levels";
        return SWR_ERR_INVALID;
    }
    return SWR_OK;
}
#undef BK

// swr_multi.inl — several devices of ONE process behind one handle (include/swr.h, swr_multi_*): what a single
// `Renderer` (renderer.rs:145-355) needs to drive all GPUs of a box. Included by swr_api.cu inside extern "C".
// Sort-first: device i owns a contiguous, cost-balanced range of tile rows; every device gets the same draw list and culls
// it per cluster against its band (k_cull); the frame is assembled in device 0's pixel buffer by the other devices'
// resolve kernels storing over NVLink peer memory (swr_peer_* with direct peer pointers: no IPC, no collective).
// One persistent host thread per device enqueues that device's work, so a frame costs the host one enqueue time, not ndev.
// (swr_api.cu includes <condition_variable>, <functional>, <mutex>, <thread> ahead of its extern "C" block)

struct swr_multi {
    int W = 0, H = 0, tiles_x = 0, tiles_y = 0, ndev = 0, mode = 0;
    std::vector<swr_ctx *> ctx;
    std::vector<std::pair<int, int>> rows;  // per device: owned tile rows [begin, end)
    std::string err;
    uint32_t frame = 0;        // peer-assembly frame number (starts at 1)
    bool balanced = false;     // bands were derived from a probe frame of the current scene
    bool have_scene = false;
    // worker pool: one thread per device, run(job) executes job(i) on every worker and waits
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<int(int)> job;
    std::vector<int> rc;
    uint64_t generation = 0;
    int remaining = 0;
    bool quit = false;
};

static void multi_worker(swr_multi *m, int i) {
    cudaSetDevice(m->ctx[i]->device);
    uint64_t seen = 0;
    for (;;) {
        std::function<int(int)> job;
        {
            std::unique_lock<std::mutex> lk(m->mu);
            m->cv_go.wait(lk, [&] { return m->quit || m->generation != seen; });
            if (m->quit) return;
            seen = m->generation;
            job = m->job;
        }
        const int r = job(i);
        {
            std::lock_guard<std::mutex> lk(m->mu);
            m->rc[i] = r;
            if (--m->remaining == 0) m->cv_done.notify_all();
        }
    }
}

// Run job(i) for every device on its own thread; returns the first failure (and keeps that context's message).
static int multi_run(swr_multi *m, std::function<int(int)> job) {
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->job = std::move(job);
        m->remaining = m->ndev;
        m->generation++;
    }
    m->cv_go.notify_all();
    {
        std::unique_lock<std::mutex> lk(m->mu);
        m->cv_done.wait(lk, [&] { return m->remaining == 0; });
    }
    for (int i = 0; i < m->ndev; i++)
        if (m->rc[i] != SWR_OK) {
            m->err = "device " + std::to_string(m->ctx[i]->device) + ": " + m->ctx[i]->err;
            return m->rc[i];
        }
    return SWR_OK;
}

// Contiguous tile-row bands of (nearly) equal cost: row cost = measured raster cycles of its tiles + a per-tile shading share.
static void multi_balance(swr_multi *m, const std::vector<uint32_t> &cycles) {
    const int ty = m->tiles_y, tx = m->tiles_x, n = m->ndev;
    double mean = 0.0;
    for (uint32_t c : cycles) mean += c;
    mean /= (double)std::max<size_t>(cycles.size(), 1);
    const double pixel_cost = 0.9 * mean + 1.0;  // measured shade : raster ratio on C3
    std::vector<double> cum(ty + 1, 0.0);
    for (int r = 0; r < ty; r++) {
        double c = pixel_cost * tx;
        for (int x = 0; x < tx; x++) c += cycles[(size_t)r * tx + x];
        cum[r + 1] = cum[r] + c;
    }
    std::vector<int> cuts(1, 0);
    for (int k = 1; k < n; k++) {
        const double target = cum[ty] * k / n;
        int r = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        if (r > 0 && std::abs(cum[std::min(r, ty)] - target) > std::abs(cum[r - 1] - target)) r--;
        const int lo = cuts.back() + ((ty - cuts.back() > n - k) ? 1 : 0);
        cuts.push_back(std::min(std::max(r, lo), ty - (n - k)));
    }
    cuts.push_back(ty);
    for (int i = 0; i < n; i++) m->rows[i] = {std::max(cuts[i], 0), std::max(cuts[i + 1], cuts[i])};
}

static int multi_apply_rows(swr_multi *m) {
    for (int i = 0; i < m->ndev; i++) {
        int rc = swr_set_tile_rows(m->ctx[i], m->rows[i].first, m->rows[i].second);
        if (rc) {
            m->err = m->ctx[i]->err;
            return rc;
        }
    }
    return SWR_OK;
}

static thread_local std::string g_multi_create_error;

const char *swr_multi_last_error(const swr_multi *m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

void swr_multi_destroy(swr_multi *m) {
    if (!m) return;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->quit = true;
    }
    m->cv_go.notify_all();
    for (std::thread &t : m->workers)
        if (t.joinable()) t.join();
    for (swr_ctx *c : m->ctx) swr_destroy(c);
    delete m;
}

swr_multi *swr_multi_create(int width, int height, const int *devices, int ndev, int mode) {
    if (!devices || ndev < 1 || ndev > 64 || mode != SWR_MULTI_SORT_FIRST) {
        g_multi_create_error = "swr_multi_create: need 1..64 devices and mode SWR_MULTI_SORT_FIRST (sort-last runs one context per rank: swr_keys_to_global / swr_keys_localize)";
        return nullptr;
    }
    for (int i = 0; i < ndev; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) {
                g_multi_create_error = "swr_multi_create: a device is listed twice";
                return nullptr;
            }
    swr_multi *m = new swr_multi();
    m->W = width;
    m->H = height;
    m->ndev = ndev;
    m->mode = mode;
    for (int i = 0; i < ndev; i++) {
        swr_ctx *c = swr_create(width, height, devices[i]);
        if (!c) {
            g_multi_create_error = std::string("swr_multi_create: device ") + std::to_string(devices[i]) + ": " + swr_last_error(nullptr);
            swr_multi_destroy(m);
            return nullptr;
        }
        m->ctx.push_back(c);
    }
    m->tiles_x = m->ctx[0]->tiles_x;
    m->tiles_y = m->ctx[0]->tiles_y;
    // peer access towards the assembling device (device 0 of the list), and the assembly protocol's control words
    for (int i = 1; i < ndev; i++) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devices[i], devices[0]);
        if (!can) {
            g_multi_create_error = "swr_multi_create: device " + std::to_string(devices[i]) + " cannot access the memory of device " + std::to_string(devices[0]) + " (no NVLink / P2P)";
            swr_multi_destroy(m);
            return nullptr;
        }
        cudaSetDevice(devices[i]);
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
            g_multi_create_error = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
            swr_multi_destroy(m);
            return nullptr;
        }
        cudaGetLastError();
    }
    unsigned char handle[SWR_PEER_HANDLE_BYTES];
    if (ndev > 1) {
        if (swr_peer_export(m->ctx[0], handle) != SWR_OK) {
            g_multi_create_error = "swr_peer_export: " + m->ctx[0]->err;
            swr_multi_destroy(m);
            return nullptr;
        }
        for (int i = 1; i < ndev; i++) swr_peer_attach(m->ctx[i], m->ctx[0]->pixels.p);
    }
    // equal bands until a scene has been probed
    m->rows.resize(ndev);
    for (int i = 0, r = 0; i < ndev; i++) {
        const int n = m->tiles_y / ndev + (i < m->tiles_y % ndev ? 1 : 0);
        m->rows[i] = {r, r + n};
        r += n;
    }
    if (multi_apply_rows(m) != SWR_OK) {
        g_multi_create_error = m->err;
        swr_multi_destroy(m);
        return nullptr;
    }
    m->rc.assign(ndev, 0);
    for (int i = 0; i < ndev; i++) m->workers.emplace_back(multi_worker, m, i);
    return m;
}

int swr_multi_device_count(const swr_multi *m) { return m ? m->ndev : 0; }
swr_ctx *swr_multi_context(swr_multi *m, int i) { return (m && i >= 0 && i < m->ndev) ? m->ctx[i] : nullptr; }

int swr_multi_tile_rows(const swr_multi *m, int i, int *row_begin, int *row_end) {
    if (!m || i < 0 || i >= m->ndev || !row_begin || !row_end) return SWR_ERR_INVALID;
    *row_begin = m->rows[i].first;
    *row_end = m->rows[i].second;
    return SWR_OK;
}

int swr_multi_set_rsqrt_table(swr_multi *m, const uint32_t *table, int mantissa_bits) {
    if (!m) return SWR_ERR_INVALID;
    return multi_run(m, [=](int i) { return swr_set_rsqrt_table(m->ctx[i], table, mantissa_bits); });
}

int swr_multi_upload_scene(swr_multi *m, const swr_scene_desc *scene) {
    if (!m || !scene) return SWR_ERR_INVALID;
    m->balanced = false;
    m->have_scene = false;
    int rc = multi_run(m, [=](int i) { return swr_upload_scene(m->ctx[i], scene); });  // replicated on every device
    m->have_scene = rc == SWR_OK;
    return rc;
}

int swr_multi_render(swr_multi *m, const swr_camera *camera, const swr_draw *draws, int ndraws) {
    if (!m || !camera || ndraws < 0 || (ndraws > 0 && !draws)) return SWR_ERR_INVALID;
    if (!m->have_scene) {
        m->err = "swr_multi_render called before swr_multi_upload_scene";
        return SWR_ERR_NO_SCENE;
    }
    int rc;
    if (!m->balanced && m->ndev > 1) {
        // probe: two full-screen visibility passes on device 0 (the second one has its work units sized by the first),
        // then bands of equal measured cost
        swr_ctx *c0 = m->ctx[0];
        if ((rc = swr_set_tile_rows(c0, 0, m->tiles_y))) return rc;
        for (int k = 0; k < 2; k++)
            if ((rc = swr_render(c0, camera, draws, ndraws, 0))) {
                m->err = c0->err;
                return rc;
            }
        std::vector<uint32_t> cycles((size_t)m->tiles_x * m->tiles_y);
        if ((rc = swr_read_tile_costs(c0, nullptr, cycles.data()))) {
            m->err = c0->err;
            return rc;
        }
        multi_balance(m, cycles);
        if ((rc = multi_apply_rows(m))) return rc;
    }
    m->balanced = true;
    return multi_run(m, [=](int i) { return swr_render(m->ctx[i], camera, draws, ndraws, 1); });
}

int swr_multi_resolve(swr_multi *m, float exposure, uint32_t *out_pixels) {
    if (!m) return SWR_ERR_INVALID;
    if (m->ndev == 1) return swr_resolve(m->ctx[0], exposure, out_pixels);
    const uint32_t f = ++m->frame;
    int rc = multi_run(m, [=](int i) {
        // When the frame is about to be handed to the host, every device first makes sure ITS part is good (a buffer-growth
        // replay re-renders it): nobody else would look at a contributor's counters while the assembler waits for it.
        // Without host output the contributors only guard themselves on the device (k_resolve_peer stays silent on a
        // frame that overflowed) and the next call that settles the frame replays it.
        if (out_pixels) {
            cudaSetDevice(m->ctx[i]->device);
            while (m->ctx[i]->slots_pending > 0) {
                int rs = settle_oldest(m->ctx[i]);
                if (rs) return rs;
            }
        }
        if (i > 0) return swr_resolve_peer(m->ctx[i], exposure, f);
        int r = swr_resolve(m->ctx[0], exposure, nullptr);
        return r ? r : swr_peer_collect(m->ctx[0], f, m->ndev - 1);
    });
    if (rc) return rc;
    swr_ctx *c0 = m->ctx[0];
    if (out_pixels) {
        cudaSetDevice(c0->device);
        cudaError_t e = cudaMemcpyAsync(out_pixels, c0->pixels.p, (size_t)m->W * m->H * 4, cudaMemcpyDeviceToHost, c0->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c0->stream);
        if (e != cudaSuccess) {
            m->err = std::string("frame read-back: ") + cudaGetErrorString(e);
            return SWR_ERR_CUDA;
        }
    }
    if ((rc = swr_peer_release(c0, f))) {
        m->err = c0->err;
        return rc;
    }
    if (out_pixels) {
        // a protocol time-out on any device surfaces here instead of as a wrong image
        for (int i = 0; i < m->ndev; i++)
            if ((rc = swr_synchronize(m->ctx[i]))) {
                m->err = "device " + std::to_string(m->ctx[i]->device) + ": " + m->ctx[i]->err;
                return rc;
            }
    }
    return SWR_OK;
}

int swr_multi_synchronize(swr_multi *m) {
    if (!m) return SWR_ERR_INVALID;
    return multi_run(m, [=](int i) { return swr_synchronize(m->ctx[i]); });
}

int swr_multi_read_tile_luminance(swr_multi *m, float *out_per_tile) {
    if (!m || !out_per_tile) return SWR_ERR_INVALID;
    std::vector<float> tmp((size_t)m->tiles_x * m->tiles_y);
    for (int i = 0; i < m->ndev; i++) {  // every device metered the tiles of its own rows
        int rc = swr_read_tile_luminance(m->ctx[i], tmp.data());
        if (rc) {
            m->err = m->ctx[i]->err;
            return rc;
        }
        const size_t t0 = (size_t)m->rows[i].first * m->tiles_x, t1 = (size_t)m->rows[i].second * m->tiles_x;
        std::copy(tmp.begin() + t0, tmp.begin() + t1, out_per_tile + t0);
    }
    return SWR_OK;
}

int swr_multi_get_stats(swr_multi *m, swr_frame_stats *out) {
    if (!m || !out) return SWR_ERR_INVALID;
    swr_frame_stats acc{};
    for (int i = 0; i < m->ndev; i++) {
        swr_frame_stats st{};
        int rc = swr_get_stats(m->ctx[i], &st);
        if (rc) {
            m->err = m->ctx[i]->err;
            return rc;
        }
        if (i == 0) acc = st;
        else {
            // every device is handed the same draw list; binned / clipped / refs are per band (a triangle that straddles
            // a band boundary is binned by both neighbours), phase times are the slowest device's
            acc.triangles_binned += st.triangles_binned;
            acc.triangles_clipped += st.triangles_clipped;
            acc.tile_refs += st.tile_refs;
            acc.clusters_culled += st.clusters_culled;
            acc.ms_setup_bin = std::max(acc.ms_setup_bin, st.ms_setup_bin);
            acc.ms_raster = std::max(acc.ms_raster, st.ms_raster);
            acc.ms_shade = std::max(acc.ms_shade, st.ms_shade);
            acc.ms_resolve = std::max(acc.ms_resolve, st.ms_resolve);
        }
    }
    *out = acc;
    return SWR_OK;
}

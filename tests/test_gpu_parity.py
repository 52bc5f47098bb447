"""GPU parity tests proper: the CUDA path (through the host mirror and the C ABI) against the CPU oracle.
Bar: visibility buffer (depth bits, seq, barycentric bits) bit-exact; RGBA8 within +-1 LSB per channel."""
import numpy as np
import pytest

import swraster_viewer_b200 as swr
from helpers import small_configs, render_gpu, render_oracle, rgba_bytes, identity_camera, triangle_scene

pytestmark = pytest.mark.gpu

RGBA_TOL_LSB = 1  # north star: +-1 LSB per channel, max and mean stated


@pytest.fixture(scope="module")
def configs():
    return small_configs()


@pytest.mark.parametrize("idx", range(7))
def test_visbuffer_bit_exact_and_colour(configs, idx):
    name, scene, spec, W, H = configs[idx]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    g = render_gpu(scene, cam, W, H)
    o = render_oracle(scene, cam, W, H)
    assert np.array_equal(g["seq"], o["seq"]), f"{name}: {np.count_nonzero(g['seq'] != o['seq'])} pixels with a different triangle id"
    assert np.array_equal(g["depth"], o["depth"]), f"{name}: depth bits differ"
    assert np.array_equal(g["bary1"].view(np.uint32), o["bary1"].view(np.uint32))
    assert np.array_equal(g["bary2"].view(np.uint32), o["bary2"].view(np.uint32))
    assert (g["seq"] != 0xFFFFFFFF).any(), "scene must cover something"
    # counters equal the oracle's
    for k in ("triangles_submitted", "vertices_submitted", "triangles_binned", "triangles_clipped", "tile_refs"):
        assert g["stats"][k] == o["stats"][k], (name, k, g["stats"][k], o["stats"][k])
    err = np.abs(rgba_bytes(g["pixels"]) - rgba_bytes(o["pixels"]))
    print(f"{name}: RGBA8 max err {err.max()} LSB, mean {err.mean():.6f} LSB, {np.count_nonzero(err)} of {err.size} channel values differ "
          f"(host rsqrt table bits = {g['rsqrt_bits']})")
    assert err.max() <= RGBA_TOL_LSB
    assert np.all((g["pixels"] & 0xFF) == 0xFF)
    # linear HDR metering value: the shader contracts value-domain a*b+c into FMAs (swr_shade.cuh), so a few ulp, not bits
    assert np.allclose(g["luminance"], o["luminance"], rtol=2e-5, atol=1e-7), "tile metering luminance differs"


def test_colour_against_exact_rsqrt_oracle(configs):
    """With the oracle's normalize() switched to exact 1/sqrt the only remaining difference is rsqrtf's
    <= 2 ulp: linear colour must agree to ~1e-5 relative, i.e. the shading logic is the same."""
    name, scene, spec, W, H = configs[5]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    g = render_gpu(scene, cam, W, H, reference_rsqrt=False)
    o = render_oracle(scene, cam, W, H, exact_rsqrt=True)
    close = np.isclose(g["color"], o["color"], rtol=2e-4, atol=2e-5)
    # the few outliers are nearest-texel / voxel-cell flips caused by the last-bit difference rsqrtf vs 1/sqrtf
    assert close.mean() > 0.999, f"only {close.mean():.5f} of the linear colour values agree"
    err = np.abs(rgba_bytes(g["pixels"]) - rgba_bytes(o["pixels"]))
    print(f"{name} (rsqrtf vs exact-rsqrt oracle): RGBA8 max err {err.max()} mean {err.mean():.6f}")
    assert err.mean() < 0.01


def test_sort_first_rows_equal_full_frame(configs):
    """Tile-row partitions rendered separately reproduce the full frame exactly (SURVEY 8e sort-first)."""
    name, scene, spec, W, H = configs[2]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    full = render_gpu(scene, cam, W, H)
    tiles_y = (H + 63) // 64
    cut = tiles_y // 2
    top = render_gpu(scene, cam, W, H, rows=(0, cut))
    bot = render_gpu(scene, cam, W, H, rows=(cut, tiles_y))
    ysplit = cut * 64
    for k in ("seq", "depth", "pixels"):
        a = full[k].reshape(H, W)
        assert np.array_equal(a[:ysplit], top[k].reshape(H, W)[:ysplit]), k
        assert np.array_equal(a[ysplit:], bot[k].reshape(H, W)[ysplit:]), k
    assert top["stats"]["tile_refs"] + bot["stats"]["tile_refs"] == full["stats"]["tile_refs"]


def test_repeatable_and_reusable_renderer(configs):
    """Same renderer, several frames: identical output every frame (atomics must not leak nondeterminism)."""
    name, scene, spec, W, H = configs[3]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    r = swr.Renderer(W, H)
    outs = []
    for _ in range(3):
        r.render_scene(scene, cam)
        buf = swr.RenderBuffer(W, H)
        r.blit_to_buffer(buf)
        outs.append((r.read_visbuffer(), buf.pixels.copy()))
    for (v, p) in outs[1:]:
        for a, b in zip(v, outs[0][0]):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert np.array_equal(p, outs[0][1])
    r.close()


def test_empty_and_fully_culled():
    W, H = 128, 64
    # a single back-facing triangle: nothing binned, sky only
    sc = triangle_scene([[(100, 100), (900, 100), (100, 900)]], W, H)
    cam = identity_camera(W, H)
    g = render_gpu(sc, cam, W, H)
    o = render_oracle(sc, cam, W, H)
    assert g["stats"]["triangles_binned"] == 0 == o["stats"]["triangles_binned"]
    assert np.all(g["seq"] == 0xFFFFFFFF) and np.array_equal(g["depth"], o["depth"])
    assert np.abs(rgba_bytes(g["pixels"]) - rgba_bytes(o["pixels"])).max() <= 1


def test_exposure_and_auto_exposure_state(configs):
    name, scene, spec, W, H = configs[0]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    r = swr.Renderer(W, H)
    r.render_scene(scene, cam)
    assert r.auto_exposure == pytest.approx(2.0)
    r.update_auto_exposure(0.0)  # dt = 0 -> alpha = 0 -> exposure unchanged (renderer.rs:287-289)
    assert r.auto_exposure == pytest.approx(2.0)
    r.update_auto_exposure(0.5)
    assert 0.05 <= r.auto_exposure <= 32.0 and r.auto_exposure != pytest.approx(2.0)
    r.close()
    # R14 end to end: a 5-frame sequence (render, meter, blit) against the oracle's update_auto_exposure restatement
    # (renderer.rs:258-290) fed with the device's per-tile metering values — bit for bit — and against the oracle's own
    # frame (its metering values, its resolve at the metered exposure): RGBA8 within 1 LSB.
    import oracle as orc
    r = swr.Renderer(W, H)
    o = orc.Oracle(W, H)
    state = np.array([2.0, 2.0, 1.0], np.float32)
    ostate = state.copy()
    buf = swr.RenderBuffer(W, H)
    for f, dt in enumerate([0.0, 1 / 60, 0.25, 1.0, 1 / 60]):
        r.render_scene(scene, cam)
        lum = r.read_tile_luminance()
        state = orc.update_auto_exposure(state, lum, dt)
        r.update_auto_exposure(dt)
        assert np.float32(r.auto_exposure).view(np.uint32) == state[0].view(np.uint32), (f, r.auto_exposure, state)
        r.blit_to_buffer(buf)
        ref = o.render(scene, cam.abi, nthreads=1)
        ostate = orc.update_auto_exposure(ostate, ref["luminance"], dt)
        assert ostate[0] == pytest.approx(state[0], rel=1e-5)
        opix = o.resolve(float(ostate[0]))
        assert np.abs(rgba_bytes(buf.pixels) - rgba_bytes(opix)).max() <= 1, f
    assert state[0] != np.float32(2.0)
    r.close()


def test_create_rejects_bad_arguments():
    core, _ = swr.load_libraries()
    assert core.swr_create(0, 64, 0) is None
    assert core.swr_create(65, 64, 0) is None  # odd width
    assert b"even" in core.swr_last_error(None)
    assert core.swr_create(64, 64, 99) is None


def alpha_scene():
    """Alpha-tested (glTF MASK) terrain with holes over two opaque objects: exercises shader.rs:40-43 / 311-329."""
    import math
    from swraster_viewer_b200 import abi, scenes
    tex = scenes.checker_noise_texture(64, 21, abi.TEX_SRGB, abi.WRAP_REPEAT, alpha_holes=True)
    mats = [scenes.Material((1, 1, 1, 1), 0.0, 0.7, flags=abi.MAT_ALPHA_TESTED, base_color_texture=0, alpha_cutoff=0.5),
            scenes.Material((0.9, 0.3, 0.2, 1), 0.0, 0.4)]
    meshes = [[scenes.height_field(40, 5, extent=4.0, height=0.6, uv_repeat=2.0, material=0)], [scenes.uv_sphere(24, 16, 1.0, 1)]]
    nodes = [scenes.Node(scenes.IDENT, 0), scenes.Node(scenes.trs((0.3, -0.9, 0.2), 1.2), 1), scenes.Node(scenes.trs((-1.8, 1.2, 0.8), 0.6), 1),
             scenes.Node(scenes.trs((0.0, 1.5, 5.5), 2.5, (1, 0, 0), 1.2), 0)]  # a second alpha-tested sheet crossing the near plane
    sc = scenes.SceneData(meshes, nodes, mats, [tex], voxel_dim=8, cube_size=32, seed=9)
    cam = scenes.CameraSpec((0.4, 3.2, 6.0), (0.0, 0.2, 0.0), math.pi / 4, float(sc.bounds_diagonal) * 2.0)
    return sc, cam


def test_alpha_tested_material():
    sc, spec = alpha_scene()
    W, H = 448, 256
    cam = swr.RenderCamera.from_spec(spec, W, H)
    g = render_gpu(sc, cam, W, H)
    o = render_oracle(sc, cam, W, H)
    assert np.array_equal(g["seq"], o["seq"]), f"{np.count_nonzero(g['seq'] != o['seq'])} pixels differ"
    assert np.array_equal(g["depth"], o["depth"])
    assert np.array_equal(g["bary1"].view(np.uint32), o["bary1"].view(np.uint32))
    # holes must actually show what is behind: some pixels inside the terrain's footprint belong to the sphere draws
    ids = o["seq"][o["seq"] != 0xFFFFFFFF] >> 3
    assert len(np.unique(ids)) > 500
    err = np.abs(rgba_bytes(g["pixels"]) - rgba_bytes(o["pixels"]))
    assert err.max() <= RGBA_TOL_LSB
    for k in ("triangles_binned", "triangles_clipped", "tile_refs"):
        assert g["stats"][k] == o["stats"][k]


def test_async_blit_matches_sync(configs):
    """swr_resolve_async / swr_wait_pixels (pipelined read-back into pinned memory) delivers the same pixels as blit_to_buffer."""
    name, scene, spec, W, H = configs[5]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    r = swr.Renderer(W, H)
    ref = swr.RenderBuffer(W, H)
    r.render_scene(scene, cam)
    r.blit_to_buffer(ref)
    bufs = [swr.RenderBuffer(W, H, pinned=True), swr.RenderBuffer(W, H, pinned=True)]
    prev = None
    for i in range(5):
        r.render_scene(scene, cam)
        tk = r.blit_to_buffer_async(bufs[i & 1])
        if prev is not None:
            r.wait_blit(prev[0])
            assert np.array_equal(prev[1].pixels, ref.pixels)
            prev[1].pixels[:] = 0
        prev = (tk, bufs[i & 1])
    r.wait_blit(prev[0])
    assert np.array_equal(prev[1].pixels, ref.pixels)
    r.close()


def _device_pixels_host(r, W, H):
    """Copy the renderer's device pixel buffer to the host on the renderer's own stream."""
    import torch
    from swraster_viewer_b200.multigpu import device_tensor
    stream = torch.cuda.ExternalStream(r.cuda_stream(), device=0)
    with torch.cuda.stream(stream):
        out = device_tensor(r.device_pixels_ptr(), W * H * 4, torch.int32, "cuda:0").cpu()
    stream.synchronize()
    return out.numpy().view(np.uint32)


def test_peer_assembly_equals_full_frame(configs):
    """Sort-first assembly through peer stores (swr_peer_*): a contributing context resolves its rows straight into the
    assembling context's pixel buffer; the assembled image equals the single-context frame, frame after frame (the
    free/done handshake is exercised three times). Two contexts on one device stand in for two ranks."""
    name, scene, spec, W, H = configs[2]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    full = render_gpu(scene, cam, W, H)["pixels"]
    tiles_y = (H + 63) // 64
    cut = tiles_y // 2
    a, b = swr.Renderer(W, H), swr.Renderer(W, H)
    a.set_tile_rows(0, cut)
    b.set_tile_rows(cut, tiles_y)
    assert len(a.peer_export()) == 64
    b.peer_attach(a.device_pixels_ptr())
    for f in (1, 2, 3):
        b.render_scene(scene, cam)
        b.resolve_peer(2.0, f)
        a.render_scene(scene, cam)
        a.resolve_device_only(2.0)
        a.peer_collect(f, 1)
        got = _device_pixels_host(a, W, H)
        a.peer_release(f)
        a.synchronize()
        b.synchronize()
        assert np.array_equal(got, full.reshape(-1)), f"frame {f}"
    b.close()
    a.close()


def test_peer_wait_times_out_instead_of_hanging(configs):
    """A contributor that runs ahead of the assembler (frame number never released) gives up after ~2 s and reports it."""
    name, scene, spec, W, H = configs[0]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    a, b = swr.Renderer(W, H), swr.Renderer(W, H)
    a.peer_export()
    b.peer_attach(a.device_pixels_ptr())
    b.render_scene(scene, cam)
    b.resolve_peer(2.0, 7)  # only frame 1 has been released
    with pytest.raises(Exception, match="timed out"):
        b.synchronize()
    b.synchronize()  # the flag is cleared: the context stays usable
    with pytest.raises(Exception, match="swr_peer_open"):
        a.resolve_peer(2.0, 1)  # the assembler does not contribute to itself
    b.close()
    a.close()


def cluster_scene():
    """Large meshes around and behind the camera under rotated, non-uniformly scaled and sheared transforms: most
    128-triangle clusters are outside the frustum, some straddle it, and the band split cuts through the rest."""
    import math
    from swraster_viewer_b200 import scenes
    mats = [scenes.Material((0.8, 0.7, 0.3, 1), 0.1, 0.6), scenes.Material((0.2, 0.5, 0.9, 1), 0.0, 0.4)]
    meshes = [[scenes.height_field(160, 11, extent=30.0, height=1.5, uv_repeat=4.0, material=0)], [scenes.uv_sphere(96, 64, 1.0, 1)]]

    def stretched(translate, sx, sy, sz, axis, angle, shear=0.0):
        m = scenes.trs(translate, 1.0, axis, angle).reshape(4, 4).T.astype(np.float64)  # row-major 4x4
        s = np.diag([sx, sy, sz, 1.0])
        s[0, 1] = shear
        return np.ascontiguousarray((m @ s).T.reshape(-1).astype(np.float32))

    def needle(eye, target, dist, thin, long):
        """Unit sphere stretched to `long` along the camera's up axis (and `thin` across), centred `dist` in front of the eye."""
        f = np.array(target, np.float64) - np.array(eye, np.float64)
        f /= np.linalg.norm(f)
        r = np.cross(f, [0.0, 1.0, 0.0])
        r /= np.linalg.norm(r)
        u = np.cross(r, f)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = r * thin, u * long, f * thin, np.array(eye) + f * dist
        return np.ascontiguousarray(m.T.reshape(-1).astype(np.float32))

    eye, target = (0.0, 1.0, 4.0), (0.5, 0.2, -4.0)
    nodes = [scenes.Node(stretched((0, -1.0, 0), 1.0, 1.0, 1.0, (0, 1, 0), 0.3), 0),               # terrain all around (and behind) the eye
             scenes.Node(stretched((2.0, 1.0, -6.0), 3.0, 0.4, 1.5, (1, 1, 0), 0.8, 0.5), 1),      # sheared ellipsoid across the frustum edge
             scenes.Node(stretched((-14.0, 2.0, -3.0), 6.0, 6.0, 0.3, (0, 0, 1), 1.1), 1),         # flat disc mostly outside to the left
             scenes.Node(stretched((0.0, 0.5, 9.0), 2.0, 2.0, 2.0, (0, 1, 0), 0.0), 1),            # behind the camera
             scenes.Node(stretched((0.3, 0.4, 0.5), 0.3, 0.3, 0.3, (0, 1, 0), 0.0), 1),            # small, really inside: drawn unclipped
             # a vertical needle: the reference sizes the node sphere with the MEAN axis scale (scene.rs:53-63), so this is
             # classified Inside and rasterised UNCLIPPED although it sticks out of the top and bottom of the frustum;
             # cluster culling may only band-cull such draws (and only in front of the eye), never frustum-cull them
             scenes.Node(needle(eye, target, 12.0, 0.2, 6.6), 1)]
    sc = scenes.SceneData(meshes, nodes, mats, [], voxel_dim=8, cube_size=32, seed=13)
    cam = scenes.CameraSpec(eye, target, math.pi / 4, float(sc.bounds_diagonal) * 2.0)
    return sc, cam


def test_cluster_culling_keeps_the_image(configs):
    """k_cull may only drop clusters that contribute nothing: full frame and every row band equal the oracle, while a
    large share of the clusters is actually rejected (frustum planes on the full frame, plus the band planes)."""
    sc, spec = cluster_scene()
    W, H = 512, 320
    cam = swr.RenderCamera.from_spec(spec, W, H)
    o = render_oracle(sc, cam, W, H)
    g = render_gpu(sc, cam, W, H)
    assert g["stats"]["clusters_culled"] > 50, g["stats"]
    for k in ("seq", "depth"):
        assert np.array_equal(g[k], o[k]), k
    assert np.array_equal(g["bary1"].view(np.uint32), o["bary1"].view(np.uint32))
    for k in ("triangles_binned", "triangles_clipped", "tile_refs"):
        assert g["stats"][k] == o["stats"][k], k
    assert np.abs(rgba_bytes(g["pixels"]) - rgba_bytes(o["pixels"])).max() <= RGBA_TOL_LSB
    tiles_y = (H + 63) // 64
    culled = []
    for row in range(tiles_y):
        b = render_gpu(sc, cam, W, H, rows=(row, row + 1))
        y0, y1 = row * 64, min(row * 64 + 64, H)
        for k in ("seq", "depth", "pixels"):
            assert np.array_equal(b[k].reshape(H, W)[y0:y1], g[k].reshape(H, W)[y0:y1]), (row, k)
        culled.append(b["stats"]["clusters_culled"])
    assert min(culled) > g["stats"]["clusters_culled"], (culled, g["stats"]["clusters_culled"])


def test_gltf_loaded_scene_matches_the_oracle(tmp_path):
    """SURVEY 8f N2 through the product path: a scene written as glTF and read back by the native loader (host/swr_gltf.hpp)
    is uploaded and rendered by the CUDA path exactly as the oracle renders the same loaded scene."""
    from swraster_viewer_b200 import gltf, scenes
    from helpers import SMALL
    sc, spec = scenes.scene_materials_test(**SMALL)
    scenes.export_gltf(sc, str(tmp_path / "scene"))
    g = gltf.load_gltf(tmp_path / "scene.gltf", environment=sc)
    W, H = 320, 192
    cam = swr.RenderCamera.from_spec(spec, W, H)
    a, b = render_gpu(g, cam, W, H), render_oracle(g, cam, W, H)
    assert np.array_equal(a["seq"], b["seq"]) and np.array_equal(a["depth"], b["depth"])
    assert np.abs(rgba_bytes(a["pixels"]) - rgba_bytes(b["pixels"])).max() <= RGBA_TOL_LSB
    assert (a["seq"] != 0xFFFFFFFF).mean() > 0.2

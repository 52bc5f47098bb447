"""Host-side logic above the C ABI (no GPU): draw-list construction against the oracle's own restatement of
renderer.rs:357-468, camera construction, rsqrt-table probing, and the sort-first partition maths over gloo."""
import ctypes as C
import math
import os
import socket

import numpy as np
import pytest

import swraster_viewer_b200 as swr
from swraster_viewer_b200 import abi, scenes
from swraster_viewer_b200.renderer import build_draws
from swraster_viewer_b200.multigpu import tile_row_ranges, balanced_row_ranges, rebalance_row_ranges
from helpers import small_configs, SMALL


def test_draw_list_matches_oracle_bitwise():
    import oracle as orc
    for name, scene, spec, W, H in small_configs():
        cam = swr.RenderCamera.from_spec(spec, W, H)
        mine, n = build_draws(scene, cam)
        o = orc.Oracle(64, 64)
        ref, m = o.build_draws(scene, cam.abi, abi.Draw)
        assert n == m and n > 0, name
        assert bytes(mine)[: n * C.sizeof(abi.Draw)] == bytes(ref)[: n * C.sizeof(abi.Draw)], name


def test_draw_list_sharding_keeps_global_ids():
    _, scene, spec, W, H = small_configs()[2]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    full, n = build_draws(scene, cam)
    got = []
    for s in range(3):
        part, m = build_draws(scene, cam, shard=s, nshards=3)
        got += [(part[i].first_triangle, part[i].primitive, part[i].flags) for i in range(m)]
    assert sorted(got) == sorted((full[i].first_triangle, full[i].primitive, full[i].flags) for i in range(n))


def test_camera_level_matches_reference_construction():
    """RenderCamera::new towards a level target: view matrix puts the target on -z, projection is glam's perspective_rh."""
    cam = swr.RenderCamera((1.0, 2.0, 9.0), (1.0, 2.0, 0.0), math.pi / 4, 640, 360, 50.0)
    V = np.array(cam.abi.view_matrix).reshape(4, 4).T
    VP = np.array(cam.abi.view_project_matrix).reshape(4, 4).T
    t = V @ np.array([1.0, 2.0, 0.0, 1.0])
    np.testing.assert_allclose(t[:3], [0, 0, -9], atol=1e-5)
    near, far = 0.5, 50.0
    c_near = VP @ np.array([1.0, 2.0, 9.0 - near, 1.0])
    c_far = VP @ np.array([1.0, 2.0, 9.0 - far, 1.0])
    assert abs(c_near[2] / c_near[3]) < 1e-5 and abs(c_far[2] / c_far[3] - 1.0) < 1e-5  # depth 0..1
    planes = np.array(cam.abi.view_clip_planes)
    np.testing.assert_allclose(planes[0], [0, 0, -1, near], atol=1e-6)
    np.testing.assert_allclose(planes[1], [0, 0, 1, far], atol=1e-6)
    assert cam.abi.one_over_width == pytest.approx(1 / 640)


def test_from_spec_aims_at_target():
    spec = scenes.CameraSpec((0.5, 3.0, 6.5), (0.0, 0.8, 0.0), math.pi / 4, 30.0)
    cam = swr.RenderCamera.from_spec(spec, 640, 360)
    VP = np.array(cam.abi.view_project_matrix).reshape(4, 4).T
    c = VP @ np.array([0.0, 0.8, 0.0, 1.0])
    assert abs(c[0] / c[3]) < 1e-4 and abs(c[1] / c[3]) < 1e-4


def test_tile_row_ranges_partition():
    for tiles_y in (17, 34, 68, 3):
        for n in (1, 2, 4, 8):
            r = tile_row_ranges(tiles_y, n)
            assert len(r) == n and r[0][0] == 0 and r[-1][1] == tiles_y
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_row_ranges_partition_and_balance():
    rng = np.random.default_rng(3)
    for tiles_y in (17, 34, 68):
        counts = (rng.random((tiles_y, 60)) ** 4 * 30000).astype(np.uint32)
        counts[tiles_y // 3] *= 6  # one very heavy row
        cost = counts.sum(1) + (0.9 * counts.mean() + 1.0) * 60
        for world in (1, 2, 4, 8):
            r = balanced_row_ranges(counts, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == tiles_y
            assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(b > a for a, b in r)
            band = [cost[a:b].sum() for a, b in r]
            even = [cost[a:b].sum() for a, b in tile_row_ranges(tiles_y, world)]
            assert max(band) <= max(even) * 1.001  # never worse than the even split


def test_rebalance_row_ranges_converges_on_the_measured_time():
    """Feedback step of the sort-first split: bands are re-cut from what they really took. Model: a band's time is its
    raster cost plus a per-row term the probe does not know (set-up + shading); a few steps level the bands."""
    rng = np.random.default_rng(5)
    tiles_y = 34
    raster = rng.random(tiles_y) ** 3 * 100 + 5
    hidden = np.linspace(30, 5, tiles_y)  # what the probe cannot see: heavier at the top of the screen
    true_ms = raster + hidden
    for world in (2, 4, 8):
        ranges = tile_row_ranges(tiles_y, world)
        first = max(true_ms[a:b].sum() for a, b in ranges)
        for _ in range(3):
            band = [true_ms[a:b].sum() for a, b in ranges]
            ranges = rebalance_row_ranges(ranges, band, raster)
            assert ranges[0][0] == 0 and ranges[-1][1] == tiles_y
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:])) and all(b > a for a, b in ranges)
        worst = max(true_ms[a:b].sum() for a, b in ranges)
        assert worst <= first * 1.001, (world, worst, first)
        assert worst <= true_ms.sum() / world + true_ms.max()  # within one row of the ideal split


def test_band_culling_is_conservative():
    """Draws dropped for a row band must have no triangle bbox inside the band: the union of the per-band lists is the
    full list and every draw kept by the oracle-equivalent full list appears in each band it can touch."""
    import ctypes as C
    _, host = swr.load_libraries()
    host.swrh_build_draws_band = getattr(host, "swrh_build_draws_band")
    name, scene, spec, W, H = small_configs()[2]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    full, n = build_draws(scene, cam)
    keys_full = {(full[i].first_triangle, full[i].primitive) for i in range(n)}
    tiles_y = (H + 63) // 64
    seen = set()
    for r0, r1 in tile_row_ranges(tiles_y, 3):
        arr = (abi.Draw * n)()
        m = host.swrh_build_draws_band(C.byref(scene.desc()), C.byref(cam.abi), arr, n, r0 * 64, r1 * 64, H)
        assert 0 < m <= n
        band = {(arr[i].first_triangle, arr[i].primitive) for i in range(m)}
        assert band <= keys_full
        seen |= band
    assert len(seen) > 0.5 * len(keys_full)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, W, H, tmp):
    import torch
    import torch.distributed as dist
    from swraster_viewer_b200.multigpu import tile_row_ranges, gather_strips
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tiles_y = (H + 63) // 64
    ranges = tile_row_ranges(tiles_y, world)
    full = torch.arange(W * H, dtype=torch.int32).reshape(H, W)
    mine = torch.zeros(H, W, dtype=torch.int32)
    y0, y1 = ranges[rank][0] * 64, min(ranges[rank][1] * 64, H)
    mine[y0:y1] = full[y0:y1]  # this rank "rendered" only its rows
    out = gather_strips(mine, ranges, H, dst=0)
    if rank == 0:
        assert torch.equal(out, full)
        open(os.path.join(tmp, "ok"), "w").write("1")
    dist.destroy_process_group()


def test_sort_first_gather_gloo_world2(tmp_path):
    """The N>1 data path (row strips -> rank 0) on CPU with gloo, world_size 2."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_gather_worker, args=(2, port, 192, 200, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def _composite_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    from swraster_viewer_b200.multigpu import composite_keys_min
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    allk = torch.randint(0, 2 ** 62, (world, 4096 * 6), generator=g, dtype=torch.int64)
    allk[:, ::7] = -1  # empty key = all ones
    mine = allk[rank].clone()
    out = composite_keys_min(mine)
    # unsigned minimum across ranks
    u = allk.numpy().view(np.uint64)
    assert np.array_equal(out.numpy().view(np.uint64), u.min(axis=0))
    if rank == 0:
        open(os.path.join(tmp, "ok"), "w").write("1")
    dist.destroy_process_group()


def test_sort_last_key_composite_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_composite_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_scene_ranges_are_validated_before_they_are_indexed():
    """node -> mesh -> primitive -> material indices are only read on the host: a bad one must raise, not index out of bounds
    (the reference would panic on the slice access)."""
    import copy
    import ctypes as C
    from swraster_viewer_b200.renderer import build_draws
    sc, spec = scenes.scene_c1_sphere(16, 8, **SMALL)
    cam = swr.RenderCamera.from_spec(spec, 128, 64)
    good = sc.desc()
    assert build_draws(sc, cam)[1] == 1

    class Wrapped:
        def __init__(self, d):
            self.d = d

        def desc(self):
            return self.d

    def variant(edit):
        d = abi.SceneDesc.from_buffer_copy(good)
        nodes = (abi.NodeDesc * d.nnodes).from_buffer_copy((abi.NodeDesc * d.nnodes).from_address(C.addressof(d.nodes.contents)))
        meshes = (abi.MeshDesc * d.nmeshes).from_buffer_copy((abi.MeshDesc * d.nmeshes).from_address(C.addressof(d.meshes.contents)))
        prims = (abi.PrimitiveDesc * d.nprimitives).from_buffer_copy((abi.PrimitiveDesc * d.nprimitives).from_address(C.addressof(d.primitives.contents)))
        edit(nodes, meshes, prims)
        d.nodes, d.meshes, d.primitives = nodes, meshes, prims
        w = Wrapped(d)
        w.keep = (nodes, meshes, prims)
        return w

    for edit, msg in [(lambda n, m, p: setattr(n[0], "mesh_index", 7), "mesh that does not exist"),
                      (lambda n, m, p: setattr(m[0], "num_primitives", 5), "beyond the primitive array"),
                      (lambda n, m, p: setattr(p[0], "material_index", 3), "material that does not exist")]:
        with pytest.raises(RuntimeError, match=msg):
            build_draws(variant(edit), cam)
    assert build_draws(variant(lambda n, m, p: setattr(n[0], "mesh_index", -1)), cam)[1] == 0  # a node without a mesh draws nothing


def test_bench_camera_fixture_matches_the_host_mirror():
    """bench.py --impl reference takes its swr_camera blocks from tests/golden/bench_cameras.json (so the CPU arm never maps
    the product libraries); the host mirror must still reproduce every one of them bit for bit from the recorded spec."""
    import json
    import os
    from swraster_viewer_b200 import scenes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fx = json.load(open(os.path.join(root, "tests", "golden", "bench_cameras.json")))
    assert set(fx) == {"c1", "c2", "c3", "c4", "c5"}
    for name, e in fx.items():
        spec = scenes.CameraSpec(tuple(e["position"]), tuple(e["look_at"]), e["fov"], e["far_plane"])
        cam = swr.RenderCamera.from_spec(spec, e["width"], e["height"])
        assert bytes(cam.abi).hex() == e["camera_hex"], name


def test_reference_arm_does_not_map_the_product_libraries():
    """the CPU arm's process must not load libswr_b200.so / libswr_host.so (only the checker)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, os; sys.argv=['bench.py','--impl','reference','--config','c1','--steps','1','--warmup','0'];"
            "sys.path.insert(0, %r); import bench; bench.main();"
            "maps=open('/proc/self/maps').read(); print('MAPPED', 'libswr_b200' in maps or 'libswr_host' in maps, 'liboracle' in maps)") % root
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "MAPPED False True" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    assert '"impl": "reference"' in out.stdout


def test_auto_exposure_matches_the_oracle_restatement_over_a_frame_sequence():
    """R14: update_auto_exposure (renderer.rs:258-290). The host mirror's metering maths against the oracle's restatement,
    bit for bit, over 5-frame sequences of per-tile luminances (state carried from frame to frame), for several tile
    counts (trim 10 %, min((n-1)/2)) and time steps, including dt = 0 (exposure must stay exactly 2.0) and odd values."""
    import oracle as orc
    _, host = swr.load_libraries()
    rng = np.random.default_rng(1234)
    for ntiles in (1, 2, 3, 9, 510, 2040, 8160):
        for dts in ([0.0] * 5, [1 / 60] * 5, [0.5, 0.0, 2.0, 1e-3, 10.0], [-1.0, 0.016, 0.016, 0.016, 0.016]):
            host_state = np.array([2.0, 2.0, 1.0], np.float32)
            orc_state = host_state.copy()
            for f, dt in enumerate(dts):
                lum = (rng.random(ntiles, dtype=np.float32) ** 3 * np.float32(8.0)).astype(np.float32)
                if f == 2:
                    lum[: max(1, ntiles // 7)] = 0.0  # unlit tiles: clamped to 1e-4 before the log
                if f == 3 and ntiles > 2:
                    lum[1] = np.float32("inf")
                    lum[2] = np.float32("nan")  # f32::max drops the NaN operand
                assert host.swrh_auto_exposure_step(host_state.ctypes.data, lum.ctypes.data, ntiles, float(dt)) == 0
                orc_state = orc.update_auto_exposure(orc_state, lum, dt)
                assert np.array_equal(host_state.view(np.uint32), orc_state.view(np.uint32)), (ntiles, dts, f, host_state, orc_state)
            if all(d == 0.0 for d in dts):
                assert host_state[0] == np.float32(2.0) and host_state[2] == np.float32(1.0)  # alpha = 1 - e^0 = 0
            assert 0.05 <= host_state[1] <= 32.0

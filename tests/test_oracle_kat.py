"""Known-answer tests that pin the ORACLE (CPU, no GPU): hand-made triangles on an identity camera against an
independent integer rasteriser written straight from SURVEY Appendix A (top-left rule, 28.4 snapping, `<=` ties).
The reference ships no tests or golden vectors, so these self-made vectors are the pin."""
import numpy as np
import pytest

from helpers import identity_camera, triangle_scene, render_oracle

W, H = 128, 64


def int_raster(tris, depths, W, H):
    """Independent model. tris: list of 3x(X,Y) 28.4 sub-pixel corners; depths: constant z per triangle.
    Returns (id map, depth map); id = triangle index, -1 = uncovered. Later triangle wins depth ties."""
    ids = -np.ones((H, W), np.int64)
    zbuf = np.full((H, W), np.inf, np.float32)
    py, px = np.mgrid[0:H, 0:W]
    cx, cy = px * 16 + 8, py * 16 + 8
    for t, ((x0, y0), (x1, y1), (x2, y2)) in enumerate(tris):
        area = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)
        if area > 0:  # renderer.rs:680
            continue
        cov = np.ones((H, W), bool)
        for (xa, ya), (xb, yb) in (((x0, y0), (x1, y1)), ((x1, y1), (x2, y2)), ((x2, y2), (x0, y0))):
            a, b = yb - ya, xa - xb
            bias = 0 if (a < 0 or (a == 0 and b > 0)) else -1  # tilerasterizer.rs:139-141
            c = xb * ya - xa * yb + bias
            cov &= (a * cx + b * cy + c) >= 0
        z = np.float32(depths[t])
        win = cov & (z <= zbuf)  # tilerasterizer.rs:516
        ids[win] = t
        zbuf[win] = z
    return ids, zbuf


def oracle_ids(tris, depths):
    sc = triangle_scene(tris, W, H, depths=[[d, d, d] for d in depths])
    o = render_oracle(sc, identity_camera(W, H), W, H)
    seq = o["seq"].astype(np.int64).reshape(H, W)
    ids = np.where(seq == 0xFFFFFFFF, -1, seq >> 3)
    return ids, o["depth"].view(np.float32).reshape(H, W), o


def check(tris, depths):
    want_ids, want_z = int_raster(tris, depths, W, H)
    got_ids, got_z, o = oracle_ids(tris, depths)
    assert np.array_equal(got_ids, want_ids), f"{np.count_nonzero(got_ids != want_ids)} pixels differ"
    assert np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))
    return o


def test_shared_diagonal_no_gap_no_overlap():
    # a quad split along its diagonal, front-facing (area <= 0 in y-down screen space)
    q = [(160, 96), (160, 800), (1500, 800), (1500, 96)]
    tris = [[q[0], q[1], q[2]], [q[0], q[2], q[3]]]
    o = check(tris, [0.5, 0.5])
    ids, _ = int_raster(tris, [0.5, 0.5], W, H)
    inside = np.zeros((H, W), bool)
    inside[6:50, 10:94] = True  # pixel centres strictly inside [160,1500)x[96,800): 16*px+8 in range
    assert np.array_equal(ids >= 0, inside)


def test_pixel_centre_on_edges_and_vertices_top_left_rule():
    # axis-aligned right triangle whose top and left edges pass exactly through pixel centres (x=168 -> px 10, y=104 -> py 6)
    tris = [[(168, 104), (168, 616), (680, 104)]]
    o = check(tris, [0.25])
    ids, _ = int_raster(tris, [0.25], W, H)
    # With y-down screen space and front faces at area <= 0, `a < 0 || (a == 0 && b > 0)` (tilerasterizer.rs:139-141)
    # makes upward edges (the hypotenuse here) and leftward horizontal edges (the top edge) inclusive, and downward
    # edges (the left edge) exclusive.
    assert ids[6, 10] == -1 and ids[20, 10] == -1   # pixel centres exactly on the left edge: excluded
    assert ids[6, 11] == 0                          # top edge: included
    assert ids[6, 42] == 0 and ids[6, 43] == -1     # top-right vertex lies on two inclusive edges
    assert ids[7, 41] == 0 and ids[7, 42] == -1     # centre exactly on the hypotenuse: included
    assert ids[38, 10] == -1                        # bottom vertex is on the exclusive left edge


def test_equal_depth_tie_later_triangle_wins():
    a = [(100, 100), (100, 900), (900, 100)]
    b = [(300, 60), (300, 700), (1200, 60)]
    o = check([a, b], [0.5, 0.5])
    ids, _ = int_raster([a, b], [0.5, 0.5], W, H)
    both = (int_raster([a], [0.5], W, H)[0] >= 0) & (int_raster([b], [0.5], W, H)[0] >= 0)
    assert both.any() and np.all(ids[both] == 1)


def test_nearer_first_submitted_still_wins():
    a = [(100, 100), (100, 900), (900, 100)]
    b = [(300, 60), (300, 700), (1200, 60)]
    check([a, b], [0.25, 0.75])
    check([a, b], [0.75, 0.25])


def test_backface_and_zero_area():
    front = [(100, 100), (100, 900), (900, 100)]
    back = [(100, 100), (900, 100), (100, 900)]
    degenerate = [(200, 200), (400, 400), (600, 600)]  # zero area is kept by the cull but covers nothing
    o = check([back, degenerate, front], [0.1, 0.1, 0.9])
    assert o["stats"]["triangles_binned"] == 2  # degenerate is binned (renderer.rs:680 keeps area == 0)


def test_fuzz_small_triangles_exact_regime():
    rng = np.random.default_rng(1234)
    tris, depths = [], []
    for _ in range(300):
        cx, cy = rng.integers(0, W * 16), rng.integers(0, H * 16)
        p = [(int(cx + rng.integers(-200, 200)), int(cy + rng.integers(-200, 200))) for _ in range(3)]
        tris.append(p)
        depths.append(float(rng.choice([0.125, 0.25, 0.5, 0.75])))
    check(tris, depths)


def test_right_and_bottom_screen_edges():
    # touches x = W*16 and y = H*16 exactly; starts in tile column 0 -> the reference's wrapped duplicate packet
    tris = [[(16, 16), (16, H * 16), (W * 16, 16)], [(W * 16, H * 16), (W * 16, 32), (32, H * 16)]]
    o = check(tris, [0.5, 0.6])
    assert o["stats"]["tile_refs"] >= 2


def test_clipped_triangle_covers_like_unclipped_on_screen():
    """A triangle hanging over the left screen border, run through the clipper (Intersecting): on-screen coverage must
    equal the unclipped integer model except possibly along edges re-snapped at the new vertices."""
    tri = [(-900, 500), (1200, 900), (1200, 200)]  # one vertex off-screen: the clipped polygon is a quad -> two fans
    sc = triangle_scene([tri], W, H, depths=[[0.5, 0.5, 0.5]])
    o = render_oracle(sc, identity_camera(W, H, intersecting=True), W, H)
    assert o["stats"]["triangles_clipped"] == 1
    got = (o["seq"].reshape(H, W) != 0xFFFFFFFF)
    want = int_raster([tri], [0.5], W, H)[0] >= 0
    assert want.sum() > 500
    assert np.count_nonzero(got != want) <= 8
    fans = np.unique(o["seq"][o["seq"] != 0xFFFFFFFF] & 7)
    assert len(fans) >= 2, "the clipped polygon must have been fanned into several triangles"


def test_parallel_schedule_same_coverage_and_depth():
    """The multi-threaded baseline schedule changes packet order only: depth is order-independent."""
    import oracle as orc
    import swraster_viewer_b200 as swr
    from helpers import small_configs
    name, scene, spec, w, h = small_configs()[2]
    cam = swr.RenderCamera.from_spec(spec, w, h)
    o = orc.Oracle(w, h)
    a = o.render(scene, cam.abi, nthreads=1)
    b = o.render(scene, cam.abi, nthreads=4)
    assert np.array_equal(a["depth"], b["depth"])
    same = a["seq"] == b["seq"]
    assert same.mean() > 0.999  # only exact depth ties may pick another packet


def test_nan_and_infinite_depths_are_never_written():
    """tilerasterizer.rs:516 `z <= current`: a NaN depth fails the comparison for every pixel (nothing is written, not even
    over the +INF clear value). A +INF vertex depth ends up the same way: the interpolator is a + b1*(z1-z0) + b2*(z2-z0)
    (util.rs:169-171) and INF - INF is NaN, so such a triangle never reaches the `INF <= INF` tie with the clear value."""
    a = [(160, 96), (160, 800), (1500, 800)]
    b = [(400, 200), (400, 900), (1800, 900)]
    c = [(96, 320), (96, 960), (900, 960)]
    tris = [a, b, c]
    want_ids, want_z = int_raster(tris, [np.nan, np.nan, 0.25], W, H)
    assert not (want_ids == 0).any() and not (want_ids == 1).any() and (want_ids == 2).any()
    with np.errstate(invalid="ignore"):
        got_ids, got_z, o = oracle_ids(tris, [np.nan, np.inf, 0.25])
    assert np.array_equal(got_ids, want_ids)
    assert np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))
    assert (got_z[got_ids == 2] == np.float32(0.25)).all() and np.isinf(got_z[got_ids == -1]).all()
    assert o["stats"]["triangles_binned"] == 3  # they are binned like any other triangle; only the depth test rejects them
    # a NaN in ONE vertex poisons the interpolated depth of the whole triangle; a huge finite depth is an ordinary depth
    sc = triangle_scene([a, c, b], W, H, depths=[[0.5, np.nan, 0.5], [0.75, 0.75, 0.75], [3.0e38, 3.0e38, 3.0e38]])
    with np.errstate(invalid="ignore"):
        o = render_oracle(sc, identity_camera(W, H), W, H)
    seq = o["seq"].astype(np.int64).reshape(H, W)
    covered = seq != 0xFFFFFFFF
    assert not (covered & ((seq >> 3) == 0)).any() and (covered & ((seq >> 3) == 1)).any() and (covered & ((seq >> 3) == 2)).any()

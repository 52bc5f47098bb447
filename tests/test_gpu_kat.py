"""Hand-made edge cases through the CUDA path vs the oracle (bit-exact visibility buffer and RGBA8)."""
import numpy as np
import pytest

from helpers import identity_camera, triangle_scene, render_gpu, render_oracle, rgba_bytes

pytestmark = pytest.mark.gpu


def compare(sc, cam, W, H):
    g = render_gpu(sc, cam, W, H)
    o = render_oracle(sc, cam, W, H)
    for k in ("seq", "depth"):
        assert np.array_equal(g[k], o[k]), f"{k}: {np.count_nonzero(g[k] != o[k])} pixels differ"
    for k in ("bary1", "bary2"):
        assert np.array_equal(g[k].view(np.uint32), o[k].view(np.uint32)), k
    assert not np.any(g["depth"] == 0xDEADBEEF)
    assert np.abs(rgba_bytes(g["pixels"]) - rgba_bytes(o["pixels"])).max() <= 1
    for k in ("triangles_binned", "triangles_clipped", "tile_refs"):
        assert g["stats"][k] == o["stats"][k], (k, g["stats"][k], o["stats"][k])
    return g, o


def random_tris(rng, n, W, H, size, front=True):
    tris, depths = [], []
    while len(tris) < n:
        cx, cy = rng.integers(-size // 4, W * 16 + size // 4), rng.integers(-size // 4, H * 16 + size // 4)
        p = [(int(cx + rng.integers(-size, size)), int(cy + rng.integers(-size, size))) for _ in range(3)]
        area = (p[1][0] - p[0][0]) * (p[2][1] - p[0][1]) - (p[2][0] - p[0][0]) * (p[1][1] - p[0][1])
        if front and area > 0:
            p = [p[0], p[2], p[1]]
        tris.append(p)
        depths.append([float(rng.uniform(0.05, 0.95)) for _ in range(3)])
    return tris, depths


def test_small_triangles_ties_and_edges():
    W, H = 128, 64
    rng = np.random.default_rng(7)
    tris, depths = random_tris(rng, 400, W, H, 200)
    depths = [[d[0]] * 3 if i % 3 else [0.5] * 3 for i, d in enumerate(depths)]  # many exact depth ties
    tris += [[(16, 16), (16, H * 16), (W * 16, 16)], [(W * 16, H * 16), (W * 16, 32), (32, H * 16)]]  # screen-edge touchers
    depths += [[0.5] * 3, [0.6] * 3]
    compare(triangle_scene(tris, W, H, depths=depths), identity_camera(W, H), W, H)


def test_large_triangles_4k_inexact_f32_chain():
    """Triangles far beyond the 2^24 exactness bound at 3840x2160: the f32 stepping chain, the coarse 16x16 reject and the
    tile-clipped chain origins must all be replayed exactly."""
    W, H = 3840, 2160
    rng = np.random.default_rng(11)
    tris, depths = random_tris(rng, 48, W, H, 30000)
    t2, d2 = random_tris(rng, 200, W, H, 1500)
    g, o = compare(triangle_scene(tris + t2, W, H, depths=depths + d2), identity_camera(W, H), W, H)
    assert (o["seq"] != 0xFFFFFFFF).mean() > 0.3


def test_8k_i32_wrap():
    """At 7680x4320 the products x*y in the edge constants exceed 2^31 and wrap (release-mode Rust): SURVEY 7.2-5."""
    W, H = 7680, 4320
    rng = np.random.default_rng(13)
    tris, depths = random_tris(rng, 24, W, H, 40000)
    # force some vertices into the far bottom-right corner where x*y > 2^31 in sub-pixels
    tris += [[(W * 16 - 50, H * 16 - 4000), (W * 16 - 3000, H * 16 - 30), (W * 16 - 20, H * 16 - 10)],
             [(60000, 50000), (60000, 69000), (122000, 50000)]]
    depths += [[0.3, 0.3, 0.3], [0.4, 0.5, 0.6]]
    compare(triangle_scene(tris, W, H, depths=depths), identity_camera(W, H), W, H)


def test_clipper_all_planes():
    """Intersecting draw: triangles poking through every frustum plane (x, y beyond +-1, z beyond [-1, 1])."""
    W, H = 640, 384
    rng = np.random.default_rng(17)
    tris, depths = random_tris(rng, 300, W, H, 6000)
    depths = [[float(rng.uniform(-1.6, 1.6)) for _ in range(3)] for _ in tris]
    g, o = compare(triangle_scene(tris, W, H, depths=depths), identity_camera(W, H, intersecting=True), W, H)
    assert o["stats"]["triangles_clipped"] > 100
    fans = np.unique(o["seq"][o["seq"] != 0xFFFFFFFF] & 7)
    assert fans.max() >= 3, "expected polygons with 5+ vertices"


def test_single_fullscreen_triangle_and_tiny_screen():
    for W, H in ((64, 64), (2, 2), (130, 66)):
        tris = [[(-W * 16, -H * 16), (-W * 16, 3 * H * 16), (3 * W * 16, -H * 16)]]
        compare(triangle_scene(tris, W, H, depths=[[0.2, 0.7, 0.4]]), identity_camera(W, H), W, H)
        compare(triangle_scene(tris, W, H, depths=[[0.2, 0.7, 0.4]]), identity_camera(W, H, intersecting=True), W, H)


def test_unorm8_exact_for_all_bytes():
    """fetch_texel's division-free byte/255 must equal the IEEE division for all 256 inputs: a 256-texel grey ramp
    texture sampled by a screen-filling quad reproduces the oracle exactly (also covers every sRGB ramp value)."""
    from swraster_viewer_b200 import abi, scenes
    W, H = 512, 64
    ramp = np.arange(256, dtype=np.uint32)
    tex = scenes.make_texture(((ramp << 24) | (ramp << 16) | (ramp << 8) | 0xFF)[None, None, :], 256, 1, abi.TEX_SRGB, abi.WRAP_CLAMP_TO_EDGE, mips=False)
    tris = [[(0, 0), (0, H * 16), (W * 16, 0)], [(W * 16, H * 16), (W * 16, 0), (0, H * 16)]]
    sc = triangle_scene(tris, W, H)
    sc.textures[0:0] = []  # keep indices: append the ramp and point the material at it
    sc.textures.append(tex)
    sc.materials[0].base_color_texture = len(sc.textures) - 1
    sc.materials[0].emissive_factor = (1.0, 1.0, 1.0)
    sc.materials[0].emissive_texture = len(sc.textures) - 1
    sc._desc = None
    g, o = compare(sc, identity_camera(W, H), W, H)
    assert np.array_equal(g["pixels"], o["pixels"])

"""Load-time bakes ON THE DEVICE (include/swr.h swr_bake_*, csrc/swr_bake.cuh) against their CPU checkers
(oracle/oracle_bake.cpp, oracle/oracle_sunvis.cpp — themselves pinned by tests/test_bakes.py and tests/test_sunvis.py), plus
the loader paths that consume them (swrh_env_bake, swrh_compute_sun_visibility, gltf.load_scene) end to end.
Tolerance: CUDA's sinf / cosf / powf differ from the C library's by <= 2 ulp, which can move a texel across an 8-bit
boundary: at most 1 LSB per channel, on a small share of the texels; SH coefficients and visibilities to 1e-5."""
import math

import numpy as np
import pytest

import oracle as orc
import swraster_viewer_b200 as swr
from swraster_viewer_b200 import abi, gltf, scenes
from helpers import SMALL, render_gpu, render_oracle, rgba_bytes
from test_bakes import cross_from_faces, unpack
from test_sunvis import shadow_scene

pytestmark = pytest.mark.gpu
F32 = np.float32


def sky_faces(size, seed):
    tex, _ = scenes.sky_cubemap(size, seed)
    return unpack(tex.data[:6 * size * size]).reshape(6, size, size, 4).astype(np.uint8)


def test_cross_layout_mips_and_structure():
    rng = np.random.default_rng(2)
    faces = rng.integers(0, 256, (6, 8, 8, 4), dtype=np.uint8)
    env = gltf.BakedEnvironment(cross_from_faces(faces), lut_size=8, specular_samples=4, voxel_dim=2)
    data, offs, ws, hs, st, typ = env.texture("cubemap")
    assert typ == abi.TEX_CUBEMAP and list(ws) == [8, 4, 2, 1] and list(st) == [64, 16, 4, 1] and list(offs) == [0, 384, 480, 504]
    assert np.array_equal(data[:384], scenes.pack_rgba8(faces).reshape(-1))  # faces in the order +X -X +Y -Y +Z -Z
    ref = scenes.make_texture(scenes.pack_rgba8(faces), 8, 8, abi.TEX_CUBEMAP, abi.WRAP_CLAMP_TO_EDGE, slices=6)  # plain 2x2 average per face
    assert np.array_equal(data, ref.data)
    sdata, soffs, sws, shs, sst, styp = env.texture("cubemap_specular")
    assert styp == abi.TEX_LINEAR and list(sws) == [8] * 4 and list(sst) == [64] * 4 and list(soffs) == [0, 384, 768, 1152] and len(sdata) == 1536
    ldata, loffs, lws, _, lst, ltyp = env.texture("brdf_lut")
    assert ltyp == abi.TEX_LINEAR and list(lws) == [8, 4, 2, 1] and list(lst) == [0, 16, 4, 1]
    with pytest.raises(gltf.GltfError, match="smaller than 4x3"):
        gltf.BakedEnvironment(np.zeros((2, 3, 4), np.uint8))


def test_device_brdf_lut_matches_the_checker():
    N = 64
    env = gltf.BakedEnvironment(np.full((3, 4, 4), 128, np.uint8), lut_size=N, specular_samples=2, voxel_dim=1)
    got = unpack(env.texture("brdf_lut")[0][:N * N])
    want = unpack(orc.bake_brdf_lut(N).reshape(-1))
    d = np.abs(got - want)
    print(f"BRDF LUT {N}x{N}: max |diff| {d.max()} LSB, {np.count_nonzero(d)} of {d.size} channel values differ")
    assert d.max() <= 1 and np.count_nonzero(d) <= 0.01 * d.size


def test_device_irradiance_sh4_matches_the_checker():
    for size, seed in ((16, 5), (64, 7)):
        faces = sky_faces(size, seed)
        env = gltf.BakedEnvironment(cross_from_faces(faces), lut_size=4, specular_samples=2, voxel_dim=1)
        want = orc.bake_irradiance_sh4(scenes.pack_rgba8(faces))
        assert np.allclose(env.irradiance_sh, want, rtol=2e-5, atol=2e-6), (size, env.irradiance_sh, want)


def test_device_prefilter_matches_the_checker():
    rng = np.random.default_rng(9)
    noise = rng.integers(0, 256, (6, 8, 8, 4), dtype=np.uint8)
    noise[..., 3] = 255
    for faces, S in ((noise, 16), (sky_faces(32, 11), 64)):
        env = gltf.BakedEnvironment(cross_from_faces(faces), lut_size=4, specular_samples=S, voxel_dim=1)
        got = unpack(env.texture("cubemap_specular")[0])
        want = unpack(orc.bake_prefilter_specular(scenes.pack_rgba8(faces), S).reshape(-1))
        assert got.shape == want.shape
        d = np.abs(got - want)
        print(f"prefilter {faces.shape[1]}^2 x {S} samples: max |diff| {d.max()} LSB, {np.count_nonzero(d)} of {d.size} channel values differ")
        assert d.max() <= 1 and np.count_nonzero(d) <= 0.01 * d.size
        assert (got[:, 3] == 255).all()


def test_voxel_grid_initialisation_and_use_as_loader_environment(tmp_path):
    faces = sky_faces(8, 3)
    env = gltf.BakedEnvironment(cross_from_faces(faces), lut_size=16, specular_samples=8, voxel_dim=3, irradiance_scale=0.25, sky_visibility=1.0, light_intensity=0.5)
    vox = env.voxels()
    assert vox.shape == (27, 4, 4)
    assert np.array_equal(vox[:, :, :3], np.broadcast_to(env.irradiance_sh * F32(0.25), (27, 4, 3)))  # gi.rs:137-144
    assert (vox[:, 0, 3] == 0.5).all() and (vox[:, 1, 3] == 1.0).all() and (vox[:, 2:, 3] == 0.0).all()  # gi.rs:145-146
    # the baked environment completes a loaded glTF file into a renderable scene: oracle and CUDA path agree on it
    sc, spec = scenes.scene_c1_sphere(segments=24, bands=16, **SMALL)
    scenes.export_gltf(sc, str(tmp_path / "s"))
    g = gltf.load_gltf(tmp_path / "s.gltf", environment=env)
    d = g.desc()
    assert d.ntextures == 3 and (d.cubemap, d.cubemap_specular, d.brdf_lut) == (0, 1, 2)
    assert np.allclose(d.voxel_grid.world_min[:], g.bounds_min) and np.allclose(d.voxel_grid.world_max[:], g.bounds_max)
    W, H = 160, 96
    cam = swr.RenderCamera.from_spec(spec, W, H)
    o = render_oracle(g, cam, W, H)
    px = np.stack([(o["pixels"] >> 24) & 255, (o["pixels"] >> 16) & 255, (o["pixels"] >> 8) & 255], -1).reshape(H, W, 3)
    covered = (o["seq"] != 0xFFFFFFFF).reshape(H, W)
    assert covered.any() and (~covered).any()
    assert px[covered].mean() > 20 and px[~covered].mean() > 60  # the sphere is lit, the sky is the bright procedural sky
    gg = render_gpu(g, cam, W, H)
    assert np.array_equal(gg["seq"], o["seq"]) and np.array_equal(gg["depth"], o["depth"])
    assert np.abs(rgba_bytes(gg["pixels"]) - rgba_bytes(o["pixels"])).max() <= 1


@pytest.mark.parametrize("lid", [False, True])
def test_device_sun_visibility_matches_the_checker(lid):
    sc, _ = shadow_scene(lid, voxel_dim=32)
    got = gltf.compute_sun_visibility(sc)
    want = orc.sun_visibility(sc)
    assert got.shape == (32, 32, 32)
    assert np.allclose(got, want, rtol=0, atol=1e-5), np.abs(got - want).max()
    assert got.min() < 0.05 and got.max() == 1.0
    if lid:
        assert ((got > 0.15) & (got < 0.35)).any()  # under the 0.5-transmission sheet: 0.5, squared by the blur


def test_device_sun_visibility_of_an_empty_scene():
    sc, _ = shadow_scene(False, voxel_dim=4)
    sc.nodes = []
    sc._desc = None
    sc.node_spheres = []
    assert np.array_equal(gltf.compute_sun_visibility(sc), np.ones((4, 4, 4), F32))  # no active voxel: no rays, no blur (gi.rs:281-284)


def test_shadow_reaches_the_frame_through_the_loader(tmp_path):
    sc, spec = shadow_scene(False, voxel_dim=40)
    scenes.export_gltf(sc, str(tmp_path / "s"))
    W, H = 256, 160
    cam = swr.RenderCamera.from_spec(spec, W, H)
    lit = gltf.load_gltf(tmp_path / "s.gltf", environment=sc)
    vis = gltf.compute_sun_visibility(lit)
    shadowed = gltf.load_gltf(tmp_path / "s.gltf", environment=sc)
    gltf.bake_sun_visibility(shadowed)
    d = shadowed.desc()
    nv = int(np.prod(d.voxel_grid.dims[:]))
    w0 = np.ctypeslib.as_array(d.voxel_grid.gi_sh4, (nv * 16,)).reshape(nv, 16)[:, 3]
    assert np.array_equal(w0, vis.reshape(-1))
    a, b = render_oracle(lit, cam, W, H), render_oracle(shadowed, cam, W, H)
    assert np.array_equal(a["seq"], b["seq"])  # visibility is untouched, only the lighting changes
    lum = lambda o: ((o["pixels"] >> 24) & 255).astype(np.int32) + ((o["pixels"] >> 16) & 255) + ((o["pixels"] >> 8) & 255)
    darker = (lum(a) - lum(b)) > 30
    assert 0.005 < darker.mean() < 0.5, darker.mean()  # a shadow patch on the ground, not the whole frame


def test_load_scene_is_the_viewers_load_path_in_one_call(tmp_path):
    """gltf.load_scene = parse + environment bake + voxel grid over the bounds + SH initialisation + sun visibility
    (main.rs:100-291); the result renders (oracle) with the default camera of main.rs:210-224."""
    from test_bakes import cross_from_faces, unpack
    sc, _ = shadow_scene(False, voxel_dim=4)
    scenes.export_gltf(sc, str(tmp_path / "s"))
    tex, _ = scenes.sky_cubemap(8, 3)
    faces = unpack(tex.data[:6 * 64]).reshape(6, 8, 8, 4).astype(np.uint8)
    scene, (pos, look, fov, far) = gltf.load_scene(tmp_path / "s.gltf", cross_from_faces(faces), grid_size=24, lut_size=16, specular_samples=8)
    d = scene.desc()
    assert tuple(d.voxel_grid.dims[:]) == (24, 24, 24) and np.allclose(d.voxel_grid.world_min[:], scene.bounds_min)
    vox = np.ctypeslib.as_array(d.voxel_grid.gi_sh4, (24 ** 3 * 16,)).reshape(-1, 4, 4)
    assert vox[:, 0, 3].min() < 0.1 and vox[:, 0, 3].max() == 1.0 and (vox[:, 1, 3] == 1.0).all()  # sun visibility in, sky visibility 1
    assert np.array_equal(vox[0, :, :3], vox[-1, :, :3]) and np.abs(vox[0, 0, :3]).min() > 0  # the same scaled SH everywhere
    assert pos[2] == pytest.approx(float(scene.bounds_center[2]) + float(scene.bounds_diagonal)) and far == pytest.approx(2 * float(scene.bounds_diagonal))
    W, H = 160, 96
    cam = swr.RenderCamera(pos, look, fov, W, H, far)
    o = render_oracle(scene, cam, W, H)
    assert (o["seq"] != 0xFFFFFFFF).mean() > 0.02


def test_load_scene_camera_rules(tmp_path):
    """main.rs:198-224: the first glTF camera wins and only its yfov is used; orthographic is a load error; none -> default."""
    import json
    from test_bakes import cross_from_faces
    sc, _ = shadow_scene(False, voxel_dim=4)
    scenes.export_gltf(sc, str(tmp_path / "s"))
    sky = cross_from_faces(np.full((6, 4, 4, 4), 200, np.uint8))
    doc = json.load(open(tmp_path / "s.gltf"))
    doc["cameras"] = [{"type": "perspective", "perspective": {"yfov": 0.6, "znear": 0.1, "zfar": 50.0}}, {"type": "orthographic", "orthographic": {"xmag": 1, "ymag": 1, "znear": 0.1, "zfar": 5}}]
    doc["nodes"].append({"camera": 0, "translation": [3, 4, 5]})
    json.dump(doc, open(tmp_path / "p.gltf", "w"))
    _, cam = gltf.load_scene(tmp_path / "p.gltf", sky, grid_size=4, lut_size=4, specular_samples=2)
    assert cam == ((0.0, 0.0, 5.0), (0.0, 0.0, 0.0), pytest.approx(0.6), 1000.0)
    doc["cameras"].reverse()
    json.dump(doc, open(tmp_path / "o.gltf", "w"))
    with pytest.raises(gltf.GltfError, match="unsupported camera type"):
        gltf.load_scene(tmp_path / "o.gltf", sky, grid_size=4, lut_size=4, specular_samples=2)


def test_prefiltered_cubemap_through_the_ggx_cache(tmp_path):
    """scene.rs:164-206: no cache -> bake (on the device) and write `cubemap.ggx`; a cache of the sky's face size -> read it,
    no bake; a cache of another size is not this sky's and is replaced."""
    rng = np.random.default_rng(9)
    faces = rng.integers(0, 256, size=(6, 8, 8, 4), dtype=np.uint8)
    cross = cross_from_faces(faces)
    p = tmp_path / "cubemap.ggx"
    first = gltf.BakedEnvironment(cross, lut_size=4, specular_samples=4, voxel_dim=1, ggx_cache=p)
    assert not first.specular_from_cache and p.exists()
    baked = first.texture("cubemap_specular")[0]
    assert np.array_equal(gltf.ggx_cache_load(p, 8, 8).ravel(), baked)
    # poison the file's texels: a second load must take them from the cache, not bake again
    marked = gltf.ggx_cache_load(p, 8, 8) ^ np.uint32(0x01000000)
    gltf.ggx_cache_save(p, marked)
    second = gltf.BakedEnvironment(cross, lut_size=4, specular_samples=4, voxel_dim=1, ggx_cache=p)
    assert second.specular_from_cache and np.array_equal(second.texture("cubemap_specular")[0], marked.ravel())
    # another sky size: the cache is not used and is rewritten for this sky
    small = cross_from_faces(faces[:, :4, :4])
    third = gltf.BakedEnvironment(small, lut_size=4, specular_samples=4, voxel_dim=1, ggx_cache=p)
    assert not third.specular_from_cache and gltf.ggx_cache_load(p, 4, 4) is not None and gltf.ggx_cache_load(p, 8, 8) is None
